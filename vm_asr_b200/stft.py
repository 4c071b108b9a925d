"""Magnitude/phase STFT and inverse: drop-ins for ``utils/stft.py`` ``wav2spectro`` (:22-68) and
``spectro2wav`` (:71-115), same argument lists and return shapes.  One CUDA kernel each
(``vmasr_stft_fwd`` / ``vmasr_istft_fwd``); ``spectro2wav`` is differentiable with respect to ``mag`` and
``phase`` through ``vmasr_istft_bwd`` (the generator loss flows through it, model/model.py:1223).
``wav2spectro`` is applied to data, not to activations, in the generator (model/model.py:424-434), so it
carries no gradient here.  Only ``spectro_scale="log2"`` (config.py:58) is implemented."""
from __future__ import annotations

from typing import Tuple

import torch

from . import _lib


def _dev(t):
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def _check_scale(scale):
    if scale != "log2":
        raise NotImplementedError("vmasr_b200 implements the log2 spectrogram scale only (config.py:58 SCALE)")


def wav2spectro(waveform: torch.Tensor, n_fft: int, hop_length: int, win_length: int,
                spectro_scale: str = "log2") -> Tuple[torch.Tensor, torch.Tensor]:
    _check_scale(spectro_scale)
    lib = _lib.load_library()
    _lib.require_cuda(waveform, "waveform")
    *other, length = waveform.shape
    wave = waveform.detach().reshape(-1, length).to(torch.float32).contiguous()
    Bsz = wave.shape[0]
    freqs, frames = n_fft // 2 + 1, 1 + length // hop_length
    mag = torch.empty((Bsz, freqs, frames), dtype=torch.float32, device=wave.device)
    phase = torch.empty_like(mag)
    with torch.cuda.device(wave.device):
        _lib.check(lib.vmasr_stft_fwd(wave.data_ptr(), mag.data_ptr(), phase.data_ptr(), Bsz, length, n_fft,
                                      hop_length, win_length, _dev(wave), _lib.current_stream_ptr(wave.device)))
    return mag.view(*other, freqs, frames), phase.view(*other, freqs, frames)


class _Spectro2Wav(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mag, phase, hop_length, win_length):
        lib = _lib.load_library()
        Bsz, freqs, frames = mag.shape
        n_fft = 2 * freqs - 2
        wave = torch.empty((Bsz, hop_length * (frames - 1)), dtype=torch.float32, device=mag.device)
        with torch.cuda.device(mag.device):
            _lib.check(lib.vmasr_istft_fwd(mag.data_ptr(), phase.data_ptr(), wave.data_ptr(), Bsz, frames, n_fft,
                                           hop_length, win_length, _dev(mag), _lib.current_stream_ptr(mag.device)))
        ctx.save_for_backward(mag, phase)
        ctx.cfg = (n_fft, hop_length, win_length)
        return wave

    @staticmethod
    def backward(ctx, dwave):
        lib = _lib.load_library()
        mag, phase = ctx.saved_tensors
        n_fft, hop_length, win_length = ctx.cfg
        Bsz, freqs, frames = mag.shape
        dwave = dwave.to(torch.float32).contiguous()
        dmag = torch.empty_like(mag)
        dphase = torch.empty_like(phase)
        with torch.cuda.device(mag.device):
            _lib.check(lib.vmasr_istft_bwd(mag.data_ptr(), phase.data_ptr(), dwave.data_ptr(), dmag.data_ptr(),
                                           dphase.data_ptr(), Bsz, frames, n_fft, hop_length, win_length, _dev(mag),
                                           _lib.current_stream_ptr(mag.device)))
        return dmag, dphase, None, None


def spectro2wav(mag: torch.Tensor, phase: torch.Tensor, n_fft: int, hop_length: int, win_length: int,
                spectro_scale: str = "log2") -> torch.Tensor:
    """``n_fft`` is accepted and, as in the reference (stft.py:86-87), re-derived from the number of bins."""
    _check_scale(spectro_scale)
    _lib.require_cuda(mag, "mag")
    _lib.require_cuda(phase, "phase")
    *other, freqs, frames = mag.shape
    in_dtype = mag.dtype
    m = mag.reshape(-1, freqs, frames).to(torch.float32).contiguous()
    p = phase.reshape(-1, freqs, frames).to(torch.float32).contiguous()
    wave = _Spectro2Wav.apply(m, p, hop_length, win_length)
    if in_dtype != torch.float32 and in_dtype.is_floating_point:
        pass  # the reference's torch.istft returns float32 for float32 spectra; half inputs are promoted here
    return wave.view(*other, wave.shape[-1])
