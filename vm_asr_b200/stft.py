"""Magnitude/phase STFT and inverse: drop-ins for ``utils/stft.py`` ``wav2spectro`` (:22-68) and
``spectro2wav`` (:71-115), same argument lists and return shapes, both differentiable like the reference's
(``torch.stft`` / ``torch.istft`` + elementwise ops): ``vmasr_stft_fwd`` / ``vmasr_stft_bwd`` and ``vmasr_istft_fwd`` /
``vmasr_istft_bwd``.  ``stft_magnitude`` is the un-normalised linear-magnitude STFT the multi-resolution loss and the LSD
metric are built on (model/loss.py:17-45, model/metric.py:5-12; ``vm_asr_b200.loss``).

``spectro_scale="log2"`` (config.py:58, every shipped config) runs entirely in the kernels.  ``"dB"`` (utils/stft.py:59-62,
100-102) needs a maximum over the whole batch (torchaudio's ``top_db`` clamp), so it is the linear-magnitude kernel plus a
few torch elementwise ops; the inverse maps dB to log2 and uses the same iSTFT kernel."""
from __future__ import annotations

import math
from typing import Tuple

import torch

from . import _lib


def _dev(t):
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def _scratch(wave_like: torch.Tensor, Bsz: int, frames: int, n_fft: int, hop: int) -> torch.Tensor:
    n = int(_lib.load_library().vmasr_stft_scratch_floats(frames, n_fft, hop))
    return torch.empty(Bsz * n, dtype=torch.float32, device=wave_like.device)


class _Wav2Spectro(torch.autograd.Function):
    @staticmethod
    def forward(ctx, wave, n_fft, hop_length, win_length):
        lib = _lib.load_library()
        Bsz, length = wave.shape
        freqs, frames = n_fft // 2 + 1, 1 + length // hop_length
        mag = torch.empty((Bsz, freqs, frames), dtype=torch.float32, device=wave.device)
        phase = torch.empty_like(mag)
        with torch.cuda.device(wave.device):
            _lib.check(lib.vmasr_stft_fwd(wave.data_ptr(), mag.data_ptr(), phase.data_ptr(), Bsz, length, n_fft,
                                          hop_length, win_length, _dev(wave), _lib.current_stream_ptr(wave.device)))
        ctx.save_for_backward(wave)
        ctx.cfg = (n_fft, hop_length, win_length)
        return mag, phase

    @staticmethod
    def backward(ctx, dmag, dphase):
        lib = _lib.load_library()
        (wave,) = ctx.saved_tensors
        n_fft, hop, win = ctx.cfg
        Bsz, length = wave.shape
        frames = 1 + length // hop
        dmag = dmag.to(torch.float32).contiguous()
        dphase = dphase.to(torch.float32).contiguous()
        dwave = torch.empty_like(wave)
        scratch = _scratch(wave, Bsz, frames, n_fft, hop)
        with torch.cuda.device(wave.device):
            _lib.check(lib.vmasr_stft_bwd(wave.data_ptr(), dmag.data_ptr(), dphase.data_ptr(), dwave.data_ptr(), scratch.data_ptr(),
                                          Bsz, length, n_fft, hop, win, _dev(wave), _lib.current_stream_ptr(wave.device)))
        return dwave, None, None, None


class _StftMagnitude(torch.autograd.Function):
    @staticmethod
    def forward(ctx, wave, n_fft, hop_length, win_length, normalized, clamp_min):
        lib = _lib.load_library()
        Bsz, length = wave.shape
        freqs, frames = n_fft // 2 + 1, 1 + length // hop_length
        mag = torch.empty((Bsz, freqs, frames), dtype=torch.float32, device=wave.device)
        with torch.cuda.device(wave.device):
            _lib.check(lib.vmasr_stft_mag_fwd(wave.data_ptr(), mag.data_ptr(), Bsz, length, n_fft, hop_length, win_length,
                                              1 if normalized else 0, float(clamp_min), _dev(wave),
                                              _lib.current_stream_ptr(wave.device)))
        ctx.save_for_backward(wave)
        ctx.cfg = (n_fft, hop_length, win_length, normalized, clamp_min)
        return mag

    @staticmethod
    def backward(ctx, dmag):
        lib = _lib.load_library()
        (wave,) = ctx.saved_tensors
        n_fft, hop, win, normalized, clamp_min = ctx.cfg
        Bsz, length = wave.shape
        frames = 1 + length // hop
        dmag = dmag.to(torch.float32).contiguous()
        dwave = torch.empty_like(wave)
        scratch = _scratch(wave, Bsz, frames, n_fft, hop)
        with torch.cuda.device(wave.device):
            _lib.check(lib.vmasr_stft_mag_bwd(wave.data_ptr(), dmag.data_ptr(), dwave.data_ptr(), scratch.data_ptr(), Bsz, length,
                                              n_fft, hop, win, 1 if normalized else 0, float(clamp_min), _dev(wave),
                                              _lib.current_stream_ptr(wave.device)))
        return dwave, None, None, None, None, None


def _flat_wave(waveform: torch.Tensor, name: str):
    _lib.require_cuda(waveform, name)
    *other, length = waveform.shape
    return waveform.reshape(-1, length).to(torch.float32).contiguous(), other


def stft_magnitude(waveform: torch.Tensor, n_fft: int, hop_length: int, win_length: int, normalized: bool = False,
                   clamp_min: float = 0.0) -> torch.Tensor:
    """``sqrt(clamp(|torch.stft(x, n_fft, hop, win, hann)|^2, min=clamp_min))`` as (..., n_fft/2+1, frames), differentiable.
    model/loss.py:30-37 (clamp 1e-7, then transposed there), model/metric.py:5-12 (clamp 0, win_length = n_fft)."""
    wave, other = _flat_wave(waveform, "waveform")
    mag = _StftMagnitude.apply(wave, n_fft, hop_length, win_length, normalized, clamp_min)
    return mag.view(*other, mag.shape[-2], mag.shape[-1])


def wav2spectro(waveform: torch.Tensor, n_fft: int, hop_length: int, win_length: int,
                spectro_scale: str = "log2") -> Tuple[torch.Tensor, torch.Tensor]:
    wave, other = _flat_wave(waveform, "waveform")
    freqs, frames = n_fft // 2 + 1, 1 + wave.shape[-1] // hop_length
    if spectro_scale == "dB":
        # utils/stft.py:59-62: AmplitudeToDB(stype="power", top_db=80)(|X|^2) = 10 log10(clamp(|X|^2, 1e-10)), floored at
        # (maximum over the whole batch) - 80; the phase is the log2 path's
        lin = _StftMagnitude.apply(wave, n_fft, hop_length, win_length, True, 0.0)
        db = 10.0 * torch.log10(torch.clamp(lin * lin, min=1e-10))
        mag = torch.maximum(db, db.amax().detach() - 80.0)
        _, phase = _Wav2Spectro.apply(wave, n_fft, hop_length, win_length)
    else:
        mag, phase = _Wav2Spectro.apply(wave, n_fft, hop_length, win_length)
    return mag.view(*other, freqs, frames), phase.view(*other, freqs, frames)


class _Spectro2Wav(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mag, phase, hop_length, win_length):
        lib = _lib.load_library()
        Bsz, freqs, frames = mag.shape
        n_fft = 2 * freqs - 2
        wave = torch.empty((Bsz, hop_length * (frames - 1)), dtype=torch.float32, device=mag.device)
        scratch = _scratch(mag, Bsz, frames, n_fft, hop_length)
        with torch.cuda.device(mag.device):
            _lib.check(lib.vmasr_istft_fwd(mag.data_ptr(), phase.data_ptr(), wave.data_ptr(), scratch.data_ptr(), Bsz, frames, n_fft,
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
    if spectro_scale not in ("log2", "dB"):
        raise ValueError(f"spectro_scale must be 'log2' or 'dB', got {spectro_scale!r}")
    _lib.require_cuda(mag, "mag")
    _lib.require_cuda(phase, "phase")
    *other, freqs, frames = mag.shape
    m = mag.reshape(-1, freqs, frames).to(torch.float32)
    if spectro_scale == "dB":
        # utils/stft.py:100-102: DB_to_amplitude(x, ref=1, power=1) ** 0.5 = 10 ** (x / 20) = 2 ** (x log2(10) / 20)
        m = m * (math.log2(10.0) / 20.0)
    p = phase.reshape(-1, freqs, frames).to(torch.float32).contiguous()
    wave = _Spectro2Wav.apply(m.contiguous(), p, hop_length, win_length)
    return wave.view(*other, wave.shape[-1])
