"""Multi-resolution STFT loss and the LSD metrics on this library's FFT kernel: drop-ins for
``model/loss.py`` ``MultiResolutionSTFTLoss`` (:137-184, with ``stft`` :17-45, ``SpectralConvergengeLoss`` :48-63,
``LogSTFTMagnitudeLoss`` :66-82, ``STFTLoss`` :85-134) and ``model/metric.py`` ``lsd`` / ``lsd_hf`` / ``lsd_lf`` (:5-12, :25-67).

The reference runs ``torch.stft`` (cuFFT + ~6 elementwise launches) six times per training step for the loss (three
resolutions, prediction and target) and four more for the metrics; here each is one ``vmasr_stft_mag_fwd`` launch and its
autograd backward one ``vmasr_stft_mag_bwd``.  Everything after the magnitude (norms, logs, means: a few reductions over
(B, F, frames)) stays on PyTorch."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .stft import stft_magnitude


def stft(x, fft_size, hop_size, win_length, window=None, emphasize_high_freq=False):
    """model/loss.py:17-45: magnitude spectrogram (B, frames, fft_size // 2 + 1) with the 1e-7 clamp under the square root.
    ``window`` is accepted for signature compatibility; the kernel builds the periodic Hann of ``win_length`` itself."""
    magnitude = stft_magnitude(x, fft_size, hop_size, win_length, normalized=False, clamp_min=1e-7).transpose(2, 1)
    if emphasize_high_freq:   # loss.py:39-43 (weights along dim 1 of the transposed tensor, as there)
        freq_weights = torch.linspace(1.0, 2.0, magnitude.size(1), device=x.device).view(1, -1, 1)
        magnitude = magnitude * freq_weights
    return magnitude


class STFTLoss(torch.nn.Module):
    def __init__(self, fft_size=1024, shift_size=120, win_length=600, window="hann_window", emphasize_high_freq=False):
        super().__init__()
        if window != "hann_window":
            raise NotImplementedError("vmasr_b200 builds the Hann window in the kernel; other windows are not implemented")
        self.fft_size, self.shift_size, self.win_length = fft_size, shift_size, win_length
        self.emphasize_high_freq = emphasize_high_freq

    def forward(self, x, y):
        x_mag = stft(x, self.fft_size, self.shift_size, self.win_length, None, self.emphasize_high_freq)
        y_mag = stft(y, self.fft_size, self.shift_size, self.win_length, None, self.emphasize_high_freq)
        sc_loss = torch.norm(y_mag - x_mag, p="fro") / torch.norm(y_mag, p="fro")     # loss.py:63
        mag_loss = F.l1_loss(torch.log(y_mag), torch.log(x_mag))                      # loss.py:82
        return sc_loss, mag_loss


class MultiResolutionSTFTLoss(torch.nn.Module):
    """Same constructor and return value as model/loss.py:137-184: ``(factor_sc * sc_loss, factor_mag * mag_loss)``."""

    def __init__(self, fft_sizes=(1024, 2048, 512), hop_sizes=(120, 240, 50), win_lengths=(600, 1200, 240), window="hann_window",
                 factor_sc=0.1, factor_mag=0.1, emphasize_high_freq=False):
        super().__init__()
        assert len(fft_sizes) == len(hop_sizes) == len(win_lengths)
        self.stft_losses = torch.nn.ModuleList(
            [STFTLoss(fs, ss, wl, window, emphasize_high_freq) for fs, ss, wl in zip(fft_sizes, hop_sizes, win_lengths)])
        self.factor_sc, self.factor_mag = factor_sc, factor_mag

    def forward(self, x, y):
        sc_loss, mag_loss = 0.0, 0.0
        for f in self.stft_losses:
            sc_l, mag_l = f(x, y)
            sc_loss = sc_loss + sc_l
            mag_loss = mag_loss + mag_l
        sc_loss = sc_loss / len(self.stft_losses)
        mag_loss = mag_loss / len(self.stft_losses)
        return self.factor_sc * sc_loss, self.factor_mag * mag_loss


# ---- model/metric.py ---------------------------------------------------------------------------------------------------
def _metric_spec(audio, n_fft=2048, hop_length=512):
    """metric.py:5-12: |torch.stft(audio, 2048, 512, window=hann(2048))|, (B, F, frames)."""
    return stft_magnitude(audio, n_fft, hop_length, n_fft, normalized=False, clamp_min=0.0)


def _log_power(audio):
    return torch.log10(_metric_spec(audio).square().clamp(1e-8))


def lsd(output, target, **kwargs):
    """metric.py:25-28."""
    sp, st = _log_power(output), _log_power(target)
    return (sp - st).square().mean(dim=1).sqrt().mean().item()


def _lsd_band(output, target, hf, high):
    sp, st = _log_power(output), _log_power(target)
    val = []
    for i in range(output.size(0)):
        hf_i = int(hf[i].item())
        band = slice(hf_i, None) if high else slice(None, hf_i)
        val.append((sp[i, band, :] - st[i, band, :]).square().mean(dim=0).sqrt().mean().item())
    return torch.tensor(val).mean().item()


def lsd_hf(output, target, hf):
    """metric.py:31-47."""
    return _lsd_band(output, target, hf, True)


def lsd_lf(output, target, hf):
    """metric.py:50-66."""
    return _lsd_band(output, target, hf, False)
