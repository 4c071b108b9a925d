"""GPU parity: STFT / iSTFT / iSTFT-backward kernels against the float64 oracle and the reference's own outputs
(tests/golden/stft.npz, produced by utils/stft.py).  Tolerance: 1e-5 absolute (BASELINE.json) on the complex
spectrum and on the waveform; on log2-magnitude and phase where |X| is not tiny (both are ill-conditioned at
|X| -> 0: d log2|X| = d|X| / (|X| ln 2), so a 1e-7 error in X moves them by more than 1e-5 once |X| < 1e-2)."""
import os

import numpy as np
import pytest
import torch

from oracle import stft_ref

pytestmark = pytest.mark.gpu
ABS_TOL = 1e-5


def _ops():
    from vm_asr_b200 import stft
    return stft


def _check_spec(mag, phase, mag_ref, phase_ref):
    X = np.exp2(mag) * np.exp(1j * phase)
    Xr = np.exp2(mag_ref) * np.exp(1j * phase_ref)
    assert np.abs(X - Xr).max() < ABS_TOL
    big = np.abs(Xr) > 1e-2
    assert np.abs(mag - mag_ref)[big].max() < 1e-4
    assert np.abs(np.angle(np.exp(1j * (phase - phase_ref))))[big].max() < 1e-4


@pytest.mark.parametrize("tag", ["48k", "16k", "nfft2048"])
def test_golden_vectors(golden_dir, tag):
    stft = _ops()
    g = np.load(os.path.join(golden_dir, "stft.npz"))
    n_fft, hop, win = (int(v) for v in g[f"{tag}_params"])
    wave = torch.from_numpy(g[f"{tag}_wave"]).cuda()
    mag, phase = stft.wav2spectro(wave, n_fft, hop, win, "log2")
    assert mag.shape == g[f"{tag}_mag"].shape
    _check_spec(mag.double().cpu().numpy(), phase.double().cpu().numpy(), g[f"{tag}_mag"].astype(np.float64),
                g[f"{tag}_phase"].astype(np.float64))
    mag2 = torch.from_numpy(g[f"{tag}_mag2"]).cuda().requires_grad_()
    phase2 = torch.from_numpy(g[f"{tag}_phase2"]).cuda().requires_grad_()
    wav2 = stft.spectro2wav(mag2, phase2, n_fft, hop, win, "log2")
    assert wav2.shape == g[f"{tag}_wav2"].shape
    assert np.abs(wav2.detach().cpu().numpy() - g[f"{tag}_wav2"]).max() < ABS_TOL
    wav2.backward(torch.from_numpy(g[f"{tag}_gw"]).cuda())
    for got, key in ((mag2.grad, "dmag2"), (phase2.grad, "dphase2")):
        ref = g[f"{tag}_{key}"].astype(np.float64)
        assert np.abs(got.double().cpu().numpy() - ref).max() / np.abs(ref).max() < 1e-4


CASES = [
    # (B, T, n_fft, hop, win)
    (4, 122640, 1024, 240, 1024),   # 48 kHz configs
    (4, 40880, 1024, 80, 1024),     # 16 kHz config
    (2, 122640, 2048, 240, 1024),   # vm_asr_48k_16k_nfft2048
    (3, 5000, 512, 120, 480),       # MultiResolutionSTFTLoss-like, win < n_fft, ragged frame count
    (1, 2000, 256, 64, 256),
    (2, 700, 128, 50, 100),
    (1, 1111, 64, 16, 64),
]


@pytest.mark.parametrize("B,T,n_fft,hop,win", CASES)
def test_vs_float64_oracle(B, T, n_fft, hop, win):
    stft = _ops()
    g = torch.Generator().manual_seed(T + n_fft)
    wave = 0.1 * torch.randn(B, 1, T, generator=g)
    mag_ref, phase_ref = stft_ref.wav2spectro(wave, n_fft, hop, win)
    mag, phase = stft.wav2spectro(wave.cuda(), n_fft, hop, win, "log2")
    assert mag.shape == mag_ref.shape == (B, 1, n_fft // 2 + 1, 1 + T // hop)
    _check_spec(mag.double().cpu().numpy(), phase.double().cpu().numpy(), mag_ref.numpy(), phase_ref.numpy())
    # inverse on a perturbed spectrogram (not the STFT of any signal), with gradients
    mag2 = (mag_ref + 0.2 * torch.randn(mag_ref.shape, generator=g, dtype=torch.float64)).float()
    phase2 = (phase_ref + 0.2 * torch.randn(mag_ref.shape, generator=g, dtype=torch.float64)).float()
    m64 = mag2.double().requires_grad_()
    p64 = phase2.double().requires_grad_()
    wav_ref = stft_ref.spectro2wav(m64, p64, n_fft, hop, win)
    gw = torch.randn(wav_ref.shape, generator=g, dtype=torch.float64)
    wav_ref.backward(gw)
    mg = mag2.cuda().requires_grad_()
    pg = phase2.cuda().requires_grad_()
    wav = stft.spectro2wav(mg, pg, n_fft, hop, win, "log2")
    assert wav.shape == wav_ref.shape
    assert np.abs(wav.detach().double().cpu().numpy() - wav_ref.detach().numpy()).max() < ABS_TOL
    wav.backward(gw.float().cuda())
    for got, ref in ((mg.grad, m64.grad), (pg.grad, p64.grad)):
        assert np.abs(got.double().cpu().numpy() - ref.numpy()).max() / ref.abs().max().item() < 1e-4


@pytest.mark.parametrize("T,n_fft,hop", [(122640, 1024, 240), (40880, 1024, 80), (122640, 2048, 240)])
def test_round_trip_full_size(T, n_fft, hop):
    """spectro2wav(wav2spectro(x)) == x on the configs' clip lengths (T is a multiple of hop)."""
    stft = _ops()
    wave = 0.1 * torch.randn(4, 1, T, device="cuda")
    mag, phase = stft.wav2spectro(wave, n_fft, hop, 1024, "log2")
    back = stft.spectro2wav(mag, phase, n_fft, hop, 1024, "log2")
    assert back.shape == wave.shape
    assert (back - wave).abs().max().item() < ABS_TOL


def test_deterministic():
    stft = _ops()
    wave = 0.1 * torch.randn(2, 1, 24000, device="cuda")
    mag, phase = stft.wav2spectro(wave, 1024, 240, 1024, "log2")
    a = stft.spectro2wav(mag, phase, 1024, 240, 1024, "log2")
    b = stft.spectro2wav(mag, phase, 1024, 240, 1024, "log2")
    assert torch.equal(a, b)


def test_errors():
    stft = _ops()
    with pytest.raises(RuntimeError):
        stft.wav2spectro(torch.randn(1, 1, 4000), 1024, 240, 1024, "log2")      # CPU tensor
    with pytest.raises(RuntimeError):
        stft.wav2spectro(torch.randn(1, 1, 4000, device="cuda"), 1000, 240, 1000, "log2")  # not a power of two
    with pytest.raises(RuntimeError):
        stft.wav2spectro(torch.randn(1, 1, 300, device="cuda"), 1024, 240, 1024, "log2")   # shorter than the reflect pad
    with pytest.raises(NotImplementedError):
        stft.wav2spectro(torch.randn(1, 1, 4000, device="cuda"), 1024, 240, 1024, "dB")
