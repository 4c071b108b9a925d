"""GPU parity: STFT / iSTFT kernels and their backward passes, the linear-magnitude STFT, the multi-resolution STFT loss and
the LSD metrics against the float64 oracle and the reference's own outputs (tests/golden/stft.npz, stft_loss.npz, produced by
utils/stft.py, model/loss.py, model/metric.py).

Tolerance: 1e-5 absolute (BASELINE.json) on the complex spectrum X and on the waveform.  The two tensors wav2spectro returns
are functions of X that are ill-conditioned at |X| -> 0, and the test states that conditioning instead of hiding it:
    mag = log2(|X| + 1e-8):  |d mag| <= |dX| / (|X| ln 2)          phase = angle(X):  |d phase| <= |dX| / |X|
(first order; the factor 1.5 below covers the second-order term for |dX| << |X|).  So with e_X = max |X - X_ref| the bars are
    |mag - mag_ref| <= 1e-5 + 1.5 e_X / (|X_ref| ln 2),      |phase - phase_ref| <= 1e-5 + 1.5 e_X / |X_ref|   for EVERY bin,
and e_X itself must be below 1e-5: wherever |X| >= 0.05 (bins that carry signal) that is 1e-5 + O(1e-6)."""
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
    e_x = np.abs(X - Xr).max()
    assert e_x < ABS_TOL
    r = np.maximum(np.abs(Xr), 1e-30)
    d_mag = np.abs(mag - mag_ref)
    d_ph = np.abs(np.angle(np.exp(1j * (phase - phase_ref))))
    tiny = np.abs(Xr) < 1e-6   # below the 1e-8 offset's reach both are noise in the reference too (log2(1e-8 + .), angle(~0))
    assert (d_mag <= ABS_TOL + 1.5 * e_x / (r * np.log(2.0)))[~tiny].all(), float(d_mag[~tiny].max())
    assert (d_ph <= ABS_TOL + 1.5 * e_x / r)[~tiny].all(), float(d_ph[~tiny].max())
    strong = np.abs(Xr) >= 0.05
    if strong.any():
        assert d_mag[strong].max() < 2 * ABS_TOL and d_ph[strong].max() < 2 * ABS_TOL


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
    with pytest.raises(ValueError):
        stft.spectro2wav(torch.randn(1, 1, 513, 8, device="cuda"), torch.randn(1, 1, 513, 8, device="cuda"), 1024, 240, 1024, "mel")


# ---------------------------------------------------------------------------------------------------------
# backward of wav2spectro, dB scale, linear-magnitude STFT, multi-resolution STFT loss, LSD
# ---------------------------------------------------------------------------------------------------------
def _gl(golden_dir):
    return np.load(os.path.join(golden_dir, "stft_loss.npz"))


@pytest.mark.parametrize("tag", ["48k", "nfft2048", "small"])
def test_wav2spectro_backward_golden(golden_dir, tag):
    """d wave of wav2spectro (vmasr_stft_bwd) against autograd through the reference's own wav2spectro."""
    stft = _ops()
    g = _gl(golden_dir)
    n_fft, hop, win = (int(v) for v in g[f"bwd_{tag}_params"])
    for which, key in (("mag", "dwave_mag"), ("phase", "dwave_phase")):
        w = torch.from_numpy(g[f"bwd_{tag}_wave"]).cuda().requires_grad_()
        mag, phase = stft.wav2spectro(w, n_fft, hop, win, "log2")
        if which == "mag":
            (mag * torch.from_numpy(g[f"bwd_{tag}_gm"]).cuda()).sum().backward()
        else:
            (phase * torch.from_numpy(g[f"bwd_{tag}_gp"]).cuda()).sum().backward()
        ref = g[f"bwd_{tag}_{key}"].astype(np.float64)
        err = np.abs(w.grad.double().cpu().numpy() - ref).max() / np.abs(ref).max()
        # both gradients are dominated by near-zero bins (d mag / d X, d phase / d X ~ 1 / |X|): fp32 reference vs fp32 kernel;
        # the well-conditioned comparison against the float64 oracle is test_wav2spectro_backward_vs_float64_oracle
        assert err < 2e-3, (which, err)


@pytest.mark.parametrize("B,T,n_fft,hop,win", [(2, 24000, 1024, 240, 1024), (1, 12000, 2048, 240, 1024), (2, 3000, 256, 64, 200)])
def test_wav2spectro_backward_vs_float64_oracle(B, T, n_fft, hop, win):
    """d wave of wav2spectro against float64 autograd through the oracle, with cotangents weighted by |X|^2 / (|X|^2 + c) so
    that the ill-conditioned near-zero bins (gradient ~ 1 / |X|) do not dominate: 1e-4 of the largest gradient entry."""
    stft = _ops()
    g = torch.Generator().manual_seed(T)
    wave = 0.1 * torch.randn(B, 1, T, generator=g)
    w64 = wave.double().requires_grad_()
    m64, p64 = stft_ref.wav2spectro(w64, n_fft, hop, win)
    pw = torch.exp2(2 * m64.detach())
    wgt = pw / (pw + 0.05 * pw.mean())
    gm = torch.randn(m64.shape, generator=g, dtype=torch.float64) * wgt
    gp = torch.randn(m64.shape, generator=g, dtype=torch.float64) * wgt
    ((m64 * gm).sum() + (p64 * gp).sum()).backward()
    wg = wave.cuda().requires_grad_()
    mag, phase = stft.wav2spectro(wg, n_fft, hop, win, "log2")
    ((mag * gm.float().cuda()).sum() + (phase * gp.float().cuda()).sum()).backward()
    err = (wg.grad.double().cpu() - w64.grad).abs().max().item() / w64.grad.abs().max().item()
    assert err < 1e-4, err


@pytest.mark.parametrize("B,T,n_fft,hop,win,normalized,clamp", [
    (2, 4800, 1024, 120, 600, False, 1e-7), (2, 4800, 2048, 240, 1200, False, 1e-7), (2, 4800, 512, 50, 240, False, 1e-7),
    (1, 30000, 2048, 512, 2048, False, 0.0), (3, 3000, 256, 64, 256, True, 0.0), (4, 122640, 1024, 240, 1024, True, 0.0),
])
def test_stft_magnitude_forward_backward(B, T, n_fft, hop, win, normalized, clamp):
    """Linear-magnitude STFT and its backward (vmasr_stft_mag_fwd / _bwd) against the float64 oracle through autograd, on the
    three resolutions of the loss, the metric's transform and a config-size clip."""
    stft = _ops()
    g = torch.Generator().manual_seed(n_fft + hop)
    wave = 0.1 * torch.randn(B, T, generator=g)
    w64 = wave.double().requires_grad_()
    ref = stft_ref.stft_magnitude(w64, n_fft, hop, win, normalized, clamp)
    cot = torch.randn(ref.shape, generator=g, dtype=torch.float64)
    (ref * cot).sum().backward()
    wg = wave.cuda().requires_grad_()
    got = stft.stft_magnitude(wg, n_fft, hop, win, normalized, clamp)
    assert got.shape == ref.shape
    scale = ref.detach().abs().max().item()
    assert (got.detach().double().cpu() - ref.detach()).abs().max().item() < 1e-5 * max(scale, 1.0)
    (got * cot.float().cuda()).sum().backward()
    err = (wg.grad.double().cpu() - w64.grad).abs().max().item() / w64.grad.abs().max().item()
    assert err < 1e-4, err


def test_multi_resolution_stft_loss_golden(golden_dir):
    """vm_asr_b200.loss.MultiResolutionSTFTLoss against the reference's own (model/loss.py:137-184): both terms and the
    gradient with respect to the prediction."""
    from vm_asr_b200 import loss
    g = _gl(golden_dir)
    x = torch.from_numpy(g["mr_x"]).cuda().requires_grad_()
    y = torch.from_numpy(g["mr_y"]).cuda()
    sc, mg = loss.MultiResolutionSTFTLoss()(x, y)
    assert abs(sc.item() - float(g["mr_sc"])) < 1e-5 * max(1.0, abs(float(g["mr_sc"])))
    assert abs(mg.item() - float(g["mr_mag"])) < 1e-5 * max(1.0, abs(float(g["mr_mag"])))
    (sc + mg).backward()
    ref = g["mr_dx"].astype(np.float64)
    assert np.abs(x.grad.double().cpu().numpy() - ref).max() / np.abs(ref).max() < 2e-4


def test_lsd_metrics_golden(golden_dir):
    from vm_asr_b200 import loss
    g = _gl(golden_dir)
    x, y = torch.from_numpy(g["mr_x"]).cuda(), torch.from_numpy(g["mr_y"]).cuda()
    hf = torch.from_numpy(g["lsd_hf_idx"])
    assert abs(loss.lsd(x, y) - float(g["lsd"])) < 1e-4
    assert abs(loss.lsd_hf(x, y, hf) - float(g["lsd_hf"])) < 1e-4
    assert abs(loss.lsd_lf(x, y, hf) - float(g["lsd_lf"])) < 1e-4


def test_db_scale_golden(golden_dir):
    """spectro_scale="dB" (utils/stft.py:59-62, 100-102; not used by the shipped configs)."""
    stft = _ops()
    g = _gl(golden_dir)
    wave = torch.from_numpy(g["db_wave"]).cuda()
    mag, phase = stft.wav2spectro(wave, 1024, 240, 1024, "dB")
    assert mag.shape == g["db_mag"].shape
    assert np.abs(mag.cpu().numpy() - g["db_mag"]).max() < 2e-3      # dB of a power: 10 log10, fp32 reference
    back = stft.spectro2wav(torch.from_numpy(g["db_mag"]).cuda(), torch.from_numpy(g["db_phase"]).cuda(), 1024, 240, 1024, "dB")
    assert np.abs(back.cpu().numpy() - g["db_back"]).max() < ABS_TOL


@pytest.mark.parametrize("n_fft,hop,win", [(1024, 240, 1024), (512, 50, 240), (2048, 1024, 2048), (256, 100, 256)])
def test_synthesis_is_deterministic(n_fft, hop, win):
    """iSTFT and the STFT backward add into the padded accumulator with red.add: at most two CTAs per sample, so bit-identical
    run to run -- also for hop << n_fft (several rounds of frames per CTA) and hop = n_fft / 2 (one round, little overlap)."""
    stft = _ops()
    T = hop * 101
    wave = (0.1 * torch.randn(3, T, device="cuda")).requires_grad_()
    outs = []
    for _ in range(3):
        wave.grad = None
        mag, phase = stft.wav2spectro(wave, n_fft, hop, win, "log2")
        back = stft.spectro2wav(mag, phase, n_fft, hop, win, "log2")
        back.square().sum().backward()
        outs.append((back.detach().clone(), wave.grad.clone()))
    for b, gr in outs[1:]:
        assert torch.equal(b, outs[0][0]) and torch.equal(gr, outs[0][1])
