"""CPU: the oracle restatements (torch and C) against the golden vectors made from the reference's own code
(oracle/make_golden.py).  This is what pins the oracle; the GPU parity tests then compare CUDA to the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import c_ref, ss2d_ref, stft_ref

CROSS_TAGS = ["sq", "rect", "odd"]
SCAN_TAGS = ["n1_full", "n1_nobias", "n1_long", "n2_g1", "n4_nosp"]
STFT_TAGS = ["48k", "16k", "nfft2048"]


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize("tag", CROSS_TAGS)
def test_cross_scan_merge_bit_exact(golden_dir, tag):
    g = _load(golden_dir, "cross_scan_merge.npz")
    x, xs, gxs, gx = (_t(g[f"{tag}_{k}"]) for k in ("x", "xs", "gxs", "gx"))
    ys, y, gy, gys = (_t(g[f"{tag}_{k}"]) for k in ("ys", "y", "gy", "gys"))
    B, C, H, W = x.shape
    assert torch.equal(ss2d_ref.cross_scan(x), xs)
    assert torch.equal(ss2d_ref.cross_scan_bwd(gxs, H, W), gx)
    assert torch.equal(ss2d_ref.cross_merge(ys), y)
    assert torch.equal(ss2d_ref.cross_merge_bwd(gy, H, W), gys)
    # C restatement
    assert np.array_equal(c_ref.cross_scan(x.numpy()), xs.numpy())
    assert np.array_equal(c_ref.cross_merge(ys.numpy().reshape(B, 4, C, H * W), H, W), y.numpy())
    assert np.array_equal(c_ref.cross_merge(gxs.numpy(), H, W).reshape(B, C, H, W), gx.numpy())


def _scan_case(g, tag):
    def opt(k):
        return _t(g[f"{tag}_{k}"]) if f"{tag}_{k}" in g.files else None

    return dict(u=_t(g[f"{tag}_u"]), delta=_t(g[f"{tag}_delta"]), A=_t(g[f"{tag}_A"]), B=_t(g[f"{tag}_B"]),
                C=_t(g[f"{tag}_C"]), D=opt("D"), bias=opt("bias"), sp=bool(g[f"{tag}_softplus"]))


@pytest.mark.parametrize("tag", SCAN_TAGS)
def test_selective_scan_forward(golden_dir, tag):
    g = _load(golden_dir, "selective_scan.npz")
    c = _scan_case(g, tag)
    out, last = ss2d_ref.selective_scan(c["u"], c["delta"], c["A"], c["B"], c["C"], c["D"], c["bias"], c["sp"],
                                        return_last_state=True)
    # same fp32 sequential arithmetic, different op grouping: a few ulp
    assert torch.allclose(out, _t(g[f"{tag}_out"]), rtol=1e-5, atol=1e-5)
    assert torch.allclose(last, _t(g[f"{tag}_last"]), rtol=1e-5, atol=1e-5)
    out64, last64, _ = c_ref.scan_fwd(c["u"], c["delta"], c["A"], c["B"], c["C"], c["D"], c["bias"], c["sp"])
    assert np.allclose(out64, g[f"{tag}_out"], rtol=2e-5, atol=2e-5)
    assert np.allclose(last64, g[f"{tag}_last"], rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("tag", SCAN_TAGS)
def test_selective_scan_backward(golden_dir, tag):
    g = _load(golden_dir, "selective_scan.npz")
    c = _scan_case(g, tag)
    gout = _t(g[f"{tag}_gout"])
    names = ["du", "ddelta", "dA", "dB", "dC", "dD", "dbias"]
    got_t = ss2d_ref.selective_scan_bwd(c["u"], c["delta"], c["A"], c["B"], c["C"], c["D"], c["bias"], c["sp"], gout)
    got_c = c_ref.scan_bwd(c["u"], c["delta"], c["A"], c["B"], c["C"], c["D"], c["bias"], c["sp"], gout)
    for name, a, b in zip(names, got_t, got_c):
        key = f"{tag}_{name}"
        if key not in g.files:
            assert a is None and b is None
            continue
        ref = g[key].astype(np.float64)
        scale = max(np.abs(ref).max(), 1e-6)
        # the golden gradients are fp32 autograd; ours are fp64 closed form
        assert np.abs(a.numpy() - ref).max() / scale < 2e-5, name
        assert np.abs(b - ref).max() / scale < 2e-5, name


def test_chunk_state_layout(golden_dir):
    """(cumulative decay, state) at every chunk end; last chunk == last state (test_selective_scan.py:114)."""
    g = _load(golden_dir, "selective_scan.npz")
    c = _scan_case(g, "n1_long")
    _, last, cs = c_ref.scan_fwd(c["u"], c["delta"], c["A"], c["B"], c["C"], c["D"], c["bias"], c["sp"], chunk=128)
    assert cs.shape == (1, 4, 3, 2)
    assert np.allclose(cs[:, :, -1, 1::2], last)


@pytest.mark.parametrize("tag", STFT_TAGS)
def test_stft_istft(golden_dir, tag):
    g = _load(golden_dir, "stft.npz")
    n_fft, hop, win = (int(v) for v in g[f"{tag}_params"])
    wave = _t(g[f"{tag}_wave"])
    mag, phase = stft_ref.wav2spectro(wave, n_fft, hop, win)
    gm, gp = g[f"{tag}_mag"].astype(np.float64), g[f"{tag}_phase"].astype(np.float64)
    assert mag.shape == gm.shape
    # compare as complex spectra (robust where |X| is tiny), then mag/phase where |X| is not tiny
    X = np.exp2(mag.numpy()) * np.exp(1j * phase.numpy())
    Xg = np.exp2(gm) * np.exp(1j * gp)
    assert np.abs(X - Xg).max() < 1e-5
    big = np.abs(Xg) > 1e-2
    assert np.abs(mag.numpy() - gm)[big].max() < 1e-4
    dphi = np.angle(np.exp(1j * (phase.numpy() - gp)))
    assert np.abs(dphi)[big].max() < 1e-4
    # inverse
    back = stft_ref.spectro2wav(_t(g[f"{tag}_mag2"]), _t(g[f"{tag}_phase2"]), n_fft, hop, win)
    assert back.shape == g[f"{tag}_wav2"].shape
    assert np.abs(back.numpy() - g[f"{tag}_wav2"]).max() < 1e-5
    # round trip of the reference itself is the identity (length hop*(frames-1))
    assert np.abs(g[f"{tag}_back"] - g[f"{tag}_wave"]).max() < 1e-5


@pytest.mark.parametrize("tag", STFT_TAGS)
def test_istft_gradient(golden_dir, tag):
    """autograd through the float64 restatement == the reference's autograd through torch.istft."""
    g = _load(golden_dir, "stft.npz")
    n_fft, hop, win = (int(v) for v in g[f"{tag}_params"])
    mag = _t(g[f"{tag}_mag2"]).double().requires_grad_()
    phase = _t(g[f"{tag}_phase2"]).double().requires_grad_()
    wav = stft_ref.spectro2wav(mag, phase, n_fft, hop, win)
    wav.backward(_t(g[f"{tag}_gw"]).double())
    for got, key in ((mag.grad, "dmag2"), (phase.grad, "dphase2")):
        ref = g[f"{tag}_{key}"].astype(np.float64)
        assert np.abs(got.numpy() - ref).max() / np.abs(ref).max() < 1e-5


def test_ss2d_core_chain_shapes():
    torch.manual_seed(0)
    B, C, H, W, N, R = 2, 4, 6, 5, 1, 1
    x = torch.randn(B, C, H, W)
    y = ss2d_ref.ss2d_core(x, torch.randn(4, R + 2 * N, C), torch.randn(4, C, R), torch.rand(4, C),
                           torch.zeros(4 * C, N), torch.ones(4 * C))
    assert y.shape == (B, C, H * W) and torch.isfinite(y).all()


def test_fused_core_plan_is_algebraically_the_reference_chain():
    """The storage-order formulation the fused SS2D kernel is planned on (no xs / ys copies, reversed
    directions read back to front, two output planes merged by one transpose-add) equals the reference's chain
    (ss2d_core, pinned above) -- non-square map, float64 accumulation."""
    torch.manual_seed(7)
    Bsz, C, H, W, N, R = 2, 6, 9, 14, 1, 2
    x = torch.randn(Bsz, C, H, W, dtype=torch.float64)
    xw, dw, db = torch.randn(4, R + 2 * N, C, dtype=torch.float64) * 0.3, torch.randn(4, C, R, dtype=torch.float64) * 0.3, torch.rand(4, C) * 0.5
    A_logs, Ds = torch.log(torch.rand(4 * C, N) + 0.5), torch.randn(4 * C)
    ref = ss2d_ref.ss2d_core(x, xw, dw, db, A_logs, Ds, dtype=torch.float64)
    got = ss2d_ref.ss2d_core_storage_order(x, xw, dw, db, A_logs, Ds, dtype=torch.float64)
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() < 1e-5 * ref.abs().max().item()   # the scan inputs pass through float32 on both sides


def test_stft_loss_oracle_against_reference(golden_dir):
    """oracle/stft_ref.py's multi-resolution STFT loss, LSD, dB scale and the gradient of wav2spectro (autograd through the
    restatement) against the reference's own model/loss.py, model/metric.py, utils/stft.py (tests/golden/stft_loss.npz)."""
    import os
    import numpy as np
    import torch
    from oracle import stft_ref
    g = np.load(os.path.join(golden_dir, "stft_loss.npz"))
    x = torch.from_numpy(g["mr_x"]).double().requires_grad_()
    y = torch.from_numpy(g["mr_y"]).double()
    sc, mg = stft_ref.multi_resolution_stft_loss(x, y)
    assert abs(sc.item() - float(g["mr_sc"])) < 1e-5 and abs(mg.item() - float(g["mr_mag"])) < 1e-5
    (sc + mg).backward()
    ref = g["mr_dx"].astype(np.float64)
    assert np.abs(x.grad.numpy() - ref).max() / np.abs(ref).max() < 1e-4
    assert abs(stft_ref.lsd(x.detach(), y).item() - float(g["lsd"])) < 1e-4
    mag, phase = stft_ref.wav2spectro(torch.from_numpy(g["db_wave"]), 1024, 240, 1024, "dB")
    assert np.abs(mag.numpy() - g["db_mag"]).max() < 1e-3
    back = stft_ref.spectro2wav(torch.from_numpy(g["db_mag"]), torch.from_numpy(g["db_phase"]), 1024, 240, 1024, "dB")
    assert np.abs(back.numpy() - g["db_back"]).max() < 1e-5
    for tag in ("48k", "nfft2048", "small"):
        n_fft, hop, win = (int(v) for v in g[f"bwd_{tag}_params"])
        w = torch.from_numpy(g[f"bwd_{tag}_wave"]).double().requires_grad_()
        m, p = stft_ref.wav2spectro(w, n_fft, hop, win)
        (m * torch.from_numpy(g[f"bwd_{tag}_gm"]).double()).sum().backward()
        ref = g[f"bwd_{tag}_dwave_mag"].astype(np.float64)
        # d mag / d X ~ 1 / |X|: the near-zero bins of a noise signal dominate this gradient, and the golden is fp32
        assert np.abs(w.grad.numpy() - ref).max() / np.abs(ref).max() < 1e-3
