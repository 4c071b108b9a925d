"""Generate ``tests/golden/*.npz`` by running the REFERENCE's own code on seeded inputs.

Run once in the build container (``/root/reference`` is mounted there and nowhere else):

    python oracle/make_golden.py

The fixtures are committed; nothing at test/bench time reads ``/root/reference``.

What is executed, unmodified, from the reference tree:
  * ``model/vmamba.py``  ``CrossScan`` / ``CrossMerge`` (forward and autograd backward), lines 27-73.
    The module imports ``timm`` and ``fvcore`` at the top (absent in this image) -- two stub modules are
    installed in ``sys.modules`` first; they are never called on this path.
  * ``kernels/selective_scan/test_selective_scan.py``  ``selective_scan_ref`` (lines 287-367), pulled out
    of the file by AST because the module itself imports a CUDA extension and runs a test at import.
    Gradients come from torch autograd through that function.
  * ``utils/stft.py``  ``wav2spectro`` / ``spectro2wav`` (log2 and dB scales; the gradient of wav2spectro by autograd).
  * ``model/loss.py``  ``MultiResolutionSTFTLoss`` (value and gradient), ``model/metric.py``  ``lsd`` / ``lsd_hf`` / ``lsd_lf``.
  * ``utils/post_processing.py``  ``unfold_audio`` / ``fold_audio``.
"""
from __future__ import annotations

import ast
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F
from einops import rearrange, repeat

REF = os.environ.get("VMASR_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _install_stubs():
    timm = types.ModuleType("timm")
    timm_models = types.ModuleType("timm.models")
    timm_layers = types.ModuleType("timm.models.layers")

    class DropPath(torch.nn.Module):
        def __init__(self, drop_prob=0.0):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            return x

    timm_layers.DropPath = DropPath
    timm_layers.trunc_normal_ = torch.nn.init.trunc_normal_
    timm.models = timm_models
    timm_models.layers = timm_layers
    fvcore = types.ModuleType("fvcore")
    fvcore_nn = types.ModuleType("fvcore.nn")
    for name in ("FlopCountAnalysis", "flop_count_str", "flop_count", "parameter_count"):
        setattr(fvcore_nn, name, None)
    fvcore.nn = fvcore_nn
    sys.modules.update({
        "timm": timm, "timm.models": timm_models, "timm.models.layers": timm_layers,
        "fvcore": fvcore, "fvcore.nn": fvcore_nn,
    })


def load_reference():
    _install_stubs()
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "model"))
    import importlib

    vmamba = importlib.import_module("vmamba")  # model/vmamba.py as a top-level module: model/__init__.py pulls in the trainer stack
    stft = importlib.import_module("utils.stft")
    src = open(os.path.join(REF, "kernels/selective_scan/test_selective_scan.py")).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "selective_scan_ref")
    scope = {"torch": torch, "F": F, "rearrange": rearrange, "repeat": repeat}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "selective_scan_ref", "exec"), scope)
    return vmamba, stft, scope["selective_scan_ref"]


def np_(t):
    return None if t is None else t.detach().cpu().numpy()


def main():
    os.makedirs(OUT, exist_ok=True)
    vmamba, stft, selective_scan_ref = load_reference()

    # ---------------- cross scan / merge -------------------------------------------------------
    cases = {"sq": (2, 3, 8, 8), "rect": (1, 2, 6, 10), "odd": (2, 2, 5, 7)}
    blob = {}
    for tag, (B, C, H, W) in cases.items():
        g = torch.Generator().manual_seed(100 + H * W)
        x = torch.randn(B, C, H, W, generator=g, requires_grad=True)
        xs = vmamba.CrossScan.apply(x)
        gxs = torch.randn(xs.shape, generator=g)
        xs.backward(gxs)
        ys = torch.randn(B, 4, C, H, W, generator=g, requires_grad=True)
        y = vmamba.CrossMerge.apply(ys)
        gy = torch.randn(y.shape, generator=g)
        y.backward(gy)
        blob.update({
            f"{tag}_x": np_(x), f"{tag}_xs": np_(xs), f"{tag}_gxs": np_(gxs), f"{tag}_gx": np_(x.grad),
            f"{tag}_ys": np_(ys), f"{tag}_y": np_(y), f"{tag}_gy": np_(gy), f"{tag}_gys": np_(ys.grad),
        })
    np.savez_compressed(os.path.join(OUT, "cross_scan_merge.npz"), **blob)

    # ---------------- selective scan -----------------------------------------------------------
    # distributions of test_selective_scan.py:593-654
    scan_cases = {
        # tag: (B, D, G, N, L, has_D, has_bias, softplus)
        "n1_full": (2, 8, 4, 1, 67, True, True, True),
        "n1_nobias": (2, 4, 2, 1, 64, False, False, False),
        "n1_long": (1, 4, 4, 1, 300, True, True, True),
        "n2_g1": (2, 6, 1, 2, 33, True, True, True),
        "n4_nosp": (1, 4, 2, 4, 40, True, False, False),
    }
    blob = {}
    for tag, (Bsz, Dm, G, N, L, has_D, has_bias, sp) in scan_cases.items():
        torch.random.manual_seed(0)
        A = (-0.5 * torch.rand(Dm, N)).requires_grad_()
        Bm = torch.randn(Bsz, G, N, L, requires_grad=True)
        Cm = torch.randn(Bsz, G, N, L, requires_grad=True)
        Dv = torch.randn(Dm, requires_grad=True) if has_D else None
        bias = (0.5 * torch.rand(Dm)).requires_grad_() if has_bias else None
        u = torch.randn(Bsz, Dm, L, requires_grad=True)
        delta = (0.5 * torch.rand(Bsz, Dm, L)).requires_grad_()
        if tag == "n1_long":  # exercise the softplus identity branch (> 20) and strongly negative inputs
            with torch.no_grad():
                delta[0, 0, :8] = 25.0
                delta[0, 1, :8] = -12.0
        out, last = selective_scan_ref(u, delta, A, Bm, Cm, Dv, delta_bias=bias, delta_softplus=sp,
                                       return_last_state=True)
        gout = torch.randn_like(out)
        out.backward(gout)
        blob.update({
            f"{tag}_u": np_(u), f"{tag}_delta": np_(delta), f"{tag}_A": np_(A), f"{tag}_B": np_(Bm),
            f"{tag}_C": np_(Cm), f"{tag}_out": np_(out), f"{tag}_last": np_(last), f"{tag}_gout": np_(gout),
            f"{tag}_du": np_(u.grad), f"{tag}_ddelta": np_(delta.grad), f"{tag}_dA": np_(A.grad),
            f"{tag}_dB": np_(Bm.grad), f"{tag}_dC": np_(Cm.grad),
            f"{tag}_softplus": np.array(sp),
        })
        if has_D:
            blob[f"{tag}_D"] = np_(Dv)
            blob[f"{tag}_dD"] = np_(Dv.grad)
        if has_bias:
            blob[f"{tag}_bias"] = np_(bias)
            blob[f"{tag}_dbias"] = np_(bias.grad)
    np.savez_compressed(os.path.join(OUT, "selective_scan.npz"), **blob)

    # ---------------- STFT / iSTFT -------------------------------------------------------------
    stft_cases = {
        # tag: (B, T, n_fft, hop, win)
        "48k": (2, 240 * 8, 1024, 240, 1024),
        "16k": (1, 80 * 20, 1024, 80, 1024),
        "nfft2048": (1, 240 * 9, 2048, 240, 1024),
    }
    blob = {}
    for tag, (Bsz, T, n_fft, hop, win) in stft_cases.items():
        g = torch.Generator().manual_seed(7 + n_fft + hop)
        wave = 0.1 * torch.randn(Bsz, 1, T, generator=g)
        mag, phase = stft.wav2spectro(wave, n_fft, hop, win, "log2")
        back = stft.spectro2wav(mag, phase, n_fft, hop, win, "log2")
        # an un-related spectrogram (not the STFT of anything) through the inverse
        mag2 = mag + 0.3 * torch.randn(mag.shape, generator=g)
        phase2 = phase + 0.3 * torch.randn(phase.shape, generator=g)
        mag2.requires_grad_()
        phase2.requires_grad_()
        wav2 = stft.spectro2wav(mag2, phase2, n_fft, hop, win, "log2")
        gw = torch.randn(wav2.shape, generator=g)
        wav2.backward(gw)
        blob.update({
            f"{tag}_wave": np_(wave), f"{tag}_mag": np_(mag), f"{tag}_phase": np_(phase),
            f"{tag}_back": np_(back), f"{tag}_mag2": np_(mag2), f"{tag}_phase2": np_(phase2),
            f"{tag}_wav2": np_(wav2), f"{tag}_gw": np_(gw), f"{tag}_dmag2": np_(mag2.grad),
            f"{tag}_dphase2": np_(phase2.grad),
            f"{tag}_params": np.array([n_fft, hop, win]),
        })
    np.savez_compressed(os.path.join(OUT, "stft.npz"), **blob)

    # ---------------- STFT backward, dB scale, multi-resolution STFT loss, LSD ----------------------------------------
    # model/loss.py and model/metric.py import torch only; run unmodified
    import importlib
    ref_loss = importlib.import_module("loss")      # model/ is on sys.path
    ref_metric = importlib.import_module("metric")
    blob = {}
    g = torch.Generator().manual_seed(2024)
    T = 4800
    x = (0.1 * torch.randn(2, T, generator=g)).requires_grad_()
    y = 0.1 * torch.randn(2, T, generator=g)
    sc, mg = ref_loss.MultiResolutionSTFTLoss()(x, y)
    (sc + mg).backward()
    blob.update(mr_x=np_(x), mr_y=np_(y), mr_sc=np_(sc), mr_mag=np_(mg), mr_dx=np_(x.grad))
    blob["lsd"] = np.array(ref_metric.lsd(x.detach(), y))
    hf = torch.tensor([300, 500])
    blob["lsd_hf"] = np.array(ref_metric.lsd_hf(x.detach(), y, hf))
    blob["lsd_lf"] = np.array(ref_metric.lsd_lf(x.detach(), y, hf))
    blob["lsd_hf_idx"] = np_(hf)
    # wav2spectro is differentiable in the reference (torch.stft + abs / log2 / angle): gradient of a random cotangent
    for tag, (n_fft, hop, win) in {"48k": (1024, 240, 1024), "nfft2048": (2048, 240, 1024), "small": (256, 64, 200)}.items():
        w = (0.1 * torch.randn(2, 1, hop * 12, generator=g)).requires_grad_()
        mag, phase = stft.wav2spectro(w, n_fft, hop, win, "log2")
        gm, gp = torch.randn(mag.shape, generator=g), torch.randn(phase.shape, generator=g)
        (mag * gm).sum().backward(retain_graph=True)
        d_from_mag = w.grad.clone()
        w.grad = None
        (phase * gp).sum().backward()
        blob.update({f"bwd_{tag}_wave": np_(w), f"bwd_{tag}_gm": np_(gm), f"bwd_{tag}_gp": np_(gp),
                     f"bwd_{tag}_dwave_mag": np_(d_from_mag), f"bwd_{tag}_dwave_phase": np_(w.grad),
                     f"bwd_{tag}_params": np.array([n_fft, hop, win])})
    wdb = 0.1 * torch.randn(2, 1, 240 * 10, generator=g)
    mag_db, phase_db = stft.wav2spectro(wdb, 1024, 240, 1024, "dB")
    back_db = stft.spectro2wav(mag_db, phase_db, 1024, 240, 1024, "dB")
    blob.update(db_wave=np_(wdb), db_mag=np_(mag_db), db_phase=np_(phase_db), db_back=np_(back_db))
    np.savez_compressed(os.path.join(OUT, "stft_loss.npz"), **blob)

    # ---------------- overlapped segments (utils/post_processing.py, torch only) --------------------------------------------
    pp = importlib.import_module("utils.post_processing")
    audio = torch.randn(2, 1, 9000, generator=g)
    seg, ov = 2500, 300
    segs = pp.unfold_audio(audio, seg, ov)
    processed = torch.tanh(segs) + 0.1 * torch.randn(segs.shape, generator=g)
    folded = pp.fold_audio(processed, audio.shape[-1], seg, ov)
    np.savez_compressed(os.path.join(OUT, "segments.npz"), audio=np_(audio), segments=np_(segs.contiguous()), processed=np_(processed),
                        folded=np_(folded), params=np.array([seg, ov]))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
