#!/usr/bin/env python
"""The block's head at the config shapes: ONE kernel (vmasr_dwconv_silu_fwd / _bwd: permute + depthwise conv 3x3 + SiLU + float32
cast + transpose, reading the x half of in_proj's (B, H, W, 2C) output in place) against what the reference runs in front of the
scan (vmamba.py:1541-1546) plus the transpose the fused core needs: permute(0, 3, 1, 2).contiguous(), cuDNN depthwise conv2d,
SiLU, .float(), transpose -- forward and forward + backward, CUDA events around CUDA graphs over rotating buffer sets.
    python tools/head_bench.py [--workload vm_asr_48k_MPD] [--dtype float32|float16]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from bench import load_peaks
from tools.shape_bench import timeit
from vm_asr_b200 import ss2d, workload as W

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="vm_asr_48k_MPD")
ap.add_argument("--dtype", default="float32")
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--xproj", action="store_true", help="with the x_proj contraction inside the kernel, against the chain + the two einsums")
args = ap.parse_args()
wl = W.WORKLOADS[args.workload]
dt = getattr(torch, args.dtype)
es = 4 if dt == torch.float32 else 2
peak, _ = load_peaks()
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(0)
B = wl.batch
for call, count in W.distinct_shapes(wl):
    C, H, Wd, L = call.d_inner, call.H, call.W, call.L
    n = B * C * L
    n_sets = max(2, min(6, int(300e6 // (n * 16)) + 1))
    sets = [dict(xz=torch.randn(B, H, Wd, 2 * C, device=dev, generator=gen).to(dt), gx=torch.randn(B, C, H, Wd, device=dev, generator=gen),
                 gxT=torch.randn(B, C, Wd, H, device=dev, generator=gen)) for _ in range(n_sets)]
    wt, bs = 0.3 * torch.randn(C, 1, 3, 3, device=dev), torch.zeros(C, device=dev)
    R = max(1, -(-(C // 2) // 16))
    xw = 0.3 * torch.randn(4, R + 2, C, device=dev)
    gxd = [torch.randn(B, 2, R + 2, L, device=dev) for _ in range(2)]

    def fused(i, grad=False):
        d = sets[i]
        ins = [d["xz"].requires_grad_(grad), wt.requires_grad_(grad), bs.requires_grad_(grad)]
        if args.xproj:
            ins.append(xw.requires_grad_(grad))
            return ss2d.ConvSiluInput.apply(ins[0][..., :C], ins[1], ins[2], ins[3]), ins, (d["gx"], d["gxT"], gxd[0], gxd[1])
        x, xT = ss2d.ConvSiluInput.apply(ins[0][..., :C], ins[1], ins[2])
        return (x, xT), ins, (d["gx"], d["gxT"])

    def chain(i, grad=False):
        d = sets[i]
        ins = [d["xz"].requires_grad_(grad), wt.requires_grad_(grad), bs.requires_grad_(grad)]
        x = ins[0][..., :C].permute(0, 3, 1, 2).contiguous()
        x = F.silu(F.conv2d(x, ins[1].to(dt), ins[2].to(dt), padding=1, groups=C)).float()
        xT = ss2d.MapTranspose.apply(x)
        if args.xproj:
            ins.append(xw.requires_grad_(grad))
            xd = [torch.einsum("bdl,kcd->bkcl", src.view(B, C, L), ins[3][par::2]) for par, src in ((0, x), (1, xT))]
            return (x, xT, xd[0], xd[1]), ins, (d["gx"], d["gxT"], gxd[0], gxd[1])
        return (x, xT), ins, (d["gx"], d["gxT"])

    def fwd_of(f):
        def run(i):
            with torch.no_grad():
                f(i)
        return run

    def both_of(f):
        def run(i):
            o, ins, g = f(i, True)
            torch.autograd.grad(o, ins, g)
        return run

    fwd_bytes, bwd_bytes = n * (es + 8), n * (es + 8 + es)
    row = dict(B=B, C=C, H=H, W=Wd, dtype=args.dtype, calls=count, x_proj=bool(args.xproj), R=R)
    for name, fn, by in (("fused_fwd", fwd_of(fused), fwd_bytes), ("chain_fwd", fwd_of(chain), None),
                         ("fused_fwd_bwd", both_of(fused), fwd_bytes + bwd_bytes), ("chain_fwd_bwd", both_of(chain), None)):
        ms = timeit(fn, args.reps, n_sets)
        row[name + "_us"] = round(ms * 1e3, 2)
        if by:
            row[name + "_GBps"] = round(by / ms / 1e6, 1)
            row[name + "_frac_of_peak"] = round(by / ms / 1e6 / peak, 3)
    row["speedup_fwd"] = round(row["chain_fwd_us"] / row["fused_fwd_us"], 2)
    row["speedup_fwd_bwd"] = round(row["chain_fwd_bwd_us"] / row["fused_fwd_bwd_us"], 2)
    print(json.dumps(row), flush=True)
    del sets
    torch.cuda.empty_cache()
