#!/usr/bin/env python
"""The block's tail at the config shapes: ONE kernel (vmasr_outnorm_gate_fwd / _bwd: merge of the core's planes + LayerNorm + cast
+ SiLU(z) + gate) against what the reference runs after the scan (vmamba.py:1497-1531, 1536-1550) on the same box -- the merge
(ours: vmasr_map_merge2), transpose(1, 2).contiguous(), nn.LayerNorm, .to(dtype), SiLU, product -- forward and forward +
backward, CUDA events around CUDA graphs over rotating buffer sets.
    python tools/tail_bench.py [--workload vm_asr_48k_MPD] [--dtype float32|float16]
One JSON line per shape; GB/s are the fused kernel's algorithmic bytes (forward: two planes + z in, out and the saved map out;
backward: dout, z, saved map in, dy, dy^T (+ its read), dz out)."""
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
    sets = [dict(planes=torch.randn(2, B, C, L, device=dev, generator=gen), z=torch.randn(B, H, Wd, C, device=dev, generator=gen).to(dt),
                 gout=torch.randn(B, H, Wd, C, device=dev, generator=gen).to(dt)) for _ in range(n_sets)]
    gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)

    def fused(i, grad=False):
        d = sets[i]
        ins = [d["planes"].requires_grad_(grad), gamma.requires_grad_(grad), beta.requires_grad_(grad), d["z"].requires_grad_(grad)]
        return ss2d.MergeNormGate.apply(*ins, H, Wd, 1e-5, True, dt), ins, d["gout"]

    def chain(i, grad=False):
        d = sets[i]
        ins = [d["planes"].requires_grad_(grad), gamma.requires_grad_(grad), beta.requires_grad_(grad), d["z"].requires_grad_(grad)]
        y = ins[0][0] + ss2d.MapTranspose.apply(ins[0][1].view(B, C, Wd, H)).view(B, C, L)   # stands in for CrossMerge's store
        y = F.layer_norm(y.transpose(1, 2).contiguous(), (C,), ins[1], ins[2], 1e-5).view(B, H, Wd, C).to(dt)
        return y * F.silu(ins[3]), ins, d["gout"]

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

    fwd_bytes = n * (2 * 4 + es + es + 4)
    bwd_bytes = n * (es + es + 4 + 4 + 4 + 4 + es)
    row = dict(B=B, C=C, H=H, W=Wd, dtype=args.dtype, calls=count)
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
