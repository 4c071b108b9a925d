#!/usr/bin/env python
"""Per-shape timing of the SS2D core: fused (vmasr_ss2d_core_fwd/bwd) against the chain of this library's three operators
(cross scan -> selective scan -> cross merge), forward and forward+backward, CUDA events around CUDA graphs over rotating
buffer sets larger than L2.  Scan inputs are synthetic (the projections are not part of the core's kernels).
    python tools/ss2d_bench.py [--workload vm_asr_48k_MPD] [--reps 10] [--pair]
One JSON line per shape: ms, fused algorithmic GB/s, "effective" GB/s against the chain's algorithmic bytes."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import load_peaks  # noqa: E402
from tools.shape_bench import timeit  # noqa: E402
from vm_asr_b200 import cross, scan, ss2d, workload as W  # noqa: E402


def make_set(B, C, H, W_, dev, gen):
    L = H * W_
    r = lambda *s: torch.randn(*s, device=dev, generator=gen)
    u = lambda *s: torch.rand(*s, device=dev, generator=gen)
    d = dict(x=r(B, C, H, W_), dts_rm=0.5 * u(B, 2, C, L), dts_cm=0.5 * u(B, 2, C, L), Bs_rm=r(B, 2, 1, L), Bs_cm=r(B, 2, 1, L),
             Cs_rm=r(B, 2, 1, L), Cs_cm=r(B, 2, 1, L), As=-0.5 * u(4 * C, 1), Ds=r(4 * C), bias=0.5 * u(4 * C), dy=r(B, C, L))
    # the chain's inputs: time-order tensors (values do not matter for timing)
    d["dts4"] = 0.5 * u(B, 4 * C, L)
    d["Bs4"], d["Cs4"] = r(B, 4, 1, L), r(B, 4, 1, L)
    # the projected form (delta generated inside the kernels): x_dbl rows (dt, B, C) per pair, dt_projs_weight (4, C, 1)
    d["xd_rm"], d["xd_cm"], d["dt_w"] = 0.5 * r(B, 2, 3, L), 0.5 * r(B, 2, 3, L), 0.5 * r(4, C, 1)
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="vm_asr_48k_MPD")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--pair", action="store_true", help="fused: the two streams' cores in one grid (ss2d pair)")
    ap.add_argument("--projected", action="store_true", help="also time the projected form (dt_rank 1, L > 2048) against the fused "
                    "core preceded by the dt einsum it replaces")
    ap.add_argument("--only", default="", help="comma-separated d_inner values to run (default: every shape)")
    args = ap.parse_args()
    wl = W.WORKLOADS[args.workload]
    peak, _ = load_peaks()
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(0)
    B = wl.batch
    only = {int(v) for v in args.only.split(",") if v}
    for call, count in W.distinct_shapes(wl):
        if only and call.d_inner not in only:
            continue
        C, H, Wd, L = call.d_inner, call.H, call.W, call.L
        D = 4 * C
        fused_fwd = 4 * (B * C * L + B * D * L + 2 * B * 4 * L + B * C * L)
        fused_bwd = 4 * (2 * B * C * L + B * D * L + 2 * B * 4 * L) + 4 * (B * C * L + B * D * L + 2 * B * 4 * L)
        chain_fwd = 4 * (5 * B * C * L + 3 * B * D * L + 2 * B * 4 * L + 5 * B * C * L)
        chain_bwd = 4 * (5 * B * C * L + 5 * B * D * L + 4 * B * 4 * L + 5 * B * C * L)
        n_sets = max(2, min(6, int(300e6 // max(chain_fwd, 1)) + 1))
        sets = [make_set(B, C, H, Wd, dev, gen) for _ in range(n_sets)]
        names = ("x", "dts_rm", "dts_cm", "Bs_rm", "Bs_cm", "Cs_rm", "Cs_cm", "As", "Ds", "bias")

        def fused_f(i, grad=False):
            d = sets[i]
            ts = [d[k].requires_grad_(grad) for k in names]
            xT = ss2d.MapTranspose.apply(ts[0])
            if args.pair:
                d2 = sets[(i + 1) % n_sets]
                ts2 = [d2[k].requires_grad_(grad) for k in names]
                xT2 = ss2d.MapTranspose.apply(ts2[0])
                ys = ss2d._SS2DScan.apply(True, 2, ts[0], xT, *ts[1:], ts2[0], xT2, *ts2[1:])
                return ys, ts + ts2, (d["dy"], d2["dy"])
            y = ss2d._SS2DScan.apply(True, 1, ts[0], xT, *ts[1:])
            return (y,), ts, (d["dy"],)

        def fused_fwd_only(i):
            with torch.no_grad():
                fused_f(i)

        def fused_both(i):
            ys, ts, dys = fused_f(i, True)
            torch.autograd.grad(ys, ts, dys)

        pnames = ("x", "xd_rm", "xd_cm", "dt_w", "As", "Ds", "bias")

        def proj_f(i, grad=False):
            d = sets[i]
            ts = [d[k].requires_grad_(grad) for k in pnames]
            xT = ss2d.MapTranspose.apply(ts[0])
            y = ss2d._SS2DScanProj.apply(True, 1, ts[0], xT, *ts[1:])
            return (y,), ts, (d["dy"],)

        def einsum_f(i, grad=False):  # what the projected form replaces: dts = einsum(dt rows, dt_projs_weight), then the fused core
            d = sets[i]
            ts = [d[k].requires_grad_(grad) for k in pnames]
            xT = ss2d.MapTranspose.apply(ts[0])
            w = ts[3]
            dts = [torch.einsum("bkrl,kdr->bkdl", xd[:, :, :1], w[par::2]) for par, xd in ((0, ts[1]), (1, ts[2]))]
            bc = [xd[:, :, j:j + 1] for j in (1, 2) for xd in (ts[1], ts[2])]
            y = ss2d._SS2DScan.apply(True, 1, ts[0], xT, dts[0], dts[1], *bc, *ts[4:])
            return (y,), ts, (d["dy"],)

        def both_of(f):
            def run(i):
                ys, ts, dys = f(i, True)
                torch.autograd.grad(ys, ts, dys)
            return run

        def fwd_of(f):
            def run(i):
                with torch.no_grad():
                    f(i)
            return run

        def chain_f(i, grad=False):
            d = sets[i]
            ts = [d[k].requires_grad_(grad) for k in ("x", "dts4", "As", "Bs4", "Cs4", "Ds", "bias")]
            xs = cross.CrossScan.apply(ts[0])
            ys = scan.SelectiveScanCore.apply(xs.view(B, D, L), ts[1], ts[2], ts[3], ts[4], ts[5], ts[6], True)
            y = cross.CrossMerge.apply(ys.view(B, 4, C, H, Wd))
            return y, ts, d["dy"]

        def chain_fwd_only(i):
            with torch.no_grad():
                chain_f(i)

        def chain_both(i):
            y, ts, dy = chain_f(i, True)
            torch.autograd.grad(y, ts, dy)

        mult = 2 if args.pair else 1
        row = dict(B=B, C=C, H=H, W=Wd, calls=count, pair=bool(args.pair))
        for name, fn, alg, eff, m in (("fused_fwd", fused_fwd_only, fused_fwd, chain_fwd, mult),
                                      ("fused_fwd_bwd", fused_both, fused_fwd + fused_bwd, chain_fwd + chain_bwd, mult),
                                      ("chain_fwd", chain_fwd_only, chain_fwd, chain_fwd, 1),
                                      ("chain_fwd_bwd", chain_both, chain_fwd + chain_bwd, chain_fwd + chain_bwd, 1)):
            ms = timeit(fn, args.reps, n_sets) / m
            row[name + "_ms"] = round(ms, 5)
            row[name + "_GBps"] = round(alg / ms / 1e6, 1)
            if name.startswith("fused"):
                row[name + "_effective_GBps"] = round(eff / ms / 1e6, 1)
        if args.projected and L > 2048 and L % 16 == 0 and not args.pair:
            for name, fn in (("proj_fwd", fwd_of(proj_f)), ("proj_fwd_bwd", both_of(proj_f)), ("einsum_fused_fwd", fwd_of(einsum_f)),
                             ("einsum_fused_fwd_bwd", both_of(einsum_f))):
                row[name + "_ms"] = round(timeit(fn, args.reps, n_sets), 5)
            row["proj_speedup_fwd"] = round(row["einsum_fused_fwd_ms"] / row["proj_fwd_ms"], 3)
            row["proj_speedup_fwd_bwd"] = round(row["einsum_fused_fwd_bwd_ms"] / row["proj_fwd_bwd_ms"], 3)
        row["speedup_fwd"] = round(row["chain_fwd_ms"] / row["fused_fwd_ms"], 3)
        row["speedup_fwd_bwd"] = round(row["chain_fwd_bwd_ms"] / row["fused_fwd_bwd_ms"], 3)
        row["fused_fwd_bwd_frac_of_peak"] = round(row["fused_fwd_bwd_GBps"] / peak, 3)
        print(json.dumps(row), flush=True)
        del sets
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
