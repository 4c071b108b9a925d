#!/usr/bin/env python
"""Same-box comparison of the cross scan / merge kernels with the reference's Triton kernels (model/csm_triton.py:311-366,
staged unmodified under oracle/_ref/py by oracle/stage_ref_py.py): outputs must be identical (scan) / equal up to the
association of the four-term sum (merge: the Triton kernel adds y1+y2+y3+y4 left to right, csm_triton.py:154), times from
CUDA graphs over rotating buffers.  One JSON line per map of the workload and dtype.
    python tools/cross_vs_triton.py [--workload vm_asr_48k_MPD] [--reps 20]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import load_peaks  # noqa: E402
from oracle import stage_ref_py  # noqa: E402
from tools.shape_bench import timeit  # noqa: E402
from vm_asr_b200 import cross, workload as W  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="vm_asr_48k_MPD")
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    if not stage_ref_py.available():
        print(json.dumps({"unavailable": "reference sources not staged (oracle/stage_ref_py.py)"}))
        return
    stage_ref_py.load("vmamba")
    import csm_triton
    wl = W.WORKLOADS[args.workload]
    peak, _ = load_peaks()
    dev = torch.device("cuda")
    B = wl.batch
    for call, count in W.distinct_shapes(wl):
        C, H, Wd, L = call.d_inner, call.H, call.W, call.L
        for dt in (torch.float32, torch.float16):
            es = 4 if dt == torch.float32 else 2
            nb = es * 5 * B * C * L
            n_sets = max(2, min(8, int(400e6 // nb) + 1))
            xs_in = [torch.randn(B, C, H, Wd, device=dev).to(dt) for _ in range(n_sets)]
            ys_in = [torch.randn(B, 4, C, H, Wd, device=dev).to(dt) for _ in range(n_sets)]
            a, b = cross.cross_scan(xs_in[0]), csm_triton.CrossScanTriton.apply(xs_in[0])
            same_scan = bool(torch.equal(a, b.view_as(a)))
            m0, m1 = cross.cross_merge(ys_in[0], H, Wd), csm_triton.CrossMergeTriton.apply(ys_in[0])
            merge_diff = float((m0.float() - m1.view_as(m0).float()).abs().max())
            row = dict(B=B, C=C, H=H, W=Wd, dtype=str(dt).split(".")[-1], calls=count, scan_identical=same_scan,
                       merge_max_abs_diff_vs_triton=merge_diff)
            for name, ours, theirs in (("cross_scan", lambda i: cross.cross_scan(xs_in[i]), lambda i: csm_triton.CrossScanTriton.apply(xs_in[i])),
                                       ("cross_merge", lambda i: cross.cross_merge(ys_in[i], H, Wd), lambda i: csm_triton.CrossMergeTriton.apply(ys_in[i]))):
                t_o, t_t = timeit(ours, args.reps, n_sets), timeit(theirs, args.reps, n_sets)
                row[name] = dict(ours_us=round(t_o * 1e3, 2), triton_us=round(t_t * 1e3, 2), speedup=round(t_t / t_o, 2),
                                 ours_GBps=round(nb / t_o / 1e6, 1), ours_frac_of_peak=round(nb / t_o / 1e6 / peak, 3))
            print(json.dumps(row), flush=True)
            del xs_in, ys_in
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
