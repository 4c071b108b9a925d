#!/usr/bin/env python
"""Where the harness step's device time goes: torch.profiler over a few eager steps, CUDA kernels grouped by name.
    python tools/harness_profile.py [--workload vm_asr_48k_MPD] [--steps 3]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from vm_asr_b200 import harness, workload as W  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="vm_asr_48k_MPD")
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    wl = W.WORKLOADS[args.workload]
    dev = torch.device("cuda")
    ts = harness.TrainStep(wl, dev)
    x, y = harness.synthetic_batch(wl, dev)
    for _ in range(3):
        ts(x, y)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(args.steps):
            ts(x, y)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=60, max_name_column_width=90))


if __name__ == "__main__":
    main()
