#!/usr/bin/env python
"""CPU only: how the host plans every SS2D call of a workload (through vmasr_scan_plan) and what wave quantisation costs:
    python tools/plan_report.py [--workload vm_asr_48k_MPD] [--sms 148]
waves = tiles / (SMs x resident CTAs per SM: 3 forward / 2 backward for the multi-chunk kernels, 3 / 2 single-chunk)."""
import argparse
import ctypes
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_plan_cpu import _params, _plan  # noqa: E402  (fake-pointer parameter blocks)
from vm_asr_b200 import workload as W  # noqa: E402

FAMILY = {0: "generic", 1: "single-chunk", 2: "multi-chunk", 3: "ring"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="vm_asr_48k_MPD")
    ap.add_argument("--sms", type=int, default=148)
    args = ap.parse_args()
    wl = W.WORKLOADS[args.workload]
    print(f"{wl.name}: batch {wl.batch}")
    print("| D | L | calls | pass | kernel family | channels/tile | tiles | slots | waves | last wave filled | wave efficiency |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for call, count in W.distinct_shapes(wl):
        for bwd in (False, True):
            rc, pl, err = _plan(_params(wl.batch, call.D, call.L, bwd=bwd), bwd)
            assert rc == 0, err
            slots = args.sms * (2 if bwd else 3)
            waves = pl["grid"] / slots
            full = math.ceil(waves)
            print(f"| {call.D} | {call.L} | {count} | {'bwd' if bwd else 'fwd'} | {FAMILY[pl['variant']]} | {pl['cpt']} | {pl['grid']} | {slots} | "
                  f"{waves:.2f} | {100 * (waves - (full - 1)):.0f} % | {100 * waves / full:.0f} % |")


if __name__ == "__main__":
    main()
