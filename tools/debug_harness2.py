#!/usr/bin/env python
"""Why does the harness step test sometimes see a loss that never moves?  Fresh TrainSteps on poisoned allocator memory;
per trial: losses, gradient statistics of the flat buffer, how far the parameters moved."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vm_asr_b200 import harness
from vm_asr_b200.workload import SS2DCall, Workload


def small():
    calls = ([SS2DCall(8, 32, 16)] * 4 + [SS2DCall(16, 16, 8)] * 4 + [SS2DCall(32, 8, 4)] * 2 + [SS2DCall(16, 16, 8)] * 2
             + [SS2DCall(4, 64, 32)] * 2 + [SS2DCall(2, 128, 64)] * 2)
    return Workload("tiny", "(test)", 2, 64 * 63, 256, 64, 256, 16000, tuple(calls))


def poison():
    junk = [torch.full((64 << 20,), float("nan"), device="cuda") for _ in range(4)]
    small_ = [torch.full((n,), float("nan"), device="cuda") for n in (1 << 8, 1 << 12, 1 << 16, 1 << 18) for _ in range(16)]
    del junk, small_


wl = small()
dev = torch.device("cuda")
for amp in (False, True):
    for trial in range(int(os.environ.get("TRIALS", "5"))):
        if os.environ.get("POISON", "1") == "1":
            poison()
        ts = harness.TrainStep(wl, dev, world=1, lr=float(os.environ.get('LR', '1e-3')), amp=amp)
        x, y = harness.synthetic_batch(wl, dev)
        p0 = torch.cat([p.detach().flatten().clone() for p in ts.net.parameters()])
        rec = []
        for step in range(int(os.environ.get("STEPS", "12"))):
            loss = ts(x, y).item()
            f = ts.grads.flat
            fin = torch.isfinite(f)
            rec.append((loss, int((~fin).sum()), float(f[fin].abs().max()) if fin.any() else float("nan"), int((f != 0).sum())))
        p1 = torch.cat([p.detach().flatten() for p in ts.net.parameters()])
        moved = (p1 - p0).abs()
        print(f"amp={amp} trial {trial}: moved max {moved.max().item():.3g} nonfinite params {int((~torch.isfinite(p1)).sum())} scale {ts.scaler.get_scale()}")
        for r in rec[:3] + rec[-2:]:
            print("    loss %.10f nonfinite grads %d gmax %.3g nonzero %d / %d" % (r + (ts.grads.flat.numel(),)))
