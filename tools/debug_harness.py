#!/usr/bin/env python
"""Diagnosis of the step harness on the tiny test workload: repeats fresh TrainSteps from one seed and reports, per trial, the
losses, the GradScaler's scale, the first non-finite activation (forward hooks on every core call) and the parameters whose
gradients are non-finite; then compares two identically seeded forward/backward passes tensor by tensor (gross run-to-run
differences point at a race or at uninitialised memory, last-bit differences at the atomics' order)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vm_asr_b200 import harness, ss2d
from vm_asr_b200.workload import SS2DCall, Workload


def small():
    calls = ([SS2DCall(8, 32, 16)] * 4 + [SS2DCall(16, 16, 8)] * 4 + [SS2DCall(32, 8, 4)] * 2 + [SS2DCall(16, 16, 8)] * 2
             + [SS2DCall(4, 64, 32)] * 2 + [SS2DCall(2, 128, 64)] * 2)
    return Workload("tiny", "(test)", 2, 64 * 63, 256, 64, 256, 16000, tuple(calls))


wl = small()
dev = torch.device("cuda")
amp = os.environ.get("AMP", "1") == "1"
trace = []
orig_pair = ss2d.ss2d_core_pair


def traced_pair(xa, pa, xb, pb, delta_softplus=True):
    ya, yb = orig_pair(xa, pa, xb, pb, delta_softplus)
    trace.append((xa.detach().float().abs().max().item(), xb.detach().float().abs().max().item(),
                  ya.detach().abs().max().item(), yb.detach().abs().max().item(),
                  bool(torch.isfinite(ya).all() and torch.isfinite(yb).all())))
    return ya, yb


ss2d.ss2d_core_pair = traced_pair
harness.ss2d.ss2d_core_pair = traced_pair

for trial in range(int(os.environ.get("TRIALS", "2"))):
    torch.manual_seed(0)
    ts = harness.TrainStep(wl, dev, world=1, amp=amp, lr=float(os.environ.get('LR', '2e-4')))
    x, y = harness.synthetic_batch(wl, dev)
    rec = []
    for step in range(24):
        trace.clear()
        loss = ts(x, y).item()
        bad = [n for n, p in ts.net.named_parameters() if not torch.isfinite(p.grad).all()]
        gmax = max(p.grad.abs().max().item() for p in ts.net.parameters() if torch.isfinite(p.grad).all()) if len(bad) < 108 else float("nan")
        first_bad = next((i for i, t in enumerate(trace) if not t[4]), None)
        rec.append((round(loss, 4), ts.scaler.get_scale(), len(bad), first_bad, f"{gmax:.3g}",
                    f"{max(t[0] for t in trace):.3g}", f"{max(t[2] for t in trace):.3g}"))
    print("trial", trial, "amp", amp)
    for r in rec:
        print("   loss %s scale %s bad_grads %s first_bad_core %s gmax %s max|x| %s max|y| %s" % r)
    if rec[-1][2]:
        print("   bad:", bad[:6])

# run-to-run comparison of one forward/backward, fp32 and amp
for use_amp in (False, True):
    outs = []
    for rep in range(3):
        torch.manual_seed(0)
        ts = harness.TrainStep(wl, dev, world=1, amp=use_amp)
        x, y = harness.synthetic_batch(wl, dev)
        trace.clear()
        loss = ts._fwd_bwd(x, y)
        torch.cuda.synchronize()
        outs.append((loss.item(), ts.grads.flat.clone(), list(trace)))
    for rep in (1, 2):
        d = (outs[rep][1] - outs[0][1]).abs()
        ref = outs[0][1].abs().max().item()
        print(f"amp={use_amp} rep {rep}: loss {outs[rep][0]:.6f} vs {outs[0][0]:.6f}; grad max diff {d.max().item():.3g} (max |g| {ref:.3g}); finite {bool(torch.isfinite(outs[rep][1]).all())}")
        for i, (ta, tb) in enumerate(zip(outs[0][2], outs[rep][2])):
            if abs(ta[2] - tb[2]) > 1e-3 * max(1.0, abs(ta[2])) or abs(ta[3] - tb[3]) > 1e-3 * max(1.0, abs(ta[3])):
                print("    core", i, "differs run to run:", ta, tb)
                break
