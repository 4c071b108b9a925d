#!/usr/bin/env python
"""Run one scan shape a few times (for ncu):  python tools/profile_one.py B D L [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import make_call_inputs
from vm_asr_b200 import scan, workload as W

B, D, L = (int(v) for v in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(0)
call = W.SS2DCall(D // 4, 1, L)
sets = []
for _ in range(3):
    inp = make_call_inputs(call, B, dev, gen)
    n_chunks = (L + 2047) // 2048
    b = dict(out=torch.empty_like(inp["u"]), x=torch.empty(B, D, n_chunks, 2, device=dev), du=torch.empty_like(inp["u"]),
             ddelta=torch.empty_like(inp["u"]), dA=torch.zeros(D, 1, device=dev), dD=torch.zeros(D, device=dev),
             dbias=torch.zeros(D, device=dev), dB=torch.zeros(B, 4, 1, L, device=dev), dC=torch.zeros(B, 4, 1, L, device=dev))
    sets.append((inp, b))
for r in range(reps):
    inp, b = sets[r % 3]
    scan.fwd_out(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["bias"], True, b["out"], b["x"])
    scan.bwd_out(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["bias"], inp["dout"], b["x"], True,
                 b["du"], b["ddelta"], b["dA"], b["dB"], b["dC"], b["dD"], b["dbias"])
torch.cuda.synchronize()
print("done")
