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
    # as the product's autograd path does: dB / dC stored where one tile spans the group, else summed into a buffer that the
    # forward launch clears as its side job
    store = scan.dbdc_store_candidate(inp["u"], inp["A"], inp["B"])
    bc = None if store else torch.full((2 * B * 4 * L,), float("nan"), device=dev)
    dB, dC = (torch.empty(B, 4, 1, L, device=dev) for _ in range(2)) if store else (bc[:B * 4 * L].view(B, 4, 1, L), bc[B * 4 * L:].view(B, 4, 1, L))
    b = dict(out=torch.empty_like(inp["u"]), x=torch.empty(B, D, n_chunks, 2, device=dev), du=torch.empty_like(inp["u"]),
             ddelta=torch.empty_like(inp["u"]), dA=torch.zeros(D, 1, device=dev), dD=torch.zeros(D, device=dev),
             dbias=torch.zeros(D, device=dev), dB=dB, dC=dC, bc=bc, flags=scan.SCAN_DBDC_STORE if store else 0)
    sets.append((inp, b))
for r in range(reps):
    inp, b = sets[r % 3]
    scan.fwd_out(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["bias"], True, b["out"], b["x"], zero=b["bc"])
    scan.bwd_out(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["bias"], inp["dout"], b["x"], True,
                 b["du"], b["ddelta"], b["dA"], b["dB"], b["dC"], b["dD"], b["dbias"], flags=b["flags"])
torch.cuda.synchronize()
assert all(torch.isfinite(b["dB"]).all() and torch.isfinite(b["dC"]).all() for _, b in sets[:min(reps, 3)])
print("done")
