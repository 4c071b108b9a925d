#!/usr/bin/env python
"""Per-CTA phase timeline of the multi-chunk scan kernels (VMASR_TUNING build, vmasr_debug_timeline):
    VMASR_B200_LIBRARY=vm_asr_b200/lib_tuning/libvmasr_b200.so python tools/timeline.py B D L
Slots (ns, %globaltimer): 0 entry, 1 parameters staged (first CTA barrier), 2 B/C arrived, 3 channel 0 arrived, 4 P1 sweep
done, 5 entering state of channel 0 ready, 6 ... of the last channel, 7 P2 sweep done; exchange warp: 8 totals of channel 0
in, 9 of the last channel, 10 look-back finished; 15 SM id.  Prints the mean duration of every phase and how the tiles of
an SM follow each other."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import make_call_inputs
from vm_asr_b200 import _lib, scan, workload as W

B, D, L = (int(v) for v in sys.argv[1:4])
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(0)
call = W.SS2DCall(D // 4, 1, L)
lib = _lib.load_library()
lib.vmasr_debug_timeline.argtypes = [ctypes.c_void_p]
lib.vmasr_debug_timeline.restype = None
inp = make_call_inputs(call, B, dev, gen)
n_chunks = (L + 2047) // 2048
b = dict(out=torch.empty_like(inp["u"]), x=torch.empty(B, D, n_chunks, 2, device=dev), du=torch.empty_like(inp["u"]),
         ddelta=torch.empty_like(inp["u"]), dA=torch.zeros(D, 1, device=dev), dD=torch.zeros(D, device=dev),
         dbias=torch.zeros(D, device=dev), dB=torch.zeros(B, 4, 1, L, device=dev), dC=torch.zeros(B, 4, 1, L, device=dev))


def fwd():
    scan.fwd_out(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["bias"], True, b["out"], b["x"])


def bwd():
    scan.bwd_out(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["bias"], inp["dout"], b["x"], True,
                 b["du"], b["ddelta"], b["dA"], b["dB"], b["dC"], b["dD"], b["dbias"])


names = {(0, 1): "entry -> parameters staged", (1, 2): "-> B/C arrived", (0, 2): "tile start -> B/C arrived (persistent)", (2, 3): "-> channel 0 arrived", (3, 4): "P1 sweep",
         (4, 5): "wait entering state ch 0", (5, 6): "P2 up to the last channel's wait", (6, 7): "P2 last channel",
         (0, 8): "[x] entry -> totals of ch 0", (8, 9): "[x] -> totals of last ch", (9, 10): "[x] -> look-back finished",
         (0, 7): "TOTAL compute warp 0", (0, 10): "TOTAL exchange warp"}
for name, fn in (("fwd", fwd), ("bwd", bwd)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    grid = 1 << 16
    buf = torch.zeros(grid, 16, dtype=torch.int64, device=dev)
    lib.vmasr_debug_timeline(buf.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    lib.vmasr_debug_timeline(None)
    t = buf.cpu().numpy()
    used = t[:, 0] > 0
    t = t[used]
    n = len(t)
    t0 = t[:, 0].min()
    print(f"== {name} B={B} D={D} L={L}: {n} tiles, kernel {e1.elapsed_time(e0) if False else e0.elapsed_time(e1) * 1e3:.1f} us, "
          f"first entry -> last exit {(max(t[:, 7].max(), t[:, 10].max()) - t0) / 1e3:.1f} us")
    for (a, c), label in names.items():
        ok = (t[:, a] > 0) & (t[:, c] > 0)
        d = (t[ok, c] - t[ok, a]) / 1e3
        if len(d) == 0:  # (the persistent kernels do not stamp slot 1)
            continue
        print(f"   {label:40s} mean {d.mean():7.2f} us   p10 {sorted(d)[len(d) // 10]:7.2f}   p90 {sorted(d)[9 * len(d) // 10]:7.2f}")
    # residency: tiles per SM over time
    import numpy as np
    sm = t[:, 15]
    end = np.maximum(t[:, 7], t[:, 10])
    busy = []
    for s in np.unique(sm):
        m = sm == s
        busy.append(((end[m] - t[m, 0]).sum()) / max(1, (end[m].max() - t[m, 0].min())))
    print(f"   mean resident tiles per SM while it is active: {np.mean(busy):.2f}; tiles per SM {n / len(np.unique(sm)):.1f}")
    # turn-around: on one SM the i-th exit frees the slot the (i + slots)-th entry takes
    gaps = []
    for s in np.unique(sm):
        m = sm == s
        ent, ex = np.sort(t[m, 0]), np.sort(end[m])
        slots = int(np.searchsorted(ent, ex[0]))  # tiles that entered before the first exit = resident slots
        if slots >= 1 and len(ent) > slots:
            gaps.append((ent[slots:] - ex[:len(ent) - slots]) / 1e3)
    if gaps:
        g = np.concatenate(gaps)
        print(f"   slot turn-around (exit of warp 0 / exchange warp -> next tile's entry on that SM): mean {g.mean():.2f} us, "
              f"p10 {np.percentile(g, 10):.2f}, p90 {np.percentile(g, 90):.2f}")
    starts = np.sort(t[:, 0] - t0) / 1e3
    print("   tile start times (us) deciles:", [round(float(starts[int(q * (n - 1) / 10)]), 1) for q in range(11)])
