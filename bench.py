#!/usr/bin/env python
"""bench.py -- measurement of the SS2D + STFT hot path on B200 (BASELINE.json metric, configs[1] by default).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One JSON line.  What each key is:

* ``value`` (GB/s, headline, continuity with round 1): one STEP = the selective-scan work of one VM-ASR training step of the
  workload -- its 34 SS2D selective-scan forwards in forward order, then the 34 backwards in reverse order, at the config's
  batch, fp32 scan IO, d_state 1, 4 B/C groups, synthetic inputs with the reference test's distributions.  Inputs resident
  in HBM, every call on its own buffers (5.4 GB touched per step >> 126 MB L2), the step captured in one CUDA graph, CUDA
  events, max over ranks.  The two generator streams' same-shape calls go out as ONE grouped launch per pair
  (``vmasr_scan_fwd_grouped``; --pairing streams | none for the alternatives).  value = algorithmic scan bytes
  / time.  N > 1: every rank runs the step on its own batch shard, no collective (the path shards by clip).
* ``roofline``: the dominant kernel (scan_bwd_pipe_kernel: every backward launch with seqlen > 2048), CUDA events around
  each launch, algorithmic bytes of the launch / time against the measured HBM copy peak.
* ``ss2d_core``: the same 34 calls as FUSED SS2D cores (CrossScan -> scan -> CrossMerge in ``vmasr_ss2d_core_fwd/bwd``, no
  xs / ys copies) against the chain of the three operators, both under autograd in one CUDA graph: fused algorithmic GB/s
  (SURVEY.md 8d fused formula) and "effective" GB/s (the chain's algorithmic bytes over the fused time).
* ``train`` / ``infer``: the step harness (vm_asr_b200/harness.py: STFT -> 34 fused cores with glue -> iSTFT -> L1 + multi-resolution STFT loss ->
  backward -> bucketed, overlapped NCCL all-reduce of the real gradients + an MPD-sized 164 MB payload -> fused AdamW),
  launched eagerly: audio-seconds per second (SURVEY.md 8d "Throughput metric"), exposed communication time.
* ``e2e``: the headline metric END TO END through the public operator API: one harness training step per step with the
  waveforms in pinned HOST memory -- upload, STFT, 34 fused cores fwd + bwd, iSTFT, loss read back to the host -- the step's
  algorithmic scan bytes / wall time (device-synchronised), max over ranks.
* ``eager``: the ``value`` step launched call by call from Python (no graph): host cost per call.
* ``stft``: wav2spectro / spectro2wav (+ backward) kernel times next to the torch.stft / torch.istft chains.
* ``cpu_baseline`` and ``--impl reference``: the oracle's C restatement of the reference's scan forward + backward on the host
  cores (one clip of every distinct SS2D shape of the workload at FULL sequence length per step, threaded over B/C groups).
Nothing here reads /root/reference.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------
# clocks: NVML polled from a thread while the timed region runs
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
        0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting",
    }

    def __init__(self, index: int, period_s: float = 0.002):
        self.index, self.period = index, period_s
        self.samples, self._stop, self._thread = [], threading.Event(), None
        self.max_mhz = None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, reasons))
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.ok:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread:
            self._stop.set()
            self._thread.join()

    def summary(self, t0, t1):
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        window = "timed"
        if not inside:
            inside, window = self.samples[-5:], "nearest"
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "window": "unavailable"}
        bits = 0
        for s in inside:
            bits |= s[2]
        reasons = sorted(name for bit, name in self.REASONS.items() if bits & bit)
        return {"sm_mhz": statistics.median(s[1] for s in inside), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(inside), "window": window}


# ---------------------------------------------------------------------------------------------------------
# the scan step (headline)
# ---------------------------------------------------------------------------------------------------------
def make_call_inputs(call, batch, device, gen, pinned=False):
    """Reference test distributions (test_selective_scan.py:593-654)."""
    D, L, G, N = call.D, call.L, 4, 1

    def rnd(*shape, kind="n"):
        if device == "cpu":
            t = torch.empty(*shape, dtype=torch.float32, pin_memory=pinned)
        else:
            t = torch.empty(*shape, dtype=torch.float32, device=device)
        return t.normal_(generator=gen) if kind == "n" else t.uniform_(generator=gen)

    return dict(
        u=rnd(batch, D, L), delta=rnd(batch, D, L, kind="u").mul_(0.5), A=rnd(D, N, kind="u").mul_(-0.5),
        B=rnd(batch, G, N, L), C=rnd(batch, G, N, L), D=rnd(D), bias=rnd(D, kind="u").mul_(0.5), dout=rnd(batch, D, L),
    )


def graph_of(fn):
    """Run ``fn`` eagerly twice on a side stream (allocates this stream's workspaces), then capture it."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
        fn()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    return graph


def time_graph(graph, steps, warmup):
    for _ in range(max(warmup, 3)):
        graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


class DeviceStep:
    """All buffers of one step on the device (distinct per call: ~5.4 GB touched per step >> 126 MB L2) and the
    CUDA graph that replays it."""

    def __init__(self, wl, device, pairing="grouped"):
        from vm_asr_b200 import scan
        self.scan, self.wl, self.device, self.pairing = scan, wl, device, pairing
        self.side = None
        gen = torch.Generator(device=device).manual_seed(1234)
        self.calls = []
        # calls whose backward STORES dB / dC (VMASR_SCAN_DBDC_STORE: one tile spans a whole B / C group -- the C = 2 maps) need
        # no zero-fill of the two; the same rule vm_asr_b200.scan.bwd applies for the tensors it allocates itself
        self.store = [c.L > 2048 and c.L % 16 == 0 and c.D // 4 <= 4 and not os.environ.get("VMASR_BENCH_NO_DBDC_STORE")
                      for c in wl.calls]
        # every other call's dB / dC buffer is cleared by the call's own FORWARD launch (vmasr_scan_params.zero_ptr: the tiles
        # spread the stores over the time they wait for their first bytes), as vm_asr_b200.scan.SelectiveScanCore does
        self.fwd_zero = not os.environ.get("VMASR_BENCH_NO_FWD_ZERO")
        acc_floats = 0
        for c, st in zip(wl.calls, self.store):
            acc_floats += c.D * 3 + (0 if st or self.fwd_zero else 2 * wl.batch * 4 * c.L)
        # one arena for what is left of the accumulated gradients (dA, dD, ddelta_bias of all calls): zeroed by ONE memset
        self.arena = torch.zeros(acc_floats, dtype=torch.float32, device=device)
        self.zero_fill_bytes = 4 * acc_floats
        off = 0

        def take(n, shape):
            nonlocal off
            t = self.arena[off:off + n].view(*shape)
            off += n
            return t

        for c, st in zip(wl.calls, self.store):
            inp = make_call_inputs(c, wl.batch, device, gen)
            n_chunks = (c.L + 2047) // 2048
            n_bc = wl.batch * 4 * c.L
            bc = None
            if st:
                dB, dC = (torch.empty(wl.batch, 4, 1, c.L, device=device) for _ in range(2))
            elif self.fwd_zero:
                bc = torch.empty(2 * n_bc, device=device)
                dB, dC = bc[:n_bc].view(wl.batch, 4, 1, c.L), bc[n_bc:].view(wl.batch, 4, 1, c.L)
            else:
                dB, dC = take(n_bc, (wl.batch, 4, 1, c.L)), take(n_bc, (wl.batch, 4, 1, c.L))
            bufs = dict(
                out=torch.empty_like(inp["u"]), x=torch.empty(wl.batch, c.D, n_chunks, 2, device=device),
                du=torch.empty_like(inp["u"]), ddelta=torch.empty_like(inp["u"]),
                dA=take(c.D, (c.D, 1)), dD=take(c.D, (c.D,)), dbias=take(c.D, (c.D,)),
                dB=dB, dC=dC, bc=bc, flags=self.scan.SCAN_DBDC_STORE if st else 0,
            )
            self.calls.append((c, inp, bufs))
        self.graph = None

    def fwd_call(self, i):
        c, inp, b = self.calls[i]
        self.scan.fwd_out(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["bias"], True, b["out"], b["x"], zero=b["bc"])

    def bwd_call(self, i):
        c, inp, b = self.calls[i]
        self.scan.bwd_out(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["bias"], inp["dout"], b["x"],
                          True, b["du"], b["ddelta"], b["dA"], b["dB"], b["dC"], b["dD"], b["dbias"], flags=b["flags"])

    # The generator's two streams (magnitude, phase) issue the same-shape SS2D call independently between their interaction
    # points (model/model.py:1124-1127, 1167-1176): calls 2j and 2j + 1 of the workload are such a pair.
    def fwd_pair(self, i):
        args, outs, zero = [], [], []
        for k in (i, i + 1):
            c, inp, b = self.calls[k]
            args.append((inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["bias"], True))
            outs.append((b["out"], b["x"]))
            zero.append(b["bc"])
        self.scan.fwd_grouped(args, outs, zero=zero)

    def bwd_pair(self, i):
        args, outs = [], []
        for k in (i, i + 1):
            c, inp, b = self.calls[k]
            args.append((inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["bias"], inp["dout"], b["x"], True, b["flags"]))
            outs.append((b["du"], b["ddelta"], b["dA"], b["dB"], b["dC"], b["dD"], b["dbias"]))
        self.scan.bwd_grouped(args, outs)

    def _two_streams(self, fn, i):
        """call i on the current stream, call i + 1 on a side stream, joined afterwards (a fork / join inside the graph)"""
        main = torch.cuda.current_stream()
        if self.side is None:
            self.side = torch.cuda.Stream()
        self.side.wait_stream(main)
        fn(i)
        with torch.cuda.stream(self.side):
            fn(i + 1)
        main.wait_stream(self.side)

    def run_eager(self):
        self.arena.zero_()
        n = len(self.calls)
        if self.pairing == "grouped":
            for i in range(0, n, 2):
                self.fwd_pair(i)
            for i in reversed(range(0, n, 2)):
                self.bwd_pair(i)
        elif self.pairing == "streams":
            for i in range(0, n, 2):
                self._two_streams(self.fwd_call, i)
            for i in reversed(range(0, n, 2)):
                self._two_streams(self.bwd_call, i)
        else:
            for i in range(n):
                self.fwd_call(i)
            for i in reversed(range(n)):
                self.bwd_call(i)

    def capture(self):
        self.graph = graph_of(self.run_eager)

    def step(self):
        self.graph.replay()

    gpu_launches_per_step = property(lambda self: len(self.calls) if self.pairing == "grouped" else 2 * len(self.calls))


def time_dominant_kernel(ds: DeviceStep, steps: int):
    """CUDA events around every launch of the dominant kernel (selective-scan backward, multi-chunk variant
    scan_bwd_pipe_kernel: every call with seqlen > 2048) on the launching stream.  The whole step is enqueued behind a device-side
    delay without any host synchronisation in between, so the event pairs bracket kernel time and not the host's
    launch latency.  With grouped pairing one launch covers the two calls of a pair.  Returns (avg_ms, avg_algorithmic_bytes,
    launches per step)."""
    grouped = ds.pairing == "grouped"
    stride = 2 if grouped else 1
    n_calls = len(ds.calls)
    idx = {i for i in range(0, n_calls, stride) if ds.calls[i][0].L > 2048}
    total_ms, total_bytes, n = 0.0, 0, 0
    B = ds.wl.batch
    for _ in range(steps):
        pairs = []
        torch.cuda.synchronize()
        torch.cuda._sleep(20_000_000)  # ~10 ms of device-side delay: the host gets ahead of the GPU
        ds.arena.zero_()
        for i in range(0, n_calls, stride):
            ds.fwd_pair(i) if grouped else ds.fwd_call(i)
        for i in reversed(range(0, n_calls, stride)):
            if i in idx:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ds.bwd_pair(i) if grouped else ds.bwd_call(i)
                e1.record()
                pairs.append((i, e0, e1))
            else:
                ds.bwd_pair(i) if grouped else ds.bwd_call(i)
        torch.cuda.synchronize()
        for i, e0, e1 in pairs:
            c = ds.calls[i][0]
            total_ms += e0.elapsed_time(e1)
            total_bytes += stride * 4 * (5 * B * c.D * c.L + 4 * B * 4 * c.L)
            n += 1
    return total_ms / n, total_bytes / n, n // steps


def time_eager(ds: DeviceStep, steps: int):
    """The step launched call by call from Python, wall clock with a device sync at both ends: host-bound when the host
    cost per call exceeds the kernel time.  Twice: through the plain entry points (validated call-site cache, outputs
    pre-allocated) and through PreparedCalls (parameter blocks filled once; what a fixed-buffer caller would use)."""
    def run(fn):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        t_host = (time.perf_counter() - t0) / steps   # host enqueue time (the device may lag behind)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / steps, t_host

    launches = ds.gpu_launches_per_step
    t_all, t_host = run(ds.run_eager)
    out = {"ms_per_step": round(t_all * 1e3, 3), "host_ms_per_step": round(t_host * 1e3, 3),
           "host_us_per_launch": round(t_host * 1e6 / launches, 2), "launches_per_step": launches,
           "note": "Python -> ctypes -> C ABI -> cudaLaunchKernelEx per launch (grouped: one launch per pair of calls); "
                   "the reference's pybind entry costs ~9 us per call (round-1 measurement on the same box)"}
    if ds.pairing == "grouped":
        plans = []
        n = len(ds.calls)
        for i in range(0, n, 2):
            args, outs, zero = [], [], []
            for k in (i, i + 1):
                c, inp, b = ds.calls[k]
                args.append((inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["bias"], True))
                outs.append((b["out"], b["x"]))
                zero.append(b["bc"])
            plans.append(ds.scan.prepare_fwd(args, outs, zero))
        for i in reversed(range(0, n, 2)):
            args, outs = [], []
            for k in (i, i + 1):
                c, inp, b = ds.calls[k]
                args.append((inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["bias"], inp["dout"], b["x"], True, b["flags"]))
                outs.append((b["du"], b["ddelta"], b["dA"], b["dB"], b["dC"], b["dD"], b["dbias"]))
            plans.append(ds.scan.prepare_bwd(args, outs))

        def prepared():
            ds.arena.zero_()
            for p in plans:
                p()

        t_all, t_host = run(prepared)
        out["prepared"] = {"ms_per_step": round(t_all * 1e3, 3), "host_ms_per_step": round(t_host * 1e3, 3),
                           "host_us_per_launch": round(t_host * 1e6 / launches, 2),
                           "host_us_per_call": round(t_host * 1e6 / (2 * launches), 2)}
    return out


# ---------------------------------------------------------------------------------------------------------
# fused SS2D core vs the chain of the three operators (both under autograd, one CUDA graph each)
# ---------------------------------------------------------------------------------------------------------
class CoreStep:
    def __init__(self, wl, device):
        from vm_asr_b200 import cross, scan, ss2d
        self.cross, self.scan, self.ss2d, self.wl = cross, scan, ss2d, wl
        gen = torch.Generator(device=device).manual_seed(77)
        B = wl.batch
        r = lambda *s: torch.randn(*s, device=device, generator=gen)
        u = lambda *s: torch.rand(*s, device=device, generator=gen)
        self.sets = []
        for c in wl.calls:
            C, H, W, L = c.d_inner, c.H, c.W, c.L
            self.sets.append(dict(call=c, x=r(B, C, H, W), dts_rm=0.5 * u(B, 2, C, L), dts_cm=0.5 * u(B, 2, C, L),
                                  Bs_rm=r(B, 2, 1, L), Bs_cm=r(B, 2, 1, L), Cs_rm=r(B, 2, 1, L), Cs_cm=r(B, 2, 1, L),
                                  As=-0.5 * u(4 * C, 1), Ds=r(4 * C), bias=0.5 * u(4 * C), dy=r(B, C, L)))
        self.names = ("x", "dts_rm", "dts_cm", "Bs_rm", "Bs_cm", "Cs_rm", "Cs_cm", "As", "Ds", "bias")

    def bytes(self):
        B = self.wl.batch
        fused = chain = 0
        for c in self.wl.calls:
            CL, DL, GL = B * c.d_inner * c.L, B * c.D * c.L, B * 4 * c.L
            fused += 4 * (CL + DL + 2 * GL + CL) + 4 * (2 * CL + DL + 2 * GL) + 4 * (CL + DL + 2 * GL)
            chain += 4 * (10 * CL + 3 * DL + 2 * GL) + 4 * (10 * CL + 5 * DL + 4 * GL)
        return fused, chain

    def fused_step(self):
        ss2d = self.ss2d
        n = len(self.sets)
        pending = []
        for i in range(0, n, 2):   # forward, the two streams' cores of a pair in one grid
            ts = []
            for d in (self.sets[i], self.sets[i + 1]):
                t = [d[k].requires_grad_(True) for k in self.names]
                xT = ss2d.MapTranspose.apply(t[0])
                ts += [t[0], xT] + t[1:]
            ys = ss2d._SS2DScan.apply(True, 2, *ts)
            pending.append((ys, ts, (self.sets[i]["dy"], self.sets[i + 1]["dy"])))
        for ys, ts, dys in reversed(pending):
            leaves = [t for t in ts if t.is_leaf]
            torch.autograd.grad(ys, leaves, dys)

    def chain_step(self):
        cross, scan = self.cross, self.scan
        B = self.wl.batch
        pending = []
        for d in self.sets:
            c = d["call"]
            C, H, W, L = c.d_inner, c.H, c.W, c.L
            x = d["x"].requires_grad_(True)
            if "dts4" not in d:   # time-order tensors of the chain (values do not matter for timing)
                d["dts4"] = torch.cat([d["dts_rm"], d["dts_cm"]], dim=1).view(B, 4 * C, L).contiguous()
                d["Bs4"] = torch.cat([d["Bs_rm"], d["Bs_cm"]], dim=1).contiguous()
                d["Cs4"] = torch.cat([d["Cs_rm"], d["Cs_cm"]], dim=1).contiguous()
            ts = [x] + [d[k].requires_grad_(True) for k in ("dts4", "As", "Bs4", "Cs4", "Ds", "bias")]
            xs = cross.CrossScan.apply(x)
            ys = scan.SelectiveScanCore.apply(xs.view(B, 4 * C, L), ts[1], ts[2], ts[3], ts[4], ts[5], ts[6], True)
            y = cross.CrossMerge.apply(ys.view(B, 4, C, H, W))
            pending.append((y, ts, d["dy"]))
        for y, ts, dy in reversed(pending):
            torch.autograd.grad(y, ts, dy)


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle's C restatement of the reference scan (forward + backward), threaded over B/C groups
# ---------------------------------------------------------------------------------------------------------
CPU_SAMPLE = ("one clip (batch index 0) of each of the workload's 6 distinct SS2D selective-scan shapes at FULL sequence length, "
              "forward + backward, oracle/scan_ref.c (float64 accumulation, plain C), one thread per B/C group slice")


class CpuStep:
    def __init__(self, wl):
        import numpy as np
        from oracle import c_ref
        from vm_asr_b200 import workload as W
        self.c_ref, self.np = c_ref, np
        c_ref.lib()
        rng = np.random.default_rng(5)
        self.jobs, self.bytes = [], 0
        for call, _count in W.distinct_shapes(wl):
            D, L, G = call.D, call.L, 4
            cpg = D // G
            f = lambda *s: rng.standard_normal(s, dtype=np.float32)
            un = lambda *s: rng.random(s, dtype=np.float32)
            u, delta, dout = f(1, D, L), 0.5 * un(1, D, L), f(1, D, L)
            A, Dv, bias = -0.5 * un(D, 1), f(D), 0.5 * un(D)
            Bm, Cm = f(1, G, 1, L), f(1, G, 1, L)
            for g in range(G):
                sl = slice(g * cpg, (g + 1) * cpg)
                self.jobs.append((u[:, sl], delta[:, sl], A[sl], Bm[:, g:g + 1], Cm[:, g:g + 1], Dv[sl], bias[sl], dout[:, sl]))
            self.bytes += 4 * (8 * D * L + 6 * G * L)
        self.threads = os.cpu_count() or 1

    def _job(self, j):
        u, delta, A, Bm, Cm, Dv, bias, dout = j
        self.c_ref.scan_fwd(u, delta, A, Bm, Cm, Dv, bias, True)
        self.c_ref.scan_bwd(u, delta, A, Bm, Cm, Dv, bias, True, dout)

    def run(self, steps, warmup):
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(self.threads) as ex:
            for _ in range(warmup):
                list(ex.map(self._job, self.jobs))
            t0 = time.perf_counter()
            for _ in range(steps):
                list(ex.map(self._job, self.jobs))
            dt = (time.perf_counter() - t0) / steps
        return self.bytes / dt / 1e9, dt * 1e3, min(self.threads, len(self.jobs))


def workload_config(wl):
    """The ``config`` object both arms print (the driver compares them)."""
    return {"workload": f"{wl.yaml}: SS2D selective-scan forward + backward of one training step (34 + 34 calls), "
                        f"batch {wl.batch} per GPU, fp32, d_state 1, 4 B/C groups"}


# ---------------------------------------------------------------------------------------------------------
def stft_times(wl, device, batch, reps=20):
    """Device times (CUDA events, calls queued behind a device-side delay) of the STFT entry points next to the torch chains."""
    from vm_asr_b200 import stft
    out = {}
    B, T, F_, Nf = batch, wl.T, wl.n_fft // 2 + 1, 1 + wl.T // wl.hop
    wave = [0.1 * torch.randn(B, 1, T, device=device) for _ in range(4)]
    mp = [stft.wav2spectro(w, wl.n_fft, wl.hop, wl.win, "log2") for w in wave]
    win = torch.hann_window(wl.win, device=device)
    k = [0]

    def rot():
        k[0] = (k[0] + 1) % 4
        return k[0]

    def ours_stft():
        for _ in range(reps):
            stft.wav2spectro(wave[rot()], wl.n_fft, wl.hop, wl.win, "log2")

    def torch_stft():
        for _ in range(reps):
            s = torch.stft(wave[rot()].reshape(-1, T), wl.n_fft, wl.hop, wl.win, win, normalized=True, return_complex=True)
            torch.log2(s.abs() + 1e-8), torch.angle(s)

    def ours_istft():
        for _ in range(reps):
            m, p = mp[rot()]
            stft.spectro2wav(m, p, wl.n_fft, wl.hop, wl.win, "log2")

    def torch_istft():
        for _ in range(reps):
            m, p = mp[rot()]
            X = torch.exp2(m.reshape(-1, F_, Nf)) * torch.exp(1j * p.reshape(-1, F_, Nf))
            torch.istft(X, wl.n_fft, wl.hop, wl.win, win, normalized=True)

    def ours_istft_fb():
        for _ in range(reps):
            m, p = mp[rot()]
            m, p = m.detach().requires_grad_(), p.detach().requires_grad_()
            w = stft.spectro2wav(m, p, wl.n_fft, wl.hop, wl.win, "log2")
            torch.autograd.grad(w, (m, p), torch.ones_like(w))

    def torch_istft_fb():
        for _ in range(reps):
            m, p = mp[rot()]
            m, p = m.detach().requires_grad_(), p.detach().requires_grad_()
            X = torch.exp2(m.reshape(-1, F_, Nf)) * torch.exp(1j * p.reshape(-1, F_, Nf))
            w = torch.istft(X, wl.n_fft, wl.hop, wl.win, win, normalized=True)
            torch.autograd.grad(w, (m, p), torch.ones_like(w))

    def time_queued(fn):
        """Device time of `reps` calls enqueued behind a device-side delay (the host runs ahead, so launch gaps do not count);
        torch.istft does not capture into a CUDA graph, so every row is timed this way."""
        fn()
        torch.cuda.synchronize()
        best = None
        for _ in range(3):
            torch.cuda._sleep(40_000_000)   # ~20 ms
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1) / reps
            best = t if best is None else min(best, t)
        return best

    nb = 4 * B * T + 8 * B * F_ * Nf
    for name, fn, nbytes in (("stft", ours_stft, nb), ("torch_stft_chain", torch_stft, nb), ("istft", ours_istft, nb),
                             ("torch_istft_chain", torch_istft, nb), ("istft_fwd_bwd", ours_istft_fb, 2 * nb + 8 * B * F_ * Nf),
                             ("torch_istft_fwd_bwd", torch_istft_fb, 2 * nb + 8 * B * F_ * Nf)):
        try:
            ms = time_queued(fn)
            out[name] = {"us": round(ms * 1e3, 2), "GBps": round(nbytes / ms / 1e6, 1)}
        except Exception as e:
            out[name] = {"error": str(e)[:120]}
    for a, b in (("stft", "torch_stft_chain"), ("istft", "torch_istft_chain"), ("istft_fwd_bwd", "torch_istft_fwd_bwd")):
        if "us" in out.get(a, {}) and "us" in out.get(b, {}):
            out[a]["vs_torch"] = round(out[b]["us"] / out[a]["us"], 2)
    out["shape"] = f"B={B} T={T} n_fft={wl.n_fft} hop={wl.hop} win={wl.win}"
    return out


# ---------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="vm_asr_48k_MPD")
    ap.add_argument("--train-steps", type=int, default=10)
    ap.add_argument("--pairing", default="grouped", choices=["grouped", "streams", "none"],
                    help="how the two generator streams' same-shape calls are issued: one grouped launch per pair (default), "
                         "two CUDA streams, or one call after the other")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the harness (train / infer / e2e)")
    ap.add_argument("--no-core", action="store_true", help="skip the fused-core-vs-chain step")
    ap.add_argument("--no-stft", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly (profiling aid)")
    args = ap.parse_args()

    from vm_asr_b200 import workload as W
    wl = W.WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    fwd_b, bwd_b = wl.scan_bytes()
    step_bytes = fwd_b + bwd_b
    config = workload_config(wl)

    if args.impl == "reference":
        if rank != 0:
            return
        cpu = CpuStep(wl)
        gbs, ms, cores = cpu.run(args.steps, max(args.warmup, 1))
        print(json.dumps({
            "impl": "reference", "metric": "ss2d_scan_fwd_bwd_GBps", "value": round(gbs, 4), "unit": "GB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": round(ms, 2), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": round(gbs, 4), "unit": "GB/s", "cores": cores, "kind": "port", "sample": CPU_SAMPLE},
            "e2e": {"value": round(gbs, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    import torch.distributed as dist
    from vm_asr_b200 import dist as vdist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the product path)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    peak, peak_src = load_peaks()

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        gbs, ms, cores = CpuStep(wl).run(3, 1)
        cpu_base = {"value": round(gbs, 4), "unit": "GB/s", "cores": cores, "kind": "port", "sample": CPU_SAMPLE,
                    "ms_per_sample_step": round(ms, 1)}

    # ---- headline: the scan step ----
    ds = DeviceStep(wl, device, args.pairing)
    if args.no_graph:
        ds.step = ds.run_eager
        ds.run_eager()
    else:
        ds.capture()
    for _ in range(max(args.warmup, 3)):
        ds.step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        ds.step()
    e1.record()
    torch.cuda.synchronize()
    t_wall1 = time.perf_counter()
    if world > 1:
        dist.barrier()
    sampler.stop()
    ms_per_step = vdist.max_over_ranks(e0.elapsed_time(e1), device) / args.steps  # timing rule: max over ranks
    value = world * step_bytes / (ms_per_step * 1e-3) / 1e9
    clocks = sampler.summary(t_wall0, t_wall1)

    k_ms, k_bytes, k_per_step = time_dominant_kernel(ds, max(2, min(args.steps, 5)))
    achieved = k_bytes / (k_ms * 1e-3) / 1e9
    eager = time_eager(ds, 5) if rank == 0 else None
    zero_fill_bytes = ds.zero_fill_bytes
    del ds
    torch.cuda.empty_cache()

    # ---- fused SS2D core vs the chain ----
    core = None
    if not args.no_core:
        cs = CoreStep(wl, device)
        fused_b, chain_b = cs.bytes()
        t_fused = time_graph(graph_of(cs.fused_step), max(3, min(args.steps, 10)), 3)
        t_chain = time_graph(graph_of(cs.chain_step), max(3, min(args.steps, 10)), 3)
        t_fused, t_chain = vdist.max_over_ranks(t_fused, device), vdist.max_over_ranks(t_chain, device)
        core = {"fused_ms_per_step": round(t_fused, 4), "chain_ms_per_step": round(t_chain, 4),
                "fused_algorithmic_GBps": round(world * fused_b / t_fused / 1e6, 1),
                "effective_GBps_vs_chain_bytes": round(world * chain_b / t_fused / 1e6, 1),
                "chain_GBps": round(world * chain_b / t_chain / 1e6, 1), "speedup_vs_chain": round(t_chain / t_fused, 3),
                "fused_frac_of_hbm_peak": round(fused_b / t_fused / 1e6 / peak, 4),
                "what": "34 SS2D cores (CrossScan -> selective scan -> CrossMerge) forward + backward under autograd in one CUDA graph; "
                        "fused: vmasr_ss2d_core_fwd/bwd, the two streams' cores of a pair in one grid, no xs / ys copies; "
                        "chain: vmasr_cross_scan -> vmasr_scan_* -> vmasr_cross_merge per call"}
        del cs
        torch.cuda.empty_cache()

    # ---- the step harness: training / inference throughput and the end-to-end number ----
    train = infer = e2e = None
    if not args.no_e2e:
        from vm_asr_b200 import harness
        host_in, host_tgt = harness.synthetic_batch(wl, device, rank, pinned=True)
        secs = wl.batch * wl.clip_seconds
        n_train = max(3, args.train_steps)

        def timed(fn, n):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(n):
                fn()
            torch.cuda.synchronize()
            return vdist.max_over_ranks((time.perf_counter() - t0) / n, device)

        def measure(graph):
            ts = harness.TrainStep(wl, device, world=world)
            if graph:
                ts.capture(host_in.to(device), host_tgt.to(device))

            def e2e_step(comm=True):
                x = host_in.to(device, non_blocking=True)
                y = host_tgt.to(device, non_blocking=True)
                return ts(x, y, comm=comm).item()   # the loss is read back to the host: the step's result

            for _ in range(3):
                e2e_step()
            t_train = timed(e2e_step, n_train)
            t_nocomm = timed(lambda: e2e_step(False), n_train) if world > 1 else t_train
            dev_in = host_in.to(device)
            t_infer = timed(lambda: ts.infer(dev_in), n_train)
            info = {"audio_sec_per_s": round(world * secs / t_train, 1), "ms_per_step": round(t_train * 1e3, 2), "steps": n_train,
                    "launch": "forward + backward replayed as one CUDA graph; copies, all-reduce, AdamW eager" if graph
                    else "eager (Python, autograd): host-bound",
                    "parameters": ts.n_params,
                    "collective": (f"NCCL all-reduce of the {ts.n_params}-float gradient buffer in {len(ts.grads.buckets)} buckets "
                                   f"+ a {harness.MPD_PARAMS}-float zero payload standing in for the MPD gradients (not built); "
                                   + ("payload overlapped with the graph replay, gradients reduced after it" if graph
                                      else "buckets issued from autograd hooks, overlapped with backward"))
                    if world > 1 else "none",
                    "exposed_comm_ms": round((t_train - t_nocomm) * 1e3, 3) if world > 1 else 0.0}
            del ts
            torch.cuda.empty_cache()
            return info, t_train, t_infer

        train_eager, _, t_infer = measure(False)
        try:
            train, t_train, _ = measure(True)
        except Exception as e:   # graph capture of the whole step is an optimisation, not a requirement of the line
            train, t_train = dict(train_eager, graph_error=str(e)[:160]), train_eager["ms_per_step"] * 1e-3
        train["eager"] = train_eager
        infer = {"audio_sec_per_s": round(world * secs / t_infer, 1), "ms_per_step": round(t_infer * 1e3, 2),
                 "what": "torch.no_grad forward of the harness (STFT -> 34 fused cores -> iSTFT), eager, inputs on the device"}
        e2e = {"value": round(world * step_bytes / t_train / 1e9, 2), "unit": "GB/s", "h2d_bytes_per_step": 2 * host_in.numel() * 4,
               "d2h_bytes_per_step": 4, "ms_per_step": round(t_train * 1e3, 2), "steps": n_train,
               "api": "vm_asr_b200.harness.TrainStep: pinned host waveforms -> device, wav2spectro, 34 x SS2D body (in_proj -> conv+SiLU+x_proj head kernel -> fused core -> merge+LayerNorm+gate tail kernel -> out_proj; paired) forward + "
                      "backward, spectro2wav (+ backward), L1 + multi-resolution STFT loss read back to the host, gradient all-reduce, AdamW; the step's "
                      "algorithmic scan bytes over its wall time"}

    stft_t = None
    if not args.no_stft and rank == 0:
        try:
            stft_t = {f"B{b}": stft_times(wl, device, b) for b in (wl.batch, 64)}
        except Exception as e:
            stft_t = {"error": str(e)[:200]}

    # DRAM bytes per launch of the dominant kernel come from an ncu pass over THIS command (tools/gpu_profile.sh ->
    # tools/launch_list.py), committed under profiles/: a run under a profiler is never a bench value, so it is not measured here
    traffic, traffic_note = None, "no ncu launch list of this command under profiles/"
    try:
        with open(os.path.join(ROOT, "profiles", "r2_dominant_kernel_traffic.json")) as f:
            tr = json.load(f)
        traffic = int(tr["traffic_bytes_per_launch"])
        traffic_note = (f"profiles/r2_dominant_kernel_traffic.json: {tr['source']}, mean of {tr['launches_averaged']} launches "
                        "(ncu pass of this same command; not measured inside this run)")
    except Exception:
        pass
    if rank == 0:
        out = {
            "metric": "ss2d_scan_fwd_bwd_GBps", "value": round(value, 1), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "run": {
                "algorithmic_bytes_per_step": step_bytes, "scan_elements_per_step": wl.scan_elements(),
                "l2": "inputs larger than L2: every call has its own buffers, ~5.4 GB touched per step vs 126 MB L2",
                "launch": "eager launches" if args.no_graph else "one CUDA graph per step", "pairing": args.pairing,
                "zero_fill_bytes_per_step": zero_fill_bytes,  # accumulated gradients (dA, dD, ddelta_bias; dB / dC unless stored)
                "parallelism": f"dp{world}: one rank per GPU on its own batch shard, no data-path collective" if world > 1 else "single",
            },
            "frac_of_hbm_peak": round(value / world / peak, 4), "hbm_peak_gbs": peak, "hbm_peak_source": peak_src,
            "clocks": clocks,
            "gpu_launches": (len(wl.calls) if args.pairing == "grouped" else 2 * len(wl.calls)) * args.steps,
            "e2e": e2e,
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": traffic,
                         "kernel": "scan_bwd_pipe_kernel<softplus> (csrc/scan_bwd_pipe.cu), every backward launch with seqlen > 2048 "
                                   "(grouped: one launch per pair of calls)",
                         "launches_per_step": k_per_step, "avg_launch_ms": round(k_ms, 5),
                         "avg_algorithmic_bytes_per_launch": int(k_bytes), "peak_source": peak_src,
                         "traffic_note": traffic_note},
            "cpu_baseline": cpu_base,
            "ss2d_core": core, "train": train, "infer": infer, "eager": eager, "stft": stft_t,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
