#!/usr/bin/env python
"""bench.py -- headline measurement of the SS2D hot path on B200.

One STEP = the selective-scan work of one VM-ASR training step of ``configs/vm_asr_48k_MPD.yaml`` (BASELINE.json
configs[1]): the 34 SS2D selective-scan forwards in forward order, then the 34 backwards in reverse order, at the
config's batch (4 clips, fp32 scan IO as the model forces, d_state 1, 4 B/C groups), on synthetic inputs with
the reference test's distributions.  Metric: algorithmic scan bytes per second, whole job.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

* ``value``  : inputs resident in HBM; the step is captured once in a CUDA graph (68 kernel launches + one memset
               that zeroes the gradient accumulators) and replayed; timed with CUDA events, max over ranks.
* ``e2e``    : the same step through the public operator API (``selective_scan_cuda_core``-style ``fwd``/``bwd``),
               every call's inputs copied from pinned host memory and its outputs copied back, inside the timed
               region.
* ``roofline``: the dominant kernel (backward, multi-chunk variant scan_bwd_pipe_kernel) timed launch by launch with CUDA events.
* ``cpu_baseline`` / ``--impl reference``: the oracle's torch restatement of the reference's pure-PyTorch path
               (``selective_scan_ref`` + autograd-equivalent closed form) on the host cores, on a bounded sample.
* N > 1 (torchrun): every rank runs the step on its own batch shard (weak scaling, no data-path collective);
               the 3.01 M-parameter generator gradient buffer that carries the scan parameters' gradients is
               all-reduced over NCCL each step, as data-parallel training would.
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

GEN_PARAMS = 3_010_000  # generator parameters (README.md:8), the data-parallel gradient payload


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
# workload buffers
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


class DeviceStep:
    """All buffers of one step on the device (distinct per call: ~5.4 GB touched per step >> 126 MB L2) and the
    CUDA graph that replays it."""

    def __init__(self, wl, device, pairing="grouped"):
        from vm_asr_b200 import scan
        self.scan, self.wl, self.device, self.pairing = scan, wl, device, pairing
        self.side = None
        gen = torch.Generator(device=device).manual_seed(1234)
        self.calls = []
        acc_floats = 0
        for c in wl.calls:
            acc_floats += c.D * 3 + 2 * wl.batch * 4 * c.L
        # one arena for every accumulated gradient (dA, dD, ddelta_bias, dB, dC of all calls): zeroed by ONE memset
        self.arena = torch.zeros(acc_floats, dtype=torch.float32, device=device)
        off = 0

        def take(n, shape):
            nonlocal off
            t = self.arena[off:off + n].view(*shape)
            off += n
            return t

        for c in wl.calls:
            inp = make_call_inputs(c, wl.batch, device, gen)
            n_chunks = (c.L + 2047) // 2048
            bufs = dict(
                out=torch.empty_like(inp["u"]), x=torch.empty(wl.batch, c.D, n_chunks, 2, device=device),
                du=torch.empty_like(inp["u"]), ddelta=torch.empty_like(inp["u"]),
                dA=take(c.D, (c.D, 1)), dD=take(c.D, (c.D,)), dbias=take(c.D, (c.D,)),
                dB=take(wl.batch * 4 * c.L, (wl.batch, 4, 1, c.L)), dC=take(wl.batch * 4 * c.L, (wl.batch, 4, 1, c.L)),
            )
            self.calls.append((c, inp, bufs))
        self.graph = None

    def fwd_call(self, i):
        c, inp, b = self.calls[i]
        self.scan.fwd_out(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["bias"], True, b["out"], b["x"])

    def bwd_call(self, i):
        c, inp, b = self.calls[i]
        self.scan.bwd_out(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["bias"], inp["dout"], b["x"],
                          True, b["du"], b["ddelta"], b["dA"], b["dB"], b["dC"], b["dD"], b["dbias"])

    # The generator's two streams (magnitude, phase) issue the same-shape SS2D call independently between their interaction
    # points (model/model.py:1124-1127, 1167-1176): calls 2j and 2j + 1 of the workload are such a pair.
    def fwd_pair(self, i):
        args, outs = [], []
        for k in (i, i + 1):
            c, inp, b = self.calls[k]
            args.append((inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["bias"], True))
            outs.append((b["out"], b["x"]))
        self.scan.fwd_grouped(args, outs)

    def bwd_pair(self, i):
        args, outs = [], []
        for k in (i, i + 1):
            c, inp, b = self.calls[k]
            args.append((inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["bias"], inp["dout"], b["x"], True))
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
        self.run_eager()  # creates this stream's carry workspace before capture
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self.run_eager()
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                self.run_eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()

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


class HostStep:
    """The step through the public API with HOST buffers: pinned inputs -> device, fwd/bwd, outputs -> pinned host."""

    def __init__(self, wl, device):
        from vm_asr_b200 import scan
        self.scan, self.wl, self.device = scan, wl, device
        gen = torch.Generator().manual_seed(99)
        self.host_in, self.host_out = [], []
        self.h2d = self.d2h = 0
        for c in wl.calls:
            inp = make_call_inputs(c, wl.batch, "cpu", gen, pinned=True)
            self.host_in.append(inp)
            outs = {}
            for name, shape in (("out", (wl.batch, c.D, c.L)), ("du", (wl.batch, c.D, c.L)), ("ddelta", (wl.batch, c.D, c.L)),
                                ("dB", (wl.batch, 4, 1, c.L)), ("dC", (wl.batch, 4, 1, c.L)), ("dA", (c.D, 1)), ("dD", (c.D,)),
                                ("dbias", (c.D,))):
                outs[name] = torch.empty(shape, dtype=torch.float32, pin_memory=True)
                self.d2h += outs[name].numel() * 4
            self.host_out.append(outs)
            self.h2d += sum(t.numel() * 4 for t in inp.values())
        self.copy_in, self.copy_out = torch.cuda.Stream(), torch.cuda.Stream()

    def step(self):
        """Copies and kernels are ordered so that PCIe runs in both directions at once: forward inputs go up in call
        order and every forward output goes down as soon as its call is done; `dout` goes up in reverse call order
        behind them and every backward call's gradients go down as soon as they exist.  Consecutive steps overlap
        the same way (the next step's uploads do not wait for this step's downloads)."""
        main = torch.cuda.current_stream()
        n = len(self.host_in)
        dev_in, saved = [None] * n, [None] * n
        ready_fwd, ready_bwd = [None] * n, [None] * n
        with torch.cuda.stream(self.copy_in):
            for i, inp in enumerate(self.host_in):
                dev_in[i] = {k: v.to(self.device, non_blocking=True) for k, v in inp.items() if k != "dout"}
                for v in dev_in[i].values():
                    v.record_stream(main)
                ready_fwd[i] = torch.cuda.Event()
                ready_fwd[i].record()
            for i in reversed(range(n)):
                dout = self.host_in[i]["dout"].to(self.device, non_blocking=True)
                dout.record_stream(main)
                dev_in[i]["dout"] = dout
                ready_bwd[i] = torch.cuda.Event()
                ready_bwd[i].record()

        def download(i, outs):
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(self.copy_out):
                self.copy_out.wait_event(ev)
                for k, v in outs.items():
                    v.record_stream(self.copy_out)
                    self.host_out[i][k].copy_(v, non_blocking=True)

        for i in range(n):
            d = dev_in[i]
            main.wait_event(ready_fwd[i])
            out, x = self.scan.fwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["bias"], True, 1)
            saved[i] = x
            download(i, {"out": out})
        for i in reversed(range(n)):
            d = dev_in[i]
            main.wait_event(ready_bwd[i])
            du, ddelta, dA, dB, dC, dD, dbias = self.scan.bwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["bias"],
                                                             d["dout"], saved[i], True, 1)
            download(i, dict(du=du, ddelta=ddelta, dA=dA, dB=dB, dC=dC, dD=dD, dbias=dbias))
        # no join here: the next step's uploads and kernels may start while this step's last gradients are still on
        # their way down (stream order keeps the pinned output buffers consistent); the timed loop ends with a device sync


# ---------------------------------------------------------------------------------------------------------
# CPU reference arm (the oracle's torch restatement of the reference's pure-PyTorch path)
# ---------------------------------------------------------------------------------------------------------
def cpu_sample_step(wl, sample_len):
    """fwd+bwd over the workload's distinct shapes truncated to their first `sample_len` positions; returns the
    algorithmic bytes processed."""
    from oracle import ss2d_ref
    from vm_asr_b200 import workload as W
    gen = torch.Generator().manual_seed(5)
    total = 0
    for call, _count in W.distinct_shapes(wl):
        L = min(call.L, sample_len)
        sub = W.SS2DCall(call.d_inner, 1, L)
        inp = make_call_inputs(sub, wl.batch, "cpu", gen)
        ss2d_ref.selective_scan(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["bias"], True)
        ss2d_ref.selective_scan_bwd(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["bias"], True,
                                    inp["dout"], dtype=torch.float32)
        total += 4 * (8 * wl.batch * call.D * L + 6 * wl.batch * 4 * L)
    return total


def run_cpu_reference(wl, steps, warmup, sample_len):
    torch.set_num_threads(os.cpu_count() or 1)
    for _ in range(warmup):
        cpu_sample_step(wl, sample_len)
    t0 = time.perf_counter()
    nbytes = 0
    for _ in range(steps):
        nbytes += cpu_sample_step(wl, sample_len)
    dt = time.perf_counter() - t0
    return nbytes / dt / 1e9, dt / steps * 1e3, torch.get_num_threads()


# ---------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="vm_asr_48k_MPD")
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--cpu-sample-len", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly (profiling aid)")
    ap.add_argument("--pairing", default="grouped", choices=["grouped", "streams", "none"],
                    help="how the two generator streams' same-shape calls are issued: one grouped launch per pair (default), "
                         "two CUDA streams, or one call after the other")
    args = ap.parse_args()

    from vm_asr_b200 import workload as W
    wl = W.WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    fwd_b, bwd_b = wl.scan_bytes()
    step_bytes = fwd_b + bwd_b
    sample_desc = (f"scan fwd+bwd over the workload's 6 distinct (B, D) shapes, sequence truncated to its first "
                   f"{args.cpu_sample_len} positions, fp32 torch ops, python loop over L (selective_scan_ref style)")

    if args.impl == "reference":
        if rank != 0:
            return
        warm = max(1, min(args.warmup, 2))
        steps = max(1, min(args.steps, 5))
        gbs, ms, cores = run_cpu_reference(wl, steps, warm, args.cpu_sample_len)
        print(json.dumps({
            "impl": "reference", "metric": "ss2d_scan_fwd_bwd_GBps", "value": round(gbs, 4), "unit": "GB/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": round(ms, 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{wl.yaml} SS2D selective-scan fwd+bwd (34 calls), batch {wl.batch}", "sample": sample_desc},
            "cpu_baseline": {"value": round(gbs, 4), "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample_desc},
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
    if rank == 0 and not args.no_cpu_baseline:
        gbs, ms, cores = run_cpu_reference(wl, 2, 1, args.cpu_sample_len)
        cpu_base = {"value": round(gbs, 4), "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample_desc,
                    "ms_per_sample_step": round(ms, 1)}

    ds = DeviceStep(wl, device, args.pairing)
    if args.no_graph:
        ds.step = ds.run_eager
        ds.run_eager()
    else:
        ds.capture()
    grad_flat = torch.zeros(GEN_PARAMS, dtype=torch.float32, device=device) if world > 1 else None

    def one_step():
        ds.step()
        if world > 1:
            return dist.all_reduce(grad_flat, async_op=True)
        return None

    for _ in range(max(args.warmup, 3)):
        h = one_step()
        if h is not None:
            h.wait()
    sampler = ClockSampler(local_rank)
    sampler.start()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    e0.record()
    handles = []
    for _ in range(args.steps):
        handles.append(one_step())
    for h in handles:
        if h is not None:
            h.wait()
    e1.record()
    torch.cuda.synchronize()
    t_wall1 = time.perf_counter()
    if world > 1:
        dist.barrier()
    sampler.stop()
    elapsed_ms = e0.elapsed_time(e1)
    elapsed_ms = vdist.max_over_ranks(elapsed_ms, device)  # timing rule: max over ranks
    ms_per_step = elapsed_ms / args.steps
    value = world * step_bytes / (ms_per_step * 1e-3) / 1e9
    clocks = sampler.summary(t_wall0, t_wall1)

    # dominant kernel, launch by launch
    k_ms, k_bytes, k_per_step = time_dominant_kernel(ds, max(2, min(args.steps, 5)))
    achieved = k_bytes / (k_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("traffic_bytes_per_launch")
        except Exception:
            traffic = None

    e2e = None
    if not args.no_e2e:
        hs = HostStep(wl, device)
        hs.step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            hs.step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        dt = vdist.max_over_ranks(dt, device)
        e2e = {"value": round(world * step_bytes / dt / 1e9, 2), "unit": "GB/s", "h2d_bytes_per_step": hs.h2d,
               "d2h_bytes_per_step": hs.d2h, "ms_per_step": round(dt * 1e3, 2), "steps": args.e2e_steps,
               "api": "vm_asr_b200.scan.fwd/bwd (selective_scan_cuda_core surface), pinned host buffers, copies on side streams"}
        del hs

    if rank == 0:
        out = {
            "metric": "ss2d_scan_fwd_bwd_GBps", "value": round(value, 1), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"{wl.yaml} SS2D selective-scan fwd+bwd of one training step (34 fwd + 34 bwd calls), "
                            f"batch {wl.batch} per GPU, d_state 1, 4 groups",
                "algorithmic_bytes_per_step": step_bytes, "scan_elements_per_step": wl.scan_elements(),
                "l2": "inputs larger than L2: every call has its own buffers, ~5.4 GB touched per step vs 126 MB L2",
                "launch": "eager launches" if args.no_graph else "one CUDA graph per step", "pairing": args.pairing, "parallelism": f"dp{world}" if world > 1 else "single",
                "collective": "NCCL all-reduce of a 3.01M-float gradient buffer per step" if world > 1 else "none",
            },
            "frac_of_hbm_peak": round(value / world / peak, 4), "hbm_peak_gbs": peak, "hbm_peak_source": peak_src,
            "audio_sec_per_s_hot_path": round(world * wl.batch * wl.clip_seconds / (ms_per_step * 1e-3), 1),
            "clocks": clocks,
            "gpu_launches": ds.gpu_launches_per_step * args.steps,
            "e2e": e2e,
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": traffic,
                         "kernel": "scan_bwd_pipe_kernel<softplus> (csrc/scan_bwd_pipe.cu), every backward call with seqlen > 2048 (18 of the 34 calls, 85 % of the backward bytes)", "launches_per_step": k_per_step,
                         "avg_launch_ms": round(k_ms, 5), "avg_algorithmic_bytes_per_launch": int(k_bytes),
                         "peak_source": peak_src},
            "cpu_baseline": cpu_base,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
