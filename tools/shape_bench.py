#!/usr/bin/env python
"""Per-shape timing of the hot-path kernels (CUDA events, rotating buffers larger than L2).
    python tools/shape_bench.py [--workload vm_asr_48k_MPD] [--reps 20] [--what scan,cross,stft]
Prints one JSON line per (kernel, shape): ms, algorithmic GB/s, fraction of the measured HBM peak."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import load_peaks, make_call_inputs  # noqa: E402
from vm_asr_b200 import cross, scan, stft, workload as W  # noqa: E402


def timeit(fn, reps, n_sets):
    """GPU time per call: `reps` calls over rotating buffer sets captured in one CUDA graph (no host launch gaps)."""
    for i in range(n_sets):
        fn(i)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(n_sets):
            fn(i)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for i in range(reps):
                fn(i % n_sets)
    torch.cuda.current_stream().wait_stream(side)
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="vm_asr_48k_MPD")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--what", default="scan,cross,stft")
    ap.add_argument("--dtype", default="float32")
    ap.add_argument("--pairs", action="store_true", help="scan only: the grouped launch of two same-shape calls, as the bench step issues them")
    args = ap.parse_args()
    wl = W.WORKLOADS[args.workload]
    peak, _ = load_peaks()
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(0)
    B = wl.batch
    dt = getattr(torch, args.dtype)
    es = 4 if dt == torch.float32 else 2
    rows = []
    for call, count in W.distinct_shapes(wl):
        D, L, C = call.D, call.L, call.d_inner
        fwd_bytes = es * (3 * B * D * L + 2 * B * 4 * L)
        bwd_bytes = es * (5 * B * D * L + 4 * B * 4 * L)
        n_sets = max(2, min(8, int(400e6 // max(fwd_bytes, 1)) + 1))  # rotate > 126 MB of distinct data
        if args.pairs:
            n_sets = 2 * max(2, n_sets // 2 + 1)
        if "scan" in args.what:
            sets = []
            for _ in range(n_sets):
                inp = make_call_inputs(call, B, dev, gen)
                for k in ("u", "delta", "B", "C", "dout"):
                    inp[k] = inp[k].to(dt)
                n_chunks = (L + 2047) // 2048
                # as the step issues them: dB / dC stored where one tile spans the group (VMASR_SCAN_DBDC_STORE), otherwise summed
                # into a buffer that the forward launch clears as its side job (fp32 fast path; zeros allocated once otherwise)
                store = dt == torch.float32 and scan.dbdc_store_candidate(inp["u"], inp["A"], inp["B"])
                bc = torch.zeros(2 * B * 4 * L, device=dev) if not store else None
                dB, dC = ((torch.empty(B, 4, 1, L, device=dev) for _ in range(2)) if store
                          else (bc[:B * 4 * L].view(B, 4, 1, L), bc[B * 4 * L:].view(B, 4, 1, L)))
                bufs = dict(out=torch.empty_like(inp["u"]), x=torch.empty(B, D, n_chunks, 2, device=dev),
                            du=torch.empty_like(inp["u"]), ddelta=torch.empty_like(inp["u"]),
                            dA=torch.zeros(D, 1, device=dev), dD=torch.zeros(D, device=dev), dbias=torch.zeros(D, device=dev),
                            dB=dB, dC=dC, bc=bc, flags=scan.SCAN_DBDC_STORE if store else 0)
                sets.append((inp, bufs))

            def f(i):
                inp, b = sets[i]
                scan.fwd_out(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["bias"], True, b["out"], b["x"], zero=b["bc"])

            def g(i):
                inp, b = sets[i]
                scan.bwd_out(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["bias"], inp["dout"], b["x"],
                             True, b["du"], b["ddelta"], b["dA"], b["dB"], b["dC"], b["dD"], b["dbias"], flags=b["flags"])

            def fp(i):
                a, b = sets[2 * i], sets[2 * i + 1]
                scan.fwd_grouped([(s_[0]["u"], s_[0]["delta"], s_[0]["A"], s_[0]["B"], s_[0]["C"], s_[0]["D"], s_[0]["bias"], True) for s_ in (a, b)],
                                 [(s_[1]["out"], s_[1]["x"]) for s_ in (a, b)], zero=[s_[1]["bc"] for s_ in (a, b)])

            def gp(i):
                a, b = sets[2 * i], sets[2 * i + 1]
                scan.bwd_grouped([(s_[0]["u"], s_[0]["delta"], s_[0]["A"], s_[0]["B"], s_[0]["C"], s_[0]["D"], s_[0]["bias"], s_[0]["dout"],
                                   s_[1]["x"], True, s_[1]["flags"]) for s_ in (a, b)],
                                 [(s_[1]["du"], s_[1]["ddelta"], s_[1]["dA"], s_[1]["dB"], s_[1]["dC"], s_[1]["dD"], s_[1]["dbias"]) for s_ in (a, b)])

            if args.pairs:
                for name, fn, nbytes in (("scan_fwd_pair", fp, 2 * fwd_bytes), ("scan_bwd_pair", gp, 2 * bwd_bytes)):
                    ms = timeit(fn, args.reps, n_sets // 2)
                    rows.append(dict(kernel=name, B=B, D=D, L=L, launches_per_step=count // 2, ms=round(ms, 5),
                                     GBps=round(nbytes / ms / 1e6, 1), frac=round(nbytes / ms / 1e6 / peak, 3)))
                    print(json.dumps(rows[-1]), flush=True)
                del sets
                torch.cuda.empty_cache()
                continue
            for name, fn, nbytes in (("scan_fwd", f, fwd_bytes), ("scan_bwd", g, bwd_bytes)):
                ms = timeit(fn, args.reps, n_sets)
                rows.append(dict(kernel=name, B=B, D=D, L=L, calls=count, ms=round(ms, 5), GBps=round(nbytes / ms / 1e6, 1),
                                 frac=round(nbytes / ms / 1e6 / peak, 3)))
                print(json.dumps(rows[-1]), flush=True)
            del sets
        if "cross" in args.what and dt == torch.float32:
            nb = 4 * 5 * B * C * L
            n_sets = max(2, min(8, int(400e6 // nb) + 1))
            xs_in = [torch.randn(B, C, call.H, call.W, device=dev) for _ in range(n_sets)]
            ys_in = [torch.randn(B, 4, C, call.H, call.W, device=dev) for _ in range(n_sets)]
            for name, fn in (("cross_scan", lambda i: cross.cross_scan(xs_in[i])),
                             ("cross_merge", lambda i: cross.cross_merge(ys_in[i], call.H, call.W))):
                ms = timeit(fn, args.reps, n_sets)
                rows.append(dict(kernel=name, B=B, C=C, H=call.H, W=call.W, calls=count, ms=round(ms, 5),
                                 GBps=round(nb / ms / 1e6, 1), frac=round(nb / ms / 1e6 / peak, 3)))
                print(json.dumps(rows[-1]), flush=True)
            del xs_in, ys_in
        torch.cuda.empty_cache()
    if "stft" in args.what:
        for Bs in (wl.batch, 64):
            F, Nf = wl.n_fft // 2 + 1, 1 + wl.T // wl.hop
            wave = [0.1 * torch.randn(Bs, 1, wl.T, device=dev) for _ in range(4)]
            mp = [stft.wav2spectro(w, wl.n_fft, wl.hop, wl.win, "log2") for w in wave]
            nb = 4 * Bs * wl.T + 8 * Bs * F * Nf
            ms = timeit(lambda i: stft.wav2spectro(wave[i], wl.n_fft, wl.hop, wl.win, "log2"), args.reps, 4)
            print(json.dumps(dict(kernel="stft_fwd", B=Bs, T=wl.T, n_fft=wl.n_fft, hop=wl.hop, ms=round(ms, 5),
                                  GBps=round(nb / ms / 1e6, 1), frac=round(nb / ms / 1e6 / peak, 3))), flush=True)
            ms = timeit(lambda i: stft.spectro2wav(mp[i][0], mp[i][1], wl.n_fft, wl.hop, wl.win, "log2"), args.reps, 4)
            print(json.dumps(dict(kernel="istft_fwd", B=Bs, T=wl.T, n_fft=wl.n_fft, hop=wl.hop, ms=round(ms, 5),
                                  GBps=round(nb / ms / 1e6, 1), frac=round(nb / ms / 1e6 / peak, 3))), flush=True)
            # torch baseline on the same box (library path: cuFFT + elementwise kernels)
            win = torch.hann_window(wl.win, device=dev)

            def torch_stft(i):
                s = torch.stft(wave[i].reshape(-1, wl.T), wl.n_fft, wl.hop, wl.win, win, normalized=True, return_complex=True)
                return torch.log2(s.abs() + 1e-8), torch.angle(s)

            ms = timeit(torch_stft, args.reps, 4)
            print(json.dumps(dict(kernel="torch.stft+log2+angle (library baseline)", B=Bs, ms=round(ms, 5),
                                  GBps=round(nb / ms / 1e6, 1))), flush=True)


if __name__ == "__main__":
    main()
