// Selective-scan forward, fast path for sequences of more than one chunk (fp32 IO, d_state 1, 16-byte aligned rows).
// Replaces selective_scan_fwd_kernel (kernels/selective_scan/csrc/selective_scan/cus/selective_scan_fwd_kernel.cuh:61-203),
// whose CTA walks the chunks of a row serially (:102); here every chunk is its own CTA (see scan.cuh / pipe.cuh).
//
// What this kernel adds to scan_fwd_tma.cu (which keeps the single-chunk shapes): no CTA-wide barrier around the carry
// exchange, and the exchange runs on a warp of its own.  ncu on the barrier version showed the warps of a CTA parked on
// `bar.sync` for most of every channel iteration while the first warp resolved the cross-chunk look-back (stall_barrier
// 7-13 cycles per issue, issue slots 20-30 % busy); a first pipelined version (outputs of channel j - 1 after the local
// scan of channel j) still had 30 % of its stall samples on the hand-off, because the aggregate a chunk needs from its
// left neighbour is published by a CTA running in lock-step with it, one cross-SM round trip (> 1 us) away.
// So the whole tile (<= 4 channels x 2048 positions, 64 KB) is RESIDENT in shared memory and the work is two sweeps:
//
//   compute warps (8; one thread owns 8 positions):
//     P1(j), j = 0..n-1   wait for channel j's u / delta (TMA bulk copies, all issued at kernel start), element-wise work,
//                in-thread scan, warp scan; the warp's aggregate goes to shared memory and the warp ARRIVES on mbarrier
//                tot[j] -- it waits for nothing.  Every output is affine in the state entering the thread,
//                y_l = Y0_l + Y1_l * h_in,  so P1 leaves just the two arrays  Y0 = C (local state) + D u  and
//                Y1 = C (local decay product)  in place of u / delta, plus two registers (the thread's warp-exclusive prefix).
//     P2(j), j = 0..n-1   wait on mbarrier in[j], y = Y0 + Y1 * h_in straight from shared memory, 128-bit stores.
//   exchange warp:
//     sweep A    per channel: wait on tot[j], combine the 8 warp aggregates, PUBLISH the chunk's aggregate, issue the
//                look-back loads (pipe.cuh) -- the loads of all channels are in flight together;
//     sweep B    per channel: reduce the look-back, write the chunk-end state to `x`, write the state entering each warp
//                to shared memory, arrive on in[j].
//
// By the time a CTA asks for its neighbours' aggregates (sweep B) it has published all of its own, and so have they.
#include <cstdlib>

#include "fast.cuh"

namespace vmasr {

constexpr int kPipeStages = 4;        // channels per tile = resident stages (u + delta, 16 KB per channel)
constexpr int kPipeThreads = 288;     // 8 compute warps + the exchange warp

// `chunk` is the chunk's index in TIME order; with REV (time runs against memory order) it sits at the mirrored place in memory.
// F1 (delta on the fly, dt_rank 1; include/vmasr_b200.h): no delta row is copied in.  While the tile loads, the second half
// of every stage is free (it only receives Y1 at the end of P1), so the dt row, B and C arrive in the second halves of stages
// 0, 1, 2 and EVERY u row is requested at kernel start (no borrowed stage).  P1(j) finds the dt row in the second half of
// stage j, forms delta = w_j * row in registers, and before its own Y1 overwrites the row hands it down to stage j + 1 (a
// thread only ever touches its own 32 bytes of a row, so the hand-down needs no synchronisation).
template <bool TAIL, bool SP, bool REV, bool F1>
__device__ __forceinline__ void scan_fwd_pipe_body(const ScanArgs &a, const TileMaps &tm, unsigned char *smem, const int chunk, const int rg,
                                                   const int half, const int ztile) {
    constexpr int NC = 256, ITEMS = 8, WPR = 8, SEG = NC * ITEMS, STAGES = kPipeStages;

    // shared memory carve-up (header 2048 bytes)
    unsigned long long *bar_full = reinterpret_cast<unsigned long long *>(smem);  // [STAGES] TMA completion
    unsigned long long *bar_bc = bar_full + STAGES;                               // B / C segment
    unsigned long long *bar_tot = bar_bc + 1;                                     // [STAGES] 8 arrivals: warp totals written
    unsigned long long *bar_in = bar_tot + STAGES;                                // [STAGES] 8 arrivals (exchange-warp lanes): entering states written
    unsigned long long *bar_free = bar_in + STAGES;                               // 8 arrivals: B / C are in registers, their slot is free
    float2 *s_tot = reinterpret_cast<float2 *>(smem + 256);                       // [STAGES][8] warp totals (p, q)
    float *s_in = reinterpret_cast<float *>(smem + 512);                          // [STAGES][8] state entering each warp
    float *s_par = reinterpret_cast<float *>(smem + 1024);                        // [3][STAGES]
    float *s_stage = reinterpret_cast<float *>(smem + 2048);                      // [STAGES][2][SEG]
    float *s_bc = s_stage + (size_t)(STAGES - 1) * 2 * SEG;                       // B, C: borrowed from the last stage (not F1)
    auto y1_slot = [&](int st) { return s_stage + (size_t)st * 2 * SEG + SEG; };  // second half of stage st
    float2 *s_exc = reinterpret_cast<float2 *>(s_stage + (size_t)STAGES * 2 * SEG);  // [STAGES][NC] warp-exclusive prefix of every thread

    const int ctile = rg % a.n_ctiles;
    const int bg = rg / a.n_ctiles;
    const int g = bg % a.ngroups;
    const int b = bg / a.ngroups;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const bool exchange = warp == WPR;  // the ninth warp
    const int L = a.seqlen;
    const int seg0 = (REV ? a.n_chunks - 1 - chunk : chunk) * SEG;  // first MEMORY position of the tile
    const int line0 = seg0 / kTileLine;                             // ... and its first line in the tensor maps
    constexpr unsigned seg_bytes = SEG * 4u;                        // a box always counts in full (lines past the end arrive as zeros)

    // half >= 0: one half of a planned tile (the launch's last round, scan_host.cu::split_last_round)
    const int c_begin = ctile * a.chan_per_tile + (half > 0 ? a.chan_per_tile >> 1 : 0);
    const int n_iter = half >= 0 ? a.chan_per_tile >> 1 : min(a.chan_per_group, c_begin + a.chan_per_tile) - c_begin;  // channels of this tile (<= STAGES)
    const int d0 = g * a.chan_per_group + c_begin;

    auto issue_stage = [&](int it) {  // lane 0 of the exchange warp only
        float *dst = s_stage + (size_t)it * 2 * SEG;
        mbar_expect_tx(&bar_full[it], 2u * seg_bytes);
        tensor_load(dst, &tm.u, line0, d0 + it, b, &bar_full[it]);
        tensor_load(dst + SEG, &tm.delta, line0, d0 + it, b, &bar_full[it]);
    };
    if (threadIdx.x == 0) {
        VMASR_TL(a, 0);
#ifdef VMASR_TUNING
        if (a.timeline) { unsigned sm_; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_)); a.timeline[(size_t)blockIdx.x * 16 + 15] = sm_; }
#endif
    }
    // the bulk copies go out first: they do not depend on the per-channel parameters staged below
    if (threadIdx.x == NC) {
#pragma unroll
        for (int i = 0; i < STAGES + 1; ++i) mbar_init(&bar_full[i], 1);
#pragma unroll
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&bar_tot[i], WPR);
            mbar_init(&bar_in[i], WPR);
        }
        mbar_init(bar_free, WPR);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_wait();  // every thread, in front of its first access to global memory
    zero_side_region(a, ztile);
    if (threadIdx.x == NC) {
        if (F1) {
            mbar_expect_tx(bar_bc, 3u * seg_bytes);
            tensor_load(y1_slot(0), &tm.delta, line0, g, b, bar_bc);  // dt row of this group (dt_rank 1: row index = group)
            tensor_load(y1_slot(1), &tm.B, line0, g, b, bar_bc);
            tensor_load(y1_slot(2), &tm.C, line0, g, b, bar_bc);
#pragma unroll
            for (int s = 0; s < STAGES; ++s)
                if (s < n_iter) {
                    mbar_expect_tx(&bar_full[s], seg_bytes);
                    tensor_load(s_stage + (size_t)s * 2 * SEG, &tm.u, line0, d0 + s, b, &bar_full[s]);
                }
        } else {
            mbar_expect_tx(bar_bc, 2u * seg_bytes);
            tensor_load(s_bc, &tm.B, line0, g, b, bar_bc);
            tensor_load(s_bc + SEG, &tm.C, line0, g, b, bar_bc);
#pragma unroll
            for (int s = 0; s < STAGES - 1; ++s)
                if (s < n_iter) issue_stage(s);
        }
    }
    if (threadIdx.x < (F1 ? 4 : 3) * n_iter) {
        const int which = threadIdx.x / n_iter, cc = threadIdx.x - which * n_iter;
        const int d = d0 + cc;
        float v;
        if (which == 0) v = __ldg(a.A + d * a.A_ds);
        else if (which == 1) v = a.D ? __ldg(a.D + d) : 0.0f;
        else if (which == 2) v = (a.delta_bias ? __ldg(a.delta_bias + d) : 0.0f) * kLog2e;
        else v = __ldg(a.dt_w + d * a.dtw_ds) * kLog2e;  // dt weight (rank 1), log2 domain like the bias
        s_par[which * STAGES + cc] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) VMASR_TL(a, 1);

    const long long seq0 = (long long)b * a.dim + d0;  // (batch, channel) row of the tile's first channel
    if (exchange) {
        // ================= exchange warp =================
        if (!F1 && STAGES - 1 < n_iter) {
            mbar_wait(bar_free, 0);  // the compute warps hold B / C in registers: the last stage is free for data now
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue_stage(STAGES - 1);
            }
        }
        const unsigned epoch = launch_epoch(a, lane);  // (also recycles the carry workspace for the next launch: pipe.cuh)
        const int n_groups16 = (a.n_chunks + 15) >> 4;
        // per channel: publish the chunk aggregate as soon as it exists and start its look-back; finish the look-back of the
        // channel before (its loads have been in flight for one P1 of the compute warps)
        auto finish = [&](int j, CarryLook &l, const Aff &cm) {
            const long long seq = seq0 + j;
            CarryEntry *l2_row = a.ws_entries2 + seq * n_groups16;
            const Aff before = shift_up1(cm, lane);
            const Aff total = {__shfl_sync(0xffffffffu, cm.p, WPR - 1), __shfl_sync(0xffffffffu, cm.q, WPR - 1)};
            bool ok;
            Aff grp = {1.0f, 0.0f};
            Aff acc = look_reduce(l, epoch, lane, ok, grp);
            if (!a.debug_nowait) acc = look_finish(l, acc, ok, l2_row, chunk, epoch, lane, grp);
            if (lane == 0) {
                if ((chunk & 15) == 15) {
                    const Aff g16 = compose(grp, total);
                    publish_entry(l2_row + (chunk >> 4), epoch, g16.p, g16.q);
                }
                reinterpret_cast<float2 *>(a.x)[seq * a.n_chunks + chunk] = make_float2(total.p * acc.p, fmaf(total.p, acc.q, total.q));
            }
            if (lane < WPR) {  // every writer releases its own store
                s_in[j * WPR + lane] = fmaf(before.p, acc.q, before.q);
                mbar_arrive(&bar_in[j]);
            }
        };
        CarryLook p_look;
        p_look.ptr = nullptr;
        p_look.e = make_uint4(0u, 0u, 0u, 0u);
        Aff p_cum = {1.0f, 0.0f};
#pragma unroll 1
        for (int j = 0; j < n_iter; ++j) {
            const long long seq = seq0 + j;
            CarryEntry *l1_row = a.ws_entries + seq * a.n_chunks;
            mbar_wait(&bar_tot[j], 0);
            if (lane == 0) VMASR_TL(a, j == 0 ? 8 : 9);
            const float2 t = (lane < WPR) ? s_tot[j * WPR + lane] : make_float2(1.0f, 0.0f);
            const Aff cum = warp_scan_up_fast<WPR>(Aff{t.x, t.y});
            if (lane == WPR - 1) publish_entry(l1_row + chunk, epoch, cum.p, cum.q);
            const CarryLook look = look_issue(l1_row, a.ws_entries2 + seq * n_groups16, chunk, lane);
            if (j >= 1) finish(j - 1, p_look, p_cum);
            p_look = look;
            p_cum = cum;
        }
        finish(n_iter - 1, p_look, p_cum);
        if (lane == 0) VMASR_TL(a, 10);
    } else {
        // ================= compute warps =================
        const int tseg = REV ? NC - 1 - (int)threadIdx.x : (int)threadIdx.x;  // this thread's 8-position segment of the tile (memory order)
        const int pos = seg0 + tseg * ITEMS;
        const int sel = swz_half(tseg), slot = swz_slot(tseg) * ITEMS;  // where the 64-byte swizzle puts this thread's 32 bytes (scan.cuh)
        int nvalid = ITEMS;
        if (TAIL) nvalid = max(0, min(ITEMS, L - pos));
        float *out_ptr = reinterpret_cast<float *>(a.out) + b * a.out_bs + (long long)d0 * a.out_ds + pos;
        const bool accum = a.accum == 1, addm = a.accum == 2;  // red.add / load-add-store (scan.cuh)

        float2 Bl[4], Cv[4];  // ln2 * B (the scan runs on dt in the log2 domain) and C of this thread's positions
        mbar_wait(bar_bc, 0);
        if (threadIdx.x == 0) VMASR_TL(a, 2);
        lds8_priv((F1 ? y1_slot(1) : s_bc) + slot, sel, Bl);
        lds8_priv((F1 ? y1_slot(2) : s_bc + SEG) + slot, sel, Cv);
        if (!F1) {
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_free);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            Bl[k] = mul2(Bl[k], f2(kLn2));
            if (TAIL) {  // positions past the end: keep the arithmetic finite
                if (2 * k >= nvalid) { Bl[k].x = 0.0f; Cv[k].x = 0.0f; }
                if (2 * k + 1 >= nvalid) { Bl[k].y = 0.0f; Cv[k].y = 0.0f; }
            }
        }

#pragma unroll 1
        for (int j = 0; j < n_iter; ++j) {
            {
                // ---- P1(j) ----
                const float Av = s_par[j];
                const float Dv = s_par[STAGES + j];
                const float bias2 = s_par[2 * STAGES + j];
                float *su = s_stage + (size_t)j * 2 * SEG + slot;
                mbar_wait(&bar_full[j], 0);
                if (threadIdx.x == 0 && j == 0) VMASR_TL(a, 3);
                float2 uv[4], dl[4], Y0[4], Y1[4];
                lds8_priv(su, sel, uv);
                lds8_priv(su + SEG, sel, dl);  // delta, or (F1) the dt row
                float wdt = kLog2e;
                if (F1) {
                    wdt = s_par[3 * STAGES + j];
                    if (j + 1 < n_iter) sts8_priv(su + 2 * SEG + SEG, sel, dl);  // hand the row down to stage j + 1
                }
                float p = 1.0f, q = 0.0f;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const int k = pair_at<REV>(kk);  // pairs in time order
                    if (TAIL) {
                        if (2 * k >= nvalid) { uv[k].x = 0.0f; dl[k].x = 0.0f; }
                        if (2 * k + 1 >= nvalid) { uv[k].y = 0.0f; dl[k].y = 0.0f; }
                    }
                    float2 dt2 = fma2(dl[k], f2(wdt), f2(bias2));
                    if (SP) {
                        float2 e, sp;
                        dt2 = softplus2_pair(dt2, e, sp);
                    }
                    const float2 da = mul2(dt2, f2(Av));
                    float2 av = make_float2(ex2_approx(da.x), ex2_approx(da.y));
                    const float2 bx = mul2(mul2(dt2, Bl[k]), uv[k]);
                    if (TAIL) {  // identity map past the end
                        if (2 * k >= nvalid) av.x = 1.0f;
                        if (2 * k + 1 >= nvalid) av.y = 1.0f;
                    }
                    float2 P, Q;
                    walk_pair<REV>(av, bx, p, q, P, Q);
                    Y0[k] = fma2(Cv[k], Q, mul2(uv[k], f2(Dv)));
                    Y1[k] = mul2(Cv[k], P);
                }
                sts8_priv(su, sel, Y0);
                sts8_priv(su + SEG, sel, Y1);
                const Aff inc = warp_scan_up_fast<32>(Aff{p, q});
                const Aff ex = shift_up1(inc, lane);
                s_exc[j * NC + threadIdx.x] = make_float2(ex.p, ex.q);  // read back by this thread in P2(j): parked in shared memory, not in
                                                                        // STAGES register pairs picked by predicated selects in a rolled loop
                if (lane == 31) s_tot[j * WPR + warp] = make_float2(inc.p, inc.q);
                __syncwarp();
                if (lane == 31) mbar_arrive(&bar_tot[j]);
            }
        }
        if (threadIdx.x == 0) VMASR_TL(a, 4);
#pragma unroll 1
        for (int j = 0; j < n_iter; ++j) {
            {
                // ---- P2(j) ----
                const float2 exv = s_exc[j * NC + threadIdx.x];
                const Aff ex = {exv.x, exv.y};
                float *o = out_ptr + (long long)j * a.out_ds;
                float2 old[4];
                if (addm && (!TAIL || nvalid == ITEMS)) ldg8(o, old);  // in flight while this warp waits for its entering state
                mbar_wait(&bar_in[j], 0);
                if (threadIdx.x == 0) VMASR_TL(a, j == 0 ? 5 : 6);
                const float h_in = fmaf(ex.p, s_in[j * WPR + warp], ex.q);
                const float *sy = s_stage + (size_t)j * 2 * SEG + slot;
                float2 Y0[4], Y1[4], y[4];
                lds8_priv(sy, sel, Y0);
                lds8_priv(sy + SEG, sel, Y1);
#pragma unroll
                for (int k = 0; k < 4; ++k) y[k] = fma2(Y1[k], f2(h_in), Y0[k]);
                if (!TAIL || nvalid == ITEMS) {
                    if (accum) {
                        red8(o, y);
                    } else {
                        if (addm) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) y[k] = add2(old[k], y[k]);
                        }
                        stg8(o, y);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (accum || addm) {
                            if (2 * k < nvalid) atomicAdd(o + 2 * k, y[k].x);
                            if (2 * k + 1 < nvalid) atomicAdd(o + 2 * k + 1, y[k].y);
                        } else {
                            if (2 * k < nvalid) o[2 * k] = y[k].x;
                            if (2 * k + 1 < nvalid) o[2 * k + 1] = y[k].y;
                        }
                    }
                }
            }
        }
        if (threadIdx.x == 0) VMASR_TL(a, 7);
    }

}

template <bool SP, bool F1>
__global__ void __launch_bounds__(kPipeThreads, 3) scan_fwd_pipe_kernel(const __grid_constant__ GroupArgs ga) {
    extern __shared__ __align__(1024) unsigned char smem_fwd_pipe[];  // swizzled tiles need 512-byte aligned slots
    pdl_launch_dependents();  // the next kernel on the stream may be scheduled while this one drains ...
    int tile;
    const int prob = group_problem(ga, tile);
    const ScanArgs &a = ga.a[prob];
    const TileMaps &tm = ga.tm[prob];
    if (a.pdl_mode & 1) pdl_wait();
    const int ztile = tile;   // ... and this one touches global memory only after its predecessor has completed: the wait is in the body,
                              // behind the CTA's own set-up (index arithmetic, mbarrier initialisation), which needs nothing from memory
    int half = -1;
    if (tile >= a.split_from) {
        half = (tile - a.split_from) & 1;
        tile = a.split_from + ((tile - a.split_from) >> 1);
    }
    const int chunk = tile / a.n_rowgroups;  // chunk-major in TIME order: a tile only waits on tiles dispatched before it
    const int rg = tile - chunk * a.n_rowgroups;
    const int mchunk = a.rev ? a.n_chunks - 1 - chunk : chunk;
    const bool tail = (mchunk + 1) * 2048 > a.seqlen;
    if (a.rev) {
        if (tail) scan_fwd_pipe_body<true, SP, true, F1>(a, tm, smem_fwd_pipe, chunk, rg, half, ztile);
        else scan_fwd_pipe_body<false, SP, true, F1>(a, tm, smem_fwd_pipe, chunk, rg, half, ztile);
    } else {
        if (tail) scan_fwd_pipe_body<true, SP, false, F1>(a, tm, smem_fwd_pipe, chunk, rg, half, ztile);
        else scan_fwd_pipe_body<false, SP, false, F1>(a, tm, smem_fwd_pipe, chunk, rg, half, ztile);
    }
}

template <bool SP, bool F1>
static int launch_fwd_pipe(const GroupArgs &ga, int grid, cudaStream_t stream) {
    const size_t smem = 2048 + sizeof(float) * ((size_t)kPipeStages * 2 * 2048) + sizeof(float2) * kPipeStages * 256;
    static PerDeviceOnce configured;  // the attribute is per function and per device
    if (!configured()) {
        if (int rc = check_cuda(cudaFuncSetAttribute(scan_fwd_pipe_kernel<SP, F1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                                "scan_fwd_pipe smem attribute"))
            return rc;
        configured() = true;
    }
    return launch_pdl(scan_fwd_pipe_kernel<SP, F1>, grid, kPipeThreads, smem, stream, "scan_fwd_pipe launch", ga);
}

// every problem: n_chunks > 1, at most kPipeStages channels per tile, same softplus flag (scan_host.cu groups them so)
int scan_fwd_pipe_dispatch(const GroupArgs &ga, int grid, cudaStream_t stream) {
    for (int i = 0; i < ga.n; ++i)
        if (ga.a[i].chan_per_tile > kPipeStages) return fail("scan_fwd_pipe: %d channels per tile (max %d)", ga.a[i].chan_per_tile, kPipeStages);
    for (int i = 0; i < ga.n; ++i)
        if ((ga.a[i].dt_rank > 0) != (ga.a[0].dt_rank > 0) || ga.a[i].dt_rank > 1) return fail("scan_fwd_pipe: mixed or unsupported dt_rank in one launch");
    if (ga.a[0].dt_rank > 0) return ga.a[0].softplus ? launch_fwd_pipe<true, true>(ga, grid, stream) : launch_fwd_pipe<false, true>(ga, grid, stream);
    return ga.a[0].softplus ? launch_fwd_pipe<true, false>(ga, grid, stream) : launch_fwd_pipe<false, false>(ga, grid, stream);
}

}  // namespace vmasr
