// Selective-scan backward, fast path for sequences of more than one chunk (fp32 IO, d_state 1, 16-byte aligned rows).
// Replaces selective_scan_bwd_kernel (kernels/selective_scan/csrc/selective_scan/cus/selective_scan_bwd_kernel.cuh:66-306).
// Maths as in scan_bwd_tma.cu (g = adjoint of the state, running right to left):
//   g_l = C_l dout_l + a_{l+1} g_{l+1};   du_l = D dout_l + g_l dt_l B_l;   ddt_l = g_l (B_l u_l + A a_l h_{l-1})
//   dA = sum g_l dt_l a_l h_{l-1};  dB_l += g_l dt_l u_l;  dC_l += dout_l h_l;  dD = sum dout u;
//   ddelta_l = ddt_l sigmoid(delta_l + bias);  ddelta_bias = sum ddelta_l;          a_l h_{l-1} = h_l - dt_l B_l u_l
//
// Same structure as scan_fwd_pipe.cu -- the tile (<= 4 channels x 2048 positions x {u, delta, dout}, 96 KB) is resident in
// shared memory, no CTA-wide barrier around the carry exchange; 8 compute warps (8 positions per thread) + 1 exchange warp:
//   compute warps:
//     P1(j), j = 0..n-1   u / delta / dout of channel j arrive by TMA (all issued at kernel start); softplus, decay, the
//                thread's aggregates of BOTH recurrences (forward state h left to right, adjoint g right to left), two
//                interleaved warp scans, warp aggregates to shared memory, ARRIVE on mbarrier tot[j].  The only
//                per-position value kept is dt (log2 domain), parked in delta's slot; the thread's warp-exclusive
//                prefixes of both scans (16 bytes per channel) wait in a thread-private shared-memory slot.  (They used
//                to be four registers per channel selected by `if (i == j)` chains: ptxas kept all sixteen in local
//                memory and moved every one of them in and out in every iteration of both loops.)  B and C live in
//                registers for the whole tile; their landing slots are the last stage's.
//     P2(j), j = 0..n-1   waits on mbarrier in[j]; re-reads u, dt, dout, recomputes the decay (one ex2) and the sigmoid
//                (1 - 2^-dt, one ex2) instead of carrying them in registers across the exchange, walks both recurrences,
//                forms the gradients; 128-bit stores of du / ddelta; dB / dC accumulate in registers over the channels.
//   exchange warp:
//     sweep A    per channel: waits on tot[j], combines the 8 warp aggregates, publishes the chunk's adjoint aggregate,
//                issues the look-back loads over the chunks to the right (pipe.cuh) and the load of the forward state
//                entering the chunk (from the `x` tensor the forward saved);
//     sweep B    per channel: reduces the look-back, writes the two states entering every warp, arrives on in[j].
//     Finally it adds the per-channel sums (dA, dD, ddelta_bias) the compute warps left in shared memory to global memory.
//
// F1 (delta on the fly, dt_rank 1; include/vmasr_b200.h): no delta rows are read and no ddelta rows are written.  Tiles take
// at most 3 channels, so the fourth stage is free: the group's dt row arrives there with B (one copy per tile), P1 forms
// delta = w_j * row in registers, and P2 folds ddelta into the gradients of the two factors -- d_dt_rows += w_j * ddelta
// accumulated over the tile's channels in a thread-private shared-memory slot (one 128-bit red per 4 positions per tile, like
// dB / dC), d_dt_weight[j] += sum ddelta * row as a fourth per-channel sum.
#include <cstdlib>

#include "fast.cuh"

namespace vmasr {

constexpr int kBPipeStages = 4;        // most channels per tile = resident stages (u + delta + dout, 24 KB per channel); 3 and 4 are built
constexpr int kBPipeThreads = 288;     // 8 compute warps + the exchange warp

// `chunk` is the chunk's index in TIME order (the forward scan's order); with REV it sits at the mirrored place in memory.
template <bool TAIL, bool SP, int STAGES, bool REV, bool F1>
__device__ __forceinline__ void scan_bwd_pipe_body(const ScanArgs &a, const TileMaps &tm, unsigned char *smem, const int chunk, const int rg) {
    constexpr int NC = 256, ITEMS = 8, WPR = 8, SEG = NC * ITEMS;

    // shared memory carve-up (header 1024 bytes; 1024 + STAGES * (24576 + 4096) bytes in all: with 4 stages exactly the
    // 113 KB a CTA may take when two share an SM)
    unsigned long long *bar_full = reinterpret_cast<unsigned long long *>(smem);  // [STAGES] TMA completion
    unsigned long long *bar_bc = bar_full + STAGES;                               // B / C segment
    unsigned long long *bar_tot = bar_bc + 1;                                     // [STAGES] 8 arrivals
    unsigned long long *bar_in = bar_tot + STAGES;                                // [STAGES] 8 arrivals (exchange-warp lanes)
    unsigned long long *bar_done = bar_in + STAGES;                                  // 8 arrivals: a compute warp has finished its sweeps
    float4 *s_tot = reinterpret_cast<float4 *>(smem + 128);                       // [STAGES][8] warp totals {p, q fwd, q adjoint, -}
    float2 *s_in = reinterpret_cast<float2 *>(smem + 640);                        // [STAGES][8] {h, g} entering each warp
    float *s_red = reinterpret_cast<float *>(smem + 896);                         // [STAGES][4] channel sums dA, dD, dbias
    float *s_par = reinterpret_cast<float *>(smem + 960);                         // [4][4]  A, D, bias * log2 e, (F1) dt weight
    float *s_stage = reinterpret_cast<float *>(smem + 1024);                      // [STAGES][3][SEG]  u, delta -> dt, dout
    float4 *s_exc = reinterpret_cast<float4 *>(s_stage + (size_t)STAGES * 3 * SEG);  // [STAGES][NC] thread-private: warp-exclusive prefixes of both scans
    float *s_b = s_stage + (size_t)(STAGES - 1) * 3 * SEG;                        // B and C land in the last stage and move to registers
    float *s_row = s_b + SEG;                                                     // F1: the dt row (the last stage stays free)
    float *s_drow = s_b + 2 * SEG;                                                // F1: its gradient, summed over the tile's channels
    float *s_c = F1 ? s_drow : s_b + SEG;                                         // (F1: the gradient slot is zeroed after C has been read)
    static_assert(!F1 || STAGES == 4, "delta on the fly keeps the dt row in the fourth stage");

    const int ctile = rg % a.n_ctiles;
    const int bg = rg / a.n_ctiles;
    const int g = bg % a.ngroups;
    const int b = bg / a.ngroups;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const bool exchange = warp == WPR;
    const int L = a.seqlen;
    const int seg0 = (REV ? a.n_chunks - 1 - chunk : chunk) * SEG;  // first MEMORY position of the tile
    const int line0 = seg0 / kTileLine;                             // ... and its first line in the tensor maps
    constexpr unsigned seg_bytes = SEG * 4u;                        // a box always counts in full (lines past the end arrive as zeros)

    const int c_begin = ctile * a.chan_per_tile;
    const int n_iter = min(a.chan_per_group, c_begin + a.chan_per_tile) - c_begin;  // <= STAGES
    const int d0 = g * a.chan_per_group + c_begin;

    auto issue_stage = [&](int it) {  // lane 0 of the exchange warp only
        float *dst = s_stage + (size_t)it * 3 * SEG;
        mbar_expect_tx(&bar_full[it], (F1 ? 2u : 3u) * seg_bytes);
        tensor_load(dst, &tm.u, line0, d0 + it, b, &bar_full[it]);
        if (!F1) tensor_load(dst + SEG, &tm.delta, line0, d0 + it, b, &bar_full[it]);
        tensor_load(dst + 2 * SEG, &tm.dout, line0, d0 + it, b, &bar_full[it]);
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
        mbar_init(bar_done, WPR);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar_bc, (F1 ? 3u : 2u) * seg_bytes);
        tensor_load(s_b, &tm.B, line0, g, b, bar_bc);
        tensor_load(s_c, &tm.C, line0, g, b, bar_bc);
        if (F1) tensor_load(s_row, &tm.delta, line0, g, b, bar_bc);  // dt_rank 1: row index = group
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s)
            if (s < n_iter) issue_stage(s);
    }
    if (threadIdx.x < STAGES * 4) s_red[threadIdx.x] = 0.0f;
    if (threadIdx.x >= 32 && threadIdx.x < 32 + (F1 ? 4 : 3) * n_iter) {
        const int i = threadIdx.x - 32;
        const int which = i / n_iter, cc = i - which * n_iter;
        const int d = d0 + cc;
        float v;
        if (which == 0) v = __ldg(a.A + d * a.A_ds);
        else if (which == 1) v = a.D ? __ldg(a.D + d) : 0.0f;
        else if (which == 2) v = (a.delta_bias ? __ldg(a.delta_bias + d) : 0.0f) * kLog2e;
        else v = __ldg(a.dt_w + d * a.dtw_ds);
        s_par[which * 4 + cc] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) VMASR_TL(a, 1);

    const long long seq0 = (long long)b * a.dim + d0;
    if (exchange) {
        // ================= exchange warp =================
        __syncthreads();  // the compute warps hold B and C in registers: the last stage is free for data now
        if (lane == 0 && STAGES - 1 < n_iter) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue_stage(STAGES - 1);
        }
        const unsigned epoch = launch_epoch(a, lane);  // (also recycles the carry workspace for the next launch: pipe.cuh)
        const int n_groups16 = (a.n_chunks + 15) >> 4;
        const int jrev = a.n_chunks - 1 - chunk;  // position of this chunk in the adjoint's scan order
        // per channel: publish the chunk's adjoint aggregate as soon as it exists and start its look-back; finish the
        // look-back of the channel before (its loads have been in flight for one P1 of the compute warps)
        auto finish = [&](int j, CarryLook &l, const Aff &cf, const Aff &cr, float hc) {
            CarryEntry *l2_row = a.ws_entries2 + (seq0 + j) * n_groups16;
            const Aff before_f = shift_up1(cf, lane);
            const Aff before_r = shift_down1(cr, lane);
            const Aff total_r = {__shfl_sync(0xffffffffu, cr.p, 0), __shfl_sync(0xffffffffu, cr.q, 0)};
            bool ok;
            Aff grp = {1.0f, 0.0f};
            Aff acc = look_reduce(l, epoch, lane, ok, grp);
            if (!a.debug_nowait) acc = look_finish(l, acc, ok, l2_row, jrev, epoch, lane, grp);
            if (lane == 0 && (jrev & 15) == 15) {
                const Aff g16 = compose(grp, total_r);
                publish_entry(l2_row + (jrev >> 4), epoch, g16.p, g16.q);
            }
            if (lane < WPR) {  // every writer releases its own store
                s_in[j * WPR + lane] = make_float2(fmaf(before_f.p, hc, before_f.q), fmaf(before_r.p, acc.q, before_r.q));
                mbar_arrive(&bar_in[j]);
            }
        };
        CarryLook p_look;
        p_look.ptr = nullptr;
        p_look.e = make_uint4(0u, 0u, 0u, 0u);
        Aff p_cf = {1.0f, 0.0f}, p_cr = {1.0f, 0.0f};
        float p_h = 0.0f;
#pragma unroll 1
        for (int j = 0; j < n_iter; ++j) {
            const long long seq = seq0 + j;
            CarryEntry *l1_row = a.ws_entries + seq * a.n_chunks;
            const float h_chunk = (chunk > 0) ? __ldg(a.x + (seq * a.n_chunks + (chunk - 1)) * 2 + 1) : 0.0f;
            mbar_wait(&bar_tot[j], 0);
            if (lane == 0) VMASR_TL(a, j == 0 ? 8 : 9);
            float4 t = make_float4(1.0f, 0.0f, 0.0f, 0.0f);
            if (lane < WPR) t = s_tot[j * WPR + lane];
            Aff cum_f = {t.x, t.y}, cum_r = {t.x, t.z};
#pragma unroll
            for (int off = 1; off < WPR; off <<= 1) {
                scan_step_down(cum_r.p, cum_r.q, off);
                scan_step_up(cum_f.p, cum_f.q, off);
            }
            if (lane == 0) publish_entry(l1_row + jrev, epoch, cum_r.p, cum_r.q);
            const CarryLook look = look_issue(l1_row, a.ws_entries2 + seq * n_groups16, jrev, lane);
            if (j >= 1) finish(j - 1, p_look, p_cf, p_cr, p_h);
            p_look = look;
            p_cf = cum_f;
            p_cr = cum_r;
            p_h = h_chunk;
        }
        finish(n_iter - 1, p_look, p_cf, p_cr, p_h);
        if (lane == 0) VMASR_TL(a, 10);
        mbar_wait(bar_done, 0);  // every compute warp is done: the channel sums are complete
        {
            const int c = lane >> 2, which = lane & 3;
            if (c < n_iter && which < (F1 ? 4 : 3)) {
                const float v = s_red[c * 4 + which];
                const int d = d0 + c;
                if (which == 0) atomicAdd(a.dA + d * a.A_ds, v);
                else if (which == 1) { if (a.dD) atomicAdd(a.dD + d, v); }
                else if (which == 2) { if (a.ddelta_bias) atomicAdd(a.ddelta_bias + d, v); }
                else atomicAdd(a.d_dt_w + d * a.dtw_ds, v);
            }
        }
    } else {
        // ================= compute warps =================
        const int tseg = REV ? NC - 1 - (int)threadIdx.x : (int)threadIdx.x;  // this thread's 8-position segment of the tile (memory order)
        const int pos = seg0 + tseg * ITEMS;
        const int sel = swz_half(tseg), slot = swz_slot(tseg) * ITEMS;  // where the 64-byte swizzle puts this thread's 32 bytes (scan.cuh)
        const bool accum = a.accum == 1, addm = a.accum == 2;  // red.add / load-add-store (scan.cuh)
        int nvalid = ITEMS;
        if (TAIL) nvalid = max(0, min(ITEMS, L - pos));
        // CTA-uniform parts of the output addresses (the thread's own part, tseg * ITEMS, is added at the stores)
        const long long du_tile = b * a.du_bs + (long long)d0 * a.du_ds + seg0;
        const long long dd_tile = F1 ? 0 : b * a.ddelta_bs + (long long)d0 * a.ddelta_ds + seg0;

        float2 Bv[4], Cv[4];  // this thread's B and C values: registers for the whole tile
        mbar_wait(bar_bc, 0);
        if (threadIdx.x == 0) VMASR_TL(a, 2);
        lds8_priv(s_b + slot, sel, Bv);
        lds8_priv(s_c + slot, sel, Cv);
        if (TAIL) {  // positions past the end contribute nothing and stay finite
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (2 * k >= nvalid) { Bv[k].x = 0.0f; Cv[k].x = 0.0f; }
                if (2 * k + 1 >= nvalid) { Bv[k].y = 0.0f; Cv[k].y = 0.0f; }
            }
        }
        __syncthreads();  // B and C are in registers: the last stage is free (the exchange warp refills it)
        if (F1) {
            const float2 zero[4] = {f2(0.0f), f2(0.0f), f2(0.0f), f2(0.0f)};
            sts8_priv(s_drow + slot, sel, zero);  // thread-private accumulator of d_dt_rows
        }
        float4 *my_exc = s_exc + threadIdx.x;  // + j * NC: the prefixes of channel j wait here between P1(j) and P2(j)
#pragma unroll 1
        for (int j = 0; j < n_iter; ++j) {
            {
                // ---- P1(j) ----
                const float Av = s_par[j];
                const float bias2 = s_par[2 * 4 + j];
                float *su = s_stage + (size_t)j * 3 * SEG + slot;
                mbar_wait(&bar_full[j], 0);
                if (threadIdx.x == 0 && j == 0) VMASR_TL(a, 3);
                float2 uv[4], dl[4], dy[4], dts[4];
                lds8_priv(su, sel, uv);
                lds8_priv(F1 ? s_row + slot : su + SEG, sel, dl);  // delta, or (F1) the dt row
                lds8_priv(su + 2 * SEG, sel, dy);
                const float wdt = F1 ? s_par[3 * 4 + j] * kLog2e : kLog2e;
                if (TAIL) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (2 * k >= nvalid) { uv[k].x = 0.0f; dl[k].x = 0.0f; dy[k].x = 0.0f; }
                        if (2 * k + 1 >= nvalid) { uv[k].y = 0.0f; dl[k].y = 0.0f; dy[k].y = 0.0f; }
                    }
                    sts8_priv(su, sel, uv);  // park the cleaned values for P2
                    sts8_priv(su + 2 * SEG, sel, dy);
                }
                float p = 1.0f, q = 0.0f, qr = 0.0f;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const int k = pair_at<REV>(kk);  // pairs in time order
                    float2 dt2 = fma2(dl[k], f2(wdt), f2(bias2));
                    if (SP) {
                        float2 e, sp;
                        dt2 = softplus2_pair(dt2, e, sp);
                    }
                    dts[k] = dt2;
                    const float2 da = mul2(dt2, f2(Av));
                    const float2 av = make_float2(ex2_approx(da.x), ex2_approx(da.y));
                    const float2 bx = mul2(mul2(dt2, f2(kLn2)), mul2(Bv[k], uv[k]));
                    const float2 cdy = mul2(Cv[k], dy[k]);
                    // local aggregates of both recurrences in one walk in time order:
                    //   forward  s -> p s + q;   adjoint (entering from the future)  G -> p G + qr,  qr = sum_i (prod_{m<=i} a_m) C_i dout_i
                    walk_pair2<REV>(av, bx, cdy, p, q, qr);
                }
                sts8_priv(su + SEG, sel, dts);
                // two independent warp scans, interleaved level by level (each level is shuffle-latency bound)
                Aff inc_f = {p, q}, inc_r = {p, qr};
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    scan_step_up(inc_f.p, inc_f.q, off);
                    scan_step_down(inc_r.p, inc_r.q, off);
                }
                const Aff exf = shift_up1(inc_f, lane), exr = shift_down1(inc_r, lane);
                my_exc[j * NC] = make_float4(exf.p, exf.q, exr.p, exr.q);
                const float qr0 = __shfl_sync(0xffffffffu, inc_r.q, 0);
                if (lane == 31) s_tot[j * WPR + warp] = make_float4(inc_f.p, inc_f.q, qr0, 0.0f);
                __syncwarp();
                if (lane == 31) mbar_arrive(&bar_tot[j]);
            }
        }
        if (threadIdx.x == 0) VMASR_TL(a, 4);
        float2 dBacc[4], dCacc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            dBacc[k] = f2(0.0f);
            dCacc[k] = f2(0.0f);
        }
#pragma unroll 1
        for (int j = 0; j < n_iter; ++j) {
            {
                // ---- P2(j) ----
                const float4 exc = my_exc[j * NC];
                const Aff exf = {exc.x, exc.y}, exr = {exc.z, exc.w};
                const float Av = s_par[j];
                const float Dv = s_par[4 + j];
                mbar_wait(&bar_in[j], 0);
                if (threadIdx.x == 0) VMASR_TL(a, j == 0 ? 5 : 6);
                const float2 in = s_in[j * WPR + warp];
                const float *su = s_stage + (size_t)j * 3 * SEG + slot;
                float2 uv[4], dts[4], dy[4];
                lds8_priv(su, sel, uv);
                lds8_priv(su + SEG, sel, dts);
                lds8_priv(su + 2 * SEG, sel, dy);
                float2 av[4], bu[4], dtn[4], hs[4], gl[4];
                // forward states of this thread's positions
                {
                    float h = fmaf(exf.p, in.x, exf.q);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const int k = pair_at<REV>(kk);  // time order
                        const float2 da = mul2(dts[k], f2(Av));
                        av[k] = make_float2(ex2_approx(da.x), ex2_approx(da.y));
                        dtn[k] = mul2(dts[k], f2(kLn2));
                        bu[k] = mul2(Bv[k], uv[k]);
                        const float2 bx = mul2(dtn[k], bu[k]);
                        walk_state<REV>(av[k], bx, h, hs[k]);
                    }
                }
                // adjoint walk, against time order
                {
                    float G = fmaf(exr.p, in.y, exr.q);
#pragma unroll
                    for (int kk = 3; kk >= 0; --kk) {
                        const int k = pair_at<REV>(kk);
                        const float2 cdy = mul2(Cv[k], dy[k]);
                        walk_adjoint<REV>(av[k], cdy, G, gl[k]);
                    }
                }
                // gradients, position pairs
                float2 du[4], ddl[4];
                float2 sA = f2(0.0f), sD = f2(0.0f), sB = f2(0.0f), sW = f2(0.0f);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 carried = fma2(mul2(dtn[k], bu[k]), f2(-1.0f), hs[k]);  // a_l h_{l-1}
                    const float2 w = mul2(gl[k], dtn[k]);
                    du[k] = fma2(w, Bv[k], mul2(dy[k], f2(Dv)));
                    const float2 ddt = mul2(gl[k], fma2(carried, f2(Av), bu[k]));  // g (B u + A a h_prev)
                    if (SP) {
                        // sigmoid(delta + bias) = 1 - exp(-softplus) = 1 - 2^(-dt2); exactly 1 above the reference's threshold.
                        // Below softplus = 1/16 the subtraction would cancel (the Mamba-style dt range 1e-3 .. 1e-1 lives there):
                        // 1 - exp(-s) = s (1 - s/2 + s^2/6 - s^3/24), next term s^4/120 < 1.3e-7 relative.
                        const float2 sn = dtn[k];  // softplus, natural units
                        float2 ser = fma2(sn, f2(-1.0f / 24.0f), f2(1.0f / 6.0f));
                        ser = fma2(sn, ser, f2(-0.5f));
                        ser = fma2(sn, ser, f2(1.0f));
                        ser = mul2(sn, ser);
                        float2 sig = make_float2(1.0f - ex2_approx(-dts[k].x), 1.0f - ex2_approx(-dts[k].y));
                        if (sn.x < 0.0625f) sig.x = ser.x;
                        if (sn.y < 0.0625f) sig.y = ser.y;
                        // (above the threshold 2^(-dt2) < 2^-28 and the subtraction already returns exactly 1)
                        ddl[k] = mul2(ddt, sig);
                    } else {
                        ddl[k] = ddt;
                    }
                    sA = fma2(w, carried, sA);
                    dBacc[k] = fma2(w, uv[k], dBacc[k]);
                    dCacc[k] = fma2(dy[k], hs[k], dCacc[k]);
                    sD = fma2(dy[k], uv[k], sD);
                    sB = add2(sB, ddl[k]);
                }
                if (F1) {  // ddelta = w * d(row) and row * d(w): neither is stored per channel
                    const float w = s_par[3 * 4 + j];
                    float2 rowv[4], dr[4];
                    lds8_priv(s_row + slot, sel, rowv);
                    lds8_priv(s_drow + slot, sel, dr);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        sW = fma2(ddl[k], rowv[k], sW);
                        dr[k] = fma2(ddl[k], f2(w), dr[k]);
                    }
                    sts8_priv(s_drow + slot, sel, dr);
                }
                {
                    float *o_du = reinterpret_cast<float *>(a.du) + (du_tile + (long long)j * a.du_ds) + tseg * ITEMS;
                    float *o_dd = F1 ? nullptr : reinterpret_cast<float *>(a.ddelta) + (dd_tile + (long long)j * a.ddelta_ds) + tseg * ITEMS;
                    if (!TAIL || nvalid == ITEMS) {
                        if (accum) {
                            red8(o_du, du);
                        } else {
                            if (addm) {  // (the load sits behind the gradient arithmetic above; registers are too scarce here to issue it earlier)
                                float2 old[4];
                                ldg8(o_du, old);
#pragma unroll
                                for (int k = 0; k < 4; ++k) du[k] = add2(old[k], du[k]);
                            }
                            stg8(o_du, du);
                        }
                        if (!F1) stg8(o_dd, ddl);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (2 * k < nvalid) {
                                if (accum || addm) atomicAdd(o_du + 2 * k, du[k].x); else o_du[2 * k] = du[k].x;
                                if (!F1) o_dd[2 * k] = ddl[k].x;
                            }
                            if (2 * k + 1 < nvalid) {
                                if (accum || addm) atomicAdd(o_du + 2 * k + 1, du[k].y); else o_du[2 * k + 1] = du[k].y;
                                if (!F1) o_dd[2 * k + 1] = ddl[k].y;
                            }
                        }
                    }
                }
                // per-channel sums: dA, dD, ddelta_bias, (F1) d_dt_weight (lanes 0, 8, 16, 24 hold them after the reduction)
                const float r = warp_sum4(sA.x + sA.y, sD.x + sD.y, sB.x + sB.y, sW.x + sW.y, lane);
                if ((lane & 7) == 0 && lane < (F1 ? 32 : 24)) atomicAdd(s_red + j * 4 + (lane >> 3), r);
            }
        }
        if (threadIdx.x == 0) VMASR_TL(a, 7);
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_done);  // this warp's shared-memory atomics are done (release)

        // dB / dC of this tile's positions, summed over the tile's channels
        float *dBg = a.dB + b * a.dB_bs + (long long)g * L;
        float *dCg = a.dC + b * a.dC_bs + (long long)g * L;
        if (F1) {
            float2 dr[4];
            lds8_priv(s_drow + slot, sel, dr);
            float *dRg = a.d_dt_rows + b * a.dtr_bs + (long long)g * a.dtr_rs + pos;
            if (!TAIL || nvalid == ITEMS) {
                red8(dRg, dr);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (2 * k < nvalid) atomicAdd(dRg + 2 * k, dr[k].x);
                    if (2 * k + 1 < nvalid) atomicAdd(dRg + 2 * k + 1, dr[k].y);
                }
            }
        }
        if (!TAIL || nvalid == ITEMS) {
            red8(dBg + pos, dBacc);
            red8(dCg + pos, dCacc);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (2 * k < nvalid) { atomicAdd(dBg + pos + 2 * k, dBacc[k].x); atomicAdd(dCg + pos + 2 * k, dCacc[k].x); }
                if (2 * k + 1 < nvalid) { atomicAdd(dBg + pos + 2 * k + 1, dBacc[k].y); atomicAdd(dCg + pos + 2 * k + 1, dCacc[k].y); }
            }
        }
    }
}

template <bool SP, int STAGES, bool F1>
__global__ void __launch_bounds__(kBPipeThreads, 2) scan_bwd_pipe_kernel(const __grid_constant__ GroupArgs ga) {
    extern __shared__ __align__(1024) unsigned char smem_bwd_pipe[];  // swizzled tiles need 512-byte aligned slots
    pdl_launch_dependents();  // the next kernel on the stream may be scheduled while this one drains ...
    pdl_wait();               // ... and this one touches global memory only after its predecessor has completed
    int tile;
    const int prob = group_problem(ga, tile);
    const ScanArgs &a = ga.a[prob];
    const TileMaps &tm = ga.tm[prob];
    // adjoint: late chunks first; block order = the adjoint's scan order, so a tile only waits on tiles dispatched before it
    const int chunk = a.n_chunks - 1 - tile / a.n_rowgroups;
    const int rg = tile % a.n_rowgroups;
    const int mchunk = a.rev ? a.n_chunks - 1 - chunk : chunk;
    const bool tail = (mchunk + 1) * 2048 > a.seqlen;
    if (a.rev) {
        if (tail) scan_bwd_pipe_body<true, SP, STAGES, true, F1>(a, tm, smem_bwd_pipe, chunk, rg);
        else scan_bwd_pipe_body<false, SP, STAGES, true, F1>(a, tm, smem_bwd_pipe, chunk, rg);
    } else {
        if (tail) scan_bwd_pipe_body<true, SP, STAGES, false, F1>(a, tm, smem_bwd_pipe, chunk, rg);
        else scan_bwd_pipe_body<false, SP, STAGES, false, F1>(a, tm, smem_bwd_pipe, chunk, rg);
    }
}

template <bool SP, int STAGES, bool F1>
static int launch_bwd_pipe(const GroupArgs &ga, int grid, cudaStream_t stream) {
    const size_t smem = 1024 + (size_t)STAGES * (3 * 2048 * sizeof(float) + 256 * sizeof(float4));
    static PerDeviceOnce configured;  // the attribute is per function and per device
    if (!configured()) {
        if (int rc = check_cuda(cudaFuncSetAttribute(scan_bwd_pipe_kernel<SP, STAGES, F1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                                "scan_bwd_pipe smem attribute"))
            return rc;
        configured() = true;
    }
    return launch_pdl(scan_bwd_pipe_kernel<SP, STAGES, F1>, grid, kBPipeThreads, smem, stream, "scan_bwd_pipe launch", ga);
}

// every problem: n_chunks > 1, at most kBPipeStages channels per tile, same softplus flag and the same side of the
// 3-channel boundary (scan_host.cu groups them so)
int scan_bwd_pipe_dispatch(const GroupArgs &ga, int grid, cudaStream_t stream) {
    for (int i = 0; i < ga.n; ++i)
        if (ga.a[i].chan_per_tile > kBPipeStages) return fail("scan_bwd_pipe: %d channels per tile (max %d)", ga.a[i].chan_per_tile, kBPipeStages);
    for (int i = 0; i < ga.n; ++i)
        if ((ga.a[i].dt_rank > 0) != (ga.a[0].dt_rank > 0) || ga.a[i].dt_rank > 1 || (ga.a[i].dt_rank > 0 && ga.a[i].chan_per_tile > 3))
            return fail("scan_bwd_pipe: mixed or unsupported dt_rank in one launch");
    if (ga.a[0].dt_rank > 0) return ga.a[0].softplus ? launch_bwd_pipe<true, 4, true>(ga, grid, stream) : launch_bwd_pipe<false, 4, true>(ga, grid, stream);
    if (ga.a[0].chan_per_tile <= 3)  // smaller tile (groups of 2 channels): 82 KB of shared memory, fewer live registers (measured +11 %)
        return ga.a[0].softplus ? launch_bwd_pipe<true, 3, false>(ga, grid, stream) : launch_bwd_pipe<false, 3, false>(ga, grid, stream);
    return ga.a[0].softplus ? launch_bwd_pipe<true, 4, false>(ga, grid, stream) : launch_bwd_pipe<false, 4, false>(ga, grid, stream);
}

}  // namespace vmasr
