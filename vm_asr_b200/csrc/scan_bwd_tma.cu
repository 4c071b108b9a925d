// Selective-scan backward, fast path for sm_100a: fp32 IO, d_state 1, 16-byte aligned rows.
// Replaces selective_scan_bwd_kernel (kernels/selective_scan/csrc/selective_scan/cus/selective_scan_bwd_kernel.cuh:66-306)
// for the shapes VM-ASR runs.  Maths as in scan_bwd.cu (g = adjoint of the state, running right to left):
//   g_l = C_l dout_l + a_{l+1} g_{l+1};   du_l = D dout_l + g_l dt_l B_l;   ddt_l = g_l (B_l u_l + A a_l h_{l-1})
//   dA = sum g_l dt_l a_l h_{l-1};  dB_l += g_l dt_l u_l;  dC_l += dout_l h_l;  dD = sum dout u;
//   ddelta_l = ddt_l sigmoid(delta_l + bias);  ddelta_bias = sum ddelta_l;          a_l h_{l-1} = h_l - dt_l B_l u_l
// Structure (same as scan_fwd_tma.cu): u / delta / dout row segments arrive by TMA bulk copies into a per-row-segment
// shared-memory ring, element-wise arithmetic runs on position pairs with packed fp32x2 instructions, the row segments of
// a CTA are independent pipelines, and the first warp of a row segment does the cross-warp combination of BOTH
// recurrences plus the cross-chunk look-back of the adjoint (the forward state entering the chunk comes from the `x`
// tensor the forward saved) and hands every warp its two incoming states through shared memory.
// dB / dC are summed over the tile's channels in registers (position pairs) and leave as one 128-bit reduction per 4
// positions per tile; the three per-channel sums are reduced with six shuffles and one shared-memory atomic per warp.
#include "fast.cuh"

namespace vmasr {

constexpr int kBwdStages = 3;
constexpr int kBwdMaxTileChannels = 64;

// REV (time runs against memory order) is supported for single-chunk sequences, which is all the host sends here.
template <int TPR, bool TAIL, bool SP, bool REV>
__device__ __forceinline__ void scan_bwd_tma_body(const ScanArgs &a, const TileMaps &tm, unsigned char *smem, const int chunk, const int rg,
                                                  const unsigned epoch) {
    constexpr int NT = 256, ITEMS = 8, STAGES = kBwdStages;
    constexpr int ROWS = NT / TPR;
    constexpr int WPR = TPR / 32;
    constexpr int SEG = TPR * ITEMS;

    // shared memory carve-up (header 2048 bytes; [384, 392) is the tile ticket of the kernel wrapper)
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem);  // [STAGES][ROWS] stage barriers, then B/C
    float2 *s_in = reinterpret_cast<float2 *>(smem + 320);                    // [8] {h, g} entering each warp
    float4 *s_tot = reinterpret_cast<float4 *>(smem + 512);                   // [2][8] warp totals {p, q fwd, q adjoint, -}
    float *s_red = reinterpret_cast<float *>(smem + 768);                     // [2][8 rows][4] per-channel sums dA, dD, dbias
    float *s_par = reinterpret_cast<float *>(smem + 1024);                    // [3][kBwdMaxTileChannels]
    float *s_bc = reinterpret_cast<float *>(smem + 2048);                     // [2][SEG]   B, C
    float *s_stage = s_bc + 2 * SEG;                                          // [STAGES][ROWS][3][SEG]  u, delta, dout
    unsigned long long *bar_bc = bars + STAGES * ROWS;

    const int ctile = rg % a.n_ctiles;
    const int bg = rg / a.n_ctiles;
    const int g = bg % a.ngroups;
    const int b = bg / a.ngroups;

    const int row = threadIdx.x / TPR;
    const int t_in_row = threadIdx.x - row * TPR;
    const int warp_in_row = t_in_row >> 5;
    const int warp_slot = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tseg = REV ? TPR - 1 - t_in_row : t_in_row;  // this thread's 8-position segment of the row segment (memory order)
    const int sel = swz_half(tseg), slot = swz_slot(tseg) * ITEMS;  // where the 64-byte swizzle puts this thread's 32 bytes (scan.cuh)
    const int L = a.seqlen;
    const int seg0 = chunk * SEG;
    const int pos = seg0 + tseg * ITEMS;
    const bool accum = a.accum == 1, addm = a.accum == 2;  // red.add / load-add-store (scan.cuh)
    const int line0 = seg0 / kTileLine;        // first line of the tile in the tensor maps
    constexpr unsigned seg_bytes = SEG * 4u;   // a box always counts in full (lines past the end arrive as zeros)
    int nvalid = ITEMS;
    if (TAIL) nvalid = max(0, min(ITEMS, L - pos));

    const int c_begin = ctile * a.chan_per_tile;
    const int n_chan = min(a.chan_per_group, c_begin + a.chan_per_tile) - c_begin;
    const int n_iter = (n_chan - row + ROWS - 1) / ROWS;  // iterations of THIS row segment (may be 0)
    const int d0 = g * a.chan_per_group + c_begin;

    float *du_ptr = reinterpret_cast<float *>(a.du) + b * a.du_bs + (long long)(d0 + row) * a.du_ds + pos;
    float *dd_ptr = reinterpret_cast<float *>(a.ddelta) + b * a.ddelta_bs + (long long)(d0 + row) * a.ddelta_ds + pos;
    const long long du_step = (long long)ROWS * a.du_ds, dd_step = (long long)ROWS * a.ddelta_ds;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < STAGES * ROWS + 1; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_wait();  // every thread, in front of its first access to global memory
    if (threadIdx.x < 64) s_red[threadIdx.x] = 0.0f;
    // the parameter loads start first, but nothing waits for them until the bulk copies below are on their way
    float par_v = 0.0f;
    if ((int)threadIdx.x < 3 * n_chan) {  // 3 * n_chan <= 192 < NT
        const int which = threadIdx.x / n_chan, cc = threadIdx.x - which * n_chan;
        const int d = d0 + cc;
        if (which == 0) par_v = __ldg(a.A + d * a.A_ds);
        else if (which == 1) par_v = a.D ? __ldg(a.D + d) : 0.0f;
        else par_v = (a.delta_bias ? __ldg(a.delta_bias + d) : 0.0f) * kLog2e;
    }
    __syncthreads();

    float *my_stage = s_stage + (size_t)row * 3 * SEG;  // + stage * ROWS * 3 * SEG
    unsigned long long *my_bars = bars + row;           // + stage * ROWS
    auto issue_stage = [&](int it) {
        const int s = it % STAGES;
        unsigned long long *bar = my_bars + s * ROWS;
        float *dst = my_stage + (size_t)s * ROWS * 3 * SEG;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the stage was written through the generic proxy (sigmoid)
        mbar_expect_tx(bar, 3u * seg_bytes);
        tensor_load(dst, &tm.u, line0, d0 + row + it * ROWS, b, bar);
        tensor_load(dst + SEG, &tm.delta, line0, d0 + row + it * ROWS, b, bar);
        tensor_load(dst + 2 * SEG, &tm.dout, line0, d0 + row + it * ROWS, b, bar);
    };
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar_bc, 2u * seg_bytes);
        tensor_load(s_bc, &tm.B, line0, g, b, bar_bc);
        tensor_load(s_bc + SEG, &tm.C, line0, g, b, bar_bc);
    }
    if (t_in_row == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s)
            if (s < n_iter) issue_stage(s);
    }
    if ((int)threadIdx.x < 3 * n_chan) {
        const int which = threadIdx.x / n_chan, cc = threadIdx.x - which * n_chan;
        s_par[which * kBwdMaxTileChannels + cc] = par_v;
    }
    __syncthreads();

    float2 Bv[4], dBacc[4], dCacc[4];
    mbar_wait(bar_bc, 0);
    lds8_priv(s_bc + slot, sel, Bv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        dBacc[j] = f2(0.0f);
        dCacc[j] = f2(0.0f);
        if (TAIL) {  // positions past the end contribute nothing and stay finite
            if (2 * j >= nvalid) Bv[j].x = 0.0f;
            if (2 * j + 1 >= nvalid) Bv[j].y = 0.0f;
        }
    }

    const int n_groups16 = (a.n_chunks + 15) >> 4;
#ifdef VMASR_TUNING
    const bool multi = a.n_chunks > 1;  // then ROWS == 1 (VMASR_SCAN_BWD=tma sends longer sequences here)
#else
    constexpr bool multi = false;  // the product library sends single-chunk sequences only (scan_host.cu::decide)
#endif
    const bool leader = warp_in_row == 0;
    const int jrev = a.n_chunks - 1 - chunk;  // position of this chunk in the adjoint's scan order

    auto flush_sums = [&](int it_done) {  // leader lanes 0..2, after a row barrier: sums of iteration it_done -> global
        float *red = s_red + ((it_done & 1) * 8 + row) * 4;
        const float v = red[lane];
        red[lane] = 0.0f;
        const int d = d0 + it_done * ROWS + row;
        if (lane == 0) atomicAdd(a.dA + d * a.A_ds, v);
        else if (lane == 1) { if (a.dD) atomicAdd(a.dD + d, v); }
        else { if (a.ddelta_bias) atomicAdd(a.ddelta_bias + d, v); }
    };

    for (int it = 0; it < n_iter; ++it) {
        const int s = it % STAGES;
        const int cc = it * ROWS + row;
        const float Av = s_par[cc];
        const float Dv = s_par[kBwdMaxTileChannels + cc];
        const float bias2 = s_par[2 * kBwdMaxTileChannels + cc];
        const long long seq = (long long)b * a.dim + d0 + cc;

        CarryLook look;
        CarryEntry *l1_row = nullptr, *l2_row = nullptr;
        float h_in = 0.0f;
        if (multi && leader) {
            l1_row = a.ws_entries + seq * a.n_chunks;
            l2_row = a.ws_entries2 + seq * n_groups16;
            look = look_issue(l1_row, l2_row, jrev, lane);
            if (chunk > 0) h_in = __ldg(a.x + (seq * a.n_chunks + (chunk - 1)) * 2 + 1);
        }
        mbar_wait(my_bars + s * ROWS, (unsigned)((it / STAGES) & 1));
        // The stage stays valid for the whole iteration (it is refilled one iteration later), so u and dout are read
        // again where they are needed instead of being held in registers, and sigmoid is parked in delta's slot.
        float *su = my_stage + (size_t)s * ROWS * 3 * SEG + slot;
        float2 dtn[4], av[4], bx[4], cdy[4];
        {
            float2 uv[4], dl[4], dy[4], Cv[4], sig[4];
            lds8_priv(su, sel, uv);
            lds8_priv(su + SEG, sel, dl);
            lds8_priv(su + 2 * SEG, sel, dy);
            lds8_priv(s_bc + SEG + slot, sel, Cv);
            if (TAIL) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (2 * j >= nvalid) { uv[j].x = 0.0f; dl[j].x = 0.0f; dy[j].x = 0.0f; Cv[j].x = 0.0f; }
                    if (2 * j + 1 >= nvalid) { uv[j].y = 0.0f; dl[j].y = 0.0f; dy[j].y = 0.0f; Cv[j].y = 0.0f; }
                }
                sts8_priv(su, sel, uv);  // park the cleaned values for the second read
                sts8_priv(su + 2 * SEG, sel, dy);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 x2 = fma2(dl[j], f2(kLog2e), f2(bias2));
                float2 dt2 = x2;
                sig[j] = f2(1.0f);
                if (SP) {
                    float2 e, sp;
                    dt2 = softplus2_pair(x2, e, sp);
                    float2 r;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(sp.x));
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(sp.y));
                    sig[j] = mul2(e, r);
                    if (x2.x > kSoftplusThr2) sig[j].x = 1.0f;
                    if (x2.y > kSoftplusThr2) sig[j].y = 1.0f;
                }
                dtn[j] = mul2(dt2, f2(kLn2));
                const float2 da = mul2(dt2, f2(Av));
                av[j] = make_float2(ex2_approx(da.x), ex2_approx(da.y));
                bx[j] = mul2(dtn[j], mul2(Bv[j], uv[j]));
                cdy[j] = mul2(Cv[j], dy[j]);
            }
            sts8_priv(su + SEG, sel, sig);
        }
        // local aggregates of both recurrences in one walk in time order:
        //   forward   s -> p s + q;      adjoint (entering from the future)  G -> p G + qr,  qr = sum_i (prod_{j<=i} a_j) C_i dout_i
        float p = 1.0f, q = 0.0f, qr = 0.0f;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = pair_at<REV>(jj);
            walk_pair2<REV>(av[j], bx[j], cdy[j], p, q, qr);
        }
        const Aff inc_f = warp_scan_up_fast<32>(Aff{p, q});
        const Aff inc_r = warp_scan_down_fast<32>(Aff{p, qr});
        const Aff exc_f = shift_up1(inc_f, lane);
        const Aff exc_r = shift_down1(inc_r, lane);

        float h_warp = 0.0f, g_warp = 0.0f;  // states entering this warp: h from the left, adjoint from the right
        if (WPR > 1) {
            float4 *tot = s_tot + (it & 1) * 8;
            if (lane == 31) *reinterpret_cast<float2 *>(&tot[warp_slot]) = make_float2(inc_f.p, inc_f.q);
            if (lane == 0) tot[warp_slot].z = inc_r.q;
            row_barrier(1 + row, TPR);  // warp totals visible; everyone is done with the stage of iteration it - 1
            if (t_in_row == 0 && it > 0 && it - 1 + STAGES < n_iter) issue_stage(it - 1 + STAGES);
            if (leader && lane < 3 && it > 0) flush_sums(it - 1);
            if (multi) {
                if (leader) {
                    float4 t = make_float4(1.0f, 0.0f, 0.0f, 0.0f);
                    if (lane < WPR) t = tot[row * WPR + lane];
                    const Aff cum_f = warp_scan_up_fast<WPR>(Aff{t.x, t.y});
                    const Aff cum_r = warp_scan_down_fast<WPR>(Aff{t.x, t.z});
                    const Aff before_f = shift_up1(cum_f, lane);
                    const Aff before_r = shift_down1(cum_r, lane);
                    const Aff total_r = {__shfl_sync(0xffffffffu, cum_r.p, 0), __shfl_sync(0xffffffffu, cum_r.q, 0)};
                    if (lane == 0) publish_entry(l1_row + jrev, epoch, total_r.p, total_r.q);
                    bool ok;
                    Aff grp = {1.0f, 0.0f};
                    Aff acc = look_reduce(look, epoch, lane, ok, grp);
                    acc = look_finish(look, acc, ok, l2_row, jrev, epoch, lane, grp);
                    if (lane == 0 && (jrev & 15) == 15) {
                        const Aff g16 = compose(grp, total_r);
                        publish_entry(l2_row + (jrev >> 4), epoch, g16.p, g16.q);
                    }
                    if (lane < WPR)
                        s_in[row * WPR + lane] = make_float2(fmaf(before_f.p, h_in, before_f.q), fmaf(before_r.p, acc.q, before_r.q));
                }
                row_barrier(1 + row, TPR);
                const float2 in = s_in[warp_slot];
                h_warp = in.x;
                g_warp = in.y;
            } else {
#pragma unroll
                for (int w = 0; w < WPR - 1; ++w) {
                    const float4 t = tot[row * WPR + w];
                    if (w < warp_in_row) h_warp = fmaf(t.x, h_warp, t.y);
                }
#pragma unroll
                for (int w = WPR - 1; w > 0; --w) {
                    const float4 t = tot[row * WPR + w];
                    if (w > warp_in_row) g_warp = fmaf(t.x, g_warp, t.z);
                }
            }
        } else {
            __syncwarp();
            if (lane == 0 && it > 0 && it - 1 + STAGES < n_iter) issue_stage(it - 1 + STAGES);
        }

        // forward states of this thread's positions
        float2 hs[4];
        {
            float h = fmaf(exc_f.p, h_warp, exc_f.q);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int j = pair_at<REV>(jj);
                walk_state<REV>(av[j], bx[j], h, hs[j]);
            }
        }
        // adjoint walk, against time order
        float2 gl[4];
        {
            float G = fmaf(exc_r.p, g_warp, exc_r.q);
#pragma unroll
            for (int jj = 3; jj >= 0; --jj) {
                const int j = pair_at<REV>(jj);
                walk_adjoint<REV>(av[j], cdy[j], G, gl[j]);
            }
        }
        // gradients, position pairs
        float2 du[4], ddl[4], uv[4], dy[4], sig[4];
        lds8_priv(su, sel, uv);
        lds8_priv(su + 2 * SEG, sel, dy);
        lds8_priv(su + SEG, sel, sig);
        float2 sA = f2(0.0f), sD = f2(0.0f), sB = f2(0.0f);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 carried = fma2(bx[j], f2(-1.0f), hs[j]);  // a_l h_{l-1}
            const float2 w = mul2(gl[j], dtn[j]);
            du[j] = fma2(w, Bv[j], mul2(dy[j], f2(Dv)));
            const float2 ddt = mul2(gl[j], fma2(carried, f2(Av), mul2(Bv[j], uv[j])));  // g (B u + A a h_prev)
            ddl[j] = mul2(ddt, sig[j]);
            sA = fma2(w, carried, sA);
            dBacc[j] = fma2(w, uv[j], dBacc[j]);
            dCacc[j] = fma2(dy[j], hs[j], dCacc[j]);
            sD = fma2(dy[j], uv[j], sD);
            sB = add2(sB, ddl[j]);
        }
        {
            float *o_du = du_ptr + it * du_step;
            float *o_dd = dd_ptr + it * dd_step;
            if (!TAIL || nvalid == ITEMS) {
                if (accum) {
                    red8(o_du, du);
                } else {
                    if (addm) {
                        float2 old[4];
                        ldg8(o_du, old);
#pragma unroll
                        for (int j = 0; j < 4; ++j) du[j] = add2(old[j], du[j]);
                    }
                    stg8(o_du, du);
                }
                stg8(o_dd, ddl);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (2 * j < nvalid) {
                        if (accum || addm) atomicAdd(o_du + 2 * j, du[j].x); else o_du[2 * j] = du[j].x;
                        o_dd[2 * j] = ddl[j].x;
                    }
                    if (2 * j + 1 < nvalid) {
                        if (accum || addm) atomicAdd(o_du + 2 * j + 1, du[j].y); else o_du[2 * j + 1] = du[j].y;
                        o_dd[2 * j + 1] = ddl[j].y;
                    }
                }
            }
        }
        // per-channel sums: dA, dD, ddelta_bias (lanes 0, 8, 16 hold them after the reduction)
        const float r = warp_sum3(sA.x + sA.y, sD.x + sD.y, sB.x + sB.y, lane);
        if ((lane & 7) == 0 && lane < 24) {
            const int slot = lane >> 3;
            if (WPR > 1) {
                atomicAdd(s_red + ((it & 1) * 8 + row) * 4 + slot, r);
            } else {
                const int d = d0 + cc;
                if (slot == 0) atomicAdd(a.dA + d * a.A_ds, r);
                else if (slot == 1) { if (a.dD) atomicAdd(a.dD + d, r); }
                else { if (a.ddelta_bias) atomicAdd(a.ddelta_bias + d, r); }
            }
        }
    }
    if (WPR > 1 && n_iter > 0) {
        row_barrier(1 + row, TPR);
        if (leader && lane < 3) flush_sums(n_iter - 1);
    }

    // dB / dC of this tile's positions, summed over the tile's channels
    float *dBg = a.dB + b * a.dB_bs + (long long)g * L;
    float *dCg = a.dC + b * a.dC_bs + (long long)g * L;
    if (ROWS > 1) {
        __syncthreads();  // every row segment is done with the staging ring: reuse it
        float2 *sB2 = reinterpret_cast<float2 *>(s_stage) + (size_t)threadIdx.x * 4;  // [ROWS][TPR][4]
        float2 *sC2 = sB2 + NT * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            sB2[j] = dBacc[j];
            sC2[j] = dCacc[j];
        }
        __syncthreads();
        if (row == 0) {
            for (int r = 1; r < ROWS; ++r) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    dBacc[j] = add2(dBacc[j], sB2[r * TPR * 4 + j]);
                    dCacc[j] = add2(dCacc[j], sC2[r * TPR * 4 + j]);
                }
            }
        }
    }
    if (row == 0) {
        if (!TAIL || nvalid == ITEMS) {
            red8(dBg + pos, dBacc);
            red8(dCg + pos, dCacc);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (2 * j < nvalid) { atomicAdd(dBg + pos + 2 * j, dBacc[j].x); atomicAdd(dCg + pos + 2 * j, dCacc[j].x); }
                if (2 * j + 1 < nvalid) { atomicAdd(dBg + pos + 2 * j + 1, dBacc[j].y); atomicAdd(dCg + pos + 2 * j + 1, dCacc[j].y); }
            }
        }
    }
    retire_tile(a);
}

template <int TPR, bool SP>
__global__ void __launch_bounds__(256, 2) scan_bwd_tma_kernel(const __grid_constant__ GroupArgs ga) {
    extern __shared__ __align__(1024) unsigned char smem_bwd_tma[];  // swizzled tiles need 512-byte aligned slots
    pdl_launch_dependents();  // the next kernel on the stream may be scheduled while this one drains ...
    constexpr int SEG = TPR * 8;
    int gtile;
    const int prob = group_problem(ga, gtile);
    const ScanArgs &a = ga.a[prob];
    const TileMaps &tm = ga.tm[prob];
    // ... and this one touches global memory only after its predecessor has completed: the wait is in the body, behind the index
    // arithmetic and the mbarrier initialisation (here already when the tile is claimed by ticket, which is a global atomic)
    if ((a.pdl_mode & 1) || a.n_chunks > 1) pdl_wait();
    unsigned tile = (unsigned)gtile, epoch = 0;
    if (a.n_chunks > 1) claim_tile(a, reinterpret_cast<unsigned *>(smem_bwd_tma + 384), tile, epoch);  // VMASR_TUNING builds only
    const int chunk = a.n_chunks - 1 - (int)(tile / a.n_rowgroups);  // adjoint: late chunks first
    const int rg = tile % a.n_rowgroups;
    const bool tail = (chunk + 1) * SEG > a.seqlen;
    if (a.rev) {  // single chunk only (host-checked)
        if (tail) scan_bwd_tma_body<TPR, true, SP, true>(a, tm, smem_bwd_tma, chunk, rg, epoch);
        else scan_bwd_tma_body<TPR, false, SP, true>(a, tm, smem_bwd_tma, chunk, rg, epoch);
    } else {
        if (tail) scan_bwd_tma_body<TPR, true, SP, false>(a, tm, smem_bwd_tma, chunk, rg, epoch);
        else scan_bwd_tma_body<TPR, false, SP, false>(a, tm, smem_bwd_tma, chunk, rg, epoch);
    }
}

static size_t scan_bwd_tma_smem(int tpr) {
    const size_t seg = (size_t)tpr * 8, rows = 256 / tpr;
    return 2048 + sizeof(float) * (2 * seg + (size_t)kBwdStages * rows * 3 * seg);
}

template <int TPR, bool SP>
static int launch_tma(const GroupArgs &ga, int grid, cudaStream_t stream) {
    const size_t smem = scan_bwd_tma_smem(TPR);
    static PerDeviceOnce configured;  // the attribute is per function and per device
    if (!configured()) {
        if (int rc = check_cuda(cudaFuncSetAttribute(scan_bwd_tma_kernel<TPR, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                                "scan_bwd_tma smem attribute"))
            return rc;
        configured() = true;
    }
    return launch_pdl(scan_bwd_tma_kernel<TPR, SP>, grid, 256, smem, stream, "scan_bwd_tma launch", ga);
}

template <bool SP>
static int dispatch_tpr(const GroupArgs &ga, int tpr, int grid, cudaStream_t stream) {
    switch (tpr) {
        case 32: return launch_tma<32, SP>(ga, grid, stream);
        case 64: return launch_tma<64, SP>(ga, grid, stream);
        case 128: return launch_tma<128, SP>(ga, grid, stream);
        default: return launch_tma<256, SP>(ga, grid, stream);
    }
}

// every problem of the group: same threads-per-row and softplus flag (scan_host.cu groups them so)
int scan_bwd_tma_dispatch(const GroupArgs &ga, int tpr, int grid, cudaStream_t stream) {
    for (int i = 0; i < ga.n; ++i) {
        if (ga.a[i].rev && ga.a[i].n_chunks > 1) return fail("scan_bwd_tma: reversed scans of more than one chunk go to the multi-chunk kernel");
        if (ga.n > 1 && ga.a[i].n_chunks > 1) return fail("scan_bwd_tma: grouped launches take single-chunk problems only");
    }
    return ga.a[0].softplus ? dispatch_tpr<true>(ga, tpr, grid, stream) : dispatch_tpr<false>(ga, tpr, grid, stream);
}

}  // namespace vmasr
