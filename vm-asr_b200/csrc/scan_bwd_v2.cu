// Selective-scan backward, fast path for sm_100a: fp32 IO, d_state 1, 16-byte aligned rows.
// Replaces selective_scan_bwd_kernel (kernels/selective_scan/csrc/selective_scan/cus/selective_scan_bwd_kernel.cuh:66-273).
// Per (batch, channel), g the adjoint of the state h:
//   g_l      = C_l dout_l + a_{l+1} g_{l+1}                      (right to left)
//   du_l     = D dout_l + g_l dt_l B_l
//   ddt_l    = g_l (B_l u_l + A a_l h_{l-1})
//   dA       = sum_l g_l dt_l a_l h_{l-1};  dB_l += g_l dt_l u_l;  dC_l += dout_l h_l;  dD = sum dout u
//   ddelta_l = ddt_l * sigmoid(delta_l + bias)  (softplus on, input <= 20);  ddelta_bias = sum_l ddelta_l
// Same persistent structure as scan_fwd_v2.cu (persist.cuh): tiles taken in scan order (here the LAST chunk first:
// the adjoint runs right to left), u / delta / dout row segments streamed through a 3-stage TMA ring three row
// passes ahead, 8 positions per thread.  The forward states are
// recomputed from the chunk-end states `x` the forward saved; the adjoint carry uses the two-level look-back.
// dB / dC of a tile's positions are summed over the tile's channels in registers and leave with one 128-bit
// reduction per 4 positions; dA / dD / ddelta_bias leave with one reduction per warp and row.
#include "persist.cuh"

namespace vmasr {

constexpr int kBwdStages = 3;
constexpr int kBwdConsumers = 256;
constexpr int kBwdItems = 8;

struct BwdSmem {
    unsigned long long full[kBwdStages], bc_full[2];
    float4 tot[2][8];  // [buffer][row * WPR + warp]: {decay, forward q, adjoint q, -}
    FeedState fs;
};
static_assert(sizeof(BwdSmem) <= 1024, "smem header");

__device__ __forceinline__ float warp_sum_v2(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

__device__ __forceinline__ void red_add4(float *p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int TPR, bool SOFTPLUS>
__global__ void __launch_bounds__(kBwdConsumers, 2) scan_bwd_v2_kernel(const __grid_constant__ ScanArgs a) {
    constexpr int ROWS = kBwdConsumers / TPR;
    constexpr int WPR = TPR / 32;
    constexpr int SEG = TPR * kBwdItems;
    constexpr int STAGE_FLOATS = 3 * ROWS * SEG;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BwdSmem &sm = *reinterpret_cast<BwdSmem *>(smem_raw);
    float *s_bc = reinterpret_cast<float *>(smem_raw + 1024);  // [2 slots][B | C][SEG]
    float *s_stage = s_bc + 2 * 2 * SEG;                      // [stages][u | delta | dout][ROWS * SEG]
    float *s_dbc = s_stage + kBwdStages * STAGE_FLOATS;       // [dB | dC][ROWS][SEG]  (ROWS > 1 only)

    const int L = a.seqlen;
    if (threadIdx.x == 0) feed_start<true, SEG, ROWS, 3, kBwdStages>(a, sm.fs, s_stage, sm.full, s_bc, sm.bc_full);
    __syncthreads();

    const int row = threadIdx.x / TPR;
    const int t_in_row = threadIdx.x - row * TPR;
    const int warp_in_row = t_in_row >> 5;
    const int lane = threadIdx.x & 31;
    const unsigned epoch = (a.n_chunks > 1) ? (*reinterpret_cast<volatile unsigned *>(a.ws_header + 2) % 0xfffffffeu + 1u) : 0u;
    const int n_groups16 = (a.n_chunks + 15) >> 4;

    Ring<kBwdStages> ring;
    int buf = 0;
    for (int n = 0;; ++n) {
        const TileDesc td = sm.fs.q[n % kQueue];
        if (td.tile < 0) break;
        const int j = td.j;          // scan-order index of the chunk (adjoint: last chunk first)
        const int chunk = td.chunk;
        const int b = td.b, d0 = td.d0, n_chan = td.n_chan;
        const int seg0 = chunk * SEG;
        const int pos = seg0 + t_in_row * kBwdItems;
        const int nvalid = max(0, min(kBwdItems, L - pos));
        const bool full_tile = seg0 + SEG <= L;
        const int n_iter = (n_chan + ROWS - 1) / ROWS;

        float Bv[kBwdItems], Cv[kBwdItems], dBacc[kBwdItems], dCacc[kBwdItems];
        mbar_wait(&sm.bc_full[n & 1], (unsigned)(n >> 1) & 1u);
        {
            const float *sb = s_bc + (n & 1) * 2 * SEG + t_in_row * kBwdItems;
            const float4 b0 = *reinterpret_cast<const float4 *>(sb), b1 = *reinterpret_cast<const float4 *>(sb + 4);
            const float4 c0 = *reinterpret_cast<const float4 *>(sb + SEG), c1 = *reinterpret_cast<const float4 *>(sb + SEG + 4);
            Bv[0] = b0.x; Bv[1] = b0.y; Bv[2] = b0.z; Bv[3] = b0.w; Bv[4] = b1.x; Bv[5] = b1.y; Bv[6] = b1.z; Bv[7] = b1.w;
            Cv[0] = c0.x; Cv[1] = c0.y; Cv[2] = c0.z; Cv[3] = c0.w; Cv[4] = c1.x; Cv[5] = c1.y; Cv[6] = c1.z; Cv[7] = c1.w;
        }
#pragma unroll
        for (int i = 0; i < kBwdItems; ++i) {
            dBacc[i] = dCacc[i] = 0.0f;
            if (!full_tile && i >= nvalid) { Bv[i] = 0.0f; Cv[i] = 0.0f; }
        }

        for (int it = 0; it < n_iter; ++it) {
            const int cc = it * ROWS + row;
            const bool active = ROWS == 1 || cc < n_chan;
            const float act = active ? 1.0f : 0.0f;
            const int ccl = active ? cc : 0;
            const int d = d0 + ccl;
            const long long seq = td.seq0 + ccl;

            if (it == 0 && threadIdx.x == 0) feed_claim<true, SEG>(a, sm.fs);  // post the id of the tile kBwdStages + 2 ahead

            CarryLook look;
            const CarryEntry *l2_row = nullptr;
            float h_in = 0.0f;
            if (a.n_chunks > 1) {
                l2_row = a.ws_entries2 + seq * n_groups16;
                look = look_issue(a.ws_entries + seq * a.n_chunks, l2_row, j, lane);
                if (chunk > 0) h_in = __ldg(a.x + (seq * a.n_chunks + (chunk - 1)) * 2 + 1);
            }
            const float Aval = __ldg(a.A + d * a.A_ds);
            const float A2 = Aval * kLog2e;
            const float Dv = a.D ? __ldg(a.D + d) : 0.0f;
            const float bias = a.delta_bias ? __ldg(a.delta_bias + d) : 0.0f;
            const float bias2 = bias * kLog2e;

            float uv[kBwdItems], dt[kBwdItems], dy[kBwdItems];
            mbar_wait(&sm.full[ring.stage], ring.phase);
            {
                const float *su = s_stage + ring.stage * STAGE_FLOATS + (active ? row : 0) * SEG + t_in_row * kBwdItems;
                const float4 u0 = *reinterpret_cast<const float4 *>(su), u1 = *reinterpret_cast<const float4 *>(su + 4);
                const float4 e0 = *reinterpret_cast<const float4 *>(su + ROWS * SEG), e1 = *reinterpret_cast<const float4 *>(su + ROWS * SEG + 4);
                const float4 y0 = *reinterpret_cast<const float4 *>(su + 2 * ROWS * SEG), y1 = *reinterpret_cast<const float4 *>(su + 2 * ROWS * SEG + 4);
                uv[0] = u0.x; uv[1] = u0.y; uv[2] = u0.z; uv[3] = u0.w; uv[4] = u1.x; uv[5] = u1.y; uv[6] = u1.z; uv[7] = u1.w;
                dt[0] = e0.x; dt[1] = e0.y; dt[2] = e0.z; dt[3] = e0.w; dt[4] = e1.x; dt[5] = e1.y; dt[6] = e1.z; dt[7] = e1.w;
                dy[0] = y0.x; dy[1] = y0.y; dy[2] = y0.z; dy[3] = y0.w; dy[4] = y1.x; dy[5] = y1.y; dy[6] = y1.z; dy[7] = y1.w;
            }

            float av[kBwdItems], sig[kBwdItems];
#pragma unroll
            for (int i = 0; i < kBwdItems; ++i) {
                if (SOFTPLUS) {
                    dt[i] = softplus2_sig(fmaf(dt[i], kLog2e, bias2), dt[i] + bias, sig[i]);
                } else {
                    dt[i] += bias;
                    sig[i] = 1.0f;
                }
                av[i] = ex2_approx(dt[i] * A2);
            }
            if (!full_tile) {
#pragma unroll
                for (int i = 0; i < kBwdItems; ++i)
                    if (i >= nvalid) { av[i] = 1.0f; dt[i] = 0.0f; sig[i] = 0.0f; uv[i] = 0.0f; dy[i] = 0.0f; }
            }
            float dD_acc = 0.0f;
            Aff loc_f = {1.0f, 0.0f};
#pragma unroll
            for (int i = 0; i < kBwdItems; ++i) {
                loc_f.q = fmaf(av[i], loc_f.q, dt[i] * uv[i] * Bv[i]);
                loc_f.p *= av[i];
                dD_acc = fmaf(dy[i], uv[i], dD_acc);
            }
            float qr = 0.0f;
#pragma unroll
            for (int i = kBwdItems - 1; i >= 0; --i) qr = av[i] * fmaf(Cv[i], dy[i], qr);
            const Aff loc_r = {loc_f.p, qr};

            Aff inc_f = warp_scan_up(loc_f, lane);
            Aff inc_r = warp_scan_down(loc_r, lane);
            Aff exc_f = {__shfl_up_sync(0xffffffffu, inc_f.p, 1), __shfl_up_sync(0xffffffffu, inc_f.q, 1)};
            Aff exc_r = {__shfl_down_sync(0xffffffffu, inc_r.p, 1), __shfl_down_sync(0xffffffffu, inc_r.q, 1)};
            if (lane == 0) exc_f = {1.0f, 0.0f};
            if (lane == 31) exc_r = {1.0f, 0.0f};
            Aff total_r;
            if (WPR > 1) {
                if (lane == 31) {
                    sm.tot[buf][row * WPR + warp_in_row].x = inc_f.p;
                    sm.tot[buf][row * WPR + warp_in_row].y = inc_f.q;
                }
                if (lane == 0) sm.tot[buf][row * WPR + warp_in_row].z = inc_r.q;
            }
            __syncthreads();  // warp totals visible; every thread has pulled this stage into registers
            if (threadIdx.x == 0) {
                if (it == 0) {  // this tile's B/C slot is free again: fetch the B/C segment of the tile after the next
                    const int next2 = sm.fs.q[(n + 2) % kQueue].tile;
                    if (next2 >= 0) feed_issue_bc<true, SEG>(a, next2, s_bc + (n & 1) * 2 * SEG, &sm.bc_full[n & 1]);
                }
                feed_issue_pass<true, SEG, ROWS, 3, kBwdStages>(a, sm.fs, s_stage, sm.full);
            }
            ring.advance();
            if (WPR > 1) {
                Aff before_f = {1.0f, 0.0f}, run = {1.0f, 0.0f};
#pragma unroll
                for (int w = 0; w < WPR; ++w) {
                    const float4 t = sm.tot[buf][row * WPR + w];
                    if (w == warp_in_row) before_f = run;
                    run = compose(run, Aff{t.x, t.y});
                }
                Aff before_r = {1.0f, 0.0f};
                run = {1.0f, 0.0f};
#pragma unroll
                for (int w = WPR - 1; w >= 0; --w) {
                    const float4 t = sm.tot[buf][row * WPR + w];
                    if (w == warp_in_row) before_r = run;
                    run = compose(run, Aff{t.x, t.z});
                }
                total_r = run;
                exc_f = compose(before_f, exc_f);
                exc_r = compose(before_r, exc_r);
                buf ^= 1;
            } else {
                total_r = {__shfl_sync(0xffffffffu, inc_r.p, 0), __shfl_sync(0xffffffffu, inc_r.q, 0)};
            }

            float g_in = 0.0f;
            if (a.n_chunks > 1) {  // one row per CTA pass in this case (TPR == 256)
                if (threadIdx.x == 0) publish_entry(a.ws_entries + seq * a.n_chunks + j, epoch, total_r.p, total_r.q);
                Aff ingroup;
                const Aff acc = look_resolve(look, l2_row, j, epoch, lane, ingroup);
                if (threadIdx.x == 0 && (j & 15) == 15) {
                    const Aff grp = compose(ingroup, total_r);
                    publish_entry(a.ws_entries2 + seq * n_groups16 + (j >> 4), epoch, grp.p, grp.q);
                }
                g_in = acc.q;
            }

            // forward states of this thread's positions
            float hs[kBwdItems];
            const float h_start = fmaf(exc_f.p, h_in, exc_f.q);
            {
                float h = h_start;
#pragma unroll
                for (int i = 0; i < kBwdItems; ++i) {
                    h = fmaf(av[i], h, dt[i] * uv[i] * Bv[i]);
                    hs[i] = h;
                }
            }
            // adjoint walk, right to left
            float G = fmaf(exc_r.p, g_in, exc_r.q);
            float dA_acc = 0.0f, dbias_acc = 0.0f;
            float du[kBwdItems], ddl[kBwdItems];
#pragma unroll
            for (int i = kBwdItems - 1; i >= 0; --i) {
                const float gl = fmaf(Cv[i], dy[i], G);
                G = av[i] * gl;
                const float carried = av[i] * (i > 0 ? hs[i - 1] : h_start);
                const float gdt = gl * dt[i];
                du[i] = fmaf(gdt, Bv[i], Dv * dy[i]);
                ddl[i] = gl * fmaf(Bv[i], uv[i], Aval * carried) * sig[i];
                dbias_acc += ddl[i];
                dA_acc = fmaf(gdt, carried, dA_acc);
                dBacc[i] = fmaf(gdt * act, uv[i], dBacc[i]);
                dCacc[i] = fmaf(dy[i] * act, hs[i], dCacc[i]);
            }
            if (active) {
                float *pu = reinterpret_cast<float *>(a.du) + b * a.du_bs + (long long)d * a.du_ds + pos;
                float *pd = reinterpret_cast<float *>(a.ddelta) + b * a.ddelta_bs + (long long)d * a.ddelta_ds + pos;
                if (full_tile || nvalid == kBwdItems) {
                    reinterpret_cast<float4 *>(pu)[0] = make_float4(du[0], du[1], du[2], du[3]);
                    reinterpret_cast<float4 *>(pu)[1] = make_float4(du[4], du[5], du[6], du[7]);
                    reinterpret_cast<float4 *>(pd)[0] = make_float4(ddl[0], ddl[1], ddl[2], ddl[3]);
                    reinterpret_cast<float4 *>(pd)[1] = make_float4(ddl[4], ddl[5], ddl[6], ddl[7]);
                } else {
#pragma unroll
                    for (int i = 0; i < kBwdItems; ++i)
                        if (i < nvalid) { pu[i] = du[i]; pd[i] = ddl[i]; }
                }
            }
            dA_acc = warp_sum_v2(dA_acc);
            dD_acc = warp_sum_v2(dD_acc);
            dbias_acc = warp_sum_v2(dbias_acc);
            if (lane == 0 && active) {
                atomicAdd(a.dA + d * a.A_ds, dA_acc);
                if (a.dD) atomicAdd(a.dD + d, dD_acc);
                if (a.ddelta_bias) atomicAdd(a.ddelta_bias + d, dbias_acc);
            }
        }

        // dB / dC of this tile's positions, summed over the tile's channels
        float *dBg = a.dB + (long long)td.bg * (long long)L + pos;
        float *dCg = a.dC + (long long)td.bg * (long long)L + pos;
        if (ROWS > 1) {
            float *sB = s_dbc + (row * TPR + t_in_row) * kBwdItems;
            float *sC = sB + ROWS * SEG;
            __syncthreads();  // previous tile's readers are done with s_dbc
            if (row > 0) {
                *reinterpret_cast<float4 *>(sB) = make_float4(dBacc[0], dBacc[1], dBacc[2], dBacc[3]);
                *reinterpret_cast<float4 *>(sB + 4) = make_float4(dBacc[4], dBacc[5], dBacc[6], dBacc[7]);
                *reinterpret_cast<float4 *>(sC) = make_float4(dCacc[0], dCacc[1], dCacc[2], dCacc[3]);
                *reinterpret_cast<float4 *>(sC + 4) = make_float4(dCacc[4], dCacc[5], dCacc[6], dCacc[7]);
            }
            __syncthreads();
            if (row == 0) {
                for (int r = 1; r < ROWS; ++r) {
#pragma unroll
                    for (int i = 0; i < kBwdItems; ++i) {
                        dBacc[i] += sB[r * SEG + i];
                        dCacc[i] += sC[r * SEG + i];
                    }
                }
            }
        }
        if (row == 0) {
            if (full_tile || nvalid == kBwdItems) {
                red_add4(dBg, dBacc[0], dBacc[1], dBacc[2], dBacc[3]);
                red_add4(dBg + 4, dBacc[4], dBacc[5], dBacc[6], dBacc[7]);
                red_add4(dCg, dCacc[0], dCacc[1], dCacc[2], dCacc[3]);
                red_add4(dCg + 4, dCacc[4], dCacc[5], dCacc[6], dCacc[7]);
            } else {
#pragma unroll
                for (int i = 0; i < kBwdItems; ++i)
                    if (i < nvalid) { atomicAdd(dBg + i, dBacc[i]); atomicAdd(dCg + i, dCacc[i]); }
            }
        }
    }
    if (a.n_chunks > 1) {
        __syncthreads();
        if (threadIdx.x == 0) retire_cta(a);
    }
}

template <int TPR, bool SOFTPLUS>
static int launch_bwd_v2(const ScanArgs &a, int grid, cudaStream_t stream) {
    constexpr int ROWS = kBwdConsumers / TPR;
    const size_t smem = 1024 + sizeof(float) * (2 * 2 * TPR * kBwdItems + kBwdStages * 3 * 2048 + (ROWS > 1 ? 2 * 2048 : 0));
    static int resident = 0;  // CTAs of this instantiation one SM holds (the round-robin deal needs the whole grid resident)
    static bool configured = false;
    if (!configured) {
        if (int rc = check_cuda(cudaFuncSetAttribute(scan_bwd_v2_kernel<TPR, SOFTPLUS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                                "scan_bwd smem attribute"))
            return rc;
        if (int rc = check_cuda(cudaFuncSetAttribute(scan_bwd_v2_kernel<TPR, SOFTPLUS>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                                     cudaSharedmemCarveoutMaxShared),
                                "scan_bwd carveout attribute"))
            return rc;
        if (int rc = check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, scan_bwd_v2_kernel<TPR, SOFTPLUS>, kBwdConsumers, smem), "occupancy query"))
            return rc;
        if (resident < 1) return fail("selective_scan: persistent kernel does not fit on an SM");
        configured = true;
    }
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    sms = sm_count(dev);
    if (grid > resident * sms) grid = resident * sms;
    scan_bwd_v2_kernel<TPR, SOFTPLUS><<<grid, kBwdConsumers, smem, stream>>>(a);
    return check_cuda(cudaGetLastError(), "scan_bwd launch");
}

template <bool SOFTPLUS>
static int bwd_by_tpr(const ScanArgs &a, const ScanPlan &pl, cudaStream_t stream) {
    switch (pl.tpr) {
        case 32: return launch_bwd_v2<32, SOFTPLUS>(a, pl.grid, stream);
        case 64: return launch_bwd_v2<64, SOFTPLUS>(a, pl.grid, stream);
        case 128: return launch_bwd_v2<128, SOFTPLUS>(a, pl.grid, stream);
        default: return launch_bwd_v2<256, SOFTPLUS>(a, pl.grid, stream);
    }
}

int scan_bwd_v2_dispatch(const ScanArgs &a, const ScanPlan &pl, cudaStream_t stream) {
    return a.softplus ? bwd_by_tpr<true>(a, pl, stream) : bwd_by_tpr<false>(a, pl, stream);
}

}  // namespace vmasr
