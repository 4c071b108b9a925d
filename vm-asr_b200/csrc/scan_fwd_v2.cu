// Selective-scan forward, fast path for sm_100a: fp32 IO, d_state 1, 16-byte aligned rows.
// Replaces selective_scan_fwd_kernel (kernels/selective_scan/csrc/selective_scan/cus/selective_scan_fwd_kernel.cuh:61-172):
//   dt = softplus(delta + delta_bias);  h_l = exp(dt*A) h_{l-1} + dt*B_l*u_l;  out_l = C_l h_l + D u_l
// and writes the per-chunk (cumulative decay, end state) tensor `x` the backward needs.
//
// Persistent kernel, 2 CTAs of 256 threads per SM (persist.cuh describes the tile feed):
//   * u / delta row segments arrive through a 4-stage shared-memory ring filled by TMA bulk copies that run four
//     row passes ahead, across tile boundaries; the tile's B / C segment arrives the same way one tile ahead;
//   * a thread pulls its 8 positions of a stage into registers and does the scan: serial over its positions,
//     shuffle scan over the warp, ONE barrier per row pass for the warps of a row (the same barrier hands the
//     stage back to the feeder), two-level look-back over the chunks whose loads were issued before the
//     arithmetic (pipe.cuh);
//   * outputs go straight from registers to HBM with 128-bit stores.
#include "persist.cuh"

namespace vmasr {

constexpr int kFwdStages = 4;
constexpr int kFwdThreads = 256;
constexpr int kItems = 8;

struct FwdSmem {
    unsigned long long full[kFwdStages], bc_full[2];
    float2 tot[2][8];  // [buffer][row * WPR + warp]
    FeedState fs;
};
static_assert(sizeof(FwdSmem) <= 1024, "smem header");

template <int TPR, bool SOFTPLUS>
__global__ void __launch_bounds__(kFwdThreads, 2) scan_fwd_v2_kernel(const __grid_constant__ ScanArgs a) {
    constexpr int ROWS = kFwdThreads / TPR;
    constexpr int WPR = TPR / 32;
    constexpr int SEG = TPR * kItems;  // positions per row segment
    constexpr int STAGE_FLOATS = 2 * ROWS * SEG;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FwdSmem &sm = *reinterpret_cast<FwdSmem *>(smem_raw);
    float *s_bc = reinterpret_cast<float *>(smem_raw + 1024);  // [2 slots][B | C][SEG]
    float *s_stage = s_bc + 2 * 2 * SEG;                      // [stages][u | delta][ROWS][SEG]

    const int L = a.seqlen;
    if (threadIdx.x == 0) feed_start<false, SEG, ROWS, 2, kFwdStages>(a, sm.fs, s_stage, sm.full, s_bc, sm.bc_full);
    __syncthreads();

    const int row = threadIdx.x / TPR;
    const int t_in_row = threadIdx.x - row * TPR;
    const int warp_in_row = t_in_row >> 5;
    const int lane = threadIdx.x & 31;
    const unsigned epoch = (a.n_chunks > 1) ? (*reinterpret_cast<volatile unsigned *>(a.ws_header + 2) % 0xfffffffeu + 1u) : 0u;
    const int n_groups16 = (a.n_chunks + 15) >> 4;

    Ring<kFwdStages> ring;
    int buf = 0;
    for (int n = 0;; ++n) {
        const TileDesc td = sm.fs.q[n % kQueue];
        if (td.tile < 0) break;
        const int chunk = td.chunk;
        const int seg0 = chunk * SEG;
        const int pos = seg0 + t_in_row * kItems;
        const int nvalid = max(0, min(kItems, L - pos));
        const bool full_tile = seg0 + SEG <= L;
        const int n_iter = (td.n_chan + ROWS - 1) / ROWS;
        float *out_tile = reinterpret_cast<float *>(a.out) + td.b * a.out_bs + (long long)td.d0 * a.out_ds + pos;

        float Bv[kItems], Cv[kItems];
        mbar_wait(&sm.bc_full[n & 1], (unsigned)(n >> 1) & 1u);
        {
            const float *sb = s_bc + (n & 1) * 2 * SEG + t_in_row * kItems;
            const float4 b0 = *reinterpret_cast<const float4 *>(sb), b1 = *reinterpret_cast<const float4 *>(sb + 4);
            const float4 c0 = *reinterpret_cast<const float4 *>(sb + SEG), c1 = *reinterpret_cast<const float4 *>(sb + SEG + 4);
            Bv[0] = b0.x; Bv[1] = b0.y; Bv[2] = b0.z; Bv[3] = b0.w; Bv[4] = b1.x; Bv[5] = b1.y; Bv[6] = b1.z; Bv[7] = b1.w;
            Cv[0] = c0.x; Cv[1] = c0.y; Cv[2] = c0.z; Cv[3] = c0.w; Cv[4] = c1.x; Cv[5] = c1.y; Cv[6] = c1.z; Cv[7] = c1.w;
        }
        if (!full_tile) {
#pragma unroll
            for (int i = 0; i < kItems; ++i)
                if (i >= nvalid) { Bv[i] = 0.0f; Cv[i] = 0.0f; }
        }

        for (int it = 0; it < n_iter; ++it) {
            const int cc = it * ROWS + row;
            const bool active = ROWS == 1 || cc < td.n_chan;
            const int ccl = active ? cc : 0;
            const int d = td.d0 + ccl;
            const long long seq = td.seq0 + ccl;

            if (it == 0 && threadIdx.x == 0) feed_claim<false, SEG>(a, sm.fs);  // post the tile kFwdStages + 2 ahead

            // look-back loads first: they fly while the row is computed
            CarryLook look;
            const CarryEntry *l2_row = nullptr;
            if (a.n_chunks > 1) {
                l2_row = a.ws_entries2 + seq * n_groups16;
                look = look_issue(a.ws_entries + seq * a.n_chunks, l2_row, chunk, lane);
            }
            const float A2 = __ldg(a.A + d * a.A_ds) * kLog2e;
            const float Dv = a.D ? __ldg(a.D + d) : 0.0f;
            const float bias = a.delta_bias ? __ldg(a.delta_bias + d) : 0.0f;
            const float bias2 = bias * kLog2e;

            float uv[kItems], dt[kItems];
            mbar_wait(&sm.full[ring.stage], ring.phase);
            {
                const float *su = s_stage + ring.stage * STAGE_FLOATS + (active ? row : 0) * SEG + t_in_row * kItems;
                const float4 u0 = *reinterpret_cast<const float4 *>(su), u1 = *reinterpret_cast<const float4 *>(su + 4);
                const float4 e0 = *reinterpret_cast<const float4 *>(su + ROWS * SEG), e1 = *reinterpret_cast<const float4 *>(su + ROWS * SEG + 4);
                uv[0] = u0.x; uv[1] = u0.y; uv[2] = u0.z; uv[3] = u0.w; uv[4] = u1.x; uv[5] = u1.y; uv[6] = u1.z; uv[7] = u1.w;
                dt[0] = e0.x; dt[1] = e0.y; dt[2] = e0.z; dt[3] = e0.w; dt[4] = e1.x; dt[5] = e1.y; dt[6] = e1.z; dt[7] = e1.w;
            }

            float av[kItems], bx[kItems];
#pragma unroll
            for (int i = 0; i < kItems; ++i) {
                float d_t;
                if (SOFTPLUS) d_t = softplus2(fmaf(dt[i], kLog2e, bias2), dt[i] + bias);
                else d_t = dt[i] + bias;
                av[i] = ex2_approx(d_t * A2);
                bx[i] = d_t * uv[i] * Bv[i];
                uv[i] *= Dv;
            }
            if (!full_tile) {
#pragma unroll
                for (int i = 0; i < kItems; ++i)
                    if (i >= nvalid) { av[i] = 1.0f; bx[i] = 0.0f; }
            }
            Aff loc = {1.0f, 0.0f};
#pragma unroll
            for (int i = 0; i < kItems; ++i) {
                loc.q = fmaf(av[i], loc.q, bx[i]);
                loc.p *= av[i];
            }
            Aff inc = warp_scan_up(loc, lane);
            Aff exc = {__shfl_up_sync(0xffffffffu, inc.p, 1), __shfl_up_sync(0xffffffffu, inc.q, 1)};
            if (lane == 0) exc = {1.0f, 0.0f};
            if (WPR > 1 && lane == 31) sm.tot[buf][row * WPR + warp_in_row] = make_float2(inc.p, inc.q);

            __syncthreads();  // warp totals visible; every thread has pulled this stage into registers

            if (threadIdx.x == 0) {
                if (it == 0) {  // this tile's B/C slot is free again: fetch the B/C segment of the tile after the next
                    const int next2 = sm.fs.q[(n + 2) % kQueue].tile;
                    if (next2 >= 0) feed_issue_bc<false, SEG>(a, next2, s_bc + (n & 1) * 2 * SEG, &sm.bc_full[n & 1]);
                }
                feed_issue_pass<false, SEG, ROWS, 2, kFwdStages>(a, sm.fs, s_stage, sm.full);
            }
            ring.advance();

            Aff total;
            if (WPR > 1) {
                Aff before = {1.0f, 0.0f};
                total = {1.0f, 0.0f};
#pragma unroll
                for (int w = 0; w < WPR; ++w) {
                    const float2 t = sm.tot[buf][row * WPR + w];
                    if (w == warp_in_row) before = total;
                    total = compose(total, Aff{t.x, t.y});
                }
                exc = compose(before, exc);
                buf ^= 1;
            } else {
                total = {__shfl_sync(0xffffffffu, inc.p, 31), __shfl_sync(0xffffffffu, inc.q, 31)};
            }

            float h_in = 0.0f, pcum_in = 1.0f;
            if (a.n_chunks > 1) {  // one row per pass in this case (TPR == 256)
                if (threadIdx.x == 0) publish_entry(a.ws_entries + seq * a.n_chunks + chunk, epoch, total.p, total.q);
                Aff ingroup;
                const Aff acc = look_resolve(look, l2_row, chunk, epoch, lane, ingroup);
                if (threadIdx.x == 0 && (chunk & 15) == 15) {
                    const Aff grp = compose(ingroup, total);
                    publish_entry(a.ws_entries2 + seq * n_groups16 + (chunk >> 4), epoch, grp.p, grp.q);
                }
                h_in = acc.q;
                pcum_in = acc.p;
            }
            if (t_in_row == 0 && active)
                reinterpret_cast<float2 *>(a.x)[seq * a.n_chunks + chunk] = make_float2(total.p * pcum_in, fmaf(total.p, h_in, total.q));

            float h = fmaf(exc.p, h_in, exc.q);
            float y[kItems];
#pragma unroll
            for (int i = 0; i < kItems; ++i) {
                h = fmaf(av[i], h, bx[i]);
                y[i] = fmaf(Cv[i], h, uv[i]);
            }
            if (active) {
                float *o = out_tile + (long long)ccl * a.out_ds;
                if (full_tile || nvalid == kItems) {
                    reinterpret_cast<float4 *>(o)[0] = make_float4(y[0], y[1], y[2], y[3]);
                    reinterpret_cast<float4 *>(o)[1] = make_float4(y[4], y[5], y[6], y[7]);
                } else {
#pragma unroll
                    for (int i = 0; i < kItems; ++i)
                        if (i < nvalid) o[i] = y[i];
                }
            }
        }
    }
    if (a.n_chunks > 1) {
        __syncthreads();
        if (threadIdx.x == 0) retire_cta(a);
    }
}

template <int TPR, bool SOFTPLUS>
static int launch_v2(const ScanArgs &a, int grid, cudaStream_t stream) {
    const size_t smem = 1024 + sizeof(float) * (2 * 2 * TPR * kItems + kFwdStages * 2 * 2048);
    static int resident = 0;  // CTAs of this instantiation one SM holds (the round-robin deal needs the whole grid resident)
    static bool configured = false;  // per instantiation; setting the attributes repeatedly is harmless
    if (!configured) {
        if (int rc = check_cuda(cudaFuncSetAttribute(scan_fwd_v2_kernel<TPR, SOFTPLUS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                                "scan_fwd smem attribute"))
            return rc;
        if (int rc = check_cuda(cudaFuncSetAttribute(scan_fwd_v2_kernel<TPR, SOFTPLUS>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                                     cudaSharedmemCarveoutMaxShared),
                                "scan_fwd carveout attribute"))
            return rc;
        if (int rc = check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, scan_fwd_v2_kernel<TPR, SOFTPLUS>, kFwdThreads, smem), "occupancy query"))
            return rc;
        if (resident < 1) return fail("selective_scan: persistent kernel does not fit on an SM");
        configured = true;
    }
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    sms = sm_count(dev);
    if (grid > resident * sms) grid = resident * sms;
    scan_fwd_v2_kernel<TPR, SOFTPLUS><<<grid, kFwdThreads, smem, stream>>>(a);
    return check_cuda(cudaGetLastError(), "scan_fwd launch");
}

template <bool SOFTPLUS>
static int by_tpr(const ScanArgs &a, const ScanPlan &pl, cudaStream_t stream) {
    switch (pl.tpr) {
        case 32: return launch_v2<32, SOFTPLUS>(a, pl.grid, stream);
        case 64: return launch_v2<64, SOFTPLUS>(a, pl.grid, stream);
        case 128: return launch_v2<128, SOFTPLUS>(a, pl.grid, stream);
        default: return launch_v2<256, SOFTPLUS>(a, pl.grid, stream);
    }
}

int scan_fwd_v2_dispatch(const ScanArgs &a, const ScanPlan &pl, cudaStream_t stream) {
    return a.softplus ? by_tpr<true>(a, pl, stream) : by_tpr<false>(a, pl, stream);
}

}  // namespace vmasr
