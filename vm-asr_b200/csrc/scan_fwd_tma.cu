// Selective-scan forward, fast path for sm_100a: fp32 IO, d_state 1, 16-byte aligned rows.
// Same maths and same tile decomposition as scan_fwd.cu (see scan.cuh); what changes is how data moves:
//   * u / delta row segments of the tile's channels are brought in by TMA bulk copies (cp.async.bulk ->
//     SASS UBLKCP) into a 2-stage shared-memory ring, completion signalled on mbarriers; the copy for
//     channel c+2 is issued as soon as channel c has been read, so HBM latency is hidden behind a whole
//     iteration of compute without spending registers on prefetch;
//   * the B / C segment shared by all channels of the tile is copied once per tile the same way;
//   * the per-channel parameters (A, D, delta_bias) of the tile are staged once in shared memory;
//   * full tiles carry no bounds checks; tiles are taken in blockIdx order (chunk-major), so a tile only
//     ever waits on carries of tiles that were dispatched before it.
// Outputs go straight from registers to HBM with 128-bit stores.
#include "pipe.cuh"

namespace vmasr {

constexpr int kMaxTileChannels = 64;

template <int TPR, bool TAIL>
__device__ __forceinline__ void scan_fwd_tma_body(const ScanArgs &a, unsigned char *smem) {
    constexpr int NT = 256, ITEMS = 8;
    constexpr int ROWS = NT / TPR;
    constexpr int WPR = TPR / 32;
    constexpr int SEG = TPR * ITEMS;  // positions per row segment (== chunk when n_chunks > 1)

    // shared memory carve-up
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem);              // [0,1] stages, [2] B/C
    float2 *s_tot = reinterpret_cast<float2 *>(smem + 32);                                // [2][ROWS][WPR]
    float *s_par = reinterpret_cast<float *>(smem + 256);                                 // [3][kMaxTileChannels]
    float *s_bc = reinterpret_cast<float *>(smem + 2048);                                 // [2][SEG]   B, C
    float *s_stage = s_bc + 2 * SEG;                                                      // [2][2][ROWS*SEG] u, delta

    const int tile = blockIdx.x;
    const int chunk = tile / a.n_rowgroups;
    const int rg = tile - chunk * a.n_rowgroups;
    const int ctile = rg % a.n_ctiles;
    const int bg = rg / a.n_ctiles;
    const int g = bg % a.ngroups;
    const int b = bg / a.ngroups;

    const int row = threadIdx.x / TPR;
    const int t_in_row = threadIdx.x - row * TPR;
    const int warp_in_row = t_in_row >> 5;
    const int lane = threadIdx.x & 31;
    const int L = a.seqlen;
    const int seg0 = chunk * SEG;                       // first position of the tile
    const int pos = seg0 + t_in_row * ITEMS;
    const int seg_len = min(SEG, L - seg0);             // valid positions in this tile (multiple of 4)
    const unsigned seg_bytes = (unsigned)seg_len * 4u;
    const bool last_warp = (warp_in_row == WPR - 1);
    int nvalid = ITEMS;
    if (TAIL) nvalid = max(0, min(ITEMS, L - pos));

    const int c_begin = ctile * a.chan_per_tile;
    const int c_end = min(a.chan_per_group, c_begin + a.chan_per_tile);
    const int n_chan = c_end - c_begin;
    const int n_iter = (n_chan + ROWS - 1) / ROWS;
    const int d0 = g * a.chan_per_group + c_begin;      // first scan channel of the tile

    const float *u_base = reinterpret_cast<const float *>(a.u) + b * a.u_bs + (long long)d0 * a.u_ds + seg0;
    const float *dl_base = reinterpret_cast<const float *>(a.delta) + b * a.delta_bs + (long long)d0 * a.delta_ds + seg0;
    float *out_ptr = reinterpret_cast<float *>(a.out) + b * a.out_bs + (long long)(d0 + row) * a.out_ds + pos;

    unsigned epoch = 0;
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (a.n_chunks > 1) epoch = *reinterpret_cast<volatile unsigned *>(a.ws_header + 2) % 0xfffffffeu + 1u;
    // per-channel parameters of the tile
    for (int i = threadIdx.x; i < 3 * n_chan; i += NT) {
        const int which = i / n_chan, cc = i - which * n_chan;
        const int d = d0 + cc;
        float v;
        if (which == 0) v = __ldg(a.A + d * a.A_ds) * kLog2e;
        else if (which == 1) v = a.D ? __ldg(a.D + d) : 0.0f;
        else v = a.delta_bias ? __ldg(a.delta_bias + d) : 0.0f;
        s_par[which * kMaxTileChannels + cc] = v;
    }
    __syncthreads();

    auto issue_stage = [&](int it) {
        // thread 0: bulk copies of iteration `it` into stage it & 1
        const int s = it & 1;
        const int rows_here = min(ROWS, n_chan - it * ROWS);
        mbar_expect_tx(&bars[s], 2u * seg_bytes * (unsigned)rows_here);
        float *dst_u = s_stage + (size_t)s * 2 * ROWS * SEG;
        float *dst_d = dst_u + ROWS * SEG;
        for (int r = 0; r < rows_here; ++r) {
            const long long ch = (long long)(it * ROWS + r);
            bulk_load(dst_u + r * SEG, u_base + ch * a.u_ds, seg_bytes, &bars[s]);
            bulk_load(dst_d + r * SEG, dl_base + ch * a.delta_ds, seg_bytes, &bars[s]);
        }
    };
    if (threadIdx.x == 0) {
        const float *Bg = reinterpret_cast<const float *>(a.B) + b * a.B_bs + g * a.B_gs + seg0;
        const float *Cg = reinterpret_cast<const float *>(a.C) + b * a.C_bs + g * a.C_gs + seg0;
        mbar_expect_tx(&bars[2], 2u * seg_bytes);
        bulk_load(s_bc, Bg, seg_bytes, &bars[2]);
        bulk_load(s_bc + SEG, Cg, seg_bytes, &bars[2]);
        issue_stage(0);
        if (n_iter > 1) issue_stage(1);
    }

    float Bv[ITEMS], Cv[ITEMS];
    mbar_wait(&bars[2], 0);
    {
        const float4 *pb = reinterpret_cast<const float4 *>(s_bc + t_in_row * ITEMS);
        const float4 *pc = reinterpret_cast<const float4 *>(s_bc + SEG + t_in_row * ITEMS);
        const float4 b0 = pb[0], b1 = pb[1], c0 = pc[0], c1 = pc[1];
        Bv[0] = b0.x; Bv[1] = b0.y; Bv[2] = b0.z; Bv[3] = b0.w; Bv[4] = b1.x; Bv[5] = b1.y; Bv[6] = b1.z; Bv[7] = b1.w;
        Cv[0] = c0.x; Cv[1] = c0.y; Cv[2] = c0.z; Cv[3] = c0.w; Cv[4] = c1.x; Cv[5] = c1.y; Cv[6] = c1.z; Cv[7] = c1.w;
    }
    if (TAIL) {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i)
            if (i >= nvalid) { Bv[i] = 0.0f; Cv[i] = 0.0f; }
    }

    const long long entry_stride = a.n_chunks;  // level-1 entries per (b, d)
    const int n_groups16 = (a.n_chunks + 15) >> 4;
    for (int it = 0; it < n_iter; ++it) {
        const int s = it & 1;
        const int cc = it * ROWS + row;          // channel index inside the tile
        const bool active = cc < n_chan;
        const int ccl = active ? cc : 0;
        const float A2 = s_par[ccl];
        const float Dv = s_par[kMaxTileChannels + ccl];
        const float bias = s_par[2 * kMaxTileChannels + ccl];
        const float bias2 = bias * kLog2e;

        // look-back loads first: they fly while the row is computed
        const long long seq = (long long)b * a.dim + d0 + ccl;
        CarryLook look;
        const CarryEntry *l2_row = nullptr;
        if (a.n_chunks > 1) {
            l2_row = a.ws_entries2 + seq * n_groups16;
            look = look_issue(a.ws_entries + seq * entry_stride, l2_row, chunk, lane);
        }
        mbar_wait(&bars[s], (unsigned)((it >> 1) & 1));
        float uv[ITEMS], dt[ITEMS];
        {
            const float *su = s_stage + (size_t)s * 2 * ROWS * SEG + (active ? row : 0) * SEG + t_in_row * ITEMS;
            const float4 *pu = reinterpret_cast<const float4 *>(su);
            const float4 *pd = reinterpret_cast<const float4 *>(su + ROWS * SEG);
            const float4 u0 = pu[0], u1 = pu[1], e0 = pd[0], e1 = pd[1];
            uv[0] = u0.x; uv[1] = u0.y; uv[2] = u0.z; uv[3] = u0.w; uv[4] = u1.x; uv[5] = u1.y; uv[6] = u1.z; uv[7] = u1.w;
            dt[0] = e0.x; dt[1] = e0.y; dt[2] = e0.z; dt[3] = e0.w; dt[4] = e1.x; dt[5] = e1.y; dt[6] = e1.z; dt[7] = e1.w;
        }
        float av[ITEMS], bx[ITEMS];
        Aff loc = {1.0f, 0.0f};
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const float x = dt[i] + bias;
            const float d = a.softplus ? softplus2(fmaf(dt[i], kLog2e, bias2), x) : x;
            av[i] = ex2_approx(d * A2);
            bx[i] = d * uv[i] * Bv[i];
            if (TAIL && i >= nvalid) { av[i] = 1.0f; bx[i] = 0.0f; }
            loc.q = fmaf(av[i], loc.q, bx[i]);
            loc.p *= av[i];
        }
        Aff inc = warp_scan_up(loc, lane);
        Aff exc = {__shfl_up_sync(0xffffffffu, inc.p, 1), __shfl_up_sync(0xffffffffu, inc.q, 1)};
        if (lane == 0) exc = {1.0f, 0.0f};
        Aff total;
        if (WPR > 1) {
            if (lane == 31) s_tot[(s * ROWS + row) * WPR + warp_in_row] = make_float2(inc.p, inc.q);
        }
        __syncthreads();  // every thread has consumed stage s (and the warp totals are visible)
        if (threadIdx.x == 0 && it + 2 < n_iter) issue_stage(it + 2);
        if (WPR > 1) {
            Aff before = {1.0f, 0.0f};
            total = {1.0f, 0.0f};
#pragma unroll
            for (int w = 0; w < WPR; ++w) {
                const float2 t = s_tot[(s * ROWS + row) * WPR + w];
                if (w == warp_in_row) before = total;
                total = compose(total, Aff{t.x, t.y});
            }
            exc = compose(before, exc);
        } else {
            total = {__shfl_sync(0xffffffffu, inc.p, 31), __shfl_sync(0xffffffffu, inc.q, 31)};
        }

        float h_in = 0.0f, pcum_in = 1.0f;
        if (a.n_chunks > 1) {  // one row per CTA in this case
            if (threadIdx.x == 0) publish_entry(a.ws_entries + seq * entry_stride + chunk, epoch, total.p, total.q);
            bool ok;
            Aff grp = {1.0f, 0.0f};
            Aff acc = look_reduce(look, epoch, lane, ok, grp);
            acc = look_finish(look, acc, ok, l2_row, chunk, epoch, lane, grp);
            if (threadIdx.x == 0 && (chunk & 15) == 15) {
                const Aff g16 = compose(grp, total);
                publish_entry(a.ws_entries2 + seq * n_groups16 + (chunk >> 4), epoch, g16.p, g16.q);
            }
            h_in = acc.q;
            pcum_in = acc.p;
        }
        if (last_warp && lane == 0 && active)
            reinterpret_cast<float2 *>(a.x)[seq * entry_stride + chunk] = make_float2(total.p * pcum_in, fmaf(total.p, h_in, total.q));

        float h = fmaf(exc.p, h_in, exc.q);
        float y[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            h = fmaf(av[i], h, bx[i]);
            y[i] = fmaf(Cv[i], h, Dv * uv[i]);
        }
        if (active) {
            float *o = out_ptr + (long long)(it * ROWS) * a.out_ds;
            if (!TAIL || nvalid == ITEMS) {
                reinterpret_cast<float4 *>(o)[0] = make_float4(y[0], y[1], y[2], y[3]);
                reinterpret_cast<float4 *>(o)[1] = make_float4(y[4], y[5], y[6], y[7]);
            } else {
#pragma unroll
                for (int i = 0; i < ITEMS; ++i)
                    if (i < nvalid) o[i] = y[i];
            }
        }
    }
    if (a.n_chunks > 1) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned prev = atomicAdd(a.ws_header + 1, 1u);
            if (prev == gridDim.x - 1) {
                a.ws_header[1] = 0u;
                a.ws_header[2] = a.ws_header[2] + 1u;
                __threadfence();
            }
        }
    }
}

template <int TPR>
__global__ void __launch_bounds__(256) scan_fwd_tma_kernel(const __grid_constant__ ScanArgs a) {
    extern __shared__ __align__(128) unsigned char smem_fwd_tma[];
    constexpr int SEG = TPR * 8;
    const int chunk = blockIdx.x / a.n_rowgroups;
    const bool tail = (chunk + 1) * SEG > a.seqlen;
    if (tail) scan_fwd_tma_body<TPR, true>(a, smem_fwd_tma);
    else scan_fwd_tma_body<TPR, false>(a, smem_fwd_tma);
}

size_t scan_fwd_tma_smem(int tpr) {
    const size_t seg = (size_t)tpr * 8, rows = 256 / tpr;
    return 2048 + sizeof(float) * (2 * seg + 2 * 2 * rows * seg);
}

template <int TPR>
static int launch_tma(const ScanArgs &a, int grid, cudaStream_t stream) {
    const size_t smem = scan_fwd_tma_smem(TPR);
    static bool configured = false;  // attribute is per function; setting it repeatedly is harmless
    if (!configured) {
        if (int rc = check_cuda(cudaFuncSetAttribute(scan_fwd_tma_kernel<TPR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                                "scan_fwd_tma smem attribute"))
            return rc;
        configured = true;
    }
    scan_fwd_tma_kernel<TPR><<<grid, 256, smem, stream>>>(a);
    return check_cuda(cudaGetLastError(), "scan_fwd_tma launch");
}

int scan_fwd_tma_dispatch(const ScanArgs &a, const ScanPlan &pl, cudaStream_t stream) {
    switch (pl.tpr) {
        case 32: return launch_tma<32>(a, pl.grid, stream);
        case 64: return launch_tma<64>(a, pl.grid, stream);
        case 128: return launch_tma<128>(a, pl.grid, stream);
        default: return launch_tma<256>(a, pl.grid, stream);
    }
}

}  // namespace vmasr
