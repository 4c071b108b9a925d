// Selective-scan forward, fast path for sm_100a: fp32 IO, d_state 1, 16-byte aligned rows.
// Replaces selective_scan_fwd_kernel (kernels/selective_scan/csrc/selective_scan/cus/selective_scan_fwd_kernel.cuh:61-203)
// for the shapes VM-ASR runs.  Same maths and tile decomposition as scan_fwd.cu (see scan.cuh); what differs:
//   * u / delta row segments come in by TMA bulk copies (cp.async.bulk -> SASS UBLKCP) into a per-row-segment
//     shared-memory ring with mbarrier completion; the copy for channel c + STAGES is issued as soon as channel c
//     has been read, so HBM latency hides behind whole iterations of compute without spending registers on it;
//   * the B / C segment shared by all channels of the tile is copied once per tile the same way;
//   * all element-wise arithmetic runs on position pairs with packed fp32x2 instructions (fast.cuh);
//   * the row segments of a CTA (ROWS = 256 / TPR of them when the sequence is short) are independent pipelines:
//     own mbarriers, own named barrier, own producer lane -- they never wait on each other;
//   * the cross-warp combination and the cross-chunk look-back are done by the FIRST WARP of the row segment only,
//     which hands every warp the state entering it (one float) through shared memory; the other warps spend no
//     instructions on it;
//   * full tiles carry no bounds checks; tiles are taken in blockIdx order (chunk-major), so a tile only ever waits
//     on carries of tiles that were dispatched before it.
// Outputs go straight from registers to HBM with 128-bit stores.
#include "fast.cuh"

namespace vmasr {

constexpr int kMaxTileChannels = 64;

// REV (time runs against memory order) is supported for single-chunk sequences, which is all the host sends here.
template <int TPR, bool TAIL, bool SP, int STAGES, bool REV>
__device__ __forceinline__ void scan_fwd_tma_body(const ScanArgs &a, const TileMaps &tm, unsigned char *smem, const int chunk, const int rg,
                                                  const int ztile) {
    constexpr int NT = 256, ITEMS = 8;
    constexpr int ROWS = NT / TPR;
    constexpr int WPR = TPR / 32;
    constexpr int SEG = TPR * ITEMS;  // positions per row segment (== chunk when n_chunks > 1)

    // shared memory carve-up (header 2048 bytes)
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem);  // [STAGES][ROWS] stage barriers, then B/C
    float2 *s_tot = reinterpret_cast<float2 *>(smem + 384);                   // [2][8] warp totals (p, q)
    float *s_in = reinterpret_cast<float *>(smem + 512);                      // [8] state entering each warp
    float *s_par = reinterpret_cast<float *>(smem + 1024);                    // [3][kMaxTileChannels]
    float *s_stage = reinterpret_cast<float *>(smem + 2048);                  // [STAGES][ROWS][2][SEG]  u, delta
    float *s_bc = s_stage + (size_t)(STAGES - 1) * ROWS * 2 * SEG;            // [2][SEG] B, C: borrowed from the last stage
    unsigned long long *bar_bc = bars + STAGES * ROWS;

    const int ctile = rg % a.n_ctiles;
    const int bg = rg / a.n_ctiles;
    const int g = bg % a.ngroups;
    const int b = bg / a.ngroups;

    const int row = threadIdx.x / TPR;
    const int t_in_row = threadIdx.x - row * TPR;
    const int warp_in_row = t_in_row >> 5;
    const int warp_slot = threadIdx.x >> 5;  // row * WPR + warp_in_row
    const int lane = threadIdx.x & 31;
    const int tseg = REV ? TPR - 1 - t_in_row : t_in_row;  // this thread's 8-position segment of the row segment (memory order)
    const int sel = swz_half(tseg), slot = swz_slot(tseg) * ITEMS;  // where the 64-byte swizzle puts this thread's 32 bytes (scan.cuh)
    const int L = a.seqlen;
    const int seg0 = chunk * SEG;                       // first position of the tile
    const int pos = seg0 + tseg * ITEMS;
    const bool accum = a.accum == 1, addm = a.accum == 2;  // red.add / load-add-store (scan.cuh)
    const int line0 = seg0 / kTileLine;                 // first line of the tile in the tensor maps
    constexpr unsigned seg_bytes = SEG * 4u;            // a box always counts in full (lines past the end arrive as zeros)
    int nvalid = ITEMS;
    if (TAIL) nvalid = max(0, min(ITEMS, L - pos));

    const int c_begin = ctile * a.chan_per_tile;
    const int n_chan = min(a.chan_per_group, c_begin + a.chan_per_tile) - c_begin;
    const int n_iter = (n_chan - row + ROWS - 1) / ROWS;  // iterations of THIS row segment (may be 0)
    const int d0 = g * a.chan_per_group + c_begin;        // first scan channel of the tile

    float *out_ptr = reinterpret_cast<float *>(a.out) + b * a.out_bs + (long long)(d0 + row) * a.out_ds + pos;
    const long long out_step = (long long)ROWS * a.out_ds;

    unsigned epoch = 0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < STAGES * ROWS + 1; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_wait();  // every thread, in front of its first access to global memory
    zero_side_region(a, ztile);
    if (a.n_chunks > 1) epoch = *reinterpret_cast<volatile unsigned *>(a.ws_header + 2) % 0xfffffffeu + 1u;
    // per-channel parameters of the tile
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

    float *my_stage = s_stage + (size_t)row * 2 * SEG;  // + stage * ROWS * 2 * SEG
    unsigned long long *my_bars = bars + row;           // + stage * ROWS
    auto issue_stage = [&](int it) {  // producer lane of the row segment: bulk copies of iteration `it`
        const int s = it % STAGES;
        unsigned long long *bar = my_bars + s * ROWS;
        float *dst = my_stage + (size_t)s * ROWS * 2 * SEG;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy reads of the slot are ordered before the refill
        mbar_expect_tx(bar, 2u * seg_bytes);
        tensor_load(dst, &tm.u, line0, d0 + row + it * ROWS, b, bar);
        tensor_load(dst + SEG, &tm.delta, line0, d0 + row + it * ROWS, b, bar);
    };
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar_bc, 2u * seg_bytes);
        tensor_load(s_bc, &tm.B, line0, g, b, bar_bc);
        tensor_load(s_bc + SEG, &tm.C, line0, g, b, bar_bc);
    }
    if (t_in_row == 0) {
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s)
            if (s < n_iter) issue_stage(s);
    }
    if ((int)threadIdx.x < 3 * n_chan) {
        const int which = threadIdx.x / n_chan, cc = threadIdx.x - which * n_chan;
        s_par[which * kMaxTileChannels + cc] = par_v;
    }
    __syncthreads();

    float2 Bl[4], Cv[4];  // ln2 * B (the scan runs on dt in the log2 domain) and C of this thread's positions
    mbar_wait(bar_bc, 0);
    lds8_priv(s_bc + slot, sel, Bl);
    lds8_priv(s_bc + SEG + slot, sel, Cv);
    __syncthreads();  // B / C are in registers: the last stage is free for data now
    if (t_in_row == 0 && STAGES - 1 < n_iter) issue_stage(STAGES - 1);
#pragma unroll
    for (int j = 0; j < 4; ++j) Bl[j] = mul2(Bl[j], f2(kLn2));

    const int n_groups16 = (a.n_chunks + 15) >> 4;
    const bool multi = a.n_chunks > 1;  // then ROWS == 1
    const bool leader = warp_in_row == 0;
    for (int it = 0; it < n_iter; ++it) {
        const int s = it % STAGES;
        const int cc = it * ROWS + row;  // channel index inside the tile
        const float Av = s_par[cc];
        const float Dv = s_par[kMaxTileChannels + cc];
        const float bias2 = s_par[2 * kMaxTileChannels + cc];
        const long long seq = (long long)b * a.dim + d0 + cc;

        // look-back loads first (leader warp): they fly while the row is computed
        CarryLook look;
        const CarryEntry *l2_row = nullptr;
        if (multi && leader) {
            l2_row = a.ws_entries2 + seq * n_groups16;
            look = look_issue(a.ws_entries + seq * a.n_chunks, l2_row, chunk, lane);
        }
        mbar_wait(my_bars + s * ROWS, (unsigned)((it / STAGES) & 1));
        float2 uv[4], av[4], bx[4];
        {
            const float *su = my_stage + (size_t)s * ROWS * 2 * SEG + slot;
            float2 dl[4];
            lds8_priv(su, sel, uv);
            lds8_priv(su + SEG, sel, dl);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 dt2 = fma2(dl[j], f2(kLog2e), f2(bias2));
                if (SP) {
                    float2 e, sp;
                    dt2 = softplus2_pair(dt2, e, sp);
                }
                const float2 da = mul2(dt2, f2(Av));
                av[j] = make_float2(ex2_approx(da.x), ex2_approx(da.y));
                bx[j] = mul2(mul2(dt2, Bl[j]), uv[j]);
            }
        }
        if (TAIL) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (2 * j >= nvalid) { av[j].x = 1.0f; bx[j].x = 0.0f; }
                if (2 * j + 1 >= nvalid) { av[j].y = 1.0f; bx[j].y = 0.0f; }
            }
        }
        Aff loc = {1.0f, 0.0f};
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = pair_at<REV>(jj);  // pairs in time order
            float2 P, Q;
            walk_pair<REV>(av[j], bx[j], loc.p, loc.q, P, Q);
        }
        const Aff inc = warp_scan_up_fast<32>(loc);
        const Aff exc = shift_up1(inc, lane);

        float h_warp;  // state entering this warp's first position
        if (WPR > 1) {
            float2 *tot = s_tot + (it & 1) * 8;
            if (lane == 31) tot[warp_slot] = make_float2(inc.p, inc.q);
            row_barrier(1 + row, TPR);  // stage s consumed by the whole row segment; warp totals visible
            if (t_in_row == 0 && it + STAGES < n_iter) issue_stage(it + STAGES);
            if (multi) {
                if (leader) {
                    const float2 t = (lane < WPR) ? tot[row * WPR + lane] : make_float2(1.0f, 0.0f);
                    const Aff cum = warp_scan_up_fast<WPR>(Aff{t.x, t.y});
                    const Aff before = shift_up1(cum, lane);
                    const Aff total = {__shfl_sync(0xffffffffu, cum.p, WPR - 1), __shfl_sync(0xffffffffu, cum.q, WPR - 1)};
                    if (lane == 0) publish_entry(a.ws_entries + seq * a.n_chunks + chunk, epoch, total.p, total.q);
                    bool ok;
                    Aff grp = {1.0f, 0.0f};
                    Aff acc = look_reduce(look, epoch, lane, ok, grp);
                    acc = look_finish(look, acc, ok, l2_row, chunk, epoch, lane, grp);
                    if (lane == 0) {
                        if ((chunk & 15) == 15) {
                            const Aff g16 = compose(grp, total);
                            publish_entry(a.ws_entries2 + seq * n_groups16 + (chunk >> 4), epoch, g16.p, g16.q);
                        }
                        reinterpret_cast<float2 *>(a.x)[seq * a.n_chunks + chunk] =
                            make_float2(total.p * acc.p, fmaf(total.p, acc.q, total.q));
                    }
                    if (lane < WPR) s_in[row * WPR + lane] = fmaf(before.p, acc.q, before.q);
                }
                row_barrier(1 + row, TPR);
                h_warp = s_in[warp_slot];
            } else {
                // single chunk: the state entering the sequence is 0; fold the totals of the warps before this one
                float hq = 0.0f, hp = 1.0f;
#pragma unroll
                for (int w = 0; w < WPR - 1; ++w) {
                    const float2 t = tot[row * WPR + w];
                    if (w < warp_in_row) {
                        hq = fmaf(t.x, hq, t.y);
                        hp *= t.x;
                    }
                }
                h_warp = hq;
                if (warp_in_row == WPR - 1 && lane == 31)
                    reinterpret_cast<float2 *>(a.x)[seq] = make_float2(hp * inc.p, fmaf(inc.p, hq, inc.q));
            }
        } else {
            __syncwarp();
            if (lane == 0 && it + STAGES < n_iter) issue_stage(it + STAGES);
            h_warp = 0.0f;
            if (lane == 31) reinterpret_cast<float2 *>(a.x)[seq] = make_float2(inc.p, inc.q);
        }

        float h = fmaf(exc.p, h_warp, exc.q);
        float2 y[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = pair_at<REV>(jj);
            float2 hh;
            walk_state<REV>(av[j], bx[j], h, hh);
            y[j] = fma2(Cv[j], hh, mul2(uv[j], f2(Dv)));
        }
        float *o = out_ptr + it * out_step;
        if (!TAIL || nvalid == ITEMS) {
            if (accum) {
                red8(o, y);
            } else {
                if (addm) {
                    float2 old[4];
                    ldg8(o, old);
#pragma unroll
                    for (int j = 0; j < 4; ++j) y[j] = add2(old[j], y[j]);
                }
                stg8(o, y);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (accum || addm) {
                    if (2 * j < nvalid) atomicAdd(o + 2 * j, y[j].x);
                    if (2 * j + 1 < nvalid) atomicAdd(o + 2 * j + 1, y[j].y);
                } else {
                    if (2 * j < nvalid) o[2 * j] = y[j].x;
                    if (2 * j + 1 < nvalid) o[2 * j + 1] = y[j].y;
                }
            }
        }
    }
    if (multi) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned prev = atomicAdd(a.ws_header + 1, 1u);
            if (prev == (unsigned)(a.n_chunks * a.n_rowgroups) - 1u) {
                a.ws_header[1] = 0u;
                a.ws_header[2] = a.ws_header[2] + 1u;
                __threadfence();
            }
        }
    }
}

template <int TPR, bool SP, int STAGES>
__global__ void __launch_bounds__(256, 3) scan_fwd_tma_kernel(const __grid_constant__ GroupArgs ga) {
    extern __shared__ __align__(1024) unsigned char smem_fwd_tma[];  // swizzled tiles need 512-byte aligned slots
    pdl_launch_dependents();  // the next kernel on the stream may be scheduled while this one drains ...
    constexpr int SEG = TPR * 8;
    int tile;
    const int prob = group_problem(ga, tile);
    const ScanArgs &a = ga.a[prob];
    const TileMaps &tm = ga.tm[prob];
    if (a.pdl_mode & 1) pdl_wait();
    const int ztile = tile;   // ... and this one touches global memory only after its predecessor has completed (wait in the body,
                              // behind the index arithmetic and the mbarrier initialisation)
    const int chunk = tile / a.n_rowgroups;
    const int rg = tile - chunk * a.n_rowgroups;
    const bool tail = (chunk + 1) * SEG > a.seqlen;
    if (a.rev) {  // single chunk only (host-checked)
        if (tail) scan_fwd_tma_body<TPR, true, SP, STAGES, true>(a, tm, smem_fwd_tma, chunk, rg, ztile);
        else scan_fwd_tma_body<TPR, false, SP, STAGES, true>(a, tm, smem_fwd_tma, chunk, rg, ztile);
    } else {
        if (tail) scan_fwd_tma_body<TPR, true, SP, STAGES, false>(a, tm, smem_fwd_tma, chunk, rg, ztile);
        else scan_fwd_tma_body<TPR, false, SP, STAGES, false>(a, tm, smem_fwd_tma, chunk, rg, ztile);
    }
}

static size_t scan_fwd_tma_smem(int stages) { return 2048 + sizeof(float) * ((size_t)stages * 2 * 2048); }

template <int TPR, bool SP, int STAGES>
static int launch_tma(const GroupArgs &ga, int grid, cudaStream_t stream) {
    const size_t smem = scan_fwd_tma_smem(STAGES);
    static PerDeviceOnce configured;  // the attribute is per function and per device
    if (!configured()) {
        if (int rc = check_cuda(cudaFuncSetAttribute(scan_fwd_tma_kernel<TPR, SP, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                                "scan_fwd_tma smem attribute"))
            return rc;
        configured() = true;
    }
    return launch_pdl(scan_fwd_tma_kernel<TPR, SP, STAGES>, grid, 256, smem, stream, "scan_fwd_tma launch", ga);
}

template <bool SP, int STAGES>
static int dispatch_tpr(const GroupArgs &ga, int tpr, int grid, cudaStream_t stream) {
    switch (tpr) {
        case 32: return launch_tma<32, SP, STAGES>(ga, grid, stream);
        case 64: return launch_tma<64, SP, STAGES>(ga, grid, stream);
        case 128: return launch_tma<128, SP, STAGES>(ga, grid, stream);
        default: return launch_tma<256, SP, STAGES>(ga, grid, stream);
    }
}

// every problem of the group: same threads-per-row and softplus flag (scan_host.cu groups them so); TMA ring of 4 stages
// (bytes in flight per CTA = 4 x 16 KB; 2 and 3 were measured slower)
int scan_fwd_tma_dispatch(const GroupArgs &ga, int tpr, int grid, cudaStream_t stream) {
    for (int i = 0; i < ga.n; ++i)
        if (ga.a[i].rev && ga.a[i].n_chunks > 1) return fail("scan_fwd_tma: reversed scans of more than one chunk go to the multi-chunk kernel");
    return ga.a[0].softplus ? dispatch_tpr<true, 4>(ga, tpr, grid, stream) : dispatch_tpr<false, 4>(ga, tpr, grid, stream);
}

}  // namespace vmasr
