// Selective-scan forward, PERSISTENT variant of the multi-chunk fast path (fp32 IO, d_state 1, 16-byte aligned rows).
// Replaces selective_scan_fwd_kernel (kernels/selective_scan/csrc/selective_scan/cus/selective_scan_fwd_kernel.cuh:61-203).
//
// scan_fwd_pipe.cu launches one CTA per tile; its ncu profile shows the CTAs of a wave moving in lock-step -- all load,
// then all compute, then all wait on each other's chunk aggregates, then all store -- so HBM idles during the compute and
// exchange phases.  Here one CTA per SM slot stays resident and pulls a STREAM of items through a 4-stage TMA ring:
//
//   tile  = (chunk, batch, B/C group, CPT consecutive channels), taken round-robin: tile = blockIdx.x + visit * gridDim.x
//           (tiles are numbered chunk-major, a CTA's tiles increase, every CTA is resident: a tile only ever waits on lower
//           tiles, and the lowest unfinished tile is always being worked on -- forward progress);
//   items = per tile one B/C item (B and C segments, 16 KB) followed by CPT channel items (u and delta segments, 16 KB).
//
// 8 compute warps + 1 exchange warp, as in scan_fwd_pipe.cu; per item k:
//   compute   P1(k): channel item -> element-wise work, in-thread scan, warp scan, Y0 / Y1 written in place of u / delta
//                    (y = Y0 + Y1 * h_in), warp aggregate to shared memory; B/C item -> B, C into registers.  ARRIVE on tot[k].
//             P2(k-LAG): wait on in[k-LAG] (complete by then); y = Y0 + Y1 * h_in; 128-bit stores.
//   exchange  wait on tot[k] (every compute warp is then past P2(k-1-LAG): that ring slot is refilled);
//             publish the chunk aggregate of item k, issue its look-back loads; finish the look-back of item k-1, write the
//             states entering each warp, arrive on in[k-1].
// The ring never drains between tiles: while a tile waits for its neighbours, the next two items are already in flight.
#include <cstdlib>

#include "fast.cuh"

namespace vmasr {

constexpr int kRingStages = 4;
constexpr int kRingThreads = 288;

// incremental decoder of a CTA's item stream
struct ItemCursor {
    int visit;   // how many tiles this CTA has started
    int r;       // 0 = B/C item, 1..CPT = channel r - 1
    int chunk, b, g, d0;  // coordinates of the current tile (d0 = first scan channel)
    __device__ __forceinline__ void set_tile(const ScanArgs &a, int tile) {
        chunk = tile / a.n_rowgroups;
        const int rg = tile - chunk * a.n_rowgroups;
        const int ctile = rg % a.n_ctiles;
        const int bg = rg / a.n_ctiles;
        g = bg % a.ngroups;
        b = bg / a.ngroups;
        d0 = g * a.chan_per_group + ctile * a.chan_per_tile;
    }
    __device__ __forceinline__ void init(const ScanArgs &a) {
        visit = 0;
        r = 0;
        set_tile(a, blockIdx.x);
    }
    __device__ __forceinline__ void next(const ScanArgs &a) {
        if (++r > a.chan_per_tile) {
            r = 0;
            ++visit;
            set_tile(a, blockIdx.x + visit * gridDim.x);
        }
    }
};

template <bool SP, int LAG>
__global__ void __launch_bounds__(kRingThreads, 3) scan_fwd_ring_kernel(const __grid_constant__ ScanArgs a, const int n_tiles) {
    constexpr int NC = 256, ITEMS = 8, WPR = 8, SEG = NC * ITEMS, STAGES = kRingStages;
    constexpr int AHEAD = STAGES - 1 - LAG;  // items in flight ahead of the one being computed
    static_assert(LAG >= 1 && AHEAD >= 1, "ring too shallow");
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned long long *bar_full = reinterpret_cast<unsigned long long *>(smem);  // [STAGES] TMA completion
    unsigned long long *bar_tot = bar_full + STAGES;                              // [STAGES] 8 arrivals
    unsigned long long *bar_in = bar_tot + STAGES;                                // [STAGES] 8 arrivals (exchange-warp lanes)
    float2 *s_tot = reinterpret_cast<float2 *>(smem + 256);                       // [STAGES][8] warp totals (p, q)
    float *s_in = reinterpret_cast<float *>(smem + 512);                          // [STAGES][8] state entering each warp
    float *s_stage = reinterpret_cast<float *>(smem + 2048);                      // [STAGES][2][SEG]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int L = a.seqlen;
    const int cpt = a.chan_per_tile;
    const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // >= 1 (grid <= n_tiles)
    const int K = my_tiles * (cpt + 1);                                                     // items of this CTA

    if (threadIdx.x == NC) {
#pragma unroll
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_tot[i], WPR);
            mbar_init(&bar_in[i], WPR);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const unsigned epoch = *reinterpret_cast<volatile unsigned *>(a.ws_header + 2) % 0xfffffffeu + 1u;
    __syncthreads();

    if (warp == WPR) {
        // ================= exchange warp (also the TMA producer) =================
        const int n_groups16 = (a.n_chunks + 15) >> 4;
        ItemCursor ic;  // cursor of the item being ISSUED
        ic.init(a);
        int issued = 0;
        auto issue_next = [&]() {  // whole warp calls it, lane 0 acts
            if (issued < K) {
                if (lane == 0) {
                    const int s = issued & (STAGES - 1);
                    const int seg0 = ic.chunk * SEG;
                    const unsigned bytes = (unsigned)min(SEG, L - seg0) * 4u;
                    float *dst = s_stage + (size_t)s * 2 * SEG;
                    const float *src0, *src1;
                    if (ic.r == 0) {
                        src0 = reinterpret_cast<const float *>(a.B) + ic.b * a.B_bs + ic.g * a.B_gs + seg0;
                        src1 = reinterpret_cast<const float *>(a.C) + ic.b * a.C_bs + ic.g * a.C_gs + seg0;
                    } else {
                        const long long d = ic.d0 + ic.r - 1;
                        src0 = reinterpret_cast<const float *>(a.u) + ic.b * a.u_bs + d * a.u_ds + seg0;
                        src1 = reinterpret_cast<const float *>(a.delta) + ic.b * a.delta_bs + d * a.delta_ds + seg0;
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the slot was written through the generic proxy
                    mbar_expect_tx(&bar_full[s], 2u * bytes);
                    bulk_load(dst, src0, bytes, &bar_full[s]);
                    bulk_load(dst + SEG, src1, bytes, &bar_full[s]);
                }
                ic.next(a);
                ++issued;
            }
        };
#pragma unroll
        for (int i = 0; i <= AHEAD; ++i) issue_next();

        ItemCursor cc;  // cursor of the item being EXCHANGED
        cc.init(a);
        // pending look-back (the previous channel item)
        bool pending = false;
        int p_slot = 0, p_chunk = 0;
        long long p_seq = 0;
        CarryLook p_look;
        p_look.ptr = nullptr;
        p_look.e = make_uint4(0u, 0u, 0u, 0u);
        Aff p_cum = {1.0f, 0.0f};
        auto finish = [&]() {
            CarryEntry *l2_row = a.ws_entries2 + p_seq * n_groups16;
            const Aff before = shift_up1(p_cum, lane);
            const Aff total = {__shfl_sync(0xffffffffu, p_cum.p, WPR - 1), __shfl_sync(0xffffffffu, p_cum.q, WPR - 1)};
            bool ok;
            Aff grp = {1.0f, 0.0f};
            Aff acc = look_reduce(p_look, epoch, lane, ok, grp);
            acc = look_finish(p_look, acc, ok, l2_row, p_chunk, epoch, lane, grp);
            if (lane == 0) {
                if ((p_chunk & 15) == 15) {
                    const Aff g16 = compose(grp, total);
                    publish_entry(l2_row + (p_chunk >> 4), epoch, g16.p, g16.q);
                }
                reinterpret_cast<float2 *>(a.x)[p_seq * a.n_chunks + p_chunk] = make_float2(total.p * acc.p, fmaf(total.p, acc.q, total.q));
            }
            if (lane < WPR) {  // every writer releases its own store
                s_in[p_slot * WPR + lane] = fmaf(before.p, acc.q, before.q);
                mbar_arrive(&bar_in[p_slot]);
            }
            pending = false;
        };
        for (int k = 0; k < K; ++k) {
            const int s = k & (STAGES - 1);
            mbar_wait(&bar_tot[s], (unsigned)((k >> 2) & 1));
            if (k >= 1) issue_next();  // every compute warp is past P2(k - 1 - LAG): its slot takes item k + AHEAD
            if (cc.r == 0) {
                if (pending) finish();
                if (lane < WPR) mbar_arrive(&bar_in[s]);  // nobody waits for it; keeps the slot's phase in step with k
            } else {
                const long long seq = (long long)cc.b * a.dim + cc.d0 + cc.r - 1;
                CarryEntry *l1_row = a.ws_entries + seq * a.n_chunks;
                const float2 t = (lane < WPR) ? s_tot[s * WPR + lane] : make_float2(1.0f, 0.0f);
                const Aff cum = warp_scan_up_fast<WPR>(Aff{t.x, t.y});
                if (lane == WPR - 1) publish_entry(l1_row + cc.chunk, epoch, cum.p, cum.q);
                const CarryLook look = look_issue(l1_row, a.ws_entries2 + seq * n_groups16, cc.chunk, lane);
                if (pending) finish();
                pending = true;
                p_slot = s;
                p_chunk = cc.chunk;
                p_seq = seq;
                p_look = look;
                p_cum = cum;
            }
            cc.next(a);
        }
        if (pending) finish();
    } else {
        // ================= compute warps =================
        ItemCursor cc;
        cc.init(a);
        const int sel = (threadIdx.x >> 2) & 1;  // bank-conflict-free access order (fast.cuh)
        float2 Bl[4], Cv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            Bl[i] = f2(0.0f);
            Cv[i] = f2(0.0f);
        }
        // what P2 needs of the last LAG items: [0] = item k, [i] = item k - i
        Aff exc[LAG + 1];
        bool is_channel[LAG + 1];
        int nv[LAG + 1];
        float *outp[LAG + 1];
#pragma unroll
        for (int i = 0; i <= LAG; ++i) {
            exc[i] = Aff{1.0f, 0.0f};
            is_channel[i] = false;
            nv[i] = ITEMS;
            outp[i] = nullptr;
        }
        int nvalid = ITEMS;          // of the current tile
        float Av = 0.0f, Dv = 0.0f, bias2 = 0.0f;  // parameters of the NEXT channel item (loaded one item ahead)

        for (int k = 0; k < K + LAG; ++k) {
            const int s = k & (STAGES - 1);
#pragma unroll
            for (int i = LAG; i > 0; --i) {
                exc[i] = exc[i - 1];
                is_channel[i] = is_channel[i - 1];
                nv[i] = nv[i - 1];
                outp[i] = outp[i - 1];
            }
            is_channel[0] = false;
            if (k < K) {
                float *su = s_stage + (size_t)s * 2 * SEG + threadIdx.x * ITEMS;
                if (cc.r == 0) {
                    // ---- B/C item: this tile's B and C into registers ----
                    nvalid = max(0, min(ITEMS, L - (cc.chunk * SEG + (int)threadIdx.x * ITEMS)));
                    {   // parameters of the tile's first channel
                        const int d = cc.d0;
                        Av = __ldg(a.A + d * a.A_ds);
                        Dv = a.D ? __ldg(a.D + d) : 0.0f;
                        bias2 = (a.delta_bias ? __ldg(a.delta_bias + d) : 0.0f) * kLog2e;
                    }
                    mbar_wait(&bar_full[s], (unsigned)((k >> 2) & 1));
                    lds8_sw(su, sel, Bl);
                    lds8_sw(su + SEG, sel, Cv);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        Bl[i] = mul2(Bl[i], f2(kLn2));
                        if (nvalid != ITEMS) {  // positions past the end: keep the arithmetic finite
                            if (2 * i >= nvalid) { Bl[i].x = 0.0f; Cv[i].x = 0.0f; }
                            if (2 * i + 1 >= nvalid) { Bl[i].y = 0.0f; Cv[i].y = 0.0f; }
                        }
                    }
                    __syncwarp();
                    if (lane == 31) mbar_arrive(&bar_tot[s]);
                } else {
                    // ---- P1(k): channel item ----
                    is_channel[0] = true;
                    const int d = cc.d0 + cc.r - 1;
                    outp[0] = reinterpret_cast<float *>(a.out) + cc.b * a.out_bs + (long long)d * a.out_ds + cc.chunk * SEG + threadIdx.x * ITEMS;
                    const float Ac = Av, Dc = Dv, bc = bias2;
                    if (cc.r < cpt) {  // parameters of the next channel of this tile
                        Av = __ldg(a.A + (d + 1) * a.A_ds);
                        Dv = a.D ? __ldg(a.D + d + 1) : 0.0f;
                        bias2 = (a.delta_bias ? __ldg(a.delta_bias + d + 1) : 0.0f) * kLog2e;
                    }
                    mbar_wait(&bar_full[s], (unsigned)((k >> 2) & 1));
                    float2 uv[4], dl[4], Y0[4], Y1[4];
                    lds8_sw(su, sel, uv);
                    lds8_sw(su + SEG, sel, dl);
                    float p = 1.0f, q = 0.0f;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (nvalid != ITEMS) {
                            if (2 * i >= nvalid) { uv[i].x = 0.0f; dl[i].x = 0.0f; }
                            if (2 * i + 1 >= nvalid) { uv[i].y = 0.0f; dl[i].y = 0.0f; }
                        }
                        float2 dt2 = fma2(dl[i], f2(kLog2e), f2(bc));
                        if (SP) {
                            float2 e, sp;
                            dt2 = softplus2_pair(dt2, e, sp);
                        }
                        const float2 da = mul2(dt2, f2(Ac));
                        float2 av = make_float2(ex2_approx(da.x), ex2_approx(da.y));
                        const float2 bx = mul2(mul2(dt2, Bl[i]), uv[i]);
                        if (nvalid != ITEMS) {  // identity map past the end
                            if (2 * i >= nvalid) av.x = 1.0f;
                            if (2 * i + 1 >= nvalid) av.y = 1.0f;
                        }
                        float2 P, Q;
                        q = fmaf(av.x, q, bx.x);
                        p *= av.x;
                        P.x = p;
                        Q.x = q;
                        q = fmaf(av.y, q, bx.y);
                        p *= av.y;
                        P.y = p;
                        Q.y = q;
                        Y0[i] = fma2(Cv[i], Q, mul2(uv[i], f2(Dc)));
                        Y1[i] = mul2(Cv[i], P);
                    }
                    sts8_priv(su, sel, Y0);
                    sts8_priv(su + SEG, sel, Y1);
                    const Aff inc = warp_scan_up_fast<32>(Aff{p, q});
                    exc[0] = shift_up1(inc, lane);
                    nv[0] = nvalid;
                    if (lane == 31) s_tot[s * WPR + warp] = make_float2(inc.p, inc.q);
                    __syncwarp();
                    if (lane == 31) mbar_arrive(&bar_tot[s]);
                }
            }
            if (is_channel[LAG]) {
                // ---- P2(k - LAG) ----
                const int sp = (k - LAG) & (STAGES - 1);
                mbar_wait(&bar_in[sp], (unsigned)(((k - LAG) >> 2) & 1));
                const float h_in = fmaf(exc[LAG].p, s_in[sp * WPR + warp], exc[LAG].q);
                const float *sy = s_stage + (size_t)sp * 2 * SEG + threadIdx.x * ITEMS;
                float2 Y0[4], Y1[4], y[4];
                lds8_priv(sy, sel, Y0);
                lds8_priv(sy + SEG, sel, Y1);
#pragma unroll
                for (int i = 0; i < 4; ++i) y[i] = fma2(Y1[i], f2(h_in), Y0[i]);
                float *o = outp[LAG];
                if (nv[LAG] == ITEMS) {
                    stg8(o, y);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (2 * i < nv[LAG]) o[2 * i] = y[i].x;
                        if (2 * i + 1 < nv[LAG]) o[2 * i + 1] = y[i].y;
                    }
                }
            }
            if (k < K) cc.next(a);
        }
    }

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

template <bool SP, int LAG>
static int launch_fwd_ring(const ScanArgs &a, int n_tiles, int device, cudaStream_t stream) {
    const size_t smem = 2048 + sizeof(float) * ((size_t)kRingStages * 2 * 2048);
    static int slots_of[64] = {};  // resident CTAs per device (every CTA of the grid must be resident: they wait on each other)
    int &slots = slots_of[device >= 0 && device < 64 ? device : 0];
    if (slots == 0) {
        if (int rc = check_cuda(cudaFuncSetAttribute(scan_fwd_ring_kernel<SP, LAG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                                "scan_fwd_ring smem attribute"))
            return rc;
        int per_sm = 0;
        if (int rc = check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan_fwd_ring_kernel<SP, LAG>, kRingThreads, smem),
                                "scan_fwd_ring occupancy"))
            return rc;
        if (per_sm < 1) return fail("scan_fwd_ring: kernel does not fit on an SM");
        slots = per_sm * sm_count(device);
    }
    const int grid = n_tiles < slots ? n_tiles : slots;
    scan_fwd_ring_kernel<SP, LAG><<<grid, kRingThreads, smem, stream>>>(a, n_tiles);
    return check_cuda(cudaGetLastError(), "scan_fwd_ring launch");
}

// n_chunks > 1, chan_per_group a multiple of chan_per_tile (scan_host.cu checks)
int scan_fwd_ring_dispatch(const ScanArgs &a, const ScanPlan &pl, int device, cudaStream_t stream) {
    // how many items the outputs trail the local scans (tuning knob: VMASR_FWD_LAG = 1 | 2)
    static const int lag = [] { const char *e = getenv("VMASR_FWD_LAG"); return e ? atoi(e) : 2; }();
    if (lag == 1) return a.softplus ? launch_fwd_ring<true, 1>(a, pl.grid, device, stream) : launch_fwd_ring<false, 1>(a, pl.grid, device, stream);
    return a.softplus ? launch_fwd_ring<true, 2>(a, pl.grid, device, stream) : launch_fwd_ring<false, 2>(a, pl.grid, device, stream);
}

}  // namespace vmasr
