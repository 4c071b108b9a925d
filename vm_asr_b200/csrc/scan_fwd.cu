// Selective-scan forward for sm_100a.
// Replaces selective_scan_fwd_kernel (kernels/selective_scan/csrc/selective_scan/cus/selective_scan_fwd_kernel.cuh:61-172):
//   dt = softplus(delta + delta_bias);  h_l = exp(dt*A) h_{l-1} + dt*B_l*u_l;  out_l = sum_n C_l h_l + D u_l
// and writes the per-chunk (cumulative decay, end state) tensor `x` the backward needs.
#include "pipe.cuh"

namespace vmasr {

template <typename T, int NT, int TPR, int ITEMS, bool N1, bool VEC>
__global__ void __launch_bounds__(NT) scan_fwd_kernel(const __grid_constant__ ScanArgs a) {
    constexpr int ROWS = NT / TPR;
    constexpr int WPR = TPR / 32;  // warps per row segment
    constexpr int CHUNK = TPR * ITEMS;
    __shared__ float2 s_tot[2][ROWS][WPR > 1 ? WPR : 1];
    __shared__ unsigned s_tile[2];

    unsigned tile, epoch;
    claim_tile(a, s_tile, tile, epoch);

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
    const int pos = chunk * CHUNK + t_in_row * ITEMS;
    const bool last_warp = (warp_in_row == WPR - 1);

    const int c_begin = ctile * a.chan_per_tile;
    const int c_end = min(a.chan_per_group, c_begin + a.chan_per_tile);
    const int n_iter = (c_end - c_begin + ROWS - 1) / ROWS;

    const T *Bg = reinterpret_cast<const T *>(a.B) + b * a.B_bs + g * a.B_gs;
    const T *Cg = reinterpret_cast<const T *>(a.C) + b * a.C_bs + g * a.C_gs;

    float Bv[ITEMS], Cv[ITEMS];
    if (N1) {
        load_items<T, ITEMS, VEC>(Bg, pos, L, Bv, 0.0f);
        load_items<T, ITEMS, VEC>(Cg, pos, L, Cv, 0.0f);
    }

    int buf = 0;
    for (int it = 0; it < n_iter; ++it) {
        const int c = c_begin + it * ROWS + row;
        const bool active = c < c_end;
        const int d = g * a.chan_per_group + (active ? c : c_begin);
        const T *u_row = reinterpret_cast<const T *>(a.u) + b * a.u_bs + d * a.u_ds;
        const T *dl_row = reinterpret_cast<const T *>(a.delta) + b * a.delta_bs + d * a.delta_ds;

        float uv[ITEMS], dt[ITEMS], y[ITEMS];
        load_items<T, ITEMS, VEC>(u_row, pos, L, uv, 0.0f);
        load_items<T, ITEMS, VEC>(dl_row, pos, L, dt, 0.0f);
        const float bias = a.delta_bias ? __ldg(a.delta_bias + d) : 0.0f;
        const float Dv = a.D ? __ldg(a.D + d) : 0.0f;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            float x = dt[i] + bias;
            float unused;
            dt[i] = a.softplus ? softplus_sig<false>(x, unused) : x;
            y[i] = Dv * uv[i];
        }

        for (int n = 0; n < (N1 ? 1 : a.dstate); ++n) {
            if (!N1) {
                load_items<T, ITEMS, VEC>(Bg + n * a.B_ns, pos, L, Bv, 0.0f);
                load_items<T, ITEMS, VEC>(Cg + n * a.C_ns, pos, L, Cv, 0.0f);
            }
            const float A2 = __ldg(a.A + d * a.A_ds + n * a.A_ns) * kLog2e;
            float av[ITEMS], bx[ITEMS];
            Aff loc = {1.0f, 0.0f};
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const bool valid = pos + i < L;
                av[i] = valid ? ex2_approx(dt[i] * A2) : 1.0f;
                bx[i] = valid ? dt[i] * uv[i] * Bv[i] : 0.0f;
                loc.q = fmaf(av[i], loc.q, bx[i]);
                loc.p *= av[i];
            }
            // exclusive prefix of this thread inside the chunk
            Aff inc = warp_scan_up(loc, lane);
            Aff exc = {__shfl_up_sync(0xffffffffu, inc.p, 1), __shfl_up_sync(0xffffffffu, inc.q, 1)};
            if (lane == 0) exc = {1.0f, 0.0f};
            Aff total = inc;  // valid in lane 31 when WPR == 1
            if (WPR > 1) {
                if (lane == 31) s_tot[buf][row][warp_in_row] = make_float2(inc.p, inc.q);
                __syncthreads();
                Aff before = {1.0f, 0.0f};
                total = {1.0f, 0.0f};
#pragma unroll
                for (int w = 0; w < WPR; ++w) {
                    const float2 t = s_tot[buf][row][w];
                    if (w == warp_in_row) before = total;
                    total = compose(total, Aff{t.x, t.y});
                }
                exc = compose(before, exc);
                buf ^= 1;
            } else {
                total = {__shfl_sync(0xffffffffu, inc.p, 31), __shfl_sync(0xffffffffu, inc.q, 31)};
            }

            // carry from the chunks before this one (n_chunks > 1 implies one row per CTA: TPR == NT)
            float h_in = 0.0f, pcum_in = 1.0f;
            if (a.n_chunks > 1) {
                const long long srow = ((long long)b * a.dim + d) * a.dstate + n;
                const int n_groups16 = (a.n_chunks + 15) >> 4;
                CarryEntry *l1_row = a.ws_entries + srow * a.n_chunks;
                CarryEntry *l2_row = a.ws_entries2 + srow * n_groups16;
                if (last_warp && lane == 0) publish_entry(l1_row + chunk, epoch, total.p, total.q);
                CarryLook look = look_issue(l1_row, l2_row, chunk, lane);
                bool ok;
                Aff grp = {1.0f, 0.0f};
                Aff acc = look_reduce(look, epoch, lane, ok, grp);
                acc = look_finish(look, acc, ok, l2_row, chunk, epoch, lane, grp);
                if (last_warp && lane == 0 && (chunk & 15) == 15) {
                    const Aff g16 = compose(grp, total);
                    publish_entry(l2_row + (chunk >> 4), epoch, g16.p, g16.q);
                }
                h_in = acc.q;
                pcum_in = acc.p;
            }
            if (last_warp && lane == 0 && active) {
                float2 *xs = reinterpret_cast<float2 *>(a.x) + (((long long)b * a.dim + d) * a.n_chunks + chunk) * a.dstate + n;
                *xs = make_float2(total.p * pcum_in, fmaf(total.p, h_in, total.q));
            }

            float h = fmaf(exc.p, h_in, exc.q);
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                h = fmaf(av[i], h, bx[i]);
                y[i] = fmaf(Cv[i], h, y[i]);
            }
        }
        if (active) {
            T *o_row = reinterpret_cast<T *>(a.out) + b * a.out_bs + d * a.out_ds;
            store_items<T, ITEMS, VEC>(o_row, pos, L, y);
        }
    }
    retire_tile(a);
}

template <typename T, int NT, int TPR, int ITEMS, bool N1, bool VEC>
static int launch(const ScanArgs &a, int grid, cudaStream_t stream) {
    scan_fwd_kernel<T, NT, TPR, ITEMS, N1, VEC><<<grid, NT, 0, stream>>>(a);
    return check_cuda(cudaGetLastError(), "scan_fwd launch");
}

template <typename T, bool N1, bool VEC>
static int dispatch_shape(const ScanArgs &a, const ScanPlan &pl, cudaStream_t stream) {
    switch (pl.tpr) {
        case 32: return launch<T, 256, 32, 8, N1, VEC>(a, pl.grid, stream);
        case 64: return launch<T, 256, 64, 8, N1, VEC>(a, pl.grid, stream);
        case 128: return launch<T, 256, 128, 8, N1, VEC>(a, pl.grid, stream);
        default: return launch<T, 256, 256, 8, N1, VEC>(a, pl.grid, stream);
    }
}

template <typename T>
static int dispatch_flags(const ScanArgs &a, const ScanPlan &pl, cudaStream_t stream) {
    const bool n1 = a.dstate == 1;
    if (n1) return pl.vec ? dispatch_shape<T, true, true>(a, pl, stream) : dispatch_shape<T, true, false>(a, pl, stream);
    return pl.vec ? dispatch_shape<T, false, true>(a, pl, stream) : dispatch_shape<T, false, false>(a, pl, stream);
}

int scan_fwd_dispatch(const ScanArgs &a, const ScanPlan &pl, int io_dtype, cudaStream_t stream) {
    switch (io_dtype) {
        case VMASR_F32: return dispatch_flags<float>(a, pl, stream);
        case VMASR_F16: return dispatch_flags<__half>(a, pl, stream);
        case VMASR_BF16: return dispatch_flags<__nv_bfloat16>(a, pl, stream);
    }
    return fail("selective_scan_fwd: unsupported io dtype %d", io_dtype);
}

}  // namespace vmasr
