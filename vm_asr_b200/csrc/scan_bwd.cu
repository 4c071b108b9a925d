// Selective-scan backward for sm_100a.
// Replaces selective_scan_bwd_kernel (kernels/selective_scan/csrc/selective_scan/cus/selective_scan_bwd_kernel.cuh:66-273).
// Per (batch, channel, state) with g the adjoint of the state:
//   g_l      = C_l dout_l + a_{l+1} g_{l+1}                         (runs right to left)
//   du_l     = D dout_l + sum_n g_l dt_l B_l
//   ddt_l    = sum_n g_l (B_l u_l + A a_l h_{l-1})                  a_l h_{l-1} = h_l - dt_l B_l u_l
//   dA       = sum_l g_l dt_l a_l h_{l-1};   dB_l += g_l dt_l u_l;   dC_l += dout_l h_l;   dD = sum dout u
//   ddelta_l = ddt_l * sigmoid(delta_l + bias)  (softplus on, input <= 20);   ddelta_bias = sum_l ddelta_l
// The forward states are recomputed per chunk from the chunk-end states `x` the forward saved; the adjoint
// carry between chunks uses the same look-back exchange as the forward, walking the chunks downwards.
#include "pipe.cuh"

namespace vmasr {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

template <int ITEMS, bool VEC>
__device__ __forceinline__ void accumulate_items(float *__restrict__ row, int pos, int len, const float (&v)[ITEMS]) {
    if (VEC && pos + ITEMS <= len) {
#pragma unroll
        for (int i = 0; i < ITEMS; i += 4)
            atomicAdd(reinterpret_cast<float4 *>(row + pos + i), make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
    } else {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i)
            if (pos + i < len) atomicAdd(row + pos + i, v[i]);
    }
}

// slot 0: dA (only cached in shared memory when dstate == 1), slot 1: dD, slot 2: ddelta_bias
template <bool N1>
__device__ __forceinline__ void flush_channel_sum(const ScanArgs &a, int slot, int d, float v) {
    if (slot == 0) {
        if (N1) atomicAdd(a.dA + d * a.A_ds, v);
    } else {
        float *dst = slot == 1 ? a.dD : a.ddelta_bias;
        if (dst) atomicAdd(dst + d, v);
    }
}

template <typename T, int NT, int TPR, int ITEMS, bool N1, bool VEC>
__global__ void __launch_bounds__(NT) scan_bwd_kernel(const __grid_constant__ ScanArgs a) {
    constexpr int ROWS = NT / TPR;
    constexpr int WPR = TPR / 32;
    constexpr int CHUNK = TPR * ITEMS;
    __shared__ float4 s_tot[2][ROWS][WPR > 1 ? WPR : 1];  // per warp: {decay, fwd q, adjoint q, -}
    __shared__ float s_red[2][ROWS][4];                    // per channel sums: dA, dD, ddelta_bias
    __shared__ float s_dbc[(ROWS > 1) ? 2 * ROWS * TPR * ITEMS : 1];
    __shared__ unsigned s_tile[2];

    if (threadIdx.x < 2 * ROWS * 4) (&s_red[0][0][0])[threadIdx.x] = 0.0f;
    __syncthreads();

    unsigned tile, epoch;
    claim_tile(a, s_tile, tile, epoch);

    const int chunk = a.n_chunks - 1 - (int)(tile / a.n_rowgroups);  // adjoint: high chunks first
    const int rg = tile % a.n_rowgroups;
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
    float *dBg = a.dB + ((long long)b * a.ngroups + g) * a.dstate * (long long)L;
    float *dCg = a.dC + ((long long)b * a.ngroups + g) * a.dstate * (long long)L;

    float Bv[ITEMS], Cv[ITEMS], dBacc[ITEMS], dCacc[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) dBacc[i] = dCacc[i] = 0.0f;
    if (N1) {
        load_items<T, ITEMS, VEC>(Bg, pos, L, Bv, 0.0f);
        load_items<T, ITEMS, VEC>(Cg, pos, L, Cv, 0.0f);
    }

    int buf = 0;
    int d_prev = -1;
    for (int it = 0; it < n_iter; ++it) {
        const int c = c_begin + it * ROWS + row;
        const bool active = c < c_end;
        const float act = active ? 1.0f : 0.0f;
        const int rbuf = it & 1;
        const int d = g * a.chan_per_group + (active ? c : c_begin);
        const T *u_row = reinterpret_cast<const T *>(a.u) + b * a.u_bs + d * a.u_ds;
        const T *dl_row = reinterpret_cast<const T *>(a.delta) + b * a.delta_bs + d * a.delta_ds;
        const T *dy_row = reinterpret_cast<const T *>(a.dout) + b * a.dout_bs + d * a.dout_ds;

        float uv[ITEMS], dt[ITEMS], sig[ITEMS], dy[ITEMS], du[ITEMS], ddt[ITEMS];
        load_items<T, ITEMS, VEC>(u_row, pos, L, uv, 0.0f);
        load_items<T, ITEMS, VEC>(dl_row, pos, L, dt, 0.0f);
        load_items<T, ITEMS, VEC>(dy_row, pos, L, dy, 0.0f);
        const float bias = a.delta_bias ? __ldg(a.delta_bias + d) : 0.0f;
        const float Dv = a.D ? __ldg(a.D + d) : 0.0f;
        float dD_acc = 0.0f, dbias_acc = 0.0f;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            float x = dt[i] + bias;
            sig[i] = 1.0f;
            dt[i] = a.softplus ? softplus_sig<true>(x, sig[i]) : x;
            du[i] = Dv * dy[i];
            ddt[i] = 0.0f;
            dD_acc = fmaf(dy[i], uv[i], dD_acc);
        }

        for (int n = 0; n < (N1 ? 1 : a.dstate); ++n) {
            if (!N1) {
                load_items<T, ITEMS, VEC>(Bg + n * a.B_ns, pos, L, Bv, 0.0f);
                load_items<T, ITEMS, VEC>(Cg + n * a.C_ns, pos, L, Cv, 0.0f);
            }
            const float Aval = __ldg(a.A + d * a.A_ds + n * a.A_ns);
            const float A2 = Aval * kLog2e;
            // state entering this chunk, saved by the forward
            float h_in = 0.0f;
            if (chunk > 0)
                h_in = __ldg(a.x + ((((long long)b * a.dim + d) * a.n_chunks + (chunk - 1)) * a.dstate + n) * 2 + 1);

            float av[ITEMS];
            Aff loc_f = {1.0f, 0.0f};
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const bool valid = pos + i < L;
                av[i] = valid ? ex2_approx(dt[i] * A2) : 1.0f;
                loc_f.q = fmaf(av[i], loc_f.q, dt[i] * uv[i] * Bv[i]);
                loc_f.p *= av[i];
            }
            float qr = 0.0f;
#pragma unroll
            for (int i = ITEMS - 1; i >= 0; --i) qr = av[i] * fmaf(Cv[i], dy[i], qr);
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
                    s_tot[buf][row][warp_in_row].x = inc_f.p;
                    s_tot[buf][row][warp_in_row].y = inc_f.q;
                }
                if (lane == 0) s_tot[buf][row][warp_in_row].z = inc_r.q;
                __syncthreads();
                if (n == 0 && t_in_row < 3 && d_prev >= 0) {
                    // everyone is past iteration it-1: hand its per-channel sums to global memory
                    const float v = s_red[rbuf ^ 1][row][t_in_row];
                    s_red[rbuf ^ 1][row][t_in_row] = 0.0f;
                    flush_channel_sum<N1>(a, t_in_row, d_prev, v);
                }
                Aff before_f = {1.0f, 0.0f}, run = {1.0f, 0.0f};
#pragma unroll
                for (int w = 0; w < WPR; ++w) {
                    const float4 t = s_tot[buf][row][w];
                    if (w == warp_in_row) before_f = run;
                    run = compose(run, Aff{t.x, t.y});
                }
                Aff before_r = {1.0f, 0.0f};
                run = {1.0f, 0.0f};
#pragma unroll
                for (int w = WPR - 1; w >= 0; --w) {
                    const float4 t = s_tot[buf][row][w];
                    if (w == warp_in_row) before_r = run;
                    run = compose(run, Aff{t.x, t.z});
                }
                total_r = run;
                exc_f = compose(before_f, exc_f);
                exc_r = compose(before_r, exc_r);
            } else {
                total_r = {__shfl_sync(0xffffffffu, inc_r.p, 0), __shfl_sync(0xffffffffu, inc_r.q, 0)};
            }

            // adjoint carry from the chunks after this one (n_chunks > 1 implies one row per CTA)
            float g_in = 0.0f;
            if (a.n_chunks > 1) {
                // the adjoint runs right to left: entries are indexed by the chunk's scan-order position j
                const long long srow = ((long long)b * a.dim + d) * a.dstate + n;
                const int n_groups16 = (a.n_chunks + 15) >> 4;
                const int j = a.n_chunks - 1 - chunk;
                CarryEntry *l1_row = a.ws_entries + srow * a.n_chunks;
                CarryEntry *l2_row = a.ws_entries2 + srow * n_groups16;
                if (last_warp && lane == 0) publish_entry(l1_row + j, epoch, total_r.p, total_r.q);
                CarryLook look = look_issue(l1_row, l2_row, j, lane);
                bool ok;
                Aff grp = {1.0f, 0.0f};
                Aff acc = look_reduce(look, epoch, lane, ok, grp);
                acc = look_finish(look, acc, ok, l2_row, j, epoch, lane, grp);
                if (last_warp && lane == 0 && (j & 15) == 15) {
                    const Aff g16 = compose(grp, total_r);
                    publish_entry(l2_row + (j >> 4), epoch, g16.p, g16.q);
                }
                g_in = acc.q;
            }

            // forward states of this thread's positions
            float hs[ITEMS];
            float h = fmaf(exc_f.p, h_in, exc_f.q);
            const float h_start = h;
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                h = fmaf(av[i], h, dt[i] * uv[i] * Bv[i]);
                hs[i] = h;
            }
            // adjoint walk, right to left
            float G = fmaf(exc_r.p, g_in, exc_r.q);
            float dA_acc = 0.0f;
            float dBn[ITEMS], dCn[ITEMS];
#pragma unroll
            for (int i = ITEMS - 1; i >= 0; --i) {
                const float gl = fmaf(Cv[i], dy[i], G);
                G = av[i] * gl;
                const float carried = av[i] * (i > 0 ? hs[i - 1] : h_start);
                const float gdt = gl * dt[i];
                du[i] = fmaf(gdt, Bv[i], du[i]);
                ddt[i] = fmaf(gl, fmaf(Bv[i], uv[i], Aval * carried), ddt[i]);
                dA_acc = fmaf(gdt, carried, dA_acc);
                if (N1) {
                    dBacc[i] = fmaf(gdt * act, uv[i], dBacc[i]);
                    dCacc[i] = fmaf(dy[i] * act, hs[i], dCacc[i]);
                } else {
                    dBn[i] = gdt * uv[i];
                    dCn[i] = dy[i] * hs[i];
                }
            }
            if (!N1 && active) {
                accumulate_items<ITEMS, VEC>(dBg + (long long)n * L, pos, L, dBn);
                accumulate_items<ITEMS, VEC>(dCg + (long long)n * L, pos, L, dCn);
            }
            dA_acc = warp_sum(dA_acc);
            if (lane == 0 && active) {
                if (WPR > 1 && N1) atomicAdd(&s_red[rbuf][row][0], dA_acc);
                else atomicAdd(a.dA + d * a.A_ds + n * a.A_ns, dA_acc);
            }
            if (WPR > 1) buf ^= 1;
        }

        float ddl[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            ddl[i] = ddt[i] * sig[i];
            dbias_acc += ddl[i];
        }
        if (active) {
            T *du_row = reinterpret_cast<T *>(a.du) + b * a.du_bs + d * a.du_ds;
            T *dd_row = reinterpret_cast<T *>(a.ddelta) + b * a.ddelta_bs + d * a.ddelta_ds;
            store_items<T, ITEMS, VEC>(du_row, pos, L, du);
            store_items<T, ITEMS, VEC>(dd_row, pos, L, ddl);
        }
        dD_acc = warp_sum(dD_acc);
        dbias_acc = warp_sum(dbias_acc);
        if (lane == 0 && active) {
            if (WPR > 1) {
                atomicAdd(&s_red[rbuf][row][1], dD_acc);
                atomicAdd(&s_red[rbuf][row][2], dbias_acc);
            } else {
                if (a.dD) atomicAdd(a.dD + d, dD_acc);
                if (a.ddelta_bias) atomicAdd(a.ddelta_bias + d, dbias_acc);
            }
        }
        d_prev = active ? d : -1;
    }

    if (WPR > 1) {
        __syncthreads();
        // sums of the final iteration
        if (t_in_row < 3 && d_prev >= 0)
            flush_channel_sum<N1>(a, t_in_row, d_prev, s_red[(n_iter - 1) & 1][row][t_in_row]);
    }

    if (N1) {
        // dB / dC of this tile's positions, summed over the tile's channels
        if (ROWS > 1) {
            float *sB = s_dbc + (row * TPR + t_in_row) * ITEMS;
            float *sC = s_dbc + ROWS * TPR * ITEMS + (row * TPR + t_in_row) * ITEMS;
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                sB[i] = dBacc[i];
                sC[i] = dCacc[i];
            }
            __syncthreads();
            if (row == 0) {
                for (int r = 1; r < ROWS; ++r) {
#pragma unroll
                    for (int i = 0; i < ITEMS; ++i) {
                        dBacc[i] += sB[r * TPR * ITEMS + i];
                        dCacc[i] += sC[r * TPR * ITEMS + i];
                    }
                }
            }
        }
        if (row == 0) {
            accumulate_items<ITEMS, VEC>(dBg, pos, L, dBacc);
            accumulate_items<ITEMS, VEC>(dCg, pos, L, dCacc);
        }
    }
    retire_tile(a);
}

template <typename T, int NT, int TPR, int ITEMS, bool N1, bool VEC>
static int launch(const ScanArgs &a, int grid, cudaStream_t stream) {
    scan_bwd_kernel<T, NT, TPR, ITEMS, N1, VEC><<<grid, NT, 0, stream>>>(a);
    return check_cuda(cudaGetLastError(), "scan_bwd launch");
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

int scan_bwd_dispatch(const ScanArgs &a, const ScanPlan &pl, int io_dtype, cudaStream_t stream) {
    switch (io_dtype) {
        case VMASR_F32: return dispatch_flags<float>(a, pl, stream);
        case VMASR_F16: return dispatch_flags<__half>(a, pl, stream);
        case VMASR_BF16: return dispatch_flags<__nv_bfloat16>(a, pl, stream);
    }
    return fail("selective_scan_bwd: unsupported io dtype %d", io_dtype);
}

}  // namespace vmasr
