// 4-direction cross scan / cross merge for sm_100a, and the two map kernels of the fused SS2D core.
// Replaces CrossScan / CrossMerge (model/vmamba.py:27-73) and triton_cross_scan / triton_cross_merge
// (model/csm_triton.py:7-154).  One read of a TILE x TILE tile of the (H, W) map feeds all four directions:
//   k=0  l = h*W + w          k=1  l = w*H + h          k=2, k=3: the same two walked backwards (L-1-l)
// The row-major pair is moved with 128-bit accesses along w, the column-major pair with 128-bit accesses
// along h after a shared-memory transpose, so every global access of every direction is coalesced (the
// Triton kernel's transposed stores are stride-H).  Pure data movement (scan) / three additions in the
// reference's association (merge): bit-exact with the PyTorch versions, for fp32, fp16 and bf16 (a 128-bit access is 4 or
// 8 elements).  Small maps (<= 32 x 32: the two deepest stages of the U-Net) take several channels per CTA.
// The grid is one-dimensional over (plane group, tile row, tile column): no limit on B * C.
//
// Fused SS2D core (ss2d.cu): directions 1 / 3 read a TRANSPOSED copy of the map (map_transpose) and the two output planes
// (row-major pair, column-major pair) are combined by map_merge2:  y = P02 + transpose(P13)  -- the outer addition of
// vmamba.py:55-60, with the same association.
#include "common.cuh"

namespace vmasr {

template <typename T> struct Vec16 { static constexpr int N = 16 / sizeof(T); };

template <typename T> __device__ __forceinline__ T add_rn(T a, T b);
template <> __device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }
// fp16/bf16 + fp16/bf16 is exact in fp32, so one rounding of the fp32 sum equals the native half add
template <> __device__ __forceinline__ __half add_rn<__half>(__half a, __half b) { return __float2half_rn(__half2float(a) + __half2float(b)); }
template <> __device__ __forceinline__ __nv_bfloat16 add_rn<__nv_bfloat16>(__nv_bfloat16 a, __nv_bfloat16 b) {
    return __float2bfloat16_rn(__bfloat162float(a) + __bfloat162float(b));
}

// 16 bytes of T as registers
template <typename T> struct alignas(16) Pack {
    T v[Vec16<T>::N];
};
template <typename T> __device__ __forceinline__ Pack<T> ld16(const T *p) {
    Pack<T> r;
    *reinterpret_cast<uint4 *>(r.v) = __ldg(reinterpret_cast<const uint4 *>(p));
    return r;
}
template <typename T> __device__ __forceinline__ void st16(T *p, const Pack<T> &r) {
    *reinterpret_cast<uint4 *>(p) = *reinterpret_cast<const uint4 *>(r.v);
}
template <typename T> __device__ __forceinline__ Pack<T> rev16(const Pack<T> &a) {
    Pack<T> r;
#pragma unroll
    for (int i = 0; i < Vec16<T>::N; ++i) r.v[i] = a.v[Vec16<T>::N - 1 - i];
    return r;
}
template <typename T> __device__ __forceinline__ Pack<T> add16(const Pack<T> &a, const Pack<T> &b) {
    Pack<T> r;
#pragma unroll
    for (int i = 0; i < Vec16<T>::N; ++i) r.v[i] = add_rn<T>(a.v[i], b.v[i]);
    return r;
}

// Tile bookkeeping shared by the vector kernels.  A CTA owns tile (th, tw) of CPB consecutive (b, c) planes.
template <int TILE, int CPB> struct TileJob {
    long long plane0;  // first plane (b * C + c)
    int n_planes, h0, w0, hv, wv;
    __device__ __forceinline__ TileJob(long long planes, int H, int W) {
        const int tw_n = (W + TILE - 1) / TILE, th_n = (H + TILE - 1) / TILE;
        long long t = blockIdx.x;
        const int tw = (int)(t % tw_n);
        t /= tw_n;
        const int th = (int)(t % th_n);
        t /= th_n;
        plane0 = t * CPB;
        n_planes = (int)min((long long)CPB, planes - plane0);
        h0 = th * TILE;
        w0 = tw * TILE;
        hv = min(TILE, H - h0);
        wv = min(TILE, W - w0);
    }
};
template <int TILE, int CPB> static long long tile_grid(long long planes, int H, int W) {
    return ((planes + CPB - 1) / CPB) * ((H + TILE - 1) / TILE) * ((W + TILE - 1) / TILE);
}

// ---- H and W multiples of the 128-bit vector length, 16-byte aligned bases ----------------------------------------
template <typename T, int TILE, int CPB>
__global__ void __launch_bounds__(256) cross_scan_vec(const T *__restrict__ x, T *__restrict__ xs, long long planes, int C, int H, int W) {
    constexpr int VE = Vec16<T>::N, PAD = TILE + (sizeof(T) == 4 ? 1 : 2);
    __shared__ T s[CPB][TILE * PAD];
    const TileJob<TILE, CPB> job(planes, H, W);
    const long long L = (long long)H * W;
    const int vpr = job.wv / VE, vpc = job.hv / VE;  // vectors per tile row / per tile column
    for (int i = threadIdx.x; i < job.n_planes * job.hv * vpr; i += 256) {
        const int pl = i / (job.hv * vpr), r = i - pl * job.hv * vpr;
        const int hl = r / vpr, wl = (r - hl * vpr) * VE;
        const long long bc = job.plane0 + pl, b = bc / C, c = bc - b * C;
        const long long l = (long long)(job.h0 + hl) * W + job.w0 + wl;
        const Pack<T> v = ld16(x + bc * L + l);
        st16(xs + ((b * 4 + 0) * C + c) * L + l, v);
        st16(xs + ((b * 4 + 2) * C + c) * L + (L - VE - l), rev16(v));
#pragma unroll
        for (int k = 0; k < VE; ++k) s[pl][hl * PAD + wl + k] = v.v[k];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < job.n_planes * job.wv * vpc; i += 256) {
        const int pl = i / (job.wv * vpc), r = i - pl * job.wv * vpc;
        const int wl = r / vpc, hl = (r - wl * vpc) * VE;
        const long long bc = job.plane0 + pl, b = bc / C, c = bc - b * C;
        Pack<T> t;
#pragma unroll
        for (int k = 0; k < VE; ++k) t.v[k] = s[pl][(hl + k) * PAD + wl];
        const long long l = (long long)(job.w0 + wl) * H + job.h0 + hl;
        st16(xs + ((b * 4 + 1) * C + c) * L + l, t);
        st16(xs + ((b * 4 + 3) * C + c) * L + (L - VE - l), rev16(t));
    }
}

template <typename T, int TILE, int CPB>
__global__ void __launch_bounds__(256) cross_merge_vec(const T *__restrict__ ys, T *__restrict__ y, long long planes, int C, int H, int W) {
    constexpr int VE = Vec16<T>::N, PAD = TILE + (sizeof(T) == 4 ? 1 : 2);
    __shared__ T s[CPB][TILE * PAD];
    const TileJob<TILE, CPB> job(planes, H, W);
    const long long L = (long long)H * W;
    const int vpr = job.wv / VE, vpc = job.hv / VE;
    // column-major pair first: (ys1 + flip ys3), transposed into shared memory
    for (int i = threadIdx.x; i < job.n_planes * job.wv * vpc; i += 256) {
        const int pl = i / (job.wv * vpc), r = i - pl * job.wv * vpc;
        const int wl = r / vpc, hl = (r - wl * vpc) * VE;
        const long long bc = job.plane0 + pl, b = bc / C, c = bc - b * C;
        const long long l = (long long)(job.w0 + wl) * H + job.h0 + hl;
        const Pack<T> a = ld16(ys + ((b * 4 + 1) * C + c) * L + l);
        const Pack<T> bb = rev16(ld16(ys + ((b * 4 + 3) * C + c) * L + (L - VE - l)));
        const Pack<T> t = add16(a, bb);
#pragma unroll
        for (int k = 0; k < VE; ++k) s[pl][(hl + k) * PAD + wl] = t.v[k];
    }
    // the row-major pair (ys0 + flip ys2) does not depend on the transposed tile: a thread's first pair of loads goes out BEFORE the
    // barrier, so the two round trips to memory overlap (small maps are one item per thread: the whole kernel is those two trips)
    const int n_row = job.n_planes * job.hv * vpr;
    auto row_pair = [&](int i, int &pl, int &hl, int &wl, long long &bc, long long &l) {
        pl = i / (job.hv * vpr);
        const int r = i - pl * job.hv * vpr;
        hl = r / vpr;
        wl = (r - hl * vpr) * VE;
        bc = job.plane0 + pl;
        const long long b = bc / C, c = bc - b * C;
        l = (long long)(job.h0 + hl) * W + job.w0 + wl;
        const Pack<T> a = ld16(ys + ((b * 4 + 0) * C + c) * L + l);
        const Pack<T> bb = rev16(ld16(ys + ((b * 4 + 2) * C + c) * L + (L - VE - l)));
        return add16(a, bb);
    };
    int pl = 0, hl = 0, wl = 0;
    long long bc = 0, l = 0;
    Pack<T> rowp;
    if ((int)threadIdx.x < n_row) rowp = row_pair(threadIdx.x, pl, hl, wl, bc, l);
    __syncthreads();
    for (int i = threadIdx.x; i < n_row; i += 256) {
        if (i != (int)threadIdx.x) rowp = row_pair(i, pl, hl, wl, bc, l);
        Pack<T> colp;
#pragma unroll
        for (int k = 0; k < VE; ++k) colp.v[k] = s[pl][hl * PAD + wl + k];
        st16(y + bc * L + l, add16(rowp, colp));
    }
}

// x (planes, H, W) -> xT (planes, W, H)
template <int TILE, int CPB>
__global__ void __launch_bounds__(256) map_transpose_vec(const float *__restrict__ x, float *__restrict__ xT, long long planes, int H, int W) {
    constexpr int VE = 4, PAD = TILE + 1;
    __shared__ float s[CPB][TILE * PAD];
    const TileJob<TILE, CPB> job(planes, H, W);
    const long long L = (long long)H * W;
    const int vpr = job.wv / VE, vpc = job.hv / VE;
    for (int i = threadIdx.x; i < job.n_planes * job.hv * vpr; i += 256) {
        const int pl = i / (job.hv * vpr), r = i - pl * job.hv * vpr;
        const int hl = r / vpr, wl = (r - hl * vpr) * VE;
        const Pack<float> v = ld16(x + (job.plane0 + pl) * L + (long long)(job.h0 + hl) * W + job.w0 + wl);
#pragma unroll
        for (int k = 0; k < VE; ++k) s[pl][hl * PAD + wl + k] = v.v[k];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < job.n_planes * job.wv * vpc; i += 256) {
        const int pl = i / (job.wv * vpc), r = i - pl * job.wv * vpc;
        const int wl = r / vpc, hl = (r - wl * vpc) * VE;
        Pack<float> t;
#pragma unroll
        for (int k = 0; k < VE; ++k) t.v[k] = s[pl][(hl + k) * PAD + wl];
        st16(xT + (job.plane0 + pl) * L + (long long)(job.w0 + wl) * H + job.h0 + hl, t);
    }
}

// y (planes, H, W) = p_rm (planes, H, W) + transpose(p_cm (planes, W, H))
template <int TILE, int CPB>
__global__ void __launch_bounds__(256) map_merge2_vec(const float *__restrict__ p_rm, const float *__restrict__ p_cm, float *__restrict__ y,
                                                     long long planes, int H, int W) {
    constexpr int VE = 4, PAD = TILE + 1;
    __shared__ float s[CPB][TILE * PAD];
    const TileJob<TILE, CPB> job(planes, H, W);
    const long long L = (long long)H * W;
    const int vpr = job.wv / VE, vpc = job.hv / VE;
    for (int i = threadIdx.x; i < job.n_planes * job.wv * vpc; i += 256) {
        const int pl = i / (job.wv * vpc), r = i - pl * job.wv * vpc;
        const int wl = r / vpc, hl = (r - wl * vpc) * VE;
        const Pack<float> t = ld16(p_cm + (job.plane0 + pl) * L + (long long)(job.w0 + wl) * H + job.h0 + hl);
#pragma unroll
        for (int k = 0; k < VE; ++k) s[pl][(hl + k) * PAD + wl] = t.v[k];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < job.n_planes * job.hv * vpr; i += 256) {
        const int pl = i / (job.hv * vpr), r = i - pl * job.hv * vpr;
        const int hl = r / vpr, wl = (r - hl * vpr) * VE;
        const long long off = (job.plane0 + pl) * L + (long long)(job.h0 + hl) * W + job.w0 + wl;
        const Pack<float> a = ld16(p_rm + off);
        Pack<float> colp;
#pragma unroll
        for (int k = 0; k < VE; ++k) colp.v[k] = s[pl][hl * PAD + wl + k];
        st16(y + off, add16(a, colp));
    }
}

// ---- any shape, any of the three dtypes: 32x32 tiles, one element per access ------------------------
template <typename T>
__global__ void __launch_bounds__(256) cross_scan_any(const T *__restrict__ x, T *__restrict__ xs, long long planes, int C, int H, int W) {
    __shared__ T s[32][33];
    const TileJob<32, 1> job(planes, H, W);
    const long long bc = job.plane0, b = bc / C, c = bc - b * C;
    const long long L = (long long)H * W;
    const int h0 = job.h0, w0 = job.w0;
    const T *src = x + bc * L;
    T *d0 = xs + ((b * 4 + 0) * C + c) * L;
    T *d1 = xs + ((b * 4 + 1) * C + c) * L;
    T *d2 = xs + ((b * 4 + 2) * C + c) * L;
    T *d3 = xs + ((b * 4 + 3) * C + c) * L;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int h = h0 + r, w = w0 + tx;
        if (h < H && w < W) {
            const long long l = (long long)h * W + w;
            const T v = src[l];
            d0[l] = v;
            d2[L - 1 - l] = v;
            s[r][tx] = v;
        }
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int w = w0 + r, h = h0 + tx;
        if (h < H && w < W) {
            const long long l = (long long)w * H + h;
            const T v = s[tx][r];
            d1[l] = v;
            d3[L - 1 - l] = v;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) cross_merge_any(const T *__restrict__ ys, T *__restrict__ y, long long planes, int C, int H, int W) {
    __shared__ T s[32][33];
    const TileJob<32, 1> job(planes, H, W);
    const long long bc = job.plane0, b = bc / C, c = bc - b * C;
    const long long L = (long long)H * W;
    const int h0 = job.h0, w0 = job.w0;
    const T *s0 = ys + ((b * 4 + 0) * C + c) * L;
    const T *s1 = ys + ((b * 4 + 1) * C + c) * L;
    const T *s2 = ys + ((b * 4 + 2) * C + c) * L;
    const T *s3 = ys + ((b * 4 + 3) * C + c) * L;
    T *dst = y + bc * L;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int w = w0 + r, h = h0 + tx;
        if (h < H && w < W) {
            const long long l = (long long)w * H + h;
            s[tx][r] = add_rn<T>(s1[l], s3[L - 1 - l]);
        }
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int h = h0 + r, w = w0 + tx;
        if (h < H && w < W) {
            const long long l = (long long)h * W + w;
            dst[l] = add_rn<T>(add_rn<T>(s0[l], s2[L - 1 - l]), s[r][tx]);
        }
    }
}

// ---- one-by-one variants (csm_triton.py:157-308): every direction has its OWN map --------------------------------------
//   scan_1b1 : x (B, 4, C, H, W) -> y (B, 4, C, L):  y0 = x0 row-major, y1 = x1 column-major, y2 / y3 = the same of x2 / x3 reversed
//   merge_1b1: the inverse permutation, y (B, 4, C, L) -> x (B, 4, C, H, W)   (no sum: four separate maps)
// Reached only from SS2D.forwardxv (vmamba.py:1618-1690), which no shipped config selects: a plain 32 x 32 tile kernel.
template <typename T, bool INVERSE>
__global__ void __launch_bounds__(256) cross_1b1_any(const T *__restrict__ src, T *__restrict__ dst, long long planes, int C, int H, int W) {
    __shared__ T s[32][33];
    const TileJob<32, 1> job(planes, H, W);  // planes = B * 4 * C; direction k = (plane / C) % 4
    const long long pl = job.plane0;
    const int k = (int)((pl / C) & 3);
    const long long L = (long long)H * W;
    const T *in = src + pl * L;
    T *out = dst + pl * L;
    const int h0 = job.h0, w0 = job.w0;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const bool rev = k >= 2;
    if (!(k & 1)) {  // row-major directions: a copy, possibly reversed
        for (int r = ty; r < 32; r += 8) {
            const int h = h0 + r, w = w0 + tx;
            if (h < H && w < W) {
                const long long l = (long long)h * W + w, m = rev ? L - 1 - l : l;
                if (INVERSE) out[l] = in[m];
                else out[m] = in[l];
            }
        }
        return;
    }
    // column-major directions through a shared-memory transpose: map position (h, w) <-> sequence index w * H + h
    for (int r = ty; r < 32; r += 8) {
        if (!INVERSE) {
            const int h = h0 + r, w = w0 + tx;
            if (h < H && w < W) s[r][tx] = in[(long long)h * W + w];
        } else {
            const int w = w0 + r, h = h0 + tx;
            if (h < H && w < W) {
                const long long l = (long long)w * H + h;
                s[tx][r] = in[rev ? L - 1 - l : l];
            }
        }
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        if (!INVERSE) {
            const int w = w0 + r, h = h0 + tx;
            if (h < H && w < W) {
                const long long l = (long long)w * H + h;
                out[rev ? L - 1 - l : l] = s[tx][r];
            }
        } else {
            const int h = h0 + r, w = w0 + tx;
            if (h < H && w < W) out[(long long)h * W + w] = s[r][tx];
        }
    }
}

template <bool INVERSE>
static int run_1b1(const void *in, void *out, int B, int C, int H, int W, int dtype, int device, void *stream_) {
    const char *who = INVERSE ? "cross_merge_1b1" : "cross_scan_1b1";
    if (!in || !out) return fail("%s: null tensor", who);
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return fail("%s: sizes must be positive (B %d C %d H %d W %d)", who, B, C, H, W);
    DeviceGuard guard(device);
    if (!guard.ok) return fail("%s: cannot select CUDA device %d", who, device);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const long long planes = (long long)B * 4 * C;
    const long long g = planes * ((H + 31) / 32) * ((W + 31) / 32);
    if (g > 0x7fffffffll) return fail("%s: %lld tiles exceed the grid limit", who, g);
    switch (dtype) {
        case VMASR_F32: cross_1b1_any<float, INVERSE><<<(unsigned)g, 256, 0, stream>>>(static_cast<const float *>(in), static_cast<float *>(out), planes, C, H, W); break;
        case VMASR_F16: cross_1b1_any<__half, INVERSE><<<(unsigned)g, 256, 0, stream>>>(static_cast<const __half *>(in), static_cast<__half *>(out), planes, C, H, W); break;
        case VMASR_BF16: cross_1b1_any<__nv_bfloat16, INVERSE><<<(unsigned)g, 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(in), static_cast<__nv_bfloat16 *>(out), planes, C, H, W); break;
        default: return fail("%s: unsupported dtype %d", who, dtype);
    }
    return check_cuda(cudaGetLastError(), who);
}

static int check_shape(const void *a, const void *b, int B, int C, int H, int W, int dtype, const char *who) {
    if (!a || !b) return fail("%s: null tensor", who);
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return fail("%s: sizes must be positive (B %d C %d H %d W %d)", who, B, C, H, W);
    if (dtype != VMASR_F32 && dtype != VMASR_F16 && dtype != VMASR_BF16) return fail("%s: unsupported dtype %d", who, dtype);
    return 0;
}

static int grid_ok(long long g, const char *who) {
    if (g > 0x7fffffffll) return fail("%s: %lld tiles exceed the grid limit", who, g);
    return 0;
}

template <bool MERGE, typename T>
static int run_cross_t(const void *in_, void *out_, int B, int C, int H, int W, cudaStream_t stream, const char *who) {
    const T *in = static_cast<const T *>(in_);
    T *out = static_cast<T *>(out_);
    constexpr int VE = Vec16<T>::N;
    const long long planes = (long long)B * C;
    const bool vec = (H % VE == 0) && (W % VE == 0) && ((reinterpret_cast<uintptr_t>(in_) | reinterpret_cast<uintptr_t>(out_)) & 15u) == 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    const long long sms = sm_count(dev);
    // maps larger than 32 x 32: tiles of 32 x 32 instead of 64 x 64 for the fp32 merge (its two dependent trips to memory want many
    // small CTAs: 7.2 vs 9.9 us on 64 x 64 maps, 11.2 vs 12.0 on 128 x 128, 19.4 vs 20.7 on 256 x 256) and for an fp32 scan of fewer than
    // three 64 x 64 tiles per SM; half precision keeps 64 x 64 (a 32-wide row is only 64 bytes).  profiles/r2_s6_cross_tiles.txt
    bool t32 = sizeof(T) == 4 && (MERGE || tile_grid<64, 1>(planes, H, W) < 3 * sms);
    if (const char *e = tuning_env("VMASR_CROSS_T32")) t32 = atoi(e) != 0;
    if (vec && H <= 32 && W <= 32) {
        // small maps: four planes per CTA when that still leaves three CTAs per SM; otherwise as many planes as give every thread
        // one 16-byte item (a 32 x 32 fp32 plane is 256 items, a 16 x 16 one 64)
        const long long items = (long long)H * W / VE;
        const long long g4 = tile_grid<32, 4>(planes, H, W), g2 = tile_grid<32, 2>(planes, H, W), g1 = tile_grid<32, 1>(planes, H, W);
        if (int rc = grid_ok(g1, who)) return rc;
        if (g4 >= 3 * sms || items <= 64) {
            if (MERGE) cross_merge_vec<T, 32, 4><<<(unsigned)g4, 256, 0, stream>>>(in, out, planes, C, H, W);
            else cross_scan_vec<T, 32, 4><<<(unsigned)g4, 256, 0, stream>>>(in, out, planes, C, H, W);
        } else if (items <= 128) {
            if (MERGE) cross_merge_vec<T, 32, 2><<<(unsigned)g2, 256, 0, stream>>>(in, out, planes, C, H, W);
            else cross_scan_vec<T, 32, 2><<<(unsigned)g2, 256, 0, stream>>>(in, out, planes, C, H, W);
        } else {
            if (MERGE) cross_merge_vec<T, 32, 1><<<(unsigned)g1, 256, 0, stream>>>(in, out, planes, C, H, W);
            else cross_scan_vec<T, 32, 1><<<(unsigned)g1, 256, 0, stream>>>(in, out, planes, C, H, W);
        }
    } else if (vec && t32) {
        const long long g = tile_grid<32, 1>(planes, H, W);
        if (int rc = grid_ok(g, who)) return rc;
        if (MERGE) cross_merge_vec<T, 32, 1><<<(unsigned)g, 256, 0, stream>>>(in, out, planes, C, H, W);
        else cross_scan_vec<T, 32, 1><<<(unsigned)g, 256, 0, stream>>>(in, out, planes, C, H, W);
    } else if (vec) {
        const long long g = tile_grid<64, 1>(planes, H, W);
        if (int rc = grid_ok(g, who)) return rc;
        if (MERGE) cross_merge_vec<T, 64, 1><<<(unsigned)g, 256, 0, stream>>>(in, out, planes, C, H, W);
        else cross_scan_vec<T, 64, 1><<<(unsigned)g, 256, 0, stream>>>(in, out, planes, C, H, W);
    } else {
        const long long g = tile_grid<32, 1>(planes, H, W);
        if (int rc = grid_ok(g, who)) return rc;
        if (MERGE) cross_merge_any<T><<<(unsigned)g, 256, 0, stream>>>(in, out, planes, C, H, W);
        else cross_scan_any<T><<<(unsigned)g, 256, 0, stream>>>(in, out, planes, C, H, W);
    }
    return check_cuda(cudaGetLastError(), who);
}

template <bool MERGE>
static int run_cross(const void *in, void *out, int B, int C, int H, int W, int dtype, int device, void *stream_) {
    const char *who = MERGE ? "cross_merge" : "cross_scan";
    if (int rc = check_shape(in, out, B, C, H, W, dtype, who)) return rc;
    DeviceGuard guard(device);
    if (!guard.ok) return fail("%s: cannot select CUDA device %d", who, device);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    switch (dtype) {
        case VMASR_F32: return run_cross_t<MERGE, float>(in, out, B, C, H, W, stream, who);
        case VMASR_F16: return run_cross_t<MERGE, __half>(in, out, B, C, H, W, stream, who);
        default: return run_cross_t<MERGE, __nv_bfloat16>(in, out, B, C, H, W, stream, who);
    }
}

// fp32 maps with H % 4 == 0 and W % 4 == 0 (what the fused SS2D core accepts); launched on the caller's device
int map_transpose_launch(const float *x, float *xT, long long planes, int H, int W, cudaStream_t stream) {
    if (H <= 32 && W <= 32) {
        const long long g = tile_grid<32, 4>(planes, H, W);
        if (int rc = grid_ok(g, "map_transpose")) return rc;
        map_transpose_vec<32, 4><<<(unsigned)g, 256, 0, stream>>>(x, xT, planes, H, W);
    } else {
        const long long g = tile_grid<64, 1>(planes, H, W);
        if (int rc = grid_ok(g, "map_transpose")) return rc;
        map_transpose_vec<64, 1><<<(unsigned)g, 256, 0, stream>>>(x, xT, planes, H, W);
    }
    return check_cuda(cudaGetLastError(), "map_transpose");
}

int map_merge2_launch(const float *p_rm, const float *p_cm, float *y, long long planes, int H, int W, cudaStream_t stream) {
    if (H <= 32 && W <= 32) {
        const long long g = tile_grid<32, 4>(planes, H, W);
        if (int rc = grid_ok(g, "map_merge2")) return rc;
        map_merge2_vec<32, 4><<<(unsigned)g, 256, 0, stream>>>(p_rm, p_cm, y, planes, H, W);
    } else {
        const long long g = tile_grid<64, 1>(planes, H, W);
        if (int rc = grid_ok(g, "map_merge2")) return rc;
        map_merge2_vec<64, 1><<<(unsigned)g, 256, 0, stream>>>(p_rm, p_cm, y, planes, H, W);
    }
    return check_cuda(cudaGetLastError(), "map_merge2");
}

}  // namespace vmasr

extern "C" int vmasr_cross_scan(const void *x, void *xs, int B, int C, int H, int W, int dtype, int device, void *stream) {
    return vmasr::run_cross<false>(x, xs, B, C, H, W, dtype, device, stream);
}
extern "C" int vmasr_cross_merge(const void *ys, void *y, int B, int C, int H, int W, int dtype, int device, void *stream) {
    return vmasr::run_cross<true>(ys, y, B, C, H, W, dtype, device, stream);
}
extern "C" int vmasr_cross_scan_1b1(const void *x, void *y, int B, int C, int H, int W, int dtype, int device, void *stream) {
    return vmasr::run_1b1<false>(x, y, B, C, H, W, dtype, device, stream);
}
extern "C" int vmasr_cross_merge_1b1(const void *y, void *x, int B, int C, int H, int W, int dtype, int device, void *stream) {
    return vmasr::run_1b1<true>(y, x, B, C, H, W, dtype, device, stream);
}
