// Head of the SS2D block fused into the core's load (second half).  SS2D.forwardv2 (model/vmamba.py:1541-1546):
//     x = x.permute(0, 3, 1, 2).contiguous()     channel-last (B, H, W, C) -> channel-first
//     x = self.conv2d(x)                         depthwise 3x3, padding 1, bias (vmamba.py:860-868)
//     x = self.act(x)                            SiLU
// and forward_corev2 then needs the map in float32 and (the fused core) its transpose.  The reference runs a permute copy, a
// cuDNN depthwise convolution, an elementwise SiLU, a cast and (here) a transpose: five kernels, ten passes.  One kernel reads the
// channel-last input once -- straight out of in_proj's (B, H, W, 2C) output: a position stride is taken, no chunk copy -- and
// writes x (B, C, H, W) and x^T (B, C, W, H) in float32.  Its backward takes the two map gradients the core's backward leaves
// (d x row-major, d x^T column-major), recomputes the convolution on a halo, and writes d input (channel-last) and per-patch
// partial sums of d weight / d bias.
//
// The convolution is depthwise, so a CTA owns a patch of PH x TW positions of a BLOCK of CB channels (no reduction over
// channels): input patch + halo in shared memory at an odd pitch per channel, so that loads from the channel-last input (lanes
// along channels) and the stencil / stores (lanes along w) are both conflict-free.  Global accesses: runs of CB elements per
// position for the channel-last tensors, 32-byte row pieces for x / d x, 128-bit pieces along h for x^T.  Bound: HBM.
//
// x_proj too (optional; SURVEY.md 8f-1: `x_dbl = einsum(xs, x_proj_weight)`, vmamba.py:1473-1475).  The projection commutes with
// the permutation of positions, so x_dbl of the row-major pair (directions 0, 2) and of the column-major pair (1, 3) are the
// same per-position contraction over channels stored in two position orders.  The CTA keeps its activations in shared memory,
// contracts them with its channel block's slice of the weight, and adds the 4 (R + 2) partial rows into x_dbl_rm / x_dbl_cm
// (plain stores when one block holds all channels).  The backward adds W^T d x_dbl to the map gradient before SiLU' and writes
// per-patch partial sums of d x_proj_weight.  The einsum's two cuBLAS launches and their passes over x and x^T are gone.
#include "common.cuh"

namespace vmasr {

struct DwArgs {
    const void *xin;   // (B, H, W, C) at position stride `ps` elements
    const float *w9;   // (C, 9)
    const float *bias; // (C) or null
    float *x, *xT;     // forward outputs
    const float *dx, *dxT;
    void *dxin;        // (B, H, W, C) contiguous
    float *dwb;        // (patches, C, 10): d weight (9) and d bias per patch
    const float *xpw, *xpb;        // x_proj weight (4, RP, C) and bias (4, RP) or null; RP = dt_rank + 2 d_state rows per direction
    float *xd_rm, *xd_cm;          // (B, 2, RP, L): x_dbl of directions (0, 2) in row-major and (1, 3) in column-major position order
    const float *dxd_rm, *dxd_cm;  // their gradients (backward)
    float *dxpw;                   // (patches, 4 RP, C): d x_proj weight per patch
    long long ps;
    int B, C, H, W, PH, TW, CB, RP;
    unsigned m_xw, m_dw, m_dh, m_dp, m_rp;  // multiply-high reciprocals of TW + 4, TW + 2, PH + 2, (PH + 2)(TW + 2), RP (backward; FastDiv)
};

__device__ __forceinline__ float silu_val(float v) { return v / (1.0f + __expf(-v)); }
__device__ __forceinline__ float silu_grad(float v) {
    const float s = 1.0f / (1.0f + __expf(-v));
    return s * (1.0f + v * (1.0f - s));
}

// patch coordinates of this CTA
struct DwJob {
    int b, h0, w0, c0, nc, patch;
    __device__ __forceinline__ DwJob(const DwArgs &a) {
        const int tiles_w = a.W / a.TW, tiles_h = a.H / a.PH, cblocks = (a.C + a.CB - 1) / a.CB;
        int t = blockIdx.x;
        const int cb = t % cblocks;
        t /= cblocks;
        patch = t;
        const int tw = t % tiles_w;
        t /= tiles_w;
        const int th = t % tiles_h;
        b = t / tiles_h;
        h0 = th * a.PH;
        w0 = tw * a.TW;
        c0 = cb * a.CB;
        nc = min(a.CB, a.C - c0);
    }
};

// i / d for i * d < 2^32 (every index of a patch is far below that): one multiply-high instead of the ~20-instruction
// division sequence; the loops below decode several flat indices per item.
static inline __host__ __device__ unsigned fastdiv_magic(unsigned d) { return d > 1 ? 0xFFFFFFFFu / d + 1u : 0u; }
struct FastDiv {
    unsigned d, m;
    __device__ __forceinline__ explicit FastDiv(unsigned d_) : d(d_), m(fastdiv_magic(d_)) {}
    __device__ __forceinline__ FastDiv(unsigned d_, unsigned m_) : d(d_), m(m_) {}  // reciprocal computed on the host
    __device__ __forceinline__ unsigned div(unsigned i) const { return d > 1 ? __umulhi(i, m) : i; }
};

// input patch with a halo of HALO positions (zeros outside the map) -> s[c][(PH + 2 HALO) x (TW + 2 HALO)], lanes along channels
template <typename T, int HALO>
__device__ __forceinline__ void load_patch(const DwArgs &a, const DwJob &j, float *s, int pitch) {
    const T *xin = static_cast<const T *>(a.xin);
    const int RW = a.TW + 2 * HALO, RH = a.PH + 2 * HALO;
    const int total = RH * RW * j.nc;
    const FastDiv dn(j.nc), drw(RW);
    constexpr int U = 8;  // loads of a batch are issued before the first store
    for (int i0 = threadIdx.x; i0 < total; i0 += 256 * U) {
        float v[U];
        int at[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int i = i0 + 256 * k;
            v[k] = 0.0f;
            at[k] = -1;
            if (i < total) {
                const int pos = dn.div(i), c = i - pos * j.nc;
                const int r = drw.div(pos), q = pos - r * RW;
                const int h = j.h0 - HALO + r, w = j.w0 - HALO + q;
                at[k] = c * pitch + pos;
                if (h >= 0 && h < a.H && w >= 0 && w < a.W) v[k] = to_f32<T>(xin[(((long long)j.b * a.H + h) * a.W + w) * a.ps + j.c0 + c]);
            }
        }
#pragma unroll
        for (int k = 0; k < U; ++k)
            if (at[k] >= 0) s[at[k]] = v[k];
    }
}

template <typename T>
__global__ void __launch_bounds__(256) dwconv_silu_fwd_kernel(const DwArgs a) {
    extern __shared__ float smem_dw[];
    const DwJob j(a);
    const int PH = a.PH, TW = a.TW, RW = TW + 2, CB = a.CB, nc = j.nc;
    const int pitch = ((PH + 2) * RW) | 1;
    const int P = PH * TW, opitch = P | 1, KR = 4 * a.RP;
    const int RQ = (2 * a.RP + 7) & ~7;       // x_dbl rows of one position order (two directions), padded to eights
    auto up4 = [](size_t v) { return (v + 3) & ~(size_t)3; };
    float *s = smem_dw;                                   // [CB][pitch]
    float *sw = s + up4((size_t)CB * pitch);              // [CB][10] weights and bias
    float *so = sw + up4((size_t)CB * 10);                // [CB][opitch] activations of the patch (x_proj only)
    float *sxw = so + up4((size_t)CB * opitch);           // [2][CB][RQ] x_proj weight slice, rows of a position order contiguous (x_proj only)
    const FastDiv dn(nc);
    for (int i = threadIdx.x; i < nc * 10; i += 256) {
        const int c = i / 10, k = i - c * 10;
        sw[i] = k < 9 ? __ldg(a.w9 + (long long)(j.c0 + c) * 9 + k) : (a.bias ? __ldg(a.bias + j.c0 + c) : 0.0f);
    }
    if (a.xpw) {
        for (int i = threadIdx.x; i < 2 * nc * RQ; i += 256) sxw[(i / (nc * RQ)) * CB * RQ + i % (nc * RQ)] = 0.0f;  // the padding rows
        __syncthreads();
        for (int i = threadIdx.x; i < KR * nc; i += 256) {
            const int kr = dn.div(i), c = i - kr * nc;
            const int k = kr / a.RP, r = kr - k * a.RP;
            sxw[((k & 1) * CB + c) * RQ + (k >> 1) * a.RP + r] = __ldg(a.xpw + (long long)kr * a.C + j.c0 + c);
        }
    }
    load_patch<T, 1>(a, j, s, pitch);
    __syncthreads();
    // item = (channel, quad of rows, column): four outputs down a column -> 32-byte row pieces of x, 128-bit pieces of x^T
    const long long L = (long long)a.H * a.W;
    const int twsh = 31 - __clz(TW), hqsh = PH == 8 ? 1 : 0;  // TW is a power of two, PH is 8 or 4 (plan_dw)
    const int items = nc << (hqsh + twsh);
    for (int i = threadIdx.x; i < items; i += 256) {
        const int c = i >> (hqsh + twsh), r = i & ((1 << (hqsh + twsh)) - 1);
        const int q = r >> twsh, w = r & (TW - 1);
        const float *wt = sw + c * 10;
        const float *src = s + c * pitch + (4 * q) * RW + w;  // top-left of the 6 x 3 window
        float wv[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) wv[k] = wt[k];
        float acc[4] = {wv[9], wv[9], wv[9], wv[9]};
#pragma unroll
        for (int rr = 0; rr < 6; ++rr) {
            const float v0 = src[rr * RW], v1 = src[rr * RW + 1], v2 = src[rr * RW + 2];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int kh = rr - k;  // output row k uses window rows k .. k + 2
                if (kh >= 0 && kh < 3) acc[k] = fmaf(wv[kh * 3 + 2], v2, fmaf(wv[kh * 3 + 1], v1, fmaf(wv[kh * 3], v0, acc[k])));
            }
        }
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = to_f32<T>(from_f32<T>(silu_val(to_f32<T>(from_f32<T>(acc[k])))));  // conv and act are tensors of the input's dtype
        const long long plane = ((long long)j.b * a.C + j.c0 + c) * L;
#pragma unroll
        for (int k = 0; k < 4; ++k) a.x[plane + (long long)(j.h0 + 4 * q + k) * a.W + j.w0 + w] = o[k];
        if (a.xT) *reinterpret_cast<float4 *>(a.xT + plane + (long long)(j.w0 + w) * a.H + j.h0 + 4 * q) = make_float4(o[0], o[1], o[2], o[3]);
        if (a.xpw) {
#pragma unroll
            for (int k = 0; k < 4; ++k) so[c * opitch + (4 * q + k) * TW + w] = o[k];
        }
    }
    if (!a.xpw) return;
    __syncthreads();
    // x_dbl rows of this channel block.  item = (position order, eight rows, position): one load of the activation and two
    // 128-bit broadcast loads of the weights feed eight FMAs per channel; even directions leave in row-major position order
    // (lanes along w), odd ones in column-major order (lanes along h)
    const bool single = a.C <= CB;
    const int chunks = RQ >> 3, phsh = PH == 8 ? 3 : 2;
    const int psh = twsh + phsh;  // P = PH * TW
    for (int i = threadIdx.x; i < 2 * chunks * P; i += 256) {
        const int qpos = i & (P - 1), pc = i >> psh;
        const int par = pc / chunks, ch = pc - par * chunks;
        int ph, pw;
        if (par) {
            pw = qpos >> phsh;
            ph = qpos & (PH - 1);
        } else {
            ph = qpos >> twsh;
            pw = qpos & (TW - 1);
        }
        const float *act = so + ph * TW + pw;
        const float4 *wk = reinterpret_cast<const float4 *>(sxw + (size_t)par * CB * RQ + ch * 8);
        float acc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = 0.0f;
        for (int c = 0; c < nc; ++c) {
            const float v = act[c * opitch];
            const float4 w0 = wk[c * (RQ >> 2)], w1 = wk[c * (RQ >> 2) + 1];
            acc[0] = fmaf(w0.x, v, acc[0]); acc[1] = fmaf(w0.y, v, acc[1]); acc[2] = fmaf(w0.z, v, acc[2]); acc[3] = fmaf(w0.w, v, acc[3]);
            acc[4] = fmaf(w1.x, v, acc[4]); acc[5] = fmaf(w1.y, v, acc[5]); acc[6] = fmaf(w1.z, v, acc[6]); acc[7] = fmaf(w1.w, v, acc[7]);
        }
        const long long pofs = par ? (long long)(j.w0 + pw) * a.H + j.h0 + ph : (long long)(j.h0 + ph) * a.W + j.w0 + pw;
        float *base = par ? a.xd_cm : a.xd_rm;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int jr = ch * 8 + u;  // (direction of this order, row)
            if (jr < 2 * a.RP) {
                const int kd = jr / a.RP, r = jr - kd * a.RP;
                float v = acc[u];
                if (j.c0 == 0 && a.xpb) v += __ldg(a.xpb + (2 * kd + par) * a.RP + r);
                float *dst = base + (((long long)j.b * 2 + kd) * a.RP + r) * L + pofs;
                if (single) *dst = v;
                else atomicAdd(dst, v);
            }
        }
    }
}

// shared-memory carve-up of the backward (floats); host and device agree through this one function
struct DwBwdLayout {
    int xpitch, dpitch, apitch;
    size_t sx, sd, sw, sacc, sxw, sdx, sxacc, sa, total;
    __host__ __device__ DwBwdLayout(int CB, int PH, int TW, int RP) {
        xpitch = ((PH + 4) * (TW + 4)) | 1;
        dpitch = ((PH + 2) * (TW + 2)) | 1;
        apitch = (PH * TW) | 1;
        const int KR = 4 * RP;
        auto up = [](size_t v) { return (v + 3) & ~(size_t)3; };
        size_t o = 0;
        sx = o; o = up(o + (size_t)CB * xpitch);
        sd = o; o = up(o + (size_t)CB * dpitch);
        sw = o; o = up(o + (size_t)CB * 10);
        sacc = o; o = up(o + (size_t)CB * 10);
        sxw = o; o = up(o + (size_t)KR * CB);
        sdx = o; o = up(o + (size_t)KR * dpitch);
        sxacc = o; o = up(o + (size_t)KR * CB);
        sa = o; o = up(o + (RP > 0 ? (size_t)CB * apitch : 0));
        total = o;
    }
};

// Backward.  pre = conv(xin) + bias (recomputed on the patch + 1), g = d x + transpose(d x^T) + W^T d x_dbl, d pre = g * silu'(pre);
//   d xin[h, w] = sum_k w[kh, kw] d pre[h - kh + 1, w - kw + 1];  d w[kh, kw] = sum d pre[h, w] xin[h + kh - 1, w + kw - 1];  d bias = sum d pre
// Four phases between three barriers; every phase hands each thread SEVERAL independent items per trip (loads first, then the
// stores), flat indices are decoded with multiply-high, and each global tensor is read with lanes along ITS fast axis:
//   A  xin patch (halo 2; lanes along channels), d x^T (halo 1; lanes along h), d x_dbl rows (even directions lanes along w, odd
//      ones along h) -> shared memory
//   B  two channels per thread: pre, g (d x read here, lanes along w), d pre -> sd; the activation of owned positions -> sa
//   C  reductions over the positions the patch OWNS: a unit is one channel's 9 + 1 stencil sums or four rows of d x_proj_weight
//      for one channel (one load of d pre / the activation feeds 10 / 4 FMAs), positions split over S slices so that every
//      thread has a unit; slices meet through shared-memory atomics
//   D  d xin (lanes along channels)
template <typename T>
__global__ void __launch_bounds__(256) dwconv_silu_bwd_kernel(const DwArgs a) {
    extern __shared__ float smem_dw[];
    const DwJob j(a);
    const int PH = a.PH, TW = a.TW, P = PH * TW, CB = a.CB, nc = j.nc;
    const int XW = TW + 4, XH = PH + 4, DW = TW + 2, DH = PH + 2, DP = DH * DW;
    const DwBwdLayout lay(CB, PH, TW, a.RP);
    const int xpitch = lay.xpitch, dpitch = lay.dpitch, apitch = lay.apitch;
    float *sx = smem_dw + lay.sx;        // [CB][xpitch]  input patch, halo 2
    float *sd = smem_dw + lay.sd;        // [CB][dpitch]  d x^T, then d pre, halo 1
    float *sw = smem_dw + lay.sw;        // [CB][10]
    float *sacc = smem_dw + lay.sacc;    // [CB][10] d weight / d bias of this patch
    const int KR = 4 * a.RP;
    float *sxw = smem_dw + lay.sxw;      // [KR][CB] x_proj weight slice (x_proj only)
    float *sdx = smem_dw + lay.sdx;      // [KR][dpitch] d x_dbl on the patch + 1 (x_proj only)
    float *sxacc = smem_dw + lay.sxacc;  // [KR][CB] d x_proj weight of this patch (x_proj only)
    float *sa = smem_dw + lay.sa;        // [CB][apitch] activation on the owned positions (x_proj only)
    const FastDiv dn(nc), dxw(XW, a.m_xw), ddw(DW, a.m_dw), ddh(DH, a.m_dh), ddp(DP, a.m_dp);
    const int tid = threadIdx.x;
    const int twsh = 31 - __clz(TW);     // TW is a power of two (plan_dw)
    const long long L = (long long)a.H * a.W;
    const bool xp = a.xpw != nullptr;

    // ---- A ----
    for (int i = tid; i < nc * 10; i += 256) {
        const int c = i / 10, k = i - c * 10;
        sw[i] = k < 9 ? __ldg(a.w9 + (long long)(j.c0 + c) * 9 + k) : (a.bias ? __ldg(a.bias + j.c0 + c) : 0.0f);
        sacc[i] = 0.0f;
    }
    if (xp) {
        for (int i = tid; i < KR * nc; i += 256) {
            const int kr = dn.div(i), c = i - kr * nc;
            sxw[kr * CB + c] = __ldg(a.xpw + (long long)kr * a.C + j.c0 + c);
            sxacc[kr * CB + c] = 0.0f;
        }
        if (nc & 1)
            for (int kr = tid; kr < KR; kr += 256) sxw[kr * CB + nc] = 0.0f;  // phase B reads channel pairs
    }
    {
        constexpr int U = 8;
        const T *xin = static_cast<const T *>(a.xin);
        const int total = XH * XW * nc;
        for (int i0 = tid; i0 < total; i0 += 256 * U) {
            float v[U];
            int at[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const int i = i0 + 256 * k;
                v[k] = 0.0f;
                at[k] = -1;
                if (i < total) {
                    const int pos = dn.div(i), c = i - pos * nc;
                    const int r = dxw.div(pos), q = pos - r * XW;
                    const int h = j.h0 - 2 + r, w = j.w0 - 2 + q;
                    at[k] = c * xpitch + pos;
                    if (h >= 0 && h < a.H && w >= 0 && w < a.W) v[k] = to_f32<T>(xin[(((long long)j.b * a.H + h) * a.W + w) * a.ps + j.c0 + c]);
                }
            }
#pragma unroll
            for (int k = 0; k < U; ++k)
                if (at[k] >= 0) sx[at[k]] = v[k];
        }
    }
    {   // d x^T on the patch + 1, lanes along h (its fast axis); zeros outside the map or when there is no d x^T
        constexpr int U = 6;
        const int total = nc * DP;
        for (int i0 = tid; i0 < total; i0 += 256 * U) {
            float v[U];
            int at[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const int i = i0 + 256 * k;
                v[k] = 0.0f;
                at[k] = -1;
                if (i < total) {
                    const int c = ddp.div(i), rem = i - c * DP;
                    const int dq = ddh.div(rem), dr = rem - dq * DH;
                    const int h = j.h0 - 1 + dr, w = j.w0 - 1 + dq;
                    at[k] = c * dpitch + dr * DW + dq;
                    if (a.dxT && h >= 0 && h < a.H && w >= 0 && w < a.W) v[k] = __ldg(a.dxT + ((long long)j.b * a.C + j.c0 + c) * L + (long long)w * a.H + h);
                }
            }
#pragma unroll
            for (int k = 0; k < U; ++k)
                if (at[k] >= 0) sd[at[k]] = v[k];
        }
    }
    if (xp) {
        constexpr int U = 4;
        const int total = KR * DP;
        const FastDiv drp(a.RP, a.m_rp);
        for (int i0 = tid; i0 < total; i0 += 256 * U) {
            float v[U];
            int at[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const int i = i0 + 256 * k;
                v[k] = 0.0f;
                at[k] = -1;
                if (i < total) {
                    const int kr = ddp.div(i), qpos = i - kr * DP;
                    const int kd = drp.div(kr), r = kr - kd * a.RP;
                    int dr, dq;
                    if (kd & 1) {  // column-major source: lanes along h
                        dq = ddh.div(qpos);
                        dr = qpos - dq * DH;
                    } else {
                        dr = ddw.div(qpos);
                        dq = qpos - dr * DW;
                    }
                    const int h = j.h0 - 1 + dr, w = j.w0 - 1 + dq;
                    at[k] = kr * dpitch + dr * DW + dq;
                    if (h >= 0 && h < a.H && w >= 0 && w < a.W) {
                        const long long row = (((long long)j.b * 2 + (kd >> 1)) * a.RP + r) * L;
                        v[k] = (kd & 1) ? __ldg(a.dxd_cm + row + (long long)w * a.H + h) : __ldg(a.dxd_rm + row + (long long)h * a.W + w);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < U; ++k)
                if (at[k] >= 0) sdx[at[k]] = v[k];
        }
    }
    __syncthreads();

    // ---- B ----  item = (channel pair, position of the patch + 1), lanes along w
    {
        constexpr int U = 2;
        const int npairs = (nc + 1) >> 1, total = npairs * DP;
        for (int i0 = tid; i0 < total; i0 += 256 * U) {
            float gx[U][2];
            int cp[U], rr[U];
            bool in[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const int i = i0 + 256 * k;
                gx[k][0] = gx[k][1] = 0.0f;
                in[k] = false;
                cp[k] = -1;
                rr[k] = 0;
                if (i < total) {
                    cp[k] = ddp.div(i);
                    rr[k] = i - cp[k] * DP;
                    const int dr = ddw.div(rr[k]), dq = rr[k] - dr * DW;
                    const int h = j.h0 - 1 + dr, w = j.w0 - 1 + dq;
                    in[k] = h >= 0 && h < a.H && w >= 0 && w < a.W;
                    if (in[k]) {
                        const long long plane = ((long long)j.b * a.C + j.c0 + 2 * cp[k]) * L + (long long)h * a.W + w;
                        gx[k][0] = __ldg(a.dx + plane);
                        if (2 * cp[k] + 1 < nc) gx[k][1] = __ldg(a.dx + plane + L);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < U; ++k) {
                if (cp[k] < 0) continue;
                const int r = rr[k], c = 2 * cp[k];
                const bool two = c + 1 < nc;
                const int dr = ddw.div(r), dq = r - dr * DW;
                float dp0 = 0.0f, dp1 = 0.0f;
                if (in[k]) {
                    const float *w0 = sw + c * 10, *w1 = w0 + (two ? 10 : 0);
                    const float *s0 = sx + c * xpitch + dr * XW + dq, *s1 = s0 + (two ? xpitch : 0);  // window of (h, w) in the halo-2 patch
                    float pre0 = w0[9], pre1 = w1[9];
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) {
                            pre0 = fmaf(w0[kh * 3 + kw], s0[kh * XW + kw], pre0);
                            pre1 = fmaf(w1[kh * 3 + kw], s1[kh * XW + kw], pre1);
                        }
                    pre0 = to_f32<T>(from_f32<T>(pre0));
                    pre1 = to_f32<T>(from_f32<T>(pre1));
                    float g0 = gx[k][0] + sd[c * dpitch + r], g1 = gx[k][1] + (two ? sd[(c + 1) * dpitch + r] : 0.0f);
                    if (xp) {  // x_dbl = W x: its gradient reaches the map through W^T
                        const float *wk = sxw + c, *dk = sdx + r;
                        for (int kr = 0; kr < KR; ++kr) {
                            const float d = dk[kr * dpitch];
                            g0 = fmaf(wk[kr * CB], d, g0);
                            g1 = fmaf(wk[kr * CB + 1], d, g1);
                        }
                    }
                    const float sg0 = 1.0f / (1.0f + __expf(-pre0)), sg1 = 1.0f / (1.0f + __expf(-pre1));
                    dp0 = g0 * sg0 * (1.0f + pre0 * (1.0f - sg0));
                    dp1 = g1 * sg1 * (1.0f + pre1 * (1.0f - sg1));
                    if (xp && dr >= 1 && dr <= PH && dq >= 1 && dq <= TW) {  // owned: keep the activation for d x_proj_weight
                        const int p = (dr - 1) * TW + dq - 1;
                        sa[c * apitch + p] = to_f32<T>(from_f32<T>(silu_val(pre0)));
                        if (two) sa[(c + 1) * apitch + p] = to_f32<T>(from_f32<T>(silu_val(pre1)));
                    }
                }
                sd[c * dpitch + r] = dp0;
                if (two) sd[(c + 1) * dpitch + r] = dp1;
            }
        }
    }
    __syncthreads();

    // ---- C ----  unit = (slice of the owned positions, kind, channel), channel fastest (odd pitches: conflict-free)
    {
        const int kinds = 1 + (xp ? (KR + 3) / 4 : 0);
        const int per = kinds * nc;
        int S = 1;
        while (S < 32 && per * S < 256 && (P >> 1) >= S * 8) S <<= 1;
        const FastDiv dper(per);
        for (int unit = tid; unit < per * S; unit += 256) {
            const int sl = dper.div(unit), rest = unit - sl * per;
            const int kind = dn.div(rest), c = rest - kind * nc;
            if (kind == 0) {
                float acc[10];
#pragma unroll
                for (int k = 0; k < 10; ++k) acc[k] = 0.0f;
                for (int p = sl; p < P; p += S) {
                    const int ph = p >> twsh, pw = p & (TW - 1);
                    const float dp = sd[c * dpitch + (ph + 1) * DW + pw + 1];
                    const float *src = sx + c * xpitch + (ph + 1) * XW + pw + 1;  // window of the owned position
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) acc[kh * 3 + kw] = fmaf(dp, src[kh * XW + kw], acc[kh * 3 + kw]);
                    acc[9] += dp;
                }
#pragma unroll
                for (int k = 0; k < 10; ++k) atomicAdd(&sacc[c * 10 + k], acc[k]);
            } else {
                const int kr0 = (kind - 1) * 4;
                float aw[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                for (int p = sl; p < P; p += S) {
                    const int ph = p >> twsh, pw = p & (TW - 1);
                    const float act = sa[c * apitch + p];
                    const float *dk = sdx + (ph + 1) * DW + pw + 1;
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (kr0 + u < KR) aw[u] = fmaf(dk[(kr0 + u) * dpitch], act, aw[u]);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (kr0 + u < KR) atomicAdd(&sxacc[(kr0 + u) * CB + c], aw[u]);
            }
        }
    }
    // ---- D ----  d input, channel-last: lanes along channels
    {
        constexpr int U = 4;
        T *dxin = static_cast<T *>(a.dxin);
        const int total = P * nc;
        for (int i0 = tid; i0 < total; i0 += 256 * U) {
            float acc[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const int i = i0 + 256 * k;
                acc[k] = 0.0f;
                if (i < total) {
                    const int pos = dn.div(i), c = i - pos * nc;
                    const int ph = pos >> twsh, pw = pos & (TW - 1);
                    const float *wt = sw + c * 10;
                    const float *src = sd + c * dpitch + ph * DW + pw;  // d pre rows ph .. ph + 2 (halo 1) = positions h - 1 .. h + 1
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) acc[k] = fmaf(wt[kh * 3 + kw], src[(2 - kh) * DW + (2 - kw)], acc[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const int i = i0 + 256 * k;
                if (i < total) {
                    const int pos = dn.div(i), c = i - pos * nc;
                    const int ph = pos >> twsh, pw = pos & (TW - 1);
                    dxin[(((long long)j.b * a.H + j.h0 + ph) * a.W + j.w0 + pw) * a.C + j.c0 + c] = from_f32<T>(acc[k]);
                }
            }
        }
    }
    __syncthreads();
    if (a.dwb) {
        float *dst = a.dwb + ((long long)j.patch * a.C + j.c0) * 10;
        for (int i = tid; i < nc * 10; i += 256) dst[i] = sacc[i];
    }
    if (xp)
        for (int i = tid; i < KR * nc; i += 256) {
            const int kr = dn.div(i), c = i - kr * nc;
            a.dxpw[((long long)j.patch * KR + kr) * a.C + j.c0 + c] = sxacc[kr * CB + c];
        }
}

// patch (PH x TW positions) x channel block: about 4 K outputs per CTA, shared memory under 100 KB, grid large enough
static int plan_dw(int batch, int C, int H, int W, bool bwd, int RP, int &PH, int &TW, int &CB, size_t &smem) {
    if (H % 4 || W % 8) return fail("dwconv_silu: H must be a multiple of 4 and W a multiple of 8 (got %d x %d)", H, W);
    PH = (H % 8 == 0) ? 8 : 4;
    CB = C < 64 ? C : 64;
    TW = 8;
    while (TW * 2 <= W && W % (TW * 2) == 0 && (long long)PH * TW * 2 * CB <= 4096) TW *= 2;
    auto grid = [&]() { return (long long)batch * (H / PH) * (W / TW) * ((C + CB - 1) / CB); };
    while (TW > 8 && grid() < 2 * 148) TW /= 2;
    while (CB > 16 && grid() < 2 * 148) CB /= 2;
    if (bwd) {  // the backward keeps four planes per channel: stay under ~56 KB so that four CTAs share an SM
        auto fits = [&](int cb, int tw, int rp) { return sizeof(float) * DwBwdLayout(cb, PH, tw, rp).total <= 56 * 1024; };
        // TW fixes the patch count, which vmasr_dwconv_patches reports without knowing RP: sized for dt_rank 1 whatever RP is
        while (CB > 16 && !fits(CB, TW, 3)) CB /= 2;
        while (TW > 8 && !fits(CB, TW, 3)) TW /= 2;
        while (CB > 16 && !fits(CB, TW, RP)) CB /= 2;  // more x_dbl rows
    }
    const int halo = bwd ? 2 : 1;
    const size_t xp = (size_t)((PH + 2 * halo) * (TW + 2 * halo)) | 1, dp = (size_t)((PH + 2) * (TW + 2)) | 1;
    smem = sizeof(float) * ((size_t)CB * xp + (size_t)CB * 20);
    if (RP > 0) smem += sizeof(float) * ((size_t)2 * CB * ((2 * RP + 7) & ~7) + (size_t)CB * (((size_t)PH * TW) | 1));
    smem += 64;  // the regions start at multiples of 16 bytes
    if (bwd) smem = sizeof(float) * DwBwdLayout(CB, PH, TW, RP).total;
    (void)dp;
    if (smem > 200 * 1024) return fail("dwconv_silu: patch does not fit shared memory");
    return 0;
}

template <typename T>
static int launch_dw(const DwArgs &a, bool bwd, size_t smem, long long grid, cudaStream_t stream) {
    auto kernel = bwd ? dwconv_silu_bwd_kernel<T> : dwconv_silu_fwd_kernel<T>;
    if (smem > 48 * 1024)
        if (int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "dwconv_silu smem attribute")) return rc;
    static PerDeviceOnce carved[2];  // the driver otherwise picks the smallest carve-out that holds ONE CTA
    if (!carved[bwd]()) {
        if (int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), "dwconv_silu carve-out")) return rc;
        carved[bwd]() = true;
    }
    kernel<<<(unsigned)grid, 256, smem, stream>>>(a);
    return check_cuda(cudaGetLastError(), bwd ? "dwconv_silu_bwd launch" : "dwconv_silu_fwd launch");
}

static int dw_run(const vmasr_dwconv_params *p, bool bwd) {
    const char *who = bwd ? "dwconv_silu_bwd" : "dwconv_silu_fwd";
    if (!p) return fail("%s: null params", who);
    if (p->batch <= 0 || p->channels <= 0 || p->H <= 0 || p->W <= 0) return fail("%s: sizes must be positive", who);
    if (p->io_dtype != VMASR_F32 && p->io_dtype != VMASR_F16 && p->io_dtype != VMASR_BF16) return fail("%s: unknown io_dtype %d", who, p->io_dtype);
    if (!p->xin || !p->weight) return fail("%s: xin and weight must be non-null", who);
    if (p->xin_pos_stride != 0 && p->xin_pos_stride < p->channels) return fail("%s: xin_pos_stride must be 0 (contiguous) or >= channels", who);
    if (!bwd) {
        if (!p->x) return fail("%s: x must be non-null", who);
        if (reinterpret_cast<uintptr_t>(p->xT) & 15u) return fail("%s: xT must be 16-byte aligned", who);
    } else {
        if (!p->dx || !p->dxin) return fail("%s: dx and dxin must be non-null", who);
    }
    DwArgs a{};
    size_t smem = 0;
    const int RP = p->x_proj_weight ? p->x_proj_rows : 0;
    if (p->x_proj_weight) {
        if (RP < 1 || RP > 64) return fail("%s: x_proj_rows must be in [1, 64]", who);
        if (!bwd && (!p->x_dbl_rm || !p->x_dbl_cm)) return fail("%s: x_dbl_rm and x_dbl_cm must be given with x_proj_weight", who);
        if (bwd && (!p->d_x_dbl_rm || !p->d_x_dbl_cm || !p->d_x_proj_weight_partial))
            return fail("%s: d_x_dbl_rm, d_x_dbl_cm and d_x_proj_weight_partial must be given with x_proj_weight", who);
    }
    if (int rc = plan_dw(p->batch, p->channels, p->H, p->W, bwd, RP, a.PH, a.TW, a.CB, smem)) return rc;
    const long long grid = (long long)p->batch * (p->H / a.PH) * (p->W / a.TW) * ((p->channels + a.CB - 1) / a.CB);
    if (grid > 0x7fffffffLL) return fail("%s: too many patches", who);
    a.xin = p->xin; a.w9 = p->weight; a.bias = p->bias; a.x = p->x; a.xT = p->xT;
    a.dx = p->dx; a.dxT = p->dxT; a.dxin = p->dxin; a.dwb = p->dwb_partial;
    a.xpw = p->x_proj_weight; a.xpb = p->x_proj_bias; a.xd_rm = p->x_dbl_rm; a.xd_cm = p->x_dbl_cm;
    a.dxd_rm = p->d_x_dbl_rm; a.dxd_cm = p->d_x_dbl_cm; a.dxpw = p->d_x_proj_weight_partial; a.RP = RP;
    a.ps = p->xin_pos_stride ? p->xin_pos_stride : p->channels;
    a.B = p->batch; a.C = p->channels; a.H = p->H; a.W = p->W;
    a.m_xw = fastdiv_magic(a.TW + 4); a.m_dw = fastdiv_magic(a.TW + 2); a.m_dh = fastdiv_magic(a.PH + 2);
    a.m_dp = fastdiv_magic((a.PH + 2) * (a.TW + 2)); a.m_rp = fastdiv_magic(RP);
    DeviceGuard guard(p->device);
    if (!guard.ok) return fail("%s: cannot select CUDA device %d", who, p->device);
    cudaStream_t stream = static_cast<cudaStream_t>(p->stream);
    switch (p->io_dtype) {
        case VMASR_F32: return launch_dw<float>(a, bwd, smem, grid, stream);
        case VMASR_F16: return launch_dw<__half>(a, bwd, smem, grid, stream);
        default: return launch_dw<__nv_bfloat16>(a, bwd, smem, grid, stream);
    }
}

}  // namespace vmasr

extern "C" int64_t vmasr_dwconv_patches(int batch, int channels, int H, int W) {
    int PH = 0, TW = 0, CB = 0;
    size_t smem = 0;
    if (batch <= 0 || channels <= 0 || H <= 0 || W <= 0) return -1;
    if (vmasr::plan_dw(batch, channels, H, W, true, 0, PH, TW, CB, smem)) return -1;
    return (int64_t)batch * (H / PH) * (W / TW);
}
extern "C" int vmasr_dwconv_channel_blocks(int batch, int channels, int H, int W) {
    int PH = 0, TW = 0, CB = 0;
    size_t smem = 0;
    if (batch <= 0 || channels <= 0 || H <= 0 || W <= 0) return -1;
    if (vmasr::plan_dw(batch, channels, H, W, false, 0, PH, TW, CB, smem)) return -1;
    return (channels + CB - 1) / CB;
}
extern "C" int vmasr_dwconv_silu_fwd(const vmasr_dwconv_params *p) { return vmasr::dw_run(p, false); }
extern "C" int vmasr_dwconv_silu_bwd(const vmasr_dwconv_params *p) { return vmasr::dw_run(p, true); }
