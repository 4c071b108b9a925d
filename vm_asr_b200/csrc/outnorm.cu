// Tail of the SS2D block fused into the merge of the core's two planes:
//     y = P_rm + transpose(P_cm)                       outer addition of CrossMerge (model/vmamba.py:57-60), association kept
//     y = LayerNorm_C(y^T)  (out_norm, (B, L, C))      vmamba.py:1527-1529  (out_norm_shape "v0": nn.LayerNorm(d_inner), eps 1e-5)
//     y = y.to(x.dtype)                                :1531
//     out = y * act(z)                                 forwardv2, :1536-1550  (z = SiLU(z) unless disable_z_act)
// The reference runs this as merge store, transpose(1, 2).contiguous(), LayerNorm, cast, SiLU and the product: about ten
// passes over a (B, C, L) map.  Here one kernel reads the two planes and z and writes out (and, for training, the merged map
// and the row statistics); its backward reads dout, z and the saved map and writes dy (row-major, what the core's backward
// takes), dz and per-tile partial sums of d gamma / d beta.
//
// Three memory orders meet: P_rm has w fastest, P_cm has h fastest, z / out have c fastest.  A CTA owns a patch of PH x TW
// positions (PH = 8 or 4 rows, TW = 8 * 2^k columns chosen so that a patch holds about 4 K elements) of ALL channels in
// shared memory, pitch P + 1 floats per channel: column accesses (lanes along positions) and row accesses (lanes along
// channels) are both conflict-free.  Global accesses: 128-bit along h for P_cm, 32-byte row pieces for P_rm / y / dy, fully
// coalesced runs of TW * C elements for z / out / dout / dz.  LayerNorm statistics are two-pass (mean, then squared
// deviations) over the shared-memory copy.  Bound: HBM (4 map passes forward, 6 backward); no tensor cores.
#include "common.cuh"

namespace vmasr {

struct OutNormArgs {
    const float *p_rm, *p_cm, *gamma, *beta;
    const void *z, *dout;
    void *out, *dz;
    float *y, *stats, *dy, *dgb;
    const float *y_in, *stats_in;
    float eps;
    int B, C, H, W, PH, TW, z_silu, c_shift;  // c_shift = log2(C) when C is a power of two, else -1
};

__device__ __forceinline__ float silu_f(float v, float &sig) {
    sig = 1.0f / (1.0f + __expf(-v));
    return v * sig;
}

// element index inside a patch row -> (pw, c)
__device__ __forceinline__ void split_pc(int e, int C, int c_shift, int &pw, int &c) {
    if (c_shift >= 0) {
        pw = e >> c_shift;
        c = e & (C - 1);
    } else {
        pw = e / C;
        c = e - pw * C;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) outnorm_fwd_kernel(const OutNormArgs a) {
    extern __shared__ float smem_on[];
    const int C = a.C, H = a.H, W = a.W, PH = a.PH, TW = a.TW;
    const int P = PH * TW, PITCH = P + 1;
    float *sy = smem_on;                      // [C][PITCH]
    float *sred = sy + (size_t)C * PITCH;     // [256]
    float *smean = sred + 256;                // [P]
    float *srstd = smean + P;                 // [P]
    const long long L = (long long)H * W;
    const int tiles_w = W / TW, tiles_h = H / PH;
    int tile = blockIdx.x;
    const int tw = tile % tiles_w;
    tile /= tiles_w;
    const int th = tile % tiles_h;
    const int b = tile / tiles_h;
    const int h0 = th * PH, w0 = tw * TW;
    const int t = threadIdx.x;

    // ---- phase 1: merge the two planes into the patch (and keep the merged map for the backward) ----
    {
        const int HQ = PH / 4, TPC = HQ * TW;  // items per channel: h-quads x columns
        const int items = C * TPC;
        constexpr int U = 4;  // items in flight per thread: every load of a batch is issued before the first use
        for (int i0 = t; i0 < items; i0 += 256 * U) {
            float4 cm[U];
            float rm[U][4];
            long long off0[U];
            int sidx[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + 256 * u;
                if (i < items) {
                    const int c = i / TPC, r = i - c * TPC;
                    const int q = r / TW, w = r - q * TW;
                    const long long plane = ((long long)b * C + c) * L;
                    cm[u] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    if (a.p_cm) cm[u] = __ldg(reinterpret_cast<const float4 *>(a.p_cm + plane + (long long)(w0 + w) * H + h0 + 4 * q));
                    off0[u] = plane + (long long)(h0 + 4 * q) * W + w0 + w;
                    sidx[u] = c * PITCH + 4 * q * TW + w;
#pragma unroll
                    for (int k = 0; k < 4; ++k) rm[u][k] = __ldg(a.p_rm + off0[u] + (long long)k * W);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (i0 + 256 * u < items) {
                    const float v[4] = {cm[u].x, cm[u].y, cm[u].z, cm[u].w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float yv = a.p_cm ? __fadd_rn(rm[u][k], v[k]) : rm[u][k];
                        sy[sidx[u] + k * TW] = yv;
                        if (a.y) a.y[off0[u] + (long long)k * W] = yv;
                    }
                }
            }
        }
    }
    __syncthreads();
    // ---- phase 2: mean and 1 / sqrt(var + eps) over the channels of every position (two passes) ----
    const int parts = P >= 256 ? 1 : 256 / P;  // P is a power of two or a multiple of 256 (host)
    for (int pass = 0; pass < 2; ++pass) {
        for (int p0 = 0; p0 < P; p0 += 256) {
            const int p = p0 + (parts == 1 ? t : t % P), part = parts == 1 ? 0 : t / P;
            float acc = 0.0f;
            if (p < P) {
                const float m = pass ? smean[p] : 0.0f;
                for (int c = part; c < C; c += parts) {
                    const float d = sy[(size_t)c * PITCH + p] - m;
                    acc += pass ? d * d : d;
                }
            }
            if (parts > 1) {
                sred[t] = acc;
                __syncthreads();
                if (t < P) {
                    acc = 0.0f;
                    for (int k = 0; k < parts; ++k) acc += sred[k * P + t];
                }
            }
            if ((parts == 1 && p < P) || (parts > 1 && t < P)) {
                const int pp = parts == 1 ? p : t;
                if (pass == 0) smean[pp] = acc / (float)C;
                else srstd[pp] = rsqrtf(acc / (float)C + a.eps);
            }
            __syncthreads();
        }
    }
    if (a.stats) {
        for (int p = t; p < P; p += 256) {
            const int ph = p / TW, pw = p - ph * TW;
            const long long pos = (long long)b * L + (long long)(h0 + ph) * W + w0 + pw;
            reinterpret_cast<float2 *>(a.stats)[pos] = make_float2(smean[p], srstd[p]);
        }
    }
    // ---- phase 3: normalise, cast, gate, store: runs of TW * C contiguous elements per patch row ----
    const T *z = static_cast<const T *>(a.z);
    T *out = static_cast<T *>(a.out);
    const int row_elems = TW * C;
    constexpr int U3 = 4;  // elements in flight per thread: the z loads of a batch are issued before the first use
    for (int ph = 0; ph < PH; ++ph) {
        const long long base = ((long long)b * L + (long long)(h0 + ph) * W + w0) * C;
        for (int e0 = t; e0 < row_elems; e0 += 256 * U3) {
            float zv[U3];
#pragma unroll
            for (int u = 0; u < U3; ++u) {
                const int e = e0 + 256 * u;
                zv[u] = 1.0f;
                if (z && e < row_elems) zv[u] = to_f32<T>(z[base + e]);
            }
#pragma unroll
            for (int u = 0; u < U3; ++u) {
                const int e = e0 + 256 * u;
                if (e >= row_elems) break;
                int pw, c;
                split_pc(e, C, a.c_shift, pw, c);
                const int p = ph * TW + pw;
                const float g = a.gamma ? __ldg(a.gamma + c) : 1.0f, be = a.beta ? __ldg(a.beta + c) : 0.0f;
                const float ln = fmaf((sy[(size_t)c * PITCH + p] - smean[p]) * srstd[p], g, be);
                float o = to_f32<T>(from_f32<T>(ln));  // y.to(x.dtype)
                if (z) {
                    float gate = zv[u];
                    if (a.z_silu) {
                        float sig;
                        gate = to_f32<T>(from_f32<T>(silu_f(gate, sig)));  // act(z) is a tensor of z's dtype in the reference
                    }
                    o *= gate;
                }
                out[base + e] = from_f32<T>(o);
            }
        }
    }
}

// Backward.  ln = xhat * gamma + beta, out = cast(ln) * g(z):
//   d ln = dout * g;  dz = dout * ln * g'(z);  d gamma = sum d ln * xhat;  d beta = sum d ln;
//   d xhat = d ln * gamma;  dy = rstd * (d xhat - mean_c(d xhat) - xhat * mean_c(d xhat * xhat))
template <typename T>
__global__ void __launch_bounds__(256) outnorm_bwd_kernel(const OutNormArgs a) {
    extern __shared__ float smem_on[];
    const int C = a.C, H = a.H, W = a.W, PH = a.PH, TW = a.TW;
    const int P = PH * TW, PITCH = P + 1;
    float *sx = smem_on;                       // [C][PITCH] xhat
    float *sd = sx + (size_t)C * PITCH;        // [C][PITCH] d xhat
    float *sred = sd + (size_t)C * PITCH;      // [512]
    float *smean = sred + 512;                 // [P] mean, then m1
    float *srstd = smean + P;                  // [P]
    float *sm2 = srstd + P;                    // [P]
    float *sgb = sm2 + P;                      // [2][C] d gamma, d beta of this patch
    const long long L = (long long)H * W;
    const int tiles_w = W / TW, tiles_h = H / PH;
    int tile = blockIdx.x;
    const int tile_id = tile;
    const int tw = tile % tiles_w;
    tile /= tiles_w;
    const int th = tile % tiles_h;
    const int b = tile / tiles_h;
    const int h0 = th * PH, w0 = tw * TW;
    const int t = threadIdx.x;

    for (int p = t; p < P; p += 256) {
        const int ph = p / TW, pw = p - ph * TW;
        const long long pos = (long long)b * L + (long long)(h0 + ph) * W + w0 + pw;
        const float2 st = __ldg(reinterpret_cast<const float2 *>(a.stats_in) + pos);
        smean[p] = st.x;
        srstd[p] = st.y;
    }
    for (int c = t; c < 2 * C; c += 256) sgb[c] = 0.0f;
    __syncthreads();
    // ---- phase 1: xhat of the patch from the saved merged map (rows of TW floats, 128-bit) ----
    {
        const int V = TW / 4, items = C * PH * V;
        constexpr int U = 4;
        for (int i0 = t; i0 < items; i0 += 256 * U) {
            float4 yv[U];
            int at[U], pp[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + 256 * u;
                at[u] = -1;
                pp[u] = 0;
                if (i < items) {
                    const int c = i / (PH * V), r = i - c * PH * V;
                    const int ph = r / V, wv = (r - ph * V) * 4;
                    yv[u] = __ldg(reinterpret_cast<const float4 *>(a.y_in + ((long long)b * C + c) * L + (long long)(h0 + ph) * W + w0 + wv));
                    pp[u] = ph * TW + wv;
                    at[u] = c * PITCH + pp[u];
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (at[u] < 0) continue;
                const int p = pp[u];
                float *dst = sx + at[u];
                dst[0] = (yv[u].x - smean[p]) * srstd[p];
                dst[1] = (yv[u].y - smean[p + 1]) * srstd[p + 1];
                dst[2] = (yv[u].z - smean[p + 2]) * srstd[p + 2];
                dst[3] = (yv[u].w - smean[p + 3]) * srstd[p + 3];
            }
        }
    }
    __syncthreads();
    // ---- phase 2: gate backward, d gamma / d beta, d xhat ----
    const T *z = static_cast<const T *>(a.z);
    const T *dout = static_cast<const T *>(a.dout);
    T *dz = static_cast<T *>(a.dz);
    const int row_elems = TW * C;
    // a thread walks elements t, t + 256, ...: its channel repeats with period K = C / gcd(C, 256) steps.  For the channel
    // counts of the configs (powers of two) K is 1 (C <= 256) or C / 256, and the thread's sums for d gamma / d beta stay in
    // registers, one shared-memory atomic per thread and channel at the end; any other C adds per element.
    const int K = (256 % C == 0) ? 1 : (C % 256 == 0 && C <= 1024 ? C / 256 : 0);
    for (int kk = 0; kk < (K ? K : 1); ++kk) {
    float acc_g = 0.0f, acc_b = 0.0f;
    const int step = K ? 256 * K : 256;
    for (int ph = 0; ph < PH; ++ph) {
        const long long base = ((long long)b * L + (long long)(h0 + ph) * W + w0) * C;
        constexpr int U2 = 4;  // elements in flight per thread
        for (int e0 = t + 256 * kk; e0 < row_elems; e0 += step * U2) {
            float gov[U2], zvv[U2];
#pragma unroll
            for (int u = 0; u < U2; ++u) {
                const int e = e0 + step * u;
                gov[u] = 0.0f;
                zvv[u] = 0.0f;
                if (e < row_elems) {
                    gov[u] = to_f32<T>(dout[base + e]);
                    if (z) zvv[u] = to_f32<T>(z[base + e]);
                }
            }
#pragma unroll
            for (int u = 0; u < U2; ++u) {
                const int e = e0 + step * u;
                if (e >= row_elems) break;
                int pw, c;
                split_pc(e, C, a.c_shift, pw, c);
                const int p = ph * TW + pw;
                const float g = a.gamma ? __ldg(a.gamma + c) : 1.0f, be = a.beta ? __ldg(a.beta + c) : 0.0f;
                const float xh = sx[(size_t)c * PITCH + p];
                const float ln = to_f32<T>(from_f32<T>(fmaf(xh, g, be)));
                const float go = gov[u];
                float dln = go;
                if (z) {
                    const float zv = zvv[u];
                    float gate = zv, dgate = 1.0f;
                    if (a.z_silu) {
                        float sig;
                        gate = to_f32<T>(from_f32<T>(silu_f(zv, sig)));
                        dgate = sig * (1.0f + zv * (1.0f - sig));
                    }
                    dln = go * gate;
                    if (dz) dz[base + e] = from_f32<T>(go * ln * dgate);
                }
                sd[(size_t)c * PITCH + p] = dln * g;
                if (K) {
                    acc_g = fmaf(dln, xh, acc_g);
                    acc_b += dln;
                } else {
                    atomicAdd(&sgb[c], dln * xh);
                    atomicAdd(&sgb[C + c], dln);
                }
            }
        }
    }
    if (K && t + 256 * kk < row_elems) {
        const int c = (t + 256 * kk) % C;
        atomicAdd(&sgb[c], acc_g);
        atomicAdd(&sgb[C + c], acc_b);
    }
    }
    __syncthreads();
    if (a.dgb) {
        float *dst = a.dgb + (long long)tile_id * 2 * C;
        for (int c = t; c < 2 * C; c += 256) dst[c] = sgb[c];
    }
    // ---- per-position sums over the channels: m1 = mean d xhat, m2 = mean d xhat * xhat ----
    const int parts = P >= 256 ? 1 : 256 / P;
    for (int p0 = 0; p0 < P; p0 += 256) {
        const int p = p0 + (parts == 1 ? t : t % P), part = parts == 1 ? 0 : t / P;
        float s1 = 0.0f, s2 = 0.0f;
        if (p < P)
            for (int c = part; c < C; c += parts) {
                const float d = sd[(size_t)c * PITCH + p];
                s1 += d;
                s2 = fmaf(d, sx[(size_t)c * PITCH + p], s2);
            }
        if (parts > 1) {
            sred[t] = s1;
            sred[256 + t] = s2;
            __syncthreads();
            if (t < P) {
                s1 = 0.0f;
                s2 = 0.0f;
                for (int k = 0; k < parts; ++k) {
                    s1 += sred[k * P + t];
                    s2 += sred[256 + k * P + t];
                }
            }
        }
        if ((parts == 1 && p < P) || (parts > 1 && t < P)) {
            const int pp = parts == 1 ? p : t;
            smean[pp] = s1 / (float)C;
            sm2[pp] = s2 / (float)C;
        }
        __syncthreads();
    }
    // ---- phase 3: dy rows ----
    {
        const int V = TW / 4, items = C * PH * V;
        for (int i = t; i < items; i += 256) {
            const int c = i / (PH * V), r = i - c * PH * V;
            const int ph = r / V, wv = (r - ph * V) * 4;
            const int p = ph * TW + wv;
            const float *xs = sx + (size_t)c * PITCH + p, *ds = sd + (size_t)c * PITCH + p;
            float4 o;
            o.x = srstd[p] * (ds[0] - smean[p] - xs[0] * sm2[p]);
            o.y = srstd[p + 1] * (ds[1] - smean[p + 1] - xs[1] * sm2[p + 1]);
            o.z = srstd[p + 2] * (ds[2] - smean[p + 2] - xs[2] * sm2[p + 2]);
            o.w = srstd[p + 3] * (ds[3] - smean[p + 3] - xs[3] * sm2[p + 3]);
            *reinterpret_cast<float4 *>(a.dy + ((long long)b * C + c) * L + (long long)(h0 + ph) * W + w0 + wv) = o;
        }
    }
}

// patch shape: PH rows x TW columns of all channels; TW the largest 8 * 2^k divisor of W that keeps the patch at <= 4096
// elements (P is then a power of two <= 256 or a multiple of 256, which the statistics pass relies on), halved while the grid
// would leave SMs without a CTA
static int plan_patch(int batch, int C, int H, int W, bool bwd, int &PH, int &TW, size_t &smem) {
    if (H % 4 || W % 8) return fail("outnorm: H must be a multiple of 4 and W a multiple of 8 (got %d x %d)", H, W);
    const size_t limit = 200 * 1024;
    const int arrays = bwd ? 2 : 1;
    for (int ph : {8, 4}) {
        if (H % ph) continue;
        int best = 0;
        for (int tw = 8; tw <= W; tw *= 2) {
            if (W % tw) break;
            const int P = ph * tw;
            if (P > 256 && P % 256) continue;
            const size_t bytes = sizeof(float) * ((size_t)arrays * C * (P + 1) + 512 + 3 * (size_t)P + 2 * (size_t)C);
            if (bytes > limit) break;
            if ((long long)P * C > 4096 && best) break;
            best = tw;
        }
        if (best) {
            while (best > 8 && (long long)batch * (H / ph) * (W / best) < 2 * 148) best /= 2;
            if (ph == 8 && best == 8 && H % 4 == 0 && (long long)batch * (H / 8) * (W / 8) < 2 * 148) ph = 4;
            PH = ph;
            TW = best;
            smem = sizeof(float) * ((size_t)arrays * C * (PH * TW + 1) + 512 + 3 * (size_t)PH * TW + 2 * (size_t)C);
            return 0;
        }
    }
    return fail("outnorm: %d channels do not fit a patch in shared memory", C);
}

template <typename T>
static int launch_outnorm(const OutNormArgs &a, bool bwd, size_t smem, long long grid, cudaStream_t stream) {
    auto kernel = bwd ? outnorm_bwd_kernel<T> : outnorm_fwd_kernel<T>;
    if (smem > 48 * 1024)
        if (int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "outnorm smem attribute")) return rc;
    kernel<<<(unsigned)grid, 256, smem, stream>>>(a);
    return check_cuda(cudaGetLastError(), bwd ? "outnorm_gate_bwd launch" : "outnorm_gate_fwd launch");
}

static int outnorm_run(const vmasr_outnorm_params *p, bool bwd) {
    const char *who = bwd ? "outnorm_gate_bwd" : "outnorm_gate_fwd";
    if (!p) return fail("%s: null params", who);
    if (p->batch <= 0 || p->channels <= 0 || p->H <= 0 || p->W <= 0) return fail("%s: sizes must be positive", who);
    if (p->io_dtype != VMASR_F32 && p->io_dtype != VMASR_F16 && p->io_dtype != VMASR_BF16) return fail("%s: unknown io_dtype %d", who, p->io_dtype);
    OutNormArgs a{};
    if (!bwd) {
        if (!p->p_rm || !p->out) return fail("%s: p_rm and out must be non-null", who);
        if ((reinterpret_cast<uintptr_t>(p->p_cm) | reinterpret_cast<uintptr_t>(p->stats)) & 15u) return fail("%s: p_cm and stats must be 16-byte aligned", who);
    } else {
        if (!p->dout || !p->y || !p->stats || !p->dy) return fail("%s: dout, y, stats, dy must be non-null", who);
        if ((reinterpret_cast<uintptr_t>(p->y) | reinterpret_cast<uintptr_t>(p->dy) | reinterpret_cast<uintptr_t>(p->stats)) & 15u)
            return fail("%s: y, dy and stats must be 16-byte aligned", who);
        if (p->z && !p->dz) return fail("%s: dz must be given when z is", who);
    }
    int PH = 0, TW = 0;
    size_t smem = 0;
    if (int rc = plan_patch(p->batch, p->channels, p->H, p->W, bwd, PH, TW, smem)) return rc;
    const long long grid = (long long)p->batch * (p->H / PH) * (p->W / TW);
    if (grid > 0x7fffffffLL) return fail("%s: too many patches", who);
    a.p_rm = p->p_rm; a.p_cm = p->p_cm; a.gamma = p->gamma; a.beta = p->beta;
    a.z = p->z; a.dout = p->dout; a.out = p->out; a.dz = p->dz;
    a.y = bwd ? nullptr : p->y; a.stats = bwd ? nullptr : p->stats; a.dy = p->dy; a.dgb = p->dgb_partial;
    a.y_in = p->y; a.stats_in = p->stats;
    a.eps = p->eps;
    a.B = p->batch; a.C = p->channels; a.H = p->H; a.W = p->W; a.PH = PH; a.TW = TW; a.z_silu = p->z_silu;
    a.c_shift = -1;
    for (int s = 0; s < 16; ++s)
        if ((1 << s) == p->channels) a.c_shift = s;
    DeviceGuard guard(p->device);
    if (!guard.ok) return fail("%s: cannot select CUDA device %d", who, p->device);
    cudaStream_t stream = static_cast<cudaStream_t>(p->stream);
    switch (p->io_dtype) {
        case VMASR_F32: return launch_outnorm<float>(a, bwd, smem, grid, stream);
        case VMASR_F16: return launch_outnorm<__half>(a, bwd, smem, grid, stream);
        default: return launch_outnorm<__nv_bfloat16>(a, bwd, smem, grid, stream);
    }
}

}  // namespace vmasr

extern "C" int64_t vmasr_outnorm_patches(int batch, int channels, int H, int W) {
    int PH = 0, TW = 0;
    size_t smem = 0;
    if (batch <= 0 || channels <= 0 || H <= 0 || W <= 0) return -1;
    if (vmasr::plan_patch(batch, channels, H, W, true, PH, TW, smem)) return -1;
    return (int64_t)batch * (H / PH) * (W / TW);
}
extern "C" int vmasr_outnorm_gate_fwd(const vmasr_outnorm_params *p) { return vmasr::outnorm_run(p, false); }
extern "C" int vmasr_outnorm_gate_bwd(const vmasr_outnorm_params *p) { return vmasr::outnorm_run(p, true); }
