// Magnitude/phase STFT, iSTFT and iSTFT-backward for sm_100a as shared-memory FFT kernels.
// Replace wav2spectro / spectro2wav (utils/stft.py:22-68, 71-115): torch.stft / torch.istft with
// normalized=True, center=True (reflect pad), periodic Hann of win_length zero-padded to n_fft, onesided,
// fused with log2(|X|+1e-8) / angle and exp2 / polar.  The reference runs ~8 library launches per call
// (pad, frame, window, cuFFT, scale, abs, log2, angle); here each transform is one kernel.
//
// Layout: a CTA of 8 warps handles 8 consecutive frames; each warp runs one real FFT of n_fft points as a
// complex FFT of M = n_fft/2 points (even/odd packing), in place in shared memory (digit-reversed load,
// radix-4 decimation-in-time stages, one radix-2 stage first when log2 M is odd).  The (F, n_frames) output
// planes have the frame index fastest, so the 8 frames of a CTA are staged in shared memory and leave (or
// enter) HBM as 32-byte rows.
#include "common.cuh"

namespace vmasr {

constexpr int kFramesPerCta = 8;
constexpr int kStageStride = 9;  // floats per k row of the staging planes (8 frames + 1 pad)

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// position of element n in the digit-reversed buffer (radix-4 digits, plus one radix-2 digit when log2 M odd)
__device__ __forceinline__ int perm_index(int n, int M, int log2m) {
    int p = 0, block = M, m = log2m;
    while (m >= 2) {
        block >>= 2;
        p += (n & 3) * block;
        n >>= 2;
        m -= 2;
    }
    if (m == 1) p += (n & 1);
    return p;
}

// in-place FFT of M complex points by one warp; tw[t] = exp(-2 pi i t / M)
__device__ __forceinline__ void warp_fft(float2 *buf, const float2 *tw, int M, int log2m, int lane) {
    int span = 1;
    if (log2m & 1) {
        for (int t = lane; t < (M >> 1); t += 32) {
            const float2 a = buf[2 * t], b = buf[2 * t + 1];
            buf[2 * t] = cadd(a, b);
            buf[2 * t + 1] = csub(a, b);
        }
        span = 2;
        __syncwarp();
    }
    while (span < M) {
        const int step = M / (span * 4);
        for (int t = lane; t < (M >> 2); t += 32) {
            const int kp = t & (span - 1);
            const int base = (t - kp) * 4 + kp;
            const float2 a = buf[base];
            const float2 b = cmul(buf[base + span], tw[kp * step]);
            const float2 c = cmul(buf[base + 2 * span], tw[2 * kp * step]);
            const float2 d = cmul(buf[base + 3 * span], tw[3 * kp * step]);
            const float2 s0 = cadd(a, c), s1 = csub(a, c), s2 = cadd(b, d);
            const float2 bd = csub(b, d);
            const float2 s3 = make_float2(bd.y, -bd.x);  // (b - d) * (-i)
            buf[base] = cadd(s0, s2);
            buf[base + span] = cadd(s1, s3);
            buf[base + 2 * span] = csub(s0, s2);
            buf[base + 3 * span] = csub(s1, s3);
        }
        span *= 4;
        __syncwarp();
    }
}

struct StftShape {
    int B, T, n_fft, hop, win, n_frames, M, log2m, F;
    int t_out;  // hop * (n_frames - 1), the iSTFT output length
};

// shared-memory carve-up common to the three kernels
struct Tables {
    float2 *tw;   // [M]    exp(-2 pi i t / M)
    float2 *tw2;  // [M+1]  exp(-pi i k / M)
    float *wtab;  // [n_fft] padded periodic Hann * n_fft^-0.5
};

__device__ __forceinline__ void build_tables(const Tables &tb, const StftShape &s) {
    const float inv_m = 1.0f / (float)s.M;
    for (int t = threadIdx.x; t < s.M; t += blockDim.x) {
        float sn, cs;
        sincospif(2.0f * (float)t * inv_m, &sn, &cs);
        tb.tw[t] = make_float2(cs, -sn);
    }
    for (int k = threadIdx.x; k <= s.M; k += blockDim.x) {
        float sn, cs;
        sincospif((float)k * inv_m, &sn, &cs);
        tb.tw2[k] = make_float2(cs, -sn);
    }
    const int left = (s.n_fft - s.win) / 2;
    const float scale = rsqrtf((float)s.n_fft);
    for (int n = threadIdx.x; n < s.n_fft; n += blockDim.x) {
        float w = 0.0f;
        if (n >= left && n < left + s.win) w = 0.5f - 0.5f * cospif(2.0f * (float)(n - left) / (float)s.win);
        tb.wtab[n] = w * scale;
    }
}

// sum over the frames covering padded sample tp of window^2 (torch.istft's envelope)
__device__ __forceinline__ float envelope(const float *wtab, const StftShape &s, int tp) {
    int f_lo = (tp - s.n_fft + s.hop) / s.hop;  // ceil((tp - n_fft + 1) / hop) for tp - n_fft + 1 > 0
    if (tp - s.n_fft + 1 <= 0) f_lo = 0;
    int f_hi = tp / s.hop;
    if (f_hi > s.n_frames - 1) f_hi = s.n_frames - 1;
    float e = 0.0f;
    for (int f = f_lo; f <= f_hi; ++f) {
        const float w = wtab[tp - f * s.hop];
        e = fmaf(w, w, e);
    }
    return e * (float)s.n_fft;  // undo the n_fft^-0.5 folded into wtab (squared)
}

__device__ __forceinline__ size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

// rfft post-processing: X[k], k in [0, M], from the packed transform Z (Z[M] == Z[0])
__device__ __forceinline__ float2 unpack_rfft(const float2 *buf, const float2 *tw2, int M, int k) {
    const float2 zk = buf[k & (M - 1)];
    const float2 zm = cconj(buf[(M - k) & (M - 1)]);
    const float2 e = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y + zm.y));
    const float2 d = csub(zk, zm);
    const float2 o = make_float2(0.5f * d.y, -0.5f * d.x);  // -i/2 * (zk - zm)
    return cadd(e, cmul(tw2[k], o));
}

// -----------------------------------------------------------------------------------------------------
// STFT forward
// -----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stft_fwd_kernel(const float *__restrict__ wave, float *__restrict__ mag,
                                                       float *__restrict__ phase, const StftShape s) {
    extern __shared__ __align__(16) unsigned char smem[];
    Tables tb;
    size_t off = 0;
    tb.tw = reinterpret_cast<float2 *>(smem + off); off = align16(off + sizeof(float2) * s.M);
    tb.tw2 = reinterpret_cast<float2 *>(smem + off); off = align16(off + sizeof(float2) * (s.M + 1));
    tb.wtab = reinterpret_cast<float *>(smem + off); off = align16(off + sizeof(float) * s.n_fft);
    float2 *bufs = reinterpret_cast<float2 *>(smem + off); off = align16(off + sizeof(float2) * s.M * kFramesPerCta);
    float *st_mag = reinterpret_cast<float *>(smem + off); off = align16(off + sizeof(float) * s.F * kStageStride);
    float *st_ph = reinterpret_cast<float *>(smem + off);

    build_tables(tb, s);
    __syncthreads();

    const int b = blockIdx.y;
    const int f0 = blockIdx.x * kFramesPerCta;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = f0 + warp;
    float2 *buf = bufs + (size_t)warp * s.M;
    const float *row = wave + (size_t)b * s.T;
    if (f < s.n_frames) {
        const int start = f * s.hop - s.n_fft / 2;
        for (int n = lane; n < s.M; n += 32) {
            int t0 = start + 2 * n, t1 = t0 + 1;
            t0 = t0 < 0 ? -t0 : (t0 >= s.T ? 2 * (s.T - 1) - t0 : t0);
            t1 = t1 < 0 ? -t1 : (t1 >= s.T ? 2 * (s.T - 1) - t1 : t1);
            buf[perm_index(n, s.M, s.log2m)] = make_float2(__ldg(row + t0) * tb.wtab[2 * n], __ldg(row + t1) * tb.wtab[2 * n + 1]);
        }
        __syncwarp();
        warp_fft(buf, tb.tw, s.M, s.log2m, lane);
        for (int k = lane; k <= s.M; k += 32) {
            const float2 X = unpack_rfft(buf, tb.tw2, s.M, k);
            st_mag[k * kStageStride + warp] = log2f(sqrtf(fmaf(X.x, X.x, X.y * X.y)) + 1e-8f);
            st_ph[k * kStageStride + warp] = atan2f(X.y, X.x);
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < s.F * kFramesPerCta; idx += blockDim.x) {
        const int k = idx >> 3, j = idx & 7;
        const int ff = f0 + j;
        if (ff < s.n_frames) {
            const size_t o = ((size_t)b * s.F + k) * s.n_frames + ff;
            mag[o] = st_mag[k * kStageStride + j];
            phase[o] = st_ph[k * kStageStride + j];
        }
    }
}

// -----------------------------------------------------------------------------------------------------
// iSTFT forward.  A CTA owns `stride_frames * hop` consecutive output samples and computes every frame that
// overlaps them (rounds of 8 frames, one per warp); the overlap-add is a gather in ascending frame order,
// so the result is deterministic.
// -----------------------------------------------------------------------------------------------------
struct IstftPlan {
    int q;              // ceil(n_fft / hop): frames overlapping one sample
    int rounds;         // rounds of 8 frames per CTA
    int stride_frames;  // frames between consecutive CTAs (16)
    int ctas_per_row;
};

__global__ void __launch_bounds__(256) istft_fwd_kernel(const float *__restrict__ mag, const float *__restrict__ phase,
                                                        float *__restrict__ wave, const StftShape s, const IstftPlan pl) {
    extern __shared__ __align__(16) unsigned char smem[];
    Tables tb;
    size_t off = 0;
    tb.tw = reinterpret_cast<float2 *>(smem + off); off = align16(off + sizeof(float2) * s.M);
    tb.tw2 = reinterpret_cast<float2 *>(smem + off); off = align16(off + sizeof(float2) * (s.M + 1));
    tb.wtab = reinterpret_cast<float *>(smem + off); off = align16(off + sizeof(float) * s.n_fft);
    float2 *bufs = reinterpret_cast<float2 *>(smem + off); off = align16(off + sizeof(float2) * s.M * kFramesPerCta);
    float *st_mag = reinterpret_cast<float *>(smem + off); off = align16(off + sizeof(float) * s.F * kStageStride);
    float *st_ph = reinterpret_cast<float *>(smem + off); off = align16(off + sizeof(float) * s.F * kStageStride);
    float *acc = reinterpret_cast<float *>(smem + off);

    build_tables(tb, s);
    const int b = blockIdx.y;
    const int cta = blockIdx.x;
    const int F0 = cta * pl.stride_frames;
    const int own = pl.stride_frames * s.hop;
    // owned padded-sample range; the first CTA also takes the samples before it
    const int a_lo = cta == 0 ? 0 : (F0 + pl.q - 1) * s.hop;
    const int a_hi = (F0 + pl.stride_frames + pl.q - 1) * s.hop;
    const int acc_len = a_hi - a_lo;  // <= own + (q-1)*hop for the first CTA
    for (int i = threadIdx.x; i < acc_len; i += blockDim.x) acc[i] = 0.0f;
    (void)own;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2 *buf = bufs + (size_t)warp * s.M;
    const int M = s.M;

    for (int r = 0; r < pl.rounds; ++r) {
        const int fr0 = F0 + r * kFramesPerCta;
        __syncthreads();  // tables ready (r == 0) / previous round's gather done
        if (fr0 >= s.n_frames) break;
        for (int idx = threadIdx.x; idx < s.F * kFramesPerCta; idx += blockDim.x) {
            const int k = idx >> 3, j = idx & 7;
            const int ff = fr0 + j;
            float m = 0.0f, p = 0.0f;
            if (ff < s.n_frames) {
                const size_t o = ((size_t)b * s.F + k) * s.n_frames + ff;
                m = __ldg(mag + o);
                p = __ldg(phase + o);
            }
            st_mag[k * kStageStride + j] = m;
            st_ph[k * kStageStride + j] = p;
        }
        __syncthreads();
        const int f = fr0 + warp;
        if (f < s.n_frames) {
            // Z'[k] = E' + i O',  E' = X[k] + conj X[M-k],  O' = (X[k] - conj X[M-k]) * exp(+pi i k / M);
            // the imaginary parts of X[0] and X[M] are ignored, as a c2r transform does.
            for (int k = lane; k <= (M >> 1); k += 32) {
                const int km = M - k;
                float sk, ck, sm, cm;
                sincosf(st_ph[k * kStageStride + warp], &sk, &ck);
                sincosf(st_ph[km * kStageStride + warp], &sm, &cm);
                const float ak = exp2f(st_mag[k * kStageStride + warp]);
                const float am = exp2f(st_mag[km * kStageStride + warp]);
                float2 xk = make_float2(ak * ck, ak * sk);
                float2 xm = make_float2(am * cm, am * sm);
                if (k == 0) { xk.y = 0.0f; xm.y = 0.0f; }
                // entry k
                {
                    const float2 e = cadd(xk, cconj(xm));
                    const float2 o = cmul(csub(xk, cconj(xm)), cconj(tb.tw2[k]));
                    const float2 z = make_float2(e.x - o.y, e.y + o.x);
                    buf[perm_index(k, M, s.log2m)] = cconj(z);
                }
                // entry M-k (distinct from k unless k == 0 or k == M/2)
                if (k != 0 && km != k) {
                    const float2 e = cadd(xm, cconj(xk));
                    const float2 o = cmul(csub(xm, cconj(xk)), cconj(tb.tw2[km]));
                    const float2 z = make_float2(e.x - o.y, e.y + o.x);
                    buf[perm_index(km, M, s.log2m)] = cconj(z);
                }
            }
            __syncwarp();
            warp_fft(buf, tb.tw, M, s.log2m, lane);
        }
        __syncthreads();
        // gather this round's frames into the owned samples, ascending frame order
        for (int i = threadIdx.x; i < acc_len; i += blockDim.x) {
            const int tp = a_lo + i;
            float v = acc[i];
#pragma unroll
            for (int j = 0; j < kFramesPerCta; ++j) {
                const int ff = fr0 + j;
                const int n = tp - ff * s.hop;
                if (ff < s.n_frames && n >= 0 && n < s.n_fft) {
                    const float2 z = bufs[(size_t)j * M + (n >> 1)];
                    const float x = (n & 1) ? -z.y : z.x;  // time sample n of conj(FFT(conj Z'))
                    v = fmaf(x, tb.wtab[n], v);
                }
            }
            acc[i] = v;
        }
    }
    __syncthreads();
    const int half = s.n_fft / 2;
    for (int i = threadIdx.x; i < acc_len; i += blockDim.x) {
        const int tp = a_lo + i;
        const int t = tp - half;
        if (t >= 0 && t < s.t_out) wave[(size_t)b * s.t_out + t] = acc[i] / envelope(tb.wtab, s, tp);
    }
}

// -----------------------------------------------------------------------------------------------------
// iSTFT backward: d wave -> d mag, d phase.  With g = d wave / envelope, zero outside the signal, framed and
// windowed like the forward STFT (zero instead of reflect padding), R = rfft(frame) * n_fft^-0.5,
// c_k = 1 for k in {0, M} else 2:
//   dRe X_k = c_k Re R_k,  dIm X_k = c_k Im R_k,   X = 2^mag (cos ph + i sin ph)
//   dmag = ln2 * 2^mag * (dRe cos ph + dIm sin ph);   dph = 2^mag * (-dRe sin ph + dIm cos ph)
// -----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) istft_bwd_kernel(const float *__restrict__ mag, const float *__restrict__ phase,
                                                        const float *__restrict__ dwave, float *__restrict__ dmag,
                                                        float *__restrict__ dphase, const StftShape s) {
    extern __shared__ __align__(16) unsigned char smem[];
    Tables tb;
    size_t off = 0;
    tb.tw = reinterpret_cast<float2 *>(smem + off); off = align16(off + sizeof(float2) * s.M);
    tb.tw2 = reinterpret_cast<float2 *>(smem + off); off = align16(off + sizeof(float2) * (s.M + 1));
    tb.wtab = reinterpret_cast<float *>(smem + off); off = align16(off + sizeof(float) * s.n_fft);
    float2 *bufs = reinterpret_cast<float2 *>(smem + off); off = align16(off + sizeof(float2) * s.M * kFramesPerCta);
    float *st_mag = reinterpret_cast<float *>(smem + off); off = align16(off + sizeof(float) * s.F * kStageStride);
    float *st_ph = reinterpret_cast<float *>(smem + off); off = align16(off + sizeof(float) * s.F * kStageStride);
    float *genv = reinterpret_cast<float *>(smem + off);  // [(8-1)*hop + n_fft]  d wave / envelope over the CTA's span

    build_tables(tb, s);
    const int b = blockIdx.y;
    const int f0 = blockIdx.x * kFramesPerCta;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int M = s.M;
    for (int idx = threadIdx.x; idx < s.F * kFramesPerCta; idx += blockDim.x) {
        const int k = idx >> 3, j = idx & 7;
        const int ff = f0 + j;
        float m = 0.0f, p = 0.0f;
        if (ff < s.n_frames) {
            const size_t o = ((size_t)b * s.F + k) * s.n_frames + ff;
            m = __ldg(mag + o);
            p = __ldg(phase + o);
        }
        st_mag[k * kStageStride + j] = m;
        st_ph[k * kStageStride + j] = p;
    }
    __syncthreads();  // tables
    const int span = (kFramesPerCta - 1) * s.hop + s.n_fft;
    const int half = s.n_fft / 2;
    for (int i = threadIdx.x; i < span; i += blockDim.x) {
        const int tp = f0 * s.hop + i;
        const int t = tp - half;
        float g = 0.0f;
        if (t >= 0 && t < s.t_out) g = __ldg(dwave + (size_t)b * s.t_out + t) / envelope(tb.wtab, s, tp);
        genv[i] = g;
    }
    __syncthreads();
    const int f = f0 + warp;
    float2 *buf = bufs + (size_t)warp * M;
    if (f < s.n_frames) {
        const float *g = genv + warp * s.hop;
        for (int n = lane; n < M; n += 32)
            buf[perm_index(n, M, s.log2m)] = make_float2(g[2 * n] * tb.wtab[2 * n], g[2 * n + 1] * tb.wtab[2 * n + 1]);
        __syncwarp();
        warp_fft(buf, tb.tw, M, s.log2m, lane);
        for (int k = lane; k <= M; k += 32) {
            float2 R = unpack_rfft(buf, tb.tw2, M, k);
            const float c = (k == 0 || k == M) ? 1.0f : 2.0f;
            R.x *= c;
            R.y *= c;
            const float m = exp2f(st_mag[k * kStageStride + warp]);
            float sn, cs;
            sincosf(st_ph[k * kStageStride + warp], &sn, &cs);
            st_mag[k * kStageStride + warp] = 0.6931471805599453f * m * fmaf(R.x, cs, R.y * sn);
            st_ph[k * kStageStride + warp] = m * fmaf(R.y, cs, -R.x * sn);
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < s.F * kFramesPerCta; idx += blockDim.x) {
        const int k = idx >> 3, j = idx & 7;
        const int ff = f0 + j;
        if (ff < s.n_frames) {
            const size_t o = ((size_t)b * s.F + k) * s.n_frames + ff;
            dmag[o] = st_mag[k * kStageStride + j];
            dphase[o] = st_ph[k * kStageStride + j];
        }
    }
}

// -----------------------------------------------------------------------------------------------------
static int make_shape(StftShape &s, int B, int T, int n_frames, int n_fft, int hop, int win, const char *who) {
    if (B <= 0) return fail("%s: batch must be positive", who);
    if (n_fft < 64 || n_fft > 2048 || (n_fft & (n_fft - 1))) return fail("%s: n_fft must be a power of two in [64, 2048], got %d", who, n_fft);
    if (hop <= 0 || hop > n_fft) return fail("%s: hop_length must be in [1, n_fft], got %d", who, hop);
    if (win <= 0 || win > n_fft) return fail("%s: win_length must be in [1, n_fft], got %d", who, win);
    s.B = B; s.T = T; s.n_fft = n_fft; s.hop = hop; s.win = win; s.n_frames = n_frames;
    s.M = n_fft / 2;
    s.log2m = 0;
    while ((1 << s.log2m) < s.M) ++s.log2m;
    s.F = s.M + 1;
    s.t_out = hop * (n_frames - 1);
    if (B > 65535) return fail("%s: batch %d exceeds the grid limit", who, B);
    return 0;
}

static size_t base_smem(const StftShape &s) {
    auto a16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    size_t off = 0;
    off = a16(off + sizeof(float2) * s.M);
    off = a16(off + sizeof(float2) * (s.M + 1));
    off = a16(off + sizeof(float) * s.n_fft);
    off = a16(off + sizeof(float2) * s.M * kFramesPerCta);
    off = a16(off + sizeof(float) * s.F * kStageStride);
    off = a16(off + sizeof(float) * s.F * kStageStride);
    return off;
}

template <typename K>
static int set_smem(K kernel, size_t bytes, const char *who) {
    if (bytes > 227 * 1024) return fail("%s: needs %zu bytes of shared memory (> 227 KB); reduce hop_length", who, bytes);
    return check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes), who);
}

}  // namespace vmasr

using namespace vmasr;

extern "C" int vmasr_stft_fwd(const float *wave, float *mag, float *phase, int B, int T, int n_fft, int hop, int win_length,
                              int device, void *stream) {
    if (!wave || !mag || !phase) return fail("stft: null tensor");
    if (T <= n_fft / 2) return fail("stft: reflect padding needs T > n_fft/2 (T %d, n_fft %d)", T, n_fft);
    StftShape s;
    if (int rc = make_shape(s, B, T, 1 + T / (hop > 0 ? hop : 1), n_fft, hop, win_length, "stft")) return rc;
    DeviceGuard guard(device);
    if (!guard.ok) return fail("stft: cannot select CUDA device %d", device);
    const size_t smem = base_smem(s);
    if (int rc = set_smem(stft_fwd_kernel, smem, "stft")) return rc;
    dim3 grid((s.n_frames + kFramesPerCta - 1) / kFramesPerCta, B);
    stft_fwd_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(wave, mag, phase, s);
    return check_cuda(cudaGetLastError(), "stft launch");
}

extern "C" int vmasr_istft_fwd(const float *mag, const float *phase, float *wave, int B, int n_frames, int n_fft, int hop,
                               int win_length, int device, void *stream) {
    if (!wave || !mag || !phase) return fail("istft: null tensor");
    if (n_frames < 2) return fail("istft: needs at least 2 frames");
    StftShape s;
    if (int rc = make_shape(s, B, hop * (n_frames - 1), n_frames, n_fft, hop, win_length, "istft")) return rc;
    DeviceGuard guard(device);
    if (!guard.ok) return fail("istft: cannot select CUDA device %d", device);
    IstftPlan pl;
    pl.q = (n_fft + hop - 1) / hop;
    pl.stride_frames = 16;
    pl.rounds = 2 + (pl.q - 1 + 7) / 8;
    const long long padded_end = (long long)n_fft / 2 + s.t_out;  // one past the last padded sample that is kept
    const long long first_end = (long long)(pl.stride_frames + pl.q - 1) * hop;
    pl.ctas_per_row = 1;
    if (padded_end > first_end) pl.ctas_per_row += (int)((padded_end - first_end + (long long)pl.stride_frames * hop - 1) / ((long long)pl.stride_frames * hop));
    const size_t acc_floats = (size_t)(pl.stride_frames + pl.q - 1) * hop;
    const size_t smem = base_smem(s) + sizeof(float) * acc_floats;
    if (int rc = set_smem(istft_fwd_kernel, smem, "istft")) return rc;
    dim3 grid(pl.ctas_per_row, B);
    istft_fwd_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(mag, phase, wave, s, pl);
    return check_cuda(cudaGetLastError(), "istft launch");
}

extern "C" int vmasr_istft_bwd(const float *mag, const float *phase, const float *dwave, float *dmag, float *dphase, int B,
                               int n_frames, int n_fft, int hop, int win_length, int device, void *stream) {
    if (!mag || !phase || !dwave || !dmag || !dphase) return fail("istft_bwd: null tensor");
    if (n_frames < 2) return fail("istft_bwd: needs at least 2 frames");
    StftShape s;
    if (int rc = make_shape(s, B, hop * (n_frames - 1), n_frames, n_fft, hop, win_length, "istft_bwd")) return rc;
    DeviceGuard guard(device);
    if (!guard.ok) return fail("istft_bwd: cannot select CUDA device %d", device);
    const size_t smem = base_smem(s) + sizeof(float) * ((size_t)(kFramesPerCta - 1) * hop + n_fft);
    if (int rc = set_smem(istft_bwd_kernel, smem, "istft_bwd")) return rc;
    dim3 grid((s.n_frames + kFramesPerCta - 1) / kFramesPerCta, B);
    istft_bwd_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(mag, phase, dwave, dmag, dphase, s);
    return check_cuda(cudaGetLastError(), "istft_bwd launch");
}
