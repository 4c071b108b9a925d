// Magnitude/phase STFT, iSTFT and their backward passes for sm_100a as shared-memory FFT kernels.
// Replace wav2spectro / spectro2wav (utils/stft.py:22-68, 71-115): torch.stft / torch.istft with
// normalized=True, center=True (reflect pad), periodic Hann of win_length zero-padded to n_fft, onesided,
// fused with log2(|X|+1e-8) / angle and exp2 / polar; and the magnitude-only, un-normalised STFT of the multi-resolution
// loss and the LSD metric (model/loss.py:17-45, model/metric.py:5-12).  The reference runs ~8 library launches per call
// (pad, frame, window, cuFFT, scale, abs, log2, angle); here a transform is one kernel (analysis) or three (synthesis).
//
// ANALYSIS (stft_fwd_kernel, istft_bwd_kernel): a CTA of 8 warps handles 8 consecutive frames; each warp runs one real FFT of
// n_fft points as a complex FFT of M = n_fft/2 points (even/odd packing), in place in shared memory (digit-reversed load,
// radix-4 decimation-in-time stages, one radix-2 stage first when log2 M is odd).  The (F, n_frames) planes have the frame
// index fastest, so the 8 frames of a CTA are staged in shared memory and leave (or enter) HBM as 32-byte rows.
//
// SYNTHESIS (synth_kernel + finalize_kernel: iSTFT forward and STFT backward): every frame is inverted exactly ONCE.  A CTA
// owns FPC consecutive frames (FPC * hop >= n_fft), overlap-adds them in shared memory in ascending frame order and adds the
// partial sums into a zero-filled PADDED accumulator (n_fft + hop (n_frames - 1) samples per clip) with red.global.add: a
// padded sample is touched by at most two CTAs, and a sum of two terms does not depend on their order, so the result is
// deterministic.  finalize_kernel turns the accumulator into the output: iSTFT divides by the window envelope and trims
// n_fft/2 on each side (torch.istft, center=True); the STFT backward folds the reflect padding back onto the signal.
// (Round 1 gave each CTA a range of OUTPUT samples and recomputed every frame that overlaps it: 24 frames per 16 owned.)
#include <mutex>

#include "common.cuh"

namespace vmasr {

constexpr int kFramesPerCta = 8;
constexpr int kStageStride = 9;  // floats per k row of the staging planes (8 frames + 1 pad)

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// ---- transcendental pieces of the epilogues / prologues (MUFU based; errors far below the 1e-5 bar of the path) ----------
// atan2 from one division and the 8-term odd polynomial of Abramowitz & Stegun 4.4.49 (|error| <= 2e-8 on [0, 1]); quadrant by
// comparisons; signed zeros as atan2f (angle(-1 - 0i) = -pi).  atan2f itself is ~70 instructions per bin.
__device__ __forceinline__ float fast_atan2(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float t = mx > 0.0f ? __fdividef(mn, mx) : 0.0f;
    const float s = t * t;
    float p = fmaf(s, 0.0028662257f, -0.0161657367f);
    p = fmaf(s, p, 0.0429096138f);
    p = fmaf(s, p, -0.0752896400f);
    p = fmaf(s, p, 0.1065626393f);
    p = fmaf(s, p, -0.1420889944f);
    p = fmaf(s, p, 0.1999355085f);
    p = fmaf(s, p, -0.3333314528f);
    float r = fmaf(t * s, p, t);
    if (ay > ax) r = 1.5707963267948966f - r;
    if (x < 0.0f) r = 3.141592653589793f - r;
    return copysignf(r, y);
}
// log2(sqrt(p) + 1e-8) with the MUFU square root and logarithm (2 ulp each)
__device__ __forceinline__ float fast_log2_mag(float p) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p));
    return lg2_approx(r + 1e-8f);
}
// sin / cos of a phase: one reduction to [-pi, pi] (phases are angles, or network outputs of order 1 .. 10), then the MUFU
__device__ __forceinline__ void fast_sincos(float x, float *sn, float *cs) {
    x = fmaf(-6.283185307179586f, rintf(x * 0.15915494309189535f), x);
    __sincosf(x, sn, cs);
}

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
    int t_out;    // hop * (n_frames - 1), the iSTFT output length
    int padded;   // n_fft + hop * (n_frames - 1): samples the frames cover
    float scale;  // n_fft^-0.5 (normalized=True) or 1
};

// shared-memory tables common to all kernels
struct Tables {
    float2 *tw;   // [M]    exp(-2 pi i t / M)
    float2 *tw2;  // [M+1]  exp(-pi i k / M)
    float *wtab;  // [n_fft] padded periodic Hann * scale
    unsigned short *perm;  // [M] perm_index(n): the digit reversal costs ~25 instructions, a frame needs M of them
};

__device__ __forceinline__ float hann_padded(const StftShape &s, int n) {
    const int left = (s.n_fft - s.win) / 2;
    return (n >= left && n < left + s.win) ? 0.5f - 0.5f * cospif(2.0f * (float)(n - left) / (float)s.win) : 0.0f;
}

// one sincospif per table entry: tw[t] = tw2[2 t]
__device__ __forceinline__ void build_tables(const Tables &tb, const StftShape &s) {
    const float inv_m = 1.0f / (float)s.M;
    for (int k = threadIdx.x; k <= s.M; k += blockDim.x) {
        float sn, cs;
        sincospif((float)k * inv_m, &sn, &cs);
        const float2 v = make_float2(cs, -sn);
        tb.tw2[k] = v;
        if (!(k & 1) && k < 2 * s.M) tb.tw[k >> 1] = v;
    }
    for (int k = s.M + 1 + (int)threadIdx.x; k < 2 * s.M; k += blockDim.x) {
        if (k & 1) continue;
        float sn, cs;
        sincospif((float)k * inv_m, &sn, &cs);
        tb.tw[k >> 1] = make_float2(cs, -sn);
    }
    for (int n = threadIdx.x; n < s.n_fft; n += blockDim.x) tb.wtab[n] = hann_padded(s, n) * s.scale;
    for (int n = threadIdx.x; n < s.M; n += blockDim.x) tb.perm[n] = (unsigned short)perm_index(n, s.M, s.log2m);
}

// sum over the frames covering padded sample tp of window^2 (torch.istft's envelope), from the analytic window
__device__ __forceinline__ float envelope(const StftShape &s, int tp) {
    int f_lo = (tp - s.n_fft + s.hop) / s.hop;  // ceil((tp - n_fft + 1) / hop) for tp - n_fft + 1 > 0
    if (tp - s.n_fft + 1 <= 0) f_lo = 0;
    int f_hi = tp / s.hop;
    if (f_hi > s.n_frames - 1) f_hi = s.n_frames - 1;
    float e = 0.0f;
    for (int f = f_lo; f <= f_hi; ++f) {
        const float w = hann_padded(s, tp - f * s.hop);
        e = fmaf(w, w, e);
    }
    return e;
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

// c2r packing: entries k and M - k of the buffer whose forward FFT (conjugated) is the real sequence with one-sided
// spectrum X (imaginary parts of X[0], X[M] ignored, as a c2r transform does): x[n] = sum_k' X[k'] e^{+2 pi i k' n / N}
// over the Hermitian extension, UN-normalised.
__device__ __forceinline__ void pack_c2r(float2 *buf, const float2 *tw2, const unsigned short *perm, int M, int k, float2 xk, float2 xm) {
    const int km = M - k;
    if (k == 0) { xk.y = 0.0f; xm.y = 0.0f; }
    {
        const float2 e = cadd(xk, cconj(xm));
        const float2 o = cmul(csub(xk, cconj(xm)), cconj(tw2[k]));
        buf[perm[k & (M - 1)]] = cconj(make_float2(e.x - o.y, e.y + o.x));
    }
    if (k != 0 && km != k) {
        const float2 e = cadd(xm, cconj(xk));
        const float2 o = cmul(csub(xm, cconj(xk)), cconj(tw2[km]));
        buf[perm[km]] = cconj(make_float2(e.x - o.y, e.y + o.x));
    }
}
// time sample n of the sequence packed by pack_c2r, after warp_fft
__device__ __forceinline__ float c2r_sample(const float2 *buf, int n) {
    const float2 z = buf[n >> 1];
    return (n & 1) ? -z.y : z.x;
}

// windowed, reflect-padded frame f of `row` into the digit-reversed FFT buffer
__device__ __forceinline__ void load_frame(float2 *buf, const float *__restrict__ row, const Tables &tb, const StftShape &s, int f, int lane) {
    const int start = f * s.hop - s.n_fft / 2;
    for (int n = lane; n < s.M; n += 32) {
        int t0 = start + 2 * n, t1 = t0 + 1;
        t0 = t0 < 0 ? -t0 : (t0 >= s.T ? 2 * (s.T - 1) - t0 : t0);
        t1 = t1 < 0 ? -t1 : (t1 >= s.T ? 2 * (s.T - 1) - t1 : t1);
        buf[tb.perm[n]] = make_float2(__ldg(row + t0) * tb.wtab[2 * n], __ldg(row + t1) * tb.wtab[2 * n + 1]);
    }
}

struct Carve {
    Tables tb;
    float2 *bufs;          // [8][M]
    float *st_a, *st_b;    // [F][kStageStride] staging planes
    float *extra;          // kernel-specific tail
};
__device__ __forceinline__ Carve carve(unsigned char *smem, const StftShape &s) {
    Carve c;
    size_t off = 0;
    c.tb.tw = reinterpret_cast<float2 *>(smem + off); off = align16(off + sizeof(float2) * s.M);
    c.tb.tw2 = reinterpret_cast<float2 *>(smem + off); off = align16(off + sizeof(float2) * (s.M + 1));
    c.tb.wtab = reinterpret_cast<float *>(smem + off); off = align16(off + sizeof(float) * s.n_fft);
    c.tb.perm = reinterpret_cast<unsigned short *>(smem + off); off = align16(off + sizeof(unsigned short) * s.M);
    c.bufs = reinterpret_cast<float2 *>(smem + off); off = align16(off + sizeof(float2) * s.M * kFramesPerCta);
    c.st_a = reinterpret_cast<float *>(smem + off); off = align16(off + sizeof(float) * s.F * kStageStride);
    c.st_b = reinterpret_cast<float *>(smem + off); off = align16(off + sizeof(float) * s.F * kStageStride);
    c.extra = reinterpret_cast<float *>(smem + off);
    return c;
}

// 8 frames x F bins between the (B, F, n_frames) planes and the staging planes, 32-byte rows
__device__ __forceinline__ void stage_in(const float *__restrict__ a, const float *__restrict__ b, float *st_a, float *st_b,
                                         const StftShape &s, int batch, int f0) {
    for (int idx = threadIdx.x; idx < s.F * kFramesPerCta; idx += blockDim.x) {
        const int k = idx >> 3, j = idx & 7;
        const int ff = f0 + j;
        float va = 0.0f, vb = 0.0f;
        if (ff < s.n_frames) {
            const size_t o = ((size_t)batch * s.F + k) * s.n_frames + ff;
            va = __ldg(a + o);
            if (b) vb = __ldg(b + o);
        }
        st_a[k * kStageStride + j] = va;
        if (b) st_b[k * kStageStride + j] = vb;
    }
}
__device__ __forceinline__ void stage_out(float *__restrict__ a, float *__restrict__ b, const float *st_a, const float *st_b,
                                          const StftShape &s, int batch, int f0) {
    for (int idx = threadIdx.x; idx < s.F * kFramesPerCta; idx += blockDim.x) {
        const int k = idx >> 3, j = idx & 7;
        const int ff = f0 + j;
        if (ff < s.n_frames) {
            const size_t o = ((size_t)batch * s.F + k) * s.n_frames + ff;
            a[o] = st_a[k * kStageStride + j];
            if (b) b[o] = st_b[k * kStageStride + j];
        }
    }
}

// -----------------------------------------------------------------------------------------------------
// STFT forward.  LINEAR = false: mag = log2(|X| + 1e-8), phase = angle(X) (utils/stft.py:65-66);
//                LINEAR = true : mag = sqrt(max(|X|^2, clamp)), no phase (model/loss.py:37, model/metric.py:11)
// -----------------------------------------------------------------------------------------------------
template <bool LINEAR>
__global__ void __launch_bounds__(256) stft_fwd_kernel(const float *__restrict__ wave, float *__restrict__ mag,
                                                       float *__restrict__ phase, const StftShape s, const float clamp) {
    extern __shared__ __align__(16) unsigned char smem[];
    const Carve c = carve(smem, s);
    build_tables(c.tb, s);
    __syncthreads();

    const int b = blockIdx.y;
    const int f0 = blockIdx.x * kFramesPerCta;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = f0 + warp;
    float2 *buf = c.bufs + (size_t)warp * s.M;
    if (f < s.n_frames) {
        load_frame(buf, wave + (size_t)b * s.T, c.tb, s, f, lane);
        __syncwarp();
        warp_fft(buf, c.tb.tw, s.M, s.log2m, lane);
        for (int k = lane; k <= s.M; k += 32) {
            const float2 X = unpack_rfft(buf, c.tb.tw2, s.M, k);
            const float p = fmaf(X.x, X.x, X.y * X.y);
            if (LINEAR) {
                c.st_a[k * kStageStride + warp] = sqrtf(fmaxf(p, clamp));
            } else {
                c.st_a[k * kStageStride + warp] = fast_log2_mag(p);
                c.st_b[k * kStageStride + warp] = fast_atan2(X.y, X.x);
            }
        }
    }
    __syncthreads();
    stage_out(mag, LINEAR ? nullptr : phase, c.st_a, c.st_b, s, b, f0);
}

// -----------------------------------------------------------------------------------------------------
// Synthesis: the frames' time signals, overlap-added into the padded accumulator.
//   MODE 0  iSTFT forward        frame spectrum X = 2^mag (cos ph + i sin ph)                 (utils/stft.py:100-111)
//   MODE 1  STFT backward, log2  G = dL/dRe X + i dL/dIm X from (d mag, d phase) of wav2spectro; X is recomputed from the wave
//   MODE 2  STFT backward, lin   G from d mag of the linear magnitude
// The adjoint of x -> X_k = scale sum_n w_n x_n e^{-2 pi i k n / N} (k = 0..M) is  d x_n = scale w_n Re sum_k G_k e^{+2 pi i k n / N}:
// a c2r transform of G with the interior bins halved (the Hermitian extension counts them twice).  `wtab` carries scale.
// -----------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) synth_kernel(const float *__restrict__ in_a, const float *__restrict__ in_b,
                                                    const float *__restrict__ wave, float *__restrict__ accum, const StftShape s,
                                                    const int rounds, const float clamp) {
    extern __shared__ __align__(16) unsigned char smem[];
    const Carve c = carve(smem, s);
    float *acc = c.extra;  // [(8 rounds - 1) hop + n_fft] this CTA's overlap-added frames
    build_tables(c.tb, s);
    const int b = blockIdx.y;
    const int F0 = blockIdx.x * kFramesPerCta * rounds;
    const int span = (kFramesPerCta * rounds - 1) * s.hop + s.n_fft;
    for (int i = threadIdx.x; i < span; i += blockDim.x) acc[i] = 0.0f;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2 *buf = c.bufs + (size_t)warp * s.M;
    const int M = s.M;

    for (int r = 0; r < rounds; ++r) {
        const int fr0 = F0 + r * kFramesPerCta;
        __syncthreads();  // tables ready (r == 0) / previous round's overlap-add done
        if (fr0 >= s.n_frames) break;
        stage_in(in_a, MODE == 2 ? nullptr : in_b, c.st_a, c.st_b, s, b, fr0);
        __syncthreads();
        const int f = fr0 + warp;
        if (f < s.n_frames) {
            if (MODE != 0) {
                // recompute the frame's spectrum, turn (d mag, d phase) into G in place of the staged values
                load_frame(buf, wave + (size_t)b * s.T, c.tb, s, f, lane);
                __syncwarp();
                warp_fft(buf, c.tb.tw, M, s.log2m, lane);
                for (int k = lane; k <= M; k += 32) {
                    const float2 X = unpack_rfft(buf, c.tb.tw2, M, k);
                    const float p = fmaf(X.x, X.x, X.y * X.y);
                    const float dm = c.st_a[k * kStageStride + warp];
                    float gre, gim;
                    if (MODE == 1) {
                        // mag = log2(r + 1e-8), phase = atan2(Im, Re), r = |X|
                        const float dp = c.st_b[k * kStageStride + warp];
                        const float r = sqrtf(p);
                        const float inv_r = r > 0.0f ? 1.0f / r : 0.0f, inv_p = p > 0.0f ? 1.0f / p : 0.0f;
                        const float gm = dm * 1.4426950408889634f / (r + 1e-8f) * inv_r;  // d mag / d Re = gm * Re
                        gre = fmaf(gm, X.x, -dp * X.y * inv_p);
                        gim = fmaf(gm, X.y, dp * X.x * inv_p);
                    } else {
                        const float gm = p > clamp ? dm * rsqrtf(p) : 0.0f;  // d sqrt(max(p, clamp)) = X / |X| above the clamp
                        gre = gm * X.x;
                        gim = gm * X.y;
                    }
                    const float half = (k == 0 || k == M) ? 1.0f : 0.5f;
                    c.st_a[k * kStageStride + warp] = gre * half;
                    c.st_b[k * kStageStride + warp] = gim * half;
                }
                __syncwarp();
            }
            for (int k = lane; k <= (M >> 1); k += 32) {
                const int km = M - k;
                float2 xk, xm;
                if (MODE == 0) {
                    float sk, ck, sm, cm;
                    fast_sincos(c.st_b[k * kStageStride + warp], &sk, &ck);
                    fast_sincos(c.st_b[km * kStageStride + warp], &sm, &cm);
                    const float ak = ex2_approx(c.st_a[k * kStageStride + warp]);
                    const float am = ex2_approx(c.st_a[km * kStageStride + warp]);
                    xk = make_float2(ak * ck, ak * sk);
                    xm = make_float2(am * cm, am * sm);
                } else {
                    xk = make_float2(c.st_a[k * kStageStride + warp], c.st_b[k * kStageStride + warp]);
                    xm = make_float2(c.st_a[km * kStageStride + warp], c.st_b[km * kStageStride + warp]);
                }
                pack_c2r(buf, c.tb.tw2, c.tb.perm, M, k, xk, xm);
            }
            __syncwarp();
            warp_fft(buf, c.tb.tw, M, s.log2m, lane);
        }
        __syncthreads();
        // this round's 8 frames into the CTA's samples, ascending frame order
        const int base = r * kFramesPerCta * s.hop;
        const int rspan = (kFramesPerCta - 1) * s.hop + s.n_fft;
        for (int i = threadIdx.x; i < rspan; i += blockDim.x) {
            float v = acc[base + i];
#pragma unroll
            for (int j = 0; j < kFramesPerCta; ++j) {
                const int n = i - j * s.hop;
                if (fr0 + j < s.n_frames && n >= 0 && n < s.n_fft) v = fmaf(c2r_sample(c.bufs + (size_t)j * M, n), c.tb.wtab[n], v);
            }
            acc[base + i] = v;
        }
    }
    __syncthreads();
    float *dst = accum + (size_t)b * s.padded + (size_t)F0 * s.hop;
    const int limit = s.padded - F0 * s.hop;
    for (int i = threadIdx.x; i < span && i < limit; i += blockDim.x) {
        const float v = acc[i];
        if (v != 0.0f) atomicAdd(dst + i, v);  // compiles to RED (result unused); at most two CTAs add to a sample
    }
}

// accumulator -> output.  ISTFT: out[t] = acc[t + n_fft/2] / envelope; otherwise fold the reflect padding back.
template <bool ISTFT>
__global__ void __launch_bounds__(256) finalize_kernel(const float *__restrict__ accum, float *__restrict__ out, const StftShape s) {
    const int b = blockIdx.y;
    const int half = s.n_fft / 2;
    const int len = ISTFT ? s.t_out : s.T;
    const float *acc = accum + (size_t)b * s.padded;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < len; t += gridDim.x * blockDim.x) {
        float v;
        if (ISTFT) {
            v = acc[t + half] / envelope(s, t + half);
        } else {
            const int p = t + half;
            v = p < s.padded ? acc[p] : 0.0f;
            if (t >= 1 && t <= half) v += acc[half - t];                                   // left reflection: pad[p] = x[half - p]
            const int pr = half + 2 * (s.T - 1) - t;                                        // right reflection
            if (t <= s.T - 2 && t >= s.T - 1 - half && pr < s.padded) v += acc[pr];
        }
        out[(size_t)b * len + t] = v;
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
    const Carve c = carve(smem, s);
    float *genv = c.extra;  // [(8-1)*hop + n_fft]  d wave / envelope over the CTA's span

    build_tables(c.tb, s);
    const int b = blockIdx.y;
    const int f0 = blockIdx.x * kFramesPerCta;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int M = s.M;
    stage_in(mag, phase, c.st_a, c.st_b, s, b, f0);
    const int span = (kFramesPerCta - 1) * s.hop + s.n_fft;
    const int half = s.n_fft / 2;
    for (int i = threadIdx.x; i < span; i += blockDim.x) {
        const int tp = f0 * s.hop + i;
        const int t = tp - half;
        float g = 0.0f;
        if (t >= 0 && t < s.t_out) g = __ldg(dwave + (size_t)b * s.t_out + t) / envelope(s, tp);
        genv[i] = g;
    }
    __syncthreads();
    const int f = f0 + warp;
    float2 *buf = c.bufs + (size_t)warp * M;
    if (f < s.n_frames) {
        const float *g = genv + warp * s.hop;
        for (int n = lane; n < M; n += 32)
            buf[c.tb.perm[n]] = make_float2(g[2 * n] * c.tb.wtab[2 * n], g[2 * n + 1] * c.tb.wtab[2 * n + 1]);
        __syncwarp();
        warp_fft(buf, c.tb.tw, M, s.log2m, lane);
        for (int k = lane; k <= M; k += 32) {
            float2 R = unpack_rfft(buf, c.tb.tw2, M, k);
            const float cf = (k == 0 || k == M) ? 1.0f : 2.0f;
            R.x *= cf;
            R.y *= cf;
            const float m = ex2_approx(c.st_a[k * kStageStride + warp]);
            float sn, cs;
            fast_sincos(c.st_b[k * kStageStride + warp], &sn, &cs);
            c.st_a[k * kStageStride + warp] = 0.6931471805599453f * m * fmaf(R.x, cs, R.y * sn);
            c.st_b[k * kStageStride + warp] = m * fmaf(R.y, cs, -R.x * sn);
        }
    }
    __syncthreads();
    stage_out(dmag, dphase, c.st_a, c.st_b, s, b, f0);
}

// -----------------------------------------------------------------------------------------------------
static int make_shape(StftShape &s, int B, int T, int n_frames, int n_fft, int hop, int win, bool normalized, const char *who) {
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
    s.padded = n_fft + hop * (n_frames - 1);
    s.scale = normalized ? 1.0f / sqrtf((float)n_fft) : 1.0f;
    if (B > 65535) return fail("%s: batch %d exceeds the grid limit", who, B);
    return 0;
}

static size_t base_smem(const StftShape &s) {
    auto a16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    size_t off = 0;
    off = a16(off + sizeof(float2) * s.M);
    off = a16(off + sizeof(float2) * (s.M + 1));
    off = a16(off + sizeof(float) * s.n_fft);
    off = a16(off + sizeof(unsigned short) * s.M);
    off = a16(off + sizeof(float2) * s.M * kFramesPerCta);
    off = a16(off + sizeof(float) * s.F * kStageStride);
    off = a16(off + sizeof(float) * s.F * kStageStride);
    return off;
}

// The dynamic shared-memory limit of a kernel is raised once per device to the most any shape can ask for (the
// attribute is sticky and cudaFuncSetAttribute is not free: round 1 called it on every launch).
constexpr size_t kMaxSmem = 227 * 1024;
static int ensure_smem(const void *kernel, size_t bytes, const char *who) {
    if (bytes > kMaxSmem) return fail("%s: needs %zu bytes of shared memory (> 227 KB); reduce hop_length", who, bytes);
    struct Entry { const void *fn; bool done[64]; };
    static std::mutex mu;
    static Entry table[16] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    std::lock_guard<std::mutex> lock(mu);
    for (Entry &e : table) {
        if (e.fn == nullptr) e.fn = kernel;
        if (e.fn != kernel) continue;
        if (!e.done[dev]) {
            if (int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem), who)) return rc;
            e.done[dev] = true;
        }
        return 0;
    }
    return check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem), who);
}

// frames one synthesis CTA owns: a multiple of 8 with FPC * hop >= n_fft (at most two CTAs touch a padded sample)
static int synth_rounds(const StftShape &s) { return ((s.n_fft + s.hop - 1) / s.hop + kFramesPerCta - 1) / kFramesPerCta; }

template <int MODE>
static int run_synth(const float *in_a, const float *in_b, const float *wave, float *out, float *scratch, const StftShape &s,
                     float clamp, cudaStream_t stream, const char *who) {
    if (!scratch) return fail("%s: scratch (B x vmasr_stft_scratch_floats) must be non-null", who);
    const int rounds = synth_rounds(s);
    const size_t smem = base_smem(s) + sizeof(float) * ((size_t)(kFramesPerCta * rounds - 1) * s.hop + s.n_fft);
    if (int rc = ensure_smem(reinterpret_cast<const void *>(&synth_kernel<MODE>), smem, who)) return rc;
    if (int rc = check_cuda(cudaMemsetAsync(scratch, 0, sizeof(float) * (size_t)s.B * s.padded, stream), who)) return rc;
    dim3 grid((s.n_frames + kFramesPerCta * rounds - 1) / (kFramesPerCta * rounds), s.B);
    synth_kernel<MODE><<<grid, 256, smem, stream>>>(in_a, in_b, wave, scratch, s, rounds, clamp);
    const int len = MODE == 0 ? s.t_out : s.T;
    dim3 fgrid((len + 1023) / 1024, s.B);
    if (MODE == 0) finalize_kernel<true><<<fgrid, 256, 0, stream>>>(scratch, out, s);
    else finalize_kernel<false><<<fgrid, 256, 0, stream>>>(scratch, out, s);
    return check_cuda(cudaGetLastError(), who);
}

}  // namespace vmasr

using namespace vmasr;

extern "C" uint64_t vmasr_stft_scratch_floats(int n_frames, int n_fft, int hop) {
    if (n_frames <= 0 || n_fft <= 0 || hop <= 0) return 0;
    return (uint64_t)n_fft + (uint64_t)hop * (uint64_t)(n_frames - 1);
}

extern "C" int vmasr_stft_fwd(const float *wave, float *mag, float *phase, int B, int T, int n_fft, int hop, int win_length,
                              int device, void *stream) {
    if (!wave || !mag || !phase) return fail("stft: null tensor");
    if (T <= n_fft / 2) return fail("stft: reflect padding needs T > n_fft/2 (T %d, n_fft %d)", T, n_fft);
    StftShape s;
    if (int rc = make_shape(s, B, T, 1 + T / (hop > 0 ? hop : 1), n_fft, hop, win_length, true, "stft")) return rc;
    DeviceGuard guard(device);
    if (!guard.ok) return fail("stft: cannot select CUDA device %d", device);
    const size_t smem = base_smem(s);
    if (int rc = ensure_smem(reinterpret_cast<const void *>(&stft_fwd_kernel<false>), smem, "stft")) return rc;
    dim3 grid((s.n_frames + kFramesPerCta - 1) / kFramesPerCta, B);
    stft_fwd_kernel<false><<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(wave, mag, phase, s, 0.0f);
    return check_cuda(cudaGetLastError(), "stft launch");
}

extern "C" int vmasr_stft_mag_fwd(const float *wave, float *mag, int B, int T, int n_fft, int hop, int win_length, int normalized,
                                  float clamp_min, int device, void *stream) {
    if (!wave || !mag) return fail("stft_mag: null tensor");
    if (T <= n_fft / 2) return fail("stft_mag: reflect padding needs T > n_fft/2 (T %d, n_fft %d)", T, n_fft);
    StftShape s;
    if (int rc = make_shape(s, B, T, 1 + T / (hop > 0 ? hop : 1), n_fft, hop, win_length, normalized != 0, "stft_mag")) return rc;
    DeviceGuard guard(device);
    if (!guard.ok) return fail("stft_mag: cannot select CUDA device %d", device);
    const size_t smem = base_smem(s);
    if (int rc = ensure_smem(reinterpret_cast<const void *>(&stft_fwd_kernel<true>), smem, "stft_mag")) return rc;
    dim3 grid((s.n_frames + kFramesPerCta - 1) / kFramesPerCta, B);
    stft_fwd_kernel<true><<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(wave, mag, nullptr, s, clamp_min);
    return check_cuda(cudaGetLastError(), "stft_mag launch");
}

extern "C" int vmasr_stft_bwd(const float *wave, const float *dmag, const float *dphase, float *dwave, float *scratch, int B, int T,
                              int n_fft, int hop, int win_length, int device, void *stream) {
    if (!wave || !dmag || !dphase || !dwave) return fail("stft_bwd: null tensor");
    if (T <= n_fft / 2) return fail("stft_bwd: reflect padding needs T > n_fft/2 (T %d, n_fft %d)", T, n_fft);
    StftShape s;
    if (int rc = make_shape(s, B, T, 1 + T / (hop > 0 ? hop : 1), n_fft, hop, win_length, true, "stft_bwd")) return rc;
    DeviceGuard guard(device);
    if (!guard.ok) return fail("stft_bwd: cannot select CUDA device %d", device);
    return run_synth<1>(dmag, dphase, wave, dwave, scratch, s, 0.0f, static_cast<cudaStream_t>(stream), "stft_bwd");
}

extern "C" int vmasr_stft_mag_bwd(const float *wave, const float *dmag, float *dwave, float *scratch, int B, int T, int n_fft, int hop,
                                  int win_length, int normalized, float clamp_min, int device, void *stream) {
    if (!wave || !dmag || !dwave) return fail("stft_mag_bwd: null tensor");
    if (T <= n_fft / 2) return fail("stft_mag_bwd: reflect padding needs T > n_fft/2 (T %d, n_fft %d)", T, n_fft);
    StftShape s;
    if (int rc = make_shape(s, B, T, 1 + T / (hop > 0 ? hop : 1), n_fft, hop, win_length, normalized != 0, "stft_mag_bwd")) return rc;
    DeviceGuard guard(device);
    if (!guard.ok) return fail("stft_mag_bwd: cannot select CUDA device %d", device);
    return run_synth<2>(dmag, nullptr, wave, dwave, scratch, s, clamp_min, static_cast<cudaStream_t>(stream), "stft_mag_bwd");
}

extern "C" int vmasr_istft_fwd(const float *mag, const float *phase, float *wave, float *scratch, int B, int n_frames, int n_fft,
                               int hop, int win_length, int device, void *stream) {
    if (!wave || !mag || !phase) return fail("istft: null tensor");
    if (n_frames < 2) return fail("istft: needs at least 2 frames");
    StftShape s;
    if (int rc = make_shape(s, B, hop * (n_frames - 1), n_frames, n_fft, hop, win_length, true, "istft")) return rc;
    DeviceGuard guard(device);
    if (!guard.ok) return fail("istft: cannot select CUDA device %d", device);
    return run_synth<0>(mag, phase, nullptr, wave, scratch, s, 0.0f, static_cast<cudaStream_t>(stream), "istft");
}

extern "C" int vmasr_istft_bwd(const float *mag, const float *phase, const float *dwave, float *dmag, float *dphase, int B,
                               int n_frames, int n_fft, int hop, int win_length, int device, void *stream) {
    if (!mag || !phase || !dwave || !dmag || !dphase) return fail("istft_bwd: null tensor");
    if (n_frames < 2) return fail("istft_bwd: needs at least 2 frames");
    StftShape s;
    if (int rc = make_shape(s, B, hop * (n_frames - 1), n_frames, n_fft, hop, win_length, true, "istft_bwd")) return rc;
    DeviceGuard guard(device);
    if (!guard.ok) return fail("istft_bwd: cannot select CUDA device %d", device);
    const size_t smem = base_smem(s) + sizeof(float) * ((size_t)(kFramesPerCta - 1) * hop + n_fft);
    if (int rc = ensure_smem(reinterpret_cast<const void *>(&istft_bwd_kernel), smem, "istft_bwd")) return rc;
    dim3 grid((s.n_frames + kFramesPerCta - 1) / kFramesPerCta, B);
    istft_bwd_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(mag, phase, dwave, dmag, dphase, s);
    return check_cuda(cudaGetLastError(), "istft_bwd launch");
}
