// Building blocks of the fast-path scan kernels (fp32 IO, d_state 1, 16-byte aligned rows; sm_100a only):
//   * packed fp32x2 arithmetic (SASS FFMA2 / FMUL2 / FADD2): a thread owns 8 consecutive positions and does all
//     element-wise work on 4 position PAIRS, so the FMA-pipe instruction count per element is halved;
//   * warp scans of affine maps whose combine step is predicated by the shuffle's own in-range predicate
//     (2 SHFL + 2 FP per level, no compare / select);
//   * softplus in the log2 domain;
//   * the 3-sums-at-once warp reduction of the backward's per-channel sums.
#pragma once
#include "pipe.cuh"

namespace vmasr {

constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float2 f2(float v) { return make_float2(v, v); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }

// 8 consecutive floats of shared memory -> 4 position pairs
__device__ __forceinline__ void lds8(const float *p, float2 (&v)[4]) {
    const float4 lo = reinterpret_cast<const float4 *>(p)[0];
    const float4 hi = reinterpret_cast<const float4 *>(p)[1];
    v[0] = make_float2(lo.x, lo.y);
    v[1] = make_float2(lo.z, lo.w);
    v[2] = make_float2(hi.x, hi.y);
    v[3] = make_float2(hi.z, hi.w);
}
__device__ __forceinline__ void ldg8(const float *p, float2 (&v)[4]) {
    const float4 lo = reinterpret_cast<const float4 *>(p)[0];
    const float4 hi = reinterpret_cast<const float4 *>(p)[1];
    v[0] = make_float2(lo.x, lo.y);
    v[1] = make_float2(lo.z, lo.w);
    v[2] = make_float2(hi.x, hi.y);
    v[3] = make_float2(hi.z, hi.w);
}
__device__ __forceinline__ void stg8(float *p, const float2 (&v)[4]) {
    reinterpret_cast<float4 *>(p)[0] = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
    reinterpret_cast<float4 *>(p)[1] = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
}

// 128-bit reductions into global memory (SASS RED.E.ADD.F32x4 on sm_100a): the fused SS2D core adds the outputs of two scan
// directions that share a memory order into one zero-filled plane, which keeps the sum exact and order-independent
// (0 + a + b == a + b whichever comes first).
__device__ __forceinline__ void red4(float *p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red8(float *p, const float2 (&v)[4]) {
    red4(p, v[0].x, v[0].y, v[1].x, v[1].y);
    red4(p + 4, v[2].x, v[2].y, v[3].x, v[3].y);
}

// ---- time order inside a thread ----------------------------------------------------------------------------------------
// A thread owns 8 consecutive MEMORY positions (4 pairs).  With REV the scan runs over the row back to front: the thread that
// is t-th in time owns the t-th segment from the END of the tile and walks its pairs 3..0, .y before .x.  Everything
// element-wise stays in memory order (packed pairs as loaded); only the serial recurrences use these helpers.
template <bool REV> __device__ __forceinline__ constexpr int pair_at(int kk) { return REV ? 3 - kk : kk; }
// forward recurrence over one pair: (p, q) <- running affine map, P / Q = its value after each position
template <bool REV>
__device__ __forceinline__ void walk_pair(const float2 av, const float2 bx, float &p, float &q, float2 &P, float2 &Q) {
    if (!REV) {
        q = fmaf(av.x, q, bx.x); p *= av.x; P.x = p; Q.x = q;
        q = fmaf(av.y, q, bx.y); p *= av.y; P.y = p; Q.y = q;
    } else {
        q = fmaf(av.y, q, bx.y); p *= av.y; P.y = p; Q.y = q;
        q = fmaf(av.x, q, bx.x); p *= av.x; P.x = p; Q.x = q;
    }
}
// state walk: h <- a h + bx, the state after each position
template <bool REV>
__device__ __forceinline__ void walk_state(const float2 av, const float2 bx, float &h, float2 &hs) {
    if (!REV) {
        h = fmaf(av.x, h, bx.x); hs.x = h;
        h = fmaf(av.y, h, bx.y); hs.y = h;
    } else {
        h = fmaf(av.y, h, bx.y); hs.y = h;
        h = fmaf(av.x, h, bx.x); hs.x = h;
    }
}
// forward + adjoint aggregates in one walk in time order: q as above, qr = sum_i (prod_{m <= i} a_m) c_i
template <bool REV>
__device__ __forceinline__ void walk_pair2(const float2 av, const float2 bx, const float2 cdy, float &p, float &q, float &qr) {
    if (!REV) {
        q = fmaf(av.x, q, bx.x); p *= av.x; qr = fmaf(p, cdy.x, qr);
        q = fmaf(av.y, q, bx.y); p *= av.y; qr = fmaf(p, cdy.y, qr);
    } else {
        q = fmaf(av.y, q, bx.y); p *= av.y; qr = fmaf(p, cdy.y, qr);
        q = fmaf(av.x, q, bx.x); p *= av.x; qr = fmaf(p, cdy.x, qr);
    }
}
// adjoint walk over one pair, AGAINST time order: g_l = c_l + G, G <- a_l g_l
template <bool REV>
__device__ __forceinline__ void walk_adjoint(const float2 av, const float2 cdy, float &G, float2 &gl) {
    if (!REV) {
        gl.y = cdy.y + G; G = av.y * gl.y;
        gl.x = cdy.x + G; G = av.x * gl.x;
    } else {
        gl.x = cdy.x + G; G = av.x * gl.x;
        gl.y = cdy.y + G; G = av.y * gl.y;
    }
}

// ---- bank-conflict-free variants ---------------------------------------------------------------------------------
// A thread owns 32 contiguous bytes of a shared-memory row; a plain 128-bit access at a 32-byte thread stride makes lanes
// t and t + 4 of every quarter-warp hit the same banks (2-way conflict: ncu counted 44 % of all shared wavefronts as
// conflicts).  With sel = (thread >> 2) & 1, lanes with sel = 1 touch their SECOND 16-byte half first: every quarter-warp
// then covers all 32 banks once.
//   lds8_sw : data laid out linearly (written by TMA): swizzled access order + a register swap, result in logical order;
//   *_priv  : data only this thread writes and reads back: logical half h lives at physical half h ^ sel, no swap.
__device__ __forceinline__ void lds8_sw(const float *p, int sel, float2 (&v)[4]) {
    const float4 a = *reinterpret_cast<const float4 *>(p + 4 * sel);
    const float4 b = *reinterpret_cast<const float4 *>(p + 4 * (1 - sel));
    const float4 lo = sel ? b : a, hi = sel ? a : b;
    v[0] = make_float2(lo.x, lo.y);
    v[1] = make_float2(lo.z, lo.w);
    v[2] = make_float2(hi.x, hi.y);
    v[3] = make_float2(hi.z, hi.w);
}
__device__ __forceinline__ void lds8_priv(const float *p, int sel, float2 (&v)[4]) {
    const float4 lo = *reinterpret_cast<const float4 *>(p + 4 * sel);
    const float4 hi = *reinterpret_cast<const float4 *>(p + 4 * (1 - sel));
    v[0] = make_float2(lo.x, lo.y);
    v[1] = make_float2(lo.z, lo.w);
    v[2] = make_float2(hi.x, hi.y);
    v[3] = make_float2(hi.z, hi.w);
}
__device__ __forceinline__ void sts8_priv(float *p, int sel, const float2 (&v)[4]) {
    *reinterpret_cast<float4 *>(p + 4 * sel) = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
    *reinterpret_cast<float4 *>(p + 4 * (1 - sel)) = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
}

// softplus of a position pair in the log2 domain.  x2 = (delta + bias) * log2(e); returns dt2 = softplus * log2(e).
// lg2(1 + e) from the MUFU above e = 1/32, the alternating series below it (1 + e would round the small e away);
// identity above the reference's threshold of 20 (selective_scan_fwd_kernel.cuh:117), tested on x2.
// `e` and `s = 1 + e` are handed back for the backward's sigmoid.
__device__ __forceinline__ float2 softplus2_pair(float2 x2, float2 &e, float2 &s) {
    constexpr float k1 = kLog2e, k2 = -0.5f * kLog2e, k3 = kLog2e / 3.0f, k4 = -0.25f * kLog2e;
    e = make_float2(ex2_approx(x2.x), ex2_approx(x2.y));
    s = add2(e, f2(1.0f));
    const float2 lg = make_float2(lg2_approx(s.x), lg2_approx(s.y));
    float2 pl = fma2(e, f2(k4), f2(k3));
    pl = fma2(e, pl, f2(k2));
    pl = fma2(e, pl, f2(k1));
    pl = mul2(e, pl);
    float2 r;
    r.x = (e.x < 0.03125f) ? pl.x : lg.x;
    r.y = (e.y < 0.03125f) ? pl.y : lg.y;
    r.x = (x2.x > kSoftplusThr2) ? x2.x : r.x;
    r.y = (x2.y > kSoftplusThr2) ? x2.y : r.y;
    return r;
}

// Inclusive scan of affine maps over the lanes of a warp, ascending lanes = ascending time.
// v <- compose(value of lane - off, v) wherever lane - off exists; the shuffle's predicate does the test.
__device__ __forceinline__ void scan_step_up(float &p, float &q, int off) {
    asm volatile(
        "{\n"
        ".reg .pred in;\n"
        ".reg .f32 tp, tq;\n"
        "shfl.sync.up.b32 tp|in, %0, %2, 0, 0xffffffff;\n"
        "shfl.sync.up.b32 tq, %1, %2, 0, 0xffffffff;\n"
        "@in fma.rn.f32 %1, %0, tq, %1;\n"
        "@in mul.rn.f32 %0, %0, tp;\n"
        "}\n"
        : "+f"(p), "+f"(q)
        : "r"(off));
}
// Same with descending lanes = ascending time (the adjoint recurrence runs right to left).
__device__ __forceinline__ void scan_step_down(float &p, float &q, int off) {
    asm volatile(
        "{\n"
        ".reg .pred in;\n"
        ".reg .f32 tp, tq;\n"
        "shfl.sync.down.b32 tp|in, %0, %2, 31, 0xffffffff;\n"
        "shfl.sync.down.b32 tq, %1, %2, 31, 0xffffffff;\n"
        "@in fma.rn.f32 %1, %0, tq, %1;\n"
        "@in mul.rn.f32 %0, %0, tp;\n"
        "}\n"
        : "+f"(p), "+f"(q)
        : "r"(off));
}
template <int WIDTH>
__device__ __forceinline__ Aff warp_scan_up_fast(Aff v) {
#pragma unroll
    for (int off = 1; off < WIDTH; off <<= 1) scan_step_up(v.p, v.q, off);
    return v;
}
template <int WIDTH>
__device__ __forceinline__ Aff warp_scan_down_fast(Aff v) {
#pragma unroll
    for (int off = 1; off < WIDTH; off <<= 1) scan_step_down(v.p, v.q, off);
    return v;
}
// value of the lane below / above, identity at the warp's edge
__device__ __forceinline__ Aff shift_up1(Aff v, int lane) {
    Aff r = {__shfl_up_sync(0xffffffffu, v.p, 1), __shfl_up_sync(0xffffffffu, v.q, 1)};
    if (lane == 0) r = Aff{1.0f, 0.0f};
    return r;
}
__device__ __forceinline__ Aff shift_down1(Aff v, int lane) {
    Aff r = {__shfl_down_sync(0xffffffffu, v.p, 1), __shfl_down_sync(0xffffffffu, v.q, 1)};
    if (lane == 31) r = Aff{1.0f, 0.0f};
    return r;
}

// Sums of three per-thread values over the warp in 6 shuffles: the first two butterfly steps fold four slots
// (the fourth is zero) down to one per lane, the remaining three steps are ordinary.  Lane L ends up holding the sum
// of slot (L >> 3) & 3 ... only lanes 0, 8, 16 (slots 0, 1, 2) are meaningful; see the caller.
__device__ __forceinline__ float warp_sum3(float v0, float v1, float v2, int lane) {
    // step 1 (xor 16): lanes with bit 4 clear keep slots {0, 1}, the others keep {2, 3}
    const bool hi = lane & 16;
    const float send0 = hi ? v0 : v2, send1 = hi ? v1 : 0.0f;
    float k0 = hi ? v2 : v0, k1 = hi ? 0.0f : v1;
    k0 += __shfl_xor_sync(0xffffffffu, send0, 16);
    k1 += __shfl_xor_sync(0xffffffffu, send1, 16);
    // step 2 (xor 8): bit 3 clear keeps the first of the pair
    const bool mid = lane & 8;
    const float send = mid ? k0 : k1;
    float k = mid ? k1 : k0;
    k += __shfl_xor_sync(0xffffffffu, send, 8);
    k += __shfl_xor_sync(0xffffffffu, k, 4);
    k += __shfl_xor_sync(0xffffffffu, k, 2);
    k += __shfl_xor_sync(0xffffffffu, k, 1);
    return k;  // lane bits (4,3): 00 -> slot 0, 01 -> slot 1, 10 -> slot 2, 11 -> slot 3 (zero)
}

// The same tree with the fourth slot in use: lanes 0, 8, 16, 24 hold the sums of v0 .. v3.
__device__ __forceinline__ float warp_sum4(float v0, float v1, float v2, float v3, int lane) {
    const bool hi = lane & 16;
    const float send0 = hi ? v0 : v2, send1 = hi ? v1 : v3;
    float k0 = hi ? v2 : v0, k1 = hi ? v3 : v1;
    k0 += __shfl_xor_sync(0xffffffffu, send0, 16);
    k1 += __shfl_xor_sync(0xffffffffu, send1, 16);
    const bool mid = lane & 8;
    const float send = mid ? k0 : k1;
    float k = mid ? k1 : k0;
    k += __shfl_xor_sync(0xffffffffu, send, 8);
    k += __shfl_xor_sync(0xffffffffu, k, 4);
    k += __shfl_xor_sync(0xffffffffu, k, 2);
    k += __shfl_xor_sync(0xffffffffu, k, 1);
    return k;
}

// named barrier over the `count` threads of one row segment (ids 1..8; id 0 is __syncthreads)
__device__ __forceinline__ void row_barrier(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

}  // namespace vmasr
