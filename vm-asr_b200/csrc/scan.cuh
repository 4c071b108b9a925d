// Selective-scan machinery shared by the forward and backward kernels.
//
// Work decomposition (differs on purpose from the reference's grid=(batch, dim) with a serial chunk loop,
// selective_scan_fwd_kernel.cuh:80-102): a CTA owns a TILE = (batch b, B/C group g, a range of the group's
// channels, one chunk of VMASR_SCAN_CHUNK positions).  The sequence axis is therefore split ACROSS CTAs and
// the carry between chunks travels through a small global exchange area with a decoupled look-back
// (publish the chunk's aggregate first, then combine the predecessors' aggregates), so one pass over HBM
// is enough.  Inside a tile every thread owns ITEMS consecutive positions and walks the tile's channels
// serially: the positions' B/C values (and, backward, their dB/dC sums) stay in registers across channels.
#pragma once
#include "common.cuh"

namespace vmasr {

constexpr float kLog2e = 1.4426950408889634f;

// Device view of one scan call.
struct ScanArgs {
    const void *u, *delta, *B, *C, *dout;
    const float *A, *D, *delta_bias;
    void *out, *du, *ddelta;
    float *x, *dA, *dB, *dC, *dD, *ddelta_bias;
    // carry exchange
    unsigned *ws_header;  // {ticket, done, epoch, pad}
    unsigned *ws_flags;
    float2 *ws_payload;
    int batch, dim, seqlen, dstate, ngroups;
    int n_chunks;         // ceil(seqlen / chunk)
    int chan_per_group;   // dim / ngroups
    int chan_per_tile;    // channels one CTA walks (multiple of ROWS)
    int n_ctiles;         // ceil(chan_per_group / chan_per_tile)
    int n_rowgroups;      // batch * ngroups * n_ctiles
    int softplus;
    long long u_bs, u_ds, delta_bs, delta_ds, A_ds, A_ns, B_bs, B_gs, B_ns, C_bs, C_gs, C_ns;
    long long out_bs, out_ds, dout_bs, dout_ds, du_bs, du_ds, ddelta_bs, ddelta_ds;
};

// (P, Q) represents the affine map  s -> P*s + Q  of a run of positions on the recurrence state.
// compose(first, then) = first applied, then `then`:  s -> Pt*(Pf*s+Qf)+Qt.
struct Aff {
    float p, q;
};
__device__ __forceinline__ Aff compose(Aff first, Aff then) { return {then.p * first.p, fmaf(then.p, first.q, then.q)}; }

// Inclusive scan over the lanes of a warp, lane 0 first (time order = lane order).
__device__ __forceinline__ Aff warp_scan_up(Aff v, int lane) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        float pp = __shfl_up_sync(0xffffffffu, v.p, off);
        float pq = __shfl_up_sync(0xffffffffu, v.q, off);
        if (lane >= off) {
            v.q = fmaf(v.p, pq, v.q);
            v.p *= pp;
        }
    }
    return v;
}
// Same, time order = descending lane order (used by the adjoint recurrence that runs right to left).
__device__ __forceinline__ Aff warp_scan_down(Aff v, int lane) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        float pp = __shfl_down_sync(0xffffffffu, v.p, off);
        float pq = __shfl_down_sync(0xffffffffu, v.q, off);
        if (lane + off < 32) {
            v.q = fmaf(v.p, pq, v.q);
            v.p *= pp;
        }
    }
    return v;
}

// ---- chunk-carry exchange --------------------------------------------------------------------------
// One entry per (batch, channel, state, chunk): payload = the chunk's own affine map (p, q), flag = epoch tag.
// A chunk publishes its map as soon as its local scan is done and never waits before publishing, so there
// is no dependency chain between chunks.  The state entering chunk i is obtained by composing the maps of
// ALL chunks before it, always in the same fixed tree (32-entry windows reduced by shuffles, windows folded
// nearest first).  Reading every predecessor instead of stopping at the first "inclusive" one costs a few
// hundred bytes per chunk and buys run-to-run bit-reproducible results.
// The workspace is zero-filled once; every launch that uses it reads the epoch from the header and the
// last CTA to finish bumps it (and rewinds the ticket counter), so flags of earlier launches are never
// mistaken for current ones and nothing has to be cleared between launches.
constexpr int kMaxWindows = 128;  // 32 chunks each: sequences up to 128*32*2048 positions

__device__ __forceinline__ void publish(const ScanArgs &a, long long entry, unsigned tag, float p, float q) {
    st_relaxed_f2(a.ws_payload + entry, make_float2(p, q));
    st_release_u32(a.ws_flags + entry, tag);
}

// Whole warp: composite map of predecessors 32*win+1 .. 32*win+32 (those that exist), nearest = lane 0.
// `step` is +1 when predecessors are the lower-numbered chunks (forward scan), -1 when they are the
// higher-numbered ones (adjoint scan).  Result valid in every lane.
__device__ __forceinline__ Aff window_map(const ScanArgs &a, long long entry0, int chunk, int step, int n_before, int win,
                                          unsigned tag, int lane) {
    const int k = win * 32 + lane + 1;
    Aff v = {1.0f, 0.0f};
    if (k <= n_before) {
        const long long e = entry0 + (long long)(chunk - step * k);
        unsigned f = ld_acquire_u32(a.ws_flags + e);
        while (f != tag) {
            __nanosleep(20);
            f = ld_acquire_u32(a.ws_flags + e);
        }
        const float2 pl = ld_relaxed_f2(a.ws_payload + e);
        v = {pl.x, pl.y};
    }
    __syncwarp();
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const float pp = __shfl_down_sync(0xffffffffu, v.p, off);
        const float pq = __shfl_down_sync(0xffffffffu, v.q, off);
        if (lane + off < 32) {  // lane+off is farther back: it is applied first
            v.q = fmaf(v.p, pq, v.q);
            v.p *= pp;
        }
    }
    return {__shfl_sync(0xffffffffu, v.p, 0), __shfl_sync(0xffffffffu, v.q, 0)};
}

// Claim a tile.  With more than one chunk per sequence the order in which tiles start matters for forward
// progress of the look-back (a tile only ever waits on tiles that started earlier), so tiles are handed
// out by an atomic ticket in chunk-major order.
__device__ __forceinline__ void claim_tile(const ScanArgs &a, unsigned *s_tile, unsigned &tile, unsigned &epoch) {
    if (a.n_chunks > 1) {
        if (threadIdx.x == 0) {
            s_tile[0] = atomicAdd(a.ws_header + 0, 1u);
            s_tile[1] = *reinterpret_cast<volatile unsigned *>(a.ws_header + 2) % 0xfffffffeu + 1u;  // epoch tag, never 0
        }
        __syncthreads();
        tile = s_tile[0];
        epoch = s_tile[1];
    } else {
        tile = blockIdx.x;
        epoch = 0;
    }
}

// Last CTA out recycles the workspace for the next launch on this stream.
__device__ __forceinline__ void retire_tile(const ScanArgs &a) {
    if (a.n_chunks > 1) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned prev = atomicAdd(a.ws_header + 1, 1u);
            if (prev == gridDim.x - 1) {
                a.ws_header[0] = 0u;
                a.ws_header[1] = 0u;
                a.ws_header[2] = a.ws_header[2] + 1u;
                __threadfence();
            }
        }
    }
}

// host side (scan_host.cu)
struct ScanPlan {
    int tpr;            // threads per row segment
    int rows;           // channel rows a CTA processes at once
    int items;          // positions per thread
    int threads;        // CTA size
    bool vec;           // 128-bit IO possible
    int grid;
};

}  // namespace vmasr
