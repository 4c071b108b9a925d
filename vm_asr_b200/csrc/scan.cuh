// Selective-scan machinery shared by the forward and backward kernels.
//
// Work decomposition (differs on purpose from the reference's grid=(batch, dim) with a serial chunk loop,
// selective_scan_fwd_kernel.cuh:80-102): a CTA owns a TILE = (batch b, B/C group g, a range of the group's
// channels, one chunk of VMASR_SCAN_CHUNK positions).  The sequence axis is therefore split ACROSS CTAs and
// the carry between chunks travels through a small global exchange area with a decoupled two-level look-back
// (publish the chunk's aggregate first, then combine the predecessors' aggregates; pipe.cuh), so one pass
// over HBM is enough.  Inside a tile every thread owns ITEMS consecutive positions and walks the tile's channels
// serially: the positions' B/C values (and, backward, their dB/dC sums) stay in registers across channels.
#pragma once
#include <cuda.h>  // CUtensorMap (type only; the encoder is fetched from the driver at run time, nothing links against libcuda)

#include "common.cuh"

namespace vmasr {

constexpr float kLog2e = 1.4426950408889634f;

struct CarryEntry;

// Device view of one scan call.
struct ScanArgs {
    const void *u, *delta, *B, *C, *dout;
    const float *A, *D, *delta_bias;
    void *out, *du, *ddelta;
    float *x, *dA, *dB, *dC, *dD, *ddelta_bias;
    // carry exchange
    unsigned *ws_header;  // {ticket, done, epoch, pad}
    CarryEntry *ws_entries;   // level 1: one entry per (batch, channel, state, chunk)
    CarryEntry *ws_entries2;  // level 2: one entry per (batch, channel, state, group of 16 chunks)
    int batch, dim, seqlen, dstate, ngroups;
    int n_chunks;         // ceil(seqlen / chunk)
    int chan_per_group;   // dim / ngroups
    int chan_per_tile;    // channels one CTA walks (multiple of ROWS)
    int n_ctiles;         // ceil(chan_per_group / chan_per_tile)
    int n_rowgroups;      // batch * ngroups * n_ctiles
    int n_tiles;          // tiles of this problem in the launch: n_chunks * n_rowgroups + the tiles split_from adds
    int split_from;       // multi-chunk fast kernels: tiles from this index on are HALF tiles (two per planned tile, chan_per_tile / 2
                          // channels each): the launch's last, partly filled round of resident CTAs then takes half as long (scan_host.cu)
    int softplus;
    int rev;              // fast kernels only: the scan runs over the row back to front (time index l <-> seqlen - 1 - l); every positional
                          // tensor (u, delta, B, C, out, dout, du, ddelta, dB, dC) keeps its memory order.  Directions 2 and 3 of SS2D.
    int accum;            // fast kernels only: `out` (forward) / `du` (backward) are added into instead of stored:
                          // 1 = 128-bit red.global.add (concurrent writers), 2 = load / add / store (this launch is the only writer)
    int dbdc_store;       // multi-chunk backward, one channel tile per group: dB / dC are stored, not added into (VMASR_SCAN_DBDC_STORE)
    int pdl_mode;         // VMASR_TUNING builds only (VMASR_PDL_X): 1 = griddepcontrol.wait at the very top of the kernel (0: behind
                          // the CTA's own set-up).  A prefetch.tensormap of the tile's maps in front of the wait was measured 0.5 % SLOWER.
    int debug_nowait;     // VMASR_TUNING builds only, timing experiment: do not wait for neighbours' aggregates -> WRONG results
    unsigned long long *timeline;  // VMASR_TUNING builds only: 16 timestamps per CTA (common.cuh), else null
    long long u_bs, u_ds, delta_bs, delta_ds, A_ds, A_ns, B_bs, B_gs, B_ns, C_bs, C_gs, C_ns;
    long long out_bs, out_ds, dout_bs, dout_ds, du_bs, du_ds, ddelta_bs, ddelta_ds;
    // delta on the fly (include/vmasr_b200.h, dt_rank > 0; multi-chunk fast kernels, rank 1): delta = dt_w[d] * dt_rows[b, g, l]
    const float *dt_w;      // (dim, dt_rank)
    float *d_dt_rows;       // (batch, ngroups * dt_rank, seqlen) with the strides below, accumulated into
    float *d_dt_w;          // like dt_w, accumulated into
    long long dtr_bs, dtr_rs, dtw_ds;
    long long dB_bs, dC_bs;  // batch strides of dB / dC (default ngroups * seqlen)
    int dt_rank;
    // side job (vmasr_scan_params.zero_ptr): tile t of this problem clears 16-byte units [t * zero_per_tile, (t + 1) * zero_per_tile)
    // of the region, cut at zero_n.  zero_n == 0: nothing to do (or the host has queued a memset instead).
    float4 *zero_ptr;
    long long zero_n, zero_per_tile;
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
// 16-byte entries {p, tag, q, tag}: an affine map with the launch's epoch tag repeated in each 8-byte half.
// An entry is valid when both tags equal the current tag; 8-byte accesses are single-copy atomic, so a
// half-written entry can never validate with stale numbers and no fence is needed on either side (the data
// validates itself, there is no separate flag to order against).  The workspace is zero-filled once; every
// launch that uses it reads the epoch from the header and bumps it once every CTA of the launch has read it (the multi-chunk
// fast kernels: launch_epoch() in pipe.cuh, at the start of a tile; the generic kernels: retire_tile(), at the end), so
// entries of earlier launches never validate and nothing has to be cleared between launches.  How the entries are
// organised (one per chunk plus one per group of 16 chunks) and combined is in pipe.cuh.
struct __align__(16) CarryEntry {
    float p;
    unsigned tag0;
    float q;
    unsigned tag1;
};

__device__ __forceinline__ uint4 load_entry(const CarryEntry *e) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(e) : "memory");
    return v;
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
            if (prev == (unsigned)(a.n_chunks * a.n_rowgroups) - 1u) {
                a.ws_header[0] = 0u;
                a.ws_header[1] = 0u;
                a.ws_header[2] = a.ws_header[2] + 1u;
                __threadfence();
            }
        }
    }
}

// One launch over up to kMaxGroup independent scan problems (same kernel variant): tiles of problem i are the block indices
// [tile_end[i - 1], tile_end[i]).  Each problem has its own carry workspace.  Used for the two streams of the generator, which
// issue same-shape calls independently (model/model.py:1167-1176), and for the four directions of the fused SS2D core.
constexpr int kMaxGroup = 8;
// TMA descriptors of one problem's positional inputs.  Each (batch, channel) row of L floats is described as a 2-D array of
// L / 16 lines of 16 floats (64 bytes) so that the copy can land in shared memory with the 64-byte swizzle: 16-byte chunk c of
// 128-byte line r goes to chunk c ^ (r & 3) of its 64-byte half.  A thread owns 32 contiguous bytes; with the swizzle every
// quarter-warp's 128-bit accesses cover all 32 banks once, with no address arithmetic beyond two per-thread constants and no
// register shuffling (round 1 copied rows linearly and paid a register select per loaded float for the same effect).
// Out-of-range lines (ragged last chunk) arrive as zeros.
struct alignas(64) TileMaps {
    CUtensorMap u, delta, dout, B, C;  // with dt_rank > 0 `delta` describes dt_rows: (L, ngroups * dt_rank rows, batch)
};
struct GroupArgs {
    TileMaps tm[kMaxGroup];
    ScanArgs a[kMaxGroup];
    int tile_end[kMaxGroup];
    int n;
};
constexpr int kTileLine = 16;  // floats per line of the tensor maps (64-byte swizzle span)

// shared-memory slot (in units of a thread's 8 floats) of the thread that owns segment `tseg`, and whether its two 16-byte
// halves are swapped: physical byte = o ^ (((o >> 7) & 3) << 4) for logical byte o = 32 tseg + 16 h
__device__ __forceinline__ int swz_slot(int tseg) { return tseg ^ ((tseg >> 3) & 1); }
__device__ __forceinline__ int swz_half(int tseg) { return (tseg >> 2) & 1; }
// problem of this CTA and its tile index inside the problem
__device__ __forceinline__ int group_problem(const GroupArgs &ga, int &tile) {
    int prob = 0, start = 0;
    while (prob + 1 < ga.n && (int)blockIdx.x >= ga.tile_end[prob]) start = ga.tile_end[prob++];
    tile = (int)blockIdx.x - start;
    return prob;
}

// side job of a forward launch (vmasr_scan_params.zero_ptr): this tile's share of the region, 128-bit stores by every thread.
// Issued at the top of the CTA: the stores drain while the tile waits for its first bytes.
__device__ __forceinline__ void zero_side_region(const ScanArgs &a, int tile) {
    if (a.zero_n == 0) return;
    const long long i0 = (long long)tile * a.zero_per_tile;
    const long long left = a.zero_n - i0;
    const int n = (int)(left < a.zero_per_tile ? (left < 0 ? 0 : left) : a.zero_per_tile);
    float4 *dst = a.zero_ptr + i0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
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
