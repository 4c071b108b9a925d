// Tile feed of the persistent scan kernels (scan_fwd_v2.cu / scan_bwd_v2.cu).
//
// A CTA is 256 threads, two per SM, and lives for the whole launch.  Work is cut into TILES = (chunk of the
// sequence, batch, B/C group, range of the group's channels), numbered in scan order (chunk-major; the adjoint
// numbers the chunks from the end).  Thread 0 is the feeder, inline with its share of the arithmetic:
//   * it posts the ids of the CTA's next tiles (dealt round-robin) in a small shared-memory queue;
//   * after the barrier of row pass k (everyone has pulled stage k % NSTAGES into registers) it issues the TMA
//     bulk copies (cp.async.bulk -> UBLKCP) of row pass k + NSTAGES into that stage, walking a cursor through the
//     queued tiles, so the copies run NSTAGES passes ahead, across tile boundaries;
//   * when a tile starts it issues the B / C segment of the tile after the next into the B/C slot just freed.
// All feeder state lives in shared memory so that it costs the other 255 threads no registers.
#pragma once
#include "pipe.cuh"

namespace vmasr {

constexpr int kQueue = 8;

// What a consumer needs to know about a tile, decoded once by the feeder.
struct __align__(16) TileDesc {
    int tile;    // -1: no more work
    int chunk;   // chunk of the sequence
    int j;       // its scan-order index (== chunk forward, n_chunks - 1 - chunk for the adjoint)
    int n_chan;  // channels in the tile
    int seq0;    // batch * dim + first channel
    int bg;      // batch * ngroups + group
    int b, d0;
};

struct FeedState {
    TileDesc q[kQueue];  // by per-CTA sequence number (mod kQueue)
    int claimed;         // sequence numbers claimed so far
    int c_m, c_it, c_niter, c_nchan, c_valid;  // cursor: tile sequence number, row pass in it, its passes / channels
    unsigned c_seg_bytes;
    const float *c_src[3];  // first row of the cursor's tile in each streamed tensor
    unsigned k_issue;       // row passes issued so far (stage = k_issue % NSTAGES)
};

struct TileCoord {
    int chunk, j, b, g, d0, n_chan, seg0;
};

template <bool REVERSE, int SEG>
__device__ __forceinline__ TileCoord decode_tile(const ScanArgs &a, int tile) {
    TileCoord t;
    const unsigned ut = (unsigned)tile;
    t.j = (int)(ut / (unsigned)a.n_rowgroups);  // scan-order index of the chunk
    t.chunk = REVERSE ? a.n_chunks - 1 - t.j : t.j;
    const unsigned rg = ut - (unsigned)t.j * (unsigned)a.n_rowgroups;
    const unsigned bg = rg / (unsigned)a.n_ctiles;
    const int ctile = (int)(rg - bg * (unsigned)a.n_ctiles);
    t.b = (int)(bg / (unsigned)a.ngroups);
    t.g = (int)(bg - (unsigned)t.b * (unsigned)a.ngroups);
    const int c_begin = ctile * a.chan_per_tile;
    t.n_chan = min(a.chan_per_group, c_begin + a.chan_per_tile) - c_begin;
    t.d0 = t.g * a.chan_per_group + c_begin;
    t.seg0 = t.chunk * SEG;
    return t;
}

// Tiles are dealt round-robin: sequence number m of CTA c is tile c + m * gridDim.x.  Tile t's look-back needs
// tiles t - k * n_rowgroups, which sit at the same or an earlier sequence number of their CTA, and a chunk
// publishes its own aggregate before it waits for anybody, so the wave of CTAs moves through the sequence
// together.  This needs every CTA of the grid resident at once: the host sizes the grid from the occupancy API.
template <bool REVERSE, int SEG>
__device__ __forceinline__ void feed_claim(const ScanArgs &a, FeedState &fs) {
    const long long t = (long long)blockIdx.x + (long long)fs.claimed * gridDim.x;
    TileDesc d;
    d.tile = -1;
    if (t < a.n_tiles) {
        const TileCoord c = decode_tile<REVERSE, SEG>(a, (int)t);
        d.tile = (int)t;
        d.chunk = c.chunk;
        d.j = c.j;
        d.n_chan = c.n_chan;
        d.seq0 = c.b * a.dim + c.d0;
        d.bg = c.b * a.ngroups + c.g;
        d.b = c.b;
        d.d0 = c.d0;
    }
    fs.q[fs.claimed % kQueue] = d;
    fs.claimed += 1;
}

// NARR streamed tensors: 0 = u, 1 = delta, 2 = dout
template <bool REVERSE, int SEG, int NARR>
__device__ __forceinline__ void feed_open_tile(const ScanArgs &a, FeedState &fs, int rows) {
    const int tile = fs.q[fs.c_m % kQueue].tile;
    fs.c_it = 0;
    fs.c_valid = tile >= 0;
    if (tile < 0) return;
    const TileCoord t = decode_tile<REVERSE, SEG>(a, tile);
    fs.c_nchan = t.n_chan;
    fs.c_niter = (t.n_chan + rows - 1) / rows;
    fs.c_seg_bytes = (unsigned)min(SEG, a.seqlen - t.seg0) * 4u;
    fs.c_src[0] = reinterpret_cast<const float *>(a.u) + t.b * a.u_bs + (long long)t.d0 * a.u_ds + t.seg0;
    fs.c_src[1] = reinterpret_cast<const float *>(a.delta) + t.b * a.delta_bs + (long long)t.d0 * a.delta_ds + t.seg0;
    if (NARR > 2) fs.c_src[2] = reinterpret_cast<const float *>(a.dout) + t.b * a.dout_bs + (long long)t.d0 * a.dout_ds + t.seg0;
}

// Issue the copies of the cursor's row pass into stage k_issue % NSTAGES and advance the cursor.
// Stage layout: [tensor][row][SEG] floats.
template <bool REVERSE, int SEG, int ROWS, int NARR, int NSTAGES>
__device__ __forceinline__ void feed_issue_pass(const ScanArgs &a, FeedState &fs, float *stages, unsigned long long *full) {
    if (!fs.c_valid) return;
    const int s = (int)(fs.k_issue % NSTAGES);
    const int it = fs.c_it;
    const int rows_here = min(ROWS, fs.c_nchan - it * ROWS);
    const unsigned seg_bytes = fs.c_seg_bytes;
    float *dst = stages + (size_t)s * (NARR * ROWS * SEG);
    mbar_expect_tx(&full[s], (unsigned)NARR * seg_bytes * (unsigned)rows_here);
    const long long ds[3] = {a.u_ds, a.delta_ds, a.dout_ds};
#pragma unroll
    for (int t = 0; t < NARR; ++t) {
        const float *src = fs.c_src[t] + (long long)(it * ROWS) * ds[t];
        if (ROWS > 1 && ds[t] == SEG && seg_bytes == SEG * 4u) {
            bulk_load(dst + t * ROWS * SEG, src, seg_bytes * (unsigned)rows_here, &full[s]);  // rows back to back in memory
        } else {
            for (int r = 0; r < rows_here; ++r) bulk_load(dst + (t * ROWS + r) * SEG, src + (long long)r * ds[t], seg_bytes, &full[s]);
        }
    }
    fs.k_issue += 1;
    fs.c_it = it + 1;
    if (fs.c_it == fs.c_niter) {
        fs.c_m += 1;
        feed_open_tile<REVERSE, SEG, NARR>(a, fs, ROWS);
    }
}

template <bool REVERSE, int SEG>
__device__ __forceinline__ void feed_issue_bc(const ScanArgs &a, int tile, float *slot, unsigned long long *bar) {
    const TileCoord t = decode_tile<REVERSE, SEG>(a, tile);
    const unsigned seg_bytes = (unsigned)min(SEG, a.seqlen - t.seg0) * 4u;
    mbar_expect_tx(bar, 2u * seg_bytes);
    bulk_load(slot, reinterpret_cast<const float *>(a.B) + t.b * a.B_bs + t.g * a.B_gs + t.seg0, seg_bytes, bar);
    bulk_load(slot + SEG, reinterpret_cast<const float *>(a.C) + t.b * a.C_bs + t.g * a.C_gs + t.seg0, seg_bytes, bar);
}

// Prologue, thread 0: barriers, first claims, B/C of the first tile, the first NSTAGES row passes.
template <bool REVERSE, int SEG, int ROWS, int NARR, int NSTAGES>
__device__ __forceinline__ void feed_start(const ScanArgs &a, FeedState &fs, float *stages, unsigned long long *full, float *bc_slots,
                                           unsigned long long *bc_full) {
    for (int s = 0; s < NSTAGES; ++s) mbar_init(&full[s], 1);
    mbar_init(&bc_full[0], 1);
    mbar_init(&bc_full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fs.claimed = 0;
    fs.k_issue = 0;
    for (int i = 0; i < NSTAGES + 2; ++i) feed_claim<REVERSE, SEG>(a, fs);
    if (fs.q[0].tile >= 0) feed_issue_bc<REVERSE, SEG>(a, fs.q[0].tile, bc_slots, &bc_full[0]);
    if (fs.q[1].tile >= 0) feed_issue_bc<REVERSE, SEG>(a, fs.q[1].tile, bc_slots + 2 * SEG, &bc_full[1]);
    fs.c_m = 0;
    feed_open_tile<REVERSE, SEG, NARR>(a, fs, ROWS);
    for (int i = 0; i < NSTAGES; ++i) feed_issue_pass<REVERSE, SEG, ROWS, NARR, NSTAGES>(a, fs, stages, full);
}

// Last CTA out recycles the carry-exchange area (done counter, epoch) for the next launch.
__device__ __forceinline__ void retire_cta(const ScanArgs &a) {
    __threadfence();
    const unsigned prev = atomicAdd(a.ws_header + 1, 1u);
    if (prev == gridDim.x - 1) {
        a.ws_header[1] = 0u;
        a.ws_header[2] = a.ws_header[2] + 1u;
        __threadfence();
    }
}

}  // namespace vmasr
