// Pieces shared by the scan kernels (sm_100a): mbarriers and TMA bulk copies (cp.async.bulk -> SASS UBLKCP),
// the two-level chunk-carry look-back, and the log2-domain softplus.
#pragma once
#include "scan.cuh"

namespace vmasr {

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy (TMA), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// tile of a 4-D tensor map (line, lines, row, batch) -> shared memory, completion counted in bytes (the FULL box) on `bar`
__device__ __forceinline__ void tensor_load(void *dst_smem, const CUtensorMap *tm, int line0, int row, int batch, unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
            smem_u32(dst_smem)),
        "l"(tm), "r"(0), "r"(line0), "r"(row), "r"(batch), "r"(smem_u32(bar))
        : "memory");
}
// ---- two-level chunk-carry look-back ---------------------------------------------------------------
// Level 1: one entry per chunk, the chunk's own affine map, published as soon as its local scan is done.
// Level 2: one entry per GROUP of 16 consecutive chunks (in scan order), the composite map of the group,
// published by the group's last chunk once it has seen the 15 level-1 entries before it.  The state entering
// chunk j (scan-order index) is the composition of  level-2 entries of every complete group before it  and
// the level-1 entries of the chunks before it inside its own group: at most 16 + 15 entries, ONE 16-byte
// load per lane, combined in a fixed shuffle tree (bit-reproducible run to run).  Publication never waits on
// anything at level 1 and only on level-1 entries at level 2, so there is no dependency chain along the
// sequence.  The TMA forward kernel issues the load at the top of the row (before any arithmetic) and validates
// it after the local scan.
struct CarryLook {
    const CarryEntry *ptr;  // this lane's entry (nullptr: lane holds the identity)
    uint4 e;
};

__device__ __forceinline__ CarryLook look_issue(const CarryEntry *l1_row, const CarryEntry *l2_row, int j, int lane) {
    CarryLook c;
    const int r = j & 15, gi = j >> 4;
    c.ptr = nullptr;
    if (lane < 16) {
        if (lane < r) c.ptr = l1_row + (j - 1 - lane);
    } else {
        const int i = lane - 16;
        if (i < gi) c.ptr = l2_row + (gi - 1 - i);
    }
    c.e = make_uint4(0u, 0u, 0u, 0u);
    if (c.ptr) c.e = load_entry(c.ptr);
    return c;
}

// Whole warp.  Shuffle-tree composition of the entries the lanes hold, whatever their state: `ok` tells whether
// every lane's entry carried the current tag.
// `ingroup` receives the composite of the level-1 entries before j inside its group (what a group's last chunk
// folds into level 2); the return value is the composite of all 31 slots (everything before chunk j when j < 272).
__device__ __forceinline__ Aff look_reduce(const CarryLook &c, unsigned tag, int lane, bool &ok, Aff &ingroup) {
    Aff v = {1.0f, 0.0f};
    bool mine = true;
    if (c.ptr) {
        mine = (c.e.y == tag) && (c.e.w == tag);
        v = Aff{__uint_as_float(c.e.x), __uint_as_float(c.e.z)};
    }
    ok = __all_sync(0xffffffffu, mine);
    // higher lanes hold maps that apply earlier; fold them in first
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const float pp = __shfl_down_sync(0xffffffffu, v.p, off);
        const float pq = __shfl_down_sync(0xffffffffu, v.q, off);
        if (lane + off < 32) {
            v.q = fmaf(v.p, pq, v.q);
            v.p *= pp;
        }
        if (off == 8) ingroup = Aff{__shfl_sync(0xffffffffu, v.p, 0), __shfl_sync(0xffffffffu, v.q, 0)};
    }
    return Aff{__shfl_sync(0xffffffffu, v.p, 0), __shfl_sync(0xffffffffu, v.q, 0)};
}

// Forward progress.  A tile waits only for entries of tiles that precede it in block-index order (chunk-major launch
// order), and the hardware hands out the blocks of a 1-D grid in ascending index order, so every tile a resident tile waits
// for is resident or finished (the same assumption CUB's decoupled look-back scan makes).  Should that ever not hold (a
// tool that serialises or reorders blocks), the wait is BOUNDED: after a few seconds of polling the kernel traps, which surfaces as
// a launch failure on the stream instead of a hang.
__device__ __forceinline__ uint4 wait_entry(const CarryEntry *p, uint4 e, unsigned tag) {
    unsigned spins = 0;
    while (e.y != tag || e.w != tag) {
        __nanosleep(32);
        e = load_entry(p);
        if (++spins > (1u << 22)) __trap();
    }
    return e;
}

// Slow path, after the chunk's own aggregate is published: wait for the entries that were not there yet, reduce
// again; then (sequences of more than 272 chunks) the remaining level-2 entries, 32 per round, nearest first.
__device__ __forceinline__ Aff look_finish(CarryLook &c, Aff acc, bool ok, const CarryEntry *l2_row, int j, unsigned tag, int lane,
                                           Aff &ingroup) {
    if (!ok) {
        if (c.ptr) c.e = wait_entry(c.ptr, c.e, tag);
        __syncwarp();
        bool again;
        acc = look_reduce(c, tag, lane, again, ingroup);
    }
    const int gi = j >> 4;
    for (int base = 16; base < gi; base += 32) {
        Aff w = {1.0f, 0.0f};
        const int i = base + lane;
        if (i < gi) {
            const CarryEntry *p = l2_row + (gi - 1 - i);
            const uint4 e = wait_entry(p, load_entry(p), tag);
            w = Aff{__uint_as_float(e.x), __uint_as_float(e.z)};
        }
        __syncwarp();
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const float pp = __shfl_down_sync(0xffffffffu, w.p, off);
            const float pq = __shfl_down_sync(0xffffffffu, w.q, off);
            if (lane + off < 32) {
                w.q = fmaf(w.p, pq, w.q);
                w.p *= pp;
            }
        }
        const Aff round = {__shfl_sync(0xffffffffu, w.p, 0), __shfl_sync(0xffffffffu, w.q, 0)};
        acc = compose(round, acc);
    }
    return acc;
}

// Epoch tag of this launch (never 0), read once per CTA by the exchange warp, and the recycling of the workspace for the next
// launch.  A CTA counts itself in `done` right AFTER it has read the epoch (the acquire load orders the two), so the CTA that
// completes the count knows that every CTA of the launch holds the current epoch and bumps it there and then: entries of this
// launch keep the old tag and never validate in the next one.  Nothing is left to do at the END of a tile, where a fence and an
// atomic round trip used to hold the CTA's shared memory for another microsecond; and only one lane waits for the header (the
// compute warps used to, before their first barrier).  Whole warp; call it after the tile's last bulk copy has been issued.
__device__ __forceinline__ unsigned launch_epoch(const ScanArgs &a, int lane) {
    unsigned raw = 0;
    if (lane == 0) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(raw) : "l"(a.ws_header + 2) : "memory");
        const unsigned prev = atomicAdd(a.ws_header + 1, 1u);
        if (prev == (unsigned)a.n_tiles - 1u) {
            a.ws_header[0] = 0u;
            a.ws_header[1] = 0u;
            a.ws_header[2] = raw + 1u;
        }
    }
    raw = __shfl_sync(0xffffffffu, raw, 0);
    return raw % 0xfffffffeu + 1u;
}

__device__ __forceinline__ void publish_entry(CarryEntry *e, unsigned tag, float p, float q) {
    asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(e), "r"(__float_as_uint(p)), "r"(tag),
                 "r"(__float_as_uint(q)), "r"(tag)
                 : "memory");
}

// softplus pieces in the log2 domain: x2 = (delta + bias) * log2(e).  Same accuracy contract as
// softplus_sig() in common.cuh (relative error < 1e-6 everywhere), fewer instructions: the threshold test
// of the reference (identity above 20, fwd_kernel.cuh:117) is done on x2.
constexpr float kSoftplusThr2 = 20.0f * 1.4426950408889634f;
__device__ __forceinline__ float softplus2(float x2, float x) {
    const float e = ex2_approx(x2);
    const float lg = lg2_approx(1.0f + e) * 0.6931471805599453f;
    const float poly = e * fmaf(-e, fmaf(-e, fmaf(-0.25f, e, 0.33333334f), 0.5f), 1.0f);
    float sp = (e < 0.03125f) ? poly : lg;
    return (x2 > kSoftplusThr2) ? x : sp;
}
__device__ __forceinline__ float softplus2_sig(float x2, float x, float &sig) {
    const float e = ex2_approx(x2);
    const float one_pe = 1.0f + e;
    const float lg = lg2_approx(one_pe) * 0.6931471805599453f;
    const float poly = e * fmaf(-e, fmaf(-e, fmaf(-0.25f, e, 0.33333334f), 0.5f), 1.0f);
    float sp = (e < 0.03125f) ? poly : lg;
    const bool big = x2 > kSoftplusThr2;
    sig = big ? 1.0f : __fdividef(e, one_pe);
    return big ? x : sp;
}

}  // namespace vmasr
