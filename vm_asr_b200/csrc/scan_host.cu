// Host side of the selective scan: argument checks (mirroring the TORCH_CHECKs of
// kernels/selective_scan/csrc/selective_scan/cus/selective_scan.cpp:165-215, 262-317), tile planning, launch.
#include <cstdlib>

#include "scan.cuh"

namespace vmasr {

int scan_fwd_dispatch(const ScanArgs &a, const ScanPlan &pl, int io_dtype, cudaStream_t stream);
int scan_bwd_dispatch(const ScanArgs &a, const ScanPlan &pl, int io_dtype, cudaStream_t stream);
int scan_fwd_tma_dispatch(const ScanArgs &a, const ScanPlan &pl, cudaStream_t stream);
int scan_bwd_tma_dispatch(const ScanArgs &a, const ScanPlan &pl, cudaStream_t stream);
int scan_fwd_pipe_dispatch(const ScanArgs &a, const ScanPlan &pl, cudaStream_t stream);
int scan_bwd_pipe_dispatch(const ScanArgs &a, const ScanPlan &pl, cudaStream_t stream);
int scan_fwd_ring_dispatch(const ScanArgs &a, const ScanPlan &pl, int device, cudaStream_t stream);
constexpr int kMaxTileChannelsHost = 64;  // scan_fwd_tma.cu stages this many channels' parameters per tile

static size_t dtype_size(int dt) { return dt == VMASR_F32 ? 4 : 2; }

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int validate(const vmasr_scan_params *p, bool bwd) {
    if (!p) return fail("selective_scan: null params");
    if (p->io_dtype != VMASR_F32 && p->io_dtype != VMASR_F16 && p->io_dtype != VMASR_BF16)
        return fail("selective_scan: u/delta/B/C must be float32, float16 or bfloat16 (selective_scan.cpp:167)");
    if (p->batch <= 0 || p->dim <= 0 || p->seqlen <= 0 || p->dstate <= 0 || p->ngroups <= 0)
        return fail("selective_scan: sizes must be positive (batch %d dim %d seqlen %d dstate %d ngroups %d)", p->batch,
                    p->dim, p->seqlen, p->dstate, p->ngroups);
    if (p->dim % p->ngroups != 0) return fail("dims should be dividable by n_groups");
    if (p->dstate > 256) return fail("selective_scan only supports state dimension <= 256");
    if (!p->u || !p->delta || !p->A || !p->B || !p->C) return fail("selective_scan: u, delta, A, B, C must be non-null");
    if (!bwd && (!p->out || !p->x)) return fail("selective_scan_fwd: out and x must be non-null");
    if (bwd) {
        if (!p->dout || !p->du || !p->ddelta || !p->dA || !p->dB || !p->dC)
            return fail("selective_scan_bwd: dout, du, ddelta, dA, dB, dC must be non-null");
        if ((p->D != nullptr) != (p->dD != nullptr)) return fail("selective_scan_bwd: dD must be given exactly when D is");
        if ((p->delta_bias != nullptr) != (p->ddelta_bias != nullptr))
            return fail("selective_scan_bwd: ddelta_bias must be given exactly when delta_bias is");
        const int n_chunks = (p->seqlen + VMASR_SCAN_CHUNK - 1) / VMASR_SCAN_CHUNK;
        if (n_chunks > 1 && !p->x) return fail("selective_scan_bwd: x (chunk states) is required when seqlen > %d", VMASR_SCAN_CHUNK);
    }
    const int n_chunks = (p->seqlen + VMASR_SCAN_CHUNK - 1) / VMASR_SCAN_CHUNK;
    if (n_chunks > 1) {
        if (!p->workspace) return fail("selective_scan: a carry workspace is required when seqlen > %d", VMASR_SCAN_CHUNK);
        const uint64_t need = vmasr_scan_workspace_bytes(p->batch, p->dim, p->seqlen, p->dstate);
        if (p->workspace_bytes < need)
            return fail("selective_scan: workspace too small (%llu < %llu bytes)", (unsigned long long)p->workspace_bytes,
                        (unsigned long long)need);
        if (!aligned16(p->workspace)) return fail("selective_scan: workspace must be 16-byte aligned");
    }
    return 0;
}

static ScanPlan make_plan(const vmasr_scan_params *p, int n_chunks, bool bwd, int &chan_per_tile, int &n_ctiles) {
    ScanPlan pl;
    pl.items = 8;
    pl.threads = 256;
    const int L = p->seqlen;
    pl.tpr = L <= 256 ? 32 : L <= 512 ? 64 : L <= 1024 ? 128 : 256;
    pl.rows = pl.threads / pl.tpr;
    const int cpg = p->dim / p->ngroups;
    // enough tiles to fill the machine a few times over, otherwise as many channels per tile as possible
    // (B/C stay in registers across a tile's channels and dB/dC need fewer atomics)
    const long long base_tiles = (long long)p->batch * p->ngroups * n_chunks;
    static const int tiles_per_sm = [] { const char *e = getenv("VMASR_SCAN_TILES_PER_SM"); const int v = e ? atoi(e) : 2; return v < 1 ? 1 : v; }();
    const long long target = (long long)tiles_per_sm * sm_count(p->device);
    const int max_ctiles = (cpg + pl.rows - 1) / pl.rows;
    long long want = (target + base_tiles - 1) / base_tiles;
    if (want < 1) want = 1;
    const long long min_ctiles = (cpg + kMaxTileChannelsHost - 1) / kMaxTileChannelsHost;
    if (want < min_ctiles) want = min_ctiles;
    if (want > max_ctiles) want = max_ctiles;
    chan_per_tile = (int)((cpg + want - 1) / want);
    {
        // Sequences of more than one chunk: the pipelined kernels keep the whole tile resident in shared memory
        // (scan_fwd_pipe.cu / scan_bwd_pipe.cu), at most 4 channels.  VMASR_SCAN_CPT = 1..4 is a tuning knob.
        static const int cap_multi = [] { const char *e = getenv("VMASR_SCAN_CPT"); const int v = e ? atoi(e) : 4; return v < 1 ? 1 : v > 4 ? 4 : v; }();
        static const int cap_multi_fwd = [] { const char *e = getenv("VMASR_SCAN_CPT_FWD"); const int v = e ? atoi(e) : 4; return v < 1 ? 1 : v > 4 ? 4 : v; }();
        const int cap = bwd ? cap_multi : (cap_multi < cap_multi_fwd ? cap_multi : cap_multi_fwd);  // measured best: 4 for both
        if (n_chunks > 1 && chan_per_tile > cap) chan_per_tile = cap;
    }
    chan_per_tile = ((chan_per_tile + pl.rows - 1) / pl.rows) * pl.rows;
    n_ctiles = (cpg + chan_per_tile - 1) / chan_per_tile;
    pl.grid = (int)(base_tiles * n_ctiles);

    const size_t es = dtype_size(p->io_dtype);
    const long long vec_elems = 16 / (long long)es;
    auto mult = [&](long long s) { return s % vec_elems == 0; };
    pl.vec = (L % vec_elems == 0) && aligned16(p->u) && aligned16(p->delta) && aligned16(p->B) && aligned16(p->C) &&
             mult(p->u_batch_stride) && mult(p->u_d_stride) && mult(p->delta_batch_stride) && mult(p->delta_d_stride) &&
             mult(p->B_batch_stride) && mult(p->B_group_stride) && mult(p->B_dstate_stride) && mult(p->C_batch_stride) &&
             mult(p->C_group_stride) && mult(p->C_dstate_stride);
    return pl;
}

static ScanArgs make_args(const vmasr_scan_params *p, int n_chunks, int chan_per_tile, int n_ctiles) {
    ScanArgs a{};
    a.u = p->u; a.delta = p->delta; a.B = p->B; a.C = p->C; a.dout = p->dout;
    a.A = p->A; a.D = p->D; a.delta_bias = p->delta_bias;
    a.out = p->out; a.du = p->du; a.ddelta = p->ddelta;
    a.x = p->x; a.dA = p->dA; a.dB = p->dB; a.dC = p->dC; a.dD = p->dD; a.ddelta_bias = p->ddelta_bias;
    if (p->workspace && n_chunks > 1) {
        // [64 B header {ticket, done, epoch}][16-byte carry entries]
        char *base = static_cast<char *>(p->workspace);
        a.ws_header = reinterpret_cast<unsigned *>(base);
        a.ws_entries = reinterpret_cast<CarryEntry *>(base + 64);
        a.ws_entries2 = a.ws_entries + (long long)p->batch * p->dim * p->dstate * n_chunks;
    }
    a.batch = p->batch; a.dim = p->dim; a.seqlen = p->seqlen; a.dstate = p->dstate; a.ngroups = p->ngroups;
    a.n_chunks = n_chunks;
    a.chan_per_group = p->dim / p->ngroups;
    a.chan_per_tile = chan_per_tile;
    a.n_ctiles = n_ctiles;
    a.n_rowgroups = p->batch * p->ngroups * n_ctiles;
    a.softplus = p->delta_softplus;
    static const int nowait = [] { const char *e = getenv("VMASR_DEBUG_NOWAIT"); return e ? atoi(e) : 0; }();
    a.debug_nowait = nowait;
    a.u_bs = p->u_batch_stride; a.u_ds = p->u_d_stride;
    a.delta_bs = p->delta_batch_stride; a.delta_ds = p->delta_d_stride;
    a.A_ds = p->A_d_stride; a.A_ns = p->A_dstate_stride;
    a.B_bs = p->B_batch_stride; a.B_gs = p->B_group_stride; a.B_ns = p->B_dstate_stride;
    a.C_bs = p->C_batch_stride; a.C_gs = p->C_group_stride; a.C_ns = p->C_dstate_stride;
    a.out_bs = p->out_batch_stride; a.out_ds = p->out_d_stride;
    a.dout_bs = p->dout_batch_stride; a.dout_ds = p->dout_d_stride;
    a.du_bs = p->du_batch_stride; a.du_ds = p->du_d_stride;
    a.ddelta_bs = p->ddelta_batch_stride; a.ddelta_ds = p->ddelta_d_stride;
    return a;
}

enum ScanVariant { kGeneric = 0, kSingleChunk = 1, kMultiChunk = 2, kRing = 3 };

// Everything the launch needs, decided on the host without touching the device (also behind vmasr_scan_plan).
static int decide(const vmasr_scan_params *p, bool bwd, ScanPlan &pl, ScanArgs &a, int &variant) {
    if (int rc = validate(p, bwd)) return rc;
    const int n_chunks = (p->seqlen + VMASR_SCAN_CHUNK - 1) / VMASR_SCAN_CHUNK;
    int chan_per_tile = 1, n_ctiles = 1;
    pl = make_plan(p, n_chunks, bwd, chan_per_tile, n_ctiles);
    a = make_args(p, n_chunks, chan_per_tile, n_ctiles);
    const size_t es = dtype_size(p->io_dtype);
    const long long vec_elems = 16 / (long long)es;
    auto mult = [&](long long s) { return s % vec_elems == 0; };
    if (!bwd) {
        pl.vec = pl.vec && aligned16(p->out) && mult(p->out_batch_stride) && mult(p->out_d_stride);
    } else {
        pl.vec = pl.vec && aligned16(p->dout) && aligned16(p->du) && aligned16(p->ddelta) && aligned16(p->dB) &&
                 aligned16(p->dC) && mult(p->dout_batch_stride) && mult(p->dout_d_stride) && mult(p->du_batch_stride) &&
                 mult(p->du_d_stride) && mult(p->ddelta_batch_stride) && mult(p->ddelta_d_stride) && (p->seqlen % 4 == 0);
    }
    // fast path: fp32, d_state 1, 16-byte aligned rows -> TMA-staged packed-fp32x2 kernels; more than one chunk: the
    // kernels with the exchange warp (scan_*_pipe.cu).  VMASR_SCAN_FWD / VMASR_SCAN_BWD = generic | tma | ring
    static const char fwd_force = [] { const char *e = getenv("VMASR_SCAN_FWD"); return e ? e[0] : '\0'; }();
    static const char bwd_force = [] { const char *e = getenv("VMASR_SCAN_BWD"); return e ? e[0] : '\0'; }();
    const char force = bwd ? bwd_force : fwd_force;
    const bool fast = p->io_dtype == VMASR_F32 && p->dstate == 1 && pl.vec && a.chan_per_tile <= kMaxTileChannelsHost;
    if (!fast || force == 'g') variant = kGeneric;
    else if (n_chunks > 1 && !bwd && force == 'r' && a.chan_per_group % a.chan_per_tile == 0) variant = kRing;
    else if (n_chunks > 1 && force != 't') variant = kMultiChunk;
    else variant = kSingleChunk;
    return 0;
}

static int run(const vmasr_scan_params *p, bool bwd) {
    ScanPlan pl;
    ScanArgs a;
    int variant = kGeneric;
    if (int rc = decide(p, bwd, pl, a, variant)) return rc;
    DeviceGuard guard(p->device);
    if (!guard.ok) return fail("selective_scan: cannot select CUDA device %d", p->device);
    cudaStream_t stream = static_cast<cudaStream_t>(p->stream);
    if (bwd) {
        if (variant == kMultiChunk) return scan_bwd_pipe_dispatch(a, pl, stream);
        if (variant == kSingleChunk) return scan_bwd_tma_dispatch(a, pl, stream);
        return scan_bwd_dispatch(a, pl, p->io_dtype, stream);
    }
    if (variant == kRing) return scan_fwd_ring_dispatch(a, pl, p->device, stream);
    if (variant == kMultiChunk) return scan_fwd_pipe_dispatch(a, pl, stream);
    if (variant == kSingleChunk) return scan_fwd_tma_dispatch(a, pl, stream);
    return scan_fwd_dispatch(a, pl, p->io_dtype, stream);
}

}  // namespace vmasr

extern "C" uint64_t vmasr_scan_workspace_bytes(int batch, int dim, int seqlen, int dstate) {
    if (batch <= 0 || dim <= 0 || seqlen <= 0 || dstate <= 0) return 0;
    const uint64_t n_chunks = ((uint64_t)seqlen + VMASR_SCAN_CHUNK - 1) / VMASR_SCAN_CHUNK;
    if (n_chunks <= 1) return 0;
    const uint64_t n_groups = (n_chunks + 15) / 16;  // level-2 entries of the persistent kernels' look-back
    const uint64_t cap = (uint64_t)batch * dim * dstate * (n_chunks + n_groups);
    const uint64_t bytes = 64 + 16 * cap;
    return (bytes + 255) / 256 * 256;
}

extern "C" int vmasr_scan_plan(const vmasr_scan_params *p, int backward, int32_t *out) {
    if (!out) return vmasr::fail("vmasr_scan_plan: null output");
    vmasr::ScanPlan pl;
    vmasr::ScanArgs a;
    int variant = 0;
    if (int rc = vmasr::decide(p, backward != 0, pl, a, variant)) return rc;
    out[0] = pl.grid;
    out[1] = variant;
    out[2] = a.chan_per_tile;
    out[3] = a.n_ctiles;
    out[4] = a.n_chunks;
    out[5] = pl.tpr;
    return 0;
}

extern "C" int vmasr_scan_fwd(const vmasr_scan_params *p) { return vmasr::run(p, false); }
extern "C" int vmasr_scan_bwd(const vmasr_scan_params *p) { return vmasr::run(p, true); }
