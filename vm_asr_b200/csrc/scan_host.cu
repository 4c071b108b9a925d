// Host side of the selective scan: argument checks (mirroring the TORCH_CHECKs of
// kernels/selective_scan/csrc/selective_scan/cus/selective_scan.cpp:165-215, 262-317), tile planning, launch of one
// call or of a GROUP of independent calls as one grid.
#include "scan.cuh"

namespace vmasr {

int scan_fwd_dispatch(const ScanArgs &a, const ScanPlan &pl, int io_dtype, cudaStream_t stream);
int scan_bwd_dispatch(const ScanArgs &a, const ScanPlan &pl, int io_dtype, cudaStream_t stream);
int scan_fwd_tma_dispatch(const GroupArgs &ga, int tpr, int grid, cudaStream_t stream);
int scan_bwd_tma_dispatch(const GroupArgs &ga, int tpr, int grid, cudaStream_t stream);
int scan_fwd_pipe_dispatch(const GroupArgs &ga, int grid, cudaStream_t stream);
int scan_bwd_pipe_dispatch(const GroupArgs &ga, int grid, cudaStream_t stream);
constexpr int kMaxTileChannelsHost = 64;  // the single-chunk kernels stage this many channels' parameters per tile
constexpr int kMultiChunkTileChannels = 4;  // the multi-chunk kernels keep the whole tile resident in shared memory

static size_t dtype_size(int dt) { return dt == VMASR_F32 ? 4 : 2; }

// ---- TMA descriptors of the fast path ------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled is a driver entry point; it is looked up through the runtime so that the library links against
// nothing but the (static) CUDA runtime.  Encoding is pure host arithmetic (no driver call).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// rows of `len` floats (unit stride, len % 16 == 0) indexed by (row, batch) with element strides -> (16, len / 16, rows, batch),
// boxes of `box_lines` lines of one row, 64-byte swizzle, zeros past the end
static int make_tile_map(CUtensorMap *tm, const void *base, long long len, long long rows, long long row_stride, long long batch,
                         long long batch_stride, int box_lines, const char *what) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return fail("selective_scan: cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[4] = {(cuuint64_t)kTileLine, (cuuint64_t)(len / kTileLine), (cuuint64_t)rows, (cuuint64_t)batch};
    // a dimension of extent 1 may carry any stride in the caller's tensor: give it the natural one
    const cuuint64_t line_bytes = kTileLine * sizeof(float);
    const cuuint64_t rs = rows > 1 ? (cuuint64_t)row_stride * sizeof(float) : line_bytes * dims[1];
    const cuuint64_t bs = batch > 1 ? (cuuint64_t)batch_stride * sizeof(float) : rs * dims[2];
    const cuuint64_t strides[3] = {line_bytes, rs, bs};
    const cuuint32_t box[4] = {(cuuint32_t)kTileLine, (cuuint32_t)box_lines, 1u, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("selective_scan: cannot describe %s to the TMA (cuTensorMapEncodeTiled error %d)", what, (int)r);
    return 0;
}

static int make_tile_maps(TileMaps &tm, const vmasr_scan_params *p, bool bwd, int box_lines) {
    const long long L = p->seqlen;
    if (int rc = make_tile_map(&tm.u, p->u, L, p->dim, p->u_d_stride, p->batch, p->u_batch_stride, box_lines, "u")) return rc;
    if (p->dt_rank > 0) {
        if (int rc = make_tile_map(&tm.delta, p->dt_rows, L, (long long)p->ngroups * p->dt_rank, p->dt_rows_row_stride, p->batch, p->dt_rows_batch_stride,
                                   box_lines, "dt_rows"))
            return rc;
    } else if (int rc = make_tile_map(&tm.delta, p->delta, L, p->dim, p->delta_d_stride, p->batch, p->delta_batch_stride, box_lines, "delta")) return rc;
    if (int rc = make_tile_map(&tm.B, p->B, L, p->ngroups, p->B_group_stride, p->batch, p->B_batch_stride, box_lines, "B")) return rc;
    if (int rc = make_tile_map(&tm.C, p->C, L, p->ngroups, p->C_group_stride, p->batch, p->C_batch_stride, box_lines, "C")) return rc;
    if (bwd)
        if (int rc = make_tile_map(&tm.dout, p->dout, L, p->dim, p->dout_d_stride, p->batch, p->dout_batch_stride, box_lines, "dout")) return rc;
    return 0;
}

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
    if (p->flags & ~(VMASR_SCAN_REVERSE | VMASR_SCAN_ACCUMULATE | VMASR_SCAN_ADD | VMASR_SCAN_DBDC_STORE)) return fail("selective_scan: unknown flag bits 0x%x", p->flags);
    if ((p->flags & VMASR_SCAN_DBDC_STORE) && !bwd) return fail("selective_scan_fwd: VMASR_SCAN_DBDC_STORE is a backward flag");
    if ((p->flags & VMASR_SCAN_ACCUMULATE) && (p->flags & VMASR_SCAN_ADD)) return fail("selective_scan: VMASR_SCAN_ACCUMULATE and VMASR_SCAN_ADD exclude each other");
    if (p->zero_bytes != 0 && (!p->zero_ptr || !aligned16(p->zero_ptr) || (p->zero_bytes & 15)))
        return fail("selective_scan: zero_ptr must be non-null and 16-byte aligned, zero_bytes a multiple of 16");
    if (p->dt_rank < 0) return fail("selective_scan: dt_rank must not be negative");
    if (!p->u || (!p->delta && p->dt_rank == 0) || !p->A || !p->B || !p->C) return fail("selective_scan: u, delta, A, B, C must be non-null");
    if (p->dt_rank > 0) {
        if (!p->dt_rows || !p->dt_weight) return fail("selective_scan: dt_rows and dt_weight must be non-null when dt_rank > 0");
        if (bwd && (!p->d_dt_rows || !p->d_dt_weight)) return fail("selective_scan_bwd: d_dt_rows and d_dt_weight must be non-null when dt_rank > 0");
    }
    if (!bwd && (!p->out || !p->x)) return fail("selective_scan_fwd: out and x must be non-null");
    if (bwd) {
        if (!p->dout || !p->du || (!p->ddelta && p->dt_rank == 0) || !p->dA || !p->dB || !p->dC)
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

// Channels per tile.  `peers` = problems launched in the same grid (their tiles fill the machine together).
//   multi-chunk sequences: the resident-tile kernels take at most 4 channels (measured best: 4);
//   single-chunk sequences: a tile walks its channels through a TMA ring, ROWS at a time, so the tile count is free.  Pick
//   the channel count that minimises  waves x (start-up + iterations per row segment), waves = ceil(tiles / resident slots):
//   a launch of 0.6 or 1.15 waves costs as much as one of 1.0 or 2.0 (profiles/r1_summary.md 4a).
static int plan_channels(const vmasr_scan_params *p, int n_chunks, bool bwd, int rows, int peers) {
    const int cpg = p->dim / p->ngroups;
    if (n_chunks > 1) {
        int cap = kMultiChunkTileChannels;
        if (const char *e = tuning_env(bwd ? "VMASR_SCAN_CPT" : "VMASR_SCAN_CPT_FWD")) cap = atoi(e) < 1 ? 1 : atoi(e) > 4 ? 4 : atoi(e);
        if (bwd && p->dt_rank > 0 && cap > 3) cap = 3;  // the fourth stage of the resident tile keeps the dt row (scan_bwd_pipe.cu)
        return cpg < cap ? cpg : cap;
    }
    const long long base_tiles = (long long)p->batch * p->ngroups * (peers < 1 ? 1 : peers);
    const long long slots = (long long)sm_count(p->device) * (bwd ? 2 : 3);  // resident CTAs: __launch_bounds__(256, 2 | 3)
    const char *legacy = tuning_env("VMASR_PLAN_LEGACY");
    if (legacy && atoi(legacy)) {  // round-1 rule: at least two tiles per SM, otherwise as many channels per tile as possible
        long long want = (2ll * sm_count(p->device) + base_tiles - 1) / base_tiles;
        const long long max_ctiles = (cpg + rows - 1) / rows, min_ctiles = (cpg + kMaxTileChannelsHost - 1) / kMaxTileChannelsHost;
        want = want < min_ctiles ? min_ctiles : want > max_ctiles ? max_ctiles : want;
        const int c = (int)((cpg + want - 1) / want);
        return ((c + rows - 1) / rows) * rows;
    }
    constexpr double kStartup = 3.0;  // tile start-up (TMA round trip, B / C, parameters) in units of one ring iteration
    int best = rows;
    double best_cost = 1e30;
    for (int c = rows; c <= kMaxTileChannelsHost; c += rows) {
        const long long ctiles = (cpg + c - 1) / c;
        const long long tiles = base_tiles * ctiles;
        const long long waves = (tiles + slots - 1) / slots;
        const double cost = (double)waves * (kStartup + (double)(c / rows));
        if (cost < best_cost - 1e-9) {
            best_cost = cost;
            best = c;
        }
        if (c >= cpg) break;
    }
    return best;
}

static ScanPlan make_plan(const vmasr_scan_params *p, int n_chunks, bool bwd, int peers, int &chan_per_tile, int &n_ctiles) {
    ScanPlan pl;
    pl.items = 8;
    pl.threads = 256;
    const int L = p->seqlen;
    pl.tpr = L <= 256 ? 32 : L <= 512 ? 64 : L <= 1024 ? 128 : 256;
    pl.rows = pl.threads / pl.tpr;
    const int cpg = p->dim / p->ngroups;
    chan_per_tile = plan_channels(p, n_chunks, bwd, pl.rows, peers);
    chan_per_tile = ((chan_per_tile + pl.rows - 1) / pl.rows) * pl.rows;
    n_ctiles = (cpg + chan_per_tile - 1) / chan_per_tile;
    pl.grid = (int)((long long)p->batch * p->ngroups * n_chunks * n_ctiles);

    const size_t es = dtype_size(p->io_dtype);
    const long long vec_elems = 16 / (long long)es;
    auto mult = [&](long long s) { return s % vec_elems == 0; };
    const bool delta_ok = p->dt_rank > 0 ? (aligned16(p->dt_rows) && mult(p->dt_rows_batch_stride) && mult(p->dt_rows_row_stride))
                                         : (aligned16(p->delta) && mult(p->delta_batch_stride) && mult(p->delta_d_stride));
    pl.vec = (L % vec_elems == 0) && aligned16(p->u) && delta_ok && aligned16(p->B) && aligned16(p->C) &&
             mult(p->u_batch_stride) && mult(p->u_d_stride) &&
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
    a.n_tiles = n_chunks * a.n_rowgroups;
    a.split_from = 0x7fffffff;
    a.softplus = p->delta_softplus;
    a.rev = (p->flags & VMASR_SCAN_REVERSE) ? 1 : 0;
    a.accum = (p->flags & VMASR_SCAN_ACCUMULATE) ? 1 : (p->flags & VMASR_SCAN_ADD) ? 2 : 0;
    a.dbdc_store = (p->flags & VMASR_SCAN_DBDC_STORE) ? 1 : 0;
    a.zero_ptr = static_cast<float4 *>(p->zero_ptr);
    a.zero_n = (long long)(p->zero_bytes / 16);
    a.zero_per_tile = 0;
    const char *pdlx = tuning_env("VMASR_PDL_X");
    a.pdl_mode = pdlx ? atoi(pdlx) : 0;
    const char *nowait = tuning_env("VMASR_DEBUG_NOWAIT");
    a.debug_nowait = nowait ? atoi(nowait) : 0;
#ifdef VMASR_TUNING
    a.timeline = debug_timeline();
#endif
    a.u_bs = p->u_batch_stride; a.u_ds = p->u_d_stride;
    a.delta_bs = p->delta_batch_stride; a.delta_ds = p->delta_d_stride;
    a.A_ds = p->A_d_stride; a.A_ns = p->A_dstate_stride;
    a.B_bs = p->B_batch_stride; a.B_gs = p->B_group_stride; a.B_ns = p->B_dstate_stride;
    a.C_bs = p->C_batch_stride; a.C_gs = p->C_group_stride; a.C_ns = p->C_dstate_stride;
    a.out_bs = p->out_batch_stride; a.out_ds = p->out_d_stride;
    a.dout_bs = p->dout_batch_stride; a.dout_ds = p->dout_d_stride;
    a.du_bs = p->du_batch_stride; a.du_ds = p->du_d_stride;
    a.ddelta_bs = p->ddelta_batch_stride; a.ddelta_ds = p->ddelta_d_stride;
    a.dt_rank = p->dt_rank;
    a.dt_w = p->dt_weight; a.d_dt_rows = p->d_dt_rows; a.d_dt_w = p->d_dt_weight;
    a.dtr_bs = p->dt_rows_batch_stride; a.dtr_rs = p->dt_rows_row_stride; a.dtw_ds = p->dt_weight_d_stride;
    a.dB_bs = p->dB_batch_stride ? p->dB_batch_stride : (long long)p->ngroups * p->dstate * p->seqlen;
    a.dC_bs = p->dC_batch_stride ? p->dC_batch_stride : (long long)p->ngroups * p->dstate * p->seqlen;
    return a;
}

enum ScanVariant { kGeneric = 0, kSingleChunk = 1, kMultiChunk = 2 };

// Multi-chunk launches run in rounds of `slots` resident CTAs (2 per SM backward, 3 forward) and a tile takes 8 - 12 us, so
// a launch of 3.46 rounds costs as much as one of 4.  When the last round is at most half full, its tiles -- the LAST tiles of
// the launch's last problem -- are cut in two (half the channels each): twice as many CTAs, still one round, about half as
// long.  Returns the number of tiles this adds; the kernels decode the halves from ScanArgs::split_from.
static int split_last_round(GroupArgs &ga, int tiles, bool bwd, int device) {
    if (const char *e = tuning_env("VMASR_SCAN_SPLIT"))
        if (!atoi(e)) return 0;
    const int slots = sm_count(device) * (bwd ? 2 : 3);
    if (tiles <= slots) return 0;
    const int rest = tiles % slots;
    ScanArgs &a = ga.a[ga.n - 1];
    const int cpt = a.chan_per_tile;
    if (rest == 0 || 2 * rest > slots || (cpt & 1) || a.chan_per_group % cpt != 0 || rest > a.n_tiles) return 0;
    if (a.dbdc_store) return 0;  // (half tiles would be two writers per dB / dC element)
    a.split_from = a.n_tiles - rest;
    a.n_tiles += rest;
    ga.tile_end[ga.n - 1] += rest;
    return rest;
}

// Everything the launch needs, decided on the host without touching the device (also behind vmasr_scan_plan).
static int decide(const vmasr_scan_params *p, bool bwd, int peers, ScanPlan &pl, ScanArgs &a, int &variant) {
    if (int rc = validate(p, bwd)) return rc;
    const int n_chunks = (p->seqlen + VMASR_SCAN_CHUNK - 1) / VMASR_SCAN_CHUNK;
    int chan_per_tile = 1, n_ctiles = 1;
    pl = make_plan(p, n_chunks, bwd, peers, chan_per_tile, n_ctiles);
    a = make_args(p, n_chunks, chan_per_tile, n_ctiles);
    const size_t es = dtype_size(p->io_dtype);
    const long long vec_elems = 16 / (long long)es;
    auto mult = [&](long long s) { return s % vec_elems == 0; };
    if (!bwd) {
        pl.vec = pl.vec && aligned16(p->out) && mult(p->out_batch_stride) && mult(p->out_d_stride);
    } else {
        const bool ddelta_ok = p->dt_rank > 0 ? aligned16(p->d_dt_rows) : (aligned16(p->ddelta) && mult(p->ddelta_batch_stride) && mult(p->ddelta_d_stride));
        pl.vec = pl.vec && aligned16(p->dout) && aligned16(p->du) && ddelta_ok && aligned16(p->dB) &&
                 aligned16(p->dC) && mult(p->dout_batch_stride) && mult(p->dout_d_stride) && mult(p->du_batch_stride) &&
                 mult(p->du_d_stride) && mult(p->dB_batch_stride) && mult(p->dC_batch_stride) && (p->seqlen % 4 == 0);
    }
    // the fast kernels stage rows through TMA descriptors whose lines are 16 floats (64-byte swizzle, scan.cuh)
    pl.vec = pl.vec && (p->seqlen % kTileLine == 0);
    // fast path: fp32, d_state 1, 16-byte aligned rows -> TMA-staged packed-fp32x2 kernels; more than one chunk: the
    // kernels with the exchange warp (scan_*_pipe.cu).  VMASR_TUNING builds: VMASR_SCAN_FWD / VMASR_SCAN_BWD = generic | tma
    const char *force_env = tuning_env(bwd ? "VMASR_SCAN_BWD" : "VMASR_SCAN_FWD");
    const char force = force_env ? force_env[0] : '\0';
    const bool fast = p->io_dtype == VMASR_F32 && p->dstate == 1 && pl.vec && a.chan_per_tile <= kMaxTileChannelsHost;
    if (!fast || force == 'g') variant = kGeneric;
    else if (n_chunks > 1 && !(force == 't' && p->flags == 0)) variant = kMultiChunk;
    else variant = kSingleChunk;
    if (p->dt_rank > 0 && (variant != kMultiChunk || p->dt_rank != 1))
        return fail("selective_scan: delta on the fly (dt_rank %d) needs dt_rank 1 on the multi-chunk fast path (float32, d_state 1, seqlen > %d and a "
                    "multiple of 16, 16-byte aligned rows and strides)", p->dt_rank, VMASR_SCAN_CHUNK);
    if (p->dB_batch_stride != 0 && variant == kGeneric) return fail("selective_scan: dB / dC batch strides need the fast path");
    if ((p->flags & VMASR_SCAN_DBDC_STORE) && (variant != kMultiChunk || a.n_ctiles != 1))
        return fail("selective_scan_bwd: VMASR_SCAN_DBDC_STORE needs the multi-chunk fast path with one channel tile per B / C group "
                    "(vmasr_scan_plan: variant 2, out[3] == 1); this call has variant %d and %d channel tiles per group", variant, a.n_ctiles);
    if (variant == kGeneric && p->flags != 0)
        return fail("selective_scan: VMASR_SCAN_REVERSE / _ACCUMULATE / _ADD need the fast path (float32, d_state 1, seqlen a multiple of 16, "
                    "16-byte aligned rows and strides)");
    return 0;
}

// kernel family + the template parameters the problems of one launch must share
static long long launch_key(int variant, const ScanArgs &a, const ScanPlan &pl, bool bwd) {
    long long k = variant * 1000 + (a.softplus ? 500 : 0);
    if (variant == kSingleChunk) k += pl.tpr;
    if (variant == kMultiChunk && bwd) k += (a.dt_rank > 0) ? 3 : (a.chan_per_tile <= 3) ? 1 : 2;
    if (variant == kMultiChunk && !bwd && a.dt_rank > 0) k += 3;
    return k;
}

int scan_run_group(int n, const vmasr_scan_params *ps, bool bwd) {
    if (n <= 0) return fail("selective_scan: empty group");
    if (n > kMaxGroup) return fail("selective_scan: at most %d problems per grouped launch (got %d)", kMaxGroup, n);
    if (!ps) return fail("selective_scan: null params");
    ScanPlan pl[kMaxGroup];
    ScanArgs a[kMaxGroup];
    int variant[kMaxGroup];
    long long key[kMaxGroup];
    for (int i = 0; i < n; ++i) {
        if (ps[i].device != ps[0].device || ps[i].stream != ps[0].stream)
            return fail("selective_scan: the problems of a grouped launch must share device and stream");
        if (int rc = decide(&ps[i], bwd, n, pl[i], a[i], variant[i])) return rc;
        key[i] = launch_key(variant[i], a[i], pl[i], bwd);
        for (int j = 0; j < i; ++j)
            if (a[i].ws_header && a[i].ws_header == a[j].ws_header)
                return fail("selective_scan: problems %d and %d of a grouped launch share a carry workspace", j, i);
    }
    DeviceGuard guard(ps[0].device);
    if (!guard.ok) return fail("selective_scan: cannot select CUDA device %d", ps[0].device);
    cudaStream_t stream = static_cast<cudaStream_t>(ps[0].stream);
    // side job (vmasr_scan_params.zero_ptr): the forward's fast kernels clear the region themselves, tile by tile; every
    // other family gets one memset in front of its launch
    for (int i = 0; i < n; ++i) {
        if (a[i].zero_n == 0 || (!bwd && variant[i] != kGeneric)) continue;
        if (int rc = check_cuda(cudaMemsetAsync(a[i].zero_ptr, 0, (size_t)a[i].zero_n * 16, stream), "selective_scan: zero-fill of the side region"))
            return rc;
        a[i].zero_n = 0;
    }
    bool done[kMaxGroup] = {};
    for (int i = 0; i < n; ++i) {
        if (done[i]) continue;
        if (variant[i] == kGeneric) {  // generic kernels (half precision, d_state > 1, unaligned views): one launch each
            done[i] = true;
            const int rc = bwd ? scan_bwd_dispatch(a[i], pl[i], ps[i].io_dtype, stream) : scan_fwd_dispatch(a[i], pl[i], ps[i].io_dtype, stream);
            if (rc) return rc;
            continue;
        }
        GroupArgs ga{};
        int grid = 0;
        for (int j = i; j < n; ++j) {
            if (done[j] || key[j] != key[i]) continue;
            done[j] = true;
            ga.a[ga.n] = a[j];
            // boxes: one 2048-position chunk (multi-chunk kernels) or one row segment of 8 positions per thread (single-chunk kernels)
            const int box_lines = (variant[i] == kMultiChunk ? VMASR_SCAN_CHUNK : pl[j].tpr * 8) / kTileLine;
            if (int rc = make_tile_maps(ga.tm[ga.n], &ps[j], bwd, box_lines)) return rc;
            grid += pl[j].grid;
            ga.tile_end[ga.n] = grid;
            ++ga.n;
        }
        if (variant[i] == kMultiChunk) grid += split_last_round(ga, grid, bwd, ps[0].device);
        for (int j = 0; j < ga.n; ++j) {  // (tile counts are final now: the split above may have added half tiles)
            const long long tiles = ga.tile_end[j] - (j ? ga.tile_end[j - 1] : 0);
            if (ga.a[j].zero_n) ga.a[j].zero_per_tile = (ga.a[j].zero_n + tiles - 1) / tiles;
        }
        for (int j = ga.n; j < kMaxGroup; ++j) ga.tile_end[j] = grid;
        int rc;
        if (variant[i] == kMultiChunk) rc = bwd ? scan_bwd_pipe_dispatch(ga, grid, stream) : scan_fwd_pipe_dispatch(ga, grid, stream);
        else rc = bwd ? scan_bwd_tma_dispatch(ga, pl[i].tpr, grid, stream) : scan_fwd_tma_dispatch(ga, pl[i].tpr, grid, stream);
        if (rc) return rc;
    }
    return 0;
}

}  // namespace vmasr

extern "C" uint64_t vmasr_scan_workspace_bytes(int batch, int dim, int seqlen, int dstate) {
    if (batch <= 0 || dim <= 0 || seqlen <= 0 || dstate <= 0) return 0;
    const uint64_t n_chunks = ((uint64_t)seqlen + VMASR_SCAN_CHUNK - 1) / VMASR_SCAN_CHUNK;
    if (n_chunks <= 1) return 0;
    const uint64_t n_groups = (n_chunks + 15) / 16;  // level-2 entries of the look-back
    const uint64_t cap = (uint64_t)batch * dim * dstate * (n_chunks + n_groups);
    const uint64_t bytes = 64 + 16 * cap;
    return (bytes + 255) / 256 * 256;
}

extern "C" int vmasr_scan_plan(const vmasr_scan_params *p, int backward, int32_t *out) {
    if (!out) return vmasr::fail("vmasr_scan_plan: null output");
    vmasr::ScanPlan pl;
    vmasr::ScanArgs a;
    int variant = 0;
    if (int rc = vmasr::decide(p, backward != 0, 1, pl, a, variant)) return rc;
    out[0] = pl.grid;
    out[1] = variant;
    out[2] = a.chan_per_tile;
    out[3] = a.n_ctiles;
    out[4] = a.n_chunks;
    out[5] = pl.tpr;
    return 0;
}

extern "C" int vmasr_scan_fwd(const vmasr_scan_params *p) { return vmasr::scan_run_group(1, p, false); }
extern "C" int vmasr_scan_bwd(const vmasr_scan_params *p) { return vmasr::scan_run_group(1, p, true); }
extern "C" int vmasr_scan_fwd_grouped(int n, const vmasr_scan_params *p) { return vmasr::scan_run_group(n, p, false); }
extern "C" int vmasr_scan_bwd_grouped(int n, const vmasr_scan_params *p) { return vmasr::scan_run_group(n, p, true); }
