// Fused SS2D core (include/vmasr_b200.h): the four directions of SS2D.forward_corev2 (model/vmamba.py:1472-1497) as four
// problems of one grouped scan launch, reading the map (and its transpose) in place and adding their outputs into two
// planes; no (B, 4, C, L) copy is ever made.  This file only assembles the scan problems; the kernels are the scan kernels.
#include <cstdlib>

#include "scan.cuh"

namespace vmasr {

int scan_run_group(int n, const vmasr_scan_params *ps, bool bwd);
int map_transpose_launch(const float *x, float *xT, long long planes, int H, int W, cudaStream_t stream);
int map_merge2_launch(const float *p_rm, const float *p_cm, float *y, long long planes, int H, int W, cudaStream_t stream);

static int ss2d_check(const vmasr_ss2d_params *p, bool bwd, const char *who) {
    if (!p) return fail("%s: null params", who);
    if (p->batch <= 0 || p->channels <= 0 || p->H <= 0 || p->W <= 0) return fail("%s: sizes must be positive", who);
    if (p->H % 4 || p->W % 4) return fail("%s: H and W must be multiples of 4 (got %d x %d)", who, p->H, p->W);
    if (!p->x || !p->xT || !p->A || !p->planes || !p->states) return fail("%s: x, xT, A, planes, states must be non-null", who);
    const bool proj = p->dt_rank != 0;  // projected form: delta generated inside the scan kernels from x_dbl and dt_weight
    if (proj) {
        if (p->dt_rank != 1 || (long long)p->H * p->W <= VMASR_SCAN_CHUNK)
            return fail("%s: the projected form needs dt_rank 1 and H * W > %d (got rank %d, %d x %d)", who, VMASR_SCAN_CHUNK, p->dt_rank, p->H, p->W);
        if (!p->dt_weight) return fail("%s: dt_weight must be non-null in the projected form", who);
        for (int k = 0; k < 4; ++k)
            if (!p->x_dbl[k]) return fail("%s: x_dbl of direction %d missing", who, k);
    } else {
        for (int k = 0; k < 4; ++k)
            if (!p->delta[k] || !p->B[k] || !p->C[k]) return fail("%s: delta / B / C of direction %d missing", who, k);
    }
    // (forward: y may be NULL -- the caller merges the two planes itself, e.g. with vmasr_outnorm_gate_fwd)
    if (bwd) {
        if (!p->dy || !p->dyT || !p->dA) return fail("%s: dy, dyT, dA must be non-null", who);
        if (proj) {
            if (!p->d_dt_weight) return fail("%s: d_dt_weight must be non-null in the projected form", who);
            for (int k = 0; k < 4; ++k)
                if (!p->d_x_dbl[k]) return fail("%s: d_x_dbl of direction %d missing", who, k);
        } else {
            if (!p->dB || !p->dC) return fail("%s: dB, dC must be non-null", who);
            for (int k = 0; k < 4; ++k)
                if (!p->ddelta[k]) return fail("%s: ddelta of direction %d missing", who, k);
        }
        if ((p->D != nullptr) != (p->dD != nullptr)) return fail("%s: dD must be given exactly when D is", who);
        if ((p->delta_bias != nullptr) != (p->ddelta_bias != nullptr)) return fail("%s: ddelta_bias must be given exactly when delta_bias is", who);
    }
    return 0;
}

static uint64_t dir_ws_bytes(const vmasr_ss2d_params *p) {
    return vmasr_scan_workspace_bytes(p->batch, p->channels, p->H * p->W, 1);
}

// How the two directions of a pair meet in their plane:
//   two passes (default): directions 0, 1 of every map in one grid with plain stores, then directions 2, 3 in a second grid
//                that adds to what the first left (VMASR_SCAN_ADD: load / add / store, one writer per element) -- no
//                zero-fill, no atomics;
//   one pass:    all four directions in one grid, each adding into a zero-filled plane with red.global.add.
// Either way a plane holds exactly y_k + y_{k+2}, one rounding, so the results are bit-identical.
static bool ss2d_one_pass() {
    const char *e = tuning_env("VMASR_SS2D_ONE_PASS");
    return e && atoi(e) != 0;
}

// scan problem of direction k of map problem p
static vmasr_scan_params direction(const vmasr_ss2d_params *p, int k, bool bwd, bool one_pass) {
    const long long C = p->channels, L = (long long)p->H * p->W, Bz = p->batch;
    const long long n_chunks = (L + VMASR_SCAN_CHUNK - 1) / VMASR_SCAN_CHUNK;
    vmasr_scan_params s{};
    s.u = (k & 1) ? p->xT : p->x;
    s.u_batch_stride = C * L;
    s.u_d_stride = L;
    const bool proj = p->dt_rank != 0;
    const long long R = p->dt_rank, xd_bs = p->x_dbl_batch_stride[k], xd_rs = p->x_dbl_row_stride[k];
    if (proj) {  // rows 0 .. R-1 of x_dbl[k] feed dt_weight, row R is B, row R + 1 is C (vmamba.py:1476)
        s.dt_rank = p->dt_rank;
        s.dt_rows = p->x_dbl[k];
        s.dt_rows_batch_stride = xd_bs;
        s.dt_rows_row_stride = xd_rs;
        s.dt_weight = p->dt_weight + k * C * R;
        s.dt_weight_d_stride = R;
    } else {
        s.delta = p->delta[k];
        s.delta_batch_stride = p->delta_batch_stride[k];
        s.delta_d_stride = p->delta_d_stride[k];
    }
    s.A = p->A + k * C;
    s.A_d_stride = 1;
    s.A_dstate_stride = 1;
    s.B = proj ? p->x_dbl[k] + R * xd_rs : p->B[k];
    s.B_batch_stride = proj ? xd_bs : p->B_batch_stride[k];
    s.B_group_stride = L;
    s.B_dstate_stride = L;
    s.C = proj ? p->x_dbl[k] + (R + 1) * xd_rs : p->C[k];
    s.C_batch_stride = proj ? xd_bs : p->C_batch_stride[k];
    s.C_group_stride = L;
    s.C_dstate_stride = L;
    s.D = p->D ? p->D + k * C : nullptr;
    s.delta_bias = p->delta_bias ? p->delta_bias + k * C : nullptr;
    float *plane = p->planes + (k & 1) * Bz * C * L;
    s.out = plane;
    s.out_batch_stride = C * L;
    s.out_d_stride = L;
    s.x = p->states + k * Bz * C * n_chunks * 2;
    if (bwd) {
        s.dout = (k & 1) ? p->dyT : p->dy;
        s.dout_batch_stride = C * L;
        s.dout_d_stride = L;
        s.du = plane;
        s.du_batch_stride = C * L;
        s.du_d_stride = L;
        s.dA = p->dA + k * C;
        if (proj) {  // d x_dbl[k] has the layout of x_dbl[k]: its rows receive d(dt rows), dB, dC
            s.d_dt_rows = p->d_x_dbl[k];
            s.d_dt_weight = p->d_dt_weight + k * C * R;
            s.dB = p->d_x_dbl[k] + R * xd_rs;
            s.dC = p->d_x_dbl[k] + (R + 1) * xd_rs;
            s.dB_batch_stride = xd_bs;
            s.dC_batch_stride = xd_bs;
        } else {
            s.ddelta = p->ddelta[k];
            s.ddelta_batch_stride = p->ddelta_batch_stride[k];
            s.ddelta_d_stride = p->ddelta_d_stride[k];
            s.dB = p->dB + k * Bz * L;
            s.dC = p->dC + k * Bz * L;
        }
        s.dD = p->dD ? p->dD + k * C : nullptr;
        s.ddelta_bias = p->ddelta_bias ? p->ddelta_bias + k * C : nullptr;
    }
    const uint64_t wsb = dir_ws_bytes(p);
    if (wsb) {
        s.workspace = static_cast<char *>(p->workspace) + k * wsb;
        s.workspace_bytes = wsb;
    }
    s.batch = p->batch;
    s.dim = p->channels;
    s.seqlen = (int)L;
    s.dstate = 1;
    s.ngroups = 1;
    s.io_dtype = VMASR_F32;
    s.delta_softplus = p->delta_softplus;
    s.device = p->device;
    s.flags = one_pass ? (VMASR_SCAN_ACCUMULATE | (k >= 2 ? VMASR_SCAN_REVERSE : 0)) : (k >= 2 ? (VMASR_SCAN_REVERSE | VMASR_SCAN_ADD) : 0);
    s.stream = p->stream;
    return s;
}

static int ss2d_run(int n, const vmasr_ss2d_params *ps, bool bwd) {
    const char *who = bwd ? "ss2d_core_bwd" : "ss2d_core_fwd";
    if (n < 1 || n > kMaxGroup / 4) return fail("%s: 1 or %d maps per call (got %d)", who, kMaxGroup / 4, n);
    if (!ps) return fail("%s: null params", who);
    const bool one_pass = ss2d_one_pass();
    vmasr_scan_params sp[kMaxGroup];  // [first pass: directions 0, 1 of every map | second pass: directions 2, 3]
    for (int i = 0; i < n; ++i) {
        const vmasr_ss2d_params *p = &ps[i];
        if (int rc = ss2d_check(p, bwd, who)) return rc;
        if (p->device != ps[0].device || p->stream != ps[0].stream) return fail("%s: the maps of one call must share device and stream", who);
        const uint64_t need = 4 * dir_ws_bytes(p);
        if (need && (!p->workspace || p->workspace_bytes < need))
            return fail("%s: workspace too small (%llu < %llu bytes)", who, (unsigned long long)p->workspace_bytes, (unsigned long long)need);
        for (int k = 0; k < 4; ++k) sp[(k >> 1) * 2 * n + 2 * i + (k & 1)] = direction(p, k, bwd, one_pass);
    }
    DeviceGuard guard(ps[0].device);
    if (!guard.ok) return fail("%s: cannot select CUDA device %d", who, ps[0].device);
    cudaStream_t stream = static_cast<cudaStream_t>(ps[0].stream);
    for (int i = 0; i < n; ++i) {
        const vmasr_ss2d_params *p = &ps[i];
        const long long planes = (long long)p->batch * p->channels, L = (long long)p->H * p->W;
        (void)L;
        if (bwd && !(p->flags & VMASR_SS2D_DYT_GIVEN))
            if (int rc = map_transpose_launch(p->dy, p->dyT, planes, p->H, p->W, stream)) return rc;
        if (one_pass)
            if (int rc = check_cuda(cudaMemsetAsync(p->planes, 0, sizeof(float) * 2 * planes * L, stream), "ss2d planes memset")) return rc;
    }
    if (one_pass) {
        if (int rc = scan_run_group(4 * n, sp, bwd)) return rc;
    } else {
        if (int rc = scan_run_group(2 * n, sp, bwd)) return rc;
        if (int rc = scan_run_group(2 * n, sp + 2 * n, bwd)) return rc;
    }
    for (int i = 0; i < n; ++i) {
        const vmasr_ss2d_params *p = &ps[i];
        const long long planes = (long long)p->batch * p->channels, L = (long long)p->H * p->W;
        float *dst = bwd ? p->dx : p->y;
        if (dst)
            if (int rc = map_merge2_launch(p->planes, p->planes + planes * L, dst, planes, p->H, p->W, stream)) return rc;
    }
    return 0;
}

static int map_entry(const float *a, const float *b, float *out, long long planes, int H, int W, int device, void *stream, bool merge) {
    const char *who = merge ? "map_merge2" : "map_transpose";
    if (!a || !out || (merge && !b)) return fail("%s: null tensor", who);
    if (planes <= 0 || H <= 0 || W <= 0) return fail("%s: sizes must be positive", who);
    if (H % 4 || W % 4) return fail("%s: H and W must be multiples of 4 (got %d x %d)", who, H, W);
    if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15u)
        return fail("%s: tensors must be 16-byte aligned", who);
    DeviceGuard guard(device);
    if (!guard.ok) return fail("%s: cannot select CUDA device %d", who, device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return merge ? map_merge2_launch(a, b, out, planes, H, W, s) : map_transpose_launch(a, out, planes, H, W, s);
}

}  // namespace vmasr

extern "C" uint64_t vmasr_ss2d_workspace_bytes(int batch, int channels, int H, int W) {
    if (batch <= 0 || channels <= 0 || H <= 0 || W <= 0) return 0;
    return 4 * vmasr_scan_workspace_bytes(batch, channels, H * W, 1);
}
extern "C" int vmasr_ss2d_core_fwd(int n, const vmasr_ss2d_params *p) { return vmasr::ss2d_run(n, p, false); }
extern "C" int vmasr_ss2d_core_bwd(int n, const vmasr_ss2d_params *p) { return vmasr::ss2d_run(n, p, true); }
extern "C" int vmasr_map_transpose(const float *x, float *xT, int64_t planes, int H, int W, int device, void *stream) {
    return vmasr::map_entry(x, nullptr, xT, planes, H, W, device, stream, false);
}
extern "C" int vmasr_map_merge2(const float *p_rm, const float *p_cm, float *y, int64_t planes, int H, int W, int device, void *stream) {
    return vmasr::map_entry(p_rm, p_cm, y, planes, H, W, device, stream, true);
}
