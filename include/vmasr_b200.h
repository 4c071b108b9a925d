/* vmasr_b200 -- C ABI of the B200 (sm_100a) implementation of VM-ASR's SS2D + STFT hot path.
 *
 * Everything here is plain C: raw device pointers, sizes, element strides, a CUDA stream handle and an
 * int return code (0 = ok, non-zero = error; text via vmasr_last_error()).  No torch types cross this
 * boundary and nothing is thrown across it.  The caller owns every buffer (inputs, outputs, the
 * chunk-state tensor and the carry workspace); the library allocates nothing and keeps no pointer after
 * a call returns.  Entry points are re-entrant and never synchronise the device: the reference calls its
 * forward from the Python main thread and its backward from an autograd worker thread
 * (model/vmamba.py:325-356), each on torch's current stream.
 *
 * The fast scan kernels are launched with the programmatic-stream-serialization attribute (their CTAs may be scheduled while
 * the previous kernel of the stream drains); each of them executes griddepcontrol.wait before its first access to global
 * memory, so stream order holds behind any predecessor.  The multi-chunk backward runs a grid of persistent CTAs sized to
 * what the device keeps resident at once (cudaOccupancyMaxActiveBlocksPerMultiprocessor x SM count).
 *
 * Each entry point names the reference interface it replaces (paths relative to the VM-ASR tree).
 */
#ifndef VMASR_B200_H
#define VMASR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VMASR_ABI_VERSION 5

#if defined(__GNUC__)
#define VMASR_API __attribute__((visibility("default")))
#else
#define VMASR_API
#endif

/* element type of u / delta / B / C / out / dout / du / ddelta and of cross-scan maps
 * (selective_scan.cpp:167-172 accepts exactly these three) */
enum vmasr_dtype { VMASR_F32 = 0, VMASR_F16 = 1, VMASR_BF16 = 2 };

/* The scan is cut into chunks of this many positions; the chunk-state tensor `x` has
 * ceil(seqlen / VMASR_SCAN_CHUNK) entries per (batch, channel) -- same figure as selective_scan.cpp:217. */
#define VMASR_SCAN_CHUNK 2048

VMASR_API int vmasr_abi_version(void);
/* Message of the last failing call on this host thread ("" if none). */
VMASR_API const char *vmasr_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Selective scan.  Replaces selective_scan_cuda_core.fwd / .bwd
 * (kernels/selective_scan/csrc/selective_scan/cus/selective_scan.cpp:157-239, 241-349, pybind :351-354).
 *
 *   u, delta, out, dout, du, ddelta : (batch, dim, seqlen), unit stride along seqlen, element type io_dtype
 *   A                               : (dim, dstate) float
 *   B, C                            : (batch, ngroups, dstate, seqlen), unit stride along seqlen, io_dtype
 *   D, delta_bias                   : (dim,) float or NULL
 *   x                               : (batch, dim, n_chunks, 2*dstate) float, contiguous: per chunk and state
 *                                     (cumulative decay, state at chunk end) -- written by fwd, read by bwd
 *   dA (dim,dstate), dD, ddelta_bias (dim,) : float, ACCUMULATED INTO (caller zero-fills, like
 *                                     selective_scan.cpp:321-327)
 *   dB, dC                          : (batch, ngroups, dstate, seqlen) float contiguous, accumulated into
 *                                     (caller zero-fills -- or lets the FORWARD call clear them through zero_ptr / zero_bytes
 *                                     below, or sets VMASR_SCAN_DBDC_STORE where the plan allows it; the reference casts
 *                                     them to io_dtype afterwards, :347)
 *   workspace                       : carry-exchange area for the cross-chunk look-back.  At least
 *                                     vmasr_scan_workspace_bytes() bytes, ZERO-FILLED ONCE when allocated and
 *                                     then reused call after call (kernels recycle it themselves, so a
 *                                     captured CUDA graph can replay).  One workspace per stream: two scans
 *                                     running concurrently must not share one.  May be NULL when
 *                                     seqlen <= VMASR_SCAN_CHUNK.
 * Strides are in elements.
 * ---------------------------------------------------------------------------------------------- */
typedef struct vmasr_scan_params {
    /* forward inputs */
    const void *u, *delta;
    const float *A;
    const void *B, *C;
    const float *D;          /* nullable */
    const float *delta_bias; /* nullable */
    /* forward outputs */
    void *out;
    float *x;
    /* backward input */
    const void *dout;
    /* backward outputs */
    void *du, *ddelta;
    float *dA, *dB, *dC;
    float *dD;          /* nullable iff D is NULL */
    float *ddelta_bias; /* nullable iff delta_bias is NULL */
    /* carry workspace */
    void *workspace;
    uint64_t workspace_bytes;
    /* sizes */
    int32_t batch, dim, seqlen, dstate, ngroups;
    /* strides (elements) */
    int64_t u_batch_stride, u_d_stride;
    int64_t delta_batch_stride, delta_d_stride;
    int64_t A_d_stride, A_dstate_stride;
    int64_t B_batch_stride, B_group_stride, B_dstate_stride;
    int64_t C_batch_stride, C_group_stride, C_dstate_stride;
    int64_t out_batch_stride, out_d_stride;
    int64_t dout_batch_stride, dout_d_stride;
    int64_t du_batch_stride, du_d_stride;
    int64_t ddelta_batch_stride, ddelta_d_stride;
    /* options */
    int32_t io_dtype;       /* enum vmasr_dtype */
    int32_t delta_softplus; /* softplus(delta + delta_bias), identity above 20 (fwd_kernel.cuh:117) */
    int32_t device;         /* CUDA device ordinal the pointers live on */
    int32_t flags;          /* VMASR_SCAN_* bits below; 0 = the reference's semantics */
    void *stream; /* cudaStream_t */
    /* ---- delta generated on the fly from the rank-R projection (model/vmamba.py:1476-1477) ----
     * With dt_rank > 0 the kernels never see a (batch, dim, seqlen) `delta`: they are handed what dt_projs_weight is applied
     * to, and form   delta[b, d, l] = sum_r dt_weight[d, r] * dt_rows[b, group(d), r, l]   tile by tile in shared memory
     * (`delta` / `ddelta` are ignored and may be NULL).  The backward returns the gradients of the two factors instead of
     * ddelta:  d_dt_rows[b, g, r, l] += sum_{d in g} dt_weight[d, r] * ddelta[b, d, l]  and
     * d_dt_weight[d, r] += sum_{b, l} ddelta[b, d, l] * dt_rows[b, g, r, l]  (both ACCUMULATED INTO; caller zero-fills).
     * Supported: the multi-chunk fast path (float32, d_state 1, seqlen > VMASR_SCAN_CHUNK and a multiple of 16, aligned
     * rows) with dt_rank 1 -- the three largest maps of every config (R = 1 wherever L >= 16384 at DIMS 16);
     * anything else fails and the caller materialises delta (vmasr_ss2d_* does that choice itself).
     *   dt_rows, d_dt_rows : (batch, ngroups, dt_rank, seqlen), unit stride along seqlen, row (g * dt_rank + r) at
     *                        dt_rows_row_stride, batch at dt_rows_batch_stride
     *   dt_weight, d_dt_weight : (dim, dt_rank), unit stride along r, channel stride dt_weight_d_stride
     *   dB_batch_stride, dC_batch_stride : 0 = contiguous (ngroups * dstate * seqlen); otherwise the batch stride of dB / dC
     *                        (the fused core lets them land next to d_dt_rows in one d x_dbl tensor) */
    const float *dt_rows;
    const float *dt_weight;
    float *d_dt_rows;
    float *d_dt_weight;
    int64_t dt_rows_batch_stride, dt_rows_row_stride;
    int64_t dt_weight_d_stride;
    int64_t dB_batch_stride, dC_batch_stride;
    int32_t dt_rank;
    int32_t reserved0;
    /* ---- side job: a region the launch ZERO-FILLS while it runs (ABI 5) ----
     * zero_ptr (16-byte aligned) / zero_bytes (a multiple of 16; 0 = none): cleared by the time the launch completes, in
     * stream order after whatever precedes the launch.  Meant for the accumulated gradients of the backward call that
     * belongs to this forward call (dB / dC, and dA / dD / ddelta_bias if they sit in the same buffer): the reference
     * zero-fills them with five memsets in front of its backward kernel (selective_scan.cpp:319-327); here the forward's
     * fast kernels spread the stores over their tiles, where they cost nothing measurable (the tiles wait for their first
     * bytes then), and every other kernel family falls back to one cudaMemsetAsync in front of its launch -- the caller
     * never needs to know which.  The region must not overlap anything the launch reads or writes. */
    void *zero_ptr;
    uint64_t zero_bytes;
} vmasr_scan_params;

/* flags (fast path only: float32, d_state 1, seqlen a multiple of 16, 16-byte aligned rows and strides; otherwise the call
 * fails).  They exist for the fused SS2D core, whose directions 2 and 3 are directions 0 and 1 run back to front
 * (model/vmamba.py:33, 54-55), and whose four outputs are summed (vmamba.py:55-60).
 *   VMASR_SCAN_REVERSE    : time runs against memory order: position l of the recurrence is element seqlen-1-l of u, delta,
 *                           B, C, out (dout, du, ddelta, dB, dC).  Every tensor keeps its memory layout; x (chunk states) is
 *                           indexed in time order.
 *   VMASR_SCAN_ACCUMULATE : `out` (forward) and `du` (backward) are ADDED INTO with 128-bit reductions (red.global.add)
 *                           instead of stored: any number of calls may add into one buffer concurrently.
 *   VMASR_SCAN_ADD        : the same sum as a plain load / add / store: cheaper, but the call must be the ONLY writer of the
 *                           buffer while it runs (stream order after whatever wrote it before).  Exclusive with ACCUMULATE.
 *   VMASR_SCAN_DBDC_STORE : backward only: dB and dC are STORED, not accumulated into -- the caller need not zero-fill them
 *                           (the reference zero-fills and accumulates with atomics, selective_scan.cpp:321-323,
 *                           bwd_kernel.cuh:216-221).  Possible when one tile covers all channels of a B / C group, i.e. when
 *                           vmasr_scan_plan(p, 1, out) reports variant 2 (multi-chunk fast path) and out[3] (channel tiles per
 *                           group) == 1: every dB / dC element then has exactly one writer.  The call FAILS otherwise
 *                           (nothing is launched), so a caller can never end up with sums on top of unset memory. */
#define VMASR_SCAN_REVERSE 1
#define VMASR_SCAN_ACCUMULATE 2
#define VMASR_SCAN_ADD 4
#define VMASR_SCAN_DBDC_STORE 8
/* most problems one grouped launch takes */
#define VMASR_SCAN_MAX_GROUP 8

VMASR_API uint64_t vmasr_scan_workspace_bytes(int batch, int dim, int seqlen, int dstate);
/* Host-side planning only (no device access, no launch): validates `p` exactly as vmasr_scan_fwd / vmasr_scan_bwd would
 * (same return codes and messages, mirroring the TORCH_CHECKs of selective_scan.cpp:165-215, 262-317) and reports how the
 * call would be run: out[0] = grid size (tiles), out[1] = kernel family (0 generic, 1 single-chunk fast path, 2 multi-chunk
 * fast path, 3 persistent ring), out[2] = channels per tile, out[3] = channel tiles per B/C group, out[4] = chunks per
 * sequence, out[5] = threads per row segment.  Pointers in `p` are only tested for null and 16-byte alignment. */
VMASR_API int vmasr_scan_plan(const vmasr_scan_params *p, int backward, int32_t *out);
VMASR_API int vmasr_scan_fwd(const vmasr_scan_params *p);
VMASR_API int vmasr_scan_bwd(const vmasr_scan_params *p);
/* n <= VMASR_SCAN_MAX_GROUP independent calls as ONE grid (same device and stream, one carry workspace EACH).  The two
 * streams of the generator issue the same-shape SS2D call independently (model/model.py:1167-1176, 1124-1127): launched
 * together they fill the machine where one call alone is a 0.6 - 1.7 wave launch.  Problems that need different kernel
 * families are launched family by family; results are identical to n separate calls. */
VMASR_API int vmasr_scan_fwd_grouped(int n, const vmasr_scan_params *p);
VMASR_API int vmasr_scan_bwd_grouped(int n, const vmasr_scan_params *p);

/* ------------------------------------------------------------------------------------------------
 * 4-direction cross scan / merge.  Replace CrossScan / CrossMerge (model/vmamba.py:27-73) and the Triton
 * kernels triton_cross_scan / triton_cross_merge (model/csm_triton.py:7-154).
 *   cross_scan : x (B, C, H, W) contiguous -> xs (B, 4, C, H*W) contiguous; also CrossMerge's backward.
 *   cross_merge: ys (B, 4, C, H*W) contiguous -> y (B, C, H*W) contiguous, association
 *                (ys0 + flip ys2) + transpose_back(ys1 + flip ys3) as vmamba.py:55-60; also CrossScan's
 *                backward.
 * Bit-exact with the PyTorch versions for all three dtypes.
 * ---------------------------------------------------------------------------------------------- */
VMASR_API int vmasr_cross_scan(const void *x, void *xs, int B, int C, int H, int W, int dtype, int device, void *stream);
VMASR_API int vmasr_cross_merge(const void *ys, void *y, int B, int C, int H, int W, int dtype, int device, void *stream);
/* One-by-one variants: replace CrossScanTriton1b1 and its kernels triton_cross_scan_1b1 / triton_cross_merge_1b1
 * (model/csm_triton.py:157-308, 369-395; reached only from SS2D.forwardxv, which no shipped config selects).
 *   cross_scan_1b1 : x (B, 4, C, H, W) -> y (B, 4, C, H*W), direction k of y is direction k's walk over map k of x
 *   cross_merge_1b1: the inverse permutation, y (B, 4, C, H*W) -> x (B, 4, C, H, W) */
VMASR_API int vmasr_cross_scan_1b1(const void *x, void *y, int B, int C, int H, int W, int dtype, int device, void *stream);
VMASR_API int vmasr_cross_merge_1b1(const void *y, void *x, int B, int C, int H, int W, int dtype, int device, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Fused SS2D core: CrossScan -> selective scan -> CrossMerge of SS2D.forward_corev2 (model/vmamba.py:1472-1497; the Triton
 * kernels model/csm_triton.py:7-154 and the scan extension in between) WITHOUT the four-fold copies `xs` (B, 4, C, L) and
 * `ys` (B, 4, C, L).  The four directions are four scan problems of ONE grid (see vmasr_scan_fwd_grouped):
 *   k = 0: the map read row-major;            k = 2: the same memory, time reversed (VMASR_SCAN_REVERSE);
 *   k = 1: the TRANSPOSED map read row-major; k = 3: the same memory, time reversed.
 * Everything positional of direction k (delta_k, B_k, C_k and their gradients) is laid out in the MEMORY order of its
 * pair -- row-major (l = h*W + w) for k = 0, 2, column-major (l = w*H + h) for k = 1, 3 -- and NOT flipped for k = 2, 3:
 * that is what the projections give when they are applied to the map and to its transpose instead of to `xs`
 * (einsum(x, x_proj_weight[k]) commutes with the permutation of positions).  The outputs of a pair meet in one plane
 * (directions 0, 1 store, directions 2, 3 run in a second grid and add to what is there: one rounding of y_k + y_{k+2}), and
 *   y = (y0 + y2) + transpose(y1 + y3)     -- the association of vmamba.py:55-60.
 * float32, d_state 1, H % 4 == 0 and W % 4 == 0 (every map of the configs); otherwise the call fails and the caller chains
 * vmasr_cross_scan / vmasr_scan_* / vmasr_cross_merge.
 *
 *   x        (B, C, H, W)            xT   (B, C, W, H) = vmasr_map_transpose(x)
 *   delta[k] (B, C, L) with strides  B[k], C[k] (B, L) with a batch stride (unit stride along L)
 *   A, D, delta_bias (4*C,) direction-major (As / Ds / dt_projs_bias.view(-1) of vmamba.py:1481-1485)
 *   y        (B, C, H, W) out        planes  2 x (B, C, L) scratch (forward) / the two gradient planes (backward)
 *   states   (4, B, C, n_chunks, 2) chunk states per direction: written by fwd, read by bwd
 *   backward: dy (B, C, H, W) in, dyT (B, C, W, H) scratch; ddelta[k] like delta[k]; dB, dC (4, B, L), dA, dD, ddelta_bias
 *   (4*C,) ACCUMULATED INTO (caller zero-fills); planes[0] = d x through directions 0, 2 (row-major), planes[1] = d xT
 *   through directions 1, 3; dx (optional) = planes[0] + transpose(planes[1]).
 *   workspace: vmasr_ss2d_workspace_bytes(), zero-filled once, one per concurrently running call (as for the scan).
 * vmasr_ss2d_core_fwd / _bwd take n = 1 or 2 parameter blocks: the generator's two streams run the same-shape core
 * independently (model/model.py:1124-1127) and share one grid.
 * ---------------------------------------------------------------------------------------------- */
typedef struct vmasr_ss2d_params {
    const float *x, *xT;
    const float *delta[4];
    int64_t delta_batch_stride[4], delta_d_stride[4];
    const float *B[4], *C[4];
    int64_t B_batch_stride[4], C_batch_stride[4];
    const float *A, *D, *delta_bias; /* D, delta_bias nullable */
    float *y;
    float *planes;
    float *states;
    /* backward */
    const float *dy;
    float *dyT;
    float *dx; /* nullable */
    float *ddelta[4];
    int64_t ddelta_batch_stride[4], ddelta_d_stride[4];
    float *dA, *dB, *dC, *dD, *ddelta_bias;
    void *workspace;
    uint64_t workspace_bytes;
    int32_t batch, channels, H, W;
    int32_t delta_softplus, device;
    void *stream;
    /* ---- projected form (dt_rank > 0): delta generated inside the scan kernels, see vmasr_scan_params ----
     * x_dbl[k] = the (R + 2) rows einsum(x_k, x_proj_weight[k]) of direction k in the memory order of its pair
     * (vmamba.py:1473-1476): rows 0..R-1 feed dt_projs_weight[k] (C, R), row R is B_k, row R + 1 is C_k.  delta[], B[], C[]
     * above are then ignored, as are ddelta[], dB, dC: the backward ACCUMULATES d x_dbl[k] (same layout; caller zero-fills)
     * and d dt_weight (4, C, R).  Needs dt_rank == 1 and H * W > VMASR_SCAN_CHUNK. */
    const float *x_dbl[4];
    int64_t x_dbl_batch_stride[4], x_dbl_row_stride[4];
    const float *dt_weight; /* (4, C, R) contiguous */
    float *d_x_dbl[4];
    float *d_dt_weight;
    int32_t dt_rank;
    int32_t flags; /* VMASR_SS2D_* bits below */
} vmasr_ss2d_params;
/* backward: dyT already holds transpose(dy) (the caller produced both, e.g. behind vmasr_outnorm_gate_bwd): skip the
 * internal transpose */
#define VMASR_SS2D_DYT_GIVEN 1

VMASR_API uint64_t vmasr_ss2d_workspace_bytes(int batch, int channels, int H, int W);
VMASR_API int vmasr_ss2d_core_fwd(int n, const vmasr_ss2d_params *p);
VMASR_API int vmasr_ss2d_core_bwd(int n, const vmasr_ss2d_params *p);
/* x (planes, H, W) -> xT (planes, W, H); y = p_rm + transpose(p_cm).  float32, H % 4 == 0, W % 4 == 0, 16-byte aligned. */
VMASR_API int vmasr_map_transpose(const float *x, float *xT, int64_t planes, int H, int W, int device, void *stream);
VMASR_API int vmasr_map_merge2(const float *p_rm, const float *p_cm, float *y, int64_t planes, int H, int W, int device, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Tail of the SS2D block fused into the merge of the core's two planes.  Replaces, in one kernel,
 *   y = CrossMerge's outer addition (model/vmamba.py:57-60)  ->  y.transpose(1, 2).contiguous()  ->  out_norm = nn.LayerNorm(C)
 *   (vmamba.py:1527-1529, out_norm_shape "v0")  ->  y.to(x.dtype) (:1531)  ->  y * act(z) (forwardv2, :1536-1550).
 *   p_rm, p_cm : (B, C, H*W) float32, the planes vmasr_ss2d_core_fwd leaves (pass y = NULL there): y0 + y2 with w fastest,
 *                y1 + y3 with h fastest.  p_cm = NULL: p_rm is an already merged map (e.g. CrossMerge's output).
 *   gamma, beta: (C) float32 LayerNorm weight / bias (NULL = 1 / 0);  eps as nn.LayerNorm (1e-5)
 *   z          : (B, H*W, C) of io_dtype, the gate BEFORE its activation when z_silu = 1 (SiLU applied inside, rounded to
 *                io_dtype as the reference's act(z) tensor is), after it when z_silu = 0; NULL = no gate
 *   out        : (B, H*W, C) of io_dtype
 *   y, stats   : forward OUTPUTS kept for the backward, (B, C, H*W) float32 merged map and (B, H*W, 2) float32 {mean, rstd};
 *                either may be NULL at inference.  The backward reads them.
 *   backward   : dout (B, H*W, C) io_dtype -> dy (B, C, H*W) float32 (row-major: what vmasr_ss2d_core_bwd takes as dy),
 *                dz (B, H*W, C) io_dtype (required when z is given), and dgb_partial (patches, 2, C) float32: per-patch sums
 *                of d gamma and d beta, every entry written (no zero-fill needed); the caller sums over the first axis
 *                (patches = vmasr_outnorm_patches()); NULL = not wanted.
 * H % 4 == 0, W % 8 == 0, float32 tensors 16-byte aligned.
 * ---------------------------------------------------------------------------------------------- */
typedef struct vmasr_outnorm_params {
    const float *p_rm, *p_cm, *gamma, *beta;
    const void *z;
    void *out;
    float *y, *stats;
    const void *dout;
    float *dy;
    void *dz;
    float *dgb_partial;
    float eps;
    int32_t batch, channels, H, W;
    int32_t io_dtype, z_silu, device;
    void *stream;
} vmasr_outnorm_params;
VMASR_API int64_t vmasr_outnorm_patches(int batch, int channels, int H, int W);
VMASR_API int vmasr_outnorm_gate_fwd(const vmasr_outnorm_params *p);
VMASR_API int vmasr_outnorm_gate_bwd(const vmasr_outnorm_params *p);

/* ------------------------------------------------------------------------------------------------
 * Head of the SS2D block fused into the core's load.  Replaces, in one kernel, SS2D.forwardv2's
 *   x.permute(0, 3, 1, 2).contiguous() -> conv2d (depthwise 3x3, padding 1, bias; model/vmamba.py:860-868) -> SiLU
 *   (vmamba.py:1541-1546) and the float32 cast / transpose the fused core needs in front of it.
 *   xin    : (B, H, W, C) of io_dtype, channel-last, xin_pos_stride elements between positions (0 = C; in_proj's (B, H, W, 2C)
 *            output is read in place with stride 2C)
 *   weight : (C, 1, 3, 3) float32 contiguous;  bias: (C) float32 or NULL
 *   x, xT  : (B, C, H, W) and (B, C, W, H) float32 -- what vmasr_ss2d_core_fwd takes (xT may be NULL).  The convolution and
 *            the activation are rounded to io_dtype where the reference holds tensors of that dtype.
 *   backward: dx (B, C, H, W), dxT (B, C, W, H) float32 (the `planes` vmasr_ss2d_core_bwd leaves; dxT may be NULL) ->
 *            dxin (B, H, W, C) of io_dtype, contiguous, and dwb_partial (patches, C, 10) float32: per-patch sums of d weight (9)
 *            and d bias, every entry written; the caller sums over patches (vmasr_dwconv_patches()); NULL = not wanted.
 * H % 4 == 0, W % 8 == 0.
 * ---------------------------------------------------------------------------------------------- */
typedef struct vmasr_dwconv_params {
    const void *xin;
    const float *weight, *bias;
    float *x, *xT;
    const float *dx, *dxT;
    void *dxin;
    float *dwb_partial;
    int64_t xin_pos_stride;
    int32_t batch, channels, H, W;
    int32_t io_dtype, device;
    void *stream;
    /* ---- x_proj in the same pass (optional; SURVEY.md 8f-1: x_dbl = einsum(xs, x_proj_weight) [+ x_proj_bias], vmamba.py:1473-1475).
     * x_proj_weight (4, x_proj_rows, C) float32 contiguous, x_proj_rows = dt_rank + 2 d_state; x_proj_bias (4, x_proj_rows) or NULL.
     * x_dbl_rm / x_dbl_cm: (B, 2, x_proj_rows, H*W) float32, the rows of directions (0, 2) in row-major and of (1, 3) in column-major
     * position order -- exactly vmasr_ss2d_params.x_dbl[] -- ACCUMULATED INTO when the channels span several CTAs
     * (vmasr_dwconv_channel_blocks() > 1): the caller zero-fills them then.
     * backward: d_x_dbl_rm / d_x_dbl_cm in, their contribution is added to the map gradient; d_x_proj_weight_partial
     * (patches, 4 * x_proj_rows, C) float32, every entry written, the caller sums over patches (d x_proj_bias is the plain sum of
     * d x_dbl over batch and positions: left to the caller). */
    const float *x_proj_weight, *x_proj_bias;
    float *x_dbl_rm, *x_dbl_cm;
    const float *d_x_dbl_rm, *d_x_dbl_cm;
    float *d_x_proj_weight_partial;
    int32_t x_proj_rows, reserved0;
} vmasr_dwconv_params;
VMASR_API int64_t vmasr_dwconv_patches(int batch, int channels, int H, int W);
/* CTAs the channels of one patch are spread over: 1 = x_dbl_rm / x_dbl_cm are plainly stored (no zero-fill needed) */
VMASR_API int vmasr_dwconv_channel_blocks(int batch, int channels, int H, int W);
VMASR_API int vmasr_dwconv_silu_fwd(const vmasr_dwconv_params *p);
VMASR_API int vmasr_dwconv_silu_bwd(const vmasr_dwconv_params *p);

/* ------------------------------------------------------------------------------------------------
 * Magnitude/phase STFT and inverse.  Replace wav2spectro / spectro2wav (utils/stft.py:22-68, 71-115),
 * i.e. torch.stft / torch.istft(normalized=True, center=True, reflect pad, periodic Hann of win_length
 * zero-padded to n_fft, onesided) fused with log2(|X|+1e-8)/angle and exp2/polar.  float32 only.
 *   wave  : (B, T) contiguous
 *   mag, phase : (B, n_fft/2+1, n_frames) contiguous, n_frames = 1 + T/hop
 *   istft output length = hop*(n_frames-1)
 *   stft_bwd  : d mag, d phase -> d wave (wav2spectro is differentiable in the reference: torch.stft, abs, log2, angle)
 *   istft_bwd : d wave -> d mag, d phase (the generator loss flows through spectro2wav, model/model.py:1223)
 *   scratch   : B * vmasr_stft_scratch_floats(n_frames, n_fft, hop) floats, caller-owned, any content: the padded
 *               overlap-add accumulator of the synthesis kernels (istft_fwd, stft_bwd, stft_mag_bwd)
 * Linear-magnitude STFT of the multi-resolution loss and the LSD metric (model/loss.py:17-45, model/metric.py:5-12):
 *   stft_mag_fwd : mag = sqrt(max(|X|^2, clamp_min)), X un-normalised unless `normalized`;  stft_mag_bwd: d mag -> d wave.
 * n_fft must be a power of two in [64, 2048]; win_length <= n_fft.
 * ---------------------------------------------------------------------------------------------- */
VMASR_API uint64_t vmasr_stft_scratch_floats(int n_frames, int n_fft, int hop);
VMASR_API int vmasr_stft_fwd(const float *wave, float *mag, float *phase, int B, int T, int n_fft, int hop,
                   int win_length, int device, void *stream);
VMASR_API int vmasr_stft_bwd(const float *wave, const float *dmag, const float *dphase, float *dwave, float *scratch, int B, int T,
                   int n_fft, int hop, int win_length, int device, void *stream);
VMASR_API int vmasr_stft_mag_fwd(const float *wave, float *mag, int B, int T, int n_fft, int hop, int win_length, int normalized,
                   float clamp_min, int device, void *stream);
VMASR_API int vmasr_stft_mag_bwd(const float *wave, const float *dmag, float *dwave, float *scratch, int B, int T, int n_fft, int hop,
                   int win_length, int normalized, float clamp_min, int device, void *stream);
VMASR_API int vmasr_istft_fwd(const float *mag, const float *phase, float *wave, float *scratch, int B, int n_frames, int n_fft,
                    int hop, int win_length, int device, void *stream);
VMASR_API int vmasr_istft_bwd(const float *mag, const float *phase, const float *dwave, float *dmag, float *dphase,
                    int B, int n_frames, int n_fft, int hop, int win_length, int device, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* VMASR_B200_H */
