"""Selective scan: the reference's operator surface on top of ``vmasr_scan_fwd`` / ``vmasr_scan_bwd``.

Mirrors, name for name:
  * the native module ``selective_scan_cuda_core`` -- ``fwd(u, delta, A, B, C, D, delta_bias, delta_softplus,
    nrows) -> [out, x]`` and ``bwd(u, delta, A, B, C, D, delta_bias, dout, x, delta_softplus, nrows) ->
    [du, ddelta, dA, dB, dC, dD, ddelta_bias]``
    (kernels/selective_scan/csrc/selective_scan/cus/selective_scan.cpp:157-164, 241-250, 351-354),
  * ``SelectiveScanCore`` (model/vmamba.py:323-356), the autograd.Function ``SS2D.forward_corev2`` calls,
  * ``selective_scan_fn`` (kernels/selective_scan/test_selective_scan.py:241-280).

Argument checks raise ``RuntimeError`` for the same conditions the reference's TORCH_CHECKs do
(selective_scan.cpp:165-215).  No CPU path, no fallback.

Beyond the reference surface: ``fwd_grouped`` / ``bwd_grouped`` launch several independent calls as one grid
(``vmasr_scan_fwd_grouped``), and ``flags`` (``SCAN_REVERSE``, ``SCAN_ACCUMULATE``) select the time-reversed /
accumulating variants the fused SS2D core is built from (``vm_asr_b200.ss2d``).

No zero-fill pass in front of the backward (the reference allocates five zero tensors, selective_scan.cpp:319-327): where one
tile spans a whole B / C group the backward STORES dB / dC (``SCAN_DBDC_STORE``; ``bwd`` asks ``vmasr_scan_plan`` and sets it by
itself), elsewhere the buffer the backward sums them into is cleared by the FORWARD launch as a side job of its tiles
(``bc_accumulator``, ``fwd(..., zero=)``, ``bwd(..., bc=)``); the autograd functions below do both.

Host cost.  A call site (same shapes, strides, dtypes, device) is validated ONCE; its filled-in parameter block is cached
per thread and later calls only refresh pointers, stream and workspace (the reference's pybind entry re-checks every call).
"""
from __future__ import annotations

import contextlib
import ctypes
import threading

import torch

from . import _lib
from ._lib import SCAN_ACCUMULATE, SCAN_ADD, SCAN_DBDC_STORE, SCAN_REVERSE, ScanParams  # noqa: F401


def _check(cond: bool, msg: str):
    if not cond:
        raise RuntimeError(msg)


def _validate(u, delta, A, B, C, D, delta_bias):
    _check(u.dtype in _lib.DTYPE_CODE, "selective_scan: input must be float32, float16 or bfloat16")
    _check(A.dtype == torch.float32, "selective_scan: A must be float32")
    for name, t in (("delta", delta), ("B", B), ("C", C)):
        _check(t.dtype == u.dtype, f"selective_scan: {name} must have the dtype of u")
    for name, t in (("u", u), ("delta", delta), ("A", A), ("B", B), ("C", C)):
        _lib.require_cuda(t, name)
    _check(u.dim() == 3, "selective_scan: u must be (batch, dim, seqlen)")
    batch, dim, seqlen = u.shape
    _check(A.dim() == 2 and A.shape[0] == dim, "selective_scan: A must be (dim, dstate)")
    dstate = A.shape[1]
    _check(B.dim() == 4 and C.dim() == 4, "selective_scan: B and C must be (batch, n_groups, dstate, seqlen)")
    ngroups = B.shape[1]
    _check(dim % ngroups == 0, "dims should be dividable by n_groups")
    _check(dstate <= 256, "selective_scan only supports state dimension <= 256")
    _check(tuple(delta.shape) == (batch, dim, seqlen), "selective_scan: delta must have the shape of u")
    _check(tuple(B.shape) == (batch, ngroups, dstate, seqlen), "selective_scan: B has the wrong shape")
    _check(tuple(C.shape) == (batch, ngroups, dstate, seqlen), "selective_scan: C has the wrong shape")
    for name, t in (("u", u), ("delta", delta), ("B", B), ("C", C)):
        _check(t.stride(-1) == 1 or t.size(-1) == 1, f"selective_scan: {name} must have unit stride along seqlen")
    for name, t in (("D", D), ("delta_bias", delta_bias)):
        if t is not None:
            _check(t.dtype == torch.float32, f"selective_scan: {name} must be float32")
            _lib.require_cuda(t, name)
            _check(tuple(t.shape) == (dim,), f"selective_scan: {name} must be (dim,)")
            _check(t.stride(-1) == 1 or t.size(-1) == 1, f"selective_scan: {name} must be contiguous")
    return batch, dim, seqlen, dstate, ngroups


def _ptr(t):
    return None if t is None else t.data_ptr()


_NO_SWITCH = contextlib.nullcontext()


def _device_of(t):
    """Device guard only when the tensor does not live on the current device (the common case costs nothing)."""
    idx = t.device.index
    return _NO_SWITCH if idx is None or idx == torch.cuda.current_device() else torch.cuda.device(t.device)


# ---- call sites: validated once, parameter block cached per thread -------------------------------------------------
class _Site:
    __slots__ = ("p", "dims", "n_chunks", "ws_bytes", "dev")


_tls = threading.local()
_SITE_CAP = 4096


def _sig(t):
    return None if t is None else (t.shape, t.stride(), t.dtype, t.device)


def _site(u, delta, A, B, C, D, delta_bias, delta_softplus, flags) -> _Site:
    cache = getattr(_tls, "sites", None)
    if cache is None:
        cache = _tls.sites = {}
    key = (_sig(u), _sig(delta), _sig(A), _sig(B), _sig(C), _sig(D), _sig(delta_bias), bool(delta_softplus), flags)
    s = cache.get(key)
    if s is not None:
        return s
    dims = _validate(u, delta, A, B, C, D, delta_bias)
    batch, dim, seqlen, dstate, ngroups = dims
    s = _Site()
    s.dims = dims
    s.n_chunks = (seqlen + _lib.SCAN_CHUNK - 1) // _lib.SCAN_CHUNK
    s.dev = u.device.index if u.device.index is not None else torch.cuda.current_device()
    s.ws_bytes = int(_lib.load_library().vmasr_scan_workspace_bytes(batch, dim, seqlen, dstate)) if s.n_chunks > 1 else 0
    p = s.p = ScanParams()
    p.batch, p.dim, p.seqlen, p.dstate, p.ngroups = dims
    p.u_batch_stride, p.u_d_stride = u.stride(0), u.stride(1)
    p.delta_batch_stride, p.delta_d_stride = delta.stride(0), delta.stride(1)
    p.A_d_stride, p.A_dstate_stride = A.stride(0), A.stride(1)
    p.B_batch_stride, p.B_group_stride, p.B_dstate_stride = B.stride(0), B.stride(1), B.stride(2)
    p.C_batch_stride, p.C_group_stride, p.C_dstate_stride = C.stride(0), C.stride(1), C.stride(2)
    p.io_dtype = _lib.DTYPE_CODE[u.dtype]
    p.delta_softplus = 1 if delta_softplus else 0
    p.device = s.dev
    p.flags = int(flags)
    if len(cache) >= _SITE_CAP:
        cache.clear()
    cache[key] = s
    return s


def _fill_inputs(s: _Site, u, delta, A, B, C, D, delta_bias, workspace=None):
    p = s.p
    p.u, p.delta, p.A, p.B, p.C = u.data_ptr(), delta.data_ptr(), A.data_ptr(), B.data_ptr(), C.data_ptr()
    p.D = None if D is None else D.data_ptr()
    p.delta_bias = None if delta_bias is None else delta_bias.data_ptr()
    p.stream = torch._C._cuda_getCurrentRawStream(s.dev)
    if s.n_chunks > 1:
        ws = workspace if workspace is not None else _lib.scan_workspace(u.device, s.ws_bytes)
        p.workspace, p.workspace_bytes = ws.data_ptr(), ws.numel()
    return p


def _set_zero_region(p, zero):
    """``vmasr_scan_params.zero_ptr / zero_bytes``: a float32 / 16-byte aligned contiguous tensor the launch clears as a side
    job (the accumulated gradients of the backward call to come), or None.  Always set: parameter blocks are reused."""
    if zero is None:
        p.zero_ptr, p.zero_bytes = None, 0
        return
    nbytes = zero.numel() * zero.element_size()
    _check(zero.is_cuda and zero.is_contiguous() and nbytes % 16 == 0 and zero.data_ptr() % 16 == 0,
           "selective_scan: the zero region must be a contiguous CUDA tensor, 16-byte aligned, a multiple of 16 bytes long")
    p.zero_ptr, p.zero_bytes = zero.data_ptr(), nbytes


def _fill_fwd(s: _Site, out, x, zero=None):
    _set_zero_region(s.p, zero)
    batch, dim, seqlen, dstate, _ = s.dims
    _check(out.dtype == _DTYPE_OF[s.p.io_dtype] and tuple(out.shape) == (batch, dim, seqlen) and (out.stride(-1) == 1 or seqlen == 1),
           "selective_scan: out must look like delta")
    _check(x.dtype == torch.float32 and x.is_contiguous() and tuple(x.shape) == (batch, dim, s.n_chunks, 2 * dstate),
           "selective_scan: x has the wrong shape")
    p = s.p
    p.out, p.x = out.data_ptr(), x.data_ptr()
    p.out_batch_stride, p.out_d_stride = out.stride(0), out.stride(1)


def _fill_bwd(s: _Site, A, dout, x, du, ddelta, dA, dB, dC, dD, ddelta_bias):
    batch, dim, seqlen, dstate, ngroups = s.dims
    _check(dout.dtype == _DTYPE_OF[s.p.io_dtype], "selective_scan: dout must have the dtype of u")
    _lib.require_cuda(dout, "dout")
    _check(tuple(dout.shape) == (batch, dim, seqlen), "selective_scan: dout has the wrong shape")
    _check(dout.stride(-1) == 1 or dout.size(-1) == 1, "selective_scan: dout must have unit stride along seqlen")
    if s.n_chunks > 1:
        _check(x is not None, "selective_scan: x is required when seqlen > 2048")
    if x is not None:
        _check(x.dtype == torch.float32 and x.is_cuda and x.is_contiguous(), "selective_scan: x must be contiguous float32 CUDA")
        _check(tuple(x.shape) == (batch, dim, s.n_chunks, 2 * dstate), "selective_scan: x has the wrong shape")
    for name, t in (("dB", dB), ("dC", dC)):
        _check(t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == (batch, ngroups, dstate, seqlen),
               f"selective_scan: {name} must be contiguous float32 (batch, n_groups, dstate, seqlen)")
    _check(dA.dtype == torch.float32 and dA.stride() == A.stride(), "selective_scan: dA must look like A")
    p = s.p
    p.zero_ptr, p.zero_bytes = None, 0
    p.dout, p.x = dout.data_ptr(), _ptr(x)
    p.du, p.ddelta, p.dA, p.dB, p.dC = du.data_ptr(), ddelta.data_ptr(), dA.data_ptr(), dB.data_ptr(), dC.data_ptr()
    p.dD, p.ddelta_bias = _ptr(dD), _ptr(ddelta_bias)
    p.dout_batch_stride, p.dout_d_stride = dout.stride(0), dout.stride(1)
    p.du_batch_stride, p.du_d_stride = du.stride(0), du.stride(1)
    p.ddelta_batch_stride, p.ddelta_d_stride = ddelta.stride(0), ddelta.stride(1)


_DTYPE_OF = {v: k for k, v in _lib.DTYPE_CODE.items()}


def fwd_out(u, delta, A, B, C, D, delta_bias, delta_softplus, out, x, flags=0, workspace=None, zero=None):
    """Launch the forward into caller-provided ``out`` (like delta) and ``x`` (batch, dim, n_chunks, 2*dstate)
    float32.  No allocation when the stream's carry workspace exists already (or ``workspace`` is given).
    ``zero``: a tensor the launch clears as a side job (see ``bc_accumulator``)."""
    _fwd_launch(_site(u, delta, A, B, C, D, delta_bias, delta_softplus, flags), u, delta, A, B, C, D, delta_bias, out, x, workspace, zero)


def _fwd_launch(s, u, delta, A, B, C, D, delta_bias, out, x, workspace=None, zero=None):
    lib = _lib.load_library()
    with _device_of(u):
        _fill_inputs(s, u, delta, A, B, C, D, delta_bias, workspace)
        _fill_fwd(s, out, x, zero)
        _lib.check(lib.vmasr_scan_fwd(ctypes.byref(s.p)))


def fwd(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, nrows=1, zero=None):
    """``selective_scan_cuda_core.fwd``: returns ``[out, x]``.  ``nrows`` is accepted and ignored, like the
    reference's core kernel does (selective_scan.cpp:163)."""
    s = _site(u, delta, A, B, C, D, delta_bias, delta_softplus, 0)
    batch, dim, seqlen, dstate, _ = s.dims
    out = torch.empty_like(delta)
    x = torch.empty((batch, dim, s.n_chunks, 2 * dstate), dtype=torch.float32, device=u.device)
    _fwd_launch(s, u, delta, A, B, C, D, delta_bias, out, x, zero=zero)
    return [out, x]


def bc_accumulator(u, A, B):
    """The buffer the backward of this call will sum dB and dC into, NOT cleared: hand it to the forward as ``zero=`` (the
    forward's kernels clear it while they wait for their first bytes: no memset pass in front of the backward) and to ``bwd``
    as ``bc=``.  None when the backward stores dB / dC anyway (``dbdc_store_candidate``)."""
    if dbdc_store_candidate(u, A, B):
        return None
    n_bc = u.shape[0] * B.shape[1] * A.shape[1] * u.shape[2]
    return torch.empty(2 * ((n_bc + 3) // 4 * 4), dtype=torch.float32, device=u.device)


def bwd_out(u, delta, A, B, C, D, delta_bias, dout, x, delta_softplus, du, ddelta, dA, dB, dC, dD, ddelta_bias, flags=0,
            workspace=None):
    """Launch the backward into caller-provided buffers.  ``dA, dB, dC, dD, ddelta_bias`` are float32 and are
    ACCUMULATED INTO (zero them first); ``dB, dC`` are (batch, n_groups, dstate, seqlen) contiguous."""
    _bwd_launch(_site(u, delta, A, B, C, D, delta_bias, delta_softplus, flags), u, delta, A, B, C, D, delta_bias, dout, x, du, ddelta,
                dA, dB, dC, dD, ddelta_bias, workspace)


def _bwd_launch(s, u, delta, A, B, C, D, delta_bias, dout, x, du, ddelta, dA, dB, dC, dD, ddelta_bias, workspace=None):
    lib = _lib.load_library()
    with _device_of(u):
        _fill_inputs(s, u, delta, A, B, C, D, delta_bias, workspace)
        _fill_bwd(s, A, dout, x, du, ddelta, dA, dB, dC, dD, ddelta_bias)
        _lib.check(lib.vmasr_scan_bwd(ctypes.byref(s.p)))


def _grad_buffers(u, A, D, delta_bias, dims, zero_bc=True, bc=None):
    """The five accumulated gradients (selective_scan.cpp:319-327 allocates five zero tensors): dB and dC share one
    buffer, the three small parameter gradients another, so that a parameter's ``.grad`` never keeps the large buffer
    alive; two memsets per call instead of five -- one when the launch stores dB / dC (``zero_bc`` False: SCAN_DBDC_STORE)."""
    batch, dim, seqlen, dstate, ngroups = dims
    n_bc = batch * ngroups * dstate * seqlen
    n_bc_pad = (n_bc + 3) // 4 * 4
    if bc is not None:  # cleared by the forward launch (bc_accumulator)
        _check(bc.dtype == torch.float32 and bc.numel() == 2 * n_bc_pad and bc.device == u.device, "selective_scan: bc has the wrong size")
        big = bc
    else:
        big = (torch.zeros if zero_bc else torch.empty)(2 * n_bc_pad, dtype=torch.float32, device=u.device)
    dB = big[:n_bc].view(batch, ngroups, dstate, seqlen)
    dC = big[n_bc_pad:n_bc_pad + n_bc].view(batch, ngroups, dstate, seqlen)
    n_a = dim * dstate
    a_dense = A.is_contiguous()
    small = torch.zeros((n_a if a_dense else 0) + 2 * dim, dtype=torch.float32, device=u.device)
    if a_dense:
        dA = small[:n_a].view(dim, dstate)
        off = n_a
    else:
        dA = torch.empty_strided(A.size(), A.stride(), dtype=torch.float32, device=u.device).zero_()
        off = 0
    dD = small[off:off + dim] if D is not None else None
    dbias = small[off + dim:off + 2 * dim] if delta_bias is not None else None
    return dA, dB, dC, dD, dbias


def dbdc_store_candidate(u, A, B) -> bool:
    """Necessary conditions of VMASR_SCAN_DBDC_STORE (include/vmasr_b200.h): float32, d_state 1, more than one chunk, and a
    B / C group narrow enough for ONE multi-chunk tile (4 channels), so that every dB / dC element has a single writer.
    The library has the last word (alignment, strides): ``_store_plan_ok`` asks it before a launch relies on the flag."""
    return (u.dtype == torch.float32 and A.shape[1] == 1 and u.shape[2] > _lib.SCAN_CHUNK and u.shape[2] % 16 == 0
            and u.shape[1] // B.shape[1] <= 4)


def _store_plan_ok(s: _Site) -> bool:
    """``vmasr_scan_plan`` on the FILLED parameter block of a site that carries SCAN_DBDC_STORE: 0 = the launch will store."""
    out = (ctypes.c_int32 * 8)()
    return _lib.load_library().vmasr_scan_plan(ctypes.byref(s.p), 1, out) == 0


def _bwd_prepare(u, delta, A, B, C, D, delta_bias, dout, x, delta_softplus, flags, workspace=None, bc=None):
    """Allocate the gradients of one backward call and fill its parameter block.  Where the plan allows it dB / dC are left
    unset and the launch stores them (no zero-fill pass over the two largest accumulated gradients)."""
    du = torch.empty_like(u)
    ddelta = torch.empty_like(delta)
    if bc is None and dbdc_store_candidate(u, A, B):
        s = _site(u, delta, A, B, C, D, delta_bias, delta_softplus, flags | SCAN_DBDC_STORE)
        dA, dB, dC, dD, dbias = _grad_buffers(u, A, D, delta_bias, s.dims, zero_bc=False)
        with _device_of(u):
            _fill_inputs(s, u, delta, A, B, C, D, delta_bias, workspace)
            _fill_bwd(s, A, dout, x, du, ddelta, dA, dB, dC, dD, dbias)
        if _store_plan_ok(s):
            return s, [du, ddelta, dA, dB, dC, dD, dbias]
        dB.zero_()  # the library would refuse the flag for this call (alignment, strides): accumulate into zeros instead
        dC.zero_()
        s = _site(u, delta, A, B, C, D, delta_bias, delta_softplus, flags)
    else:
        s = _site(u, delta, A, B, C, D, delta_bias, delta_softplus, flags)
        dA, dB, dC, dD, dbias = _grad_buffers(u, A, D, delta_bias, s.dims, bc=bc)
    with _device_of(u):
        _fill_inputs(s, u, delta, A, B, C, D, delta_bias, workspace)
        _fill_bwd(s, A, dout, x, du, ddelta, dA, dB, dC, dD, dbias)
    return s, [du, ddelta, dA, dB, dC, dD, dbias]


def bwd(u, delta, A, B, C, D, delta_bias, dout, x=None, delta_softplus=False, nrows=1, bc=None):
    """``selective_scan_cuda_core.bwd``: returns ``[du, ddelta, dA, dB, dC, dD, ddelta_bias]``.  ``bc``: the buffer of
    ``bc_accumulator`` that the forward call cleared (used up by this call: never pass it twice)."""
    s, res = _bwd_prepare(u, delta, A, B, C, D, delta_bias, dout, x, delta_softplus, 0, bc=bc)
    with _device_of(u):
        _lib.check(_lib.load_library().vmasr_scan_bwd(ctypes.byref(s.p)))
    res[3], res[4] = res[3].to(B.dtype), res[4].to(C.dtype)
    return res


# ---- delta generated inside the kernels (include/vmasr_b200.h, dt_rank > 0) -------------------------------
def _projected_params(u, dt_rows, dt_weight, A, B, C, D, delta_bias, delta_softplus, flags):
    """Parameter block of the projected form: ``delta[b, d, l] = dt_weight[d, 0] * dt_rows[b, group(d), 0, l]`` is formed tile
    by tile inside the multi-chunk fast kernels (vmamba.py:1476-1477 with dt_rank 1); no (batch, dim, seqlen) delta exists."""
    batch, dim, seqlen, dstate, ngroups = _validate(u, u, A, B, C, D, delta_bias)
    _check(dt_rows.dtype == torch.float32 and dt_rows.is_cuda and tuple(dt_rows.shape) == (batch, ngroups, 1, seqlen) and dt_rows.stride(-1) == 1,
           "selective_scan: dt_rows must be float32 CUDA (batch, n_groups, 1, seqlen) with unit stride along seqlen")
    _check(dt_weight.dtype == torch.float32 and dt_weight.is_cuda and tuple(dt_weight.shape) == (dim, 1),
           "selective_scan: dt_weight must be float32 CUDA (dim, 1)")
    p = ScanParams()
    p.batch, p.dim, p.seqlen, p.dstate, p.ngroups = batch, dim, seqlen, dstate, ngroups
    p.u, p.A, p.B, p.C = u.data_ptr(), A.data_ptr(), B.data_ptr(), C.data_ptr()
    p.D, p.delta_bias = _ptr(D), _ptr(delta_bias)
    p.u_batch_stride, p.u_d_stride = u.stride(0), u.stride(1)
    p.A_d_stride, p.A_dstate_stride = A.stride(0), A.stride(1)
    p.B_batch_stride, p.B_group_stride, p.B_dstate_stride = B.stride(0), B.stride(1), B.stride(2)
    p.C_batch_stride, p.C_group_stride, p.C_dstate_stride = C.stride(0), C.stride(1), C.stride(2)
    p.dt_rows, p.dt_weight, p.dt_rank = dt_rows.data_ptr(), dt_weight.data_ptr(), 1
    p.dt_rows_batch_stride, p.dt_rows_row_stride, p.dt_weight_d_stride = dt_rows.stride(0), dt_rows.stride(1), dt_weight.stride(0)
    p.io_dtype = _lib.DTYPE_CODE[u.dtype]
    p.delta_softplus = 1 if delta_softplus else 0
    p.device = u.device.index if u.device.index is not None else torch.cuda.current_device()
    p.flags = int(flags)
    p.stream = torch._C._cuda_getCurrentRawStream(p.device)
    n_chunks = (seqlen + _lib.SCAN_CHUNK - 1) // _lib.SCAN_CHUNK
    if n_chunks > 1:
        ws = _lib.scan_workspace(u.device, int(_lib.load_library().vmasr_scan_workspace_bytes(batch, dim, seqlen, dstate)))
        p.workspace, p.workspace_bytes = ws.data_ptr(), ws.numel()
    return p, (batch, dim, seqlen, dstate, ngroups), n_chunks


def fwd_projected(u, dt_rows, dt_weight, A, B, C, D=None, delta_bias=None, delta_softplus=True, flags=0):
    """Forward with delta generated on the fly.  u (batch, dim, L) float32, dt_rows (batch, n_groups, 1, L), dt_weight (dim, 1);
    the rest as ``fwd``.  Multi-chunk fast path only (float32, d_state 1, L > 2048 and a multiple of 16): anything else raises
    and the caller materialises delta.  Returns ``[out, x]``."""
    p, (batch, dim, seqlen, dstate, _), n_chunks = _projected_params(u, dt_rows, dt_weight, A, B, C, D, delta_bias, delta_softplus, flags)
    out = torch.zeros_like(u) if flags & (_lib.SCAN_ACCUMULATE | 4) else torch.empty_like(u)
    x = torch.empty((batch, dim, n_chunks, 2 * dstate), dtype=torch.float32, device=u.device)
    p.out, p.x = out.data_ptr(), x.data_ptr()
    p.out_batch_stride, p.out_d_stride = out.stride(0), out.stride(1)
    with _device_of(u):
        _lib.check(_lib.load_library().vmasr_scan_fwd(ctypes.byref(p)))
    return [out, x]


def bwd_projected(u, dt_rows, dt_weight, A, B, C, D, delta_bias, dout, x, delta_softplus=True, flags=0):
    """Backward of ``fwd_projected``: returns ``[du, d_dt_rows, d_dt_weight, dA, dB, dC, dD, ddelta_bias]`` -- the gradients
    of the two factors of delta instead of a (batch, dim, L) ddelta."""
    p, dims, n_chunks = _projected_params(u, dt_rows, dt_weight, A, B, C, D, delta_bias, delta_softplus, flags)
    _check(tuple(dout.shape) == tuple(u.shape) and dout.dtype == u.dtype and dout.stride(-1) == 1, "selective_scan: dout must look like u")
    du = torch.empty_like(u)
    d_rows = torch.zeros_like(dt_rows, memory_format=torch.contiguous_format)
    _check(d_rows.stride() == dt_rows.stride(), "selective_scan: dt_rows must be contiguous for the backward")
    d_w = torch.zeros_like(dt_weight)
    dA, dB, dC, dD, dbias = _grad_buffers(u, A, D, delta_bias, dims)
    p.dout, p.x = dout.data_ptr(), x.data_ptr()
    p.dout_batch_stride, p.dout_d_stride = dout.stride(0), dout.stride(1)
    p.du, p.du_batch_stride, p.du_d_stride = du.data_ptr(), du.stride(0), du.stride(1)
    p.d_dt_rows, p.d_dt_weight = d_rows.data_ptr(), d_w.data_ptr()
    p.dA, p.dB, p.dC, p.dD, p.ddelta_bias = dA.data_ptr(), dB.data_ptr(), dC.data_ptr(), _ptr(dD), _ptr(dbias)
    with _device_of(u):
        _lib.check(_lib.load_library().vmasr_scan_bwd(ctypes.byref(p)))
    return [du, d_rows, d_w, dA, dB, dC, dD, dbias]


# ---- grouped launches -------------------------------------------------------------------------------------------------
def _group_workspaces(sites, device):
    """One buffer from the stream's workspace, cut into 256-byte aligned regions (one per problem)."""
    total = sum(s.ws_bytes for s in sites)
    if total == 0:
        return [None] * len(sites)
    ws = _lib.scan_workspace(device, total, layout=tuple(s.ws_bytes for s in sites))
    out, off = [], 0
    for s in sites:
        out.append(ws[off:off + s.ws_bytes] if s.ws_bytes else None)
        off += s.ws_bytes
    return out


def fwd_grouped(calls, outs=None, zero=None):
    """``calls``: list of ``(u, delta, A, B, C, D, delta_bias, delta_softplus[, flags])`` tuples, at most
    ``SCAN_MAX_GROUP`` of them, all on one device.  One launch per kernel family (normally one).  Returns a list of
    ``[out, x]``; ``outs`` may give pre-allocated ``(out, x)`` pairs, ``zero`` one tensor (or None) per call that the call's
    tiles clear as a side job (``bc_accumulator``)."""
    _check(0 < len(calls) <= _lib.SCAN_MAX_GROUP, f"fwd_grouped: 1..{_lib.SCAN_MAX_GROUP} calls")
    sites, args = [], []
    for c in calls:
        u, delta, A, B, C, D, bias, sp = c[:8]
        flags = c[8] if len(c) > 8 else 0
        # a site's parameter block is reused between calls: equal call sites inside one group need their own copy
        s = _site(u, delta, A, B, C, D, bias, sp, flags)
        sites.append(s)
        args.append((u, delta, A, B, C, D, bias))
    device = calls[0][0].device
    wss = _group_workspaces(sites, device)
    results = []
    n = len(calls)
    arr = (ScanParams * n)()
    size = ctypes.sizeof(ScanParams)
    with _device_of(calls[0][0]):
        for i, (s, a) in enumerate(zip(sites, args)):
            batch, dim, seqlen, dstate, _ = s.dims
            if outs is not None:
                out, x = outs[i]
            else:
                out = torch.empty_like(a[1])
                x = torch.empty((batch, dim, s.n_chunks, 2 * dstate), dtype=torch.float32, device=device)
            _fill_inputs(s, *a, workspace=wss[i])
            _fill_fwd(s, out, x, None if zero is None else zero[i])
            ctypes.memmove(ctypes.addressof(arr[i]), ctypes.addressof(s.p), size)
            results.append([out, x])
        _lib.check(_lib.load_library().vmasr_scan_fwd_grouped(n, arr))
    return results


def bwd_grouped(calls, outs=None):
    """``calls``: list of ``(u, delta, A, B, C, D, delta_bias, dout, x, delta_softplus[, flags])``; returns a list of
    ``[du, ddelta, dA, dB, dC, dD, ddelta_bias]`` (dB / dC float32).  ``outs`` may give pre-allocated 7-tuples (the five
    accumulated ones zero-filled by the caller)."""
    _check(0 < len(calls) <= _lib.SCAN_MAX_GROUP, f"bwd_grouped: 1..{_lib.SCAN_MAX_GROUP} calls")
    device = calls[0][0].device
    sites = []
    for c in calls:
        u, delta, A, B, C, D, bias, dout, x, sp = c[:10]
        flags = c[10] if len(c) > 10 else 0
        sites.append(_site(u, delta, A, B, C, D, bias, sp, flags))
    wss = _group_workspaces(sites, device)
    n = len(calls)
    arr = (ScanParams * n)()
    size = ctypes.sizeof(ScanParams)
    results = []
    with _device_of(calls[0][0]):
        for i, (s, c) in enumerate(zip(sites, calls)):
            u, delta, A, B, C, D, bias, dout, x, sp = c[:10]
            if outs is not None:
                res = list(outs[i])
                du, ddelta, dA, dB, dC, dD, dbias = res
                _fill_inputs(s, u, delta, A, B, C, D, bias, workspace=wss[i])
                _fill_bwd(s, A, dout, x, du, ddelta, dA, dB, dC, dD, dbias)
            else:  # (dB / dC stored instead of zero-filled and accumulated where the plan allows it)
                s, res = _bwd_prepare(u, delta, A, B, C, D, bias, dout, x, sp, s.p.flags, workspace=wss[i])
            ctypes.memmove(ctypes.addressof(arr[i]), ctypes.addressof(s.p), size)
            results.append(res)
        _lib.check(_lib.load_library().vmasr_scan_bwd_grouped(n, arr))
    return results


class PreparedCalls:
    """One (grouped) launch with its parameter blocks filled ONCE -- for call sites that launch the same tensors again and
    again (a captured step, a benchmark loop, an inference server's fixed buffers).  The tensors must stay alive and keep
    their storage; a call only refreshes the stream handle and the carry-workspace pointers and launches: no validation, no
    allocation, no per-tensor work.  ``prepare_fwd`` / ``prepare_bwd`` build it from the arguments of ``fwd_grouped`` /
    ``bwd_grouped`` with pre-allocated outputs."""

    def __init__(self, kind, calls, outs, zero=None):
        _check(outs is not None and len(outs) == len(calls), "prepare: pre-allocated outputs are required")
        _check(0 < len(calls) <= _lib.SCAN_MAX_GROUP, f"prepare: 1..{_lib.SCAN_MAX_GROUP} calls")
        lib = _lib.load_library()
        self._fn = lib.vmasr_scan_fwd_grouped if kind == "fwd" else lib.vmasr_scan_bwd_grouped
        self._keep = (calls, outs, zero)
        self.device = calls[0][0].device
        self._dev = self.device.index if self.device.index is not None else torch.cuda.current_device()
        n = self.n = len(calls)
        self.arr = (ScanParams * n)()
        size = ctypes.sizeof(ScanParams)
        self._ws_bytes, self._total = [], 0
        for i, (c, o) in enumerate(zip(calls, outs)):
            if kind == "fwd":
                u, delta, A, B, C, D, bias, sp = c[:8]
                flags = c[8] if len(c) > 8 else 0
                s = _site(u, delta, A, B, C, D, bias, sp, flags)
                _fill_inputs(s, u, delta, A, B, C, D, bias, workspace=torch.empty(0) if s.ws_bytes else None)
                _fill_fwd(s, o[0], o[1], None if zero is None else zero[i])
            else:
                u, delta, A, B, C, D, bias, dout, x, sp = c[:10]
                flags = c[10] if len(c) > 10 else 0
                s = _site(u, delta, A, B, C, D, bias, sp, flags)
                _fill_inputs(s, u, delta, A, B, C, D, bias, workspace=torch.empty(0) if s.ws_bytes else None)
                _fill_bwd(s, A, dout, x, *o)
            ctypes.memmove(ctypes.addressof(self.arr[i]), ctypes.addressof(s.p), size)
            self._ws_bytes.append(s.ws_bytes)
            self._total += s.ws_bytes
        self._layout = tuple(self._ws_bytes)

    def __call__(self):
        stream = torch._C._cuda_getCurrentRawStream(self._dev)
        arr = self.arr
        if self._total:
            base = _lib.scan_workspace(self.device, self._total, layout=self._layout).data_ptr()
            for i in range(self.n):
                p = arr[i]
                p.stream = stream
                if self._ws_bytes[i]:
                    p.workspace, p.workspace_bytes = base, self._ws_bytes[i]
                    base += self._ws_bytes[i]
        else:
            for i in range(self.n):
                arr[i].stream = stream
        rc = self._fn(self.n, arr)
        if rc:
            _lib.check(rc)


def prepare_fwd(calls, outs, zero=None) -> PreparedCalls:
    return PreparedCalls("fwd", calls, outs, zero)


def prepare_bwd(calls, outs) -> PreparedCalls:
    return PreparedCalls("bwd", calls, outs)


class SelectiveScanCore(torch.autograd.Function):
    """Drop-in for ``model.vmamba.SelectiveScanCore`` (vmamba.py:323-356); same argument list, the trailing
    ``nrows, backnrows, oflex`` are accepted and ignored exactly as there."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, nrows=1, backnrows=1, oflex=True):
        ctx.delta_softplus = delta_softplus
        _site(u, delta, A, B, C, D, delta_bias, delta_softplus, 0)  # (validates: reference-style errors before anything else)
        ctx.bc = bc_accumulator(u, A, B) if any(ctx.needs_input_grad) else None
        out, x = fwd(u, delta, A, B, C, D, delta_bias, delta_softplus, 1, zero=ctx.bc)
        ctx.save_for_backward(u, delta, A, B, C, D, delta_bias, x)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dout, *args):
        u, delta, A, B, C, D, delta_bias, x = ctx.saved_tensors
        if dout.stride(-1) != 1:
            dout = dout.contiguous()
        bc, ctx.bc = ctx.bc, None  # (a second backward over a retained graph gets fresh zeros)
        du, ddelta, dA, dB, dC, dD, ddelta_bias = bwd(u, delta, A, B, C, D, delta_bias, dout, x, ctx.delta_softplus, 1, bc=bc)
        return (du, ddelta, dA, dB, dC, dD, ddelta_bias, None, None, None, None)


class _SelectiveScanFn(torch.autograd.Function):
    """The functional form with ``return_last_state`` (test_selective_scan.py:24-135)."""

    @staticmethod
    def forward(ctx, u, delta, A, B, C, D, delta_bias, delta_softplus, return_last_state):
        if u.stride(-1) != 1:
            u = u.contiguous()
        if delta.stride(-1) != 1:
            delta = delta.contiguous()
        if B.stride(-1) != 1:
            B = B.contiguous()
        if C.stride(-1) != 1:
            C = C.contiguous()
        ctx.squeeze_B = B.dim() == 3
        ctx.squeeze_C = C.dim() == 3
        if ctx.squeeze_B:
            B = B.unsqueeze(1)
        if ctx.squeeze_C:
            C = C.unsqueeze(1)
        ctx.d_dtype = None if D is None else D.dtype
        ctx.bias_dtype = None if delta_bias is None else delta_bias.dtype
        if D is not None:
            D = D.float().contiguous()
        if delta_bias is not None:
            delta_bias = delta_bias.float().contiguous()
        ctx.delta_softplus = delta_softplus
        _site(u, delta, A, B, C, D, delta_bias, delta_softplus, 0)  # (validates: reference-style errors before anything else)
        ctx.bc = bc_accumulator(u, A, B) if any(ctx.needs_input_grad) else None
        out, x = fwd(u, delta, A, B, C, D, delta_bias, delta_softplus, 1, zero=ctx.bc)
        ctx.save_for_backward(u, delta, A, B, C, D, delta_bias, x)
        if return_last_state:
            last_state = x[:, :, -1, 1::2]
            ctx.mark_non_differentiable(last_state)
            return out, last_state
        return out

    @staticmethod
    def backward(ctx, dout, *args):
        u, delta, A, B, C, D, delta_bias, x = ctx.saved_tensors
        if dout.stride(-1) != 1:
            dout = dout.contiguous()
        bc, ctx.bc = ctx.bc, None
        du, ddelta, dA, dB, dC, dD, dbias = bwd(u, delta, A, B, C, D, delta_bias, dout, x, ctx.delta_softplus, 1, bc=bc)
        if ctx.squeeze_B:
            dB = dB.squeeze(1)
        if ctx.squeeze_C:
            dC = dC.squeeze(1)
        if dD is not None and ctx.d_dtype is not None:
            dD = dD.to(ctx.d_dtype)
        if dbias is not None and ctx.bias_dtype is not None:
            dbias = dbias.to(ctx.bias_dtype)
        return du, ddelta, dA, dB, dC, dD, dbias, None, None


def selective_scan_fn(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, return_last_state=False):
    """``selective_scan_fn(u, delta, A, B, C, D, delta_bias, delta_softplus)`` with its autograd backward.
    ``B``/``C`` may be (batch, dstate, seqlen) (one group) or (batch, n_groups, dstate, seqlen)."""
    return _SelectiveScanFn.apply(u, delta, A, B, C, D, delta_bias, delta_softplus, return_last_state)
