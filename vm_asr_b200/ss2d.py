"""The SS2D core of a VSS block, ``SS2D.forward_corev2`` (model/vmamba.py:1472-1497):

    xs = CrossScan(x) -> x_dbl = einsum(xs, x_proj_weight) -> dts = einsum(dts, dt_projs_weight)
       -> ys = SelectiveScanCore(xs, dts, -exp(A_logs), Bs, Cs, Ds, dt_projs_bias, delta_softplus=True) -> y = CrossMerge(ys)

``ss2d_core`` runs it FUSED (``vmasr_ss2d_core_fwd`` / ``_bwd``): the four-fold copies ``xs`` and ``ys`` are never made.

* The projections commute with the permutation of positions, so they are applied to the map itself (directions 0, 2) and to
  its transpose (directions 1, 3) instead of to ``xs``: ``delta_k``, ``B_k``, ``C_k`` come out in the MEMORY order of
  their pair (row-major / column-major), not flipped, at a quarter of the einsum work on the input side.
* The scan kernels read ``x`` / ``x^T`` in place; directions 2, 3 run time-reversed over the same memory; the outputs of a
  pair are added into one zero-filled plane; ``y = (y0 + y2) + transpose(y1 + y3)`` (the association of vmamba.py:55-60).
* The two small einsums stay on PyTorch/cuBLAS (they are not on the named path), as do autograd's sums.

``ss2d_core_chain`` is the unfused chain of the library's three operators (what round 1 shipped); it remains the path for
maps the fused kernels do not take (H or W not a multiple of 4) and the reference point of the fused path's tests.
``ss2d_core_pair`` runs the cores of the generator's two streams (same shapes, independent: model/model.py:1124-1127) as
one grid.
"""
from __future__ import annotations

import ctypes  # noqa: F401

import torch

from . import _lib
from .cross import CrossMerge, CrossScan
from .scan import SelectiveScanCore

from ._lib import SS2DParams


def _dev(t):
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def map_transpose(x: torch.Tensor) -> torch.Tensor:
    """(..., H, W) float32 contiguous -> (..., W, H), one tiled pass (``vmasr_map_transpose``)."""
    lib = _lib.load_library()
    _lib.require_cuda(x, "x")
    if x.dtype != torch.float32:
        raise RuntimeError("map_transpose: float32 only")
    x = x.contiguous()
    *lead, H, W = x.shape
    out = x.new_empty(*lead, W, H)
    planes = x.numel() // (H * W)
    with torch.cuda.device(x.device):
        _lib.check(lib.vmasr_map_transpose(x.data_ptr(), out.data_ptr(), planes, H, W, _dev(x), _lib.current_stream_ptr(x.device)))
    return out


def map_merge2(p_rm: torch.Tensor, p_cm: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """p_rm (..., H*W) + transpose(p_cm (..., W*H)) -> (..., H*W): the outer addition of CrossMerge (vmamba.py:57-60)."""
    lib = _lib.load_library()
    p_rm, p_cm = p_rm.contiguous(), p_cm.contiguous()
    out = torch.empty_like(p_rm)
    planes = p_rm.numel() // (H * W)
    with torch.cuda.device(p_rm.device):
        _lib.check(lib.vmasr_map_merge2(p_rm.data_ptr(), p_cm.data_ptr(), out.data_ptr(), planes, H, W, _dev(p_rm),
                                        _lib.current_stream_ptr(p_rm.device)))
    return out


class MapTranspose(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return map_transpose(x)

    @staticmethod
    def backward(ctx, g):
        return map_transpose(g)


_PER_MAP = 11  # tensors per map of _SS2DScan
# first argument of the two scan Functions: bit 0 = delta_softplus (so plain True / False keep their meaning), bit 1 = return the
# core's two PLANES (2, B, C, L) -- y0 + y2 with w fastest, y1 + y3 with h fastest -- instead of the merged map: the caller
# merges them itself (MergeNormGate below fuses that merge with the block's LayerNorm / gate tail), and the backward then
# receives (dy, dy^T) stacked the same way
_SOFTPLUS, _PLANES = 1, 2


def _fill(p: SS2DParams, x, xT, dts_rm, dts_cm, Bs_rm, Bs_cm, Cs_rm, Cs_cm, As, Ds, bias, softplus):
    Bsz, C, H, W = x.shape
    p.x, p.xT = x.data_ptr(), xT.data_ptr()
    for k in range(4):
        pair, j = (dts_rm, Bs_rm, Cs_rm) if k % 2 == 0 else (dts_cm, Bs_cm, Cs_cm), k // 2
        d, b, c = pair[0][:, j], pair[1][:, j], pair[2][:, j]
        p.delta[k], p.delta_batch_stride[k], p.delta_d_stride[k] = d.data_ptr(), d.stride(0), d.stride(1)
        p.B[k], p.B_batch_stride[k] = b.data_ptr(), b.stride(0)
        p.C[k], p.C_batch_stride[k] = c.data_ptr(), c.stride(0)
    p.A, p.D, p.delta_bias = As.data_ptr(), Ds.data_ptr(), bias.data_ptr()
    p.batch, p.channels, p.H, p.W = Bsz, C, H, W
    p.delta_softplus = 1 if softplus else 0
    p.device = _dev(x)
    p.stream = _lib.current_stream_ptr(x.device)


def _check_map(x, xT, dts_rm, dts_cm, Bs_rm, Bs_cm, Cs_rm, Cs_cm, As, Ds, bias):
    Bsz, C, H, W = x.shape
    L = H * W
    ok = x.is_contiguous() and xT.is_contiguous() and tuple(xT.shape) == (Bsz, C, W, H)
    for t in (dts_rm, dts_cm):
        ok = ok and tuple(t.shape) == (Bsz, 2, C, L) and t.stride(-1) == 1
    for t in (Bs_rm, Bs_cm, Cs_rm, Cs_cm):
        ok = ok and tuple(t.shape) == (Bsz, 2, 1, L) and t.stride(-1) == 1
    ok = ok and tuple(As.shape) == (4 * C, 1) and As.is_contiguous() and Ds.numel() == 4 * C and bias.numel() == 4 * C
    for t in (x, xT, dts_rm, dts_cm, Bs_rm, Bs_cm, Cs_rm, Cs_cm, As, Ds, bias):
        ok = ok and t.is_cuda and t.dtype == torch.float32
    if not ok:
        raise RuntimeError("ss2d_core: fused core expects float32 CUDA tensors x (B,C,H,W), xT (B,C,W,H), dts (B,2,C,L), "
                           "Bs / Cs (B,2,1,L) with unit stride along L, As (4C,1), Ds / delta_bias (4C)")


class _SS2DScan(torch.autograd.Function):
    """CrossScan -> selective scan -> CrossMerge of ``n_maps`` maps in one grid.  Per map, in order:
    x (B,C,H,W), xT (B,C,W,H), dts_rm, dts_cm (B,2,C,L), Bs_rm, Bs_cm, Cs_rm, Cs_cm (B,2,1,L), As (4C,1), Ds (4C), delta_bias (4C);
    index 0 / 1 of a ``_rm`` tensor is direction 0 / 2, of a ``_cm`` tensor direction 1 / 3.  Returns y (B,C,L) per map."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, softplus, n_maps, *tensors):
        lib = _lib.load_library()
        want_planes, softplus = bool(int(softplus) & _PLANES), bool(int(softplus) & _SOFTPLUS)
        arr = (SS2DParams * n_maps)()
        keep, ys = [], []
        dev = tensors[0].device
        wbytes = [int(lib.vmasr_ss2d_workspace_bytes(*tensors[m * _PER_MAP].shape)) for m in range(n_maps)]
        ws = _lib.scan_workspace(dev, sum(wbytes), layout=("ss2d",) + tuple(wbytes)) if sum(wbytes) else None
        off = 0
        for m in range(n_maps):
            t = tensors[m * _PER_MAP:(m + 1) * _PER_MAP]
            _check_map(*t)
            x = t[0]
            Bsz, C, H, W = x.shape
            L = H * W
            n_chunks = (L + _lib.SCAN_CHUNK - 1) // _lib.SCAN_CHUNK
            y = None if want_planes else torch.empty((Bsz, C, L), dtype=torch.float32, device=dev)
            planes = torch.empty((2, Bsz, C, L), dtype=torch.float32, device=dev)
            states = torch.empty((4, Bsz, C, n_chunks, 2), dtype=torch.float32, device=dev)
            p = arr[m]
            _fill(p, *t, softplus)
            p.y, p.planes, p.states = (None if want_planes else y.data_ptr()), planes.data_ptr(), states.data_ptr()
            if wbytes[m]:
                p.workspace, p.workspace_bytes = ws.data_ptr() + off, wbytes[m]
                off += wbytes[m]
            keep.append((planes, states))
            ys.append(planes if want_planes else y)
        with torch.cuda.device(dev):
            _lib.check(lib.vmasr_ss2d_core_fwd(n_maps, arr))
        ctx.softplus, ctx.n_maps, ctx.want_planes = softplus, n_maps, want_planes
        ctx.save_for_backward(*tensors, *[k[1] for k in keep])
        return tuple(ys) if n_maps > 1 else ys[0]

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, *dys):
        lib = _lib.load_library()
        n_maps = ctx.n_maps
        saved = ctx.saved_tensors
        tensors, states_all = saved[:n_maps * _PER_MAP], saved[n_maps * _PER_MAP:]
        arr = (SS2DParams * n_maps)()
        dev = tensors[0].device
        wbytes = [int(lib.vmasr_ss2d_workspace_bytes(*tensors[m * _PER_MAP].shape)) for m in range(n_maps)]
        ws = _lib.scan_workspace(dev, sum(wbytes), layout=("ss2d",) + tuple(wbytes)) if sum(wbytes) else None
        off = 0
        grads, keep = [], []
        for m in range(n_maps):
            t = tensors[m * _PER_MAP:(m + 1) * _PER_MAP]
            x = t[0]
            Bsz, C, H, W = x.shape
            L = H * W
            if ctx.want_planes:   # the gradient of the planes IS (dy, dy^T)
                g = dys[m].to(torch.float32).contiguous()
                dy, dyT = g[0], g[1]
            else:
                dy = dys[m].to(torch.float32).contiguous()
                dyT = torch.empty((Bsz, C, L), dtype=torch.float32, device=dev)
            planes = torch.empty((2, Bsz, C, L), dtype=torch.float32, device=dev)
            ddts = torch.empty((2, Bsz, 2, C, L), dtype=torch.float32, device=dev)   # [rm | cm]
            dBC = torch.zeros((2, 4, Bsz, L), dtype=torch.float32, device=dev)       # [dB | dC], direction-major
            small = torch.zeros((3, 4 * C), dtype=torch.float32, device=dev)         # dA, dD, ddelta_bias
            p = arr[m]
            _fill(p, *t, ctx.softplus)
            p.planes, p.states = planes.data_ptr(), states_all[m].data_ptr()
            p.dy, p.dyT, p.dx = dy.data_ptr(), dyT.data_ptr(), None
            p.flags = _lib.SS2D_DYT_GIVEN if ctx.want_planes else 0
            for k in range(4):
                d = ddts[k % 2][:, k // 2]
                p.ddelta[k], p.ddelta_batch_stride[k], p.ddelta_d_stride[k] = d.data_ptr(), d.stride(0), d.stride(1)
            p.dA, p.dD, p.ddelta_bias = small[0].data_ptr(), small[1].data_ptr(), small[2].data_ptr()
            p.dB, p.dC = dBC[0].data_ptr(), dBC[1].data_ptr()
            if wbytes[m]:
                p.workspace, p.workspace_bytes = ws.data_ptr() + off, wbytes[m]
                off += wbytes[m]
            keep.append((dy, dyT))
            # (4, B, L) direction-major -> the (B, 2, 1, L) layout of the forward inputs: k = 2 j + parity
            dB4, dC4 = dBC[0].view(2, 2, Bsz, L), dBC[1].view(2, 2, Bsz, L)
            pair = lambda g, par: g[:, par].permute(1, 0, 2).unsqueeze(2)
            grads += [planes[0].view(Bsz, C, H, W), planes[1].view(Bsz, C, W, H), ddts[0], ddts[1],
                      pair(dB4, 0), pair(dB4, 1), pair(dC4, 0), pair(dC4, 1),
                      small[0].view(4 * C, 1), small[1].view_as(t[9]), small[2].view_as(t[10])]
        with torch.cuda.device(dev):
            _lib.check(lib.vmasr_ss2d_core_bwd(n_maps, arr))
        return (None, None, *grads)


_PER_MAP_PROJ = 8  # tensors per map of _SS2DScanProj


def _check_map_proj(x, xT, xdbl_rm, xdbl_cm, dt_w, As, Ds, bias):
    Bsz, C, H, W = x.shape
    L = H * W
    ok = x.is_contiguous() and xT.is_contiguous() and tuple(xT.shape) == (Bsz, C, W, H) and L > _lib.SCAN_CHUNK
    for t in (xdbl_rm, xdbl_cm):
        ok = ok and tuple(t.shape) == (Bsz, 2, 3, L) and t.is_contiguous()
    ok = ok and tuple(dt_w.shape) == (4, C, 1) and dt_w.is_contiguous()
    ok = ok and tuple(As.shape) == (4 * C, 1) and As.is_contiguous() and Ds.numel() == 4 * C and bias.numel() == 4 * C
    for t in (x, xT, xdbl_rm, xdbl_cm, dt_w, As, Ds, bias):
        ok = ok and t.is_cuda and t.dtype == torch.float32
    if not ok:
        raise RuntimeError("ss2d_core: the projected fused core expects float32 CUDA tensors x (B,C,H,W), xT (B,C,W,H), "
                           "contiguous x_dbl (B,2,3,L) per pair, dt_projs_weight (4,C,1), As (4C,1), Ds / delta_bias (4C) "
                           f"and H*W > {_lib.SCAN_CHUNK}")


def _fill_proj(p: SS2DParams, x, xT, xdbl_rm, xdbl_cm, dt_w, As, Ds, bias, softplus):
    Bsz, C, H, W = x.shape
    p.x, p.xT = x.data_ptr(), xT.data_ptr()
    for k in range(4):
        xd = (xdbl_rm if k % 2 == 0 else xdbl_cm)[:, k // 2]
        p.x_dbl[k], p.x_dbl_batch_stride[k], p.x_dbl_row_stride[k] = xd.data_ptr(), xd.stride(0), xd.stride(1)
    p.dt_weight, p.dt_rank = dt_w.data_ptr(), 1
    p.A, p.D, p.delta_bias = As.data_ptr(), Ds.data_ptr(), bias.data_ptr()
    p.batch, p.channels, p.H, p.W = Bsz, C, H, W
    p.delta_softplus = 1 if softplus else 0
    p.device = _dev(x)
    p.stream = _lib.current_stream_ptr(x.device)


class _SS2DScanProj(torch.autograd.Function):
    """``_SS2DScan`` with the dt projection inside the scan kernels (vmamba.py:1476-1477): the kernels are
    handed ``x_dbl`` itself -- rows (dt, B, C) of each direction in the memory order of its pair -- and ``dt_projs_weight``,
    and form ``delta = w_d * dt_row`` tile by tile.  No (B, 4C, L) ``delta`` is written or read, and the backward returns
    ``d x_dbl`` (d dt-row, dB, dC side by side) and ``d dt_projs_weight`` instead of a (B, 4C, L) ``ddelta``.  dt_rank 1 and
    more than one 2048-chunk only (the three largest maps of every config).  Per map, in order:
    x (B,C,H,W), xT (B,C,W,H), x_dbl_rm, x_dbl_cm (B,2,3,L), dt_projs_weight (4,C,1), As (4C,1), Ds (4C), delta_bias (4C)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, softplus, n_maps, *tensors):
        lib = _lib.load_library()
        want_planes, softplus = bool(int(softplus) & _PLANES), bool(int(softplus) & _SOFTPLUS)
        arr = (SS2DParams * n_maps)()
        keep, ys = [], []
        dev = tensors[0].device
        wbytes = [int(lib.vmasr_ss2d_workspace_bytes(*tensors[m * _PER_MAP_PROJ].shape)) for m in range(n_maps)]
        ws = _lib.scan_workspace(dev, sum(wbytes), layout=("ss2d",) + tuple(wbytes))
        off = 0
        for m in range(n_maps):
            t = tensors[m * _PER_MAP_PROJ:(m + 1) * _PER_MAP_PROJ]
            _check_map_proj(*t)
            Bsz, C, H, W = t[0].shape
            L = H * W
            n_chunks = (L + _lib.SCAN_CHUNK - 1) // _lib.SCAN_CHUNK
            y = None if want_planes else torch.empty((Bsz, C, L), dtype=torch.float32, device=dev)
            planes = torch.empty((2, Bsz, C, L), dtype=torch.float32, device=dev)
            states = torch.empty((4, Bsz, C, n_chunks, 2), dtype=torch.float32, device=dev)
            p = arr[m]
            _fill_proj(p, *t, softplus)
            p.y, p.planes, p.states = (None if want_planes else y.data_ptr()), planes.data_ptr(), states.data_ptr()
            p.workspace, p.workspace_bytes = ws.data_ptr() + off, wbytes[m]
            off += wbytes[m]
            keep.append((planes, states))
            ys.append(planes if want_planes else y)
        with torch.cuda.device(dev):
            _lib.check(lib.vmasr_ss2d_core_fwd(n_maps, arr))
        ctx.softplus, ctx.n_maps, ctx.want_planes = softplus, n_maps, want_planes
        ctx.save_for_backward(*tensors, *[k[1] for k in keep])
        return tuple(ys) if n_maps > 1 else ys[0]

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, *dys):
        lib = _lib.load_library()
        n_maps = ctx.n_maps
        saved = ctx.saved_tensors
        tensors, states_all = saved[:n_maps * _PER_MAP_PROJ], saved[n_maps * _PER_MAP_PROJ:]
        arr = (SS2DParams * n_maps)()
        dev = tensors[0].device
        wbytes = [int(lib.vmasr_ss2d_workspace_bytes(*tensors[m * _PER_MAP_PROJ].shape)) for m in range(n_maps)]
        ws = _lib.scan_workspace(dev, sum(wbytes), layout=("ss2d",) + tuple(wbytes))
        off = 0
        grads, keep = [], []
        for m in range(n_maps):
            t = tensors[m * _PER_MAP_PROJ:(m + 1) * _PER_MAP_PROJ]
            Bsz, C, H, W = t[0].shape
            L = H * W
            if ctx.want_planes:   # the gradient of the planes IS (dy, dy^T)
                g = dys[m].to(torch.float32).contiguous()
                dy, dyT = g[0], g[1]
            else:
                dy = dys[m].to(torch.float32).contiguous()
                dyT = torch.empty((Bsz, C, L), dtype=torch.float32, device=dev)
            planes = torch.empty((2, Bsz, C, L), dtype=torch.float32, device=dev)
            dxdbl = torch.zeros((2, Bsz, 2, 3, L), dtype=torch.float32, device=dev)   # [rm | cm], accumulated into
            small = torch.zeros((4, 4 * C), dtype=torch.float32, device=dev)          # dA, dD, ddelta_bias, d dt_weight
            p = arr[m]
            _fill_proj(p, *t, ctx.softplus)
            p.planes, p.states = planes.data_ptr(), states_all[m].data_ptr()
            p.dy, p.dyT, p.dx = dy.data_ptr(), dyT.data_ptr(), None
            p.flags = _lib.SS2D_DYT_GIVEN if ctx.want_planes else 0
            for k in range(4):
                p.d_x_dbl[k] = dxdbl[k % 2][:, k // 2].data_ptr()
            p.dA, p.dD, p.ddelta_bias, p.d_dt_weight = (small[i].data_ptr() for i in range(4))
            p.workspace, p.workspace_bytes = ws.data_ptr() + off, wbytes[m]
            off += wbytes[m]
            keep.append((dy, dyT))
            grads += [planes[0].view(Bsz, C, H, W), planes[1].view(Bsz, C, W, H), dxdbl[0], dxdbl[1],
                      small[3].view(4, C, 1), small[0].view(4 * C, 1), small[1].view_as(t[6]), small[2].view_as(t[7])]
        with torch.cuda.device(dev):
            _lib.check(lib.vmasr_ss2d_core_bwd(n_maps, arr))
        return (None, None, *grads)


class MergeNormGate(torch.autograd.Function):
    """The block's tail fused into the merge of the core's planes (``vmasr_outnorm_gate_fwd`` / ``_bwd``):
    ``y = P_rm + transpose(P_cm)`` (vmamba.py:57-60) -> ``out_norm = LayerNorm(C)`` on (B, L, C) (:1527-1529) -> ``.to(dtype)``
    (:1531) -> ``* act(z)`` (forwardv2 :1536-1550).  planes (2, B, C, L) float32, gamma / beta (C), z (B, H, W, C) or None
    (dtype of z = dtype of the result; without z: ``out_dtype``).  Returns (B, H, W, C)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, planes, gamma, beta, z, H, W, eps, z_silu, out_dtype):
        lib = _lib.load_library()
        _, Bsz, C, L = planes.shape
        dev = planes.device
        io = z.dtype if z is not None else out_dtype
        if planes.dtype != torch.float32 or not planes.is_contiguous() or L != H * W or io not in _lib.DTYPE_CODE:
            raise RuntimeError("MergeNormGate: planes must be contiguous float32 (2, B, C, H*W); io dtype float32 / float16 / bfloat16")
        if z is not None:
            z = z.contiguous()
            if tuple(z.shape) != (Bsz, H, W, C):
                raise RuntimeError("MergeNormGate: z must be (B, H, W, C)")
        need = any(ctx.needs_input_grad[:4])
        out = torch.empty((Bsz, H, W, C), dtype=io, device=dev)
        y = torch.empty((Bsz, C, L), dtype=torch.float32, device=dev) if need else None
        stats = torch.empty((Bsz, L, 2), dtype=torch.float32, device=dev) if need else None
        g32 = None if gamma is None else gamma.detach().to(torch.float32).contiguous()
        b32 = None if beta is None else beta.detach().to(torch.float32).contiguous()
        p = _lib.OutNormParams()
        p.p_rm, p.p_cm = planes[0].data_ptr(), planes[1].data_ptr()
        p.gamma, p.beta = (None if g32 is None else g32.data_ptr()), (None if b32 is None else b32.data_ptr())
        p.z, p.out = (None if z is None else z.data_ptr()), out.data_ptr()
        p.y, p.stats = (None if y is None else y.data_ptr()), (None if stats is None else stats.data_ptr())
        p.eps, p.batch, p.channels, p.H, p.W = float(eps), Bsz, C, H, W
        p.io_dtype, p.z_silu, p.device = _lib.DTYPE_CODE[io], int(bool(z_silu)), _dev(planes)
        p.stream = _lib.current_stream_ptr(dev)
        with torch.cuda.device(dev):
            _lib.check(lib.vmasr_outnorm_gate_fwd(ctypes.byref(p)))
        ctx.save_for_backward(y, stats, z, g32, b32)
        ctx.dims = (Bsz, C, H, W, float(eps), bool(z_silu), io, gamma is not None, beta is not None)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dout):
        lib = _lib.load_library()
        y, stats, z, g32, b32 = ctx.saved_tensors
        Bsz, C, H, W, eps, z_silu, io, has_g, has_b = ctx.dims
        dev, L = y.device, H * W
        dout = dout.to(io).contiguous()
        g = torch.empty((2, Bsz, C, L), dtype=torch.float32, device=dev)     # (dy, dy^T): the gradient of the planes
        dz = torch.empty_like(z) if z is not None else None
        patches = int(lib.vmasr_outnorm_patches(Bsz, C, H, W))
        part = torch.empty((patches, 2, C), dtype=torch.float32, device=dev)
        p = _lib.OutNormParams()
        p.gamma, p.beta = (None if g32 is None else g32.data_ptr()), (None if b32 is None else b32.data_ptr())
        p.z, p.dz = (None if z is None else z.data_ptr()), (None if dz is None else dz.data_ptr())
        p.y, p.stats, p.dout, p.dy, p.dgb_partial = y.data_ptr(), stats.data_ptr(), dout.data_ptr(), g[0].data_ptr(), part.data_ptr()
        p.eps, p.batch, p.channels, p.H, p.W = eps, Bsz, C, H, W
        p.io_dtype, p.z_silu, p.device = _lib.DTYPE_CODE[io], int(z_silu), _dev(y)
        p.stream = _lib.current_stream_ptr(dev)
        with torch.cuda.device(dev):
            _lib.check(lib.vmasr_outnorm_gate_bwd(ctypes.byref(p)))
            _lib.check(lib.vmasr_map_transpose(g[0].data_ptr(), g[1].data_ptr(), Bsz * C, H, W, _dev(y), p.stream))
        sums = part.sum(0)
        return g, (sums[0] if has_g else None), (sums[1] if has_b else None), dz, None, None, None, None, None


class ConvSiluInput(torch.autograd.Function):
    """The block's head fused into the core's load (``vmasr_dwconv_silu_fwd`` / ``_bwd``): channel-last
    ``xin (B, H, W, C)`` -- e.g. the x half of in_proj's output, read in place through its position stride -- ->
    ``permute -> conv2d (depthwise 3x3, padding 1) -> SiLU`` (vmamba.py:1541-1546) -> ``(x (B, C, H, W), x^T (B, C, W, H))`` in
    float32, the two tensors the fused core reads.  weight (C, 1, 3, 3), bias (C) or None.

    With ``x_proj_weight`` (4, R + 2, C) [and ``x_proj_bias`` (4, R + 2)] the same kernel also forms
    ``x_dbl = einsum(xs, x_proj_weight)`` (vmamba.py:1473-1475) -- the rows of directions (0, 2) in row-major and of (1, 3) in
    column-major position order, ``(B, 2, R + 2, L)`` each, what ``_SS2DScanProj`` takes -- and returns
    ``(x, x^T, x_dbl_rm, x_dbl_cm)``: the einsum's cuBLAS launches and their passes over x and x^T disappear."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, xin, weight, bias, x_proj_weight=None, x_proj_bias=None):
        lib = _lib.load_library()
        _lib.require_cuda(xin, "xin")
        if xin.dim() != 4 or xin.dtype not in _lib.DTYPE_CODE or tuple(weight.shape[1:]) != (1, 3, 3) or weight.shape[0] != xin.shape[3]:
            raise RuntimeError("ConvSiluInput: xin (B, H, W, C) float32 / float16 / bfloat16, weight (C, 1, 3, 3)")
        Bsz, H, W, C = xin.shape
        ps = xin.stride(2)
        if not (xin.stride(3) == 1 and ps >= C and xin.stride(1) == W * ps and xin.stride(0) == H * W * ps):
            xin = xin.contiguous()
            ps = C
        dev = xin.device
        w32 = weight.detach().to(torch.float32).contiguous()
        b32 = None if bias is None else bias.detach().to(torch.float32).contiguous()
        x = torch.empty((Bsz, C, H, W), dtype=torch.float32, device=dev)
        xT = torch.empty((Bsz, C, W, H), dtype=torch.float32, device=dev)
        p = _lib.DwConvParams()
        p.xin, p.weight, p.bias = xin.data_ptr(), w32.data_ptr(), (None if b32 is None else b32.data_ptr())
        p.x, p.xT, p.xin_pos_stride = x.data_ptr(), xT.data_ptr(), ps
        p.batch, p.channels, p.H, p.W = Bsz, C, H, W
        p.io_dtype, p.device, p.stream = _lib.DTYPE_CODE[xin.dtype], _dev(xin), _lib.current_stream_ptr(dev)
        xw32 = xb32 = xd = None
        if x_proj_weight is not None:
            if x_proj_weight.dim() != 3 or x_proj_weight.shape[0] != 4 or x_proj_weight.shape[2] != C:
                raise RuntimeError("ConvSiluInput: x_proj_weight must be (4, R + 2, C)")
            RP = x_proj_weight.shape[1]
            xw32 = x_proj_weight.detach().to(torch.float32).contiguous()
            xb32 = None if x_proj_bias is None else x_proj_bias.detach().to(torch.float32).contiguous()
            # [rm | cm]; partial sums are added into them when the channels of a patch span several CTAs
            alloc = torch.zeros if int(lib.vmasr_dwconv_channel_blocks(Bsz, C, H, W)) > 1 else torch.empty
            xd = alloc((2, Bsz, 2, RP, H * W), dtype=torch.float32, device=dev)
            p.x_proj_weight, p.x_proj_bias = xw32.data_ptr(), (None if xb32 is None else xb32.data_ptr())
            p.x_dbl_rm, p.x_dbl_cm, p.x_proj_rows = xd[0].data_ptr(), xd[1].data_ptr(), RP
        with torch.cuda.device(dev):
            _lib.check(lib.vmasr_dwconv_silu_fwd(ctypes.byref(p)))
        ctx.save_for_backward(xin, w32, b32, xw32)
        ctx.ps, ctx.has_bias, ctx.has_xb = ps, bias is not None, x_proj_bias is not None
        if xd is not None:
            return x, xT, xd[0], xd[1]
        return x, xT

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dx, dxT, dxd_rm=None, dxd_cm=None):
        lib = _lib.load_library()
        xin, w32, b32, xw32 = ctx.saved_tensors
        Bsz, H, W, C = xin.shape
        dev = xin.device
        dx, dxT = dx.to(torch.float32).contiguous(), dxT.to(torch.float32).contiguous()
        dxin = torch.empty((Bsz, H, W, C), dtype=xin.dtype, device=dev)
        patches = int(lib.vmasr_dwconv_patches(Bsz, C, H, W))
        part = torch.empty((patches, C, 10), dtype=torch.float32, device=dev)
        p = _lib.DwConvParams()
        p.xin, p.weight, p.bias = xin.data_ptr(), w32.data_ptr(), (None if b32 is None else b32.data_ptr())
        p.dx, p.dxT, p.dxin, p.dwb_partial, p.xin_pos_stride = dx.data_ptr(), dxT.data_ptr(), dxin.data_ptr(), part.data_ptr(), ctx.ps
        p.batch, p.channels, p.H, p.W = Bsz, C, H, W
        p.io_dtype, p.device, p.stream = _lib.DTYPE_CODE[xin.dtype], _dev(xin), _lib.current_stream_ptr(dev)
        d_xw = d_xb = xpart = None
        if xw32 is not None:
            RP = xw32.shape[1]
            dxd_rm, dxd_cm = dxd_rm.to(torch.float32).contiguous(), dxd_cm.to(torch.float32).contiguous()
            xpart = torch.empty((patches, 4 * RP, C), dtype=torch.float32, device=dev)
            p.x_proj_weight, p.x_proj_rows = xw32.data_ptr(), RP
            p.d_x_dbl_rm, p.d_x_dbl_cm, p.d_x_proj_weight_partial = dxd_rm.data_ptr(), dxd_cm.data_ptr(), xpart.data_ptr()
        with torch.cuda.device(dev):
            _lib.check(lib.vmasr_dwconv_silu_bwd(ctypes.byref(p)))
        sums = part.sum(0)
        if xw32 is not None:
            d_xw = xpart.sum(0).view(4, RP, C)
            if ctx.has_xb:   # rows k = 2 j + parity: d x_proj_bias[k, r] = sum over batch and positions of d x_dbl
                d_xb = torch.stack([dxd_rm.sum((0, 3)), dxd_cm.sum((0, 3))], dim=1).reshape(4, RP)
        return dxin, sums[:, :9].reshape(C, 1, 3, 3), (sums[:, 9] if ctx.has_bias else None), d_xw, d_xb


def outnorm_fusable(x: torch.Tensor, N: int) -> bool:
    """the fused tail applies: fused core + W a multiple of 8 + a patch of all channels fits shared memory"""
    if not _fusable(x, N) or x.shape[3] % 8:
        return False
    return int(_lib.load_library().vmasr_outnorm_patches(x.shape[0], x.shape[1], x.shape[2], x.shape[3])) > 0


def ss2d_core_out(x: torch.Tensor, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds, out_norm_weight, out_norm_bias,
                  z: torch.Tensor | None = None, z_silu: bool = True, eps: float = 1e-5, delta_softplus: bool = True,
                  x_proj_bias=None, projected: bool | None = None, xT: torch.Tensor | None = None, out_dtype=None, xd=None) -> torch.Tensor:
    """``forward_corev2`` INCLUDING its tail and the gate of ``forwardv2`` (vmamba.py:1472-1531, 1536-1550) for the configs'
    layout (channel_first False, out_norm = nn.LayerNorm): x (B, C, H, W) -> (B, H, W, C) in the dtype of z (of x without a
    gate).  The core's planes go straight into ``MergeNormGate``: the merged map is written once (for the backward) and never
    read back in the forward; transpose, LayerNorm, cast, SiLU and the product are not separate passes."""
    N = A_logs.shape[1]
    if not outnorm_fusable(x, N):
        raise RuntimeError("ss2d_core_out: needs a CUDA map with H % 4 == 0, W % 8 == 0, d_state 1 and channels that fit a patch")
    if projected is None:
        projected = _projectable(x, dt_projs_weight, N)
    mode = (_SOFTPLUS if delta_softplus else 0) | _PLANES
    if projected:
        planes = _SS2DScanProj.apply(mode, 1, *_prepare_proj(x, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds, x_proj_bias, xT, xd))
    else:
        planes = _SS2DScan.apply(mode, 1, *_prepare(x, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds, x_proj_bias, xT, xd))
    return MergeNormGate.apply(planes, out_norm_weight, out_norm_bias, z, x.shape[2], x.shape[3], eps, z_silu, out_dtype or x.dtype)


def ss2d_block_core(x_cl: torch.Tensor, conv_weight, conv_bias, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds,
                    out_norm_weight, out_norm_bias, z: torch.Tensor | None = None, z_silu: bool = True, eps: float = 1e-5,
                    delta_softplus: bool = True, x_proj_bias=None, projected: bool | None = None, fuse_x_proj: bool = True) -> torch.Tensor:
    """Everything of ``SS2D.forwardv2`` between in_proj and out_proj (vmamba.py:1536-1550 with forward_corev2 inside) for the
    configs' layout: channel-last x_cl (B, H, W, C) -- a strided view of in_proj's output is read in place -- -> depthwise
    conv 3x3 + SiLU + permute + transpose (``ConvSiluInput``) -> fused core -> merge + LayerNorm + cast + gate
    (``MergeNormGate``) -> (B, H, W, C).  With ``fuse_x_proj`` the head kernel also forms ``x_dbl`` (the x_proj einsum of
    vmamba.py:1473-1475), and where ``dt_rank`` is 1 the scan kernels form ``delta`` themselves: the block then runs on four
    kernels of this library and nothing else -- no ``xs`` / ``ys`` / ``dts`` copies, no einsum, no separate permute, activation,
    transpose, normalisation or gate passes."""
    xd = None
    if fuse_x_proj:
        x, xT, xd_rm, xd_cm = ConvSiluInput.apply(x_cl, conv_weight, conv_bias, x_proj_weight, x_proj_bias)
        xd = (xd_rm, xd_cm)
    else:
        x, xT = ConvSiluInput.apply(x_cl, conv_weight, conv_bias)
    return ss2d_core_out(x, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds, out_norm_weight, out_norm_bias, z=z,
                         z_silu=z_silu, eps=eps, delta_softplus=delta_softplus, x_proj_bias=x_proj_bias, projected=projected,
                         xT=xT, out_dtype=x_cl.dtype, xd=xd)


def ss2d_block_core_pair(x_cl_a, blk_a, x_cl_b, blk_b, z_a=None, z_b=None, z_silu: bool = True, eps: float = 1e-5,
                         delta_softplus: bool = True):
    """``ss2d_block_core`` for the generator's two streams (same shapes, independent until their interaction point,
    model/model.py:1124-1131): the two heads and the two tails are launched one after the other, the two cores' scans as ONE
    grid.  ``blk_*`` = (conv_weight, conv_bias, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds, out_norm_weight,
    out_norm_bias[, x_proj_bias]).  Returns (out_a, out_b), each (B, H, W, C)."""
    Bsz, H, W, C = x_cl_a.shape
    proj = all(_projectable(x.permute(0, 3, 1, 2), blk[3], blk[5].shape[1]) for x, blk in ((x_cl_a, blk_a), (x_cl_b, blk_b)))
    args = []
    for x_cl, blk in ((x_cl_a, blk_a), (x_cl_b, blk_b)):
        xpb = blk[9] if len(blk) > 9 else None
        if not outnorm_fusable(x_cl.permute(0, 3, 1, 2), blk[5].shape[1]):
            raise RuntimeError("ss2d_block_core_pair: needs CUDA maps with H % 4 == 0, W % 8 == 0, d_state 1")
        x, xT, xd_rm, xd_cm = ConvSiluInput.apply(x_cl, blk[0], blk[1], blk[2], xpb)
        args += list((_prepare_proj if proj else _prepare)(x, blk[2], blk[3], blk[4], blk[5], blk[6], xpb, xT, (xd_rm, xd_cm)))
    mode = (_SOFTPLUS if delta_softplus else 0) | _PLANES
    planes = (_SS2DScanProj if proj else _SS2DScan).apply(mode, 2, *args)
    outs = []
    for pl, blk, z, x_cl in ((planes[0], blk_a, z_a, x_cl_a), (planes[1], blk_b, z_b, x_cl_b)):
        outs.append(MergeNormGate.apply(pl, blk[7], blk[8], z, H, W, eps, z_silu, x_cl.dtype))
    return tuple(outs)


def _projections(x, xT, x_proj_weight, x_proj_bias, dt_projs_weight, R, N, xd=None):
    """x_dbl and dts of vmamba.py:1473-1477 in memory order: the row-major pair from the map, the column-major pair from
    its transpose.  Returns dts (B,2,C,L), Bs, Cs (B,2,N,L) per pair; Bs / Cs are VIEWS of x_dbl (no contiguous copies).
    ``xd`` = (x_dbl_rm, x_dbl_cm) when the head kernel has formed them already."""
    Bsz, C, H, W = x.shape
    L = H * W
    out = []
    for par, src in ((0, x.view(Bsz, C, L)), (1, xT.view(Bsz, C, L))):
        if xd is not None:
            x_dbl = xd[par]
        else:
            x_dbl = torch.einsum("bdl,kcd->bkcl", src, x_proj_weight[par::2])
            if x_proj_bias is not None:
                x_dbl = x_dbl + x_proj_bias[par::2].view(1, 2, -1, 1)
        dts, Bs, Cs = torch.split(x_dbl, [R, N, N], dim=2)
        dts = torch.einsum("bkrl,kdr->bkdl", dts, dt_projs_weight[par::2])
        out.append((dts, Bs, Cs))
    return out


def _fusable(x, N):
    return x.is_cuda and x.dim() == 4 and x.shape[2] % 4 == 0 and x.shape[3] % 4 == 0 and N == 1


def _projectable(x, dt_projs_weight, N):
    """delta can be generated inside the scan kernels: fused core, dt_rank 1, more than one chunk of 16-float lines"""
    L = x.shape[2] * x.shape[3]
    return _fusable(x, N) and dt_projs_weight.shape[2] == 1 and L > _lib.SCAN_CHUNK and L % 16 == 0


def _prepare_proj(x, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds, x_proj_bias, xT=None, xd=None):
    """inputs of _SS2DScanProj: x_dbl (vmamba.py:1473-1475) of the two pairs in memory order; the dt projection (:1477) is
    left to the kernels"""
    Bsz, C, H, W = x.shape
    L = H * W
    x32 = x.to(torch.float32).contiguous()
    xT = MapTranspose.apply(x32) if xT is None else xT
    if xd is None:
        xd = []
        for par, src in ((0, x32.view(Bsz, C, L)), (1, xT.view(Bsz, C, L))):
            x_dbl = torch.einsum("bdl,kcd->bkcl", src, x_proj_weight[par::2].to(torch.float32))
            if x_proj_bias is not None:
                x_dbl = x_dbl + x_proj_bias[par::2].view(1, 2, -1, 1)
            xd.append(x_dbl.to(torch.float32).contiguous())
    As = -torch.exp(A_logs.to(torch.float))
    return (x32, xT, xd[0], xd[1], dt_projs_weight.to(torch.float).contiguous(), As.contiguous(), Ds.to(torch.float).contiguous(),
            dt_projs_bias.reshape(-1).to(torch.float).contiguous())


def _prepare(x, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds, x_proj_bias, xT=None, xd=None):
    K, _, R = dt_projs_weight.shape
    N = A_logs.shape[1]
    x32 = x.to(torch.float32).contiguous()
    xT = MapTranspose.apply(x32) if xT is None else xT
    (dts_rm, Bs_rm, Cs_rm), (dts_cm, Bs_cm, Cs_cm) = _projections(x32, xT, x_proj_weight, x_proj_bias, dt_projs_weight, R, N, xd)
    # force_fp32 (vmamba.py:1487-1491; a no-op outside autocast); einsum is free to return any strides: the kernels need
    # unit stride along L only (the reference makes Bs / Cs / dts contiguous unconditionally, vmamba.py:1480-1483)
    f = lambda t: t.to(torch.float32) if t.stride(-1) == 1 else t.to(torch.float32).contiguous()
    As = -torch.exp(A_logs.to(torch.float))                           # vmamba.py:1481
    return (x32, xT, f(dts_rm).contiguous(), f(dts_cm).contiguous(), f(Bs_rm), f(Bs_cm), f(Cs_rm), f(Cs_cm), As.contiguous(),
            Ds.to(torch.float).contiguous(), dt_projs_bias.reshape(-1).to(torch.float).contiguous())


def ss2d_core(x: torch.Tensor, x_proj_weight: torch.Tensor, dt_projs_weight: torch.Tensor, dt_projs_bias: torch.Tensor,
              A_logs: torch.Tensor, Ds: torch.Tensor, delta_softplus: bool = True, force_fp32: bool = True,
              x_proj_bias: torch.Tensor | None = None, fused: bool | None = None, projected: bool | None = None) -> torch.Tensor:
    """x (B, C, H, W) -> y (B, C, H*W), float32.  Parameter layouts as in ``SS2D.__initv2__`` (vmamba.py:772-850):
    x_proj_weight (K=4, R + 2N, C), dt_projs_weight (K, C, R), dt_projs_bias (K, C), A_logs (K*C, N), Ds (K*C),
    x_proj_bias (K, R + 2N) or None (vmamba.py:1474-1475).  ``fused=None`` picks the fused core whenever it applies;
    ``projected=None`` additionally leaves the dt projection to the scan kernels wherever they take it (dt_rank 1, H*W > 2048)."""
    if x.dim() != 4:
        raise RuntimeError("ss2d_core: expected (B, C, H, W)")
    N = A_logs.shape[1]
    if fused is None:
        fused = _fusable(x, N)
    if not fused:
        return ss2d_core_chain(x, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds, delta_softplus, force_fp32, x_proj_bias)
    if not _fusable(x, N):
        raise RuntimeError("ss2d_core: the fused core needs a CUDA map with H, W multiples of 4 and d_state 1")
    if projected is None:
        projected = _projectable(x, dt_projs_weight, N)
    if projected:
        if not _projectable(x, dt_projs_weight, N):
            raise RuntimeError("ss2d_core: the projected form needs the fused core, dt_rank 1 and H*W > 2048 (a multiple of 16)")
        return _SS2DScanProj.apply(delta_softplus, 1, *_prepare_proj(x, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds, x_proj_bias))
    return _SS2DScan.apply(delta_softplus, 1, *_prepare(x, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds, x_proj_bias))


def ss2d_core_pair(x_a, params_a, x_b, params_b, delta_softplus: bool = True, projected: bool | None = None):
    """The cores of the generator's two streams as one grid.  ``params_*`` = (x_proj_weight, dt_projs_weight, dt_projs_bias,
    A_logs, Ds[, x_proj_bias]).  Returns (y_a, y_b).  ``projected`` as in ``ss2d_core`` (both maps or neither)."""
    for x, prm in ((x_a, params_a), (x_b, params_b)):
        if not _fusable(x, prm[3].shape[1]):
            raise RuntimeError("ss2d_core_pair: the fused core needs CUDA maps with H, W multiples of 4 and d_state 1")
    can = all(_projectable(x, prm[1], prm[3].shape[1]) for x, prm in ((x_a, params_a), (x_b, params_b)))
    if projected is None:
        projected = can
    if projected and not can:
        raise RuntimeError("ss2d_core_pair: the projected form needs dt_rank 1 and H*W > 2048 (a multiple of 16) on both maps")
    args = []
    for x, prm in ((x_a, params_a), (x_b, params_b)):
        bias = prm[5] if len(prm) > 5 else None
        args += list((_prepare_proj if projected else _prepare)(x, prm[0], prm[1], prm[2], prm[3], prm[4], bias))
    return (_SS2DScanProj if projected else _SS2DScan).apply(delta_softplus, 2, *args)


def ss2d_core_chain(x: torch.Tensor, x_proj_weight: torch.Tensor, dt_projs_weight: torch.Tensor, dt_projs_bias: torch.Tensor,
                    A_logs: torch.Tensor, Ds: torch.Tensor, delta_softplus: bool = True, force_fp32: bool = True,
                    x_proj_bias: torch.Tensor | None = None) -> torch.Tensor:
    """The same core as a chain of the three operators, statement for statement vmamba.py:1472-1497 (materialises ``xs``
    and ``ys``)."""
    Bsz, C, H, W = x.shape
    K, _, R = dt_projs_weight.shape
    N = A_logs.shape[1]
    L = H * W
    xs = CrossScan.apply(x)                                           # vmamba.py:1472
    x_dbl = torch.einsum("bkdl,kcd->bkcl", xs, x_proj_weight)         # :1473
    if x_proj_bias is not None:
        x_dbl = x_dbl + x_proj_bias.view(1, K, -1, 1)                 # :1474-1475
    dts, Bs, Cs = torch.split(x_dbl, [R, N, N], dim=2)                # :1476
    dts = torch.einsum("bkrl,kdr->bkdl", dts, dt_projs_weight)        # :1477
    xs = xs.view(Bsz, -1, L)
    dts = dts.contiguous().view(Bsz, -1, L)
    As = -torch.exp(A_logs.to(torch.float))                           # :1481
    Bs, Cs = Bs.contiguous(), Cs.contiguous()
    Ds = Ds.to(torch.float)
    delta_bias = dt_projs_bias.view(-1).to(torch.float)
    if force_fp32:                                                    # :1487-1491
        xs, dts, Bs, Cs = xs.to(torch.float), dts.to(torch.float), Bs.to(torch.float), Cs.to(torch.float)
    ys = SelectiveScanCore.apply(xs, dts, As, Bs, Cs, Ds, delta_bias, delta_softplus)   # :1493-1495
    return CrossMerge.apply(ys.view(Bsz, K, -1, H, W))                # :1497
