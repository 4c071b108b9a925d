"""4-direction cross scan / merge: drop-ins for ``CrossScan`` / ``CrossMerge`` (model/vmamba.py:27-73) and for
the Triton variants the shipped configs select, ``CrossScanTriton`` / ``CrossMergeTriton``
(model/csm_triton.py:311-366).  Both directions of both autograd functions are the two CUDA kernels
``vmasr_cross_scan`` / ``vmasr_cross_merge`` (each op's backward is the other op's forward)."""
from __future__ import annotations

import torch

from . import _lib


def _run(kind: str, src: torch.Tensor, dst: torch.Tensor, B, C, H, W):
    lib = _lib.load_library()
    _lib.require_cuda(src, kind)
    if src.dtype not in _lib.DTYPE_CODE:
        raise RuntimeError(f"{kind}: dtype must be float32, float16 or bfloat16, got {src.dtype}")
    fn = lib.vmasr_cross_scan if kind == "cross_scan" else lib.vmasr_cross_merge
    dev = src.device.index if src.device.index is not None else torch.cuda.current_device()
    with torch.cuda.device(src.device):
        _lib.check(fn(src.data_ptr(), dst.data_ptr(), B, C, H, W, _lib.DTYPE_CODE[src.dtype], dev,
                      _lib.current_stream_ptr(src.device)))
    return dst


def cross_scan(x: torch.Tensor) -> torch.Tensor:
    """(B, C, H, W) -> (B, 4, C, H*W)"""
    if x.dim() != 4:
        raise RuntimeError("cross_scan: expected (B, C, H, W)")
    B, C, H, W = x.shape
    x = x.contiguous()  # csm_triton.py:324
    return _run("cross_scan", x, x.new_empty((B, 4, C, H * W)), B, C, H, W)


def cross_merge(ys: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """(B, 4, C, H*W) or (B, 4, C, H, W) -> (B, C, H*W)"""
    B, K, C = ys.shape[:3]
    if K != 4:
        raise RuntimeError("cross_merge: expected 4 directions")
    ys = ys.contiguous()  # csm_triton.py:353
    return _run("cross_merge", ys, ys.new_empty((B, C, H * W)), B, C, H, W)


class CrossScan(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: torch.Tensor):
        B, C, H, W = x.shape
        ctx.shape = (B, C, H, W)
        return cross_scan(x)

    @staticmethod
    def backward(ctx, ys: torch.Tensor):
        B, C, H, W = ctx.shape
        return cross_merge(ys, H, W).view(B, -1, H, W)


class CrossMerge(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ys: torch.Tensor):
        B, K, D, H, W = ys.shape
        ctx.shape = (H, W)
        return cross_merge(ys, H, W)

    @staticmethod
    def backward(ctx, x: torch.Tensor):
        H, W = ctx.shape
        B, C, L = x.shape
        return cross_scan(x.view(B, C, H, W)).view(B, 4, C, H, W)


def _run_1b1(kind: str, src: torch.Tensor, B, C, H, W):
    lib = _lib.load_library()
    _lib.require_cuda(src, kind)
    if src.dtype not in _lib.DTYPE_CODE:
        raise RuntimeError(f"{kind}: dtype must be float32, float16 or bfloat16, got {src.dtype}")
    src = src.contiguous()
    dst = torch.empty_like(src)
    fn = lib.vmasr_cross_scan_1b1 if kind == "cross_scan_1b1" else lib.vmasr_cross_merge_1b1
    dev = src.device.index if src.device.index is not None else torch.cuda.current_device()
    with torch.cuda.device(src.device):
        _lib.check(fn(src.data_ptr(), dst.data_ptr(), B, C, H, W, _lib.DTYPE_CODE[src.dtype], dev, _lib.current_stream_ptr(src.device)))
    return dst


class CrossScanTriton1b1(torch.autograd.Function):
    """Drop-in for ``model/csm_triton.py`` ``CrossScanTriton1b1`` (:369-395): x (B, 4, C, H, W), one map per direction ->
    (B, 4, C, H*W); the backward is the inverse permutation (``triton_cross_merge_1b1``)."""

    @staticmethod
    def forward(ctx, x: torch.Tensor):
        B, K, C, H, W = x.shape
        if K != 4:
            raise RuntimeError("cross_scan_1b1: expected 4 directions")
        ctx.shape = (B, C, H, W)
        return _run_1b1("cross_scan_1b1", x, B, C, H, W).view(B, 4, C, -1)

    @staticmethod
    def backward(ctx, y: torch.Tensor):
        B, C, H, W = ctx.shape
        return _run_1b1("cross_merge_1b1", y.contiguous().view(B, 4, C, H, W), B, C, H, W)


# the names SS2D.__initv2__ binds for forward_type "v5" (vmamba.py:842-848)
CrossScanTriton = CrossScan
CrossMergeTriton = CrossMerge
