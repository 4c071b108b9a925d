"""The SS2D core of a VSS block, ``SS2D.forward_corev2`` (model/vmamba.py:1472-1497), on this library's operators:

    xs = CrossScan(x) -> x_dbl = einsum(xs, x_proj_weight) -> dts = einsum(dts, dt_projs_weight)
       -> ys = SelectiveScanCore(xs, dts, -exp(A_logs), Bs, Cs, Ds, dt_projs_bias, delta_softplus=True) -> y = CrossMerge(ys)

The two small einsums stay on PyTorch/cuBLAS (they are not on the named path); the scan inputs are cast to fp32 as the
reference does with ``force_fp32`` (vmamba.py:1487-1491).  This is a CHAIN of the library's kernels, differentiable end
to end through their autograd functions; a single kernel that reads the map through the four index maps and writes the
merged map (no ``xs`` / ``ys`` copies) is the next row of SURVEY.md 8(f), not this function."""
from __future__ import annotations

import torch

from .cross import CrossMerge, CrossScan
from .scan import SelectiveScanCore


def ss2d_core(x: torch.Tensor, x_proj_weight: torch.Tensor, dt_projs_weight: torch.Tensor, dt_projs_bias: torch.Tensor,
              A_logs: torch.Tensor, Ds: torch.Tensor, delta_softplus: bool = True, force_fp32: bool = True) -> torch.Tensor:
    """x (B, C, H, W) -> y (B, C, H*W).  Parameter layouts as in ``SS2D.__initv2__`` (vmamba.py:772-850):
    x_proj_weight (K=4, R + 2N, C), dt_projs_weight (K, C, R), dt_projs_bias (K, C), A_logs (K*C, N), Ds (K*C)."""
    if x.dim() != 4:
        raise RuntimeError("ss2d_core: expected (B, C, H, W)")
    Bsz, C, H, W = x.shape
    K, _, R = dt_projs_weight.shape
    N = A_logs.shape[1]
    L = H * W
    xs = CrossScan.apply(x)                                           # vmamba.py:1472
    x_dbl = torch.einsum("bkdl,kcd->bkcl", xs, x_proj_weight)         # :1473
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
