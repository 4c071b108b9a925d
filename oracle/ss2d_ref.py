"""ORACLE (test infrastructure, not product code).

CPU restatement, in plain torch, of the reference's SS2D hot path:

* ``cross_scan`` / ``cross_scan_bwd``   -- reference ``model/vmamba.py:27-47``  (``CrossScan``)
* ``cross_merge`` / ``cross_merge_bwd`` -- reference ``model/vmamba.py:50-73``  (``CrossMerge``)
* ``selective_scan``                    -- reference ``kernels/selective_scan/test_selective_scan.py:287-367``
                                           (``selective_scan_ref``; real ``A``, variable grouped ``B``/``C``)
* ``selective_scan_bwd``                -- the gradients ``selective_scan_bwd_kernel`` produces
                                           (``kernels/selective_scan/csrc/selective_scan/cus/selective_scan_bwd_kernel.cuh:125-272``),
                                           written out as the closed-form adjoint recurrence.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
module.  The shipped operators in ``vm_asr_b200/`` never do: they call the CUDA library or raise.

Parity of this restatement is PINNED: ``tests/test_oracle_golden.py`` checks every function here against
fixtures under ``tests/golden/`` that ``oracle/make_golden.py`` produced by importing and running the
reference's own code (``/root/reference``) in the build container.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# 4-direction cross scan / merge (index maps only; must be bit-exact)
# ----------------------------------------------------------------------------------------------
def cross_scan(x: torch.Tensor) -> torch.Tensor:
    """(B, C, H, W) -> (B, 4, C, H*W).

    Direction 0 walks the map row-major, direction 1 column-major (position ``l = w*H + h``),
    directions 2 and 3 are directions 0 and 1 walked backwards (vmamba.py:31-36).
    """
    B, C, H, W = x.shape
    row_major = x.reshape(B, C, H * W)
    col_major = x.permute(0, 1, 3, 2).reshape(B, C, H * W)
    return torch.stack(
        [row_major, col_major, row_major.flip(-1), col_major.flip(-1)], dim=1
    ).contiguous()


def cross_merge(ys: torch.Tensor) -> torch.Tensor:
    """(B, 4, C, H, W) -> (B, C, H*W).

    The association of the three additions follows vmamba.py:55-60 exactly:
    ``(ys0 + flip(ys2)) + transpose_back(ys1 + flip(ys3))`` -- that is what makes the result
    bit-reproducible.
    """
    B, K, C, H, W = ys.shape
    assert K == 4
    L = H * W
    ys = ys.reshape(B, 4, C, L)
    fwd_pair = ys[:, 0:2] + ys[:, 2:4].flip(-1)
    col_part = fwd_pair[:, 1].reshape(B, C, W, H).permute(0, 1, 3, 2).reshape(B, C, L)
    return fwd_pair[:, 0] + col_part


def cross_scan_bwd(g_xs: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """Gradient of ``cross_scan``: (B, 4, C, L) -> (B, C, H, W)  (vmamba.py:39-47).

    It is the same arithmetic as ``cross_merge``.
    """
    B, K, C, L = g_xs.shape
    return cross_merge(g_xs.reshape(B, K, C, H, W)).reshape(B, C, H, W)


def cross_merge_bwd(g_y: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """Gradient of ``cross_merge``: (B, C, L) -> (B, 4, C, H, W)  (vmamba.py:62-73)."""
    B, C, L = g_y.shape
    return cross_scan(g_y.reshape(B, C, H, W)).reshape(B, 4, C, H, W)


# ----------------------------------------------------------------------------------------------
# selective scan, sequential in L
# ----------------------------------------------------------------------------------------------
def _expand_groups(t: torch.Tensor, dim: int) -> torch.Tensor:
    """(B, G, N, L) -> (B, D, N, L): channel d uses group d // (D/G)."""
    return t.repeat_interleave(dim // t.shape[1], dim=1)


def effective_delta(delta, delta_bias, delta_softplus, dtype=torch.float32):
    dt = delta.to(dtype)
    if delta_bias is not None:
        dt = dt + delta_bias.to(dtype)[None, :, None]
    if delta_softplus:
        dt = F.softplus(dt)  # threshold 20, like the kernel (fwd_kernel.cuh:117)
    return dt


def selective_scan(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False,
                   return_last_state=False, dtype=torch.float32):
    """out[b,d,l] = sum_n C[b,g,n,l] * h[b,d,n,l] + D[d]*u[b,d,l],
    h[.,l] = exp(dt*A) * h[.,l-1] + dt * B * u,  dt = softplus(delta + delta_bias).

    u, delta: (B, D, L); A: (D, N); B, C: (B, G, N, L) or (B, N, L); D, delta_bias: (D,).
    ``dtype`` is the accumulation type (float32 mirrors the reference; float64 is used by the tests
    as the tighter yardstick).  The result is cast back to ``u.dtype``.
    """
    in_dtype = u.dtype
    if B.dim() == 3:
        B = B[:, None]
    if C.dim() == 3:
        C = C[:, None]
    Bsz, Dm, L = u.shape
    N = A.shape[1]
    uf = u.to(dtype)
    dt = effective_delta(delta, delta_bias, delta_softplus, dtype)
    Af = A.to(dtype)
    Bf = _expand_groups(B.to(dtype), Dm)
    Cf = _expand_groups(C.to(dtype), Dm)
    decay = torch.exp(dt[:, :, None, :] * Af[None, :, :, None])          # (B, D, N, L)
    drive = (dt * uf)[:, :, None, :] * Bf                                 # (B, D, N, L)
    h = torch.zeros(Bsz, Dm, N, dtype=dtype)
    y = torch.empty(Bsz, Dm, L, dtype=dtype)
    for l in range(L):
        h = decay[..., l] * h + drive[..., l]
        y[..., l] = (h * Cf[..., l]).sum(-1)
    if D is not None:
        y = y + uf * D.to(dtype)[None, :, None]
    out = y.to(in_dtype)
    return (out, h) if return_last_state else out


def selective_scan_bwd(u, delta, A, B, C, D, delta_bias, delta_softplus, dout, dtype=torch.float64):
    """Closed-form gradients of ``selective_scan`` (sequential adjoint recurrence).

    Returns (du, ddelta, dA, dB, dC, dD, ddelta_bias) in ``dtype``; dD / ddelta_bias are None when the
    corresponding input is None.  Follows the quantities of bwd_kernel.cuh:198-207,259-268.
    """
    if B.dim() == 3:
        B = B[:, None]
    if C.dim() == 3:
        C = C[:, None]
    Bsz, Dm, L = u.shape
    G = B.shape[1]
    N = A.shape[1]
    per = Dm // G
    uf = u.to(dtype)
    dy = dout.to(dtype)
    pre = delta.to(dtype)
    if delta_bias is not None:
        pre = pre + delta_bias.to(dtype)[None, :, None]
    dt = F.softplus(pre) if delta_softplus else pre
    Af = A.to(dtype)
    Bf = _expand_groups(B.to(dtype), Dm)
    Cf = _expand_groups(C.to(dtype), Dm)
    decay = torch.exp(dt[:, :, None, :] * Af[None, :, :, None])
    drive = (dt * uf)[:, :, None, :] * Bf
    # forward states
    hs = torch.empty(Bsz, Dm, N, L, dtype=dtype)
    h = torch.zeros(Bsz, Dm, N, dtype=dtype)
    for l in range(L):
        h = decay[..., l] * h + drive[..., l]
        hs[..., l] = h
    # adjoint states g[l] = C[l]*dy[l] + decay[l+1]*g[l+1]
    gs = torch.empty_like(hs)
    g = torch.zeros(Bsz, Dm, N, dtype=dtype)
    for l in range(L - 1, -1, -1):
        g = Cf[..., l] * dy[:, :, None, l] + g
        gs[..., l] = g
        g = decay[..., l] * g
    carried = hs - drive                      # decay[l] * h[l-1]
    du = (gs * Bf).sum(2) * dt
    if D is not None:
        du = du + dy * D.to(dtype)[None, :, None]
    ddt = (gs * (Bf * uf[:, :, None, :] + Af[None, :, :, None] * carried)).sum(2)
    dA = (gs * carried * dt[:, :, None, :]).sum(dim=(0, 3))
    dB_full = gs * (dt * uf)[:, :, None, :]
    dC_full = hs * dy[:, :, None, :]
    dB = dB_full.reshape(Bsz, G, per, N, L).sum(2)
    dC = dC_full.reshape(Bsz, G, per, N, L).sum(2)
    dD = (dy * uf).sum(dim=(0, 2)) if D is not None else None
    if delta_softplus:
        # d softplus(x)/dx = sigmoid(x) (identity branch above the threshold 20)
        ddelta = torch.where(pre <= 20.0, ddt * torch.sigmoid(pre), ddt)
    else:
        ddelta = ddt
    dbias = ddelta.sum(dim=(0, 2)) if delta_bias is not None else None
    return du, ddelta, dA, dB, dC, dD, dbias


# ----------------------------------------------------------------------------------------------
# the SS2D core chain the way SS2D.forward_corev2 strings it together (vmamba.py:1472-1497)
# ----------------------------------------------------------------------------------------------
def ss2d_core(x, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds, dtype=torch.float32, x_proj_bias=None):
    """x: (B, C, H, W) -> y: (B, C, H*W).  CrossScan -> two einsums -> scan -> CrossMerge (x_proj_bias: vmamba.py:1474-1475)."""
    Bsz, C, H, W = x.shape
    K, _, R = dt_projs_weight.shape
    N = A_logs.shape[1]
    L = H * W
    xs = cross_scan(x)
    x_dbl = torch.einsum("bkdl,kcd->bkcl", xs, x_proj_weight)
    if x_proj_bias is not None:
        x_dbl = x_dbl + x_proj_bias.view(1, K, -1, 1)
    dts, Bs, Cs = torch.split(x_dbl, [R, N, N], dim=2)
    dts = torch.einsum("bkrl,kdr->bkdl", dts, dt_projs_weight)
    As = -torch.exp(A_logs.float())
    ys = selective_scan(
        xs.reshape(Bsz, K * C, L).float(), dts.reshape(Bsz, K * C, L).float(), As,
        Bs.contiguous().float(), Cs.contiguous().float(), Ds.float(),
        dt_projs_bias.reshape(-1).float(), True, dtype=dtype,
    )
    return cross_merge(ys.reshape(Bsz, K, C, H, W))


def dwconv_silu(x_cl, weight, bias):
    """Head of the block for the configs' layout (forwardv2, vmamba.py:1541-1546): channel-last (B, H, W, C) ->
    permute -> depthwise conv 3x3 (padding 1, vmamba.py:860-868) -> SiLU -> (B, C, H, W)."""
    x = x_cl.permute(0, 3, 1, 2).contiguous()                                      # :1541-1542
    x = F.conv2d(x, weight, bias, padding=1, groups=x.shape[1])                    # :1543-1544
    return F.silu(x)                                                               # :1545


def out_norm_gate(y, gamma, beta, z, H, W, eps=1e-5, z_silu=True, out_dtype=None):
    """Tail of the block for the configs' layout: forward_corev2's out_norm branch (vmamba.py:1525-1531; channel_first False,
    out_norm_shape "v0", out_norm = nn.LayerNorm(d_inner)) followed by the gate of forwardv2 (:1536-1550).
    y (B, C, L) merged map -> (B, H, W, C); z (B, H, W, C) or None."""
    Bsz, C, L = y.shape
    y = y.transpose(dim0=1, dim1=2).contiguous()                         # :1527
    y = F.layer_norm(y, (C,), gamma, beta, eps).view(Bsz, H, W, -1)      # :1528
    if out_dtype is not None:
        y = y.to(out_dtype)                                              # :1531
    if z is not None:
        if z_silu:
            z = F.silu(z)                                                # :1538-1539
        y = y * z                                                        # :1549-1550
    return y


def ss2d_core_storage_order(x, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds, dtype=torch.float32):
    """The SS2D core computed the way the planned fused kernel will, as a check of that plan's
    algebra: no `xs` / `ys` copies.  Directions 0/2 work on the map in row-major STORAGE order, directions 1/3 on one
    transposed copy (column-major storage order); `delta`, `B`, `C` come from un-replicated einsums over the storage order;
    the time-reversed directions (2, 3) read every positional tensor of their group back to front and write their outputs
    back to front; outputs are summed per storage order (y0 + y2, y1 + y3) and the column-major plane is transposed back once.
    Must equal `ss2d_core` (the reference's chain, vmamba.py:1472-1497) up to summation order of the einsums."""
    Bsz, C, H, W = x.shape
    K, _, R = dt_projs_weight.shape
    N = A_logs.shape[1]
    L = H * W
    src = [x.reshape(Bsz, C, L), x.transpose(2, 3).reshape(Bsz, C, L)]   # row-major map, column-major map (one copy)
    As = -torch.exp(A_logs.float()).reshape(K, C, N)
    Dk = Ds.float().reshape(K, C)
    bias = dt_projs_bias.float().reshape(K, C)
    planes = [None, None]
    for k in range(K):
        s = src[k & 1]                                                     # storage order of this direction's pair
        x_dbl = torch.einsum("bdl,cd->bcl", s, x_proj_weight[k])           # (B, R + 2N, L), storage order
        dts, Bs, Cs = torch.split(x_dbl, [R, N, N], dim=1)
        dts = torch.einsum("brl,dr->bdl", dts, dt_projs_weight[k])         # (B, C, L), storage order
        rev = k >= 2
        tf = (lambda t: t.flip(-1)) if rev else (lambda t: t)              # reversed directions: back to front
        out = selective_scan(tf(s).float(), tf(dts).float(), As[k], tf(Bs).unsqueeze(1).float().contiguous(),
                             tf(Cs).unsqueeze(1).float().contiguous(), Dk[k], bias[k], True, dtype=dtype)
        out = tf(out)                                                      # written back to front: storage order again
        planes[k & 1] = out if planes[k & 1] is None else planes[k & 1] + out   # y0 + y2, y1 + y3
    y_cm = planes[1].reshape(Bsz, C, W, H).transpose(2, 3).reshape(Bsz, C, L)
    return planes[0] + y_cm


def scan_algorithmic_bytes(Bsz, Dm, L, G, N, itemsize=4):
    """bytes the scan must move, forward and backward."""
    fwd = itemsize * (3 * Bsz * Dm * L + 2 * Bsz * G * N * L)
    bwd = itemsize * (5 * Bsz * Dm * L + 4 * Bsz * G * N * L)
    return fwd, bwd
