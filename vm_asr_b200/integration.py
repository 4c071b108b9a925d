"""Reference-side rebinding as code: ``install(vmamba_module, ...)`` makes the reference's
``SS2D`` / ``VSSBlock`` run on this library without touching its sources.

* ``fused=False``: rebind the operator names the forward-type table captures when a module is CONSTRUCTED
  (``SelectiveScanCore``, ``CrossScan`` / ``CrossMerge`` and their ``*Triton`` variants: model/vmamba.py:772-850), i.e. the
  chain of three operators, statement for statement ``forward_corev2``.
* ``fused=True`` (default): additionally replace ``SS2D.forward_corev2`` (vmamba.py:1377-1531) by a version whose
  CrossScan -> einsums -> selective scan -> CrossMerge part is ``vm_asr_b200.ss2d.ss2d_core`` (the fused core; maps it does
  not take fall back to the chain inside ``ss2d_core``), followed by the same ``out_norm`` tail (vmamba.py:1514-1531); and,
  with ``fused_tail`` (default), ``SS2D.forwardv2`` (vmamba.py:1533-1552) by a version that hands the gate ``z`` to the core so
  that merge, LayerNorm, cast, SiLU(z) and the product are one kernel (``ss2d.ss2d_core_out``).

Call it BEFORE building the model (the table binds names at construction).  Parameters, state-dict keys and module
structure are untouched."""
from __future__ import annotations

from . import cross, scan, ss2d, stft


def _fused_forward_corev2(self, x=None, x_proj_weight=None, x_proj_bias=None, dt_projs_weight=None, dt_projs_bias=None,
                          A_logs=None, Ds=None, delta_softplus=True, out_norm=None, out_norm_shape="v0", channel_first=False,
                          to_dtype=True, force_fp32=False, **kwargs):
    """Drop-in for SS2D.forward_corev2; ``SelectiveScan`` / ``CrossScan`` / ``CrossMerge`` / ``nrows`` / ``no_einsum`` keyword
    arguments of the forward-type table are accepted and ignored (the fused core replaces all three operators)."""
    out_norm = getattr(self, "out_norm", None)
    out_norm_shape = getattr(self, "out_norm_shape", "v0")
    B, D, H, W = x.shape
    y = ss2d.ss2d_core(x, self.x_proj_weight, self.dt_projs_weight, self.dt_projs_bias, self.A_logs, self.Ds,
                       delta_softplus=delta_softplus, force_fp32=True, x_proj_bias=getattr(self, "x_proj_bias", None))
    if self.channel_first:                                     # vmamba.py:1514-1521
        y = y.view(B, -1, H, W)
        if out_norm_shape in ["v1"]:
            y = out_norm(y)
        else:
            y = out_norm(y.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
        return y.to(x.dtype) if to_dtype else y
    if out_norm_shape in ["v1"]:                               # vmamba.py:1523-1524
        y = out_norm(y.view(B, -1, H, W)).permute(0, 2, 3, 1)
    else:                                                      # vmamba.py:1525-1529
        y = out_norm(y.transpose(dim0=1, dim1=2).contiguous()).view(B, H, W, -1)
    return y.to(x.dtype) if to_dtype else y


def _fused_forwardv2(self, x, **kwargs):
    """Drop-in for SS2D.forwardv2 (vmamba.py:1533-1552).  Where the fused tail applies -- channel-last layout, out_norm =
    nn.LayerNorm (what the configs build), a map ``ss2d.outnorm_fusable`` takes -- the core's planes go straight into the merge
    + LayerNorm + cast + SiLU(z) + gate kernel (``ss2d.ss2d_core_out``); otherwise statement for statement the reference."""
    import torch.nn as nn

    with_dconv = self.d_conv > 1
    x = self.in_proj(x)
    z = None
    if not self.disable_z:
        x, z = x.chunk(2, dim=(1 if self.channel_first else -1))
    core_is_v2 = getattr(self.forward_core, "func", None) is not None and self.forward_core.func.__name__ in ("forward_corev2", "_fused_forward_corev2")
    tail_ok = core_is_v2 and not self.channel_first and isinstance(self.out_norm, nn.LayerNorm) and self.out_norm.elementwise_affine
    # head too: depthwise 3x3 + SiLU read from the channel-last in_proj output in place, writing x and x^T for the core
    if (tail_ok and with_dconv and self.d_conv == 3 and isinstance(self.act, nn.SiLU) and x.is_cuda
            and ss2d.outnorm_fusable(x.permute(0, 3, 1, 2), self.A_logs.shape[1])):
        y = ss2d.ss2d_block_core(x, self.conv2d.weight, self.conv2d.bias, self.x_proj_weight, self.dt_projs_weight, self.dt_projs_bias,
                                 self.A_logs, self.Ds, self.out_norm.weight, self.out_norm.bias, z=z, z_silu=not self.disable_z_act,
                                 eps=self.out_norm.eps, x_proj_bias=getattr(self, "x_proj_bias", None))
        return self.dropout(self.out_proj(y))
    if not self.channel_first:
        x = x.permute(0, 3, 1, 2).contiguous()
    if with_dconv:
        x = self.conv2d(x)
    x = self.act(x)
    fuse_tail = tail_ok and ss2d.outnorm_fusable(x, self.A_logs.shape[1])
    if fuse_tail:
        y = ss2d.ss2d_core_out(x, self.x_proj_weight, self.dt_projs_weight, self.dt_projs_bias, self.A_logs, self.Ds,
                               self.out_norm.weight, self.out_norm.bias, z=z, z_silu=not self.disable_z_act,
                               eps=self.out_norm.eps, x_proj_bias=getattr(self, "x_proj_bias", None))
    else:
        if z is not None and not self.disable_z_act:
            z = self.act(z)
        y = self.forward_core(x)
        if z is not None:
            y = y * z
    return self.dropout(self.out_proj(y))


def install(vmamba_module, model_module=None, fused: bool = True, fused_tail: bool = True):
    """``vmamba_module``: the imported ``model.vmamba``; ``model_module``: the imported ``model.model`` (its ``wav2spectro`` /
    ``spectro2wav`` names are rebound when given)."""
    vmamba_module.SelectiveScanCore = scan.SelectiveScanCore          # vmamba.py:323
    vmamba_module.CrossScanTriton = cross.CrossScanTriton              # csm_triton.py:311 (forward_type v5)
    vmamba_module.CrossMergeTriton = cross.CrossMergeTriton            # csm_triton.py:340
    vmamba_module.CrossScan = cross.CrossScan                          # vmamba.py:27 (forward_type v2)
    vmamba_module.CrossMerge = cross.CrossMerge                        # vmamba.py:50
    if fused:
        vmamba_module.SS2D.forward_corev2 = _fused_forward_corev2      # vmamba.py:1377
        if fused_tail:
            vmamba_module.SS2D.forwardv2 = _fused_forwardv2            # vmamba.py:1533 (LayerNorm + gate tail fused into the merge)
    if model_module is not None:
        model_module.wav2spectro, model_module.spectro2wav = stft.wav2spectro, stft.spectro2wav   # utils/stft.py:22, 71
    return vmamba_module
