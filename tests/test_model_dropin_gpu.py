"""GPU: model-level drop-in.  The reference's own ``SS2D`` module (model/vmamba.py:544-1552, staged
unmodified under oracle/_ref/py by oracle/stage_ref_py.py) is built twice from the same seed:

  * REFERENCE: forward_type "v2" -- its pure-PyTorch ``CrossScan`` / ``CrossMerge`` (vmamba.py:27-73) and its own CUDA
    selective scan (``selective_scan_cuda_core`` rebuilt for sm_100a under oracle/_ref; when that extension is not there, the
    oracle's pure-PyTorch ``selective_scan_ref`` restatement);
  * OURS: after ``vm_asr_b200.integration.install``, forward_type "v5" (what the configs select),
    once with the chain of three operators and once with the fused core.

``SS2D.forwardv2`` (in_proj, depthwise conv, SiLU, core, out_norm, gate, out_proj: vmamba.py:1533-1552) and its backward must
agree: output, input gradient and every parameter gradient."""
import copy

import numpy as np
import pytest
import torch

from oracle import stage_ref_py

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not stage_ref_py.available(), reason="reference sources not staged (oracle/stage_ref_py.py)")]


def _rel(a, b):
    a, b = a.detach().double().cpu().numpy(), b.detach().double().cpu().numpy()
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def _reference_module(vm, d_model, seed):
    torch.manual_seed(seed)
    has_ext = "selective_scan_cuda_core" in vm.__dict__ or hasattr(vm, "selective_scan_cuda_core")
    if not has_ext:
        # no rebuilt extension on this box: bind the reference's autograd wrapper to the oracle's restatement of selective_scan_ref
        from oracle import ss2d_ref

        class _RefScan(torch.autograd.Function):
            @staticmethod
            def forward(ctx, u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, nrows=1, backnrows=1, oflex=True):
                with torch.enable_grad():
                    leaves = [t.detach().requires_grad_() for t in (u, delta, A, B, C, D, delta_bias)]
                    out = ss2d_ref.selective_scan(*leaves, delta_softplus)
                ctx.leaves, ctx.out = leaves, out
                return out.detach()

            @staticmethod
            def backward(ctx, dout):
                grads = torch.autograd.grad(ctx.out, ctx.leaves, dout)
                return (*grads, None, None, None, None)

        vm.SelectiveScanCore = _RefScan
    return vm.SS2D(d_model=d_model, d_state=1, ssm_ratio=2.0, dt_rank="auto", d_conv=3, conv_bias=True, forward_type="v2").cuda()


@pytest.mark.parametrize("d_model,H,W,fused", [(16, 32, 32, False), (16, 32, 32, True), (8, 64, 48, True), (32, 16, 16, True)])
def test_ss2d_module_drop_in(d_model, H, W, fused):
    from vm_asr_b200 import integration
    vm_ref = stage_ref_py.load("vmamba")
    ref = _reference_module(vm_ref, d_model, seed=0)
    vm_ours = integration.install(stage_ref_py.load("vmamba"), fused=fused)
    torch.manual_seed(0)
    ours = vm_ours.SS2D(d_model=d_model, d_state=1, ssm_ratio=2.0, dt_rank="auto", d_conv=3, conv_bias=True, forward_type="v5").cuda()
    ours.load_state_dict(copy.deepcopy(ref.state_dict()))          # state-dict keys are untouched by the rebinding
    assert [k for k, _ in ours.named_parameters()] == [k for k, _ in ref.named_parameters()]

    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, H, W, d_model, generator=g)                 # channel-last, as VSSBlock feeds it (vmamba.py:1826-1837)
    gy = torch.randn(2, H, W, d_model, generator=g).cuda()
    xr, xo = x.cuda().requires_grad_(), x.cuda().requires_grad_()
    yr, yo = ref(xr), ours(xo)
    assert yo.shape == yr.shape
    assert _rel(yo, yr) < 2e-4
    yr.backward(gy)
    yo.backward(gy)
    assert _rel(xo.grad, xr.grad) < 5e-4
    for (name, po), (_, pr) in zip(ours.named_parameters(), ref.named_parameters()):
        assert po.grad is not None, name
        assert _rel(po.grad, pr.grad) < 2e-3, (name, _rel(po.grad, pr.grad))
