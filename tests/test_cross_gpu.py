"""GPU parity: cross scan / cross merge must be BIT-EXACT with the reference's PyTorch versions
(model/vmamba.py:27-73), forward and backward, for every dtype and for odd / non-square maps."""
import os

import numpy as np
import pytest
import torch

from oracle import c_ref, ss2d_ref

pytestmark = pytest.mark.gpu


def _ops():
    from vm_asr_b200 import cross
    return cross


@pytest.mark.parametrize("tag", ["sq", "rect", "odd"])
def test_golden_vectors(golden_dir, tag):
    cross = _ops()
    g = np.load(os.path.join(golden_dir, "cross_scan_merge.npz"))
    t = lambda k: torch.from_numpy(g[f"{tag}_{k}"]).cuda()
    x = t("x").requires_grad_()
    xs = cross.CrossScan.apply(x)
    assert torch.equal(xs, t("xs"))
    xs.backward(t("gxs"))
    assert torch.equal(x.grad, t("gx"))
    ys = t("ys").requires_grad_()
    y = cross.CrossMerge.apply(ys)
    assert torch.equal(y, t("y"))
    y.backward(t("gy"))
    assert torch.equal(ys.grad, t("gys"))


SHAPES = [
    (2, 3, 56, 57),      # the odd shape of CHECKS.check_csm_triton (vmamba.py:2560)
    (1, 2, 64, 64), (2, 4, 128, 32), (1, 3, 20, 132), (2, 2, 4, 4), (1, 1, 1, 7), (1, 2, 9, 1),
    (4, 2, 512, 512), (8, 2, 1024, 512), (4, 16, 256, 256), (4, 256, 16, 16), (8, 64, 128, 64),
    # the VSSM32 maps (DIMS 32, batch 8) and the two small-map families that take several channels per CTA
    (8, 32, 256, 256), (8, 64, 128, 128), (8, 128, 64, 64), (8, 256, 32, 32), (8, 512, 16, 16), (4, 128, 32, 32),
    (3, 7, 24, 32), (2, 5, 8, 16),   # vector path with a partial plane group / tile
    (70000, 1, 4, 4),                # more planes than the old gridDim.z limit (65535)
    (1, 889, 8, 8), (1, 1779, 4, 4),  # small maps, two / four planes per CTA (enough CTAs for three per SM) with a partial last group
]


@pytest.mark.parametrize("B,C,H,W", SHAPES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_bit_exact_vs_oracle(B, C, H, W, dtype):
    cross = _ops()
    torch.manual_seed(H * 1000 + W)
    x = torch.randn(B, C, H, W).to(dtype)
    ys = torch.randn(B, 4, C, H, W).to(dtype)
    xs_ref = ss2d_ref.cross_scan(x)
    y_ref = ss2d_ref.cross_merge(ys)
    assert torch.equal(cross.cross_scan(x.cuda()).cpu(), xs_ref)
    assert torch.equal(cross.cross_merge(ys.cuda(), H, W).cpu(), y_ref)
    if dtype == torch.float32 and B * C * H * W < 1 << 22:
        assert np.array_equal(c_ref.cross_scan(x.numpy()), xs_ref.numpy())


def test_merge_of_scan_is_4x_and_linearity_full_size():
    """Size-independent properties at the largest config shape: merge(scan(x)) == 4x exactly."""
    cross = _ops()
    x = torch.randn(8, 2, 1024, 512, device="cuda")
    xs = cross.cross_scan(x)
    y = cross.cross_merge(xs, 1024, 512)
    assert torch.equal(y.view_as(x), 4 * x)
    # each direction is a permutation: sorted values agree
    assert torch.equal(xs[:, 3].sort(-1).values, x.flatten(2).sort(-1).values)


def test_non_contiguous_input_is_made_contiguous():
    cross = _ops()
    x = torch.randn(2, 8, 32, 48, device="cuda")[:, ::2]
    assert torch.equal(cross.cross_scan(x).cpu(), ss2d_ref.cross_scan(x.cpu().contiguous()))


def test_errors():
    cross = _ops()
    with pytest.raises(RuntimeError):
        cross.cross_scan(torch.randn(1, 2, 4, 4))           # CPU tensor
    with pytest.raises(RuntimeError):
        cross.cross_scan(torch.randn(1, 2, 4, 4, device="cuda").double())


def test_matches_the_references_triton_kernels():
    """CrossScanTriton / CrossMergeTriton (model/csm_triton.py:311-366, staged by oracle/stage_ref_py.py) on the same device:
    the scan is a permutation, so identical; the merge differs only by the association of its four-term sum."""
    from oracle import stage_ref_py
    if not stage_ref_py.available():
        pytest.skip("reference sources not staged")
    stage_ref_py.load("vmamba")
    import csm_triton
    from vm_asr_b200 import cross
    for shape, dt in (((2, 8, 64, 64), torch.float32), ((2, 5, 56, 57), torch.float32), ((1, 16, 128, 64), torch.float16), ((4, 64, 16, 16), torch.bfloat16)):
        x = torch.randn(*shape, device="cuda").to(dt)
        a, b = cross.cross_scan(x), csm_triton.CrossScanTriton.apply(x)
        assert torch.equal(a, b.view_as(a))
        ys = torch.randn(shape[0], 4, *shape[1:], device="cuda").to(dt)
        m0, m1 = cross.cross_merge(ys, shape[2], shape[3]), csm_triton.CrossMergeTriton.apply(ys)
        tol = {torch.float32: 1e-5, torch.float16: 2e-2, torch.bfloat16: 1.3e-1}[dt]   # a couple of ulps of a sum of four N(0, 1) values
        assert (m0.float() - m1.view_as(m0).float()).abs().max().item() < tol


@pytest.mark.parametrize("B,C,H,W", [(2, 3, 8, 8), (1, 5, 56, 57), (2, 4, 64, 32), (1, 2, 33, 100)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_one_by_one_variant(B, C, H, W, dtype):
    """CrossScanTriton1b1 (csm_triton.py:369-395): every direction walks its OWN map; defined here by the pure-PyTorch
    CrossScan applied to each of the four maps (direction k of CrossScan(x_k)); the backward is the inverse permutation."""
    cross = _ops()
    x = torch.randn(B, 4, C, H, W, generator=torch.Generator().manual_seed(H + W)).to(dtype)
    ref = torch.stack([ss2d_ref.cross_scan(x[:, k])[:, k] for k in range(4)], dim=1)
    xg = x.cuda().requires_grad_()
    y = cross.CrossScanTriton1b1.apply(xg)
    assert y.shape == (B, 4, C, H * W) and torch.equal(y.detach().cpu(), ref)
    gy = torch.randn(B, 4, C, H * W, generator=torch.Generator().manual_seed(1)).to(dtype)
    y.backward(gy.cuda())
    xs = [x[:, k].float().requires_grad_() for k in range(4)]
    gref = torch.autograd.grad(torch.stack([ss2d_ref.cross_scan(xr)[:, k] for k, xr in enumerate(xs)], dim=1), xs, gy.float())
    assert torch.equal(xg.grad.float().cpu(), torch.stack(gref, dim=1))
