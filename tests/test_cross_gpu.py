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
