"""CPU: the batched segment cut / cross-fade of vm_asr_b200/segment.py against the oracle's restatement of
utils/post_processing.py and against values computed by the reference's own functions (tests/golden/segments.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import segment_ref


@pytest.mark.parametrize("B,C,T,seg,ov", [(1, 1, 81760, 40880, 2000), (2, 1, 30000, 8000, 500), (3, 2, 1000, 300, 149),
                                          (1, 1, 5000, 1000, 600), (2, 1, 777, 256, 255 // 2), (1, 1, 400, 400, 10)])
def test_fold_matches_the_reference_loop(B, C, T, seg, ov):
    from vm_asr_b200 import segment
    x = torch.randn(B, C, T, dtype=torch.float64, generator=torch.Generator().manual_seed(T))
    segs = segment.unfold_audio(x, seg, ov)
    assert torch.equal(segs, segment_ref.unfold_audio(x, seg, ov))
    work = segs * 1.5 + 0.25                     # "processed" segments
    got = segment.fold_audio(work, T, seg, ov)
    ref = segment_ref.fold_audio(work, T, seg, ov)
    assert got.shape == ref.shape
    assert torch.allclose(got, ref, rtol=0, atol=1e-12)


def test_golden_from_the_reference(golden_dir):
    from vm_asr_b200 import segment
    g = np.load(os.path.join(golden_dir, "segments.npz"))
    seg, ov = (int(v) for v in g["params"])
    audio = torch.from_numpy(g["audio"])
    segs = segment.unfold_audio(audio, seg, ov)
    assert np.array_equal(segs.numpy(), g["segments"])
    out = segment.fold_audio(torch.from_numpy(g["processed"]), audio.shape[-1], seg, ov)
    assert np.abs(out.numpy() - g["folded"]).max() < 1e-6


def test_overlap_must_be_smaller_than_segment():
    from vm_asr_b200 import segment
    with pytest.raises(ValueError):
        segment.fold_audio(torch.zeros(1, 1, 2, 10), 20, 10, 10)
