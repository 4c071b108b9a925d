"""GPU: the step harness (vm_asr_b200/harness.py) -- STFT -> pairs of fused SS2D cores -> iSTFT under autograd, flat gradient
buffer, fused AdamW -- on a small workload: every parameter receives a finite gradient through the library's kernels, the
paired (grouped) and unpaired paths agree, a few steps reduce the loss, inference runs without autograd."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _small_workload():
    from vm_asr_b200.workload import SS2DCall, Workload
    calls = ([SS2DCall(8, 32, 16)] * 4 + [SS2DCall(16, 16, 8)] * 4 + [SS2DCall(32, 8, 4)] * 2 + [SS2DCall(16, 16, 8)] * 2
             + [SS2DCall(4, 64, 32)] * 2 + [SS2DCall(2, 128, 64)] * 2)
    return Workload("tiny", "(test)", 2, 64 * 63, 256, 64, 256, 16000, tuple(calls))


def test_train_step_runs_and_learns():
    """fp32 step: finite gradients for every parameter, all living in the flat buffer, and the loss goes down.  The toy net is
    cubic in its activations (harness.py), so the LEARNING check runs without autocast; the autocast + GradScaler plumbing has
    its own short test below."""
    from vm_asr_b200 import harness
    wl = _small_workload()
    dev = torch.device("cuda")
    ts = harness.TrainStep(wl, dev, world=1, lr=1e-3, amp=False)
    x, y = harness.synthetic_batch(wl, dev)
    losses = [ts(x, y).item() for _ in range(30)]
    assert all(l == l and l < 1e6 for l in losses), losses
    assert min(losses[8:]) < losses[0] - 1e-3, losses
    for name, p in ts.net.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), name
        assert p.grad.data_ptr() >= ts.grads.flat.data_ptr()
    nz = sum(int(p.grad.abs().sum() > 0) for p in ts.net.parameters())
    assert nz >= 0.9 * len(list(ts.net.parameters()))
    out = ts.infer(x)
    assert out.shape == x.shape and not out.requires_grad


def test_train_step_under_autocast():
    """fp16 autocast + GradScaler as the reference trains (config.py:217, trainer/trainer.py:106-107): the first steps give finite
    losses and finite (unscaled) gradients, and the scaler keeps a usable scale."""
    from vm_asr_b200 import harness
    wl = _small_workload()
    dev = torch.device("cuda")
    ts = harness.TrainStep(wl, dev, world=1, lr=2e-4, amp=True)
    x, y = harness.synthetic_batch(wl, dev)
    losses = [ts(x, y).item() for _ in range(6)]
    assert all(l == l and l < 1e6 for l in losses), losses
    assert ts.scaler.get_scale() >= 1024.0 / 8, ts.scaler.get_scale()
    out = ts.infer(x)
    assert out.shape == x.shape and torch.isfinite(out).all()


def test_paired_and_unpaired_harness_agree():
    from vm_asr_b200 import harness
    wl = _small_workload()
    dev = torch.device("cuda")
    a = harness.HotPathNet(wl, pair=True).to(dev)
    b = harness.HotPathNet(wl, pair=False).to(dev)
    b.load_state_dict(a.state_dict())
    x, _ = harness.synthetic_batch(wl, dev)
    ya, yb = a(x), b(x)
    assert torch.allclose(ya, yb, rtol=1e-5, atol=1e-6)
    ya.square().mean().backward()
    yb.square().mean().backward()
    for (n, p), q in zip(a.named_parameters(), b.parameters()):
        if p.grad is None:   # block parameters of a call whose map takes the bare-core path (W % 8 != 0)
            assert q.grad is None, n
            continue
        assert torch.allclose(p.grad, q.grad, rtol=2e-3, atol=1e-6), n


def test_block_mode_runs_whole_ss2d_bodies():
    """block=True (the default): every call whose map qualifies runs in_proj -> head kernel -> fused core -> tail kernel -> out_proj
    (ss2d_block_core_pair); paired and unpaired agree, every block parameter (conv, LayerNorm, projections) gets a finite
    gradient; block=False keeps the bare cores between normalisations"""
    from vm_asr_b200 import harness
    wl = _small_workload()
    dev = torch.device("cuda")
    a = harness.HotPathNet(wl, pair=True, block=True).to(dev)
    b = harness.HotPathNet(wl, pair=False, block=True).to(dev)
    b.load_state_dict(a.state_dict())
    names = [n for n, _ in a.named_parameters()]
    assert any("conv_weight" in n for n in names) and any("in_proj" in n for n in names) and any("norm_weight" in n for n in names)
    x, _ = harness.synthetic_batch(wl, dev)
    ya, yb = a(x), b(x)
    assert torch.isfinite(ya).all() and torch.allclose(ya, yb, rtol=1e-5, atol=1e-6)
    ya.square().mean().backward()
    yb.square().mean().backward()
    used = 0
    for (n, p), q in zip(a.named_parameters(), b.parameters()):
        if p.grad is None:   # block parameters of a call whose map takes the bare-core path (W % 8 != 0)
            assert q.grad is None and ".core." not in n, n
            continue
        assert torch.isfinite(p.grad).all(), n
        assert torch.allclose(p.grad, q.grad, rtol=2e-3, atol=1e-6), n
        used += int("conv_weight" in n and p.grad.abs().sum() > 0)
    assert used >= 2
    c = harness.HotPathNet(wl, block=False).to(dev)
    assert not any("conv_weight" in n for n, _ in c.named_parameters())
    assert torch.isfinite(c(x)).all()


def test_graph_captured_step_matches_eager():
    """TrainStep.capture(): forward + backward as one CUDA graph gives the same losses as the eager step."""
    from vm_asr_b200 import harness
    wl = _small_workload()
    dev = torch.device("cuda")
    x, y = harness.synthetic_batch(wl, dev)
    torch.manual_seed(0)
    a = harness.TrainStep(wl, dev, world=1)
    b = harness.TrainStep(wl, dev, world=1)
    b.net.load_state_dict(a.net.state_dict())
    b.capture(x, y)
    la = [a(x, y).item() for _ in range(4)]
    lb = [b(x, y).item() for _ in range(4)]
    for u, v in zip(la, lb):
        assert abs(u - v) < 1e-4 * max(1.0, abs(u)), (la, lb)


def test_batched_long_clip_inference_matches_segment_by_segment():
    """segment.infer_long: all segments of all clips as ONE batch through the harness generator, cross-faded on the device,
    equals the reference's procedure (one segment at a time, tester.py:106-131; the oracle's fold) -- and reports an RTF from
    CUDA events."""
    from oracle import segment_ref
    from vm_asr_b200 import harness, segment
    wl = _small_workload()
    dev = torch.device("cuda")
    net = harness.HotPathNet(wl).to(dev).eval()
    T, ov = wl.T, 200
    long_clip = 0.1 * torch.randn(2, 1, 2 * (T - ov) + T + 37, generator=torch.Generator().manual_seed(3)).to(dev)
    out, info = segment.infer_long(net, long_clip, None, segment_length=T, overlap=ov, sample_rate=wl.sr)
    assert out.shape == long_clip.shape and info["segments"] == 2 * 3 and info["rtf"] > 0
    segs = segment_ref.unfold_audio(long_clip, T, ov)
    done = torch.zeros_like(segs)
    with torch.no_grad():
        for i in range(segs.shape[2]):
            done[:, :, i] = net(segs[:, :, i].contiguous())
    ref = segment_ref.fold_audio(done.cpu(), long_clip.shape[-1], T, ov)
    assert torch.allclose(out.cpu(), ref, rtol=1e-4, atol=1e-5)
    short, info1 = segment.infer_long(net, long_clip[..., :T].contiguous(), None, segment_length=T, overlap=ov, sample_rate=wl.sr)
    assert short.shape[-1] == T and info1["segments"] == 2
