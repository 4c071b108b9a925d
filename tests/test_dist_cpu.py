"""N > 1 host logic on CPU: world_size-2 gloo processes (no GPU).  The path shards by batch with no exchange step
(SURVEY.md 8e); the only collective is the gradient sum of the parameters the shards share.  Checked here with the
oracle standing in for the kernels: shard -> per-rank fwd/bwd -> all-reduce == the full-batch result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vm_asr_b200 import dist as vdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_range_covers_batch_exactly():
    for n in (0, 1, 4, 7, 8, 32, 33):
        for world in (1, 2, 3, 4, 8):
            spans = [vdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    # the configs' weak scaling: world * B_local clips -> [r * B_local, (r + 1) * B_local)
    assert vdist.shard_range(8 * 4, 3, 8) == (12, 16)
    with pytest.raises(ValueError):
        vdist.shard_range(4, 2, 2)


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import ss2d_ref
        torch.manual_seed(0)  # same full batch on every rank, each takes its shard
        Bsz, D, L, G = 4, 8, 96, 4
        u, delta = torch.randn(Bsz, D, L), 0.5 * torch.rand(Bsz, D, L)
        A, Dv, bias = -0.5 * torch.rand(D, 1), torch.randn(D), 0.5 * torch.rand(D)
        Bm, Cm = torch.randn(Bsz, G, 1, L), torch.randn(Bsz, G, 1, L)
        dout = torch.randn(Bsz, D, L)
        full = ss2d_ref.selective_scan_bwd(u, delta, A, Bm, Cm, Dv, bias, True, dout, dtype=torch.float64)
        sh = lambda t: vdist.shard_batch(t, rank, world)
        mine = ss2d_ref.selective_scan_bwd(sh(u), sh(delta), A, sh(Bm), sh(Cm), Dv, bias, True, sh(dout), dtype=torch.float64)
        # per-clip gradients (du, ddelta, dB, dC) are this rank's rows of the full result: no collective
        lo, hi = vdist.shard_range(Bsz, rank, world)
        for i in (0, 1, 3, 4):
            np.testing.assert_allclose(mine[i].numpy(), full[i][lo:hi].numpy(), rtol=1e-12, atol=1e-12)
        # shared-parameter gradients (dA, dD, ddelta_bias) need the sum over ranks
        shared = [mine[2].clone().float(), mine[5].clone().float(), mine[6].clone().float()]
        vdist.allreduce_grads_(shared)
        for got, ref in zip(shared, (full[2], full[5], full[6])):
            np.testing.assert_allclose(got.numpy(), ref.float().numpy(), rtol=1e-5, atol=1e-5)
        # mean variant and the timing reduction
        ones = [torch.full((3,), float(rank + 1))]
        vdist.allreduce_grads_(ones, mean=True)
        assert torch.allclose(ones[0], torch.full((3,), sum(range(1, world + 1)) / world))
        assert vdist.max_over_ranks(10.0 + rank) == 10.0 + world - 1
        assert vdist.whole_job_throughput([5.0] * world, 2.0) == 2.5 * world
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_batch_shards_plus_allreduce_equal_full_batch(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_single_process_is_a_no_op():
    g = [torch.ones(4)]
    vdist.allreduce_grads_(g)
    assert torch.equal(g[0], torch.ones(4))
    assert vdist.max_over_ranks(3.5) == 3.5


def _flat_worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vm_asr_b200.harness import FlatGrads
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4))
        unused = torch.nn.Parameter(torch.ones(7))   # never receives a gradient (cf. the reference's unused phase decoder)
        params = list(net.parameters()) + [unused]
        fg = FlatGrads(params, bucket_floats=100, payload_floats=1000, payload_chunks=3)
        assert len(fg.buckets) >= 3 and fg.enabled
        x_all = torch.randn(8, 16)
        for _ in range(2):   # two steps: the hooks re-arm
            fg.zero()
            fg.start_payload()
            lo, hi = vdist.shard_range(8, rank, world)
            net(x_all[lo:hi]).square().sum().backward()
            fg.finish(world)
        ref = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4))
        ref.load_state_dict(net.state_dict())
        (ref(x_all).square().sum() / world).backward()   # mean over ranks of the per-shard sums
        for p, q in zip(net.parameters(), ref.parameters()):
            np.testing.assert_allclose(p.grad.numpy(), q.grad.numpy(), rtol=1e-5, atol=1e-6)
        assert torch.equal(unused.grad, torch.zeros(7))
        assert all(p.grad.data_ptr() >= fg.flat.data_ptr() for p in params)   # gradients are views of the flat buffer
        open(os.path.join(tmp, f"flat{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_flat_bucketed_gradient_allreduce(tmp_path):
    """The harness's gradient plumbing (one flat buffer, buckets reduced from autograd hooks as they fill, a payload going out
    first, a bucket whose parameter got no gradient) on two gloo ranks: equal to the full-batch gradient divided by world."""
    world, port = 2, _free_port()
    mp.spawn(_flat_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"flat{r}").exists() for r in range(world))
