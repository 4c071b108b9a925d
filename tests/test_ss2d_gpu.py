"""GPU parity of the FUSED SS2D core (vmasr_ss2d_core_fwd / _bwd, vm_asr_b200.ss2d) -- SURVEY.md 8 row a9:
CrossScan -> selective scan -> CrossMerge of SS2D.forward_corev2 (model/vmamba.py:1472-1497) without the xs / ys copies.

Checked against (i) the oracle chain cross_scan -> C float64 scan -> cross_merge on every config map, outputs and all
gradients; (ii) the float64 torch oracle of the whole core (projections included) through autograd; (iii) the unfused chain
of this library's own operators; and, bit for bit, (iv) the merge association (y0 + y2) + transpose(y1 + y3) of
vmamba.py:55-60 and (v) the two map kernels."""
import numpy as np
import pytest
import torch

from oracle import c_ref, ss2d_ref

pytestmark = pytest.mark.gpu

REL_FP32 = 1e-4


def rel_err(got, ref):
    got = got.detach().double().cpu().numpy() if torch.is_tensor(got) else np.asarray(got, dtype=np.float64)
    ref = ref.detach().double().cpu().numpy() if torch.is_tensor(ref) else np.asarray(ref, dtype=np.float64)
    return np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30)


@pytest.mark.parametrize("shape", [(3, 5, 64, 64), (2, 7, 128, 36), (1, 3, 16, 16), (10, 300, 8, 12), (2, 2, 1024, 512)])
def test_map_transpose_and_merge2_bit_exact(shape):
    from vm_asr_b200 import ss2d
    torch.manual_seed(0)
    x = torch.randn(*shape, device="cuda")
    xT = ss2d.map_transpose(x)
    assert torch.equal(xT, x.transpose(-1, -2).contiguous())
    H, W = shape[-2:]
    q = torch.randn_like(xT)
    y = ss2d.map_merge2(x.flatten(-2), q.flatten(-2), H, W)
    assert torch.equal(y, (x + q.transpose(-1, -2)).flatten(-2))


def _time_to_memory_order(t):
    """(B, 4, C, L) per-direction tensors in TIME order (the reference's layout) -> the (rm, cm) pairs (B, 2, C, L) in
    memory order the fused core takes: directions 2 / 3 un-flipped."""
    rm = torch.stack([t[:, 0], t[:, 2].flip(-1)], dim=1).contiguous()
    cm = torch.stack([t[:, 1], t[:, 3].flip(-1)], dim=1).contiguous()
    return rm, cm


def _memory_to_time_order(rm, cm):
    return torch.stack([rm[:, 0], cm[:, 0], rm[:, 1].flip(-1), cm[:, 1].flip(-1)], dim=1)


def _scan_inputs(Bsz, C, H, W, seed=0):
    torch.manual_seed(seed)
    L = H * W
    x = torch.randn(Bsz, C, H, W)
    dts = 0.5 * torch.rand(Bsz, 4, C, L)
    Bs, Cs = torch.randn(Bsz, 4, 1, L), torch.randn(Bsz, 4, 1, L)
    As = -0.5 * torch.rand(4 * C, 1)
    Ds, bias = torch.randn(4 * C), 0.5 * torch.rand(4 * C)
    dy = torch.randn(Bsz, C, L)
    return x, dts, Bs, Cs, As, Ds, bias, dy


def _run_fused(x, dts, Bs, Cs, As, Ds, bias, dy=None):
    from vm_asr_b200 import ss2d
    dev = "cuda"
    xg = x.to(dev).requires_grad_(dy is not None)
    xT = ss2d.MapTranspose.apply(xg)
    leaves = []
    for t in (dts, Bs, Cs):
        rm, cm = _time_to_memory_order(t)
        leaves += [rm.to(dev).requires_grad_(dy is not None), cm.to(dev).requires_grad_(dy is not None)]
    dts_rm, dts_cm, Bs_rm, Bs_cm, Cs_rm, Cs_cm = leaves
    par = [t.to(dev).requires_grad_(dy is not None) for t in (As, Ds, bias)]
    y = ss2d._SS2DScan.apply(True, 1, xg, xT, dts_rm, dts_cm, Bs_rm, Bs_cm, Cs_rm, Cs_cm, *par)
    if dy is None:
        return y
    y.backward(dy.to(dev))
    torch.cuda.synchronize()
    g = lambda a, b: _memory_to_time_order(a.grad, b.grad)
    return y, dict(dx=xg.grad, ddelta=g(dts_rm, dts_cm), dB=g(Bs_rm, Bs_cm), dC=g(Cs_rm, Cs_cm), dA=par[0].grad, dD=par[1].grad,
                   dbias=par[2].grad)


def _run_oracle(x, dts, Bs, Cs, As, Ds, bias, dy):
    Bsz, C, H, W = x.shape
    L = H * W
    u = ss2d_ref.cross_scan(x).reshape(Bsz, 4 * C, L).numpy()
    delta = dts.reshape(Bsz, 4 * C, L).numpy()
    out, _, _ = c_ref.scan_fwd(u, delta, As.numpy(), Bs.numpy(), Cs.numpy(), Ds.numpy(), bias.numpy(), True)
    y = ss2d_ref.cross_merge(torch.from_numpy(out).reshape(Bsz, 4, C, H, W))
    dys = ss2d_ref.cross_merge_bwd(dy, H, W).reshape(Bsz, 4 * C, L).numpy()
    du, ddelta, dA, dB, dC, dD, dbias = c_ref.scan_bwd(u, delta, As.numpy(), Bs.numpy(), Cs.numpy(), Ds.numpy(), bias.numpy(), True, dys)
    dx = ss2d_ref.cross_scan_bwd(torch.from_numpy(du).reshape(Bsz, 4, C, L), H, W).reshape(Bsz, C, H, W)
    return y, dict(dx=dx, ddelta=ddelta.reshape(Bsz, 4, C, L), dB=dB, dC=dC, dA=dA, dD=dD, dbias=dbias)


# every SS2D map of the BASELINE.json configs at its full size, plus small / ragged-chunk / non-square ones
CONFIG_MAPS = [
    (4, 2, 512, 512), (4, 16, 256, 256), (4, 32, 128, 128), (4, 64, 64, 64), (4, 128, 32, 32), (4, 256, 16, 16),
    (8, 2, 1024, 512), (8, 16, 512, 256), (8, 32, 256, 128), (8, 64, 128, 64), (8, 128, 64, 32), (8, 256, 32, 16),
    (8, 32, 256, 256), (8, 64, 128, 128), (8, 128, 64, 64), (8, 256, 32, 32), (8, 512, 16, 16), (8, 2, 512, 512),
]
SMALL_MAPS = [(2, 3, 8, 12), (1, 5, 36, 60), (2, 6, 44, 52), (1, 4, 100, 84), (2, 1, 64, 40)]


@pytest.mark.parametrize("Bsz,C,H,W", SMALL_MAPS + CONFIG_MAPS)
def test_fused_core_against_oracle_chain(Bsz, C, H, W):
    """Outputs and every gradient of the fused core vs cross_scan -> float64 C scan -> cross_merge (and their adjoints)."""
    inp = _scan_inputs(Bsz, C, H, W)
    y, grads = _run_fused(*inp)
    y_ref, g_ref = _run_oracle(*inp)
    assert rel_err(y, y_ref) < REL_FP32
    for name in ("dx", "ddelta", "dB", "dC", "dA", "dD", "dbias"):
        # the three per-channel sums are fp32 atomic sums over batch x L terms that cancel (tests/test_scan_gpu.py: SUM_MAXNORM_TOL)
        tol = 3e-4 if name in ("dA", "dD", "dbias") else REL_FP32
        assert rel_err(grads[name], g_ref[name]) < tol, f"{name}: {rel_err(grads[name], g_ref[name])}"


@pytest.mark.parametrize("Bsz,C,H,W", [(2, 4, 16, 16), (2, 8, 64, 48), (1, 4, 128, 96)])
def test_merge_association_is_the_references(Bsz, C, H, W):
    """y of the fused core == (y0 + y2) + transpose(y1 + y3) BIT FOR BIT (model/vmamba.py:55-60), the four direction
    outputs taken from the same kernels launched one by one without accumulation."""
    from vm_asr_b200 import scan, ss2d
    x, dts, Bs, Cs, As, Ds, bias, _ = _scan_inputs(Bsz, C, H, W, seed=4)
    L = H * W
    y = _run_fused(x, dts, Bs, Cs, As, Ds, bias)
    xg = x.cuda()
    xT = ss2d.map_transpose(xg)
    (dts_rm, dts_cm), (Bs_rm, Bs_cm), (Cs_rm, Cs_cm) = (_time_to_memory_order(t) for t in (dts, Bs, Cs))
    outs = []
    for k in range(4):
        u = (xT if k & 1 else xg).view(Bsz, C, L)
        pick = lambda rm, cm: (cm if k & 1 else rm)[:, k // 2].cuda().contiguous()
        out = torch.empty(Bsz, C, L, device="cuda")
        xs = torch.empty(Bsz, C, (L + 2047) // 2048, 2, device="cuda")
        sl = slice(k * C, (k + 1) * C)
        scan.fwd_out(u, pick(dts_rm, dts_cm), As[sl].cuda(), pick(Bs_rm, Bs_cm).unsqueeze(1), pick(Cs_rm, Cs_cm).unsqueeze(1),
                     Ds[sl].cuda(), bias[sl].cuda(), True, out, xs, flags=scan.SCAN_REVERSE if k >= 2 else 0)
        outs.append(out)
    rowp = outs[0] + outs[2]
    colp = (outs[1] + outs[3]).view(Bsz, C, W, H).transpose(2, 3).reshape(Bsz, C, L)
    assert torch.equal(y, rowp + colp)


def _core_params(C, R, N=1, seed=2):
    torch.manual_seed(seed)
    return (torch.randn(4, R + 2 * N, C) * 0.3, torch.randn(4, C, R) * 0.3, torch.rand(4, C) * 0.5,
            torch.log(torch.rand(4 * C, N) + 0.5), torch.ones(4 * C))


@pytest.mark.parametrize("Bsz,C,H,W,R", [(2, 8, 72, 64, 2), (2, 4, 16, 16, 1), (1, 16, 32, 48, 1), (1, 6, 128, 60, 3)])
def test_ss2d_core_fused_forward_backward(Bsz, C, H, W, R):
    """vm_asr_b200.ss2d.ss2d_core (projections in memory order + fused kernels) against the float64 torch oracle of
    forward_corev2, outputs and gradients of the map and of every parameter."""
    from vm_asr_b200 import ss2d
    names = ("x", "xw", "dw", "db", "A_logs", "Ds")
    vals = (torch.randn(Bsz, C, H, W, generator=torch.Generator().manual_seed(5)),) + _core_params(C, R)
    dy = torch.randn(Bsz, C, H * W, generator=torch.Generator().manual_seed(6))
    ref_in = [v.double().requires_grad_() for v in vals]
    ref = ss2d_ref.ss2d_core(*ref_in, dtype=torch.float64)
    ref.backward(dy.double())
    got_in = [v.cuda().requires_grad_() for v in vals]
    got = ss2d.ss2d_core(*got_in, fused=True)
    got.backward(dy.cuda())
    chain_in = [v.cuda().requires_grad_() for v in vals]
    chain = ss2d.ss2d_core_chain(*chain_in)
    chain.backward(dy.cuda())
    assert rel_err(got, ref) < 2e-4  # the einsums run in fp32 on the GPU; summation order differs
    assert rel_err(got, chain) < 2e-5
    for n, g, c, r in zip(names, got_in, chain_in, ref_in):
        assert rel_err(g.grad, r.grad) < 5e-4, n
        assert rel_err(g.grad, c.grad) < 1e-4, n


def test_x_proj_bias():
    """x_proj_bias (vmamba.py:1474-1475) in both paths."""
    from vm_asr_b200 import ss2d
    Bsz, C, H, W, R = 2, 4, 24, 20, 2
    x = torch.randn(Bsz, C, H, W, generator=torch.Generator().manual_seed(7)).cuda()
    xw, dw, db, A_logs, Ds = (t.cuda() for t in _core_params(C, R))
    xb = (0.2 * torch.randn(4, R + 2, generator=torch.Generator().manual_seed(8))).cuda()
    a = ss2d.ss2d_core(x, xw, dw, db, A_logs, Ds, x_proj_bias=xb, fused=True)
    b = ss2d.ss2d_core_chain(x, xw, dw, db, A_logs, Ds, x_proj_bias=xb)
    c = ss2d.ss2d_core_chain(x, xw, dw, db, A_logs, Ds)
    assert rel_err(a, b) < 2e-5 and rel_err(b, c) > 1e-3


def test_pair_equals_two_single_calls():
    """ss2d_core_pair (the generator's two streams in one grid) is bit-identical to two single calls."""
    from vm_asr_b200 import ss2d
    outs = {}
    for mode in ("single", "pair"):
        maps, prms = [], []
        for s in (11, 12):
            x = torch.randn(2, 8, 64, 80, generator=torch.Generator().manual_seed(s)).cuda().requires_grad_()
            prm = [t.cuda().requires_grad_() for t in _core_params(8, 1, seed=s)]
            maps.append(x)
            prms.append(prm)
        if mode == "single":
            ys = [ss2d.ss2d_core(x, *p, fused=True) for x, p in zip(maps, prms)]
        else:
            ys = ss2d.ss2d_core_pair(maps[0], prms[0], maps[1], prms[1])
        (ys[0].sum() + (ys[1] * ys[1]).sum()).backward()
        outs[mode] = (ys, maps, prms)
    for a, b in zip(outs["single"][0], outs["pair"][0]):
        assert torch.equal(a, b)
    for a, b in zip(outs["single"][1], outs["pair"][1]):
        assert rel_err(b.grad, a.grad) < 1e-6


def test_fused_core_allocates_no_fourfold_copies():
    """xs and ys (4 map-sizes each) are never allocated: the fused call's peak extra memory stays below 4 map-sizes
    (y + 2 planes + chunk states), while the chain needs more than 8."""
    from vm_asr_b200 import ss2d
    Bsz, C, H, W = 4, 16, 256, 256
    x, dts, Bs, Cs, As, Ds, bias, _ = _scan_inputs(Bsz, C, H, W)
    map_bytes = Bsz * C * H * W * 4
    xg = x.cuda()
    xT = ss2d.map_transpose(xg)
    args = []
    for t in (dts, Bs, Cs):
        rm, cm = _time_to_memory_order(t)
        args += [rm.cuda(), cm.cuda()]
    par = [t.cuda() for t in (As, Ds, bias)]
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    with torch.no_grad():
        y = ss2d._SS2DScan.apply(True, 1, xg, xT, *args, *par)
    torch.cuda.synchronize()
    peak = torch.cuda.max_memory_allocated() - base
    assert peak < 3.6 * map_bytes, peak / map_bytes
    assert y.shape == (Bsz, C, H * W)


def test_fused_core_rejects_unsupported_maps():
    from vm_asr_b200 import ss2d
    x = torch.randn(1, 2, 6, 10, device="cuda")
    prm = [t.cuda() for t in _core_params(2, 1)]
    with pytest.raises(RuntimeError):
        ss2d.ss2d_core(x, *prm, fused=True)
    y = ss2d.ss2d_core(x, *prm)  # falls back to the chain of this library's operators
    assert y.shape == (1, 2, 60)


# ---- delta generated inside the scan kernels (vmamba.py:1476-1477) --------------------------------------------
# maps of more than one chunk with dt_rank 1: ragged last chunk (72x64, 100x84), 2 / 3 / 7 / 16 channels (tiles of 2, 3, 3+3+1 and
# 5x3+1 channels), and the three largest maps of the 48 kHz config (R = 1 wherever L >= 16384) + the largest of all
PROJ_MAPS = [(2, 2, 64, 48), (1, 3, 72, 64), (2, 16, 64, 80), (1, 7, 100, 84), (4, 2, 512, 512), (4, 16, 256, 256), (4, 32, 128, 128),
             (8, 2, 1024, 512)]


@pytest.mark.parametrize("Bsz,C,H,W", PROJ_MAPS)
def test_projected_core_matches_materialised_delta(Bsz, C, H, W):
    """ss2d_core with the dt projection inside the kernels (x_dbl and dt_projs_weight handed to vmasr_ss2d_core_*) against the
    same fused core fed a materialised delta, outputs and the gradients of the map and of every parameter; small maps also
    against the float64 torch oracle of forward_corev2."""
    from vm_asr_b200 import ss2d
    names = ("x", "xw", "dw", "db", "A_logs", "Ds")
    vals = (torch.randn(Bsz, C, H, W, generator=torch.Generator().manual_seed(15)),) + _core_params(C, 1, seed=16)
    xb = 0.2 * torch.randn(4, 3, generator=torch.Generator().manual_seed(17))
    dy = torch.randn(Bsz, C, H * W, generator=torch.Generator().manual_seed(18))
    runs = {}
    for projected in (True, False):
        ins = [v.cuda().requires_grad_() for v in vals]
        bias = xb.cuda().requires_grad_()
        y = ss2d.ss2d_core(*ins, x_proj_bias=bias, fused=True, projected=projected)
        y.backward(dy.cuda())
        runs[projected] = (y, ins + [bias])
    assert rel_err(runs[True][0], runs[False][0]) < 2e-5
    for n, g, c in zip(names + ("x_proj_bias",), runs[True][1], runs[False][1]):
        assert rel_err(g.grad, c.grad) < 1e-4, (n, rel_err(g.grad, c.grad))
    if Bsz * C * H * W <= 1 << 17:
        ref_in = [v.double().requires_grad_() for v in vals]
        rb = xb.double().requires_grad_()
        ref = ss2d_ref.ss2d_core(*ref_in, x_proj_bias=rb, dtype=torch.float64)
        ref.backward(dy.double())
        assert rel_err(runs[True][0], ref) < 2e-4
        for n, g, r in zip(names + ("x_proj_bias",), runs[True][1], ref_in + [rb]):
            assert rel_err(g.grad, r.grad) < 5e-4, n


def test_projected_pair_and_activation_memory():
    """Both streams' cores in one grid in the projected form equal two single calls bit for bit; and the projected form never
    allocates a (B, 4C, L) delta: the forward's extra memory stays below the map-sized buffers it needs (x^T, y, two planes)
    plus the 12 rows of x_dbl."""
    from vm_asr_b200 import ss2d
    Bsz, C, H, W = 2, 16, 128, 128
    maps = [torch.randn(Bsz, C, H, W, generator=torch.Generator().manual_seed(s)).cuda() for s in (21, 22)]
    prms = [[t.cuda() for t in _core_params(C, 1, seed=s)] for s in (23, 24)]
    singles = [ss2d.ss2d_core(x, *p, projected=True) for x, p in zip(maps, prms)]
    pair = ss2d.ss2d_core_pair(maps[0], prms[0], maps[1], prms[1])
    for a, b in zip(singles, pair):
        assert torch.equal(a, b)
    map_bytes = Bsz * C * H * W * 4
    del singles, pair
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    with torch.no_grad():
        y = ss2d.ss2d_core(maps[0], *prms[0], projected=True)
    torch.cuda.synchronize()
    peak = torch.cuda.max_memory_allocated() - base
    assert peak < (4 + 12 / C + 0.5) * map_bytes, peak / map_bytes   # a materialised delta alone is 4 map sizes more
    assert y.shape == (Bsz, C, H * W)


@pytest.mark.parametrize("batch,dim,ngroups,L,rev", [(2, 8, 4, 4096 + 512, False), (1, 12, 2, 3 * 2048, True), (2, 20, 4, 8192 + 16, False)])
def test_scan_level_projected_form(batch, dim, ngroups, L, rev):
    """vmasr_scan_fwd / _bwd with dt_rank 1 and several B / C groups (dt row index = group), forward and time-reversed,
    against the same kernels fed delta = dt_weight * dt_rows: out, chunk states, du, dA, dB, dC, dD, ddelta_bias, and the two
    factor gradients d_dt_rows = sum_d w_d ddelta_d, d_dt_weight = sum_{b,l} ddelta * row."""
    from vm_asr_b200 import scan
    g = torch.Generator().manual_seed(31)
    cpg = dim // ngroups
    u = torch.randn(batch, dim, L, generator=g).cuda()
    rows = (0.5 * torch.randn(batch, ngroups, 1, L, generator=g)).cuda()
    w = (0.5 * torch.randn(dim, 1, generator=g)).cuda()
    A = (-0.5 * torch.rand(dim, 1, generator=g)).cuda()
    Bm, Cm = torch.randn(batch, ngroups, 1, L, generator=g).cuda(), torch.randn(batch, ngroups, 1, L, generator=g).cuda()
    D, bias = torch.randn(dim, generator=g).cuda(), (0.5 * torch.rand(dim, generator=g)).cuda()
    dout = torch.randn(batch, dim, L, generator=g).cuda()
    flags = scan.SCAN_REVERSE if rev else 0
    delta = (rows[:, :, 0].repeat_interleave(cpg, dim=1) * w.view(1, dim, 1)).contiguous()
    out_ref = torch.empty_like(u)
    x_ref = torch.empty(batch, dim, (L + 2047) // 2048, 2, device="cuda")
    scan.fwd_out(u, delta, A, Bm, Cm, D, bias, True, out_ref, x_ref, flags=flags)
    out, x = scan.fwd_projected(u, rows, w, A, Bm, Cm, D, bias, True, flags=flags)
    assert rel_err(out, out_ref) < 2e-5 and rel_err(x, x_ref) < 2e-5
    du_r, dd_r = torch.empty_like(u), torch.empty_like(u)
    dA_r, dB_r, dC_r, dD_r, db_r = scan._grad_buffers(u, A, D, bias, (batch, dim, L, 1, ngroups))
    scan.bwd_out(u, delta, A, Bm, Cm, D, bias, dout, x_ref, True, du_r, dd_r, dA_r, dB_r, dC_r, dD_r, db_r, flags=flags)
    du, d_rows, d_w, dA, dB, dC, dD, db = scan.bwd_projected(u, rows, w, A, Bm, Cm, D, bias, dout, x, True, flags=flags)
    d_rows_ref = (dd_r.double() * w.double().view(1, dim, 1)).view(batch, ngroups, cpg, L).sum(2).unsqueeze(2)
    d_w_ref = (dd_r.double() * rows[:, :, 0].double().repeat_interleave(cpg, dim=1)).sum((0, 2)).view(dim, 1)
    for name, got, ref, tol in (("du", du, du_r, 1e-4), ("dA", dA, dA_r, 3e-4), ("dB", dB, dB_r, 1e-4), ("dC", dC, dC_r, 1e-4),
                                ("dD", dD, dD_r, 3e-4), ("dbias", db, db_r, 3e-4), ("d_dt_rows", d_rows, d_rows_ref, 1e-4),
                                ("d_dt_weight", d_w, d_w_ref, 3e-4)):
        assert rel_err(got, ref) < tol, (name, rel_err(got, ref))
    # anything the multi-chunk fast kernels do not take is refused, not silently materialised
    with pytest.raises(RuntimeError):
        scan.fwd_projected(u[..., :1024].contiguous(), rows[..., :1024].contiguous(), w, A, Bm[..., :1024].contiguous(),
                           Cm[..., :1024].contiguous(), D, bias, True)


# ---- merge + LayerNorm + cast + SiLU(z) gate in one kernel (vmamba.py:1525-1531, 1536-1550) -------------------
def _tail_oracle(planes, gamma, beta, z, H, W, z_silu, out_dtype, dtype=torch.float64):
    """P_rm + transpose(P_cm) (vmamba.py:57-60) -> oracle tail, in float64 (or at ``dtype``) with autograd"""
    _, Bsz, C, L = planes.shape
    y = planes[0] + planes[1].view(Bsz, C, W, H).transpose(2, 3).reshape(Bsz, C, L)
    return ss2d_ref.out_norm_gate(y, gamma, beta, z, H, W, 1e-5, z_silu, out_dtype)


TAIL_CASES = [(2, 2, 64, 64), (1, 16, 32, 48), (2, 32, 16, 16), (1, 64, 8, 16), (1, 256, 16, 16), (1, 512, 16, 16), (1, 6, 12, 24),
              (2, 24, 20, 40), (4, 2, 512, 512), (4, 16, 256, 256), (8, 512, 16, 16)]


@pytest.mark.parametrize("Bsz,C,H,W", TAIL_CASES)
@pytest.mark.parametrize("gate", ["silu", "plain", "none"])
def test_merge_norm_gate_fp32(Bsz, C, H, W, gate):
    """forward and every gradient (both planes, gamma, beta, z) against the float64 oracle"""
    from vm_asr_b200 import ss2d
    if gate != "silu" and Bsz * C * H * W > 1 << 18:
        pytest.skip("large maps once")
    g = torch.Generator().manual_seed(40)
    L = H * W
    planes = torch.randn(2, Bsz, C, L, generator=g)
    gamma, beta = 1.0 + 0.3 * torch.randn(C, generator=g), 0.3 * torch.randn(C, generator=g)
    z = None if gate == "none" else torch.randn(Bsz, H, W, C, generator=g)
    gout = torch.randn(Bsz, H, W, C, generator=g)
    ins = [t.cuda().requires_grad_() if t is not None else None for t in (planes, gamma, beta, z)]
    out = ss2d.MergeNormGate.apply(ins[0], ins[1], ins[2], ins[3], H, W, 1e-5, gate == "silu", torch.float32)
    out.backward(gout.cuda())
    ref_in = [t.double().requires_grad_() if t is not None else None for t in (planes, gamma, beta, z)]
    ref = _tail_oracle(ref_in[0], ref_in[1], ref_in[2], ref_in[3], H, W, gate == "silu", None)
    ref.backward(gout.double())
    assert out.shape == (Bsz, H, W, C)
    # two channels: xhat = +-1 / sqrt(1 + eps / var) with var = (y0 - y1)^2 / 4 -- where the two values nearly coincide the fp32
    # subtraction in front of a large rstd is ill-conditioned (the reference's fp32 LayerNorm has the same property)
    cond = 10.0 if C == 2 else 1.0
    assert rel_err(out, ref) < 1e-5 * cond
    for name, a, r in zip(("planes", "gamma", "beta", "z"), ins, ref_in):
        if a is not None:
            tol = 2e-4 if name in ("gamma", "beta") else 2e-5 * cond   # gamma / beta: sums over B * L terms in fp32
            assert rel_err(a.grad, r.grad) < tol, (name, rel_err(a.grad, r.grad))


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 2e-3), (torch.bfloat16, 1.6e-2)])
@pytest.mark.parametrize("Bsz,C,H,W", [(2, 32, 16, 16), (1, 16, 32, 48), (2, 2, 64, 64)])
def test_merge_norm_gate_half(Bsz, C, H, W, dtype, tol):
    """AMP: z and the result are half tensors (y.to(x.dtype), act(z) rounded to half, half product), the LayerNorm is fp32;
    against the same statements in torch on the GPU at the same dtypes"""
    from vm_asr_b200 import ss2d
    g = torch.Generator().manual_seed(41)
    planes = torch.randn(2, Bsz, C, H * W, generator=g).cuda()
    gamma, beta = (1.0 + 0.3 * torch.randn(C, generator=g)).cuda(), (0.3 * torch.randn(C, generator=g)).cuda()
    z = torch.randn(Bsz, H, W, C, generator=g).cuda().to(dtype)
    gout = torch.randn(Bsz, H, W, C, generator=g).cuda().to(dtype)
    ins = [planes.clone().requires_grad_(), gamma.clone().requires_grad_(), beta.clone().requires_grad_(), z.clone().requires_grad_()]
    out = ss2d.MergeNormGate.apply(*ins, H, W, 1e-5, True, dtype)
    out.backward(gout)
    ref_in = [planes.clone().requires_grad_(), gamma.clone().requires_grad_(), beta.clone().requires_grad_(), z.clone().requires_grad_()]
    ref = _tail_oracle(*ref_in, H, W, True, dtype)
    ref.backward(gout)
    assert out.dtype == dtype and rel_err(out, ref) < tol
    for name, a, r in zip(("planes", "gamma", "beta", "z"), ins, ref_in):
        assert rel_err(a.grad, r.grad) < 2 * tol, (name, rel_err(a.grad, r.grad))


@pytest.mark.parametrize("Bsz,C,H,W,R", [(2, 8, 72, 64, 2), (2, 4, 16, 16, 1), (1, 16, 64, 80, 1), (4, 16, 256, 256, 1)])
def test_ss2d_core_out_matches_core_plus_tail(Bsz, C, H, W, R):
    """ss2d_core_out (planes of the fused core -> MergeNormGate) against ss2d_core followed by the oracle's tail in torch on the
    GPU: output and the gradients of the map, the gate, the LayerNorm and every parameter of the core"""
    from vm_asr_b200 import ss2d
    g = torch.Generator().manual_seed(42)
    vals = (torch.randn(Bsz, C, H, W, generator=g),) + _core_params(C, R, seed=43)
    gamma, beta = 1.0 + 0.3 * torch.randn(C, generator=g), 0.3 * torch.randn(C, generator=g)
    z = torch.randn(Bsz, H, W, C, generator=g)
    gout = torch.randn(Bsz, H, W, C, generator=g).cuda()
    a_in = [v.cuda().requires_grad_() for v in vals + (gamma, beta, z)]
    out = ss2d.ss2d_core_out(*a_in[:6], a_in[6], a_in[7], z=a_in[8])
    out.backward(gout)
    b_in = [v.cuda().requires_grad_() for v in vals + (gamma, beta, z)]
    y = ss2d.ss2d_core(*b_in[:6], fused=True)
    ref = ss2d_ref.out_norm_gate(y, b_in[6], b_in[7], b_in[8], H, W)
    ref.backward(gout)
    assert rel_err(out, ref) < 2e-5
    for n, a, r in zip(("x", "xw", "dw", "db", "A_logs", "Ds", "gamma", "beta", "z"), a_in, b_in):
        assert rel_err(a.grad, r.grad) < 2e-4, (n, rel_err(a.grad, r.grad))


def test_merge_norm_gate_inference_keeps_nothing():
    """without autograd the merged map and the statistics are not written: the tail allocates its output only"""
    from vm_asr_b200 import ss2d
    Bsz, C, H, W = 2, 32, 64, 64
    planes = torch.randn(2, Bsz, C, H * W, device="cuda")
    gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    z = torch.randn(Bsz, H, W, C, device="cuda")
    torch.cuda.synchronize()
    base = torch.cuda.memory_allocated()
    torch.cuda.reset_peak_memory_stats()
    with torch.no_grad():
        out = ss2d.MergeNormGate.apply(planes, gamma, beta, z, H, W, 1e-5, True, torch.float32)
    torch.cuda.synchronize()
    assert torch.cuda.max_memory_allocated() - base <= out.numel() * 4 + (1 << 20)
    ref = _tail_oracle(planes, gamma, beta, z, H, W, True, None, dtype=torch.float32)
    assert rel_err(out, ref) < 1e-5


# ---- permute + depthwise conv 3x3 + SiLU + transpose in one kernel (vmamba.py:1541-1546) ------------------------
HEAD_CASES = [(2, 16, 16, 8), (1, 32, 48, 16), (2, 64, 64, 2), (1, 12, 24, 6), (1, 16, 16, 256), (1, 8, 16, 70), (2, 128, 128, 32),
              (4, 512, 512, 2)]


@pytest.mark.parametrize("Bsz,H,W,C", HEAD_CASES)
@pytest.mark.parametrize("strided", [False, True])
def test_conv_silu_input_fp32(Bsz, H, W, C, strided):
    """x and x^T against the float64 oracle (x^T bit-identical to transpose(x)); d input, d weight, d bias with gradients arriving
    through BOTH outputs; `strided`: the input is the first half of a (B, H, W, 2C) tensor, read in place"""
    from vm_asr_b200 import ss2d
    if strided and Bsz * H * W * C > 1 << 20:
        pytest.skip("large maps once")
    g = torch.Generator().manual_seed(50)
    full = torch.randn(Bsz, H, W, 2 * C if strided else C, generator=g)
    wt, bs = 0.4 * torch.randn(C, 1, 3, 3, generator=g), 0.3 * torch.randn(C, generator=g)
    gx, gxT = torch.randn(Bsz, C, H, W, generator=g), torch.randn(Bsz, C, W, H, generator=g)
    leaf = full.cuda().requires_grad_()
    w_, b_ = wt.cuda().requires_grad_(), bs.cuda().requires_grad_()
    x, xT = ss2d.ConvSiluInput.apply(leaf[..., :C], w_, b_)
    assert torch.equal(xT, x.transpose(2, 3).contiguous())
    ((x * gx.cuda()).sum() + (xT * gxT.cuda()).sum()).backward()
    rleaf, rw, rb = full.double().requires_grad_(), wt.double().requires_grad_(), bs.double().requires_grad_()
    rx = ss2d_ref.dwconv_silu(rleaf[..., :C], rw, rb)
    ((rx * gx.double()).sum() + (rx.transpose(2, 3) * gxT.double()).sum()).backward()
    assert rel_err(x, rx) < 1e-5
    assert rel_err(leaf.grad, rleaf.grad) < 2e-5
    assert rel_err(w_.grad, rw.grad) < 2e-4 and rel_err(b_.grad, rb.grad) < 2e-4   # sums over B * H * W terms in fp32


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 2e-3), (torch.bfloat16, 1.6e-2)])
def test_conv_silu_input_half(dtype, tol):
    """AMP: the input is a half tensor; conv and activation are rounded to half as the reference's tensors are, x / x^T come out
    in float32 (the scan is forced to fp32, vmamba.py:1487-1491); against the same statements in torch on the GPU"""
    from vm_asr_b200 import ss2d
    Bsz, H, W, C = 2, 32, 32, 24
    g = torch.Generator().manual_seed(51)
    xin = torch.randn(Bsz, H, W, C, generator=g).cuda().to(dtype)
    wt, bs = (0.4 * torch.randn(C, 1, 3, 3, generator=g)).cuda(), (0.3 * torch.randn(C, generator=g)).cuda()
    gx = torch.randn(Bsz, C, H, W, generator=g).cuda()
    a = [xin.clone().requires_grad_(), wt.clone().requires_grad_(), bs.clone().requires_grad_()]
    x, xT = ss2d.ConvSiluInput.apply(*a)
    (x * gx).sum().backward()
    r = [xin.clone().requires_grad_(), wt.clone().requires_grad_(), bs.clone().requires_grad_()]
    rx = ss2d_ref.dwconv_silu(r[0], r[1].to(dtype), r[2].to(dtype)).float()
    (rx * gx).sum().backward()
    assert x.dtype == torch.float32 and rel_err(x, rx) < tol
    assert a[0].grad.dtype == dtype
    for n, u, v in zip(("xin", "weight", "bias"), a, r):
        assert rel_err(u.grad, v.grad) < 2 * tol, (n, rel_err(u.grad, v.grad))


@pytest.mark.parametrize("Bsz,H,W,C,R", [(2, 16, 16, 8, 1), (1, 32, 48, 16, 2), (2, 64, 64, 2, 1), (1, 8, 16, 70, 3), (1, 16, 16, 256, 8),
                                         (2, 128, 128, 32, 1)])
def test_conv_silu_input_forms_x_dbl(Bsz, H, W, C, R):
    """the head kernel with x_proj inside (vmamba.py:1473-1475): x_dbl of the row-major and of the column-major pair against the
    einsums on the oracle's activations in float64 (70 and 256 channels span several CTAs: atomic partial sums), and the
    gradients of input, conv, x_proj_weight and x_proj_bias with gradients arriving through all four outputs"""
    from vm_asr_b200 import ss2d
    g = torch.Generator().manual_seed(54)
    L = H * W
    xin = torch.randn(Bsz, H, W, C, generator=g)
    wt, bs = 0.4 * torch.randn(C, 1, 3, 3, generator=g), 0.3 * torch.randn(C, generator=g)
    xw, xb = 0.3 * torch.randn(4, R + 2, C, generator=g), 0.2 * torch.randn(4, R + 2, generator=g)
    ups = [torch.randn(Bsz, C, H, W, generator=g), torch.randn(Bsz, C, W, H, generator=g),
           torch.randn(Bsz, 2, R + 2, L, generator=g), torch.randn(Bsz, 2, R + 2, L, generator=g)]
    a = [t.cuda().requires_grad_() for t in (xin, wt, bs, xw, xb)]
    outs = ss2d.ConvSiluInput.apply(*a)
    sum((o * u.cuda()).sum() for o, u in zip(outs, ups)).backward()
    r = [t.double().requires_grad_() for t in (xin, wt, bs, xw, xb)]
    rx = ss2d_ref.dwconv_silu(r[0], r[1], r[2])
    rxT = rx.transpose(2, 3)
    rd_rm = torch.einsum("bdl,kcd->bkcl", rx.reshape(Bsz, C, L), r[3][0::2]) + r[4][0::2].view(1, 2, -1, 1)
    rd_cm = torch.einsum("bdl,kcd->bkcl", rxT.reshape(Bsz, C, L), r[3][1::2]) + r[4][1::2].view(1, 2, -1, 1)
    sum((o * u.double()).sum() for o, u in zip((rx, rxT, rd_rm, rd_cm), ups)).backward()
    assert rel_err(outs[0], rx) < 1e-5 and rel_err(outs[2], rd_rm) < 2e-5 and rel_err(outs[3], rd_cm) < 2e-5
    for n, u, v in zip(("xin", "conv_w", "conv_b", "x_proj_w", "x_proj_b"), a, r):
        assert rel_err(u.grad, v.grad) < 2e-4, (n, rel_err(u.grad, v.grad))


@pytest.mark.parametrize("Bsz,C,H,W,R", [(2, 8, 72, 64, 2), (1, 16, 64, 80, 1), (2, 32, 16, 16, 2)])
def test_ss2d_block_core_matches_separate_statements(Bsz, C, H, W, R):
    """ss2d_block_core (head kernel -> fused core -> tail kernel) against oracle head -> ss2d_core -> oracle tail in torch on the GPU:
    output and the gradients of the input, the gate and every parameter (conv, core, LayerNorm)"""
    from vm_asr_b200 import ss2d
    g = torch.Generator().manual_seed(52)
    xz = torch.randn(Bsz, H, W, 2 * C, generator=g)                      # in_proj's output: x half and z half side by side
    wt, bs = 0.4 * torch.randn(C, 1, 3, 3, generator=g), 0.3 * torch.randn(C, generator=g)
    core = _core_params(C, R, seed=53)
    gamma, beta = 1.0 + 0.3 * torch.randn(C, generator=g), 0.3 * torch.randn(C, generator=g)
    gout = torch.randn(Bsz, H, W, C, generator=g).cuda()
    names = ("xz", "conv_w", "conv_b", "xw", "dw", "db", "A_logs", "Ds", "gamma", "beta")
    vals = (xz, wt, bs) + core + (gamma, beta)
    a = [v.cuda().requires_grad_() for v in vals]
    xa, za = a[0].chunk(2, dim=-1)
    out = ss2d.ss2d_block_core(xa, a[1], a[2], *a[3:8], a[8], a[9], z=za)
    out.backward(gout)
    b = [v.cuda().requires_grad_() for v in vals]
    xb, zb = b[0].chunk(2, dim=-1)
    y = ss2d.ss2d_core(ss2d_ref.dwconv_silu(xb, b[1], b[2]), *b[3:8], fused=True)
    ref = ss2d_ref.out_norm_gate(y, b[8], b[9], zb, H, W)
    ref.backward(gout)
    assert rel_err(out, ref) < 2e-5
    for n, u, v in zip(names, a, b):
        assert rel_err(u.grad, v.grad) < 2e-4, (n, rel_err(u.grad, v.grad))


@pytest.mark.parametrize("R", [1, 2])
def test_block_core_pair_equals_two_single_calls(R):
    """ss2d_block_core_pair (the generator's two streams: heads and tails one after the other, the two cores' scans in ONE grid)
    against two ss2d_block_core calls: same output bits, same gradients of inputs, gates and every parameter"""
    from vm_asr_b200 import ss2d
    Bsz, C, H, W = 2, 8, 64, 80
    res = {}
    for mode in ("single", "pair"):
        ins, blks = [], []
        for s in (61, 62):
            g = torch.Generator().manual_seed(s)
            xz = torch.randn(Bsz, H, W, 2 * C, generator=g).cuda().requires_grad_()
            wt, bs = 0.4 * torch.randn(C, 1, 3, 3, generator=g), 0.3 * torch.randn(C, generator=g)
            gamma, beta = 1.0 + 0.3 * torch.randn(C, generator=g), 0.3 * torch.randn(C, generator=g)
            blk = [t.cuda().requires_grad_() for t in (wt, bs) + _core_params(C, R, seed=s) + (gamma, beta)]
            ins.append(xz)
            blks.append(blk)
        halves = [t.chunk(2, dim=-1) for t in ins]
        if mode == "single":
            ys = [ss2d.ss2d_block_core(h[0], *blk, z=h[1]) for h, blk in zip(halves, blks)]
        else:
            ys = ss2d.ss2d_block_core_pair(halves[0][0], blks[0], halves[1][0], blks[1], z_a=halves[0][1], z_b=halves[1][1])
        (ys[0].sum() + (ys[1] * ys[1]).sum()).backward()
        res[mode] = (ys, ins, blks)
    for a, b in zip(res["single"][0], res["pair"][0]):
        assert torch.equal(a, b)
    for a, b in zip(res["single"][1], res["pair"][1]):
        assert rel_err(b.grad, a.grad) < 1e-6
    for pa, pb in zip(res["single"][2], res["pair"][2]):
        for u, v in zip(pa, pb):
            assert rel_err(v.grad, u.grad) < 1e-5
