"""CPU: the host side of the selective scan -- argument checks (the reference's TORCH_CHECKs, selective_scan.cpp:165-215,
262-317) and tile planning -- through ``vmasr_scan_plan``, which validates and plans exactly like ``vmasr_scan_fwd`` /
``vmasr_scan_bwd`` but touches no device.  Pointers are fake (only tested for null / alignment)."""
import ctypes

import pytest

GENERIC, SINGLE, MULTI, RING = 0, 1, 2, 3
SM = 148  # sm_count() falls back to 148 when there is no device


def _params(B, D, L, G=4, N=1, dtype=0, bwd=False, base=0x10000000, misalign=0):
    from vm_asr_b200 import _lib
    p = _lib.ScanParams()
    es = 4 if dtype == 0 else 2
    n_chunks = (L + 2047) // 2048
    ptr = [base]

    def fake(nbytes):
        v = ptr[0]
        ptr[0] += (nbytes + 255) // 256 * 256
        return v

    p.u, p.delta = fake(B * D * L * es) + misalign, fake(B * D * L * es)
    p.A, p.D, p.delta_bias = fake(D * N * 4), fake(D * 4), fake(D * 4)
    p.B, p.C = fake(B * G * N * L * es), fake(B * G * N * L * es)
    p.out, p.x = fake(B * D * L * es), fake(B * D * n_chunks * 2 * N * 4)
    if bwd:
        p.dout, p.du, p.ddelta = fake(B * D * L * es), fake(B * D * L * es), fake(B * D * L * es)
        p.dA, p.dB, p.dC = fake(D * N * 4), fake(B * G * N * L * 4), fake(B * G * N * L * 4)
        p.dD, p.ddelta_bias = fake(D * 4), fake(D * 4)
    lib = _lib.load_library()
    p.workspace_bytes = lib.vmasr_scan_workspace_bytes(B, D, L, N)
    p.workspace = fake(p.workspace_bytes) if p.workspace_bytes else 0
    p.batch, p.dim, p.seqlen, p.dstate, p.ngroups = B, D, L, N, G
    for name in ("u", "delta", "out", "dout", "du", "ddelta"):
        setattr(p, name + "_batch_stride", D * L)
        setattr(p, name + "_d_stride", L)
    p.A_d_stride, p.A_dstate_stride = N, 1
    for name in ("B", "C"):
        setattr(p, name + "_batch_stride", G * N * L)
        setattr(p, name + "_group_stride", N * L)
        setattr(p, name + "_dstate_stride", L)
    p.io_dtype, p.delta_softplus, p.device = dtype, 1, 0
    return p


def _plan(p, bwd=False):
    from vm_asr_b200 import _lib
    lib = _lib.load_library()
    out = (ctypes.c_int32 * 6)()
    rc = lib.vmasr_scan_plan(ctypes.byref(p), int(bwd), out)
    return rc, dict(zip(("grid", "variant", "cpt", "n_ctiles", "n_chunks", "tpr"), out)), lib.vmasr_last_error().decode()


# every SS2D call shape of the four configs: B, D = 4 d_inner, L
CONFIG_CALLS = [(4, 8, 262144), (4, 64, 65536), (4, 128, 16384), (4, 256, 4096), (4, 512, 1024), (4, 1024, 256),
                (8, 8, 524288), (8, 64, 131072), (8, 128, 32768), (8, 256, 8192), (8, 512, 2048), (8, 1024, 512),
                (8, 128, 65536), (8, 2048, 256)]


@pytest.mark.parametrize("B,D,L", CONFIG_CALLS)
@pytest.mark.parametrize("bwd", [False, True])
def test_config_shapes_take_the_fast_paths(B, D, L, bwd):
    rc, pl, err = _plan(_params(B, D, L, bwd=bwd), bwd)
    assert rc == 0, err
    n_chunks = (L + 2047) // 2048
    cpg = D // 4
    assert pl["n_chunks"] == n_chunks
    assert pl["variant"] == (MULTI if n_chunks > 1 else SINGLE)
    assert pl["n_ctiles"] == -(-cpg // pl["cpt"])
    assert pl["grid"] == B * 4 * n_chunks * pl["n_ctiles"]
    if n_chunks > 1:
        assert pl["tpr"] == 256 and 1 <= pl["cpt"] <= 4          # the tile must fit the resident shared-memory stages
    else:
        rows = 256 // pl["tpr"]
        assert pl["tpr"] * 8 >= L and pl["cpt"] % rows == 0 and pl["cpt"] <= 64
        assert pl["grid"] >= min(2 * SM, B * 4 * -(-cpg // rows)) // 2  # enough tiles for the machine when the shape has them


def test_everything_else_takes_the_generic_kernels():
    assert _plan(_params(2, 8, 4096, dtype=1))[1]["variant"] == GENERIC          # fp16 IO
    assert _plan(_params(2, 8, 4096, dtype=2, bwd=True), True)[1]["variant"] == GENERIC
    assert _plan(_params(2, 8, 4096, N=4))[1]["variant"] == GENERIC               # d_state > 1
    assert _plan(_params(1, 4, 65, G=2))[1]["variant"] == GENERIC                 # odd length: scalar IO
    assert _plan(_params(2, 8, 4096, misalign=4))[1]["variant"] == GENERIC        # unaligned view


def test_argument_checks_mirror_the_reference():
    rc, _, err = _plan(_params(2, 10, 256, G=4))
    assert rc != 0 and "dividable by n_groups" in err                              # selective_scan.cpp:190
    rc, _, err = _plan(_params(1, 4, 64, N=300))
    assert rc != 0 and "state dimension <= 256" in err                             # selective_scan.cpp:191
    p = _params(2, 8, 256)
    p.io_dtype = 7
    rc, _, err = _plan(p)
    assert rc != 0 and "float32, float16 or bfloat16" in err                       # selective_scan.cpp:167
    p = _params(2, 8, 256)
    p.B = 0
    assert _plan(p)[0] != 0
    p = _params(2, 8, 4096, bwd=True)
    p.x = 0
    rc, _, err = _plan(p, True)
    assert rc != 0 and "x (chunk states) is required" in err                       # selective_scan.cpp:310
    p = _params(2, 8, 4096)
    p.workspace = 0
    rc, _, err = _plan(p)
    assert rc != 0 and "workspace" in err
    p = _params(2, 8, 4096, bwd=True)
    p.dD = 0
    rc, _, err = _plan(p, True)
    assert rc != 0 and "dD must be given exactly when D is" in err


# ---- session 6: stored dB / dC and the zero-fill side job (host rules only; the kernels are tested in test_scan_gpu.py) -------------
DBDC_STORE = 8


@pytest.mark.parametrize("B,D,L,ok", [
    (4, 8, 262144, True),     # the C = 2 maps: one tile of two channels spans the group
    (2, 16, 4112, True),      # four channels per group, ragged last chunk
    (4, 64, 65536, False),    # sixteen channels per group: four channel tiles write every dB / dC element
    (4, 8, 1024, False),      # single chunk: no multi-chunk tile
    (1, 8, 4100, False),      # length not a multiple of 16: generic kernels
])
def test_dbdc_store_needs_one_channel_tile_per_group(B, D, L, ok):
    """VMASR_SCAN_DBDC_STORE is honoured exactly where every dB / dC element has one writer, and refused (nothing launched)
    otherwise -- the Python rule that decides whether to ASK agrees with the library on the config shapes."""
    import torch
    from vm_asr_b200 import scan
    p = _params(B, D, L, bwd=True)
    p.flags = DBDC_STORE
    rc, pl, err = _plan(p, True)
    assert (rc == 0) == ok, err
    if ok:
        assert pl["variant"] == MULTI and pl["n_ctiles"] == 1
    else:
        assert "VMASR_SCAN_DBDC_STORE" in err or "fast path" in err
    meta = lambda *s: torch.empty(*s, device="meta")
    cand = scan.dbdc_store_candidate(meta(B, D, L), meta(D, 1), meta(B, 4, 1, L))
    assert cand or not ok            # the candidate rule is necessary ...
    if L % 16 == 0:
        assert cand == ok            # ... and exact on aligned float32 calls


def test_dbdc_store_is_a_backward_flag():
    p = _params(4, 8, 8192)
    p.flags = DBDC_STORE
    rc, _, err = _plan(p, False)
    assert rc != 0 and "backward flag" in err


@pytest.mark.parametrize("ptr,nbytes,ok", [(0x20000000, 4096, True), (0x20000000, 16, True), (0, 0, True),
                                           (0x20000004, 4096, False), (0x20000000, 4100, False), (0, 64, False)])
@pytest.mark.parametrize("bwd", [False, True])
def test_zero_region_arguments(ptr, nbytes, ok, bwd):
    """zero_ptr / zero_bytes: 16-byte aligned, a multiple of 16 bytes, non-null when non-empty -- for every kernel family
    (the fast forward kernels clear it themselves, the others get a memset in front)."""
    for shape in ((4, 64, 16384), (4, 512, 1024), (1, 8, 132)):
        p = _params(*shape, bwd=bwd)
        p.zero_ptr, p.zero_bytes = ptr, nbytes
        rc, _, err = _plan(p, bwd)
        assert (rc == 0) == ok, err
        if not ok:
            assert "zero_ptr" in err
