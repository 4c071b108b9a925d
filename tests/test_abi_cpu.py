"""CPU: the C-ABI library loads and exports every symbol include/vmasr_b200.h declares; host-side logic that
needs no GPU (argument checks, workspace sizing, the package alias)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "vmasr_b200.h")).read()
    return sorted(set(re.findall(r"VMASR_API\s+[\w\s\*]+?\b(vmasr_\w+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = _declared_symbols()
    for must in ("vmasr_scan_fwd", "vmasr_scan_bwd", "vmasr_cross_scan", "vmasr_cross_merge", "vmasr_stft_fwd",
                 "vmasr_istft_fwd", "vmasr_istft_bwd", "vmasr_scan_workspace_bytes", "vmasr_last_error"):
        assert must in syms


def test_library_loads_and_exports_every_declared_symbol():
    import vm_asr_b200
    from vm_asr_b200 import _lib
    lib = vm_asr_b200.load_library()
    raw = ctypes.CDLL(vm_asr_b200.library_path())
    for name in _declared_symbols():
        assert hasattr(raw, name), name
        assert name in _lib.EXPORTS, f"{name} missing from the ctypes binding table"
    assert lib.vmasr_abi_version() == _lib.ABI_VERSION


def test_workspace_sizing():
    import vm_asr_b200
    lib = vm_asr_b200.load_library()
    assert lib.vmasr_scan_workspace_bytes(4, 8, 2048, 1) == 0          # single chunk: no exchange needed
    small = lib.vmasr_scan_workspace_bytes(4, 8, 2049, 1)
    big = lib.vmasr_scan_workspace_bytes(4, 8, 262144, 1)
    assert 0 < small < big
    assert big >= 4 * 8 * 128 * 16                                      # 16 bytes per (b, d, n, chunk) entry
    assert lib.vmasr_scan_workspace_bytes(0, 8, 4096, 1) == 0


@pytest.mark.parametrize("struct,mirror", [("vmasr_scan_params", "ScanParams"), ("vmasr_ss2d_params", "SS2DParams"),
                                           ("vmasr_outnorm_params", "OutNormParams"), ("vmasr_dwconv_params", "DwConvParams")])
def test_struct_layout_matches_header(struct, mirror):
    """every ctypes mirror has the fields of its struct in the header, in order (array members by name)"""
    from vm_asr_b200 import _lib
    text = open(os.path.join(ROOT, "include", "vmasr_b200.h")).read()
    start = text.index("typedef struct %s {" % struct) + len("typedef struct %s {" % struct)
    body = text[start:text.index("} %s;" % struct)]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    quals = {"const", "void", "float", "int32_t", "int64_t", "uint64_t"}
    names = []
    for decl in body.split(";"):
        parts = re.sub(r"\[\d+\]", "", decl).replace("*", " ").replace(",", " ").split()
        names += [p for p in parts if p not in quals]
    assert names == [f[0] for f in getattr(_lib, mirror)._fields_]


def test_operators_refuse_cpu_tensors():
    """No CPU fallback: every operator raises on host tensors."""
    from vm_asr_b200 import cross, scan, stft
    u = torch.zeros(1, 4, 8)
    with pytest.raises(RuntimeError):
        scan.selective_scan_fn(u, u, torch.zeros(4, 1), torch.zeros(1, 1, 1, 8), torch.zeros(1, 1, 1, 8))
    with pytest.raises(RuntimeError):
        cross.CrossScan.apply(torch.zeros(1, 2, 4, 4))
    with pytest.raises(RuntimeError):
        cross.CrossMerge.apply(torch.zeros(1, 4, 2, 4, 4))
    with pytest.raises(RuntimeError):
        stft.wav2spectro(torch.zeros(1, 1, 4096), 1024, 240, 1024, "log2")
    with pytest.raises(RuntimeError):
        stft.spectro2wav(torch.zeros(1, 1, 513, 8), torch.zeros(1, 1, 513, 8), 1024, 240, 1024, "log2")


def test_argument_checks_mirror_reference():
    """Host-side checks fire before any device work (selective_scan.cpp:165-215)."""
    from vm_asr_b200 import scan
    u = torch.zeros(1, 4, 8)
    with pytest.raises(RuntimeError, match="float32, float16 or bfloat16"):
        scan.fwd(u.double(), u.double(), torch.zeros(4, 1), torch.zeros(1, 1, 1, 8).double(), torch.zeros(1, 1, 1, 8).double())
    with pytest.raises(RuntimeError, match="A must be float32"):
        scan.fwd(u, u, torch.zeros(4, 1).half(), torch.zeros(1, 1, 1, 8), torch.zeros(1, 1, 1, 8))


def test_header_is_valid_c_and_the_library_links_from_c(tmp_path):
    """include/vmasr_b200.h compiles as C99 (-Wall -Wextra -pedantic) and a plain C program links the shared library and
    gets plans, sizes and error text back without a GPU (tests/c/abi_smoke.c)."""
    import shutil
    import subprocess
    import vm_asr_b200
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    vm_asr_b200.load_library()
    libdir = os.path.dirname(vm_asr_b200.library_path())
    exe = str(tmp_path / "abi_smoke")
    build = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                            os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-L", libdir, "-lvmasr_b200",
                            "-Wl,-rpath," + libdir, "-o", exe], capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 0, (run.returncode, run.stdout, run.stderr)
    assert "family 2" in run.stdout
