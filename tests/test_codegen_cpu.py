"""CPU: register and spill budgets of the four fast scan kernels, read from the object files with ``cuobjdump -res-usage`` (no GPU).

The multi-chunk backward runs at its register cap (96 registers: two CTAs of 320 threads per SM) and the forward at 72 (three CTAs of
288): a few innocent lines at the top of a kernel can push the allocator into spilling inside the sweeps.  That happened once -- ten
lines of tensor-map prefetch doubled the backward's ``LDL`` / ``STL`` count (30 -> 58) and cost 3.5 % of the bench step before it showed
in a measurement (profiles/r2_summary.md 7) -- so the budgets are pinned here.  Skipped when the library has not been built in-tree."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "vm_asr_b200", "lib", "obj")


def _usage(name):
    path = os.path.join(OBJ, name)
    if not os.path.exists(path) or shutil.which("cuobjdump") is None:
        pytest.skip("object files / cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True, check=True).stdout
    rows = []
    fn = None
    for line in out.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            fn = m.group(1)
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and fn:
            rows.append((fn, *map(int, m.groups())))
            fn = None
    assert rows, out[:400]
    return rows


def _spills(name, kernel_substr):
    path = os.path.join(OBJ, name)
    fns = [r[0] for r in _usage(name) if kernel_substr in r[0]]
    assert fns
    total = {}
    for fn in fns:
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn, path], capture_output=True, text=True, check=True).stdout
        total[fn] = len(re.findall(r"\b(?:LDL|STL)\b", sass))
    return total


def test_forward_pipe_kernel_three_ctas_per_sm_no_spills():
    for fn, reg, stack, _, local in _usage("scan_fwd_pipe.o"):
        assert reg * 288 * 3 <= 65536, (fn, reg)     # __launch_bounds__(288, 3)
        assert stack == 0 and local == 0, (fn, stack, local)


def test_backward_pipe_kernel_two_ctas_per_sm_bounded_spills():
    for fn, reg, stack, _, local in _usage("scan_bwd_pipe.o"):
        assert reg * 320 * 2 <= 65536, (fn, reg)     # __launch_bounds__(320, 2)
        assert local == 0, (fn, local)
        f1 = "ELb1EEEvNS" in fn  # delta on the fly: one more row per tile
        assert stack <= (64 if f1 else 24), (fn, stack)
    # the variant the bench step runs (softplus, four stages, materialised delta): spill instructions in the whole kernel
    spills = _spills("scan_bwd_pipe.o", "scan_bwd_pipe_kernelILb1ELi4ELb0")
    assert all(n <= 32 for n in spills.values()), spills


def test_single_chunk_kernels_keep_their_occupancy():
    for fn, reg, stack, _, local in _usage("scan_fwd_tma.o"):
        assert reg * 256 * 3 <= 65536 and stack <= 32 and local == 0, (fn, reg, stack, local)   # __launch_bounds__(256, 3)
    for fn, reg, stack, _, local in _usage("scan_bwd_tma.o"):
        assert reg <= 128 and stack == 0 and local == 0, (fn, reg, stack, local)                # two CTAs of 256 per SM
