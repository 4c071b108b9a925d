"""ORACLE (test infrastructure): stage the reference's Python SS2D module next to its rebuilt CUDA extension so that the
model-level drop-in test can run on the GPU box, where /root/reference does not exist.

    python oracle/stage_ref_py.py      (run by __graft_entry__.build() in the build container)

Copies, unmodified, ``model/vmamba.py`` and ``model/csm_triton.py`` from the reference tree into ``oracle/_ref/py/``
(git-ignored, like the extension built by oracle/build_ref_cuda.py: it travels with the snapshot, it is never committed).
``load()`` imports the staged module with stubs for the two packages it pulls in at import time and never calls on this
path (timm, fvcore)."""
from __future__ import annotations

import importlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref", "py")
REF = os.environ.get("VMASR_REFERENCE", "/root/reference")
FILES = ("model/vmamba.py", "model/csm_triton.py")


def stage() -> bool:
    if not os.path.isdir(REF):
        return False
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(REF, f), os.path.join(DST, os.path.basename(f)))
    return True


def available() -> bool:
    return all(os.path.exists(os.path.join(DST, os.path.basename(f))) for f in FILES)


def load(name: str = "vmamba"):
    """Import a FRESH copy of the staged module (each test rebinds names in its own copy)."""
    from oracle.make_golden import _install_stubs
    _install_stubs()
    for p in (DST, os.path.join(HERE, "_ref")):   # the rebuilt selective_scan_cuda_core extension lives in oracle/_ref
        if p not in sys.path:
            sys.path.insert(0, p)
    sys.modules.pop(name, None)
    return importlib.import_module(name)


if __name__ == "__main__":
    print("staged" if stage() else f"reference tree not found at {REF}")
