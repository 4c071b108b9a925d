"""ORACLE (test infrastructure, not product code): ctypes loader for ``oracle/scan_ref.c``.

Builds ``oracle/_build/libvmasr_oracle.so`` on first use (gcc, < 1 s).  Used by the parity tests at
sizes where the torch restatement in ``ss2d_ref.py`` would take minutes.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    return os.path.join(_HERE, "_build", "libvmasr_oracle.so")


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libvmasr_oracle.so")
        src = os.path.join(_HERE, "scan_ref.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build()
        _LIB = ctypes.CDLL(path)
    return _LIB


def _f32(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def scan_fwd(u, delta, A, Bm, Cm, Dv=None, bias=None, softplus=False, chunk=0):
    """float64 results: out (B,D,L), last (B,D,N), chunk_state (B,D,n_chunks,2N) or None."""
    u, delta, A, Bm, Cm, Dv, bias = map(_f32, (u, delta, A, Bm, Cm, Dv, bias))
    B, D, L = u.shape
    N = A.shape[1]
    G = Bm.shape[1]
    out = np.empty((B, D, L), np.float64)
    last = np.empty((B, D, N), np.float64)
    cs = np.empty((B, D, (L + chunk - 1) // chunk, 2 * N), np.float64) if chunk else None
    lib().vmasr_ref_scan_fwd(_p(u), _p(delta), _p(A), _p(Bm), _p(Cm), _p(Dv), _p(bias), int(bool(softplus)),
                             B, D, L, N, G, _p(out), _p(last), _p(cs), int(chunk))
    return out, last, cs


def scan_bwd(u, delta, A, Bm, Cm, Dv, bias, softplus, dout):
    """float64 gradients (du, ddelta, dA, dB, dC, dD, dbias)."""
    u, delta, A, Bm, Cm, Dv, bias, dout = map(_f32, (u, delta, A, Bm, Cm, Dv, bias, dout))
    B, D, L = u.shape
    N = A.shape[1]
    G = Bm.shape[1]
    du = np.empty((B, D, L), np.float64)
    dd = np.empty((B, D, L), np.float64)
    dA = np.empty((D, N), np.float64)
    dB = np.empty((B, G, N, L), np.float64)
    dC = np.empty((B, G, N, L), np.float64)
    dD = np.empty((D,), np.float64) if Dv is not None else None
    db = np.empty((D,), np.float64) if bias is not None else None
    lib().vmasr_ref_scan_bwd(_p(u), _p(delta), _p(A), _p(Bm), _p(Cm), _p(Dv), _p(bias), int(bool(softplus)),
                             _p(dout), B, D, L, N, G, _p(du), _p(dd), _p(dA), _p(dB), _p(dC), _p(dD), _p(db))
    return du, dd, dA, dB, dC, dD, db


def cross_scan(x):
    x = _f32(x)
    B, C, H, W = x.shape
    xs = np.empty((B, 4, C, H * W), np.float32)
    lib().vmasr_ref_cross_scan(_p(x), _p(xs), B, C, H, W)
    return xs


def cross_merge(ys, H, W):
    ys = _f32(ys)
    B, K, C = ys.shape[:3]
    y = np.empty((B, C, H * W), np.float32)
    lib().vmasr_ref_cross_merge(_p(ys), _p(y), B, C, H, W)
    return y
