"""ctypes binding of ``libvmasr_b200.so`` (C ABI: ``include/vmasr_b200.h``).

There is no fallback: if the CUDA library has not been built, or a tensor is not on a CUDA device, the
operators raise.  Build with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C vm_asr_b200/csrc``.
"""
from __future__ import annotations

import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# VMASR_B200_LIBRARY: developer switch to load another build of the same library (e.g. `make TUNING=1 OUT=../lib_tuning`)
_LIB_PATH = os.environ.get("VMASR_B200_LIBRARY") or os.path.join(_HERE, "lib", "libvmasr_b200.so")
_lib = None
_lock = threading.Lock()

VMASR_F32, VMASR_F16, VMASR_BF16 = 0, 1, 2
SCAN_REVERSE, SCAN_ACCUMULATE, SCAN_ADD, SCAN_DBDC_STORE = 1, 2, 4, 8
SCAN_MAX_GROUP = 8
SS2D_DYT_GIVEN = 1
ABI_VERSION = 5
SCAN_CHUNK = 2048
DTYPE_CODE = {torch.float32: VMASR_F32, torch.float16: VMASR_F16, torch.bfloat16: VMASR_BF16}

_vp, _i32, _i64, _u64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64


class ScanParams(ctypes.Structure):
    """Mirror of ``vmasr_scan_params`` (include/vmasr_b200.h)."""

    _fields_ = (
        [(n, _vp) for n in ("u", "delta", "A", "B", "C", "D", "delta_bias", "out", "x", "dout", "du", "ddelta",
                            "dA", "dB", "dC", "dD", "ddelta_bias", "workspace")]
        + [("workspace_bytes", _u64)]
        + [(n, _i32) for n in ("batch", "dim", "seqlen", "dstate", "ngroups")]
        + [(n, _i64) for n in (
            "u_batch_stride", "u_d_stride", "delta_batch_stride", "delta_d_stride", "A_d_stride", "A_dstate_stride",
            "B_batch_stride", "B_group_stride", "B_dstate_stride", "C_batch_stride", "C_group_stride", "C_dstate_stride",
            "out_batch_stride", "out_d_stride", "dout_batch_stride", "dout_d_stride", "du_batch_stride", "du_d_stride",
            "ddelta_batch_stride", "ddelta_d_stride")]
        + [(n, _i32) for n in ("io_dtype", "delta_softplus", "device", "flags")]
        + [("stream", _vp)]
        + [(n, _vp) for n in ("dt_rows", "dt_weight", "d_dt_rows", "d_dt_weight")]
        + [(n, _i64) for n in ("dt_rows_batch_stride", "dt_rows_row_stride", "dt_weight_d_stride", "dB_batch_stride",
                               "dC_batch_stride")]
        + [("dt_rank", _i32), ("reserved0", _i32)]
        + [("zero_ptr", _vp), ("zero_bytes", _u64)]
    )


class SS2DParams(ctypes.Structure):
    """Mirror of ``vmasr_ss2d_params`` (include/vmasr_b200.h)."""

    _fields_ = [
        ("x", _vp), ("xT", _vp),
        ("delta", _vp * 4), ("delta_batch_stride", _i64 * 4), ("delta_d_stride", _i64 * 4),
        ("B", _vp * 4), ("C", _vp * 4), ("B_batch_stride", _i64 * 4), ("C_batch_stride", _i64 * 4),
        ("A", _vp), ("D", _vp), ("delta_bias", _vp),
        ("y", _vp), ("planes", _vp), ("states", _vp),
        ("dy", _vp), ("dyT", _vp), ("dx", _vp),
        ("ddelta", _vp * 4), ("ddelta_batch_stride", _i64 * 4), ("ddelta_d_stride", _i64 * 4),
        ("dA", _vp), ("dB", _vp), ("dC", _vp), ("dD", _vp), ("ddelta_bias", _vp),
        ("workspace", _vp), ("workspace_bytes", _u64),
        ("batch", _i32), ("channels", _i32), ("H", _i32), ("W", _i32),
        ("delta_softplus", _i32), ("device", _i32),
        ("stream", _vp),
        ("x_dbl", _vp * 4), ("x_dbl_batch_stride", _i64 * 4), ("x_dbl_row_stride", _i64 * 4),
        ("dt_weight", _vp), ("d_x_dbl", _vp * 4), ("d_dt_weight", _vp),
        ("dt_rank", _i32), ("flags", _i32),
    ]


class OutNormParams(ctypes.Structure):
    """Mirror of ``vmasr_outnorm_params`` (include/vmasr_b200.h)."""

    _fields_ = (
        [(n, _vp) for n in ("p_rm", "p_cm", "gamma", "beta", "z", "out", "y", "stats", "dout", "dy", "dz", "dgb_partial")]
        + [("eps", ctypes.c_float)]
        + [(n, _i32) for n in ("batch", "channels", "H", "W", "io_dtype", "z_silu", "device")]
        + [("stream", _vp)]
    )


class DwConvParams(ctypes.Structure):
    """Mirror of ``vmasr_dwconv_params`` (include/vmasr_b200.h)."""

    _fields_ = (
        [(n, _vp) for n in ("xin", "weight", "bias", "x", "xT", "dx", "dxT", "dxin", "dwb_partial")]
        + [("xin_pos_stride", _i64)]
        + [(n, _i32) for n in ("batch", "channels", "H", "W", "io_dtype", "device")]
        + [("stream", _vp)]
        + [(n, _vp) for n in ("x_proj_weight", "x_proj_bias", "x_dbl_rm", "x_dbl_cm", "d_x_dbl_rm", "d_x_dbl_cm", "d_x_proj_weight_partial")]
        + [("x_proj_rows", _i32), ("reserved0", _i32)]
    )


EXPORTS = {
    "vmasr_abi_version": (ctypes.c_int, []),
    "vmasr_last_error": (ctypes.c_char_p, []),
    "vmasr_scan_workspace_bytes": (_u64, [ctypes.c_int] * 4),
    "vmasr_scan_plan": (ctypes.c_int, [ctypes.POINTER(ScanParams), ctypes.c_int, ctypes.POINTER(ctypes.c_int32)]),
    "vmasr_scan_fwd": (ctypes.c_int, [ctypes.POINTER(ScanParams)]),
    "vmasr_scan_bwd": (ctypes.c_int, [ctypes.POINTER(ScanParams)]),
    "vmasr_scan_fwd_grouped": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ScanParams)]),
    "vmasr_scan_bwd_grouped": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ScanParams)]),
    "vmasr_ss2d_workspace_bytes": (_u64, [ctypes.c_int] * 4),
    "vmasr_ss2d_core_fwd": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(SS2DParams)]),
    "vmasr_ss2d_core_bwd": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(SS2DParams)]),
    "vmasr_map_transpose": (ctypes.c_int, [_vp, _vp, _i64, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp]),
    "vmasr_map_merge2": (ctypes.c_int, [_vp, _vp, _vp, _i64, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp]),
    "vmasr_outnorm_patches": (_i64, [ctypes.c_int] * 4),
    "vmasr_outnorm_gate_fwd": (ctypes.c_int, [ctypes.POINTER(OutNormParams)]),
    "vmasr_outnorm_gate_bwd": (ctypes.c_int, [ctypes.POINTER(OutNormParams)]),
    "vmasr_dwconv_patches": (_i64, [ctypes.c_int] * 4),
    "vmasr_dwconv_channel_blocks": (ctypes.c_int, [ctypes.c_int] * 4),
    "vmasr_dwconv_silu_fwd": (ctypes.c_int, [ctypes.POINTER(DwConvParams)]),
    "vmasr_dwconv_silu_bwd": (ctypes.c_int, [ctypes.POINTER(DwConvParams)]),
    "vmasr_cross_scan": (ctypes.c_int, [_vp, _vp] + [ctypes.c_int] * 6 + [_vp]),
    "vmasr_cross_merge": (ctypes.c_int, [_vp, _vp] + [ctypes.c_int] * 6 + [_vp]),
    "vmasr_cross_scan_1b1": (ctypes.c_int, [_vp, _vp] + [ctypes.c_int] * 6 + [_vp]),
    "vmasr_cross_merge_1b1": (ctypes.c_int, [_vp, _vp] + [ctypes.c_int] * 6 + [_vp]),
    "vmasr_stft_scratch_floats": (_u64, [ctypes.c_int] * 3),
    "vmasr_stft_fwd": (ctypes.c_int, [_vp, _vp, _vp] + [ctypes.c_int] * 6 + [_vp]),
    "vmasr_stft_bwd": (ctypes.c_int, [_vp] * 5 + [ctypes.c_int] * 6 + [_vp]),
    "vmasr_stft_mag_fwd": (ctypes.c_int, [_vp, _vp] + [ctypes.c_int] * 6 + [ctypes.c_float, ctypes.c_int, _vp]),
    "vmasr_stft_mag_bwd": (ctypes.c_int, [_vp] * 4 + [ctypes.c_int] * 6 + [ctypes.c_float, ctypes.c_int, _vp]),
    "vmasr_istft_fwd": (ctypes.c_int, [_vp] * 4 + [ctypes.c_int] * 6 + [_vp]),
    "vmasr_istft_bwd": (ctypes.c_int, [_vp] * 5 + [ctypes.c_int] * 6 + [_vp]),
}


def library_path() -> str:
    return _LIB_PATH


def load_library():
    """Load the shared library once; raise (never fall back) when it is missing."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(_LIB_PATH):
                    raise RuntimeError(
                        f"vmasr_b200: CUDA library not built ({_LIB_PATH} missing). Run `make -C vm_asr_b200/csrc` "
                        "(needs nvcc, sm_100a). There is no CPU or PyTorch fallback for these operators.")
                lib = ctypes.CDLL(_LIB_PATH)
                for name, (res, args) in EXPORTS.items():
                    fn = getattr(lib, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = lib
    return _lib


def check(rc: int):
    if rc != 0:
        msg = load_library().vmasr_last_error().decode("utf-8", "replace")
        raise RuntimeError(msg or f"vmasr_b200 call failed with code {rc}")


def require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (vmasr_b200 has no CPU path)")


def current_stream_ptr(device: torch.device) -> int:
    idx = device.index if device.index is not None else torch.cuda.current_device()
    return torch._C._cuda_getCurrentRawStream(idx)


# ---- carry workspace of the scan ------------------------------------------------------------------------------------
# One per (device, stream, "is this stream being captured into a CUDA graph"): zero-filled once, recycled by the kernels
# (epoch tags), so two scans that run one after the other on a stream share it and nothing is cleared between launches.
#   * a buffer that has been handed out is NEVER freed: when a larger one is needed the old one is retired, not released,
#     because its address may be baked into a captured graph;
#   * calls made while the stream is capturing get a workspace of their own (allocated inside the capture, i.e. from the
#     graph's pool; its zero-fill is a node of the graph), so replaying the graph -- on whatever stream -- never shares a
#     workspace with eager calls;
#   * a buffer is cut into regions (one per problem of a grouped launch) in one way only: the kernels keep a header at the
#     start of every region, so each layout has a buffer of its own;
#   * scans that may run CONCURRENTLY (two streams, two graphs replayed side by side) must not share one: give each its own
#     stream, or capture each graph after `reset_capture_workspaces()`.
_workspaces = {}
_retired = []


def scan_workspace(device: torch.device, nbytes: int, layout=None) -> torch.Tensor:
    """``layout``: how the caller cuts the buffer into regions (grouped launches, the fused SS2D core).  Every region starts
    with its own header, so a buffer must only ever be cut ONE way: each layout gets a buffer of its own."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, torch._C._cuda_getCurrentRawStream(idx), torch.cuda.is_current_stream_capturing(), layout)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        size = max(int(nbytes), 1 << 20)
        if ws is not None:
            size = max(size, 2 * ws.numel())
            _retired.append(ws)
        ws = torch.zeros(size, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def reset_capture_workspaces():
    """Forget (not free) the workspaces handed to captured calls: the next capture allocates its own."""
    for key in [k for k in _workspaces if k[2]]:
        _retired.append(_workspaces.pop(key))
