"""Batched inference over long clips: drop-ins for ``utils/post_processing.py`` ``unfold_audio`` /
``fold_audio`` (:4-33) and the segment loop of ``trainer/tester.py`` (:106-140) / ``trainer/inferencer.py`` (:84-103).

The reference cuts a long clip into overlapping segments, runs the generator ONE SEGMENT AT A TIME (batch 1), averages the
overlaps with a Python loop over segments, and times the whole thing with ``time.time()`` without synchronising the device.
Here the segments of all clips go through the generator as one batch (every kernel on the hot path is independent across
the batch index, so this is the "embarrassingly parallel" inference of the north star), the cross-fade is two index
assignments and one division on the device -- deterministic: even and odd segments never overlap among themselves -- and
the real-time factor comes from CUDA events."""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch


def unfold_audio(audio: torch.Tensor, segment_length: int, overlap: int) -> torch.Tensor:
    """(B, C, T) -> (B, C, n_segments, segment_length), a view (post_processing.py:4-9); a tail shorter than a step is
    dropped, as there."""
    return audio.unfold(dimension=-1, size=segment_length, step=segment_length - overlap)


def fold_audio(segments: torch.Tensor, total_length: int, segment_length: int, overlap: int) -> torch.Tensor:
    """(B, C, n_segments, segment_length) -> (B, C, total_length): overlap-average of post_processing.py:12-33 (sum of the
    segments divided by the number of segments covering each sample, 1 where none does), without the loop over segments."""
    step = segment_length - overlap
    if step <= 0:
        raise ValueError("overlap must be smaller than the segment length")
    B, C, n, seg = segments.shape
    if seg != segment_length:
        raise ValueError("segments have the wrong length")
    dev = segments.device
    flat = segments.reshape(B * C, n, seg)
    t = torch.arange(seg, device=dev)
    out = torch.zeros(B * C, total_length, dtype=segments.dtype, device=dev)
    count = torch.zeros(total_length, dtype=segments.dtype, device=dev)
    # segments i, i + k, i + 2k, ... with k * step >= segment_length are disjoint: each group is ONE index assignment
    k = -(-segment_length // step)
    for r in range(min(k, n)):
        idx = (torch.arange(r, n, k, device=dev)[:, None] * step + t[None, :]).reshape(-1)
        keep = idx < total_length
        part = torch.zeros_like(out)
        part[:, idx[keep]] = flat[:, r::k].reshape(B * C, -1)[:, keep]
        out += part
        c = torch.zeros_like(count)
        c[idx[keep]] = 1
        count += c
    count[count == 0] = 1
    return (out / count).view(B, C, total_length)


@torch.no_grad()
def infer_long(generator: Callable[..., torch.Tensor], wave_input: torch.Tensor, highcut=None, segment_length: int = 122640,
               overlap: int = 2000, sample_rate: int = 48000, max_batch: Optional[int] = None,
               pad_length: int = 0) -> Tuple[torch.Tensor, dict]:
    """``wave_input`` (B, C, T) on the device.  Clips no longer than a segment go through ``generator`` as they are
    (tester.py:92-105); longer ones are unfolded, run as ONE batch of B * n_segments segments (``max_batch`` bounds the
    batch), and folded back.  Returns the output and ``{"rtf", "rtf_reciprocal", "device_seconds", "segments"}`` with the time
    measured by CUDA events around the whole computation (the reference's RTF, tester.py:97-100, 133-136, uses the host
    clock without a device synchronisation)."""
    B, C, T = wave_input.shape
    call = (lambda w: generator(w, highcut)) if highcut is not None else generator
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if T <= segment_length:
        out, n_seg = call(wave_input), 1
    else:
        segs = unfold_audio(wave_input, segment_length, overlap)                 # (B, C, n, seg) view
        n_seg = segs.shape[2]
        batch = segs.permute(0, 2, 1, 3).reshape(B * n_seg, C, segment_length)   # every segment of every clip
        if max_batch is None or batch.shape[0] <= max_batch:
            done = call(batch)
        else:
            done = torch.cat([call(batch[i:i + max_batch]) for i in range(0, batch.shape[0], max_batch)], dim=0)
        done = done.reshape(B, n_seg, C, segment_length).permute(0, 2, 1, 3)
        out = fold_audio(done, T, segment_length, overlap)
    e1.record()
    e1.synchronize()
    seconds = e0.elapsed_time(e1) * 1e-3
    audio_seconds = B * (T - pad_length) / float(sample_rate)
    rtf = seconds / audio_seconds
    return out, {"rtf": rtf, "rtf_reciprocal": 1.0 / rtf, "device_seconds": seconds, "segments": B * n_seg}
