"""ORACLE (test infrastructure, not product code): restatement of ``utils/post_processing.py`` (:4-33) -- the overlapping
segment cut and the loop that averages the overlaps -- for the CPU tests of ``vm_asr_b200/segment.py``."""
import torch


def unfold_audio(audio, segment_length, overlap):
    return audio.unfold(dimension=-1, size=segment_length, step=segment_length - overlap)   # post_processing.py:6-8


def fold_audio(segments, total_length, segment_length, overlap):
    step = segment_length - overlap                                                          # :14
    batch_size, channels, num_segments, _ = segments.size()
    reconstructed = torch.zeros(batch_size, channels, total_length, dtype=segments.dtype)    # :20-21
    count = torch.zeros(batch_size, channels, total_length, dtype=segments.dtype)
    for i in range(num_segments):                                                            # :24-28
        start = i * step
        end = start + segment_length
        reconstructed[:, :, start:end] += segments[:, :, i]
        count[:, :, start:end] += 1
    count[count == 0] = 1                                                                    # :31
    return reconstructed / count
