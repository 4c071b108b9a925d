"""vm_asr_b200 -- the SS2D + STFT hot path of VM-ASR on B200 (sm_100a).

Python operator surface (the reference is Python) over the C ABI of ``include/vmasr_b200.h``; the CUDA library is
``vm_asr_b200/lib/libvmasr_b200.so``, built from ``vm_asr_b200/csrc``.  No CPU path, no fallback."""
from ._lib import library_path, load_library  # noqa: F401
