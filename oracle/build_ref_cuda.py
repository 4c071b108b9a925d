#!/usr/bin/env python
"""TEST INFRASTRUCTURE, not product code.  Builds the REFERENCE's own selective-scan CUDA extension
(`selective_scan_cuda_core`: kernels/selective_scan/csrc/selective_scan/cus/{selective_scan.cpp, selective_scan_core_fwd.cu,
selective_scan_core_bwd.cu}, flags of kernels/selective_scan/setup.py:115-131) for sm_100a, from the sources where they lie
under /root/reference, into oracle/_ref/ (git-ignored; it travels to the GPU box with the snapshot).  Nothing is copied into
the repository.  Only tests/test_vs_reference_cuda_gpu.py loads the result: it checks this library's kernels against the
reference's on the same GPU and times both.  The reference ships SASS for sm_70/80/90 only (setup.py:61-67), so its wheel cannot
run on B200; the sources compile unchanged for sm_100a.

    python oracle/build_ref_cuda.py        # a few minutes of nvcc; skipped when the reference tree is absent or it is built"""
import glob
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/kernels/selective_scan/csrc/selective_scan"
OUT = os.path.join(HERE, "_ref")


def main() -> int:
    if glob.glob(os.path.join(OUT, "selective_scan_cuda_core*.so")):
        print("oracle/_ref: reference extension already built")
        return 0
    if not os.path.isdir(REF):
        print("oracle/_ref: reference sources not present, nothing built")
        return 0
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load
    load(
        name="selective_scan_cuda_core",
        sources=[os.path.join(REF, "cus", f) for f in ("selective_scan.cpp", "selective_scan_core_fwd.cu", "selective_scan_core_bwd.cu")],
        extra_include_paths=[REF],
        extra_cflags=["-O3", "-std=c++17"],
        extra_cuda_cflags=["-O3", "-std=c++17", "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
                           "-U__CUDA_NO_BFLOAT16_OPERATORS__", "-U__CUDA_NO_BFLOAT16_CONVERSIONS__",
                           "-U__CUDA_NO_BFLOAT162_OPERATORS__", "-U__CUDA_NO_BFLOAT162_CONVERSIONS__",
                           "--expt-relaxed-constexpr", "--expt-extended-lambda", "--use_fast_math", "-lineinfo"],
        build_directory=OUT,
        verbose=False,
    )
    print("oracle/_ref: built", glob.glob(os.path.join(OUT, "*.so")))
    return 0


if __name__ == "__main__":
    sys.exit(main())
