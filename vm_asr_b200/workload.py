"""Hot-path workload tables of the BASELINE.json configs.

Every generator forward of VM-ASR makes 34 SS2D calls (17 VSS blocks x 2 streams), 2 STFTs and 1 iSTFT;
the backward repeats each SS2D call once and adds the iSTFT backward.  The shapes below were read off the
reference generator instantiated with each config's constants (d_state 1, ssm_ratio 2, 4 scan directions:
scan channels D = 4 * d_inner, B/C groups G = 4).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple


@dataclass(frozen=True)
class SS2DCall:
    d_inner: int
    H: int
    W: int

    @property
    def L(self):
        return self.H * self.W

    @property
    def D(self):
        return 4 * self.d_inner


@dataclass(frozen=True)
class Workload:
    name: str
    yaml: str
    batch: int
    T: int            # samples per clip
    n_fft: int
    hop: int
    win: int
    sr: int
    calls: Tuple[SS2DCall, ...]   # in forward order

    @property
    def clip_seconds(self):
        return 2.555  # DATA.SEGMENT (config.py:50)

    def scan_elements(self):
        return sum(self.batch * c.D * c.L for c in self.calls)

    def scan_bytes(self, itemsize=4, G=4, N=1):
        """(forward, backward) algorithmic bytes of all scan calls: s*(3BDL + 2BGNL), s*(5BDL + 4BGNL)."""
        fwd = sum(itemsize * (3 * self.batch * c.D * c.L + 2 * self.batch * G * N * c.L) for c in self.calls)
        bwd = sum(itemsize * (5 * self.batch * c.D * c.L + 4 * self.batch * G * N * c.L) for c in self.calls)
        return fwd, bwd


def _order(dims0: int, H0: int, W0: int) -> Tuple[SS2DCall, ...]:
    """Forward order of the 34 calls for base width ``dims0`` and full-resolution grid H0 x W0 (after dropping
    the DC bin): patch embed halves the grid twice; 4 encoder stages (2 blocks x 2 streams each) halve it
    further; the decoder mirrors stages 3, 2, 1 (4 calls each) and ends with 2 calls at the first stage's size;
    the output layer runs 2 calls at half width / double resolution and 2 calls with d_inner 2 at full
    resolution.  Reproduces the per-config inventory of SURVEY.md 8a."""
    d = dims0 * 2  # d_inner of the first encoder stage (ssm_ratio 2)
    h, w = H0 // 4, W0 // 4
    enc = []
    for stage in range(4):
        enc += [SS2DCall(d << stage, h >> stage, w >> stage)] * 4
    dec = []
    for stage in (3, 2, 1):
        dec += [SS2DCall(d << stage, h >> stage, w >> stage)] * 4
    dec += [SS2DCall(d, h, w)] * 2
    out = [SS2DCall(d // 2, h * 2, w * 2)] * 2 + [SS2DCall(2, h * 4, w * 4)] * 2
    calls = tuple(enc + dec + out)
    assert len(calls) == 34
    return calls


WORKLOADS = {
    "vm_asr_16k": Workload("vm_asr_16k", "configs/vm_asr_16k.yaml", 4, 40880, 1024, 80, 1024, 16000, _order(16, 512, 512)),
    "vm_asr_48k_MPD": Workload("vm_asr_48k_MPD", "configs/vm_asr_48k_MPD.yaml", 4, 122640, 1024, 240, 1024, 48000,
                               _order(16, 512, 512)),
    "vm_asr_48k_16k_nfft2048": Workload("vm_asr_48k_16k_nfft2048", "configs/vm_asr_48k_16k_nfft2048.yaml", 8, 122640, 2048,
                                        240, 1024, 48000, _order(16, 1024, 512)),
    "vm_asr_48k_16k_MPD_VSSM32": Workload("vm_asr_48k_16k_MPD_VSSM32", "configs/vm_asr_48k_16k_MPD_VSSM32.yaml", 8, 122640,
                                          1024, 240, 1024, 48000, _order(32, 512, 512)),
}


def distinct_shapes(wl: Workload) -> List[Tuple[SS2DCall, int]]:
    seen = {}
    for c in wl.calls:
        seen[c] = seen.get(c, 0) + 1
    return list(seen.items())
