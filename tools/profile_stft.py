#!/usr/bin/env python
"""Run the STFT entry points a few times (for ncu):  python tools/profile_stft.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vm_asr_b200 import stft
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
w = (0.1 * torch.randn(B, 1, 122640, device="cuda")).requires_grad_()
for _ in range(3):
    w.grad = None
    m, p = stft.wav2spectro(w, 1024, 240, 1024, "log2")
    back = stft.spectro2wav(m, p, 1024, 240, 1024, "log2")
    back.square().sum().backward()
torch.cuda.synchronize()
print("done")
