#!/usr/bin/env python
"""One fused SS2D core forward + backward on a small map (with and without the fused LayerNorm / gate tail), and the STFT entry points (for compute-sanitizer / ncu):
    python tools/profile_fused.py B C H W"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vm_asr_b200 import ss2d, stft, loss
B, C, H, W = (int(v) for v in sys.argv[1:5])
dev = torch.device("cuda")
torch.manual_seed(0)
x = torch.randn(B, C, H, W, device=dev, requires_grad=True)
R = 1
prm = [torch.randn(4, R + 2, C, device=dev) * 0.3, torch.randn(4, C, R, device=dev) * 0.3, torch.rand(4, C, device=dev) * 0.5,
       torch.zeros(4 * C, 1, device=dev), torch.ones(4 * C, device=dev)]
prm = [p.requires_grad_() for p in prm]
y = ss2d.ss2d_core(x, *prm, fused=True)
y.square().mean().backward()
# the block's tail fused into the merge (vmasr_outnorm_gate_fwd / _bwd), fp32 and fp16 gate; with W % 8 == 0 only
if W % 8 == 0:
    for dt in (torch.float32, torch.float16):
        gam, bet = torch.ones(C, device=dev, requires_grad=True), torch.zeros(C, device=dev, requires_grad=True)
        z = torch.randn(B, H, W, C, device=dev).to(dt).requires_grad_()
        o = ss2d.ss2d_core_out(x, *prm, gam, bet, z=z)
        o.float().square().mean().backward()
    # the whole block between in_proj and out_proj: head (conv + SiLU + x_proj) -> core -> tail
    xz = torch.randn(B, H, W, 2 * C, device=dev, requires_grad=True)
    cw, cb = (0.3 * torch.randn(C, 1, 3, 3, device=dev)).requires_grad_(), torch.zeros(C, device=dev, requires_grad=True)
    xa, za = xz.chunk(2, dim=-1)
    ob = ss2d.ss2d_block_core(xa, cw, cb, *prm, gam, bet, z=za)
    ob.square().mean().backward()
ya, yb = ss2d.ss2d_core_pair(x.detach(), [p.detach() for p in prm], x.detach() * 0.5, [p.detach() for p in prm])
w = (0.1 * torch.randn(2, 1, 240 * 20, device=dev)).requires_grad_()
m, p = stft.wav2spectro(w, 1024, 240, 1024, "log2")
back = stft.spectro2wav(m, p, 1024, 240, 1024, "log2")
sc, mg = loss.MultiResolutionSTFTLoss()(back.flatten(0, -2), w.detach().flatten(0, -2) * 0.9)
(sc + mg + back.square().mean()).backward()
torch.cuda.synchronize()
print("done")
