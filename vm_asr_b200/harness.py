"""Step harness of the hot path (SURVEY.md 8d "Throughput metric", 8e): the SHAPE of one VM-ASR generator step, with every
hot-path operator on this library's kernels and nothing else of the model rebuilt.

    wave --wav2spectro--> (mag, phase) --drop DC bin--> two streams of 17 SS2D cores each (the 34 calls of the config, at the
    config's map sizes, in the generator's order: model/model.py:1103-1227) --add DC bin back--> spectro2wav --> wave_out

Between two SS2D cores the real generator runs in_proj / depthwise conv / LayerNorm / MLP / patch merging on cuBLAS / cuDNN
(out of scope: BASELINE.json north_star keeps them on PyTorch).  Here they are replaced by the cheapest differentiable glue
that produces a map of the next call's shape -- residual add, a one-group normalisation, nearest / average resampling and a
1x1 convolution when the channel count changes -- so that one autograd graph links all 34 calls, the STFT and the iSTFT, and
the two streams interact after every pair as in the reference (``mag = mag + phase; phase = phase + mag``, model.py:1129-1131).
The cores' parameters use the reference's initialisers (vmamba.py:1204-1267).

``TrainStep`` adds what a data-parallel trainer adds (trainer/trainer.py:138-156 has no multi-GPU path at all, README.md:31):
  * all gradients live in ONE flat buffer, cut into buckets; a bucket's NCCL all-reduce is issued on a side stream from the
    autograd hook of its last parameter, so it overlaps the rest of the backward;
  * the multi-period discriminator is not built here; its 41.09 M-float gradient payload is carried as a
    zero buffer of that size, all-reduced in chunks alongside -- stated as synthetic wherever it is reported;
  * fused AdamW on the real parameters.
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import ss2d, stft
from .loss import MultiResolutionSTFTLoss
from .workload import Workload

MPD_PARAMS = 41_090_000  # MultiPeriodDiscriminator(hidden=32), SURVEY.md 8e


class SS2DCoreParams(nn.Module):
    """Parameters of one SS2D core, initialised as SS2D.__initv2__ does (vmamba.py:772-850 with dt_init :1204-1236,
    A_log_init :1241-1255, D_init :1258-1267); d_state 1, 4 directions, dt_rank = ceil(d_model / 16) with d_model = d_inner / 2."""

    def __init__(self, d_inner: int, generator: torch.Generator):
        super().__init__()
        K, N = 4, 1
        R = max(1, math.ceil((d_inner // 2) / 16))
        self.d_inner, self.R = d_inner, R
        bound = 1.0 / math.sqrt(d_inner)
        self.x_proj_weight = nn.Parameter((torch.rand(K, R + 2 * N, d_inner, generator=generator) * 2 - 1) * bound)
        std = R ** -0.5
        self.dt_projs_weight = nn.Parameter((torch.rand(K, d_inner, R, generator=generator) * 2 - 1) * std)
        dt = torch.exp(torch.rand(K, d_inner, generator=generator) * (math.log(0.1) - math.log(0.001)) + math.log(0.001)).clamp(min=1e-4)
        self.dt_projs_bias = nn.Parameter(dt + torch.log(-torch.expm1(-dt)))
        self.A_logs = nn.Parameter(torch.zeros(K * d_inner, N))   # log(arange(1, N + 1)) with N = 1
        self.Ds = nn.Parameter(torch.ones(K * d_inner))

    def tensors(self):
        return (self.x_proj_weight, self.dt_projs_weight, self.dt_projs_bias, self.A_logs, self.Ds)


class VSSBlockParams(nn.Module):
    """What SS2D holds around its core (vmamba.py:853-890): in_proj (d -> 2 d_inner: the x half and the gate z), the depthwise
    conv 3x3, out_norm (LayerNorm) and out_proj; here d = d_inner (the harness keeps its streams at d_inner channels)."""

    def __init__(self, d_inner: int, generator: torch.Generator):
        super().__init__()
        self.core = SS2DCoreParams(d_inner, generator)
        self.in_proj = nn.Linear(d_inner, 2 * d_inner, bias=False)
        self.out_proj = nn.Linear(d_inner, d_inner, bias=False)
        self.conv_weight = nn.Parameter((torch.rand(d_inner, 1, 3, 3, generator=generator) * 2 - 1) / 3.0)
        self.conv_bias = nn.Parameter(torch.zeros(d_inner))
        self.norm_weight = nn.Parameter(torch.ones(d_inner))
        self.norm_bias = nn.Parameter(torch.zeros(d_inner))
        with torch.no_grad():
            for lin in (self.in_proj, self.out_proj):
                bound = 1.0 / math.sqrt(lin.in_features)
                lin.weight.copy_((torch.rand(lin.weight.shape, generator=generator) * 2 - 1) * bound)

    def block_tensors(self):
        c = self.core
        return (self.conv_weight, self.conv_bias, c.x_proj_weight, c.dt_projs_weight, c.dt_projs_bias, c.A_logs, c.Ds,
                self.norm_weight, self.norm_bias)


class HotPathNet(nn.Module):
    """STFT -> 17 pairs of SS2D cores (magnitude stream, phase stream) -> iSTFT, see the module docstring."""

    def __init__(self, wl: Workload, seed: int = 123, pair: bool = True, block: bool = True):
        """``block``: every core call is a whole SS2D module body -- in_proj, then head (conv + SiLU + x_proj) -> core -> tail
        (LayerNorm + gate) on this library's four kernels (``ss2d.ss2d_block_core``), then out_proj -- instead of the bare core
        between normalisations."""
        super().__init__()
        self.wl, self.pair, self.block = wl, pair, block
        gen = torch.Generator().manual_seed(seed)
        assert len(wl.calls) % 2 == 0
        self.steps = [wl.calls[i] for i in range(0, len(wl.calls), 2)]   # calls 2j, 2j + 1: the two streams' same-shape pair
        self.cores = nn.ModuleList()
        self.glue = nn.ModuleList()
        c_prev = 1
        for call in self.steps:
            make = VSSBlockParams if block else SS2DCoreParams
            self.cores.append(nn.ModuleList([make(call.d_inner, gen), make(call.d_inner, gen)]))
            if call.d_inner != c_prev:
                convs = nn.ModuleList([nn.Conv2d(c_prev, call.d_inner, 1), nn.Conv2d(c_prev, call.d_inner, 1)])
            else:
                convs = nn.ModuleList()
            self.glue.append(convs)
            c_prev = call.d_inner
        self.head = nn.ModuleList([nn.Conv2d(c_prev, 1, 1), nn.Conv2d(c_prev, 1, 1)])
        # the glue is initialised from the same generator as the cores (not from the global RNG: two harnesses built in one
        # process are identical), the heads small: the net starts close to the identity STFT -> iSTFT, as a residual generator does
        with torch.no_grad():
            for conv in [c for convs in self.glue for c in convs] + list(self.head):
                bound = 1.0 / math.sqrt(conv.in_channels)
                conv.weight.copy_((torch.rand(conv.weight.shape, generator=gen) * 2 - 1) * bound)
                conv.bias.zero_()
            for conv in self.head:
                conv.weight.mul_(0.05)

    @staticmethod
    def _rms(y):
        """per-clip RMS normalisation: one multi-block reduction and one elementwise pass (F.group_norm with one group runs a
        single CTA per clip: 7 ms of a 30 ms step when it stood here)"""
        y = y.float()
        return y * torch.rsqrt(y.square().mean(dim=(1, 2, 3), keepdim=True) + 1e-6)

    @staticmethod
    def _resample(x, H, W):
        h, w = x.shape[-2:]
        if (h, w) == (H, W):
            return x
        if h >= H and w >= W and h % H == 0 and w % W == 0:
            return F.avg_pool2d(x, (h // H, w // W))
        return F.interpolate(x, size=(H, W), mode="nearest")

    def forward(self, wave: torch.Tensor) -> torch.Tensor:
        wl = self.wl
        Bsz = wave.shape[0]
        mag, phase = stft.wav2spectro(wave, wl.n_fft, wl.hop, wl.win, "log2")   # (B, 1, F, Nf); model.py:424-434
        dc = (mag[..., :1, :], phase[..., :1, :])
        streams = [mag[..., 1:, :], phase[..., 1:, :]]                            # model.py:1110-1111
        residual_mag = streams[0]
        for call, cores, convs in zip(self.steps, self.cores, self.glue):
            xs = []
            for s in range(2):
                x = self._resample(streams[s], call.H, call.W)
                if len(convs):
                    x = convs[s](x)
                xs.append(x.contiguous())
            # VSSBlock: x + SS2D(LN(x)), and SS2D normalises its own output (out_norm) -- vmamba.py:1826-1837, 1527-1531.  The
            # core is cubic in the scale of its input (B, C and u are all linear in it), so the pre-normalisation is what keeps
            # the activations (and the fp16 gradients under autocast) in range, exactly as in the reference.
            xn = [self._rms(x) for x in xs]
            if self.block and call.H % 4 == 0 and call.W % 8 == 0:
                # SS2D.forwardv2 (vmamba.py:1533-1552): in_proj -> [head -> core -> tail on this library] -> out_proj
                xz = [cores[s].in_proj(xn[s].permute(0, 2, 3, 1)) for s in range(2)]          # (B, H, W, 2 d_inner), channel-last
                halves = [t.chunk(2, dim=-1) for t in xz]
                if self.pair:
                    ys = ss2d.ss2d_block_core_pair(halves[0][0], cores[0].block_tensors(), halves[1][0], cores[1].block_tensors(),
                                                   z_a=halves[0][1], z_b=halves[1][1])
                else:
                    ys = [ss2d.ss2d_block_core(halves[s][0], *cores[s].block_tensors(), z=halves[s][1]) for s in range(2)]
                outs = [(xs[s] + cores[s].out_proj(ys[s]).permute(0, 3, 1, 2)) * 0.7071067811865476 for s in range(2)]
                m = 0.5 * (outs[0] + outs[1])
                streams = [m, 0.5 * (outs[1] + m)]
                continue
            if self.block:
                cores = [c.core for c in cores]
            if self.pair and call.H % 4 == 0 and call.W % 4 == 0:
                ys = ss2d.ss2d_core_pair(xn[0], cores[0].tensors(), xn[1], cores[1].tensors())
            else:
                ys = [ss2d.ss2d_core(xn[s], *cores[s].tensors()) for s in range(2)]
            outs = []
            for s in range(2):
                y = ys[s].view(Bsz, call.d_inner, call.H, call.W)
                outs.append((xs[s] + self._rms(y)) * 0.7071067811865476)
            # the streams interact after every pair as in model.py:1129-1131 (mag = mag + phase; phase = phase + mag); AVERAGED
            # here: the reference's blocks renormalise what they add, this glue does not, and 17 plain sums grow like 3^17 --
            # log2-magnitudes of that size overflow exp2 in spectro2wav (or underflow it to an all-zero wave with zero gradient)
            m = 0.5 * (outs[0] + outs[1])
            streams = [m, 0.5 * (outs[1] + m)]
        full_h, full_w = residual_mag.shape[-2:]
        out = [self.head[s](self._resample(streams[s], full_h, full_w)) for s in range(2)]
        mag_out = torch.cat([dc[0], out[0] + residual_mag], dim=-2)              # model.py:1205-1215
        phase_out = torch.cat([dc[1], out[1]], dim=-2)
        wav = stft.spectro2wav(mag_out, phase_out, wl.n_fft, wl.hop, wl.win, "log2")   # model.py:436-445
        return wav[..., : wave.shape[-1]]


class FlatGrads:
    """All parameter gradients as views of one flat fp32 buffer, cut into buckets of ``bucket_floats``; the all-reduce of a
    bucket is issued from the post-accumulate hook of the LAST of its parameters to receive a gradient."""

    def __init__(self, params: List[nn.Parameter], bucket_floats: int = 1 << 20, payload_floats: int = 0, payload_chunks: int = 4):
        self.params = [p for p in params if p.requires_grad]
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.buckets = []   # (lo, hi, pending count)
        off, lo, members = 0, 0, 0
        self._bucket_of = {}
        # buckets in REVERSE registration order: the last layers' gradients are ready first
        for p in reversed(self.params):
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            self._bucket_of[p] = len(self.buckets)
            off += n
            members += 1
            if off - lo >= bucket_floats:
                self.buckets.append([lo, off, members])
                lo, members = off, 0
        if members:
            self.buckets.append([lo, off, members])
        self.payload = torch.zeros(payload_floats, dtype=torch.float32, device=dev) if payload_floats else None
        self.payload_chunks = payload_chunks
        self.comm_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None   # CPU (gloo) in the unit tests
        self._pending = [b[2] for b in self.buckets]
        self._handles = []
        self.enabled = dist.is_initialized() and dist.get_world_size() > 1
        for p in self.params:
            p.register_post_accumulate_grad_hook(self._hook)

    def _on_comm_stream(self):
        import contextlib
        if self.comm_stream is None:
            return contextlib.nullcontext()
        self.comm_stream.wait_stream(torch.cuda.current_stream())
        return torch.cuda.stream(self.comm_stream)

    def arm(self):
        """host-side bookkeeping of a new step (what a CUDA-graph replay cannot do)"""
        self._pending = [b[2] for b in self.buckets]
        self._handles = []

    def zero(self):
        self.flat.zero_()
        self.arm()

    def start_payload(self):
        """The discriminator-sized payload goes out first, in chunks, on the communication stream (it has no dependency on
        this step's backward): it overlaps the whole backward."""
        if not self.enabled or self.payload is None:
            return
        with self._on_comm_stream():
            for chunk in self.payload.chunk(self.payload_chunks):
                self._handles.append(dist.all_reduce(chunk, async_op=True))

    def _hook(self, p):
        if not self.enabled:
            return
        b = self._bucket_of[p]
        self._pending[b] -= 1
        if self._pending[b] == 0:
            lo, hi, _ = self.buckets[b]
            with self._on_comm_stream():
                self._handles.append(dist.all_reduce(self.flat[lo:hi], async_op=True))

    def finish(self, world: int):
        if not self.enabled:
            return
        # buckets with a parameter that received no gradient this step (the reference has such parameters: the phase decoder
        # is never run with CONCAT_SKIP, model/model.py:1186-1187) were never triggered: their zeros are reduced now
        for b, left in enumerate(self._pending):
            if left > 0:
                lo, hi, _ = self.buckets[b]
                with self._on_comm_stream():
                    self._handles.append(dist.all_reduce(self.flat[lo:hi], async_op=True))
                self._pending[b] = 0
        for h in self._handles:
            h.wait()
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self.flat.div_(world)


class TrainStep:
    """forward -> L1 + multi-resolution STFT loss -> backward (bucketed all-reduce overlapped) -> fused AdamW.  ``comm=False`` runs the
    same step without any collective (to state the exposed communication time).

    ``capture()`` records forward + backward (gradients into the flat buffer) as ONE CUDA graph over static input buffers: the
    step is then a copy of the batch into those buffers, a graph replay, the all-reduce and the optimizer.  The eager step is
    host-bound (a few hundred small PyTorch launches of glue per step); the graph is what the device can do.  In graph mode
    the discriminator-sized payload still overlaps the whole replay (it is issued first, on the communication stream); the
    real gradients (a few MB) are reduced after the replay."""

    def __init__(self, wl: Workload, device, world: int = 1, pair: bool = True, mpd_payload: bool = True, lr: float = 1e-3,
                 amp: bool = True):
        self.wl, self.world, self.device = wl, world, device
        # the reference trains under fp16 autocast with a GradScaler (config.py:217, trainer/trainer.py:106-107, 138): the two small
        # einsums of the core and the glue then run on tensor cores, the scan is forced to fp32 (vmamba.py:1487-1491) as here
        self.amp = amp
        self.scaler = torch.amp.GradScaler("cuda", init_scale=1024.0, enabled=amp)
        self.net = HotPathNet(wl, pair=pair).to(device)
        if world > 1:
            for p in self.net.parameters():
                dist.broadcast(p.data, 0)
        self.grads = FlatGrads(list(self.net.parameters()), payload_floats=MPD_PARAMS if (mpd_payload and world > 1) else 0)
        self.opt = torch.optim.AdamW(self.net.parameters(), lr=lr, weight_decay=0.0, fused=True)   # config.py:131-154
        self.n_params = sum(p.numel() for p in self.net.parameters())
        self.graph = None
        # generator loss of the reference step without the adversarial terms (trainer/trainer.py:318-333): L1 + the
        # multi-resolution STFT loss with factors 0.5 / 0.5 (config.py:176-191), on this library's STFT kernel
        self.stft_loss = MultiResolutionSTFTLoss(factor_sc=0.5, factor_mag=0.5)

    def loss_fn(self, out, target):
        sc, mag = self.stft_loss(out.flatten(0, -2), target.flatten(0, -2))
        return (out - target).abs().mean() + sc + mag

    def _fwd_bwd(self, wave_in, wave_target):
        self.grads.flat.zero_()
        with torch.autocast("cuda", dtype=torch.float16, enabled=self.amp):
            out = self.net(wave_in)
        loss = self.loss_fn(out.float(), wave_target)
        self.scaler.scale(loss).backward()
        return loss

    def capture(self, wave_in: torch.Tensor, wave_target: torch.Tensor):
        g = self.grads
        was, g.enabled = g.enabled, False        # no collectives from the autograd hooks while capturing
        self.static_in, self.static_tgt = torch.empty_like(wave_in), torch.empty_like(wave_target)
        self.static_in.copy_(wave_in)
        self.static_tgt.copy_(wave_target)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self._fwd_bwd(self.static_in, self.static_tgt)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                self.static_loss = self._fwd_bwd(self.static_in, self.static_tgt)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = graph
        g.enabled = was

    def __call__(self, wave_in: torch.Tensor, wave_target: torch.Tensor, comm: bool = True) -> torch.Tensor:
        g = self.grads
        was = g.enabled
        g.enabled = was and comm
        g.arm()
        if self.graph is not None:
            self.static_in.copy_(wave_in, non_blocking=True)
            self.static_tgt.copy_(wave_target, non_blocking=True)
            g.start_payload()
            hooks, g.enabled = g.enabled, False
            self.graph.replay()
            g.enabled = hooks
            loss = self.static_loss
        else:
            g.start_payload()
            loss = self._fwd_bwd(wave_in, wave_target)
        g.finish(self.world)     # reduces the buckets the hooks did not (all of them in graph mode)
        g.enabled = was
        self.scaler.step(self.opt)     # unscales inside the fused optimizer, skips the step on inf / nan (identical on every rank
        self.scaler.update()           # after the all-reduce), no host synchronisation
        return loss

    @torch.no_grad()
    def infer(self, wave_in: torch.Tensor) -> torch.Tensor:
        with torch.autocast("cuda", dtype=torch.float16, enabled=self.amp):
            return self.net(wave_in).float()


def synthetic_batch(wl: Workload, device, rank: int = 0, pinned: bool = False):
    """wave_target = 0.1 randn(B, 1, T); the input is a crudely band-limited copy (box filter: the data loader's
    resampling chain is out of scope)."""
    gen = torch.Generator().manual_seed(1000 + rank)
    target = 0.1 * torch.randn(wl.batch, 1, wl.T, generator=gen)
    k = 6
    inp = F.avg_pool1d(F.pad(target, (k // 2, k - 1 - k // 2), mode="replicate"), k, stride=1)
    if pinned:
        return inp.pin_memory(), target.pin_memory()
    return inp.to(device), target.to(device)
