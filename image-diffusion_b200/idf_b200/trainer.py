"""One optimizer step of the reference's UNet trainer (trainers/diffusion_trainer.py:141-187) as a fixed kernel
sequence, captured into a CUDA graph:

  KL reparametrisation + add_noise -> Unet forward -> MSE loss + d(eps) -> Unet backward (flat fp32 gradients)
  -> [data parallel: bucketed NCCL all-reduce overlapped with the rest of the backward] -> global-norm clip -> Adam
  -> re-pack of the bf16 operand copies of the weights.

Parameters live in ONE flat fp32 buffer laid out like the engine's gradient buffer (the module's nn.Parameters are
views into it, so state_dict / checkpoints are unchanged); Adam and the norm reduction are single launches over it.
The random draws of a step (noise, timesteps, class-drop mask, reparametrisation noise) come from torch's generator
exactly where the reference draws them, or are injected for parity tests.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from . import native, ops

F32 = torch.float32
HYPER_SLOTS = 8


class GradBuckets:
    """Data-parallel gradient averaging over contiguous ranges of the flat gradient buffer. The ranges follow the
    backward's completion order (UnetTrainEngine.bucket_ends), so bucket k can be all-reduced on a side stream while
    the backward kernels of the later stages still run. SUM all-reduce; the 1/world factor is folded into the clip
    and Adam kernels (grad_div)."""

    def __init__(self, flat_grad: torch.Tensor, bucket_ends, group=None, min_bucket_elems: int = 8 * 1024 * 1024):
        self.flat, self.group = flat_grad, group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        # merge small stages so that every all-reduce moves at least min_bucket_elems (launch latency vs overlap)
        self.ranges, lo, last_end = {}, 0, 0
        for stage, (_, end) in enumerate(bucket_ends):
            last = stage == len(bucket_ends) - 1
            if end - lo >= min_bucket_elems or last:
                self.ranges[stage] = (lo, end)
                lo = end
            last_end = end
        assert last_end == flat_grad.numel() and lo == last_end
        self.side = torch.cuda.Stream() if flat_grad.is_cuda else None
        self.works = []
        self.enabled = True  # False: skip the exchange (bench.py's "what does the all-reduce cost a step" experiment)
        self.elem_bytes, self.dtype_name = flat_grad.element_size(), str(flat_grad.dtype).replace("torch.", "")

    def plan(self):
        return [self.ranges[s] for s in sorted(self.ranges)]

    def on_stage_done(self, stage: int):
        if self.world == 1 or stage not in self.ranges or not self.enabled:
            return
        lo, hi = self.ranges[stage]
        view = self.flat[lo:hi]
        if self.side is None:  # CPU (gloo) path used by the host-logic tests
            dist.all_reduce(view, group=self.group)
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        with torch.cuda.stream(self.side):
            self.side.wait_event(ev)
            dist.all_reduce(view, group=self.group)
        done = torch.cuda.Event()
        done.record(self.side)
        self.works.append(done)

    def finish(self):
        for ev in self.works:
            torch.cuda.current_stream().wait_event(ev)
        self.works.clear()


class DiffusionTrainStep:
    def __init__(self, unet, scheduler, batch: int, latent_shape=(3, 32, 32), clip_grad: float | None = 1.0,
                 cond_drop_prob: float = 0.15, sample_latents: bool = True, betas=(0.9, 0.999), eps: float = 1e-8,
                 group=None, use_graph: bool = True, data_parallel: bool = True):
        """data_parallel=False keeps the step local even when a process group is initialised (single-rank reference
        runs next to a data-parallel job must not enter collectives the other ranks never call)."""
        self.unet = unet
        dev = unet.in_conv.weight.device
        if dev.type != "cuda":
            raise RuntimeError("DiffusionTrainStep: CUDA (sm_100a) required; there is no CPU path")
        self.dev, self.B, self.shape = dev, batch, (batch, *latent_shape)
        self.sched = scheduler._on(dev)
        self.clip_grad, self.cond_drop_prob, self.sample_latents = clip_grad, cond_drop_prob, sample_latents
        self.betas, self.eps = betas, eps
        self.eng = unet.train_engine()
        self._flatten_params()
        n = self.eng.flat_numel
        self.exp_avg = torch.zeros(n, device=dev, dtype=F32)
        self.exp_avg_sq = torch.zeros(n, device=dev, dtype=F32)
        self.step_count = 0
        self.buckets = GradBuckets(self.eng.flat_grad, self.eng.bucket_ends, group)
        if not data_parallel:
            self.buckets.world = 1
        self.world = self.buckets.world
        if self.world > 1:
            # replicas start from rank 0's parameters and optimizer state (what torch DDP does at construction);
            # only gradients are exchanged afterwards
            for buf in (self.flat_param, self.exp_avg, self.exp_avg_sq):
                dist.broadcast(buf, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            self.eng.pkey = None
        C = latent_shape[0]
        z = lambda *s, dt=F32: torch.zeros(*s, device=dev, dtype=dt)
        self.latents = z(batch, (2 if sample_latents else 1) * C, *latent_shape[1:])
        self.labels = z(batch, dt=torch.int64)
        self.reparam_noise = z(*self.shape) if sample_latents else None
        self.noise, self.x_noise, self.pred, self.dpred = z(*self.shape), z(*self.shape), z(*self.shape), z(*self.shape)
        self.t = z(batch, dt=torch.int64)
        self.mask = z(batch)
        self.loss = z(1)
        self.norm_clip = z(2)
        self.hyper = z(4)
        # {lr, 1 - beta1^k, sqrt(1 - beta2^k)} of step k travel through a RING of pinned slots: the H2D copy of slot j
        # is asynchronous, and the host may be several steps ahead of the GPU, so a slot is rewritten only after the
        # event recorded behind its previous copy has completed (a single reused pinned buffer would let step k read
        # step k+n's values).
        self.hyper_ring = [torch.zeros(4, dtype=F32).pin_memory() for _ in range(HYPER_SLOTS)]
        self.hyper_events = [None] * HYPER_SLOTS
        self.sq_scratch = z(2048)
        # One rank: the whole step is ONE CUDA graph. Several ranks: the step is captured as a CHAIN of graphs cut at
        # the gradient-bucket boundaries; the NCCL all-reduce of bucket k is issued eagerly on a side stream between
        # the launches of segments k and k+1 and overlaps the remaining backward segments (collectives captured
        # inside a graph with side-stream fork/join deadlocked under torch 2.11 + NCCL 2.28, measured on 2 x B200).
        self.use_graph = use_graph
        self.graph = None
        self.segments = None
        self.launches_per_step = None

    def _flatten_params(self):
        """Moves every parameter into one flat fp32 buffer in the gradient layout; the nn.Parameters become views."""
        eng = self.eng
        self.flat_param = torch.zeros(eng.flat_numel, device=self.dev, dtype=F32)
        with torch.no_grad():
            for name in eng.grad_names:
                p = eng.params[name]
                view = self.flat_param[eng.goff[name]:eng.goff[name] + p.numel()].view(p.shape)
                view.copy_(p.detach())
                p.data = view
        eng.pkey = None

    # ---------------------------------------------------------------------------------------------
    def draw(self, generator=None):
        """The step's random draws, in the reference's order (diffusion_trainer.py:153,160-161,167)."""
        if self.reparam_noise is not None:
            self.reparam_noise.normal_(generator=generator)
        self.noise.normal_(generator=generator)
        self.t.random_(0, self.sched.num_steps, generator=generator)
        self.mask.copy_((torch.rand(self.B, device=self.dev, generator=generator) > self.cond_drop_prob).to(F32))

    def _fwd_bwd(self):
        ops.reparam_add_noise(self.latents, self.reparam_noise, self.noise, self.t, self.sched, self.x_noise)
        self.eng.forward(self.x_noise, self.t, self.labels, self.mask, self.pred)
        ops.mse_loss_grad(self.pred, self.noise, self.dpred, self.loss)
        self.eng.backward(self.dpred, on_stage_done=self.buckets.on_stage_done if self.world > 1 else None)

    def _update(self):
        gd = float(self.world)
        clip = None
        if self.clip_grad is not None:
            ops.grad_norm_clip(self.eng.flat_grad, self.norm_clip, self.sq_scratch, self.clip_grad, grad_div=gd)
            clip = self.norm_clip
        ops.adam_step(self.flat_param, self.eng.flat_grad, self.exp_avg, self.exp_avg_sq, self.hyper, self.betas[0],
                      self.betas[1], self.eps, grad_div=gd, clip2=clip)
        self.eng.prepare(force=True)

    def _whole_step(self):
        self._fwd_bwd()
        if self.world > 1:
            self.buckets.finish()
        self._update()

    def _capture_segments(self):
        """world > 1: graphs [fwd + bwd stages up to bucket 0], [.. bucket 1], ..., [clip + Adam + re-pack]."""
        stream = torch.cuda.Stream()
        segs, state = [], {"g": None, "pool": torch.cuda.graph_pool_handle()}

        def begin():
            g = torch.cuda.CUDAGraph()
            g.capture_begin(pool=state["pool"])
            state["g"] = g

        def cut(bucket):
            state["g"].capture_end()
            segs.append((state["g"], bucket))
            begin()

        def on_stage(stage):
            if stage in self.buckets.ranges:
                cut(self.buckets.ranges[stage])

        torch.cuda.synchronize()
        stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(stream):
            begin()
            ops.reparam_add_noise(self.latents, self.reparam_noise, self.noise, self.t, self.sched, self.x_noise)
            self.eng.forward(self.x_noise, self.t, self.labels, self.mask, self.pred)
            ops.mse_loss_grad(self.pred, self.noise, self.dpred, self.loss)
            self.eng.backward(self.dpred, on_stage_done=on_stage)
            # the last backward stage always closes a bucket, so a fresh segment is open here: the update
            self._update()
            state["g"].capture_end()
            segs.append((state["g"], None))
        torch.cuda.current_stream().wait_stream(stream)
        torch.cuda.synchronize()
        self.segments = segs

    def _replay_segments(self):
        b = self.buckets
        for g, bucket in self.segments:
            if bucket is None:
                b.finish()  # the update segment needs every bucket averaged
            g.replay()
            if bucket is not None and b.enabled:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream())
                with torch.cuda.stream(b.side):
                    b.side.wait_event(ev)
                    dist.all_reduce(b.flat[bucket[0]:bucket[1]], group=b.group)
                    done = torch.cuda.Event()
                    done.record(b.side)
                b.works.append(done)

    def _ensure_graph(self):
        if self.graph is not None or self.segments is not None or not self.use_graph:
            return
        if self.world > 1:
            keep = (self.flat_param.clone(), self.exp_avg.clone(), self.exp_avg_sq.clone())
            self.hyper.copy_(torch.tensor(ops.adam_hyper(0.0, 1, *self.betas)))
            self._whole_step()  # warm-up (eager): allocates workspaces, sets kernel attributes, initialises NCCL
            torch.cuda.synchronize()
            before = native.launch_count
            self._capture_segments()
            self.launches_per_step = native.launch_count - before
            for dst, src in zip((self.flat_param, self.exp_avg, self.exp_avg_sq), keep):
                dst.copy_(src)
            self.eng.prepare(force=True)
            return
        keep = (self.flat_param.clone(), self.exp_avg.clone(), self.exp_avg_sq.clone())
        self.hyper.copy_(torch.tensor(ops.adam_hyper(0.0, 1, *self.betas)))
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._whole_step()  # warm-up: allocates workspaces, sets kernel attributes
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        before = native.launch_count
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._whole_step()
        self.launches_per_step = native.launch_count - before
        self.graph = g
        for dst, src in zip((self.flat_param, self.exp_avg, self.exp_avg_sq), keep):
            dst.copy_(src)
        self.eng.prepare(force=True)

    def step(self, latents: torch.Tensor, labels: torch.Tensor, lr: float, generator=None, draw: bool = True):
        """One training step on a batch of stored latents (mean || logvar, or plain latents) and class labels.
        Returns the device scalar loss (no host sync)."""
        self._ensure_graph()
        self.latents.copy_(latents)
        self.labels.copy_(labels)
        if draw:
            self.draw(generator)
        self.step_count += 1
        slot = self.step_count % HYPER_SLOTS
        if self.hyper_events[slot] is not None:
            self.hyper_events[slot].synchronize()
        host = self.hyper_ring[slot]
        host.copy_(torch.tensor(ops.adam_hyper(lr, self.step_count, *self.betas)))
        self.hyper.copy_(host, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.hyper_events[slot] = ev
        # the fused Adam kernel writes the parameters through a raw pointer (no torch version bump): tell the
        # inference engines / cached samplers of this module that their packed weights are stale
        self.unet._weights_epoch = getattr(self.unet, "_weights_epoch", 0) + 1
        if self.graph is not None:
            self.graph.replay()
        elif self.segments is not None:
            self._replay_segments()
        else:
            self._whole_step()
        return self.loss

    @property
    def grad_norm(self) -> torch.Tensor:
        """Global gradient norm of the last step (device scalar), as clip_grad_norm_ returns it."""
        return self.norm_clip[0]
