"""Batch-sharded sampling across the GPUs of one node (BASELINE config 3): one process per GPU, contiguous shards of
the latent batch, no per-step cross-GPU traffic — the only collectives are the optional final gather of decoded
images. Randomness is keyed by the GLOBAL micro-batch index, so the images do not depend on the number of ranks."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int, rank: int, align: int = 1):
    """Contiguous, balanced [lo, hi) of `total` items for `rank`, with boundaries on multiples of `align`."""
    if total % align:
        raise ValueError(f"total {total} is not a multiple of the micro-batch {align}")
    units = total // align
    base, extra = divmod(units, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo * align, hi * align


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def gather_shards(local: torch.Tensor, total: int, align: int = 1, group=None) -> torch.Tensor:
    """All-gathers variable-length contiguous shards (dim 0) into the full tensor on every rank."""
    rank, world = world_info()
    if world == 1:
        return local
    sizes = [shard_bounds(total, world, r, align) for r in range(world)]
    longest = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros(longest, *local.shape[1:], dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)


class ShardedSampler:
    """Samples `total` images = class labels[i], guidance scale cfg[i] for i in this rank's shard, in micro-batches.

    Micro-batch g (global samples [g*mb, (g+1)*mb)) draws x_T and every step's noise from its own generator seeded
    with (seed, g): the result is identical for any world size (SURVEY.md §8e shard invariance)."""

    def __init__(self, diffusion, labels: torch.Tensor, cfg_scales: torch.Tensor, micro_batch: int, seed: int = 0):
        self.d, self.labels, self.cfg, self.mb, self.seed = diffusion, labels, cfg_scales, micro_batch, seed
        self.total = labels.shape[0]

    @torch.no_grad()
    def run(self, steps=None, decode: bool = True, decode_events=None) -> torch.Tensor:
        """decode_events: optional list that receives a (start, end) CUDA-event pair around every decode call."""
        from .sampler import CfgSampler
        rank, world = world_info()
        lo, hi = shard_bounds(self.total, world, rank, self.mb)
        dev = self.labels.device
        outs = []
        sampler = None
        sched_steps = list(reversed(range(self.d.scheduler.num_steps))) if steps is None else list(steps)
        for start in range(lo, hi, self.mb):
            lab, cfg = self.labels[start:start + self.mb], self.cfg[start:start + self.mb]
            if sampler is None:
                # one sampler (one captured graph) serves every micro-batch of the shard - and every later run of the
                # same micro-batch size on this Diffusion object (labels / scales are rewritten in place)
                cache = self.d.__dict__.setdefault("_shard_samplers", {})
                sampler = cache.get(self.mb)
                if sampler is None:
                    sampler = cache[self.mb] = CfgSampler(self.d.unet, self.d.scheduler, lab, cfg, self.d.latent_shape)
            sampler.set_conditioning(lab, cfg)
            gen = torch.Generator(device=dev).manual_seed(self.seed * 1000003 + start // self.mb)
            x_T = torch.randn(self.mb, *self.d.latent_shape, device=dev, generator=gen)
            sampler.set_latent(x_T)
            z = torch.empty_like(x_T)
            for i in sched_steps:
                if i > 0:
                    z.normal_(generator=gen)
                sampler.step(i, noise=z)
            lat = sampler.latent.clone()
            if decode and decode_events is not None:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
            outs.append(self.d.vae.decode(lat, quantize=self.d.vae.architecture["bottleneck"] == "vq") if decode else lat)
            if decode and decode_events is not None:
                ev[1].record()
                decode_events.append(ev)
        if not outs:
            shape = (0, *self.d.latent_shape)
            return torch.empty(shape, device=dev)
        return torch.cat(outs, dim=0)
