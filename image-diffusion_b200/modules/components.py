"""B200-native counterparts of the reference's modules/components.py pieces that sit on the hot path:
`Scheduler` (components.py:364-424) and `Codebook` (components.py:249-315). The layer classes of the reference
(ConvBlock, DiffusionBlock, ...) have no counterpart here: the engines in idf_b200.engine execute whole networks
from a flat parameter tree."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from idf_b200 import ops


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what}: this implementation runs on CUDA (sm_100a) only; there is no CPU path")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


class Scheduler:
    """DDPM noise schedule with the reference's constructor, attributes and methods.

    type="linear" is linear in sqrt(beta) (components.py:389-392); type="cosine" follows components.py:380-387.
    The tables are built once on the host side; add_noise / sample_prev_timestep are single fused kernels.
    """

    def __init__(self, num_steps: int, beta_start: float = 0.0001, beta_end: float = 0.02, type: str = "linear",
                 device: str = "cpu"):
        self.num_steps, self.beta_start, self.beta_end, self.type = num_steps, beta_start, beta_end, type
        if type == "linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_steps) ** 2
        elif type == "cosine":
            s = 8e-3
            grid = torch.arange(num_steps + 1, dtype=torch.float32) / num_steps
            f = torch.cos((grid + s) / (1 + s) * math.pi / 2).pow(2)
            a_hat = f / f[0]
            betas = torch.clip(1 - a_hat[1:] / a_hat[:-1], min=0, max=0.999)
        else:
            raise ValueError(f"unknown schedule type {type!r}")
        self.betas = betas.to(device)
        self.alphas = 1.0 - self.betas
        self.alpha_cum_prod = torch.cumprod(self.alphas, dim=0)
        self.sqrt_alpha_cum_prod = torch.sqrt(self.alpha_cum_prod)
        self.sqrt_one_minus_alpha_cum_prod = torch.sqrt(1 - self.alpha_cum_prod)

    def _on(self, device):
        if self.betas.device != device:
            for n in ("betas", "alphas", "alpha_cum_prod", "sqrt_alpha_cum_prod", "sqrt_one_minus_alpha_cum_prod"):
                setattr(self, n, getattr(self, n).to(device))
        return self

    def add_noise(self, x, noise: torch.Tensor, t: torch.Tensor):
        """sqrt(acp[t]) * x + sqrt(1 - acp[t]) * noise with per-sample t (components.py:399-403)."""
        _require_cuda(x, "Scheduler.add_noise")
        self._on(x.device)
        out = torch.empty(x.shape, device=x.device, dtype=torch.float32)
        ops.add_noise(_f32c(x), _f32c(noise), t.to(torch.int64).contiguous(), self, out)
        return out

    def sample_prev_timestep(self, xt: torch.Tensor, noise_pred: torch.Tensor, t: torch.Tensor):
        """One ancestral step, returns (x_{t-1}, x0) like components.py:405-424. As in the reference the noise is
        drawn from the global generator only when t[0] != 0 (which costs the same host sync as the reference);
        Diffusion.sample uses the sync-free fused CFG step instead."""
        _require_cuda(xt, "Scheduler.sample_prev_timestep")
        self._on(xt.device)
        xt_c, eps = _f32c(xt), _f32c(noise_pred)
        z = torch.randn_like(xt_c) if int(t[0]) != 0 else xt_c
        out, x0 = torch.empty_like(xt_c), torch.empty_like(xt_c)
        zero = torch.zeros(xt_c.shape[0], device=xt.device, dtype=torch.float32)
        ops.cfg_posterior_step(xt_c, eps, eps, z, zero, t.to(torch.int64).contiguous(), self, out, x0)
        return out, x0


    def ddim_step(self, xt: torch.Tensor, noise_pred: torch.Tensor, t: int, t_prev: int, eta: float = 0.0,
                  noise: torch.Tensor | None = None):
        """Strided step x_t -> x_{t_prev} (t_prev < 0: final step) on this schedule's tables, returns (x_prev, x0):
        Song et al. (DDIM) eq. 12; not in the reference, whose only sampler is the 1-step ancestral update above."""
        _require_cuda(xt, "Scheduler.ddim_step")
        self._on(xt.device)
        xt_c, eps = _f32c(xt), _f32c(noise_pred)
        if eta > 0.0 and t_prev >= 0 and noise is None:
            noise = torch.randn_like(xt_c)
        out, x0 = torch.empty_like(xt_c), torch.empty_like(xt_c)
        zero = torch.zeros(xt_c.shape[0], device=xt.device, dtype=torch.float32)
        tt = torch.tensor([t], device=xt.device, dtype=torch.int64)
        tp = torch.tensor([t_prev], device=xt.device, dtype=torch.int64)
        ops.cfg_ddim_step(xt_c, eps, eps, noise if noise is not None else xt_c, zero, tt, tp, self, out, eta=eta,
                          x0_out=x0)
        return out, x0


class Codebook(nn.Module):
    """VQ codebook with EMA statistics; same parameters/buffers and return triple as components.py:249-315.

    The nearest-code search is the warp-shuffle argmin kernel (no distance matrix, no per-batch codebook copy) and
    reproduces torch.cdist + argmin bit for bit. The training-mode EMA bookkeeping (components.py:284-298) is
    host-orchestrated tensor plumbing outside the hot path.
    """

    def __init__(self, size: int, dim: int, beta: float, gamma: float, epsilon: float = 1e-5):
        super().__init__()
        self.embeddings = nn.Module()
        self.embeddings.register_parameter("weight", nn.Parameter(torch.empty(size, dim).uniform_(-1 / size, 1 / size)))
        self.size, self.dim, self.beta, self.gamma, self.epsilon = size, dim, beta, gamma, epsilon
        self.register_buffer("ema_cluster_size", torch.zeros(size))
        self.ema_w = nn.Parameter(torch.empty(size, dim).uniform_(-1 / size, 1 / size))

    def quantize(self, x: torch.Tensor):
        """x fp32 (B, dim, H, W) -> (z_q NCHW fp32, indices int64 (B*H*W,))."""
        _require_cuda(x, "Codebook")
        z = _f32c(x)
        B, _, H, W = z.shape
        idx = torch.empty(B * H * W, device=z.device, dtype=torch.int64)
        zq = torch.empty_like(z)
        ops.vq_argmin(z, _f32c(self.embeddings.weight), idx, zq)
        return zq, idx

    def forward(self, x):
        zq, idx = self.quantize(x)
        z = _f32c(x)
        if not self.training:
            # eval: commitment loss, straight-through output and perplexity in one fused kernel pair (no torch arithmetic)
            return ops.vq_loss_perplexity(z, zq, idx, self.size, self.beta)
        with torch.no_grad():  # EMA codebook update (components.py:284-298): host-orchestrated, off the hot path
            flat = z.permute(0, 2, 3, 1).reshape(-1, self.dim)
            counts = torch.bincount(idx, minlength=self.size).to(torch.float32)
            self.ema_cluster_size = self.ema_cluster_size * self.gamma + (1 - self.gamma) * counts
            n = torch.sum(self.ema_cluster_size)
            self.ema_cluster_size = (self.ema_cluster_size + self.epsilon) / (n + self.size * self.epsilon) * n
            dw = torch.zeros_like(self.ema_w).index_add_(0, idx, flat)
            self.ema_w = nn.Parameter(self.ema_w * self.gamma + (1 - self.gamma) * dw)
            self.embeddings.weight = nn.Parameter(self.ema_w / self.ema_cluster_size.unsqueeze(1))
        quant_loss = self.beta * torch.mean((zq - z) ** 2)
        quant_out = z + (zq - z).detach()  # straight-through (components.py:305)
        probs = torch.bincount(idx, minlength=self.size).to(torch.float32) / idx.numel()
        perplexity = torch.exp(-torch.sum(probs * torch.log(probs + 1e-6)))
        return quant_out, quant_loss, perplexity
