"""Drop-in `VAE` (reference modules/vae.py:11-144): KL or VQ bottleneck, same 13-argument constructor, attributes,
state_dict keys, error behaviour and checkpoint format; encoder/decoder run on idf_b200.engine.VaeEngine."""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from idf_b200 import ops
from idf_b200.engine import VaeEngine
from idf_b200.spec import register_tree, vae_param_spec

from .components import Codebook


class VAE(nn.Module):

    def __init__(self, in_channels: int, channels: list, z_dim: int, bottleneck: str, codebook_size: int,
                 codebook_beta: float, codebook_gamma: float, enc_num_res_blocks: int, dec_num_res_blocks: int,
                 attn_resolutions: list, num_heads: int, init_resolution: int, num_groups: int):
        super().__init__()
        self.bottleneck = bottleneck
        self.architecture = dict(in_channels=in_channels, channels=channels, z_dim=z_dim, bottleneck=bottleneck,
                                 codebook_size=codebook_size, codebook_beta=codebook_beta,
                                 codebook_gamma=codebook_gamma, enc_num_res_blocks=enc_num_res_blocks,
                                 dec_num_res_blocks=dec_num_res_blocks, attn_resolutions=attn_resolutions,
                                 num_heads=num_heads, init_resolution=init_resolution, num_groups=num_groups)
        spec = vae_param_spec(self.architecture)
        register_tree(self, type(spec)((k, v) for k, v in spec.items() if not k.startswith("codebook.")))
        self.codebook = Codebook(codebook_size, z_dim, codebook_beta, codebook_gamma) if bottleneck == "vq" else None
        self._engines = {}

    def _engine(self, key) -> VaeEngine:
        dev = next(self.parameters()).device
        key = key + (str(dev),)
        eng = self._engines.get(key)
        if eng is None:
            eng = self._engines[key] = VaeEngine(self, self.architecture, dev)
        return eng

    def _check(self, x, what):
        if not x.is_cuda:
            raise RuntimeError(f"VAE.{what}: CUDA (sm_100a) tensors required; there is no CPU path")
        if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("VAE under autograd: stage-1 training is out of scope; use torch.no_grad()")

    def forward(self, x, return_metrics=False):
        z, loss, perplexity = self.encode(x, sample=self.bottleneck == "kl")
        x_hat = self.decode(z)
        return (x_hat, loss, perplexity) if return_metrics else x_hat

    def encode(self, x, sample=False):
        if self.bottleneck == "vq" and sample:
            raise ValueError("Cannot sample from the VQ model!")
        self._check(x, "encode")
        xin = x.detach().to(torch.float32).contiguous()
        B, _, H, W = xin.shape
        f = 2 ** (len(self.architecture["channels"]) - 1)
        zc = self.architecture["z_dim"] * (1 if self.bottleneck == "vq" else 2)
        z = torch.empty(B, zc, H // f, W // f, device=x.device, dtype=torch.float32)
        self._engine(("enc", B, H, W)).encode(xin, z)
        if self.bottleneck == "vq":
            return self.codebook(z)
        # KL loss (+ reparametrised sample) in one fused kernel pair (vae.py:99-113); the noise comes from torch's
        # global CUDA generator exactly where the reference draws it (randn_like(mean))
        noise = torch.randn(B, zc // 2, H // f, W // f, device=x.device, dtype=torch.float32) if sample else None
        kl_loss, zs = ops.kl_loss_reparam(z, noise)
        return (zs if sample else z), kl_loss, 0.0

    def decode(self, z, quantize=False):
        if self.bottleneck == "kl" and quantize:
            raise ValueError("Cannot quantize in the KL model!")
        self._check(z, "decode")
        zin = z.detach().to(torch.float32).contiguous()
        if quantize:
            zin, _ = self.codebook.quantize(zin)
        B, _, H, W = zin.shape
        f = 2 ** (len(self.architecture["channels"]) - 1)
        out = torch.empty(B, self.architecture["in_channels"], H * f, W * f, device=z.device, dtype=torch.float32)
        self._engine(("dec", B, H, W)).decode(zin, out)
        return out

    @classmethod
    def from_checkpoint(cls, path=None, checkpoint=None):
        if path is None and checkpoint is None:
            raise ValueError("Either `path` or `checkpoint` must be specified.")
        if path is not None:
            checkpoint = torch.load(path)
        model = cls(**checkpoint["architecture"])
        model.load_state_dict({k.replace("_orig_mod.", ""): v for k, v in checkpoint["vae"].items()})
        return model

    def to_checkpoint(self, path):
        folder = os.path.dirname(path)
        if folder:
            os.makedirs(folder, exist_ok=True)
        torch.save({"vae": self.state_dict(), "architecture": self.architecture}, path)
