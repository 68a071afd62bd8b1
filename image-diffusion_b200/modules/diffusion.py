"""Drop-in `Diffusion` (reference modules/diffusion.py:11-105): owns vae + unet + scheduler + class names, samples
with classifier-free guidance and decodes; same constructor, sample() contract and bundle-checkpoint format."""
from __future__ import annotations

import os

import torch

from idf_b200.sampler import CfgSampler
from modules.components import Scheduler
from modules.unet import Unet
from modules.vae import VAE


class Diffusion:

    def __init__(self, vae: VAE, unet: Unet, scheduler: Scheduler, classes: str, device: str = "cuda"):
        self.vae, self.unet, self.scheduler = vae, unet, scheduler
        self.classes = classes.split(",") if isinstance(classes, str) else list(classes)
        self.device = device
        arch = vae.architecture
        side = arch["init_resolution"] // 2 ** (len(arch["channels"]) - 1)
        self.latent_shape = (unet.architecture["z_dim"], side, side)
        self._samplers = {}

    @torch.no_grad()
    def sample(self, cfg_scales, num_images: int = 10, seed: int = None, *, steps=None, sampler: str = "ddpm",
               eta: float = 0.0):
        """`len(classes)` x `len(cfg_scales)` images (list) or `len(classes)` x `num_images` images (int scale),
        returned as the unclamped fp32 decoder output (N, 3, R, R) like diffusion.py:31-60.

        Class and scale are paired exactly as the reference pairs them (image k gets class k % B and scale
        cfg[k % C], diffusion.py:44,49). RNG: x_T and each step's noise are drawn from the global CUDA generator in
        the reference's order, so a seed reproduces the reference's draws on the same device.

        Beyond the reference (keyword-only, defaults reproduce it): `steps` = a decreasing subset of the schedule and
        `sampler="ddim"` (+ `eta`) run the strided sampler on the same kernels."""
        assert self.device == "cuda" and torch.cuda.is_available(), "You need a GPU to sample images."
        if seed is not None:
            torch.manual_seed(seed)
        B = len(self.classes)
        C = len(cfg_scales) if isinstance(cfg_scales, list) else num_images
        scales = cfg_scales if isinstance(cfg_scales, list) else [cfg_scales] * num_images
        cfg = torch.tensor(B * scales, device=self.device)
        xt = torch.randn(B * C, *self.latent_shape, device=self.device)
        labels = torch.tensor(list(range(B)) * C, device=self.device)
        if sampler == "ddpm" and steps is not None:
            raise ValueError("Diffusion.sample: the ancestral DDPM update advances one timestep at a time; pass "
                             "sampler='ddim' (eta=1.0 for the ancestral variance) to use a subset of the steps")
        key = (B * C, tuple(labels.tolist()), tuple(cfg.tolist()), sampler, float(eta))
        smp = self._samplers.get(key)
        if smp is None:
            self._samplers = {key: CfgSampler(self.unet, self.scheduler, labels, cfg, self.latent_shape, kind=sampler,
                                              eta=eta)}
            smp = self._samplers[key]
        xt = smp.run(xt, steps=steps)
        return self.vae.decode(xt, quantize=self.vae.architecture["bottleneck"] == "vq")

    @classmethod
    def from_checkpoint(cls, path: str, device: str = "cuda"):
        ck = torch.load(path)
        vae = VAE.from_checkpoint(checkpoint=ck["v"]).to(device).eval()
        unet = Unet.from_checkpoint(checkpoint=ck["u"]).to(device).eval()
        s = ck["scheduler"]
        scheduler = Scheduler(s["num_steps"], s["beta_start"], s["beta_end"], s["type"], device)
        return cls(vae, unet, scheduler, ck["classes"], device)

    def to_checkpoint(self, path: str):
        s = self.scheduler
        bundle = {
            "v": {"vae": self.vae.state_dict(), "architecture": self.vae.architecture},
            "u": {"unet": self.unet.state_dict(), "architecture": self.unet.architecture},
            "scheduler": {"num_steps": s.num_steps, "beta_start": s.beta_start, "beta_end": s.beta_end,
                          "type": s.type},
            "classes": self.classes,
        }
        folder = os.path.dirname(path)
        if folder:
            os.makedirs(folder, exist_ok=True)
        torch.save(bundle, path)
