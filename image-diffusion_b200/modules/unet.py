"""Drop-in `Unet` (reference modules/unet.py:13-159): same constructor, attributes, state_dict keys and checkpoint
format; forward() executes the fixed kernel sequence of idf_b200.engine.UnetEngine."""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from idf_b200.engine import UnetEngine
from idf_b200.spec import register_tree, unet_param_spec


class Unet(nn.Module):

    def __init__(self, z_dim: int, channels: list, mid_channels: list, time_dim: int, num_res_layers: int,
                 num_heads: int, num_groups: int, num_classes: int):
        super().__init__()
        self.time_dim, self.num_classes = time_dim, num_classes
        self.architecture = dict(z_dim=z_dim, channels=channels, mid_channels=mid_channels, time_dim=time_dim,
                                 num_res_layers=num_res_layers, num_heads=num_heads, num_groups=num_groups,
                                 num_classes=num_classes)
        register_tree(self, unet_param_spec(self.architecture))
        self._engines = {}

    def engine(self, batch: int, height: int, width: int) -> UnetEngine:
        dev = self.in_conv.weight.device
        key = (batch, height, width, str(dev))
        eng = self._engines.get(key)
        if eng is None:
            eng = self._engines[key] = UnetEngine(self, self.architecture, dev)
        return eng

    def forward(self, x, timestep, context=None, context_mask=None):
        """eps = Unet(x, t, y[, mask]) as unet.py:103-136: x (B, z, H, W), timestep (B,) int64, context (B,) int64 or
        None (unconditional), context_mask (B, 1) multiplies the class embedding row (0 = dropped)."""
        if not x.is_cuda:
            raise RuntimeError("Unet.forward: CUDA (sm_100a) tensors required; there is no CPU path")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and self.training:
            raise NotImplementedError("Unet.forward under autograd: the backward kernels are not built yet "
                                      "(inference / sampling only in this round); wrap the call in torch.no_grad()")
        B, _, H, W = x.shape
        xin = x.detach().to(torch.float32).contiguous()
        t = timestep.to(device=x.device, dtype=torch.int64).contiguous()
        ctx = None if context is None else context.to(device=x.device, dtype=torch.int64).contiguous()
        mask = None
        if context is not None and context_mask is not None:
            mask = context_mask.to(device=x.device, dtype=torch.float32).reshape(-1).contiguous()
        out = torch.empty_like(xin)
        self.engine(B, H, W).run(xin, t, ctx, mask, None, out)
        return out

    @classmethod
    def from_checkpoint(cls, path=None, checkpoint=None):
        if path is None and checkpoint is None:
            raise ValueError("Either `path` or `checkpoint` must be specified.")
        if path is not None:
            checkpoint = torch.load(path)
        model = cls(**checkpoint["architecture"])
        # torch.compile leaves an `_orig_mod.` prefix on every key (unet.py:147-148)
        model.load_state_dict({k.replace("_orig_mod.", ""): v for k, v in checkpoint["unet"].items()})
        return model

    def to_checkpoint(self, path):
        folder = os.path.dirname(path)
        if folder:
            os.makedirs(folder, exist_ok=True)
        torch.save({"unet": self.state_dict(), "architecture": self.architecture}, path)
