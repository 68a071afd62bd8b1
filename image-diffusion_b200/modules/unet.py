"""Drop-in `Unet` (reference modules/unet.py:13-159): same constructor, attributes, state_dict keys and checkpoint
format; forward() executes the fixed kernel sequence of idf_b200.engine.UnetEngine."""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from idf_b200.engine import UnetEngine
from idf_b200.spec import register_tree, unet_param_spec
from idf_b200.train_engine import UnetTrainEngine


class _UnetTrainFn(torch.autograd.Function):
    """Autograd bridge for the reference trainer (trainers/diffusion_trainer.py:169-173): forward and backward both
    run the C-ABI kernel sequences of UnetTrainEngine; parameter gradients come back as fp32 tensors in PyTorch layout
    (copies of the engine's flat gradient buffer, so .grad accumulation semantics hold)."""

    @staticmethod
    def forward(ctx, module, x, t, context, mask, *params):
        eng = module.train_engine()
        out = torch.empty_like(x)
        eng.forward(x, t, context, mask, out)
        ctx.eng = eng
        ctx.names = [n for n, _ in module.named_parameters()]
        return out

    @staticmethod
    def backward(ctx, grad_out):
        eng = ctx.eng
        eng.backward(grad_out.to(torch.float32).contiguous())
        return (None, None, None, None, None) + tuple(eng.gv[n].clone() for n in ctx.names)


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
        self._train_engine = None

    def train_engine(self) -> UnetTrainEngine:
        dev = self.in_conv.weight.device
        if self._train_engine is None or self._train_engine.device != dev:
            self._train_engine = UnetTrainEngine(self, self.architecture, dev)
        return self._train_engine

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
        B, _, H, W = x.shape
        xin = x.detach().to(torch.float32).contiguous()
        t = timestep.to(device=x.device, dtype=torch.int64).contiguous()
        ctx = None if context is None else context.to(device=x.device, dtype=torch.int64).contiguous()
        mask = None
        if context is not None and context_mask is not None:
            mask = context_mask.to(device=x.device, dtype=torch.float32).reshape(-1).contiguous()
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return _UnetTrainFn.apply(self, xin, t, ctx, mask, *self.parameters())
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
