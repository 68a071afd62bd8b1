"""The callers and data formats either side of the hot path, as thin host code over the same kernels:

  f1  dataset-scale latent extraction (scripts/prepare_dataset.py:81-109): uint8 NHWC images -> VAE encoder ->
      fp16 `.npy` wire format [M, zc, 32, 32] (zc = 6: mean || logvar for the KL model)
  f2  the rest of the UNet trainer around the fused step (trainers/diffusion_trainer.py:102-217): LR warm-up,
      epoch loop, checkpoints in the reference's `save_checkpoint` / `load_checkpoint` format (modules/util.py:81-108)
      with a torch.optim.Adam-compatible optimizer state
  f3  the output stage of scripts/sample_grid.py:44-47: clamp -> image grid -> uint8 (PNG when Pillow is present)
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import native
from .trainer import DiffusionTrainStep


# ---------------------------------------------------------------------------------------------
# f1: latent extraction
# ---------------------------------------------------------------------------------------------
@torch.no_grad()
def extract_latents(vae, images, batch_size: int = 256, out: np.ndarray | None = None) -> np.ndarray:
    """images: uint8 array-like [M, H, W, 3] (e.g. an np.load(..., mmap_mode="r") memmap). Returns / fills the fp16
    buffer [M, zc, H/4, W/4] exactly like prepare_dataset.py:95-109 (x / 127.5 - 1, NHWC -> NCHW, vae.encode(sample=
    False), .half()). Host staging is pinned and double-buffered: the copy of batch i+1 overlaps the encode of batch i."""
    dev = next(vae.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("extract_latents: CUDA (sm_100a) required; there is no CPU path")
    M, H, W, C = images.shape
    copy_stream = torch.cuda.Stream(device=dev)
    stage = [torch.empty(batch_size, H, W, C, dtype=torch.uint8).pin_memory() for _ in range(2)]
    dimg = [torch.empty(batch_size, H, W, C, dtype=torch.uint8, device=dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    x = torch.empty(batch_size, C, H, W, device=dev, dtype=torch.float32)

    def upload(i, slot):
        n = min(batch_size, M - i)
        stage[slot][:n].copy_(torch.from_numpy(np.ascontiguousarray(images[i:i + n])))
        with torch.cuda.stream(copy_stream):
            dimg[slot][:n].copy_(stage[slot][:n], non_blocking=True)
            ready[slot].record(copy_stream)
        return n

    n_next = upload(0, 0) if M > 0 else 0
    slot = 0
    for i in range(0, M, batch_size):
        n = n_next
        torch.cuda.current_stream().wait_event(ready[slot])
        native.call("idf_u8_nhwc_to_f32_nchw", dimg[slot].data_ptr(), x.data_ptr(), n, H, W, C, 1.0 / 127.5, -1.0)
        z, _, _ = vae.encode(x[:n], sample=False)
        zh = z.to(torch.float16).cpu()  # synchronises: the staging slot used two iterations ago is free again
        if out is None:
            out = np.zeros((M, *zh.shape[1:]), dtype=np.float16)
        out[i:i + n] = zh.numpy()
        if i + batch_size < M:
            n_next = upload(i + batch_size, slot ^ 1)
        slot ^= 1
    return out


# ---------------------------------------------------------------------------------------------
# f3: sample_grid output stage
# ---------------------------------------------------------------------------------------------
def image_grid(images: torch.Tensor, nrow: int, padding: int = 2) -> np.ndarray:
    """[N, 3, H, W] float images in [-1, 1] (unclamped, as Diffusion.sample returns them) -> uint8 [GH, GW, 3] grid
    laid out like torchvision.utils.make_grid(images, nrow) followed by sample_grid.py:45's clamp and (x + 1) / 2."""
    x = images.detach().float().cpu().clamp(-1.0, 1.0)
    N, C, H, W = x.shape
    ncol = min(nrow, N)
    rows = (N + ncol - 1) // ncol
    grid = torch.zeros(C, rows * (H + padding) + padding, ncol * (W + padding) + padding)
    for k in range(N):
        r, c = divmod(k, ncol)
        y0, x0 = r * (H + padding) + padding, c * (W + padding) + padding
        grid[:, y0:y0 + H, x0:x0 + W] = x[k]
    return ((grid.permute(1, 2, 0) + 1.0) * 127.5).round().clamp(0, 255).to(torch.uint8).numpy()


def save_grid(images: torch.Tensor, path: str, nrow: int) -> np.ndarray:
    g = image_grid(images, nrow)
    folder = os.path.dirname(path)
    if folder:
        os.makedirs(folder, exist_ok=True)
    try:
        from PIL import Image
        Image.fromarray(g).save(path)
    except ImportError:  # no Pillow in the image: keep the pixels
        np.save(os.path.splitext(path)[0] + ".npy", g)
    return g


# ---------------------------------------------------------------------------------------------
# f2: trainer around the fused step
# ---------------------------------------------------------------------------------------------
def warmup_lr(step: int, learning_rate: float, warmup_steps: int) -> float:
    """diffusion_trainer.py:133-138: linear from lr/100 to lr over warmup_steps, then constant."""
    if step < warmup_steps:
        min_lr = learning_rate / 100
        return min_lr + (learning_rate - min_lr) * (step / warmup_steps)
    return learning_rate


def optimizer_state_dict(ts: DiffusionTrainStep, lr: float) -> dict:
    """torch.optim.Adam-compatible state_dict (parameter order = unet.parameters()) from the flat moment buffers, so
    a checkpoint written here resumes in the reference trainer and vice versa."""
    eng = ts.eng
    names = [n for n, _ in ts.unet.named_parameters()]
    state = {}
    if ts.step_count > 0:
        for idx, n in enumerate(names):
            off, cnt, shape = eng.goff[n], eng.params[n].numel(), eng.params[n].shape
            state[idx] = {"step": torch.tensor(float(ts.step_count)),
                          "exp_avg": ts.exp_avg[off:off + cnt].view(shape).clone(),
                          "exp_avg_sq": ts.exp_avg_sq[off:off + cnt].view(shape).clone()}
    group = dict(lr=lr, betas=tuple(ts.betas), eps=ts.eps, weight_decay=0, amsgrad=False, maximize=False, foreach=None,
                 capturable=False, differentiable=False, fused=None, decoupled_weight_decay=False,
                 params=list(range(len(names))))
    return {"state": state, "param_groups": [group]}


def load_optimizer_state_dict(ts: DiffusionTrainStep, sd: dict) -> None:
    eng = ts.eng
    names = [n for n, _ in ts.unet.named_parameters()]
    steps = set()
    ts.exp_avg.zero_()
    ts.exp_avg_sq.zero_()
    for idx, st in sd.get("state", {}).items():
        n = names[int(idx)]
        off, cnt = eng.goff[n], eng.params[n].numel()
        ts.exp_avg[off:off + cnt].copy_(st["exp_avg"].reshape(-1))
        ts.exp_avg_sq[off:off + cnt].copy_(st["exp_avg_sq"].reshape(-1))
        steps.add(int(float(st["step"])))
    if len(steps) > 1:
        raise ValueError(f"load_optimizer_state_dict: per-parameter step counts differ ({sorted(steps)[:4]}...)")
    ts.step_count = steps.pop() if steps else 0
    g = sd["param_groups"][0]
    ts.betas, ts.eps = tuple(g["betas"]), g["eps"]


def save_train_checkpoint(path: str, ts: DiffusionTrainStep, epoch: int, lr: float) -> None:
    """modules/util.py:81-93 layout: {"unet", "optim", "epoch", "architecture"}."""
    folder = os.path.dirname(path)
    if folder:
        os.makedirs(folder, exist_ok=True)
    torch.save({"unet": {k: v.detach().clone() for k, v in ts.unet.state_dict().items()},
                "optim": optimizer_state_dict(ts, lr), "epoch": epoch, "architecture": ts.unet.architecture}, path)


def load_train_checkpoint(path: str, ts: DiffusionTrainStep) -> int:
    """modules/util.py:96-108 (incl. the torch.compile `_orig_mod.` prefix); returns the stored epoch."""
    ck = torch.load(path, map_location=ts.dev, weights_only=False)
    sd = {k.replace("_orig_mod.", ""): v for k, v in ck["unet"].items()}
    with torch.no_grad():
        for k, p in ts.unet.state_dict().items():  # copy INTO the flat-buffer views (keeps them views)
            p.copy_(sd[k])
    ts.eng.prepare(force=True)
    if ck.get("optim") is not None:
        load_optimizer_state_dict(ts, ck["optim"])
    return ck["epoch"]


class DiffusionTrainer:
    """Epoch loop of trainers/diffusion_trainer.py:102-217 around DiffusionTrainStep: per-step LR warm-up, loss /
    gradient-norm read-back at the logging interval only (no per-step host sync), one checkpoint per epoch."""

    def __init__(self, unet, scheduler, batch_size: int, learning_rate: float, warmup_steps: int, epochs: int,
                 clip_grad: float | None = 1.0, cond_drop_prob: float = 0.15, ae_type: str = "kl",
                 checkpoints_dir: str | None = None, checkpoint: str | None = None, log_interval: int = 50, group=None,
                 latent_shape=(3, 32, 32)):
        self.lr, self.warmup_steps, self.epochs = learning_rate, warmup_steps, epochs
        self.batch_size, self.checkpoints_dir, self.log_interval = batch_size, checkpoints_dir, log_interval
        self.step_fn = DiffusionTrainStep(unet, scheduler, batch_size, latent_shape, clip_grad=clip_grad,
                                          cond_drop_prob=cond_drop_prob, sample_latents=(ae_type == "kl"), group=group)
        self.curr_epoch = load_train_checkpoint(checkpoint, self.step_fn) + 1 if checkpoint else 0
        self.history = []

    def fit(self, latents: np.ndarray | torch.Tensor, labels: np.ndarray | torch.Tensor, generator=None):
        """latents [M, zc, h, w] (fp16 wire format accepted), labels [M]; shuffled mini-batches, last partial batch
        dropped (the step is captured for a fixed batch size)."""
        ts = self.step_fn
        lat = torch.as_tensor(np.asarray(latents)) if not torch.is_tensor(latents) else latents
        lab = torch.as_tensor(np.asarray(labels)).long() if not torch.is_tensor(labels) else labels.long()
        steps_per_epoch = lat.shape[0] // self.batch_size
        for epoch in range(self.curr_epoch, self.epochs):
            perm = torch.randperm(lat.shape[0], generator=generator)
            acc = []
            for step in range(steps_per_epoch):
                idx = perm[step * self.batch_size:(step + 1) * self.batch_size]
                adjusted = epoch * steps_per_epoch + step
                lr = warmup_lr(adjusted, self.lr, self.warmup_steps)
                loss = ts.step(lat[idx].to(ts.dev, non_blocking=True).float(), lab[idx].to(ts.dev, non_blocking=True), lr)
                if (adjusted + 1) % self.log_interval == 0 or step == steps_per_epoch - 1:
                    acc.append((adjusted, float(loss), float(ts.grad_norm), lr))
            self.history += acc
            if self.checkpoints_dir is not None:
                save_train_checkpoint(os.path.join(self.checkpoints_dir, f"unet-epoch-{epoch:02}.pt"), ts, epoch,
                                      warmup_lr((epoch + 1) * steps_per_epoch, self.lr, self.warmup_steps))
            self.curr_epoch = epoch + 1
        return self.history
