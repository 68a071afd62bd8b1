"""Runs a few eager (non-graph) training steps at batch 48 — the command profiled with ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "image-diffusion_b200"))
import torch
from bench import build_models, BATCH
from idf_b200.trainer import DiffusionTrainStep
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
unet, _, sched = build_models("cuda")
unet.train()
ts = DiffusionTrainStep(unet, sched, BATCH, (3, 32, 32), use_graph=False)
lat = torch.randn(BATCH, 6, 32, 32, device="cuda"); lab = torch.randint(0, 3, (BATCH,), device="cuda")
for k in range(steps):
    if k == steps - 1:  # ncu --profile-from-start off: exactly ONE step is profiled
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    ts.step(lat, lab, 1e-4)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(ts.loss))
