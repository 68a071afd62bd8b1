"""Runs a few eager (non-graph) CFG sampling steps at batch 48 — the command profiled with ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "image-diffusion_b200"))
import torch
from bench import build_models, BATCH
from idf_b200.sampler import CfgSampler
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
with torch.no_grad():
    unet, vae, sched = build_models("cuda")
    labels = torch.tensor([0, 1, 2] * (BATCH // 3), device="cuda")
    s = CfgSampler(unet, sched, labels, torch.full((BATCH,), 3, device="cuda"), (3, 32, 32), use_graph=False)
    s.set_latent(torch.randn(BATCH, 3, 32, 32, device="cuda"))
    for k in range(steps):
        if k == steps - 1:  # ncu --profile-from-start off: exactly ONE step is profiled
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        s.step(999 - k)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    if len(sys.argv) > 2:
        vae.decode(s.latent.clone())
        torch.cuda.synchronize()
print("ok")
