"""GroupNorm(+SiLU) kernel timings at the UNet / VAE shapes: slab kernel vs whole-row kernels; embedding-table build."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "image-diffusion_b200"))
import torch
from idf_b200 import ops

def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

for B, HW, C in ((96, 1024, 256), (96, 1024, 128), (96, 1024, 512), (96, 256, 384), (96, 64, 512), (48, 16384, 128), (48, 4096, 256)):
    x = torch.randn(B * HW, C, device="cuda").to(torch.bfloat16)
    # a producer-like kernel writes x right before (L2-warm, as in the real step): copy from a twin buffer
    src = x.clone()
    y = torch.empty_like(x)
    g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    ws = torch.empty(B * ((HW + 255) // 256) * 32 * 2 + 16, device="cuda")
    a = t(lambda: ops.groupnorm_silu(x, y, g, b, B, HW, C, 32, True))
    r = t(lambda: ops.groupnorm_silu_rows(x, y, g, b, B, HW, C, 32, True, ws))
    def warm_a(): x.copy_(src); ops.groupnorm_silu(x, y, g, b, B, HW, C, 32, True)
    def warm_r(): x.copy_(src); ops.groupnorm_silu_rows(x, y, g, b, B, HW, C, 32, True, ws)
    def cp(): x.copy_(src)
    c0 = t(cp); wa = t(warm_a) - c0; wr = t(warm_r) - c0
    mb = 2 * x.numel() * 2 / 1e6
    print(f"B={B} HW={HW} C={C} ({mb:.0f} MB r+w): slab {a:6.1f} us ({mb / a:.2f} TB/s)  rows {r:6.1f} us | after a producer copy: slab {wa:6.1f}  rows {wr:6.1f}")

from bench import build_models
from idf_b200.sampler import CfgSampler
with torch.no_grad():
    unet, vae, sched = build_models("cuda")
    labels = torch.tensor([0, 1, 2] * 16, device="cuda")
    s = CfgSampler(unet, sched, labels, torch.full((48,), 3, device="cuda"), (3, 32, 32))
    s._ensure_table(); torch.cuda.synchronize()
    eng = s.engine
    def build():
        eng.__dict__["cfg_tables"].clear(); s._ensure_table()
    print(f"embedding table for 1000 timesteps x 4 class rows: {t(build, n=3) / 1e3:.2f} ms")
