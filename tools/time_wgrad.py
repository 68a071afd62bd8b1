"""Times idf_conv2d_wgrad on a few layer shapes (CUDA events around 5 back-to-back launches)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "image-diffusion_b200"))
import torch
from idf_b200 import ops
shapes = [(48, 256, 256, 32, 9), (48, 512, 128, 32, 9), (48, 384, 384, 16, 9), (48, 512, 512, 8, 9), (48, 256, 768, 32, 1)]
ws = torch.empty(48 * 1024 * 1024, device="cuda")
for B, Cin, Cout, H, taps in shapes:
    M = B * H * H
    x = torch.randn(M, Cin, device="cuda").to(torch.bfloat16)
    dy = torch.randn(M, Cout, device="cuda").to(torch.bfloat16)
    k = 3 if taps == 9 else 1
    g = torch.empty(Cout, Cin, k, k, device="cuda")
    grid = (B, H, H) if taps == 9 else (1, 1, M)
    for _ in range(2):
        ops.conv_wgrad(x, grid, Cin, taps, dy, Cout, g, ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.conv_wgrad(x, grid, Cin, taps, dy, Cout, g, ws)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 5
    gf = 2.0 * M * Cin * Cout * taps / 1e9
    print(f"B={B} Cin={Cin} Cout={Cout} H={H} taps={taps}: {us:7.1f} us  {gf / us * 1e3:7.0f} TF/s")
