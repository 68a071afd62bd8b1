"""Intercept and slope of the persistent igemm at the 4x4-stage size: GPU time of one launch against K."""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "image-diffusion_b200"))
import torch
from idf_b200 import ops
dev = "cuda"
for M, N in ((1536, 512), (6144, 512), (24576, 256)):
    for K in (64, 128, 256, 512, 1024, 2048, 4096):
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(torch.bfloat16)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        run = lambda: ops.igemm([(a, (1, 1, M), K, 1)], w, N, out)
        for _ in range(3): run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            torch.cuda._sleep(int(3e7))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); run(); run(); run(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / 4)
        print(f"M={M} N={N} K={K}: {sorted(ts)[3]:.2f} us per launch (4 back-to-back)")
