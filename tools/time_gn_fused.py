"""Per-layer A/B on one B200: conv (+ time bias) followed by the standalone GroupNorm kernel vs the same conv with the
GroupNorm-fused epilogue (gn_mode 1: normalised output only; gn_mode 2: raw + normalised), for the conv1 / conv2 shapes
of the UNet's 32x32 and 16x16 stages at the bench batch (96 = CFG-doubled 48)."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "image-diffusion_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from idf_b200 import ops  # noqa: E402

DEV = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 96
# (H, Cin of the 3x3 segment, skip channels of the 1x1 segment, Cout)
SHAPES = [(32, 128, 0, 256), (32, 256, 128, 256), (32, 256, 0, 256), (32, 256, 256, 256), (32, 512, 0, 128), (32, 128, 512, 128),
          (32, 128, 0, 128), (32, 128, 128, 128), (16, 256, 0, 384), (16, 384, 256, 384), (16, 384, 0, 384), (16, 384, 384, 384),
          (16, 768, 0, 256), (16, 256, 768, 256), (16, 256, 0, 256), (16, 256, 256, 256),
          (8, 384, 0, 512), (8, 512, 384, 512), (8, 512, 0, 512), (8, 512, 512, 512),
          (8, 1024, 0, 384), (8, 384, 1024, 384), (8, 384, 0, 384), (8, 384, 384, 384)]


def timeit(fn, n=12, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


tot = [0.0, 0.0, 0.0]
for H, cin, cskip, cout in SHAPES:
    M = B * H * H
    x = torch.randn(M, cin, device=DEV).to(torch.bfloat16)
    segs = [(x, (B, H, H), cin, 9)]
    K = 9 * cin + cskip
    if cskip:
        segs.append((torch.randn(M, cskip, device=DEV).to(torch.bfloat16), (B, H, H), cskip, 1))
    w = (torch.randn(cout, K, device=DEV) / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(cout, device=DEV)
    table = torch.randn(B, cout, device=DEV)
    gamma, beta = torch.ones(cout, device=DEV), torch.zeros(cout, device=DEV)
    y, h = torch.empty(M, cout, device=DEV, dtype=torch.bfloat16), torch.empty(M, cout, device=DEV, dtype=torch.bfloat16)
    ws = torch.zeros(ops.gn_workspace_bytes(B, M, cout), device=DEV, dtype=torch.uint8)
    silu = cskip == 0
    t_conv = timeit(lambda: ops.igemm(segs, w, cout, y, bias=bias, rowbias=table))
    t_gn = timeit(lambda: ops.groupnorm_silu(y, h, gamma, beta, B, H * H, cout, 32, silu))
    t_both = timeit(lambda: (ops.igemm(segs, w, cout, y, bias=bias, rowbias=table),
                             ops.groupnorm_silu(y, h, gamma, beta, B, H * H, cout, 32, silu)))
    g1 = dict(gamma=gamma, beta=beta, groups=32, silu=silu, ws=ws)
    t_f1 = timeit(lambda: ops.igemm(segs, w, cout, h, bias=bias, rowbias=table, gn=g1))
    g2 = dict(g1, out=h)
    t_f2 = timeit(lambda: ops.igemm(segs, w, cout, y, bias=bias, rowbias=table, gn=g2))
    tot[0] += t_both; tot[1] += t_f1; tot[2] += t_f2
    print(f"{H}x{H} K={K:5d} N={cout:4d}: conv {t_conv:6.1f} + GN {t_gn:5.1f} us (back to back {t_both:6.1f}) | fused mode 1 "
          f"{t_f1:6.1f} | mode 2 {t_f2:6.1f} us", flush=True)
print(f"sum: separate {tot[0]:.0f} us, mode 1 {tot[1]:.0f} us, mode 2 {tot[2]:.0f} us")
