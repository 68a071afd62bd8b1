"""Intra-kernel timeline of the GroupNorm-fused igemm epilogue (CTA 0): run with the trace build,
  python image-diffusion_b200/csrc/build.py --variant gntrace -DIDF_GN_TRACE && \
  IDF_B200_LIB=image-diffusion_b200/idf_b200/libidf_b200_gntrace.so python tools/trace_gn.py H Cin Cout [mode] [batch]
Prints per tile the SM-clock intervals of the MMA-issuing thread and of the epilogue's elected thread."""
import ctypes
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "image-diffusion_b200"))
import numpy as np
import torch
from idf_b200 import ops

H, cin, cout = (int(v) for v in sys.argv[1:4])
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 1
B = int(sys.argv[5]) if len(sys.argv) > 5 else 96
M, K = B * H * H, 9 * cin
x = torch.randn(M, cin, device="cuda").to(torch.bfloat16)
w = (torch.randn(cout, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
bias, table = torch.randn(cout, device="cuda"), torch.randn(B, cout, device="cuda")
gamma, beta = torch.ones(cout, device="cuda"), torch.zeros(cout, device="cuda")
y, h = (torch.empty(M, cout, device="cuda", dtype=torch.bfloat16) for _ in range(2))
ws = torch.zeros(ops.gn_workspace_bytes(B, M, cout), device="cuda", dtype=torch.uint8)
g = dict(gamma=gamma, beta=beta, groups=32, silu=True, ws=ws)
if mode == 2:
    g["out"] = h
for _ in range(3):
    ops.igemm([(x, (B, H, H), cin, 9)], w, cout, y, bias=bias, rowbias=table, gn=g)
torch.cuda.synchronize()
ptr = int(open("/tmp/idf_gn_trace_ptr").read())
NT, NE = 8, 8
buf = (ctypes.c_longlong * (2 * NT * NE))()
ctypes.CDLL("libcudart.so.12").cudaMemcpy(buf, ctypes.c_void_p(ptr), ctypes.c_size_t(2 * NT * NE * 8), 2)
a = np.array(buf).reshape(2, NT, NE).astype(np.int64)
t0 = a[a > 0].min()
print(f"{H}x{H} Cin {cin} Cout {cout} mode {mode} batch {B} (clocks relative to the first stamp)")
for t in range(NT):
    m, e = a[0, t], a[1, t]
    if m[2] == 0 and e[7] == 0:
        continue
    print(f"tile {t}: MMA wait acc {m[1] - m[0]:6d} | mainloop {m[2] - m[1]:6d} (ends {m[2] - t0:7d}) || epilogue: wait acc {e[1] - e[0]:6d} | "
          f"pass 1 {e[2] - e[1]:5d} | publish {e[3] - e[2]:5d} | raw pass {e[4] - e[3]:5d} | poll + statistics {e[6] - e[4]:6d} | "
          f"norm pass {e[7] - e[6]:5d} | total {e[7] - e[1]:6d} (ends {e[7] - t0:7d})")
