"""Intra-kernel timeline of attention_pipe_kernel (CTA 0): run with the trace build of the library,
  python image-diffusion_b200/csrc/build.py --trace && IDF_B200_LIB=image-diffusion_b200/idf_b200/libidf_b200_trace.so \
  python tools/trace_attn.py [head_dim]
Prints, per key block, the SM-clock stamps of the MMA-issuing thread and of lane 0 of each softmax group's first warp."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "image-diffusion_b200"))
import numpy as np
import torch
from idf_b200 import ops

hd = int(sys.argv[1]) if len(sys.argv) > 1 else 32
B, T, heads = 96, 1024, 8
C = heads * hd; M = B * T
qkv = torch.randn(M, 3 * C, device="cuda").to(torch.bfloat16)
out = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attention_qkv(qkv, out, M, T, heads, hd)
torch.cuda.synchronize()
ptr = int(open("/tmp/idf_attn_trace_ptr").read())
NB, NE = 32, 8
buf = (ctypes.c_longlong * (3 * NB * NE))()
ctypes.CDLL("libcudart.so.12").cudaMemcpy(buf, ctypes.c_void_p(ptr), ctypes.c_size_t(3 * NB * NE * 8), 2)
a = np.array(buf).reshape(3, NB, NE)
t0 = a[a > 0].min()
names = {0: "mma", 1: "softmax g0", 2: "softmax g1"}
evn = {0: ["S_0 issue", "S_1 issue", "wait P_0", "wait P_1", "P_0 ready", "P_1 ready", "PV_0 issued", "PV_1 issued"],
       1: ["wait S", "S in TMEM", "S in regs", "max done / wait PV(e-1)", "PV(e-1) done", "exp issued", "P handed over", ""]}
ev = []
for who in range(3):
    for j in range(NB):
        for e in range(NE):
            if a[who, j, e]:
                ev.append((int(a[who, j, e] - t0), names[who], j, evn[min(who, 1)][e]))
for t, w, j, e in sorted(ev):
    if 8 <= j < 14:
        print(f"{t:8d}  {w:12s} e={j:2d}  {e}")
# per-phase durations of the softmax groups (steady state: blocks 8..23)
for g in (1, 2):
    d = a[g, 8:24].astype(np.int64)
    print(f"{names[g]}: period {np.diff(d[:, 0]).mean():.0f} clk | wait S {np.mean(d[:, 1] - d[:, 0]):.0f} | TMEM->regs "
          f"{np.mean(d[:, 2] - d[:, 1]):.0f} | max {np.mean(d[:, 3] - d[:, 2]):.0f} | wait PV {np.mean(d[:, 4] - d[:, 3]):.0f} | "
          f"exp {np.mean(d[:, 5] - d[:, 4]):.0f} | st wait + hand over {np.mean(d[:, 6] - d[:, 5]):.0f}")
m = a[0, 8:24].astype(np.int64)
print(f"mma: wait P_0 {np.mean(m[:, 4] - m[:, 2]):.0f} | P_0 ready -> PV_0 issued {np.mean(m[:, 6] - m[:, 4]):.0f} | "
      f"wait P_1 {np.mean(m[:, 5] - m[:, 3]):.0f} | P_1 ready -> PV_1 issued {np.mean(m[:, 7] - m[:, 5]):.0f}")
