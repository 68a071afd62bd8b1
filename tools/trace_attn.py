import os, sys, ctypes
os.environ["IDF_ATTN_DBG"] = "16"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "image-diffusion_b200"))
import torch
from idf_b200 import ops
B, T, heads, hd = 96, 1024, 8, 32
C = heads * hd; M = B * T
qk = torch.randn(M, 2 * C, device="cuda").to(torch.bfloat16)
vt = torch.randn(C, M, device="cuda").to(torch.bfloat16)
out = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
for _ in range(3): ops.attention(qk, vt, out, M, T, heads, hd)
torch.cuda.synchronize()
ptr = int(open("/tmp/idf_attn_trace_ptr").read())
n = 5 * 16 * 8
buf = (ctypes.c_longlong * n)()
cudart = ctypes.CDLL("libcudart.so.12")
cudart.cudaMemcpy(buf, ctypes.c_void_p(ptr), ctypes.c_size_t(n * 8), 2)
import numpy as np
a = np.array(buf).reshape(5, 16, 8)
t0 = a[4, 0, 2]
print("kernel entry -> setup done:", a[4, 0, 0] - t0, " -> exit barrier:", a[4, 0, 1] - t0)
names = {0: "producer", 1: "mma", 2: "softmax g0", 3: "softmax g1"}
evn = {0: ["K issued", "V slot free"], 1: ["k_full ok", "S issued", "p_full ok", "v_full ok", "PV issued"],
       2: ["wait s_full", "s_full ok", "S in regs", "wait o_full", "o_full ok", "O folded", "p_full arrive"]}
ev = []
for who in range(4):
    for j in range(16):
        for e in range(8):
            if a[who, j, e]:
                ev.append((a[who, j, e] - t0, names[who], j, evn[min(who, 2)][e]))
for t, w, j, e in sorted(ev):
    print(f"{t:8d}  {w:12s} j={j}  {e}")
