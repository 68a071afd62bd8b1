"""One shape of idf_attention_fwd_qkv, a few launches (ncu target)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "image-diffusion_b200"))
import torch
from idf_b200 import ops
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
hd = int(sys.argv[2]) if len(sys.argv) > 2 else 32
B, heads = 96, 8
C = heads * hd; M = B * T
qkv = torch.randn(M, 3 * C, device="cuda").to(torch.bfloat16)
out = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
for _ in range(5): ops.attention_qkv(qkv, out, M, T, heads, hd)
torch.cuda.synchronize()
print("ok")
