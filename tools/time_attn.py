import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "image-diffusion_b200"))
import torch
from idf_b200 import ops
B, T, heads, hd = 96, 1024, 8, int(sys.argv[1]) if len(sys.argv) > 1 else 32
C = heads * hd; M = B * T
qk = torch.randn(M, 2 * C, device="cuda").to(torch.bfloat16)
vt = torch.randn(C, M, device="cuda").to(torch.bfloat16)
out = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
for _ in range(3): ops.attention(qk, vt, out, M, T, heads, hd)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.attention(qk, vt, out, M, T, heads, hd)
e1.record(); torch.cuda.synchronize()
print(f"dbg={os.environ.get('IDF_ATTN_DBG','0')} hd={hd}: {e0.elapsed_time(e1)/10*1e3:.1f} us")
