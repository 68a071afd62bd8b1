"""Times idf_attention_fwd_qkv (token-major QKV) at the UNet's attention shapes and reports the error against an
fp32 torch softmax attention on the same bf16 inputs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "image-diffusion_b200"))
import torch
from idf_b200 import ops
torch.manual_seed(0)
B, heads = 96, 8
for T, hd, sc in ((1024, 32, 1.0), (1024, 16, 1.0), (256, 48, 1.0), (256, 32, 3.0)):
    C = heads * hd; M = B * T
    qkv = (torch.randn(M, 3 * C, device="cuda") * sc).to(torch.bfloat16)
    out = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
    for _ in range(3): ops.attention_qkv(qkv, out, M, T, heads, hd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.attention_qkv(qkv, out, M, T, heads, hd)
    e1.record(); torch.cuda.synchronize()
    nb = 8
    q, k, v = [t.reshape(B, T, heads, hd)[:nb].permute(0, 2, 1, 3).float() for t in qkv.split(C, dim=1)]
    ref = torch.softmax(q @ k.transpose(-1, -2) / hd ** 0.5, dim=-1) @ v
    got = out.reshape(B, T, heads, hd)[:nb].permute(0, 2, 1, 3).float()
    rel = ((got - ref).norm() / ref.norm()).item()
    print(f"T={T} hd={hd} scale={sc}: {e0.elapsed_time(e1) / 10 * 1e3:7.1f} us  rel-rms {rel:.2e}")
