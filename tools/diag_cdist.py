"""Diagnostic (GPU): which fp32 evaluation order reproduces torch.cdist's matmul path bit for bit?"""
import itertools
import torch

torch.manual_seed(0)
dev = "cuda"
B, P, R, D = 4, 1024, 1024, 3
x = torch.randn(B, P, D, device=dev)
e = (torch.rand(R, D, device=dev) * 2 - 1) / R
e2 = torch.randn(R, D, device=dev)

def norms_candidates(v):
    s = v * v
    return {
        "(0+1)+2": (s[..., 0] + s[..., 1]) + s[..., 2],
        "(0+2)+1": (s[..., 0] + s[..., 2]) + s[..., 1],
        "0+(1+2)": s[..., 0] + (s[..., 1] + s[..., 2]),
        "fma_seq": torch.addcmul(torch.addcmul(s[..., 0], v[..., 1], v[..., 1]), v[..., 2], v[..., 2]),
    }

for name, v in (("x", x), ("e_small", e), ("e_normal", e2)):
    ref = v.pow(2).sum(-1)
    for k, c in norms_candidates(v).items():
        print(f"norm[{name}] {k}: mismatches {(c != ref).sum().item()} / {ref.numel()}")

def fma(a, b, c):  # fp32 fma emulated in fp64 (exact product, one rounding of the sum + one to fp32)
    return (a.double() * b.double() + c.double()).float()

for cname, cb in (("small", e), ("normal", e2)):
    x1n = x.pow(2).sum(-1, keepdim=True)
    x2n = cb.pow(2).sum(-1, keepdim=True)
    x1_ = torch.cat([x.mul(-2), x1n, torch.ones_like(x1n)], -1)            # (B,P,5)
    x2_ = torch.cat([cb, torch.ones_like(x2n), x2n], -1)[None].repeat(B, 1, 1)  # (B,R,5)
    ref = x1_.matmul(x2_.mT)
    a = x1_[:, :, None, :]  # B,P,1,5
    b = x2_[:, None, :, :]  # B,1,R,5
    for perm in [(0, 1, 2, 3, 4), (4, 3, 2, 1, 0), (3, 4, 0, 1, 2), (0, 1, 2, 4, 3)]:
        acc = torch.zeros(B, P, R, device=dev)
        for k in perm:
            acc = fma(a[..., k], b[..., k], acc)
        print(f"matmul[{cname}] fma order {perm}: mismatches {(acc != ref).sum().item()} / {ref.numel()}")
    # non-fused multiply-add, sequential
    acc = torch.zeros(B, P, R, device=dev)
    for k in range(5):
        acc = acc + a[..., k] * b[..., k]
    print(f"matmul[{cname}] mul+add seq: mismatches {(acc != ref).sum().item()}")
    # pairwise: (p0+p1) + (p2+p3) + p4 with fma
    d = torch.cdist(x, cb[None].repeat(B, 1, 1))
    print(f"cdist == sqrt(clamp(ref)) : {(d != ref.clamp_min(0).sqrt()).sum().item()} mismatches")
    print("allow_tf32 matmul:", torch.backends.cuda.matmul.allow_tf32, torch.get_float32_matmul_precision())
