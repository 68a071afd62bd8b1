"""Diagnostic (GPU): |x|^2 evaluation order inside torch.cdist when x is the permuted NCHW view the Codebook uses."""
import torch
dev = "cuda"
torch.manual_seed(1)
for (B, H) in ((4, 32), (256, 32), (2, 16)):
    z = torch.randn(B, 3, H, H, device=dev)
    x = z.permute(0, 2, 3, 1).reshape(B, H * H, 3)
    print("B", B, "H", H, "x strides", x.stride(), "contig", x.is_contiguous())
    p2 = x.pow(2)
    print(" pow strides", p2.stride())
    ref = p2.sum(-1)
    s = x * x
    cands = {"(0+1)+2": (s[..., 0] + s[..., 1]) + s[..., 2], "(0+2)+1": (s[..., 0] + s[..., 2]) + s[..., 1],
             "0+(1+2)": s[..., 0] + (s[..., 1] + s[..., 2])}
    for k, c in cands.items():
        print(f"  strided norm {k}: mismatches {(c != ref).sum().item()} / {ref.numel()}")
    xc = x.contiguous()
    refc = xc.pow(2).sum(-1)
    print("  contiguous-vs-strided norm mismatches:", (refc != ref).sum().item())
    # does cdist itself agree between the two layouts?
    e = (torch.rand(1024, 3, device=dev) * 2 - 1) / 1024
    d1 = torch.cdist(x, e[None].repeat(B, 1, 1))
    d2 = torch.cdist(xc, e[None].repeat(B, 1, 1))
    print("  cdist strided vs contiguous: value mismatches", (d1 != d2).sum().item(), "argmin mismatches",
          (d1.argmin(-1) != d2.argmin(-1)).sum().item())
