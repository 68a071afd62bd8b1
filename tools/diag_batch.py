"""Diagnostic (GPU): run-to-run determinism and batch invariance of the UNet engine, layer by layer."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "image-diffusion_b200"))
import torch
from oracle import ref_path as O
from modules.unet import Unet
torch.set_grad_enabled(False)
dev = "cuda"
sd = O.seeded_state_dict(O.unet_param_shapes(O.UNET_ARCH), 2018)
m = Unet(**O.UNET_ARCH); m.load_state_dict(sd); m = m.to(dev).eval()
g = torch.Generator().manual_seed(77)
x = torch.randn(3, 3, 32, 32, generator=g).to(dev)
t = torch.full((3,), 640, device=dev)
ctx = torch.tensor([0, 1, 2], device=dev)

def run(xx, tt, cc, mask):
    e = m.engine(xx.shape[0], 32, 32)
    e.taps = {}
    out = m(xx, tt, cc, mask)
    taps = e.taps
    e.taps = None
    return out, taps

a, ta = run(x, t, ctx, None)
b, tb = run(x, t, ctx, None)
print("run-to-run identical:", torch.equal(a, b), "max diff", (a - b).abs().max().item())
for k in ta:
    if not torch.equal(ta[k], tb[k]):
        print("  first nondeterministic tap:", k, (ta[k] - tb[k]).abs().max().item()); break
c, tc = run(torch.cat([x, x]), torch.cat([t, t]), torch.cat([ctx, ctx]), torch.tensor([[1.], [1.], [1.], [0.], [0.], [0.]], device=dev))
print("B=6 cond half vs B=3: rel", ((c[:3] - a).norm() / a.norm()).item())
for k in ta:
    A, C = ta[k], tc[k]
    if k == "table":
        d = (A - C[:3]).abs().max().item()
    elif k.endswith(".qk") or True:
        rows = A.shape[0]
        d = (A - C[:rows]).abs().max().item() if A.shape[1] == C.shape[1] else float("nan")
    print(f"  {k:40s} shape {tuple(A.shape)} maxdiff {d:.3e}  (scale {A.abs().max().item():.2f})")
    if d > 1e-2 * A.abs().max().item() and d == d:
        pass
