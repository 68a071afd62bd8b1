"""Per-parameter gradient error of the kernel training path vs the fp32 oracle (debugging aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "image-diffusion_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
import test_train_step_gpu as T
from oracle import ref_path as Oref
arch = T.MID_ARCH if (len(sys.argv) < 2 or sys.argv[1] == "mid") else Oref.UNET_ARCH
B, res = (4, 16) if arch is T.MID_ARCH else (3, 32)
O, m, sd, x, noise, t, c, mask = T._setup(arch, B, res, 11)
loss = torch.nn.MSELoss()(m(x, t, context=c, context_mask=mask), noise)
loss.backward()
ref_loss, ref = T._oracle_grads(O, arch, sd, x, noise, t, c, mask)
eng = m.train_engine()
for k in eng.grad_names:
    g, r = dict(m.named_parameters())[k].grad, ref[k]
    rel = ((g - r).norm() / r.norm().clamp_min(1e-20)).item()
    flag = "  <<<<" if rel > 0.1 else ""
    print(f"{rel:10.3e} {r.norm().item():10.3e} {k}{flag}")
