#!/bin/bash
cat > /tmp/vae_time.py <<'PY'
import os, sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "image-diffusion_b200")
from idf_b200.spec import VAE_KL_ARCH, VAE_VQ_ARCH
from modules.vae import VAE
def t(fn, n=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
with torch.no_grad():
    torch.manual_seed(2018)
    vae = VAE(**VAE_KL_ARCH).cuda().eval()
    z = torch.randn(48, 3, 32, 32, device="cuda"); img = torch.rand(48, 3, 128, 128, device="cuda") * 2 - 1
    d, e = t(lambda: vae.decode(z)), t(lambda: vae.encode(img))
    print(f"KL decode b48 {d:.3f} ms ({48*65.633/d:.0f} TF/s)  encode b48 {e:.3f} ms ({48*141.3/e:.0f} TF/s)")
    vq = VAE(**VAE_VQ_ARCH).cuda().eval()
    img = torch.rand(256, 3, 128, 128, device="cuda") * 2 - 1
    f = t(lambda: vq(img), n=3)
    print(f"VQ forward b256 {f:.2f} ms ({256*206.933/f:.0f} TF/s, {256/f*1e3:.0f} img/s)")
PY
for cfg in "IDF_GN_SUB_MB=0" "IDF_GN_SUB_MB=32" "IDF_GN_SUB_MB=48" "IDF_GN_SUB_MB=64" "IDF_GN_SUB_MB=48 IDF_GN_ROWS=0" "IDF_GN_SUB_MB=0"; do
  echo "== $cfg"; env $cfg timeout 300 python /tmp/vae_time.py 2>&1 | tail -n 2
done
