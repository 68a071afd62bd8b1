"""Round-2 exploratory timings on one B200 (not a bench): CFG step time vs micro-batch, the torch oracle on the same
GPU (fp32 eager / bf16 autocast), VAE decode / encode / VQ forward at the BASELINE batch sizes."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "image-diffusion_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from idf_b200.sampler import CfgSampler  # noqa: E402
from modules.components import Scheduler  # noqa: E402
from modules.unet import Unet  # noqa: E402
from modules.vae import VAE  # noqa: E402
from oracle import ref_path as O  # noqa: E402

dev = "cuda:0"
out = {}


def ev_time(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    torch.manual_seed(2018)
    unet = Unet(**O.UNET_ARCH).to(dev).eval()
    sched = Scheduler(1000, device=dev)
    for N in (48, 64, 96, 128, 192, 256):
        labels = torch.tensor(([0, 1, 2] * N)[:N], device=dev)
        cfg = torch.full((N,), 3, device=dev)
        s = CfgSampler(unet, sched, labels, cfg, (3, 32, 32))
        s.set_latent(torch.randn(N, 3, 32, 32, device=dev))
        k = [999]

        def step():
            s.step(k[0])
            k[0] = k[0] - 1 if k[0] > 1 else 999

        ms = ev_time(step, n=30, warm=5)
        out[f"cfg_step_mb{N}"] = {"ms": ms, "img_steps_per_s": N / ms * 1e3, "pct_sustained": N / ms * 1e3 * 45.51e9 / 1373.8e12}
        print(N, out[f"cfg_step_mb{N}"], flush=True)
        del s
        unet._engines.clear()
        torch.cuda.empty_cache()

    # torch oracle on the same GPU
    usd = {k: v.to(dev) for k, v in unet.state_dict().items()}
    N = 48
    x = torch.randn(N, 3, 32, 32, device=dev)
    z = torch.randn(N, 3, 32, 32, device=dev)
    labels = torch.tensor(([0, 1, 2] * N)[:N], device=dev)
    cfg = torch.full((N,), 3, device=dev)
    st = O.SchedulerTables(1000, device=dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    f = lambda: O.cfg_sample(usd, O.UNET_ARCH, st, x, labels, cfg, [z], steps=[500])
    ms = ev_time(f, n=5, warm=2)
    out["torch_fp32_eager_cfg_step_mb48"] = {"ms": ms, "img_steps_per_s": N / ms * 1e3}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ms = ev_time(f, n=5, warm=2)
    out["torch_bf16_autocast_cfg_step_mb48"] = {"ms": ms, "img_steps_per_s": N / ms * 1e3}
    print(out, flush=True)

    # VAE
    torch.manual_seed(2018)
    vae = VAE(**O.VAE_KL_ARCH).to(dev).eval()
    z0 = torch.randn(48, 3, 32, 32, device=dev)
    ms = ev_time(lambda: vae.decode(z0), n=5, warm=2)
    out["kl_decode_b48"] = {"ms": ms, "tflops": 48 * 65.633e9 / ms / 1e9}
    img = torch.rand(48, 3, 128, 128, device=dev) * 2 - 1
    ms = ev_time(lambda: vae.encode(img), n=5, warm=2)
    out["kl_encode_b48"] = {"ms": ms, "tflops": 48 * 141.3e9 / ms / 1e9}
    print(out, flush=True)
    vq = VAE(**O.VAE_VQ_ARCH).to(dev).eval()
    for B in (64, 256):
        img = torch.rand(B, 3, 128, 128, device=dev) * 2 - 1
        t0 = time.time()
        try:
            ms = ev_time(lambda: vq(img), n=2, warm=1)
            out[f"vq_forward_b{B}"] = {"ms": ms, "img_per_s": B / ms * 1e3, "tflops": B * (141.3 + 65.633) * 1e9 / ms / 1e9,
                                      "mem_gb": torch.cuda.max_memory_allocated() / 1e9}
        except Exception as e:  # noqa: BLE001
            out[f"vq_forward_b{B}"] = {"error": repr(e)[:300]}
        print(B, out[f"vq_forward_b{B}"], time.time() - t0, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "r2_explore.json"), "w") as fh:
    json.dump(out, fh, indent=1)
