#!/bin/bash
# round-2 experiment 2: full GPU tests, VAE engine variants (whole-row GroupNorm, tensor-core tail conv), vq / shard bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/x2_tests.log 2>&1; echo "tests rc=$?"; tail -n 12 gpurun_out/x2_tests.log | cut -c1-300
grep -h "rel-RMS\|update difference\|PSNR" gpurun_out/x2_tests.log | head -40
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
for cfg in "IDF_GN_ROWS=0 IDF_VAE_TC_TAIL=0" "IDF_GN_ROWS=0 IDF_VAE_TC_TAIL=1" "IDF_GN_ROWS=1 IDF_GN_L2_MB=24" "IDF_GN_ROWS=1 IDF_GN_L2_MB=48" "IDF_GN_ROWS=1 IDF_GN_L2_MB=80" "IDF_GN_ROWS=1 IDF_GN_L2_MB=100000" "IDF_VAE_GRAPH=0"; do
  echo "== $cfg"; env $cfg timeout 300 python /tmp/vae_time.py 2>&1 | tail -n 2
done
timeout 300 python bench.py --workload vq --steps 5 --warmup 3 > gpurun_out/x2_vq.json 2> gpurun_out/x2_vq.err; echo "vq rc=$?"; tail -n 3 gpurun_out/x2_vq.err | cut -c1-300
timeout 300 python bench.py --workload shard --micro-batch 128 --sample-steps 20 --total 3072 > gpurun_out/x2_shard.json 2> gpurun_out/x2_shard.err; echo "shard rc=$?"; tail -n 3 gpurun_out/x2_shard.err | cut -c1-300
python - <<'PY'
import json
for f in ("x2_vq", "x2_shard"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], {k: v for k, v in d["config"].items() if k in ("stages", "tflops_per_gpu", "decode_ms_total", "job_ms_with_decode")})
        if "kernel_breakdown_ms_per_step" in d: print({k: v["ms"] for k, v in d["kernel_breakdown_ms_per_step"].items()})
    except Exception as e: print(f, "failed", e)
PY
