#!/bin/bash
# final one-GPU pass of round 2: GPU tests, smoke, every bench workload, reference arm. Outputs gpurun_out/f1_*.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/f1_tests.log 2>&1; echo "tests rc=$?"; tail -n 5 gpurun_out/f1_tests.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/f1_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/f1_smoke.log | cut -c1-300
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/f1_bench.json 2> gpurun_out/f1_bench.err; echo "bench rc=$?"; tail -n 2 gpurun_out/f1_bench.err | cut -c1-300
timeout 400 python bench.py --impl reference --steps 6 --warmup 2 > gpurun_out/f1_reference.json 2> gpurun_out/f1_reference.err; echo "reference rc=$?"
timeout 300 python bench.py --workload train --steps 20 --warmup 5 > gpurun_out/f1_train.json 2> gpurun_out/f1_train.err; echo "train rc=$?"; tail -n 2 gpurun_out/f1_train.err | cut -c1-300
timeout 300 python bench.py --workload vq --steps 5 --warmup 3 > gpurun_out/f1_vq.json 2> gpurun_out/f1_vq.err; echo "vq rc=$?"; tail -n 2 gpurun_out/f1_vq.err | cut -c1-300
timeout 300 python bench.py --workload shard --total 4096 --micro-batch 128 --sample-steps 50 > gpurun_out/f1_shard.json 2> gpurun_out/f1_shard.err; echo "shard rc=$?"; tail -n 2 gpurun_out/f1_shard.err | cut -c1-300
python - <<'PY'
import json
def L(f):
    try: return json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    except Exception as e: return {"err": str(e)}
b = L("f1_bench"); print("sample", b.get("value"), b.get("ms_per_step"), "e2e", b.get("e2e", {}).get("value"), b.get("parity"), {k: v for k, v in b.get("config", {}).items() if k in ("kl_decode_ms_batch48", "full_job_s", "pct_tensor_peak_sustained")}, b.get("gpu_torch_baseline", {}).get("bf16_autocast"), b.get("cpu_baseline", {}).get("value"))
print({k: v["ms"] for k, v in b.get("kernel_breakdown_ms_per_step", {}).items()}); r = b.get("roofline", {}); print({k: r.get(k) for k in ("achieved", "frac", "frac_of_sustained_peak", "traffic", "share_of_step")})
r = L("f1_reference"); print("reference", r.get("value"), r.get("steps"), r.get("config", {}).get("per_gpu_batch"), r.get("cpu_baseline", {}).get("cores"))
t = L("f1_train"); print("train", t.get("value"), t.get("ms_per_step"), t.get("config", {}).get("pct_tensor_peak_sustained")); print({k: v["ms"] for k, v in list(t.get("kernel_breakdown_ms_per_step", {}).items())[:12]})
v = L("f1_vq"); print("vq", v.get("value"), v.get("ms_per_step"), v.get("e2e", {}).get("value"), v.get("config", {}).get("stages"))
s = L("f1_shard"); print("shard", s.get("value"), s.get("e2e", {}).get("value"), s.get("config", {}).get("job_ms_with_decode"), s.get("clocks"))
PY
