#!/bin/bash
# A/B timings of round-2 changes (one GPU): attention exp-phase token, two-stream small stages, VAE engine.
mkdir -p gpurun_out
for tok in 0 1; do
  echo "== IDF_ATTN_TOKEN=$tok"; IDF_ATTN_TOKEN=$tok timeout 200 python tools/time_attn_qkv.py 2>&1 | tail -n 5
done
for sp in 1 2; do
  echo "== IDF_SPLIT_SMALL=$sp"
  IDF_SPLIT_SMALL=$sp timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-full-job --no-torch-baseline > gpurun_out/x1_bench_split$sp.json 2> gpurun_out/x1_bench_split$sp.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/x1_bench_split$sp.json"))
    print("ms_per_step", d["ms_per_step"], "value", d["value"], "parity", d["parity"]["rel_rms"], "decode_ms", d["config"]["kl_decode_ms_batch48"])
    print({k: v["ms"] for k, v in d["kernel_breakdown_ms_per_step"].items()})
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/x1_bench_split$sp.err").read()[-1500:])
PY
done
echo "== IDF_SPLIT_SMALL=2 IDF_SPLIT_HW=256"
IDF_SPLIT_HW=256 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-full-job --no-torch-baseline 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', d['ms_per_step'])"
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "batched or qkv_attention" 2>&1 | tail -n 5
timeout 600 python -m pytest tests/test_measured_configs_gpu.py -q -x -k "kl_decode or attention or unet_forward" 2>&1 | tail -n 8
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/x1_tests.log 2>&1; echo "tests rc=$?"; tail -n 30 gpurun_out/x1_tests.log | cut -c1-300
