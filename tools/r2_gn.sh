#!/bin/bash
# GroupNorm-fused igemm epilogue: kernel tests, per-layer A/B, A/B of the sampling step (IDF_GN_FUSE = 0 / 2)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gn_fused_gpu.py -x -q --timeout=120 > gpurun_out/gn_tests.log 2>&1; echo "gn tests rc=$?"; tail -n 15 gpurun_out/gn_tests.log | cut -c1-400
timeout 300 python tools/time_gn_fused.py 96 > gpurun_out/gn_layers.txt 2>&1; echo "layers rc=$?"; cat gpurun_out/gn_layers.txt | cut -c1-200
for f in ${FUSE_LIST:-0 2}; do
  IDF_GN_FUSE=$f timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-full-job --no-torch-baseline > gpurun_out/gn_bench_$f.json 2> gpurun_out/gn_bench_$f.err; echo "bench fuse=$f rc=$?"
  python - <<PY
import json
try:
    b = json.loads(open("gpurun_out/gn_bench_$f.json").read().strip().splitlines()[-1])
    print("fuse=$f", b.get("value"), b.get("ms_per_step"), b.get("parity", {}).get("rel_rms"), b.get("gpu_launches"), {k: v["ms"] for k, v in b.get("kernel_breakdown_ms_per_step", {}).items()})
except Exception as e:
    print("fuse=$f failed", e); print(open("gpurun_out/gn_bench_$f.err").read()[-1500:])
PY
done
