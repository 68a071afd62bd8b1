#!/bin/bash
mkdir -p gpurun_out
for v in 1 0 1 0; do echo "== IDF_GN_RES=$v"; IDF_GN_RES=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-full-job --no-torch-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', round(d['ms_per_step'], 4), {k: v['ms'] for k, v in list(d['kernel_breakdown_ms_per_step'].items())[:3]})"; done
timeout 300 python -m pytest tests/test_round2_gpu.py tests/test_kernels_gpu.py -q --timeout=600 2>&1 | tail -n 3
