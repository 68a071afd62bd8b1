#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_modules_gpu.py tests/test_round2_gpu.py tests/test_measured_configs_gpu.py -q --timeout=600 > gpurun_out/x10_tests.log 2>&1; echo "tests rc=$?"; tail -n 6 gpurun_out/x10_tests.log | cut -c1-300
for v in 1 0; do echo "== IDF_TC_TAIL=$v"; IDF_TC_TAIL=$v IDF_VAE_TC_TAIL=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-full-job --no-torch-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', round(d['ms_per_step'], 4), 'parity', d['parity']['rel_rms'], 'decode', d['config']['kl_decode_ms_batch48'], {k: v['ms'] for k, v in d['kernel_breakdown_ms_per_step'].items()})"; done
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -n 2
