#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_kernels_gpu.py tests/test_train_step_gpu.py -q --timeout=600 2>&1 | tail -n 4 | cut -c1-300
for v in 1 0 1 0; do echo "== IDF_TRAIN_SIDE=$v"; IDF_TRAIN_SIDE=$v timeout 300 python bench.py --workload train --steps 20 --warmup 5 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', round(d['ms_per_step'], 4), 'loss', d['config']['loss'], 'e2e', round(d['e2e']['value'], 1))"; done
timeout 300 python -m pytest tests/test_measured_configs_gpu.py tests/test_round2_gpu.py -q -k "train" --timeout=600 2>&1 | tail -n 3
