#!/bin/bash
for v in 0 1 0 1; do echo "== IDF_IGEMM_BALANCE=$v"; IDF_IGEMM_BALANCE=$v timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-full-job --no-torch-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', round(d['ms_per_step'], 4), 'parity', d['parity']['rel_rms'], {k: v['ms'] for k, v in list(d['kernel_breakdown_ms_per_step'].items())[:3]}, d['clocks'])"; done
