#!/bin/bash
for v in 0 1 0 1; do echo "== IDF_IGEMM_BALANCE=$v"; IDF_IGEMM_BALANCE=$v timeout 300 python bench.py --workload shard --total 2048 --micro-batch 128 --sample-steps 50 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), d['clocks'])"; done
