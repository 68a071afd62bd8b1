#!/bin/bash
# one-GPU validation pass: GPU tests, smoke, the four bench workloads. Logs in gpurun_out/.
mkdir -p gpurun_out
T=${1:-v1}
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -n 25 gpurun_out/${T}_tests.log | cut -c1-400
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 4 gpurun_out/${T}_smoke.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/${T}_bench.err | cut -c1-400
timeout 300 python bench.py --workload train --steps 20 --warmup 5 > gpurun_out/${T}_train.json 2> gpurun_out/${T}_train.err; echo "train rc=$?"; tail -n 3 gpurun_out/${T}_train.err | cut -c1-400
timeout 300 python bench.py --workload vq --steps 5 --warmup 3 > gpurun_out/${T}_vq.json 2> gpurun_out/${T}_vq.err; echo "vq rc=$?"; tail -n 3 gpurun_out/${T}_vq.err | cut -c1-400
for mb in 48 96 128 256; do
timeout 300 python bench.py --workload shard --micro-batch $mb --sample-steps 20 --total 3072 > gpurun_out/${T}_shard_mb$mb.json 2> gpurun_out/${T}_shard_mb$mb.err; echo "shard mb=$mb rc=$?"; tail -n 3 gpurun_out/${T}_shard_mb$mb.err | cut -c1-400
done
