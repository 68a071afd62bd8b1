#!/bin/bash
# two-GPU check of the final code: default workload (one batch of 48 per rank) and the sharded 4096-latent job
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29493 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/m2_bench_2gpu.json 2> gpurun_out/m2_bench_2gpu.err; echo "bench2 rc=$?"; tail -n 2 gpurun_out/m2_bench_2gpu.err | cut -c1-300; cut -c1-600 gpurun_out/m2_bench_2gpu.json
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29494 bench.py --workload shard --gpus 2 --total 4096 --micro-batch 128 --sample-steps 50 > gpurun_out/m2_shard_2gpu.json 2> gpurun_out/m2_shard_2gpu.err; echo "shard2 rc=$?"; tail -n 2 gpurun_out/m2_shard_2gpu.err | cut -c1-300; cut -c1-600 gpurun_out/m2_shard_2gpu.json
