#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29490 tools/ddp_check.py > gpurun_out/m_ddp_check.log 2>&1
rc=$?; echo "ddp_check rc=$rc"; grep "ddp_check" gpurun_out/m_ddp_check.log | cut -c1-300; tail -n 3 gpurun_out/m_ddp_check.log | cut -c1-200
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29491 bench.py --workload train --gpus 2 --steps 20 --warmup 5 > gpurun_out/m_train_2gpu.json 2> gpurun_out/m_train_2gpu.err; echo "train2 rc=$?"; tail -n 2 gpurun_out/m_train_2gpu.err | cut -c1-300; cut -c1-900 gpurun_out/m_train_2gpu.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29492 bench.py --workload shard --gpus 2 --total 4096 --micro-batch 128 --sample-steps 20 > gpurun_out/m_shard_2gpu.json 2> gpurun_out/m_shard_2gpu.err; echo "shard2 rc=$?"; tail -n 2 gpurun_out/m_shard_2gpu.err | cut -c1-300; cut -c1-700 gpurun_out/m_shard_2gpu.json
