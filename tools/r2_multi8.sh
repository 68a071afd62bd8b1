#!/bin/bash
# 8 x B200 box: data-parallel training step and batch-sharded sampling at 8 and 4 ranks (2 ranks: tools/r2_ddp2.sh).
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv > gpurun_out/m_smi.txt 2>&1
for n in 8 4; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --workload train --gpus $n --steps 20 --warmup 5 > gpurun_out/m_train_${n}gpu.json 2> gpurun_out/m_train_${n}gpu.err
  echo "train n=$n rc=$?"; tail -n 2 gpurun_out/m_train_${n}gpu.err | cut -c1-300
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --workload shard --gpus $n --total 4096 --micro-batch 128 --sample-steps 20 > gpurun_out/m_shard_${n}gpu.json 2> gpurun_out/m_shard_${n}gpu.err
  echo "shard n=$n rc=$?"; tail -n 2 gpurun_out/m_shard_${n}gpu.err | cut -c1-300
done
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 bench.py --workload vq --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/m_vq_8gpu.json 2> gpurun_out/m_vq_8gpu.err; echo "vq n=8 rc=$?"
python - <<'PY'
import json
for w in ("train", "shard", "vq"):
    for n in (4, 8):
        try:
            d = json.loads(open(f"gpurun_out/m_{w}_{n}gpu.json").read().strip().splitlines()[-1])
            print(w, n, round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 3), d["config"].get("allreduce"), d["clocks"])
        except Exception as e:
            print(w, n, "failed", e)
PY
