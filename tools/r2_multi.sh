#!/bin/bash
# Multi-GPU measurements on one 8 x B200 box: data-parallel training step at 1/2/4/8 ranks (NCCL gradient all-reduce),
# batch-sharded sampling of 4096 latents at 1/2/4/8 ranks. Logs and JSON lines in gpurun_out/m_*.
mkdir -p gpurun_out
run() { n=$1; shift; if [ "$n" = "1" ]; then python "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) "$@"; fi; }
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv > gpurun_out/m_smi.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29490 tools/ddp_check.py > gpurun_out/m_ddp_check.log 2>&1
rc=$?; echo "ddp_check rc=$rc"; tail -n 6 gpurun_out/m_ddp_check.log | cut -c1-300
if [ $rc -ne 0 ]; then exit 1; fi
for n in 8 4 2 1; do
  timeout 400 bash -c "$(declare -f run); run $n bench.py --workload train --gpus $n --steps 20 --warmup 5" > gpurun_out/m_train_${n}gpu.json 2> gpurun_out/m_train_${n}gpu.err
  echo "train n=$n rc=$?"; tail -n 2 gpurun_out/m_train_${n}gpu.err | cut -c1-300
done
for n in 8 4 2 1; do
  timeout 400 bash -c "$(declare -f run); run $n bench.py --workload shard --gpus $n --total 4096 --micro-batch 128 --sample-steps 20" > gpurun_out/m_shard_${n}gpu.json 2> gpurun_out/m_shard_${n}gpu.err
  echo "shard n=$n rc=$?"; tail -n 2 gpurun_out/m_shard_${n}gpu.err | cut -c1-300
done
python - <<'PY'
import json
for w in ("train", "shard"):
    for n in (1, 2, 4, 8):
        try:
            d = json.loads(open(f"gpurun_out/m_{w}_{n}gpu.json").read().strip().splitlines()[-1])
            print(w, n, round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 3), d["config"].get("allreduce"), d["clocks"])
        except Exception as e:
            print(w, n, "failed", e)
PY
