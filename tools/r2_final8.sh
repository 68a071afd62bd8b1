#!/bin/bash
# final 8 x B200 pass: batch-sharded 4096-latent job (50 steps) at 8 / 4 / 2 ranks, training step at 8 / 4 / 2 ranks,
# default workload at 8 ranks. Outputs gpurun_out/f8_*.
mkdir -p gpurun_out
T() { n=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
for n in 8 4 2; do
  T $n bench.py --workload shard --gpus $n --total 4096 --micro-batch 128 --sample-steps 50 > gpurun_out/f8_shard_${n}gpu.json 2> gpurun_out/f8_shard_${n}gpu.err; echo "shard n=$n rc=$?"
  T $n bench.py --workload train --gpus $n --steps 20 --warmup 5 > gpurun_out/f8_train_${n}gpu.json 2> gpurun_out/f8_train_${n}gpu.err; echo "train n=$n rc=$?"
done
T 8 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/f8_sample_8gpu.json 2> gpurun_out/f8_sample_8gpu.err; echo "sample n=8 rc=$?"
timeout 200 python -m pytest tests/test_multi_gpu.py -q 2>&1 | tail -n 2
python - <<'PY'
import json
for w in ("shard", "train", "sample"):
    for n in (2, 4, 8):
        try:
            d = json.loads(open(f"gpurun_out/f8_{w}_{n}gpu.json").read().strip().splitlines()[-1])
            print(w, n, round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), d["config"].get("allreduce"), d["clocks"])
        except Exception as e:
            print(w, n, "failed", e)
PY
