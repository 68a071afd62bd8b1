#!/bin/bash
# BASELINE configs[2] on one GPU: micro-batch sweep of the sharded job (3072 latents, 30 steps, per-sample CFG scales 1,3,5,7,9)
mkdir -p gpurun_out
for mb in 48 96 128 256; do
  timeout 300 python bench.py --workload shard --total 3072 --micro-batch $mb --sample-steps 30 > gpurun_out/sweep_shard_mb$mb.json 2>/dev/null
done
python - <<'PY'
import json
out = {}
for mb in (48, 96, 128, 256):
    d = json.loads(open(f"gpurun_out/sweep_shard_mb{mb}.json").read().strip().splitlines()[-1])
    out[mb] = {"img_steps_per_s_sampling": d["value"], "img_steps_per_s_whole_job": d["e2e"]["value"], "job_ms": d["config"]["job_ms_with_decode"], "decode_ms": d["config"]["decode_ms_total"], "clocks": d["clocks"]}
    print(mb, out[mb])
json.dump({"what": "bench.py --workload shard --total 3072 --sample-steps 30 on one B200, per-sample CFG scales cycling through 1,3,5,7,9; sustained (power-capped) regime", "by_micro_batch": out}, open("gpurun_out/sweep_shard.json", "w"), indent=1)
PY
