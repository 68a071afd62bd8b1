#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/x4_tests.log 2>&1; echo "tests rc=$?"; tail -n 8 gpurun_out/x4_tests.log | cut -c1-300
for v in 0 1; do echo "== IDF_ATTN_F16P=$v"; IDF_ATTN_F16P=$v timeout 200 python tools/time_attn_qkv.py 2>&1 | tail -n 4; done
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/x4_bench.json 2> gpurun_out/x4_bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/x4_bench.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open("gpurun_out/x4_bench.json").read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "parity", d["parity"]["rel_rms"])
print({k: v for k, v in d["config"].items() if k in ("kl_decode_ms_batch48", "kl_decode_tflops", "full_job_s", "pct_tensor_peak_sustained")})
print({k: v["ms"] for k, v in d["kernel_breakdown_ms_per_step"].items()})
print(d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["traffic"])
PY
