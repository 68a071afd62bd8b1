#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_round2_gpu.py tests/test_measured_configs_gpu.py -q --timeout=600 2>&1 | tail -n 6 | cut -c1-300
IDF_B200_LIB=$PWD/image-diffusion_b200/idf_b200/libidf_b200_trace.so timeout 200 python tools/trace_attn.py 32 > gpurun_out/x3_trace_hd32.txt 2>&1; tail -n 80 gpurun_out/x3_trace_hd32.txt
timeout 300 python bench.py --workload vq --steps 5 --warmup 3 > gpurun_out/x3_vq.json 2> gpurun_out/x3_vq.err; echo "vq rc=$?"; tail -n 3 gpurun_out/x3_vq.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open("gpurun_out/x3_vq.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["stages"], d["roofline"]["achieved"], d["gpu_launches"])
print({k: v["ms"] for k, v in d["kernel_breakdown_ms_per_step"].items()})
PY
tools/r2_profile.sh p2
