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
for v in poly2 poly4; do echo "== lib $v"; IDF_B200_LIB=$PWD/image-diffusion_b200/idf_b200/libidf_b200_$v.so timeout 200 python tools/time_attn_qkv.py 2>&1 | tail -n 4; done
IDF_B200_LIB=$PWD/image-diffusion_b200/idf_b200/libidf_b200_poly4t.so timeout 200 python tools/trace_attn.py 32 > gpurun_out/x4_trace_poly4_hd32.txt 2>&1; tail -n 4 gpurun_out/x4_trace_poly4_hd32.txt
for v in poly2 poly4; do echo "== bench with lib $v"; IDF_B200_LIB=$PWD/image-diffusion_b200/idf_b200/libidf_b200_$v.so timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-full-job --no-torch-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', d['ms_per_step'], 'parity', d['parity']['rel_rms'], {k: v['ms'] for k, v in list(d['kernel_breakdown_ms_per_step'].items())[:3]})"; done
timeout 300 python bench.py --workload shard --total 4096 --micro-batch 128 --sample-steps 50 > gpurun_out/x4_shard_1gpu.json 2> gpurun_out/x4_shard.err; echo "shard rc=$?"; python -c "
import json
d = json.loads(open('gpurun_out/x4_shard_1gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['config']['job_ms_with_decode'], d['config']['decode_ms_total'], d['clocks'])"
