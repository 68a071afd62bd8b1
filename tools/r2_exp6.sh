#!/bin/bash
mkdir -p gpurun_out
L=$PWD/image-diffusion_b200/idf_b200
timeout 300 python -m pytest tests/test_round2_gpu.py -q --timeout=600 > gpurun_out/x6_round2_tests.log 2>&1; tail -n 3 gpurun_out/x6_round2_tests.log; grep -n "^E  \|rel-RMS" gpurun_out/x6_round2_tests.log | head -20
echo "== base"; timeout 200 python tools/time_attn_qkv.py 2>&1 | tail -n 4
echo "== lazymax"; IDF_B200_LIB=$L/libidf_b200_lazymax.so timeout 200 python tools/time_attn_qkv.py 2>&1 | tail -n 4
IDF_B200_LIB=$L/libidf_b200_lazymaxt.so timeout 200 python tools/trace_attn.py 32 > gpurun_out/x6_trace_lazymax_hd32.txt 2>&1; tail -n 3 gpurun_out/x6_trace_lazymax_hd32.txt
echo "== tests lazymax"; IDF_B200_LIB=$L/libidf_b200_lazymax.so timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_measured_configs_gpu.py tests/test_train_kernels_gpu.py tests/test_modules_gpu.py -q --timeout=600 -k "attention or unet or chain or sampl" 2>&1 | tail -n 4
echo "== bench lazymax"; IDF_B200_LIB=$L/libidf_b200_lazymax.so timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-full-job --no-torch-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', d['ms_per_step'], 'parity', d['parity']['rel_rms'], {k: v['ms'] for k, v in list(d['kernel_breakdown_ms_per_step'].items())[:3]})"
echo "== bench base"; timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-full-job --no-torch-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', d['ms_per_step'], 'parity', d['parity']['rel_rms'], {k: v['ms'] for k, v in list(d['kernel_breakdown_ms_per_step'].items())[:3]})"
