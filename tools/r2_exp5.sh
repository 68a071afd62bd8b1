#!/bin/bash
mkdir -p gpurun_out
L=$PWD/image-diffusion_b200/idf_b200
echo "== base"; timeout 200 python tools/time_attn_qkv.py 2>&1 | tail -n 4
echo "== nolsum"; IDF_B200_LIB=$L/libidf_b200_nolsum.so timeout 200 python tools/time_attn_qkv.py 2>&1 | tail -n 4
IDF_B200_LIB=$L/libidf_b200_nolsumt.so timeout 200 python tools/trace_attn.py 32 > gpurun_out/x5_trace_nolsum_hd32.txt 2>&1; tail -n 3 gpurun_out/x5_trace_nolsum_hd32.txt
IDF_B200_LIB=$L/libidf_b200_nolsumt.so timeout 200 python tools/trace_attn.py 16 > gpurun_out/x5_trace_nolsum_hd16.txt 2>&1; tail -n 3 gpurun_out/x5_trace_nolsum_hd16.txt
IDF_B200_LIB=$L/libidf_b200_trace.so timeout 200 python tools/trace_attn.py 16 > gpurun_out/x5_trace_base_hd16.txt 2>&1; tail -n 3 gpurun_out/x5_trace_base_hd16.txt
echo "== bench nolsum"; IDF_B200_LIB=$L/libidf_b200_nolsum.so timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-full-job --no-torch-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', d['ms_per_step'], 'parity', d['parity']['rel_rms'], {k: v['ms'] for k, v in list(d['kernel_breakdown_ms_per_step'].items())[:3]})"
timeout 300 python -m pytest tests/test_round2_gpu.py tests/test_kernels_gpu.py -q --timeout=600 2>&1 | tail -n 4
IDF_B200_LIB=$L/libidf_b200_nolsum.so timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_measured_configs_gpu.py tests/test_train_kernels_gpu.py -q --timeout=600 -k "attention" 2>&1 | tail -n 4
