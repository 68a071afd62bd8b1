#!/bin/bash
# Runs every kernel-level GPU test group in its own process (a device-side trap poisons the CUDA context, so
# isolation keeps the remaining groups informative). Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
rc_all=0
for k in conv3x3_igemm split_k rowbias linear_residual small_m qkv_attention groupnorm embed small_channel downsample cfg_posterior vq_argmin; do
  timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "$k" > gpurun_out/kt_$k.log 2>&1
  rc=$?
  echo "== $k rc=$rc"; tail -n 15 gpurun_out/kt_$k.log | cut -c1-300
  [ $rc -ne 0 ] && rc_all=1
done
exit $rc_all
