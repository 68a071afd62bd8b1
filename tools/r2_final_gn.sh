#!/bin/bash
# final one-GPU pass of round 2: everything of r2_final1.sh (GPU tests, smoke, every bench workload, reference arm), then
# the per-layer A/B and the intra-kernel trace of the GroupNorm-fused epilogue. Outputs gpurun_out/f1_*, gn_layers.txt, gn_trace.txt
tools/r2_final1.sh
timeout 300 python tools/time_gn_fused.py 96 > gpurun_out/gn_layers.txt 2>&1; echo "layers rc=$?"; tail -n 1 gpurun_out/gn_layers.txt
rm -f gpurun_out/gn_trace.txt
export IDF_B200_LIB=image-diffusion_b200/idf_b200/libidf_b200_gntrace.so
for args in "32 128 256 1" "32 128 256 2" "32 256 256 1" "32 128 128 1" "8 384 512 1"; do
  timeout 120 python tools/trace_gn.py $args >> gpurun_out/gn_trace.txt 2>&1 || echo "trace $args failed"
done
grep -c tile gpurun_out/gn_trace.txt
