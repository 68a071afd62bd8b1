#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/gn_trace.txt
FUSE_LIST=2 tools/r2_gn.sh
export IDF_B200_LIB=image-diffusion_b200/idf_b200/libidf_b200_gntrace.so
for args in "32 128 256 1" "32 128 256 2" "32 128 128 1"; do
  timeout 120 python tools/trace_gn.py $args >> gpurun_out/gn_trace.txt 2>&1 || echo "trace $args failed"
done
cat gpurun_out/gn_trace.txt | cut -c1-300
