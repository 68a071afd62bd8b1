#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gn_fused_gpu.py -x -q --timeout=120 > gpurun_out/gn_tests.log 2>&1; echo "gn tests rc=$?"; tail -n 6 gpurun_out/gn_tests.log | cut -c1-300
timeout 900 python -m pytest tests/test_modules_gpu.py tests/test_round2_gpu.py -q --timeout=300 > gpurun_out/gn_modules.log 2>&1; echo "module tests rc=$?"; tail -n 6 gpurun_out/gn_modules.log | cut -c1-300
