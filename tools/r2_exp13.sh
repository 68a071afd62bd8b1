#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_measured_configs_gpu.py tests/test_round2_gpu.py tests/test_modules_gpu.py tests/test_train_step_gpu.py -q -s --timeout=600 > gpurun_out/x13_tests_verbose.log 2>&1; echo rc=$?
grep -h "rel-RMS\|PSNR\|update difference\|parity:\|differ" gpurun_out/x13_tests_verbose.log | cut -c1-220
tail -n 2 gpurun_out/x13_tests_verbose.log
