#!/bin/bash
# final one-GPU pass after the GroupNorm-fused epilogue: everything of r2_final1.sh, then the ncu launch list of one
# eager sampling step and a --set full capture of the fused conv kernels. Outputs gpurun_out/f1_* and gpurun_out/p3_*.
tools/r2_final1.sh
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 300 python tools/profile_step.py 3 > gpurun_out/p3_step_plain.log 2>&1 || { echo "profile_step failed"; tail gpurun_out/p3_step_plain.log; exit 0; }
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/p3_launches_sample.csv \
  python tools/profile_step.py 3 > gpurun_out/p3_ncu_sample.log 2>&1; echo "ncu sample list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"igemm_persist" -c 6 \
  -o gpurun_out/p3_full_sample -f python tools/profile_step.py 3 > gpurun_out/p3_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/p3_* | cut -c1-150
