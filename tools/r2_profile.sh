#!/bin/bash
# ncu evidence of the round-2 code (one GPU). Every profiled command first runs WITHOUT ncu and must exit 0.
# Outputs in gpurun_out/p2_*: launch lists (duration + DRAM bytes per launch), --set full reports of the top kernels.
mkdir -p gpurun_out
T=${1:-p2}
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 300 python tools/profile_step.py 3 > gpurun_out/${T}_step_plain.log 2>&1 || { echo "profile_step failed"; tail gpurun_out/${T}_step_plain.log; exit 1; }
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_launches_sample.csv \
  python tools/profile_step.py 3 > gpurun_out/${T}_ncu_sample.log 2>&1; echo "ncu sample list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"igemm_persist|attention_pipe|groupnorm_kernel" -c 14 \
  -o gpurun_out/${T}_full_sample -f python tools/profile_step.py 3 > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 300 python tools/profile_train_step.py 2 > gpurun_out/${T}_train_plain.log 2>&1 && \
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_launches_train.csv \
  python tools/profile_train_step.py 2 > gpurun_out/${T}_ncu_train.log 2>&1; echo "ncu train list rc=$?"
cat > /tmp/vq_once.py <<'PY'
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "image-diffusion_b200")
import os
os.environ["IDF_VAE_GRAPH"] = "0"
from idf_b200.spec import VAE_VQ_ARCH
from modules.vae import VAE
with torch.no_grad():
    torch.manual_seed(2018)
    vq = VAE(**VAE_VQ_ARCH).cuda().eval()
    img = torch.rand(64, 3, 128, 128, device="cuda") * 2 - 1
    vq(img); torch.cuda.synchronize()
    torch.cuda.profiler.start(); vq(img); torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("ok")
PY
timeout 300 python /tmp/vq_once.py > gpurun_out/${T}_vq_plain.log 2>&1 && \
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_launches_vq_b64.csv \
  python /tmp/vq_once.py > gpurun_out/${T}_ncu_vq.log 2>&1; echo "ncu vq list rc=$?"
ls -la gpurun_out/${T}_*
