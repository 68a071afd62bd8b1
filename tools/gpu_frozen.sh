#!/bin/bash
# usage: tools/gpu_frozen.sh <tag> <gpurun-timeout-s> [--gpus N] -- '<command run from the frozen copy>'
# Freezes the current tree under .frozen/<tag>/ (travels with the gpurun snapshot, which is taken only when the call
# leaves the queue) so the working tree can keep changing while the call waits. Outputs: gpurun_out/ of the main tree.
tag=$1; tmo=$2; shift 2
extra=()
while [ "$1" != "--" ]; do extra+=("$1"); shift; done
shift
cmd="$1"
dst=/root/repo/.frozen/$tag
rm -rf "$dst"; mkdir -p "$dst"
tar -C /root/repo --exclude=./.git --exclude=./gpurun_out --exclude=./.frozen --exclude='./image-diffusion_b200/csrc/build' \
    --exclude='__pycache__' --exclude=./.pytest_cache -cf - . | tar -C "$dst" -xf -
/root/repo/tools/gpurun_retry.sh --timeout "$tmo" "${extra[@]}" -- "cd .frozen/$tag && mkdir -p gpurun_out && ( $cmd ); cp -r gpurun_out/. \$GRAFT_REPO_ROOT/gpurun_out/ 2>/dev/null; true"
rc=$?
rm -rf "$dst"
exit $rc
