#!/bin/bash
set -u
mkdir -p gpurun_out
bash tools/gpu_call_m.sh "$@" 2>&1 | cut -c1-520
H="python tools/bench_kernels.py --steps 3 --only head_probs"
$H > gpurun_out/plain_k3h.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'focal_loss_head_kernel' -s 3 -c 1 -o gpurun_out/r02_k3h_tma --force-overwrite $H > gpurun_out/ncu_k3h.log 2>&1
tail -n 2 gpurun_out/ncu_k3h.log
