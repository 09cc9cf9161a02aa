#!/bin/bash
# ncu --set full of the conv-layout loss kernel (probabilities and logits) and the packed K3, each after a plain run
set -u
mkdir -p gpurun_out
H="python tools/bench_head_layout.py --steps 3"
$H > gpurun_out/plain_k3h.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'focal_loss_head_kernel' -s 2 -c 1 -o gpurun_out/r02_k3h --force-overwrite $H > gpurun_out/ncu_k3h.log 2>&1
tail -2 gpurun_out/ncu_k3h.log
B2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode --no-configs"
$B2 > gpurun_out/plain_loss2.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'focal_loss_kernel' -s 2 -c 1 -o gpurun_out/r02_loss_packed --force-overwrite $B2 > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log
ls -la gpurun_out/*.ncu-rep | tail -3
