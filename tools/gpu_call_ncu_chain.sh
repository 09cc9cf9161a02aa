#!/bin/bash
# ncu --set full of the post-filter chain (K5/K6) on the trained-like config-4 batch; raw page exported on the box
set -u
mkdir -p gpurun_out
D="python tools/bench_detect.py --mu -10.5 --steps 3 --warmup 3"
$D > gpurun_out/plain_chain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'rank_sort_kernel|nms_fused_kernel|select_fused_kernel' -s 9 -c 3 -o gpurun_out/r02_chain --force-overwrite $D > gpurun_out/ncu_chain.log 2>&1
ncu -i gpurun_out/r02_chain.ncu-rep --page raw --csv > gpurun_out/r02_chain_raw.csv 2>/dev/null
tail -3 gpurun_out/ncu_chain.log; ls -la gpurun_out/r02_chain_raw.csv gpurun_out/r02_chain.ncu-rep
