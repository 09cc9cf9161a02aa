#!/bin/bash
# round-2 profiling pass (1 GPU): launch lists + ncu --set full captures; every ncu run follows a plain run of the same command
set -u
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-decode --no-configs"
$B > gpurun_out/plain_loss.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_loss_launches.csv $B > gpurun_out/ncu_l.log 2>&1
B2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode --no-configs"
$B2 > gpurun_out/plain_loss2.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'focal_loss_kernel|gt_scatter_kernel' -c 2 -o gpurun_out/r02_loss --force-overwrite $B2 > gpurun_out/ncu_f.log 2>&1
D="python tools/bench_detect.py --mu -10.5 --steps 10 --warmup 3"
$D > gpurun_out/plain_det.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_detect_sparse_launches.csv $D > gpurun_out/ncu_d.log 2>&1
P="python tools/profile_predict.py --mu -9.5 --calls 3"
$P > gpurun_out/plain_pred.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_predict8k_launches.csv $P > gpurun_out/ncu_p.log 2>&1
$P > gpurun_out/plain_pred2.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'nms_resolve_stream_kernel|radix_sort_kernel|nms_mask_kernel' -s 3 -c 3 -o gpurun_out/r02_predict8k --force-overwrite $P > gpurun_out/ncu_pf.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_*.csv 2>/dev/null | tail; tail -2 gpurun_out/ncu_f.log gpurun_out/ncu_pf.log
