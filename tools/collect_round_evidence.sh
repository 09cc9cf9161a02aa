#!/bin/bash
# One-GPU evidence pass of a round: tests, smoke, bench lines, ncu launch list + full captures, side benches.
# Run under gpurun from the repo root; everything lands in gpurun_out/ (summaries are then made with tools/summarise_profiles.py).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --steps 200 --warmup 20 > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-decode > gpurun_out/plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_loss_launches.csv \
      python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-decode > gpurun_out/ncu_l.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode > gpurun_out/plain2.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'focal_loss_kernel|gt_scatter_kernel' -c 2 -o gpurun_out/r01_final --force-overwrite \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode > gpurun_out/ncu_f.log 2>&1
python tools/bench_head_layout.py > gpurun_out/bench_head_layout.json 2> gpurun_out/bhl.err
python tools/bench_head_layout.py --logits >> gpurun_out/bench_head_layout.json 2>> gpurun_out/bhl.err
python tools/bench_head_layout.py --steps 2 > gpurun_out/plain3.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:focal_loss_head_kernel -c 1 -o gpurun_out/r01_head_final --force-overwrite \
      python tools/bench_head_layout.py --steps 2 > gpurun_out/ncu_h.log 2>&1
python tools/bench_detect.py --head > gpurun_out/detect_dense.json 2>/dev/null
python tools/bench_detect.py --mu -10.5 --head > gpurun_out/detect_sparse.json 2>/dev/null
python tools/bench_api.py > gpurun_out/bench_api.jsonl 2>/dev/null
python tools/bench_logits.py > gpurun_out/bench_logits.json 2>/dev/null
python tools/bench_distill.py > gpurun_out/bench_distill.json 2>/dev/null
python tools/sweep_configs.py > gpurun_out/sweep_configs.jsonl 2>/dev/null
cut -c1-400 gpurun_out/bench_r1_n1.json; echo; cat gpurun_out/bench_r1_ref.json | cut -c1-300; tail -2 gpurun_out/bench_r1_n1.err
