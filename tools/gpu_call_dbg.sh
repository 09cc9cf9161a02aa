#!/bin/bash
# scratch GPU call: K4 append A/B (one atomic per block vs per warp)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv,noheader
show() { python -c "import json,sys;d=json.load(open('$1'));print('$2',round(d['ms_per_step'],4),{k:round(v,4) for k,v in d['stage_ms'].items()},round(d['roofline']['frac'],4))"; }
for a in 1 0 1 0; do
CLDET_K4_BLOCK_APPEND=$a timeout 300 python tools/bench_detect.py --mu -4 > gpurun_out/det_dense_blk$a.json 2>/dev/null; show gpurun_out/det_dense_blk$a.json "mu=-4 block_append=$a"
CLDET_K4_BLOCK_APPEND=$a timeout 300 python tools/bench_detect.py --mu -10.5 > gpurun_out/det_sparse_blk$a.json 2>/dev/null; show gpurun_out/det_sparse_blk$a.json "mu=-10.5 block_append=$a"
done
timeout 900 python -m pytest tests/test_detect_gpu.py tests/test_guards_gpu.py tests/test_robustness_gpu.py -m gpu -x -q 2>&1 | tail -5 | cut -c1-400
