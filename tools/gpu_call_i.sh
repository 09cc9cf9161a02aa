#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for mu in -10.5 -4.0; do
timeout 120 python tools/bench_detect.py --mu $mu > gpurun_out/det_$mu.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/det_$mu.json'));print('mu=$mu',d['ms_per_step'],d['stage_ms'],d['roofline']['frac'])"
done
timeout 200 python tools/bench_head_layout.py > gpurun_out/bench_head_layout.json 2> gpurun_out/bhl.err; cat gpurun_out/bench_head_layout.json | cut -c1-600
timeout 200 python tools/bench_head_layout.py --logits >> gpurun_out/bench_head_layout.json 2>> gpurun_out/bhl.err; tail -1 gpurun_out/bench_head_layout.json | cut -c1-600
timeout 200 python tools/bench_logits.py > gpurun_out/bench_logits.json 2>/dev/null; cat gpurun_out/bench_logits.json | cut -c1-600
timeout 200 python tools/bench_distill.py > gpurun_out/bench_distill.json 2>/dev/null; cat gpurun_out/bench_distill.json | cut -c1-400
