#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for mu in -10.5 -4.0; do
python tools/bench_detect.py --mu $mu > gpurun_out/detect_fused_$mu.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/detect_fused_$mu.json'));print('cluster-fused mu=$mu',d['ms_per_step'],d['stage_ms'])"
done
