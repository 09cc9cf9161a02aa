#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for mu in -10.5 -4.0; do
timeout 120 python tools/bench_detect.py --mu $mu > gpurun_out/detect_pdl_$mu.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/detect_pdl_$mu.json'));print('pdl    mu=$mu',d['ms_per_step'],d['stage_ms'])"
CLDET_NO_PDL=1 timeout 120 python tools/bench_detect.py --mu $mu > gpurun_out/detect_nopdl_$mu.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/detect_nopdl_$mu.json'));print('no-pdl mu=$mu',d['ms_per_step'],d['stage_ms'])"
done
for v in 0 1; do
CLDET_NO_PDL=$v timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-decode --no-configs > gpurun_out/bench_pdl$v.json 2> gpurun_out/bench_pdl$v.err; echo "bench NO_PDL=$v rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench_pdl$v.json'))
print({k:d[k] for k in ('value','ms_per_step','host_enqueue_us_per_step')}, d['roofline']['kernel_ms'], d['roofline']['assign_kernel_ms'])"
done
timeout 120 python tools/bench_host.py; timeout 200 python tools/bench_api.py 2>/dev/null
