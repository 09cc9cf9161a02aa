#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for mu in -10.5 -4.0; do
python tools/bench_detect.py --mu $mu > gpurun_out/detect_fused_$mu.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/detect_fused_$mu.json'));print('fused  mu=$mu',d['ms_per_step'],d['stage_ms'])"
CLDET_SELECT_MULTI=1 python tools/bench_detect.py --mu $mu > gpurun_out/detect_multi_$mu.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/detect_multi_$mu.json'));print('multi  mu=$mu',d['ms_per_step'],d['stage_ms'])"
done
CLDET_NMS_RESOLVE=stream python tools/bench_detect.py --mu -10.5 > gpurun_out/detect_stream.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/detect_stream.json'));print('stream mu=-10.5',d['ms_per_step'],d['stage_ms'])"
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-decode > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_d.json'))
print({k:d[k] for k in ('value','ms_per_step','host_enqueue_us_per_step')}, d['roofline']['kernel_ms'], d['roofline']['assign_kernel_ms'])
print(json.dumps({k:(v['ms_per_step'],v['value']) for k,v in d['configs'].items()}))
PY
