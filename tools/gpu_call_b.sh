#!/bin/bash
# round-2 GPU pass B (1 GPU): tests, host-overhead probe, bench, resolve-kernel A/B
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25
python __graft_entry__.py smoke 2>&1 | tail -1
python tools/bench_host.py > gpurun_out/bench_host.json 2> gpurun_out/bench_host.err; echo "host rc=$?"; cat gpurun_out/bench_host.json; tail -3 gpurun_out/bench_host.err
python tools/bench_detect.py --mu -10.5 > gpurun_out/detect_sparse_default.json 2>/dev/null; cut -c1-900 gpurun_out/detect_sparse_default.json; echo
CLDET_NMS_RESOLVE=stream python tools/bench_detect.py --mu -10.5 > gpurun_out/detect_sparse_stream.json 2>/dev/null; cut -c1-900 gpurun_out/detect_sparse_stream.json; echo
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_b.json'))
print({k:d[k] for k in ('value','ms_per_step','host_enqueue_us_per_step')}, d['roofline']['kernel_ms'], d['e2e']['value'])
print(json.dumps({k:(v['ms_per_step'],v['value']) for k,v in d['configs'].items()}))
print(json.dumps(d['decode'].get('predict_batch1_reference_mode'))[:3000])
print(json.dumps(d['decode'].get('nms_vs_torchvision')))
PY
