#!/bin/bash
# packed-arithmetic (FFMA2) loss kernels: loss-side parity tests, then the three f1 benches and the main line
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_loss_gpu.py tests/test_guards_gpu.py tests/test_robustness_gpu.py -m gpu -x -q 2>&1 | tail -6
timeout 200 python tools/bench_head_layout.py > gpurun_out/bench_head_layout.json 2> gpurun_out/bhl.err; cat gpurun_out/bench_head_layout.json | cut -c1-600
timeout 200 python tools/bench_head_layout.py --logits >> gpurun_out/bench_head_layout.json 2>> gpurun_out/bhl.err; tail -1 gpurun_out/bench_head_layout.json | cut -c1-600
timeout 200 python tools/bench_logits.py > gpurun_out/bench_logits.json 2>/dev/null; cat gpurun_out/bench_logits.json | cut -c1-600
timeout 300 python bench.py --no-decode --no-cpu-baseline > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_j.json'))
print({k:d[k] for k in ('value','ms_per_step','host_enqueue_us_per_step')}, d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks'])
print(json.dumps({k:(round(v['ms_per_step'],4),round(v['value']),round(v['roofline_frac'],3)) for k,v in d['configs'].items()}))
PY
