#!/bin/bash
# K3h launch-shape variants through the C ABI (two interleaved rounds) + the new reference-snapshot tests and predict legs
set -u
mkdir -p gpurun_out
: > gpurun_out/ab_k3h.jsonl
for round in 1 2; do
for v in in-tree "$@"; do
  if [ "$v" = in-tree ]; then timeout 200 python tools/bench_kernels.py --only head_probs,head_logits 2>/dev/null | tee -a gpurun_out/ab_k3h.jsonl
  else CLDET_LIBRARY=build/variants/libcldet_$v.so timeout 200 python tools/bench_kernels.py --only head_probs,head_logits 2>/dev/null | tee -a gpurun_out/ab_k3h.jsonl; fi
done; done
timeout 1200 python -m pytest tests/test_reference_snapshot.py tests/test_detect_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python - <<'PY' 2>&1 | tail -4
import json, sys, torch
sys.path.insert(0, '.')
from tools.bench_detect import measure_predict
dev = torch.device('cuda', 0)
for mu in (-10.5, -9.5):
    r = measure_predict(dev, mu, cpu_images=2)
    print(json.dumps({k: r.get(k) for k in ('candidates_per_image', 'device_resident_ms_per_image', 'gpu_eager_ms_per_image', 'reference_predict_same_gpu', 'reference_predict_cpu', 'cpu_baseline')})[:1500])
PY
