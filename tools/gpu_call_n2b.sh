#!/bin/bash
# 2-GPU pass: sharded-loss correctness (ragged shards -> NCCL gather; equal shards -> fused peer exchange; straggler) + bench (both arms)
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
$TR tools/check_sharded.py 2>&1 | grep -v "^W\|^\[W\|^$" | tail -3
CLDET_EQUAL_SHARDS=1 $TR tools/check_sharded.py 2>&1 | grep -v "^W\|^\[W\|^$" | tail -3
CLDET_EQUAL_SHARDS=1 CLDET_STRAGGLER=1 $TR tools/check_sharded.py 2>&1 | grep -v "^W\|^\[W\|^$" | tail -4
$TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_r2_n2.err
$TR bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r2_n2_ref.json 2> gpurun_out/bench_r2_n2_ref.err; echo "ref rc=$?"
cut -c1-900 gpurun_out/bench_r2_n2_ref.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_n2.json'))
print({k:d[k] for k in ('value','ms_per_step','host_enqueue_us_per_step','n_gpus')}, d['roofline']['kernel_ms'], d['e2e']['value'], d['config']['collective'])
print(json.dumps(d.get('sharded_parity')))
print(json.dumps({k:(v['ms_per_step'],v['value'],v['roofline_frac']) for k,v in d['configs'].items()}))
PY
timeout 600 python -m pytest tests/test_reference_snapshot.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python - <<'PY' 2>&1 | tail -5
import json, sys, torch
sys.path.insert(0, '.')
from tools.bench_detect import measure_predict
dev = torch.device('cuda', 0)
for mu in (-10.5, -9.5):
    r = measure_predict(dev, mu, cpu_images=0)
    print(json.dumps({k: r[k] for k in ('candidates_per_image', 'e2e_ms_per_image', 'device_resident_ms_per_image', 'device_resident_per_call_ms', 'gpu_eager_ms_per_image')}))
PY
