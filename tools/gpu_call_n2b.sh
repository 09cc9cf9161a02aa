#!/bin/bash
# 2-GPU pass on the final tree: sharded-loss correctness (equal shards -> fused peer exchange) + bench with the driver's launch line
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
CLDET_EQUAL_SHARDS=1 $TR tools/check_sharded.py 2>&1 | grep -v "^W\|^\[W\|^$" | tail -3
$TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_r2_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_n2.json'))
print({k:d[k] for k in ('value','ms_per_step','host_enqueue_us_per_step','n_gpus')}, d['roofline']['kernel_ms'], d['e2e']['value'], d['config']['collective'])
print(json.dumps(d.get('sharded_parity')))
print(json.dumps({k:(v['ms_per_step'],v['value'],v['roofline_frac']) for k,v in d['configs'].items()}))
PY
