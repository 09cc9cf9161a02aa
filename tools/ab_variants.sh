#!/bin/bash
# A/B of libcldet build variants (tools/build_variant.py) on the default bench workload; one line per variant.
# usage: tools/ab_variants.sh name[:ENV=VAL] ...   -> gpurun_out/ab_variants.jsonl
mkdir -p gpurun_out
: > gpurun_out/ab_variants.jsonl
for spec in "$@"; do
  name="${spec%%:*}"; envs=""
  if [[ "$spec" == *:* ]]; then envs="${spec#*:}"; fi
  out=$(env CLDET_LIBRARY=build/variants/libcldet_${name}.so $envs python bench.py --no-cpu-baseline --no-decode --steps 200 --warmup 20 2>&1 | tail -1)
  python - "$spec" "$out" <<'PY' | tee -a gpurun_out/ab_variants.jsonl
import json, sys
spec, out = sys.argv[1], sys.argv[2]
try:
    d = json.loads(out)
    print(json.dumps({'variant': spec, 'ms_per_step': round(d['ms_per_step'], 5), 'loss_kernel_ms': round(d['roofline']['kernel_ms'], 5),
                      'assign_ms': round(d['roofline']['assign_kernel_ms'], 5), 'frac': round(d['roofline']['frac'], 4),
                      'sm_mhz': d['clocks']['sm_mhz'], 'reasons': d['clocks']['reasons']}))
except Exception as e:
    print(json.dumps({'variant': spec, 'error': repr(e), 'tail': out[-300:]}))
PY
done
