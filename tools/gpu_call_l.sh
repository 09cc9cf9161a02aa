#!/bin/bash
# bulk-copy (TMA) sweep of the conv-layout loss kernel: parity tests under a timeout, then the f1 benches
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_loss_gpu.py tests/test_guards_gpu.py tests/test_robustness_gpu.py -m gpu -x -q 2>&1 | tail -6
timeout 200 python tools/bench_head_layout.py > gpurun_out/bench_head_layout.json 2> gpurun_out/bhl.err; cat gpurun_out/bench_head_layout.json | cut -c1-600
timeout 200 python tools/bench_head_layout.py --logits >> gpurun_out/bench_head_layout.json 2>> gpurun_out/bhl.err; tail -1 gpurun_out/bench_head_layout.json | cut -c1-600
tail -3 gpurun_out/bhl.err
