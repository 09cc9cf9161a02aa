#!/bin/bash
# state-of-the-tree pass (1 GPU): GPU tests, smoke, the default bench line + reference arm, side benches
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 200 python tools/bench_kernels.py > gpurun_out/bench_kernels.json 2>gpurun_out/bk.err; cat gpurun_out/bench_kernels.json
timeout 300 python tools/bench_api.py > gpurun_out/bench_api.jsonl 2>gpurun_out/ba.err; cat gpurun_out/bench_api.jsonl | cut -c1-700
timeout 300 python tools/bench_detect.py --head > gpurun_out/detect_dense.json 2>gpurun_out/bd.err; cut -c1-900 gpurun_out/detect_dense.json
timeout 300 python tools/bench_detect.py --mu -10.5 --head > gpurun_out/detect_sparse.json 2>>gpurun_out/bd.err; cut -c1-900 gpurun_out/detect_sparse.json
cut -c1-1500 gpurun_out/bench_default.json; echo; cut -c1-600 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_default.err
