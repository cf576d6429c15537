#!/bin/bash
# one measurement step (run under gpurun): conv tests, per-layer bench of the given variants, the bench line
set -u
mkdir -p gpurun_out
TAG=${1:-x}
VARS=${2:-auto,pair}
timeout 600 python -m pytest tests/test_gpu_conv.py -q -x 2>&1 | tail -3
timeout 900 python tools/conv_bench.py --variants $VARS --reps 7 --out gpurun_out/conv_bench_$TAG.json > gpurun_out/conv_bench_$TAG.log 2>&1; echo "conv_bench rc=$?"
tail -1 gpurun_out/conv_bench_$TAG.log
timeout 600 python bench.py --no-cpu-baseline --no-library-baseline --breakdown gpurun_out/breakdown_$TAG.json > gpurun_out/bench_$TAG.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench_$TAG.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],4), "frac", round(d["roofline"]["frac"],4), "whole", round(d["roofline"]["frac_whole_step"],4), d["parity"]["graph_rows_equal_eager_api_chain"], d["clocks"])'
