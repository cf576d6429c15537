#!/bin/bash
# CTA-pair kernel against the plan's current choice, per layer, batch 64 (run under gpurun)
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -q -k "cta_pair" 2>&1 | tail -3
timeout 900 python tools/conv_bench.py --variants auto,halo,pair --reps 7 --out gpurun_out/conv_bench_pair.json > gpurun_out/conv_bench_pair.log 2>&1; echo "bench rc=$?"
tail -3 gpurun_out/conv_bench_pair.log
