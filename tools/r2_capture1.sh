#!/bin/bash
# Round-2 first capture on one B200 (run under gpurun): GPU tests, bench line + per-op breakdown, reference arm, ncu launch list.
set -u
mkdir -p gpurun_out
TAG=${1:-r2a}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --breakdown gpurun_out/breakdown_$TAG.json > gpurun_out/bench_$TAG.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench_$TAG.log | cut -c1-1500
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1; echo "bench ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -c 900 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-library-baseline --min-seconds 0.01 > gpurun_out/ncu_list_$TAG.log 2>&1
echo "ncu list rc=$?"
