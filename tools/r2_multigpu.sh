#!/bin/bash
# Multi-GPU leg of round 2 (run under gpurun --gpus N): the H2D ceiling, BASELINE config 3 (strong scaling, host gather,
# equality with the 1-GPU result) and the weak-scaling bench line, all at N ranks.
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR tools/h2d_ceiling.py > gpurun_out/h2d_ceiling_${N}gpu.log 2>&1; echo "h2d rc=$?"; tail -1 gpurun_out/h2d_ceiling_${N}gpu.log
timeout 300 python tools/h2d_ceiling.py > gpurun_out/h2d_ceiling_1gpu_on_${N}box.log 2>&1; tail -1 gpurun_out/h2d_ceiling_1gpu_on_${N}box.log
timeout 900 $TR bench.py --config3 --gpus $N > gpurun_out/config3_${N}gpu.log 2>&1; echo "config3 rc=$?"; tail -1 gpurun_out/config3_${N}gpu.log
timeout 900 $TR bench.py --gpus $N --no-cpu-baseline --no-library-baseline > gpurun_out/bench_${N}gpu.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_${N}gpu.log | cut -c1-1500
