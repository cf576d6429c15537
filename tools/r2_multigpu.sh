#!/bin/bash
# Multi-GPU leg of round 2 (run under gpurun --gpus 8): the pinned-H2D ceiling of the box, then BASELINE config 3 (strong
# scaling over 4096 images, host gather, equality with the 1-GPU result) at 2 / 4 / 8 ranks.
set -u
mkdir -p gpurun_out
TR() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
nvidia-smi topo -m > gpurun_out/topo_8gpu.txt 2>&1
lscpu | head -25 > gpurun_out/lscpu_8gpu.txt 2>&1
for N in 8 4 2; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N tools/h2d_ceiling.py > gpurun_out/h2d_ceiling_${N}gpu.log 2>&1; echo "h2d $N rc=$?"; tail -1 gpurun_out/h2d_ceiling_${N}gpu.log | cut -c1-400
done
timeout 200 python tools/h2d_ceiling.py > gpurun_out/h2d_ceiling_1gpu.log 2>&1; tail -1 gpurun_out/h2d_ceiling_1gpu.log | cut -c1-400
for N in 8 4 2; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --config3 --gpus $N > gpurun_out/config3_${N}gpu.log 2>&1; echo "config3 $N rc=$?"; tail -1 gpurun_out/config3_${N}gpu.log | cut -c1-900
done
timeout 300 python bench.py --config3 --gpus 1 > gpurun_out/config3_1gpu.log 2>&1; echo "config3 1 rc=$?"; tail -1 gpurun_out/config3_1gpu.log | cut -c1-600
