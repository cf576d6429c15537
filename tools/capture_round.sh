#!/bin/bash
# Round capture on one B200 (run under gpurun): GPU tests, the bench lines, the ncu launch list of the bench command and
# one `--set full` capture of the dominant conv launches.  Outputs land in gpurun_out/ (copy summaries to profiles/).
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log
python bench.py --breakdown gpurun_out/breakdown_$TAG.json > gpurun_out/bench_$TAG.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench_$TAG.log | cut -c1-600
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1; echo "bench ref rc=$?"
tail -1 gpurun_out/bench_ref_$TAG.log | cut -c1-400
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -c 700 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_$TAG.log 2>&1
echo "ncu list rc=$?"
python tools/profile_ops.py --ops head.cls.0.0,backbone.dark4.1.m.0.cv1,backbone.dark2.1.cv1,backbone.dark2.1.m.0.cv1,neck.h2.cv1 > gpurun_out/plain_ops_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:^conv_" -s 64 -c 5 -f -o gpurun_out/prof_$TAG \
    python tools/profile_ops.py --ops head.cls.0.0,backbone.dark4.1.m.0.cv1,backbone.dark2.1.cv1,backbone.dark2.1.m.0.cv1,neck.h2.cv1 > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
