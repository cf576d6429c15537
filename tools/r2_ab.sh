#!/bin/bash
# Round-2 A/B on one B200 (run under gpurun): GPU tests, the bench line, then the env toggles of this round's conv changes.
set -u
mkdir -p gpurun_out
TAG=${1:-r2a}
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --breakdown gpurun_out/breakdown_$TAG.json > gpurun_out/bench_$TAG.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench_$TAG.log | cut -c1-900
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-library-baseline --min-seconds 0.5 --breakdown gpurun_out/breakdown_${TAG}_$name.json > gpurun_out/bench_${TAG}_$name.log 2>&1
  echo "$name rc=$? $(tail -1 gpurun_out/bench_${TAG}_$name.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"],4), round(d["roofline"]["frac_whole_step"],4), d["breakdown_ms"])' 2>&1 | tail -1)"
}
run old TOD_L2PROMO=256 TOD_SNAKE=0
run promo TOD_SNAKE=0
run snake TOD_L2PROMO=256
run head0 TOD_FUSE_HEAD0=1
run nopdl TOD_PDL=0
