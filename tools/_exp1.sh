run() { local name=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --no-library-baseline --min-seconds 0.5 $EXTRA > gpurun_out/exp_$name.log 2>&1; echo "$name rc=$? $(tail -1 gpurun_out/exp_$name.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["config"]["single_pass_latency_ms"], d["clocks"]["sm_mhz"])' 2>&1 | tail -1)"; }
EXTRA="" run base TOD_FUSE_HEAD0=1
EXTRA="" run nowait TOD_FUSE_HEAD0=1 TOD_PDL_NOWAIT=1 TOD_BENCH_SKIP_PARITY=1
EXTRA="--plans 1" run base_p1 TOD_FUSE_HEAD0=1
EXTRA="--plans 1" run nowait_p1 TOD_FUSE_HEAD0=1 TOD_PDL_NOWAIT=1 TOD_BENCH_SKIP_PARITY=1
EXTRA="--plans 3" run base_p3 TOD_FUSE_HEAD0=1
EXTRA="--plans 3" run nowait_p3 TOD_FUSE_HEAD0=1 TOD_PDL_NOWAIT=1 TOD_BENCH_SKIP_PARITY=1
