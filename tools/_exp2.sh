timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
run() { local name=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --no-library-baseline --min-seconds 0.5 --breakdown gpurun_out/breakdown_exp_$name.json > gpurun_out/exp_$name.log 2>&1; echo "$name rc=$? $(tail -1 gpurun_out/exp_$name.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], round(d["roofline"]["frac"],4), round(d["roofline"]["frac_whole_step"],4), d["breakdown_ms"]["conv"])' 2>&1 | tail -1)"; }
run flat1 TOD_FLAT=1
run flat0 TOD_FLAT=0
python bench.py --config3 --images 1024 2>&1 | tail -1 | cut -c1-900
