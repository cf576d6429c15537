#!/bin/bash
# A/B of environment switches on the bench line (run under gpurun): name=ENV... triples
set -u
mkdir -p gpurun_out
run() {
  local name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-library-baseline --min-seconds 1.0 > gpurun_out/ab_$name.log 2>&1
  echo "$name rc=$? $(tail -1 gpurun_out/ab_$name.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("value", round(d["value"]), "best", round(d["timing"]["value_best_repeat"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],4), "clk", d["clocks"]["sm_mhz"], "W", d["clocks"].get("power_w_median"), d["parity"]["graph_rows_equal_eager_api_chain"])' 2>&1 | tail -1)"
}
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}
  run $name ${envs//,/ }
done
