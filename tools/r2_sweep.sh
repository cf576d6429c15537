#!/bin/bash
# sweep (plans, SM budget) on the bench line (run under gpurun): args = "plans:budget" pairs
set -u
mkdir -p gpurun_out
for spec in "$@"; do
  P=${spec%%:*}; Bud=${spec#*:}
  TOD_SM_BUDGET=$Bud timeout 300 python bench.py --no-cpu-baseline --no-library-baseline --plans $P --depth ${DEPTH:-4} > gpurun_out/sweep_p${P}_b${Bud}.log 2>&1
  echo "plans $P budget $Bud: $(tail -1 gpurun_out/sweep_p${P}_b${Bud}.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'best', round(d['timing']['value_best_repeat']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],4), 'clk', d['clocks']['sm_mhz'], 'W', d['clocks'].get('power_w_median'), d['parity']['graph_rows_equal_eager_api_chain'])" 2>&1 | tail -1)"
done
