"""Per-launch roofline table of one batch pass from `bench.py --breakdown` (CUDA events around each launch, eager) and,
optionally, the in-graph CTA timeline (tools/timeline.py --out): markdown to stdout.
usage: make_roofline_table.py gpurun_out/breakdown_X.json [gpurun_out/timeline_X.json] > profiles/rN_roofline_table.md"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
bd = json.load(open(sys.argv[1]))
tl = json.load(open(sys.argv[2])) if len(sys.argv) > 2 else None
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
HBM, TF, TFB = pk["hbm_gbs"], pk["bf16_tflops_sustained"], pk["bf16_tflops"]
line = bd["line"]
print(f"# Roofline per launch of one batch pass ({line['config']['workload']})\n")
print(f"Source: `bench.py --breakdown` on one B200 (CUDA events around each launch of an eager pass, median of 5) -- `{os.path.basename(sys.argv[1])}`.  ")
print(f"Peaks: `MEASURED_PEAKS.json` -- HBM copy {HBM:.0f} GB/s, cuBLAS bf16 sustained {TF:.0f} TFLOP/s (burst {TFB:.0f}).  ")
print("Algorithmic bytes = input activations read once + output written once (+ residual); FLOPs = 2·B·Ho·Wo·Cout·Cin·k².  ")
print("`bound` = the roofline that gives the longer time; `frac` = bound time / measured time.  Eager launches include ~7 us of launch /"
      " event overhead each (tools/fixed_cost.py: single vs chained launches); the captured graph's step is shorter than their sum.\n")
print("| op | shape | µs | TFLOP/s | % bf16 peak | GB/s | % HBM peak | bound | frac |")
print("|---|---|---:|---:|---:|---:|---:|---|---:|")
tot_ms = tot_bound = 0.0
for r in bd["ops"]:
    us = r["ms"] * 1e3
    if r["kind"] == "conv":
        t_tc, t_hbm = r["gflop"] / TF, r["mbytes"] / HBM          # ms
        bound = "tensor" if t_tc >= t_hbm else "hbm"
        b = max(t_tc, t_hbm)
        tot_ms += r["ms"]
        tot_bound += b
        print(f"| {r['op']} | {r['shape']} | {us:.1f} | {r['tflops']:.0f} | {100 * r['tflops'] / TF:.0f} | {r['mbytes'] / r['ms']:.0f} | "
              f"{100 * r['mbytes'] / r['ms'] / HBM:.0f} | {bound} | {b / r['ms']:.2f} |")
    elif "mbytes" in r:
        print(f"| {r['op']} | {r.get('shape', r['kind'])} | {us:.1f} | | | {r['mbytes'] / r['ms']:.0f} | {100 * r['mbytes'] / r['ms'] / HBM:.0f} | hbm | "
              f"{r['mbytes'] / HBM / r['ms']:.2f} |")
    else:
        print(f"| {r['op']} | {r['kind']} | {us:.1f} | | | | | | |")
rf = line["roofline"]
print(f"\nConv launches: {tot_ms * 1e3:.0f} µs measured against {tot_bound * 1e3:.0f} µs of per-layer roofline time = {tot_bound / tot_ms:.2f}; "
      f"conv FLOPs / summed launch time = {rf['achieved']:.0f} TFLOP/s = {rf['frac']:.3f} of the sustained bf16 peak "
      f"({rf['frac_vs_burst']:.3f} of burst); conv FLOPs / captured-graph step = {rf['achieved_whole_step']:.0f} TFLOP/s = "
      f"{rf['frac_whole_step']:.3f} ({rf['frac_whole_step_vs_burst']:.3f} of burst).  Step {line['ms_per_step']:.3f} ms, "
      f"{line['value']:.0f} images/s device-resident, {line['e2e']['value']:.0f} end to end; SM clock {line['clocks']['sm_mhz']} MHz "
      f"(nvidia-smi, {line['clocks']['reasons']}).")
if tl:
    print("\n## Inside the captured graph (tools/timeline.py: one record per CTA, two plans alternating)\n")
    print("`pdl` = CTA start -> programmatic-launch wait returned; `fill` = -> first accumulator complete; `steady` = first -> last accumulator; "
          "`drain` = last accumulator -> CTA exit; `wOp` / `wAcc` = share of the MMA role's lifetime spent waiting for operands (A / B landed) / for a "
          "free accumulator (the epilogue); GHz = SM cycle counter / globaltimer over the CTA.\n")
    print("| launch | kernel | CTAs/launch | span µs (first CTA start -> last CTA end) | µs per CTA | pdl | fill | steady | drain | wOp % | wAcc % | GHz |")
    print("|---:|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    reps = tl["args"]["replays"] // tl["args"]["plans"]
    seen = set()
    for r in tl["launches"]:
        if r["name"] in ("?",):
            continue
        key = r["id"]
        if key in seen:
            continue
        seen.add(key)
        print(f"| {r['id']} | {r['name']} | {r['ctas'] // max(reps, 1)} | {r.get('span_us', 0):.1f} | {r['cta_us_mean']:.1f} | {r['wait_us']:.1f} | {r['fill_us']:.1f} | "
              f"{r.get('steady_us', 0):.1f} | {r.get('drain_us', 0):.1f} | {r.get('mma_wait_operands_pct', 0):.0f} | {r.get('mma_wait_acc_pct', 0):.0f} | {r.get('ghz', 0):.2f} |")
        if r["name"].startswith("nms segments"):
            break
