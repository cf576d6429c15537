"""Key counters of an `ncu --set full` report, one block per captured launch, labelled with the layer names given in
launch order.  usage: summarize_ncu_full.py report.ncu-rep [name1,name2,...]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
labels = sys.argv[2].split(",") if len(sys.argv) > 2 else []
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_active.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
for n_, r in enumerate(rows[2:]):
    d = dict(zip(hdr, zip(units, r)))
    print("-" * 100)
    li = n_ - (len(rows) - 2 - len(labels))      # labels name the LAST len(labels) launches of the capture
    if labels and 0 <= li < len(labels):
        print(f"  layer: {labels[li]}")
    elif labels:
        print("  layer: (warm pass launch before the selected ops)")
    for k in want:
        if k in d and d[k][1] not in ("", "no data"):
            print(f"  {k:90s} {d[k][1]:>18s} {d[k][0]}")
    try:
        cyc = float(d["sm__cycles_active.avg"][1].replace(",", ""))
        t = d.get("TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", d.get("sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg"))
        if t and t[1] not in ("", "no data"):
            tc = float(t[1].replace(",", "")) / 4.0      # the counter sums the four sub-partitions
            print(f"  => tensor pipe active {tc:.0f} of {cyc:.0f} SM-active cycles = {100 * tc / cyc:.1f} %")
    except Exception as e:
        pass
