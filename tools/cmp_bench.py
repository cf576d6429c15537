"""Compare two conv_bench JSON files: per-layer best-of and totals.  usage: cmp_bench.py old.json new.json"""
import json, sys
old = {r['op']: r for r in json.load(open(sys.argv[1]))['rows']}
new = json.load(open(sys.argv[2]))['rows']
tb = tn = 0
for r in new:
    o = old.get(r['op'], {})
    vs = [k for k in ('v1', 'halo', 'auto') if r.get(k) is not None]
    ob = min([o[k] for k in ('v1', 'halo', 'auto') if o.get(k) is not None] or [0])
    nb = min(r[k] for k in vs)
    tb += ob; tn += nb
    bad = max((r.get(k + '_bad') or 0) for k in vs)
    print(f"{r['op']:28s} {r['shape']:30s} " + " ".join(f"{k} {o.get(k, 0) or 0:6.1f}->{r[k]:6.1f}" for k in vs) +
          f"  best {ob:6.1f}->{nb:6.1f}  {r['gflop'] / nb * 1e3:5.0f}TF {r['mbytes'] / nb:4.2f}TB/s" + (f"  BAD {bad}" if bad else ""))
print(f"best-of total: {tb:.1f} -> {tn:.1f} us")
