"""Timeline of the captured detector graph (tod_debug_set_timeline): one record per CTA of every instrumented kernel --
launch id, SM, start / end in %globaltimer ns -- over a few graph replays.  Answers what ncu (which serialises kernels)
cannot: how busy the SMs are INSIDE the graph, how much consecutive kernels overlap, and each kernel's share of the step.

usage: timeline.py [--batch 64] [--size 640] [--scale s] [--plans 2] [--replays 6] [--out gpurun_out/timeline.json]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transparent_object_detection_b200 import BaseModel, synth          # noqa: E402
from transparent_object_detection_b200._lib import check                # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--scale", default="s")
    ap.add_argument("--plans", type=int, default=2)
    ap.add_argument("--replays", type=int, default=8)
    ap.add_argument("--out", default="gpurun_out/timeline.json")
    a = ap.parse_args()
    C_, d, m = synth.SCALES[a.scale]
    model = BaseModel(80, C_, d, m).eval()
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.make_state_dict(80, C_, d, m, seed=0).items()})
    dev = torch.device("cuda", 0)
    engs = [model.engine(a.batch, a.size, a.size, dev, instance=i) for i in range(a.plans)]
    L = engs[0].L
    cap = 1 << 20
    buf = torch.zeros(2 + 12 * cap, dtype=torch.int64, device=dev)
    buf[1] = cap
    check(L.tod_debug_set_timeline(buf.data_ptr()), "set timeline")
    graphs = [e.graph_for("u8", 0, 0.05, 0.5) for e in engs]            # captured WITH the tags (the warm-up pass records too)
    n_launch = L.tod_debug_timeline_launches()
    names = [L.tod_debug_timeline_name(i).decode() for i in range(n_launch)]
    for j, e in enumerate(engs):
        e.input_buffer("u8", 0).copy_(torch.from_numpy(synth.make_images_u8(a.batch, a.size, a.size, seed=3 + j)).to(dev))
    streams = [torch.cuda.Stream(dev) for _ in range(a.plans)]

    def replay(n):
        main_s = torch.cuda.current_stream(dev)
        for s in streams:
            s.wait_stream(main_s)
        for k in range(n):
            with torch.cuda.stream(streams[k % a.plans]):
                graphs[k % a.plans].replay()
        for s in streams:
            main_s.wait_stream(s)

    for _ in range(30):                                                   # reach the sustained clock / power state
        replay(20)
    torch.cuda.synchronize()
    buf[0] = 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    replay(a.replays)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    n = int(buf[0])
    rec = buf[2:2 + 12 * min(n, cap)].view(-1, 12).cpu().numpy()
    check(L.tod_debug_set_timeline(None), "clear timeline")
    ids, sm, t0, t1 = (rec[:, 0] & 0xffffffff).astype(np.int64), (rec[:, 1] & 0xffff).astype(np.int64), rec[:, 2].astype(np.int64), rec[:, 3].astype(np.int64)
    cyc_end = (rec[:, 1].astype(np.uint64) >> np.uint64(16)).astype(np.int64)          # SM cycle counter (48 bits) at CTA end
    # SM clock under this load: consecutive CTA ends on the same SM, cycles / globaltimer ns
    ghz = []
    ghz_num, ghz_den = {}, {}     # per launch id: cycles / ns over the intervals that END with one of its CTAs
    for s_ in range(int(sm.max()) + 1):
        sel = np.nonzero(sm == s_)[0]
        o = sel[np.argsort(t1[sel])]
        dt, dc = np.diff(t1[o]), np.diff(cyc_end[o])
        ok = (dt > 20000) & (dc > 0)
        if ok.any():
            ghz.append(float(dc[ok].sum() / dt[ok].sum()))
        for k in np.nonzero((dt > 3000) & (dc > 0))[0]:
            i_ = int(ids[o[k + 1]])
            ghz_num[i_] = ghz_num.get(i_, 0) + int(dc[k])
            ghz_den[i_] = ghz_den.get(i_, 0) + int(dt[k])
    if ghz:
        print(f"SM clock during the replays (cycle counter / globaltimer between CTA ends, per SM): median {np.median(ghz):.3f} GHz, "
              f"min {min(ghz):.3f}, max {max(ghz):.3f}")
    t_ready, t_acc = rec[:, 4].astype(np.int64), rec[:, 5].astype(np.int64)
    t_mma, t_last = rec[:, 6].astype(np.int64), rec[:, 7].astype(np.int64)
    T0, T1 = int(t0.min()), int(t1.max())
    span = T1 - T0
    n_sm = int(sm.max()) + 1
    print(f"{a.replays} replays over {a.plans} plan(s): {ms:.3f} ms by events, {span / 1e6:.3f} ms first CTA start -> last CTA end, "
          f"{n} CTA records, {n_sm} SMs, {ms / a.replays:.3f} ms per step")
    # ---- SM occupancy: fraction of the span each SM hosts at least one CTA / the sum of CTA residencies
    busy_union, busy_sum = np.zeros(n_sm), np.zeros(n_sm)
    for s in range(n_sm):
        sel = sm == s
        iv = sorted(zip(t0[sel], t1[sel]))
        cur_a, cur_b = None, None
        for x, y in iv:
            busy_sum[s] += y - x
            if cur_b is None or x > cur_b:
                if cur_b is not None:
                    busy_union[s] += cur_b - cur_a
                cur_a, cur_b = x, y
            else:
                cur_b = max(cur_b, y)
        if cur_b is not None:
            busy_union[s] += cur_b - cur_a
    print(f"SM has >= 1 resident CTA for {100 * busy_union.mean() / span:.1f} % of the span (min {100 * busy_union.min() / span:.1f} %, "
          f"max {100 * busy_union.max() / span:.1f} %); summed CTA residency / span = {busy_sum.mean() / span:.2f} CTAs per SM on average")
    # ---- per launch id: CTAs, per-CTA time, first start -> last end, summed SM-time share
    rows = []
    total_cta_ns = float((t1 - t0).sum())
    for i in np.unique(ids):
        sel = ids == i
        dur = (t1[sel] - t0[sel]).astype(np.float64)
        per_replay = sel.sum() / a.replays * a.plans if False else sel.sum()
        rdy, acc = t_ready[sel], t_acc[sel]
        wait_us = float(np.where(rdy > 0, rdy - t0[sel], 0).mean() / 1e3)          # CTA start -> programmatic-launch wait returned
        fill_us = float(np.where((acc > 0) & (rdy > 0), acc - rdy, 0).mean() / 1e3)  # -> first complete accumulator
        # in-graph duration of one launch: the records of an id form one cluster per replay of its graph
        o_ = np.argsort(t0[sel])
        ts0, ts1 = t0[sel][o_], t1[sel][o_]
        cuts = np.nonzero(np.diff(ts0) > 400000)[0] + 1          # a new replay starts > 0.4 ms after the previous CTA start
        spans = [float(ts1[a_:b_].max() - ts0[a_:b_].min()) / 1e3 for a_, b_ in zip(np.r_[0, cuts], np.r_[cuts, len(ts0)]) if b_ > a_]
        span_us = float(np.median(spans)) if spans else 0.0
        mc, wo, wa = rec[sel, 8].astype(np.float64), rec[sel, 9].astype(np.float64), rec[sel, 10].astype(np.float64)
        has = mc > 0
        wait_op_pct = float(100 * wo[has].sum() / mc[has].sum()) if has.any() else 0.0
        wait_acc_pct = float(100 * wa[has].sum() / mc[has].sum()) if has.any() else 0.0
        mma_end, last = t_mma[sel], t_last[sel]
        steady_us = float(np.where((last > 0) & (acc > 0), last - acc, 0).mean() / 1e3)      # first -> last accumulator complete
        drain_us = float(np.where(last > 0, t1[sel] - last, 0).mean() / 1e3)                # last accumulator -> CTA end
        mma_tail_us = float(np.where(mma_end > 0, t1[sel] - mma_end, 0).mean() / 1e3)       # MMA role done -> CTA end
        rows.append({"mma_wait_operands_pct": wait_op_pct, "mma_wait_acc_pct": wait_acc_pct, "steady_us": steady_us, "drain_us": drain_us, "mma_tail_us": mma_tail_us, "span_us": span_us,
                     "ghz": (ghz_num[int(i)] / ghz_den[int(i)]) if ghz_den.get(int(i)) else 0.0,
                     "id": int(i), "name": names[i] if i < len(names) else "?", "ctas": int(sel.sum()),
                     "cta_us_mean": float(dur.mean() / 1e3), "cta_us_max": float(dur.max() / 1e3), "wait_us": wait_us, "fill_us": fill_us,
                     "sm_time_share": float(dur.sum() / total_cta_ns), "sm_us_per_step": float(dur.sum() / 1e3 / n_sm / a.replays * 1.0)})
    # one graph = one set of ids per plan; fold the plans' copies of the same layer together by name order
    print(f"{'launch':>6s} {'kernel':40s} {'CTAs':>7s} {'span us':>8s} {'us/CTA':>8s} {'max':>8s} {'pdl-wait':>9s} {'fill':>6s} {'steady':>7s} {'drain':>6s} {'wOp%':>5s} {'wAcc%':>5s} {'GHz':>5s} {'SM-us/step':>11s} {'share':>7s}")
    for r in rows:
        print(f"{r['id']:6d} {r['name']:40s} {r['ctas']:7d} {r['span_us']:8.1f} {r['cta_us_mean']:8.1f} {r['cta_us_max']:8.1f} {r['wait_us']:9.1f} {r['fill_us']:6.1f} "
              f"{r['steady_us']:7.1f} {r['drain_us']:6.1f} {r['mma_wait_operands_pct']:5.1f} {r['mma_wait_acc_pct']:5.1f} {r['ghz']:5.2f} {r['sm_us_per_step']:11.1f} {100 * r['sm_time_share']:6.2f}%")
    by_name = {}
    for r in rows:
        by_name.setdefault(r["name"], 0.0)
        by_name[r["name"]] += r["sm_us_per_step"]
    tot = sum(by_name.values())
    print(f"sum of per-SM CTA residency per step: {tot:.1f} us vs {1e3 * ms / a.replays:.1f} us per step by events "
          f"(ratio {tot / (1e3 * ms / a.replays):.2f}: > 1 means CTAs of different kernels share SMs)")
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    json.dump({"args": vars(a), "ms": ms, "span_ns": span, "launches": rows,
               "sm_busy_union_frac": (busy_union / span).tolist(), "sm_busy_sum_frac": (busy_sum / span).tolist()}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
