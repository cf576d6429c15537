#!/usr/bin/env python
"""Benchmark of the detector inference hot path: images/sec @640x640 (forward + decode + NMS).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scale s] [--batch 64] [--size 640]

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE); every rank runs the same batch-sharded work
(weak scaling, no data-path collective); rank 0 prints ONE JSON line.

  value        whole-job images/s with the uint8 image batch already resident in HBM (CUDA-graph replay of
               network + decode + NMS), device-timed, max over ranks
  e2e          same metric through the public API Detector.submit / Detector.collect (two batches in flight): pinned
               host uint8 (B, H, W, 3) batch -> H2D -> graph -> D2H of counts and kept rows, every step;
               e2e.f32_input is the synchronous Detector.detect on the reference's float32 (B, 3, H, W) tensor
  roofline     the dominant kernel (the tcgen05 conv kernels, all conv launches of one pass) against the measured
               bf16 tensor peak: algorithmic conv FLOPs / summed conv launch time (CUDA events, eager pass);
               traffic = measured DRAM bytes per conv launch from the committed ncu pass (profiles/conv_dram_traffic.json)
  cpu_baseline the CPU oracle (port of the reference path) on this box's host cores, bounded sample
  --impl reference   times the CPU oracle alone (the reference itself cannot travel to the GPU box)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec @640^2 (fwd+decode+NMS)"
CONF, IOU = 0.05, 0.5          # reference EvalCallback defaults (utils/callbacks.py:102-104)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def bind_to_gpu_numa_node(local_rank: int):
    """Best effort: run this rank (and first-touch its pinned upload buffers) on the NUMA node its GPU hangs off.  With 8
    ranks each re-reading its host batches at ~30 GB/s, uploads that cross the socket interconnect are what bounds the
    end-to-end number.  Returns the node or None."""
    try:
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id if hasattr(torch.cuda.get_device_properties(local_rank), "pci_bus_id") else None
        if bus is None:
            import pynvml as nv
            nv.nvmlInit()
            bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(local_rank)).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:          # nvml prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def cpu_oracle_rate(scale: str, size: int, images: int, min_seconds: float, max_runs: int):
    """images/s of the CPU oracle (forward fp32 + decode_box + NMS) on all host cores."""
    from oracle import detector_oracle as O, synth
    C_, d, m = synth.SCALES[scale]
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in synth.make_state_dict(80, C_, d, m, seed=0).items()}
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    x = torch.from_numpy(synth.images_u8_to_f32(synth.make_images_u8(images, size, size, seed=3)))
    O.detect(sd, x[:1], 80, d, (size, size), True, CONF, IOU)          # warm-up
    times = []
    t_all = time.perf_counter()
    while len(times) < max_runs and (time.perf_counter() - t_all < min_seconds or not times):
        t0 = time.perf_counter()
        O.detect(sd, x, 80, d, (size, size), True, CONF, IOU)
        times.append(time.perf_counter() - t0)
    return images / float(np.median(times)), cores, len(times)


def run_reference(args, rank, world):
    if rank != 0:
        return
    images = args.ref_images
    times = []
    from oracle import detector_oracle as O, synth
    C_, d, m = synth.SCALES[args.scale]
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in synth.make_state_dict(80, C_, d, m, seed=0).items()}
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    x = torch.from_numpy(synth.images_u8_to_f32(synth.make_images_u8(images, args.size, args.size, seed=3)))
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        O.detect(sd, x, 80, d, (args.size, args.size), True, CONF, IOU)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    v = images / (ms / 1e3)
    line = {"impl": "reference", "metric": METRIC.replace("640", str(args.size)), "value": v, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"scale {args.scale} detector, {args.size}x{args.size}, nc 80, conf {CONF} iou {IOU}; "
                                   f"each step = {images} images (bounded sample of the batch-{args.batch} workload)"},
            "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": f"{images} images per step, {args.steps} steps, CPU oracle (fp32 torch + numpy NMS)"},
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", default="s")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--ref-images", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--attention", action="store_true",
                    help="current-source topology: CBAM x 14 + SelfAttention (SURVEY 8 row f1; not the BASELINE config)")
    ap.add_argument("--plans", type=int, default=2, help="independent plans / streams alternating in the device-resident leg")
    ap.add_argument("--depth", type=int, default=4, help="batches in flight in the end-to-end leg (Detector.pipeline_depth)")
    ap.add_argument("--breakdown", default="", help="write the per-op eager timing table to this JSON file")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch.distributed as dist
    from transparent_object_detection_b200 import synth
    from transparent_object_detection_b200 import BaseModel, Detector

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    C_, d, m = synth.SCALES[args.scale]
    model = BaseModel(80, C_, d, m, attention=args.attention).eval()
    sd_np = synth.make_state_dict(80, C_, d, m, seed=0)
    if args.attention:
        sd_np.update(synth.make_attention_state_dict(80, C_, d, m, seed=0))
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()})
    det = Detector(model, (args.size, args.size), confidence=CONF, nms_iou=IOU, letterbox_image=True, pipeline_depth=args.depth)
    B = args.batch
    # rank-distinct synthetic uint8 batches (seed 3 + 17*rank + j), two pinned host copies for the e2e leg
    hosts = [torch.from_numpy(synth.make_images_u8(B, args.size, args.size, seed=3 + 17 * rank + j)).pin_memory() for j in range(2)]
    # two independent plans (own activation arena + graph) on two streams, replayed alternately: consecutive batches
    # overlap the NMS tail of one with the stem / first layers of the next -- the same pipelining Detector.submit uses
    NP = max(1, args.plans)
    engs = [model.engine(B, args.size, args.size, dev, instance=i) for i in range(NP)]
    eng = engs[0]
    graphs = [e.graph_for("u8", 0, CONF, IOU) for e in engs]
    for j, e in enumerate(engs):
        e.input_buffer("u8", 0).copy_(hosts[j & 1])
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(dev) for _ in range(NP)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def replay_many(n):
        main = torch.cuda.current_stream(dev)
        for st_ in streams:
            st_.wait_stream(main)
        for k in range(n):
            with torch.cuda.stream(streams[k % NP]):
                graphs[k % NP].replay()
        for st_ in streams:
            main.wait_stream(st_)

    sampler = ClockSampler(local_rank)
    # ---------------------------------------------------------------- device-resident throughput
    replay_many(args.warmup)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    replay_many(args.steps)
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    # one batch alone (no overlap with a neighbour): latency of a pass
    graphs[0].replay()
    torch.cuda.synchronize()
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record()
    graphs[0].replay()
    l1.record()
    torch.cuda.synchronize()
    pass_latency_ms = l0.elapsed_time(l1)
    # ---------------------------------------------------------------- end to end (host buffers, public API, pipeline_depth in flight)
    for j in range(max(args.warmup, det.pipeline_depth)):     # every plan instance captures its graph on first use
        det.detect(hosts[j & 1])
    barrier()
    d2h = 0
    t0 = time.perf_counter()
    inflight = []
    for i in range(args.steps):
        inflight.append(det.submit(hosts[i & 1], (args.size, args.size)))   # shapes at submit: rows un-letterboxed on the device
        if len(inflight) == det.pipeline_depth:
            pend = inflight.pop(0)
            det.collect(pend)
            d2h = pend.d2h_bytes
    for pend in inflight:
        det.collect(pend)
        d2h = pend.d2h_bytes
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3      # host clock: the region starts and ends on the host by definition
    barrier()
    # the reference's float32 tensor through the synchronous call (same pixels)
    hosts_f32 = torch.from_numpy(synth.images_u8_to_f32(hosts[0].numpy())).pin_memory()
    for _ in range(det.pipeline_depth):          # each plan instance captures its float32-input graph on first use
        det.detect(hosts_f32)
    barrier()
    t0 = time.perf_counter()
    nf = max(3, args.steps // 4)
    for _ in range(nf):
        det.detect(hosts_f32)
    torch.cuda.synchronize()
    e2e_f32_ms = (time.perf_counter() - t0) * 1e3 / nf
    sampler.stop_flag = True
    t = torch.tensor([dev_ms, e2e_ms, e2e_f32_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, e2e_f32_ms = float(t[0]), float(t[1]), float(t[2])

    # ---------------------------------------------------------------- per-op eager pass (kernel-level roofline)
    table = []
    if rank == 0:
        x_u8 = eng.input_buffer("u8", 0)
        eng.run_network(x_u8, fused_decode=eng.fuse_head_decode)
        if not eng.fuse_head_decode:
            eng.run_decode(False, False, True)
        eng.run_nms(CONF, IOU)
        torch.cuda.synchronize()
        import ctypes as C
        from transparent_object_detection_b200._lib import check
        st = torch.cuda.current_stream().cuda_stream
        reps = 3
        acc = {}
        for rep in range(reps):
            evs = []
            for kind, name, payload in eng.ops:
                if kind == "conv" and (name in eng.tail_skip or (eng.fuse_head_decode and name in eng.tail_box_skip)):
                    continue                                                # fused into the previous conv's epilogue
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                if kind == "conv":
                    if eng.fuse_head_decode and name in eng.tail_box:
                        tb, fb = eng.tail_box[name]
                        check(eng.L.tod_conv2d_tail1x1_box_decode(C.byref(payload), C.byref(tb), C.byref(fb), st), name)
                    elif name in eng.tail_fuse:
                        check(eng.L.tod_conv2d_tail1x1(C.byref(payload), C.byref(eng.tail_fuse[name][0]), st), name)
                    elif eng.fuse_head_decode and name in eng.head_fuse:      # what the captured graph runs
                        check(eng.L.tod_conv2d_head_decode(C.byref(payload), C.byref(eng.head_fuse[name]), st), name)
                    else:
                        check(eng.L.tod_conv2d_nhwc_bf16(C.byref(payload), st), name)
                elif kind == "stem":
                    w, bb, out = payload
                    check(eng.L.tod_stem_conv_nhwc_u8(x_u8.data_ptr(), w.data_ptr(), bb.data_ptr(), out.ptr, B, args.size,
                                                      args.size, C_, out.pitch, st), name)
                elif kind == "cbam":
                    check(eng.L.tod_cbam_nhwc_bf16(C.byref(payload), st), name)
                elif kind == "attn":
                    eng._run_attention(payload, st)
                else:
                    buf, c_ = payload
                    check(eng.L.tod_sppf_pool_nhwc_bf16(buf.ptr, B, buf.h, buf.w, c_, buf.pitch, st), name)
                b.record()
                evs.append((kind, name, payload, a, b))
            if not eng.fuse_head_decode:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); eng.run_decode(False, False, True); b.record()
                evs.append(("decode", "head_decode", None, a, b))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); eng.run_nms(CONF, IOU); b.record()
            evs.append(("nms", "nms(sort+segments+compact)", None, a, b))
            torch.cuda.synchronize()
            for kind, name, payload, a, b in evs:
                acc.setdefault(name, [kind, payload, []])[2].append(a.elapsed_time(b))
        for name, (kind, payload, ts) in acc.items():
            row = {"op": name, "kind": kind, "ms": float(np.median(ts))}
            if kind == "conv":
                dd = payload
                ho, wo = dd.hin // dd.stride, dd.win // dd.stride
                row["gflop"] = 2.0 * B * ho * wo * dd.cout * dd.cin * dd.ksize ** 2 / 1e9
                row["shape"] = f"{dd.cin}->{dd.cout} k{dd.ksize} s{dd.stride} @{ho}x{wo}"
                if name in eng.tail_fuse or (eng.fuse_head_decode and name in eng.tail_box):   # + the fused 1x1 conv's FLOPs
                    row["gflop"] += 2.0 * B * ho * wo * 64 * 64 / 1e9
                    row["shape"] += " + 64->64 k1 (fused tail)"
                row["tflops"] = row["gflop"] / row["ms"]
            table.append(row)

    if rank == 0:
        pk = peaks()
        total_images = B * world * args.steps
        value = total_images / (dev_ms / 1e3)
        e2e = total_images / (e2e_ms / 1e3)
        conv_rows = [r for r in table if r["kind"] == "conv"]
        conv_ms = sum(r["ms"] for r in conv_rows)
        conv_gflop = sum(r["gflop"] for r in conv_rows)
        all_ms = sum(r["ms"] for r in table)
        achieved = conv_gflop / conv_ms if conv_ms > 0 else 0.0          # TFLOP/s (GFLOP / ms)
        peak = pk["bf16_tflops_sustained"]
        traffic = None                                     # DRAM bytes per conv launch from the committed ncu pass
        tpath = os.path.join(ROOT, "profiles", "conv_dram_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = float(json.load(open(tpath))["avg_bytes_per_launch"])
            except Exception:
                traffic = None
        roof = {"bound": "tensor", "kernel": "conv_halo_tcgen05 / conv_igemm_tcgen05 (all conv launches of one pass)", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": pk["source"] + " (sustained cuBLAS bf16)",
                "launches": len(conv_rows), "flop_per_launch_avg": conv_gflop * 1e9 / max(len(conv_rows), 1),
                "avg_launch_ms": conv_ms / max(len(conv_rows), 1), "conv_share_of_eager_pass": conv_ms / all_ms if all_ms else None}
        cpu = None
        if not args.no_cpu_baseline and world == 1:     # N = 1 only: the other ranks' barrier spin would share the cores
            v, cores, runs = cpu_oracle_rate(args.scale, args.size, 8, 10.0, 6)
            cpu = {"value": v, "unit": "images/s", "cores": cores, "kind": "port",
                   "sample": f"8 images per run x {runs} runs of the CPU oracle (fp32 torch forward + decode + numpy NMS), median"}
        line = {"metric": METRIC.replace("640", str(args.size)), "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"scale {args.scale} detector (BaseModel(80,{C_},{d},{m})), batch {B} per GPU, {args.size}x{args.size}, "
                                       f"nc 80, conf {CONF} iou {IOU}, random-init weights, uint8 NHWC images (/255 fused into the stem)"
                                       + (", CURRENT-SOURCE topology (CBAM x 14 + SelfAttention)" if args.attention else ""),
                           "timing": f"CUDA events around K graph replays, {NP} independent plans alternating on {NP} streams (batch i's NMS tail "
                                     "overlaps batch i+1's first layers); activations per pass (~2.5 GB) exceed L2 (126 MB)",
                           "single_pass_latency_ms": pass_latency_ms},
                "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": int(hosts[0].numel()), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": e2e_ms / args.steps, "api": f"Detector.submit/collect, {det.pipeline_depth} batches in flight, pinned uint8 host batch"
                               + (f" (rank bound to its GPU's NUMA node: {numa_node})" if numa_node is not None else ""),
                        "f32_input": {"value": B * world / (e2e_f32_ms / 1e3), "ms_per_step": e2e_f32_ms,
                                      "h2d_bytes_per_step": int(hosts_f32.numel() * 4),
                                      "api": "Detector.detect (synchronous) on the reference's float32 (B,3,H,W) tensor"}},
                "gpu_launches": eng.launches_per_pass * args.steps,
                "clocks": sampler.result(), "roofline": roof, "cpu_baseline": cpu,
                "breakdown_ms": {"conv": conv_ms, "stem": sum(r["ms"] for r in table if r["kind"] == "stem"),
                                 "pool": sum(r["ms"] for r in table if r["kind"] == "pool"),
                                 "cbam": sum(r["ms"] for r in table if r["kind"] == "cbam"),
                                 "attn": sum(r["ms"] for r in table if r["kind"] == "attn"),
                                 "decode": sum(r["ms"] for r in table if r["kind"] == "decode"),
                                 "nms": sum(r["ms"] for r in table if r["kind"] == "nms")},
                "conv_tflops_whole_pass": eng.conv_flops / 1e12 / (dev_ms / args.steps / 1e3)}
        if args.breakdown:
            os.makedirs(os.path.dirname(os.path.abspath(args.breakdown)), exist_ok=True)
            json.dump({"line": line, "ops": table}, open(args.breakdown, "w"), indent=1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
