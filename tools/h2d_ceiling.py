"""Host-to-device upload ceiling of this box: every rank (one per GPU, torchrun) copies pinned host batches of the bench's
size (64 x 640 x 640 x 3 uint8 = 78.6 MB) to its GPU back to back, all ranks at once -- no kernels, no result copies.
If the aggregate GB/s here equals what the end-to-end bench pulls at N = 8, the e2e scaling loss is the host's memory /
PCIe fabric, not this repo's pipeline.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_ceiling.py [--mb 78.6] [--seconds 2]
"""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bytes", type=int, default=64 * 640 * 640 * 3)
    ap.add_argument("--seconds", type=float, default=2.0)
    ap.add_argument("--buffers", type=int, default=2)
    a = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    hosts = [torch.randint(0, 256, (a.bytes,), dtype=torch.uint8).pin_memory() for _ in range(a.buffers)]
    devs = [torch.empty(a.bytes, dtype=torch.uint8, device=dev) for _ in range(a.buffers)]
    st = torch.cuda.Stream(dev)

    def burst(n):
        with torch.cuda.stream(st):
            for i in range(n):
                devs[i % a.buffers].copy_(hosts[i % a.buffers], non_blocking=True)
        st.synchronize()

    burst(4)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < a.seconds:
        burst(8)
        n += 8
    dt = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    rate = torch.tensor([n * a.bytes / dt / 1e9], dtype=torch.float64, device=dev)
    rates = [torch.zeros_like(rate) for _ in range(world)]
    if world > 1:
        dist.all_gather(rates, rate)
    else:
        rates = [rate]
    if rank == 0:
        per = [float(r) for r in rates]
        print(json.dumps({"test": "pinned host -> device copies only, all ranks at once", "n_gpus": world, "bytes_per_copy": a.bytes,
                          "gb_per_s_per_rank": per, "gb_per_s_total": sum(per), "images_per_s_equivalent": sum(per) * 1e9 / (640 * 640 * 3),
                          "host_cores": len(os.sched_getaffinity(0))}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
