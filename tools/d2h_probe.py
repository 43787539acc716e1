#!/usr/bin/env python3
"""Concurrent device->host bandwidth of all ranks into pinned memory (what bounds the end-to-end number at N > 1),
with and without binding each rank to the CPUs local to its GPU before the pinned buffer is allocated.
    python -m torch.distributed.run --nproc-per-node N tools/d2h_probe.py
"""
import os
import time

import torch
import torch.distributed as dist


def local_cpus(dev_index):
    try:
        p = torch.cuda.get_device_properties(dev_index)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bus}"
        node = open(f"{base}/numa_node").read().strip()
        cpus = open(f"{base}/local_cpulist").read().strip()
        return bus, node, cpus
    except Exception as exc:
        return "?", "?", repr(exc)


def parse_cpulist(s):
    out = set()
    for part in s.split(","):
        if "-" in part:
            a, b = part.split("-"); out.update(range(int(a), int(b) + 1))
        elif part.strip().isdigit():
            out.add(int(part))
    return out


def measure(dev, n_bytes, reps=8):
    host = torch.empty(n_bytes // 4, dtype=torch.float32).pin_memory()
    host.zero_()
    src = torch.empty(n_bytes // 4, dtype=torch.float32, device=dev)
    host.copy_(src, non_blocking=True); torch.cuda.synchronize(dev)
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        host.copy_(src, non_blocking=True)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    dist.barrier()
    return n_bytes * reps / dt / 1e9


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("gloo")
    bus, node, cpus = local_cpus(local)
    allowed = os.sched_getaffinity(0)
    if rank == 0:
        os.system("nvidia-smi topo -m 2>&1 | head -20; ls /sys/devices/system/node | head; nproc")
    before = measure(dev, 1 << 30)
    want = parse_cpulist(cpus) & allowed
    bound = False
    if want:
        os.sched_setaffinity(0, want)
        bound = True
    after = measure(dev, 1 << 30)
    res = [None] * world
    dist.all_gather_object(res, (rank, bus, node, cpus, len(allowed), bound, round(before, 1), round(after, 1)))
    if rank == 0:
        for r in res:
            print(r)
        print("aggregate GB/s before", round(sum(r[6] for r in res), 1), "after binding", round(sum(r[7] for r in res), 1))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
