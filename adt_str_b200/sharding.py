"""Batch partitioning across the GPUs of one box.

Segments are independent (every ``SynthDrum.__call__`` has its own draws, mix and
normalisation; every log-mel row depends on one segment), so the path shards with
**no data-path collective**: each rank plans, renders and featurises a contiguous
slice of the batch against its own replica of the one-shot bank - what
``DistributedSampler`` under HF Trainer / accelerate already does for the model
(reference ``README.md:45,53-57``, ``train.py:307``).  The only cross-rank traffic is
the host-side reduction of benchmark statistics below.
"""
from __future__ import annotations

import random
from typing import Tuple


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced ``[lo, hi)`` slice of ``n_items`` for ``rank``."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def rank_rng(seed: int, rank: int) -> random.Random:
    """Per-rank ``random.Random`` (timbre / mixup draws) - distinct streams, reproducible."""
    return random.Random(seed * 1_000_003 + rank)


def reduce_stats(units: float, elapsed_ms: float) -> Tuple[float, float]:
    """(sum of units over ranks, max of elapsed over ranks) via torch.distributed when
    initialised; identity for a single process."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return units, elapsed_ms
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    u = torch.tensor([units], dtype=torch.float64, device=dev)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(u.item()), float(t.item())


def bind_to_local_cpus(device_index: int, rank: int = 0, world: int = 1):
    """Pin this process to the CPUs that are local to its GPU (sysfs ``local_cpulist`` of the PCI device), and among
    them to this rank's share when several ranks sit on the same node - BEFORE pinned host buffers are allocated, so
    that first touch puts them on the GPU's NUMA node and the planner threads of different ranks do not compete for
    cores.  Returns the CPU set chosen (None when sysfs gives nothing, e.g. in a container without the topology)."""
    import os
    import torch
    try:
        p = torch.cuda.get_device_properties(device_index)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        text = open(f"/sys/bus/pci/devices/{bus}/local_cpulist").read().strip()
    except Exception:
        return None
    cpus = set()
    for part in text.split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        elif part.strip().isdigit():
            cpus.add(int(part))
    allowed = sorted(cpus & os.sched_getaffinity(0))
    if not allowed:
        return None
    everything = len(allowed) == len(os.sched_getaffinity(0))
    if world > 1 and everything and len(allowed) >= 2 * world:   # one node for all GPUs: an even share per rank
        lo, hi = shard_range(len(allowed), rank, world)
        allowed = allowed[lo:hi]
    try:
        os.sched_setaffinity(0, set(allowed))
    except OSError:
        return None
    return allowed
