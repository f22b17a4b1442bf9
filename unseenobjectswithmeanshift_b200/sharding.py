"""Batch sharding across one-process-per-GPU replicas.

Every image (and every crop) is independent through the whole hot path, so the
multi-GPU form is: split the batch, one replica per rank, NO data-path collective. The only
collectives are the timing barrier / max-over-ranks used by bench.py.
"""
import os

import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous [begin, end) of ``total`` items owned by ``rank``; sizes differ by at most 1."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def init_from_env(backend=None):
    """(rank, local_rank, world). Initialises torch.distributed when launched by torchrun."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"))
    return rank, local_rank, world


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value, device="cpu"):
    """max of a python float over all ranks (identity when not distributed)."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def pin_to_gpu_numa(local_rank):
    """Restrict this process to the CPUs of the NUMA node its GPU hangs off (sysfs: the PCI device's ``numa_node``), so
    that pinned host buffers allocated afterwards are first-touched on that node and the H2D / D2H copies of the ranks
    of one box do not all cross the same socket (round 1: 8 ranks on NUMA 0, end-to-end scaling efficiency 0.35).
    Returns a one-line description for the bench record; never raises (returns the reason when it cannot pin)."""
    try:
        p = torch.cuda.get_device_properties(local_rank)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return f"gpu {local_rank} ({bus}): no NUMA affinity reported"
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return f"gpu {local_rank} ({bus}): NUMA node {node} has no CPU this process may use"
        os.sched_setaffinity(0, allowed)
        return f"gpu {local_rank} ({bus}) -> NUMA node {node}, {len(allowed)} CPUs"
    except Exception as e:  # informational: sysfs layout, containers, cpusets differ
        return f"not pinned: {e!r}"
