"""Scan-sharded data parallelism: scans are independent, so a stream of scans is split by scan across
ranks and nothing is exchanged on the data path.  ``torch.distributed`` (NCCL on the GPU box, gloo in the CPU tests) is
only used for the barrier and for reducing timings / counters."""
from __future__ import annotations


def shard_indices(n_items, rank, world):
    """Indices of the scans rank ``rank`` owns: i % world == rank (the order inside a shard is the stream order)."""
    assert 0 <= rank < world
    return list(range(rank, n_items, world))


def shard_batches(n_items, rank, world, batch):
    """The rank's indices cut into engine batches of at most ``batch`` scans."""
    idx = shard_indices(n_items, rank, world)
    return [idx[i:i + batch] for i in range(0, len(idx), batch)]


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def all_reduce_scalar(value, op="max", device="cpu"):
    """max / sum of a python float over the ranks (identity when not distributed)."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())


def whole_job_rate(units_this_rank, seconds_this_rank, device="cpu"):
    """Whole-job throughput: all ranks' units divided by the slowest rank's time."""
    total = all_reduce_scalar(units_this_rank, "sum", device)
    slowest = all_reduce_scalar(seconds_this_rank, "max", device)
    return total / slowest


def bind_to_gpu_numa_node(device_index):
    """Pin the calling thread (and the threads / pinned allocations it creates afterwards) to the CPUs next to GPU
    ``device_index``: with one process per GPU the pinned staging buffers then sit in the memory of the socket the GPU
    hangs off, and the H2D / D2H DMA of 8 ranks does not funnel through one memory controller.  Returns the number of
    CPUs the thread may run on afterwards, or None if NVML cannot tell (the affinity is left alone)."""
    import os
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(device_index).uuid)
        handle = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        before = os.sched_getaffinity(0)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        after = os.sched_getaffinity(0)
        if not after or not (after & before):          # NVML named CPUs outside our cpuset: keep what we had
            os.sched_setaffinity(0, before)
            return None
        return len(after)
    except Exception:
        return None
