"""Sharding of independent streams / recordings across GPUs, and the end-of-run counter gather.

Streams share no state (each StreamingReceiver owns its ring, sums, DC mean and flags, app.js:707-746; each
decodeReceivedSignal call is pure), so the hot path has NO collective: rank r of W owns a contiguous range of
stream indices.  The only exchange is a sum of a few int64 counters (frames decoded, frame errors, CRC errors,
payload bytes, samples) after the timed region — one all_reduce over NCCL (NVLink/NVSwitch) on GPU boxes,
gloo in the CPU tests.
"""
COUNTER_KEYS = ("framesDecoded", "frameErrors", "falsePositives", "receivedCount", "crcErrors", "duplicates",
                "crcValidFrames", "payloadBytes", "samples")


def shard_range(n_items, rank, world):
    """Contiguous, balanced split: the first n_items % world ranks get one extra item."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def owner_of(item, n_items, world):
    base, extra = divmod(n_items, world)
    cut = extra * (base + 1)
    if item < cut:
        return item // (base + 1)
    return extra + (item - cut) // base if base else world - 1


def gather_counters(counters, group=None, device=None):
    """Sum a dict of integer counters over all ranks; identity when torch.distributed is not initialised."""
    import torch
    import torch.distributed as dist

    keys = sorted(counters)
    if not (dist.is_available() and dist.is_initialized()):
        return {k: int(counters[k]) for k in keys}
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor([int(counters[k]) for k in keys], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return dict(zip(keys, [int(v) for v in t.cpu()]))


def max_over_ranks(value, group=None, device=None):
    """Max of a float over ranks (device time of the slowest rank)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def bind_to_gpu_numa_node(device_index):
    """Pin the calling process to the CPUs of the NUMA node its GPU hangs off, so that page-locked host buffers allocated
    afterwards (amb_host_alloc) are node-local and the host->device copies of several ranks do not cross the socket
    interconnect.  Linux only; returns {"node": n, "cpus": count} or None when the topology cannot be read (nothing changed)."""
    import os
    import subprocess

    try:
        out = subprocess.run(["nvidia-smi", "-i", str(device_index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().splitlines()[0].strip().lower()
        # nvidia-smi prints an 8-digit PCI domain, sysfs uses 4
        dom, rest = out.split(":", 1)
        node = int(open(f"/sys/bus/pci/devices/{dom[-4:]}:{rest}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return {"node": node, "cpus": len(allowed)}
    except Exception:
        return None
