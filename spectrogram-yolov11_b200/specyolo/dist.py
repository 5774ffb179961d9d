"""One-process-per-GPU plumbing for the sharded inference path: images / bursts are independent
units, so a global batch is split contiguously across ranks, every rank runs its own replica of the detector
on its shard, and nothing crosses NVLink on the data path.  torch.distributed (NCCL on GPUs, gloo in the CPU
tests) is used only for the barrier around a timed region, the max-over-ranks of the elapsed time and the
gather of per-rank detection counts.
"""
from __future__ import annotations

import os
from typing import List, Tuple

import torch
import torch.distributed as dist


def env_rank() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) as set by torch.distributed.run; (0, 1, 0) when run directly."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of `n_items` owned by `rank`; the first n_items % world ranks get one extra item."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def init(backend: str | None = None, device: torch.device | None = None) -> Tuple[int, int]:
    """Initialise the default process group from the environment when WORLD_SIZE > 1. Returns (rank, world)."""
    rank, world, _ = env_rank()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return rank, world


def barrier() -> None:
    if dist.is_initialized():
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def max_over_ranks(value: float, device: torch.device | str = "cpu") -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device: torch.device | str = "cpu") -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_counts(counts: torch.Tensor) -> List[torch.Tensor]:
    """Per-image detection counts of every rank, in global image order (rank 0's shard first)."""
    if not dist.is_initialized():
        return [counts]
    world = dist.get_world_size()
    sizes = [torch.zeros(1, dtype=torch.int64, device=counts.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([counts.numel()], dtype=torch.int64, device=counts.device))
    mx = int(max(int(s) for s in sizes))
    pad = torch.zeros(mx, dtype=counts.dtype, device=counts.device)
    pad[: counts.numel()] = counts
    outs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return [o[: int(s)] for o, s in zip(outs, sizes)]


def gather_detections(dets: List[torch.Tensor], device: torch.device | str = "cpu") -> List[torch.Tensor]:
    """Per-image [n_i, 6] detection tensors of this rank's shard -> the detections of the WHOLE global batch in input
    order on every rank.  Two collectives on results only (counts, then rows padded to the largest per-rank total); the
    data path itself has none."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return list(dets)
    world = dist.get_world_size()
    counts = torch.tensor([int(d.shape[0]) for d in dets], dtype=torch.int64, device=device)
    per_rank = gather_counts(counts)
    rows = torch.cat([d.to(device=device, dtype=torch.float32).reshape(-1, 6) for d in dets]) if dets else \
        torch.zeros((0, 6), dtype=torch.float32, device=device)
    mx = max(int(c.sum()) for c in per_rank)
    pad = torch.zeros((max(mx, 1), 6), dtype=torch.float32, device=device)
    pad[: rows.shape[0]] = rows
    outs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    result: List[torch.Tensor] = []
    for o, c in zip(outs, per_rank):
        off = 0
        for n in c.tolist():
            result.append(o[off: off + n].cpu())
            off += n
    return result


def predict_sharded(predict_fn, batch: torch.Tensor, device: torch.device | str | None = None) -> List[torch.Tensor]:
    """One GLOBAL batch (images [B,3,H,W] or IQ bursts [B,L], the same tensor on every rank, e.g. pinned host memory) ->
    per-image detections [n_i, 6] of all B units in input order, on every rank (contiguous shards, one
    replica per GPU, results concatenated on the host in input order).

    `predict_fn(shard) -> List[Results | Tensor[n,6]]` is the per-rank replica: `yolo.predict` for images,
    `yolo.predict_iq` for bursts.  Ranks whose shard is empty (B < world) run nothing."""
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    lo, hi = shard_range(int(batch.shape[0]), rank, world)
    dets: List[torch.Tensor] = []
    if hi > lo:
        for r in predict_fn(batch[lo:hi]):
            d = r.boxes.data if hasattr(r, "boxes") else r
            dets.append(torch.as_tensor(d).reshape(-1, 6))
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if (
            dist.is_initialized() and dist.get_backend() == "nccl") else "cpu"
    return gather_detections(dets, device)


def bind_to_gpu_numa(device_index: int) -> dict:
    """Pin this process (and, through first-touch, the pinned host buffers it allocates afterwards) to the NUMA node
    the GPU hangs off.  At 8 GPUs the end-to-end path is bound by the pinned-host -> device copies (78.6 MB per 64-image
    batch per GPU, ~25 GB/s each, >200 GB/s through host memory in aggregate): a rank whose staging buffers live on the
    other socket pays the inter-socket link on every batch.  Linux only, no libnuma: the node comes from sysfs
    (/sys/bus/pci/devices/<bus id>/numa_node), the CPU list from /sys/devices/system/node/node<N>/cpulist, the binding is
    os.sched_setaffinity.  Returns what was done (for the bench line); a box without NUMA information is left alone."""
    info = {"gpu": device_index, "numa_node": None, "cpus": None, "bound": False}
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id \
            if hasattr(torch.cuda.get_device_properties(device_index), "pci_bus_id") else None
        dom = getattr(torch.cuda.get_device_properties(device_index), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(device_index), "pci_device_id", 0)
        if bus is None:
            return info
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        with open(path) as f:
            node = int(f.read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus"], info["bound"] = len(allowed), True
    except (OSError, ValueError, AttributeError, RuntimeError):
        pass
    return info
