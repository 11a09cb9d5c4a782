"""Multi-GPU plumbing: one process per GPU, independent audio streams sharded across ranks.

The reference's only parallel mode is a process pool with one pipeline per file
(``src/diart/inference.py:435-559``): streams share nothing (centroids live in the pipeline instance,
``blocks/diarization.py:146-155``), so the data path needs **no collective**; ``torch.distributed`` is used
only for the barrier and for reducing timings / counters (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist


def env_rank() -> Dict[str, int]:
    return {"rank": int(os.environ.get("RANK", "0")), "world": int(os.environ.get("WORLD_SIZE", "1")),
            "local": int(os.environ.get("LOCAL_RANK", "0"))}


def init(backend: Optional[str] = None, device: Optional[torch.device] = None) -> Dict[str, int]:
    """Initialises the default process group from the torchrun environment (no-op for world size 1)."""
    info = env_rank()
    if info["world"] > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kwargs = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kwargs)
    return info


def shard_streams(num_streams: int, world: int, rank: int) -> List[int]:
    """Stream s runs on rank ``s mod world``."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("invalid rank / world size")
    return [s for s in range(num_streams) if s % world == rank]


def barrier(device: Optional[torch.device] = None):
    if device is not None and device.type == "cuda":
        torch.cuda.synchronize(device)
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    """Multi-GPU timings are reported as the maximum over ranks."""
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def throughput(chunks_this_rank: int, seconds_this_rank: float, step_seconds: float = 0.5,
               device: Optional[torch.device] = None) -> float:
    """Whole-job stream-seconds per second: all ranks' chunks over the slowest rank's time."""
    total = sum_over_ranks(float(chunks_this_rank), device)
    return total * step_seconds / max_over_ranks(seconds_this_rank, device)


def gather_maps(local_maps: Sequence[torch.Tensor]) -> List[List[torch.Tensor]]:
    """Collects every rank's per-stream speaker maps on all ranks (result bookkeeping only; sizes may differ)."""
    if not dist.is_initialized():
        return [list(local_maps)]
    out: List[Optional[List[torch.Tensor]]] = [None] * dist.get_world_size()
    dist.all_gather_object(out, [m.cpu() for m in local_maps])
    return out  # type: ignore[return-value]
