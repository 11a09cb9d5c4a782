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


class SharedIdentity:
    """Shared global-speaker table across ranks (extension beyond the reference).

    After every pipeline step call :meth:`sync` with the step's speaker maps: this rank's centroid changes are
    exported, all ranks' records are exchanged with ONE all-gather (NCCL on GPUs; ~82 KB per rank), merged in rank
    order by every rank with the same deterministic rule (``csrc/cluster.cu``), and the maps are rewritten where a
    centre this rank created was merged into / moved to another global index.  All ranks hold bit-identical tables
    afterwards.  ``world == 1`` degenerates to a local merge (no collective)."""

    def __init__(self, clustering, group=None):
        import ctypes as C

        from . import _lib

        self._C, self._lib = C, _lib
        self.clustering, self.group = clustering, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def sync_submitted(self, pipeline_handle) -> None:
        """The same exchange for steps submitted with ``dg_pipeline_submit*`` (call after every submit, before that step's
        collect): export, all-gather and merge are stream-ordered behind the clustering of the submitted steps and ahead of
        the next one -- the one-step-at-a-time protocol -- while the networks of the following steps keep running.  No host
        synchronisation; the maps are relabelled in their device slots."""
        lib, clu = self._lib.lib(), self.clustering
        device = clu.device
        n = lib.dg_cluster_record_len(clu._h)
        if getattr(self, "_rec", None) is None or self._rec.numel() != n:
            self._rec = torch.empty(n, dtype=torch.float64, device=device)
            self._gathered = torch.empty(self.world * n, dtype=torch.float64, device=device) if self.world > 1 else self._rec
        # the exchange runs on its OWN stream: on the caller's stream the all-gather of step i (which waits for the clustering
        # of step i) would sit in front of the next submit's start event and de-pipeline the networks
        if getattr(self, "_comm", None) is None:
            self._comm = torch.cuda.Stream(device)
        with torch.cuda.device(device), torch.cuda.stream(self._comm):
            stream = self._C.c_void_p(self._comm.cuda_stream)
            self._lib.check(lib.dg_pipeline_identity_export(pipeline_handle, self._rec.data_ptr(), stream))
            if self.world > 1:
                dist.all_gather_into_tensor(self._gathered, self._rec, group=self.group)
            self._lib.check(lib.dg_pipeline_identity_merge(pipeline_handle, self._gathered.data_ptr(), self.world, self.rank,
                                                           stream))

    def sync(self, maps: torch.Tensor) -> torch.Tensor:
        lib, clu = self._lib.lib(), self.clustering
        if clu._h is None:
            raise self._lib.DiartB200Error("SharedIdentity.sync before the first clustering step")
        device = clu.device
        n = lib.dg_cluster_record_len(clu._h)
        rec = torch.empty(n, dtype=torch.float64, device=device)
        stream = self._lib.stream_ptr(device)
        with torch.cuda.device(device):
            self._lib.check(lib.dg_cluster_export_delta(clu._h, rec.data_ptr(), stream))
            if self.world > 1:
                gathered = torch.empty(self.world * n, dtype=torch.float64, device=device)
                dist.all_gather_into_tensor(gathered, rec, group=self.group)
            else:
                gathered = rec
            assert maps.dtype == torch.int32 and maps.is_contiguous() and maps.device == device
            self._lib.check(lib.dg_cluster_merge(clu._h, gathered.data_ptr(), self.world, self.rank, maps.data_ptr(),
                                                 maps.numel(), stream))
        return maps
