"""
Multi-GPU plumbing for nn-classification: one process per GPU (torchrun), windows sharded in
contiguous blocks, one exchange step at the end over NCCL (NVLink/NVSwitch).

The reference has no distributed code at all (single process, GPUs hidden --
reference genomad/modules/nn_classification.py:8); what has to be preserved is its result:
``tf.math.segment_mean`` over ALL windows of a contig in FASTA order (nn_classification.py:319-320).
Windows are independent until that mean, so the only collective is on per-window probabilities
(12 B/window) or per-contig partial sums (16 B/contig):

  * gather_window_probs : all_gather of the [W_local, 3] shards -> every rank holds [W, 3] in window
                          order; the segment mean is then computed exactly as on one GPU (bitwise
                          identical outputs for any world size).  Default.
  * allreduce_partials  : each rank reduces its own shard to [n_contigs, 4] = (sum p, count) with
                          gnm_segment_sum, then one all_reduce(SUM); used when contigs are long and
                          n_contigs << W (BASELINE config 4).  Order-free: differs from the gather
                          variant by fp32 re-association only (~1e-7).

Everything here is backend-agnostic (tensors in, tensors out) so the logic is covered on CPU with
gloo at world_size 2 (tests/test_dist_gloo.py); in production the backend is NCCL.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np


@dataclass
class DistInfo:
    rank: int = 0
    world_size: int = 1
    local_rank: int = 0

    @property
    def is_main(self) -> bool:
        return self.rank == 0


def dist_info_from_env() -> DistInfo:
    return DistInfo(int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)),
                    int(os.environ.get("LOCAL_RANK", 0)))


def init_process_group_if_needed(backend: Optional[str] = None) -> DistInfo:
    """Initialise torch.distributed from the torchrun environment (no-op for a single process)."""
    info = dist_info_from_env()
    if info.world_size > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(info.local_rank)
                dist.init_process_group(backend=backend, device_id=torch.device("cuda", info.local_rank))
            else:
                dist.init_process_group(backend=backend)
    return info


def broadcast_object(obj, info: DistInfo, src: int = 0):
    """Rank `src`'s Python object on every rank (no-op for a single process).  Used for control-flow decisions that
    depend on the file system: only rank 0 looks, everybody follows."""
    if info.world_size == 1:
        return obj
    import torch.distributed as dist
    box = [obj if info.rank == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def barrier(info: DistInfo) -> None:
    if info.world_size > 1:
        import torch.distributed as dist
        dist.barrier()


def shard_bounds(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced block of [0, n_items) owned by `rank` (first n % world ranks get one extra)."""
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def local_offsets(offsets: np.ndarray, start: int, end: int) -> np.ndarray:
    """Contig window offsets [n_contigs+1] (global) -> offsets into the local shard [start, end)."""
    return (np.clip(offsets.astype(np.int64), start, end) - start).astype(np.int32)


def gather_window_probs(local_probs, n_total: int, world_size: int, group=None):
    """all_gather of contiguous shards of unequal length -> [n_total, 3] in global window order."""
    import torch
    import torch.distributed as dist
    if world_size == 1:
        return local_probs
    max_len = -(-n_total // world_size)
    buf = torch.zeros((max_len, 3), dtype=local_probs.dtype, device=local_probs.device)
    buf[: local_probs.shape[0]] = local_probs
    out = torch.empty((world_size * max_len, 3), dtype=local_probs.dtype, device=local_probs.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    parts = []
    for r in range(world_size):
        s, e = shard_bounds(n_total, world_size, r)
        parts.append(out[r * max_len: r * max_len + (e - s)])
    return torch.cat(parts, dim=0)


def allreduce_partials(partials, world_size: int, group=None):
    """[n_contigs, 4] per-rank (sum p0, sum p1, sum p2, count) -> global; then mean = sums / count."""
    import torch.distributed as dist
    if world_size > 1:
        dist.all_reduce(partials, op=dist.ReduceOp.SUM, group=group)
    return partials


def finish_mean(partials):
    cnt = partials[:, 3:4].clamp(min=1)
    return partials[:, :3] / cnt
