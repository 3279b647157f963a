"""Multi-GPU plumbing: utterances are independent units, so a batch is
partitioned BY UTTERANCE across the ranks of one box - no collective inside the sampling loop - and
only the finished mels are exchanged once at the end (NCCL over NVLink on GPUs, gloo in CPU tests).

The reference has no multi-GPU inference at all (B = 1 per segment, inference/ds_acoustic.py:209-219);
this module is new and deliberately tiny: a deterministic longest-first bin-packing and one gather.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def partition_by_length(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Greedy longest-first bin packing on total frames.  Deterministic: ties broken by utterance index,
    then by rank.  Returns, per rank, the ORIGINAL indices of its utterances in ascending order."""
    if world_size < 1:
        raise ValueError('world_size must be >= 1')
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    load = [0] * world_size
    parts: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        parts[r].append(i)
        load[r] += int(lengths[i])
    return [sorted(p) for p in parts]


def gather_mels(local: torch.Tensor, local_index: Sequence[int], total: int, group=None,
                dst: Optional[int] = 0) -> Optional[torch.Tensor]:
    """Gathers per-rank results [B_local, T, M] (same T, M on every rank; padded batches) into
    [total, T, M] ordered by original utterance index.  One collective: an all_gather on a buffer padded to
    the largest local batch (<= 85 MB per rank at the largest sweep point).  ``dst=None`` returns the full
    tensor on every rank, otherwise only on rank ``dst`` (None elsewhere)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        out = local.new_empty((total,) + tuple(local.shape[1:]))
        out[torch.as_tensor(list(local_index), device=local.device, dtype=torch.long)] = local
        return out
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = local.device
    counts = [torch.zeros(1, dtype=torch.long, device=dev) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([local.shape[0]], dtype=torch.long, device=dev), group=group)
    counts = [int(c) for c in counts]
    bmax = max(counts)
    pad = local.new_zeros((bmax,) + tuple(local.shape[1:]))
    pad[:local.shape[0]] = local
    idx = torch.full((bmax,), -1, dtype=torch.long, device=dev)
    idx[:local.shape[0]] = torch.as_tensor(list(local_index), dtype=torch.long, device=dev)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    idxs = [torch.empty_like(idx) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    dist.all_gather(idxs, idx, group=group)
    if dst is not None and rank != dst:
        return None
    out = local.new_empty((total,) + tuple(local.shape[1:]))
    for b, i, n in zip(bufs, idxs, counts):
        if n:
            out[i[:n]] = b[:n]
    return out
