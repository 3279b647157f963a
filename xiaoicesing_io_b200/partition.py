"""Multi-GPU plumbing: utterances are independent units, so a batch is
partitioned BY UTTERANCE across the ranks of one box - no collective inside the sampling loop - and
only the finished mels are exchanged once at the end: ONE gather to one rank (NCCL over NVLink on GPUs, gloo in CPU tests).

The reference has no multi-GPU inference at all (B = 1 per segment, inference/ds_acoustic.py:209-219);
this module is new and deliberately tiny: a deterministic longest-first bin-packing and one gather into pre-allocated
buffers (no host synchronisation on the hot path).
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


class MelGather:
    """The ONE exchange of the multi-GPU path: the finished mels of every rank -> rank ``dst`` (or every rank), ordered by
    original utterance index.  Everything that does not depend on the data is prepared ONCE here - every rank derives every
    other rank's utterance list from the same deterministic partition, so no counts or indices travel and nothing on the hot
    path synchronises with the host: ``__call__`` is one ``gather`` (NCCL over NVLink on GPUs, gloo in CPU tests) into
    pre-allocated buffers plus one index copy per rank on the destination.

    ``parts``: per-rank lists of original utterance indices (``partition_by_length`` or any fixed assignment);
    ``shape``: the per-utterance shape, e.g. (T, M)."""

    def __init__(self, parts: Sequence[Sequence[int]], shape: Sequence[int], device, dtype=torch.float32, group=None,
                 dst: Optional[int] = 0):
        self.group, self.dst = group, dst
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if len(parts) != self.world:
            raise ValueError(f'partition has {len(parts)} parts for world size {self.world}')
        self.parts = [list(p) for p in parts]
        self.total = sum(len(p) for p in self.parts)
        self.bmax = max((len(p) for p in self.parts), default=0)
        self.shape = tuple(shape)
        self.n_local = len(self.parts[self.rank])
        self.pad = torch.zeros((self.bmax,) + self.shape, device=device, dtype=dtype)
        receives = dst is None or self.rank == dst
        self.out = torch.empty((self.total,) + self.shape, device=device, dtype=dtype) if receives else None
        self.bufs = [torch.empty_like(self.pad) for _ in range(self.world)] if (receives and self.world > 1) else None
        self.idx = [torch.as_tensor(p, device=device, dtype=torch.long) for p in self.parts] if receives else None

    def __call__(self, local: torch.Tensor) -> Optional[torch.Tensor]:
        if tuple(local.shape) != (self.n_local,) + self.shape:
            raise ValueError(f'local result has shape {tuple(local.shape)}, expected {(self.n_local,) + self.shape}')
        if self.world == 1:
            self.out.index_copy_(0, self.idx[0], local)
            return self.out
        self.pad[:self.n_local].copy_(local)
        if self.dst is None:
            dist.all_gather(self.bufs, self.pad, group=self.group)
        else:
            dist.gather(self.pad, gather_list=self.bufs if self.rank == self.dst else None, dst=self.dst, group=self.group)
            if self.rank != self.dst:
                return None
        for b, i in zip(self.bufs, self.idx):
            if i.numel():
                self.out.index_copy_(0, i, b[:i.numel()])
        return self.out


def gather_mels(local: torch.Tensor, local_index: Sequence[int], total: int, group=None,
                dst: Optional[int] = 0, parts: Optional[Sequence[Sequence[int]]] = None) -> Optional[torch.Tensor]:
    """One-shot convenience wrapper around ``MelGather`` ([B_local, T, M] per rank -> [total, T, M] by original index).  With
    ``parts`` (every rank's index list, known from the partition) nothing but the mels is exchanged; without it the index lists
    are exchanged first (one ``all_gather_object``: a host round trip - pass ``parts``, or keep a ``MelGather``, on a hot path)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        out = local.new_empty((total,) + tuple(local.shape[1:]))
        out[torch.as_tensor(list(local_index), device=local.device, dtype=torch.long)] = local
        return out
    if parts is None:
        parts = [None] * dist.get_world_size(group)
        dist.all_gather_object(parts, list(local_index), group=group)
    g = MelGather(parts, tuple(local.shape[1:]), local.device, local.dtype, group=group, dst=dst)
    if g.total != total:
        raise ValueError(f'partition covers {g.total} utterances, expected {total}')
    return g(local)
