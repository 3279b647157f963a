"""World-size-2 gloo test of the multi-GPU host logic on the CPU (no kernels involved)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from xiaoicesing_io_b200.partition import MelGather, gather_mels, partition_by_length


def test_partition_is_balanced_and_complete():
    lengths = [690, 345, 2584, 1292, 690, 345, 345, 1292, 120]
    for w in (1, 2, 4, 8):
        parts = partition_by_length(lengths, w)
        assert sorted(i for p in parts for i in p) == list(range(len(lengths)))
        loads = [sum(lengths[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(lengths)
    assert partition_by_length([], 2) == [[], []]
    with pytest.raises(ValueError):
        partition_by_length([1], 0)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, lengths, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        parts = partition_by_length(lengths, world)
        mine = parts[rank]
        # "sampling result" of utterance i is a tensor filled with i (utterances are independent units)
        local = torch.stack([torch.full((5, 3), float(i)) for i in mine]) if mine else torch.zeros((0, 5, 3))
        full = gather_mels(local, mine, len(lengths), dst=0)                     # index lists exchanged
        everywhere = gather_mels(local, mine, len(lengths), dst=None, parts=parts)
        g = MelGather(parts, (5, 3), 'cpu', dst=0)                                # the prepared gather of the hot path, used twice
        again = [g(local), g(local * 1.0)][-1]
        ok = everywhere is not None and all(float(everywhere[i, 0, 0]) == i for i in range(len(lengths)))
        if rank == 0:
            ok = ok and full is not None and all(float(full[i, 0, 0]) == i for i in range(len(lengths)))
            ok = ok and again is not None and torch.equal(again, full)
        else:
            ok = ok and full is None and again is None
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_gather_world2_gloo():
    lengths = [690, 345, 2584, 1292, 690]          # odd count -> unequal local batches
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lengths, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_gather_single_process():
    local = torch.arange(6.).reshape(3, 2, 1)
    out = gather_mels(local, [2, 0, 1], 3)
    assert torch.equal(out[2], local[0]) and torch.equal(out[0], local[1])
