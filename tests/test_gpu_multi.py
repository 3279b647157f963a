"""Two-process test of the multi-GPU path ON HARDWARE: ``partition_by_length`` -> each rank samples ITS utterances with the
product's tensor-core path -> ``MelGather`` to rank 0 -> the gathered result equals the single-process run of the whole batch
BIT FOR BIT (utterances are independent units; no collective inside the loop).

With two or more GPUs (``gpurun --gpus 2``) the ranks use cuda:0 / cuda:1 and the exchange is NCCL; on a one-GPU box both
ranks share cuda:0 and the finished mels travel through gloo (NCCL refuses two ranks on one device) - the sampling itself runs
on the GPU either way."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

B_TOTAL, T, H, M, STEPS = 7, 300, 256, 128, 5


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model(dev, precision):
    import xiaoicesing_io_b200 as P
    from oracle import denoisers as OD
    from oracle import weights as OW
    cfg = OD.WaveNetCfg(num_layers=6, num_channels=256, dilation_cycle_length=4)
    P.hparams.clear()
    P.hparams.update(hidden_size=H, schedule_type='linear', use_shallow_diffusion=False, diff_speedup=1000 // STEPS,
                     diff_accelerator='unipc', infer=False, b2s_precision=precision)
    model = P.GaussianDiffusion(M, backbone_type='wavenet', backbone_args=dict(num_layers=6, num_channels=256, dilation_cycle_length=4),
                                spec_min=[-12.], spec_max=[0.])
    model.denoise_fn.load_state_dict(OW.make_state_dict(cfg, seed=0, sigma_w=0.01), strict=True)
    return model.to(dev).eval()


def _inputs():
    g = torch.Generator().manual_seed(77)
    cond = torch.randn((B_TOTAL, T, H), generator=g)
    noise = torch.randn((B_TOTAL, 1, M, T), generator=g)
    return cond, noise


def _sample(model, cond, noise, idx, dev):
    it = iter([noise[idx]])
    model._noise_source = lambda shape: next(it).to(dev)
    return model(cond[idx].to(dev), infer=True)


def _worker(rank, world, port, n_gpus, precision, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dev = torch.device('cuda', rank if n_gpus >= world else 0)
    torch.cuda.set_device(dev)
    nccl = n_gpus >= world
    dist.init_process_group('nccl' if nccl else 'gloo', rank=rank, world_size=world, **({'device_id': dev} if nccl else {}))
    try:
        from xiaoicesing_io_b200.partition import MelGather, partition_by_length
        cond, noise = _inputs()
        parts = partition_by_length([T] * B_TOTAL, world)          # 7 utterances on 2 ranks: 4 + 3
        mine = parts[rank]
        model = _model(dev, precision)
        local = _sample(model, cond, noise, mine, dev).contiguous()
        if nccl:
            full = MelGather(parts, (T, M), dev, dst=0)(local)
        else:
            full = MelGather(parts, (T, M), 'cpu', dst=0)(local.cpu())
        ok, err = True, 0.0
        if rank == 0:
            whole = _sample(model, cond, noise, list(range(B_TOTAL)), dev)
            ok = bool(torch.equal(full.cpu(), whole.cpu()))
            err = float((full.cpu() - whole.cpu()).abs().max())
        else:
            ok = full is None
        q.put((rank, ok, err, 'nccl' if nccl else 'gloo'))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('precision', ['fp16', 'fp32'])
def test_two_rank_partition_reproduces_single_process_bitwise(precision):
    assert torch.cuda.is_available(), 'gpu tests need a CUDA device'
    n_gpus = torch.cuda.device_count()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_gpus, precision, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    print('two-rank partition test:', res)
    assert all(ok for _, ok, _, _ in res), res
