"""The whole inference chain on the GPU (SURVEY section 8 rows f-3 + f-2 + the path + f-1): ``.ds`` segments -> preprocess_input ->
FastSpeech2 encoder -> ConvNeXt aux decoder -> shallow DDIM sampling -> NSF-HiFiGAN -> one waveform, through
``xiaoicesing_io_b200.infer.DiffSingerAcousticInfer``.  The batched route (ragged batches of segments, per-segment seeded noise) must give
every segment the bits of the reference's one-segment-per-call loop (inference/ds_acoustic.py:209-219) run on the same kernels."""
import numpy as np
import pytest
import torch

from oracle.make_golden import _ds_param

pytestmark = pytest.mark.gpu


def _build(dev):
    import xiaoicesing_io_b200 as P
    from oracle import vocoder as OV
    smin, smax = [-12.] * 128, [0.] * 128
    P.hparams.clear()
    P.hparams.update(hidden_size=256, enc_layers=2, enc_ffn_kernel_size=3, ffn_act='gelu', num_heads=2, use_pos_embed=True, rel_pos=True,
                     use_rope=True, dropout=0.1, use_spk_id=False, num_spk=1, schedule_type='linear', infer=False,
                     use_shallow_diffusion=True, K_step_infer=100, diff_speedup=10, diff_accelerator='ddim', timesteps=1000, K_step=100,
                     spec_min=smin, spec_max=smax, diffusion_type='ddpm', backbone_type='wavenet',
                     backbone_args=dict(num_layers=6, num_channels=256, dilation_cycle_length=4),
                     shallow_diffusion_args=dict(train_aux_decoder=True, train_diffusion=True, val_gt_start=False, aux_decoder_grad=0.1,
                                                 aux_decoder_arch='convnext', aux_decoder_args=dict(num_channels=256, num_layers=2, kernel_size=7)),
                     hop_size=512, audio_sample_rate=44100, use_energy_embed=True, use_key_shift_embed=True, use_speed_embed=True,
                     augmentation_args=dict(random_pitch_shifting=dict(range=[-5., 5.]), random_time_stretching=dict(range=[0.5, 2.])),
                     b2s_precision='fp16', mel_base='e')
    _, vocab = _ds_param(0, 4)
    torch.manual_seed(8)
    model = P.DiffSingerAcoustic(len(vocab) + 1, 128)
    g = torch.Generator().manual_seed(9)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith('gamma'):
                p.copy_(0.2 + 0.3 * torch.rand(p.shape, generator=g))
            elif n.startswith('diffusion') and n.endswith('output_projection.weight') and p.dim() == 3 and p.shape[0] == 128:
                p.copy_(0.01 * torch.randn(p.shape, generator=g))
    model = model.to(dev).eval()
    gen = P.vocoder.Generator(dict(num_mels=128, sampling_rate=44100, upsample_rates=[8, 8, 2, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4, 4],
                                   upsample_initial_channel=512, resblock='1', resblock_kernel_sizes=[3, 7, 11],
                                   resblock_dilation_sizes=[[1, 3, 5]] * 3))
    gen.load_state_dict(OV.random_state_dict(OV.NsfHifiGanCfg(), 7), strict=True)
    voc = P.NsfHifiGAN(gen.to(dev).eval())
    return P, model, voc, vocab


def test_ds_to_wav_batched_equals_the_per_segment_loop(tmp_path):
    dev = torch.device('cuda:0')
    P, model, voc, vocab = _build(dev)
    params, offset = [], 0.0
    for i, n_ph in enumerate((9, 5, 14, 7, 11)):
        p, _ = _ds_param(700 + i, n_ph, energy=(-60., -10., 0.011), velocity=(0.6, 1.8, 0.05))
        p['offset'] = offset
        p['gender'] = (-0.5, 0.0, 0.3, None, 1.0)[i]
        if p['gender'] is None:
            del p['gender']
        if i % 2 == 0:
            p['seed'] = 100 + i
        dur = sum(float(v) for v in p['ph_dur'].split())
        offset += dur * (0.9 if i == 1 else 1.1)                # segment 2 starts inside segment 1: cross-fade; the others after a gap
        params.append(p)
    inf = P.infer.DiffSingerAcousticInfer(model, voc, vocab_list=vocab, device=dev)
    seq = inf.infer_segments(params, seed=11, batched=False)
    bat = inf.infer_segments(params, seed=11, batched=True)
    for i, (a, b) in enumerate(zip(seq, bat)):
        T = P.segments.segment_frames(params[i], inf.timestep)
        assert a['mel'].shape == b['mel'].shape == (1, T, 128) and a['f0'].shape == (1, T)
        assert bool(torch.isfinite(a['mel']).all()) and float(a['mel'].abs().max()) > 0
        assert torch.equal(a['mel'], b['mel']), f'segment {i}: the batched route differs from the one-segment-per-call loop'
        assert torch.equal(a['f0'], b['f0']) and a['offset'] == b['offset'] == params[i]['offset']
    wav = inf.run_inference(params, tmp_path, 'song', seed=11)
    assert (tmp_path / 'song.wav').exists() and wav.dtype == np.float64 and np.isfinite(wav).all() and np.abs(wav).max() <= 1.0
    last = round(params[-1]['offset'] * 44100) + seq[-1]['mel'].shape[1] * 512
    assert wav.shape[0] == last
    mel_list = inf.run_inference(params, tmp_path, 'song', seed=11, save_mel=True)
    loaded = torch.load(tmp_path / 'song.mel.pt')
    assert len(loaded) == 5 and all(torch.equal(x['mel'], y['mel']) for x, y in zip(loaded, mel_list))
    # the .mel.pt route of scripts/vocode.py gives a waveform of the same length
    wav2 = P.segments.vocode_segments(loaded, voc, 44100)
    assert wav2.shape == wav.shape


def _dist_worker(rank, world, port, n_gpus, q):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dev = torch.device('cuda', rank if n_gpus >= world else 0)
    torch.cuda.set_device(dev)
    nccl = n_gpus >= world
    dist.init_process_group('nccl' if nccl else 'gloo', rank=rank, world_size=world, **({'device_id': dev} if nccl else {}))
    try:
        P, model, voc, vocab = _build(dev)
        params = []
        for i, n_ph in enumerate((9, 5, 14, 7, 11, 6)):
            p, _ = _ds_param(800 + i, n_ph, energy=(-60., -10., 0.011))
            p['offset'], p['seed'] = 3.0 * i, 40 + i
            params.append(p)
        inf = P.infer.DiffSingerAcousticInfer(model, voc, vocab_list=vocab, device=dev)
        got = inf.infer_segments(params, batched=True)
        ok = True
        if rank == 0:
            dist.destroy_process_group()
            whole = inf.infer_segments(params, batched=True)            # the same project on one rank
            ok = len(got) == len(whole) == 6 and all(torch.equal(a['mel'], b['mel']) and torch.equal(a['f0'], b['f0']) for a, b in zip(got, whole))
        else:
            ok = got is None
        q.put((rank, ok))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def test_ds_driver_across_two_ranks_equals_one_rank():
    """``infer_segments`` under torch.distributed: segments partitioned by length over two ranks (cuda:0 / cuda:1 with NCCL when the box has
    two GPUs, else both on cuda:0 with gloo), gathered once on rank 0 - bit-identical to the one-rank run, None on the other rank."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    n_gpus = torch.cuda.device_count()
    procs = [ctx.Process(target=_dist_worker, args=(r, 2, port, n_gpus, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)], res
