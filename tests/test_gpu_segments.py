"""GPU test of the batched segment driver (xiaoicesing_io_b200/segments.py): a RAGGED batch of .ds segments gives every segment
the bits of its own B = 1 run (the reference's inference shape, inference/ds_acoustic.py:209-219) - per-utterance lengths inside
the whole-stack kernels, per-segment seeded noise, CUDA-graph replay of the bucketed shape."""
import json

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TIMESTEP = 512 / 44100
H, M = 256, 128


def _model(dev, precision, acc='unipc', steps=5, L=6, cycle=4):
    import xiaoicesing_io_b200 as P
    from oracle import denoisers as OD
    from oracle import weights as OW
    cfg = OD.WaveNetCfg(num_layers=L, num_channels=256, dilation_cycle_length=cycle)
    P.hparams.clear()
    P.hparams.update(hidden_size=H, schedule_type='linear', use_shallow_diffusion=False, diff_speedup=1000 // steps,
                     diff_accelerator=acc, infer=False, b2s_precision=precision)
    model = P.GaussianDiffusion(M, backbone_type='wavenet', backbone_args=dict(num_layers=L, num_channels=256, dilation_cycle_length=cycle),
                                spec_min=[-12.], spec_max=[0.])
    model.denoise_fn.load_state_dict(OW.make_state_dict(cfg, seed=0, sigma_w=0.01), strict=True)
    return model.to(dev).eval()


def _project(frames):
    """A .ds project whose segments have exactly the requested frame counts."""
    segs = []
    for i, n in enumerate(frames):
        total = (n - 0.25) * TIMESTEP                       # round(total / timestep + 0.5) == n
        segs.append({'offset': 2.0 * i, 'ph_seq': 'a b', 'ph_dur': f'{total * 0.4:.7f} {total * 0.6:.7f}', 'seed': 1000 + i})
    return segs


def _cond_fn(param, n):
    g = torch.Generator().manual_seed(int(param['seed']))
    return torch.randn((n, H), generator=g), None, torch.full((n,), 220.0)


@pytest.mark.parametrize('acc,cycle', [('unipc', 4), ('ddim', 5)])
def test_ragged_batch_equals_per_segment_runs_bitwise(acc, cycle):
    from xiaoicesing_io_b200 import segments as SG
    dev = torch.device('cuda:0')
    frames = [300, 129, 690, 50, 257, 691, 128]
    params = _project(frames)
    assert [SG.segment_frames(p, TIMESTEP) for p in params] == frames
    model = _model(dev, 'fp16', acc=acc, cycle=cycle)
    res = SG.sample_segments(model, params, _cond_fn, TIMESTEP, dev, max_batch_frames=4 * 768)
    assert sorted(res) == list(range(len(frames)))
    for i, n in enumerate(frames):
        cond, _, _ = _cond_fn(params[i], n)
        noise = SG.seeded_noise((1, 1, M, n), SG.segment_seed(params[i]), dev)
        solo = model(cond[None].to(dev), infer=True, initial_noise=noise).cpu()        # the reference's shape: B = 1, T = frames
        assert tuple(res[i]['mel'].shape) == (1, n, M) and res[i]['offset'] == 2.0 * i
        assert torch.equal(res[i]['mel'], solo), (acc, i, n, float((res[i]['mel'] - solo).abs().max()))
    # second and third call: the bucketed shapes are captured as CUDA graphs and replayed - same bits
    again = SG.sample_segments(model, params, _cond_fn, TIMESTEP, dev, max_batch_frames=4 * 768)
    third = SG.sample_segments(model, params, _cond_fn, TIMESTEP, dev, max_batch_frames=4 * 768)
    for i in range(len(frames)):
        assert torch.equal(again[i]['mel'], res[i]['mel']) and torch.equal(third[i]['mel'], res[i]['mel']), i


def test_ragged_batches_are_rejected_loudly_off_the_whole_stack_path():
    import xiaoicesing_io_b200 as P
    from xiaoicesing_io_b200 import segments as SG
    dev = torch.device('cuda:0')
    model = _model(dev, 'fp32')
    with pytest.raises(P.B2SError):
        SG.sample_segments(model, _project([100, 60]), _cond_fn, TIMESTEP, dev)


def test_mel_pt_written_in_segment_order(tmp_path):
    from xiaoicesing_io_b200 import segments as SG
    dev = torch.device('cuda:0')
    frames = [200, 64, 333]
    params = _project(frames)
    (tmp_path / 'p.ds').write_text(json.dumps(params))
    model = _model(dev, 'fp16')
    entries = SG.sample_segments_distributed(model, SG.load_ds(tmp_path / 'p.ds'), _cond_fn, TIMESTEP, dev)
    SG.save_mel_pt(tmp_path / 'p.mel.pt', entries)
    back = torch.load(tmp_path / 'p.mel.pt')
    assert [b['mel'].shape[1] for b in back] == frames and [b['offset'] for b in back] == [0.0, 2.0, 4.0]
    assert all(b['f0'].shape == (1, n) for b, n in zip(back, frames))
    assert all(bool(torch.isfinite(b['mel']).all()) for b in back)


def test_seeded_noise_equals_the_reference_reseeding_rule():
    """ds_acoustic.py:212-217 reseeds the process-wide generators before every segment; seeded_noise draws the same bits from a
    private generator (and leaves the global stream alone)."""
    from xiaoicesing_io_b200 import segments as SG
    dev = torch.device('cuda:0')
    for seed, shape in ((0, (1, 1, 128, 690)), (123456789, (1, 2, 24, 37)), (0xffffffff, (1, 1, 128, 1))):
        torch.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)
        want = torch.randn(shape, device=dev)
        torch.manual_seed(99)
        before = torch.cuda.get_rng_state(dev).clone()
        got = SG.seeded_noise(shape, seed, dev)
        assert torch.equal(got, want), seed
        assert torch.equal(torch.cuda.get_rng_state(dev), before)
