""".ds segment -> model inputs: ``DiffSingerAcousticInfer.preprocess_input`` / ``load_speaker_mix`` of
xiaoicesing_io_b200/infer.py against outputs of the UNMODIFIED reference methods (tests/golden/ds_*.npz, oracle/make_golden.py),
BIT-exact: tokens, mel2ph, f0 and variance curves, key shift, speed, speaker mix.  Plus the driver logic around it with stand-in models
(the GPU pipeline is tested in tests/test_gpu_infer_driver.py)."""
import types

import numpy as np
import pytest
import torch

import golden_util as GU

DS = GU.fixture_names('ds_')


def _infer(fx, model=None, vocoder=None):
    import xiaoicesing_io_b200 as P
    P.hparams.clear()
    P.hparams.update(fx.meta['hparams'])
    model = model if model is not None else types.SimpleNamespace(parameters=lambda: iter([torch.zeros(1)]))
    return P.infer.DiffSingerAcousticInfer(model, vocoder, vocab_list=fx.meta['vocab'], spk_map=fx.meta['spk_map'], device='cpu')


@pytest.mark.parametrize('name', DS)
def test_preprocess_input_is_bit_exact(name):
    fx = GU.Fixture(name)
    batch = _infer(fx).preprocess_input(fx.meta['param'])
    assert sorted(batch.keys()) == fx.meta['keys']
    for k in fx.meta['keys']:
        ref = fx[k]
        assert batch[k].shape == ref.shape and batch[k].dtype == ref.dtype, (k, batch[k].shape, ref.shape, batch[k].dtype, ref.dtype)
        assert torch.equal(batch[k], ref), k
    # the frame count the segment driver plans with equals the preprocessed length
    import xiaoicesing_io_b200 as P
    hp = fx.meta['hparams']
    assert P.segments.segment_frames(fx.meta['param'], hp['hop_size'] / hp['audio_sample_rate']) == batch['mel2ph'].shape[1]


def test_length_regulator_matches_the_mask_formulation():
    """tts_modules.py:299-311 restated literally (the [B, L, T] mask) against the repeat_interleave implementation."""
    from xiaoicesing_io_b200.infer import length_regulator
    g = torch.Generator().manual_seed(0)
    dur = torch.randint(0, 6, (3, 9), generator=g)
    pad = torch.zeros(3, 9, dtype=torch.bool)
    pad[1, 6:] = True
    d = dur * (1 - pad.long())
    token_idx = torch.arange(1, 10)[None, :, None]
    cs = torch.cumsum(d, 1)
    prev = torch.nn.functional.pad(cs, [1, -1])
    pos = torch.arange(int(d.sum(-1).max()))[None, None]
    ref = (token_idx * ((pos >= prev[:, :, None]) & (pos < cs[:, :, None])).long()).sum(1)
    assert torch.equal(length_regulator(dur, pad), ref)


def test_token_encoder_and_curve_edges():
    from xiaoicesing_io_b200.infer import TokenTextEncoder, resample_align_curve
    enc = TokenTextEncoder(['b', 'a', 'SP'])
    assert enc.encode('SP a <PAD> b') == [1, 2, 0, 3] and len(enc) == 4 and enc.decode([1, 0, 3]) == 'SP <PAD> b'
    with pytest.raises(ValueError):
        enc.encode('zz')
    c = resample_align_curve(np.array([1., 3.], np.float32), 0.1, 0.05, 5)            # 2 interpolated points, padded with the last
    assert c.dtype == np.float32 and np.allclose(c, [1., 2., 2., 2., 2.])
    assert len(resample_align_curve(np.arange(50, dtype=np.float32), 0.01, 0.005, 7)) == 7


def test_run_inference_wav_and_mel_pt_with_stand_ins(tmp_path):
    """The reference's loop (batched=False): per-segment reseeding, .mel.pt layout, silence / cross-fade assembly, file names."""
    fx = GU.Fixture('ds_preprocess_plain')
    seeds = []

    class Model:
        def parameters(self):
            return iter([torch.zeros(1)])

        def __call__(self, tokens, mel2ph, f0, infer, **kw):
            seeds.append(torch.initial_seed())
            return types.SimpleNamespace(diff_out=f0[..., None].repeat(1, 1, 4) * 1e-3)

    class Voc:
        device = torch.device('cpu')

        def spec2wav_torch(self, mel, f0=None):
            return mel[0, :, 0].repeat_interleave(2) * 0.5

    inf = _infer(fx, Model(), Voc())
    p0 = dict(fx.meta['param'], offset=0.0, seed=7)
    p1 = dict(fx.meta['param'], offset=0.5)
    entries = inf.run_inference([p0, p1], tmp_path, 'song', seed=3, save_mel=True, batched=False)
    assert seeds == [7, 3] and (tmp_path / 'song.mel.pt').exists()
    loaded = torch.load(tmp_path / 'song.mel.pt')
    assert [e['offset'] for e in loaded] == [0.0, 0.5] and loaded[0]['mel'].shape == entries[0]['mel'].shape
    T = entries[0]['mel'].shape[1]
    assert loaded[0]['f0'].shape == (1, T)
    wav = inf.run_inference([p0, p1], tmp_path, 'song', num_runs=2, save_mel=False, batched=False)
    assert (tmp_path / 'song-000.wav').exists() and (tmp_path / 'song-001.wav').exists()
    sr = fx.meta['hparams']['audio_sample_rate']
    assert wav.shape[0] == round(0.5 * sr) + 2 * T                                   # second segment placed at its offset
