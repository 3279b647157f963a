"""CPU tests of the batched segment driver's host logic (xiaoicesing_io_b200/segments.py): .ds parsing, the reference's frame
count and reseeding rules, ragged-batch planning, the .mel.pt layout.  No kernels involved."""
import json

import numpy as np
import pytest
import torch

from xiaoicesing_io_b200 import segments as SG

TIMESTEP = 512 / 44100


def _ds(tmp_path, n=5, as_list=True):
    rng = np.random.default_rng(0)
    segs = []
    for i in range(n):
        durs = rng.uniform(0.05, 0.6, size=int(rng.integers(3, 30)))
        segs.append({'offset': round(3.5 * i, 3), 'text': 'AP x SP', 'ph_seq': ' '.join(['a'] * len(durs)),
                     'ph_dur': ' '.join(f'{d:.4f}' for d in durs), 'f0_seq': '160.3 160.3', 'f0_timestep': '0.005'})
    if n > 1:
        segs[1]['seed'] = 0x1_0000_0005                  # wider than 32 bits: the reference masks it
    p = tmp_path / 'demo.ds'
    p.write_text(json.dumps(segs if as_list else segs[0]), encoding='utf-8')
    return p, segs


def test_load_ds_list_and_single_dict(tmp_path):
    p, segs = _ds(tmp_path)
    assert SG.load_ds(p) == segs
    p1, segs1 = _ds(tmp_path, n=1, as_list=False)
    assert SG.load_ds(p1) == [segs1[0]]
    (tmp_path / 'empty.ds').write_text('[]')
    with pytest.raises(ValueError):
        SG.load_ds(tmp_path / 'empty.ds')


def test_segment_frames_follows_the_reference_formula(tmp_path):
    """ds_acoustic.py:79-83: ph_acc = round(cumsum(ph_dur) / timestep + 0.5); frames = ph_acc[-1] (= sum of the per-phone frame
    counts the length regulator expands)."""
    _, segs = _ds(tmp_path, n=8)
    for s in segs:
        d = np.array(s['ph_dur'].split(), np.float32)
        acc = np.round(np.cumsum(d, dtype=np.float32) / np.float32(TIMESTEP) + np.float32(0.5)).astype(np.int64)
        assert SG.segment_frames(s, TIMESTEP) == int(acc[-1]) == int(np.diff(acc, prepend=0).sum())
    assert SG.segment_frames({'ph_dur': ''}, TIMESTEP) == 0


def test_segment_seed_rule(tmp_path):
    _, segs = _ds(tmp_path)
    assert SG.segment_seed(segs[1]) == 5                                  # & 0xffffffff (ds_acoustic.py:212-214)
    assert SG.segment_seed(segs[0]) is None and SG.segment_seed(segs[0], seed=7) == 7
    assert SG.segment_seed(segs[1], seed=7) == 5                          # the segment's own seed wins


def test_plan_batches_is_complete_bounded_and_deterministic():
    lengths = [690, 345, 2584, 1292, 690, 345, 345, 1292, 120, 0, 50, 700]
    for budget, bmax in ((16 * 704, 64), (4 * 704, 3), (100, 8)):
        batches = SG.plan_batches(lengths, budget, bmax)
        flat = sorted(i for b in batches for i in b)
        assert flat == [i for i in range(len(lengths)) if lengths[i] > 0]
        for b in batches:
            T = -(-max(lengths[i] for i in b) // 128) * 128
            assert len(b) <= bmax and (len(b) == 1 or len(b) * T <= budget)
            assert all(lengths[b[0]] >= lengths[i] for i in b)             # padded to its longest member only
        assert batches == SG.plan_batches(lengths, budget, bmax)


def test_mel_pt_layout_roundtrip(tmp_path):
    entries = [{'offset': 0.5 * i, 'mel': torch.randn(1, 10 + i, 128), 'f0': torch.rand(1, 10 + i), 'extra': 1} for i in range(3)]
    SG.save_mel_pt(tmp_path / 'x.mel.pt', entries)
    back = torch.load(tmp_path / 'x.mel.pt')
    assert isinstance(back, list) and len(back) == 3
    for e, b in zip(entries, back):
        assert set(b) == {'offset', 'mel', 'f0'} and b['offset'] == e['offset']
        assert torch.equal(b['mel'], e['mel']) and torch.equal(b['f0'], e['f0'])
    assert SG.real_time_factor(entries, 1.0, TIMESTEP) == pytest.approx(1.0 / (33 * TIMESTEP))


def test_vocode_segments_assembly_follows_the_reference(tmp_path):
    """Silence up to a segment's offset, cross-fade on overlap (scripts/vocode.py:64-84 with utils/infer_utils.py:89-96), 16-bit WAV
    (infer_utils.py:99-104) - with a stand-in vocoder (the assembly is host code; the vocoder itself is tested on the GPU)."""
    import numpy as np
    from scipy.io import wavfile
    from xiaoicesing_io_b200 import segments as S

    class Voc:
        device = torch.device('cpu')

        def spec2wav_torch(self, mel, f0=None):
            return (mel[0, :, :1] * torch.ones(1, 4)).reshape(-1) * f0[0].repeat_interleave(4)

    sr = 100
    entries = [dict(offset=0.05, mel=torch.full((1, 10, 3), 0.5), f0=torch.ones(1, 10)),
               dict(offset=0.30, mel=torch.full((1, 8, 3), -0.25), f0=torch.ones(1, 8)),        # starts at 30 < 5 + 40: overlaps by 15
               dict(offset=1.00, mel=torch.full((1, 5, 3), 0.1), f0=torch.ones(1, 5))]           # after a gap
    out = S.vocode_segments(entries, Voc(), sr)
    # the reference's loop, restated with its own cross_fade formula
    ref, cur = np.zeros(0), 0
    for e in entries:
        w = Voc().spec2wav_torch(e['mel'], f0=e['f0']).numpy()
        sil = round(e['offset'] * sr) - cur
        if sil >= 0:
            ref = np.concatenate([ref, np.zeros(sil), w])
        else:
            idx = cur + sil
            fade = ref.shape[0] - idx
            k = np.linspace(0, 1.0, num=fade, endpoint=True)
            ref = np.concatenate([ref[:idx], (1 - k) * ref[idx:] + k * w[:fade], w[fade:]])
        cur = cur + sil + w.shape[0]
    assert out.shape == ref.shape == (120,) and np.allclose(out, ref)
    assert out[:5].tolist() == [0.] * 5 and abs(out[5] - 0.5) < 1e-7 and abs(out[44] + 0.25) < 1e-7 and abs(out[-1] - 0.1) < 1e-7
    S.save_wav(out, tmp_path / 'a.wav', sr)
    rate, pcm = wavfile.read(tmp_path / 'a.wav')
    assert rate == sr and pcm.dtype == np.int16 and pcm[5] == int(0.5 * 32767) and pcm[-1] == int(0.1 * 32767)

