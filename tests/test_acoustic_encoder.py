"""FastSpeech2 acoustic encoder: host-side checks that need no GPU."""
import pytest
import torch

import golden_util as GU


def _model(meta):
    import xiaoicesing_io_b200 as P
    P.hparams.clear()
    P.hparams.update(meta['hparams'])
    P.hparams.update(meta['flags'])
    P.hparams.update(use_pos_embed=True, rel_pos=True, use_rope=True, dropout=0.1)
    return P.FastSpeech2Acoustic(meta['vocab'])


@pytest.mark.parametrize('name', GU.fixture_names('enc_'))
def test_state_dict_is_drop_in(name):
    """Parameter names and shapes equal the reference's (the fixture holds the reference module's state dict, including the
    rotary frequencies that appear once per layer)."""
    fx = GU.Fixture(name)
    m = _model(fx.meta)
    assert set(m.state_dict()) == set(fx.sd)
    m.load_state_dict(fx.sd, strict=True)
    ref_freqs = fx.sd['encoder.layers.0.op.self_attn.rotary_embed.freqs']
    fresh = _model(fx.meta).state_dict()['encoder.layers.0.op.self_attn.rotary_embed.freqs']
    assert torch.equal(fresh, ref_freqs), 'the default rotary frequencies must be the reference library\'s'


def test_unsupported_configurations_raise_at_construction():
    import xiaoicesing_io_b200 as P
    base = dict(hidden_size=64, enc_layers=1, enc_ffn_kernel_size=3, ffn_act='gelu', num_heads=2, use_pos_embed=True, use_rope=True,
                use_spk_id=False, num_spk=1)
    for bad in (dict(use_rope=False), dict(ffn_act='swiglu')):
        P.hparams.clear()
        P.hparams.update(base)
        P.hparams.update(bad)
        with pytest.raises(NotImplementedError):
            P.FastSpeech2Acoustic(10)
    P.hparams.clear()
    P.hparams.update(base)
    m = P.FastSpeech2Acoustic(10)
    with pytest.raises(P.B2SError):
        m(torch.ones(1, 3, dtype=torch.long), torch.ones(1, 5, dtype=torch.long), torch.ones(1, 5))
