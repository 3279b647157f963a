"""ConvNeXt aux decoder: host-side checks that need no GPU - the module is a drop-in parameter
container for the reference's checkpoints, and fails loudly without a CUDA device."""
import pytest
import torch

import golden_util as GU


def _model(meta):
    import xiaoicesing_io_b200 as P
    return P.AuxDecoderAdaptor(in_dims=meta['in_dims'], out_dims=meta['out_dims'], num_feats=meta['num_feats'],
                               spec_min=meta['spec_min'], spec_max=meta['spec_max'], aux_decoder_arch='convnext',
                               aux_decoder_args=meta['args'])


@pytest.mark.parametrize('name', GU.fixture_names('aux_'))
def test_state_dict_is_drop_in(name):
    """Parameter names and shapes equal the reference's (the fixture holds the reference decoder's state dict); the adaptor's
    spec_min / spec_max buffers are non-persistent like the reference's (aux_decoder/__init__.py:47-48)."""
    fx = GU.Fixture(name)
    m = _model(fx.meta)
    m.decoder.load_state_dict(fx.sd, strict=True)
    assert not any(k.startswith('spec_') for k in m.state_dict())
    assert set(m.state_dict()) == {'decoder.' + k for k in fx.sd}


def test_registry_and_kwarg_filtering():
    import xiaoicesing_io_b200 as P
    dec = P.build_aux_decoder(256, 128, 'convnext', dict(num_channels=128, num_layers=2, kernel_size=5, dropout_rate=0.3,
                                                          not_an_argument=1))
    assert isinstance(dec, P.AUX_DECODERS['convnext']) and len(dec.conv) == 2 and dec.inconv.kernel_size == (5,)
    with pytest.raises(KeyError):
        P.build_aux_decoder(256, 128, 'unet', {})


def test_cpu_module_raises():
    import xiaoicesing_io_b200 as P
    dec = P.build_aux_decoder(256, 128, 'convnext', dict(num_channels=128, num_layers=1))
    with pytest.raises(P.B2SError):
        dec(torch.zeros(1, 10, 256))
