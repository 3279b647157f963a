"""CPU tests of the PRODUCT's host logic (no GPU, no kernels): the sampler compiler in
``xiaoicesing_io_b200.schedules`` must reproduce the reference's samplers when its programs are
executed with the ORACLE denoiser, and the drop-in modules must keep the reference's state-dict."""
import pytest
import torch

import golden_util as GU
import product_util as PU
from oracle import denoisers as OD
from oracle import samplers as OS

DF = [n for n in GU.fixture_names() if not n.startswith(('bb_', 'aux_', 'enc_', 'voc_', 'ds_'))]


def _maxabs(a, b):
    return float((a.double() - b.double()).abs().max())


@pytest.mark.parametrize('name', DF)
def test_program_matches_reference_fixture(name):
    fx = GU.Fixture(name)
    model = PU.build_model(fx)                       # CPU construction: parameters only
    prog = model.build_program()
    m = fx.meta
    M, nf, smin, smax, clamps = GU.variance_geometry(m)
    cfg = GU.backbone_cfg(m['ctor']['backbone_type'], m['ctor']['backbone_args'], M, nf, m['hidden_size'])
    denoise = OD.make_denoiser(fx.sd, cfg, torch.float64)
    cond = fx['condition'].transpose(1, 2).double()
    draws = fx['draws']
    src = fx['src_spec'] if 'src_spec' in fx else None
    x_start = model._source_to_state(src)
    shape = tuple(draws[0].shape)
    assert prog.n_draws == m['n_draws'] - 1
    x = PU.run_program_cpu(prog, denoise, cond, shape, draws[0], x_start, draws[1:])
    x = x.transpose(2, 3).squeeze(1).float()
    out = model.denorm_spec(x)
    ref = GU.expected_outputs(fx)
    outs = out if isinstance(out, list) else [out]
    refs = ref if isinstance(ref, list) else [ref]
    for o, r in zip(outs, refs):
        assert o.shape == r.shape
        scale = max(1.0, float(r.abs().max()))
        # float64 host coefficients + fp64 denoiser vs the fp32 reference
        assert _maxabs(o, r) <= 1e-3 * scale / 3, (name, _maxabs(o, r), scale)


@pytest.mark.parametrize('name', ['gd_unipc_10', 'gd_dpmsolver_5', 'gd_pndm_10', 'rf_rk5_4_lynx', 'gd_ddpm_full_T30'])
def test_program_matches_oracle_fp64(name):
    """Same comparison against the oracle's own fp64 run: isolates the host coefficient tables
    (both sides fp64 arithmetic -> only table rounding differs)."""
    fx = GU.Fixture(name)
    model = PU.build_model(fx)
    prog = model.build_program()
    truth = GU.run_oracle_diffusion(fx, dtype=torch.float64)
    m = fx.meta
    M, nf, smin, smax, clamps = GU.variance_geometry(m)
    cfg = GU.backbone_cfg(m['ctor']['backbone_type'], m['ctor']['backbone_args'], M, nf, m['hidden_size'])
    denoise = OD.make_denoiser(fx.sd, cfg, torch.float64)
    cond = fx['condition'].transpose(1, 2).double()
    draws = fx['draws']
    x = PU.run_program_cpu(prog, denoise, cond, tuple(draws[0].shape), draws[0], None, draws[1:])
    out = OS.denorm_spec(x.transpose(2, 3).squeeze(1), smin.double(), smax.double())
    scale = max(1.0, float(truth.abs().max()))
    assert _maxabs(out, truth) <= 2e-5 * scale


@pytest.mark.parametrize('name', GU.fixture_names('bb_') + ['gd_ddim_10', 'rf_euler_4_lynx', 'var_multi_unipc_10'])
def test_state_dict_is_drop_in(name):
    """Reference checkpoints load with strict=True: identical parameter names and shapes."""
    fx = GU.Fixture(name)
    if fx.meta['kind'] == 'backbone':
        net = PU.build_backbone(fx)
    else:
        model = PU.build_model(fx)
        net = getattr(model, model.backbone_attr)
    sd = net.state_dict()
    assert set(sd.keys()) == set(fx.sd.keys())
    for k, v in sd.items():
        assert v.shape == fx.sd[k].shape
        assert torch.equal(v.cpu(), fx.sd[k])


def test_nfe_counts_match_reference():
    """NFE per sampler as probed on the reference: DDIM 20, PNDM 21, DPM 20, UniPC 20."""
    import xiaoicesing_io_b200 as P
    P.hparams.clear()
    P.hparams.update(hidden_size=16, schedule_type='linear', use_shallow_diffusion=False, diff_speedup=50)
    args = dict(out_dims=8, backbone_type='wavenet', backbone_args=dict(num_layers=1, num_channels=16),
                spec_min=[-12.], spec_max=[0.])
    m = P.GaussianDiffusion(**args)
    for acc, n in (('ddim', 20), ('pndm', 21), ('dpm-solver', 20), ('unipc', 20)):
        P.hparams['diff_accelerator'] = acc
        assert m.build_program().n_nfe == n, acc
    P.hparams.update(diff_speedup=1)
    assert m.build_program().n_nfe == 1000 and m.build_program().n_draws == 1000
    P.hparams.update(diff_speedup=7)
    with pytest.raises(AssertionError):
        m.build_program()
    P.hparams.update(diff_speedup=50, diff_accelerator='euler')
    with pytest.raises(ValueError):
        m.build_program()
    # model-time inputs: DDIM-20 950..0; DPM-Solver++/UniPC-20 999.0, 949.05, ...
    P.hparams['diff_accelerator'] = 'ddim'
    assert m.build_program().t_values[:3] == [950.0, 900.0, 850.0]
    P.hparams['diff_accelerator'] = 'unipc'
    tv = m.build_program().t_values
    assert abs(tv[0] - 999.0) < 1e-3 and abs(tv[1] - 949.05) < 1e-3 and abs(tv[-1] - 49.95) < 1e-2


def test_no_cpu_fallback():
    """The product path fails loudly on CPU tensors instead of falling back."""
    import xiaoicesing_io_b200 as P
    from xiaoicesing_io_b200._cabi import B2SError
    fx = GU.Fixture('gd_ddim_10')
    model = PU.build_model(fx)
    with pytest.raises(B2SError):
        model(fx['condition'], infer=True)
    net = PU.build_backbone(GU.Fixture('bb_wavenet_int_t'))
    bb = GU.Fixture('bb_wavenet_int_t')
    with pytest.raises(B2SError):
        net(bb['spec'], bb['t'], bb['cond'])
