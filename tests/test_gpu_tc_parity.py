"""GPU parity of the 16-bit tensor-core path (tcgen05 kernels) against the CPU oracle (fp32 restatement of
the reference) on seeded inputs, through the product's public API.

Tolerance (BASELINE.json north_star): max-abs de-normalised mel error <= 2e-2 for the 16-bit path - asserted ABSOLUTE for
fp16 operands, the product's 16-bit path and the bench dtype; bf16 operands are an opt-in mode whose error is reported and
only guarded against regressions (tests/tol16.py).  Under random init the sampled |mel| reaches ~300 (SURVEY.md H4:
eps_hat ~ 0, so x0 ~ x_T / sqrt(alpha_bar_T)), which is why sigma_w of the output projection is fixed at 0.01 and reported
with every number.  Every measurement is appended to gpurun_out/parity_report.jsonl before it is asserted."""
import json
import os

import pytest
import torch

from oracle import denoisers as OD
from test_gpu_parity import _maxabs, _oracle_and_product

pytestmark = pytest.mark.gpu

from tol16 import check16


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'gpu tests need a CUDA device'
    yield torch.device('cuda:0')
    import xiaoicesing_io_b200 as P
    P.hparams.pop('b2s_precision', None)


def _report(**kw):
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'parity_report.jsonl'), 'a') as f:
        f.write(json.dumps(kw) + '\n')
    print(kw)


@pytest.mark.parametrize('precision', ['bf16', 'fp16'])
def test_backbone_single_call_wavenet_20x256(precision, dev):
    """One denoiser evaluation (the unit the sampling loop repeats) at the full config-1 size."""
    import xiaoicesing_io_b200 as P
    from oracle import weights as OW
    cfg = OD.WaveNetCfg()
    P.hparams.clear()
    P.hparams.update(hidden_size=cfg.hidden_size, b2s_precision=precision)
    net = P.build_backbone(cfg.in_dims, 1, 'wavenet', dict(num_layers=20, num_channels=256, dilation_cycle_length=4))
    sd = OW.make_state_dict(cfg, seed=0, sigma_w=0.01)
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).eval()
    g = torch.Generator().manual_seed(11)
    B, T = 2, 345
    spec = torch.randn((B, 1, cfg.in_dims, T), generator=g)
    cond = torch.randn((B, cfg.hidden_size, T), generator=g)
    t = torch.tensor([437, 12])
    out = net(spec.to(dev), t.to(dev), cond.to(dev)).cpu()
    ref = OD.make_denoiser(sd, cfg)(spec, t, cond)
    err, scale = _maxabs(out, ref), float(ref.abs().max())
    _report(test='backbone_single_call', precision=precision, max_abs=err, ref_absmax=scale)
    assert err <= 2e-2 * max(1.0, scale)


@pytest.mark.parametrize('precision', ['bf16', 'fp16'])
@pytest.mark.parametrize('acc', ['ddim', 'dpm-solver', 'unipc'])
def test_config1_tensor_core(acc, precision, dev):
    cfg = OD.WaveNetCfg()
    out, ref = _oracle_and_product(
        cfg, dict(use_shallow_diffusion=False, diff_speedup=50, diff_accelerator=acc, b2s_precision=precision),
        {}, 1, 690, dev)
    err, scale = _maxabs(out, ref), float(ref.abs().max())
    _report(test='config1', sampler=acc, precision=precision, max_abs=err, ref_absmax=scale, sigma_w=0.01)
    check16(precision, err, scale)


@pytest.mark.parametrize('precision', ['bf16', 'fp16'])
def test_config2_shape_tensor_core(precision, dev):
    cfg = OD.WaveNetCfg()
    out, ref = _oracle_and_product(
        cfg, dict(use_shallow_diffusion=True, K_step_infer=40, diff_speedup=1, b2s_precision=precision),
        dict(k_step=40), 3, 173, dev, n_draws=41, src=True)
    err, scale = _maxabs(out, ref), float(ref.abs().max())
    _report(test='config2_shape_ddpm40', precision=precision, max_abs=err, ref_absmax=scale, sigma_w=0.01)
    check16(precision, err, scale)


@pytest.mark.parametrize('precision', ['bf16', 'fp16'])
def test_config5_shape_tensor_core(precision, dev):
    cfg = OD.WaveNetCfg(num_channels=512)
    out, ref = _oracle_and_product(
        cfg, dict(use_shallow_diffusion=False, diff_speedup=100, diff_accelerator='unipc', b2s_precision=precision),
        {}, 2, 131, dev)
    err, scale = _maxabs(out, ref), float(ref.abs().max())
    _report(test='config5_shape_unipc10_C512', precision=precision, max_abs=err, ref_absmax=scale, sigma_w=0.01)
    check16(precision, err, scale)


@pytest.mark.parametrize('precision', ['bf16', 'fp16'])
def test_config3_shape_lynxnet_tensor_core(precision, dev):
    """BASELINE config 3 shape on the tensor cores: rectified-flow Euler 20 with LYNXNet (C=1024 shrunk to 256 so that the
    oracle finishes in seconds), strong_cond, depthwise k=31."""
    import xiaoicesing_io_b200 as P
    from oracle import samplers as OS
    from oracle import weights as OW
    cfg = OD.LYNXNetCfg(num_channels=256, num_layers=6, kernel_size=31, strong_cond=True, hidden_size=256)
    P.hparams.clear()
    P.hparams.update(hidden_size=256, use_shallow_diffusion=False, sampling_algorithm='euler', sampling_steps=20,
                     infer=False, b2s_precision=precision)
    bargs = dict(num_layers=6, num_channels=256, kernel_size=31, strong_cond=True)
    model = P.RectifiedFlow(128, backbone_type='lynxnet', backbone_args=bargs, spec_min=[-12.], spec_max=[0.])
    sd = OW.make_state_dict(cfg, seed=0, sigma_w=0.01)
    model.velocity_fn.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    g = torch.Generator().manual_seed(7)
    B, T = 2, 211
    condition = torch.randn((B, T, 256), generator=g)
    noise0 = torch.randn((B, 1, 128, T), generator=g)
    draws = iter([noise0])
    model._noise_source = lambda shape: next(draws).to(dev)
    out = model(condition.to(dev), infer=True).cpu()
    x = OS.rectified_flow_inference(OD.make_denoiser(sd, cfg), condition.transpose(1, 2), t_start=0.,
                                    use_shallow=False, algorithm='euler', steps=20, noise0=noise0)
    ref = OS.denorm_spec(x, torch.tensor(-12.), torch.tensor(0.))
    err, scale = _maxabs(out, ref), float(ref.abs().max())
    _report(test='config3_shape_lynx_euler20', precision=precision, max_abs=err, ref_absmax=scale, sigma_w=0.01)
    check16(precision, err, scale)


@pytest.mark.parametrize('precision', ['bf16', 'fp16'])
def test_lynxnet_weak_cond_gelu_single_call(precision, dev):
    """strong_cond=False: GELU after the input projection (lynxnet.py:142-143), SiLU activation, k=7."""
    import xiaoicesing_io_b200 as P
    from oracle import weights as OW
    cfg = OD.LYNXNetCfg(num_channels=128, num_layers=3, kernel_size=7, strong_cond=False, hidden_size=64, activation='SiLU',
                        in_dims=64)
    P.hparams.clear()
    P.hparams.update(hidden_size=64, b2s_precision=precision)
    net = P.build_backbone(64, 1, 'lynxnet', dict(num_layers=3, num_channels=128, kernel_size=7, strong_cond=False,
                                                   activation='SiLU'))
    sd = OW.make_state_dict(cfg, seed=1, sigma_w=0.05)
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).eval()
    g = torch.Generator().manual_seed(3)
    B, T = 3, 77
    spec = torch.randn((B, 1, 64, T), generator=g)
    cond = torch.randn((B, 64, T), generator=g)
    t = torch.tensor([500.5])
    out = net(spec.to(dev), t.to(dev), cond.to(dev)).cpu()
    ref = OD.make_denoiser(sd, cfg)(spec, t, cond)
    err, scale = _maxabs(out, ref), float(ref.abs().max())
    _report(test='lynx_weak_gelu_single_call', precision=precision, max_abs=err, ref_absmax=scale)
    assert err <= 2e-2 * max(1.0, scale)


@pytest.mark.parametrize('precision', ['bf16', 'fp16'])
@pytest.mark.parametrize('acc', ['dpm-solver', 'unipc'])
def test_config4_variance_predictors_tensor_core(acc, precision, dev):
    """BASELINE config 4 on the tensor cores: the multi-variance predictor (2 curves x 24 repeat bins = 48 packed bins,
    WaveNet 10 x 192, configs/variance.yaml:95-100) under DPM-Solver++ / UniPC 10 steps.  C = 192 runs the two-kernel
    tensor-core path (N = 384 is not a multiple of the 256-column tile, K = 48 < 64, N = 48 output bins)."""
    import xiaoicesing_io_b200 as P
    from oracle import samplers as OS
    from oracle import weights as OW
    cfg = OD.WaveNetCfg(in_dims=24, n_feats=2, num_layers=10, num_channels=192, dilation_cycle_length=4, hidden_size=256)
    P.hparams.clear()
    P.hparams.update(hidden_size=256, schedule_type='linear', use_shallow_diffusion=False, diff_speedup=100,
                     diff_accelerator=acc, infer=False, b2s_precision=precision)
    model = P.MultiVarianceDiffusion(ranges=[(-96., -12.), (-96., -20.)], clamps=[(-96., -12.), (-96., -20.)], repeat_bins=24,
                                     backbone_type='wavenet',
                                     backbone_args=dict(num_layers=10, num_channels=192, dilation_cycle_length=4))
    sd = OW.make_state_dict(cfg, seed=0, sigma_w=0.01)
    model.denoise_fn.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    g = torch.Generator().manual_seed(21)
    B, T = 3, 257
    cond = torch.randn((B, 256, T), generator=g)
    noise0 = torch.randn((B, 2, 24, T), generator=g)
    draws = iter([noise0])
    model._noise_source = lambda shape: next(draws).to(dev)
    x = model.inference(cond.to(dev), B, None, dev).cpu()                      # normalised [B, F, T, M]
    sch = OS.DiffusionSchedule(1000, 'linear')
    ref = OS.gaussian_diffusion_inference(OD.make_denoiser(sd, cfg), sch, cond, k_step=1000, timesteps=1000, use_shallow=False,
                                          K_step_infer=1000, speedup=100, accelerator=acc, noise0=noise0, x_start=None,
                                          step_noise=[])
    assert tuple(x.shape) == tuple(ref.shape)
    raw_err, raw_scale = _maxabs(x, ref), float(ref.abs().max())
    # the model's output are the CURVES: de-normalise (ddpm.py:415-421: (x + 1) / 2 * (vmax - vmin) + vmin), mean over the 24 repeat
    # bins; compared BEFORE the clamp (random-init sampling leaves the clamp range, which would make the comparison vacuous)
    lo = torch.tensor([-96., -96.]).reshape(1, 2, 1, 1)
    hi = torch.tensor([-12., -20.]).reshape(1, 2, 1, 1)
    curves = lambda t: ((t.double() + 1) / 2 * (hi - lo) + lo).mean(-1)
    err, scale = float((curves(x) - curves(ref)).abs().max()), float(curves(ref).abs().max())
    _report(test='config4_multivariance_10', sampler=acc, precision=precision, max_abs=err, ref_absmax=scale,
            max_abs_normalised_bins=raw_err, normalised_absmax=raw_scale)
    check16(precision, err, scale)


def _bf16_backbone(dev, stack, L=4, fuse_io=False, in_dims=128, n_feats=1, stack_t=False, cycle=4, precision='bf16', stack3=False,
                   head3=True):
    import xiaoicesing_io_b200 as P
    from oracle import weights as OW
    cfg = OD.WaveNetCfg(in_dims=in_dims, n_feats=n_feats, num_layers=L, num_channels=256, dilation_cycle_length=cycle)
    P.hparams.clear()
    P.hparams.update(hidden_size=cfg.hidden_size, b2s_precision=precision, b2s_stack=stack, b2s_fuse_io=fuse_io, b2s_stack_t=stack_t,
                     b2s_stack3=stack3, b2s_stack3_head=head3)
    net = P.build_backbone(cfg.in_dims, n_feats, 'wavenet', dict(num_layers=L, num_channels=256, dilation_cycle_length=cycle))
    net.load_state_dict(OW.make_state_dict(cfg, seed=0, sigma_w=0.01), strict=True)
    return net.to(dev).eval()


@pytest.mark.parametrize('precision', ['fp16', 'bf16'])
@pytest.mark.parametrize('channels,L,cycle,in_dims,n_feats,B,T', [(192, 10, 4, 24, 2, 7, 690), (192, 4, 5, 128, 1, 3, 257),
                                                                  (136, 3, 2, 64, 1, 2, 130)])
def test_narrow_wavenet_on_the_padded_whole_stack_kernel(channels, L, cycle, in_dims, n_feats, B, T, precision, dev):
    """Narrow WaveNets (config 4's 192-channel variance predictor) run on the 256-channel whole-stack kernel with zero-padded
    channels (engine.py: a padded channel is exactly 0 everywhere).  Checked against the CPU oracle at the model's own width, and -
    where the unpadded two-kernel path exists (C % 64 == 0) - against that path to 16-bit rounding; padded output bins do not
    exist (the head's rows are not padded), and the session must really have taken the whole-stack path."""
    import xiaoicesing_io_b200 as P
    from oracle import weights as OW
    cfg = OD.WaveNetCfg(in_dims=in_dims, n_feats=n_feats, num_layers=L, num_channels=channels, dilation_cycle_length=cycle)
    sd = OW.make_state_dict(cfg, seed=2, sigma_w=0.01)
    g = torch.Generator().manual_seed(channels + T)
    spec = torch.randn((B, n_feats, in_dims, T), generator=g)
    cond = torch.randn((B, cfg.hidden_size, T), generator=g)
    t = torch.arange(B, dtype=torch.float32) * 29 + 3
    outs = {}
    for pad in (True, False):
        if not pad and channels % 64:
            continue
        P.hparams.clear()
        P.hparams.update(hidden_size=cfg.hidden_size, b2s_precision=precision, b2s_pad_channels=pad)
        net = P.build_backbone(in_dims, n_feats, 'wavenet', dict(num_layers=L, num_channels=channels, dilation_cycle_length=cycle))
        net.load_state_dict(sd, strict=True)
        net = net.to(dev).eval()
        outs[pad] = net(spec.to(dev), t.to(dev), cond.to(dev)).cpu()
        eng = net._engine()
        assert eng.C == (256 if pad else channels) and eng.C0 == channels
        sess = eng.begin(cond.to(dev).transpose(1, 2).contiguous(), t.to(dev), per_row_t=True)
        assert sess.stack3 == pad and sess.head3 == pad
    # the narrow-model mode of the whole-stack kernels (the all-zero fourth K slab skipped, GEMM1's second half with N = 128) adds
    # and multiplies exact zeros less: bit-identical to the plain padded run, with and without the skip / head kernel
    for head3 in (True, False):
        res = []
        for narrow in (True, False):
            P.hparams.clear()
            P.hparams.update(hidden_size=cfg.hidden_size, b2s_precision=precision, b2s_pad_channels=True, b2s_narrow_slabs=narrow,
                             b2s_stack3_head=head3)
            net = P.build_backbone(in_dims, n_feats, 'wavenet', dict(num_layers=L, num_channels=channels, dilation_cycle_length=cycle))
            net.load_state_dict(sd, strict=True)
            net = net.to(dev).eval()
            assert net._engine().C_used == (192 if narrow else 256)
            res.append(net(spec.to(dev), t.to(dev), cond.to(dev)))
        assert torch.equal(res[0], res[1]), (head3, float((res[0] - res[1]).abs().max()))
        if head3:
            assert torch.equal(res[0].cpu(), outs[True])
    ref = OD.make_denoiser(sd, cfg)(spec, t.long(), cond)
    eps = 2 ** -8 if precision == 'bf16' else 2 ** -11
    scale = float(ref.abs().max())
    err = _maxabs(outs[True], ref)
    _report(test='narrow_wavenet_padded', channels=channels, L=L, precision=precision, max_abs_err=err, ref_scale=scale)
    assert err <= 2e-2 * max(1.0, scale), (err, scale)
    if False in outs:
        d = _maxabs(outs[True], outs[False])
        assert d <= 6 * eps * scale, (d, scale)


def test_whole_stack_path_randomised_shapes(dev):
    """Randomised shapes through the whole-stack kernels (fixed seed): batch sizes that need one / several chained groups, utterance
    lengths from 1 frame to 9 tiles (clusters of 2 / 4 / 6 / 8 and utterances spanning clusters), ragged lengths, 2-20 layers,
    dilation cycles 1-5, 24-128 packed bins.  Each case: finite, equal to itself on a second run (no race), and equal to the
    per-layer kernels within 16-bit rounding of the layer inputs."""
    import random
    import xiaoicesing_io_b200 as P
    rnd = random.Random(1234)
    for case in range(14):
        L = rnd.choice([2, 3, 5, 8, 20])
        cycle = rnd.choice([1, 2, 4, 5])
        in_dims = rnd.choice([24, 64, 128])
        T = rnd.choice([1, 37, 128, 129, 300, 511, 690, 1023, 1100])
        B = rnd.choice([1, 2, 5, 17, 33])
        if B * T > 24000:
            B = max(1, 24000 // T)
        g = torch.Generator().manual_seed(case)
        spec = torch.randn((B, 1, in_dims, T), generator=g).to(dev)
        cond = torch.randn((B, 256, T), generator=g).to(dev)
        t = (torch.arange(B, dtype=torch.float32) * 3 + 1).to(dev)
        net = _bf16_backbone(dev, stack=True, L=L, cycle=cycle, in_dims=in_dims, precision='fp16', stack3=True)
        a = net(spec, t, cond)
        a2 = net(spec, t, cond)
        b = _bf16_backbone(dev, stack=False, L=L, cycle=cycle, in_dims=in_dims, precision='fp16')(spec, t, cond)
        tag = (case, B, T, L, cycle, in_dims)
        assert bool(torch.isfinite(a).all()), tag
        assert torch.equal(a, a2), tag
        scale = float(b.abs().max())
        assert float((a - b).abs().max()) <= 4 * 2 ** -11 * max(scale, 1e-3), (tag, float((a - b).abs().max()), scale)
        # ragged: every utterance's valid frames equal the run of that utterance alone at its own length
        if T >= 130 and B >= 2:
            P.hparams.update(b2s_stack=True, b2s_stack3=True)          # (the per-layer reference above switched them off)
            lens = [T] + [rnd.randrange(1, T + 1) for _ in range(B - 1)]
            sess_out = _ragged_eval(net, spec, t, cond, lens)
            for bi in (1, B - 1):
                n = lens[bi]
                solo = net(spec[bi:bi + 1, ..., :n].contiguous(), t[bi:bi + 1], cond[bi:bi + 1, :, :n].contiguous())
                assert torch.equal(sess_out[bi, ..., :n], solo[0]), (tag, bi, n)


def _ragged_eval(net, spec, t, cond, lens):
    """One backbone evaluation of a ragged batch through the engine session (the public ragged entry is the sampler's lengths=)."""
    from xiaoicesing_io_b200 import _cabi as C
    B, F_, M, T = spec.shape
    dev = spec.device
    eng = net._engine()
    sess = eng.begin(cond.transpose(1, 2).contiguous(), t, per_row_t=True, lens=torch.tensor(lens, dtype=torch.int32, device=dev))
    x_tm = torch.empty((B * T, F_ * M), device=dev)
    C.transpose(spec.reshape(B, F_ * M, T).contiguous(), x_tm, B, F_ * M, T)
    out_tm = torch.empty_like(x_tm)
    sess.eval(x_tm, 0, out_tm)
    out = torch.empty((B, F_ * M, T), device=dev)
    C.transpose(out_tm, out, B, T, F_ * M)
    return out.reshape(B, F_, M, T)


@pytest.mark.parametrize('B,T', [(40, 690), (70, 345), (150, 129)])
def test_chained_utterance_groups_do_not_change_a_bit(B, T, dev):
    """Batches larger than one launch run as several utterance groups per evaluation; by default a group's layer kernel does not
    wait for the previous group's skip / head tail (b2s_tc_wavenet_denoiser3_chained: late griddepcontrol.wait).  Same bits as the
    plain back-to-back launches, single call and a whole sampling loop (eager and CUDA-graph replay)."""
    import xiaoicesing_io_b200 as P
    g = torch.Generator().manual_seed(B + T)
    spec = torch.randn((B, 1, 128, T), generator=g).to(dev)
    cond = torch.randn((B, 256, T), generator=g).to(dev)
    t = (torch.arange(B, dtype=torch.float32) * 7 + 3).to(dev)
    outs = []
    for chain in (True, False):
        net = _bf16_backbone(dev, stack=True, L=6, precision='fp16', stack3=True)
        P.hparams['b2s_chain_groups'] = chain
        sess = net._engine().begin(cond.transpose(1, 2).contiguous(), t, per_row_t=True)
        assert sess.stack3 and sess.head3 and B > sess.stack_group, 'the shape must need several groups'
        outs.append(net(spec, t, cond))
    assert bool(torch.isfinite(outs[0]).all())
    assert torch.equal(outs[0], outs[1])
    from oracle import weights as OW
    res = []
    for chain in (True, False):
        cfg = OD.WaveNetCfg(num_layers=4)
        P.hparams.clear()
        P.hparams.update(hidden_size=cfg.hidden_size, schedule_type='linear', use_shallow_diffusion=False, diff_speedup=100,
                         diff_accelerator='ddim', b2s_precision='fp16', b2s_chain_groups=chain)
        m = P.GaussianDiffusion(out_dims=128, num_feats=1, timesteps=1000, k_step=1000, backbone_type='wavenet',
                                backbone_args=dict(num_layers=4, num_channels=256, dilation_cycle_length=4),
                                spec_min=[-12.0], spec_max=[0.0])
        m.denoise_fn.load_state_dict(OW.make_state_dict(cfg, seed=0, sigma_w=0.01), strict=True)
        m = m.to(dev).eval()
        runs = []
        for _ in range(3):                                  # eager, capture, replay
            torch.manual_seed(5)
            runs.append(m(cond.transpose(1, 2).contiguous(), infer=True))
        assert torch.equal(runs[0], runs[1]) and torch.equal(runs[0], runs[2])
        res.append(runs[0])
    assert torch.equal(res[0], res[1])


@pytest.mark.parametrize('B,T', [(30, 690), (5, 129), (1, 19500)])
def test_stack_kernel_grouping_and_fallback_match_per_layer_path(B, T, dev):
    """The whole-stack kernel needs every tile resident: batches with more tiles than SMs are split by utterance into
    several launches (B=30 x 690 -> 2 groups), and an utterance that alone exceeds the SM count (19500 frames = 154 tiles)
    falls back to the per-layer kernels.  All of them must agree with the per-layer path bit for bit, also with one
    diffusion step per utterance (d_stride != 0)."""
    g = torch.Generator().manual_seed(B * 7 + T)
    spec = torch.randn((B, 1, 128, T), generator=g).to(dev)
    cond = torch.randn((B, 256, T), generator=g).to(dev)
    for t in (torch.tensor([437.0]), torch.arange(B, dtype=torch.float32) * 13 + 5):
        a = _bf16_backbone(dev, stack=True)(spec, t.to(dev), cond)
        b = _bf16_backbone(dev, stack=False)(spec, t.to(dev), cond)
        assert bool(torch.isfinite(a).all())
        assert torch.equal(a, b), (B, T, float((a - b).abs().max()))


@pytest.mark.parametrize('precision', ['bf16', 'fp16'])
@pytest.mark.parametrize('B,T,L,cycle,in_dims,n_feats', [(30, 690, 4, 4, 128, 1), (5, 129, 6, 4, 128, 1), (16, 690, 20, 4, 128, 1),
                                                        (3, 257, 5, 5, 64, 1), (2, 130, 4, 4, 24, 2), (1, 1292, 4, 4, 128, 1),
                                                        (40, 345, 4, 4, 128, 1)])
def test_stack3_matches_per_layer_path(B, T, L, cycle, in_dims, n_feats, precision, dev):
    """The third whole-stack kernel (b2s_stack3, the default: cta_group::2 pairs, y resident in shared memory, residual stream in
    TMEM, halo rows through distributed shared memory, deferred skip GEMM) against the per-layer kernels: same 16-bit operands,
    but the fp32 residual stream is carried as 2^(l/2)-scaled accumulator, so the two agree to 16-bit rounding of y, not bit for
    bit.  Shapes: several launches per batch (30 x 690, 40 x 345), clusters of 6 / 4 / 2 tiles, an utterance spanning two
    clusters (1292 frames), dilation 16 (cycle 5), 64 and 48 packed bins, per-utterance diffusion steps."""
    g = torch.Generator().manual_seed(B * 7 + T)
    spec = torch.randn((B, n_feats, in_dims, T), generator=g).to(dev)
    cond = torch.randn((B, 256, T), generator=g).to(dev)
    eps = 2 ** -8 if precision == 'bf16' else 2 ** -11
    for t in (torch.tensor([437.0]), torch.arange(B, dtype=torch.float32) * 13 + 5):
        net = _bf16_backbone(dev, stack=True, L=L, cycle=cycle, in_dims=in_dims, n_feats=n_feats, precision=precision, stack3=True)
        a = net(spec, t.to(dev), cond)
        b = _bf16_backbone(dev, stack=False, L=L, cycle=cycle, in_dims=in_dims, n_feats=n_feats, precision=precision)(spec, t.to(dev), cond)
        assert bool(torch.isfinite(a).all())
        scale = float(b.abs().max())
        err = float((a - b).abs().max())
        assert err <= 4 * eps * scale, (B, T, L, precision, err, scale)
        # the skip sum + head on extra CTAs of the same launch (default) against the three separate GEMM launches: same operands,
        # same accumulation order
        c = _bf16_backbone(dev, stack=True, L=L, cycle=cycle, in_dims=in_dims, n_feats=n_feats, precision=precision, stack3=True,
                           head3=False)(spec, t.to(dev), cond)
        assert float((a - c).abs().max()) <= 0.25 * eps * scale, (B, T, L, precision, float((a - c).abs().max()), scale)


@pytest.mark.parametrize('B,T,in_dims,n_feats', [(16, 690, 128, 1), (30, 300, 128, 1), (3, 257, 64, 1), (2, 130, 24, 2), (1, 50, 128, 1)])
def test_one_launch_denoiser_matches_separate_stem_and_head(B, T, in_dims, n_feats, dev):
    """b2s_fuse_io: stem and head run inside the persistent stack kernel (one launch per evaluation).  Same operands, same
    accumulation order -> the result must equal the stem GEMM + stack kernel + two head GEMMs bit for bit, for 128 / 64 /
    48 packed bins, several utterance groups and per-utterance diffusion steps."""
    g = torch.Generator().manual_seed(B * 11 + T)
    spec = torch.randn((B, n_feats, in_dims, T), generator=g).to(dev)
    cond = torch.randn((B, 256, T), generator=g).to(dev)
    for t in (torch.tensor([437.0]), torch.arange(B, dtype=torch.float32) * 13 + 5):
        a = _bf16_backbone(dev, True, fuse_io=True, in_dims=in_dims, n_feats=n_feats)(spec, t.to(dev), cond)
        b = _bf16_backbone(dev, True, fuse_io=False, in_dims=in_dims, n_feats=n_feats)(spec, t.to(dev), cond)
        assert bool(torch.isfinite(a).all())
        assert torch.equal(a, b), (B, T, in_dims, float((a - b).abs().max()))


@pytest.mark.parametrize('B,T,L,cycle', [(2, 300, 4, 4), (16, 690, 20, 4), (3, 81, 5, 5), (1, 40, 3, 2), (5, 129, 6, 4), (7, 1000, 4, 4)])
@pytest.mark.parametrize('precision', ['bf16', 'fp16'])
def test_transposed_stack_matches_row_stack(B, T, L, cycle, precision, dev):
    """b2s_stack_t: the residual stack with channels on the tensor-core M axis and 32..80-frame tiles on N (x in registers,
    skip sum in TMEM).  Same operands as the 128-row stack kernel; only the fp32 summation order of the skip path differs
    (one TMEM accumulator over all layers instead of a per-layer fp32 add), so the outputs agree to fp32 rounding of the
    16-bit head inputs.  Shapes: one launch with 144 tiles of 80 (config 2), odd tile counts (dummy CTA), dilation 16
    (cycle 5), tiles shorter than one frame tile, per-utterance diffusion steps."""
    from xiaoicesing_io_b200 import _cabi
    if not _cabi.HAS_EXPERIMENTS:
        pytest.skip('the transposed stack kernel is an experiment: B2S_BUILD_EXPERIMENTS=1 build only')
    g = torch.Generator().manual_seed(B * 5 + T)
    spec = torch.randn((B, 1, 128, T), generator=g).to(dev)
    cond = torch.randn((B, 256, T), generator=g).to(dev)
    for t in (torch.tensor([437.0]), torch.arange(B, dtype=torch.float32) * 13 + 5):
        net = _bf16_backbone(dev, stack=True, L=L, stack_t='always', cycle=cycle, precision=precision)
        a = net(spec, t.to(dev), cond)
        b = _bf16_backbone(dev, stack=True, L=L, stack_t=False, cycle=cycle, precision=precision)(spec, t.to(dev), cond)
        assert bool(torch.isfinite(a).all())
        scale = float(b.abs().max())
        err = float((a - b).abs().max())
        assert err <= 4e-3 * scale, (B, T, L, err, scale)


@pytest.mark.parametrize('precision', ['bf16', 'fp16'])
@pytest.mark.parametrize('acc', ['dpm-solver', 'unipc'])
def test_config4_pitch_predictor_full_size(acc, precision, dev):
    """BASELINE config 4, the PITCH predictor at its real size (configs/variance.yaml:67-77): PitchDiffusion over 64 repeat bins,
    WaveNet 20 x 256 with dilation cycle 5 (dilations up to 16: the widest halo of the whole-stack kernel), DPM-Solver++ /
    UniPC 10 steps, 8 utterances x 690 frames, against the oracle on the GPU in strict fp32."""
    import xiaoicesing_io_b200 as P
    from oracle import samplers as OS
    from oracle import weights as OW
    cfg = OD.WaveNetCfg(in_dims=64, n_feats=1, num_layers=20, num_channels=256, dilation_cycle_length=5, hidden_size=256)
    P.hparams.clear()
    P.hparams.update(hidden_size=256, schedule_type='linear', use_shallow_diffusion=False, diff_speedup=100,
                     diff_accelerator=acc, infer=False, b2s_precision=precision)
    model = P.PitchDiffusion(vmin=-8., vmax=8., cmin=-12., cmax=12., repeat_bins=64, backbone_type='wavenet',
                             backbone_args=dict(num_layers=20, num_channels=256, dilation_cycle_length=5))
    sd = OW.make_state_dict(cfg, seed=0, sigma_w=0.01)
    model.denoise_fn.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    g = torch.Generator().manual_seed(31)
    B, T = 8, 690
    cond = torch.randn((B, 256, T), generator=g)
    noise0 = torch.randn((B, 1, 64, T), generator=g)
    draws = iter([noise0])
    model._noise_source = lambda shape: next(draws).to(dev)
    x = model.inference(cond.to(dev), B, None, dev)                            # normalised [B, T, 64]
    out = model.denorm_spec(x).cpu()                                           # the pitch-delta curve [B, T], clipped to [-12, 12]
    P.hparams.pop('b2s_precision', None)
    sd_gpu = {k: v.to(dev) for k, v in sd.items()}
    denoise = lambda xx, t, c: OD.wavenet_forward(sd_gpu, cfg, xx, t.to(dev), c)
    sch = OS.DiffusionSchedule(1000, 'linear')
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            xr = OS.gaussian_diffusion_inference(denoise, sch, cond.to(dev), k_step=1000, timesteps=1000, use_shallow=False,
                                                 K_step_infer=1000, speedup=100, accelerator=acc, noise0=noise0.to(dev),
                                                 x_start=None, step_noise=[])
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    assert tuple(x.shape) == tuple(xr.shape)
    # reference de-normalisation (ddpm.py:415-421, 441-445): (x + 1) / 2 * (vmax - vmin) + vmin, mean over the bins, clip
    ref = (((xr.cpu() + 1) / 2) * 16. - 8.).mean(-1).clamp(-12., 12.)
    raw_err, raw_scale = _maxabs(x, xr), float(xr.abs().max())
    err, scale = _maxabs(out, ref), float(ref.abs().max())
    _report(test='config4_pitch_20x256_cycle5', sampler=acc, precision=precision, max_abs=err, ref_absmax=scale,
            max_abs_normalised_bins=raw_err, normalised_absmax=raw_scale)
    check16(precision, err, scale)
    # the un-averaged, un-clipped bins in de-normalised units (slope 8 per normalised unit): the same bound
    check16(precision, 8. * raw_err, 8. * raw_scale)


@pytest.mark.parametrize('precision', ['bf16', 'fp16'])
def test_pndm_full_width_tensor_core(precision, dev):
    """PNDM / PLMS (ddpm.py:169-204, :323-333) on the 16-bit path at the full WaveNet 20 x 256 width, one 8-s utterance
    (the reference's PNDM only runs for B = 1), 20 steps = 21 evaluations."""
    cfg = OD.WaveNetCfg()
    out, ref = _oracle_and_product(
        cfg, dict(use_shallow_diffusion=False, diff_speedup=50, diff_accelerator='pndm', b2s_precision=precision), {}, 1, 690, dev)
    err, scale = _maxabs(out, ref), float(ref.abs().max())
    _report(test='pndm20_wavenet_20x256', precision=precision, max_abs=err, ref_absmax=scale, sigma_w=0.01)
    check16(precision, err, scale)


@pytest.mark.parametrize('precision', ['bf16', 'fp16'])
@pytest.mark.parametrize('alg', ['rk4', 'rk2', 'rk5'])
def test_reflow_runge_kutta_full_width_tensor_core(alg, precision, dev):
    """Rectified flow with the Runge-Kutta integrators (reflow.py:72-102) on the 16-bit path at full width: WaveNet 20 x 256
    velocity network, 10 steps (40 / 20 / 60 evaluations), 4 utterances x 690 frames, oracle on the GPU in strict fp32."""
    import xiaoicesing_io_b200 as P
    from oracle import samplers as OS
    from oracle import weights as OW
    cfg = OD.WaveNetCfg()
    P.hparams.clear()
    P.hparams.update(hidden_size=256, use_shallow_diffusion=False, sampling_algorithm=alg, sampling_steps=10, infer=False,
                     b2s_precision=precision)
    model = P.RectifiedFlow(128, backbone_type='wavenet', backbone_args=dict(num_layers=20, num_channels=256, dilation_cycle_length=4),
                            spec_min=[-12.], spec_max=[0.])
    sd = OW.make_state_dict(cfg, seed=0, sigma_w=0.01)
    model.velocity_fn.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    g = torch.Generator().manual_seed(17)
    B, T = 4, 690
    condition = torch.randn((B, T, 256), generator=g)
    noise0 = torch.randn((B, 1, 128, T), generator=g)
    draws = iter([noise0])
    model._noise_source = lambda shape: next(draws).to(dev)
    out = model(condition.to(dev), infer=True).cpu()
    P.hparams.pop('b2s_precision', None)
    sd_gpu = {k: v.to(dev) for k, v in sd.items()}
    velocity = lambda x, t, c: OD.wavenet_forward(sd_gpu, cfg, x, t.to(dev), c)
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            x = OS.rectified_flow_inference(velocity, condition.transpose(1, 2).to(dev), t_start=0., use_shallow=False,
                                            algorithm=alg, steps=10, noise0=noise0.to(dev))
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    ref = OS.denorm_spec(x.cpu(), torch.tensor(-12.), torch.tensor(0.))
    err, scale = _maxabs(out, ref), float(ref.abs().max())
    _report(test=f'reflow_{alg}_10_wavenet_20x256', precision=precision, max_abs=err, ref_absmax=scale, sigma_w=0.01)
    check16(precision, err, scale)
