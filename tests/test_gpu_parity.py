"""GPU parity tests (run with -m gpu on a B200).  Everything here goes through the C ABI of
libb2s.so via the product's public API and is compared with (a) the committed fixtures produced by
the unmodified reference and (b) the CPU oracle on seeded inputs.

Tolerance (BASELINE.json north_star): max-abs de-normalised mel error <= 1e-3 for the fp32 path."""
import math

import pytest
import torch

import golden_util as GU
import product_util as PU
from oracle import denoisers as OD
from oracle import samplers as OS
from oracle import weights as OW

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-3
BB = GU.fixture_names('bb_')
DF = [n for n in GU.fixture_names() if not n.startswith(('bb_', 'aux_', 'enc_', 'voc_', 'ds_'))]


def _maxabs(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max())


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'gpu tests need a CUDA device'
    import xiaoicesing_io_b200 as P
    P.hparams.pop('b2s_precision', None)
    return torch.device('cuda:0')


def _inject(model, draws, dev):
    it = iter(draws)
    model._noise_source = lambda shape: next(it).to(dev)


@pytest.mark.parametrize('name', BB)
def test_backbone_forward_vs_reference_fixture(name, dev):
    fx = GU.Fixture(name)
    net = PU.build_backbone(fx, dev)
    out = net(fx['spec'].to(dev), fx['t'].to(dev), fx['cond'].to(dev))
    ref = fx['out']
    assert out.shape == ref.shape
    assert _maxabs(out, ref) <= 2e-5 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize('name', DF)
def test_sampling_vs_reference_fixture(name, dev):
    fx = GU.Fixture(name)
    model = PU.build_model(fx, dev)
    _inject(model, fx['draws'], dev)
    kw = {}
    if 'src_spec' in fx:
        kw['src_spec'] = fx['src_spec'].to(dev)
    out = model(fx['condition'].to(dev), infer=True, **kw)
    ref = GU.expected_outputs(fx)
    outs = out if isinstance(out, (list, tuple)) else [out]
    refs = ref if isinstance(ref, list) else [ref]
    assert len(outs) == len(refs)
    for o, r in zip(outs, refs):
        assert tuple(o.shape) == tuple(r.shape)
        # 1e-3 absolute; the cosine-schedule fixtures blow up to |mel| ~ 4e5 under random init, where one
        # fp32 ulp is already 0.03 - there the bound is 16 ulps of the largest magnitude.
        tol = max(FP32_TOL, 16 * 1.1920929e-07 * float(r.abs().max()))
        assert _maxabs(o, r) <= tol, (name, _maxabs(o, r), float(r.abs().max()))


def _oracle_and_product(cfg, hp, sampler_kw, B, T, dev, seed=1234, sigma_w=0.01, n_draws=1, src=False, oracle_on_gpu=False):
    import xiaoicesing_io_b200 as P
    P.hparams.clear()
    P.hparams.update(hidden_size=cfg.hidden_size, schedule_type='linear', infer=False, **hp)
    bargs = dict(num_layers=cfg.num_layers, num_channels=cfg.num_channels,
                 dilation_cycle_length=cfg.dilation_cycle_length)
    model = P.GaussianDiffusion(cfg.in_dims, backbone_type='wavenet', backbone_args=bargs,
                                spec_min=[-12.], spec_max=[0.], **sampler_kw)
    sd = OW.make_state_dict(cfg, seed=0, sigma_w=sigma_w)
    model.denoise_fn.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    g = torch.Generator().manual_seed(seed)
    condition = torch.randn((B, T, cfg.hidden_size), generator=g)
    draws = [torch.randn((B, 1, cfg.in_dims, T), generator=g) for _ in range(n_draws)]
    src_spec = (torch.rand((B, T, cfg.in_dims), generator=g) * 12 - 12) if src else None
    _inject(model, draws, dev)
    out = model(condition.to(dev), src_spec=None if src_spec is None else src_spec.to(dev), infer=True).cpu()
    sch = OS.DiffusionSchedule(sampler_kw.get('timesteps', 1000), 'linear')
    x_start = None
    if src:
        x_start = OS.norm_spec(src_spec, torch.tensor(-12.), torch.tensor(0.)).transpose(-2, -1)[:, None]
    use_shallow = hp.get('use_shallow_diffusion', False)
    k_step = sampler_kw.get('k_step', 1000) if use_shallow else sampler_kw.get('timesteps', 1000)
    denoise, to = OD.make_denoiser(sd, cfg), (lambda t: t)
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    if oracle_on_gpu:
        # the oracle is plain torch code over a state dict: with its tensors on the GPU (TF32 off) it is a strict-fp32 truth
        # that finishes FULL-SIZE batches in seconds (the CPU oracle needs minutes for 16 x 690 frames)
        sd_gpu = {k: v.to(dev) for k, v in sd.items()}
        denoise = lambda x, t, c: OD.wavenet_forward(sd_gpu, cfg, x, t.to(dev), c)
        to = lambda t: t.to(dev)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            x = OS.gaussian_diffusion_inference(
                denoise, sch, to(condition.transpose(1, 2)), k_step=k_step,
                timesteps=sampler_kw.get('timesteps', 1000), use_shallow=use_shallow,
                K_step_infer=hp.get('K_step_infer', k_step), speedup=hp['diff_speedup'],
                accelerator=hp.get('diff_accelerator', 'ddim'), noise0=to(draws[0]),
                x_start=None if x_start is None else to(x_start), step_noise=[to(d) for d in draws[1:]])
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    ref = OS.denorm_spec(x.cpu(), torch.tensor(-12.), torch.tensor(0.))
    return out, ref


@pytest.mark.parametrize('acc', ['ddim', 'dpm-solver', 'unipc'])
def test_config1_full_size_wavenet_20x256(acc, dev):
    """BASELINE config 1: WaveNet 20x256, 128 mel, one 8-s utterance (690 frames), 20 steps."""
    cfg = OD.WaveNetCfg()
    out, ref = _oracle_and_product(cfg, dict(use_shallow_diffusion=False, diff_speedup=50, diff_accelerator=acc),
                                   {}, 1, 690, dev)
    err = _maxabs(out, ref)
    print(f'config1 {acc}: max-abs {err:.3e}, |mel|max {float(ref.abs().max()):.1f}')
    assert err <= FP32_TOL


@pytest.mark.parametrize('K', [8, 400])
@pytest.mark.parametrize('precision,tol', [('fp32', FP32_TOL), ('bf16', 2e-2), ('fp16', 2e-2)])
def test_config2_full_batch_against_gpu_oracle(precision, tol, K, dev):
    """BASELINE config 2 at its FULL size: 16 utterances x 690 frames, WaveNet 20x256, shallow ancestral sampling with
    per-step noise, K = 8 and the headline's K_step = 400.  The oracle runs on the GPU in strict fp32 (TF32 off) on the
    same pre-drawn noise (401 draws of [16, 1, 128, 690] for K = 400)."""
    cfg = OD.WaveNetCfg()
    out, ref = _oracle_and_product(
        cfg, dict(use_shallow_diffusion=True, K_step_infer=K, diff_speedup=1, b2s_precision=precision), dict(k_step=K),
        16, 690, dev, n_draws=K + 1, src=True, oracle_on_gpu=True)
    import xiaoicesing_io_b200 as P
    P.hparams.pop('b2s_precision', None)
    err = _maxabs(out, ref)
    print(f'config2 full batch ddpm K={K} {precision}: max-abs {err:.3e}, |mel|max {float(ref.abs().max()):.1f}')
    assert err <= tol


def test_config2_shape_shallow_ddpm_ragged_T(dev):
    """BASELINE config 2 at a size the oracle finishes in seconds: shallow DDPM full-step (K=40 of 1000),
    WaveNet 20x256, B=3, T=173 (not a tile multiple), per-step ancestral noise."""
    cfg = OD.WaveNetCfg()
    out, ref = _oracle_and_product(
        cfg, dict(use_shallow_diffusion=True, K_step_infer=40, diff_speedup=1), dict(k_step=40),
        3, 173, dev, n_draws=41, src=True)
    err = _maxabs(out, ref)
    print(f'config2-shape ddpm K=40: max-abs {err:.3e}, |mel|max {float(ref.abs().max()):.1f}')
    assert err <= FP32_TOL


def test_config5_shape_wavenet_20x512_unipc(dev):
    cfg = OD.WaveNetCfg(num_channels=512)
    out, ref = _oracle_and_product(cfg, dict(use_shallow_diffusion=False, diff_speedup=100, diff_accelerator='unipc'),
                                   {}, 2, 131, dev)
    err = _maxabs(out, ref)
    print(f'config5-shape unipc-10 C=512: max-abs {err:.3e}, |mel|max {float(ref.abs().max()):.1f}')
    assert err <= FP32_TOL


def test_config3_shape_lynxnet_reflow(dev):
    """BASELINE config 3 shape: rectified-flow Euler 20 with LYNXNet (C=1024 shrunk to 256 for the oracle)."""
    import xiaoicesing_io_b200 as P
    cfg = OD.LYNXNetCfg(num_channels=256, num_layers=6, kernel_size=31, strong_cond=True, hidden_size=256)
    P.hparams.clear()
    P.hparams.update(hidden_size=256, use_shallow_diffusion=False, sampling_algorithm='euler', sampling_steps=20,
                     infer=False)
    bargs = dict(num_layers=6, num_channels=256, kernel_size=31, strong_cond=True)
    model = P.RectifiedFlow(128, backbone_type='lynxnet', backbone_args=bargs, spec_min=[-12.], spec_max=[0.])
    sd = OW.make_state_dict(cfg, seed=0, sigma_w=0.01)
    model.velocity_fn.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    g = torch.Generator().manual_seed(7)
    B, T = 2, 211
    condition = torch.randn((B, T, 256), generator=g)
    noise0 = torch.randn((B, 1, 128, T), generator=g)
    _inject(model, [noise0], dev)
    out = model(condition.to(dev), infer=True).cpu()
    x = OS.rectified_flow_inference(OD.make_denoiser(sd, cfg), condition.transpose(1, 2), t_start=0.,
                                    use_shallow=False, algorithm='euler', steps=20, noise0=noise0)
    ref = OS.denorm_spec(x, torch.tensor(-12.), torch.tensor(0.))
    err = _maxabs(out, ref)
    print(f'config3-shape lynx euler-20: max-abs {err:.3e}, |mel|max {float(ref.abs().max()):.1f}')
    assert err <= FP32_TOL


def test_seeded_rng_order_matches_torch(dev):
    """Without injection the product draws with torch.randn in the reference's order: initial noise
    first, then one draw per ancestral step -> a seeded run is reproducible and equals the injected run."""
    fx = GU.Fixture('gd_ddpm_shallow_K12')
    model = PU.build_model(fx, dev)
    src = fx['src_spec'].to(dev)
    cond = fx['condition'].to(dev)
    torch.manual_seed(99)
    a = model(cond, src_spec=src, infer=True)
    torch.manual_seed(99)
    shape = tuple(fx['draws'][0].shape)
    draws = [torch.randn(shape, device=dev) for _ in range(13)]
    _inject(model, draws, dev)
    b = model(cond, src_spec=src, infer=True)
    assert torch.equal(a, b)


def test_batch_rows_are_independent(dev):
    """No op on the path mixes utterances: sampling a batch equals sampling each
    utterance alone, bit for bit - the property the multi-GPU partition relies on."""
    fx = GU.Fixture('gd_unipc_10')
    model = PU.build_model(fx, dev)
    draws = fx['draws']
    _inject(model, draws, dev)
    full = model(fx['condition'].to(dev), infer=True)
    for i in range(fx['condition'].shape[0]):
        _inject(model, [d[i:i + 1] for d in draws], dev)
        one = model(fx['condition'][i:i + 1].to(dev), infer=True)
        assert torch.equal(one[0], full[i])


def test_cuda_graph_replay_equals_eager(dev):
    """The second call with the same shapes captures the whole loop (tables, noise draws, every kernel) as a
    CUDA graph; seeded replays must reproduce the eager result bit for bit (same Philox consumption)."""
    import xiaoicesing_io_b200 as P
    from xiaoicesing_io_b200.core import _sampling
    fx = GU.Fixture('gd_ddpm_shallow_K12')
    model = PU.build_model(fx, dev)
    P.hparams['b2s_cuda_graph'] = True
    _sampling.clear_graph_cache()
    src, cond = fx['src_spec'].to(dev), fx['condition'].to(dev)
    outs = []
    for _ in range(3):                      # eager, capture + replay, replay
        torch.manual_seed(5)
        outs.append(model(cond, src_spec=src, infer=True).clone())
    assert any(not isinstance(v, str) for v in _sampling._GRAPH_CACHE.values()), 'the loop was never captured'
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    torch.manual_seed(6)
    other = model(cond, src_spec=src, infer=True)
    assert not torch.equal(other, outs[0])
    _sampling.clear_graph_cache()


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
@pytest.mark.parametrize('graph', [False, True])
def test_launch_structure_switches_do_not_change_results(precision, graph, dev):
    """``b2s_overlap_noise`` (per-step noise draws on a side stream, a parallel branch of the captured graph),
    ``b2s_fuse_cast`` (the sampler update writes the 16-bit denoiser input itself) and ``b2s_fuse_update`` (the update runs
    inside the whole-denoiser launch; bf16 / fp16 sessions only) only change WHICH launches run WHERE:
    seeded ancestral sampling must give bit-identical mels with the switches on and off, eagerly and from a graph replay
    (same draw order on the default generator, same arithmetic)."""
    import xiaoicesing_io_b200 as P
    from xiaoicesing_io_b200.core import _sampling
    outs = []
    for overlap, fuse_cast, fuse_update in [(True, True, True), (False, True, True), (True, True, False), (False, True, False),
                                           (True, False, False), (False, False, False)]:
        P.hparams.clear()
        P.hparams.update(hidden_size=256, schedule_type='linear', use_shallow_diffusion=True, K_step_infer=9, diff_speedup=1,
                         infer=False, b2s_precision=precision, b2s_cuda_graph=graph, b2s_overlap_noise=overlap,
                         b2s_fuse_cast=fuse_cast, b2s_fuse_update=fuse_update)
        torch.manual_seed(0)
        model = P.GaussianDiffusion(128, k_step=9, backbone_type='wavenet',
                                    backbone_args=dict(num_layers=4, num_channels=256, dilation_cycle_length=4),
                                    spec_min=[-12.], spec_max=[0.])
        torch.nn.init.normal_(model.denoise_fn.output_projection.weight, std=0.01)
        model = model.to(dev).eval()
        g = torch.Generator().manual_seed(3)
        cond = torch.randn((3, 150, 256), generator=g).to(dev)
        src = (torch.rand((3, 150, 128), generator=g) * 12 - 12).to(dev)
        _sampling.clear_graph_cache()
        for _ in range(2 if graph else 1):          # with graphs: the second call captures and replays
            torch.manual_seed(11)
            out = model(cond, src_spec=src, infer=True).clone()
        outs.append(out)
        assert bool(torch.isfinite(out).all())
    _sampling.clear_graph_cache()
    P.hparams.pop('b2s_precision', None)
    for o in outs[1:]:
        assert torch.equal(o, outs[0])


def test_edge_cases_empty_and_tiny_batches(dev):
    """Edge cases: a single frame, T smaller than every dilation, and an empty batch (no kernel may fault)."""
    import xiaoicesing_io_b200 as P
    for precision in ('fp32', 'bf16'):
        P.hparams.clear()
        P.hparams.update(hidden_size=256, schedule_type='linear', use_shallow_diffusion=False, diff_speedup=200,
                         diff_accelerator='ddim', infer=False, b2s_precision=precision, b2s_cuda_graph=False)
        torch.manual_seed(0)
        model = P.GaussianDiffusion(128, backbone_type='wavenet',
                                    backbone_args=dict(num_layers=5, num_channels=256, dilation_cycle_length=5),
                                    spec_min=[-12.], spec_max=[0.])
        torch.nn.init.normal_(model.denoise_fn.output_projection.weight, std=0.01)
        model = model.to(dev).eval()
        for (B, T) in [(1, 1), (2, 3), (3, 129)]:
            out = model(torch.randn(B, T, 256, device=dev), infer=True)
            assert tuple(out.shape) == (B, T, 128) and bool(torch.isfinite(out).all()), (precision, B, T)
        out = model(torch.randn(0, 7, 256, device=dev), infer=True)
        assert tuple(out.shape) == (0, 7, 128)
    P.hparams.pop('b2s_precision', None)


def test_cpu_inputs_are_rejected_loudly(dev):
    """No CPU fallback: a CPU condition tensor raises instead of silently computing somewhere else."""
    import xiaoicesing_io_b200 as P
    fx = GU.Fixture('gd_ddim_5')
    model = PU.build_model(fx, dev)
    with pytest.raises(P._cabi.B2SError):
        model(fx['condition'], infer=True)
