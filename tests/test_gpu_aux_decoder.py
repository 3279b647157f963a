"""GPU parity of the ConvNeXt aux decoder (producer of x_start) against the CPU oracle
(oracle/aux_decoder.py, pinned by tests/golden/aux_*.npz) through the product's public API.

Tolerance: the de-normalised mel within 2e-2 ABSOLUTE for fp16 operands (north_star's 16-bit bound; x_start is afterwards noised
to t = K_step, so this is far tighter than the path needs); bf16 reported with the regression guard of tests/tol16.py."""
import pytest
import torch

from oracle import aux_decoder as OA
from tol16 import check16

pytestmark = pytest.mark.gpu


def _random_sd(cfg: OA.ConvNeXtCfg, seed):
    import xiaoicesing_io_b200 as P
    torch.manual_seed(seed)
    m = P.AuxDecoderAdaptor(in_dims=cfg.in_dims, out_dims=cfg.out_dims, num_feats=cfg.num_feats, spec_min=None, spec_max=None,
                            aux_decoder_arch='convnext',
                            aux_decoder_args=dict(num_channels=cfg.num_channels, num_layers=cfg.num_layers, kernel_size=cfg.kernel_size))
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith('gamma'):
                p.copy_(0.2 + 0.3 * torch.rand(p.shape, generator=g))          # trained layer scales are O(0.1 .. 1), the init 1e-6
            elif n.endswith('norm.weight') or n.endswith('norm.bias'):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    return {k: v.detach().clone() for k, v in m.decoder.state_dict().items()}


@pytest.mark.parametrize('precision', ['fp16', 'bf16'])
@pytest.mark.parametrize('case', ['acoustic_default', 'feats2_k5', 'ragged_tiles', 'train_mode'])
def test_aux_decoder_against_oracle(case, precision):
    import xiaoicesing_io_b200 as P
    dev = torch.device('cuda:0')
    if case == 'acoustic_default':        # configs/acoustic.yaml:101-105: 512 channels, 6 layers, k = 7, 128 mel bins, H = 256
        cfg, B, T, smin, smax, infer = OA.ConvNeXtCfg(), 2, 300, [-12.] * 128, [0.] * 128, True
    elif case == 'feats2_k5':
        cfg = OA.ConvNeXtCfg(in_dims=64, out_dims=8, num_feats=2, num_channels=128, num_layers=3, kernel_size=5)
        B, T, smin, smax, infer = 3, 131, [[-10.] * 8, [-4.] * 8], [[2.] * 8, [6.] * 8], True
    elif case == 'ragged_tiles':          # utterances shorter than the kernel / not a multiple of any tile; 1 frame
        cfg = OA.ConvNeXtCfg(in_dims=64, out_dims=16, num_feats=1, num_channels=128, num_layers=2, kernel_size=7)
        B, T, smin, smax, infer = 5, 3, [-12.] * 16, [0.] * 16, True
    else:
        cfg, B, T, smin, smax, infer = OA.ConvNeXtCfg(num_channels=256, num_layers=2), 1, 257, [-12.] * 128, [0.] * 128, False
    sd = _random_sd(cfg, 5)
    P.hparams.clear()
    P.hparams.update(b2s_precision=precision)
    model = P.AuxDecoderAdaptor(in_dims=cfg.in_dims, out_dims=cfg.out_dims, num_feats=cfg.num_feats, spec_min=smin, spec_max=smax,
                                aux_decoder_arch='convnext',
                                aux_decoder_args=dict(num_channels=cfg.num_channels, num_layers=cfg.num_layers, kernel_size=cfg.kernel_size))
    model.decoder.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    g = torch.Generator().manual_seed(9)
    cond = torch.randn((B, T, cfg.in_dims), generator=g)
    out = model(cond.to(dev), infer=infer)
    ref = OA.aux_adaptor_forward(sd, cfg, cond, smin, smax, infer=infer, dtype=torch.float64)
    assert out.shape == ref.shape
    assert bool(torch.isfinite(out).all())
    err, scale = float((out.double().cpu() - ref).abs().max()), float(ref.abs().max())
    print(dict(test='aux_decoder', case=case, precision=precision, max_abs=err, ref_absmax=scale))
    check16(precision, err, scale)
    # a second call re-uses the packed weights; a weight update is picked up
    out2 = model(cond.to(dev), infer=infer)
    assert torch.equal(out, out2)
    with torch.no_grad():
        model.decoder.outconv.bias.add_(1.0)
    out3 = model(cond.to(dev), infer=infer)
    k = 1.0 if not infer else float((torch.tensor(smax).reshape(-1)[0] - torch.tensor(smin).reshape(-1)[0]) / 2)
    assert abs(float((out3 - out).mean()) - k) < 1e-3 * max(1.0, k) or cfg.num_feats > 1


def test_empty_batch():
    import xiaoicesing_io_b200 as P
    P.hparams.clear()
    dec = P.build_aux_decoder(64, 16, 'convnext', dict(num_channels=128, num_layers=1)).cuda()
    assert dec(torch.zeros(0, 10, 64, device='cuda')).shape == (0, 10, 16)
    assert dec(torch.zeros(2, 0, 64, device='cuda')).shape == (2, 0, 16)


def test_aux_decoder_feeds_shallow_diffusion_like_the_reference_toplevel():
    """The reference's acoustic inference (modules/toplevel.py:94-102): ``aux_mel = aux_decoder(condition, infer=True)`` then
    ``mel = diffusion(condition, src_spec=aux_mel, infer=True)`` (shallow diffusion from x_start = aux_mel noised to K_step).
    Both stages on the B200 kernels (fp16 operands) against the chain of their oracles, same injected noise."""
    import xiaoicesing_io_b200 as P
    from oracle import denoisers as OD, samplers as OS, weights as OW
    dev = torch.device('cuda:0')
    B, T = 2, 300
    acfg = OA.ConvNeXtCfg(num_channels=256, num_layers=3)
    asd = _random_sd(acfg, 21)
    wcfg = OD.WaveNetCfg()
    wsd = OW.make_state_dict(wcfg, seed=0, sigma_w=0.01)
    smin, smax = [-12.] * 128, [0.] * 128
    P.hparams.clear()
    P.hparams.update(hidden_size=256, schedule_type='linear', infer=False, use_shallow_diffusion=True, K_step_infer=100,
                     diff_speedup=10, diff_accelerator='ddim', b2s_precision='fp16')
    aux = P.AuxDecoderAdaptor(in_dims=256, out_dims=128, num_feats=1, spec_min=smin, spec_max=smax, aux_decoder_arch='convnext',
                              aux_decoder_args=dict(num_channels=256, num_layers=3, kernel_size=7))
    aux.decoder.load_state_dict(asd, strict=True)
    aux = aux.to(dev).eval()
    diff = P.GaussianDiffusion(128, k_step=100, backbone_type='wavenet',
                               backbone_args=dict(num_layers=20, num_channels=256, dilation_cycle_length=4), spec_min=smin, spec_max=smax)
    diff.denoise_fn.load_state_dict(wsd, strict=True)
    diff = diff.to(dev).eval()
    g = torch.Generator().manual_seed(77)
    cond = torch.randn((B, T, 256), generator=g)
    noise0 = torch.randn((B, 1, 128, T), generator=g)
    diff._noise_source = lambda shape: noise0.to(dev)
    aux_mel = aux(cond.to(dev), infer=True)
    mel = diff(cond.to(dev), src_spec=aux_mel, infer=True).cpu()
    # oracle chain
    ref_aux = OA.aux_adaptor_forward(asd, acfg, cond, smin, smax, infer=True)
    sch = OS.DiffusionSchedule(1000, 'linear')
    x_start = OS.norm_spec(ref_aux, torch.tensor(-12.), torch.tensor(0.)).transpose(-2, -1)[:, None]
    with torch.no_grad():
        x = OS.gaussian_diffusion_inference(OD.make_denoiser(wsd, wcfg), sch, cond.transpose(1, 2), k_step=100, timesteps=1000,
                                            use_shallow=True, K_step_infer=100, speedup=10, accelerator='ddim', noise0=noise0,
                                            x_start=x_start, step_noise=[])
    ref = OS.denorm_spec(x, torch.tensor(-12.), torch.tensor(0.))
    e_aux = float((aux_mel.cpu().double() - ref_aux.double()).abs().max())
    e_mel = float((mel.double() - ref.double()).abs().max())
    print(dict(test='aux_then_shallow_diffusion', aux_max_abs=e_aux, mel_max_abs=e_mel, ref_absmax=float(ref.abs().max())))
    assert e_aux <= 2e-2 and e_mel <= 2e-2, (e_aux, e_mel)


def test_graphed_launches_policy():
    """_graphs.GraphedLaunches: first call with a key eager, second captures, later calls replay with fresh inputs; a new key does not
    disturb an existing graph; the LRU bound holds; b2s_cuda_graph = False bypasses it."""
    import xiaoicesing_io_b200 as P
    from xiaoicesing_io_b200._graphs import GraphedLaunches
    P.hparams.clear()
    dev = torch.device('cuda:0')
    calls = []

    def fn(inp):
        calls.append(torch.cuda.is_current_stream_capturing())
        return inp[0] * 2 + 1

    gl = GraphedLaunches(max_graphs=2)
    xs = [torch.full((4,), float(i), device=dev) for i in range(5)]
    outs = [gl(('k', 4), [x], fn) for x in xs]
    assert calls == [False, True]                       # eager once, captured once, then replays only
    for i, o in enumerate(outs):
        assert torch.equal(o, xs[i] * 2 + 1)
    outs[2].add_(100)                                   # the caller owns what it gets
    assert torch.equal(gl(('k', 4), [xs[1]], fn), xs[1] * 2 + 1)
    for key in ('a', 'b', 'c'):
        gl((key,), [xs[0]], fn), gl((key,), [xs[0]], fn)
    assert len(gl._graphs) == 2 and ('k', 4) not in gl._graphs
    P.hparams['b2s_cuda_graph'] = False
    n = len(calls)
    gl(('c',), [xs[0]], fn)
    assert len(calls) == n + 1 and calls[-1] is False
    P.hparams.clear()


def test_graph_cache_does_not_thrash_on_ragged_shapes():
    """A caller cycling through more shapes than the graph cache holds (one segment per call over a ragged project) must stop paying
    for captures that are evicted before their first replay; a shape asked for twice in a row is still captured and replayed."""
    import xiaoicesing_io_b200 as P
    from xiaoicesing_io_b200._graphs import GraphedLaunches
    P.hparams.clear()
    g = GraphedLaunches(max_graphs=2)
    calls = []

    def fn(inp):
        calls.append(torch.cuda.is_current_stream_capturing())
        return inp[0] * 2

    x = {n: torch.full((n,), 1.0, device='cuda') for n in range(1, 8)}
    for _ in range(2):                                       # two passes over 7 shapes: first sights, then captures (cache of 2 thrashes)
        for n in range(1, 8):
            assert torch.equal(g((n,), [x[n]], fn), x[n] * 2)
    captures_before = sum(calls)
    assert g._evicted_unused >= 2
    for _ in range(3):                                       # further passes: no capture at all, everything runs from the host
        for n in range(1, 8):
            assert torch.equal(g((n,), [x[n]], fn), x[n] * 2)
    assert sum(calls) == captures_before
    n_calls = len(calls)
    assert (6,) not in g._graphs                            # (the two graphs captured before the guard closed stay cached and replay)
    for _ in range(4):                                       # a steady shape: second sight in a row captures, then replays (fn not called)
        assert torch.equal(g((6,), [x[6]], fn), x[6] * 2)
    assert sum(calls) == captures_before + 1 and len(calls) <= n_calls + 2
    assert g._evicted_unused == 0
