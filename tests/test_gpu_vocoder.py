"""GPU parity of the NSF-HiFiGAN vocoder (mel + f0 -> waveform) against the CPU oracle
(oracle/vocoder.py, pinned by tests/golden/voc_*.npz) through the product's public API and through the C ABI.

Tolerance: the waveform lives in (-1, 1).  fp16 conv operands with fp32 accumulation / residual streams are predicted to cost
~2e-3 (tests/test_vocoder.py::test_operand_rounding_budget): asserted <= 1e-2 ABSOLUTE.  bf16 operands (8 mantissa bits, predicted
2-3e-2) are reported and guarded at 6e-2.  The fp32 pieces (harmonic source, source convs, conv_post) are asserted at 1e-4 / 1e-5."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import golden_util as GU
from oracle import vocoder as OV
from test_vocoder import voc_cfg

pytestmark = pytest.mark.gpu

TOL = {'fp16': 1e-2, 'bf16': 6e-2}
DEFAULT_H = dict(num_mels=128, sampling_rate=44100, upsample_rates=[8, 8, 2, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4, 4],
                 upsample_initial_channel=512, resblock='1', resblock_kernel_sizes=[3, 7, 11],
                 resblock_dilation_sizes=[[1, 3, 5]] * 3)


def _gen(h, sd, precision):
    import xiaoicesing_io_b200 as P
    P.hparams.clear()
    P.hparams.update(b2s_precision=precision)
    gen = P.vocoder.Generator(dict(h))
    gen.load_state_dict(sd, strict=True)
    return gen.cuda().eval()


def _inputs(B, T, M, hop, seed, dim=9):
    g = torch.Generator().manual_seed(seed)
    mel = torch.randn(B, M, T, generator=g) * 1.5 - 4
    f0 = 110 * 2 ** (2 * torch.rand(B, T, generator=g))
    f0[:, T // 3: T // 3 + max(1, T // 5)] = 0
    return mel, f0, torch.rand(1, 1, dim, generator=g), torch.randn(B, T * hop, dim, generator=g)


@pytest.mark.parametrize('precision', ['fp16', 'bf16'])
@pytest.mark.parametrize('name', GU.fixture_names('voc_'))
def test_fixture_against_reference_output(name, precision):
    fx = GU.Fixture(name)
    gen = _gen(fx.meta['h'], fx.sd, precision)
    ri = fx['rand_ini'].cuda() if 'rand_ini' in fx else None
    nz = fx['noise'].cuda() if 'noise' in fx else None
    out = gen(fx['mel'].cuda(), fx['f0'].cuda(), rand_ini=ri, noise=nz)
    assert out.shape == fx['out'].shape and bool(torch.isfinite(out).all())
    err = float((out.cpu() - fx['out']).abs().max())
    print(dict(test='vocoder_fixture', case=name, precision=precision, max_abs=err, ref_absmax=float(fx['out'].abs().max())))
    assert err <= TOL[precision], (name, precision, err)


@pytest.mark.parametrize('precision', ['fp16', 'bf16'])
@pytest.mark.parametrize('B,T', [(2, 37), (1, 150), (3, 1)])
def test_default_geometry_against_oracle(B, T, precision):
    """The public 44.1 kHz model's geometry (512 channels, hop 512, residual kernels 3 / 7 / 11) with fan-in scaled random weights."""
    cfg = OV.NsfHifiGanCfg()
    sd = OV.random_state_dict(cfg, 7)
    mel, f0, ri, nz = _inputs(B, T, 128, 512, 11)
    with torch.no_grad():
        ref = OV.generator_forward(sd, cfg, mel, f0, ri, nz)
    gen = _gen(DEFAULT_H, sd, precision)
    out = gen(mel.cuda(), f0.cuda(), rand_ini=ri.cuda(), noise=nz.cuda())
    assert out.shape == ref.shape == (B, 1, T * 512) and bool(torch.isfinite(out).all())
    err = float((out.cpu() - ref).abs().max())
    print(dict(test='vocoder_default', B=B, T=T, precision=precision, max_abs=err, ref_absmax=float(ref.abs().max()), ref_std=float(ref.std())))
    assert err <= TOL[precision], (B, T, precision, err)
    # the time-major entry with the log10 scale (NsfHifiGAN.spec2wav_torch) gives the same samples as scaling by hand
    import xiaoicesing_io_b200 as P
    P.hparams['mel_base'] = 10
    voc = P.NsfHifiGAN(gen)
    y = voc.spec2wav_torch((mel / 2.30259).transpose(1, 2).contiguous().cuda(), f0=f0.cuda(), rand_ini=ri.cuda(), noise=nz.cuda())
    assert y.shape == (B * T * 512,)
    assert float((y.reshape(B, 1, -1) - out).abs().max()) <= 5e-3


def test_graph_replay_and_seeded_draws():
    """Third call with a shape replays a CUDA graph: same bits as the eager call.  Without rand_ini / noise the product draws them
    with torch.rand / torch.randn in the reference's order (models.py:145, :165) from the device generator."""
    cfg = OV.NsfHifiGanCfg()
    sd = OV.random_state_dict(cfg, 8)
    mel, f0, ri, nz = _inputs(2, 20, 128, 512, 12)
    gen = _gen(DEFAULT_H, sd, 'fp16')
    args = (mel.cuda(), f0.cuda())
    outs = [gen(*args, rand_ini=ri.cuda(), noise=nz.cuda()) for _ in range(4)]
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    torch.manual_seed(123)
    a = gen(*args)
    torch.manual_seed(123)
    ri2 = torch.rand(1, 1, 9, device='cuda')
    nz2 = torch.randn(2, 20 * 512, 9, device='cuda')
    assert torch.equal(a, gen(*args, rand_ini=ri2, noise=nz2))
    # the residual blocks run on parallel streams (forked / joined inside the graph): same bits as one stream
    import xiaoicesing_io_b200 as P
    P.hparams['b2s_voc_streams'] = False
    gen._engine()._graphs.clear()
    assert torch.equal(gen(*args, rand_ini=ri.cuda(), noise=nz.cuda()), outs[0])
    P.hparams.pop('b2s_voc_streams')
    # utterances of a batch are independent: a batch of two copies gives two identical waveforms, equal to the B = 1 run
    m2, f2, n2 = mel[:1].repeat(2, 1, 1).cuda(), f0[:1].repeat(2, 1).cuda(), nz[:1].repeat(2, 1, 1).cuda()
    o2 = gen(m2, f2, rand_ini=ri.cuda(), noise=n2)
    o1 = gen(m2[:1].contiguous(), f2[:1].contiguous(), rand_ini=ri.cuda(), noise=n2[:1].contiguous())
    assert torch.equal(o2[0], o2[1]) and torch.equal(o2[0], o1[0])


def test_empty_batch_and_bad_shapes():
    import xiaoicesing_io_b200 as P
    gen = _gen(DEFAULT_H, OV.random_state_dict(OV.NsfHifiGanCfg(), 9), 'fp16')
    assert gen(torch.zeros(0, 128, 5, device='cuda'), torch.zeros(0, 5, device='cuda')).shape == (0, 1, 5 * 512)
    with pytest.raises(P.B2SError):
        gen(torch.zeros(1, 64, 5, device='cuda'), torch.zeros(1, 5, device='cuda'))
    with pytest.raises(P.B2SError):
        gen(torch.zeros(1, 128, 5, device='cuda'), torch.zeros(1, 6, device='cuda'))


# ---- the pieces, through the C ABI ---------------------------------------------------------------------------------------------
def test_source_kernels_against_oracle():
    from xiaoicesing_io_b200 import _cabi as C
    cfg = OV.NsfHifiGanCfg()
    B, T, upp, dim = 3, 61, 512, 9
    g = torch.Generator().manual_seed(3)
    f0 = 80 * 2 ** (3 * torch.rand(B, T, generator=g))
    f0[0, 10:20] = 0
    f0[2, -1] = 0
    ri, nz = torch.rand(1, 1, dim, generator=g), torch.randn(B, T * upp, dim, generator=g)
    w, b = torch.randn(dim, generator=g), torch.randn(1, generator=g) * 0.1
    ref = OV.source_module({'m_source.l_linear.weight': w[None], 'm_source.l_linear.bias': b}, cfg, f0, upp, ri, nz)[..., 0]
    f0d, phase, out = f0.cuda(), torch.empty(B, T, device='cuda'), torch.empty(B, T * upp, device='cuda')
    C.voc_phase(f0d, phase, B, T, 44100, upp, 0)
    C.voc_source(f0d, phase, ri.reshape(-1).cuda(), nz.cuda(), w.cuda(), b.cuda(), out, B, T, upp, dim, 44100, 0.1, 0.003, 0.)
    err = float((out.cpu() - ref).abs().max())
    print(dict(test='voc_source', max_abs=err))
    assert err <= 2e-4          # the phase prefix sum is a warp scan in fp32 (torch's CPU cumsum accumulates in double): ~1e-5 cycles x 9
    # mini_nsf
    cfg2 = OV.NsfHifiGanCfg(mini_nsf=True)
    ref2 = OV.fast_sine_gen(cfg2, f0)[:, 0]
    upp2, sr2 = 64, 44100 / 8
    out2 = torch.empty(B, T * upp2, device='cuda')
    C.voc_phase(f0d, phase, B, T, sr2, upp2, 1)
    C.voc_source(f0d, phase, None, None, None, None, out2, B, T, upp2, 0, sr2, 0., 0., 0.)
    err2 = float((out2.cpu() - ref2).abs().max())
    print(dict(test='voc_source_mini', max_abs=err2))
    assert err2 <= 2e-4


@pytest.mark.parametrize('precision', ['fp16', 'bf16'])
@pytest.mark.parametrize('Cin,N,k,dil,B,T', [(64, 64, 11, 5, 2, 300), (128, 128, 7, 3, 1, 1000), (256, 256, 3, 1, 3, 130),
                                             (512, 2048, 3, 1, 2, 77), (64, 128, 3, 1, 2, 4100), (128, 512, 7, 1, 1, 9)])
def test_conv1d_dil_and_residual(Cin, N, k, dil, B, T, precision):
    from xiaoicesing_io_b200 import _cabi as C
    bf = precision == 'bf16'
    hd = C.HALF_DTYPES[precision]
    g = torch.Generator().manual_seed(Cin + k)
    a = torch.randn(B, T, Cin, generator=g).to(hd)
    W = (torch.randn(N, Cin, k, generator=g) / np.sqrt(Cin * k)).to(hd)
    bias = torch.randn(N, generator=g) * 0.1
    ref = F.conv1d(a.float().transpose(1, 2).double(), W.double(), bias.double(), padding=(k // 2) * dil, dilation=dil).transpose(1, 2)
    Wg = W.permute(0, 2, 1).reshape(N, k * Cin).contiguous().cuda()
    ad, bd = a.cuda(), bias.cuda()
    out = torch.empty(B * T, N, device='cuda')
    out_h = torch.empty(B * T, N, device='cuda', dtype=hd)
    C.tc_conv1d_dil(ad, Wg, bd, out, N, out_h, N, B, T, Cin, N, k, dil, C.ACT_LRELU, bf)
    exp = F.leaky_relu(ref, 0.1).reshape(B * T, N)
    err = float((out.double().cpu() - exp).abs().max())
    assert err <= 2e-4, err                                   # exact products of 16-bit operands, fp32 accumulation order only
    assert float((out_h.double().cpu() - exp).abs().max()) <= (2e-2 if bf else 3e-3)
    if Cin == N:
        xs = torch.randn(B * T, N, generator=g)
        x = torch.full((B * T, N), 7.0, device='cuda')
        y_h = torch.empty(B * T, N, device='cuda', dtype=hd)
        C.tc_conv1d_residual(ad, Wg, bd, xs.cuda(), x, y_h, 0.1, B, T, Cin, N, k, dil, bf)
        exp2 = xs.double() + ref.reshape(B * T, N)
        assert float((x.double().cpu() - exp2).abs().max()) <= 2e-4
        assert float((y_h.double().cpu() - F.leaky_relu(exp2, 0.1)).abs().max()) <= (4e-2 if bf else 5e-3)
        C.tc_conv1d_residual(ad, Wg, bd, None, x, None, 0.1, B, T, Cin, N, k, dil, bf)        # in place, no copy
        assert float((x.double().cpu() - (exp2 + ref.reshape(B * T, N))).abs().max()) <= 4e-4


def test_elementwise_pieces():
    from xiaoicesing_io_b200 import _cabi as C
    g = torch.Generator().manual_seed(5)
    B, T, Cp, Cr = 2, 1000, 64, 16
    xs = [torch.randn(B * T, Cp, generator=g) for _ in range(3)]
    for t in xs:
        t[:, Cr:] = 0
    xd = [t.cuda() for t in xs]
    mean = (xs[0] + xs[1] + xs[2]) / 3
    out_h = torch.empty(B * T, Cp, device='cuda', dtype=torch.float16)
    C.voc_avg_act(xd, out_h, 0.1, False)
    assert torch.equal(out_h.cpu(), F.leaky_relu(mean, 0.1).half())
    # conv_post + tanh
    W, b0 = torch.randn(1, Cr, 7, generator=g) * 0.2, torch.randn(1, generator=g) * 0.1
    ref = torch.tanh(F.conv1d(F.leaky_relu(mean.reshape(B, T, Cp)[:, :, :Cr].transpose(1, 2), 0.01), W, b0, padding=3))[:, 0]
    wav = torch.empty(B, T, device='cuda')
    C.voc_post(xd, W[0].t().contiguous().cuda(), b0.cuda(), wav, B, T, Cr, Cp, 7, 0.01)
    assert float((wav.cpu() - ref).abs().max()) <= 1e-5
    # strided source conv (noise_convs) + the 16-bit leaky-ReLU copy
    s, n_src = 8, T * 8
    src = torch.randn(B, n_src, generator=g)
    Wn, bn = torch.randn(Cp, 1, 2 * s, generator=g) * 0.3, torch.randn(Cp, generator=g) * 0.1
    x0 = torch.randn(B * T, Cp, generator=g)
    ref_x = x0 + F.conv1d(src[:, None], Wn, bn, stride=s, padding=s // 2).transpose(1, 2).reshape(B * T, Cp)
    x, lx = x0.cuda(), torch.empty(B * T, Cp, device='cuda', dtype=torch.float16)
    C.voc_source_add(x, lx, src.cuda(), Wn[:, 0].t().contiguous().cuda(), bn.cuda(), B, T, Cp, 2 * s, s, s // 2, n_src, 0.1, False)
    assert float((x.cpu() - ref_x).abs().max()) <= 1e-5
    assert float((lx.float().cpu() - F.leaky_relu(ref_x, 0.1)).abs().max()) <= 4e-3
    # scaled cast
    m = torch.randn(999, generator=g)
    o = torch.empty(999, device='cuda', dtype=torch.bfloat16)
    C.cast_scale_h(m.cuda(), o, 2.30259, True)
    assert torch.equal(o.cpu(), (m * 2.30259).bfloat16())


def test_kernels_stay_inside_their_outputs():
    """Every output buffer of the vocoder's entry points sits between two sentinel regions that must come back untouched (the pool has
    no compute-sanitizer): odd row counts, partial tiles, folded-row shapes."""
    from xiaoicesing_io_b200 import _cabi as C
    guards = []

    def guarded(shape, dtype=torch.float32, pad=4096):
        n = int(np.prod(shape))
        flat = torch.full((pad + n + pad,), 7.0, device='cuda', dtype=dtype)
        guards.append((flat, pad, n))
        return flat[pad:pad + n].view(shape)

    def check(what):
        torch.cuda.synchronize()
        for flat, pad, n in guards:
            assert bool((flat[:pad] == 7).all()) and bool((flat[pad + n:] == 7).all()), what
        guards.clear()

    g = torch.Generator().manual_seed(2)
    for (B, T, Cin, N, k, dil) in [(3, 77, 64, 64, 7, 3), (1, 130, 256, 256, 3, 1), (2, 5, 128, 384, 3, 1), (5, 1, 64, 128, 11, 5)]:
        a = torch.randn(B, T, Cin, generator=g).half().cuda()
        W = (torch.randn(N, k * Cin, generator=g) * 0.05).half().cuda()
        bias = torch.randn(N, generator=g).cuda()
        of, oh = guarded((B * T, N)), guarded((B * T, N), torch.float16)
        C.tc_conv1d_dil(a, W, bias, of, N, oh, N, B, T, Cin, N, k, dil, C.ACT_LRELU, False)
        check(('conv1d_dil', B, T, Cin, N))
        if Cin == N:
            x, yh = guarded((B * T, N)), guarded((B * T, N), torch.float16)
            x.zero_()
            C.tc_conv1d_residual(a, W, bias, None, x, yh, 0.1, B, T, Cin, N, k, dil, False)
            check(('conv1d_residual', B, T, N))
    B, T, Cp = 3, 333, 16
    xs = [torch.randn(B * T, Cp, generator=g).cuda() for _ in range(3)]
    out_h = guarded((B * T, Cp), torch.float16)
    C.voc_avg_act(xs, out_h, 0.1, False)
    wav = guarded((B, T))
    C.voc_post(xs, torch.randn(7, Cp, generator=g).cuda(), torch.zeros(1, device='cuda'), wav, B, T, Cp, Cp, 7, 0.01)
    check('avg_act / post')
    x, lx = guarded((B * T, Cp)), guarded((B * T, Cp), torch.float16)
    x.zero_()
    src = torch.randn(B, T * 2, generator=g).cuda()
    C.voc_source_add(x, lx, src, torch.randn(4, Cp, generator=g).cuda(), torch.zeros(Cp, device='cuda'), B, T, Cp, 4, 2, 1, T * 2, 0.1, False)
    check('source_add')
    f0 = (100 + 200 * torch.rand(B, T, generator=g)).cuda()
    phase, har = guarded((B, T)), guarded((B, T * 8))
    C.voc_phase(f0, phase, B, T, 44100, 8, 0)
    C.voc_source(f0, phase, torch.rand(9, generator=g).cuda(), torch.randn(B, T * 8, 9, generator=g).cuda(), torch.randn(9, generator=g).cuda(),
                 torch.zeros(1, device='cuda'), har, B, T, 8, 9, 44100, 0.1, 0.003, 0.)
    check('phase / source')
    assert bool(torch.isfinite(har).all()) and bool(torch.isfinite(wav).all())


@pytest.mark.parametrize('h', [
    dict(num_mels=128, sampling_rate=44100, upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4], upsample_initial_channel=256,
         resblock='1', resblock_kernel_sizes=[3, 7], resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5]]),                  # hop 256, two blocks
    dict(num_mels=80, sampling_rate=24000, upsample_rates=[8, 8, 4], upsample_kernel_sizes=[16, 16, 8], upsample_initial_channel=128,
         resblock='2', resblock_kernel_sizes=[3, 5, 7], resblock_dilation_sizes=[[1, 2], [2, 6], [3, 12]]),            # ResBlock2, 80 mels
    dict(num_mels=128, sampling_rate=44100, upsample_rates=[8, 4, 2, 2, 2, 2], upsample_kernel_sizes=[16, 8, 4, 4, 4, 4],
         upsample_initial_channel=1024, resblock='1', resblock_kernel_sizes=[3, 11], resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5]],
         mini_nsf=True),                                                                                               # 1024 channels, mini_nsf
], ids=['hop256', 'resblock2_80mel', 'wide_mini_nsf'])
def test_other_geometries_against_oracle(h):
    """Geometries other than the public 44.1 kHz one: another hop size, ResBlock2 with even dilations and 80 mel bins (padded to 128
    GEMM columns), a 1024-channel first stage (two N tiles) with the mini_nsf source - every fold / padding decision of vocoder.py."""
    cfg = OV.NsfHifiGanCfg(**{k: (tuple(tuple(x) if isinstance(x, list) else x for x in v) if isinstance(v, list) else v)
                              for k, v in h.items()})
    sd = OV.random_state_dict(cfg, 21)
    hop = int(np.prod(h['upsample_rates']))
    B, T = 2, 29
    mel, f0, ri, nz = _inputs(B, T, h['num_mels'], hop, 31)
    with torch.no_grad():
        ref = OV.generator_forward(sd, cfg, mel, f0, ri, nz)
    gen = _gen(h, sd, 'fp16')
    kw = {} if h.get('mini_nsf') else dict(rand_ini=ri.cuda(), noise=nz.cuda())
    out = gen(mel.cuda(), f0.cuda(), **kw)
    err = float((out.cpu() - ref).abs().max())
    print(dict(test='vocoder_geometry', upsample_rates=h['upsample_rates'], max_abs=err, ref_absmax=float(ref.abs().max())))
    assert out.shape == ref.shape == (B, 1, T * hop) and err <= TOL['fp16'], err
