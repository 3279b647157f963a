"""NSF-HiFiGAN vocoder: the HOST logic of xiaoicesing_io_b200/vocoder.py without a GPU - operand packing (zero-padded channels, the
transposed convs as dense convs over u * C phase columns, weight-norm folding), buffer hand-offs and the launch order - by running
the product's launch sequence with torch stand-ins for the C-ABI calls (each stand-in restates what the C entry point computes,
include/b2s.h) and comparing with the unmodified reference's fixtures.  The CUDA kernels themselves are checked by the -m gpu tests."""
import contextlib

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import golden_util as GU
from oracle import vocoder as OV
from test_vocoder import voc_cfg


def _lrelu(x, s):
    return torch.where(x > 0, x, x * s)


def _conv(a, W, B, T, Cin, N, k, dil):
    a3 = a.reshape(B, T, Cin).transpose(1, 2).float()
    w = W.float().reshape(N, k, Cin).permute(0, 2, 1)
    return F.conv1d(a3, w, padding=(k // 2) * dil, dilation=dil).transpose(1, 2).reshape(B * T, N)


@pytest.fixture
def cpu_kernels(monkeypatch):
    import xiaoicesing_io_b200 as P
    from xiaoicesing_io_b200 import _cabi as C
    from xiaoicesing_io_b200 import vocoder as V

    def tc_conv1d_dil(a, W, b, of, ldo, oh, ldoh, B, T, Cin, N, k, dil, act, bf):
        v = _conv(a, W, B, T, Cin, N, k, dil) + b
        v = _lrelu(v, 0.1) if act == C.ACT_LRELU else v
        if of is not None:
            of.view(-1)[:] = v.reshape(-1)
        if oh is not None:
            oh.view(-1)[:] = v.reshape(-1).to(oh.dtype)

    def tc_conv1d_residual(a, W, b, xs, x, yh, sl, B, T, Cin, N, k, dil, bf):
        assert yh is None or yh.data_ptr() != a.data_ptr(), 'y_h aliases the conv input'
        v = (x if xs is None else xs).reshape(B * T, N) + _conv(a, W, B, T, Cin, N, k, dil) + b       # the kernel sees [B * T, N] rows
        x.view(-1)[:] = v.reshape(-1)
        if yh is not None:
            yh.view(-1)[:] = _lrelu(v, sl).to(yh.dtype).reshape(-1)

    def cast_scale_h(i, o, sc, bf):
        o.view(-1)[:] = (i.reshape(-1) * sc).to(o.dtype)

    def voc_phase(f0, ph, B, T, sr, upp, mini):
        s0 = f0 / sr
        last = s0 * upp
        if mini:
            ds0 = F.pad(s0[:, 1:] - s0[:, :-1], (0, 1))
            last = last + 0.5 * ds0 * upp * (upp - 1) / upp
        acc = (torch.fmod(last + 0.5, 1.0) - 0.5).cumsum(1).fmod(1.0)
        ph[:] = F.pad(acc[:, :-1], (1, 0))

    def voc_source(f0, ph, ini, nz, w, b, out, B, T, upp, dim, sr, amp, std, thr):
        n = torch.arange(1, upp + 1)
        s0 = (f0 / sr)[..., None]
        if dim == 0:
            ds0 = F.pad(s0[:, 1:] - s0[:, :-1], (0, 0, 0, 1))
            out[:] = torch.sin(2 * np.pi * (s0 * n + 0.5 * ds0 * n * (n - 1) / upp + ph[..., None])).reshape(B, -1)
            return
        rad = (s0 * n + ph[..., None]).reshape(B, -1, 1) * torch.arange(1, dim + 1).reshape(1, 1, -1)
        i2 = ini.clone()
        i2[0] = 0
        uv = (f0 > thr).float().repeat_interleave(upp, 1)[..., None]
        v = torch.sin(2 * np.pi * (rad + i2)) * amp * uv + (uv * std + (1 - uv) * amp / 3) * nz.reshape(B, -1, dim)
        out[:] = torch.tanh(v @ w + b)

    def voc_source_add(x, lx, src, Wt, b, B, T, Cp, K, st, pad, n_src, sl, bf):
        if K > 0:
            y = F.conv1d(src.reshape(B, 1, n_src), Wt.t()[:, None, :], b, stride=st, padding=pad)
            assert y.shape[-1] == T
            x += y.transpose(1, 2).reshape(B * T, Cp)
        lx[:] = _lrelu(x, sl).to(lx.dtype)

    def _mean(xs):
        s = xs[0].clone()
        for t in xs[1:]:
            s = s + t
        return s / len(xs)

    def voc_avg_act(xs, o, sl, bf):
        o[:] = _lrelu(_mean(xs), sl).to(o.dtype)

    def voc_post(xs, W, b0, wav, B, T, Cc, Cp, k, sl):
        a = _lrelu(_mean(xs), sl).reshape(B, T, Cp)[:, :, :Cc].transpose(1, 2)
        wav[:] = torch.tanh(F.conv1d(a, W.t()[None], b0, padding=k // 2))[:, 0]

    for fn in (tc_conv1d_dil, tc_conv1d_residual, cast_scale_h, voc_phase, voc_source, voc_source_add, voc_avg_act, voc_post):
        monkeypatch.setattr(C, fn.__name__, fn)
    monkeypatch.setattr(C, 'require_cuda', lambda t, name, dtype=torch.float32: t)
    monkeypatch.setattr(C, 'HALF_DTYPES', {'bf16': torch.float32, 'fp16': torch.float32})      # "16-bit" buffers kept exact: logic only
    monkeypatch.setattr(V._VocoderEngine, '_guard', staticmethod(lambda dev: contextlib.nullcontext()))
    monkeypatch.setitem(P.hparams, 'b2s_cuda_graph', False)
    monkeypatch.setitem(P.hparams, 'b2s_voc_streams', False)
    return V


@pytest.mark.parametrize('name', GU.fixture_names('voc_'))
def test_launch_sequence_reproduces_the_reference(cpu_kernels, name):
    fx = GU.Fixture(name)
    gen = cpu_kernels.Generator(dict(fx.meta['h']))
    gen.load_state_dict(fx.sd, strict=True)
    out = gen.forward_rows(fx['mel'].transpose(1, 2).contiguous(), fx['f0'], 1.0, rand_ini=fx['rand_ini'] if 'rand_ini' in fx else None,
                           noise=fx['noise'] if 'noise' in fx else None)
    assert (out - fx['out'][:, 0]).abs().max().item() < 5e-6


def test_launch_sequence_default_geometry(cpu_kernels):
    """The public 44.1 kHz geometry (512 channels, rates 8-8-2-2-2, kernels 3 / 7 / 11): every stage's padding and phase packing."""
    cfg = OV.NsfHifiGanCfg()
    sd = OV.random_state_dict(cfg, 7)
    g = torch.Generator().manual_seed(1)
    B, T = 2, 10
    mel = torch.randn(B, 128, T, generator=g) * 1.5 - 4
    f0 = 110 * 2 ** (2 * torch.rand(B, T, generator=g))
    f0[:, 4:7] = 0
    ri, nz = torch.rand(1, 1, 9, generator=g), torch.randn(B, T * 512, 9, generator=g)
    with torch.no_grad():
        ref = OV.generator_forward(sd, cfg, mel, f0, ri, nz)
    h = dict(num_mels=128, sampling_rate=44100, upsample_rates=[8, 8, 2, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4, 4],
             upsample_initial_channel=512, resblock='1', resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5]] * 3)
    gen = cpu_kernels.Generator(h)
    gen.load_state_dict(sd, strict=True)
    out = gen.forward_rows(mel.transpose(1, 2).contiguous(), f0, rand_ini=ri, noise=nz)
    assert (out - ref[:, 0]).abs().max().item() < 2e-5
    # log10 mels: spec2wav_torch scales by 2.30259 (vocoders/nsf_hifigan.py:60-64)
    import xiaoicesing_io_b200 as P
    P.hparams['mel_base'] = '10'
    try:
        voc = P.NsfHifiGAN(gen)
        y = voc.spec2wav_torch(mel.transpose(1, 2).contiguous(), f0=f0, rand_ini=ri, noise=nz)
        ref10 = OV.spec2wav(sd, cfg, mel.transpose(1, 2), f0, mel_base='10', rand_ini=ri, noise=nz)
    finally:
        P.hparams.pop('mel_base', None)
    assert y.shape == ref10.shape and (y - ref10).abs().max().item() < 2e-5


def test_weight_norm_checkpoints_load(cpu_kernels):
    """A reference checkpoint stores weight_g / weight_v (models.py:225, :233, :248); loading folds them (what load_model +
    remove_weight_norm leave, :29-32)."""
    fx = GU.Fixture('voc_nsf_resblock1')
    sd = {}
    g = torch.Generator().manual_seed(3)
    for k, v in fx.sd.items():
        normed = k.endswith('.weight') and (k.startswith(('conv_pre', 'ups', 'resblocks', 'conv_post')))
        if normed:                                     # g = ||w|| along dim 0 and v parallel to w  =>  g * v / ||v|| = w
            sd[k[:-6] + 'weight_g'] = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, 1, 1)
            sd[k[:-6] + 'weight_v'] = v * (0.5 + torch.rand(v.shape[0], 1, 1, generator=g))
        else:
            sd[k] = v
    gen = cpu_kernels.Generator(dict(fx.meta['h']))
    gen.load_state_dict(sd, strict=True)
    for k, v in fx.sd.items():
        assert torch.allclose(gen.state_dict()[k], v, atol=1e-6), k


def test_cpu_module_raises():
    import xiaoicesing_io_b200 as P
    fx = GU.Fixture('voc_mini_nsf')
    gen = P.vocoder.Generator(dict(fx.meta['h']))
    gen.load_state_dict(fx.sd, strict=True)
    with pytest.raises(P.B2SError):
        gen.forward_rows(fx['mel'].transpose(1, 2).contiguous(), fx['f0'])


GEOMETRIES = {
    'hop256': dict(num_mels=128, sampling_rate=44100, upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4],
                   upsample_initial_channel=256, resblock='1', resblock_kernel_sizes=[3, 7], resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5]]),
    'resblock2_80mel': dict(num_mels=80, sampling_rate=24000, upsample_rates=[8, 8, 4], upsample_kernel_sizes=[16, 16, 8],
                            upsample_initial_channel=128, resblock='2', resblock_kernel_sizes=[3, 5, 7],
                            resblock_dilation_sizes=[[1, 2], [2, 6], [3, 12]]),
    'wide_mini_nsf': dict(num_mels=128, sampling_rate=44100, upsample_rates=[8, 4, 2, 2, 2, 2], upsample_kernel_sizes=[16, 8, 4, 4, 4, 4],
                          upsample_initial_channel=1024, resblock='1', resblock_kernel_sizes=[3, 11],
                          resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5]], mini_nsf=True),
}


@pytest.mark.parametrize('name', sorted(GEOMETRIES))
def test_launch_sequence_other_geometries(cpu_kernels, name):
    """Fold / padding decisions for other hop sizes, ResBlock2 with even dilations, 80 mel bins, a 1024-channel first stage, mini_nsf."""
    h = GEOMETRIES[name]
    cfg = OV.NsfHifiGanCfg(**{k: (tuple(tuple(x) if isinstance(x, list) else x for x in v) if isinstance(v, list) else v) for k, v in h.items()})
    sd = OV.random_state_dict(cfg, 21)
    hop = int(np.prod(h['upsample_rates']))
    g = torch.Generator().manual_seed(31)
    B, T = 2, 7
    mel = torch.randn(B, h['num_mels'], T, generator=g) * 1.5 - 4
    f0 = 110 * 2 ** (2 * torch.rand(B, T, generator=g))
    f0[:, 2:4] = 0
    ri, nz = torch.rand(1, 1, 9, generator=g), torch.randn(B, T * hop, 9, generator=g)
    with torch.no_grad():
        ref = OV.generator_forward(sd, cfg, mel, f0, ri, nz)
    gen = cpu_kernels.Generator(dict(h))
    gen.load_state_dict(sd, strict=True)
    kw = {} if h.get('mini_nsf') else dict(rand_ini=ri, noise=nz)
    out = gen.forward_rows(mel.transpose(1, 2).contiguous(), f0, **kw)
    assert out.shape == (B, T * hop) and (out - ref[:, 0]).abs().max().item() < 3e-5


@pytest.mark.parametrize('k,d,f,C', [(3, 1, 4, 16), (11, 5, 16, 16), (7, 3, 8, 32), (11, 1, 2, 64), (5, 2, 4, 16), (3, 5, 1, 64), (11, 3, 4, 64)])
def test_folded_conv_equals_the_conv(k, d, f, C):
    """The row-folding identity on its own: a k-tap conv with dilation d over [T, C] equals the dense conv with the packed operand over
    [T / f, f * C] (zero padding outside the utterance), for every fold the planner can pick and some it would not."""
    from xiaoicesing_io_b200.vocoder import _fold_conv
    g = torch.Generator().manual_seed(k * 100 + d * 10 + f)
    T = 16 * f + 3 * f
    W, b = torch.randn(C, C, k, generator=g), torch.randn(C, generator=g)
    x = torch.randn(2, T, C, generator=g)
    ref = F.conv1d(x.transpose(1, 2), W, b, padding=(k // 2) * d, dilation=d).transpose(1, 2)                      # [B, T, C]
    Wg, bg, ks, dil = _fold_conv(W, b, d, f, C)
    rows = x.reshape(2, T // f, f * C)
    out = _conv(rows.reshape(-1, f * C), Wg, 2, T // f, f * C, f * C, ks, dil) + bg                                # the GEMM's view
    assert dil == (d if f == 1 else 1) and out.shape == (2 * T // f, f * C)
    assert (out.reshape(2, T, C) - ref).abs().max().item() < 1e-5 * ref.abs().max().item()       # fp32 summation order only


@pytest.mark.parametrize('u,k,f_in,Ci,Co', [(8, 16, 1, 64, 32), (2, 4, 2, 32, 16), (4, 8, 1, 64, 64), (2, 4, 4, 16, 16), (3, 9, 2, 32, 32)])
def test_folded_transposed_conv_equals_conv_transpose(u, k, f_in, Ci, Co):
    """A transposed conv with stride u as a dense conv over input rows of f_in samples whose columns are the f_in * u output phases."""
    from xiaoicesing_io_b200.vocoder import _fold_conv_transpose
    g = torch.Generator().manual_seed(u * 100 + k + f_in)
    T = 12 * f_in
    W, b = torch.randn(Ci, Co, k, generator=g), torch.randn(Co, generator=g)
    x = torch.randn(2, T, Ci, generator=g)
    ref = F.conv_transpose1d(x.transpose(1, 2), W, b, stride=u, padding=(k - u) // 2).transpose(1, 2)               # [B, T * u, Co]
    Wg, bg, ks = _fold_conv_transpose(W, b, u, (k - u) // 2, f_in, Ci, Co)
    out = _conv(x.reshape(-1, f_in * Ci), Wg, 2, T // f_in, f_in * Ci, f_in * u * Co, ks, 1) + bg
    assert ref.shape == (2, T * u, Co)
    assert (out.reshape(2, T * u, Co) - ref).abs().max().item() < 1e-5 * ref.abs().max().item()
