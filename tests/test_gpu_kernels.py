"""GPU unit tests of individual libb2s entry points against plain torch expressions of the same op
(edge cases: empty input, ragged sizes, aliasing, non-multiple-of-tile shapes)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def C():
    assert torch.cuda.is_available()
    from xiaoicesing_io_b200 import _cabi
    return _cabi


def test_transpose_ragged(C):
    for (b, r, c) in [(1, 1, 1), (3, 37, 129), (2, 128, 690), (5, 33, 31)]:
        x = torch.randn(b, r, c, device='cuda')
        y = torch.empty(b, c, r, device='cuda')
        C.transpose(x, y, b, r, c)
        assert torch.equal(y, x.transpose(1, 2).contiguous())


def test_transpose_empty(C):
    x = torch.empty(0, 4, 4, device='cuda')
    C.transpose(x, x, 0, 4, 4)


@pytest.mark.parametrize('n', [1, 3, 4, 1023, 4096 * 17 + 2])
def test_lincomb(C, n):
    srcs = [torch.randn(n + 8, device='cuda')[:n] for _ in range(5)]   # views keep 16B alignment of the base
    srcs = [s.clone() for s in srcs]
    coef = torch.tensor([0.5, -1.25, 3.0, 1e-3, 7.0], device='cuda')
    want = sum(c * s for c, s in zip(coef.tolist(), [s.double() for s in srcs]))
    dst = torch.empty(n, device='cuda')
    C.lincomb(dst, srcs, coef)
    assert float((dst.double() - want).abs().max()) < 1e-5
    # aliasing: dst is also a source
    keep = srcs[0].clone()
    C.lincomb(srcs[0], srcs, coef)
    assert torch.equal(srcs[0], dst)
    srcs[0].copy_(keep)


def test_lincomb_rejects_too_many_sources(C):
    x = torch.zeros(16, device='cuda')
    with pytest.raises(C.B2SError):
        C.lincomb(x, [x] * 9, torch.zeros(9, device='cuda'))


def test_sinusoid_matches_reference_formula(C):
    import math
    t = torch.tensor([0.0, 3.0, 949.05, 999.0, 437.25], device='cuda')
    for dim in (32, 256, 1024):
        out = torch.empty(t.numel(), dim, device='cuda')
        C.sinusoid(t, out, t.numel(), dim)
        half = dim // 2
        emb = math.log(10000) / (half - 1)
        freq = torch.exp(torch.arange(half) * -emb)
        arg = t.cpu()[:, None] * freq[None, :]
        want = torch.cat((arg.sin(), arg.cos()), -1)
        # freq = exp(.) differs by 1 fp32 ulp between libm implementations; times t ~ 1e3 that is ~6e-5 in phase
        assert float((out.cpu() - want).abs().max()) < 1e-4


@pytest.mark.parametrize('M,N,K', [(1, 4, 4), (130, 36, 24), (257, 512, 256), (20, 1024, 256), (1000, 48, 192)])
def test_linear(C, M, N, K):
    A = torch.randn(M, K, device='cuda')
    W = torch.randn(N, K, device='cuda') / K ** 0.5
    b = torch.randn(N, device='cuda')
    out = torch.empty(M, N, device='cuda')
    for act, fn in ((C.ACT_NONE, lambda v: v), (C.ACT_RELU, torch.relu),
                    (C.ACT_MISH, torch.nn.functional.mish), (C.ACT_GELU, torch.nn.functional.gelu)):
        C.linear(A, K, W, K, b, out, N, M, N, K, alpha=0.5, act=act)
        want = fn(0.5 * (A.double() @ W.double().t()) + b.double())
        assert float((out.double() - want).abs().max()) < 2e-5


def test_gate_edges_and_dilations(C):
    """Zero padding is applied to y = x + step embedding, per utterance."""
    import torch.nn.functional as F
    B, T, Cc = 3, 45, 32
    for d in (1, 2, 4, 8, 16):
        y = torch.randn(B, T, Cc, device='cuda')
        Wref = torch.randn(2 * Cc, Cc, 3, device='cuda') / (3 * Cc) ** 0.5
        cond = torch.randn(B * T, 2 * Cc + 8, device='cuda')          # ld_cond > 2C on purpose
        perm = torch.stack([torch.arange(Cc), torch.arange(Cc) + Cc], 1).reshape(-1).cuda()
        Wd = Wref[perm].permute(0, 2, 1).reshape(2 * Cc, 3 * Cc).contiguous()
        z = torch.empty(B * T, Cc, device='cuda')
        C.wavenet_gate(y, Wd, cond, 2 * Cc + 8, z, B, T, Cc, d)
        conv = F.conv1d(y.transpose(1, 2).double(), Wref.double(), padding=d, dilation=d)     # [B, 2C, T]
        c = cond[:, :2 * Cc].reshape(B, T, Cc, 2).double()                                     # interleaved g/f
        g = conv[:, :Cc].transpose(1, 2) + c[..., 0]
        f = conv[:, Cc:].transpose(1, 2) + c[..., 1]
        want = torch.sigmoid(g) * torch.tanh(f)
        assert float((z.reshape(B, T, Cc).double() - want).abs().max()) < 1e-5, d
