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


@pytest.mark.parametrize('n', [4, 1023, 4096 + 3, 16 * 690 * 128])
@pytest.mark.parametrize('bf16', [True, False])
def test_lincomb_h_equals_lincomb_then_cast(C, n, bf16):
    """The fused update (+ 16-bit copy + flag reset) is bit-identical to the update followed by b2s_cast_f32_h_reset."""
    hd = torch.bfloat16 if bf16 else torch.float16
    srcs = [torch.randn(n, device='cuda') for _ in range(3)]
    coef = torch.tensor([0.75, -1.5, 0.031], device='cuda')
    want = torch.empty(n, device='cuda')
    C.lincomb(want, srcs, coef)
    want_h = torch.empty(n, device='cuda', dtype=hd)
    flags_a = torch.full((37,), 5, device='cuda', dtype=torch.int32)
    C.cast_h(want, want_h, bf16, reset_flags=flags_a)
    got = torch.empty(n, device='cuda')
    got_h = torch.full((n,), 9.0, device='cuda', dtype=hd)
    flags_b = torch.full((37,), 5, device='cuda', dtype=torch.int32)
    C.lincomb_h(got, srcs, coef, got_h, bf16, reset_flags=flags_b)
    assert torch.equal(got, want)
    assert torch.equal(got_h.view(torch.int16), want_h.view(torch.int16))
    assert int(flags_a.abs().sum()) == 0 and int(flags_b.abs().sum()) == 0
    # dst aliases a source (the ancestral update x <- a x + b eps + c z)
    keep = srcs[0].clone()
    C.lincomb_h(srcs[0], srcs, coef, got_h, bf16)
    assert torch.equal(srcs[0], want) and torch.equal(got_h.view(torch.int16), want_h.view(torch.int16))
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


# ------------------------------------------------------------------------------------------------------
# 16-bit tensor-core path (tcgen05 / TMEM / TMA).  Reference = the same op in fp64 on the SAME 16-bit-rounded
# operands, so the only differences are fp32 accumulation order and the 16-bit rounding of the outputs.
# ------------------------------------------------------------------------------------------------------
HALF = [('bf16', torch.bfloat16, 2 ** -8), ('fp16', torch.float16, 2 ** -11)]


@pytest.mark.parametrize('kind,hd,eps', HALF)
@pytest.mark.parametrize('M,N,K', [(1, 32, 64), (130, 48, 48), (257, 512, 256), (1000, 128, 192), (690 * 3, 10240, 256)])
def test_tc_linear(C, kind, hd, eps, M, N, K):
    bf = kind == 'bf16'
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device='cuda').to(hd)
    W = (torch.randn(N, K, device='cuda') / K ** 0.5).to(hd)
    b = torch.randn(N, device='cuda')
    dvec = torch.randn(N, device='cuda')
    want = 0.5 * (A.double() @ W.double().t()) + b.double()
    out = torch.full((M, N), float('nan'), device='cuda')
    out_h = torch.zeros((M, N), device='cuda', dtype=hd) if N % 8 == 0 else None
    y_h = torch.zeros((M, N), device='cuda', dtype=hd) if N % 8 == 0 else None
    C.tc_linear(A, K, M, M, W, K, b, N, K, bf, alpha=0.5, out_f32=out, ldo=N, out_h=out_h, ldoh=N, y_h=y_h, ldy=N,
                dvec=dvec, d_stride=0)
    scale = float(want.abs().max())
    assert float((out.double() - want).abs().max()) < 2e-5 * max(1.0, scale)
    if out_h is not None:
        assert float((out_h.double() - want).abs().max()) < eps * scale
        assert float((y_h.double() - (want + dvec.double())).abs().max()) < eps * (scale + 4)
    # activation variant
    C.tc_linear(A, K, M, M, W, K, b, N, K, bf, alpha=0.5, act=C.ACT_RELU, out_f32=out, ldo=N)
    assert float((out.double() - torch.relu(want)).abs().max()) < 2e-5 * max(1.0, scale)


@pytest.mark.parametrize('kind,hd,eps', HALF)
def test_tc_gate_edges_and_dilations(C, kind, hd, eps):
    import torch.nn.functional as F
    bf = kind == 'bf16'
    for (B, T, Cc) in [(3, 45, 64), (2, 300, 256)]:
        for d in (1, 2, 4, 8, 16):
            torch.manual_seed(d)
            y = torch.randn(B, T, Cc, device='cuda').to(hd)
            Wref = (torch.randn(2 * Cc, Cc, 3, device='cuda') / (3 * Cc) ** 0.5).to(hd)
            cond = torch.randn(B * T, 2 * Cc + 8, device='cuda').to(hd)      # ld_cond > 2C on purpose
            perm = torch.stack([torch.arange(Cc), torch.arange(Cc) + Cc], 1).reshape(-1).cuda()
            Wd = Wref[perm].permute(0, 2, 1).reshape(2 * Cc, 3 * Cc).contiguous()
            z = torch.zeros(B * T, Cc, device='cuda', dtype=hd)
            C.tc_wavenet_gate(y, Wd, cond, 2 * Cc + 8, z, B, T, Cc, d, bf)
            conv = F.conv1d(y.transpose(1, 2).double(), Wref.double(), padding=d, dilation=d)
            c = cond[:, :2 * Cc].reshape(B, T, Cc, 2).double()
            g = conv[:, :Cc].transpose(1, 2) + c[..., 0]
            f = conv[:, Cc:].transpose(1, 2) + c[..., 1]
            want = torch.sigmoid(g) * torch.tanh(f)
            # tanh.approx (2^-11 relative) + 16-bit rounding of z
            assert float((z.reshape(B, T, Cc).double() - want).abs().max()) < 2e-3 + eps, (B, T, Cc, d)


@pytest.mark.parametrize('kind,hd,eps', HALF)
def test_tc_out_resskip(C, kind, hd, eps):
    bf = kind == 'bf16'
    B, T, Cc = 2, 333, 128
    rows = B * T
    torch.manual_seed(3)
    z = torch.randn(rows, Cc, device='cuda').to(hd)
    Wo = (torch.randn(2 * Cc, Cc, device='cuda') / Cc ** 0.5).to(hd)
    bo = torch.randn(2 * Cc, device='cuda')
    x0 = torch.randn(rows, Cc, device='cuda')
    skip0 = torch.randn(rows, Cc, device='cuda')
    dvec = torch.randn(B, Cc, device='cuda')
    o = z.double() @ Wo.double().t() + bo.double()
    x_want = (x0.double() + o[:, :Cc]) / 2 ** 0.5
    y_want = x_want + dvec.double().repeat_interleave(T, 0)
    for first in (True, False):
        x, skip = x0.clone(), skip0.clone()
        y = torch.zeros(rows, Cc, device='cuda', dtype=hd)
        sh = torch.zeros(rows, Cc, device='cuda', dtype=hd)
        C.tc_wavenet_out(z, Wo, bo, x, y, skip, sh, dvec, Cc, first, B, T, Cc, bf)
        s_want = o[:, Cc:] + (0 if first else skip0.double())
        assert float((x.double() - x_want).abs().max()) < 2e-5
        assert float((skip.double() - s_want).abs().max()) < 2e-5
        assert float((y.double() - y_want).abs().max()) < eps * float(y_want.abs().max())
        assert float((sh.double() - s_want).abs().max()) < eps * float(s_want.abs().max())


@pytest.mark.parametrize('kind,hd,eps', HALF)
def test_tc_fused_layer_matches_gate_plus_out(C, kind, hd, eps):
    """The fused per-layer kernel (z kept in shared memory) against the two-kernel tensor-core path and fp64."""
    import torch.nn.functional as F
    bf = kind == 'bf16'
    Cc = 256
    for (B, T, d, first, last) in [(2, 300, 1, True, False), (3, 130, 4, False, False), (1, 690, 8, False, True),
                                   (2, 50, 16, False, False)]:
        rows = B * T
        torch.manual_seed(T + d)
        y = torch.randn(rows, Cc, device='cuda').to(hd)
        Wref = (torch.randn(2 * Cc, Cc, 3, device='cuda') / (3 * Cc) ** 0.5).to(hd)
        perm = torch.stack([torch.arange(Cc), torch.arange(Cc) + Cc], 1).reshape(-1).cuda()
        Wd = Wref[perm].permute(0, 2, 1).reshape(2 * Cc, 3 * Cc).contiguous()
        ldc = 2 * Cc * 3                                              # a 3-layer table, this layer in the middle
        cond_all = torch.randn(rows, ldc, device='cuda').to(hd)
        cond = cond_all[:, 2 * Cc:]
        Wo = (torch.randn(2 * Cc, Cc, device='cuda') / Cc ** 0.5).to(hd)
        bo = torch.randn(2 * Cc, device='cuda')
        x0 = torch.randn(rows, Cc, device='cuda')
        skip0 = torch.randn(rows, Cc, device='cuda')
        dvec = torch.randn(Cc, device='cuda')
        # two-kernel path
        z = torch.zeros(rows, Cc, device='cuda', dtype=hd)
        xa, sa = x0.clone(), skip0.clone()
        ya = torch.zeros(rows, Cc, device='cuda', dtype=hd)
        sha = torch.zeros(rows, Cc, device='cuda', dtype=hd)
        C.tc_wavenet_gate(y, Wd, cond, ldc, z, B, T, Cc, d, bf)
        C.tc_wavenet_out(z, Wo, bo, xa, None if last else ya, sa, sha if last else None, None if last else dvec, 0, first,
                         B, T, Cc, bf)
        # fused
        xb, sb = x0.clone(), skip0.clone()
        yb = torch.zeros(rows, Cc, device='cuda', dtype=hd)
        shb = torch.zeros(rows, Cc, device='cuda', dtype=hd)
        C.tc_wavenet_layer(y, Wd, cond, ldc, Wo, bo, xb, None if last else yb, sb, shb if last else None,
                           None if last else dvec, 0, first, B, T, Cc, d, bf)
        torch.cuda.synchronize()
        tag = (kind, B, T, d)
        # the two paths add the 3 x C conv terms in different orders (the gate GEMM walks channel slabs x taps, the fused kernel
        # taps x channel slabs), so a z value may round to the neighbouring 16-bit number: a few eps in x and skip
        assert float((xa - xb).abs().max()) < 4 * eps, tag
        assert float((sa - sb).abs().max()) < 4 * eps, tag
        assert float((ya.float() - yb.float()).abs().max()) <= eps * 8, tag
        assert float((sha.float() - shb.float()).abs().max()) <= eps * 8, tag
        # fp64 reference on the same 16-bit operands (z rounded to 16 bits in between)
        conv = F.conv1d(y.reshape(B, T, Cc).transpose(1, 2).double(), Wref.double(), padding=d, dilation=d)
        c = cond[:, :2 * Cc].reshape(B, T, Cc, 2).double()
        g = conv[:, :Cc].transpose(1, 2) + c[..., 0]
        f = conv[:, Cc:].transpose(1, 2) + c[..., 1]
        zz = (torch.sigmoid(g) * torch.tanh(f)).reshape(rows, Cc)
        o = zz @ Wo.double().t() + bo.double()
        x_want = (x0.double() + o[:, :Cc]) / 2 ** 0.5
        s_want = o[:, Cc:] + (0 if first else skip0.double())
        assert float((xb.double() - x_want).abs().max()) < 0.02, tag     # z rounding (2^-9) x 256-term dot products
        assert float((sb.double() - s_want).abs().max()) < 0.03, tag


@pytest.mark.parametrize('kind,hd,eps', HALF)
def test_tc_cond_table_layer_major(C, kind, hd, eps):
    bf = kind == 'bf16'
    rows, H, L, N2 = 333, 256, 5, 128
    torch.manual_seed(1)
    cond = torch.randn(rows, H, device='cuda').to(hd)
    W = (torch.randn(L * N2, H, device='cuda') / H ** 0.5).to(hd)
    b = torch.randn(L * N2, device='cuda')
    tab = torch.zeros(L, rows, N2, device='cuda', dtype=hd)
    C.tc_cond_table(cond, rows, W, b, L, N2, H, tab, bf)
    want = (cond.double() @ W.double().t() + b.double()).reshape(rows, L, N2).permute(1, 0, 2)
    assert float((tab.double() - want).abs().max()) < eps * float(want.abs().max())


def _tile_cond(cond, B, T, tpb):
    """[L, B*T, N2] -> the tile/chunk-major layout of b2s_tc_cond_table_tiled (include/b2s.h), built with torch ops."""
    L, rows, N2 = cond.shape
    pad = torch.zeros(L, B, tpb * 128, N2, device=cond.device, dtype=cond.dtype)
    pad[:, :, :T] = cond.reshape(L, B, T, N2)
    v = pad.reshape(L, B, tpb, 128, N2 // 32, 4, 8)              # l, b, tile, row, chunk, piece, elem
    return v.permute(0, 1, 2, 4, 5, 3, 6).contiguous()           # l, b, tile, chunk, piece, row, elem


@pytest.mark.parametrize('kind,hd,eps', HALF)
def test_tc_cond_table_tiled(C, kind, hd, eps):
    bf = kind == 'bf16'
    B, T, H, L, N2 = 3, 300, 256, 4, 512
    tpb = (-(-T // 128) + 1) & ~1
    torch.manual_seed(2)
    cond = torch.randn(B * T, H, device='cuda').to(hd)
    W = (torch.randn(L * N2, H, device='cuda') / H ** 0.5).to(hd)
    b = torch.randn(L * N2, device='cuda')
    tab = torch.zeros(L, B * tpb * 128, N2, device='cuda', dtype=hd)
    C.tc_cond_table_tiled(cond, B, T, W, b, L, N2, H, tab, bf)
    want = (cond.double() @ W.double().t() + b.double()).reshape(B * T, L, N2).permute(1, 0, 2)
    want_t = _tile_cond(want, B, T, tpb).reshape(L, B, tpb, N2 // 32, 4, 128, 8)
    got = tab.reshape(L, B, tpb, N2 // 32, 4, 128, 8).double()
    valid = torch.zeros(tpb * 128, dtype=torch.bool, device='cuda')
    valid[:T] = True
    m = valid.reshape(1, 1, tpb, 1, 1, 128, 1).expand_as(got)
    assert float((got[m] - want_t[m]).abs().max()) < eps * float(want.abs().max())


@pytest.mark.parametrize('kind,hd,eps', HALF)
def test_tc_stack_matches_per_layer_kernels(C, kind, hd, eps):
    """The persistent whole-stack kernel (tile-to-tile hand-off through release/acquire flags) must reproduce the
    sequence of per-layer fused kernels: same operands, same accumulation order -> same bits up to 16-bit y rounding."""
    bf = kind == 'bf16'
    Cc, L = 256, 6
    dil = [1, 2, 4, 8, 16, 1]
    for (B, T) in [(2, 300), (3, 690), (1, 100), (5, 129)]:
        rows = B * T
        torch.manual_seed(B * 1000 + T)
        y0 = torch.randn(rows, Cc, device='cuda').to(hd)
        Wd = (torch.randn(L, 2 * Cc, 3 * Cc, device='cuda') / (3 * Cc) ** 0.5).to(hd)
        Wo = (torch.randn(L, 2 * Cc, Cc, device='cuda') / Cc ** 0.5).to(hd)
        bo = torch.randn(L, 2 * Cc, device='cuda')
        cond = torch.randn(L, rows, 2 * Cc, device='cuda').to(hd)
        x0 = torch.randn(rows, Cc, device='cuda')
        dvec = torch.randn(B, L * Cc, device='cuda')               # per-utterance step embeddings (d_stride != 0)
        for per_row in (False, True):
            ds = L * Cc if per_row else 0
            # per-layer reference
            xa, sa = x0.clone(), torch.zeros(rows, Cc, device='cuda')
            ya, yb = y0.clone(), torch.zeros(rows, Cc, device='cuda', dtype=hd)
            sha = torch.zeros(rows, Cc, device='cuda', dtype=hd)
            for l in range(L):
                last = l == L - 1
                C.tc_wavenet_layer(ya, Wd[l], cond[l], 2 * Cc, Wo[l], bo[l], xa, None if last else yb, sa, sha if last else None,
                                   None if last else dvec[0, (l + 1) * Cc:], ds, l == 0, B, T, Cc, dil[l], bf)
                ya, yb = yb, ya
            # whole stack
            xb, sb = x0.clone(), torch.full((rows, Cc), float('nan'), device='cuda')
            y0b, y1b = y0.clone(), torch.zeros(rows, Cc, device='cuda', dtype=hd)
            shb = torch.zeros(rows, Cc, device='cuda', dtype=hd)
            tpb = (-(-T // 128) + 1) & ~1
            flags = torch.zeros(B * tpb, device='cuda', dtype=torch.int32)
            tiled = _tile_cond(cond, B, T, tpb)
            C.tc_wavenet_stack(y0b, y1b, Wd, tiled, 2 * Cc, B * tpb * 128 * 2 * Cc, Wo, bo, xb, sb, shb, dvec, ds, dil, B, T, Cc,
                               flags, bf)
            torch.cuda.synchronize()
            tag = (kind, B, T, per_row)
            assert int(flags.min()) == L, tag
            assert torch.equal(xa, xb), (tag, float((xa - xb).abs().max()))
            assert torch.equal(sa, sb), (tag, float((sa - sb).abs().max()))
            assert torch.equal(sha, shb), tag


def test_tc_stack_rejects_oversized_grid(C):
    Cc, L, B, T = 256, 2, 40, 690                                    # 240 tiles > 148 SMs
    z = torch.zeros(8, device='cuda')
    with pytest.raises(C.B2SError):
        C.tc_wavenet_stack(z, z[4:], z, z, 2 * Cc, 8, z, z, z, z, None, z, 0, [1, 2], B, T, Cc, z.int(), True)


def _stack3_reference(xin, Win, b_in, Wd, cond, Wres3, bres, dvec, dil, B, T, hd, per_row):
    """fp64 restatement of wavenet.py:86-96 on the SAME 16-bit operands, with y and z rounded to 16 bits where the kernel
    rounds them.  Returns (z of every layer [L, rows, C], final x)."""
    L, C2, _ = Wd.shape
    Cc = C2 // 2
    rows = B * T
    r16 = lambda t: t.to(hd).double()
    x = torch.relu(xin.double() @ Win.double().t() + b_in.double())
    zs = []
    for l in range(L):
        d = (dvec[:, l * Cc:(l + 1) * Cc] if per_row else dvec[:1, l * Cc:(l + 1) * Cc]).double()
        y = r16((x.reshape(B, T, Cc) + d[:, None, :]).float()).reshape(B, T, Cc)
        pre = cond[l].double().reshape(B, T, C2).clone()
        for tap in range(3):
            sh = (tap - 1) * dil[l]
            ys = torch.zeros_like(y)
            if sh < 0:
                ys[:, -sh:] = y[:, :T + sh]
            elif sh > 0:
                ys[:, :T - sh] = y[:, sh:]
            else:
                ys = y
            pre += ys @ Wd[l, :, tap * Cc:(tap + 1) * Cc].double().t()
        z = r16((torch.sigmoid(pre[..., 0::2]) * torch.tanh(pre[..., 1::2])).float()).reshape(rows, Cc)
        zs.append(z)
        w = Wres3[l].double() / 2.0 ** (0.5 * l)
        x = (x + z @ w.t() + bres[l].double()) / 2 ** 0.5
    return torch.stack(zs, 0), x


@pytest.mark.parametrize('kind,hd,eps', HALF)
def test_tc_stack3_against_fp64(C, kind, hd, eps):
    """Third whole-stack kernel (cta_group::2 pairs, resident y tile with halo, residual stream in TMEM, z tiles TMA-stored
    for the deferred skip GEMM): every layer's z against an fp64 restatement on the same 16-bit operands."""
    bf = kind == 'bf16'
    Cc, L, MF = 256, 6, 128
    dil = [1, 2, 4, 8, 16, 1]
    for (B, T) in [(2, 300), (3, 690), (1, 100), (5, 129), (1, 128), (2, 261), (1, 1292)]:
        rows = B * T
        torch.manual_seed(B * 1000 + T)
        xin = torch.randn(rows, MF, device='cuda').to(hd)
        Win = (torch.randn(Cc, MF, device='cuda') / MF ** 0.5).to(hd)
        b_in = torch.randn(Cc, device='cuda') * 0.1
        Wd = (torch.randn(L, 2 * Cc, 3 * Cc, device='cuda') / (3 * Cc) ** 0.5).to(hd)
        Wres = torch.randn(L, Cc, Cc, device='cuda') / Cc ** 0.5
        bres = torch.randn(L, Cc, device='cuda') * 0.1
        sc = 2.0 ** (0.5 * torch.arange(L, device='cuda', dtype=torch.float64))
        Wres3 = (Wres.double() * sc[:, None, None]).float().to(hd)
        b3 = bres.double() * sc[:, None]
        bsum = (torch.cumsum(b3, 0) - b3).float().contiguous()
        cond = torch.randn(L, rows, 2 * Cc, device='cuda').to(hd)
        dvec = torch.randn(B, L * Cc, device='cuda')
        tpb = (-(-T // 128) + 1) & ~1
        tiled = _tile_cond(cond, B, T, tpb)
        for per_row in (False, True):
            ds = L * Cc if per_row else 0
            flags = torch.zeros(B * tpb, device='cuda', dtype=torch.int32)
            ye0 = torch.zeros(rows, Cc, device='cuda', dtype=hd)
            ye1 = torch.zeros(rows, Cc, device='cuda', dtype=hd)
            z_all = torch.full((L, rows, Cc), float('nan'), device='cuda', dtype=hd)
            C.tc_wavenet_stack3(xin, MF, Win, MF, b_in, Wd, tiled, B * tpb * 128 * 2 * Cc, Wres3, bsum, dvec, ds, dil, ye0, ye1,
                                z_all, rows * Cc, B, T, Cc, flags, bf)
            torch.cuda.synchronize()
            z_want, _ = _stack3_reference(xin, Win, b_in, Wd, cond, Wres3, bres, dvec, dil, B, T, hd, per_row)
            tag = (kind, B, T, per_row)
            assert int(flags.max()) <= L, (tag, flags.tolist())     # only tiles at a cluster boundary inside an utterance publish
            assert bool(torch.isfinite(z_all.float()).all()), tag
            errs = [float((z_all[l].double() - z_want[l]).abs().max()) for l in range(L)]
            print('stack3', tag, ['%.1e' % v for v in errs])
            # z in (-1, 1); one 16-bit rounding of y moves a pre-activation by ~eps * |y| * sqrt(3C) * |w| ~ a few eps
            assert max(errs) < 24 * eps, (tag, errs)


@pytest.mark.parametrize('F_', [1, 2])
def test_spec_norm_denorm_fused_with_layout(C, F_):
    """b2s_spec_norm_f32 / b2s_spec_denorm_f32: norm_spec / denorm_spec (ddpm.py:379-383) fused with the layout
    change between [B, T, M] / [B, F, T, M] and the time-major sampler state [B*T, F*M]."""
    B, T, M = 3, 77, 24
    torch.manual_seed(F_)
    lo = (torch.rand(F_, M, device='cuda') * -10 - 2)
    hi = lo + torch.rand(F_, M, device='cuda') * 8 + 1
    spec = torch.randn((B, T, M) if F_ == 1 else (B, F_, T, M), device='cuda') * 5
    state = torch.full((B * T, F_ * M), float('nan'), device='cuda')
    C.spec_norm(spec, lo.reshape(-1).contiguous(), hi.reshape(-1).contiguous(), state, B, F_, T, M)
    l4, h4 = (lo.reshape(1, 1, M), hi.reshape(1, 1, M)) if F_ == 1 else (lo.reshape(1, F_, 1, M), hi.reshape(1, F_, 1, M))
    want = (spec - l4) / (h4 - l4) * 2 - 1                                          # the reference's expression
    want_tm = want.reshape(B * T, M) if F_ == 1 else want.permute(0, 2, 1, 3).reshape(B * T, F_ * M)
    assert torch.equal(state, want_tm)
    back = torch.full_like(spec, float('nan'))
    C.spec_denorm(state, lo.reshape(-1).contiguous(), hi.reshape(-1).contiguous(), back, B, F_, T, M)
    want_back = (want + 1) / 2 * (h4 - l4) + l4
    assert torch.equal(back, want_back)
    C.spec_norm(spec[:0], lo.reshape(-1), hi.reshape(-1), state[:0], 0, F_, T, M)   # empty batch: no launch, no error


@pytest.mark.parametrize('kind,hd,eps', HALF)
@pytest.mark.parametrize('B,T,inner,K', [(3, 130, 192, 31), (2, 19, 64, 31), (1, 704, 2048, 31), (2, 257, 128, 7), (2, 128, 70, 31),
                                         (1, 1, 64, 31), (2, 300, 256, 1)])
def test_lynx_dwconv_h_against_fp64(C, kind, hd, eps, B, T, inner, K):
    """Depthwise conv along time + bias + PReLU on 16-bit activations (lynxnet.py:57-58): the tensor-core Toeplitz kernel
    (inner % 64 == 0; taps rounded to 16 bits like every weight of that path) and the register-window fallback (inner = 70,
    fp32 taps), tiles that end inside an utterance, utterances shorter than the kernel (zero padding on both sides), kernel
    sizes below 31, frames of the NEXT utterance never leaking in."""
    import torch.nn.functional as F
    bf = kind == 'bf16'
    torch.manual_seed(T + inner + K)
    g = torch.randn(B, T, inner, device='cuda').to(hd)
    w = torch.randn(inner, K, device='cuda') / K ** 0.5
    bias = torch.randn(inner, device='cuda')
    slope = torch.rand(inner, device='cuda')
    out = torch.full((B, T, inner), float('nan'), device='cuda').to(hd)
    C.lynx_dwconv_h(g, w.t().contiguous(), bias, slope, out, B, T, inner, K, 0, bf)
    conv = F.conv1d(g.double().transpose(1, 2), w.double()[:, None, :], bias.double(), padding=K // 2, groups=inner)
    want = torch.where(conv >= 0, conv, slope.double()[None, :, None] * conv).transpose(1, 2)
    assert bool(torch.isfinite(out.float()).all())
    err = (out.double() - want).abs()
    mass = F.conv1d(g.double().abs().transpose(1, 2), w.double().abs()[:, None, :], None, padding=K // 2, groups=inner).transpose(1, 2)
    assert float((err - eps * (want.abs() + mass)).max()) < 1e-5, float(err.max())


@pytest.mark.parametrize('M,N,K', [(256 * 10, 512, 320), (256 * 42, 512, 128), (256 * 5 + 77, 1024, 192), (256 * 3, 192, 64), (128, 256, 512)])
def test_tc_linear_last_wave_split_is_bit_identical(C, M, N, K):
    """The generic GEMM cuts the tiles of its last, partial wave (or of a launch smaller than the chip) along N into 2 - 4 narrower
    sub-tiles.  A sub-tile accumulates over K in the same order as the whole tile, so the result must equal -
    bit for bit - the same product computed column block by column block in separate launches (whose own tile counts split differently
    or not at all), and an fp64 reference within rounding.  Shapes: 20 tile pairs (all split by 2), 84 = one full wave + 10 (split by 4),
    ragged rows, a 192-wide N tile, a single tile pair."""
    torch.manual_seed(M + N)
    A = torch.randn(M, K, device='cuda').half()
    W = (torch.randn(N, K, device='cuda') / K ** 0.5).half()
    b = torch.randn(N, device='cuda')
    out = torch.full((M, N), float('nan'), device='cuda')
    out_h = torch.zeros((M, N), device='cuda', dtype=torch.float16)
    C.tc_linear(A, K, M, M, W, K, b, N, K, False, act=C.ACT_GELU, out_f32=out, ldo=N, out_h=out_h, ldoh=N)
    want = torch.nn.functional.gelu(A.double() @ W.double().t() + b.double())
    assert float((out.double() - want).abs().max()) < 2e-5 * max(1.0, float(want.abs().max()))
    for nb in (64, 128):
        if N % nb:
            continue
        parts = torch.full((M, N), float('nan'), device='cuda')
        for n0 in range(0, N, nb):
            C.tc_linear(A, K, M, M, W[n0:n0 + nb].contiguous(), K, b[n0:n0 + nb].contiguous(), nb, K, False, act=C.ACT_GELU,
                        out_f32=parts[:, n0:], ldo=N)
        assert torch.equal(parts, out), (M, N, K, nb)
    assert torch.equal(out_h, out.half())
