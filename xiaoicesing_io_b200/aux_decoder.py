"""ConvNeXt aux decoder on the B200 kernels: the producer of ``x_start`` for shallow diffusion (the
step BEFORE the sampling loop; reference modules/aux_decoder/convnext.py:58-87 and modules/aux_decoder/__init__.py:28-70).

Same names as the reference (``AUX_DECODERS``, ``build_aux_decoder``, ``ConvNeXtDecoder``, ``AuxDecoderAdaptor``), same
constructor signatures and parameter names, so ``load_state_dict(strict=True)`` of a reference checkpoint works.  The modules
only HOLD parameters; ``forward`` is a sequence of libb2s launches on time-major rows (r = b * T + t):

    cast            condition fp32 -> 16 bit
    b2s_tc_conv1d   inconv as ONE tcgen05 GEMM over the k taps            -> x (fp32 residual stream) + 16-bit copy
    per block:      b2s_lynx_dwconv_h (k = 7, tensor-core Toeplitz form)   -> b2s_layernorm_hh (eps 1e-6)
                    b2s_tc_linear (C -> 4C, erf-GELU epilogue)             -> b2s_tc_linear_residual_scaled (x += gamma * (...))
    b2s_tc_conv1d   outconv; with ``infer`` the adaptor's denorm_spec is folded into its weights and bias

There is no CUDA-core fp32 variant of this path: operands are fp16 (bf16 when ``hparams['b2s_precision'] == 'bf16'``), accumulation
and the residual stream are fp32.  No CPU fallback.
"""
from __future__ import annotations

import inspect

import torch
from torch import nn

from . import _cabi as C
from ._graphs import GraphedLaunches
from .hparams import hparams


class ConvNeXtBlock(nn.Module):
    """Parameter container of one block (convnext.py:8-37): dwconv k=7, norm (eps 1e-6), pwconv1, pwconv2, gamma."""

    def __init__(self, dim: int, intermediate_dim: int, layer_scale_init_value=None, drop_out: float = 0.0):
        super().__init__()
        self.dwconv = nn.Conv1d(dim, dim, kernel_size=7, padding=3, groups=dim)
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, intermediate_dim)
        self.pwconv2 = nn.Linear(intermediate_dim, dim)
        self.gamma = (nn.Parameter(layer_scale_init_value * torch.ones(dim)) if layer_scale_init_value is not None
                      and layer_scale_init_value > 0 else None)


class _ConvNeXtEngine:
    """Packed 16-bit weights + the launch sequence; repacked when a parameter's version / storage changes."""

    def __init__(self, net: 'ConvNeXtDecoder'):
        self.net = net
        self._version = None
        self._graphs = GraphedLaunches()                        # CUDA graphs of the 27-launch sequence, per (weights, B, T)

    def _ver(self):
        ps = list(self.net.parameters())
        return tuple((p._version, p.data_ptr(), p.dtype) for p in ps) + (str(ps[0].device), hparams.get('b2s_precision'))

    def pack(self):
        v = self._ver()
        if v == self._version:
            return
        net = self.net
        dev = net.inconv.weight.device
        if dev.type != 'cuda':
            raise C.B2SError('the aux decoder lives on the CPU; this path has no CPU fallback - move the module to a CUDA device')
        self.bf16 = hparams.get('b2s_precision') == 'bf16'
        hd = C.HALF_DTYPES['bf16' if self.bf16 else 'fp16']
        f = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        h = lambda t: t.to(hd).contiguous()
        Cc, H, k = net.num_channels, net.in_dims, net.kernel_size
        if H % 64 or Cc % 128:
            raise C.B2SError(f'the tensor-core aux decoder needs in_dims % 64 == 0 and num_channels % 128 == 0 (got {H}, {Cc})')
        if net.out_dims % 4:
            raise C.B2SError(f'the tensor-core aux decoder needs out_dims % 4 == 0 (got {net.out_dims})')
        # Conv1d weight [N, Cin, k] -> GEMM operand [N, k * Cin], column = tap * Cin + c
        conv_w = lambda w: w.detach().permute(0, 2, 1).reshape(w.shape[0], -1)
        self.w_in = h(f(conv_w(net.inconv.weight)))
        self.b_in = f(net.inconv.bias)
        self.blocks = []
        for blk in net.conv:
            g = f(blk.gamma) if blk.gamma is not None else None
            self.blocks.append(dict(
                wdw=f(blk.dwconv.weight[:, 0, :].t()), bdw=f(blk.dwconv.bias),          # K-major [7, C]
                ln_g=f(blk.norm.weight), ln_b=f(blk.norm.bias), eps=float(blk.norm.eps),
                w1=h(f(blk.pwconv1.weight)), b1=f(blk.pwconv1.bias), w2=h(f(blk.pwconv2.weight)), b2=f(blk.pwconv2.bias), gamma=g))
        self.w_out_raw = f(conv_w(net.outconv.weight))
        self.b_out_raw = f(net.outconv.bias)
        self.w_out = h(self.w_out_raw)
        self._denorm_key = None
        self.device, self.hd = dev, hd
        self._version = v

    def out_weights(self, scale, shift):
        """outconv operands with ``x * scale + shift`` per output column folded in (denorm_spec, aux_decoder/__init__.py:52-55)."""
        if scale is None:
            return self.w_out, self.b_out_raw
        key = (scale.data_ptr(), shift.data_ptr(), scale._version, shift._version)
        if key != self._denorm_key:
            s = scale.to(self.device, torch.float32).reshape(-1)
            t = shift.to(self.device, torch.float32).reshape(-1)
            self._w_out_dn = (self.w_out_raw * s[:, None]).to(self.hd).contiguous()
            self._b_out_dn = (self.b_out_raw * s + t).contiguous()
            self._denorm_key = key
        return self._w_out_dn, self._b_out_dn

    @torch.no_grad()
    def forward(self, cond: torch.Tensor, scale=None, shift=None) -> torch.Tensor:
        self.pack()
        C.require_cuda(cond, 'condition')
        net, bf, dev, hd = self.net, self.bf16, self.device, self.hd
        B, T, H = cond.shape
        if H != net.in_dims:
            raise C.B2SError(f'condition has {H} channels, the aux decoder expects in_dims={net.in_dims}')
        Cc, k, N = net.num_channels, net.kernel_size, net.out_dims
        rows = B * T
        if rows == 0:
            return torch.empty((B, T, N), device=dev)
        w_out, b_out = self.out_weights(scale, shift)

        def launches(inp):
            cond_, = inp
            out = torch.empty((B, T, N), device=dev)
            c_h = torch.empty((rows, H), device=dev, dtype=hd)
            C.cast_h(cond_, c_h, bf)
            x = torch.empty((rows, Cc), device=dev)
            x_h = torch.empty((rows, Cc), device=dev, dtype=hd)
            d_h = torch.empty((rows, Cc), device=dev, dtype=hd)
            n_h = torch.empty((rows, Cc), device=dev, dtype=hd)
            C.tc_conv1d(c_h, self.w_in, self.b_in, x, Cc, x_h, Cc, B, T, H, Cc, k, C.ACT_NONE, bf)
            g_h = None
            for blk in self.blocks:
                inner = blk['w1'].shape[0]
                if g_h is None or g_h.shape[1] != inner:
                    g_h = torch.empty((rows, inner), device=dev, dtype=hd)
                C.lynx_dwconv_h(x_h, blk['wdw'], blk['bdw'], None, d_h, B, T, Cc, 7, -1, bf)                   # act < 0: none
                C.layernorm_hh(d_h, blk['ln_g'], blk['ln_b'], n_h, rows, Cc, blk['eps'], bf)
                C.tc_linear(n_h, Cc, rows, 0, blk['w1'], Cc, blk['b1'], inner, Cc, bf, act=C.ACT_GELU, out_h=g_h, ldoh=inner)
                C.tc_linear_residual_scaled(g_h, blk['w2'], blk['b2'], blk['gamma'], x, x_h, rows, Cc, inner, bf)
            C.tc_conv1d(x_h, w_out, b_out, out, N, None, 0, B, T, Cc, N, k, C.ACT_NONE, bf)
            return out

        with torch.cuda.device(dev):
            key = (self._version, B, T, w_out.data_ptr(), b_out.data_ptr())
            return self._graphs(key, [cond.float().contiguous()], launches)


class ConvNeXtDecoder(nn.Module):
    """Reference modules/aux_decoder/convnext.py:58-87: ``forward(x [B, T, in_dims], infer) -> [B, T, out_dims]``."""

    def __init__(self, in_dims, out_dims, /, *, num_channels=512, num_layers=6, kernel_size=7, dropout_rate=0.1):
        super().__init__()
        self.in_dims, self.out_dims = in_dims, out_dims
        self.num_channels, self.num_layers, self.kernel_size = num_channels, num_layers, kernel_size
        pad = (kernel_size - 1) // 2
        self.inconv = nn.Conv1d(in_dims, num_channels, kernel_size, stride=1, padding=pad)
        self.conv = nn.ModuleList(ConvNeXtBlock(dim=num_channels, intermediate_dim=num_channels * 4, layer_scale_init_value=1e-6,
                                                drop_out=dropout_rate) for _ in range(num_layers))
        self.outconv = nn.Conv1d(num_channels, out_dims, kernel_size, stride=1, padding=pad)

    def _engine(self) -> _ConvNeXtEngine:
        eng = self.__dict__.get('_b2s_engine')
        if eng is None:
            eng = self.__dict__['_b2s_engine'] = _ConvNeXtEngine(self)
        return eng

    # noinspection PyUnusedLocal
    def forward(self, x, infer=False):
        return self._engine().forward(x)


AUX_DECODERS = {'convnext': ConvNeXtDecoder}


def build_aux_decoder(in_dims: int, out_dims: int, aux_decoder_arch: str, aux_decoder_args: dict) -> nn.Module:
    """aux_decoder/__init__.py:16-22 (filter_kwargs: only the arguments the class accepts)."""
    cls = AUX_DECODERS[aux_decoder_arch]
    accepted = inspect.signature(cls.__init__).parameters
    return cls(in_dims, out_dims, **{k: v for k, v in aux_decoder_args.items() if k in accepted})


class AuxDecoderAdaptor(nn.Module):
    """Reference modules/aux_decoder/__init__.py:28-70: ``forward(condition [B, T, H], infer) -> [B, T, M]`` or ``[B, F, T, M]``
    (a transposed view, like the reference's), de-normalised when ``infer`` - here inside the outconv GEMM."""

    def __init__(self, in_dims: int, out_dims: int, num_feats: int, spec_min: list, spec_max: list, aux_decoder_arch: str,
                 aux_decoder_args: dict):
        super().__init__()
        self.decoder = build_aux_decoder(in_dims=in_dims, out_dims=out_dims * num_feats, aux_decoder_arch=aux_decoder_arch,
                                         aux_decoder_args=aux_decoder_args)
        self.out_dims = out_dims
        self.n_feats = num_feats
        if spec_min is not None and spec_max is not None:
            self.register_buffer('spec_min', torch.FloatTensor(spec_min)[None, None, :].transpose(-3, -2), persistent=False)
            self.register_buffer('spec_max', torch.FloatTensor(spec_max)[None, None, :].transpose(-3, -2), persistent=False)

    def _denorm_columns(self):
        # per output column f * M + m of the decoder: k = (max - min) / 2, b = (max + min) / 2  (:52-55)
        k = ((self.spec_max - self.spec_min) / 2.).reshape(self.n_feats if self.spec_max.dim() == 4 else 1, -1)
        b = ((self.spec_max + self.spec_min) / 2.).reshape(k.shape[0], -1)
        F_, M = self.n_feats, self.out_dims
        return k.expand(F_, M).reshape(-1).contiguous(), b.expand(F_, M).reshape(-1).contiguous()

    def forward(self, condition, infer=False):
        eng = self.decoder._engine()
        if infer:
            if self.__dict__.get('_dn') is None or self._dn[2] != (self.spec_min.data_ptr(), self.spec_min.device):
                k, b = self._denorm_columns()
                self.__dict__['_dn'] = (k, b, (self.spec_min.data_ptr(), self.spec_min.device))
            x = eng.forward(condition, self._dn[0], self._dn[1])
        else:
            x = eng.forward(condition)
        if self.n_feats > 1:
            x = x.reshape(-1, x.shape[1], self.n_feats, self.out_dims).transpose(1, 2)
        return x
