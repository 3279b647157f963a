"""NSF-HiFiGAN vocoder on the B200 kernels: mel + f0 -> waveform, the step AFTER the sampling loop (SURVEY.md section 8 row f-1;
reference modules/vocoders/nsf_hifigan.py:16-104 and modules/nsf_hifigan/models.py:18-299).

Same names as the reference (``Generator``, ``ResBlock1``, ``ResBlock2``, ``SineGen``, ``SourceModuleHnNSF``, ``load_model``,
``NsfHifiGAN`` with ``spec2wav_torch`` / ``spec2wav``), the same config keys (the vocoder's ``config.json``) and the same parameter
names, so a reference checkpoint loads with ``strict=True`` - in the weight-norm form it is stored in (``weight_g`` / ``weight_v``,
folded at load time: what ``load_model`` + ``remove_weight_norm`` leave, models.py:29-32) or in the plain form.  The modules only HOLD
parameters; ``forward`` is a sequence of libb2s launches on TIME-MAJOR rows (r = b * T_i + t at the stage's sample rate) with the
rows folded so that every GEMM row is at least 64 wide (see "folded rows" below; widths that cannot fold are zero-padded to a
multiple of 64 - a padded channel has zero weights and a zero bias, so it is exactly 0 everywhere):

    cast (x 2.30259 for log10 mels)     b2s_cast_scale_f32_h
    conv_pre + leaky_relu               b2s_tc_conv1d_dil (7 taps, ONE tcgen05 GEMM, lrelu epilogue, 16-bit out)
    harmonic source                     b2s_voc_phase + b2s_voc_source (SineGen + tanh(linear), or fastsinegen for mini_nsf)
    per upsampling stage:
      ConvTranspose1d(stride u)         b2s_tc_conv1d_dil: a transposed conv with stride u IS a small dense conv over the INPUT
                                        frames whose N = u * C output columns are the u phases - row [t_in, u * C] of the GEMM output
                                        is rows t_in * u .. t_in * u + u - 1 of the upsampled [T * u, C] stream, no scatter, no zeros
      + noise_conv(source), lrelu copy  b2s_voc_source_add
      residual blocks                   b2s_tc_conv1d_dil (dilated conv, lrelu epilogue, 16-bit out) -> b2s_tc_conv1d_residual
                                        (x_j <- x_j + conv + b on the fp32 stream, 16-bit leaky_relu(x_j) for the next pair)
      mean of the blocks + lrelu        b2s_voc_avg_act (16-bit input of the next transposed conv)
    conv_post + tanh                    b2s_voc_post (fp32, reads the blocks' streams directly)

Operands of the dense convolutions are fp16 (bf16 when ``hparams['b2s_precision'] == 'bf16'``); accumulation, the residual streams,
the source module, conv_post and tanh are fp32.  No CPU fallback.

Batches: every utterance of a call has the same number of frames (the reference vocodes one segment per call,
inference/ds_acoustic.py:227-236); convolutions never read across utterances.
"""
from __future__ import annotations

import contextlib
import json
import pathlib

import numpy as np
import torch
from torch import nn

from . import _cabi as C
from ._graphs import GraphedLaunches
from .hparams import hparams

LRELU_SLOPE = 0.1                                               # models.py:15


class AttrDict(dict):
    """modules/nsf_hifigan/env.py: attribute access, missing keys read as None."""

    def __getattr__(self, name):
        return self.get(name)

    __setattr__ = dict.__setitem__


def get_padding(kernel_size, dilation=1):                       # modules/nsf_hifigan/utils.py:12-13
    return int((kernel_size * dilation - dilation) / 2)


def _pad64(c: int) -> int:
    return (c + 63) // 64 * 64


class ResBlock1(nn.Module):
    """Parameter container (models.py:36-58): convs1 (dilated) and convs2 (dilation 1), ``len(dilation)`` pairs."""

    def __init__(self, h, channels, kernel_size=3, dilation=(1, 3, 5)):
        super().__init__()
        self.h, self.kernel_size, self.dilation = h, kernel_size, tuple(dilation)
        self.convs1 = nn.ModuleList(nn.Conv1d(channels, channels, kernel_size, 1, dilation=d, padding=get_padding(kernel_size, d))
                                    for d in dilation)
        self.convs2 = nn.ModuleList(nn.Conv1d(channels, channels, kernel_size, 1, dilation=1, padding=get_padding(kernel_size, 1))
                                    for _ in dilation)

    def remove_weight_norm(self):
        pass


class ResBlock2(nn.Module):
    """Parameter container (models.py:71-88)."""

    def __init__(self, h, channels, kernel_size=3, dilation=(1, 3)):
        super().__init__()
        self.h, self.kernel_size, self.dilation = h, kernel_size, tuple(dilation)
        self.convs = nn.ModuleList(nn.Conv1d(channels, channels, kernel_size, 1, dilation=d, padding=get_padding(kernel_size, d))
                                   for d in dilation)

    def remove_weight_norm(self):
        pass


class SineGen(nn.Module):
    """models.py:104-132: no parameters; the arithmetic runs in b2s_voc_phase / b2s_voc_source."""

    def __init__(self, samp_rate, harmonic_num=0, sine_amp=0.1, noise_std=0.003, voiced_threshold=0):
        super().__init__()
        self.sine_amp, self.noise_std, self.harmonic_num = sine_amp, noise_std, harmonic_num
        self.dim = harmonic_num + 1
        self.sampling_rate, self.voiced_threshold = samp_rate, voiced_threshold


class SourceModuleHnNSF(nn.Module):
    """models.py:176-195."""

    def __init__(self, sampling_rate, harmonic_num=0, sine_amp=0.1, add_noise_std=0.003, voiced_threshold=0):
        super().__init__()
        self.sine_amp, self.noise_std = sine_amp, add_noise_std
        self.l_sin_gen = SineGen(sampling_rate, harmonic_num, sine_amp, add_noise_std, voiced_threshold)
        self.l_linear = nn.Linear(harmonic_num + 1, 1)
        self.l_tanh = nn.Tanh()


def _fold_weight_norm(state_dict, prefix):
    """``weight_g`` / ``weight_v`` (torch.nn.utils.weight_norm, dim 0) or ``parametrizations.weight.original0/1`` -> ``weight``."""
    for key in [k for k in state_dict if k.startswith(prefix)]:
        for g_name, v_name in (('weight_g', 'weight_v'), ('parametrizations.weight.original0', 'parametrizations.weight.original1')):
            if key.endswith('.' + g_name):
                base = key[:-len(g_name)]
                g, v = state_dict.pop(key), state_dict.pop(base + v_name)
                norm = v.float().reshape(v.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (v.dim() - 1)))
                state_dict[base + 'weight'] = (v.float() * (g.float() / norm)).to(v.dtype)


# ---- folded rows ---------------------------------------------------------------------------------------------------------------------
# The late stages are narrow (64, 32, 16 channels at 1/4 .. 1/1 of the waveform rate): as GEMMs over rows of C channels they would
# need zero-padded channels (4x the HBM traffic and 4x the MMAs at C = 16).  A time-major stream [T, C] is the SAME memory as
# [T / f, f * C]: f consecutive samples form one row of f * C "super channels", and a k-tap conv with dilation d over samples is a
# dense conv over super rows, out[(q, r), co] = sum_{tap, ci} W[co, ci, tap] x[(q + delta, r'), ci] with f * delta + r' = r + (tap - k//2) * d,
# whose weight matrix [f * C, taps' * f * C] holds the original taps at the matching (r, r', delta) and zeros elsewhere.  The MMA count
# per sample falls with f (taps' ~ 2 * ceil((k//2) * d / f) + 1 instead of k at f times the rows per MMA), nothing is padded, and
# at f * C = 256 the MMAs run at full width.  A transposed conv with stride u is the same construction with u output phases per input
# sample.  Packing is host-side only; the kernels see plain dense convs.

def _fold_choices(lay: int, R: int):
    """Folds f (samples per GEMM row) a [T * R, lay] stream supports for EVERY T: f | R, f * lay a multiple of 64, at most 256 wide
    (f = 1: any width that is a multiple of 64)."""
    return [f_ for f_ in (1, 2, 4, 8, 16, 32, 64) if R % f_ == 0 and (f_ * lay) % 64 == 0 and (f_ == 1 or f_ * lay <= 256)]


def _fold_taps(k: int, d: int, f_: int):
    """Taps of the dense conv over rows of f samples (f = 1: the k dilated taps themselves, the GEMM kernel shifts its A tiles by d)."""
    if f_ == 1:
        return k
    c = k // 2
    mr = max(-((-c * d) // f_), (f_ - 1 + c * d) // f_)
    return 2 * mr + 1


def _best_fold(k: int, d: int, lay: int, R: int) -> int:
    """The fold with the fewest tensor-core nanoseconds per sample: taps'(f) K groups of f * lay / 64 blocks x 4 MMAs per 256 rows of
    f samples; an MMA of width N = f * lay costs ~27 + 0.29 N ns (measured: 64 ns at N = 128, 101 ns at N = 256)."""
    best = None
    cap = int(hparams.get('b2s_voc_max_fold', 0) or 0)          # measurement switch: folds above the cap are not considered (0: no cap)
    choices = _fold_choices(lay, R)
    if cap:
        choices = [f_ for f_ in choices if f_ <= cap] or choices[:1]
    for f_ in choices:
        ks = _fold_taps(k, d, f_)
        if ks > 63:
            continue
        n = min(f_ * lay, 256)
        cost = ks * (f_ * lay / 64) * 4 * (27.0 + 0.29 * n) * max(1, f_ * lay // 256) / f_
        if best is None or cost < best[0]:
            best = (cost, f_)
    if best is None:
        raise C.B2SError(f'no GEMM row layout for a conv with k={k}, dilation {d} over {lay}-wide rows')
    return best[1]


def _fold_conv(W, b, d: int, f_: int, lay: int):
    """Conv1d weight [Co, Ci, k] (dilation d, 'same' padding) -> (operand [f * lay, taps' * f * lay], bias [f * lay], taps', the
    dilation left to the GEMM kernel: d at f = 1, else 1)."""
    Co, Ci, k = W.shape
    c, ks = k // 2, _fold_taps(k, d, f_)
    mr = ks // 2
    Wg = torch.zeros((f_, lay, ks, f_, lay), device=W.device, dtype=torch.float32)
    for r in range(f_):
        for tap in range(k):
            s_ = r + (tap - c) * d
            Wg[r, :Co, (s_ // f_ + mr) if f_ > 1 else tap, s_ % f_, :Ci] = W[:, :, tap]
    bg = torch.zeros((f_, lay), device=W.device, dtype=torch.float32)
    bg[:, :Co] = b
    return Wg.reshape(f_ * lay, ks * f_ * lay), bg.reshape(-1).contiguous(), ks, (d if f_ == 1 else 1)


def _fold_conv_transpose(W, b, u: int, pad: int, f_in: int, lay_in: int, lay_out: int):
    """ConvTranspose1d weight [Ci, Co, k] (stride u, padding pad, k - 2 pad = u) over input rows of f_in samples x lay_in channels ->
    (operand [f_in * u * lay_out, taps' * f_in * lay_in], bias, taps'): output row q holds samples (q f_in + a) u + r, a < f_in, r < u;
    out[(a, r), co] = sum x[(q + delta, a'), ci] W[ci, co, (a - a' - f_in delta) u + r + pad]."""
    Ci, Co, k = W.shape
    span = [dl for dl in range(-64, 65) if any(0 <= (a - a2 - f_in * dl) * u + r + pad < k
                                               for a in range(f_in) for a2 in range(f_in) for r in range(u))]
    mr = max(abs(dl) for dl in span)
    ks = 2 * mr + 1
    Wg = torch.zeros((f_in, u, lay_out, ks, f_in, lay_in), device=W.device, dtype=torch.float32)
    for a in range(f_in):
        for r in range(u):
            for j in range(ks):
                for a2 in range(f_in):
                    kk = (a - a2 - f_in * (j - mr)) * u + r + pad
                    if 0 <= kk < k:
                        Wg[a, r, :Co, j, a2, :Ci] = W[:, :, kk].t()
    bg = torch.zeros((f_in, u, lay_out), device=W.device, dtype=torch.float32)
    bg[:, :, :Co] = b
    return Wg.reshape(f_in * u * lay_out, ks * f_in * lay_in), bg.reshape(-1).contiguous(), ks


class Generator(nn.Module):
    """Reference modules/nsf_hifigan/models.py:206-299: ``forward(x [B, num_mels, T], f0 [B, T]) -> [B, 1, T * hop]``.

    ``rand_ini`` / ``noise`` (keyword-only) replace the two random draws of SineGen (models.py:147, :170); when absent they are drawn
    with ``torch.rand(1, 1, dim)`` and ``torch.randn(B, T * hop, dim)`` on the model's device in the reference's order, so a seeded
    call consumes the same generator stream as the reference on the same device."""

    def __init__(self, h):
        super().__init__()
        h = h if isinstance(h, AttrDict) else AttrDict(h)
        self.h = h
        self.num_kernels = len(h.resblock_kernel_sizes)
        self.num_upsamples = len(h.upsample_rates)
        self.mini_nsf = bool(h.mini_nsf)
        if self.mini_nsf:
            self.source_sr = h.sampling_rate / int(np.prod(h.upsample_rates[2:]))
            self.upp = int(np.prod(h.upsample_rates[:2]))
        else:
            self.source_sr = h.sampling_rate
            self.upp = int(np.prod(h.upsample_rates))
            self.m_source = SourceModuleHnNSF(sampling_rate=h.sampling_rate, harmonic_num=8)
            self.noise_convs = nn.ModuleList()
        self.conv_pre = nn.Conv1d(h.num_mels, h.upsample_initial_channel, 7, 1, padding=3)
        self.ups = nn.ModuleList()
        self.resblocks = nn.ModuleList()
        resblock = ResBlock1 if h.resblock == '1' else ResBlock2
        ch = h.upsample_initial_channel
        for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
            ch //= 2
            self.ups.append(nn.ConvTranspose1d(2 * ch, ch, k, u, padding=(k - u) // 2))
            for k2, d in zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes):
                self.resblocks.append(resblock(h, ch, k2, d))
            if not self.mini_nsf:
                if i + 1 < len(h.upsample_rates):
                    stride_f0 = int(np.prod(h.upsample_rates[i + 1:]))
                    self.noise_convs.append(nn.Conv1d(1, ch, kernel_size=stride_f0 * 2, stride=stride_f0, padding=stride_f0 // 2))
                else:
                    self.noise_convs.append(nn.Conv1d(1, ch, kernel_size=1))
            elif i == 1:
                self.source_conv = nn.Conv1d(1, ch, 1)
        self.conv_post = nn.Conv1d(ch, 1, 7, 1, padding=3)
        self._register_load_state_dict_pre_hook(lambda sd, prefix, *a: _fold_weight_norm(sd, prefix))

    def remove_weight_norm(self):
        """models.py:291-299.  Nothing to do: weight norm is folded when a checkpoint is loaded."""

    def _engine(self) -> '_VocoderEngine':
        eng = self.__dict__.get('_b2s_engine')
        if eng is None:
            eng = self.__dict__['_b2s_engine'] = _VocoderEngine(self)
        return eng

    def invalidate(self):
        """Call after writing parameters through detached views (``p.data`` storage swaps are detected, in-place writes are too)."""
        self._engine()._version = None

    def forward(self, x, f0, *, rand_ini=None, noise=None):
        C.require_cuda(x.contiguous(), 'mel')
        B, M, T = x.shape
        rows = torch.empty((B, T, M), device=x.device, dtype=torch.float32)
        if B * T:
            with torch.cuda.device(x.device):
                C.transpose(x.contiguous(), rows, B, M, T)
        return self._engine().forward(rows, f0, 1.0, rand_ini, noise).unsqueeze(1)

    def forward_rows(self, mel, f0, scale: float = 1.0, *, rand_ini=None, noise=None):
        """Time-major entry: ``mel [B, T, num_mels]`` (what the acoustic model produces) x ``scale`` -> ``[B, T * hop]``; no transpose."""
        return self._engine().forward(mel, f0, scale, rand_ini, noise)


class _VocoderEngine:
    """Packed operands + the launch sequence; repacked when a parameter's version / storage changes."""

    def __init__(self, net: Generator):
        self.net = net
        self._version = None
        self._graphs = GraphedLaunches(max_graphs=4)

    def _ver(self):
        ps = list(self.net.parameters())
        return tuple((p._version, p.data_ptr(), p.dtype) for p in ps) + (str(ps[0].device), hparams.get('b2s_precision'),
                                                                           hparams.get('b2s_voc_max_fold'))

    def _side_streams(self, n):
        ss = self.__dict__.setdefault('_streams', [])
        while len(ss) < n:
            ss.append(torch.cuda.Stream(device=self.device))
        return ss[:n]

    @staticmethod
    def _guard(dev):
        """The CUDA device context of every launch; there is no other device to run on."""
        if dev.type != 'cuda':
            raise C.B2SError('the vocoder lives on the CPU; this path has no CPU fallback - move the module to a CUDA device')
        return torch.cuda.device(dev)

    # ---- operand packing -------------------------------------------------------------------------------
    def pack(self):
        v = self._ver()
        if v == self._version:
            return
        net, h = self.net, self.net.h
        dev = net.conv_pre.weight.device
        self._guard(dev)
        self.bf16 = hparams.get('b2s_precision') == 'bf16'
        hd = C.HALF_DTYPES['bf16' if self.bf16 else 'fp16']
        f = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()

        def conv_operand(w, cin_p, n_p):
            """Conv1d weight [N, Cin, k] -> zero-padded 16-bit GEMM operand [n_p, k * cin_p], column = tap * cin_p + c."""
            N, Cin, k = w.shape
            out = torch.zeros((n_p, k, cin_p), device=dev, dtype=torch.float32)
            out[:N, :, :Cin] = f(w).permute(0, 2, 1)
            return out.reshape(n_p, k * cin_p).to(hd).contiguous()

        def padded(b, n_p):
            out = torch.zeros(n_p, device=dev, dtype=torch.float32)
            out[:b.numel()] = f(b).reshape(-1)
            return out

        self.Mp = _pad64(h.num_mels)
        c0 = h.upsample_initial_channel
        self.c0p = _pad64(c0)
        self.w_pre, self.b_pre = conv_operand(net.conv_pre.weight, self.Mp, self.c0p), padded(net.conv_pre.bias, self.c0p)
        self.stages = []
        ch, nk = c0, net.num_kernels
        n_up = len(h.upsample_rates)
        lay_in, R_in = self.c0p, 1                  # channel layout (row width) and samples per mel frame of the stream entering a stage
        for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
            cin = ch
            ch //= 2
            R = R_in * u
            # Row width of this stage's streams: the REAL channel count when rows can be folded into >= 64-wide GEMM rows (see
            # _fold_choices), else zero-padded to a multiple of 64
            lay = ch if ch % 4 == 0 and _fold_choices(ch, R) else _pad64(ch)
            if (k - u) % 2:
                raise C.B2SError(f'upsample kernel {k} / rate {u}: the transposed conv does not produce exactly T * {u} samples '
                                 f'(the reference would fail at x + x_source too)')
            f_in = _fold_choices(lay_in, R_in)[0]
            w_up, b_up, ks_up = _fold_conv_transpose(f(net.ups[i].weight), f(net.ups[i].bias), u, (k - u) // 2, f_in, lay_in, lay)
            st = dict(u=u, f_in=f_in, lay_in=lay_in, ch=ch, lay=lay, ks=ks_up, w_up=w_up.to(hd).contiguous(), b_up=b_up, src=None, blocks=[])
            src_conv = None
            if not net.mini_nsf:
                src_conv = net.noise_convs[i]
                s = int(np.prod(h.upsample_rates[i + 1:])) if i + 1 < n_up else 1
            elif i == 1:
                src_conv, s = net.source_conv, 1        # the source runs at sr / prod(rates[2:]) (models.py:216-217) and so does stage 1
            if src_conv is not None:
                K = src_conv.weight.shape[-1]
                Wt = torch.zeros((K, lay), device=dev, dtype=torch.float32)
                Wt[:, :ch] = f(src_conv.weight)[:, 0, :].t()
                st['src'] = dict(K=K, stride=s if K > 1 else 1, pad=(s // 2) if K > 1 else 0, Wt=Wt.contiguous(),
                                 b=padded(src_conv.bias, lay))

            def folded(conv, d):
                fold = _best_fold(conv.weight.shape[-1], d, lay, R)
                w, b, ks, dil = _fold_conv(f(conv.weight), f(conv.bias), d, fold, lay)
                return dict(w=w.to(hd).contiguous(), b=b, ks=ks, f=fold, dil=dil)

            for j in range(nk):
                blk = net.resblocks[i * nk + j]
                if blk.kernel_size % 2 == 0:
                    raise C.B2SError(f'residual-block kernel size {blk.kernel_size}: only odd sizes keep the length (models.py:39-41)')
                if isinstance(blk, ResBlock1):
                    st['blocks'].append(dict(kind=1, convs=[(folded(c1, d), folded(c2, 1)) for c1, c2, d in zip(blk.convs1, blk.convs2, blk.dilation)]))
                else:
                    st['blocks'].append(dict(kind=2, convs=[folded(c, d) for c, d in zip(blk.convs, blk.dilation)]))
            self.stages.append(st)
            lay_in, R_in = lay, R
        if nk > 4:
            raise C.B2SError(f'{nk} residual blocks per stage: b2s_voc_avg_act / b2s_voc_post take at most 4')
        self.c_last = ch
        self.w_post = f(net.conv_post.weight)[0].t().contiguous()                  # [7, C]
        self.b_post = f(net.conv_post.bias)
        if not net.mini_nsf:
            sg = net.m_source.l_sin_gen
            self.src = dict(dim=sg.dim, amp=sg.sine_amp, std=sg.noise_std, thr=sg.voiced_threshold,
                            w=f(net.m_source.l_linear.weight).reshape(-1), b=f(net.m_source.l_linear.bias))
        self.device, self.hd = dev, hd
        self._version = v

    # ---- launch sequence -------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, mel, f0, scale, rand_ini, noise):
        self.pack()
        net, h, bf, dev, hd = self.net, self.net.h, self.bf16, self.device, self.hd
        mel = C.require_cuda(mel.contiguous(), 'mel')
        f0 = C.require_cuda(f0.contiguous(), 'f0')
        B, T, M = mel.shape
        if M != h.num_mels or tuple(f0.shape) != (B, T):
            raise C.B2SError(f'mel {tuple(mel.shape)} / f0 {tuple(f0.shape)}: expected [B, T, {h.num_mels}] and [B, T]')
        hop = int(np.prod(h.upsample_rates))
        if B * T == 0:
            return torch.empty((B, T * hop), device=dev)
        with self._guard(dev):
            if not net.mini_nsf:
                dim = self.src['dim']
                if rand_ini is None:
                    rand_ini = torch.rand(1, 1, dim, device=dev)                                         # models.py:145
                if noise is None:
                    noise = torch.randn(B, T * net.upp, dim, device=dev)                                 # models.py:165
                rand_ini = C.require_cuda(rand_ini.reshape(-1).contiguous(), 'rand_ini')
                noise = C.require_cuda(noise.contiguous(), 'noise')
                if rand_ini.numel() != dim or noise.numel() != B * T * net.upp * dim:
                    raise C.B2SError(f'rand_ini needs {dim} values and noise [B, T * {net.upp}, {dim}]')
            else:
                rand_ini = noise = None
            key = (self._version, B, T, float(scale), bool(hparams.get('b2s_voc_streams', True)))      # launch-structure switches are part of the key
            return self._graphs(key, [mel, f0, rand_ini, noise], lambda inp: self._launches(inp, B, T, float(scale)))

    def _launches(self, inp, B, T, scale):
        mel, f0, rand_ini, noise = inp
        net, h, bf, dev, hd = self.net, self.net.h, self.bf16, self.device, self.hd
        M, Mp = h.num_mels, self.Mp
        # ---- mel -> 16-bit rows
        if M == Mp:
            mel_h = torch.empty((B * T, Mp), device=dev, dtype=hd)
            C.cast_scale_h(mel, mel_h, scale, bf)
        else:                                                   # toy geometries only (num_mels is 128 in every shipped vocoder): pad
            tmp = torch.empty((B * T, M), device=dev, dtype=hd)
            C.cast_scale_h(mel, tmp, scale, bf)
            mel_h = torch.zeros((B * T, Mp), device=dev, dtype=hd)
            mel_h[:, :M] = tmp
        # ---- harmonic source at the waveform rate (mini_nsf: at sr / prod(rates[2:])).  Nothing before the first stage's source conv
        # needs it: it runs on a side stream next to the mel cast, conv_pre and the first transposed conv, joined where it is read
        upp = net.upp
        phase = torch.empty((B, T), device=dev)
        har = torch.empty((B, T * upp), device=dev)
        use_streams = bool(hparams.get('b2s_voc_streams', True))
        src_branch, src_stream = contextlib.nullcontext(), None
        if use_streams:
            src_stream = self._side_streams(1)[0]
            src_stream.wait_stream(torch.cuda.current_stream())
            src_branch = torch.cuda.stream(src_stream)
        with src_branch:
            C.voc_phase(f0, phase, B, T, net.source_sr, upp, net.mini_nsf)
            if net.mini_nsf:
                C.voc_source(f0, phase, None, None, None, None, har, B, T, upp, 0, net.source_sr, 0., 0., 0.)
            else:
                s = self.src
                C.voc_source(f0, phase, rand_ini, noise, s['w'], s['b'], har, B, T, upp, s['dim'], net.source_sr, s['amp'], s['std'], s['thr'])
        # ---- conv_pre (+ the leaky ReLU in front of the first transposed conv)
        a_h = torch.empty((B * T, self.c0p), device=dev, dtype=hd)
        C.tc_conv1d_dil(mel_h, self.w_pre, self.b_pre, None, 0, a_h, self.c0p, B, T, Mp, self.c0p, 7, 1, C.ACT_LRELU, bf)
        Ti = T
        wav = None
        for i, st in enumerate(self.stages):
            u, lay, fi = st['u'], st['lay'], st['f_in']
            x = torch.empty((B * Ti * u, lay), device=dev)
            n_up = fi * u * lay
            C.tc_conv1d_dil(a_h, st['w_up'], st['b_up'], x, n_up, None, 0, B, Ti // fi, fi * st['lay_in'], n_up, st['ks'], 1, C.ACT_NONE, bf)
            Ti *= u
            rows = B * Ti
            lx_h = torch.empty((rows, lay), device=dev, dtype=hd)
            sc = st['src']
            if src_stream is not None and sc is not None:
                torch.cuda.current_stream().wait_stream(src_stream)       # the harmonic source is complete (first use joins the branch)
                src_stream = None
            if sc is None:
                C.voc_source_add(x, lx_h, None, None, None, B, Ti, lay, 0, 1, 0, 0, LRELU_SLOPE, bf)
            else:
                C.voc_source_add(x, lx_h, har, sc['Wt'], sc['b'], B, Ti, lay, sc['K'], sc['stride'], sc['pad'], T * upp, LRELU_SLOPE, bf)
            def conv_act(inp, cv, out_h):                                      # 16-bit leaky_relu(conv + b)
                w_ = cv['f'] * lay
                C.tc_conv1d_dil(inp, cv['w'], cv['b'], None, 0, out_h, w_, B, Ti // cv['f'], w_, w_, cv['ks'], cv['dil'], C.ACT_LRELU, bf)

            def conv_res(inp, cv, x_src, xj, y_h):                             # xj <- x_src + conv + b, y_h <- 16-bit leaky_relu(xj)
                w_ = cv['f'] * lay
                C.tc_conv1d_residual(inp, cv['w'], cv['b'], x_src, xj, y_h, LRELU_SLOPE, B, Ti // cv['f'], w_, w_, cv['ks'], cv['dil'], bf)

            # The residual blocks of a stage are independent (models.py:279-284): each runs on its own stream, forked from and joined
            # to the caller's (captured as parallel branches of the CUDA graph).  One utterance fills 22 .. 86 CTA pairs per conv, so
            # three convs side by side keep the 74 pairs busy; large batches simply queue.  All buffers are allocated and freed on the
            # caller's stream, around the fork / join.
            nb = len(st['blocks'])
            xs = [torch.empty((rows, lay), device=dev) for _ in range(nb)]
            bufs = [[torch.empty((rows, lay), device=dev, dtype=hd) for _ in range(2 if blk['kind'] == 1 else 3)] for blk in st['blocks']]
            side = self._side_streams(nb - 1) if (nb > 1 and hparams.get('b2s_voc_streams', True)) else []
            if side:
                main = torch.cuda.current_stream()
                fork = torch.cuda.Event()
                fork.record(main)
            for j, blk in enumerate(st['blocks']):
                branch = contextlib.nullcontext()
                if side and j > 0:
                    side[j - 1].wait_event(fork)
                    branch = torch.cuda.stream(side[j - 1])
                with branch:
                    xj, n = xs[j], len(blk['convs'])
                    if blk['kind'] == 1:                                                                  # models.py:60-68
                        t_h, l_a = bufs[j]
                        for m, (c1, c2) in enumerate(blk['convs']):
                            conv_act(lx_h if m == 0 else l_a, c1, t_h)
                            conv_res(t_h, c2, x if m == 0 else None, xj, l_a if m + 1 < n else None)
                    else:                                                                                 # models.py:90-95
                        pp = bufs[j]
                        for m, cv in enumerate(blk['convs']):
                            conv_res(lx_h if m == 0 else pp[(m + 1) & 1], cv, x if m == 0 else None, xj, pp[m & 1] if m + 1 < n else None)
            for stream in side:
                main.wait_stream(stream)
            if i + 1 < len(self.stages):
                a_h = torch.empty((rows, lay), device=dev, dtype=hd)
                C.voc_avg_act(xs, a_h, LRELU_SLOPE, bf)                                                   # :285, :271
            else:
                wav = torch.empty((B, Ti), device=dev)
                C.voc_post(xs, self.w_post, self.b_post, wav, B, Ti, self.c_last, lay, 7, 0.01)         # :286-288 (default slope)
            del x, lx_h, bufs, xs
        if src_stream is not None:
            torch.cuda.current_stream().wait_stream(src_stream)
        return wav


def load_model(model_path: pathlib.Path, device='cuda'):
    """models.py:18-33: ``config.json`` next to the checkpoint, ``cp_dict['generator']``; returns ``(generator, h)``."""
    model_path = pathlib.Path(model_path)
    with open(model_path.with_name('config.json')) as fh:
        h = AttrDict(json.loads(fh.read()))
    generator = Generator(h)
    cp_dict = torch.load(model_path, map_location='cpu')
    generator.load_state_dict(cp_dict['generator'])
    generator.eval()
    generator.remove_weight_norm()
    del cp_dict
    return generator.to(device), h


class NsfHifiGAN:
    """Reference modules/vocoders/nsf_hifigan.py:16-104 (``hparams['vocoder_ckpt']``, ``mel_base``, the mismatch warnings)."""

    def __init__(self, generator: Generator = None, h=None):
        if generator is None:
            model_path = pathlib.Path(hparams['vocoder_ckpt'])
            if not model_path.exists():
                raise FileNotFoundError(f"NSF-HiFiGAN vocoder model is not found at '{model_path}'. "
                                        'Please follow instructions in docs/BestPractices.md#vocoders to get one.')
            print(f'| Load HifiGAN: {model_path}')
            generator, h = load_model(model_path)
        self.model, self.h = generator, (h if h is not None else generator.h)

    @property
    def device(self):
        return next(self.model.parameters()).device

    def to_device(self, device):
        self.model.to(device)

    def get_device(self):
        return self.device

    def _check(self):
        for hk, vk in (('audio_sample_rate', 'sampling_rate'), ('audio_num_mel_bins', 'num_mels'), ('fft_size', 'n_fft'),
                       ('win_size', 'win_size'), ('hop_size', 'hop_size'), ('fmin', 'fmin'), ('fmax', 'fmax')):
            if hk in hparams and self.h.get(vk) is not None and self.h[vk] != hparams[hk]:
                print(f"Mismatch parameters: hparams['{hk}']=", hparams[hk], '!=', self.h[vk], '(vocoder)')

    @staticmethod
    def _scale():
        mel_base = hparams.get('mel_base', 10)
        if mel_base != 'e':
            assert mel_base in [10, '10'], "mel_base must be 'e', '10' or 10."
            return 2.30259                                      # log10 to log mel (nsf_hifigan.py:60-64)
        return 1.0

    def spec2wav_torch(self, mel, **kwargs):                    # mel: [B, T, bins] on the device
        self._check()
        f0 = kwargs.get('f0')
        if f0 is None:
            raise C.B2SError('NSF-HiFiGAN needs f0 (the reference calls Generator.forward(c) without it and fails)')
        return self.model.forward_rows(mel.float(), f0.float(), self._scale(), rand_ini=kwargs.get('rand_ini'),
                                       noise=kwargs.get('noise')).reshape(-1)

    def spec2wav(self, mel, **kwargs):                          # numpy [T, bins] -> numpy [T * hop]
        dev = self.device
        f0 = kwargs.get('f0')
        if f0 is None:
            raise C.B2SError('NSF-HiFiGAN needs f0')
        y = self.spec2wav_torch(torch.as_tensor(mel, dtype=torch.float32, device=dev)[None],
                                f0=torch.as_tensor(f0, dtype=torch.float32, device=dev)[None])
        return y.cpu().numpy()
