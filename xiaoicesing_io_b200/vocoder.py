"""NSF-HiFiGAN vocoder on the B200 kernels: mel + f0 -> waveform, the step AFTER the sampling loop (SURVEY.md section 8 row f-1;
reference modules/vocoders/nsf_hifigan.py:16-104 and modules/nsf_hifigan/models.py:18-299).

Same names as the reference (``Generator``, ``ResBlock1``, ``ResBlock2``, ``SineGen``, ``SourceModuleHnNSF``, ``load_model``,
``NsfHifiGAN`` with ``spec2wav_torch`` / ``spec2wav``), the same config keys (the vocoder's ``config.json``) and the same parameter
names, so a reference checkpoint loads with ``strict=True`` - in the weight-norm form it is stored in (``weight_g`` / ``weight_v``,
folded at load time: what ``load_model`` + ``remove_weight_norm`` leave, models.py:29-32) or in the plain form.  The modules only HOLD
parameters; ``forward`` is a sequence of libb2s launches on TIME-MAJOR rows (r = b * T_i + t at the stage's sample rate) with the
channels zero-padded to a multiple of 64 (a padded channel has zero weights and a zero bias, so it is exactly 0 everywhere):

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
        return tuple((p._version, p.data_ptr(), p.dtype) for p in ps) + (str(ps[0].device), hparams.get('b2s_precision'))

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
        for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
            cin, cin_p = ch, _pad64(ch)
            ch //= 2
            cp = _pad64(ch)
            if (k - u) % 2:
                raise C.B2SError(f'upsample kernel {k} / rate {u}: the transposed conv does not produce exactly T * {u} samples '
                                 f'(the reference would fail at x + x_source too)')
            pad = (k - u) // 2
            # ConvTranspose1d as a dense conv over the input frames: out[q * u + r] = sum_delta x[q + delta] . W[:, :, r + pad - delta * u]
            deltas = [d for d in range(-16, 17) if any(0 <= r + pad - d * u < k for r in range(u))]
            mr = max(abs(d) for d in deltas)
            ks = 2 * mr + 1
            W = f(net.ups[i].weight)                                              # [cin, ch, k]
            Wg = torch.zeros((u, cp, ks, cin_p), device=dev, dtype=torch.float32)
            for r in range(u):
                for j in range(ks):
                    kk = r + pad - (j - mr) * u
                    if 0 <= kk < k:
                        Wg[r, :ch, j, :cin] = W[:, :, kk].t()
            bg = torch.zeros((u, cp), device=dev, dtype=torch.float32)
            bg[:, :ch] = f(net.ups[i].bias)
            st = dict(u=u, cin_p=cin_p, ch=ch, cp=cp, ks=ks, w_up=Wg.reshape(u * cp, ks * cin_p).to(hd).contiguous(),
                      b_up=bg.reshape(-1).contiguous(), src=None, blocks=[])
            src_conv = None
            if not net.mini_nsf:
                src_conv = net.noise_convs[i]
                s = int(np.prod(h.upsample_rates[i + 1:])) if i + 1 < n_up else 1
            elif i == 1:
                src_conv, s = net.source_conv, 1        # the source runs at sr / prod(rates[2:]) (models.py:216-217) and so does stage 1
            if src_conv is not None:
                K = src_conv.weight.shape[-1]
                Wt = torch.zeros((K, cp), device=dev, dtype=torch.float32)
                Wt[:, :ch] = f(src_conv.weight)[:, 0, :].t()
                st['src'] = dict(K=K, stride=s if K > 1 else 1, pad=(s // 2) if K > 1 else 0, Wt=Wt.contiguous(),
                                 b=padded(src_conv.bias, cp))
            for j in range(nk):
                blk = net.resblocks[i * nk + j]
                if isinstance(blk, ResBlock1):
                    pairs = [(conv_operand(c1.weight, cp, cp), padded(c1.bias, cp), conv_operand(c2.weight, cp, cp), padded(c2.bias, cp), d)
                             for c1, c2, d in zip(blk.convs1, blk.convs2, blk.dilation)]
                    st['blocks'].append(dict(kind=1, k=blk.kernel_size, convs=pairs))
                else:
                    convs = [(conv_operand(c.weight, cp, cp), padded(c.bias, cp), d) for c, d in zip(blk.convs, blk.dilation)]
                    st['blocks'].append(dict(kind=2, k=blk.kernel_size, convs=convs))
                if blk.kernel_size % 2 == 0:
                    raise C.B2SError(f'residual-block kernel size {blk.kernel_size}: only odd sizes keep the length (models.py:39-41)')
            self.stages.append(st)
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
            key = (self._version, B, T, float(scale))
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
        # ---- harmonic source at the waveform rate (mini_nsf: at sr / prod(rates[2:]))
        upp = net.upp
        phase = torch.empty((B, T), device=dev)
        C.voc_phase(f0, phase, B, T, net.source_sr, upp, net.mini_nsf)
        har = torch.empty((B, T * upp), device=dev)
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
            u, cp = st['u'], st['cp']
            x = torch.empty((B * Ti * u, cp), device=dev)
            C.tc_conv1d_dil(a_h, st['w_up'], st['b_up'], x, u * cp, None, 0, B, Ti, st['cin_p'], u * cp, st['ks'], 1, C.ACT_NONE, bf)
            Ti *= u
            rows = B * Ti
            lx_h = torch.empty((rows, cp), device=dev, dtype=hd)
            sc = st['src']
            if sc is None:
                C.voc_source_add(x, lx_h, None, None, None, B, Ti, cp, 0, 1, 0, 0, LRELU_SLOPE, bf)
            else:
                C.voc_source_add(x, lx_h, har, sc['Wt'], sc['b'], B, Ti, cp, sc['K'], sc['stride'], sc['pad'], T * upp, LRELU_SLOPE, bf)
            t_h = torch.empty((rows, cp), device=dev, dtype=hd)
            l_a = torch.empty((rows, cp), device=dev, dtype=hd)
            l_b = None
            xs = []
            for blk in st['blocks']:
                xj = torch.empty((rows, cp), device=dev)
                k, n = blk['k'], len(blk['convs'])
                if blk['kind'] == 1:                                                                      # models.py:60-68
                    for m, (w1, b1, w2, b2, d) in enumerate(blk['convs']):
                        C.tc_conv1d_dil(lx_h if m == 0 else l_a, w1, b1, None, 0, t_h, cp, B, Ti, cp, cp, k, d, C.ACT_LRELU, bf)
                        C.tc_conv1d_residual(t_h, w2, b2, x if m == 0 else None, xj, l_a if m + 1 < n else None, LRELU_SLOPE,
                                             B, Ti, cp, cp, k, 1, bf)
                else:                                                                                     # models.py:90-95
                    if l_b is None:
                        l_b = torch.empty((rows, cp), device=dev, dtype=hd)
                    pp = (l_a, l_b)
                    for m, (w, b, d) in enumerate(blk['convs']):
                        C.tc_conv1d_residual(lx_h if m == 0 else pp[(m + 1) & 1], w, b, x if m == 0 else None, xj,
                                             pp[m & 1] if m + 1 < n else None, LRELU_SLOPE, B, Ti, cp, cp, k, d, bf)
                xs.append(xj)
            if i + 1 < len(self.stages):
                a_h = torch.empty((rows, cp), device=dev, dtype=hd)
                C.voc_avg_act(xs, a_h, LRELU_SLOPE, bf)                                                   # :285, :271
            else:
                wav = torch.empty((B, Ti), device=dev)
                C.voc_post(xs, self.w_post, self.b_post, wav, B, Ti, self.c_last, cp, 7, 0.01)          # :286-288 (default slope)
            del x, lx_h, t_h, l_a, l_b, xs
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
