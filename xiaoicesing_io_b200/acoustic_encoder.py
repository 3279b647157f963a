"""FastSpeech2 acoustic encoder on the B200 kernels: the producer of ``condition [B, T, H]`` for the sampling path (SURVEY.md
section 8 row f-2, the step BEFORE the path; reference modules/fastspeech/acoustic_encoder.py:14-109, modules/fastspeech/
tts_modules.py:353-428, modules/commons/common_layers.py:120-263).

``FastSpeech2Acoustic(vocab_size)`` reads the same ``hparams`` keys as the reference (hidden_size, enc_layers, enc_ffn_kernel_size,
ffn_act, num_heads, use_pos_embed, use_rope, use_spk_id, num_spk, use_*_embed) and keeps the reference's parameter names, so
``load_state_dict(strict=True)`` of a reference checkpoint works.  Implemented: the rotary-position configuration configs/acoustic.yaml
ships (``use_pos_embed`` and ``use_rope``), GELU feed-forward; anything else raises at construction.

``forward(txt_tokens, mel2ph, f0, key_shift=None, speed=None, spk_embed_id=None, **variances)`` is a sequence of libb2s launches:

    b2s_enc_mel2ph_to_dur, b2s_enc_embed                     x = (sqrt(H) E[token] + dur_embed) * keep        token rows
    per layer:  b2s_layernorm_h -> b2s_tc_linear (QKV, tcgen05) -> b2s_enc_rope -> b2s_enc_attention
                -> b2s_tc_linear_residual (out_proj + residual) -> b2s_enc_mask_rows
                b2s_layernorm_h -> b2s_tc_conv1d (k-tap conv FFN as one GEMM, k^-0.5 folded in, GELU epilogue)
                -> b2s_tc_linear_residual (ffn_2 + residual) -> b2s_enc_mask_rows
    b2s_enc_layernorm_mask                                   padded table [B, L + 1, H]
    b2s_enc_assemble                                         gather by mel2ph + speaker / pitch / variance / key-shift / speed embeds

GEMM operands are fp16 (bf16 under ``hparams['b2s_precision'] == 'bf16'``), the residual stream, LayerNorm statistics, rotary
embedding, attention scores and softmax are fp32.  No CPU fallback.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _cabi as C
from ._graphs import GraphedLaunches
from .hparams import hparams

VARIANCE_NAMES = ('energy', 'breathiness', 'voicing', 'tension')     # the reference's order (acoustic_encoder.py:34-41)


class _RotaryEmbedding(nn.Module):
    """Parameter container of modules/commons/rotary_embedding_torch.py:89-135 (freqs_for='lang', theta 10000): ``freqs`` [dim / 2]."""

    def __init__(self, dim: int, theta: float = 10000.0):
        super().__init__()
        self.freqs = nn.Parameter(1. / (theta ** (torch.arange(0, dim, 2)[:(dim // 2)].float() / dim)), requires_grad=False)


class _SelfAttention(nn.Module):
    def __init__(self, embed_dim, rotary_embed):
        super().__init__()
        self.in_proj = nn.Linear(embed_dim, embed_dim * 3, bias=False)
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=False)
        self.rotary_embed = rotary_embed


class _FFN(nn.Module):
    def __init__(self, hidden, filt, kernel_size):
        super().__init__()
        self.ffn_1 = nn.Conv1d(hidden, filt, kernel_size, padding=kernel_size // 2)
        self.ffn_2 = nn.Linear(filt, hidden)
        nn.init.xavier_uniform_(self.ffn_2.weight)
        nn.init.constant_(self.ffn_2.bias, 0.)


class _EncSALayer(nn.Module):
    def __init__(self, c, kernel_size, rotary_embed):
        super().__init__()
        self.layer_norm1 = nn.LayerNorm(c)
        self.self_attn = _SelfAttention(c, rotary_embed)
        self.layer_norm2 = nn.LayerNorm(c)
        self.ffn = _FFN(c, 4 * c, kernel_size)


class _EncoderLayer(nn.Module):
    def __init__(self, c, kernel_size, rotary_embed):
        super().__init__()
        self.op = _EncSALayer(c, kernel_size, rotary_embed)


class FastSpeech2Encoder(nn.Module):
    """Parameter container of tts_modules.py:353-383 (rotary configuration)."""

    def __init__(self, hidden_size, num_layers, ffn_kernel_size=9, ffn_act='gelu', dropout=None, num_heads=2, use_pos_embed=True,
                 rel_pos=True, use_rope=False):
        super().__init__()
        if not (use_pos_embed and use_rope):
            raise NotImplementedError('the B200 acoustic encoder implements the rotary-position configuration (use_pos_embed and '
                                      'use_rope, configs/acoustic.yaml:71) only')
        if ffn_act != 'gelu':
            raise NotImplementedError(f"the B200 acoustic encoder implements ffn_act='gelu' only (got {ffn_act!r})")
        self.hidden_size, self.num_layers, self.num_heads, self.ffn_kernel_size = hidden_size, num_layers, num_heads, ffn_kernel_size
        rotary = _RotaryEmbedding(hidden_size // num_heads)        # ONE module shared by all layers, like the reference (:362-373)
        self.layers = nn.ModuleList(_EncoderLayer(hidden_size, ffn_kernel_size, rotary) for _ in range(num_layers))
        self.layer_norm = nn.LayerNorm(hidden_size)


def _linear1(h):
    lin = nn.Linear(1, h)
    nn.init.xavier_uniform_(lin.weight)
    nn.init.constant_(lin.bias, 0.)
    return lin


class FastSpeech2Acoustic(nn.Module):
    """Reference modules/fastspeech/acoustic_encoder.py:14-109."""

    def __init__(self, vocab_size):
        super().__init__()
        H = hparams['hidden_size']
        self.hidden_size, self.vocab_size = H, vocab_size
        self.txt_embed = nn.Embedding(vocab_size, H, padding_idx=0)
        nn.init.normal_(self.txt_embed.weight, mean=0, std=H ** -0.5)
        nn.init.constant_(self.txt_embed.weight[0], 0)
        self.dur_embed = _linear1(H)
        self.encoder = FastSpeech2Encoder(hidden_size=H, num_layers=hparams['enc_layers'], ffn_kernel_size=hparams['enc_ffn_kernel_size'],
                                          ffn_act=hparams['ffn_act'], dropout=hparams.get('dropout'), num_heads=hparams['num_heads'],
                                          use_pos_embed=hparams['use_pos_embed'], rel_pos=hparams.get('rel_pos', False),
                                          use_rope=hparams.get('use_rope', False))
        self.pitch_embed = _linear1(H)
        self.variance_embed_list = [n for n in VARIANCE_NAMES if hparams.get(f'use_{n}_embed', False)]
        self.use_variance_embeds = len(self.variance_embed_list) > 0
        if self.use_variance_embeds:
            self.variance_embeds = nn.ModuleDict({n: _linear1(H) for n in self.variance_embed_list})
        self.use_key_shift_embed = hparams.get('use_key_shift_embed', False)
        if self.use_key_shift_embed:
            self.key_shift_embed = _linear1(H)
        self.use_speed_embed = hparams.get('use_speed_embed', False)
        if self.use_speed_embed:
            self.speed_embed = _linear1(H)
        self.use_spk_id = hparams['use_spk_id']
        if self.use_spk_id:
            self.spk_embed = nn.Embedding(hparams['num_spk'], H)
            nn.init.normal_(self.spk_embed.weight, mean=0, std=H ** -0.5)
        self.__dict__['_graphs'] = GraphedLaunches()           # not a sub-module: CUDA graphs of the launch sequence per input shape

    # ------------------------------------------------------------------------------------------------------------------
    def _pack(self):
        ps = list(self.parameters())
        ver = tuple((p._version, p.data_ptr(), p.dtype) for p in ps) + (str(ps[0].device), hparams.get('b2s_precision'))
        st = self.__dict__.get('_b2s_packed')
        if st is not None and st['ver'] == ver:
            return st
        dev = ps[0].device
        if dev.type != 'cuda':
            raise C.B2SError('the acoustic encoder lives on the CPU; this path has no CPU fallback - move the module to a CUDA device')
        H, enc = self.hidden_size, self.encoder
        if H % 64 or (H // enc.num_heads) % 4 or H // enc.num_heads > 128:
            raise C.B2SError(f'the tensor-core acoustic encoder needs hidden_size % 64 == 0 and a head dim that is a multiple of 4, at '
                             f'most 128 (got hidden_size {H}, {enc.num_heads} heads)')
        bf16 = hparams.get('b2s_precision') == 'bf16'
        hd = C.HALF_DTYPES['bf16' if bf16 else 'fp16']
        f = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        h = lambda t: f(t).to(hd).contiguous()
        k = enc.ffn_kernel_size
        layers = []
        for lay in enc.layers:
            op = lay.op
            w1 = op.ffn.ffn_1.weight.detach().permute(0, 2, 1).reshape(op.ffn.ffn_1.weight.shape[0], -1) * k ** -0.5     # [4H, k*H], x k^-0.5
            layers.append(dict(
                ln1=(f(op.layer_norm1.weight), f(op.layer_norm1.bias)), ln2=(f(op.layer_norm2.weight), f(op.layer_norm2.bias)),
                w_in=h(op.self_attn.in_proj.weight), w_out=h(op.self_attn.out_proj.weight), freqs=f(op.self_attn.rotary_embed.freqs),
                w1=h(w1), b1=f(op.ffn.ffn_1.bias * k ** -0.5), w2=h(op.ffn.ffn_2.weight), b2=f(op.ffn.ffn_2.bias)))
        scal = [(self.pitch_embed, None)]
        scal += [(self.variance_embeds[n], n) for n in self.variance_embed_list]
        n_var_first, n_var = 1, len(self.variance_embed_list)
        if self.use_key_shift_embed:
            scal.append((self.key_shift_embed, 'key_shift'))
        if self.use_speed_embed:
            scal.append((self.speed_embed, 'speed'))
        st = dict(ver=ver, dev=dev, bf16=bf16, hd=hd, layers=layers, zero_bias=torch.zeros(H, device=dev),
                  E=f(self.txt_embed.weight), w_dur=f(self.dur_embed.weight[:, 0]), b_dur=f(self.dur_embed.bias),
                  ln=(f(enc.layer_norm.weight), f(enc.layer_norm.bias), float(enc.layer_norm.eps)),
                  scal=[(f(m.weight[:, 0]), f(m.bias), name) for m, name in scal], n_var_first=n_var_first, n_var=n_var,
                  spk=f(self.spk_embed.weight) if self.use_spk_id else None)
        self.__dict__['_b2s_packed'] = st
        return st

    @torch.no_grad()
    def forward(self, txt_tokens, mel2ph, f0, key_shift=None, speed=None, spk_embed_id=None, **kwargs):
        st = self._pack()
        dev, bf, hd = st['dev'], st['bf16'], st['hd']
        for name, t in (('txt_tokens', txt_tokens), ('mel2ph', mel2ph), ('f0', f0)):
            if not t.is_cuda:
                raise C.B2SError(f'{name} must be a CUDA tensor: this path has no CPU fallback (got {t.device})')
        B, L = txt_tokens.shape
        T = mel2ph.shape[1]
        H, enc = self.hidden_size, self.encoder
        nh, k = enc.num_heads, enc.ffn_kernel_size
        rows = B * L
        if B * T == 0:
            return torch.empty((B, T, H), device=dev)
        vals = []
        for w, b_, name in st['scal']:
            if name is None:
                v = f0
            elif name == 'key_shift':
                v = key_shift
            elif name == 'speed':
                v = speed
            else:
                v = kwargs.get(name)
            if v is None:
                raise C.B2SError(f'the acoustic encoder was built with the {name} embedding: pass {name}=[B, T]')
            vals.append(v.to(device=dev, dtype=torch.float32).expand(B, T).contiguous())
        spk = None
        if self.use_spk_id:
            mix = kwargs.get('spk_mix_embed')
            if mix is not None:                                  # [B, 1, H] or per frame [B, T, H] (acoustic_encoder.py:93-97)
                mix = mix.to(device=dev, dtype=torch.float32)
                spk = mix.reshape(B, H).contiguous() if mix.shape[1] == 1 else mix.expand(B, T, H).contiguous()
            else:
                spk = st['spk'][spk_embed_id.to(dev).reshape(-1)].contiguous()                # [B, H] rows of the embedding table

        def launches(inp):
            tok, m2p, spk_, *vals_ = inp
            cond = torch.empty((B, T, H), device=dev)
            dur = torch.empty((B, L), device=dev)
            C.enc_mel2ph_to_dur(m2p, dur, B, T, L)
            table = torch.empty((B, L + 1, H), device=dev)
            x = torch.empty((rows, H), device=dev)
            keep = torch.empty((rows,), device=dev)
            if rows:
                C.enc_embed(tok, dur, st['E'], st['w_dur'], st['b_dur'], x, keep, rows, H, self.vocab_size)
                n_h = torch.empty((rows, H), device=dev, dtype=hd)
                a_h = torch.empty((rows, H), device=dev, dtype=hd)
                qkv = torch.empty((rows, 3 * H), device=dev)
                f_h = torch.empty((rows, 4 * H), device=dev, dtype=hd)
                for lay in st['layers']:
                    C.layernorm_h(x, lay['ln1'][0], lay['ln1'][1], n_h, rows, H, bf)
                    C.tc_linear(n_h, H, rows, 0, lay['w_in'], H, None, 3 * H, H, bf, out_f32=qkv, ldo=3 * H)
                    C.enc_rope(qkv, lay['freqs'], B, L, H, nh)
                    C.enc_attention(qkv, keep, a_h, B, L, H, nh, bf)
                    C.tc_linear_residual(a_h, lay['w_out'], st['zero_bias'], x, rows, H, H, bf)
                    C.enc_mask_rows(x, keep, rows, H)
                    C.layernorm_h(x, lay['ln2'][0], lay['ln2'][1], n_h, rows, H, bf)
                    C.tc_conv1d(n_h, lay['w1'], lay['b1'], None, 0, f_h, 4 * H, B, L, H, 4 * H, k, C.ACT_GELU, bf)
                    C.tc_linear_residual(f_h, lay['w2'], lay['b2'], x, rows, H, 4 * H, bf)
                    C.enc_mask_rows(x, keep, rows, H)
            C.enc_layernorm_mask(x, st['ln'][0], st['ln'][1], keep, table, B, L, H, st['ln'][2])
            C.enc_assemble(table, m2p, spk_, list(vals_), [w for w, _, _ in st['scal']], [b2 for _, b2, _ in st['scal']],
                           st['n_var_first'], st['n_var'], cond, B, T, L, H)
            return cond

        with torch.cuda.device(dev):
            tok = txt_tokens.to(torch.int64).contiguous()
            m2p = mel2ph.to(torch.int64).contiguous()
            key = (st['ver'], B, L, T, None if spk is None else spk.dim())
            return self._graphs(key, [tok, m2p, spk, *vals], launches)
        with torch.cuda.device(dev):
            tok = txt_tokens.to(torch.int64).contiguous()
            m2p = mel2ph.to(torch.int64).contiguous()
            dur = torch.empty((B, L), device=dev)
            C.enc_mel2ph_to_dur(m2p, dur, B, T, L)
            table = torch.empty((B, L + 1, H), device=dev)
            x = torch.empty((rows, H), device=dev)
            keep = torch.empty((rows,), device=dev)
            if rows:
                C.enc_embed(tok, dur, st['E'], st['w_dur'], st['b_dur'], x, keep, rows, H, self.vocab_size)
                n_h = torch.empty((rows, H), device=dev, dtype=hd)
                a_h = torch.empty((rows, H), device=dev, dtype=hd)
                qkv = torch.empty((rows, 3 * H), device=dev)
                f_h = torch.empty((rows, 4 * H), device=dev, dtype=hd)
                for lay in st['layers']:
                    C.layernorm_h(x, lay['ln1'][0], lay['ln1'][1], n_h, rows, H, bf)
                    C.tc_linear(n_h, H, rows, 0, lay['w_in'], H, None, 3 * H, H, bf, out_f32=qkv, ldo=3 * H)
                    C.enc_rope(qkv, lay['freqs'], B, L, H, nh)
                    C.enc_attention(qkv, keep, a_h, B, L, H, nh, bf)
                    C.tc_linear_residual(a_h, lay['w_out'], st['zero_bias'], x, rows, H, H, bf)
                    C.enc_mask_rows(x, keep, rows, H)
                    C.layernorm_h(x, lay['ln2'][0], lay['ln2'][1], n_h, rows, H, bf)
                    C.tc_conv1d(n_h, lay['w1'], lay['b1'], None, 0, f_h, 4 * H, B, L, H, 4 * H, k, C.ACT_GELU, bf)
                    C.tc_linear_residual(f_h, lay['w2'], lay['b2'], x, rows, H, 4 * H, bf)
                    C.enc_mask_rows(x, keep, rows, H)
            C.enc_layernorm_mask(x, st['ln'][0], st['ln'][1], keep, table, B, L, H, st['ln'][2])
            vals = []
            for w, b, name in st['scal']:
                if name is None:
                    v = f0
                elif name == 'key_shift':
                    v = key_shift
                elif name == 'speed':
                    v = speed
                else:
                    v = kwargs.get(name)
                if v is None:
                    raise C.B2SError(f'the acoustic encoder was built with the {name} embedding: pass {name}=[B, T]')
                vals.append(v.to(device=dev, dtype=torch.float32).expand(B, T).contiguous())
            spk = None
            if self.use_spk_id:
                mix = kwargs.get('spk_mix_embed')
                if mix is not None:
                    if mix.shape[1] != 1:
                        raise C.B2SError('per-frame speaker mixes (spk_mix_embed [B, T, H]) are not implemented; pass [B, 1, H]')
                    spk = mix.to(device=dev, dtype=torch.float32).reshape(B, H).contiguous()
                else:
                    spk = st['spk'][spk_embed_id.to(dev).reshape(-1)].contiguous()            # [B, H] rows of the embedding table
            C.enc_assemble(table, m2p, spk, vals, [w for w, _, _ in st['scal']], [b for _, b, _ in st['scal']], st['n_var_first'],
                           st['n_var'], cond, B, T, L, H)
        return cond
