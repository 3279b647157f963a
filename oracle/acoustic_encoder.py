"""Oracle (TEST INFRASTRUCTURE): CPU restatement of the reference's FastSpeech2 acoustic encoder - the producer of the
``condition [B, T, H]`` tensor the sampling path consumes, SURVEY.md section 8 row f-2 - as plain functions over a state dict with
the reference's parameter names.  Rotary-position (``use_rope``) configuration only, the one configs/acoustic.yaml ships.

Follows, line by line:
    modules/fastspeech/acoustic_encoder.py:79-109   FastSpeech2Acoustic.forward (+ forward_variance_embedding :61-77)
    modules/fastspeech/tts_modules.py:345-351       mel2ph_to_dur
    modules/fastspeech/tts_modules.py:385-428       FastSpeech2Encoder.forward_embedding / forward
    modules/commons/common_layers.py:216-263        EncSALayer.forward (pre-LN attention + pre-LN conv FFN, padding mask)
    modules/commons/common_layers.py:152-213        MultiheadSelfAttentionWithRoPE.forward
    modules/commons/common_layers.py:120-149        TransformerFFNLayer.forward (conv k, * k^-0.5, GELU, linear)
    modules/commons/rotary_embedding_torch.py:36-75, :174-188   rotate_half / apply_rotary_emb / rotate_queries_or_keys

Pinned by tests/golden/enc_*.npz (outputs of the unmodified reference, oracle/make_golden.py).  Only tests/, smoke() and bench.py's
CPU legs may import this module; the product never does.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List

import torch
import torch.nn.functional as F


@dataclass
class AcousticEncoderCfg:
    vocab_size: int = 60
    hidden_size: int = 256
    enc_layers: int = 4
    num_heads: int = 2
    ffn_kernel_size: int = 3              # configs/acoustic.yaml:70 (base.yaml:31 says 9)
    ffn_act: str = 'gelu'
    variance_embeds: List[str] = field(default_factory=list)     # of energy / breathiness / voicing / tension, in that order
    use_key_shift_embed: bool = False
    use_speed_embed: bool = False
    use_spk_id: bool = False
    num_spk: int = 1


def _c(sd, name, dtype):
    return sd[name].to(dtype)


def mel2ph_to_dur(mel2ph, T_txt):
    """tts_modules.py:345-351: frames per token (token indices in mel2ph are 1-based, 0 = padding frame)."""
    B = mel2ph.shape[0]
    dur = mel2ph.new_zeros(B, T_txt + 1).scatter_add(1, mel2ph, torch.ones_like(mel2ph))
    return dur[:, 1:]


def rotate_queries_or_keys(freqs_param, t):
    """t [B, heads, L, D]; rotary_embedding_torch.py:174-188 with seq_dim -2, offset 0, scale 1 (interpolate_factor 1)."""
    L = t.shape[-2]
    seq = torch.arange(L, dtype=t.dtype)
    fr = torch.einsum('n,f->nf', seq.to(freqs_param.dtype), freqs_param)            # :forward
    fr = fr.repeat_interleave(2, dim=-1).to(t.dtype)                               # '... n -> ... (n r)', r = 2
    x = t.reshape(*t.shape[:-1], -1, 2)
    rot = torch.stack((-x[..., 1], x[..., 0]), dim=-1).reshape(t.shape)            # rotate_half :36-40
    return t * fr.cos() + rot * fr.sin()                                           # apply_rotary_emb :71


def enc_sa_layer(sd, cfg: AcousticEncoderCfg, i: int, x, padding_mask, dtype):
    """x [B, L, H], padding_mask [B, L] bool (True = padding); common_layers.py:237-263."""
    p = f'encoder.layers.{i}.op.'
    H, nh = cfg.hidden_size, cfg.num_heads
    hd = H // nh
    keep = (1 - padding_mask.to(dtype))[..., None]
    res = x
    h = F.layer_norm(x, (H,), _c(sd, p + 'layer_norm1.weight', dtype), _c(sd, p + 'layer_norm1.bias', dtype))
    B, L, _ = h.shape
    q, k, v = torch.split(F.linear(h, _c(sd, p + 'self_attn.in_proj.weight', dtype)), H, dim=-1)      # :179 (bias=False)
    q = q.view(B, L, nh, hd).transpose(1, 2)
    k = k.view(B, L, nh, hd).transpose(1, 2)
    v = v.view(B, L, nh, hd).transpose(1, 2)
    fr = _c(sd, p + 'self_attn.rotary_embed.freqs', dtype)
    q, k = rotate_queries_or_keys(fr, q), rotate_queries_or_keys(fr, k)                               # :187-189
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(hd)                                     # :192
    scores = scores.masked_fill(padding_mask[:, None, None, :], float('-inf'))                       # :195-198
    a = torch.matmul(F.softmax(scores, dim=-1), v)                                                    # :201-205
    a = a.transpose(1, 2).contiguous().view(B, L, H)
    a = F.linear(a, _c(sd, p + 'self_attn.out_proj.weight', dtype))                                   # :211
    x = (res + a) * keep                                                                              # :254-255
    res = x
    h = F.layer_norm(x, (H,), _c(sd, p + 'layer_norm2.weight', dtype), _c(sd, p + 'layer_norm2.bias', dtype))
    ks = cfg.ffn_kernel_size
    f = F.conv1d(h.transpose(1, 2), _c(sd, p + 'ffn.ffn_1.weight', dtype), _c(sd, p + 'ffn.ffn_1.bias', dtype), padding=ks // 2)
    f = f.transpose(1, 2) * ks ** -0.5                                                                # :143-144
    assert cfg.ffn_act == 'gelu'
    f = F.gelu(f)
    f = F.linear(f, _c(sd, p + 'ffn.ffn_2.weight', dtype), _c(sd, p + 'ffn.ffn_2.bias', dtype))
    return (res + f) * keep                                                                           # :261-262


def encoder_forward(sd, cfg: AcousticEncoderCfg, txt_embed, dur_embed, padding_mask, dtype):
    """FastSpeech2Encoder.forward with use_pos_embed and use_rope (embed_positions is None): tts_modules.py:385-428."""
    H = cfg.hidden_size
    keep = (1 - padding_mask.to(dtype))[..., None]
    x = (math.sqrt(H) * txt_embed + dur_embed) * keep                                                 # :387-389, :415
    for i in range(cfg.enc_layers):
        x = enc_sa_layer(sd, cfg, i, x, padding_mask, dtype) * keep                                   # :418
    return F.layer_norm(x, (H,), _c(sd, 'encoder.layer_norm.weight', dtype), _c(sd, 'encoder.layer_norm.bias', dtype)) * keep


def acoustic_encoder_forward(sd, cfg: AcousticEncoderCfg, txt_tokens, mel2ph, f0, key_shift=None, speed=None, spk_embed_id=None,
                             variances=None, dtype=torch.float32, spk_mix_embed=None):
    """FastSpeech2Acoustic.forward (acoustic_encoder.py:79-109): tokens [B, L] int64, mel2ph [B, T] int64, f0 [B, T] ->
    condition [B, T, H]."""
    H = cfg.hidden_size
    txt_embed = F.embedding(txt_tokens, _c(sd, 'txt_embed.weight', dtype))                            # :84
    dur = mel2ph_to_dur(mel2ph, txt_tokens.shape[1]).to(dtype)                                        # :85
    dur_embed = F.linear(dur[:, :, None], _c(sd, 'dur_embed.weight', dtype), _c(sd, 'dur_embed.bias', dtype))
    enc = encoder_forward(sd, cfg, txt_embed, dur_embed, txt_tokens == 0, dtype)                      # :87
    enc = F.pad(enc, [0, 0, 1, 0])                                                                    # :89
    cond = torch.gather(enc, 1, mel2ph[..., None].repeat([1, 1, H]))                                  # :90-91
    if cfg.use_spk_id:
        if spk_mix_embed is not None:                                                                 # :93-96 ([B, 1 or T, H])
            cond = cond + spk_mix_embed.to(dtype)
        else:
            cond = cond + F.embedding(spk_embed_id, _c(sd, 'spk_embed.weight', dtype))[:, None, :]    # :98-99
    f0_mel = (1 + f0.to(dtype) / 700).log()                                                           # :101
    cond = cond + F.linear(f0_mel[:, :, None], _c(sd, 'pitch_embed.weight', dtype), _c(sd, 'pitch_embed.bias', dtype))
    if cfg.variance_embeds:                                                                           # :62-67: stacked, summed, then added
        ve = torch.stack([F.linear(variances[name].to(dtype)[:, :, None], _c(sd, f'variance_embeds.{name}.weight', dtype),
                                   _c(sd, f'variance_embeds.{name}.bias', dtype)) for name in cfg.variance_embeds], dim=-1).sum(-1)
        cond = cond + ve
    if cfg.use_key_shift_embed:
        cond = cond + F.linear(key_shift.to(dtype)[:, :, None], _c(sd, 'key_shift_embed.weight', dtype), _c(sd, 'key_shift_embed.bias', dtype))
    if cfg.use_speed_embed:
        cond = cond + F.linear(speed.to(dtype)[:, :, None], _c(sd, 'speed_embed.weight', dtype), _c(sd, 'speed_embed.bias', dtype))
    return cond
