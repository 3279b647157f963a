"""Oracle (TEST INFRASTRUCTURE): generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden            # rewrites tests/golden/

Every fixture holds the backbone state dict (reference parameter names), the
inputs (condition, optional src_spec, every random draw the reference made, in
draw order) and the reference's output, plus a JSON ``meta`` blob describing the
hparams / constructor arguments.  Sizes are kept tiny so the fixtures stay a few
hundred KB in total.  The reference publishes no golden vectors of its own
(SURVEY.md section 4); these are outputs of the reference code itself.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

WN_SMALL = dict(num_layers=3, num_channels=32, dilation_cycle_length=2)
WN_CYC = dict(num_layers=5, num_channels=32, dilation_cycle_length=5)        # dilations 1..16
LX_SMALL = dict(num_layers=2, num_channels=32, expansion_factor=2, kernel_size=31, strong_cond=True)
LX_WEAK = dict(num_layers=2, num_channels=32, expansion_factor=2, kernel_size=7, strong_cond=False, activation='SiLU')
HIDDEN = 24
SIGMA_W = 0.05
WEIGHT_SEED = 100


def _reinit(backbone, seed):
    g = torch.Generator().manual_seed(seed)
    w = backbone.output_projection.weight
    with torch.no_grad():
        w.copy_(torch.randn(w.shape, generator=g) * SIGMA_W)
        # exercise affine / PReLU parameters that default to constants
        for name, p in backbone.named_parameters():
            if name.endswith('net.0.weight') or name == 'norm.weight':
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith('net.0.bias') or name == 'norm.bias':
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith('net.5.weight'):
                p.add_(0.05 * torch.randn(p.shape, generator=g))


def _save(name, meta, arrays, sd):
    os.makedirs(OUT_DIR, exist_ok=True)
    # state dicts are shared between cases built from the same weight seed: store each once
    sd_np = {k: v.detach().numpy() for k, v in sd.items()}
    h = hashlib.sha1()
    for k in sorted(sd_np):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd_np[k]).tobytes())
    sd_name = f'weights_{h.hexdigest()[:12]}'
    sd_path = os.path.join(OUT_DIR, sd_name + '.npz')
    if not os.path.exists(sd_path):
        np.savez_compressed(sd_path, **sd_np)
    meta = dict(meta, weights=sd_name)
    payload = {f'in.{k}': (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrays.items()}
    payload['meta'] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(OUT_DIR, name + '.npz')
    np.savez_compressed(path, **payload)
    print(f'{name}: {os.path.getsize(path) / 1024:.1f} KiB')


def backbone_case(name, btype, bargs, in_dims, n_feats, B, T, t, seed):
    ref = ref_loader.load()
    ref.hparams.update(hidden_size=HIDDEN)
    torch.manual_seed(seed)
    net = ref.backbones.build_backbone(in_dims, n_feats, btype, bargs).eval()
    _reinit(net, seed + 1)
    g = torch.Generator().manual_seed(seed + 2)
    spec = torch.randn((B, n_feats, in_dims, T), generator=g)
    cond = torch.randn((B, HIDDEN, T), generator=g)
    with torch.no_grad():
        out = net(spec, t, cond)
    meta = dict(kind='backbone', backbone_type=btype, backbone_args=bargs, in_dims=in_dims, n_feats=n_feats,
                hidden_size=HIDDEN)
    _save(name, meta, dict(spec=spec, t=t, cond=cond, out=out), net.state_dict())


def diffusion_case(name, cls_name, ctor, hp, B, T, seed, src=None, n_draws=1, variance_inputs=None):
    """Runs ``model(condition, src_spec=src, infer=True)`` of a ddpm.py / reflow.py class."""
    ref = ref_loader.load()
    ref.hparams.clear()
    ref.hparams.update(hidden_size=HIDDEN, schedule_type='linear', infer=False)
    ref.hparams.update(hp)
    mod = ref.ddpm if hasattr(ref.ddpm, cls_name) else ref.reflow
    torch.manual_seed(WEIGHT_SEED)
    model = getattr(mod, cls_name)(**ctor).eval()
    bb = model.denoise_fn if hasattr(model, 'denoise_fn') else model.velocity_fn
    _reinit(bb, WEIGHT_SEED + 1)
    g = torch.Generator().manual_seed(seed + 2)
    condition = torch.randn((B, T, HIDDEN), generator=g)
    arrays = dict(condition=condition)
    if src == 'mel':
        smin, smax = ctor['spec_min'][0], ctor['spec_max'][0]
        src_spec = torch.rand((B, T, ctor['out_dims']), generator=g) * (smax - smin) + smin
        arrays['src_spec'] = src_spec
    elif src == 'curve':
        src_spec = torch.randn((B, T), generator=g) * 3
        arrays['src_spec'] = src_spec
    elif src == 'curves':
        src_spec = [torch.randn((B, T), generator=g) * 20 - 50 for _ in ctor['ranges']]
        for i, s in enumerate(src_spec):
            arrays[f'src_spec{i}'] = s
    else:
        src_spec = None
    F_, M_ = model.num_feats, model.out_dims
    torch.manual_seed(seed + 3)
    with torch.no_grad():
        out = model(condition, src_spec=src_spec, infer=True)
    torch.manual_seed(seed + 3)
    draws = torch.stack([torch.randn(B, F_, M_, T) for _ in range(n_draws)])
    arrays['draws'] = draws
    if isinstance(out, (list, tuple)):
        for i, o in enumerate(out):
            arrays[f'out{i}'] = o
    else:
        arrays['out'] = out
    meta = dict(kind='diffusion', cls=cls_name, ctor=ctor, hparams=hp, hidden_size=HIDDEN, n_draws=n_draws)
    _save(name, meta, arrays, bb.state_dict())


def aux_case(name, args, out_dims, n_feats, spec_min, spec_max, B, T, seed, infer=True):
    """Runs ``AuxDecoderAdaptor(condition, infer)`` of modules/aux_decoder/__init__.py (ConvNeXt decoder, the producer of x_start)."""
    ref_loader.load()
    from modules.aux_decoder import AuxDecoderAdaptor
    torch.manual_seed(seed)
    model = AuxDecoderAdaptor(in_dims=HIDDEN, out_dims=out_dims, num_feats=n_feats, spec_min=spec_min, spec_max=spec_max,
                              aux_decoder_arch='convnext', aux_decoder_args=args).eval()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for pname, p in model.named_parameters():
            if pname.endswith('gamma'):                    # layer scale is initialised to 1e-6: make the blocks matter
                p.copy_(0.5 + 0.5 * torch.rand(p.shape, generator=g))
            elif pname.endswith('norm.weight') or pname.endswith('norm.bias'):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    condition = torch.randn((B, T, HIDDEN), generator=g)
    with torch.no_grad():
        out = model(condition, infer=infer)
    meta = dict(kind='aux_decoder', args=args, in_dims=HIDDEN, out_dims=out_dims, num_feats=n_feats, spec_min=spec_min,
                spec_max=spec_max, infer=infer)
    _save(name, meta, dict(condition=condition, out=out), model.decoder.state_dict())


def enc_case(name, hp, vocab, B, L, T, seed, extras=()):
    """Runs ``FastSpeech2Acoustic(txt_tokens, mel2ph, f0, ...)`` of modules/fastspeech/acoustic_encoder.py (the producer of the
    condition tensor), rotary-position configuration."""
    ref = ref_loader.load()
    ref.hparams.clear()
    ref.hparams.update(hidden_size=32, enc_layers=2, enc_ffn_kernel_size=3, ffn_act='gelu', dropout=0.1, num_heads=2,
                       use_pos_embed=True, rel_pos=True, use_rope=True, use_spk_id=False, num_spk=1)
    ref.hparams.update(hp)
    from modules.fastspeech.acoustic_encoder import FastSpeech2Acoustic
    torch.manual_seed(seed)
    model = FastSpeech2Acoustic(vocab).eval()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for pname, p in model.named_parameters():
            if pname.endswith('bias') or 'layer_norm' in pname:       # biases are initialised to 0, LayerNorm to (1, 0): exercise them
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    lens = torch.randint(max(2, L // 2), L + 1, (B,), generator=g)
    lens[0] = L
    tokens = torch.zeros((B, L), dtype=torch.long)
    mel2ph = torch.zeros((B, T), dtype=torch.long)
    for b in range(B):
        n = int(lens[b])
        tokens[b, :n] = torch.randint(1, vocab, (n,), generator=g)
        frames = int(T if b == 0 else torch.randint(T // 2, T + 1, (1,), generator=g))
        cuts = torch.sort(torch.randint(0, frames + 1, (n - 1,), generator=g)).values
        bounds = torch.cat([torch.tensor([0]), cuts, torch.tensor([frames])])
        for j in range(n):
            mel2ph[b, int(bounds[j]):int(bounds[j + 1])] = j + 1          # 1-based token index per frame, 0 = padding frame
    f0 = 100 + 300 * torch.rand((B, T), generator=g)
    kw, arrays = {}, dict(txt_tokens=tokens, mel2ph=mel2ph, f0=f0)
    for e in extras:
        if e == 'spk':
            kw['spk_embed_id'] = torch.randint(0, hp['num_spk'], (B,), generator=g)
            arrays['spk_embed_id'] = kw['spk_embed_id']
        else:
            kw[e] = torch.randn((B, T), generator=g)
            arrays[e] = kw[e]
    with torch.no_grad():
        out = model(tokens, mel2ph, f0, **kw)
    arrays['out'] = out
    meta = dict(kind='acoustic_encoder', hparams={k: ref.hparams[k] for k in ('hidden_size', 'enc_layers', 'enc_ffn_kernel_size',
                'ffn_act', 'num_heads', 'use_spk_id', 'num_spk')}, vocab=vocab, extras=list(extras),
                flags={k: bool(ref.hparams.get(k, False)) for k in ('use_energy_embed', 'use_breathiness_embed', 'use_voicing_embed',
                                                                   'use_tension_embed', 'use_key_shift_embed', 'use_speed_embed')})
    _save(name, meta, arrays, model.state_dict())


def enc_main():
    enc_case('enc_fs2_plain', {}, 12, 2, 7, 23, 400)
    enc_case('enc_fs2_all_embeds_k9', dict(enc_layers=3, enc_ffn_kernel_size=9, use_spk_id=True, num_spk=3, use_energy_embed=True,
                                           use_breathiness_embed=True, use_key_shift_embed=True, use_speed_embed=True),
             20, 3, 11, 41, 401, extras=('spk', 'energy', 'breathiness', 'key_shift', 'speed'))


def aux_main():
    aux_case('aux_convnext_mel', dict(num_channels=32, num_layers=2, kernel_size=7, dropout_rate=0.1), 16, 1,
             [-12.] * 16, [0.] * 16, 2, 37, 300)
    aux_case('aux_convnext_feats2_k5', dict(num_channels=48, num_layers=3, kernel_size=5, dropout_rate=0.0), 8, 2,
             [[-10.] * 8, [-4.] * 8], [[2.] * 8, [6.] * 8], 3, 23, 301)
    aux_case('aux_convnext_train_mode_output', dict(num_channels=32, num_layers=1, kernel_size=7), 16, 1,
             [-12.] * 16, [0.] * 16, 1, 9, 302, infer=False)


def voc_case(name, h, B, T, seed, f0_kind='mixed'):
    """Runs ``Generator(h)(mel, f0)`` of modules/nsf_hifigan/models.py:206-289 (NSF-HiFiGAN, mel + f0 -> waveform) after
    ``remove_weight_norm()`` (what load_model does, :31-32), with the fan-in scaled random weights of
    ``oracle.vocoder.random_state_dict`` (the reference's own N(0, 0.01) init would make every block a near-identity), and records
    the two random draws of SineGen in draw order."""
    from oracle import vocoder as ov
    voc = ref_loader.load_vocoder()
    cfg = ov.NsfHifiGanCfg(**h)
    hd = voc.AttrDict(dict(h, upsample_rates=list(cfg.upsample_rates), upsample_kernel_sizes=list(cfg.upsample_kernel_sizes),
                           resblock_kernel_sizes=list(cfg.resblock_kernel_sizes),
                           resblock_dilation_sizes=[list(d) for d in cfg.resblock_dilation_sizes]))
    torch.manual_seed(seed)
    gen = voc.models.Generator(hd).eval()
    gen.remove_weight_norm()
    sd = ov.random_state_dict(cfg, seed + 1)
    missing = gen.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(seed + 2)
    mel = torch.randn((B, cfg.num_mels, T), generator=g) * 1.5 - 4.0
    f0 = 110.0 * 2 ** (2 * torch.rand((B, T), generator=g))                 # 110 .. 440 Hz
    if f0_kind == 'mixed':
        f0[:, T // 3: T // 3 + max(1, T // 4)] = 0.                        # an unvoiced stretch
        f0[-1, -2:] = 0.
    arrays = dict(mel=mel, f0=f0)
    torch.manual_seed(seed + 3)
    with torch.no_grad():
        out = gen(mel, f0)
    if not cfg.mini_nsf:
        torch.manual_seed(seed + 3)                                         # the same two draws, in the reference's order (:147, :170)
        arrays['rand_ini'] = torch.rand(1, 1, cfg.harmonic_num + 1)
        arrays['noise'] = torch.randn(B, T * cfg.hop, cfg.harmonic_num + 1)
    arrays['out'] = out
    _save(name, dict(kind='vocoder', h=h), arrays, sd)


def voc_main():
    voc_case('voc_nsf_resblock1', dict(num_mels=16, sampling_rate=16000, upsample_rates=(4, 2, 2), upsample_kernel_sizes=(8, 4, 4),
                                      upsample_initial_channel=32, resblock='1', resblock_kernel_sizes=(3, 7),
                                      resblock_dilation_sizes=((1, 3, 5), (1, 3, 5))), 2, 11, 500)
    voc_case('voc_nsf_resblock2', dict(num_mels=8, sampling_rate=22050, upsample_rates=(4, 4), upsample_kernel_sizes=(8, 8),
                                      upsample_initial_channel=16, resblock='2', resblock_kernel_sizes=(3, 5),
                                      resblock_dilation_sizes=((1, 3), (1, 3))), 1, 7, 501)
    voc_case('voc_mini_nsf', dict(num_mels=8, sampling_rate=16000, upsample_rates=(2, 2, 2), upsample_kernel_sizes=(4, 4, 4),
                                 upsample_initial_channel=32, resblock='1', resblock_kernel_sizes=(3,),
                                 resblock_dilation_sizes=((1, 3, 5),), mini_nsf=True), 2, 9, 502)


def ds_case(name, hp, param, vocab, spk_map, seed):
    """Runs the UNMODIFIED ``DiffSingerAcousticInfer.preprocess_input`` (inference/ds_acoustic.py:68-166, with
    ``BaseSVSInfer.load_speaker_mix``, basics/base_svs_infer.py:37-122) on a synthetic ``.ds`` segment.  The methods are called
    unbound on a stand-in object that carries exactly the attributes they read (the real constructor loads checkpoints)."""
    import types
    da = ref_loader.load_acoustic_infer()
    ref = ref_loader.load()
    ref.hparams.clear()
    ref.hparams.update(hp)
    cls = da.DiffSingerAcousticInfer
    me = types.SimpleNamespace(device='cpu', timestep=hp['hop_size'] / hp['audio_sample_rate'],
                               ph_encoder=da.TokenTextEncoder(vocab_list=vocab), lr=da.LengthRegulator(), spk_map=spk_map,
                               variances_to_embed={v for v in ('energy', 'breathiness', 'voicing', 'tension') if hp.get(f'use_{v}_embed')})
    me.load_speaker_mix = types.MethodType(cls.load_speaker_mix, me)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        batch = cls.preprocess_input(me, param, idx=0)
    arrays = {k: v for k, v in batch.items()}
    meta = dict(kind='ds_preprocess', hparams=hp, param=param, vocab=vocab, spk_map=spk_map, keys=sorted(batch.keys()))
    _save(name, meta, arrays, {'unused': torch.zeros(1)})


def _ds_param(seed, n_ph, **extra):
    g = np.random.RandomState(seed)
    vocab = ['AP', 'SP', 'a', 'ai', 'b', 'ch', 'e', 'i', 'n', 'sh', 'u', 'zh']
    ph = ['SP'] + [vocab[2 + int(g.randint(len(vocab) - 2))] for _ in range(n_ph - 2)] + ['AP']
    dur = g.uniform(0.04, 0.37, n_ph)
    total = float(dur.sum())
    f0_ts = 0.005
    n_f0 = int(total / f0_ts) + 7
    f0 = 220.0 * 2 ** (0.5 * np.sin(np.arange(n_f0) * 0.013) + 0.02 * g.randn(n_f0))
    p = dict(offset=1.25, ph_seq=' '.join(ph), ph_dur=' '.join('%.6f' % d for d in dur), f0_seq=' '.join('%.1f' % v for v in f0),
             f0_timestep=str(f0_ts))

    def curve(lo, hi, ts):
        n = int(total / ts) + 3
        return ' '.join('%.3f' % v for v in g.uniform(lo, hi, n)), str(ts)
    for key, (lo, hi, ts) in extra.items():
        if key in ('energy', 'breathiness', 'voicing', 'tension', 'velocity', 'gender'):
            p[key], p[f'{key}_timestep'] = curve(lo, hi, ts)
    return p, vocab


def ds_main():
    base = dict(hop_size=512, audio_sample_rate=44100, use_spk_id=False, use_key_shift_embed=False, use_speed_embed=False)
    p, vocab = _ds_param(600, 9)
    ds_case('ds_preprocess_plain', dict(base), p, vocab, None, 600)
    aug = dict(random_pitch_shifting=dict(range=[-5., 5.]), random_time_stretching=dict(range=[0.5, 2.]))
    p, vocab = _ds_param(601, 14, energy=(-60., -10., 0.011), breathiness=(-80., -30., 0.02), velocity=(0.3, 2.4, 0.05), gender=(-1.2, 1.2, 0.03))
    p['spk_mix'] = {'alto': 0.7, 'tenor': 0.5}
    ds_case('ds_preprocess_all_static_mix', dict(base, use_spk_id=True, use_key_shift_embed=True, use_speed_embed=True, use_energy_embed=True,
                                                 use_breathiness_embed=True, augmentation_args=aug), p, vocab, {'alto': 0, 'tenor': 3, 'bass': 1}, 601)
    p, vocab = _ds_param(602, 6, voicing=(-70., -5., 0.0116), tension=(-3., 3., 0.01))
    p['gender'] = -0.4
    n = 40
    p['spk_mix'] = {'alto': ' '.join('%.3f' % v for v in np.linspace(0.1, 0.9, n)), 'bass': 0.25}
    p['spk_mix_timestep'] = '0.05'
    ds_case('ds_preprocess_dynamic_mix_static_gender', dict(base, hop_size=256, audio_sample_rate=22050, use_spk_id=True, use_key_shift_embed=True,
                                                            use_speed_embed=True, use_voicing_embed=True, use_tension_embed=True,
                                                            augmentation_args=aug), p, vocab, {'alto': 0, 'tenor': 3, 'bass': 1}, 602)


def main():
    aux_main()
    enc_main()
    voc_main()
    ds_main()
    # ---- backbone forward ------------------------------------------------------------------
    backbone_case('bb_wavenet_int_t', 'wavenet', WN_SMALL, 16, 1, 2, 37, torch.tensor([950, 3]), 10)
    backbone_case('bb_wavenet_float_t1', 'wavenet', WN_CYC, 16, 1, 3, 41, torch.tensor([437.25]), 11)
    backbone_case('bb_wavenet_feats2', 'wavenet', WN_SMALL, 8, 2, 2, 19, torch.tensor([12.5, 700.0]), 12)
    backbone_case('bb_lynxnet_strong', 'lynxnet', LX_SMALL, 16, 1, 2, 37, torch.tensor([950, 3]), 13)
    backbone_case('bb_lynxnet_weak_silu', 'lynxnet', LX_WEAK, 8, 2, 2, 23, torch.tensor([333.5]), 14)

    mel = dict(out_dims=16, num_feats=1, spec_min=[-12.], spec_max=[0.])
    wn = dict(backbone_type='wavenet', backbone_args=WN_SMALL)
    lx = dict(backbone_type='lynxnet', backbone_args=LX_SMALL)
    base = dict(use_shallow_diffusion=False, diff_speedup=1, diff_accelerator='ddim')

    # ---- GaussianDiffusion, every sampler ---------------------------------------------------
    # full-depth DDPM on a short schedule (timesteps=30): 30 ancestral steps, 31 draws
    diffusion_case('gd_ddpm_full_T30', 'GaussianDiffusion', dict(mel, timesteps=30, k_step=30, **wn),
                   dict(base), 2, 21, 20, n_draws=31)
    # shallow DDPM: K_step=12 of 1000, q_sample start from src_spec
    diffusion_case('gd_ddpm_shallow_K12', 'GaussianDiffusion', dict(mel, timesteps=1000, k_step=12, **wn),
                   dict(base, use_shallow_diffusion=True, K_step_infer=12), 2, 21, 21, src='mel', n_draws=13)
    for acc, b in (('ddim', 2), ('pndm', 1), ('dpm-solver', 2), ('unipc', 2)):
        tag = acc.replace('-', '')
        diffusion_case(f'gd_{tag}_10', 'GaussianDiffusion', dict(mel, timesteps=1000, k_step=1000, **wn),
                       dict(base, diff_speedup=100, diff_accelerator=acc), b, 21, 22)
        diffusion_case(f'gd_{tag}_5', 'GaussianDiffusion', dict(mel, timesteps=1000, k_step=1000, **wn),
                       dict(base, diff_speedup=200, diff_accelerator=acc), b, 17, 23)
    # shallow + accelerated (schedule is betas[:t_max], N = 400)
    for acc in ('ddim', 'dpm-solver', 'unipc'):
        tag = acc.replace('-', '')
        diffusion_case(f'gd_{tag}_shallow400_8', 'GaussianDiffusion', dict(mel, timesteps=1000, k_step=400, **wn),
                       dict(base, use_shallow_diffusion=True, K_step_infer=400, diff_speedup=50,
                            diff_accelerator=acc), 2, 19, 24, src='mel')
    # cosine schedule: DPM-Solver clips lambda (total_N 996), UniPC does not
    for acc in ('dpm-solver', 'unipc'):
        tag = acc.replace('-', '')
        diffusion_case(f'gd_{tag}_cosine_10', 'GaussianDiffusion', dict(mel, timesteps=1000, k_step=1000, **wn),
                       dict(base, schedule_type='cosine', diff_speedup=100, diff_accelerator=acc), 2, 19, 25)
    # LYNXNet under DDIM
    diffusion_case('gd_ddim_lynx_10', 'GaussianDiffusion', dict(mel, timesteps=1000, k_step=1000, **lx),
                   dict(base, diff_speedup=100, diff_accelerator='ddim'), 2, 33, 26)

    # ---- RectifiedFlow ---------------------------------------------------------------------
    rf = dict(mel, time_scale_factor=1000)
    for alg in ('euler', 'rk2', 'rk4', 'rk5'):
        diffusion_case(f'rf_{alg}_4_lynx', 'RectifiedFlow', dict(rf, t_start=0., **lx),
                       dict(use_shallow_diffusion=False, sampling_algorithm=alg, sampling_steps=4), 2, 33, 30)
    diffusion_case('rf_euler_shallow_6_lynx', 'RectifiedFlow', dict(rf, t_start=0.4, **lx),
                   dict(use_shallow_diffusion=True, T_start_infer=0.4, sampling_algorithm='euler', sampling_steps=6),
                   2, 33, 31, src='mel')
    diffusion_case('rf_euler_5_wavenet', 'RectifiedFlow', dict(rf, t_start=0., **wn),
                   dict(use_shallow_diffusion=False, sampling_algorithm='euler', sampling_steps=5), 2, 21, 32)

    # ---- variance models (repeat bins, clamps, mean over bins) ------------------------------
    diffusion_case('var_pitch_ddim_10', 'PitchDiffusion',
                   dict(vmin=-8., vmax=8., cmin=-12., cmax=12., repeat_bins=8, timesteps=1000, k_step=1000, **wn),
                   dict(base, diff_speedup=100, diff_accelerator='ddim'), 2, 21, 40)
    diffusion_case('var_multi_unipc_10', 'MultiVarianceDiffusion',
                   dict(ranges=[(-96., -12.), (-96., -20.)], clamps=[(-96., -12.), (-96., -20.)], repeat_bins=6,
                        timesteps=1000, k_step=1000, **wn),
                   dict(base, diff_speedup=100, diff_accelerator='unipc'), 2, 21, 41)
    diffusion_case('var_pitch_reflow_euler_4', 'PitchRectifiedFlow',
                   dict(vmin=-8., vmax=8., cmin=-12., cmax=12., repeat_bins=8, time_scale_factor=1000, **wn),
                   dict(use_shallow_diffusion=False, sampling_algorithm='euler', sampling_steps=4), 2, 21, 42)
    diffusion_case('var_multi_reflow_rk4_3', 'MultiVarianceRectifiedFlow',
                   dict(ranges=[(-96., -12.), (-96., -20.)], clamps=[(-96., -12.), None], repeat_bins=6,
                        time_scale_factor=1000, **wn),
                   dict(use_shallow_diffusion=False, sampling_algorithm='rk4', sampling_steps=3), 2, 21, 43)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'aux':
        aux_main()                                          # only the aux-decoder fixtures
    elif len(sys.argv) > 1 and sys.argv[1] == 'ds':
        ds_main()                                           # only the .ds preprocessing fixtures
    elif len(sys.argv) > 1 and sys.argv[1] == 'voc':
        voc_main()                                          # only the vocoder fixtures
    elif len(sys.argv) > 1 and sys.argv[1] == 'enc':
        enc_main()                                          # only the acoustic-encoder fixtures
    else:
        main()
