"""GPU parity of the FastSpeech2 acoustic encoder (producer of the condition tensor) against the CPU oracle
(oracle/acoustic_encoder.py, pinned by tests/golden/enc_*.npz) through the product's public API.  GEMM operands are 16 bits, the
rest fp32: the condition must agree within 2e-2 absolute (it is cast to 16 bits by the sampling path anyway)."""
import pytest
import torch

from oracle import acoustic_encoder as OE
from tol16 import check16

pytestmark = pytest.mark.gpu


def _inputs(cfg, B, L, T, seed):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(max(2, L // 2), L + 1, (B,), generator=g)
    lens[0] = L
    tokens = torch.zeros((B, L), dtype=torch.long)
    mel2ph = torch.zeros((B, T), dtype=torch.long)
    for b in range(B):
        n = int(lens[b])
        tokens[b, :n] = torch.randint(1, cfg.vocab_size, (n,), generator=g)
        frames = T if b == 0 else int(torch.randint(T // 2, T + 1, (1,), generator=g))
        cuts = torch.sort(torch.randint(0, frames + 1, (n - 1,), generator=g)).values
        bounds = torch.cat([torch.tensor([0]), cuts, torch.tensor([frames])])
        for j in range(n):
            mel2ph[b, int(bounds[j]):int(bounds[j + 1])] = j + 1
    f0 = 100 + 300 * torch.rand((B, T), generator=g)
    extra = {n: torch.randn((B, T), generator=g) for n in cfg.variance_embeds}
    ks = torch.randn((B, T), generator=g) if cfg.use_key_shift_embed else None
    sp = torch.randn((B, T), generator=g) if cfg.use_speed_embed else None
    spk = torch.randint(0, cfg.num_spk, (B,), generator=g) if cfg.use_spk_id else None
    return tokens, mel2ph, f0, extra, ks, sp, spk


@pytest.mark.parametrize('precision', ['fp16', 'bf16'])
@pytest.mark.parametrize('case', ['acoustic_default', 'all_embeds_k9', 'tiny'])
def test_acoustic_encoder_against_oracle(case, precision):
    import xiaoicesing_io_b200 as P
    dev = torch.device('cuda:0')
    if case == 'acoustic_default':       # configs/acoustic.yaml + base.yaml: H 256, 4 layers, 2 heads, 3-tap FFN conv
        cfg, B, L, T = OE.AcousticEncoderCfg(), 3, 61, 690
    elif case == 'all_embeds_k9':
        cfg = OE.AcousticEncoderCfg(vocab_size=40, hidden_size=128, enc_layers=2, num_heads=2, ffn_kernel_size=9,
                                    variance_embeds=['energy', 'breathiness', 'tension'], use_key_shift_embed=True,
                                    use_speed_embed=True, use_spk_id=True, num_spk=5)
        B, L, T = 4, 23, 257
    else:                                # one token, one frame / a handful
        cfg, B, L, T = OE.AcousticEncoderCfg(hidden_size=64, enc_layers=1, num_heads=4), 2, 2, 3
    P.hparams.clear()
    P.hparams.update(hidden_size=cfg.hidden_size, enc_layers=cfg.enc_layers, enc_ffn_kernel_size=cfg.ffn_kernel_size, ffn_act='gelu',
                     num_heads=cfg.num_heads, use_pos_embed=True, rel_pos=True, use_rope=True, dropout=0.1, use_spk_id=cfg.use_spk_id,
                     num_spk=cfg.num_spk, use_key_shift_embed=cfg.use_key_shift_embed, use_speed_embed=cfg.use_speed_embed,
                     b2s_precision=precision, **{f'use_{n}_embed': True for n in cfg.variance_embeds})
    torch.manual_seed(3)
    model = P.FastSpeech2Acoustic(cfg.vocab_size)
    g = torch.Generator().manual_seed(4)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith('bias') or 'layer_norm' in n:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.to(dev).eval()
    tokens, mel2ph, f0, extra, ks, sp, spk = _inputs(cfg, B, L, T, 11)
    kw = {n: v.to(dev) for n, v in extra.items()}
    out = model(tokens.to(dev), mel2ph.to(dev), f0.to(dev), key_shift=None if ks is None else ks.to(dev),
                speed=None if sp is None else sp.to(dev), spk_embed_id=None if spk is None else spk.to(dev), **kw)
    ref = OE.acoustic_encoder_forward(sd, cfg, tokens, mel2ph, f0, key_shift=ks, speed=sp, spk_embed_id=spk, variances=extra,
                                      dtype=torch.float64)
    assert out.shape == ref.shape and bool(torch.isfinite(out).all())
    err, scale = float((out.double().cpu() - ref).abs().max()), float(ref.abs().max())
    print(dict(test='acoustic_encoder', case=case, precision=precision, max_abs=err, ref_absmax=scale))
    check16(precision, err, scale)
    if cfg.use_spk_id:       # speaker mixes (spk_mix_embed, acoustic_encoder.py:93-96): one row per utterance and one row per frame
        E = sd['spk_embed.weight']
        wgt = torch.rand((B, T, 1), generator=torch.Generator().manual_seed(2))
        for mix in (E[spk][:, None, :] * 0.3 + E[(spk + 1) % cfg.num_spk][:, None, :] * 0.7,
                    E[spk][:, None, :] * wgt + E[(spk + 1) % cfg.num_spk][:, None, :] * (1 - wgt)):
            o2 = model(tokens.to(dev), mel2ph.to(dev), f0.to(dev), key_shift=ks.to(dev), speed=sp.to(dev), spk_mix_embed=mix.to(dev), **kw)
            r2 = OE.acoustic_encoder_forward(sd, cfg, tokens, mel2ph, f0, key_shift=ks, speed=sp, variances=extra, dtype=torch.float64,
                                             spk_mix_embed=mix)
            check16(precision, float((o2.double().cpu() - r2).abs().max()), float(r2.abs().max()))
    # padding frames (mel2ph == 0) carry only the frame-rate embeddings; a second call reproduces the bits
    assert torch.equal(out, model(tokens.to(dev), mel2ph.to(dev), f0.to(dev), key_shift=None if ks is None else ks.to(dev),
                                  speed=None if sp is None else sp.to(dev), spk_embed_id=None if spk is None else spk.to(dev), **kw))


def test_whole_acoustic_model_tokens_to_mel():
    """``DiffSingerAcoustic`` (modules/toplevel.py:32-102): tokens -> FastSpeech2 encoder -> ConvNeXt aux decoder -> shallow DDIM sampling,
    every stage on the B200 kernels, against the chain of the three oracles with the same injected noise; frames with mel2ph == 0
    come out as zeros like the reference's."""
    import xiaoicesing_io_b200 as P
    from oracle import aux_decoder as OA, denoisers as OD, samplers as OS
    dev = torch.device('cuda:0')
    B, L, T = 2, 31, 300
    ecfg = OE.AcousticEncoderCfg(vocab_size=30, enc_layers=2)
    acfg = OA.ConvNeXtCfg(num_channels=256, num_layers=2)
    wcfg = OD.WaveNetCfg(num_layers=6)
    smin, smax = [-12.] * 128, [0.] * 128
    P.hparams.clear()
    P.hparams.update(hidden_size=256, enc_layers=2, enc_ffn_kernel_size=3, ffn_act='gelu', num_heads=2, use_pos_embed=True, rel_pos=True,
                     use_rope=True, dropout=0.1, use_spk_id=False, num_spk=1, schedule_type='linear', infer=False,
                     use_shallow_diffusion=True, K_step_infer=100, diff_speedup=10, diff_accelerator='ddim', timesteps=1000, K_step=100,
                     spec_min=smin, spec_max=smax, diffusion_type='ddpm', backbone_type='wavenet',
                     backbone_args=dict(num_layers=6, num_channels=256, dilation_cycle_length=4),
                     shallow_diffusion_args=dict(train_aux_decoder=True, train_diffusion=True, val_gt_start=False, aux_decoder_grad=0.1,
                                                 aux_decoder_arch='convnext',
                                                 aux_decoder_args=dict(num_channels=256, num_layers=2, kernel_size=7)),
                     b2s_precision='fp16')
    torch.manual_seed(8)
    model = P.DiffSingerAcoustic(30, 128)
    g = torch.Generator().manual_seed(9)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith('gamma'):
                p.copy_(0.2 + 0.3 * torch.rand(p.shape, generator=g))
            elif n.startswith('diffusion') and n.endswith('output_projection.weight') and p.dim() == 3 and p.shape[0] == 128:
                p.copy_(0.01 * torch.randn(p.shape, generator=g))
            elif n.startswith('fs2') and (n.endswith('bias') or 'layer_norm' in n):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.to(dev).eval()
    tokens, mel2ph, f0, _, _, _, _ = _inputs(ecfg, B, L, T, 13)
    noise0 = torch.randn((B, 1, 128, T), generator=g)
    model.diffusion._noise_source = lambda shape: noise0.to(dev)
    out = model(tokens.to(dev), mel2ph.to(dev), f0.to(dev), infer=True)
    # oracle chain
    sub = lambda pre: {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
    cond = OE.acoustic_encoder_forward(sub('fs2.'), ecfg, tokens, mel2ph, f0)
    keep = (mel2ph > 0).float()[:, :, None]
    aux = OA.aux_adaptor_forward(sub('aux_decoder.decoder.'), acfg, cond, smin, smax, infer=True) * keep
    sch = OS.DiffusionSchedule(1000, 'linear')
    x_start = OS.norm_spec(aux, torch.tensor(-12.), torch.tensor(0.)).transpose(-2, -1)[:, None]
    with torch.no_grad():
        x = OS.gaussian_diffusion_inference(OD.make_denoiser(sub('diffusion.denoise_fn.'), wcfg), sch, cond.transpose(1, 2), k_step=100,
                                            timesteps=1000, use_shallow=True, K_step_infer=100, speedup=10, accelerator='ddim',
                                            noise0=noise0, x_start=x_start, step_noise=[])
    ref = OS.denorm_spec(x, torch.tensor(-12.), torch.tensor(0.)) * keep
    e_aux = float((out.aux_out.cpu().double() - aux.double()).abs().max())
    e_mel = float((out.diff_out.cpu().double() - ref.double()).abs().max())
    print(dict(test='tokens_to_mel', aux_max_abs=e_aux, mel_max_abs=e_mel, ref_absmax=float(ref.abs().max())))
    assert e_aux <= 2e-2 and e_mel <= 2e-2, (e_aux, e_mel)
    pad = (mel2ph == 0)
    assert pad.any() and float(out.diff_out.cpu()[pad].abs().max()) == 0.0


def test_whole_acoustic_model_training_branch_forward_values():
    """``infer=False`` (modules/toplevel.py:103-120): the aux decoder's normalised prediction and (x_recon, noise) of the denoiser on
    q_sample(gt_mel) - forward values for validation losses."""
    import xiaoicesing_io_b200 as P
    dev = torch.device('cuda:0')
    smin, smax = [-12.] * 128, [0.] * 128
    P.hparams.clear()
    P.hparams.update(hidden_size=256, enc_layers=1, enc_ffn_kernel_size=3, ffn_act='gelu', num_heads=2, use_pos_embed=True, rel_pos=True,
                     use_rope=True, use_spk_id=False, num_spk=1, schedule_type='linear', infer=False, use_shallow_diffusion=True,
                     timesteps=1000, K_step=100, spec_min=smin, spec_max=smax, diffusion_type='ddpm', backbone_type='wavenet',
                     backbone_args=dict(num_layers=4, num_channels=256, dilation_cycle_length=4),
                     shallow_diffusion_args=dict(train_aux_decoder=True, train_diffusion=True, val_gt_start=False, aux_decoder_grad=0.1,
                                                 aux_decoder_arch='convnext',
                                                 aux_decoder_args=dict(num_channels=128, num_layers=1, kernel_size=7)),
                     b2s_precision='fp16')
    torch.manual_seed(8)
    model = P.DiffSingerAcoustic(30, 128).to(dev).eval()
    tokens, mel2ph, f0, _, _, _, _ = _inputs(OE.AcousticEncoderCfg(vocab_size=30), 2, 17, 150, 5)
    gt = (torch.rand(2, 150, 128) * 12 - 12).to(dev)
    out = model(tokens.to(dev), mel2ph.to(dev), f0.to(dev), gt_mel=gt, infer=False)
    x_recon, noise = out.diff_out
    assert tuple(out.aux_out.shape) == (2, 150, 128) and x_recon.shape == noise.shape == (2, 1, 128, 150)
    assert bool(torch.isfinite(out.aux_out).all()) and bool(torch.isfinite(x_recon).all())
    with pytest.raises(P.B2SError):
        model(tokens.to(dev), mel2ph.to(dev), f0.to(dev), infer=False)
