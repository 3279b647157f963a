"""Wall-clock time of a whole synthetic `.ds` project through the inference driver (xiaoicesing_io_b200.infer): N segments of 2 - 30 s ->
preprocess_input -> FastSpeech2 encoder -> ConvNeXt aux decoder -> shallow DDIM (20 evaluations of WaveNet 20x256) -> NSF-HiFiGAN ->
one cross-faded waveform + 16-bit WAV, batched through the segment driver and as the reference's one-segment-per-call loop.
Run under gpurun: python scripts/ds_project_bench.py [n_segments]"""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                   # noqa: E402  (VOCODER_H)
import xiaoicesing_io_b200 as P                # noqa: E402

VOCAB = ['AP', 'SP'] + [f'ph{i}' for i in range(57)]


def project(n, seed=0):
    g = np.random.RandomState(seed)
    params, offset = [], 0.0
    for _ in range(n):
        seconds = float(g.uniform(2.0, 30.0))
        n_ph = max(3, int(seconds * 6))
        dur = g.uniform(0.05, 0.3, n_ph)
        dur *= seconds / dur.sum()
        n_f0 = int(seconds / 0.005) + 5
        f0 = 220.0 * 2 ** (0.5 * np.sin(np.arange(n_f0) * 0.01))
        params.append(dict(offset=offset, ph_seq=' '.join(['SP'] + [VOCAB[2 + int(g.randint(57))] for _ in range(n_ph - 2)] + ['AP']),
                           ph_dur=' '.join('%.5f' % d for d in dur), f0_seq=' '.join('%.1f' % v for v in f0), f0_timestep='0.005',
                           seed=int(g.randint(1 << 30))))
        offset += seconds + 0.5
    return params


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    dev = torch.device('cuda:0')
    P.hparams.clear()
    P.hparams.update(hidden_size=256, enc_layers=4, enc_ffn_kernel_size=3, ffn_act='gelu', num_heads=2, use_pos_embed=True, rel_pos=True,
                     use_rope=True, dropout=0.1, use_spk_id=False, num_spk=1, schedule_type='linear', infer=False,
                     use_shallow_diffusion=True, K_step_infer=400, diff_speedup=20, diff_accelerator='ddim', timesteps=1000, K_step=400,
                     spec_min=[-12.0] * 128, spec_max=[0.0] * 128, diffusion_type='ddpm', backbone_type='wavenet',
                     backbone_args=dict(num_layers=20, num_channels=256, dilation_cycle_length=4),
                     shallow_diffusion_args=dict(train_aux_decoder=True, train_diffusion=True, val_gt_start=False, aux_decoder_grad=0.1,
                                                 aux_decoder_arch='convnext', aux_decoder_args=dict(num_channels=512, num_layers=6, kernel_size=7)),
                     hop_size=512, audio_sample_rate=44100, b2s_precision='fp16', mel_base='e')
    torch.manual_seed(0)
    model = P.DiffSingerAcoustic(len(VOCAB) + 1, 128)
    with torch.no_grad():
        torch.nn.init.normal_(model.diffusion.denoise_fn.output_projection.weight, std=0.01)
        for blk in model.aux_decoder.decoder.conv:
            blk.gamma.fill_(0.5)
    model = model.to(dev).eval()
    voc = P.NsfHifiGAN(P.vocoder.Generator(dict(bench.VOCODER_H)).to(dev).eval())
    inf = P.infer.DiffSingerAcousticInfer(model, voc, vocab_list=VOCAB, device=dev)
    params = project(n)
    audio_s = sum(P.segments.segment_frames(p, inf.timestep) for p in params) * inf.timestep
    out = {'segments': n, 'audio_seconds': round(audio_s, 1)}
    with tempfile.TemporaryDirectory() as tmp:
        for name, batched in (('batched', True), ('one_segment_per_call', False)):
            r = {}
            for it in range(3):              # pass 0: first sight of every shape (eager), pass 1: graph captures, pass 2: steady state (replays)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                entries = inf.infer_segments(params, batched=batched)
                torch.cuda.synchronize()
                t1 = time.perf_counter()
                wav = P.segments.vocode_segments(entries, voc, 44100, device=dev)
                torch.cuda.synchronize()
                t2 = time.perf_counter()
                P.segments.save_wav(wav, os.path.join(tmp, 'song.wav'), 44100)
                t3 = time.perf_counter()
                r[f'pass{it}'] = {'mel_s': round(t1 - t0, 3), 'vocode_s': round(t2 - t1, 3), 'wav_write_s': round(t3 - t2, 3),
                                  'total_s': round(t3 - t0, 3), 'rtf': round((t3 - t0) / audio_s, 6)}
            r['samples'], r['finite'] = int(wav.shape[0]), bool(np.isfinite(wav).all())
            out[name] = r
    print(json.dumps(out))


if __name__ == '__main__':
    main()
