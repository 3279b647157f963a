"""The reference's acoustic inference driver on the B200 modules: ``.ds`` segment -> model inputs -> mel -> waveform
(reference inference/ds_acoustic.py:23-246, basics/base_svs_infer.py:37-122, utils/infer_utils.py:41-53,
utils/text_encoder.py:7-32, modules/fastspeech/tts_modules.py:278-311).

``DiffSingerAcousticInfer`` keeps the reference's method names and the dictionaries they exchange - ``preprocess_input(param) -> batch``
(tokens, mel2ph, f0, variance curves, key_shift, speed, speaker mix), ``forward_model(batch) -> mel``, ``run_vocoder(mel, f0=)``,
``run_inference(params, out_dir, title, num_runs, seed, save_mel)`` - but takes its collaborators as objects (the acoustic model
``xiaoicesing_io_b200.DiffSingerAcoustic``, the vocoder ``xiaoicesing_io_b200.NsfHifiGAN``, the phoneme list, the speaker map)
instead of reading a work directory; checkpoint / dictionary / CLI handling stay with the reference (out of scope).

Two ways through ``run_inference``:
  * ``batched=False``: the reference's loop - one segment per call, reseeded per segment (ds_acoustic.py:209-219);
  * ``batched=True`` (default): the segment driver (segments.py) - conditions and x_start per segment from ``fs2`` / ``aux_decoder``,
    then the sampler over RAGGED batches of segments with per-segment seeded noise; every segment receives the bits of its own
    B = 1 run for the deterministic samplers, at several times the throughput.  Under ``torch.distributed`` (one process per GPU)
    the segments are partitioned across the ranks by length; rank 0 receives every mel, vocodes and writes, the others return None.
Host-side preprocessing is numpy / torch on small per-segment arrays (a few hundred tokens, a few thousand frames), as in the reference.
"""
from __future__ import annotations

import pathlib
from collections import OrderedDict
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import segments as S
from ._cabi import B2SError
from .hparams import hparams

VARIANCE_CHECKLIST = ['energy', 'breathiness', 'voicing', 'tension']      # modules/fastspeech/param_adaptor.py:10
PAD, PAD_INDEX = '<PAD>', 0                                                # utils/text_encoder.py:3-4


class TokenTextEncoder:
    """utils/text_encoder.py:7-44: ids are 1 + the index in the SORTED vocabulary, 0 is padding."""

    def __init__(self, vocab_list):
        self.vocab_list = sorted(vocab_list)
        self._index = {ph: i + 1 for i, ph in enumerate(self.vocab_list)}

    def encode(self, sentence):
        phones = sentence.strip().split() if isinstance(sentence, str) else sentence
        try:
            return [self._index[ph] if ph != PAD else PAD_INDEX for ph in phones]
        except KeyError as e:
            raise ValueError(f'{e.args[0]!r} is not in list') from None        # list.index's error, like the reference

    def decode(self, ids, strip_padding=False):
        if strip_padding:
            ids = np.trim_zeros(ids)
        return ' '.join(self.vocab_list[i - 1] if i >= 1 else PAD for i in list(ids))

    @property
    def vocab_size(self):
        return len(self.vocab_list) + 1

    def __len__(self):
        return self.vocab_size


def resample_align_curve(points: np.ndarray, original_timestep: float, target_timestep: float, align_length: int) -> np.ndarray:
    """utils/infer_utils.py:41-53: linear interpolation onto the frame grid, cut or padded with the last value to ``align_length``."""
    t_max = (len(points) - 1) * original_timestep
    curve = np.interp(np.arange(0, t_max, target_timestep), original_timestep * np.arange(len(points)), points).astype(points.dtype)
    delta = align_length - len(curve)
    if delta < 0:
        curve = curve[:align_length]
    elif delta > 0:
        curve = np.concatenate((curve, np.full(delta, fill_value=curve[-1])), axis=0)
    return curve


def length_regulator(dur: torch.Tensor, dur_padding: Optional[torch.Tensor] = None) -> torch.Tensor:
    """LengthRegulator.forward (tts_modules.py:278-311): ``dur [B, L]`` frames per token -> ``mel2ph [B, T]`` (1-based token index
    per frame, 0 = padding), T = the longest utterance.  Same result as the reference's [B, L, T] mask, built by repeat_interleave."""
    dur = dur.long()
    if dur_padding is not None:
        dur = dur * (1 - dur_padding.long())
    B, L = dur.shape
    T = int(dur.sum(-1).max()) if B else 0
    mel2ph = torch.zeros((B, T), dtype=torch.long, device=dur.device)
    tok = torch.arange(1, L + 1, device=dur.device)
    for b in range(B):
        row = torch.repeat_interleave(tok, dur[b].clamp(min=0))
        mel2ph[b, :row.numel()] = row
    return mel2ph


class DiffSingerAcousticInfer:
    """inference/ds_acoustic.py:23-246 on the B200 modules."""

    def __init__(self, model, vocoder=None, *, vocab_list: Sequence[str], spk_map: Optional[Dict[str, int]] = None, device=None):
        self.model = model
        self.vocoder = vocoder
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        self.timestep = hparams['hop_size'] / hparams['audio_sample_rate']                   # base_svs_infer.py:26
        self.variances_to_embed = {v for v in VARIANCE_CHECKLIST if hparams.get(f'use_{v}_embed', False)}   # ds_acoustic.py:30-39
        self.ph_encoder = TokenTextEncoder(vocab_list=vocab_list)
        self.spk_map = spk_map
        if hparams.get('use_spk_id', False):
            assert isinstance(spk_map, dict) and len(spk_map) > 0, 'Invalid or empty speaker map!'
            assert len(spk_map) == len(set(spk_map.values())), 'Duplicate speaker id in speaker map!'

    # ---- ds segment -> model inputs ----------------------------------------------------------------------------------------------
    def load_speaker_mix(self, param_src: dict, summary_dst: dict, mix_mode: str = 'frame', mix_length: int = None
                         ) -> Tuple[torch.Tensor, torch.Tensor]:
        """basics/base_svs_infer.py:37-122 -> ``spk_mix_id [1, 1, N]``, ``spk_mix_value [1, T or 1, N]`` (normalised)."""
        assert mix_mode == 'token' or mix_mode == 'frame'
        param_key = 'spk_mix' if mix_mode == 'frame' else 'ph_spk_mix'
        solo_key = 'spk' if mix_mode == 'frame' else 'ph_spk'
        mix = param_src.get(param_key)
        if mix is None:
            mix = {next(iter(self.spk_map.keys())): 1.0}
        else:
            for name in mix:
                assert name in self.spk_map, f'Speaker \'{name}\' not found.'
        dynamic = False
        if len(mix) == 1:
            summary_dst[solo_key] = list(mix.keys())[0]
        elif any(isinstance(v, str) for v in mix.values()):
            summary_dst[param_key] = f'dynamic({"|".join(mix.keys())})'
            dynamic = True
        else:
            summary_dst[param_key] = 'static(' + '|'.join(f'{n}:{"%.3f" % mix[n]}' for n in mix) + ')'
        ids, values = [], []
        if dynamic:
            for name, v in mix.items():
                ids.append(self.spk_map[name])
                if isinstance(v, str):
                    if mix_mode == 'token':
                        cur = v.split()
                        assert len(cur) == mix_length, ('Speaker mix checks failed. In dynamic token-level mix, '
                                                        'number of proportion values must equal number of tokens.')
                        cur = torch.from_numpy(np.array(cur, 'float32')).to(self.device)[None]
                    else:
                        cur = torch.from_numpy(resample_align_curve(np.array(v.split(), 'float32'),
                                                                    original_timestep=float(param_src['spk_mix_timestep']),
                                                                    target_timestep=self.timestep, align_length=mix_length)).to(self.device)[None]
                    assert torch.all(cur >= 0.), f'Speaker mix checks failed.\nProportions of speaker \'{name}\' on some {mix_mode}s are negative.'
                else:
                    assert v >= 0., f'Speaker mix checks failed.\nProportion of speaker \'{name}\' is negative.'
                    cur = torch.full((1, mix_length), fill_value=v, dtype=torch.float32, device=self.device)
                values.append(cur)
            spk_mix_id = torch.LongTensor(ids).to(self.device)[None, None]
            spk_mix_value = torch.stack(values, dim=2)
            total = torch.sum(spk_mix_value, dim=2, keepdim=True)
            assert torch.all(total > 0.), 'Speaker mix checks failed.\nProportions of speaker mix on some frames sum to zero.'
            spk_mix_value = spk_mix_value / total
        else:
            for name, v in mix.items():
                ids.append(self.spk_map[name])
                assert v >= 0., f'Speaker mix checks failed.\nProportion of speaker \'{name}\' is negative.'
                values.append(v)
            spk_mix_id = torch.LongTensor(ids).to(self.device)[None, None]
            spk_mix_value = torch.FloatTensor(values).to(self.device)[None, None]
            total = spk_mix_value.sum()
            assert total > 0., 'Speaker mix checks failed.\nProportions of speaker mix sum to zero.'
            spk_mix_value = spk_mix_value / total
        return spk_mix_id, spk_mix_value

    def _curve(self, param, seq_key, timestep_key, length):
        return torch.from_numpy(resample_align_curve(np.array(param[seq_key].split(), np.float32), original_timestep=float(param[timestep_key]),
                                                     target_timestep=self.timestep, align_length=length)).to(self.device)[None]

    def preprocess_input(self, param: dict, idx: int = 0, verbose: bool = False) -> dict:
        """ds_acoustic.py:68-166: one ``.ds`` segment -> the batch dict of the model inputs (B = 1)."""
        batch, summary = {}, OrderedDict()
        txt_tokens = torch.LongTensor([self.ph_encoder.encode(param['ph_seq'])]).to(self.device)                 # :76
        batch['tokens'] = txt_tokens
        ph_dur = torch.from_numpy(np.array(param['ph_dur'].split(), np.float32)).to(self.device)                  # :79
        ph_acc = torch.round(torch.cumsum(ph_dur, dim=0) / self.timestep + 0.5).long()                           # :80
        durations = torch.diff(ph_acc, dim=0, prepend=torch.LongTensor([0]).to(self.device))[None]               # :81
        mel2ph = length_regulator(durations, txt_tokens == 0)                                                    # :82
        batch['mel2ph'] = mel2ph
        length = mel2ph.size(1)
        summary['tokens'], summary['frames'], summary['seconds'] = txt_tokens.size(1), length, '%.2f' % (length * self.timestep)
        if hparams.get('use_spk_id', False):                                                                     # :90-95
            batch['spk_mix_id'], batch['spk_mix_value'] = self.load_speaker_mix(param, summary, 'frame', length)
        batch['f0'] = self._curve(param, 'f0_seq', 'f0_timestep', length)                                                           # :97-102
        for v_name in VARIANCE_CHECKLIST:                                                                        # :104-112
            if v_name in self.variances_to_embed:
                batch[v_name] = self._curve(param, v_name, f'{v_name}_timestep', length)
                summary[v_name] = 'manual'
        if hparams.get('use_key_shift_embed', False):                                                            # :114-136
            shift_min, shift_max = hparams['augmentation_args']['random_pitch_shifting']['range']
            gender = param.get('gender')
            if gender is None:
                gender = 0.
            if isinstance(gender, (int, float, bool)):
                summary['gender'] = f'static({gender:.3f})'
                value = gender * shift_max if gender >= 0 else gender * abs(shift_min)
                batch['key_shift'] = torch.FloatTensor([value]).to(self.device)[:, None]
            else:
                summary['gender'] = 'dynamic'
                g = resample_align_curve(np.array(gender.split(), np.float32), original_timestep=float(param['gender_timestep']),
                                         target_timestep=self.timestep, align_length=length)
                mask = g >= 0
                seq = g * (mask * shift_max + (1 - mask) * abs(shift_min))
                batch['key_shift'] = torch.clip(torch.from_numpy(seq.astype(np.float32)).to(self.device)[None], min=shift_min, max=shift_max)
        if hparams.get('use_speed_embed', False):                                                                # :138-156
            if param.get('velocity') is None:
                summary['velocity'] = 'default'
                batch['speed'] = torch.FloatTensor([1.]).to(self.device)[:, None]
            else:
                summary['velocity'] = 'manual'
                speed_min, speed_max = hparams['augmentation_args']['random_time_stretching']['range']
                sp = resample_align_curve(np.array(param['velocity'].split(), np.float32), original_timestep=float(param['velocity_timestep']),
                                          target_timestep=self.timestep, align_length=length)
                batch['speed'] = torch.clip(torch.from_numpy(sp.astype(np.float32)).to(self.device)[None], min=speed_min, max=speed_max)
        if verbose:
            print(f'[{idx}]\t' + ', '.join(f'{k}: {v}' for k, v in summary.items()))
        return batch

    # ---- model inputs -> mel -> waveform ---------------------------------------------------------------------------------------------
    def _spk_mix_embed(self, sample):
        if not hparams.get('use_spk_id', False):
            return None
        return torch.sum(self.model.fs2.spk_embed(sample['spk_mix_id']) * sample['spk_mix_value'].unsqueeze(3), dim=2, keepdim=False)   # :177-181

    def _model_kwargs(self, sample):
        kw = {v: sample.get(v) for v in self.variances_to_embed}
        kw.update(key_shift=sample.get('key_shift'), speed=sample.get('speed'), spk_mix_embed=self._spk_mix_embed(sample))
        return kw

    @torch.no_grad()
    def forward_model(self, sample):
        """ds_acoustic.py:168-190: ``model(tokens, mel2ph, f0, variances..., infer=True).diff_out`` - mel [1, T, M]."""
        return self.model(sample['tokens'], mel2ph=sample['mel2ph'], f0=sample['f0'], infer=True, **self._model_kwargs(sample)).diff_out

    @torch.no_grad()
    def run_vocoder(self, spec, **kwargs):
        """ds_acoustic.py:192-195."""
        if self.vocoder is None:
            raise B2SError('no vocoder: pass one to DiffSingerAcousticInfer or use save_mel=True')
        return self.vocoder.spec2wav_torch(spec, **kwargs)[None]

    @torch.no_grad()
    def _cond_fn(self, batches):
        """For the segment driver: condition [T, H] and x_start [T, M] of one segment from the encoder / aux decoder (the stages before
        the sampler in DiffSingerAcoustic.forward, modules/toplevel.py:88-96), f0 [T]."""
        index = {id(p): b for p, b in batches}

        def cond_fn(param, frames):
            sample = index[id(param)]
            condition = self.model.fs2(sample['tokens'], sample['mel2ph'], sample['f0'], **self._model_kwargs(sample))
            src = self.model.aux_decoder(condition, infer=True)[0] if getattr(self.model, 'use_shallow_diffusion', False) else None
            return condition[0], src, sample['f0'][0]
        return cond_fn

    @torch.no_grad()
    def infer_segments(self, params: Sequence[dict], seed: int = -1, batched: bool = True, **driver_kw):
        """``params`` -> the ``.mel.pt`` entries ``{'offset', 'mel' [1, T, M] (CPU), 'f0' [1, T] (CPU)}`` in segment order."""
        batches = [self.preprocess_input(p, idx=i) for i, p in enumerate(params)]
        if batched:
            import torch.distributed as dist
            cond_fn = self._cond_fn(list(zip(params, batches)))
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                # one process per GPU: the segments are partitioned by length, every rank runs the producers and the sampler for ITS
                # segments only (no collective inside the loop), the finished entries are gathered once on rank 0; None elsewhere
                res = S.sample_segments_distributed(self.model.diffusion, params, cond_fn, self.timestep, self.device, seed=seed, **driver_kw)
                if res is None:
                    return None
            else:
                got = S.sample_segments(self.model.diffusion, params, cond_fn, self.timestep, self.device, seed=seed, **driver_kw)
                res = [got[i] for i in range(len(params))]
            return [dict(offset=res[i]['offset'], mel=res[i]['mel'], f0=batches[i]['f0'].cpu()) for i in range(len(params))]
        out = []
        for param, batch in zip(params, batches):
            s = S.segment_seed(param, seed)                                                                       # :212-217
            if s is not None:
                torch.manual_seed(s)
                torch.cuda.manual_seed_all(s)
            out.append(dict(offset=param.get('offset', 0.), mel=self.forward_model(batch).cpu(), f0=batch['f0'].cpu()))
        return out

    def run_inference(self, params, out_dir: pathlib.Path = None, title: str = None, num_runs: int = 1, seed: int = -1,
                      save_mel: bool = False, batched: bool = True):
        """ds_acoustic.py:197-246: writes ``<title>.wav`` (or ``<title>.mel.pt``; ``<title>-000...`` for several runs) and returns the
        last run's result (the waveform as float64 numpy, or the ``.mel.pt`` list)."""
        out_dir = pathlib.Path(out_dir)
        out_dir.mkdir(parents=True, exist_ok=True)
        suffix = '.wav' if not save_mel else '.mel.pt'
        result = None
        for i in range(num_runs):
            entries = self.infer_segments(params, seed=seed, batched=batched)
            if entries is None:                                  # a rank other than 0 of a multi-GPU run: rank 0 vocodes and writes
                continue
            path = out_dir / (f'{title}-{str(i).zfill(3)}{suffix}' if num_runs > 1 else title + suffix)
            if save_mel:
                result = entries
                print(f'| save mel: {path}')
                S.save_mel_pt(path, entries)
            else:
                result = S.vocode_segments(entries, self.vocoder, hparams['audio_sample_rate'], device=self.device)
                print(f'| save audio: {path}')
                S.save_wav(result, path, hparams['audio_sample_rate'])
        return result
