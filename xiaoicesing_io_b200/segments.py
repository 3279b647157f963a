"""Batched segment driver: what feeds the fast sampling path in real use.

The reference synthesises a ``.ds`` project ONE SEGMENT AT A TIME (inference/ds_acoustic.py:189-246: B = 1 per segment, reseeded
per segment at :212-217, results collected as ``{'offset', 'mel', 'f0'}`` and saved with ``torch.save`` as ``<title>.mel.pt``
at :220-225, :243).  One 8-second segment occupies 6 of the 148 SMs; this module

  1. reads the ``.ds`` JSON (scripts/infer.py:129-133: a list of segment dicts, or one dict) and derives every segment's frame
     count exactly as ``preprocess_input`` does (ds_acoustic.py:79-83: ``round(cumsum(ph_dur) / timestep + 0.5)``);
  2. buckets the segments by length into RAGGED batches - padded to a multiple of 128 frames (one tile of the whole-stack kernel,
     so captured CUDA graphs are re-used across calls) under a frame budget;
  3. runs each batch through the sampler with ``lengths=`` (frames beyond a segment's length are the conv's zero padding, exactly
     as in a batch of its own) and ``initial_noise=`` drawn PER SEGMENT from the segment's seed with the reference's own reseeding
     rule - so every segment receives the bits of its own B = 1 run (deterministic samplers: DDIM / PNDM / DPM-Solver++ / UniPC /
     rectified flow; the ancestral sampler draws fresh noise per step from the global stream);
  4. on several GPUs partitions the segments by length (``partition_by_length``), every rank sampling its share with no collective,
     and gathers the results once;
  5. writes the reference's ``.mel.pt`` layout;
  6. vocodes a ``.mel.pt`` list into one waveform (``vocode_segments``: scripts/vocode.py:64-84, ds_acoustic.py:227-236 - silence up
     to a segment's offset, linear cross-fade where segments overlap) and writes the 16-bit WAV.

The linguistic encoder that turns a segment into ``condition [T, H]`` (modules/fastspeech) is outside the
hot path: the caller passes ``cond_fn(segment, frames) -> (condition [T, H], src_spec [T, M] or None, f0 [T] or None)``.
"""
from __future__ import annotations

import json
import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .partition import partition_by_length

TILE = 128     # frames per tile of the whole-stack kernel: ragged batches are padded to a multiple of it
GRAPH_BUCKETS = 24   # CUDA graphs kept while a project is sampled: one per padded length (30 s = 21 tiles), see sample_segments


def load_ds(path) -> List[dict]:
    """A ``.ds`` file: a JSON list of segment dicts or a single dict (scripts/infer.py:129-133)."""
    with open(path, 'r', encoding='utf-8') as f:
        params = json.load(f)
    if not isinstance(params, list):
        params = [params]
    if not params:
        raise ValueError(f'{path}: the project has no segments')
    return params


def segment_frames(param: dict, timestep: float) -> int:
    """Frames of one segment, as ``preprocess_input`` derives them (ds_acoustic.py:79-83): the last entry of
    ``round(cumsum(ph_dur) / timestep + 0.5)``, computed in fp32 like the reference."""
    ph_dur = torch.tensor([float(v) for v in str(param['ph_dur']).split()], dtype=torch.float32)
    if ph_dur.numel() == 0:
        return 0
    ph_acc = torch.round(torch.cumsum(ph_dur, dim=0) / timestep + 0.5).long()
    return int(ph_acc[-1])


def segment_seed(param: dict, seed: int = -1) -> Optional[int]:
    """The seed the reference would set before this segment (ds_acoustic.py:212-217), or None (no reseeding)."""
    if 'seed' in param:
        return int(param['seed']) & 0xffff_ffff
    if seed >= 0:
        return seed & 0xffff_ffff
    return None


def plan_batches(lengths: Sequence[int], max_batch_frames: int = 16 * 704, max_batch_size: int = 64,
                 tile: int = TILE) -> List[List[int]]:
    """Groups segment indices into ragged batches: longest first, a batch is padded to ``ceil(longest / tile) * tile`` frames and
    grows while ``size * padded_T <= max_batch_frames`` and ``size <= max_batch_size``.  Zero-length segments are dropped.
    Deterministic; every non-empty segment appears in exactly one batch."""
    order = sorted((i for i in range(len(lengths)) if int(lengths[i]) > 0), key=lambda i: (-int(lengths[i]), i))
    batches: List[List[int]] = []
    cur: List[int] = []
    cur_T = 0
    for i in order:
        if not cur:
            cur, cur_T = [i], -(-int(lengths[i]) // tile) * tile
            continue
        if len(cur) < max_batch_size and (len(cur) + 1) * cur_T <= max_batch_frames:
            cur.append(i)
        else:
            batches.append(cur)
            cur, cur_T = [i], -(-int(lengths[i]) // tile) * tile
    if cur:
        batches.append(cur)
    return batches


_GENERATORS: Dict[str, torch.Generator] = {}


def seeded_noise(shape: Tuple[int, ...], seed: Optional[int], device) -> torch.Tensor:
    """The first noise draw of a B = 1 run of the reference on ``device`` after its per-segment reseeding:
    ``torch.manual_seed(seed); torch.cuda.manual_seed_all(seed); randn(1, F, M, T, device=device)`` (ds_acoustic.py:212-217,
    ddpm.py:227).  Drawn from a private generator seeded the same way (same Philox seed, offset 0 -> the same bits; checked by
    tests/test_gpu_segments.py) so that a batch of segments does not reseed the process-wide generators once per segment (two
    global reseeds cost more host time than the draw).  ``seed`` None: drawn from the current global stream."""
    if seed is None:
        return torch.randn(shape, device=device)
    device = torch.device(device)
    g = _GENERATORS.get(str(device))
    if g is None:
        g = _GENERATORS[str(device)] = torch.Generator(device=device)
    g.manual_seed(seed)
    return torch.randn(shape, device=device, generator=g)


CondFn = Callable[[dict, int], Tuple[torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor]]]


MAX_PENDING = 16     # batches in flight (enqueued or finished) whose pinned host copies have not been handed out yet


def _drain(pending, params, out) -> None:
    for seg, lens, mel_h, f0s, done in pending:
        done.synchronize()                                     # this batch's device-to-host copy has landed
        for k, i in enumerate(seg):
            out[i] = {'offset': float(params[i].get('offset', 0.)), 'mel': mel_h[k:k + 1, :lens[k]].clone(),
                      'f0': None if f0s[k] is None else f0s[k].reshape(1, -1).float().cpu()}


@torch.no_grad()
def sample_segments(model, params: Sequence[dict], cond_fn: CondFn, timestep: float, device, seed: int = -1,
                    max_batch_frames: int = 16 * 704, max_batch_size: int = 64, indices: Optional[Sequence[int]] = None
                    ) -> Dict[int, dict]:
    """Runs the acoustic sampler over the segments ``indices`` (default: all) in ragged batches.  Returns
    ``{segment index: {'offset', 'mel' [1, T, M] (CPU), 'f0' [1, T] or None}}`` - the entries the reference appends to its
    ``.mel.pt`` list (ds_acoustic.py:220-225)."""
    device = torch.device(device)
    idx = list(range(len(params))) if indices is None else list(indices)
    frames = {i: segment_frames(params[i], timestep) for i in idx}
    out: Dict[int, dict] = {}
    F_, M = model.num_feats, model.out_dims
    pending = []          # (segments, lengths, mel on the device, f0s): results stay on the device until every batch is enqueued,
    #                       so the host prepares batch i + 1 (conditions, seeded noise) while the GPU samples batch i
    from .hparams import hparams
    saved_cap = hparams.get('b2s_graph_cache')
    hparams['b2s_graph_cache'] = max(GRAPH_BUCKETS, int(saved_cap or 0))
    try:
        for batch in plan_batches([frames[i] for i in idx], max_batch_frames, max_batch_size):
            seg = [idx[j] for j in batch]
            lens = [frames[i] for i in seg]
            T = -(-max(lens) // TILE) * TILE
            # ONE batch shape per padded length: the batch always has the capacity of its bucket, unused slots are utterances of
            # length 0 (lengths= makes them the conv's zero padding; a launch holds every tile of the batch anyway, so they cost no
            # time) - a project of any size replays at most one captured graph per bucket instead of capturing one per (B, T)
            cap = max(len(seg), min(max_batch_size, max(1, max_batch_frames // T)))
            conds, srcs, f0s = [], [], []
            for i in seg:
                c, s, f0 = cond_fn(params[i], frames[i])
                if c.shape[0] != frames[i]:
                    raise ValueError(f'segment {i}: cond_fn returned {c.shape[0]} frames, the segment has {frames[i]}')
                conds.append(c)
                srcs.append(s)
                f0s.append(f0)
            H = conds[0].shape[1]
            condition = torch.zeros((cap, T, H), device=device)
            noise = torch.zeros((cap, F_, M, T), device=device)
            src = None
            if any(s is not None for s in srcs):
                src = torch.zeros((cap, T, M), device=device)
            for k, i in enumerate(seg):
                condition[k, :lens[k]] = conds[k].to(device, non_blocking=True)
                noise[k, ..., :lens[k]] = seeded_noise((1, F_, M, lens[k]), segment_seed(params[i], seed), device)[0]
                if src is not None and srcs[k] is not None:
                    src[k, :lens[k]] = srcs[k].to(device, non_blocking=True)
            lengths = torch.tensor(lens + [0] * (cap - len(seg)), dtype=torch.int32)
            mel = model(condition, src_spec=src, infer=True, lengths=lengths, initial_noise=noise).float()
            # asynchronous copy into pinned host memory, stream-ordered behind this batch: the host never waits for a batch here
            host = torch.empty(mel.shape, dtype=torch.float32, pin_memory=True)
            host.copy_(mel, non_blocking=True)
            pending.append((seg, lens, host, f0s, torch.cuda.current_stream(device).record_event()))
            if len(pending) >= 2 * MAX_PENDING:                # bound the device memory held by finished batches
                _drain(pending[:MAX_PENDING], params, out)
                del pending[:MAX_PENDING]
    finally:
        if saved_cap is None:
            hparams.pop('b2s_graph_cache', None)
        else:
            hparams['b2s_graph_cache'] = saved_cap
    _drain(pending, params, out)
    for i in idx:
        if i not in out:                                        # zero-length segment
            out[i] = {'offset': float(params[i].get('offset', 0.)), 'mel': torch.zeros((1, 0, M)), 'f0': torch.zeros((1, 0))}
    return out


@torch.no_grad()
def sample_segments_distributed(model, params: Sequence[dict], cond_fn: CondFn, timestep: float, device, seed: int = -1,
                                group=None, dst: int = 0, **kw) -> Optional[List[dict]]:
    """``sample_segments`` over the ranks of ``torch.distributed``: segments are partitioned by length (no collective inside the
    loop), the finished entries are gathered ONCE on rank ``dst`` (``gather_object``: variable-length results, off the hot loop)
    and returned there in segment order; None elsewhere.  Single process: the same list."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        res = sample_segments(model, params, cond_fn, timestep, device, seed, **kw)
        return [res[i] for i in range(len(params))]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    parts = partition_by_length([segment_frames(p, timestep) for p in params], world)
    mine = sample_segments(model, params, cond_fn, timestep, device, seed, indices=parts[rank], **kw)
    gathered = [None] * world if rank == dst else None
    dist.gather_object(mine, gathered, dst=dst, group=group)
    if rank != dst:
        return None
    merged: Dict[int, dict] = {}
    for g in gathered:
        merged.update(g)
    return [merged[i] for i in range(len(params))]


def save_mel_pt(path, entries: Sequence[dict]) -> None:
    """The reference's ``.mel.pt`` layout (ds_acoustic.py:220-225, :243): ``torch.save`` of the list of
    ``{'offset': float, 'mel': [1, T, M], 'f0': [1, T]}`` in segment order (read back by scripts/vocode.py)."""
    torch.save([{'offset': e['offset'], 'mel': e['mel'], 'f0': e['f0']} for e in entries], path)


def real_time_factor(entries: Sequence[dict], seconds: float, timestep: float) -> float:
    """Wall seconds per second of synthesised audio."""
    audio = sum(e['mel'].shape[1] for e in entries) * timestep
    return math.inf if audio == 0 else seconds / audio


def cross_fade(a: np.ndarray, b: np.ndarray, idx: int) -> np.ndarray:
    """utils/infer_utils.py:89-96: ``b`` starts at sample ``idx`` of ``a``; the overlap is a linear fade from ``a`` to ``b``."""
    result = np.zeros(idx + b.shape[0])
    fade_len = a.shape[0] - idx
    result[:idx] = a[:idx]
    k = np.linspace(0, 1.0, num=fade_len, endpoint=True)
    result[idx: a.shape[0]] = (1 - k) * a[idx:] + k * b[:fade_len]
    result[a.shape[0]:] = b[fade_len:]
    return result


def vocode_segments(entries: Sequence[dict], vocoder, sample_rate: int, device=None) -> np.ndarray:
    """``.mel.pt`` entries ``{'offset', 'mel' [1, T, M], 'f0' [1, T]}`` -> one waveform (float64 numpy, like the reference's):
    every segment through ``vocoder.spec2wav_torch(mel, f0=f0)`` (one call per segment, ds_acoustic.py:185-187), placed at
    ``round(offset * sample_rate)`` with silence before it or a linear cross-fade into the previous segment (scripts/vocode.py:64-84,
    ds_acoustic.py:227-236, utils/infer_utils.py:89-96).

    Same samples as the reference's loop, without its quadratic cost: the reference grows the result with ``np.append`` / a fresh
    ``cross_fade`` array per segment (two copies of everything so far: 2.5 s of host time for 48 segments / 12 minutes of audio); here
    the segments are vocoded first (device-to-host copies queued behind the kernels), then written once into a preallocated array with
    the identical fade arithmetic.  Segment lengths rarely repeat, so CUDA-graph capture is switched off for the loop when the project
    has more distinct lengths than the vocoder's graph cache holds (a capture costs more than the launches it would save)."""
    from .hparams import hparams
    device = device if device is not None else vocoder.device
    distinct = len({int(e['mel'].shape[1]) for e in entries})
    saved = hparams.get('b2s_cuda_graph', None)
    if distinct > 4:
        hparams['b2s_cuda_graph'] = False
    try:
        wavs = []
        for e in entries:
            y = vocoder.spec2wav_torch(e['mel'].to(device), f0=e['f0'].to(device))
            host = torch.empty(y.shape, dtype=y.dtype, pin_memory=True) if y.is_cuda else y
            if y.is_cuda:
                host.copy_(y, non_blocking=True)
            wavs.append(host)
        if any(w.is_pinned() for w in wavs):
            torch.cuda.synchronize(device)
    finally:
        if distinct > 4:
            if saved is None:
                hparams.pop('b2s_cuda_graph', None)
            else:
                hparams['b2s_cuda_graph'] = saved
    # placement: a segment always starts at round(offset * sr) (silence before it, or a fade from there to the end of what exists)
    starts = [round(e.get('offset', 0.) * sample_rate) for e in entries]
    ends, cur, ordered = [], 0, True
    for s0, w in zip(starts, wavs):
        ordered = ordered and s0 >= 0 and (s0 >= cur or s0 + w.shape[0] >= cur)
        cur = s0 + w.shape[0]
        ends.append(cur)
    if not ordered:                                   # a segment nested inside its predecessor: the reference's literal loop (it raises)
        result, current_length = np.zeros(0), 0
        for s0, w in zip(starts, wavs):
            w = w.numpy()
            silent_length = s0 - current_length
            if silent_length >= 0:
                result = np.append(np.append(result, np.zeros(silent_length)), w)
            else:
                result = cross_fade(result, w, current_length + silent_length)
            current_length = current_length + silent_length + w.shape[0]
        return result
    result = np.zeros(ends[-1] if ends else 0)
    cur = 0
    for s0, w in zip(starts, wavs):
        w = w.numpy()
        if s0 >= cur:
            result[s0:s0 + w.shape[0]] = w
        else:
            fade_len = cur - s0
            k = np.linspace(0, 1.0, num=fade_len, endpoint=True)
            result[s0:cur] = (1 - k) * result[s0:cur] + k * w[:fade_len]
            result[cur:s0 + w.shape[0]] = w[fade_len:]
        cur = s0 + w.shape[0]
    return result


def save_wav(wav: np.ndarray, path, sr: int, norm: bool = False) -> None:
    """utils/infer_utils.py:99-104: 16-bit PCM (``wav * 32767`` truncated to int16)."""
    from scipy.io import wavfile
    wav = np.array(wav, dtype=np.float64)
    if norm:
        wav = wav / np.abs(wav).max()
    wavfile.write(str(path), sr, (wav * 32767).astype(np.int16))

