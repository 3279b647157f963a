"""Shared base of GaussianDiffusion / RectifiedFlow on the B200 path: spec (de)normalisation, the
``forward(condition, gt_spec, src_spec, infer)`` entry, buffer + noise plumbing and the program run."""
from __future__ import annotations

from collections import OrderedDict
from typing import Callable, Optional

import torch
from torch import nn

from .. import _cabi as C
from ..backbones.wavenet import _time_major_cond
from ..engine import CompiledProgram, allocate_buffers, run_program
from ..hparams import hparams
from ..schedules import NOISE0, XSTART, Program


def _bounds(values, out_dims):
    """spec_min / spec_max list -> [1, 1, M] (one feature) or [1, F, 1, M] (ddpm.py:103-109)."""
    return torch.FloatTensor(values)[None, None, :out_dims].transpose(-3, -2)


class SamplerBase(nn.Module):
    """Holds the backbone under ``backbone_attr`` ('denoise_fn' / 'velocity_fn') and runs programs."""

    backbone_attr = 'denoise_fn'

    def _init_common(self, out_dims, num_feats, spec_min, spec_max, persistent_bounds):
        self.out_dims = out_dims
        self.num_feats = num_feats
        self.register_buffer('spec_min', _bounds(spec_min, out_dims), persistent=persistent_bounds)
        self.register_buffer('spec_max', _bounds(spec_max, out_dims), persistent=persistent_bounds)
        self._noise_source = None               # test hook: callable(shape) -> N(0,1) tensor [B,F,M,T]

    # ---- spec normalisation (ddpm.py:379-383, reflow.py:140-144) -------------------------------------
    def norm_spec(self, x):
        return (x - self.spec_min) / (self.spec_max - self.spec_min) * 2 - 1

    def denorm_spec(self, x):
        return (x + 1) / 2 * (self.spec_max - self.spec_min) + self.spec_min

    def _source_to_state(self, src_spec):
        """src_spec [B,T,M] / [B,F,T,M] (or curves) -> normalised [B,F,M,T], or None."""
        if src_spec is None:
            return None
        spec = self.norm_spec(src_spec).transpose(-2, -1)
        return spec[:, None] if self.num_feats == 1 else spec

    # ---- program execution -------------------------------------------------------------------------
    def build_program(self) -> Program:                      # pragma: no cover - abstract
        raise NotImplementedError

    def _run(self, cond, b, start, device, lengths=None, initial_noise=None, denorm=None):
        return sample(getattr(self, self.backbone_attr), self.build_program(), cond, b, self.num_feats,
                      self.out_dims, start, device, noise_source=self._noise_source, lengths=lengths, noise0=initial_noise,
                      denorm=denorm)

    def _fused_norm(self) -> bool:
        """True when (de)normalisation is exactly the base expression (no repeat-bin / clip / multi-curve mixin): then norm_spec,
        denorm_spec and the layout changes around the sampler run as ONE kernel each (b2s_spec_norm_f32 / b2s_spec_denorm_f32)."""
        return type(self).norm_spec is SamplerBase.norm_spec and type(self).denorm_spec is SamplerBase.denorm_spec

    def _training_forward(self, spec, cond, b, device):     # pragma: no cover - abstract
        raise NotImplementedError

    def forward(self, condition, gt_spec=None, src_spec=None, infer=True, lengths=None, initial_noise=None):
        """condition [B, T, H] -> de-normalised sample [B, T, M] / [B, F, T, M] (or curves).

        Two extensions over the reference's signature (both default to its behaviour), used by the batched segment driver
        (``xiaoicesing_io_b200.segments``): ``lengths`` [B] - the batch is RAGGED, utterance b has lengths[b] <= T valid frames and
        every valid frame gets the bits it would get in a batch of its own (frames beyond are undefined); ``initial_noise``
        [B, F, M, T] - the first noise draw (ddpm.py:227, reflow.py:105), e.g. drawn per segment from the segment's seed."""
        cond = condition.transpose(1, 2)                     # [B, H, T] view, as the reference hands it on
        b, device = condition.shape[0], condition.device
        if infer:
            kw = {}
            if lengths is not None:
                kw['lengths'] = lengths
            if initial_noise is not None:
                kw['initial_noise'] = initial_noise
            if self._fused_norm() and condition.is_cuda and type(self).inference in _BASE_INFERENCE:
                # acoustic models: norm_spec + transpose and transpose + denorm_spec fused into one kernel each
                # the registered bounds are [1, 1, K] / [1, F, 1, K] with K = 1 (one value for all bins) or M: flat [F*M] for the kernel
                flat = lambda v: v.reshape(self.num_feats, v.shape[-1]).expand(self.num_feats, self.out_dims).contiguous().reshape(-1)
                lo, hi = flat(self.spec_min), flat(self.spec_max)
                start = None
                if src_spec is not None:
                    T, M, F_ = condition.shape[1], self.out_dims, self.num_feats
                    start = TimeMajorState(torch.empty((b * T, F_ * M), device=device))
                    with torch.cuda.device(device):
                        C.spec_norm(src_spec.to(device=device, dtype=torch.float32).contiguous(), lo, hi, start.t, b, F_, T, M)
                return self._run(cond, b, start, device, denorm=(lo, hi), **kw)
            x = self.inference(cond, b, self._source_to_state(src_spec), device, **kw)
            return self.denorm_spec(x)
        return self._training_forward(self._source_to_state(gt_spec), cond, b, device)


class TimeMajorState:
    """A normalised start state that is ALREADY in the sampler's time-major layout [B*T, F*M] (b2s_spec_norm_f32 wrote it): the
    reference's [B, F, M, T] round trip (ddpm.py:370-373) is skipped."""
    def __init__(self, t: torch.Tensor):
        self.t = t


_BASE_INFERENCE: set = set()     # the stock ``inference`` methods of GaussianDiffusion / RectifiedFlow (filled by core/__init__)


class _GraphedLoop:
    """The WHOLE sampling call - hoisted cond / step tables, initial noise draw, every denoiser evaluation,
    every sampler update and per-step noise draw - captured once as a CUDA graph over static buffers and
    replayed with a single launch.  Noise comes from ``torch.randn`` inside the graph (graph-safe Philox:
    a replay consumes the default generator exactly like the eager loop, so seeded runs stay reproducible)."""

    def __init__(self, eng, cp: CompiledProgram, B, T, H, F_, M, device, ragged=False, ext_noise=False, start_tm=False):
        prog = cp.prog
        self.cp = cp                     # the coefficient / model-time tables are read by the captured kernels
        self.cond = torch.empty((B, T, H), device=device)
        self.start_tm = start_tm         # x_start arrives time-major [B*T, F*M] (fused norm_spec): a copy instead of a transpose
        self.x_start = (torch.empty((B * T, F_ * M) if start_tm else (B, F_ * M, T), device=device)) if prog.needs_x_start else None
        self.lens = torch.full((B,), T, device=device, dtype=torch.int32) if ragged else None      # static inputs of the graph
        self.noise0 = torch.empty((B, F_ * M, T), device=device) if ext_noise else None
        shape = (B, F_ * M, T)

        def body():
            sess = eng.begin(self.cond, cp.t_values, per_row_t=False, lens=self.lens, table_owner=cp)
            bufs = allocate_buffers(prog, B * T, F_ * M, device)

            def load(src, dst):
                C.transpose(src, dst, B, F_ * M, T)

            noise0 = self.noise0 if ext_noise else torch.randn(shape, device=device)
            if NOISE0 in bufs:
                load(noise0, bufs[NOISE0])
            if prog.needs_x_start:
                if start_tm:
                    bufs[XSTART].copy_(self.x_start)
                else:
                    load(self.x_start, bufs[XSTART])
            self.keep = (sess, bufs)     # every buffer a captured kernel touches stays owned by this object
            return run_program(cp, sess, bufs, lambda j, dst: load(torch.randn(shape, device=device), dst))

        self.graph = torch.cuda.CUDAGraph()
        n0 = C.N_CALLS
        with torch.cuda.graph(self.graph):
            self.out = body()
        self.n_launches = C.N_CALLS - n0        # kernels of libb2s.so in one replay (torch.randn nodes not counted)

    def run(self, cond_bth, x_start_bfmt, lens=None, noise0=None):
        # (the caller holds torch.cuda.device(device): replay goes to the device the graph was captured on)
        self.cond.copy_(cond_bth)
        if self.lens is not None:
            self.lens.copy_(lens)
        if self.noise0 is not None:
            self.noise0.copy_(noise0.reshape(self.noise0.shape))
        if self.x_start is not None:
            self.x_start.copy_(x_start_bfmt.reshape(self.x_start.shape))
        self.graph.replay()
        return self.out              # static buffer: the caller copies / de-normalises it before the next replay


# Captured graphs, least recently used first.  A graph pins its engine, its session and a private memory pool of hundreds of MB,
# so only a few are kept; the keys seen ONCE (a graph is captured on the second call with a key) live in their own small LRU so
# that a stream of new shapes (variable-length batches) never evicts a live graph.
_GRAPH_CACHE: 'OrderedDict[tuple, object]' = OrderedDict()
_GRAPH_CACHE_MAX = 4          # default; hparams['b2s_graph_cache'] overrides (the segment driver asks for one graph per length bucket)
_SEEN_KEYS: 'OrderedDict[tuple, None]' = OrderedDict()
_SEEN_KEYS_MAX = 64
# hparams that change WHICH kernels a sampling call launches: part of the graph key (toggling one after capture must not replay
# the old launch structure)
_STRUCTURE_HPARAMS = ('b2s_stack', 'b2s_stack3', 'b2s_stack3_head', 'b2s_stack_t', 'b2s_stack_t_tile', 'b2s_fuse_io', 'b2s_fuse_update',
                      'b2s_overlap_noise', 'b2s_fuse_cast', 'b2s_defer_skip', 'b2s_lynx_fold_cond', 'b2s_pad_channels', 'b2s_chain_groups', 'b2s_narrow_slabs')


def _program_key(prog: Program):
    return hash((tuple((op.kind, op.dst, op.src, op.t_index, tuple(op.terms), op.draw) for op in prog.ops),
                 tuple(prog.t_values)))


_THRASH = {'evicted_unused': 0, 'last_key': None}


def clear_graph_cache():
    _GRAPH_CACHE.clear()
    _THRASH.update(evicted_unused=0, last_key=None)
    _SEEN_KEYS.clear()


def sample(backbone, prog: Program, cond_bht: torch.Tensor, b: int, num_feats: int, out_dims: int,
           x_start: Optional[torch.Tensor], device,
           noise_source: Optional[Callable[[tuple], torch.Tensor]] = None, lengths=None, noise0=None,
           denorm=None) -> torch.Tensor:
    """Runs ``prog`` with ``backbone`` as the denoiser.  Returns [B, T, M] or [B, F, T, M]
    (the transpose of ddpm.py:350 / reflow.py:137 is free: the state is time-major already).

    ``noise_source(shape)`` must return a standard-normal tensor of ``shape`` = (B, F, M, T); the
    default is ``torch.randn`` on ``device``, called in the reference's order (initial draw first,
    then one draw per ancestral step), so a seeded run consumes the same Philox stream as the
    reference on the same device.

    With the default noise source and ``hparams['b2s_cuda_graph']`` (default True) the second call
    with the same (backbone weights, program, B, T) captures the whole loop as a CUDA graph; from
    then on a sampling call is one graph launch.
    """
    if device is None:
        device = cond_bht.device
    device = torch.device(device)
    if device.type != 'cuda':
        raise C.B2SError(f'sampling needs a CUDA device (got {device}); this path has no CPU fallback')
    B, H, T = cond_bht.shape
    if B != b:
        raise C.B2SError(f'batch size mismatch: cond has {B} utterances, b={b}')
    F_, M = num_feats, out_dims
    shape = (B, F_, M, T)
    lens = None
    if lengths is not None:
        lens = torch.as_tensor(lengths).reshape(-1).to(device=device, dtype=torch.int32)
        if lens.numel() != B:
            raise C.B2SError(f'lengths must have one entry per utterance (got {lens.numel()} for B={B})')
    if noise0 is not None and tuple(noise0.shape) != shape:
        raise C.B2SError(f'initial_noise must have shape {shape} (got {tuple(noise0.shape)})')
    with torch.cuda.device(device):         # every launch below (and the current stream) belongs to the tensors' device
        return _sample_on_device(backbone, prog, cond_bht, B, H, T, F_, M, shape, x_start, device, noise_source, lens, noise0,
                                 denorm)


def _sample_on_device(backbone, prog, cond_bht, B, H, T, F_, M, shape, x_start, device, noise_source, lens=None, noise0_in=None,
                      denorm=None):
    eng = backbone._engine()
    eng.pack()
    if prog.needs_x_start:
        assert x_start is not None, 'Missing shallow diffusion source.'
    cond_bth = _time_major_cond(cond_bht.float())

    start_tm = isinstance(x_start, TimeMajorState)
    if start_tm:
        x_start = x_start.t

    def finish(x, static=False):
        if denorm is not None:                              # denorm_spec + layout change in one kernel (also un-aliases a static buffer)
            out = torch.empty((B, T, M) if F_ == 1 else (B, F_, T, M), device=device)
            if B * T:
                C.spec_denorm(x, denorm[0], denorm[1], out, B, F_, T, M)
            return out
        if static:
            x = x.clone()
        x = x.reshape(B, T, F_, M)
        if F_ == 1:
            return x[:, :, 0, :]                            # [B, T, M]
        return x.permute(0, 2, 1, 3).contiguous()           # [B, F, T, M]

    if B * T == 0:                                          # empty batch / zero frames: nothing to launch
        return finish(torch.zeros((0, F_ * M), device=device))

    graph_key = None
    if noise_source is None and hparams.get('b2s_cuda_graph', True):
        structure = tuple(repr(hparams.get(k)) for k in _STRUCTURE_HPARAMS)
        graph_key = (id(backbone), eng._packed_version, eng.precision, _program_key(prog), B, T, F_, M, str(device), structure,
                     lens is not None, noise0_in is not None, start_tm)
        entry = _GRAPH_CACHE.get(graph_key)
        cache_max = max(1, int(hparams.get('b2s_graph_cache', _GRAPH_CACHE_MAX)))
        # Thrash guard: a caller that cycles through more shapes than the cache holds (one segment per call over a ragged project)
        # would capture every shape on its second sight and lose the graph before its third - captures cost more than the launches
        # they save.  Once a cache's worth of graphs in a row has been evicted without a single replay, nothing new is captured (the
        # cached graphs keep replaying, everything else is launched from the host) until a key is asked for twice IN A ROW - a steady
        # workload - which is captured and lifts the guard.
        steady = _THRASH['last_key'] == graph_key
        thrashing = _THRASH['evicted_unused'] >= cache_max and not steady
        _THRASH['last_key'] = graph_key
        if entry is None and graph_key in _SEEN_KEYS and not thrashing:       # second call with this key: capture
            del _SEEN_KEYS[graph_key]
            if steady:
                _THRASH['evicted_unused'] = 0
            while len(_GRAPH_CACHE) >= cache_max:
                _, old = _GRAPH_CACHE.popitem(last=False)   # least recently used graph
                _THRASH['evicted_unused'] = _THRASH['evicted_unused'] + 1 if getattr(old, 'replays', 0) == 0 else 0
            entry = _GraphedLoop(eng, CompiledProgram(prog, device), B, T, H, F_, M, device, ragged=lens is not None,
                                 ext_noise=noise0_in is not None, start_tm=start_tm)
            entry.replays = -1                              # the run below is the capture's own
            _GRAPH_CACHE[graph_key] = entry
        if entry is not None:
            entry.replays = getattr(entry, 'replays', 0) + 1
            _GRAPH_CACHE.move_to_end(graph_key)
            xs = None if x_start is None else x_start.to(device=device, dtype=torch.float32)
            n0 = None if noise0_in is None else noise0_in.to(device=device, dtype=torch.float32)
            return finish(entry.run(cond_bth, xs, lens, n0), static=True)
    if noise_source is None:
        noise_source = lambda s: torch.randn(s, device=device)
    cp = CompiledProgram(prog, device)
    sess = eng.begin(cond_bth, cp.t_values, per_row_t=False, lens=lens, table_owner=cp)
    bufs = allocate_buffers(prog, B * T, F_ * M, device)

    def load(src_bfmt, dst):
        C.transpose(src_bfmt.to(device=device, dtype=torch.float32).reshape(B, F_ * M, T).contiguous(), dst,
                    B, F_ * M, T)

    noise0 = noise0_in if noise0_in is not None else noise_source(shape)      # always drawn first (ddpm.py:227, reflow.py:105)
    if NOISE0 in bufs:
        load(noise0, bufs[NOISE0])
    if prog.needs_x_start:
        if start_tm:
            bufs[XSTART].copy_(x_start)
        else:
            load(x_start, bufs[XSTART])

    x = run_program(cp, sess, bufs, lambda j, dst: load(noise_source(shape), dst))   # [B*T, F*M]
    if graph_key is not None:
        _SEEN_KEYS[graph_key] = None
        while len(_SEEN_KEYS) > _SEEN_KEYS_MAX:
            _SEEN_KEYS.popitem(last=False)
    return finish(x)
