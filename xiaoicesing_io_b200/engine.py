"""Device-side orchestration of the hot path: weight packing, the per-utterance hoisted tables, one
denoiser evaluation as a sequence of libb2s launches, and the sampler-program executor.

Data layout in HBM (all time-major, row r = b*T + t):
    state / eps / history buffers   [B*T, M*F]   fp32
    x (residual stream), y = x + step embedding, z (gated), skip accumulator   [B*T, C]
    cond table   [B*T, L*2C] (WaveNet, gate/filter interleaved, both conv biases folded in)
                 [B*T, L*C]  (LYNXNet)                      - computed ONCE per utterance batch
    step table   [K, L*C]    one row per denoiser evaluation of the whole sampling loop
                             - computed ONCE per sampling call (all K model times are known up front)

The reference recomputes the conditioner projection (wavenet.py:35, lynxnet.py:77-82) and the step
embedding MLP (wavenet.py:89-90, :34) inside every denoiser call; its own ONNX exporter hoists the
former (utils/onnx_helper.py:231-314), which is the precedent for doing it here.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional

import torch

from . import _cabi as C
from .hparams import hparams
from .schedules import NOISE0, XSTART, Program, Z

PRECISIONS = ('fp32', 'bf16', 'fp16')


def _dev_f32(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


# =====================================================================================================
# WaveNet
# =====================================================================================================
class WaveNetEngine:
    """Packed weights + launch sequence for ``WaveNet.forward`` (reference wavenet.py:75-107)."""

    def __init__(self, net, precision: str = 'fp32'):
        if precision not in PRECISIONS:
            raise ValueError(f'unknown precision {precision!r}; available: {PRECISIONS}')
        self.net = net
        self.precision = precision
        self.C0 = net.num_channels                            # the model's channel count (step-embedding MLP, algorithmic FLOPs)
        self.L = net.num_layers
        self.MF = net.in_dims * net.n_feats
        self.H = net.hidden_size
        self.dilations = [layer.dilation for layer in net.residual_layers]
        # Narrow WaveNets (the 192-channel variance predictor) run on the 256-channel whole-stack kernel with
        # ZERO-PADDED channels: a padded channel has zero stem / conv / conditioner / residual / skip rows and zero step embedding,
        # so it stays exactly 0 through every layer (gate: sigmoid(0) * tanh(0) = 0) and contributes exactly 0 to every real
        # channel - the same result, (256 / C)^2 more tensor-core work, but one persistent kernel instead of 2 launches per layer.
        self.C = self.C0
        if (precision != 'fp32' and hparams.get('b2s_pad_channels', True) and self.C0 < C.FUSED_LAYER_CHANNELS
                and self.C0 % 8 == 0 and (self.C0 > 128 or self.C0 % 64) and max(self.dilations) <= 16 and self.L <= 32
                and self.MF <= 256):
            self.C = C.FUSED_LAYER_CHANNELS
        # channels the whole-stack kernels are told about: 192 when the model fits three of the four 64-channel K slabs (the zero
        # fourth slab of every GEMM is skipped, bit-identical results), else the layout's 256
        self.C_used = 192 if (self.C != self.C0 and self.C0 <= 192 and hparams.get('b2s_narrow_slabs', True)) else self.C
        self._packed_version = None
        self.device = None

    # -- weights ------------------------------------------------------------------------------------
    def _version(self):
        # version counters catch in-place updates (optimizer steps, load_state_dict, p.data.copy_ does NOT bump them but the
        # storage pointer / dtype part catches p.data = t, .half(), .to()); see _B2SBackbone.invalidate() for the rest
        ps = list(self.net.parameters())
        return tuple((p._version, p.data_ptr(), p.dtype) for p in ps) + (str(ps[0].device),)

    def pack(self, force=False):
        v = self._version()
        if not force and v == self._packed_version:
            return
        net = self.net
        dev = next(net.parameters()).device
        if dev.type != 'cuda':
            raise C.B2SError('the denoiser lives on the CPU; this path has no CPU fallback - move the module to a CUDA device')
        Cc, C0, L = self.C, self.C0, self.L
        f = lambda t: _dev_f32(t, dev)

        def P(t, modes):
            # channel padding C0 -> C (no-op when equal): per dim 'c' = a channel axis, 'g' = a [gate | filter] or
            # [residual | skip] axis of 2*C0 rows (each half padded on its own), '-' = left alone
            t = t.detach()
            if Cc == C0:
                return t
            for d, m in enumerate(modes):
                if m == 'c':
                    shp = list(t.shape); shp[d] = Cc - C0
                    t = torch.cat([t, t.new_zeros(shp)], d)
                elif m == 'g':
                    a, b = t.split(C0, d)
                    shp = list(a.shape); shp[d] = Cc - C0
                    z = t.new_zeros(shp)
                    t = torch.cat([a, z, b, z], d)
            return t

        self.device = dev
        self.w_in = f(P(net.input_projection.weight[:, :, 0], 'c-'))         # [C, MF]
        self.b_in = f(P(net.input_projection.bias, 'c'))
        self.w_mlp0, self.b_mlp0 = f(net.mlp[0].weight), f(net.mlp[0].bias)  # [4C0, C0]
        self.w_mlp2, self.b_mlp2 = f(net.mlp[2].weight), f(net.mlp[2].bias)  # [C0, 4C0]
        layers = net.residual_layers
        # all L diffusion projections as one [L*C, C0] matrix -> one GEMM for the whole step table
        self.w_dp = f(torch.cat([P(l.diffusion_projection.weight, 'c-') for l in layers], 0))
        self.b_dp = f(torch.cat([P(l.diffusion_projection.bias, 'c') for l in layers], 0))
        # gate/filter interleave: packed row 2j = gate j (reference row j), 2j+1 = filter j (row C+j)
        perm = torch.stack([torch.arange(Cc), torch.arange(Cc) + Cc], 1).reshape(-1).to(dev)
        wc, bc, wd = [], [], []
        for l in layers:
            wc.append(P(l.conditioner_projection.weight[:, :, 0], 'g-')[perm])          # [2C, H]
            bc.append(P(l.conditioner_projection.bias + l.dilated_conv.bias, 'g')[perm])   # both biases folded
            w = P(l.dilated_conv.weight, 'gc-')[perm]                         # [2C, C, 3]
            wd.append(w.permute(0, 2, 1).reshape(2 * Cc, 3 * Cc))             # column = tap*C + c
        self.w_cond = f(torch.cat(wc, 0))                                     # [L*2C, H]
        self.b_cond = f(torch.cat(bc, 0))
        self.w_dil = f(torch.stack(wd, 0))                                    # [L, 2C, 3C]
        self.w_out = f(torch.stack([P(l.output_projection.weight[:, :, 0], 'gc') for l in layers], 0))   # [L, 2C, C]
        self.b_out = f(torch.stack([P(l.output_projection.bias, 'g') for l in layers], 0))               # [L, 2C]
        self.w_sp, self.b_sp = f(P(net.skip_projection.weight[:, :, 0], 'cc')), f(P(net.skip_projection.bias, 'c'))
        self.w_fin, self.b_fin = f(P(net.output_projection.weight[:, :, 0], '-c')), f(net.output_projection.bias)
        if self.precision != 'fp32':
            if Cc % 64:
                raise C.B2SError(f'the {self.precision} tensor-core path needs num_channels % 64 == 0 (got {Cc}); '
                                 f"use b2s_precision='fp32'")
            hd = C.HALF_DTYPES[self.precision]
            self.bf16 = self.precision == 'bf16'
            h = lambda t: t.to(hd).contiguous()
            self.w_in_h = h(_pad_cols(self.w_in, 8))
            self.w_cond_h, self.w_dil_h, self.w_out_h = h(self.w_cond), h(self.w_dil), h(self.w_out)
            self.w_sp_h, self.w_fin_h = h(self.w_sp), h(self.w_fin)
            # deferred skip sum (two-kernel path): residual rows per layer, all skip rows as ONE [C, L*C] matrix (K = l*C + k)
            self.w_res_h = h(self.w_out[:, :Cc, :])
            self.b_res = self.b_out[:, :Cc].contiguous()
            self.w_skipcat_h = h(self.w_out[:, Cc:, :].permute(1, 0, 2).reshape(Cc, L * Cc))
            self.b_skip_sum = self.b_out[:, Cc:].sum(0).contiguous()
            # third stack kernel (b2s_tc_wavenet_stack3): the residual stream lives in TMEM as X_l = 2^(l/2) x_l - (bias terms),
            # so layer l's residual rows carry 2^(l/2) and the biases become one exclusive prefix sum per layer
            sc = 2.0 ** (0.5 * torch.arange(L, device=dev, dtype=torch.float64))
            w_res3 = self.w_out[:, :Cc, :].double() * sc[:, None, None]
            b_res3 = self.b_out[:, :Cc].double() * sc[:, None]
            self.w_res3_h = h(w_res3.float())
            self.bsum3 = (torch.cumsum(b_res3, 0) - b_res3).float().contiguous()              # [L, C]: sum over k < l
            self.res3_ok = bool(torch.isfinite(self.w_res3_h.float()).all())                  # fp16 range (2^(l/2) <= 2^15.5 at L = 32)
            self.w_skip3_h = h(self.w_out[:, Cc:, :])                                          # [L, C, C]: skip rows per layer
            self.w_dil_t_h = self.w_cond_t_h = self.b_cond_t = None
            if C0 == C.FUSED_LAYER_CHANNELS and C.HAS_EXPERIMENTS:
                # transposed stack kernel (b2s_tc_wavenet_stack_t): row 256h + 128g + c = (g ? filter : gate) of channel 128h + c,
                # i.e. gate and filter of a channel land on the same TMEM lane of two neighbouring accumulator blocks
                n = torch.arange(2 * Cc, device=dev)
                perm_t = 128 * (n // 256) + (n % 128) + Cc * ((n % 256) // 128)
                wc_t, bc_t, wd_t = [], [], []
                for l in layers:
                    wc_t.append(l.conditioner_projection.weight[:, :, 0][perm_t])
                    bc_t.append((l.conditioner_projection.bias + l.dilated_conv.bias)[perm_t])
                    wd_t.append(l.dilated_conv.weight[perm_t].permute(0, 2, 1).reshape(2 * Cc, 3 * Cc))
                self.w_cond_t_h = h(f(torch.cat(wc_t, 0)))
                self.b_cond_t = f(torch.cat(bc_t, 0))
                self.w_dil_t_h = h(f(torch.stack(wd_t, 0)))
        self._packed_version = v

    # -- per-call tables ---------------------------------------------------------------------------
    def step_table(self, t_values: torch.Tensor) -> torch.Tensor:
        """t_values [K] fp32 (device) -> [K, L*C]: diffusion_projection_l(mlp(sinusoid(t))) for all l."""
        return _cached_step_table(self, t_values, self._step_table)

    def _step_table(self, t_values: torch.Tensor) -> torch.Tensor:
        K = t_values.numel()
        Cc, LC = self.C0, self.L * self.C                    # the MLP has the model's width; the table the (padded) kernels' width
        dev = self.device
        sin = torch.empty((K, Cc), device=dev)
        C.sinusoid(t_values, sin, K, Cc)
        h = torch.empty((K, 4 * Cc), device=dev)
        C.linear(sin, Cc, self.w_mlp0, Cc, self.b_mlp0, h, 4 * Cc, K, 4 * Cc, Cc, act=C.ACT_MISH)
        e = torch.empty((K, Cc), device=dev)
        C.linear(h, 4 * Cc, self.w_mlp2, 4 * Cc, self.b_mlp2, e, Cc, K, Cc, 4 * Cc)
        tab = torch.empty((K, LC), device=dev)
        C.linear(e, Cc, self.w_dp, Cc, self.b_dp, tab, LC, K, LC, Cc)
        return tab

    def cond_table(self, cond_bth: torch.Tensor) -> torch.Tensor:
        """cond [B, T, H] (time-major, contiguous) -> [B*T, L*2C], ONE batched GEMM for all layers."""
        B, T, H = cond_bth.shape
        N = self.L * 2 * self.C
        tab = torch.empty((B * T, N), device=self.device)
        C.linear(cond_bth, H, self.w_cond, H, self.b_cond, tab, N, B * T, N, H)
        return tab

    def begin(self, cond_bth: torch.Tensor, t_values: torch.Tensor, per_row_t: bool = False, lens=None, table_owner=None) -> 'WaveNetSession':
        """``lens``: optional int32 device tensor [B] - a RAGGED batch padded to T; frames at or beyond lens[b] are treated as the
        conv's zero padding (only the whole-stack tensor-core path implements this).  ``table_owner``: the compiled sampler program
        ``t_values`` belongs to - the step table (a function of the weights and of the program's model times only) is then computed
        once and kept on it instead of once per sampling call."""
        self.pack()
        C.require_cuda(cond_bth, 'cond')
        C.require_cuda(t_values, 't_values')
        self._table_owner = table_owner
        try:
            if self.precision != 'fp32':
                return WaveNetSessionTC(self, cond_bth, t_values, per_row_t, lens)
            if lens is not None:
                raise C.B2SError("ragged batches (lengths=...) need the 16-bit whole-stack path (b2s_precision 'fp16' / 'bf16', 256 channels, "
                                 "dilations <= 16); batch equal-length segments instead")
            return WaveNetSession(self, cond_bth, t_values, per_row_t)
        finally:
            self._table_owner = None


def _cached_step_table(eng, t_values, compute):
    """The step table of (engine weights, sampler program): three tiny fp32 GEMMs that cost ~0.3 ms per sampling call - 5 % of a
    one-utterance DDIM-20 call - although nothing in them changes between calls.  Kept on the compiled program (``table_owner``),
    keyed by the engine and its packed-weight version; never created while a CUDA graph is being captured (a captured fill would
    only be valid after that graph's replay), so the first, eager call of a shape fills it and the captured graphs read it."""
    owner = getattr(eng, '_table_owner', None)
    if owner is None:
        return compute(t_values)
    cache = owner.__dict__.setdefault('_b2s_step_tables', {})
    key = (id(eng), eng._packed_version)
    tab = cache.get(key)
    if tab is None:
        tab = compute(t_values)
        if not torch.cuda.is_current_stream_capturing():
            if len(cache) >= 4:
                cache.clear()
            cache[key] = tab
    return tab


def _pad_cols(w: torch.Tensor, mult: int) -> torch.Tensor:
    """Zero-pads the K (last) dimension to a multiple of ``mult`` (16B-aligned rows for TMA)."""
    k = w.shape[-1]
    if k % mult == 0:
        return w
    return torch.nn.functional.pad(w, (0, mult - k % mult))


class WaveNetSession:
    """Everything that is constant across the K denoiser evaluations of one sampling call."""

    def __init__(self, eng: WaveNetEngine, cond_bth, t_values, per_row_t):
        self.eng = eng
        B, T, H = cond_bth.shape
        if H != eng.H:
            raise C.B2SError(f'condition has {H} channels, backbone expects hidden_size={eng.H}')
        if per_row_t and t_values.numel() != B:
            raise C.B2SError('per-row step embedding needs one time value per utterance')
        self.B, self.T = B, T
        self.rows = B * T
        self.per_row_t = per_row_t
        self.cond = eng.cond_table(cond_bth)
        self.dtab = eng.step_table(t_values)                 # [K, L*C]
        dev = eng.device
        rows, Cc = self.rows, eng.C
        self.x = torch.empty((rows, Cc), device=dev)
        self.y = torch.empty((rows, Cc), device=dev)
        self.z = torch.empty((rows, Cc), device=dev)
        self.skip = torch.empty((rows, Cc), device=dev)
        self.h = torch.empty((rows, Cc), device=dev)

    def _dvec(self, k, l):
        LC = self.eng.L * self.eng.C
        if self.per_row_t:
            return self.dtab[0, l * self.eng.C:], LC
        return self.dtab[k, l * self.eng.C:], 0

    def eval(self, x_in: torch.Tensor, k: int, out: torch.Tensor):
        """One denoiser call: x_in [B*T, MF] -> out [B*T, MF] using step-table row k."""
        e = self.eng
        B, T, rows, Cc, L, MF = self.B, self.T, self.rows, e.C, e.L, e.MF
        d0, ds = self._dvec(k, 0)
        # input_projection + ReLU (wavenet.py:86-88); y = x + d_0 (wavenet.py:36)
        C.linear(x_in, MF, e.w_in, MF, e.b_in, self.x, Cc, rows, Cc, MF, act=C.ACT_RELU,
                 y=self.y, dvec=d0, d_stride=ds, T=T)
        ldc = L * 2 * Cc
        for l in range(L):
            C.wavenet_gate(self.y, e.w_dil[l], self.cond[:, l * 2 * Cc:], ldc, self.z, B, T, Cc, e.dilations[l])
            if l + 1 < L:
                dn, ds = self._dvec(k, l + 1)
                C.wavenet_out(self.z, e.w_out[l], e.b_out[l], self.x, self.y, self.skip, dn, ds, l == 0, B, T, Cc)
            else:
                C.wavenet_out(self.z, e.w_out[l], e.b_out[l], self.x, None, self.skip, None, 0, l == 0, B, T, Cc)
        # sum(skip)/sqrt(L) -> skip_projection -> ReLU -> output_projection (wavenet.py:96-99)
        C.linear(self.skip, Cc, e.w_sp, Cc, e.b_sp, self.h, Cc, rows, Cc, Cc, alpha=1.0 / math.sqrt(L), act=C.ACT_RELU)
        C.linear(self.h, Cc, e.w_fin, Cc, e.b_fin, out, MF, rows, MF, Cc)

    @property
    def launches_per_eval(self) -> int:
        return 1 + 2 * self.eng.L + 2


class WaveNetSessionTC:
    """16-bit tensor-core session (tcgen05 kernels): same launch structure as ``WaveNetSession`` with
    bf16 / fp16 MMA operands (y, z, hoisted cond table, weights) and fp32 residual stream, skip sum,
    biases, step embeddings and sampler state."""

    def __init__(self, eng: WaveNetEngine, cond_bth, t_values, per_row_t, lens=None):
        self.eng = eng
        B, T, H = cond_bth.shape
        self.lens = lens
        if lens is not None and (lens.dtype != torch.int32 or not lens.is_cuda or lens.numel() != B):
            raise C.B2SError('lens must be an int32 CUDA tensor with one entry per utterance')
        if H != eng.H:
            raise C.B2SError(f'condition has {H} channels, backbone expects hidden_size={eng.H}')
        if per_row_t and t_values.numel() != B:
            raise C.B2SError('per-row step embedding needs one time value per utterance')
        if H % 8 or eng.MF % 8:
            raise C.B2SError(f'the tensor-core path needs hidden_size and in_dims*n_feats to be multiples of 8 '
                             f'(got {H}, {eng.MF})')
        self.B, self.T, self.rows = B, T, B * T
        self.per_row_t = per_row_t
        dev, hd, bf = eng.device, C.HALF_DTYPES[eng.precision], eng.bf16
        rows, Cc, L = self.rows, eng.C, eng.L
        self.dtab = eng.step_table(t_values)                 # [K, L*C] fp32 (tiny; CUDA-core GEMMs)
        # hoisted conditioner projection of ALL layers: one tensor-core GEMM, 16-bit table [rows, L*2C]
        cond_h = torch.empty((rows, H), device=dev, dtype=hd)
        C.cast_h(cond_bth, cond_h, bf)
        self._cond_h, self._hd = cond_h, hd
        self.cond = None          # built below once the path (stack / per-layer) is known
        self.xin_h = torch.empty((rows, eng.MF), device=dev, dtype=hd)
        self.x = torch.empty((rows, Cc), device=dev)
        self.skip = torch.empty((rows, Cc), device=dev)
        self.y_h = torch.empty((rows, Cc), device=dev, dtype=hd)
        self.fused = Cc == C.FUSED_LAYER_CHANNELS           # one fused kernel per layer (z never leaves the SM)
        self.y2_h = torch.empty((rows, Cc), device=dev, dtype=hd) if self.fused else None
        self.z_h = None if self.fused else torch.empty((rows, Cc), device=dev, dtype=hd)
        # whole-stack persistent kernel: every 128-frame tile of a launch must be resident, so the batch is split by
        # utterance into groups of `stack_group` (0 = an utterance alone exceeds the SM count -> per-layer kernels)
        tpb = (-(-T // 128) + 1) & ~1
        self.stack_group = (C.lib.b2s_tc_wavenet_stack_max_tiles() // tpb) if (self.fused and hparams.get('b2s_stack', True)) else 0
        # third stack kernel (default): resident y tile (halo <= 16 rows), residual stream in TMEM, deferred skip sum; its launch
        # capacity comes from cudaOccupancyMaxActiveClusters for the cluster size it will use for this T
        self.stack3 = False
        if (self.stack_group and hparams.get('b2s_stack3', True) and not hparams.get('b2s_stack_t', False) and eng.res3_ok
                and max(eng.dilations) <= C.lib.b2s_tc_wavenet_stack3_halo() and eng.MF <= 256):
            g3 = C.lib.b2s_tc_wavenet_stack3_max_tiles(T, int(bf)) // tpb
            if g3 > 0:
                self.stack3, self.stack_group = True, g3
        # ... with the skip sum + head (wavenet.py:96-99) in a second kernel that is released (programmatic dependent launch) as soon
        # as every layer tile is resident: it runs on the SMs the layer tiles leave idle and follows them through the z-tile flags
        self.head3 = bool(self.stack3 and hparams.get('b2s_stack3_head', True) and eng.MF % 16 == 0
                          and C.lib.b2s_tc_wavenet_denoiser3_max_utterances(T, int(bf)) >= self.stack_group)
        if lens is not None and not self.stack3:
            raise C.B2SError('ragged batches (lengths=...) need the whole-stack kernel (256 channels, dilations <= 16, b2s_stack3 on); '
                             'batch equal-length segments instead')
        self.flags = torch.zeros((2 * B * tpb,), device=dev, dtype=torch.int32) if self.stack_group else None     # tile flags | z flags
        self.tpb = tpb
        if hparams.get('b2s_stack_t', False):
            C.require_experiments("hparams['b2s_stack_t'] (the transposed stack kernel)")
        self.tgroups = self._plan_transposed(cond_h) if (self.stack_group and hparams.get('b2s_stack_t', False)) else None
        if self.tgroups:
            self.flags = torch.zeros((max(B * tpb, sum(g[4] for g in self.tgroups)),), device=dev, dtype=torch.int32)
        self.flags_alt = torch.zeros_like(self.flags) if self.flags is not None else None
        if self.stack_group and not self.tgroups:
            # one table per utterance group, in the tile/chunk-major layout the stack kernel reads with coalesced loads
            self.cond_groups = []
            for b0 in range(0, B, self.stack_group):
                nb = min(B, b0 + self.stack_group) - b0
                tab = torch.empty((L, nb * tpb * 128, 2 * Cc), device=dev, dtype=hd)
                C.tc_cond_table_tiled(cond_h[b0 * T:], nb, T, eng.w_cond_h, eng.b_cond, L, 2 * Cc, H, tab, bf)
                self.cond_groups.append(tab)
        elif not self.tgroups:
            self.cond = torch.empty((L, rows, 2 * Cc), device=dev, dtype=hd)     # layer-major: one contiguous slab per layer
            C.tc_cond_table(cond_h, rows, eng.w_cond_h, eng.b_cond, L, 2 * Cc, H, self.cond, bf)
        del self._cond_h
        # two-kernel path: the L skip outputs are summed by ONE K = L*C GEMM after the last layer (no per-layer fp32 skip RMW)
        self.defer_skip = (not self.fused) and hparams.get('b2s_defer_skip', True) and rows * L * Cc * 2 <= 8e9
        # layer-major: contiguous rows per layer (zero-initialised for narrow models: their fourth K slab is never written, and the
        # deferred skip GEMM of the non-head3 path multiplies it with zero weights)
        alloc = torch.zeros if eng.C_used != Cc else torch.empty
        self.z_all = alloc((L, rows, Cc), device=dev, dtype=hd) if (self.defer_skip or self.stack3) else None
        self.skip_h = torch.empty((rows, Cc), device=dev, dtype=hd)
        self.h_h = torch.empty((rows, Cc), device=dev, dtype=hd)

    def _dvec(self, k, l):
        LC = self.eng.L * self.eng.C
        if self.per_row_t:
            return self.dtab[0, l * self.eng.C:], LC
        return self.dtab[k, l * self.eng.C:], 0

    T_TILES = (32, 48, 64, 80)

    def _plan_transposed(self, cond_h):
        """Utterance groups for the transposed stack kernel (``b2s_stack_t``, OFF by default; frame tiles of 32..80 instead of
        128, so a small batch occupies every SM): [(b0, b1, NT, retiled cond table, tiles, flag offset)] or None.
        Every tile of a launch must be resident.  ``True``: only when the whole batch fits ONE launch; ``'always'``: any batch.
        Measured: a tcgen05 SS-mode instruction costs the same ~128 cycles for N = 48 .. 80 as for
        N = 256 (the 128 x 16 A slice is re-read from shared memory every instruction), so the 4x larger instruction count of
        this formulation makes it SLOWER than the 128-row kernel despite 144 busy SMs (18.0 vs 20.8 M frame*NFE/s)."""
        e = self.eng
        B, T, Cc, L, H, bf = self.B, self.T, e.C, e.L, e.H, e.bf16
        if e.w_dil_t_h is None or max(e.dilations) > 16:
            return None
        cap = C.lib.b2s_tc_wavenet_stack_max_tiles()
        tiles = lambda nb, nt: C.lib.b2s_tc_wavenet_stack_t_tiles(nb, T, nt)
        per = max(nb for nb in range(0, B + 1) if tiles(nb, self.T_TILES[-1]) <= cap)     # utterances per launch at the largest tile
        if per == 0 or (per < B and hparams.get('b2s_stack_t', False) != 'always'):
            return None
        hd = C.HALF_DTYPES[e.precision]
        dev = cond_h.device
        groups, f0 = [], 0
        for b0 in range(0, B, per):
            b1 = min(B, b0 + per)
            nb = b1 - b0
            nt = next(t for t in self.T_TILES if t >= max(e.dilations) and tiles(nb, t) <= cap)
            nt = int(hparams.get('b2s_stack_t_tile', nt))
            n_tiles = tiles(nb, nt)
            tmp = torch.empty((L, nb * T, 2 * Cc), device=dev, dtype=hd)
            C.tc_cond_table(cond_h[b0 * T:], nb * T, e.w_cond_t_h, e.b_cond_t, L, 2 * Cc, H, tmp, bf)
            tab = torch.empty((L * n_tiles * nt * 2 * Cc,), device=dev, dtype=hd)
            C.tc_cond_retile(tmp, L, nb, T, 2 * Cc, nt, tab)
            del tmp
            groups.append((b0, b1, nt, tab, n_tiles, f0))
            f0 += n_tiles
        return groups

    def half_sink(self):
        """(16-bit input buffer, tile flags of the NEXT evaluation, bf16): a sampler update that produces the next denoiser
        input can write its 16-bit copy (and re-arm the flags) itself - ``eval(..., precast=True)`` then skips the cast."""
        return self.xin_h, self.flags, self.eng.bf16

    def _chain_bits(self, b0: int, b1: int) -> int:
        """Several utterance groups per evaluation: a group's layer kernel does not wait for the previous group's skip / head tail
        (b2s_tc_wavenet_denoiser3_chained); hparams['b2s_chain_groups'] = False restores the plain back-to-back launches."""
        if not hparams.get('b2s_chain_groups', True):
            return 0
        return (1 if b0 > 0 else 0) | (2 if b1 < self.B else 0)

    @property
    def can_fuse_update(self) -> bool:
        return (C.HAS_EXPERIMENTS and bool(self.stack_group) and not self.tgroups and not self.stack3
                and hparams.get('b2s_fuse_io', True) and self.eng.MF <= 256)

    def eval_update(self, x_in, k, srcs, coef, x_dst, precast=False):
        """One launch: denoiser evaluation + ``x_dst <- sum coef[i] * srcs[i]`` (``None`` in srcs = the evaluation's output),
        the 16-bit copy of x_dst for the next evaluation, and the next launch's tile flags re-armed (two flag buffers alternate)."""
        e = self.eng
        B, T, Cc, L, MF, bf = self.B, self.T, e.C, e.L, e.MF, e.bf16
        if not precast:
            C.cast_h(x_in, self.xin_h, bf, reset_flags=self.flags)
        LC = L * Cc
        dv = self.dtab.reshape(-1) if self.per_row_t else self.dtab[k]
        for gi, b0 in enumerate(range(0, B, self.stack_group)):
            b1 = min(B, b0 + self.stack_group)
            r0 = b0 * T
            tab = self.cond_groups[gi]
            C.tc_wavenet_denoiser_update(self.xin_h[r0:], MF, e.w_in_h, e.w_in_h.shape[1], e.b_in, self.y_h[r0:], self.y2_h[r0:],
                                         e.w_dil_h, tab, tab.shape[1] * 2 * Cc, e.w_out_h, e.b_out, self.x[r0:], self.skip[r0:],
                                         dv[b0 * LC:] if self.per_row_t else dv, LC if self.per_row_t else 0, e.dilations,
                                         e.w_sp_h, e.b_sp, e.w_fin_h, e.b_fin, b1 - b0, T, Cc, self.flags[b0 * self.tpb:],
                                         self.flags_alt[b0 * self.tpb:], [None if t is None else t[r0:] for t in srcs], coef,
                                         x_dst[r0:], self.xin_h[r0:], bf)
        self.flags, self.flags_alt = self.flags_alt, self.flags

    def eval(self, x_in: torch.Tensor, k: int, out: torch.Tensor, precast: bool = False):
        e = self.eng
        B, T, rows, Cc, L, MF, bf = self.B, self.T, self.rows, e.C, e.L, e.MF, e.bf16
        if not precast:
            C.cast_h(x_in, self.xin_h, bf, reset_flags=self.flags)   # also re-arms the tile flags of the persistent kernel
        if self.tgroups:
            LC = L * Cc
            d0, ds = self._dvec(k, 0)
            C.tc_linear(self.xin_h, MF, rows, T, e.w_in_h, e.w_in_h.shape[1], e.b_in, Cc, MF, bf, act=C.ACT_RELU,
                        out_f32=self.x, ldo=Cc, y_h=self.y_h, ldy=Cc, dvec=d0, d_stride=ds)
            dv = self.dtab.reshape(-1) if self.per_row_t else self.dtab[k]
            for (b0, b1, nt, tab, n_tiles, f0) in self.tgroups:
                r0 = b0 * T
                C.tc_wavenet_stack_t(self.y_h[r0:], self.y2_h[r0:], e.w_dil_t_h, tab, e.w_out_h, e.b_out, self.x[r0:],
                                     self.skip_h[r0:], dv[b0 * LC:] if self.per_row_t else dv, LC if self.per_row_t else 0,
                                     e.dilations, b1 - b0, T, Cc, nt, self.flags[f0:], bf)
            C.tc_linear(self.skip_h, Cc, rows, T, e.w_sp_h, Cc, e.b_sp, Cc, Cc, bf, alpha=1.0 / math.sqrt(L),
                        act=C.ACT_RELU, out_h=self.h_h, ldoh=Cc)
            C.tc_linear(self.h_h, Cc, rows, T, e.w_fin_h, Cc, e.b_fin, MF, Cc, bf, out_f32=out, ldo=MF)
            return
        if self.stack3:
            # stem + residual stack in ONE launch per utterance group (z_l of every layer -> z_all), then the deferred skip GEMM
            # (K = L*C) and the two head GEMMs - or, with head3, all of it in that one launch
            LC = L * Cc
            dv = self.dtab.reshape(-1) if self.per_row_t else self.dtab[k]
            nfl = B * self.tpb
            for gi, b0 in enumerate(range(0, B, self.stack_group)):
                b1 = min(B, b0 + self.stack_group)
                r0 = b0 * T
                tab = self.cond_groups[gi]
                if self.head3:
                    C.tc_wavenet_denoiser3(self.xin_h[r0:], MF, e.w_in_h, e.w_in_h.shape[1], e.b_in, e.w_dil_h, tab, tab.shape[1] * 2 * Cc,
                                           e.w_res3_h, e.bsum3, dv[b0 * LC:] if self.per_row_t else dv, LC if self.per_row_t else 0,
                                           e.dilations, self.y_h[r0:], self.y2_h[r0:], self.z_all[:, r0:], rows * Cc, e.w_skip3_h,
                                           e.b_skip_sum, e.w_sp_h, e.b_sp, e.w_fin_h, e.b_fin, out[r0:], b1 - b0, T, e.C_used,
                                           self.flags[b0 * self.tpb:], self.flags[nfl + b0 * self.tpb:], bf,
                                           lens=None if self.lens is None else self.lens[b0:],
                                           chain=self._chain_bits(b0, b1))
                    continue
                C.tc_wavenet_stack3(self.xin_h[r0:], MF, e.w_in_h, e.w_in_h.shape[1], e.b_in, e.w_dil_h, tab, tab.shape[1] * 2 * Cc,
                                    e.w_res3_h, e.bsum3, dv[b0 * LC:] if self.per_row_t else dv, LC if self.per_row_t else 0,
                                    e.dilations, self.y_h[r0:], self.y2_h[r0:], self.z_all[:, r0:], rows * Cc, b1 - b0, T, e.C_used,
                                    self.flags[b0 * self.tpb:], bf, lens=None if self.lens is None else self.lens[b0:])
            if self.head3:
                return
            C.tc_skip_sum(self.z_all, e.w_skipcat_h, e.b_skip_sum, self.skip_h, rows, Cc, L, bf)
            C.tc_linear(self.skip_h, Cc, rows, T, e.w_sp_h, Cc, e.b_sp, Cc, Cc, bf, alpha=1.0 / math.sqrt(L),
                        act=C.ACT_RELU, out_h=self.h_h, ldoh=Cc)
            C.tc_linear(self.h_h, Cc, rows, T, e.w_fin_h, Cc, e.b_fin, MF, Cc, bf, out_f32=out, ldo=MF)
            return
        if self.stack_group and hparams.get('b2s_fuse_io', True) and MF <= 256:
            # ONE launch per utterance group: stem + residual stack + head inside the persistent kernel
            LC = L * Cc
            dv = self.dtab.reshape(-1) if self.per_row_t else self.dtab[k]
            for gi, b0 in enumerate(range(0, B, self.stack_group)):
                b1 = min(B, b0 + self.stack_group)
                r0 = b0 * T
                tab = self.cond_groups[gi]
                C.tc_wavenet_denoiser(self.xin_h[r0:], MF, e.w_in_h, e.w_in_h.shape[1], e.b_in, self.y_h[r0:], self.y2_h[r0:],
                                      e.w_dil_h, tab, tab.shape[1] * 2 * Cc, e.w_out_h, e.b_out, self.x[r0:], self.skip[r0:],
                                      dv[b0 * LC:] if self.per_row_t else dv, LC if self.per_row_t else 0, e.dilations,
                                      e.w_sp_h, e.b_sp, e.w_fin_h, e.b_fin, out[r0:], b1 - b0, T, Cc, self.flags[b0 * self.tpb:], bf)
            return
        d0, ds = self._dvec(k, 0)
        C.tc_linear(self.xin_h, MF, rows, T, e.w_in_h, e.w_in_h.shape[1], e.b_in, Cc, MF, bf, act=C.ACT_RELU,
                    out_f32=self.x, ldo=Cc, y_h=self.y_h, ldy=Cc, dvec=d0, d_stride=ds)
        ldc = 2 * Cc
        if self.stack_group:
            LC = L * Cc
            dv = self.dtab.reshape(-1) if self.per_row_t else self.dtab[k]     # per-row: utterance b's row starts at b * L*C
            for gi, b0 in enumerate(range(0, B, self.stack_group)):
                b1 = min(B, b0 + self.stack_group)
                r0 = b0 * T
                tab = self.cond_groups[gi]
                C.tc_wavenet_stack(self.y_h[r0:], self.y2_h[r0:], e.w_dil_h, tab, ldc, tab.shape[1] * ldc, e.w_out_h,
                                   e.b_out, self.x[r0:], self.skip[r0:], self.skip_h[r0:],
                                   dv[b0 * LC:] if self.per_row_t else dv, LC if self.per_row_t else 0, e.dilations,
                                   b1 - b0, T, Cc, self.flags[b0 * self.tpb:], bf)
        ya, yb = self.y_h, self.y2_h
        for l in range(0 if not self.stack_group else L, L):
            last = l + 1 == L
            dn, ds = (None, 0) if last else self._dvec(k, l + 1)
            if self.fused:
                C.tc_wavenet_layer(ya, e.w_dil_h[l], self.cond[l], ldc, e.w_out_h[l], e.b_out[l], self.x,
                                   None if last else yb, self.skip, self.skip_h if last else None, dn, ds, l == 0,
                                   B, T, Cc, e.dilations[l], bf)
                ya, yb = yb, ya                            # ping-pong: neighbours still read the halo of ya
                continue
            if self.defer_skip:
                zl = self.z_all[l]
                C.tc_wavenet_gate(self.y_h, e.w_dil_h[l], self.cond[l], ldc, zl, B, T, Cc, e.dilations[l], bf)
                C.tc_wavenet_res(zl, Cc, e.w_res_h[l], e.b_res[l], self.x, None if last else self.y_h, dn, ds, B, T, Cc, bf)
                continue
            C.tc_wavenet_gate(self.y_h, e.w_dil_h[l], self.cond[l], ldc, self.z_h, B, T, Cc, e.dilations[l], bf)
            C.tc_wavenet_out(self.z_h, e.w_out_h[l], e.b_out[l], self.x, None if last else self.y_h, self.skip,
                             self.skip_h if last else None, dn, ds, l == 0, B, T, Cc, bf)
        if self.defer_skip:
            C.tc_skip_sum(self.z_all, e.w_skipcat_h, e.b_skip_sum, self.skip_h, rows, Cc, L, bf)
        C.tc_linear(self.skip_h, Cc, rows, T, e.w_sp_h, Cc, e.b_sp, Cc, Cc, bf, alpha=1.0 / math.sqrt(L),
                    act=C.ACT_RELU, out_h=self.h_h, ldoh=Cc)
        C.tc_linear(self.h_h, Cc, rows, T, e.w_fin_h, Cc, e.b_fin, MF, Cc, bf, out_f32=out, ldo=MF)

    @property
    def launches_per_eval(self) -> int:
        if self.tgroups:
            return 1 + 1 + len(self.tgroups) + 2                   # cast, stem, transposed stack launches, 2 head GEMMs
        if self.stack3:
            # cast (+ flag reset), stack launches (+ skip GEMM, 2 head GEMMs unless they run inside the launch)
            return 1 + -(-self.B // self.stack_group) + (0 if self.head3 else 3)
        if self.stack_group and hparams.get('b2s_fuse_io', True) and self.eng.MF <= 256:
            return 1 + -(-self.B // self.stack_group)              # cast (+ flag reset), one denoiser launch per utterance group
        if self.stack_group:
            return 1 + 1 + -(-self.B // self.stack_group) + 2      # cast (+ flag reset), stem, stack launches, 2 head GEMMs
        return 2 + (1 if self.fused else 2) * self.eng.L + 2 + (1 if self.defer_skip else 0)

    def dominant_kernel(self, w=None):
        """(name, algorithmic FLOPs per launch, callable launching it once per layer) for bench.py's roofline."""
        e = self.eng
        B, T, Cc, L = self.B, self.T, e.C, e.L
        if self.tgroups:
            b0, b1, nt, tab, n_tiles, f0 = self.tgroups[0]
            flops = 2.0 * (b1 - b0) * T * 8 * Cc * Cc * L
            dv = self.dtab[0]

            def launch_all():
                self.flags.zero_()
                for (b0, b1, nt, tab, n_tiles, f0) in self.tgroups:
                    r0 = b0 * T
                    C.tc_wavenet_stack_t(self.y_h[r0:], self.y2_h[r0:], e.w_dil_t_h, tab, e.w_out_h, e.b_out, self.x[r0:],
                                         self.skip_h[r0:], dv, 0, e.dilations, b1 - b0, T, Cc, nt, self.flags[f0:], e.bf16)
            return (f'wavenet_stack_t_kernel<{nt}, {e.precision}> (b2s_tc_wavenet_stack_t, {L} layers per launch, {n_tiles} tiles of {nt} frames)',
                    flops, launch_all, len(self.tgroups))
        if self.stack3:
            # algorithmic FLOPs of this launch: stem + conv (6C^2) + residual half of the output projection (C^2; the last layer's
            # is not needed) per frame; the skip half (C^2 per layer) and the head run in the same launch with head3, else in the
            # deferred b2s_tc_skip_sum / head GEMM launches
            nb = min(self.stack_group, B)
            Ca = e.C0                                          # ALGORITHMIC: the model's width, not the zero-padded one
            per_frame = e.MF * Ca + L * 6 * Ca * Ca + (L - 1) * Ca * Ca
            if self.head3:
                per_frame += L * Ca * Ca + Ca * Ca + Ca * e.MF
            flops = 2.0 * nb * T * per_frame
            dv = self.dtab[0]
            nfl = B * self.tpb
            out = torch.empty((self.rows, e.MF), device=self.y_h.device)

            def launch_all():
                self.flags.zero_()
                for gi, b0 in enumerate(range(0, B, self.stack_group)):
                    b1 = min(B, b0 + self.stack_group)
                    r0 = b0 * T
                    tab = self.cond_groups[gi]
                    if self.head3:
                        C.tc_wavenet_denoiser3(self.xin_h[r0:], e.MF, e.w_in_h, e.w_in_h.shape[1], e.b_in, e.w_dil_h, tab,
                                               tab.shape[1] * 2 * Cc, e.w_res3_h, e.bsum3, dv, 0, e.dilations, self.y_h[r0:], self.y2_h[r0:],
                                               self.z_all[:, r0:], self.rows * Cc, e.w_skip3_h, e.b_skip_sum, e.w_sp_h, e.b_sp, e.w_fin_h,
                                               e.b_fin, out[r0:], b1 - b0, T, e.C_used, self.flags[b0 * self.tpb:],
                                               self.flags[nfl + b0 * self.tpb:], e.bf16, chain=self._chain_bits(b0, b1))
                    else:
                        C.tc_wavenet_stack3(self.xin_h[r0:], e.MF, e.w_in_h, e.w_in_h.shape[1], e.b_in, e.w_dil_h, tab, tab.shape[1] * 2 * Cc,
                                            e.w_res3_h, e.bsum3, dv, 0, e.dilations, self.y_h[r0:], self.y2_h[r0:], self.z_all[:, r0:],
                                            self.rows * Cc, b1 - b0, T, e.C_used, self.flags[b0 * self.tpb:], e.bf16)
            what = (f'b2s_tc_wavenet_denoiser3: stem + {L} layers + skip sum + head per launch' if self.head3
                    else f'b2s_tc_wavenet_stack3: stem + {L} layers per launch, skip sum deferred')
            return (f'wavenet_stack3_kernel<{e.precision}> ({what})', flops, launch_all, -(-B // self.stack_group))
        if self.stack_group:
            flops = 2.0 * min(self.stack_group, B) * T * 8 * Cc * Cc * L      # the whole residual stack of one group per launch
            dv = self.dtab[0]

            def launch_all():
                self.flags.zero_()
                for gi, b0 in enumerate(range(0, B, self.stack_group)):
                    b1 = min(B, b0 + self.stack_group)
                    r0 = b0 * T
                    tab = self.cond_groups[gi]
                    C.tc_wavenet_stack(self.y_h[r0:], self.y2_h[r0:], e.w_dil_h, tab, 2 * Cc, tab.shape[1] * 2 * Cc,
                                       e.w_out_h, e.b_out, self.x[r0:], self.skip[r0:], self.skip_h[r0:], dv, 0, e.dilations,
                                       b1 - b0, T, Cc, self.flags[b0 * self.tpb:], e.bf16)
            return (f'wavenet_stack_kernel<{e.precision}> (b2s_tc_wavenet_stack, {L} layers per launch)', flops, launch_all,
                    -(-B // self.stack_group))
        if self.fused:
            flops = 2.0 * self.rows * 8 * Cc * Cc              # conv 6C^2 + output projection 2C^2 MACs per frame
            dv = self.dtab[0, :Cc]

            def launch_all():
                ya, yb = self.y_h, self.y2_h
                for l in range(L):
                    C.tc_wavenet_layer(ya, e.w_dil_h[l], self.cond[l], 2 * Cc, e.w_out_h[l], e.b_out[l],
                                       self.x, yb, self.skip, None, dv, 0, l == 0, B, T, Cc, e.dilations[l], e.bf16)
                    ya, yb = yb, ya
            return f'wavenet_layer_kernel<{e.precision}> (b2s_tc_wavenet_layer)', flops, launch_all, L
        flops = 2.0 * self.rows * (3 * Cc) * (2 * Cc)

        def launch_all():
            for l in range(L):
                C.tc_wavenet_gate(self.y_h, e.w_dil_h[l], self.cond[l], 2 * Cc, self.z_h, B, T, Cc, e.dilations[l], e.bf16)
        return f'tc_gemm_kernel<EPI_GATE,{e.precision}> (b2s_tc_wavenet_gate)', flops, launch_all, L


# =====================================================================================================
# LYNXNet
# =====================================================================================================
class LYNXNetEngine:
    """Packed weights + launch sequence for ``LYNXNet.forward`` (reference lynxnet.py:128-163)."""

    ACT_CODES = {'PReLU': 0, 'SiLU': C.ACT_SILU, 'ReLU': C.ACT_RELU}

    def __init__(self, net, precision: str = 'fp32'):
        if precision not in PRECISIONS:
            raise ValueError(f'unknown precision {precision!r}; available: {PRECISIONS}')
        self.net = net
        self.precision = precision
        self.C = net.num_channels
        self.L = net.num_layers
        self.MF = net.in_dims * net.n_feats
        self.H = net.hidden_size
        self.inner = net.num_channels * net.expansion_factor
        self.ksize = net.kernel_size
        self.strong = bool(net.strong_cond)
        self.act = self.ACT_CODES[net.activation]
        self._packed_version = None
        self.device = None

    def _version(self):
        # version counters catch in-place updates (optimizer steps, load_state_dict, p.data.copy_ does NOT bump them but the
        # storage pointer / dtype part catches p.data = t, .half(), .to()); see _B2SBackbone.invalidate() for the rest
        ps = list(self.net.parameters())
        return tuple((p._version, p.data_ptr(), p.dtype) for p in ps) + (str(ps[0].device),)

    def pack(self, force=False):
        v = self._version()
        if not force and v == self._packed_version:
            return
        net = self.net
        dev = next(net.parameters()).device
        if dev.type != 'cuda':
            raise C.B2SError('the denoiser lives on the CPU; this path has no CPU fallback - move the module to a CUDA device')
        f = lambda t: _dev_f32(t, dev)
        self.device = dev
        inner = self.inner
        self.w_in, self.b_in = f(net.input_projection.weight[:, :, 0]), f(net.input_projection.bias)
        de = net.diffusion_embedding
        self.w_e1, self.b_e1 = f(de[1].weight), f(de[1].bias)
        self.w_e3, self.b_e3 = f(de[3].weight), f(de[3].bias)
        layers = net.residual_layers
        self.w_dp = f(torch.cat([l.diffusion_projection.weight[:, :, 0] for l in layers], 0))   # [L*C, C]
        self.b_dp = f(torch.cat([l.diffusion_projection.bias for l in layers], 0))
        self.w_cond = f(torch.cat([l.conditioner_projection.weight[:, :, 0] for l in layers], 0))   # [L*C, H]
        self.b_cond = f(torch.cat([l.conditioner_projection.bias for l in layers], 0))
        perm = torch.stack([torch.arange(inner), torch.arange(inner) + inner], 1).reshape(-1).to(dev)
        self.ln_g = f(torch.stack([l.convmodule.net[0].weight for l in layers], 0))
        self.ln_b = f(torch.stack([l.convmodule.net[0].bias for l in layers], 0))
        self.w_up = f(torch.stack([l.convmodule.net[2].weight[:, :, 0][perm] for l in layers], 0))   # [L, 2*inner, C]
        self.b_up = f(torch.stack([l.convmodule.net[2].bias[perm] for l in layers], 0))
        self.w_dw = f(torch.stack([l.convmodule.net[4].weight[:, 0, :] for l in layers], 0))         # [L, inner, k]
        self.b_dw = f(torch.stack([l.convmodule.net[4].bias for l in layers], 0))
        if self.act == 0:
            self.slope = f(torch.stack([l.convmodule.net[5].weight for l in layers], 0))            # [L, inner]
        else:
            self.slope = None
        self.w_down = f(torch.stack([l.convmodule.net[6].weight[:, :, 0] for l in layers], 0))      # [L, C, inner]
        self.b_down = f(torch.stack([l.convmodule.net[6].bias for l in layers], 0))
        self.norm_g, self.norm_b = f(net.norm.weight), f(net.norm.bias)
        self.w_fin, self.b_fin = f(net.output_projection.weight[:, :, 0]), f(net.output_projection.bias)
        if self.precision != 'fp32':
            if self.C % 32 or inner % 32:
                raise C.B2SError(f'the {self.precision} tensor-core path needs num_channels and num_channels*expansion_factor '
                                 f"to be multiples of 32 (got {self.C}, {inner}); use b2s_precision='fp32'")
            hd = C.HALF_DTYPES[self.precision]
            self.bf16 = self.precision == 'bf16'
            h = lambda t: t.to(hd).contiguous()
            self.w_in_h, self.w_cond_h, self.w_up_h = h(self.w_in), h(self.w_cond), h(self.w_up)
            self.w_down_h, self.w_fin_h = h(self.w_down), h(self.w_fin)
            self.w_dw_t = self.w_dw.transpose(1, 2).contiguous()          # [L, k, inner]: k-major for coalesced loads
        self._packed_version = v

    def step_table(self, t_values):
        return _cached_step_table(self, t_values, self._step_table)

    def _step_table(self, t_values):
        K = t_values.numel()
        Cc, dev = self.C, self.device
        sin = torch.empty((K, Cc), device=dev)
        C.sinusoid(t_values, sin, K, Cc)
        h = torch.empty((K, 4 * Cc), device=dev)
        C.linear(sin, Cc, self.w_e1, Cc, self.b_e1, h, 4 * Cc, K, 4 * Cc, Cc, act=C.ACT_GELU)
        e = torch.empty((K, Cc), device=dev)
        C.linear(h, 4 * Cc, self.w_e3, 4 * Cc, self.b_e3, e, Cc, K, Cc, 4 * Cc)
        tab = torch.empty((K, self.L * Cc), device=dev)
        C.linear(e, Cc, self.w_dp, Cc, self.b_dp, tab, self.L * Cc, K, self.L * Cc, Cc)
        return tab

    def cond_table(self, cond_bth):
        B, T, H = cond_bth.shape
        N = self.L * self.C
        tab = torch.empty((B * T, N), device=self.device)
        C.linear(cond_bth, H, self.w_cond, H, self.b_cond, tab, N, B * T, N, H)
        return tab

    def begin(self, cond_bth, t_values, per_row_t=False, lens=None, table_owner=None):
        if lens is not None:
            raise C.B2SError('ragged batches (lengths=...) are implemented for the WaveNet whole-stack path only; batch equal-length segments')
        self.pack()
        C.require_cuda(cond_bth, 'cond')
        C.require_cuda(t_values, 't_values')
        self._table_owner = table_owner
        try:
            if self.precision != 'fp32':
                return LYNXNetSessionTC(self, cond_bth, t_values, per_row_t)
            return LYNXNetSession(self, cond_bth, t_values, per_row_t)
        finally:
            self._table_owner = None


class LYNXNetSession:
    def __init__(self, eng: LYNXNetEngine, cond_bth, t_values, per_row_t):
        self.eng = eng
        B, T, H = cond_bth.shape
        if H != eng.H:
            raise C.B2SError(f'condition has {H} channels, backbone expects hidden_size={eng.H}')
        if per_row_t and t_values.numel() != B:
            raise C.B2SError('per-row step embedding needs one time value per utterance')
        self.B, self.T, self.rows = B, T, B * T
        self.per_row_t = per_row_t
        self.cond = eng.cond_table(cond_bth)
        self.dtab = eng.step_table(t_values)
        dev = eng.device
        self.x = torch.empty((self.rows, eng.C), device=dev)
        self.h = torch.empty((self.rows, eng.C), device=dev)
        self.g = torch.empty((self.rows, eng.inner), device=dev)
        self.p = torch.empty((self.rows, eng.inner), device=dev)

    def _dvec(self, k, l):
        LC = self.eng.L * self.eng.C
        if self.per_row_t:
            return self.dtab[0, l * self.eng.C:], LC
        return self.dtab[k, l * self.eng.C:], 0

    def eval(self, x_in, k, out):
        e = self.eng
        B, T, rows, Cc, L, MF, inner = self.B, self.T, self.rows, e.C, e.L, e.MF, e.inner
        # input projection (+ exact GELU unless strong_cond, lynxnet.py:141-143)
        C.linear(x_in, MF, e.w_in, MF, e.b_in, self.x, Cc, rows, Cc, MF, act=C.ACT_NONE if e.strong else C.ACT_GELU)
        ldc = L * Cc
        for l in range(L):
            dv, ds = self._dvec(k, l)
            C.lynx_prenorm(self.x, self.cond[:, l * Cc:], ldc, dv, ds, e.ln_g[l], e.ln_b[l], self.h, B, T, Cc, e.strong)
            C.lynx_glu(self.h, e.w_up[l], e.b_up[l], self.g, rows, Cc, inner)
            C.lynx_dwconv(self.g, e.w_dw[l], e.b_dw[l], None if e.slope is None else e.slope[l], self.p, B, T, inner,
                          e.ksize, e.act)
            C.linear_residual(self.p, e.w_down[l], e.b_down[l], self.x, rows, Cc, inner)
        C.layernorm(self.x, e.norm_g, e.norm_b, self.h, rows, Cc)
        C.linear(self.h, Cc, e.w_fin, Cc, e.b_fin, out, MF, rows, MF, Cc)

    @property
    def launches_per_eval(self) -> int:
        return 1 + 4 * self.eng.L + 2


class LYNXNetSessionTC:
    """16-bit tensor-core session of LYNXNet (lynxnet.py:128-163): the three pointwise convolutions of every layer and
    the stem / head run on tcgen05 (SwiGLU and residual-add fused into the GEMM epilogues); LayerNorm and the depthwise
    k=31 conv are HBM-bound kernels with 16-bit I/O and fp32 math; the residual stream x stays fp32."""

    def __init__(self, eng: LYNXNetEngine, cond_bth, t_values, per_row_t):
        self.eng = eng
        B, T, H = cond_bth.shape
        if H != eng.H:
            raise C.B2SError(f'condition has {H} channels, backbone expects hidden_size={eng.H}')
        if per_row_t and t_values.numel() != B:
            raise C.B2SError('per-row step embedding needs one time value per utterance')
        if H % 8 or eng.MF % 8:
            raise C.B2SError(f'the tensor-core path needs hidden_size and in_dims*n_feats to be multiples of 8 '
                             f'(got {H}, {eng.MF})')
        self.B, self.T, self.rows = B, T, B * T
        self.per_row_t = per_row_t
        dev, hd, bf = eng.device, C.HALF_DTYPES[eng.precision], eng.bf16
        rows, Cc, L, inner = self.rows, eng.C, eng.L, eng.inner
        self.dtab = eng.step_table(t_values)
        cond_h = torch.empty((rows, H), device=dev, dtype=hd)
        C.cast_h(cond_bth, cond_h, bf)
        self.cond = torch.empty((L, rows, Cc), device=dev, dtype=hd)             # layer-major hoisted cond projection
        C.tc_cond_table(cond_h, rows, eng.w_cond_h, eng.b_cond, L, Cc, H, self.cond, bf)
        self.xin_h = torch.empty((rows, eng.MF), device=dev, dtype=hd)
        self.x = torch.empty((rows, Cc), device=dev)
        self.h_h = torch.empty((rows, Cc), device=dev, dtype=hd)
        self.g_h = torch.empty((rows, inner), device=dev, dtype=hd)
        self.p_h = torch.empty((rows, inner), device=dev, dtype=hd)

    def _dvec(self, k, l):
        LC = self.eng.L * self.eng.C
        if self.per_row_t:
            return self.dtab[0, l * self.eng.C:], LC
        return self.dtab[k, l * self.eng.C:], 0

    def half_sink(self):
        return self.xin_h, None, self.eng.bf16

    def eval(self, x_in, k, out, precast: bool = False):
        e = self.eng
        B, T, rows, Cc, L, MF, inner, bf = self.B, self.T, self.rows, e.C, e.L, e.MF, e.inner, e.bf16
        if not precast:
            C.cast_h(x_in, self.xin_h, bf)
        C.tc_linear(self.xin_h, MF, rows, T, e.w_in_h, MF, e.b_in, Cc, MF, bf,
                    act=C.ACT_NONE if e.strong else C.ACT_GELU, out_f32=self.x, ldo=Cc)
        # strong_cond: layer l+1's front_cond_inject (x += cond_{l+1}) rides on layer l's residual epilogue, so only the first
        # LayerNorm kernel reads the cond table and writes x back (halves the HBM traffic of the other L-1)
        fold = bool(e.strong) and hparams.get('b2s_lynx_fold_cond', True)
        for l in range(L):
            dv, ds = self._dvec(k, l)
            C.lynx_prenorm_h(self.x, None if (fold and l > 0) else self.cond[l], Cc, dv, ds, e.ln_g[l], e.ln_b[l], self.h_h, B, T, Cc,
                             e.strong, bf)
            C.tc_lynx_glu(self.h_h, e.w_up_h[l], e.b_up[l], self.g_h, rows, Cc, inner, bf)
            C.lynx_dwconv_h(self.g_h, e.w_dw_t[l], e.b_dw[l], None if e.slope is None else e.slope[l], self.p_h, B, T, inner,
                            e.ksize, e.act, bf)
            if fold and l + 1 < L:
                C.tc_linear_residual_cond(self.p_h, e.w_down_h[l], e.b_down[l], self.x, self.cond[l + 1], Cc, rows, Cc, inner, bf)
            else:
                C.tc_linear_residual(self.p_h, e.w_down_h[l], e.b_down[l], self.x, rows, Cc, inner, bf)
        C.layernorm_h(self.x, e.norm_g, e.norm_b, self.h_h, rows, Cc, bf)
        C.tc_linear(self.h_h, Cc, rows, T, e.w_fin_h, Cc, e.b_fin, MF, Cc, bf, out_f32=out, ldo=MF)

    @property
    def launches_per_eval(self) -> int:
        return 2 + 4 * self.eng.L + 2

    def dominant_kernel(self, w=None):
        """(name, algorithmic FLOPs per launch, callable launching it once per layer, launches) for bench.py's roofline: the
        SwiGLU up-projection GEMM (lynxnet.py:55-56), 2/3 of the layer's FLOPs."""
        e = self.eng
        flops = 2.0 * self.rows * e.C * (2 * e.inner)

        def launch_all():
            for l in range(e.L):
                C.tc_lynx_glu(self.h_h, e.w_up_h[l], e.b_up[l], self.g_h, self.rows, e.C, e.inner, e.bf16)
        return f'tc_gemm_cg2_kernel<EPI_SWIGLU,{e.precision}> (b2s_tc_lynx_glu)', flops, launch_all, e.L


# =====================================================================================================
# sampler-program executor
# =====================================================================================================
class CompiledProgram:
    """A Program bound to a device: flat coefficient table in HBM, per-op offsets."""

    def __init__(self, prog: Program, device):
        self.prog = prog
        flat: List[float] = []
        self.offsets: List[int] = []
        for op in prog.ops:
            self.offsets.append(len(flat))
            if op.kind == 'lin':
                flat.extend(c for _, c in op.terms)
                while len(flat) % 4:                       # keep rows 16B aligned
                    flat.append(0.0)
        self.coef = torch.tensor(flat if flat else [0.0], dtype=torch.float64).to(torch.float32).to(device)
        self.t_values = torch.tensor(prog.t_values, dtype=torch.float32, device=device)
        self.n_lin = sum(1 for op in prog.ops if op.kind == 'lin')


def _fusable_update(prog: Program, i: int):
    """ops[i] is an nfe.  Returns (index of the step's noise op or None, index of the lin op) when the update that follows
    can run inside the denoiser launch: ``x <- c0 x + c1 eps_hat (+ c2 z)`` written back to the evaluation's own input
    buffer (ancestral and DDIM-type steps, ddpm.py:149-167), with eps_hat not needed by anything else."""
    ops = prog.ops
    nfe = ops[i]
    j, noise = i + 1, None
    if j < len(ops) and ops[j].kind == 'noise':
        noise, j = j, j + 1
    if j >= len(ops) or ops[j].kind != 'lin':
        return None
    lin = ops[j]
    names = [b for b, _ in lin.terms]
    if lin.dst != nfe.src or nfe.dst not in names or len(names) > 3 or len(set(names)) != len(names):
        return None
    zname = ops[noise].dst if noise is not None else None
    if zname in (nfe.src, nfe.dst) or any(b not in (nfe.src, nfe.dst, zname) for b in names):
        return None
    for op in ops[j + 1:]:                      # eps_hat is never materialised: nobody may read it afterwards
        reads = [op.src] if op.kind == 'nfe' else ([b for b, _ in op.terms] if op.kind == 'lin' else [])
        if nfe.dst in reads:
            return None
        if op.dst == nfe.dst:
            break
    else:
        if prog.result == nfe.dst:
            return None
    return noise, j


def run_program(cp: CompiledProgram, session, bufs: Dict[str, torch.Tensor],
                draw_noise: Optional[Callable[[int, torch.Tensor], None]] = None) -> torch.Tensor:
    """Executes the program.  ``bufs`` maps buffer names to [B*T, MF] fp32 device tensors (NOISE0 /
    XSTART pre-filled by the caller); ``draw_noise(j, dst)`` fills ``dst`` with the j-th per-step draw.

    Launch structure (none of it changes a result bit - ``test_launch_structure_switches_do_not_change_results``):
    * ``b2s_fuse_update`` (OFF by default): an evaluation followed by ``x <- c0 x + c1 eps_hat (+ c2 z)`` is ONE launch - the
      update runs in the head epilogue of the whole-denoiser kernel, which also writes the next evaluation's 16-bit input and
      re-arms the next launch's tile flags; an ancestral step is then one kernel on the critical path.  Measured 1 % SLOWER
      than the separate update launch (the update's operand loads are a latency chain in 96 CTAs' tails, the separate kernel
      streams them on all 148 SMs in 6 us), so it stays a tested switch.
    * ``b2s_overlap_noise``: per-step noise draws run on a side stream, concurrently with a denoiser launch that does not
      need them (the draw of step s+1 during the fused launch of step s, into the other of two noise buffers).  The draws keep
      their call order, so the random stream is unchanged; under CUDA-graph capture the fork / join is a parallel branch.
    * ``b2s_fuse_cast``: any other update that feeds an evaluation writes its 16-bit input itself."""
    prog = cp.prog
    ops = prog.ops
    dev = bufs[prog.result].device
    main = torch.cuda.current_stream()
    side = None
    overlap = hparams.get('b2s_overlap_noise', True)
    use_sink = hasattr(session, 'half_sink') and hparams.get('b2s_fuse_cast', True)
    fuse_upd = hparams.get('b2s_fuse_update', False) and getattr(session, 'can_fuse_update', False)
    done, precast = set(), set()
    pre_drawn: Dict[int, torch.Tensor] = {}     # noise op index -> buffer being filled on the side stream
    zb, zflip = None, 0
    for i, (op, off) in enumerate(zip(ops, cp.offsets)):
        if i in done:
            continue
        if op.kind == 'nfe':
            fus = _fusable_update(prog, i) if fuse_upd else None
            if fus is not None:
                ni, j = fus
                lin = ops[j]
                zsel = None
                if ni is not None:
                    if zb is None:
                        bufs['__noise_alt'] = torch.empty_like(bufs[ops[ni].dst])
                        zb = [bufs[ops[ni].dst], bufs['__noise_alt']]
                    if ni in pre_drawn:
                        zsel = pre_drawn.pop(ni)
                        main.wait_stream(side)
                    else:
                        zsel, zflip = zb[zflip], zflip ^ 1
                        draw_noise(ops[ni].draw, zsel)
                    done.add(ni)
                srcs = [None if b == op.dst else (zsel if (ni is not None and b == ops[ni].dst) else bufs[b]) for b, _ in lin.terms]
                ev = None
                if overlap:
                    ev = torch.cuda.Event()
                    ev.record(main)                     # everything before this launch (the reader of the other noise buffer) is done
                n = len(lin.terms)
                session.eval_update(bufs[op.src], op.t_index, srcs, cp.coef[cp.offsets[j]:cp.offsets[j] + n], bufs[lin.dst],
                                    precast=i in precast)
                done.add(j)
                k = j + 1
                if k < len(ops) and ops[k].kind == 'nfe' and ops[k].src == lin.dst:
                    precast.add(k)
                    f2 = _fusable_update(prog, k) if overlap else None
                    if f2 is not None and f2[0] is not None and zb is not None:
                        if side is None:
                            side = torch.cuda.Stream(device=dev)
                        ztarget, zflip = zb[zflip], zflip ^ 1
                        side.wait_event(ev)
                        with torch.cuda.stream(side):
                            draw_noise(ops[f2[0]].draw, ztarget)
                        pre_drawn[f2[0]] = ztarget
                continue
            nxt = ops[i + 1] if i + 1 < len(ops) else None
            joined = False
            if nxt is not None and nxt.kind == 'noise' and nxt.dst not in (op.src, op.dst) and overlap:
                if side is None:
                    side = torch.cuda.Stream(device=dev)
                side.wait_stream(main)                      # earlier readers of the noise buffer are done
                with torch.cuda.stream(side):
                    draw_noise(nxt.draw, bufs[nxt.dst])
                done.add(i + 1)
                joined = True
            if i in precast:
                session.eval(bufs[op.src], op.t_index, bufs[op.dst], precast=True)
            else:
                session.eval(bufs[op.src], op.t_index, bufs[op.dst])
            if joined:
                main.wait_stream(side)
        elif op.kind == 'lin':
            n = len(op.terms)
            nxt = ops[i + 1] if i + 1 < len(ops) else None
            if use_sink and nxt is not None and nxt.kind == 'nfe' and nxt.src == op.dst:
                xin_h, flags, bf = session.half_sink()
                C.lincomb_h(bufs[op.dst], [bufs[b] for b, _ in op.terms], cp.coef[off:off + n], xin_h, bf, reset_flags=flags)
                precast.add(i + 1)
            else:
                C.lincomb(bufs[op.dst], [bufs[b] for b, _ in op.terms], cp.coef[off:off + n])
        else:
            draw_noise(op.draw, bufs[op.dst])
    return bufs[prog.result]


def allocate_buffers(prog: Program, rows: int, mf: int, device) -> Dict[str, torch.Tensor]:
    return {name: torch.empty((rows, mf), device=device) for name in prog.buffers()}
