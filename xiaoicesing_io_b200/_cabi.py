"""ctypes binding of libb2s.so (the C ABI declared in include/b2s.h).

This is the ONLY way the Python host code reaches the GPU kernels; there is no CPU or PyTorch fallback.
If the shared library is missing the import fails loudly.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p

import torch

from . import _build

_vp, _i, _f, _i64 = c_void_p, c_int, c_float, c_int64

# name -> argtypes (all return int except where noted); mirrors include/b2s.h line by line
_SIGNATURES = {
    'b2s_transpose_f32': [_vp, _vp, _i, _i, _i, _vp],
    'b2s_spec_norm_f32': [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    'b2s_spec_denorm_f32': [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    'b2s_sampler_lincomb_f32': [_vp, ctypes.POINTER(_vp), _vp, _i, _i64, _vp],
    'b2s_sampler_lincomb_f32_h': [_vp, ctypes.POINTER(_vp), _vp, _i, _i64, _vp, _i, _vp, _i, _vp],
    'b2s_sinusoid_f32': [_vp, _vp, _i, _i, _vp],
    'b2s_linear_f32': [_vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp, _vp, _i, _i, _vp],
    'b2s_wavenet_gate_f32': [_vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _vp],
    'b2s_wavenet_out_f32': [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    'b2s_lynx_prenorm_f32': [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    'b2s_layernorm_f32': [_vp, _vp, _vp, _vp, _i, _i, _vp],
    'b2s_lynx_glu_f32': [_vp, _vp, _vp, _vp, _i, _i, _i, _vp],
    'b2s_lynx_dwconv_f32': [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    'b2s_linear_residual_f32': [_vp, _vp, _vp, _vp, _i, _i, _i, _vp],
    # 16-bit tensor-core path
    'b2s_cast_f32_h': [_vp, _vp, _i64, _i, _vp],
    'b2s_cast_f32_h_reset': [_vp, _vp, _i64, _vp, _i, _i, _vp],
    'b2s_tc_linear': [_vp, _i, _i, _i, _vp, _i, _vp, _i, _i, _f, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _vp],
    'b2s_tc_cond_table': [_vp, _i, _vp, _vp, _i, _i, _i, _vp, _i, _vp],
    'b2s_tc_cond_table_tiled': [_vp, _i, _i, _vp, _vp, _i, _i, _i, _vp, _i, _vp],
    'b2s_tc_wavenet_gate_ld': [_vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp],
    'b2s_tc_skip_sum': [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    'b2s_tc_wavenet_res': [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    'b2s_tc_wavenet_gate': [_vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp],
    'b2s_tc_wavenet_out': [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    'b2s_tc_wavenet_layer': [_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    'b2s_tc_wavenet_stack': [_vp, _vp, _vp, _vp, _i, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i, ctypes.POINTER(_i), _i, _i, _i, _i, _vp,
                             _i, _vp],
    'b2s_tc_wavenet_stack_t_tiles': [_i, _i, _i],
    'b2s_tc_cond_retile': [_vp, _i, _i, _i, _i, _i, _vp, _vp],
    'b2s_tc_wavenet_stack_t': [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, ctypes.POINTER(_i), _i, _i, _i, _i, _i, _vp, _i, _vp],
    'b2s_tc_wavenet_denoiser': [_vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i, ctypes.POINTER(_i), _i,
                                _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _i, _vp],
    'b2s_tc_wavenet_denoiser_update': [_vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i, ctypes.POINTER(_i), _i,
                                       _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _i, ctypes.POINTER(_vp), _vp, _vp, _vp, _i, _vp],
    'b2s_tc_wavenet_stack3': [_vp, _i, _vp, _i, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i, ctypes.POINTER(_i), _i, _vp, _vp, _vp, _i64,
                              _i, _i, _i, _vp, _vp, _i, _vp],
    'b2s_tc_wavenet_denoiser3': [_vp, _i, _vp, _i, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i, ctypes.POINTER(_i), _i, _vp, _vp, _vp, _i64,
                                 _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _vp],
    'b2s_tc_wavenet_denoiser3_chained': [_vp, _i, _vp, _i, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i, ctypes.POINTER(_i), _i, _vp, _vp, _vp, _i64,
                                         _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp],
    'b2s_tc_lynx_glu': [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    'b2s_tc_linear_residual_cond': [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    'b2s_tc_linear_residual': [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    'b2s_lynx_prenorm_h': [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    'b2s_layernorm_h': [_vp, _vp, _vp, _vp, _i, _i, _i, _vp],
    'b2s_lynx_dwconv_h': [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    'b2s_tc_conv1d': [_vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    'b2s_layernorm_hh': [_vp, _vp, _vp, _vp, _i, _i, _f, _i, _vp],
    'b2s_tc_linear_residual_scaled': [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    'b2s_enc_mel2ph_to_dur': [_vp, _vp, _i, _i, _i, _vp],
    'b2s_enc_embed': [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp],
    'b2s_enc_rope': [_vp, _vp, _i, _i, _i, _i, _vp],
    'b2s_enc_attention': [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    'b2s_enc_mask_rows': [_vp, _vp, _i, _i, _vp],
    'b2s_enc_layernorm_mask': [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp],
    'b2s_tc_conv1d_dil': [_vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    'b2s_tc_conv1d_residual': [_vp, _vp, _vp, _vp, _vp, _vp, _f, _i, _i, _i, _i, _i, _i, _i, _vp],
    'b2s_voc_phase': [_vp, _vp, _i, _i, _f, _i, _i, _vp],
    'b2s_voc_source': [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _f, _f, _vp],
    'b2s_voc_source_add': [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _i, _vp],
    'b2s_voc_avg_act': [ctypes.POINTER(_vp), _i, _vp, _i64, _f, _i, _vp],
    'b2s_voc_post': [ctypes.POINTER(_vp), _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp],
    'b2s_cast_scale_f32_h': [_vp, _vp, _i64, _f, _i, _vp],
    'b2s_masked_loss_f32': [_vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp],
    'b2s_enc_assemble': [_vp, _vp, _vp, _i, ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp), _i, _i, _i, _vp, _i, _i, _i, _i, _vp],
}

# entry points of B2S_BUILD_EXPERIMENTS=1 builds only (measured-and-rejected variants)
EXPERIMENTAL = ('b2s_tc_wavenet_stack_t_tiles', 'b2s_tc_cond_retile', 'b2s_tc_wavenet_stack_t', 'b2s_tc_wavenet_denoiser_update')

EXPORTED_SYMBOLS = ['b2s_masked_loss_workspace_bytes', 'b2s_abi_version', 'b2s_last_error', 'b2s_has_experiments', 'b2s_tc_wavenet_stack_max_tiles', 'b2s_tc_wavenet_stack3_halo', 'b2s_tc_wavenet_stack3_max_tiles', 'b2s_tc_wavenet_denoiser3_max_utterances', *[n for n in _SIGNATURES if n not in EXPERIMENTAL]]

ACT_NONE, ACT_RELU, ACT_MISH, ACT_GELU, ACT_SILU, ACT_LRELU = 0, 1, 2, 3, 4, 5


class B2SError(RuntimeError):
    pass


def _load():
    path = os.environ.get('B2S_LIB', _build.LIB_PATH)       # B2S_LIB: A/B measurements of two builds on the same GPU box
    if not os.path.exists(path):
        raise ImportError(
            f'{path} is missing: the CUDA extension has not been built.  Run '
            f'`python -c "import __graft_entry__ as g; g.build()"` (needs nvcc).  '
            f'There is no CPU fallback for this path.')
    lib = ctypes.CDLL(path)
    lib.b2s_abi_version.restype = c_int
    lib.b2s_abi_version.argtypes = []
    lib.b2s_last_error.restype = c_char_p
    lib.b2s_last_error.argtypes = []
    lib.b2s_tc_wavenet_stack_max_tiles.restype = c_int
    lib.b2s_tc_wavenet_stack_max_tiles.argtypes = []
    lib.b2s_tc_wavenet_stack3_halo.restype = c_int
    lib.b2s_tc_wavenet_stack3_halo.argtypes = []
    lib.b2s_tc_wavenet_stack3_max_tiles.restype = c_int
    lib.b2s_tc_wavenet_stack3_max_tiles.argtypes = [c_int, c_int]
    lib.b2s_tc_wavenet_denoiser3_max_utterances.restype = c_int
    lib.b2s_tc_wavenet_denoiser3_max_utterances.argtypes = [c_int, c_int]
    lib.b2s_masked_loss_workspace_bytes.restype = c_int
    lib.b2s_masked_loss_workspace_bytes.argtypes = []
    lib.b2s_has_experiments.restype = c_int
    lib.b2s_has_experiments.argtypes = []
    for name, args in _SIGNATURES.items():
        if name in EXPERIMENTAL and not lib.b2s_has_experiments():
            continue
        fn = getattr(lib, name)
        fn.restype = c_int
        fn.argtypes = args
    return lib


lib = _load()
LIB_PATH = _build.LIB_PATH
HAS_EXPERIMENTS = bool(lib.b2s_has_experiments())


def require_experiments(what: str):
    if not HAS_EXPERIMENTS:
        raise B2SError(f'{what} is a measured-and-rejected experiment (DESIGN.md section 3.3): rebuild libb2s.so with '
                       f'B2S_BUILD_EXPERIMENTS=1 to use it')


N_CALLS = 0     # successful kernel-launching C-ABI calls so far (every entry point of the sampling path launches exactly one kernel; the
#                 whole-denoiser entry launches two and b2s_masked_loss_f32, outside the sampling loop, two)


def check(rc: int, what: str = ''):
    global N_CALLS
    N_CALLS += 1
    if rc != 0:
        raise B2SError(f'{what} failed (code {rc}): {lib.b2s_last_error().decode()}')


def ptr(t):
    """Device pointer of a tensor (or None)."""
    if t is None:
        return None
    return c_void_p(t.data_ptr())


def stream_ptr():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(t: torch.Tensor, name: str, dtype=torch.float32):
    if not t.is_cuda:
        raise B2SError(f'{name} must be a CUDA tensor: this path has no CPU fallback (got {t.device})')
    if t.dtype != dtype:
        raise B2SError(f'{name} must be {dtype} (got {t.dtype})')
    if not t.is_contiguous():
        raise B2SError(f'{name} must be contiguous')
    return t


# ---- thin typed wrappers (argument order = include/b2s.h) ------------------------------------------
def transpose(inp, out, batch, rows, cols):
    check(lib.b2s_transpose_f32(ptr(inp), ptr(out), batch, rows, cols, stream_ptr()), 'b2s_transpose_f32')


def spec_norm(spec, lo, hi, state, B, F, T, M):
    check(lib.b2s_spec_norm_f32(ptr(spec), ptr(lo), ptr(hi), ptr(state), B, F, T, M, stream_ptr()), 'b2s_spec_norm_f32')


def spec_denorm(state, lo, hi, spec, B, F, T, M):
    check(lib.b2s_spec_denorm_f32(ptr(state), ptr(lo), ptr(hi), ptr(spec), B, F, T, M, stream_ptr()), 'b2s_spec_denorm_f32')


def lincomb(dst, srcs, coef_dev):
    """dst = sum_i coef_dev[i] * srcs[i]; coef_dev is a device fp32 tensor (view) of len(srcs) entries."""
    arr = (_vp * len(srcs))(*[s.data_ptr() for s in srcs])
    check(lib.b2s_sampler_lincomb_f32(ptr(dst), arr, ptr(coef_dev), len(srcs), dst.numel(), stream_ptr()),
          'b2s_sampler_lincomb_f32')


def lincomb_h(dst, srcs, coef_dev, out_h, bf16, reset_flags=None):
    """lincomb + the 16-bit copy of dst the tensor-core denoiser reads next (+ tile-flag reset), one launch."""
    arr = (_vp * len(srcs))(*[s.data_ptr() for s in srcs])
    nf = 0 if reset_flags is None else reset_flags.numel()
    check(lib.b2s_sampler_lincomb_f32_h(ptr(dst), arr, ptr(coef_dev), len(srcs), dst.numel(), ptr(out_h), int(bf16),
                                        ptr(reset_flags), nf, stream_ptr()), 'b2s_sampler_lincomb_f32_h')


def sinusoid(t, out, n, dim):
    check(lib.b2s_sinusoid_f32(ptr(t), ptr(out), n, dim, stream_ptr()), 'b2s_sinusoid_f32')


def linear(A, lda, W, ldw, bias, out, ldo, M, N, K, alpha=1.0, act=ACT_NONE, y=None, dvec=None, d_stride=0, T=0):
    check(lib.b2s_linear_f32(ptr(A), lda, ptr(W), ldw, ptr(bias), ptr(out), ldo, M, N, K, alpha, act,
                             ptr(y), ptr(dvec), d_stride, T, stream_ptr()), 'b2s_linear_f32')


def wavenet_gate(y, Wd, cond, ld_cond, z, B, T, C, dilation):
    check(lib.b2s_wavenet_gate_f32(ptr(y), ptr(Wd), ptr(cond), ld_cond, ptr(z), B, T, C, dilation, stream_ptr()),
          'b2s_wavenet_gate_f32')


def wavenet_out(z, Wo, bo, x, y_next, skip, dvec_next, d_stride, first, B, T, C):
    check(lib.b2s_wavenet_out_f32(ptr(z), ptr(Wo), ptr(bo), ptr(x), ptr(y_next), ptr(skip), ptr(dvec_next),
                                  d_stride, int(first), B, T, C, stream_ptr()), 'b2s_wavenet_out_f32')


def lynx_prenorm(x, cond, ld_cond, dvec, d_stride, gamma, beta, h, B, T, C, strong):
    check(lib.b2s_lynx_prenorm_f32(ptr(x), ptr(cond), ld_cond, ptr(dvec), d_stride, ptr(gamma), ptr(beta), ptr(h),
                                   B, T, C, int(strong), stream_ptr()), 'b2s_lynx_prenorm_f32')


def layernorm(x, gamma, beta, h, rows, C):
    check(lib.b2s_layernorm_f32(ptr(x), ptr(gamma), ptr(beta), ptr(h), rows, C, stream_ptr()), 'b2s_layernorm_f32')


def lynx_glu(h, W, bias, g, rows, C, inner):
    check(lib.b2s_lynx_glu_f32(ptr(h), ptr(W), ptr(bias), ptr(g), rows, C, inner, stream_ptr()), 'b2s_lynx_glu_f32')


def lynx_dwconv(g, Wdw, bias, slope, p, B, T, inner, ksize, act):
    check(lib.b2s_lynx_dwconv_f32(ptr(g), ptr(Wdw), ptr(bias), ptr(slope), ptr(p), B, T, inner, ksize, act,
                                  stream_ptr()), 'b2s_lynx_dwconv_f32')


def linear_residual(p, W, bias, x, rows, C, inner):
    check(lib.b2s_linear_residual_f32(ptr(p), ptr(W), ptr(bias), ptr(x), rows, C, inner, stream_ptr()),
          'b2s_linear_residual_f32')


# ---- 16-bit tensor-core path ------------------------------------------------------------------------
HALF_DTYPES = {'bf16': torch.bfloat16, 'fp16': torch.float16}


def cast_h(inp, out, bf16, reset_flags=None):
    """fp32 -> 16-bit cast; ``reset_flags`` (int32 tensor) is zeroed in the same launch."""
    if reset_flags is None:
        check(lib.b2s_cast_f32_h(ptr(inp), ptr(out), inp.numel(), int(bf16), stream_ptr()), 'b2s_cast_f32_h')
    else:
        check(lib.b2s_cast_f32_h_reset(ptr(inp), ptr(out), inp.numel(), ptr(reset_flags), reset_flags.numel(), int(bf16),
                                       stream_ptr()), 'b2s_cast_f32_h_reset')


def tc_linear(A, lda, rows, T, W, ldw, bias, N, K, bf16, alpha=1.0, act=ACT_NONE, out_f32=None, ldo=0, out_h=None,
              ldoh=0, y_h=None, ldy=0, dvec=None, d_stride=0):
    check(lib.b2s_tc_linear(ptr(A), lda, rows, T, ptr(W), ldw, ptr(bias), N, K, alpha, act, ptr(out_f32), ldo,
                            ptr(out_h), ldoh, ptr(y_h), ldy, ptr(dvec), d_stride, int(bf16), stream_ptr()),
          'b2s_tc_linear')


def tc_cond_table(cond_h, rows, Wc_h, bc, L, N2, H, table_h, bf16):
    check(lib.b2s_tc_cond_table(ptr(cond_h), rows, ptr(Wc_h), ptr(bc), L, N2, H, ptr(table_h), int(bf16), stream_ptr()),
          'b2s_tc_cond_table')


def tc_cond_table_tiled(cond_h, B, T, Wc_h, bc, L, N2, H, table_h, bf16):
    check(lib.b2s_tc_cond_table_tiled(ptr(cond_h), B, T, ptr(Wc_h), ptr(bc), L, N2, H, ptr(table_h), int(bf16),
                                      stream_ptr()), 'b2s_tc_cond_table_tiled')


def tc_wavenet_gate(y_h, Wd_h, cond_h, ld_cond, z_h, B, T, C, dilation, bf16):
    check(lib.b2s_tc_wavenet_gate(ptr(y_h), ptr(Wd_h), ptr(cond_h), ld_cond, ptr(z_h), B, T, C, dilation, int(bf16),
                                  stream_ptr()), 'b2s_tc_wavenet_gate')


def tc_wavenet_out(z_h, Wo_h, bo, x, y_next_h, skip, skip_h, dvec_next, d_stride, first, B, T, C, bf16):
    check(lib.b2s_tc_wavenet_out(ptr(z_h), ptr(Wo_h), ptr(bo), ptr(x), ptr(y_next_h), ptr(skip), ptr(skip_h),
                                 ptr(dvec_next), d_stride, int(first), B, T, C, int(bf16), stream_ptr()),
          'b2s_tc_wavenet_out')


FUSED_LAYER_CHANNELS = 256


def tc_wavenet_layer(y_h, Wd_h, cond_h, ld_cond, Wo_h, bo, x, y_next_h, skip, skip_h, dvec_next, d_stride, first, B, T,
                     C, dilation, bf16):
    check(lib.b2s_tc_wavenet_layer(ptr(y_h), ptr(Wd_h), ptr(cond_h), ld_cond, ptr(Wo_h), ptr(bo), ptr(x), ptr(y_next_h),
                                   ptr(skip), ptr(skip_h), ptr(dvec_next), d_stride, int(first), B, T, C, dilation,
                                   int(bf16), stream_ptr()), 'b2s_tc_wavenet_layer')


def tc_wavenet_stack(y0_h, y1_h, Wd_h, cond_h, ld_cond, cond_layer_stride, Wo_h, bo, x, skip, skip_h, dvec, d_stride,
                     dilations, B, T, C, flags, bf16):
    L = len(dilations)
    dil = (_i * L)(*dilations)
    check(lib.b2s_tc_wavenet_stack(ptr(y0_h), ptr(y1_h), ptr(Wd_h), ptr(cond_h), ld_cond, cond_layer_stride, ptr(Wo_h),
                                   ptr(bo), ptr(x), ptr(skip), ptr(skip_h), ptr(dvec), d_stride, dil, L, B, T, C,
                                   ptr(flags), int(bf16), stream_ptr()), 'b2s_tc_wavenet_stack')


def tc_wavenet_denoiser_update(xin_h, MF, Win_h, ld_win, b_in, y0_h, y1_h, Wd_h, cond_h, cond_layer_stride, Wo_h, bo, x, skip, dvec,
                               d_stride, dilations, Wsp_h, b_sp, Wfin_h, b_fin, B, T, C, flags, flags_next, srcs, coef, x_out,
                               x_out_h, bf16):
    """srcs: fp32 tensors or None (None = this evaluation's output), in term order; coef: device fp32 view of len(srcs)."""
    require_experiments('b2s_tc_wavenet_denoiser_update')
    L = len(dilations)
    dil = (_i * L)(*dilations)
    arr = (_vp * len(srcs))(*[None if t is None else t.data_ptr() for t in srcs])
    check(lib.b2s_tc_wavenet_denoiser_update(ptr(xin_h), MF, ptr(Win_h), ld_win, ptr(b_in), ptr(y0_h), ptr(y1_h), ptr(Wd_h),
                                             ptr(cond_h), cond_layer_stride, ptr(Wo_h), ptr(bo), ptr(x), ptr(skip), ptr(dvec),
                                             d_stride, dil, L, ptr(Wsp_h), ptr(b_sp), ptr(Wfin_h), ptr(b_fin), B, T, C, ptr(flags),
                                             ptr(flags_next), len(srcs), arr, ptr(coef), ptr(x_out), ptr(x_out_h), int(bf16),
                                             stream_ptr()), 'b2s_tc_wavenet_denoiser_update')


def tc_wavenet_gate_ld(y_h, Wd_h, cond_h, ld_cond, z_h, ld_z, B, T, C, dilation, bf16):
    check(lib.b2s_tc_wavenet_gate_ld(ptr(y_h), ptr(Wd_h), ptr(cond_h), ld_cond, ptr(z_h), ld_z, B, T, C, dilation, int(bf16),
                                     stream_ptr()), 'b2s_tc_wavenet_gate_ld')


def tc_wavenet_res(z_h, ld_z, Wres_h, b_res, x, y_next_h, dvec_next, d_stride, B, T, C, bf16):
    check(lib.b2s_tc_wavenet_res(ptr(z_h), ld_z, ptr(Wres_h), ptr(b_res), ptr(x), ptr(y_next_h), ptr(dvec_next), d_stride, B, T, C,
                                 int(bf16), stream_ptr()), 'b2s_tc_wavenet_res')


def tc_skip_sum(z_all_h, Wcat_h, bias, out_h, rows, C, L, bf16):
    check(lib.b2s_tc_skip_sum(ptr(z_all_h), ptr(Wcat_h), ptr(bias), ptr(out_h), rows, C, L, int(bf16), stream_ptr()), 'b2s_tc_skip_sum')


def tc_cond_retile(table_h, L, B, T, n2, NT, out_h):
    require_experiments('b2s_tc_cond_retile')
    check(lib.b2s_tc_cond_retile(ptr(table_h), L, B, T, n2, NT, ptr(out_h), stream_ptr()), 'b2s_tc_cond_retile')


def tc_wavenet_stack_t(y0_h, y1_h, Wd_h, cond_t, Wo_h, bo, x, skip_h, dvec, d_stride, dilations, B, T, C, NT, flags, bf16):
    require_experiments('b2s_tc_wavenet_stack_t')
    L = len(dilations)
    dil = (_i * L)(*dilations)
    check(lib.b2s_tc_wavenet_stack_t(ptr(y0_h), ptr(y1_h), ptr(Wd_h), ptr(cond_t), ptr(Wo_h), ptr(bo), ptr(x), ptr(skip_h),
                                     ptr(dvec), d_stride, dil, L, B, T, C, NT, ptr(flags), int(bf16), stream_ptr()),
          'b2s_tc_wavenet_stack_t')


def tc_wavenet_denoiser(xin_h, MF, Win_h, ld_win, b_in, y0_h, y1_h, Wd_h, cond_h, cond_layer_stride, Wo_h, bo, x, skip, dvec,
                        d_stride, dilations, Wsp_h, b_sp, Wfin_h, b_fin, out, B, T, C, flags, bf16):
    L = len(dilations)
    dil = (_i * L)(*dilations)
    check(lib.b2s_tc_wavenet_denoiser(ptr(xin_h), MF, ptr(Win_h), ld_win, ptr(b_in), ptr(y0_h), ptr(y1_h), ptr(Wd_h), ptr(cond_h),
                                      cond_layer_stride, ptr(Wo_h), ptr(bo), ptr(x), ptr(skip), ptr(dvec), d_stride, dil, L,
                                      ptr(Wsp_h), ptr(b_sp), ptr(Wfin_h), ptr(b_fin), ptr(out), B, T, C, ptr(flags), int(bf16),
                                      stream_ptr()), 'b2s_tc_wavenet_denoiser')


def tc_wavenet_stack3(xin_h, MF, Win_h, ld_win, b_in, Wd_h, cond_h, cond_layer_stride, Wres_h, bsum, dvec, d_stride, dilations,
                      yedge0_h, yedge1_h, z_all_h, z_layer_stride, B, T, C, flags, bf16, lens=None):
    L = len(dilations)
    dil = (_i * L)(*dilations)
    check(lib.b2s_tc_wavenet_stack3(ptr(xin_h), MF, ptr(Win_h), ld_win, ptr(b_in), ptr(Wd_h), ptr(cond_h), cond_layer_stride,
                                    ptr(Wres_h), ptr(bsum), ptr(dvec), d_stride, dil, L, ptr(yedge0_h), ptr(yedge1_h),
                                    ptr(z_all_h), z_layer_stride, B, T, C, ptr(flags), ptr(lens), int(bf16), stream_ptr()),
          'b2s_tc_wavenet_stack3')


def tc_wavenet_denoiser3(xin_h, MF, Win_h, ld_win, b_in, Wd_h, cond_h, cond_layer_stride, Wres_h, bsum, dvec, d_stride, dilations,
                         yedge0_h, yedge1_h, z_all_h, z_layer_stride, Wskip_h, bss, Wsp_h, b_sp, Wfin_h, b_fin, out, B, T, C, flags,
                         zflags, bf16, lens=None, chain=0):
    """``chain``: bit 0 - this launch follows another utterance group of the same evaluation, bit 1 - another group follows."""
    L = len(dilations)
    dil = (_i * L)(*dilations)
    check(lib.b2s_tc_wavenet_denoiser3_chained(ptr(xin_h), MF, ptr(Win_h), ld_win, ptr(b_in), ptr(Wd_h), ptr(cond_h), cond_layer_stride,
                                               ptr(Wres_h), ptr(bsum), ptr(dvec), d_stride, dil, L, ptr(yedge0_h), ptr(yedge1_h),
                                               ptr(z_all_h), z_layer_stride, ptr(Wskip_h), ptr(bss), ptr(Wsp_h), ptr(b_sp), ptr(Wfin_h),
                                               ptr(b_fin), ptr(out), B, T, C, ptr(flags), ptr(zflags), ptr(lens), int(bf16), int(chain),
                                               stream_ptr()),
          'b2s_tc_wavenet_denoiser3')


def tc_lynx_glu(h_h, W_h, bias, g_h, rows, C, inner, bf16):
    check(lib.b2s_tc_lynx_glu(ptr(h_h), ptr(W_h), ptr(bias), ptr(g_h), rows, C, inner, int(bf16), stream_ptr()),
          'b2s_tc_lynx_glu')


def tc_linear_residual(p_h, W_h, bias, x, rows, C, inner, bf16):
    check(lib.b2s_tc_linear_residual(ptr(p_h), ptr(W_h), ptr(bias), ptr(x), rows, C, inner, int(bf16), stream_ptr()),
          'b2s_tc_linear_residual')


def tc_linear_residual_cond(p_h, W_h, bias, x, cond_next_h, ld_cond, rows, C, inner, bf16):
    check(lib.b2s_tc_linear_residual_cond(ptr(p_h), ptr(W_h), ptr(bias), ptr(x), ptr(cond_next_h), ld_cond, rows, C, inner,
                                          int(bf16), stream_ptr()), 'b2s_tc_linear_residual_cond')


def lynx_prenorm_h(x, cond_h, ld_cond, dvec, d_stride, gamma, beta, h_h, B, T, C, strong, bf16):
    check(lib.b2s_lynx_prenorm_h(ptr(x), ptr(cond_h), ld_cond, ptr(dvec), d_stride, ptr(gamma), ptr(beta), ptr(h_h), B, T, C,
                                 int(strong), int(bf16), stream_ptr()), 'b2s_lynx_prenorm_h')


def layernorm_h(x, gamma, beta, h_h, rows, C, bf16):
    check(lib.b2s_layernorm_h(ptr(x), ptr(gamma), ptr(beta), ptr(h_h), rows, C, int(bf16), stream_ptr()), 'b2s_layernorm_h')


def tc_conv1d(a_h, W_h, bias, out_f32, ldo, out_h, ldoh, B, T, Cin, N, ksize, act, bf16):
    check(lib.b2s_tc_conv1d(ptr(a_h), ptr(W_h), ptr(bias), ptr(out_f32), ldo, ptr(out_h), ldoh, B, T, Cin, N, ksize, act, int(bf16),
                            stream_ptr()), 'b2s_tc_conv1d')


def layernorm_hh(in_h, gamma, beta, out_h, rows, C, eps, bf16):
    check(lib.b2s_layernorm_hh(ptr(in_h), ptr(gamma), ptr(beta), ptr(out_h), rows, C, float(eps), int(bf16), stream_ptr()),
          'b2s_layernorm_hh')


def tc_linear_residual_scaled(p_h, W_h, bias, gamma, x, x_h, rows, C, inner, bf16):
    check(lib.b2s_tc_linear_residual_scaled(ptr(p_h), ptr(W_h), ptr(bias), ptr(gamma), ptr(x), ptr(x_h), rows, C, inner, int(bf16),
                                            stream_ptr()), 'b2s_tc_linear_residual_scaled')


def enc_mel2ph_to_dur(mel2ph, dur, B, T, L):
    check(lib.b2s_enc_mel2ph_to_dur(ptr(mel2ph), ptr(dur), B, T, L, stream_ptr()), 'b2s_enc_mel2ph_to_dur')


def enc_embed(tokens, dur, E, w_dur, b_dur, x, keep, rows, H, vocab):
    check(lib.b2s_enc_embed(ptr(tokens), ptr(dur), ptr(E), ptr(w_dur), ptr(b_dur), ptr(x), ptr(keep), rows, H, vocab, stream_ptr()),
          'b2s_enc_embed')


def enc_rope(qkv, freqs, B, L, H, heads):
    check(lib.b2s_enc_rope(ptr(qkv), ptr(freqs), B, L, H, heads, stream_ptr()), 'b2s_enc_rope')


def enc_attention(qkv, keep, out_h, B, L, H, heads, bf16):
    check(lib.b2s_enc_attention(ptr(qkv), ptr(keep), ptr(out_h), B, L, H, heads, int(bf16), stream_ptr()), 'b2s_enc_attention')


def enc_mask_rows(x, keep, rows, H):
    check(lib.b2s_enc_mask_rows(ptr(x), ptr(keep), rows, H, stream_ptr()), 'b2s_enc_mask_rows')


def enc_layernorm_mask(x, gamma, beta, keep, enc, B, L, H, eps):
    check(lib.b2s_enc_layernorm_mask(ptr(x), ptr(gamma), ptr(beta), ptr(keep), ptr(enc), B, L, H, float(eps), stream_ptr()),
          'b2s_enc_layernorm_mask')


def enc_assemble(enc, mel2ph, spk, vals, ws, biases, n_var_first, n_var, cond, B, T, L, H):
    n = len(vals)
    mk = lambda ts: (_vp * n)(*[t.data_ptr() for t in ts])
    per_frame = int(spk is not None and spk.dim() == 3)
    check(lib.b2s_enc_assemble(ptr(enc), ptr(mel2ph), ptr(spk), per_frame, mk(vals), mk(ws), mk(biases), n, n_var_first, n_var, ptr(cond), B, T, L, H,
                               stream_ptr()), 'b2s_enc_assemble')


def lynx_dwconv_h(g_h, Wdw, bias, slope, p_h, B, T, inner, ksize, act, bf16):
    check(lib.b2s_lynx_dwconv_h(ptr(g_h), ptr(Wdw), ptr(bias), ptr(slope), ptr(p_h), B, T, inner, ksize, act, int(bf16),
                                stream_ptr()), 'b2s_lynx_dwconv_h')


# ---- NSF-HiFiGAN vocoder ---------------------------------------------------------------------------
def tc_conv1d_dil(a_h, W_h, bias, out_f32, ldo, out_h, ldoh, B, T, Cin, N, ksize, dil, act, bf16):
    check(lib.b2s_tc_conv1d_dil(ptr(a_h), ptr(W_h), ptr(bias), ptr(out_f32), ldo, ptr(out_h), ldoh, B, T, Cin, N, ksize, dil, act,
                                int(bf16), stream_ptr()), 'b2s_tc_conv1d_dil')


def tc_conv1d_residual(a_h, W_h, bias, x_src, x, y_h, y_slope, B, T, Cin, N, ksize, dil, bf16):
    check(lib.b2s_tc_conv1d_residual(ptr(a_h), ptr(W_h), ptr(bias), ptr(x_src), ptr(x), ptr(y_h), float(y_slope), B, T, Cin, N, ksize,
                                     dil, int(bf16), stream_ptr()), 'b2s_tc_conv1d_residual')


def voc_phase(f0, phase, B, T, sr, upp, mini_nsf):
    check(lib.b2s_voc_phase(ptr(f0), ptr(phase), B, T, float(sr), upp, int(mini_nsf), stream_ptr()), 'b2s_voc_phase')


def voc_source(f0, phase, rand_ini, noise, w, bias, out, B, T, upp, dim, sr, sine_amp, noise_std, thr):
    check(lib.b2s_voc_source(ptr(f0), ptr(phase), ptr(rand_ini), ptr(noise), ptr(w), ptr(bias), ptr(out), B, T, upp, dim, float(sr),
                             float(sine_amp), float(noise_std), float(thr), stream_ptr()), 'b2s_voc_source')


def voc_source_add(x, lx_h, src, Wt, bias, B, T, Cp, ksize, stride, pad, n_src, slope, bf16):
    check(lib.b2s_voc_source_add(ptr(x), ptr(lx_h), ptr(src), ptr(Wt), ptr(bias), B, T, Cp, ksize, stride, pad, n_src, float(slope),
                                 int(bf16), stream_ptr()), 'b2s_voc_source_add')


def voc_avg_act(xs, out_h, slope, bf16):
    arr = (_vp * len(xs))(*[t.data_ptr() for t in xs])
    check(lib.b2s_voc_avg_act(arr, len(xs), ptr(out_h), xs[0].numel(), float(slope), int(bf16), stream_ptr()), 'b2s_voc_avg_act')


def voc_post(xs, W, b0, wav, B, T, C, Cp, ksize, slope):
    arr = (_vp * len(xs))(*[t.data_ptr() for t in xs])
    check(lib.b2s_voc_post(arr, len(xs), ptr(W), ptr(b0), ptr(wav), B, T, C, Cp, ksize, float(slope), stream_ptr()), 'b2s_voc_post')


def cast_scale_h(inp, out, scale, bf16):
    check(lib.b2s_cast_scale_f32_h(ptr(inp), ptr(out), inp.numel(), float(scale), int(bf16), stream_ptr()), 'b2s_cast_scale_f32_h')


def masked_loss(a, b, mask, t_weights, l1, out):
    """out[0] = mean(w * loss(a * mask, b * mask)) over a, b [B, F, M, T]; mask [B, T, 1 or M] or None; t_weights [B] or None."""
    B, F, M, T = a.shape
    ws = torch.empty(lib.b2s_masked_loss_workspace_bytes() // 8, dtype=torch.float64, device=a.device)
    check(lib.b2s_masked_loss_f32(ptr(a), ptr(b), ptr(mask), 0 if mask is None else mask.shape[-1], ptr(t_weights), B, F, M, T, int(l1),
                                  ptr(ws), ptr(out), stream_ptr()), 'b2s_masked_loss_f32')

