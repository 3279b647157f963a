"""The process-global config dict the hot path reads (reference: utils/hparams.py:13).

Inside the reference tree the reference's own dict object is re-used, so ``set_hparams()`` there
configures this package too - that *is* the drop-in contract.  Stand-alone, a
local dict with the same role is used.  Keys read on the path: hidden_size, schedule_type,
use_shallow_diffusion, K_step_infer, diff_speedup, diff_accelerator, T_start_infer,
sampling_algorithm, sampling_steps, infer; plus this package's own ``b2s_precision``
('fp32' | 'fp16' | 'bf16'; unset: ``$B2S_PRECISION``, else 'fp32' - true fp32 FFMA kernels, parity <= 1e-3; 'fp16' is the
16-bit tensor-core path bench.py measures, parity <= 2e-2 on every config; 'bf16' is the same path with bf16 operands for
checkpoints whose activations leave the fp16 range) and ``b2s_cuda_graph`` (bool, default True).
"""
try:                                    # pragma: no cover - only inside the reference tree
    from utils.hparams import hparams   # type: ignore
except Exception:                       # noqa: BLE001
    hparams = {}


def set_hparams(**kw):
    hparams.update(kw)
    return hparams
