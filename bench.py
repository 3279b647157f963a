#!/usr/bin/env python
"""bench.py - headline benchmark of the B200 sampling hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--precision fp32|bf16|fp16]

Workload (BASELINE.json configs[1], the configuration the metric is quoted on):
    shallow-diffusion DDPM, K_step = 400 full-step ancestral sampling (400 denoiser evaluations),
    WaveNet 20 layers x 256 channels, 128 mel bins, hidden 256, batch 16 utterances x 690 frames (8 s)
    PER GPU (weak scaling: utterances are independent units, no collective inside the loop; the only
    exchange is the final mel gather), synthetic condition / shallow source, random-init weights.

One "step" = one full sampling call over the batch.  Metric: denoised mel-frames x NFE per second.
    value : inputs resident in HBM, `model(condition_dev, src_spec=src_dev, infer=True)`
    e2e   : the same public call from PINNED HOST buffers, H2D + D2H inside the timed region
Also reported: roofline of the dominant kernel (CUDA events, the shipped launch at the bench shapes), the CPU baseline
(bounded sample on the host cores), the reference's eager PyTorch code on the SAME GPU (`gpu_eager_baseline`: the number
SURVEY.md section 8d calls "the one to beat"), short runs of the other BASELINE configs (`secondary`), SM clocks during
the timed region.

`--impl reference` times the reference on the host CPU on the same metric, one bounded sample per step: the UNMODIFIED
reference from baseline/_ref (baseline/install_reference.sh; `cpu_baseline.kind = "reference"`), or the oracle port when that
copy is absent (`kind = "port"`).  That arm imports nothing of the product.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (B per GPU, T, K_step, sampler)
    'config2': dict(B=16, T=690, k_step=400, layers=20, channels=256, mel=128, hidden=256, cycle=4,
                    desc='shallow DDPM K_step=400 full-step, WaveNet 20x256, 128 mel, B=16 x T=690 per GPU'),
    # secondary workloads (not the headline; `--workload NAME` for the tables in DESIGN.md)
    'config3': dict(kind='lynx_reflow', B=64, T=690, k_step=20, layers=6, channels=1024, mel=128, hidden=256, cycle=0,
                    desc='rectified-flow Euler 20 steps, LYNXNet 6x1024 (E=2, k=31, strong_cond), 128 mel, B=64 x T=690 per GPU'),
    'config5': dict(kind='wavenet_unipc', B=32, T=690, k_step=20, layers=20, channels=512, mel=128, hidden=256, cycle=4,
                    desc='UniPC 20 steps, WaveNet 20x512, 128 mel, B=32 x T=690 per GPU'),
    'config4': dict(kind='wavenet_variance', B=64, T=690, k_step=10, layers=10, channels=192, mel=48, out=2, hidden=256, cycle=4,
                    desc='variance multi-predictor (energy + breathiness, 2 x 24 repeat bins), DPM-Solver++ 10 steps, WaveNet 10x192, '
                         'B=64 x T=690 per GPU'),
    'config4_pitch': dict(kind='wavenet_pitch', B=64, T=690, k_step=10, layers=20, channels=256, mel=64, out=1, hidden=256, cycle=5,
                          desc='variance pitch predictor (PitchDiffusion, 64 repeat bins), UniPC 10 steps, WaveNet 20x256 (cycle 5), '
                               'B=64 x T=690 per GPU'),
    'config1': dict(kind='wavenet_ddim', B=1, T=690, k_step=20, layers=20, channels=256, mel=128, hidden=256, cycle=4,
                    desc='DDIM 20 steps (speedup 50), WaveNet 20x256, 128 mel, one 8-s utterance (690 frames)'),
}
SIGMA_W = 0.01


def flops_per_frame_nfe(L, C, MF, kind='wavenet'):
    """F_step(WaveNet) = 2*[L*8C^2 + MF*C + C^2 + C*MF];
    F_step(LYNXNet) = 2*[L*(C*2EC + EC*C + EC*k) + 2*MF*C], E = 2, k = 31."""
    if kind.startswith('lynx'):
        E, k = 2, 31
        return 2 * (L * (C * 2 * E * C + E * C * C + E * C * k) + 2 * MF * C)
    return 2 * (L * 8 * C * C + MF * C + C * C + C * MF)


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(bf16_burst=p['bf16_tflops'], bf16_sustained=p.get('bf16_tflops_sustained', p['bf16_tflops']),
                    hbm=p['hbm_gbs'], source='measured (MEASURED_PEAKS.json)')
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source='fallback (B200_PROFILING.md)')


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile('w', suffix='.csv', delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.idx}', f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except Exception:                                   # noqa: BLE001
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:                                   # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), parts[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['no samples'])
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


def make_model(w, precision, device, cuda_graph=True):
    import xiaoicesing_io_b200 as P
    kind = w.get('kind', 'wavenet_ddpm_shallow')
    P.hparams.clear()
    if kind == 'lynx_reflow':
        P.hparams.update(hidden_size=w['hidden'], use_shallow_diffusion=False, sampling_algorithm='euler',
                         sampling_steps=w['k_step'], infer=False, b2s_precision=precision, b2s_cuda_graph=cuda_graph)
        torch.manual_seed(0)
        model = P.RectifiedFlow(
            w['mel'], backbone_type='lynxnet',
            backbone_args=dict(num_layers=w['layers'], num_channels=w['channels'], kernel_size=31, strong_cond=True),
            spec_min=[-12.], spec_max=[0.])
        torch.nn.init.normal_(model.velocity_fn.output_projection.weight, std=SIGMA_W)
        return model.to(device).eval()
    if kind == 'wavenet_variance':
        # configs/variance.yaml:95-100: MultiVarianceDiffusion over 2 curves x 24 repeat bins, WaveNet 10 x 192
        P.hparams.update(hidden_size=w['hidden'], schedule_type='linear', use_shallow_diffusion=False,
                         diff_speedup=1000 // w['k_step'], diff_accelerator='dpm-solver', infer=False, b2s_precision=precision,
                         b2s_cuda_graph=cuda_graph)
        torch.manual_seed(0)
        model = P.MultiVarianceDiffusion(ranges=[(-96., -12.), (-96., -20.)], clamps=[(-96., -12.), (-96., -20.)], repeat_bins=24,
                                         backbone_type='wavenet',
                                         backbone_args=dict(num_layers=w['layers'], num_channels=w['channels'],
                                                            dilation_cycle_length=w['cycle']))
        torch.nn.init.normal_(model.denoise_fn.output_projection.weight, std=SIGMA_W)
        return model.to(device).eval()
    if kind == 'wavenet_pitch':
        # configs/variance.yaml:67-77: PitchDiffusion over 64 repeat bins, WaveNet 20 x 256, dilation cycle 5
        P.hparams.update(hidden_size=w['hidden'], schedule_type='linear', use_shallow_diffusion=False,
                         diff_speedup=1000 // w['k_step'], diff_accelerator='unipc', infer=False, b2s_precision=precision,
                         b2s_cuda_graph=cuda_graph)
        torch.manual_seed(0)
        model = P.PitchDiffusion(vmin=-8., vmax=8., cmin=-12., cmax=12., repeat_bins=w['mel'], backbone_type='wavenet',
                                 backbone_args=dict(num_layers=w['layers'], num_channels=w['channels'],
                                                    dilation_cycle_length=w['cycle']))
        torch.nn.init.normal_(model.denoise_fn.output_projection.weight, std=SIGMA_W)
        return model.to(device).eval()
    if kind in ('wavenet_unipc', 'wavenet_ddim'):
        P.hparams.update(hidden_size=w['hidden'], schedule_type='linear', use_shallow_diffusion=False,
                         diff_speedup=1000 // w['k_step'], diff_accelerator='unipc' if kind == 'wavenet_unipc' else 'ddim',
                         infer=False, b2s_precision=precision, b2s_cuda_graph=cuda_graph)
        torch.manual_seed(0)
        model = P.GaussianDiffusion(
            w['mel'], timesteps=1000, k_step=1000, backbone_type='wavenet',
            backbone_args=dict(num_layers=w['layers'], num_channels=w['channels'], dilation_cycle_length=w['cycle']),
            spec_min=[-12.], spec_max=[0.])
        torch.nn.init.normal_(model.denoise_fn.output_projection.weight, std=SIGMA_W)
        return model.to(device).eval()
    P.hparams.update(hidden_size=w['hidden'], schedule_type='linear', use_shallow_diffusion=True,
                     K_step_infer=w['k_step'], diff_speedup=1, diff_accelerator='ddim', infer=False,
                     b2s_precision=precision, b2s_cuda_graph=cuda_graph)
    torch.manual_seed(0)
    model = P.GaussianDiffusion(
        w['mel'], timesteps=1000, k_step=w['k_step'], backbone_type='wavenet',
        backbone_args=dict(num_layers=w['layers'], num_channels=w['channels'], dilation_cycle_length=w['cycle']),
        spec_min=[-12.], spec_max=[0.])
    # the reference zero-initialises this weight (wavenet.py:73); re-draw it or the output is constant
    torch.nn.init.normal_(model.denoise_fn.output_projection.weight, std=SIGMA_W)
    return model.to(device).eval()


def synth_inputs(w, seed):
    g = torch.Generator().manual_seed(seed)
    condition = torch.randn((w['B'], w['T'], w['hidden']), generator=g)
    src_spec = torch.rand((w['B'], w['T'], w['mel']), generator=g) * 12 - 12
    return condition, src_spec


# ------------------------------------------------------------------------------------------------------
# CPU arm: the reference (baseline/_ref) or, when it is absent, the oracle port - on the host cores.  No product imports.
# ------------------------------------------------------------------------------------------------------
def oracle_state_dict(w):
    """Random-init weights of the workload's WaveNet from the oracle's seeded recipe (oracle/weights.py), fp32 CPU."""
    from oracle import denoisers as OD
    from oracle import weights as OW
    cfg = OD.WaveNetCfg(in_dims=w['mel'], n_feats=1, num_layers=w['layers'], num_channels=w['channels'],
                        dilation_cycle_length=w['cycle'], hidden_size=w['hidden'])
    return cfg, OW.make_state_dict(cfg, seed=0, sigma_w=SIGMA_W)


def reference_model(w, device='cpu'):
    """The UNMODIFIED reference's GaussianDiffusion (modules/core/ddpm.py:55) with its own WaveNet, or None when the reference
    is not installed (baseline/_ref absent and no /root/reference)."""
    from oracle import ref_loader
    if not ref_loader.available():
        return None
    ref = ref_loader.load()
    ref.hparams.update(hidden_size=w['hidden'], schedule_type='linear', use_shallow_diffusion=True, K_step_infer=w['k_step'],
                       diff_speedup=1, diff_accelerator='ddim', infer=False)
    model = ref.ddpm.GaussianDiffusion(
        w['mel'], timesteps=1000, k_step=w['k_step'], backbone_type='wavenet',
        backbone_args=dict(num_layers=w['layers'], num_channels=w['channels'], dilation_cycle_length=w['cycle']),
        spec_min=[-12.], spec_max=[0.])
    _, sd = oracle_state_dict(w)
    model.denoise_fn.load_state_dict(sd, strict=True)
    return model.to(device).eval()


def ancestral_sample_throughput(w, n_utt, n_nfe, repeats, warmup, device='cpu', autocast=None):
    """frame*NFE/s of the reference's ancestral sampler on a bounded sample: ``n_utt`` utterances x T frames, the first ``n_nfe``
    steps of the K_step = 400 chain (ddpm.py:346-349: x = p_sample(x, t, cond)).  Runs the installed reference when there is
    one (kind 'reference'), else the oracle port (kind 'port').  Returns (throughput, times, kind)."""
    from oracle import denoisers as OD
    from oracle import samplers as OS
    g = torch.Generator().manual_seed(5)
    T = w['T']
    cond = torch.randn((n_utt, w['hidden'], T), generator=g).to(device)
    x0 = torch.randn((n_utt, 1, w['mel'], T), generator=g).to(device)
    model = reference_model(w, device)
    if model is not None:
        kind = 'reference'

        def one_pass():
            x = x0
            for j in range(n_nfe):
                i = w['k_step'] - 1 - j
                x = model.p_sample(x, torch.full((n_utt,), i, device=device, dtype=torch.long), cond)
            return x
    else:
        kind = 'port'
        cfg, sd = oracle_state_dict(w)
        sd = {k: v.to(device) for k, v in sd.items()}
        sch = OS.DiffusionSchedule(1000, 'linear')
        noises = [torch.randn((n_utt, 1, w['mel'], T), generator=g).to(device) for _ in range(n_nfe)]

        def one_pass():
            x = x0
            for j in range(n_nfe):
                i = w['k_step'] - 1 - j
                eps = OD.wavenet_forward(sd, cfg, x, torch.full((n_utt,), i, dtype=torch.long, device=device), cond)
                xr = float(sch.sqrt_recip_alphas_cumprod[i]) * x - float(sch.sqrt_recipm1_alphas_cumprod[i]) * eps
                mean = float(sch.posterior_mean_coef1[i]) * xr + float(sch.posterior_mean_coef2[i]) * x
                x = mean + math.exp(0.5 * float(sch.posterior_log_variance_clipped[i])) * noises[j]
            return x
    cuda = str(device).startswith('cuda')
    times = []
    with torch.no_grad():
        for it in range(warmup + repeats):
            if cuda:
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            if autocast is not None:
                with torch.autocast('cuda', dtype=autocast):
                    one_pass()
            else:
                one_pass()
            if cuda:
                torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    per = n_utt * T * n_nfe
    return per / statistics.median(times), times, kind


def run_reference_arm(args, w):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    n_utt, n_nfe = 4, 4
    t0 = time.perf_counter()
    thr, times, kind = ancestral_sample_throughput(w, n_utt, n_nfe, args.steps, args.warmup)
    sample = f'{n_utt} utterances x {w["T"]} frames x first {n_nfe} of {w["k_step"]} ancestral steps per step'
    line = {
        'impl': 'reference', 'metric': 'denoised mel-frames x NFE per second', 'value': thr, 'unit': 'frame*NFE/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * statistics.median(times), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': w['desc'], 'sample': sample},
        'cpu_baseline': {'value': thr, 'unit': 'frame*NFE/s', 'cores': torch.get_num_threads(), 'kind': kind,
                         'sample': sample},
        'e2e': {'value': thr, 'unit': 'frame*NFE/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'wall_s': time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def dominant_kernel_roofline(model, w, precision, reps=5):
    """Times the dominant kernel of one denoiser evaluation in isolation, at the bench shapes, with CUDA
    events on the launching stream: all L per-layer launches back to back, ``reps`` times."""
    from xiaoicesing_io_b200 import _cabi as C
    eng = getattr(model, model.backbone_attr)._engine()
    eng.pack()
    dev = eng.device
    B, T, Cc, L = w['B'], w['T'], w['channels'], w['layers']
    cond = torch.randn((B, T, w['hidden']), device=dev)
    tvals = torch.tensor([float(w['k_step'] - 1)], device=dev)
    sess = eng.begin(cond, tvals)
    if not hasattr(sess, 'dominant_kernel') and not (precision == 'fp32' and hasattr(sess, 'y')):
        return None                                   # no per-kernel roofline for this backbone / precision
    x_in = torch.randn((B * T, w['mel']), device=dev)
    out = torch.empty_like(x_in)
    sess.eval(x_in, 0, out)                       # fills y / z with realistic values
    rows = B * T
    p = peaks()
    if precision == 'fp32':
        name = 'sgemm_fused_kernel<EPI_GATE,CONV> (b2s_wavenet_gate_f32)'
        flops = 2.0 * rows * (3 * Cc) * (2 * Cc)
        n_launch = L

        def launch_all():
            for l in range(L):
                C.wavenet_gate(sess.y, eng.w_dil[l], sess.cond[:, l * 2 * Cc:], L * 2 * Cc, sess.z, B, T, Cc,
                               eng.dilations[l])
    else:
        name, flops, launch_all, n_launch = sess.dominant_kernel(w)
    # a launch from Python costs ~15 us of host time, more than the kernel itself: time the launches as a CUDA graph
    for _ in range(2):
        launch_all()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(reps):
            launch_all()
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (reps * n_launch)
    # DRAM traffic per launch: read from THIS round's `ncu --set full` capture of the same kernel at the same shapes when one is
    # committed (profiles/r02_ncu_full_<kernel>.csv, `ncu --page raw --csv`), else null - never a constant from an older kernel
    traffic = None
    kern = name.split('<')[0].split(' ')[0]
    prof = os.path.join(ROOT, 'profiles', f'r02_ncu_full_{kern}.csv')
    if os.path.exists(prof) and (B, T, Cc, L) == (16, 690, 256, 20):
        try:
            import csv
            rws = list(csv.reader(open(prof)))
            scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
            hdr = next(i for i, r in enumerate(rws) if 'dram__bytes_read.sum' in r)
            cols = [rws[hdr].index(k) for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum')]
            traffic = sum(float(rws[hdr + 2][c].replace(',', '')) * scale[rws[hdr + 1][c]] for c in cols)
        except Exception:                                   # noqa: BLE001
            traffic = None
    achieved = flops / (ms * 1e-3) / 1e12
    return {'bound': 'tensor', 'kernel': name, 'achieved': achieved, 'peak': p['bf16_burst'], 'unit': 'TFLOP/s',
            'frac': achieved / p['bf16_burst'], 'frac_of_sustained_peak': achieved / p['bf16_sustained'], 'traffic': traffic,
            'avg_launch_ms': ms,
            'flops_per_launch': flops, 'peak_source': p['source'], 'launches_per_eval': sess.launches_per_eval,
            'note': 'algorithmic FLOPs of the launch / avg CUDA-event duration of the launches replayed back to back as a CUDA graph'}


def time_workload(model, w, dev, steps, warmup):
    """(frame*NFE/s, ms per sampling call, NFE per call) of one workload, inputs resident, CUDA events."""
    cond_h, src_h = synth_inputs(w, seed=1000)
    cond_d, src_d = cond_h.to(dev), src_h.to(dev)
    kind = w.get('kind', 'wavenet_ddpm_shallow')
    if kind != 'wavenet_ddpm_shallow':
        src_d = None
    variance = kind in ('wavenet_variance', 'wavenet_pitch')
    run = (lambda: model(cond_d, infer=True)) if variance else (lambda: model(cond_d, src_spec=src_d, infer=True))
    for _ in range(warmup):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    nfe = model.build_program().n_nfe
    return w['B'] * w['T'] * nfe / (ms * 1e-3), ms, nfe


def run_secondary(args, dev):
    """Short runs (3 warm-up + 3 timed sampling calls each) of the BASELINE configs the headline is NOT quoted on, so that they are
    driver-run numbers too; each with the roofline of its own dominant kernel."""
    import xiaoicesing_io_b200 as P
    out = {}
    saved = dict(P.hparams)
    for name in ('config1', 'config3', 'config4_pitch', 'config4', 'config5'):
        w = dict(WORKLOADS[name])
        try:
            model = make_model(w, args.precision, dev)
            value, ms, nfe = time_workload(model, w, dev, steps=3, warmup=3)
            kind = w.get('kind', 'wavenet')
            F = flops_per_frame_nfe(w['layers'], w['channels'], w['mel'] if kind != 'wavenet_variance' else w['mel'], kind)
            roof = dominant_kernel_roofline(model, w, args.precision, reps=3)
            p = peaks()
            out[name] = {'workload': w['desc'], 'value': value, 'unit': 'frame*NFE/s', 'ms_per_call': ms, 'nfe_per_call': nfe,
                         'rtf': (ms * 1e-3) / (w['B'] * w['T'] * 512 / 44100.0),
                         'tflops_algorithmic': value * F / 1e12, 'frac_of_bf16_peak': value * F / 1e12 / p['bf16_burst'],
                         'roofline': None if roof is None else {k: roof[k] for k in ('kernel', 'achieved', 'peak', 'frac', 'avg_launch_ms')}}
            del model
        except Exception as ex:                             # noqa: BLE001
            out[name] = {'workload': w['desc'], 'error': f'{type(ex).__name__}: {ex}'}
        torch.cuda.empty_cache()
    try:
        out['aux_decoder'] = time_aux_decoder(args.precision, dev)
    except Exception as ex:                                 # noqa: BLE001
        out['aux_decoder'] = {'error': f'{type(ex).__name__}: {ex}'}
    try:
        out['tokens_to_mel_one_utterance'] = time_tokens_to_mel(args.precision, dev)
    except Exception as ex:                                 # noqa: BLE001
        out['tokens_to_mel_one_utterance'] = {'error': f'{type(ex).__name__}: {ex}'}
    torch.cuda.empty_cache()
    try:
        out['acoustic_encoder'] = time_acoustic_encoder(args.precision, dev)
    except Exception as ex:                                 # noqa: BLE001
        out['acoustic_encoder'] = {'error': f'{type(ex).__name__}: {ex}'}
    torch.cuda.empty_cache()
    try:
        out['vocoder'] = time_vocoder(args.precision, dev)
    except Exception as ex:                                 # noqa: BLE001
        out['vocoder'] = {'error': f'{type(ex).__name__}: {ex}'}
    torch.cuda.empty_cache()
    P.hparams.clear()
    P.hparams.update(saved)
    return out


def time_tokens_to_mel(precision, dev, L=64, T=690, reps=10, extra_hparams=None):
    """The reference's real inference shape end to end (modules/toplevel.py:79-102, inference/ds_acoustic.py:189-246): ONE 8-second
    utterance, phoneme tokens -> FastSpeech2 encoder -> ConvNeXt aux decoder -> shallow DDIM sampling (K_step 400, speedup 20 = 20
    evaluations of WaveNet 20x256) -> mel, every stage on the B200 kernels (xiaoicesing_io_b200.DiffSingerAcoustic).  Wall clock per
    call with a host synchronisation, CUDA graphs replayed."""
    import xiaoicesing_io_b200 as P
    P.hparams.clear()
    P.hparams.update(hidden_size=256, enc_layers=4, enc_ffn_kernel_size=3, ffn_act='gelu', num_heads=2, use_pos_embed=True, rel_pos=True,
                     use_rope=True, dropout=0.1, use_spk_id=False, num_spk=1, schedule_type='linear', infer=False,
                     use_shallow_diffusion=True, K_step_infer=400, diff_speedup=20, diff_accelerator='ddim', timesteps=1000, K_step=400,
                     spec_min=[-12.0] * 128, spec_max=[0.0] * 128, diffusion_type='ddpm', backbone_type='wavenet',
                     backbone_args=dict(num_layers=20, num_channels=256, dilation_cycle_length=4),
                     shallow_diffusion_args=dict(train_aux_decoder=True, train_diffusion=True, val_gt_start=False, aux_decoder_grad=0.1,
                                                 aux_decoder_arch='convnext',
                                                 aux_decoder_args=dict(num_channels=512, num_layers=6, kernel_size=7)),
                     b2s_precision=precision if precision != 'fp32' else 'fp16')
    P.hparams.update(extra_hparams or {})
    torch.manual_seed(0)
    m = P.DiffSingerAcoustic(60, 128)
    with torch.no_grad():
        torch.nn.init.normal_(m.diffusion.denoise_fn.output_projection.weight, std=0.01)
        for blk in m.aux_decoder.decoder.conv:
            blk.gamma.fill_(0.5)
    m = m.to(dev).eval()
    g = torch.Generator().manual_seed(1)
    tokens = torch.randint(1, 60, (1, L), generator=g).to(dev)
    mel2ph = (torch.arange(T)[None, :] * L // T + 1).contiguous().to(dev)
    f0 = (100 + 300 * torch.rand((1, T), generator=g)).to(dev)
    for _ in range(4):
        m(tokens, mel2ph, f0, infer=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        mel = m(tokens, mel2ph, f0, infer=True).diff_out
        torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / reps
    res = {'workload': f'one utterance ({L} tokens, {T} frames = 8 s): encoder 4x256 -> aux decoder 6x512 -> shallow DDIM 20 evaluations of '
                       f'WaveNet 20x256 -> mel', 'ms_per_call': ms, 'rtf': ms * 1e-3 / (T * 512 / 44100.0),
           'finite': bool(torch.isfinite(mel).all())}
    try:                                                     # ... and on to the waveform (NSF-HiFiGAN, public 44.1 kHz geometry)
        P.hparams['mel_base'] = 'e'
        voc = P.NsfHifiGAN(P.vocoder.Generator(dict(VOCODER_H)).to(dev).eval())
        for _ in range(4):
            voc.spec2wav_torch(m(tokens, mel2ph, f0, infer=True).diff_out, f0=f0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            wav = voc.spec2wav_torch(m(tokens, mel2ph, f0, infer=True).diff_out, f0=f0)
            torch.cuda.synchronize()
        ms_w = (time.perf_counter() - t0) * 1e3 / reps
        res['tokens_to_waveform'] = {'ms_per_call': ms_w, 'rtf': ms_w * 1e-3 / (T * 512 / 44100.0), 'samples': int(wav.numel()),
                                     'finite': bool(torch.isfinite(wav).all())}
    except Exception as ex:                                  # noqa: BLE001
        res['tokens_to_waveform'] = {'error': f'{type(ex).__name__}: {ex}'}
    return res


def time_acoustic_encoder(precision, dev, B=16, L=64, T=690, reps=5, extra_hparams=None):
    """The other step BEFORE the path: the FastSpeech2 acoustic encoder that produces the condition
    tensor (configs/acoustic.yaml + base.yaml: H 256, 4 layers, 2 heads, 3-tap conv FFN, rotary positions), once per utterance
    batch - config 2's batch with 64 phoneme tokens per utterance.  Random-init weights; device-timed with CUDA events."""
    import xiaoicesing_io_b200 as P
    P.hparams.clear()
    P.hparams.update(hidden_size=256, enc_layers=4, enc_ffn_kernel_size=3, ffn_act='gelu', num_heads=2, use_pos_embed=True, rel_pos=True,
                     use_rope=True, dropout=0.1, use_spk_id=False, num_spk=1, b2s_precision=precision if precision != 'fp32' else 'fp16')
    P.hparams.update(extra_hparams or {})
    torch.manual_seed(0)
    m = P.FastSpeech2Acoustic(60).to(dev).eval()
    g = torch.Generator().manual_seed(1)
    tokens = torch.randint(1, 60, (B, L), generator=g).to(dev)
    mel2ph = (torch.arange(T)[None, :] * L // T + 1).expand(B, T).contiguous().to(dev)
    f0 = (100 + 300 * torch.rand((B, T), generator=g)).to(dev)
    for _ in range(3):
        m(tokens, mel2ph, f0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        m(tokens, mel2ph, f0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return {'workload': f'FastSpeech2 acoustic encoder 4x256 (2 heads, rotary, conv-3 FFN), B={B} x {L} tokens -> {T} frames: the '
                        f'condition tensor, once per batch', 'ms_per_call': ms, 'frames_per_s': B * T / (ms * 1e-3),
            'launches_per_call': 2 + 10 * 4 + 2}


VOCODER_H = dict(num_mels=128, sampling_rate=44100, upsample_rates=[8, 8, 2, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4, 4],
                 upsample_initial_channel=512, resblock='1', resblock_kernel_sizes=[3, 7, 11],
                 resblock_dilation_sizes=[[1, 3, 5]] * 3)


def time_vocoder(precision, dev, T=690, reps=3, extra_hparams=None):
    """The step AFTER the path: NSF-HiFiGAN at the public 44.1 kHz geometry (512 channels, hop 512,
    residual kernels 3 / 7 / 11), mel + f0 -> waveform for one 8-s utterance (the reference's shape, inference/ds_acoustic.py:227-236)
    and for a batch of 8.  Fan-in scaled random weights; device-timed with CUDA events; next to it the unmodified reference's Generator
    in eager PyTorch fp32 on the same GPU when baseline/_ref is installed."""
    import xiaoicesing_io_b200 as P
    P.hparams.clear()
    P.hparams.update(b2s_precision=precision if precision != 'fp32' else 'fp16')
    P.hparams.update(extra_hparams or {})
    torch.manual_seed(0)
    gen = P.vocoder.Generator(dict(VOCODER_H))
    with torch.no_grad():
        # fan-in scaled weights that keep the activations O(1) through the five stages (the reference's own N(0, 0.01) init would make
        # every residual block a near-identity): same recipe as the parity tests' weights
        rate = {f'ups.{i}.': u for i, u in enumerate(VOCODER_H['upsample_rates'])}
        for n, p in gen.named_parameters():
            if n.endswith('weight') and p.dim() == 3:
                up = next((u for k, u in rate.items() if n.startswith(k)), None)
                fan_in = p.shape[1] * p.shape[2] if up is None else p.shape[0] * p.shape[2] / up
                p.normal_(0.0, (0.5 if '.convs2.' in n else 1.0) / math.sqrt(fan_in))
            elif n.endswith('bias'):
                p.normal_(0.0, 0.05)
    gen = gen.to(dev).eval()
    sd = {k: v.detach().clone() for k, v in gen.state_dict().items()}
    ref_gen = None
    try:
        from oracle import ref_loader                        # the unmodified reference (baseline/_ref), eager PyTorch on this GPU
        if ref_loader.available():
            voc = ref_loader.load_vocoder()
            ref_gen = voc.models.Generator(voc.AttrDict(dict(VOCODER_H))).eval()
            import contextlib, io
            with contextlib.redirect_stdout(io.StringIO()):
                ref_gen.remove_weight_norm()
            ref_gen.load_state_dict(sd, strict=True)
            ref_gen = ref_gen.to(dev)
    except Exception:                                        # noqa: BLE001
        ref_gen = None
    hop = 512
    # dense-conv MACs per mel frame: conv_pre, transposed convs (their non-zero taps), residual blocks
    macs, ch, rate = 7 * 128 * 512, 512, 1
    for u, k in zip(VOCODER_H['upsample_rates'], VOCODER_H['upsample_kernel_sizes']):
        macs += rate * k * ch * (ch // 2)
        ch //= 2
        rate *= u
        macs += rate * ch * ch * sum(6 * kk for kk in VOCODER_H['resblock_kernel_sizes'])
    # algorithmic HBM bytes per mel frame: per stage element (rate x channels) the transposed conv writes the fp32 stream (4 B), the source
    # conv updates it and writes the 16-bit copy (8 + 2), every conv pair reads / writes 16-bit tiles around the fp32 read-modify-write
    # (2 + 2, then 2 + 4 + 4 + 2; no 16-bit copy after a block's last pair), the block mean reads 3 streams and writes 16 bits (12 + 2)
    elems, ch, rate = 0, 512, 1
    for u in VOCODER_H['upsample_rates']:
        ch //= 2
        rate *= u
        elems += rate * ch
    bytes_per_frame = elems * (14 + 3 * (16 + 16 + 14) + 14)
    pk = peaks()
    res = {'workload': f'NSF-HiFiGAN 44.1 kHz geometry (512 ch, hop 512, resblock kernels 3/7/11), mel [B, {T}, 128] + f0 -> {T * hop} samples '
                       f'per utterance, once per utterance', 'gflop_per_utterance': 2.0 * macs * T / 1e9,
           'hbm_gb_algorithmic_per_utterance': bytes_per_frame * T / 1e9}

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    for B in (1, 8):
        mel = torch.randn((B, T, 128), device=dev) * 1.5 - 4.0
        f0 = 110.0 * 2 ** (2 * torch.rand((B, T), device=dev))
        ri, nz = torch.rand(1, 1, 9, device=dev), torch.randn(B, T * hop, 9, device=dev)
        ms = timed(lambda: gen.forward_rows(mel, f0, rand_ini=ri, noise=nz))
        r = {'ms_per_call': ms, 'samples_per_s': B * T * hop / (ms * 1e-3), 'rtf': (ms * 1e-3) / (B * T * hop / 44100.0),
             'tflops_algorithmic': 2.0 * macs * T * B / (ms * 1e-3) / 1e12}
        r['frac_of_bf16_peak'] = r['tflops_algorithmic'] / pk['bf16_burst']
        r['hbm_gbs_algorithmic'] = bytes_per_frame * T * B / (ms * 1e-3) / 1e9          # B = 1 lives in the 126 MB L2: not an HBM figure there
        r['frac_of_hbm_peak'] = r['hbm_gbs_algorithmic'] / pk['hbm']
        if ref_gen is not None:
            mel_c = mel.transpose(1, 2).contiguous()
            with torch.no_grad():
                ms_ref = timed(lambda: ref_gen(mel_c, f0))    # torch defaults: cuDNN convolutions may use TF32
                tf32 = torch.backends.cudnn.allow_tf32
                torch.backends.cudnn.allow_tf32 = False        # the parity figure is against true fp32 arithmetic
                try:
                    torch.manual_seed(1)
                    a = ref_gen(mel_c, f0)
                finally:
                    torch.backends.cudnn.allow_tf32 = tf32
                torch.manual_seed(1)
                b = gen(mel_c, f0)                            # same draws from the device generator, in the reference's order
            r.update(eager_reference_ms=ms_ref, speedup_vs_eager=ms_ref / ms, max_abs_vs_eager_reference_fp32=float((a - b).abs().max()))
        res[f'B{B}'] = r
    return res


def time_aux_decoder(precision, dev, B=16, T=690, reps=5, extra_hparams=None):
    """The step BEFORE the path: the ConvNeXt aux decoder that produces x_start for shallow diffusion
    (configs/acoustic.yaml:101-105: 512 channels, 6 blocks, k = 7), once per utterance batch - config 2's batch.  Random-init
    weights with layer scale 0.5; device-timed with CUDA events."""
    import xiaoicesing_io_b200 as P
    P.hparams.clear()
    P.hparams.update(b2s_precision=precision if precision != 'fp32' else 'fp16')
    P.hparams.update(extra_hparams or {})
    torch.manual_seed(0)
    m = P.AuxDecoderAdaptor(in_dims=256, out_dims=128, num_feats=1, spec_min=[-12.0] * 128, spec_max=[0.0] * 128,
                            aux_decoder_arch='convnext', aux_decoder_args=dict(num_channels=512, num_layers=6, kernel_size=7))
    with torch.no_grad():
        for blk in m.decoder.conv:
            blk.gamma.fill_(0.5)
    m = m.to(dev).eval()
    cond = torch.randn((B, T, 256), device=dev)
    for _ in range(3):
        m(cond, infer=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        m(cond, infer=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    macs = 7 * 256 * 512 + 6 * (2 * 512 * 2048 + 7 * 512) + 7 * 512 * 128        # per frame: inconv, blocks (pointwise + depthwise), outconv
    tf = 2.0 * macs * B * T / (ms * 1e-3) / 1e12
    return {'workload': f'ConvNeXt aux decoder 6x512 (k=7), 256 -> 128 mel, B={B} x T={T}: x_start for shallow diffusion, once per batch',
            'ms_per_call': ms, 'frames_per_s': B * T / (ms * 1e-3), 'tflops_algorithmic': tf,
            'frac_of_bf16_peak': tf / peaks()['bf16_burst'], 'launches_per_call': 2 + 4 * 6 + 1}


def run_b200_arm(args, w):
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device - the B200 arm has no CPU fallback (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    import xiaoicesing_io_b200 as P
    from xiaoicesing_io_b200.partition import MelGather

    model = make_model(w, args.precision, dev, cuda_graph=not args.no_graph)
    if args.fuse_io is not None:
        P.hparams['b2s_fuse_io'] = bool(args.fuse_io)
    if args.overlap_noise is not None:
        P.hparams['b2s_overlap_noise'] = bool(args.overlap_noise)
    for kv in args.hparam or []:
        k, v = kv.split('=', 1)
        P.hparams[k] = json.loads(v)
    cond_h, src_h = synth_inputs(w, seed=1000 + rank)          # every rank owns its utterances (weak scaling)
    cond_h, src_h = cond_h.pin_memory(), src_h.pin_memory()
    cond_d, src_d = cond_h.to(dev), src_h.to(dev)
    B, T = w['B'], w['T']
    prog = model.build_program()
    nfe = prog.n_nfe
    parts = [list(range(r * B, (r + 1) * B)) for r in range(world)]     # weak scaling: every rank owns B utterances

    shallow = w.get('kind', 'wavenet_ddpm_shallow') == 'wavenet_ddpm_shallow'
    if not shallow:
        src_d = None

    variance = w.get('kind') in ('wavenet_variance', 'wavenet_pitch')
    n_out = w.get('out', w['mel'])
    def run(c, s):
        if not variance:
            return model(c, src_spec=s, infer=True)
        r = model(c, infer=True)
        return torch.stack(r, -1) if isinstance(r, (list, tuple)) else r.unsqueeze(-1)

    # the ONLY exchange: one gather of the finished mels to rank 0 into pre-allocated buffers (prepared once: no counts, no
    # indices, no host synchronisation per step), then ONE device-to-host copy into a pinned buffer
    out_shape = (T, n_out)
    gather = MelGather(parts, out_shape, dev, dst=0) if world > 1 else None
    host_out = torch.empty((world * B,) + out_shape, dtype=torch.float32).pin_memory() if rank == 0 else None

    def step_resident():
        mel = run(cond_d, src_d)
        if gather is not None:
            gather(mel.contiguous())
        return mel

    def step_e2e():
        c = cond_h.to(dev, non_blocking=True)
        s = src_h.to(dev, non_blocking=True) if shallow else None
        mel = run(c, s)
        full = gather(mel.contiguous()) if gather is not None else mel
        if full is not None and host_out is not None:
            host_out.copy_(full.reshape(host_out.shape), non_blocking=True)     # completes before the timed region's final sync
        return full

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sampler=None):
        for _ in range(warmup):
            fn()
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        clocks = sampler.stop() if sampler else None
        ms = max(e0.elapsed_time(e1), 0.0)
        ms = max(ms, 0.0)
        t = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), clocks

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_total, wall_ms, clocks = timed(step_resident, args.steps, args.warmup, sampler)
    e2e_ms, e2e_wall_ms, _ = timed(step_e2e, args.steps, max(1, args.warmup // 3))
    units = world * B * T * nfe * args.steps
    value = units / (ms_total * 1e-3)
    e2e_value = units / (max(e2e_ms, e2e_wall_ms) * 1e-3)      # host copies: wall clock bounds it from above

    roof = cpu = eager = secondary = None
    if rank == 0:
        roof = dominant_kernel_roofline(model, w, args.precision)
        if world == 1 and not args.no_cpu_baseline and shallow:
            torch.set_num_threads(os.cpu_count() or 1)
            n_utt, n_nfe = 4, 4
            thr, _, kind = ancestral_sample_throughput(w, n_utt, n_nfe, repeats=3, warmup=1)
            cpu = {'value': thr, 'unit': 'frame*NFE/s', 'cores': torch.get_num_threads(), 'kind': kind,
                   'sample': f'{n_utt} utterances x {T} frames x first {n_nfe} of {w["k_step"]} ancestral steps, median of 3'}
            # the reference's eager PyTorch code on THIS GPU, full batch: the number SURVEY.md section 8d calls "the one to beat"
            n_nfe_g = 8
            e32, _, kind_g = ancestral_sample_throughput(w, B, n_nfe_g, repeats=3, warmup=2, device=dev)
            e16, _, _ = ancestral_sample_throughput(w, B, n_nfe_g, repeats=3, warmup=2, device=dev, autocast=torch.bfloat16)
            eager = {'fp32': e32, 'autocast_bf16': e16, 'unit': 'frame*NFE/s', 'kind': kind_g,
                     'sample': f'{B} utterances x {T} frames x first {n_nfe_g} of {w["k_step"]} ancestral steps, torch eager on cuda '
                               f'(cudnn/cublas defaults), median of 3',
                     'speedup_over_fp32': value / e32, 'speedup_over_autocast_bf16': value / e16}
        if world == 1 and not args.no_secondary and args.workload == 'config2' and not (args.k_step or args.batch or args.frames):
            secondary = run_secondary(args, dev)
    if rank == 0:
        sess_launches = (roof or {}).get('launches_per_eval', (1 if args.precision == 'fp32' else 2) + 2 * w['layers'] + 2)
        n_lin = sum(1 for op in prog.ops if op.kind == 'lin')
        n_noise = prog.n_draws
        # 16-bit path: an update that feeds the next evaluation writes the 16-bit input itself (no cast launch there)
        n_precast = 0 if (args.precision == 'fp32' or not P.hparams.get('b2s_fuse_cast', True)) else sum(
            1 for i, op in enumerate(prog.ops)
            if op.kind == 'nfe' and i > 0 and prog.ops[i - 1].kind == 'lin' and prog.ops[i - 1].dst == op.src)
        launches_per_step = nfe * sess_launches - n_precast + n_lin + n_noise + 2 + 4   # + start transposes + tables
        if not args.no_graph:
            from xiaoicesing_io_b200.core import _sampling as _S
            counted = [g.n_launches for g in _S._GRAPH_CACHE.values() if hasattr(g, 'n_launches')]
            if counted:
                launches_per_step = counted[-1]          # counted while the replayed graph was captured
        F = flops_per_frame_nfe(w['layers'], w['channels'], w['mel'], w.get('kind', 'wavenet'))
        p = peaks()
        line = {
            'metric': 'denoised mel-frames x NFE per second', 'value': value, 'unit': 'frame*NFE/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_total / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': {'fp32': 'f32', 'bf16': 'bf16', 'fp16': 'f16'}[args.precision], 'data': 'synthetic',
            'config': {'workload': w['desc'], 'nfe_per_step': nfe, 'global_batch': world * B, 'frames': T,
                       'parallelism': f'utterance-partition x{world}', 'precision': args.precision, 'cuda_graph': not args.no_graph,
                       'l2': ('hoisted cond table %.0f MB + weights/activations per GPU vs the 126 MB L2; no explicit flush '
                              '(the table is streamed once per denoiser evaluation)'
                              % (B * T * w['layers'] * (1 if w.get('kind', '').startswith('lynx') else 2) * w['channels']
                                 * (4 if args.precision == 'fp32' else 2) / 1e6)),
                       'sigma_w': SIGMA_W},
            'e2e': {'value': e2e_value, 'unit': 'frame*NFE/s', 'ms_per_step': max(e2e_ms, e2e_wall_ms) / args.steps,
                    'h2d_bytes_per_step': int(cond_h.numel() * 4 + (src_h.numel() * 4 if shallow else 0)) * world,
                    'd2h_bytes_per_step': int(world * B * T * n_out * 4)},
            'gpu_launches': int(launches_per_step * args.steps),
            'rtf': (ms_total / args.steps * 1e-3) / (world * B * T * 512 / 44100.0),
            'tflops_algorithmic': value * F / 1e12,
            'frac_of_bf16_peak': value * F / 1e12 / (world * p['bf16_sustained']),
            'roofline': roof, 'cpu_baseline': cpu, 'gpu_eager_baseline': eager, 'secondary': secondary, 'clocks': clocks,
            'wall_ms_per_step': wall_ms / args.steps,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _ensure_built():
    """libb2s.so is git-ignored: build it when it is missing (local rank 0 builds, the other ranks wait for the file)."""
    lib = os.path.join(ROOT, 'xiaoicesing_io_b200', 'libb2s.so')
    if os.path.exists(lib):
        return
    if int(os.environ.get('LOCAL_RANK', '0')) == 0:
        import __graft_entry__ as entry
        entry._build_module().build()
    else:
        t0 = time.time()
        while not os.path.exists(lib) and time.time() - t0 < 900:
            time.sleep(1.0)
        time.sleep(2.0)


def main():
    _ensure_built()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--precision', default=os.environ.get('B2S_PRECISION', 'fp16'), choices=['fp32', 'bf16', 'fp16'],
                    help='fp16 (default): the 16-bit tensor-core path that meets the 2e-2 parity bound on every config')
    ap.add_argument('--workload', default='config2', choices=sorted(WORKLOADS))
    ap.add_argument('--k-step', type=int, default=None, help='override K_step (debug only; invalidates the metric)')
    ap.add_argument('--batch', type=int, default=None, help='override utterances per GPU (sweeps only; not the headline config)')
    ap.add_argument('--frames', type=int, default=None, help='override frames per utterance (sweeps only)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-secondary', action='store_true', help='skip the short runs of the other BASELINE configs')
    ap.add_argument('--fuse-io', type=int, default=None, help='override hparams b2s_fuse_io (A/B switch)')
    ap.add_argument('--overlap-noise', type=int, default=None, help='override hparams b2s_overlap_noise (A/B switch)')
    ap.add_argument('--hparam', action='append', help='extra hparams entry key=json (A/B switches, e.g. b2s_fuse_cast=false)')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel from the host instead of replaying the captured CUDA graph')
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.k_step:
        w['k_step'] = args.k_step
        w['desc'] += f' [DEBUG K_step={args.k_step}]'
    if args.batch or args.frames:
        w['B'] = args.batch or w['B']
        w['T'] = args.frames or w['T']
        w['desc'] += f' [SWEEP OVERRIDE B={w["B"]} T={w["T"]}]'
    if args.impl == 'reference':
        run_reference_arm(args, w)
    else:
        run_b200_arm(args, w)


if __name__ == '__main__':
    main()
