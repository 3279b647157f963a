"""Oracle (TEST INFRASTRUCTURE): CPU restatement of the sampling loops.

Every sampler takes the denoiser as a closure ``denoise(x[B,F,M,T], t[B or 1], cond)``
and takes ALL random draws as explicit tensors (``noise0`` = the initial draw of
``inference()``, ``step_noise[i]`` = the i-th per-step draw of DDPM ancestral
sampling), because CPU and CUDA generators differ (SURVEY.md section 8a gotcha 5).
Solver scalars are computed with fp32 ``torch`` ops exactly where the reference
does so; pass ``dtype=torch.float64`` for the truth run.
"""
from __future__ import annotations

import math
from collections import deque

import numpy as np
import torch


# --------------------------------------------------------------------------
# schedule buffers  (ddpm.py:28-52, 64-101)
# --------------------------------------------------------------------------
def linear_beta_schedule(timesteps, max_beta=0.01):
    # NB: the YAML's max_beta is never forwarded (ddpm.py:67) -> always 1e-4 .. 0.01
    return np.linspace(1e-4, max_beta, timesteps)


def cosine_beta_schedule(timesteps, s=0.008):
    steps = timesteps + 1
    x = np.linspace(0, steps, steps)
    ac = np.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = 1 - (ac[1:] / ac[:-1])
    return np.clip(betas, a_min=0, a_max=0.999)


class DiffusionSchedule:
    """The registered buffers of ``GaussianDiffusion.__init__`` (ddpm.py:64-101): float64 NumPy -> fp32."""

    def __init__(self, timesteps=1000, schedule_type='linear', betas=None, dtype=torch.float32):
        if betas is None:
            betas = {'linear': linear_beta_schedule, 'cosine': cosine_beta_schedule}[schedule_type](timesteps)
        betas = np.asarray(betas, dtype=np.float64)
        alphas = 1. - betas
        ac = np.cumprod(alphas, axis=0)
        ac_prev = np.append(1., ac[:-1])
        f = lambda a: torch.tensor(a, dtype=torch.float32).to(dtype)
        self.timesteps = len(betas)
        self.betas = f(betas)
        self.alphas_cumprod = f(ac)
        self.sqrt_alphas_cumprod = f(np.sqrt(ac))
        self.sqrt_one_minus_alphas_cumprod = f(np.sqrt(1. - ac))
        self.sqrt_recip_alphas_cumprod = f(np.sqrt(1. / ac))
        self.sqrt_recipm1_alphas_cumprod = f(np.sqrt(1. / ac - 1))
        pv = betas * (1. - ac_prev) / (1. - ac)
        self.posterior_log_variance_clipped = f(np.log(np.maximum(pv, 1e-20)))
        self.posterior_mean_coef1 = f(betas * np.sqrt(ac_prev) / (1. - ac))
        self.posterior_mean_coef2 = f((1. - ac_prev) * np.sqrt(alphas) / (1. - ac))


def q_sample(sch: DiffusionSchedule, x_start, t: int, noise):
    """ddpm.py:206-210 with a scalar step index."""
    return sch.sqrt_alphas_cumprod[t] * x_start + sch.sqrt_one_minus_alphas_cumprod[t] * noise


# --------------------------------------------------------------------------
# DDPM ancestral / DDIM / PLMS   (ddpm.py:123-204)
# --------------------------------------------------------------------------
def _tvec(i, b):
    return torch.full((b,), i, dtype=torch.long)


def sample_ddpm(denoise, sch, x, cond, t_max, step_noise):
    """ddpm.py:346-349 + p_sample :149-156.  ``step_noise[j]`` is the draw of the j-th iteration
    (i = t_max-1-j); one is consumed every step INCLUDING t = 0 where it is masked out."""
    b = x.shape[0]
    for j, i in enumerate(reversed(range(0, t_max))):
        eps = denoise(x, _tvec(i, b), cond)
        x_recon = sch.sqrt_recip_alphas_cumprod[i] * x - sch.sqrt_recipm1_alphas_cumprod[i] * eps
        mean = sch.posterior_mean_coef1[i] * x_recon + sch.posterior_mean_coef2[i] * x
        nonzero = 0.0 if i == 0 else 1.0
        x = mean + nonzero * (0.5 * sch.posterior_log_variance_clipped[i]).exp() * step_noise[j].to(x.dtype)
    return x


def sample_ddim(denoise, sch, x, cond, t_max, interval):
    """ddpm.py:334-343 + p_sample_ddim :158-167."""
    b = x.shape[0]
    for i in reversed(range(0, t_max, interval)):
        a_t = sch.alphas_cumprod[i]
        a_prev = sch.alphas_cumprod[max(i - interval, 0)]
        eps = denoise(x, _tvec(i, b), cond)
        x = a_prev.sqrt() * (x / a_t.sqrt() + (((1 - a_prev) / a_prev).sqrt() - ((1 - a_t) / a_t).sqrt()) * eps)
    return x


def sample_plms(denoise, sch, x, cond, t_max, interval):
    """ddpm.py:323-333 + p_sample_plms :169-204.

    The reference crashes for B > 1 (``max(t - interval, 0)`` on a [B] tensor, :192); the restatement
    uses the scalar step index so it is well-defined for any B with the B = 1 semantics.
    """
    b = x.shape[0]
    hist = deque(maxlen=4)

    def x_pred(x, eps, i):
        a_t = sch.alphas_cumprod[i]
        a_prev = sch.alphas_cumprod[max(i - interval, 0)]
        a_t_sq, a_prev_sq = a_t.sqrt(), a_prev.sqrt()
        delta = (a_prev - a_t) * ((1 / (a_t_sq * (a_t_sq + a_prev_sq))) * x
                                  - 1 / (a_t_sq * (((1 - a_prev) * a_t).sqrt() + ((1 - a_t) * a_prev).sqrt())) * eps)
        return x + delta

    for i in reversed(range(0, t_max, interval)):
        eps = denoise(x, _tvec(i, b), cond)
        if len(hist) == 0:
            xp = x_pred(x, eps, i)
            eps_prev = denoise(xp, _tvec(max(i - interval, 0), b), cond)
            eps_prime = (eps + eps_prev) / 2
        elif len(hist) == 1:
            eps_prime = (3 * eps - hist[-1]) / 2
        elif len(hist) == 2:
            eps_prime = (23 * eps - 16 * hist[-1] + 5 * hist[-2]) / 12
        else:
            eps_prime = (55 * eps - 59 * hist[-1] + 37 * hist[-2] - 9 * hist[-3]) / 24
        x = x_pred(x, eps_prime, i)
        hist.append(eps)
    return x


# --------------------------------------------------------------------------
# VP noise schedule for the ODE solvers
# --------------------------------------------------------------------------
def _interp(x, xp, yp):
    """Piece-wise linear interpolation with end-segment extrapolation; same arithmetic as
    ``interpolate_fn`` (dpm_solver_pytorch.py:1253-1292) for a scalar query."""
    K = xp.shape[0]
    idx = int(torch.searchsorted(xp, x, right=False))     # number of keypoints < x
    s = min(max(idx - 1, 0), K - 2)
    return yp[s] + (x - xp[s]) * (yp[s + 1] - yp[s]) / (xp[s + 1] - xp[s])


class NoiseScheduleVPDiscrete:
    """``NoiseScheduleVP('discrete', betas=...)``.

    ``clip=True``  -> dpm_solver_pytorch.py:98-125 (numerical_clip_alpha, may shorten the array);
    ``clip=False`` -> uni_pc.py:74-86 (no clipping).
    """

    def __init__(self, betas, clip: bool, dtype=torch.float32):
        betas = betas.to(dtype)
        log_alphas = 0.5 * torch.log(1 - betas).cumsum(dim=0)
        if clip:
            log_sigmas = 0.5 * torch.log(1. - torch.exp(2. * log_alphas))
            lambs = log_alphas - log_sigmas
            idx = int(torch.searchsorted(torch.flip(lambs, [0]), torch.tensor(-5.1, dtype=dtype)))
            if idx > 0:
                log_alphas = log_alphas[:-idx]
        self.dtype = dtype
        self.log_alpha_array = log_alphas
        self.total_N = log_alphas.shape[0]
        self.T = 1.
        self.t_array = torch.linspace(0., 1., self.total_N + 1)[1:].to(dtype)   # fp32 linspace, then cast

    def log_alpha(self, t):
        return _interp(t, self.t_array, self.log_alpha_array)

    def alpha(self, t):
        return torch.exp(self.log_alpha(t))

    def sigma(self, t):
        return torch.sqrt(1. - torch.exp(2. * self.log_alpha(t)))

    def lam(self, t):
        la = self.log_alpha(t)
        return la - 0.5 * torch.log(1. - torch.exp(2. * la))

    def model_time(self, t):
        """model_wrapper.get_model_input_time (dpm_solver_pytorch.py:271-280)."""
        return (t - 1. / self.total_N) * self.total_N

    def time_steps(self, steps):
        """get_time_steps('time_uniform'): fp32 linspace on the host (dpm_solver_pytorch.py:474)."""
        return torch.linspace(self.T, 1. / self.total_N, steps + 1).to(self.dtype)


def _data_pred(denoise, ns, x, t, cond):
    """data_prediction_fn (dpm_solver_pytorch.py:434-442 / uni_pc.py:282-291): x0 = (x - sigma eps)/alpha."""
    b = x.shape[0]
    eps = denoise(x, ns.model_time(t).reshape(1).expand(b), cond)
    return (x - ns.sigma(t) * eps) / ns.alpha(t)


def sample_dpm_solver_pp(denoise, betas, x, cond, steps, order=2, dtype=torch.float32):
    """DPM-Solver++ multistep, time_uniform, solver_type 'dpmsolver'
    (dpm_solver_pytorch.py:1171-1213, :547-580, :796-831) as called from ddpm.py:246-284."""
    assert order == 2 and steps >= order
    ns = NoiseScheduleVPDiscrete(betas, clip=True, dtype=dtype)
    ts = ns.time_steps(steps)
    t_prev = [ts[0]]
    m_prev = [_data_pred(denoise, ns, x, ts[0], cond)]

    def first_update(x, s, t, m_s):
        h = ns.lam(t) - ns.lam(s)
        return ns.sigma(t) / ns.sigma(s) * x - ns.alpha(t) * torch.expm1(-h) * m_s

    def second_update(x, t):
        m1, m0 = m_prev[-2], m_prev[-1]
        t1, t0 = t_prev[-2], t_prev[-1]
        l1, l0, lt = ns.lam(t1), ns.lam(t0), ns.lam(t)
        h_0 = l0 - l1
        h = lt - l0
        r0 = h_0 / h
        D1_0 = (1. / r0) * (m0 - m1)
        phi_1 = torch.expm1(-h)
        a_t = ns.alpha(t)
        return (ns.sigma(t) / ns.sigma(t0)) * x - (a_t * phi_1) * m0 - 0.5 * (a_t * phi_1) * D1_0

    # step 1 (order 1)
    t = ts[1]
    x = first_update(x, t_prev[-1], t, m_prev[-1])
    t_prev.append(t)
    m_prev.append(_data_pred(denoise, ns, x, t, cond))
    for step in range(order, steps + 1):
        t = ts[step]
        step_order = min(order, steps + 1 - step) if steps < 10 else order      # lower_order_final, :1198
        if step_order == 1:
            x = first_update(x, t_prev[-1], t, m_prev[-1])
        else:
            x = second_update(x, t)
        t_prev[0], m_prev[0] = t_prev[1], m_prev[1]
        t_prev[-1] = t
        if step < steps:                                                       # no final model eval, :1212
            m_prev[-1] = _data_pred(denoise, ns, x, t, cond)
    return x


def sample_unipc(denoise, betas, x, cond, steps, order=2, dtype=torch.float32):
    """UniPC bh2, data prediction, multistep, time_uniform, lower_order_final
    (uni_pc.py:590-672 + multistep_uni_pc_bh_update :471-588) as called from ddpm.py:285-322."""
    assert order == 2 and steps >= order
    ns = NoiseScheduleVPDiscrete(betas, clip=False, dtype=dtype)
    ts = ns.time_steps(steps)
    t_prev = [ts[0]]
    m_prev = [_data_pred(denoise, ns, x, ts[0], cond)]

    def bh_update(x, t, order, use_corrector):
        t0 = t_prev[-1]
        m0 = m_prev[-1]
        l0, lt = ns.lam(t0), ns.lam(t)
        s0, st = ns.sigma(t0), ns.sigma(t)
        a_t = torch.exp(ns.log_alpha(t))
        h = lt - l0
        rks, D1s = [], []
        for i in range(1, order):
            ti = t_prev[-(i + 1)]
            mi = m_prev[-(i + 1)]
            rk = (ns.lam(ti) - l0) / h
            rks.append(rk)
            D1s.append((mi - m0) / rk)
        rks.append(torch.tensor(1., dtype=dtype))
        rks = torch.stack([r.to(dtype) for r in rks])
        hh = -h
        h_phi_1 = torch.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        fact = 1
        B_h = torch.expm1(hh)                                    # bh2, :511-512
        R, bvec = [], []
        for i in range(1, order + 1):
            R.append(torch.pow(rks, i - 1))
            bvec.append(h_phi_k * fact / B_h)
            fact *= (i + 1)
            h_phi_k = h_phi_k / hh - 1 / fact
        R = torch.stack(R)
        bvec = torch.stack(bvec)
        rhos_p = torch.tensor([0.5], dtype=dtype) if len(D1s) > 0 else None       # order 2 simplified, :531-532
        if use_corrector:
            rhos_c = torch.tensor([0.5], dtype=dtype) if order == 1 else torch.linalg.solve(R, bvec)
        x_t_ = st / s0 * x - a_t * h_phi_1 * m0
        pred_res = sum(r * d for r, d in zip(rhos_p, D1s)) if len(D1s) > 0 else 0
        x_t = x_t_ - a_t * B_h * pred_res
        m_t = None
        if use_corrector:
            m_t = _data_pred(denoise, ns, x_t, t, cond)
            corr_res = sum(r * d for r, d in zip(rhos_c[:-1], D1s)) if len(D1s) > 0 else 0
            x_t = x_t_ - a_t * B_h * (corr_res + rhos_c[-1] * (m_t - m0))
        return x_t, m_t

    for step in range(1, order):
        t = ts[step]
        x, m = bh_update(x, t, step, True)
        t_prev.append(t)
        m_prev.append(m)
    for step in range(order, steps + 1):
        t = ts[step]
        step_order = min(order, steps + 1 - step)                # lower_order_final always on, :636-637
        x, m = bh_update(x, t, step_order, use_corrector=(step < steps))
        t_prev[0], m_prev[0] = t_prev[1], m_prev[1]
        t_prev[-1] = t
        if step < steps:
            m_prev[-1] = m                                        # predictor-point model output reused, :645-658
    return x


# --------------------------------------------------------------------------
# GaussianDiffusion.inference   (ddpm.py:221-351)
# --------------------------------------------------------------------------
def gaussian_diffusion_inference(denoise, sch: DiffusionSchedule, cond, *, k_step, timesteps,
                                 use_shallow, K_step_infer, speedup, accelerator, noise0,
                                 x_start=None, step_noise=None, dtype=torch.float32):
    """Returns x [B,T,M] or [B,F,T,M] (normalised domain), as ``inference`` does."""
    depth = K_step_infer if K_step_infer is not None else k_step
    if speedup > 0:
        assert depth % speedup == 0
    noise = noise0.to(dtype)
    t_max = min(depth, k_step) if use_shallow else k_step
    if t_max >= timesteps:
        x = noise
    elif t_max > 0:
        assert x_start is not None, 'Missing shallow diffusion source.'
        x = q_sample(sch, x_start.to(dtype), t_max - 1, noise)
    else:
        assert x_start is not None, 'Missing shallow diffusion source.'
        x = x_start.to(dtype)
    if speedup > 1 and t_max > 0:
        if accelerator == 'dpm-solver':
            x = sample_dpm_solver_pp(denoise, sch.betas[:t_max], x, cond, t_max // speedup, dtype=dtype)
        elif accelerator == 'unipc':
            x = sample_unipc(denoise, sch.betas[:t_max], x, cond, t_max // speedup, dtype=dtype)
        elif accelerator == 'pndm':
            x = sample_plms(denoise, sch, x, cond, t_max, speedup)
        elif accelerator == 'ddim':
            x = sample_ddim(denoise, sch, x, cond, t_max, speedup)
        else:
            raise ValueError(f'Unsupported acceleration algorithm for DDPM: {accelerator}.')
    else:
        x = sample_ddpm(denoise, sch, x, cond, t_max, step_noise)
    x = x.transpose(2, 3)
    return x.squeeze(1)


# --------------------------------------------------------------------------
# RectifiedFlow.inference   (reflow.py:66-138)
# --------------------------------------------------------------------------
def rectified_flow_inference(velocity, cond, *, t_start, use_shallow, algorithm, steps, noise0,
                             time_scale_factor=1000, x_end=None, dtype=torch.float32):
    noise = noise0.to(dtype)
    if use_shallow and t_start > 0:
        assert x_end is not None, 'Missing shallow diffusion source.'
        if t_start >= 1.:
            t_start = 1.
            x = x_end.to(dtype)
        else:
            x = t_start * x_end.to(dtype) + (1 - t_start) * noise
    else:
        t_start = 0.
        x = noise
    tsf = time_scale_factor
    v = lambda x, t: velocity(x, tsf * t, cond)
    if t_start < 1:
        dt = (1.0 - t_start) / max(1, steps)
        dts = torch.tensor([dt]).to(dtype)            # fp32 tensor in the reference (reflow.py:132)
        if algorithm not in ('euler', 'rk2', 'rk4', 'rk5'):
            raise ValueError(f'Unsupported algorithm for Rectified Flow: {algorithm}.')
        for i in range(steps):
            t = t_start + i * dts                     # shape (1,)
            if algorithm == 'euler':
                x = x + v(x, t) * dt
            elif algorithm == 'rk2':
                k1 = v(x, t)
                k2 = v(x + 0.5 * k1 * dt, t + 0.5 * dt)
                x = x + k2 * dt
            elif algorithm == 'rk4':
                k1 = v(x, t)
                k2 = v(x + 0.5 * k1 * dt, t + 0.5 * dt)
                k3 = v(x + 0.5 * k2 * dt, t + 0.5 * dt)
                k4 = v(x + k3 * dt, t + dt)
                x = x + (k1 + 2 * k2 + 2 * k3 + k4) * dt / 6
            else:
                k1 = v(x, t)
                k2 = v(x + 0.25 * k1 * dt, t + 0.25 * dt)
                k3 = v(x + 0.125 * (k2 + k1) * dt, t + 0.25 * dt)
                k4 = v(x + 0.5 * (-k2 + 2 * k3) * dt, t + 0.5 * dt)
                k5 = v(x + 0.0625 * (3 * k1 + 9 * k4) * dt, t + 0.75 * dt)
                k6 = v(x + (-3 * k1 + 2 * k2 + 12 * k3 - 12 * k4 + 8 * k5) * dt / 7, t + dt)
                x = x + (7 * k1 + 32 * k3 + 12 * k4 + 32 * k5 + 7 * k6) * dt / 90
    x = x.transpose(2, 3)
    return x.squeeze(1)


# --------------------------------------------------------------------------
# norm / denorm   (ddpm.py:379-383, 403-421; reflow.py:140-144)
# --------------------------------------------------------------------------
def norm_spec(x, spec_min, spec_max):
    return (x - spec_min) / (spec_max - spec_min) * 2 - 1


def denorm_spec(x, spec_min, spec_max):
    return (x + 1) / 2 * (spec_max - spec_min) + spec_min
