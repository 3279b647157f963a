"""Host-side sampler compiler: turns a (sampler, schedule, step count) into a flat PROGRAM of
``nfe`` / ``lin`` / ``noise`` ops over named device buffers.

Every sampler on the reference's path is a sequence of denoiser evaluations interleaved with LINEAR
combinations of a handful of tensors whose scalar coefficients are known before the loop starts
(SURVEY.md section 7.1 item 5).  The reference recomputes those scalars on the device every step with
dozens of tiny kernels (``extract`` gathers, ``interpolate_fn`` -> ``torch.sort``, 2x2
``linalg.solve``); here they are computed ONCE on the host in float64 and uploaded as a coefficient
table, so the device loop is ``denoiser kernels + one elementwise kernel`` per step and can be
captured in a CUDA graph.

Reference being restated (algorithms, not code):
  DDPM ancestral  ddpm.py:123-156          DDIM   ddpm.py:158-167       PNDM/PLMS ddpm.py:169-204
  DPM-Solver++ 2M dpm_solver_pytorch.py:98-125, 271-280, 434-442, 474, 547-580, 796-831, 1171-1213
  UniPC bh2       uni_pc.py:74-86, 282-291, 471-588, 590-672
  reflow Euler/RK reflow.py:66-102, 104-138

Model-time values fed to the step embedding are computed with the SAME fp32 expressions as the
reference (``torch.linspace`` in fp32, ``(t - 1/N) * N``, ``t_start + i * dts``), because the
sinusoidal embedding amplifies time differences.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

# ---- well-known buffer names ----------------------------------------------------------------------
X = 'x'                # sampler state
NOISE0 = 'noise0'      # the initial draw of inference() (ddpm.py:227, reflow.py:105), time-major
XSTART = 'x_start'     # normalised shallow-diffusion source, time-major
Z = 'z'                # per-step ancestral noise


@dataclass
class Op:
    kind: str                                   # 'nfe' | 'lin' | 'noise'
    dst: str
    src: Optional[str] = None                   # nfe: input buffer
    t_index: int = -1                           # nfe: row of Program.t_values
    terms: List[Tuple[str, float]] = field(default_factory=list)   # lin: dst = sum coef * buf
    draw: int = -1                              # noise: index of the per-step draw


@dataclass
class Program:
    ops: List[Op] = field(default_factory=list)
    t_values: List[float] = field(default_factory=list)   # model time of every nfe (fp32 values)
    n_draws: int = 0                                        # per-step noise draws (after the initial one)
    needs_noise0: bool = True
    needs_x_start: bool = False
    result: str = X

    # -- builder helpers
    def nfe(self, src: str, t_model: float, dst: str):
        self.t_values.append(float(t_model))
        self.ops.append(Op('nfe', dst, src=src, t_index=len(self.t_values) - 1))

    def lin(self, dst: str, terms):
        terms = [(b, float(c)) for b, c in terms if float(c) != 0.0 or b == dst]
        # merge duplicates
        merged: Dict[str, float] = {}
        for b, c in terms:
            merged[b] = merged.get(b, 0.0) + c
        self.ops.append(Op('lin', dst, terms=list(merged.items())))

    def noise(self, dst: str):
        self.ops.append(Op('noise', dst, draw=self.n_draws))
        self.n_draws += 1

    @property
    def n_nfe(self) -> int:
        return len(self.t_values)

    def buffers(self) -> List[str]:
        names: List[str] = []
        for op in self.ops:
            for b in [op.dst, op.src] + [t[0] for t in op.terms]:
                if b is not None and b not in names:
                    names.append(b)
        return names


# ---------------------------------------------------------------------------------------------------
# DDPM schedule buffers in float64 (ddpm.py:64-101)
# ---------------------------------------------------------------------------------------------------
def linear_beta_schedule(timesteps, max_beta=0.01):
    """ddpm.py:28-33.  ``max_beta`` from the YAML never reaches this function (ddpm.py:67)."""
    return np.linspace(1e-4, max_beta, timesteps)


def cosine_beta_schedule(timesteps, s=0.008):
    """ddpm.py:36-46."""
    steps = timesteps + 1
    x = np.linspace(0, steps, steps)
    ac = np.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = 1 - (ac[1:] / ac[:-1])
    return np.clip(betas, a_min=0, a_max=0.999)


beta_schedule = {'cosine': cosine_beta_schedule, 'linear': linear_beta_schedule}


class DiffusionTables:
    """float64 versions of the registered buffers; ``fp32()`` gives the tensors the module registers."""

    def __init__(self, betas: np.ndarray):
        betas = np.asarray(betas, dtype=np.float64)
        self.betas = betas
        alphas = 1. - betas
        ac = np.cumprod(alphas, axis=0)
        ac_prev = np.append(1., ac[:-1])
        self.alphas_cumprod = ac
        self.alphas_cumprod_prev = ac_prev
        self.sqrt_alphas_cumprod = np.sqrt(ac)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1. - ac)
        self.log_one_minus_alphas_cumprod = np.log(1. - ac)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1. / ac)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1. / ac - 1)
        pv = betas * (1. - ac_prev) / (1. - ac)
        self.posterior_variance = pv
        self.posterior_log_variance_clipped = np.log(np.maximum(pv, 1e-20))
        self.posterior_mean_coef1 = betas * np.sqrt(ac_prev) / (1. - ac)
        self.posterior_mean_coef2 = (1. - ac_prev) * np.sqrt(alphas) / (1. - ac)

    BUFFER_NAMES = ('betas', 'alphas_cumprod', 'alphas_cumprod_prev', 'sqrt_alphas_cumprod',
                    'sqrt_one_minus_alphas_cumprod', 'log_one_minus_alphas_cumprod', 'sqrt_recip_alphas_cumprod',
                    'sqrt_recipm1_alphas_cumprod', 'posterior_variance', 'posterior_log_variance_clipped',
                    'posterior_mean_coef1', 'posterior_mean_coef2')

    def fp32(self):
        return {n: torch.tensor(getattr(self, n), dtype=torch.float32) for n in self.BUFFER_NAMES}

    def rounded(self):
        """The tables as the reference sees them: float64 math on the fp32-rounded buffers."""
        r = DiffusionTables.__new__(DiffusionTables)
        for n in self.BUFFER_NAMES:
            setattr(r, n, np.asarray(getattr(self, n), dtype=np.float32).astype(np.float64))
        return r


# ---------------------------------------------------------------------------------------------------
# initial state (ddpm.py:227-242)
# ---------------------------------------------------------------------------------------------------
def _ddpm_start(prog: Program, tb: DiffusionTables, t_max: int, timesteps: int):
    if t_max >= timesteps:
        prog.lin(X, [(NOISE0, 1.0)])
    elif t_max > 0:
        prog.needs_x_start = True
        i = t_max - 1
        prog.lin(X, [(XSTART, tb.sqrt_alphas_cumprod[i]), (NOISE0, tb.sqrt_one_minus_alphas_cumprod[i])])
    else:
        prog.needs_x_start = True
        prog.lin(X, [(XSTART, 1.0)])


def build_ddpm(tb: DiffusionTables, t_max: int, timesteps: int) -> Program:
    """Ancestral sampling, ``reversed(range(t_max))`` (ddpm.py:346-349, 149-156)."""
    tb = tb.rounded()
    p = Program()
    _ddpm_start(p, tb, t_max, timesteps)
    for i in reversed(range(0, t_max)):
        p.nfe(X, float(i), 'eps')
        p.noise(Z)                                   # drawn every step, t == 0 included (ddpm.py:153)
        c1, c2 = tb.posterior_mean_coef1[i], tb.posterior_mean_coef2[i]
        cx = c1 * tb.sqrt_recip_alphas_cumprod[i] + c2
        ce = -c1 * tb.sqrt_recipm1_alphas_cumprod[i]
        cz = 0.0 if i == 0 else math.exp(0.5 * tb.posterior_log_variance_clipped[i])
        p.lin(X, [(X, cx), ('eps', ce), (Z, cz)])
    return p


def build_ddim(tb: DiffusionTables, t_max: int, timesteps: int, interval: int) -> Program:
    """ddpm.py:334-343, 158-167.  The t = 0 step is an exact identity but still one NFE."""
    tb = tb.rounded()
    p = Program()
    _ddpm_start(p, tb, t_max, timesteps)
    ac = tb.alphas_cumprod
    for i in reversed(range(0, t_max, interval)):
        a_t, a_prev = ac[i], ac[max(i - interval, 0)]
        p.nfe(X, float(i), 'eps')
        cx = math.sqrt(a_prev) / math.sqrt(a_t)
        ce = math.sqrt(a_prev) * (math.sqrt((1 - a_prev) / a_prev) - math.sqrt((1 - a_t) / a_t))
        p.lin(X, [(X, cx), ('eps', ce)])
    return p


def build_plms(tb: DiffusionTables, t_max: int, timesteps: int, interval: int) -> Program:
    """PNDM / PLMS (ddpm.py:323-333, 169-204).  Well-defined for any batch size (the reference itself
    raises for B > 1 at ddpm.py:192; per-utterance semantics are those of its B = 1 run)."""
    tb = tb.rounded()
    p = Program()
    _ddpm_start(p, tb, t_max, timesteps)
    ac = tb.alphas_cumprod

    def xpred_coefs(i):
        a_t, a_prev = ac[i], ac[max(i - interval, 0)]
        a_t_sq, a_prev_sq = math.sqrt(a_t), math.sqrt(a_prev)
        kx = 1.0 / (a_t_sq * (a_t_sq + a_prev_sq))
        ke = 1.0 / (a_t_sq * (math.sqrt((1 - a_prev) * a_t) + math.sqrt((1 - a_t) * a_prev)))
        return 1.0 + (a_prev - a_t) * kx, -(a_prev - a_t) * ke

    hist: List[str] = []                 # most recent last, at most 3 older eps are ever used
    ring = ['eps0', 'eps1', 'eps2', 'eps3']
    n = 0
    for i in reversed(range(0, t_max, interval)):
        cur = ring[n % 4]
        p.nfe(X, float(i), cur)
        cx, ce = xpred_coefs(i)
        if len(hist) == 0:
            p.lin('x_pred', [(X, cx), (cur, ce)])
            p.nfe('x_pred', float(max(i - interval, 0)), 'eps_prev')
            w = [(cur, 0.5), ('eps_prev', 0.5)]
        elif len(hist) == 1:
            w = [(cur, 3 / 2), (hist[-1], -1 / 2)]
        elif len(hist) == 2:
            w = [(cur, 23 / 12), (hist[-1], -16 / 12), (hist[-2], 5 / 12)]
        else:
            w = [(cur, 55 / 24), (hist[-1], -59 / 24), (hist[-2], 37 / 24), (hist[-3], -9 / 24)]
        p.lin(X, [(X, cx)] + [(b, ce * c) for b, c in w])
        hist.append(cur)
        hist = hist[-3:]
        n += 1
    return p


# ---------------------------------------------------------------------------------------------------
# VP noise schedule (discrete), float64 evaluation of the reference's fp32 tables
# ---------------------------------------------------------------------------------------------------
class VPSchedule:
    """``NoiseScheduleVP('discrete', betas=betas[:t_max])``.

    clip=True: DPM-Solver's ``numerical_clip_alpha`` (dpm_solver_pytorch.py:114-125);
    clip=False: UniPC (uni_pc.py:74-86)."""

    def __init__(self, betas_fp32: torch.Tensor, clip: bool):
        b = betas_fp32.detach().to('cpu', torch.float32)
        log_alphas = 0.5 * torch.log(1 - b).cumsum(dim=0)                 # fp32, as the reference
        if clip:
            log_sigmas = 0.5 * torch.log(1. - torch.exp(2. * log_alphas))
            lambs = log_alphas - log_sigmas
            idx = int(torch.searchsorted(torch.flip(lambs, [0]), torch.tensor(-5.1)))
            if idx > 0:
                log_alphas = log_alphas[:-idx]
        self.total_N = int(log_alphas.shape[0])
        self.log_alpha_array = log_alphas.double().numpy()
        self.t_array = torch.linspace(0., 1., self.total_N + 1)[1:].double().numpy()   # fp32 linspace values

    def log_alpha(self, t: float) -> float:
        xp, yp = self.t_array, self.log_alpha_array
        K = len(xp)
        idx = int(np.searchsorted(xp, t, side='left'))
        s = min(max(idx - 1, 0), K - 2)
        return float(yp[s] + (t - xp[s]) * (yp[s + 1] - yp[s]) / (xp[s + 1] - xp[s]))

    def alpha(self, t):
        return math.exp(self.log_alpha(t))

    def sigma(self, t):
        return math.sqrt(1. - math.exp(2. * self.log_alpha(t)))

    def lam(self, t):
        la = self.log_alpha(t)
        return la - 0.5 * math.log(1. - math.exp(2. * la))

    def time_steps(self, steps: int):
        """fp32 ``torch.linspace(T, 1/N, steps+1)`` (dpm_solver_pytorch.py:474) and the fp32 model times
        ``(t - 1/N) * N`` (:278).  Returns (float64 list of the fp32 t values, fp32 model-time list)."""
        ts = torch.linspace(1., 1. / self.total_N, steps + 1)                # fp32
        tm = (ts - 1. / self.total_N) * self.total_N                         # fp32, same expression
        return [float(v) for v in ts.double()], [float(v) for v in tm]


def _x0_terms(ns: VPSchedule, t: float, xbuf: str, ebuf: str):
    """x0 = (x - sigma_t eps) / alpha_t   (data_prediction_fn)."""
    a, s = ns.alpha(t), ns.sigma(t)
    return [(xbuf, 1.0 / a), (ebuf, -s / a)]


def _vp_start(prog, tb, t_max, timesteps):
    _ddpm_start(prog, tb, t_max, timesteps)


def build_dpm_solver_pp(tb: DiffusionTables, betas_fp32: torch.Tensor, t_max: int, timesteps: int, steps: int,
                        order: int = 2) -> Program:
    """DPM-Solver++ multistep order 2, time_uniform, 'dpmsolver' type, as called at ddpm.py:246-284."""
    assert order == 2, 'the reference calls DPM-Solver++ with order=2'
    assert steps >= order
    p = Program()
    _vp_start(p, tb.rounded(), t_max, timesteps)
    ns = VPSchedule(betas_fp32[:t_max], clip=True)
    ts, tm = ns.time_steps(steps)
    m = ['m0', 'm1']                  # ring of x0-predictions
    # init
    p.nfe(X, tm[0], 'eps')
    p.lin(m[0], _x0_terms(ns, ts[0], X, 'eps'))
    t_prev = [ts[0]]
    m_prev = [m[0]]

    def first_update(s, t, ms):
        h = ns.lam(t) - ns.lam(s)
        p.lin(X, [(X, ns.sigma(t) / ns.sigma(s)), (ms, -ns.alpha(t) * math.expm1(-h))])

    def second_update(t):
        m1b, m0b = m_prev[-2], m_prev[-1]
        t1, t0 = t_prev[-2], t_prev[-1]
        l1, l0, lt = ns.lam(t1), ns.lam(t0), ns.lam(t)
        h0, h = l0 - l1, lt - l0
        r0 = h0 / h
        ap = ns.alpha(t) * math.expm1(-h)
        # x = (s_t/s_0) x - ap m0 - 0.5 ap (m0 - m1)/r0
        p.lin(X, [(X, ns.sigma(t) / ns.sigma(t0)), (m0b, -ap - 0.5 * ap / r0), (m1b, 0.5 * ap / r0)])

    # step 1
    first_update(t_prev[-1], ts[1], m_prev[-1])
    p.nfe(X, tm[1], 'eps')
    p.lin(m[1], _x0_terms(ns, ts[1], X, 'eps'))
    t_prev.append(ts[1])
    m_prev.append(m[1])
    for step in range(order, steps + 1):
        t = ts[step]
        step_order = min(order, steps + 1 - step) if steps < 10 else order     # lower_order_final (:1198)
        if step_order == 1:
            first_update(t_prev[-1], t, m_prev[-1])
        else:
            second_update(t)
        t_prev = [t_prev[1], t]
        if step < steps:                                                        # no final model eval (:1212)
            new = m_prev[0]                                                     # recycle the oldest buffer
            p.nfe(X, tm[step], 'eps')
            p.lin(new, _x0_terms(ns, t, X, 'eps'))
            m_prev = [m_prev[1], new]
        else:
            m_prev = [m_prev[1], m_prev[1]]
    return p


def build_unipc(tb: DiffusionTables, betas_fp32: torch.Tensor, t_max: int, timesteps: int, steps: int,
                order: int = 2) -> Program:
    """UniPC bh2, data prediction, multistep, order 2, lower_order_final (ddpm.py:285-322)."""
    assert order == 2, 'the reference calls UniPC with order=2'
    assert steps >= order
    p = Program()
    _vp_start(p, tb.rounded(), t_max, timesteps)
    ns = VPSchedule(betas_fp32[:t_max], clip=False)
    ts, tm = ns.time_steps(steps)
    ring = ['m0', 'm1', 'm2']
    p.nfe(X, tm[0], 'eps')
    p.lin(ring[0], _x0_terms(ns, ts[0], X, 'eps'))
    t_prev = [ts[0]]
    m_prev = [ring[0]]
    free = [ring[1], ring[2]]

    def bh_update(t, tmodel, order_, use_corrector):
        """Emits the ops of multistep_uni_pc_bh_update (uni_pc.py:471-588); returns the buffer holding
        the model value at the predictor point (or None)."""
        t0 = t_prev[-1]
        m0 = m_prev[-1]
        l0, lt = ns.lam(t0), ns.lam(t)
        h = lt - l0
        a_t = ns.alpha(t)
        hh = -h
        h_phi_1 = math.expm1(hh)
        B_h = math.expm1(hh)                                        # bh2 (:511-512)
        rks = []
        d1_terms = []                                               # D1 = (m_i - m0)/rk as (buf, coef) lists
        for i in range(1, order_):
            ti, mi = t_prev[-(i + 1)], m_prev[-(i + 1)]
            rk = (ns.lam(ti) - l0) / h
            rks.append(rk)
            d1_terms.append([(mi, 1.0 / rk), (m0, -1.0 / rk)])
        rks.append(1.0)
        # R, b (:517-523)
        R = np.array([[rk ** (i - 1) for rk in rks] for i in range(1, order_ + 1)], dtype=np.float64)
        bvec = []
        h_phi_k = h_phi_1 / hh - 1
        fact = 1
        for i in range(1, order_ + 1):
            bvec.append(h_phi_k * fact / B_h)
            fact *= (i + 1)
            h_phi_k = h_phi_k / hh - 1 / fact
        bvec = np.array(bvec, dtype=np.float64)
        base = [(X, ns.sigma(t) / ns.sigma(t0)), (m0, -a_t * h_phi_1)]   # x_t_ (:548-551)
        # predictor (:553-558); order 2 uses rho_p = 1/2 (:531-532), order 1 has no D1 term
        pred = list(base)
        if d1_terms:
            rhos_p = [0.5] if order_ == 2 else list(np.linalg.solve(R[:-1, :-1], bvec[:-1]))
            for rho, d1 in zip(rhos_p, d1_terms):
                pred += [(b, -a_t * B_h * rho * c) for b, c in d1]
        if not use_corrector:
            p.lin(X, pred)
            return None
        p.lin('x_pred', pred)
        p.nfe('x_pred', tmodel, 'eps')
        m_t = free.pop(0)
        p.lin(m_t, _x0_terms(ns, t, 'x_pred', 'eps'))
        rhos_c = [0.5] if order_ == 1 else list(np.linalg.solve(R, bvec))   # (:541-544)
        corr = list(base)
        for rho, d1 in zip(rhos_c[:-1], d1_terms):
            corr += [(b, -a_t * B_h * rho * c) for b, c in d1]
        corr += [(m_t, -a_t * B_h * rhos_c[-1]), (m0, a_t * B_h * rhos_c[-1])]
        p.lin(X, corr)
        return m_t

    for step in range(1, order):
        m_t = bh_update(ts[step], tm[step], step, True)
        t_prev.append(ts[step])
        m_prev.append(m_t)
    for step in range(order, steps + 1):
        step_order = min(order, steps + 1 - step)                  # lower_order_final always (:636-637)
        m_t = bh_update(ts[step], tm[step], step_order, use_corrector=(step < steps))
        t_prev = [t_prev[1], ts[step]]
        if step < steps:
            free.append(m_prev[0])                                   # oldest history buffer is recycled
            m_prev = [m_prev[1], m_t]                                # predictor-point model value reused (:645-658)
    return p


# ---------------------------------------------------------------------------------------------------
# rectified flow (reflow.py:66-138)
# ---------------------------------------------------------------------------------------------------
def build_reflow(algorithm: str, steps: int, t_start: float, use_shallow: bool, time_scale_factor: float) -> Program:
    p = Program()
    if use_shallow and t_start > 0:
        p.needs_x_start = True
        if t_start >= 1.:
            t_start = 1.
            p.lin(X, [(XSTART, 1.0)])
        else:
            p.lin(X, [(XSTART, t_start), (NOISE0, 1 - t_start)])
    else:
        t_start = 0.
        p.lin(X, [(NOISE0, 1.0)])
    if t_start >= 1:                                          # reflow.py:115-118: x_end is returned before the algorithm is looked up
        return p
    if algorithm not in ('euler', 'rk2', 'rk4', 'rk5'):
        raise ValueError(f'Unsupported algorithm for Rectified Flow: {algorithm}.')
    dt = (1.0 - t_start) / max(1, steps)
    dts = torch.tensor([dt])                                  # fp32 (reflow.py:132)
    tsf = float(time_scale_factor)

    def tmodel(i, frac):
        t = t_start + i * dts                                 # fp32 tensor arithmetic, as the reference
        if frac:
            t = t + frac * dt
        return float(tsf * t)

    for i in range(steps):
        if algorithm == 'euler':
            p.nfe(X, tmodel(i, 0), 'k1')
            p.lin(X, [(X, 1.0), ('k1', dt)])
        elif algorithm == 'rk2':
            p.nfe(X, tmodel(i, 0), 'k1')
            p.lin('xs', [(X, 1.0), ('k1', 0.5 * dt)])
            p.nfe('xs', tmodel(i, 0.5), 'k2')
            p.lin(X, [(X, 1.0), ('k2', dt)])
        elif algorithm == 'rk4':
            p.nfe(X, tmodel(i, 0), 'k1')
            p.lin('xs', [(X, 1.0), ('k1', 0.5 * dt)])
            p.nfe('xs', tmodel(i, 0.5), 'k2')
            p.lin('xs', [(X, 1.0), ('k2', 0.5 * dt)])
            p.nfe('xs', tmodel(i, 0.5), 'k3')
            p.lin('xs', [(X, 1.0), ('k3', dt)])
            p.nfe('xs', tmodel(i, 1.0), 'k4')
            p.lin(X, [(X, 1.0), ('k1', dt / 6), ('k2', 2 * dt / 6), ('k3', 2 * dt / 6), ('k4', dt / 6)])
        else:  # rk5
            p.nfe(X, tmodel(i, 0), 'k1')
            p.lin('xs', [(X, 1.0), ('k1', 0.25 * dt)])
            p.nfe('xs', tmodel(i, 0.25), 'k2')
            p.lin('xs', [(X, 1.0), ('k2', 0.125 * dt), ('k1', 0.125 * dt)])
            p.nfe('xs', tmodel(i, 0.25), 'k3')
            p.lin('xs', [(X, 1.0), ('k2', -0.5 * dt), ('k3', 1.0 * dt)])
            p.nfe('xs', tmodel(i, 0.5), 'k4')
            p.lin('xs', [(X, 1.0), ('k1', 0.0625 * 3 * dt), ('k4', 0.0625 * 9 * dt)])
            p.nfe('xs', tmodel(i, 0.75), 'k5')
            p.lin('xs', [(X, 1.0), ('k1', -3 * dt / 7), ('k2', 2 * dt / 7), ('k3', 12 * dt / 7),
                         ('k4', -12 * dt / 7), ('k5', 8 * dt / 7)])
            p.nfe('xs', tmodel(i, 1.0), 'k6')
            p.lin(X, [(X, 1.0), ('k1', 7 * dt / 90), ('k3', 32 * dt / 90), ('k4', 12 * dt / 90),
                      ('k5', 32 * dt / 90), ('k6', 7 * dt / 90)])
    return p


def build_gaussian_program(tb: DiffusionTables, betas_fp32: torch.Tensor, *, timesteps: int, k_step: int,
                           use_shallow: bool, K_step_infer: Optional[int], speedup: int, accelerator: str) -> Program:
    """Dispatch of ``GaussianDiffusion.inference`` (ddpm.py:221-351)."""
    depth = K_step_infer if K_step_infer is not None else k_step
    if speedup > 0:
        assert depth % speedup == 0, f'Acceleration ratio must be a factor of diffusion depth {depth}.'
    t_max = min(depth, k_step) if use_shallow else k_step
    if speedup > 1 and t_max > 0:
        if accelerator == 'dpm-solver':
            return build_dpm_solver_pp(tb, betas_fp32, t_max, timesteps, t_max // speedup)
        if accelerator == 'unipc':
            return build_unipc(tb, betas_fp32, t_max, timesteps, t_max // speedup)
        if accelerator == 'pndm':
            return build_plms(tb, t_max, timesteps, speedup)
        if accelerator == 'ddim':
            return build_ddim(tb, t_max, timesteps, speedup)
        raise ValueError(f'Unsupported acceleration algorithm for DDPM: {accelerator}.')
    return build_ddpm(tb, t_max, timesteps)
