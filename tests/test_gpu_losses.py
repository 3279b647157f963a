"""Training-branch losses, forward values (forward half): b2s_masked_loss_f32 through
``xiaoicesing_io_b200.DiffusionLoss`` / ``RectifiedFlowLoss`` against the reference's formulas (modules/losses/diff_loss.py:17-37,
reflow_loss.py:18-50, restated literally below with torch ops in float64 for the yardstick and float32 as the reference runs them)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_loss(a, b, non_padding, t, kind, log_norm, dtype):
    a, b = a.to(dtype), b.to(dtype)
    if non_padding is not None:
        m = non_padding.to(dtype).transpose(1, 2).unsqueeze(1)                          # diff_loss.py:19-21
        a, b = a * m, b * m
    el = (a - b).abs() if kind == 'l1' else (a - b) ** 2
    if t is not None and log_norm:                                                      # reflow_loss.py:26-33
        eps = 1e-7
        tt = torch.clip(t.float(), 0 + eps, 1 - eps)
        w = 0.398942 / tt / (1 - tt) * torch.exp(-0.5 * torch.log(tt / (1 - tt)) ** 2) + eps
        el = w[:, None, None, None].to(dtype) * el
    return el.mean()


@pytest.mark.parametrize('kind', ['l1', 'l2'])
@pytest.mark.parametrize('B,F,M,T,mask_m', [(3, 1, 128, 257, 1), (2, 2, 16, 33, 16), (1, 1, 8, 5, 0), (16, 1, 128, 690, 1)])
def test_losses_match_the_reference_formulas(B, F, M, T, mask_m, kind):
    import xiaoicesing_io_b200 as P
    g = torch.Generator().manual_seed(B * 1000 + T)
    a, b = torch.randn(B, F, M, T, generator=g), torch.randn(B, F, M, T, generator=g)
    mask = None
    if mask_m:
        mask = (torch.rand(B, T, mask_m, generator=g) > 0.2).float()
        mask[:, T // 2:, :] *= (torch.arange(B)[:, None, None] % 2 == 0).float()      # padded tails on every other utterance
    t = 0.05 + 0.9 * torch.rand(B, generator=g)
    t[0] = 0.0                                                                        # clipped to eps
    d = lambda x: None if x is None else x.cuda()
    for cls, tt, log_norm in ((P.DiffusionLoss, None, False), (P.RectifiedFlowLoss, t, True), (P.RectifiedFlowLoss, t, False)):
        mod = cls(kind) if cls is P.DiffusionLoss else cls(kind, log_norm=log_norm)
        out = mod(a.cuda(), b.cuda(), d(mask)) if cls is P.DiffusionLoss else mod(a.cuda(), b.cuda(), t.cuda(), d(mask))
        ref64 = float(_ref_loss(a, b, mask, tt, kind, log_norm, torch.float64))
        ref32 = float(_ref_loss(a, b, mask, tt, kind, log_norm, torch.float32))
        got = float(out)
        assert out.shape == () and abs(got - ref64) <= 2e-6 * max(1.0, abs(ref64)) + 4 * abs(ref32 - ref64), (cls.__name__, kind, got, ref64, ref32)
        again = mod(a.cuda(), b.cuda(), d(mask)) if cls is P.DiffusionLoss else mod(a.cuda(), b.cuda(), t.cuda(), d(mask))
        assert float(again) == got                                                    # deterministic


def test_validation_loss_of_the_training_branch():
    """``diffusion(condition, gt_spec, infer=False)`` -> (x_recon, noise) -> DiffusionLoss: the reference's validation step
    (training/acoustic_task.py run_model) end to end on the GPU; finite and near the loss of an untrained denoiser (~1 for l2)."""
    import xiaoicesing_io_b200 as P
    P.hparams.clear()
    P.hparams.update(hidden_size=256, schedule_type='linear', b2s_precision='fp16')
    m = P.GaussianDiffusion(128, backbone_type='wavenet', backbone_args=dict(num_layers=4, num_channels=256, dilation_cycle_length=4),
                            spec_min=[-12.], spec_max=[0.]).cuda().eval()
    torch.manual_seed(0)
    cond, gt = torch.randn(2, 200, 256, device='cuda'), torch.rand(2, 200, 128, device='cuda') * 12 - 12
    x_recon, noise = m(cond, gt_spec=gt, infer=False)
    non_padding = torch.ones(2, 200, 1, device='cuda')
    non_padding[1, 150:] = 0
    loss = P.DiffusionLoss('l2')(x_recon, noise, non_padding)
    assert bool(torch.isfinite(loss)) and 0.2 < float(loss) < 3.0


def test_bad_arguments_raise():
    import xiaoicesing_io_b200 as P
    with pytest.raises(P.B2SError):
        P.DiffusionLoss('l1')(torch.zeros(1, 1, 4, 4), torch.zeros(1, 1, 4, 4))             # CPU tensors
    with pytest.raises(P.B2SError):
        P.DiffusionLoss('l1')(torch.zeros(1, 1, 4, 4, device='cuda'), torch.zeros(1, 1, 4, 4, device='cuda'), torch.ones(1, 4, 3, device='cuda'))
    with pytest.raises(NotImplementedError):
        P.DiffusionLoss('huber')
