"""Times b2s_lynx_dwconv_h alone at config 3's shape (B = 64, T = 704, inner = 2048, K = 31) and prints its error against fp64."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from xiaoicesing_io_b200 import _cabi as C

B, T, inner, K = 64, 704, 2048, 31
for hd, bf in ((torch.float16, False), (torch.bfloat16, True)):
    torch.manual_seed(0)
    g = torch.randn(B, T, inner, device='cuda').to(hd)
    w = torch.randn(inner, K, device='cuda') / K ** 0.5
    bias = torch.randn(inner, device='cuda')
    slope = torch.rand(inner, device='cuda')
    out = torch.empty_like(g)
    wt = w.t().contiguous()
    for _ in range(3):
        C.lynx_dwconv_h(g, wt, bias, slope, out, B, T, inner, K, 0, bf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        C.lynx_dwconv_h(g, wt, bias, slope, out, B, T, inner, K, 0, bf)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    nb = 2
    conv = F.conv1d(g[:nb].double().transpose(1, 2), w.double()[:, None, :], bias.double(), padding=K // 2, groups=inner)
    want = torch.where(conv >= 0, conv, slope.double()[None, :, None] * conv).transpose(1, 2)
    err = (out[:nb].double() - want).abs()
    print(f'{hd}: {us:.1f} us per call, {2 * g.numel() * 2 / us / 1e6:.2f} TB/s; max err {float(err.max()):.3e} (|out| max {float(want.abs().max()):.2f})',
          flush=True)
