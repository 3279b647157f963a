"""GPU probe: time the 16-bit depthwise-conv kernel (and the LayerNorm kernel) alone at the config-3 shapes."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from xiaoicesing_io_b200 import _cabi as C

dev = torch.device('cuda:0')
B, T, inner, k, Cc = 64, 690, 2048, 31, 1024
g = torch.randn(B * T, inner, device=dev).to(torch.bfloat16)
p = torch.empty_like(g)
w = torch.randn(k, inner, device=dev)
b = torch.randn(inner, device=dev)
sl = torch.rand(inner, device=dev)


def timeit(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps):
            fn()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


us = timeit(lambda: C.lynx_dwconv_h(g, w, b, sl, p, B, T, inner, k, 0, True))
print(f'dwconv_h  B={B} T={T} inner={inner} k={k}: {us:8.1f} us   {2 * g.numel() * 2 / us / 1e3:7.1f} GB/s   '
      f'{g.numel() * k / us / 1e6:6.2f} TFMA/s')
x = torch.randn(B * T, Cc, device=dev)
cond = torch.randn(B * T, Cc, device=dev).to(torch.bfloat16)
dv = torch.randn(Cc, device=dev)
ga, be = torch.ones(Cc, device=dev), torch.zeros(Cc, device=dev)
h = torch.empty(B * T, Cc, device=dev, dtype=torch.bfloat16)
us = timeit(lambda: C.lynx_prenorm_h(x, cond, Cc, dv, 0, ga, be, h, B, T, Cc, True, True))
print(f'prenorm_h rows={B * T} C={Cc}: {us:8.1f} us   {(x.numel() * 8 + cond.numel() * 2 + h.numel() * 2) / us / 1e3:7.1f} GB/s')
