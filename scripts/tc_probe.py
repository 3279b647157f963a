"""GPU probe: time the tcgen05 GEMM kernel over a (rows, N, K) grid to separate fixed per-launch cost, per-tile
cost and per-k-block cost.  Prints one line per case (CUDA events, isolated back-to-back launches)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from xiaoicesing_io_b200 import _cabi as C

dev = torch.device('cuda:0')
hd = torch.bfloat16


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


print('case rows N K tiles kblocks us TFLOPs cycles_per_tile_kblock')
for rows in (128 * 148, 11040, 128 * 148 * 2):
    for N in (256, 128, 512):
        for K in (64, 256, 768, 3072):
            A = torch.randn(rows, K, device=dev).to(hd)
            W = torch.randn(N, K, device=dev).to(hd)
            out = torch.empty(rows, N, device=dev, dtype=hd)
            fn = lambda: C.tc_linear(A, K, rows, rows, W, K, None, N, K, True, out_h=out, ldoh=N)
            us = timeit(fn)
            tiles = -(-rows // 128) * -(-N // 256)
            rounds = -(-tiles // 148)
            kb = K // 64
            print('linear', rows, N, K, tiles, kb, f'{us:8.2f}', f'{2.0 * rows * N * K / us / 1e6:8.1f}',
                  f'{us * 1.965e3 / (rounds * kb):8.0f}')
# gate kernel at the bench shape
B, T, Cc = 16, 690, 256
y = torch.randn(B * T, Cc, device=dev).to(hd)
Wd = torch.randn(2 * Cc, 3 * Cc, device=dev).to(hd)
cond = torch.randn(B * T, 2 * Cc, device=dev).to(hd)
z = torch.empty(B * T, Cc, device=dev, dtype=hd)
us = timeit(lambda: C.tc_wavenet_gate(y, Wd, cond, 2 * Cc, z, B, T, Cc, 2, True))
print('gate', B * T, 2 * Cc, 3 * Cc, 96 * 2, 12, f'{us:8.2f}', f'{2.0 * B * T * 2 * Cc * 3 * Cc / us / 1e6:8.1f}')
zz = torch.randn(B * T, Cc, device=dev).to(hd)
Wo = torch.randn(2 * Cc, Cc, device=dev).to(hd)
bo = torch.randn(2 * Cc, device=dev)
x = torch.randn(B * T, Cc, device=dev)
skip = torch.randn(B * T, Cc, device=dev)
dv = torch.randn(Cc, device=dev)
us = timeit(lambda: C.tc_wavenet_out(zz, Wo, bo, x, y, skip, None, dv, 0, False, B, T, Cc, True))
print('out', B * T, 2 * Cc, Cc, 87 * 2, 4, f'{us:8.2f}', f'{2.0 * B * T * 2 * Cc * Cc / us / 1e6:8.1f}')
# empty-kernel launch floor
us = timeit(lambda: C.cast_h(x[:8], y[:8], True))
print('cast8 (launch floor)', f'{us:8.2f}')
