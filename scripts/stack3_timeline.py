"""GPU profiling helper: phase timeline of CTA 2 of the third whole-stack WaveNet kernel (in-kernel globaltimer stamps).
Needs a profiling build:  B2S_BUILD_TLOG=1 B2S_LIB_OUT=libb2s_tlog.so python xiaoicesing_io_b200/_build.py --force
and B2S_LIB=<that file> when running this script."""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import xiaoicesing_io_b200 as P
from xiaoicesing_io_b200 import _cabi as C

B = int(os.environ.get('TL_B', 16))
T = int(os.environ.get('TL_T', 690))
dev = torch.device('cuda:0')
P.hparams.clear()
P.hparams.update(hidden_size=256, b2s_precision=os.environ.get('TL_PREC', 'bf16'))
torch.manual_seed(0)
CH = int(os.environ.get('TL_C', 256))          # 192: the narrow-model (KS = 3) variant on the padded layout
net = P.build_backbone(128, 1, 'wavenet', dict(num_layers=20, num_channels=CH, dilation_cycle_length=4)).to(dev).eval()
torch.nn.init.normal_(net.output_projection.weight, std=0.01)
eng = net._engine(); eng.pack()
sess = eng.begin(torch.randn(B, T, 256, device=dev), torch.tensor([399.0], device=dev))
assert sess.stack3
x_in = torch.randn(B * T, 128, device=dev); out = torch.empty_like(x_in)
for _ in range(3): sess.eval(x_in, 0, out)
tlog = torch.zeros(32 * 16 + 8, dtype=torch.int64, device=dev)
C.lib.b2s_debug_set_stack_tlog(ctypes.c_void_p(tlog.data_ptr()))
sess.eval(x_in, 0, out); torch.cuda.synchronize()
C.lib.b2s_debug_set_stack_tlog(None)
raw = tlog.cpu()
t = raw[:320].reshape(20, 16).double()
ck = raw[512:520].tolist()
if ck[2] > ck[0]:
    print(f'# kernel span {(ck[2] - ck[0]) / 1e3:.1f} us, {ck[3] - ck[1]} SM clocks -> effective SM clock {(ck[3] - ck[1]) / (ck[2] - ck[0]) * 1e3:.0f} MHz')
names = ['MMA yready', 'MMA halo', 'G1h0 done', 'EPI1h0 done', 'G1h1 done', 'EPI1h1 done', 'EPI2 waits', 'EPI2 loop', 'EPI2 yready', 'SKIP issued',
         'G2 done', 'y_next done', 'IO ydone', 'IO flag out', 'IO flags in', 'MMA issued']
t0 = t[2, 0]
print(f'# B={B} T={T}; CTA 2; microseconds since layer 2 "MMA yready"; IO columns are indexed by the layer whose INPUT they move')
print('layer ' + ' '.join(f'{n[:11]:>12s}' for n in names))
for l in range(2, 9):
    print(f'{l:5d} ' + ' '.join((f'{(t[l, i] - t0) / 1e3:12.2f}' if t[l, i] > 0 else f'{"-":>12s}') for i in range(len(names))))
if ck[4] > 0:
    print(f'# skip/head kernel (CTA 0): starts {(ck[4] - ck[0]) / 1e3:.1f} us after the layer kernel (CTA 2), ends {(ck[5] - ck[2]) / 1e3:.1f} us after it; its layer-l MMAs issued at (us after layer kernel start):',
          [round(float(t[l, 9] - ck[0]) / 1e3, 1) for l in range(20)])
    print('#   layer kernel: MMA yready of layer l at', [round(float(t[l, 0] - ck[0]) / 1e3, 1) for l in range(20)])
print('per-layer period (us):', [round(float(t[l + 1, 0] - t[l, 0]) / 1e3, 2) for l in range(2, 17)])
