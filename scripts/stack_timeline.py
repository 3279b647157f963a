"""GPU profiling helper: phase timeline of one CTA of the whole-stack WaveNet kernel (in-kernel globaltimer stamps).
Needs a profiling build:  B2S_BUILD_TLOG=1 python xiaoicesing_io_b200/_build.py --force  (then rebuild without it)."""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import xiaoicesing_io_b200 as P
from xiaoicesing_io_b200 import _cabi as C

dev = torch.device('cuda:0')
P.hparams.clear()
P.hparams.update(hidden_size=256, b2s_precision='bf16')
torch.manual_seed(0)
net = P.build_backbone(128, 1, 'wavenet', dict(num_layers=20, num_channels=256, dilation_cycle_length=4)).to(dev).eval()
torch.nn.init.normal_(net.output_projection.weight, std=0.01)
eng = net._engine(); eng.pack()
B, T = 16, 690
sess = eng.begin(torch.randn(B, T, 256, device=dev), torch.tensor([399.0], device=dev))
x_in = torch.randn(B * T, 128, device=dev); out = torch.empty_like(x_in)
for _ in range(3): sess.eval(x_in, 0, out)
tlog = torch.zeros(20 * 16, dtype=torch.int64, device=dev)
C.lib.b2s_debug_set_stack_tlog(ctypes.c_void_p(tlog.data_ptr()))
sess.eval(x_in, 0, out); torch.cuda.synchronize()
C.lib.b2s_debug_set_stack_tlog(None)
t = tlog.cpu().reshape(20, 16).double()
if getattr(sess, 'tgroups', None):
    names = ['W pre issued', 'own EPI2 seen', 'flags+Y issued', 'Y landed', 'G1b0 issued', 'G1b1 issued', 'G1b2 issued', 'G1b3 issued',
             'G2res issued', 'G2skip issued', 'G1h0 done', 'G1h1 done', 'EPI1h0 done', 'EPI1h1 done', 'G2res done', 'EPI2+flag']
    t0 = t[2, 0]
    print('layer ' + ' '.join(f'{n[:13]:>14s}' for n in names))
    for l in range(2, 9):
        print(f'{l:5d} ' + ' '.join(f'{(t[l, i] - t0) / 1e3:14.2f}' for i in range(len(names))))
    print('per-layer period (us):', [(float(t[l + 1, 15] - t[l, 15]) / 1e3) for l in range(2, 12)])
    sys.exit(0)
names = ['y ready/A issued', 'G1h0 1st slab', 'G1h1 1st slab', 'G1h0 done', 'EPI1h0 done', 'G1h1 done', 'EPI1h1 done',
         'G2res done', 'EPI2res done+flag', 'G2skip done', 'EPI2skip done', 'res chunk0', 'res chunk1', 'res chunk2', 'res chunk3',
         'after fence']
t0 = t[2, 0]
print('layer ' + ' '.join(f'{n[:12]:>13s}' for n in names))
for l in range(2, 9):
    print(f'{l:5d} ' + ' '.join(f'{(t[l, i] - t0) / 1e3:13.2f}' for i in range(len(names))))
print('per-layer period (us):', [(float(t[l + 1, 0] - t[l, 0]) / 1e3) for l in range(2, 12)])
