"""GPU diagnostic: does torch.randn under CUDA-graph replay reproduce the eager Philox stream?"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import torch
dev = torch.device('cuda:0')
shape = (2, 1, 16, 37)
torch.manual_seed(5)
eager = [torch.randn(shape, device=dev) for _ in range(4)]
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
torch.manual_seed(5)
with torch.cuda.graph(g):
    outs = [torch.randn(shape, device=dev) for _ in range(4)]
torch.manual_seed(5)
g.replay()
torch.cuda.synchronize()
print('randn graph vs eager equal:', [bool(torch.equal(a, b)) for a, b in zip(eager, outs)])
torch.manual_seed(5)
g.replay(); torch.cuda.synchronize()
print('second replay equal:', [bool(torch.equal(a, b)) for a, b in zip(eager, outs)])

import golden_util as GU, product_util as PU
import xiaoicesing_io_b200 as P
from xiaoicesing_io_b200.core import _sampling
fx = GU.Fixture('gd_ddpm_shallow_K12')
model = PU.build_model(fx, dev)
src, cond = fx['src_spec'].to(dev), fx['condition'].to(dev)
P.hparams['b2s_cuda_graph'] = False
torch.manual_seed(5); e1 = model(cond, src_spec=src, infer=True).clone()
torch.manual_seed(5); e2 = model(cond, src_spec=src, infer=True).clone()
print('eager deterministic:', bool(torch.equal(e1, e2)))
P.hparams['b2s_cuda_graph'] = True
_sampling.clear_graph_cache()
outs = []
for i in range(4):
    torch.manual_seed(5)
    outs.append(model(cond, src_spec=src, infer=True).clone())
    print(i, 'vs eager maxabs', float((outs[-1] - e1).abs().max()), 'cache', [type(v).__name__ for v in _sampling._GRAPH_CACHE.values()])
