set -x
python -m pytest tests/test_mapper_gpu.py -m gpu -x -q -k "pca" 2>&1 | tail -15
python - <<'PY'
import torch, time, numpy
from sctda_b200 import mapper, synth, _lib
ctx=_lib.context(0)
for (N,G) in ((20000,5000),(100000,2000)):
    X = synth.expression_genes_torch(G, N, seed=synth.SEED+11, device='cuda').T.contiguous()
    mapper.pca_lens_device(X, ctx=ctx); torch.cuda.synchronize()
    t=time.perf_counter(); lens, info = mapper.pca_lens_device(X, ctx=ctx, return_info=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print(N,G,'pca lens ms', 1e3*dt, info)
    Xc = X - X.mean(dim=0, keepdim=True)
    w,v = torch.linalg.eigh(Xc.T@Xc)
    comps = v[:,[-1,-2]].T
    idx = comps.abs().argmax(dim=1); comps = comps*torch.sign(comps[torch.arange(2),idx])[:,None]
    ref = Xc@comps.T
    print('max abs diff vs eigh', float((ref-lens).abs().max()), 'scale', float(ref.abs().max()), 'eig', w[-1].item(), w[-2].item(), w[-3].item(), w[-9].item())
    del X, Xc
PY
