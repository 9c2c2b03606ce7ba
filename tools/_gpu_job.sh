set -x
python -m pytest tests/test_mapper_gpu.py -m gpu -x -q -k "linkage or histgap or identical" 2>&1 | tail -15
python -m pytest tests/test_conn_gpu.py -m gpu -x -q 2>&1 | tail -5
python - <<'PY'
import torch, time
from sctda_b200 import mapper, synth, _lib
ctx=_lib.context(0)
N,G=20000,5000
X = synth.expression_genes_torch(G, N, seed=synth.SEED+11, device='cuda').T.contiguous()
lens = mapper.pca_lens_device(X, ctx=ctx)
for link in ('single','ward','average'):
    mapper.build_graph_device(ctx, X, lens, 30, 3.0, metric='correlation', linkage=link)
    prof={}; res=mapper.build_graph_device(ctx, X, lens, 30, 3.0, metric='correlation', linkage=link, profile=prof)
    print(link, 'build ms', round(prof['total_ms'],1), 'V', res.V, 'E', res.E, flush=True)
PY
