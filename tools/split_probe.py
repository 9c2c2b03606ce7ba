"""Opt-in split-precision (fp16x3) tcgen05 contraction against the FP64 DMMA kernel: error of the distance
matrix, contraction time and the Mapper graph at a given shape.

    python tools/tf32_probe.py [cells genes metric resolution gain]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy  # noqa: E402
import torch  # noqa: E402


def main():
    from sctda_b200 import _lib, mapper, synth
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    g = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    metric = sys.argv[3] if len(sys.argv) > 3 else 'correlation'
    r = int(sys.argv[4]) if len(sys.argv) > 4 else 20
    gain = float(sys.argv[5]) if len(sys.argv) > 5 else 3.0
    ctx = _lib.context(0)
    dev = torch.device('cuda', 0)
    X = synth.expression_matrix(n, g, seed=synth.SEED)[0]
    Xd = mapper._padded_device_table(X, dev)
    out = {'cells': n, 'genes': g, 'metric': metric}
    ctx.set_kernel_timing(True)
    nd = min(n, 8192)
    D = {}
    for mode in ('fp64', 'fp16x3'):
        with ctx.precision(mode):
            for _ in range(2):
                D[mode] = mapper.pairwise_distances(ctx, Xd[:nd], metric)
            torch.cuda.synchronize()
            out['pairdist_ms_' + mode] = ctx.last_kernel_ms()
    err = (D['fp64'] - D['fp16x3']).abs()
    out['pairdist_max_abs_err'] = float(err.max())
    out['pairdist_max_rel_err'] = float((err / D['fp64'].clamp_min(1e-300))[D['fp64'] > 0].max())
    out['pairdist_mean'] = float(D['fp64'].mean())
    del D, err
    lens = mapper.pca_lens_device(Xd).cpu().numpy()
    res = {}
    for mode in ('fp64', 'fp16x3'):
        with ctx.precision(mode):
            for _ in range(2):
                prof = {}
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                res[mode] = mapper.build_graph_device(ctx, Xd, lens, r, gain, metric=metric, profile=prof)
                e1.record()
                torch.cuda.synchronize()
            out['build_ms_' + mode] = e0.elapsed_time(e1)
            out['gemm_ms_' + mode] = ctx.last_kernel_ms()
            out['profile_' + mode] = {k: round(v, 3) for k, v in prof.items() if isinstance(v, (int, float))}
    a, b = res['fp64'], res['fp16x3']
    out['V'] = [int(a.V), int(b.V)]
    out['E'] = [int(a.E), int(b.E)]
    out['labels_identical'] = bool(a.labels.shape == b.labels.shape and (a.labels == b.labels).all())
    out['bins_with_different_K'] = int((a.bin_K != b.bin_K).sum())
    out['graph_identical'] = bool(a.V == b.V and a.E == b.E and (a.node_members == b.node_members).all()
                                  and (a.edges == b.edges).all())
    print(json.dumps(out))


if __name__ == '__main__':
    main()
