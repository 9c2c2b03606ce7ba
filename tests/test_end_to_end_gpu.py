"""BASELINE.json configs[0] — the reference's own CPU-runnable case — end to end through the
reference-facing classes on a B200, next to the CPU oracles on the same files (tools/config1.py):
1,000 cells x 2,000 genes, TopologicalRepresentation(lens='mds', metric='correlation')
.save(resolution=20, gain=2), UnrootedGraph.save(n) with the reference's RNG stream
(/root/reference/scTDA/main.py:735-778, 1053-1086).  n = 100 permutations keeps the oracle's share
of the test under a minute; the graph, the gene table and the p-values must be those of the oracle."""
import argparse
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config1_graph_table_and_pvalues_match_the_oracle():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import config1
    out = config1.run(argparse.Namespace(cells=1000, genes=2000, perms=100, resolution=20, gain=2.0, oracle=True))
    assert out['graph_nodes_identical'] and out['graph_edges_identical']
    assert out['table_rows_identical']
    assert out['rows'] > 500                                    # most genes pass the expression filter
    assert out['p_values_wrong'] == 0
    assert out['p_values_exact'] + out['p_values_tied_genes'] == out['rows']
    assert out['p_values_tied_genes'] <= out['rows'] // 100     # ties only on degenerate genes
    assert out['max_rel_err_connectivity'] < 1e-12
