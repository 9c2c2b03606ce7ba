"""sctda_b200 — B200-native implementation of the scTDA hot path.

`import sctda_b200 as scTDA` gives the names of the reference package that lie on the rebuilt path
(/root/reference/scTDA/__init__.py:1-10): TopologicalRepresentation, UnrootedGraph, RootedGraph,
benjamini_hochberg, plus write_preprocess_tables (the file formats of Preprocess.save).  The interactive
Preprocess class, ParseAyasdiGraph, compare_results and the drawing helpers are out of scope.  The names are resolved lazily so that importing the package does not
load torch or the CUDA library.
"""

_LAZY = {
    'TopologicalRepresentation': ('mapper', 'TopologicalRepresentation'),
    'UnrootedGraph': ('graph', 'UnrootedGraph'),
    'RootedGraph': ('graph', 'RootedGraph'),
    'benjamini_hochberg': ('graph', 'benjamini_hochberg'),
    'apply_lens': ('mapper', 'apply_lens'),
    'mapper_graph': ('mapper', 'mapper_graph'),
    'write_preprocess_tables': ('formats', 'write_preprocess_tables'),
}

__all__ = sorted(_LAZY)


def __getattr__(name):
    if name in _LAZY:
        import importlib
        mod, attr = _LAZY[name]
        return getattr(importlib.import_module('.' + mod, __name__), attr)
    raise AttributeError('module %r has no attribute %r' % (__name__, name))
