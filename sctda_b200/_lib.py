"""ctypes binding of include/sctda_b200.h.

There is no CPU fallback: if the shared library is missing, or no B200 is visible when a
context is created, the calls raise.  Torch is used only as the owner of device memory and
streams; no torch type crosses the C-ABI (plain pointers and sizes).
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'csrc', 'libsctda_b200.so')

c_i32 = ctypes.c_int32
c_i64 = ctypes.c_int64
c_f64 = ctypes.c_double
c_ptr = ctypes.c_void_p

SCTDA_OK = 0
ERR_NAMES = {-1: 'SCTDA_ERR_INVALID', -2: 'SCTDA_ERR_CUDA', -3: 'SCTDA_ERR_NOMEM', -4: 'SCTDA_ERR_UNSUPPORTED'}


class SctdaError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, '%s (%d): %s' % (ERR_NAMES.get(code, 'error'), code, msg))
        self.code = code


class Graph(ctypes.Structure):
    """struct sctda_graph."""
    _fields_ = [('V', c_i32), ('S', c_i32), ('Ncells', c_i32), ('Mtot', c_i32),
                ('d_node_off', c_ptr), ('d_node_cols', c_ptr), ('d_adj_off', c_ptr),
                ('d_adj_idx', c_ptr), ('d_adj_w', c_ptr), ('d_samples', c_ptr)]


# name -> (restype, argtypes); every symbol include/sctda_b200.h declares
SIGNATURES = {
    'sctda_abi_version': (c_i32, []),
    'sctda_ctx_create': (c_i32, [c_i32, ctypes.POINTER(c_ptr)]),
    'sctda_ctx_destroy': (c_i32, [c_ptr]),
    'sctda_last_error': (ctypes.c_char_p, [c_ptr]),
    'sctda_workspace_bytes': (c_i64, [c_ptr]),
    'sctda_launch_count': (c_i64, [c_ptr]),
    'sctda_set_precision': (c_i32, [c_ptr, c_i32]),
    'sctda_get_precision': (c_i32, [c_ptr]),
    'sctda_set_linkage': (c_i32, [c_ptr, c_i32]),
    'sctda_get_linkage': (c_i32, [c_ptr]),
    'sctda_set_kernel_timing': (c_i32, [c_ptr, c_i32]),
    'sctda_last_kernel_ms': (c_i32, [c_ptr, ctypes.POINTER(ctypes.c_float)]),
    'sctda_row_normalize_f64': (c_i32, [c_ptr, c_i32, c_i32, c_ptr, c_i64, c_i32, c_ptr, c_i64, c_ptr, c_ptr]),
    'sctda_gram_gather_f64': (c_i32, [c_ptr, c_i32, c_ptr, c_i64, c_i32, c_ptr, c_i32, c_ptr, c_ptr, c_i64, c_ptr]),
    'sctda_pairdist_rows_f64': (c_i32, [c_ptr, c_i32, c_i32, c_ptr, c_i64, c_ptr, c_i32, c_i32, c_i32, c_ptr, c_i64,
                                        c_ptr]),
    'sctda_pca_lens_f64': (c_i32, [c_ptr, c_i32, c_i32, c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_i32, c_f64, c_ptr]),
    'sctda_smacof_pass': (c_i32, [c_ptr, c_i32, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr]),
    'sctda_cover_count': (c_i32, [c_ptr, c_i32, c_ptr, c_i64, c_i32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    'sctda_cover_fill': (c_i32, [c_ptr, c_i32, c_i32, c_ptr, c_ptr, c_ptr]),
    'sctda_bin_cluster': (c_i32, [c_ptr, c_i32, c_ptr, c_i64, c_i32, c_ptr, c_ptr, c_i32, c_i32, c_ptr, c_ptr, c_ptr,
                                  c_ptr]),
    'sctda_nodes_count': (c_i32, [c_ptr, c_i32, c_ptr, c_ptr, c_ptr]),
    'sctda_nodes_fill': (c_i32, [c_ptr, c_i32, c_ptr, c_ptr, c_ptr, c_ptr, c_i32, c_ptr, c_ptr, c_ptr, c_ptr]),
    'sctda_nerve_count': (c_i32, [c_ptr, c_i32, c_i32, c_ptr, c_ptr, c_i64, c_ptr, c_ptr]),
    'sctda_nerve_fill': (c_i32, [c_ptr, c_i32, c_ptr, c_ptr]),
    'sctda_csr_select': (c_i32, [c_ptr, c_i32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    'sctda_allgather_csr': (c_i32, [c_ptr, c_ptr, c_i32, c_ptr, c_i64, c_i64, c_ptr, c_i32, c_i32, c_i32, c_ptr, c_ptr, c_ptr,
                                    c_ptr, c_ptr, c_ptr, c_ptr]),
    'sctda_gather_results': (c_i32, [c_ptr, c_ptr, c_i32, c_ptr, c_i64, c_i64, c_ptr, c_ptr]),
    'sctda_jsd_matrix': (c_i32, [c_ptr, c_i32, c_i32, c_ptr, c_i64, c_ptr, c_i64, c_ptr]),
    'sctda_mt19937_shuffle_rows': (c_i32, [c_ptr, ctypes.POINTER(c_i32), c_i32, c_i32, c_ptr]),
    'sctda_parse_table_f64': (c_i32, [ctypes.c_char_p, c_i64, c_i32, c_i32, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_i64,
                                      ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), c_i32]),
    'sctda_format_table_f64': (c_i32, [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_i32, c_i32, c_ptr, c_i64, c_ptr, c_i32]),
    'sctda_root_scores': (c_i32, [c_ptr, c_i32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    'sctda_selftest_div': (c_i32, [c_ptr, c_i64, ctypes.POINTER(c_i64)]),
    'sctda_probe_l2_read_gbs': (c_i32, [c_ptr, ctypes.POINTER(c_f64)]),
    'sctda_gemm_tile_fp64': (c_i32, []),
    'sctda_selftest_log1p': (c_i32, [c_ptr, c_i64, ctypes.POINTER(c_i64)]),
    'sctda_node_stats': (c_i32, [c_ptr, ctypes.POINTER(Graph), c_i32, c_ptr, c_i64, c_i32, c_ptr,
                                 c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    'sctda_conn_perm_test_ordered': (c_i32, [c_ptr, ctypes.POINTER(Graph), c_i32, c_ptr, c_i64, c_i32,
                                             c_i32, c_ptr, c_ptr, c_ptr, c_f64, c_ptr, c_ptr, c_ptr, c_ptr]),
    'sctda_selftest_pm_fast': (c_i32, [c_ptr, c_i64, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64)]),
    'sctda_conn_perm_test': (c_i32, [c_ptr, ctypes.POINTER(Graph), c_i32, c_ptr, c_i64, c_i32,
                                     c_i32, c_ptr, c_i64, c_ptr, c_f64, c_ptr, c_ptr, c_ptr, c_ptr]),
}

_lib = None


def load():
    """Loads libsctda_b200.so (built by sctda_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError('%s is missing: run `python -m sctda_b200.build` (there is no CPU fallback)' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


class Context(object):
    """One sctda_ctx on one GPU."""

    def __init__(self, device=0):
        self.lib = load()
        h = c_ptr()
        rc = self.lib.sctda_ctx_create(int(device), ctypes.byref(h))
        if rc != SCTDA_OK:
            raise SctdaError(rc, 'sctda_ctx_create(device=%d) failed: a B200 (sm_100a) GPU is required; '
                                 'there is no CPU fallback' % device)
        self.handle = h
        self.device = int(device)

    def check(self, rc):
        if rc != SCTDA_OK:
            raise SctdaError(rc, self.lib.sctda_last_error(self.handle).decode())

    def launch_count(self):
        return int(self.lib.sctda_launch_count(self.handle))

    timing_enabled = False

    def set_kernel_timing(self, enabled=True):
        self.check(self.lib.sctda_set_kernel_timing(self.handle, 1 if enabled else 0))
        self.timing_enabled = bool(enabled)

    PRECISIONS = {'fp64': 0, 'fp16x3': 1, '3xtf32': 1}      # '3xtf32': the north star's name for the same 22-bit split

    def set_precision(self, mode):
        """'fp64' (default: graphs identical to the oracle) or 'fp16x3' (opt-in tcgen05 contraction of
        (hi, lo) half-precision operand pairs, alias '3xtf32'; see sctda_set_precision in the header)."""
        if mode not in self.PRECISIONS:
            raise ValueError("precision must be 'fp64' or 'fp16x3', got %r" % (mode,))
        self.check(self.lib.sctda_set_precision(self.handle, self.PRECISIONS[mode]))

    def precision(self, mode):
        """Context manager: `with ctx.precision('3xtf32'): ...` (restores the previous mode)."""
        import contextlib

        @contextlib.contextmanager
        def scope():
            old = self.get_precision()
            self.set_precision(mode)
            try:
                yield self
            finally:
                self.set_precision(old)
        return scope()

    LINKAGES = {'single': 0, 'average': 1, 'ward': 2, 'complete': 3}

    def linkage(self, mode):
        """Context manager: `with ctx.linkage('ward'): ...` — the linkage of the per-patch agglomerative clustering
        (sctda_set_linkage; 'single' is the default), restored on exit."""
        import contextlib
        if mode not in self.LINKAGES:
            raise ValueError("linkage must be one of %s, got %r" % (sorted(self.LINKAGES), mode))

        @contextlib.contextmanager
        def scope():
            old = int(self.lib.sctda_get_linkage(self.handle))
            self.check(self.lib.sctda_set_linkage(self.handle, self.LINKAGES[mode]))
            try:
                yield self
            finally:
                self.check(self.lib.sctda_set_linkage(self.handle, old))
        return scope()

    def get_precision(self):
        v = int(self.lib.sctda_get_precision(self.handle))
        return 'fp64' if v == 0 else 'fp16x3'

    def last_kernel_ms(self):
        ms = ctypes.c_float()
        self.check(self.lib.sctda_last_kernel_ms(self.handle, ctypes.byref(ms)))
        return float(ms.value)

    def probe_l2_read_gbs(self):
        v = c_f64()
        self.check(self.lib.sctda_probe_l2_read_gbs(self.handle, ctypes.byref(v)))
        return float(v.value)

    def gemm_tile_fp64(self):
        return int(self.lib.sctda_gemm_tile_fp64())

    def workspace_bytes(self):
        return int(self.lib.sctda_workspace_bytes(self.handle))

    def close(self):
        if self.handle is not None:
            self.lib.sctda_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_contexts = {}


def context(device=0):
    """Process-wide context per device."""
    if device not in _contexts:
        _contexts[device] = Context(device)
    return _contexts[device]
