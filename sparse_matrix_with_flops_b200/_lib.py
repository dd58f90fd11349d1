"""ctypes binding of the C-ABI in include/b200_spgemm.h.

This module is plumbing only: it loads ``libb200spgemm.so`` (built in-tree by
``sparse_matrix_with_flops_b200/csrc/Makefile``) and declares the argument types of every
entry point.  There is no Python or CPU implementation of the hot path: if the library is
missing, or if ``b200_init`` finds no CUDA device, the import / call fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200spgemm.so")

c_int_p = C.POINTER(C.c_int)
c_double_p = C.POINTER(C.c_double)
c_ll_p = C.POINTER(C.c_longlong)
# b200_block_fn: int fn(void* user, int row_lo, int row_hi, int* IC, int* JC, double* C, int nnzC)
block_fn = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, c_int_p, c_int_p, c_double_p, C.c_int)
csr_t = C.c_void_p


class Stats(C.Structure):
    _fields_ = [
        ("ms_total", C.c_double),
        ("ms_flops", C.c_double),
        ("ms_symbolic", C.c_double),
        ("ms_numeric", C.c_double),
        ("ms_other", C.c_double),
        ("products", C.c_longlong),
        ("nnz_out", C.c_longlong),
        ("nnz_unpruned", C.c_longlong),
        ("launches", C.c_int),
        ("bins_rows", C.c_int * 16),
        ("ms_sym_bin", C.c_double * 16),
        ("ms_num_bin", C.c_double * 16),
        ("sym_bin_rows", C.c_longlong * 16),
        ("sym_bin_products", C.c_longlong * 16),
        ("sym_bin_nnzA", C.c_longlong * 16),
        ("num_bin_products", C.c_longlong * 16),
        ("num_bin_nnzA", C.c_longlong * 16),
        ("num_bin_nnzC", C.c_longlong * 16),
        ("part_kernel", C.c_int),
        ("part_count", C.c_int),
        ("range_items", C.c_int),
        ("ranges", C.c_int),
        ("row_tiles", C.c_int),
    ]

    def as_dict(self):
        d = {}
        for k, t in self._fields_:
            v = getattr(self, k)
            d[k] = list(v) if hasattr(v, "__len__") else v
        return d


# name -> (restype, argtypes); kept in the order of include/b200_spgemm.h
SIGNATURES = {
    "b200_init": (C.c_int, [C.c_int]),
    "b200_finalize": (C.c_int, []),
    "b200_options_reload": (C.c_int, []),
    "b200_set_topk": (C.c_int, [C.c_int]),
    "b200_last_error": (C.c_char_p, []),
    "b200_stream": (C.c_int, [C.POINTER(C.c_void_p)]),
    "b200_device_info": (C.c_int, [c_int_p, c_ll_p, C.c_char_p, C.c_int]),
    "b200_spgemm_csr": (C.c_int, [c_int_p, c_int_p, c_double_p, C.c_int, c_int_p, c_int_p, c_double_p,
                                  C.c_int, C.POINTER(c_int_p), C.POINTER(c_int_p), C.POINTER(c_double_p),
                                  c_int_p, C.c_int, C.c_int, C.c_int]),
    "b200_rmcl_onestep_csr": (C.c_int, [c_int_p, c_int_p, c_double_p, C.c_int, c_int_p, c_int_p,
                                        c_double_p, C.c_int, C.POINTER(c_int_p), C.POINTER(c_int_p),
                                        C.POINTER(c_double_p), c_int_p, C.c_int, C.c_int, C.c_int,
                                        c_double_p]),
    "b200_rmcl_iter": (C.c_int, [C.c_int, C.c_double, c_int_p, c_int_p, c_double_p, C.c_int, c_int_p,
                                 c_int_p, c_double_p, C.c_int, C.POINTER(c_int_p), C.POINTER(c_int_p),
                                 C.POINTER(c_double_p), c_int_p, C.c_int, c_int_p, c_double_p]),
    "b200_spgemm_csr_stream": (C.c_int, [c_int_p, c_int_p, c_double_p, C.c_int, c_int_p, c_int_p,
                                         c_double_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong,
                                         block_fn, C.c_void_p]),
    "b200_csr_upload": (C.c_int, [c_int_p, c_int_p, c_double_p, C.c_int, C.c_int, C.c_int,
                                  C.POINTER(csr_t)]),
    "b200_csr_info": (C.c_int, [csr_t, c_int_p, c_int_p, c_ll_p]),
    "b200_csr_download": (C.c_int, [csr_t, C.POINTER(c_int_p), C.POINTER(c_int_p),
                                    C.POINTER(c_double_p), c_int_p]),
    "b200_csr_download_rows": (C.c_int, [csr_t, C.c_int, C.c_int, C.POINTER(c_int_p),
                                         C.POINTER(c_int_p), C.POINTER(c_double_p), c_int_p]),
    "b200_csr_free": (C.c_int, [csr_t]),
    "b200_csr_sort_rows": (C.c_int, [csr_t]),
    "b200_coo_to_csr": (C.c_int, [c_int_p, c_int_p, c_double_p, C.c_longlong, C.c_int, C.c_int, C.c_int,
                                  C.POINTER(csr_t)]),
    "b200_csr_device_ptrs": (C.c_int, [csr_t, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                       C.POINTER(C.c_void_p)]),
    "b200_spgemm_device": (C.c_int, [csr_t, csr_t, C.POINTER(csr_t), C.POINTER(Stats)]),
    "b200_spgemm_device_rows": (C.c_int, [csr_t, csr_t, C.c_int, C.c_int, C.POINTER(csr_t),
                                          C.POINTER(Stats)]),
    "b200_spgemm_device_stream": (C.c_int, [csr_t, csr_t, C.c_longlong, block_fn, C.c_void_p]),
    "b200_rmcl_step_device": (C.c_int, [csr_t, csr_t, C.POINTER(csr_t), c_double_p, C.POINTER(Stats)]),
    "b200_rmcl_step_device_rows": (C.c_int, [csr_t, csr_t, C.c_int, C.c_int, C.POINTER(csr_t),
                                             c_double_p, C.POINTER(Stats)]),
    "b200_flops_prefix": (C.c_int, [csr_t, csr_t, c_ll_p]),
    "b200_cost_prefix": (C.c_int, [csr_t, csr_t, C.c_longlong, c_ll_p]),
    "b200_equal_partition64": (C.c_int, [c_ll_p, C.c_int, C.c_int, c_int_p]),
    "b200_csr_row_argmax": (C.c_int, [csr_t, c_int_p]),
    "b200_csr_concat_rows": (C.c_int, [C.POINTER(csr_t), C.c_int, C.POINTER(csr_t)]),
    "b200_csr_column_stripe": (C.c_int, [csr_t, C.c_int, C.c_int, C.POINTER(csr_t)]),
    "b200_csr_concat_cols": (C.c_int, [C.POINTER(csr_t), C.c_int, C.POINTER(csr_t)]),
    "b200_comm_unique_id": (C.c_int, [C.c_char_p]),
    "b200_comm_init": (C.c_int, [C.c_int, C.c_int, C.c_char_p]),
    "b200_comm_destroy": (C.c_int, []),
    "b200_rmcl_iter_sharded": (C.c_int, [C.c_int, C.c_double, csr_t, C.POINTER(csr_t), c_int_p,
                                         c_double_p, c_double_p]),
    "b200_rmcl_iter_sharded_stats": (C.c_int, [C.c_int, C.c_double, csr_t, C.POINTER(csr_t), c_int_p,
                                               c_double_p, c_double_p, c_ll_p]),
    "b200_host_free": (None, [C.c_void_p]),
    "b200_host_cache_info": (C.c_int, [c_ll_p, c_int_p, c_ll_p]),
    "b200_host_cache_drop": (C.c_int, []),
    "b200_host_cache_pin": (C.c_int, [C.c_int]),
    "b200_host_cache_stats": (C.c_int, [c_ll_p, c_ll_p, c_ll_p, c_int_p]),
}

# harness-only generators: libb200synth.so (include/b200_synth.h), host code
SYNTH_PATH = os.path.join(_HERE, "libb200synth.so")
SYNTH_SIGNATURES = {
    "b200_synth_rmat": (C.c_int, [C.c_int, C.c_int, C.c_ulonglong, C.c_int, c_int_p,
                                  C.POINTER(c_int_p), C.POINTER(c_int_p), C.POINTER(c_double_p), c_ll_p]),
    "b200_synth_stencil27": (C.c_int, [C.c_int, C.c_int, C.c_int, c_int_p, C.POINTER(c_int_p),
                                       C.POINTER(c_int_p), C.POINTER(c_double_p), c_ll_p]),
    "b200_synth_planted": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_ulonglong, c_int_p,
                                     C.POINTER(c_int_p), C.POINTER(c_int_p), C.POINTER(c_double_p),
                                     c_ll_p, C.POINTER(c_int_p)]),
}

_lib = None
_synth = None


def load_synth():
    """Load the generator library (host code only; no CUDA, no product code)."""
    global _synth
    if _synth is not None:
        return _synth
    if not os.path.exists(SYNTH_PATH):
        raise ImportError(f"{SYNTH_PATH} is missing: build it with `make -C sparse_matrix_with_flops_b200/csrc`")
    lib = C.CDLL(SYNTH_PATH)
    for name, (res, args) in SYNTH_SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib.free.restype = None
    lib.free.argtypes = [C.c_void_p]
    _synth = lib
    return lib


def load():
    """Load the shared library (once) and declare every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C sparse_matrix_with_flops_b200/csrc` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no fallback path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class B200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"b200 error {code}: {msg}")
        self.code = code


def check(rc):
    if rc != 0:
        raise B200Error(rc, load().b200_last_error().decode(errors="replace"))
