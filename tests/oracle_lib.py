"""ctypes access to the checker: oracle/liboracle.so (C restatement) and, when present,
oracle/_ref/libref.so (the unmodified reference).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref.so")

ip = C.POINTER(C.c_int)
dp = C.POINTER(C.c_double)
lp = C.POINTER(C.c_int64)


def _i(a):
    return a.ctypes.data_as(ip)


def _d(a):
    return a.ctypes.data_as(dp)


def build_oracle():
    if not os.path.exists(ORACLE_SO):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "all"])
    if not os.path.exists(REF_SO) and os.path.isdir("/root/reference/nlibs"):
        subprocess.call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])


_o = None
_r = None


def oracle():
    global _o
    if _o is None:
        build_oracle()
        _o = C.CDLL(ORACLE_SO)
        _o.oracle_compute_threshold.restype = C.c_double
        _o.oracle_compute_threshold.argtypes = [C.c_double, C.c_double]
        _o.oracle_chaos.restype = C.c_double
    return _o


def have_ref():
    build_oracle()
    return os.path.exists(REF_SO)


def ref():
    global _r
    if _r is None:
        _r = C.CDLL(REF_SO)
        _r.ref_compute_threshold.restype = C.c_double
        _r.ref_compute_threshold.argtypes = [C.c_double, C.c_double]
        _r.ref_thresh_prune_normalize.restype = C.c_double
        _r.ref_thresh_prune_normalize.argtypes = [C.c_double, ip, dp, ip, ip, dp]
        _r.ref_spgemm_timed.restype = C.c_double
    return _r


def _take(lib_free, ptr, n, dtype):
    out = np.ctypeslib.as_array(ptr, shape=(max(n, 1),))[:n].astype(dtype, copy=True)
    lib_free(C.cast(ptr, C.c_void_p))
    return out


class M:
    """Minimal host CSR for the checker (int32 / float64 numpy)."""

    def __init__(self, I, J, V, rows, cols):
        self.I = np.ascontiguousarray(I, dtype=np.int32)
        self.J = np.ascontiguousarray(J, dtype=np.int32)
        self.V = np.ascontiguousarray(V, dtype=np.float64)
        self.rows, self.cols = rows, cols

    @property
    def nnz(self):
        return int(self.I[self.rows])

    def copy(self):
        return M(self.I.copy(), self.J.copy(), self.V.copy(), self.rows, self.cols)


def from_csr(c):
    return M(c.rowPtr, c.colInd, c.values, c.rows, c.cols)


# ---- oracle wrappers ------------------------------------------------------------------------

def o_spgemm(A, B):
    o = oracle()
    IC, JC, Cv, n = ip(), ip(), dp(), C.c_int()
    rc = o.oracle_spgemm(_i(A.I), _i(A.J), _d(A.V), _i(B.I), _i(B.J), _d(B.V), A.rows, B.cols,
                         C.byref(IC), C.byref(JC), C.byref(Cv), C.byref(n))
    assert rc == 0, rc
    return M(_take(o.oracle_free, IC, A.rows + 1, np.int32), _take(o.oracle_free, JC, n.value, np.int32),
             _take(o.oracle_free, Cv, n.value, np.float64), A.rows, B.cols)


def o_rmcl_onestep(A, B):
    o = oracle()
    IC, JC, Cv, n = ip(), ip(), dp(), C.c_int()
    unp = C.c_int64()
    rc = o.oracle_rmcl_onestep(_i(A.I), _i(A.J), _d(A.V), _i(B.I), _i(B.J), _d(B.V), A.rows, B.cols,
                               C.byref(IC), C.byref(JC), C.byref(Cv), C.byref(n), C.byref(unp))
    assert rc == 0, rc
    out = M(_take(o.oracle_free, IC, A.rows + 1, np.int32), _take(o.oracle_free, JC, n.value, np.int32),
            _take(o.oracle_free, Cv, n.value, np.float64), A.rows, B.cols)
    out.nnz_unpruned = unp.value
    return out


def o_rmcl_iter(G, T, max_iter, eps=0.0):
    o = oracle()
    IM, JM, Mv, n, it = ip(), ip(), dp(), C.c_int(), C.c_int()
    hist = np.zeros(max(1, max_iter))
    rc = o.oracle_rmcl_iter(max_iter, C.c_double(eps), _i(G.I), _i(G.J), _d(G.V), _i(T.I), _i(T.J), _d(T.V),
                            G.rows, C.byref(IM), C.byref(JM), C.byref(Mv), C.byref(n), C.byref(it), _d(hist))
    assert rc == 0, rc
    out = M(_take(o.oracle_free, IM, G.rows + 1, np.int32), _take(o.oracle_free, JM, n.value, np.int32),
            _take(o.oracle_free, Mv, n.value, np.float64), G.rows, G.cols)
    return out, it.value, hist[:it.value].copy()


def o_make_ordered(m):
    oracle().oracle_make_ordered(_i(m.I), _i(m.J), _d(m.V), m.rows)
    return m


def o_flops_prefix(A, B):
    out = np.zeros(A.rows + 1, dtype=np.int64)
    oracle().oracle_flops_prefix(_i(A.I), _i(A.J), _i(B.I), A.rows, out.ctypes.data_as(lp))
    return out


def o_equal_partition64(prefix, parts):
    prefix = np.ascontiguousarray(prefix, dtype=np.int64)
    ends = np.zeros(parts + 1, dtype=np.int32)
    oracle().oracle_equal_partition64(prefix.ctypes.data_as(lp), prefix.shape[0] - 1, parts, _i(ends))
    return ends


def o_set_topk(k):
    """Opt-in top-k of the checker's rMCL epilogue (0 = off); mirrors b200_set_topk."""
    oracle().oracle_set_topk(int(k))


def o_chaos(m):
    return float(oracle().oracle_chaos(_i(m.I), _d(m.V), m.rows))


def o_row_argmax(m):
    lab = np.zeros(m.rows, dtype=np.int32)
    oracle().oracle_row_argmax(_i(m.I), _i(m.J), _d(m.V), m.rows, _i(lab))
    return lab


def o_row_epilogue(cols, vals):
    cols = np.ascontiguousarray(cols, dtype=np.int32).copy()
    vals = np.ascontiguousarray(vals, dtype=np.float64).copy()
    k = oracle().oracle_rmcl_row_epilogue(len(vals), _i(cols), _d(vals))
    return cols[:k], vals[:k]


def o_pcsr_split(m, c):
    bp = np.zeros(c + 1, dtype=np.int32)
    rp = np.zeros(c * (m.rows + 1), dtype=np.int32)
    J = np.zeros(max(1, m.nnz), dtype=np.int32)
    V = np.zeros(max(1, m.nnz), dtype=np.float64)
    oracle().oracle_pcsr_split(_i(m.I), _i(m.J), _d(m.V), m.rows, m.cols, c, _i(bp), _i(rp), _i(J), _d(V))
    return bp, rp.reshape(c, m.rows + 1), J[:m.nnz], V[:m.nnz]


def o_rmcl_init(er, ec, n):
    o = oracle()
    er = np.ascontiguousarray(er, dtype=np.int32)
    ec = np.ascontiguousarray(ec, dtype=np.int32)
    I, J, V, nnz = ip(), ip(), dp(), C.c_int()
    o.oracle_rmcl_init(_i(er), _i(ec), len(er), n, C.byref(I), C.byref(J), C.byref(V), C.byref(nnz))
    return M(_take(o.oracle_free, I, n + 1, np.int32), _take(o.oracle_free, J, nnz.value, np.int32),
             _take(o.oracle_free, V, nnz.value, np.float64), n, n)


# ---- reference wrappers (only when oracle/_ref/libref.so exists) -------------------------------

def r_spgemm(A, B, variant=3):
    r = ref()
    IC, JC, Cv, n = ip(), ip(), dp(), C.c_int()
    r.ref_spgemm(variant, _i(A.I), _i(A.J), _d(A.V), A.nnz, _i(B.I), _i(B.J), _d(B.V), B.nnz,
                 C.byref(IC), C.byref(JC), C.byref(Cv), C.byref(n), A.rows, A.cols, B.cols, 512)
    return M(_take(r.ref_free, IC, A.rows + 1, np.int32), _take(r.ref_free, JC, n.value, np.int32),
             _take(r.ref_free, Cv, n.value, np.float64), A.rows, B.cols)


def r_rmcl_onestep(A, B, variant=2):
    r = ref()
    IC, JC, Cv, n = ip(), ip(), dp(), C.c_int()
    r.ref_rmcl_onestep(variant, _i(A.I), _i(A.J), _d(A.V), A.nnz, _i(B.I), _i(B.J), _d(B.V), B.nnz,
                       C.byref(IC), C.byref(JC), C.byref(Cv), C.byref(n), A.rows, A.cols, B.cols, 512)
    return M(_take(r.ref_free, IC, A.rows + 1, np.int32), _take(r.ref_free, JC, n.value, np.int32),
             _take(r.ref_free, Cv, n.value, np.float64), A.rows, B.cols)


def r_rmcl_iter(G, T, max_iter, run_option=4):
    r = ref()
    IM, JM, Mv, n = ip(), ip(), dp(), C.c_int()
    ms = C.c_double()
    r.ref_rmcl_iter(run_option, max_iter, _i(G.I), _i(G.J), _d(G.V), _i(T.I), _i(T.J), _d(T.V), G.rows,
                    C.byref(IM), C.byref(JM), C.byref(Mv), C.byref(n), C.byref(ms))
    out = M(_take(r.ref_free, IM, G.rows + 1, np.int32), _take(r.ref_free, JM, n.value, np.int32),
            _take(r.ref_free, Mv, n.value, np.float64), G.rows, G.cols)
    out.ms = ms.value
    return out


def r_make_ordered(m):
    ref().ref_make_ordered(_i(m.I), _i(m.J), _d(m.V), m.rows, m.cols)
    return m


def r_flops_prefix(A, B):
    out = np.zeros(A.rows + 1, dtype=np.int64)
    ref().ref_flops_prefix(_i(A.I), _i(A.J), _i(B.I), _i(B.J), A.rows, B.cols, out.ctypes.data_as(lp))
    return out


def r_equal_partition64(prefix, parts):
    prefix = np.ascontiguousarray(prefix, dtype=np.int64).copy()
    ends = np.zeros(parts + 1, dtype=np.int32)
    ref().ref_equal_partition64(prefix.ctypes.data_as(lp), prefix.shape[0] - 1, parts, _i(ends))
    return ends


def r_rmcl_init(er, ec, n):
    r = ref()
    er = np.ascontiguousarray(er, dtype=np.int32)
    ec = np.ascontiguousarray(ec, dtype=np.int32)
    I, J, V, nnz = ip(), ip(), dp(), C.c_int()
    r.ref_rmcl_init(_i(er), _i(ec), len(er), n, C.byref(I), C.byref(J), C.byref(V), C.byref(nnz))
    return M(_take(r.ref_free, I, n + 1, np.int32), _take(r.ref_free, J, nnz.value, np.int32),
             _take(r.ref_free, V, nnz.value, np.float64), n, n)


def r_pcsr_split(m, c):
    bp = np.zeros(c + 1, dtype=np.int32)
    rp = np.zeros(c * (m.rows + 1), dtype=np.int32)
    J = np.zeros(max(1, m.nnz), dtype=np.int32)
    V = np.zeros(max(1, m.nnz), dtype=np.float64)
    ref().ref_pcsr_split(_i(m.I), _i(m.J), _d(m.V), m.rows, m.cols, c, _i(bp), _i(rp), _i(J), _d(V))
    return bp, rp.reshape(c, m.rows + 1), J[:m.nnz], V[:m.nnz]


# ---- strict comparator: rowPtr exact, sorted colInd exact, |dv| <= tol*|v| ------

def assert_same(got, want, tol=1e-12, what=""):
    """`got`/`want` are M with rows sorted by column."""
    assert got.rows == want.rows and got.cols == want.cols, f"{what}: shape"
    np.testing.assert_array_equal(got.I, want.I, err_msg=f"{what}: rowPtr differs")
    np.testing.assert_array_equal(got.J, want.J, err_msg=f"{what}: colInd differs")
    if want.nnz:
        rel = np.abs(got.V - want.V) / np.maximum(np.abs(want.V), 1e-300)
        worst = float(rel.max())
        assert worst <= tol, f"{what}: max relative value error {worst:.3e} > {tol}"
