"""Host-side mirror of the reference's container / driver interface for the hot path.

Names, argument meaning and error behaviour follow the reference (paths relative to the
reference root) so that tests read like the reference's own:

* ``CSR``            — struct CSR, nlibs/CSR.h:23-379 (rowPtr / colInd / values / rows / cols / nnz)
* ``CSR.flops_spmm`` / ``omp_spmm`` / ``somp_spmm`` / ``spmm`` — nlibs/CSR.cc:59-71,108-194
* ``CSR.staticOmpRmclOneStep`` / ``ompRmclOneStep``           — nlibs/CSR.cc:251-276
* ``CSR.toGpuCSR`` / ``DeviceCSR.toCpuCSR`` / ``deviceDispose`` — nlibs/CSR.cc:342-379
* ``gpuSpMMWrapper`` / ``gpuRmclIter``                         — nlibs/gpus/gpu_csr_kernel.h:5-6
* ``rmclInit`` / ``RMCL``                                       — nlibs/qrmcl.cc:126-164

Everything computes on the GPU through the C-ABI (``_lib``); nothing here is a CPU
implementation.  Output rows have ascending columns (the reference sorts later with
``makeOrdered``).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import Stats, c_double_p, c_int_p, check, csr_t

_inited = {"device": None}


def init(device=0):
    """b200_init: select the CUDA device. Raises B200Error when there is no GPU."""
    lib = _lib.load()
    check(lib.b200_init(int(device)))
    _inited["device"] = int(device)


def reload_options():
    """b200_options_reload: re-read the B200_* developer switches from the environment (the
    library reads them once, in b200_init)."""
    _ensure_init()
    check(_lib.load().b200_options_reload())


def set_topk(k):
    """b200_set_topk: keep at most k entries per row in every rMCL step (0 = off, the default)."""
    _ensure_init()
    check(_lib.load().b200_set_topk(int(k)))


def _ensure_init():
    if _inited["device"] is None:
        init(0)


def _ip(a):
    return a.ctypes.data_as(c_int_p)


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _take(ptr, count, dtype):
    """Copy a malloc()'d result block into numpy and free() it (CSR::dispose semantics)."""
    lib = _lib.load()
    if count > 0:
        out = np.ctypeslib.as_array(ptr, shape=(count,)).astype(dtype, copy=True)
    else:
        out = np.zeros(0, dtype=dtype)
    lib.b200_host_free(C.cast(ptr, C.c_void_p))
    return out


class CSR:
    """Host CSR with the reference's field names (nlibs/CSR.h:23-50)."""

    def __init__(self, values, colInd, rowPtr, rows, cols, nnz=None):
        self.rowPtr = np.ascontiguousarray(rowPtr, dtype=np.int32)
        self.colInd = np.ascontiguousarray(colInd, dtype=np.int32)
        self.values = np.ascontiguousarray(values, dtype=np.float64)
        self.rows, self.cols = int(rows), int(cols)
        self.nnz = int(self.rowPtr[self.rows]) if nnz is None else int(nnz)
        assert self.rowPtr.shape[0] == self.rows + 1

    # ---- SpGEMM: all reference variants compute the same product ----------
    def _mul(self, B, rmcl=False):
        assert self.cols == B.rows  # nlibs/CSR.cc:183
        _ensure_init()
        lib = _lib.load()
        IC, JC, Cv = c_int_p(), c_int_p(), c_double_p()
        nnzC = C.c_int(0)
        if rmcl:
            chaos = C.c_double(0.0)
            check(lib.b200_rmcl_onestep_csr(_ip(self.rowPtr), _ip(self.colInd), _dp(self.values), self.nnz,
                                            _ip(B.rowPtr), _ip(B.colInd), _dp(B.values), B.nnz,
                                            C.byref(IC), C.byref(JC), C.byref(Cv), C.byref(nnzC),
                                            self.rows, self.cols, B.cols, C.byref(chaos)))
        else:
            check(lib.b200_spgemm_csr(_ip(self.rowPtr), _ip(self.colInd), _dp(self.values), self.nnz,
                                      _ip(B.rowPtr), _ip(B.colInd), _dp(B.values), B.nnz,
                                      C.byref(IC), C.byref(JC), C.byref(Cv), C.byref(nnzC),
                                      self.rows, self.cols, B.cols))
        out = CSR(_take(Cv, nnzC.value, np.float64), _take(JC, nnzC.value, np.int32),
                  _take(IC, self.rows + 1, np.int32), self.rows, B.cols, nnzC.value)
        if rmcl:
            out.chaos = chaos.value
        return out

    def spmm_blocks(self, B, on_block, block_products=0):
        """C = self x B as consecutive row blocks (b200_spgemm_csr_stream): for products whose nnz
        exceeds the reference's int CSR.  `on_block(row_lo, row_hi, CSR_block)` is called in row
        order; a block's rowPtr starts at 0.  A truthy return value stops the product
        (B200Error with code 8)."""
        assert self.cols == B.rows
        _ensure_init()
        lib = _lib.load()
        failure = []

        def tramp(_user, lo, hi, IC, JC, Cv, nnz):
            try:
                blk = CSR(_take(Cv, nnz, np.float64), _take(JC, nnz, np.int32),
                          _take(IC, hi - lo + 1, np.int32), hi - lo, B.cols, nnz)
                return 1 if on_block(lo, hi, blk) else 0
            except BaseException as e:      # never unwind through the C frames
                failure.append(e)
                return 1

        cb = _lib.block_fn(tramp)
        rc = lib.b200_spgemm_csr_stream(_ip(self.rowPtr), _ip(self.colInd), _dp(self.values), self.nnz,
                                        _ip(B.rowPtr), _ip(B.colInd), _dp(B.values), B.nnz,
                                        self.rows, self.cols, B.cols, int(block_products), cb, None)
        if failure:
            raise failure[0]
        check(rc)

    def flops_spmm(self, B, stride=512):
        """CSR::flops_spmm (nlibs/CSR.cc:182-194) — `stride` is accepted and unused."""
        return self._mul(B)

    omp_spmm = flops_spmm     # nlibs/CSR.cc:122-134
    somp_spmm = flops_spmm    # nlibs/CSR.cc:108-120
    spmm = flops_spmm         # nlibs/CSR.cc:59-71

    def staticOmpRmclOneStep(self, B, thread_datas=None, stride=512):
        """CSR::staticOmpRmclOneStep (nlibs/CSR.cc:265-276): self = Mgt, B = Mt."""
        return self._mul(B, rmcl=True)

    ompRmclOneStep = staticOmpRmclOneStep  # nlibs/CSR.cc:251-263

    # ---- container helpers --------------------------------------------------------------
    def deepCopy(self):
        return CSR(self.values.copy(), self.colInd.copy(), self.rowPtr.copy(), self.rows, self.cols, self.nnz)

    def makeOrdered(self):
        """CSR::makeOrdered (nlibs/CSR.cc:73-86): sort each row by column (host, numpy)."""
        if self.nnz:
            rowid = np.repeat(np.arange(self.rows, dtype=np.int64), np.diff(self.rowPtr))
            order = np.lexsort((self.colInd, rowid))
            self.colInd = self.colInd[order]
            self.values = self.values[order]
        return self

    def rowCount(self, i):
        return int(self.rowPtr[i + 1] - self.rowPtr[i])

    def toGpuCSR(self):
        """CSR::toGpuCSR (nlibs/CSR.cc:342-354)."""
        _ensure_init()
        lib = _lib.load()
        h = csr_t()
        check(lib.b200_csr_upload(_ip(self.rowPtr), _ip(self.colInd), _dp(self.values), self.rows,
                                  self.cols, self.nnz, C.byref(h)))
        return DeviceCSR(h)

    def dispose(self):
        self.rowPtr = self.colInd = self.values = None


class DeviceCSR:
    """Device-resident CSR handle (the role of the device-pointer CSR of nlibs/CSR.cc:342-379)."""

    def __init__(self, handle):
        self.handle = handle

    def info(self):
        lib = _lib.load()
        r, c, z = C.c_int(), C.c_int(), C.c_longlong()
        check(lib.b200_csr_info(self.handle, C.byref(r), C.byref(c), C.byref(z)))
        return r.value, c.value, z.value

    rows = property(lambda self: self.info()[0])
    cols = property(lambda self: self.info()[1])
    nnz = property(lambda self: self.info()[2])

    def toCpuCSR(self, row_lo=None, row_hi=None):
        """CSR::toCpuCSR (nlibs/CSR.cc:358-371); optionally only rows [row_lo,row_hi)."""
        lib = _lib.load()
        rows, cols, _ = self.info()
        I, J, V = c_int_p(), c_int_p(), c_double_p()
        nnz = C.c_int(0)
        if row_lo is None:
            check(lib.b200_csr_download(self.handle, C.byref(I), C.byref(J), C.byref(V), C.byref(nnz)))
            m = rows
        else:
            check(lib.b200_csr_download_rows(self.handle, row_lo, row_hi, C.byref(I), C.byref(J),
                                             C.byref(V), C.byref(nnz)))
            m = row_hi - row_lo
        return CSR(_take(V, nnz.value, np.float64), _take(J, nnz.value, np.int32),
                   _take(I, m + 1, np.int32), m, cols, nnz.value)

    def column_stripe(self, col_lo, col_hi):
        """b200_csr_column_stripe: B[:, col_lo:col_hi) with local columns — one PCSR block
        (nlibs/PCSR.cc:3-56) on the device."""
        h = csr_t()
        check(_lib.load().b200_csr_column_stripe(self.handle, int(col_lo), int(col_hi), C.byref(h)))
        return DeviceCSR(h)

    def makeOrdered(self):
        """CSR::makeOrdered (nlibs/CSR.cc:73-86) on the device."""
        check(_lib.load().b200_csr_sort_rows(self.handle))
        return self

    def deviceDispose(self):
        """CSR::deviceDispose (nlibs/CSR.cc:374-378)."""
        if self.handle:
            _lib.load().b200_csr_free(self.handle)
            self.handle = None

    def row_argmax(self):
        lib = _lib.load()
        lab = np.empty(self.rows, dtype=np.int32)
        check(lib.b200_csr_row_argmax(self.handle, _ip(lab)))
        return lab


def gpuSpMMWrapper(dA, dB, row_lo=None, row_hi=None, want_stats=False):
    """gpuSpMMWrapper (nlibs/gpus/gpu_csr_kernel.cu:128-172): device in, device out."""
    lib = _lib.load()
    h = csr_t()
    st = Stats()
    if row_lo is None:
        check(lib.b200_spgemm_device(dA.handle, dB.handle, C.byref(h), C.byref(st)))
    else:
        check(lib.b200_spgemm_device_rows(dA.handle, dB.handle, row_lo, row_hi, C.byref(h), C.byref(st)))
    out = DeviceCSR(h)
    return (out, st.as_dict()) if want_stats else out


class DevicePCSR:
    """The reference's PCSR (nlibs/PCSR.h:5-101) on the device: c column stripes of width
    ceil(cols / c), block b with local columns; leftMultiply = A x every stripe, glued back row
    by row (PCSR::leftMultiply, correctTests/pcsrTest.cc:7-19)."""

    def __init__(self, dB, c):
        rows, cols, _ = dB.info()
        self.rows, self.cols, self.c = rows, cols, int(c)
        self.stride = (cols + self.c - 1) // self.c
        self.blocks = [dB.column_stripe(b * self.stride, min(cols, (b + 1) * self.stride))
                       for b in range(self.c) if b * self.stride < cols]

    def leftMultiply(self, dA):
        parts = [gpuSpMMWrapper(dA, blk) for blk in self.blocks]
        arr = (csr_t * len(parts))(*[p.handle for p in parts])
        h = csr_t()
        check(_lib.load().b200_csr_concat_cols(arr, len(parts), C.byref(h)))
        for p in parts:
            p.deviceDispose()
        return DeviceCSR(h)

    def dispose(self):
        for blk in self.blocks:
            blk.deviceDispose()
        self.blocks = []


def gpuRmclOneStep(dMgt, dMt, row_lo=None, row_hi=None, want_stats=False):
    """gpuRmclOneStepWrapper (nlibs/gpus/gpu_csr_kernel.cu:243-279). Returns (newMt, chaos)."""
    lib = _lib.load()
    h = csr_t()
    st = Stats()
    chaos = C.c_double(0.0)
    if row_lo is None:
        check(lib.b200_rmcl_step_device(dMgt.handle, dMt.handle, C.byref(h), C.byref(chaos), C.byref(st)))
    else:
        check(lib.b200_rmcl_step_device_rows(dMgt.handle, dMt.handle, row_lo, row_hi, C.byref(h),
                                             C.byref(chaos), C.byref(st)))
    out = DeviceCSR(h)
    return (out, chaos.value, st.as_dict()) if want_stats else (out, chaos.value)


def gpuRmclIter(maxIter, Mgt, Mt, eps=0.0):
    """gpuRmclIter (nlibs/gpus/gpu_csr_kernel.cu:281-312): host Mgt, Mt in; final Mt out.

    Returns (Mt, iters_done, chaos_history)."""
    _ensure_init()
    lib = _lib.load()
    I, J, V = c_int_p(), c_int_p(), c_double_p()
    nnz, iters = C.c_int(0), C.c_int(0)
    hist = np.zeros(max(1, maxIter), dtype=np.float64)
    check(lib.b200_rmcl_iter(int(maxIter), float(eps), _ip(Mgt.rowPtr), _ip(Mgt.colInd), _dp(Mgt.values),
                             Mgt.nnz, _ip(Mt.rowPtr), _ip(Mt.colInd), _dp(Mt.values), Mt.nnz,
                             C.byref(I), C.byref(J), C.byref(V), C.byref(nnz), Mgt.rows, C.byref(iters),
                             _dp(hist)))
    out = CSR(_take(V, nnz.value, np.float64), _take(J, nnz.value, np.int32),
              _take(I, Mgt.rows + 1, np.int32), Mgt.rows, Mt.cols, nnz.value)
    return out, iters.value, hist[:iters.value].copy()


# ---- multi-GPU rMCL (one process per GPU; not in the reference) -----------------

def comm_unique_id():
    """b200_comm_unique_id: the 128-byte NCCL id, created on rank 0 and shipped by the host program."""
    buf = C.create_string_buffer(128)
    check(_lib.load().b200_comm_unique_id(buf))
    return buf.raw


def comm_init(rank, nranks, uid):
    _ensure_init()
    check(_lib.load().b200_comm_init(int(rank), int(nranks), C.create_string_buffer(uid, 128)))


def comm_destroy():
    check(_lib.load().b200_comm_destroy())


def gpuRmclIterSharded(maxIter, dMgt, dMt, eps=0.0, want_counts=False):
    """b200_rmcl_iter_sharded: every rank holds Mgt and Mt (device); rank r computes the r-th
    flops-balanced row block each iteration, the pruned blocks are all-gathered and chaos is
    max-reduced.  dMt is replaced by the final Mt (sorted rows).  Returns
    (iters_done, chaos_history, ms_per_iteration)."""
    lib = _lib.load()
    iters = C.c_int(0)
    hist = np.zeros(max(1, maxIter), dtype=np.float64)
    ms = np.zeros(max(1, maxIter), dtype=np.float64)
    if not isinstance(dMt.handle, csr_t):
        dMt.handle = csr_t(dMt.handle)
    counts = np.zeros(5 * max(1, maxIter), dtype=np.int64)
    check(lib.b200_rmcl_iter_sharded_stats(int(maxIter), float(eps), dMgt.handle, C.byref(dMt.handle),
                                           C.byref(iters), _dp(hist), _dp(ms),
                                           counts.ctypes.data_as(_lib.c_ll_p) if want_counts else None))
    out = (iters.value, hist[:iters.value].copy(), ms[:iters.value].copy())
    if want_counts:   # per iteration: products, nnz(new Mt), unpruned nnz, row tiles, launches
        out += (counts.reshape(-1, 5)[:iters.value].copy(),)
    return out


def rmclInit(rows_idx, cols_idx, n):
    """rmclInit (nlibs/qrmcl.cc:126-134) on a duplicate-free edge list: self loops, sorted rows,
    values 1/rowcount.  Input preparation on the host (numpy); not part of the timed path."""
    r = np.asarray(rows_idx, dtype=np.int64)
    c = np.asarray(cols_idx, dtype=np.int64)
    has = np.zeros(n, dtype=bool)
    has[r[r == c]] = True
    missing = np.nonzero(~has)[0]
    r = np.concatenate([r, missing])
    c = np.concatenate([c, missing])
    order = np.lexsort((c, r))
    r, c = r[order], c[order]
    rowPtr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowPtr, r + 1, 1)
    rowPtr = np.cumsum(rowPtr)
    cnt = np.diff(rowPtr)
    vals = np.repeat(1.0 / np.maximum(cnt, 1), cnt)
    return CSR(vals, c.astype(np.int32), rowPtr.astype(np.int32), n, n)


COO_DEDUP, COO_SELF_LOOPS, COO_NORMALISE = 1, 2, 4


def cooToGpuCSR(rows_idx, cols_idx, vals, rows, cols, flags=0):
    """COO -> device CSR built on the device (b200_coo_to_csr): COO::toCSR (nlibs/COO.cc:222-235)
    plus, by flag, duplicate removal (COO.cc:237-266), addSelfLoopIfNeeded (COO.cc:160-188) and
    averAndNormRowQValue (nlibs/CSR.cc:88-95).  `vals` may be None (all ones)."""
    _ensure_init()
    lib = _lib.load()
    r = np.ascontiguousarray(rows_idx, dtype=np.int32)
    c = np.ascontiguousarray(cols_idx, dtype=np.int32)
    assert r.shape == c.shape and r.ndim == 1
    v = None if vals is None else np.ascontiguousarray(vals, dtype=np.float64)
    h = csr_t()
    check(lib.b200_coo_to_csr(_ip(r), _ip(c), None if v is None else _dp(v), r.shape[0], rows, cols,
                              flags, C.byref(h)))
    return DeviceCSR(h)


def rmclInitDevice(rows_idx, cols_idx, n, dedup=False):
    """rmclInit (nlibs/qrmcl.cc:126-134) on the device: self loops, (row, col) order, 1/rowcount."""
    return cooToGpuCSR(rows_idx, cols_idx, None, n, n,
                       COO_SELF_LOOPS | COO_NORMALISE | (COO_DEDUP if dedup else 0))


def RMCL(Mt0, maxIters=5, eps=0.0):
    """RMCL (nlibs/qrmcl.cc:136-164) with RunOptions::GPU semantics, starting from an
    rmclInit()'ed matrix instead of a file: Mgt = Mt.deepCopy(); loop; return Mt."""
    Mgt = Mt0.deepCopy()
    Mt, iters, hist = gpuRmclIter(maxIters, Mgt, Mt0, eps)
    Mt.iters_done, Mt.chaos_hist = iters, hist
    return Mt


# ---- flops analysis / partition ------------------------------------------------------------

def flops_prefix(dA, dB):
    """dynamic_omp_CSR_flops (nlibs/flops_csr_kernel.cc:14-31): exclusive prefix, [rows]=P."""
    lib = _lib.load()
    out = np.empty(dA.rows + 1, dtype=np.int64)
    check(lib.b200_flops_prefix(dA.handle, dB.handle, out.ctypes.data_as(_lib.c_ll_p)))
    return out


def cost_prefix(dA, dB, row_charge=-1):
    """b200_cost_prefix: prefix of products + row_charge per heavy row (the device analogue of
    static_omp's footprint, nlibs/static_omp_csr_kernel.cc:28-95); -1 = the library's default."""
    lib = _lib.load()
    rows = dA.info()[0]
    out = np.zeros(rows + 1, dtype=np.int64)
    check(lib.b200_cost_prefix(dA.handle, dB.handle, int(row_charge), out.ctypes.data_as(_lib.c_ll_p)))
    return out


def arrayEqualPartition64(prefix, nparts):
    """arrayEqualPartition64 (nlibs/tools/util.cc:123-135)."""
    lib = _lib.load()
    prefix = np.ascontiguousarray(prefix, dtype=np.int64)
    ends = np.empty(nparts + 1, dtype=np.int32)
    check(lib.b200_equal_partition64(prefix.ctypes.data_as(_lib.c_ll_p), prefix.shape[0] - 1, nparts, _ip(ends)))
    return ends


# ---- synthetic inputs --------------------------------------------------------

# ---- synthetic inputs: libb200synth.so (harness code; the product library is not involved) ----

def _take_plain(lib, ptr, count, dtype):
    out = (np.ctypeslib.as_array(ptr, shape=(count,)).astype(dtype, copy=True) if count > 0
           else np.zeros(0, dtype=dtype))
    lib.free(C.cast(ptr, C.c_void_p))
    return out


def _synth_check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed with code %d" % (what, rc))


def _synth_out(lib, n, I, J, V, nnz):
    return CSR(_take_plain(lib, V, nnz, np.float64), _take_plain(lib, J, nnz, np.int32),
               _take_plain(lib, I, n + 1, np.int32), n, n, nnz)


def synth_rmat(scale, edge_factor=16, seed=12345, symmetrise=False):
    lib = _lib.load_synth()
    n, nnz = C.c_int(), C.c_longlong()
    I, J, V = c_int_p(), c_int_p(), c_double_p()
    _synth_check(lib.b200_synth_rmat(scale, edge_factor, seed, int(symmetrise), C.byref(n), C.byref(I),
                                     C.byref(J), C.byref(V), C.byref(nnz)), "b200_synth_rmat")
    return _synth_out(lib, n.value, I, J, V, nnz.value)


def synth_stencil27(gx, gy, gz):
    lib = _lib.load_synth()
    n, nnz = C.c_int(), C.c_longlong()
    I, J, V = c_int_p(), c_int_p(), c_double_p()
    _synth_check(lib.b200_synth_stencil27(gx, gy, gz, C.byref(n), C.byref(I), C.byref(J), C.byref(V),
                                          C.byref(nnz)), "b200_synth_stencil27")
    return _synth_out(lib, n.value, I, J, V, nnz.value)


def synth_planted(n, nblocks, intra=16, inter=2, seed=12345, want_labels=False):
    lib = _lib.load_synth()
    rows, nnz = C.c_int(), C.c_longlong()
    I, J, V, L = c_int_p(), c_int_p(), c_double_p(), c_int_p()
    _synth_check(lib.b200_synth_planted(n, nblocks, intra, inter, seed, C.byref(rows), C.byref(I), C.byref(J),
                                        C.byref(V), C.byref(nnz), C.byref(L) if want_labels else None),
                 "b200_synth_planted")
    out = _synth_out(lib, rows.value, I, J, V, nnz.value)
    if want_labels:
        return out, _take_plain(lib, L, n, np.int32)
    return out
