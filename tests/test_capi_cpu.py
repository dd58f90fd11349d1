"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/b200_spgemm.h declares, refuses to compute without a GPU, and its host-only pieces
(partition arithmetic, generators) agree with the checker.  No device compute here."""
import ctypes as C
import os
import subprocess
import sys
import re

import numpy as np
import pytest

import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported(smf):
    header = open(os.path.join(ROOT, "include", "b200_spgemm.h")).read()
    declared = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", header))
    declared -= {"b200_csr_t"}
    assert len(declared) >= 25
    lib = smf._lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/b200_spgemm.h but not exported"
    assert declared == set(smf._lib.SIGNATURES), declared ^ set(smf._lib.SIGNATURES)
    # the harness-only generator library has its own header; none of it lives in the product
    header = open(os.path.join(ROOT, "include", "b200_synth.h")).read()
    declared = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", header))
    synth = smf._lib.load_synth()
    for name in sorted(declared):
        assert hasattr(synth, name), f"{name} declared in include/b200_synth.h but not exported"
        assert not hasattr(lib, name), f"{name}: generators must stay out of the product library"
    assert declared == set(smf._lib.SYNTH_SIGNATURES)


def test_no_cpu_fallback(smf):
    """Without a CUDA device b200_init must fail and compute entry points must refuse."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = smf._lib.load()
    assert lib.b200_init(0) == 5  # B200_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.b200_last_error()
    A = smf.synth_rmat(5, 4, 1, True)
    with pytest.raises(smf._lib.B200Error):
        A.flops_spmm(A)
    with pytest.raises(smf._lib.B200Error):
        A.spmm_blocks(A, lambda lo, hi, blk: 0)


def test_host_block_cache_policy():
    """b200_host_free keeps large malloc blocks up to B200_HOST_CACHE_GB (smallest evicted first,
    a block smaller than everything kept is not worth an eviction), frees small ones, and
    B200_HOST_CACHE_GB=0 turns it off.  Host logic only; runs in a subprocess because the limit
    is read once."""
    prog = r"""
import ctypes as C, sys
sys.path.insert(0, %r)
from sparse_matrix_with_flops_b200 import _lib
lib = _lib.load(); libc = C.CDLL(None)
libc.malloc.restype = C.c_void_p; libc.malloc.argtypes = [C.c_size_t]
def info():
    b, n, l = C.c_longlong(), C.c_int(), C.c_longlong()
    assert lib.b200_host_cache_info(C.byref(b), C.byref(n), C.byref(l)) == 0
    return b.value, n.value, l.value
MB = 1 << 20
out = []
for mb in (100, 150, 80, 1):
    lib.b200_host_free(C.c_void_p(libc.malloc(mb * MB)))
    out.append(info()[:2])
lib.b200_host_free(None)
lib.b200_host_cache_drop()
out.append(info())
print(out)
""" % ROOT
    def run(gb, pin="0"):
        env = dict(os.environ, B200_HOST_CACHE_GB=gb, B200_HOST_PIN=pin)
        r = subprocess.run([sys.executable, "-c", prog], capture_output=True, text=True, env=env, timeout=120)
        assert r.returncode == 0, r.stderr
        return eval(r.stdout.strip())
    MB = 1 << 20
    a = run("0.2")                       # 214 MB
    assert a[0][1] == 1 and a[0][0] >= 100 * MB
    assert a[1][1] == 1 and 150 * MB <= a[1][0] < 160 * MB      # the 100 MB block made room
    assert a[2] == a[1] and a[3] == a[1]                        # 80 MB not kept, 1 MB freed
    assert a[4][:2] == (0, 0) and abs(a[4][2] - 0.2 * (1 << 30)) < 2
    z = run("0")
    assert all(x[:2] == (0, 0) for x in z)
    # page-locked mode without a device context: blocks are kept as ordinary ones
    assert run("0.2", pin="1") == a


def test_product_does_not_touch_the_oracle():
    """Nothing under sparse_matrix_with_flops_b200/ may reference oracle/."""
    pkg = os.path.join(ROOT, "sparse_matrix_with_flops_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboracle" not in text and "oracle_" not in text and "libref" not in text, f


@pytest.mark.parametrize("parts", [1, 2, 3, 8, 64])
def test_equal_partition_matches_oracle(smf, parts):
    rng = np.random.default_rng(parts)
    for n in (1, 2, 17, 1000):
        f = rng.integers(0, 50, size=n)
        f[rng.integers(0, n)] += 5000  # a hub row
        pre = np.concatenate([[0], np.cumsum(f)]).astype(np.int64)
        assert np.array_equal(smf.arrayEqualPartition64(pre, parts), ol.o_equal_partition64(pre, parts))
    pre = np.zeros(11, dtype=np.int64)  # all-empty product
    assert np.array_equal(smf.arrayEqualPartition64(pre, parts), ol.o_equal_partition64(pre, parts))


def _check_rmclinit_semantics(A):
    assert A.rowPtr[0] == 0 and A.rowPtr[-1] == A.nnz
    for i in range(A.rows):
        cols = A.colInd[A.rowPtr[i]:A.rowPtr[i + 1]]
        assert np.all(np.diff(cols) > 0), "sorted, duplicate-free"
        assert i in cols, "self loop"
        assert np.all(A.values[A.rowPtr[i]:A.rowPtr[i + 1]] == 1.0 / len(cols))


def test_generators(smf):
    A = smf.synth_rmat(7, 8, 12345, True)
    assert A.rows == 128
    _check_rmclinit_semantics(A)
    B = smf.synth_rmat(7, 8, 12345, True)
    assert np.array_equal(A.colInd, B.colInd)  # deterministic
    dense = np.zeros((128, 128), dtype=bool)
    dense[np.repeat(np.arange(128), np.diff(A.rowPtr)), A.colInd] = True
    assert np.array_equal(dense, dense.T), "symmetrised"
    S = smf.synth_stencil27(4, 3, 5)
    assert S.rows == 60
    _check_rmclinit_semantics(S)
    cnt = np.diff(S.rowPtr)
    assert cnt.max() == 27 and cnt.min() == 8
    P, lab = smf.synth_planted(300, 3, 5, 1, 7, want_labels=True)
    _check_rmclinit_semantics(P)
    assert list(np.unique(lab)) == [0, 1, 2]


def test_rmclinit_host_mirror_matches_oracle(smf, golden):
    m = smf.rmclInit(golden["t2_edges_r"], golden["t2_edges_c"], 3)
    assert np.array_equal(m.rowPtr, golden["t2_A_I"]) and np.array_equal(m.colInd, golden["t2_A_J"])
    assert np.array_equal(m.values, golden["t2_A_V"])


def test_cpp_ingest_matches_reference_golden(golden, tmp_path):
    """include/b200_nlibs.hpp: COO::readSNAPFile (edge list and symmetric MatrixMarket),
    orderedAndDuplicatesRemoving, rmclInit and process_args — host code, compiled here with g++.
    The t2 graph must give the rmclInit matrix the UNMODIFIED reference produced (golden t2_A)."""
    import subprocess
    exe = str(tmp_path / "ingest_test.x")
    subprocess.check_call(["g++", "-O1", "-std=c++11", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "ingest_test.cc"), "-o", exe,
                           "-L" + os.path.join(ROOT, "sparse_matrix_with_flops_b200"), "-lb200spgemm",
                           "-Wl,-rpath," + os.path.join(ROOT, "sparse_matrix_with_flops_b200")])

    # orderedAndDuplicatesRemoving against golden vectors of the unmodified reference
    # (tests/golden/make_golden_dedup.py): repeated pairs summed, the NEW nnz returned
    gd = np.load(os.path.join(ROOT, "tests", "golden", "golden_dedup_v1.npz"))
    for tag in ("w", "t"):
        text = "".join("%d %d %.17g\n" % t for t in zip(gd[tag + "_in_r"], gd[tag + "_in_c"], gd[tag + "_in_v"]))
        out = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "t2_edges.snap"), "1", "2",
                              str(gd[tag + "_shape"][0]), str(gd[tag + "_shape"][1])],
                             input=text, capture_output=True, text=True, timeout=60)
        assert out.returncode == 0, out.stderr
        L = [ln.split() for ln in out.stdout.splitlines()]
        ret, nn = (int(x) for x in [x for x in L if x[0] == "dedup"][0][1:])
        assert ret == nn == int(gd[tag + "_ret"][0])
        rows_ = [x for x in L if x[0] == "d"]
        assert [int(x[1]) for x in rows_] == list(gd[tag + "_out_r"]) and [int(x[2]) for x in rows_] == list(gd[tag + "_out_c"])
        assert np.array_equal(np.array([float(x[3]) for x in rows_]).view(np.int64), gd[tag + "_out_v"].view(np.int64))

    def run(path, trans, c=2):
        out = subprocess.run([exe, path, str(trans), str(c)], capture_output=True, text=True, timeout=60)
        assert out.returncode == 0, out.stderr
        L = [ln.split() for ln in out.stdout.splitlines()]
        return {"pcsr": {"stride": int([x for x in L if x[0] == "pcsr"][0][2]),
                         "bp": [(int(x[1]), int(x[2])) for x in L if x[0] == "bp"],
                         "bv": [(int(x[1]), int(x[2]), float(x[3])) for x in L if x[0] == "bv"]},"coo": [x for x in L if x[0] == "coo"][0], "removed": int([x for x in L if x[0] == "removed"][0][1]),
                "I": np.array([int(x[1]) for x in L if x[0] == "p"], dtype=np.int32),
                "J": np.array([int(x[1]) for x in L if x[0] == "v"], dtype=np.int32),
                "V": np.array([float(x[2]) for x in L if x[0] == "v"]),
                "edges": [(int(x[1]), int(x[2]), float(x[3])) for x in L if x[0] == "e"],
                "opts": [x for x in L if x[0] == "opts"][0]}

    t2 = run(os.path.join(ROOT, "tests", "golden", "t2_edges.snap"), 1)   # transposed, the reference's default
    assert t2["coo"][1:] == ["3", "3", "4"] and t2["removed"] == 0
    assert np.array_equal(t2["I"], golden["t2_A_I"]) and np.array_equal(t2["J"], golden["t2_A_J"])
    assert np.array_equal(t2["V"].view(np.int64), golden["t2_A_V"].view(np.int64))
    sym = run(os.path.join(ROOT, "tests", "golden", "sym4.mtx"), 1)
    # 1-based, mirrored off-diagonals: (0,0) (1,0) (0,1) (2,1) (1,2) (3,3) (3,0) (0,3)
    assert sorted((r, c) for r, c, _ in sym["edges"]) == [(0, 0), (0, 1), (0, 3), (1, 0), (1, 2), (2, 1), (3, 0), (3, 3)]
    assert dict(((r, c), v) for r, c, v in sym["edges"])[(0, 3)] == 3.0
    # rmclInit adds the missing self loops (vertices 1 and 2) and sets 1/rowcount
    assert list(sym["I"]) == [0, 3, 6, 8, 10] and np.allclose(sym["V"][:3], 1 / 3)
    assert t2["opts"][1:] == ["some.snap", "7", "4", "64", "1"]    # SOMP == 4 (nlibs/qrmcl.h:8)
    # PCSR split of the rmclInit matrix against the checker's restatement of nlibs/PCSR.cc:3-56
    for res, c in ((sym, 2), (run(os.path.join(ROOT, "tests", "golden", "sym4.mtx"), 1, 3), 3)):
        Mx = ol.M(res["I"], res["J"], res["V"], len(res["I"]) - 1, len(res["I"]) - 1)
        bp, rp, J, V = ol.o_pcsr_split(Mx, c)
        assert res["pcsr"]["stride"] == (Mx.cols + c - 1) // c
        got_rp = np.array([v for _, v in res["pcsr"]["bp"]], dtype=np.int32).reshape(c, Mx.rows + 1)
        assert np.array_equal(got_rp, rp)
        assert [x[1] for x in res["pcsr"]["bv"]] == list(J) and np.array_equal(np.array([x[2] for x in res["pcsr"]["bv"]]), V)
