"""Golden vectors for COO::orderedAndDuplicatesRemoving (nlibs/COO.cc:237-266) from the UNMODIFIED
reference (oracle/_ref/libref.so, through oracle/ref_shim.cc: ref_coo_dedup).

Run in the build container (where /root/reference exists):
    make -C oracle ref && python tests/golden/make_golden_dedup.py
Writes tests/golden/golden_dedup_v1.npz (committed).

The reference sorts with an UNSTABLE std::sort on (row, col) and adds the values of a run of equal
pairs front to back, so for a pair repeated three or more times its result depends on the order
the sort happens to leave.  The weighted case below therefore repeats pairs at most twice
(a + b is commutative), and the triple case uses values whose sums are exact in any order."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402


def ref_dedup(r, c, v, rows, cols):
    lib = ol.ref()
    n = len(r)
    ro, co, vo = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.float64)
    ret = C.c_int(0)
    nn = lib.ref_coo_dedup(ol._i(r), ol._i(c), ol._d(v), n, rows, cols, ol._i(ro), ol._i(co), ol._d(vo), C.byref(ret))
    assert ret.value == nn
    return ro[:nn].copy(), co[:nn].copy(), vo[:nn].copy(), ret.value


def main():
    assert ol.have_ref(), "oracle/_ref/libref.so missing: make -C oracle ref"
    rng = np.random.default_rng(2026)
    d = {}
    # weighted: 4000 distinct pairs of a 300 x 900 matrix, 900 of them given a second time
    rows, cols, m = 300, 900, 4000
    key = rng.choice(rows * cols, m, replace=False)
    r, c = (key // cols).astype(np.int32), (key % cols).astype(np.int32)
    v = rng.standard_normal(m)
    r2 = np.concatenate([r, r[:900]]).astype(np.int32)
    c2 = np.concatenate([c, c[:900]]).astype(np.int32)
    v2 = np.concatenate([v, rng.standard_normal(900)])
    perm = rng.permutation(len(r2))
    r2, c2, v2 = r2[perm].copy(), c2[perm].copy(), v2[perm].copy()
    d["w_shape"] = np.array([rows, cols], np.int32)
    d["w_in_r"], d["w_in_c"], d["w_in_v"] = r2, c2, v2
    d["w_out_r"], d["w_out_c"], d["w_out_v"], ret = ref_dedup(r2, c2, v2, rows, cols)
    d["w_ret"] = np.array([ret], np.int32)
    # triples with dyadic values: every summation order gives the same bits
    r3 = np.array([2, 0, 2, 1, 2, 0, 1, 1], np.int32)
    c3 = np.array([1, 3, 1, 0, 1, 3, 0, 2], np.int32)
    v3 = np.array([0.5, 1.0, 0.25, 2.0, 0.125, -1.0, 4.0, 8.0])
    d["t_shape"] = np.array([3, 4], np.int32)
    d["t_in_r"], d["t_in_c"], d["t_in_v"] = r3, c3, v3
    d["t_out_r"], d["t_out_c"], d["t_out_v"], ret = ref_dedup(r3, c3, v3, 3, 4)
    d["t_ret"] = np.array([ret], np.int32)
    out = os.path.join(HERE, "golden_dedup_v1.npz")
    np.savez_compressed(out, **d)
    print("wrote", out, {k: v.shape for k, v in d.items()})
    print("triples ->", list(zip(d["t_out_r"], d["t_out_c"], d["t_out_v"])), "ret", d["t_ret"])


if __name__ == "__main__":
    main()
