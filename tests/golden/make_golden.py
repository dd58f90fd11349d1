"""Generate golden input/output vectors from the UNMODIFIED reference (oracle/_ref/libref.so).

Run in the build container (where /root/reference exists):
    make -C oracle ref && python tests/golden/make_golden.py
Writes tests/golden/golden_v1.npz (committed).  The tests then pin oracle/oracle.c against
these vectors everywhere, including on machines where the reference is absent.

Inputs:
  * t2      — the reference's own fixture t2.snap (3 vertices, 4 edges), read the way
              COO::readSNAPFile does by default (transposed, nlibs/COO.cc:142-148)
  * mtx4a/b — the two 4x4 MatrixMarket fixtures under mindex2-cuda/test_dir_dat/ (entries
              typed in below; mtx4a with toAbs() as nGpuSpMM.cc does)
  * rmat8, stencil543, planted200 — small synthetic graphs from this repo's generators (the
              generated CSR itself is stored, so the vectors do not depend on the generator)
  * known answers kept in the reference's (commented-out) tests: tests/util_test.cc:21-28
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as ol  # noqa: E402
import sparse_matrix_with_flops_b200 as smf  # noqa: E402


def coo_to_M(r, c, v, n):
    order = np.lexsort((c, r))
    r, c, v = np.asarray(r)[order], np.asarray(c)[order], np.asarray(v, dtype=np.float64)[order]
    I = np.zeros(n + 1, dtype=np.int32)
    np.add.at(I, r + 1, 1)
    return ol.M(np.cumsum(I).astype(np.int32), c.astype(np.int32), v, n, n)


def put(d, name, m):
    d[name + "_I"], d[name + "_J"], d[name + "_V"] = m.I, m.J, m.V
    d[name + "_shape"] = np.array([m.rows, m.cols], dtype=np.int32)


def main():
    assert ol.have_ref(), "oracle/_ref/libref.so missing: make -C oracle ref"
    d = {}
    cases = {}
    # t2.snap: "3 4" then edges from->to: (0,0) (0,1) (1,1) (2,2); transposed read
    er, ec = np.array([0, 1, 1, 2]), np.array([0, 0, 1, 2])  # row=to, col=from
    cases["t2"] = ol.r_rmcl_init(er, ec, 3)
    d["t2_edges_r"], d["t2_edges_c"] = er.astype(np.int32), ec.astype(np.int32)
    # test.mtx (1-based in the file), toAbs()
    r = np.array([1, 2, 2, 2, 3, 4]) - 1
    c = np.array([2, 1, 3, 4, 1, 4]) - 1
    v = np.abs([-.0008109343960956759, -2.6727333604411684e-6, 1.478460142719664e-6,
                -8.662701342500658e-6, 2.3063597321124548e-5, -8.147218938641685e-7])
    cases["mtx4a"] = coo_to_M(r, c, v, 4)
    r = np.array([1, 1, 2, 2, 2, 3, 4, 4]) - 1
    c = np.array([2, 4, 1, 3, 4, 1, 1, 4]) - 1
    v = [1.2, 2.0, -2.65, 1.4, -8.66, 2.30, 9.1, -8.14]
    cases["mtx4b"] = coo_to_M(r, c, v, 4)
    cases["rmat8"] = ol.from_csr(smf.synth_rmat(8, 8, 12345, True))
    cases["stencil543"] = ol.from_csr(smf.synth_stencil27(5, 4, 3))
    cases["planted200"] = ol.from_csr(smf.synth_planted(200, 4, 6, 2, 12345))
    for name, A in cases.items():
        put(d, name + "_A", A)
        # SpGEMM A*A, all four reference variants must agree bitwise; keep raw + ordered
        raw = ol.r_spgemm(A, A, 3)
        for variant in (0, 1, 2):
            other = ol.r_spgemm(A, A, variant)
            assert np.array_equal(raw.I, other.I) and np.array_equal(raw.J, other.J) and \
                np.array_equal(raw.V, other.V), (name, variant)
        put(d, name + "_AA_raw", raw)
        put(d, name + "_AA_sorted", ol.r_make_ordered(raw.copy()))
        d[name + "_flops"] = ol.r_flops_prefix(A, A)
        for parts in (2, 3, 8):
            d[f"{name}_ends{parts}"] = ol.r_equal_partition64(d[name + "_flops"], parts)
        # one rMCL step and a 6-iteration loop (SOMP == OMP == SEQ == SFOMP bitwise)
        if name != "mtx4b":  # negative values are outside rMCL's domain
            step = ol.r_rmcl_onestep(A, A, 2)
            step1 = ol.r_rmcl_onestep(A, A, 1)
            assert np.array_equal(step.V, step1.V) and np.array_equal(step.J, step1.J)
            put(d, name + "_step_raw", step)
            put(d, name + "_step_sorted", ol.r_make_ordered(step.copy()))
            loop = ol.r_rmcl_iter(A, A, 6, 4)
            loop_seq = ol.r_rmcl_iter(A, A, 6, 0)
            assert np.array_equal(loop.V, loop_seq.V) and np.array_equal(loop.J, loop_seq.J)
            put(d, name + "_iter6_raw", loop)
            put(d, name + "_iter6_sorted", ol.r_make_ordered(loop.copy()))
        bp, rp, J, V = ol.r_pcsr_split(A, 2)
        d[name + "_pcsr2_bp"], d[name + "_pcsr2_rp"], d[name + "_pcsr2_J"], d[name + "_pcsr2_V"] = bp, rp, J, V
    # computeThreshold on a grid, incl. the floor and the cap
    avg = np.array([0.0, 1e-9, 1e-4, 0.01, 0.1, 0.25, 0.5, 0.9, 1.0, 0.3, 2e-8])
    mx = np.array([0.0, 2e-9, 3e-4, 0.5, 0.1, 1.0, 0.5, 1.0, 1.0, 0.31, 3e-8])
    d["thr_avg"], d["thr_max"] = avg, mx
    d["thr_out"] = np.array([ol.ref().ref_compute_threshold(a, m) for a, m in zip(avg, mx)])
    # util_test.cc:21-28 known answer
    iv = np.array([4.0, 3.0, -2.0, 0.0])
    ov = np.zeros(4)
    ol.ref().ref_inflation_r2(ol._d(iv), 4, ol._d(ov))
    assert list(ov) == [16.0, 9.0, 4.0, 0.0]
    d["infl_in"], d["infl_out"] = iv, ov
    out = os.path.join(HERE, "golden_v1.npz")
    np.savez_compressed(out, **d)
    print("wrote", out, os.path.getsize(out), "bytes;", len(d), "arrays")


if __name__ == "__main__":
    main()
