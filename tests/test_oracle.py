"""Pins the checker: oracle/oracle.c against the golden vectors produced by the unmodified
reference (tests/golden/make_golden.py) and, where oracle/_ref/libref.so is present, against
the reference itself on fresh seeded inputs.  CPU only."""
import numpy as np
import pytest

import oracle_lib as ol

CASES = ["t2", "mtx4a", "mtx4b", "rmat8", "stencil543", "planted200"]
RMCL_CASES = [c for c in CASES if c != "mtx4b"]


def get(g, name):
    r, c = g[name + "_shape"]
    return ol.M(g[name + "_I"], g[name + "_J"], g[name + "_V"], int(r), int(c))


def bitwise(a, b, what):
    assert np.array_equal(a.I, b.I), what + " rowPtr"
    assert np.array_equal(a.J, b.J), what + " colInd"
    assert np.array_equal(a.V.view(np.int64), b.V.view(np.int64)), what + " values (bitwise)"


@pytest.mark.parametrize("name", CASES)
def test_spgemm_matches_golden_bitwise(golden, name):
    A = get(golden, name + "_A")
    got = ol.o_spgemm(A, A)
    bitwise(got, get(golden, name + "_AA_raw"), name + " raw (first-touch order)")
    bitwise(ol.o_make_ordered(got), get(golden, name + "_AA_sorted"), name + " sorted")


@pytest.mark.parametrize("name", CASES)
def test_flops_and_partition_match_golden(golden, name):
    A = get(golden, name + "_A")
    pre = ol.o_flops_prefix(A, A)
    assert np.array_equal(pre, golden[name + "_flops"])
    for parts in (2, 3, 8):
        assert np.array_equal(ol.o_equal_partition64(pre, parts), golden[f"{name}_ends{parts}"])


@pytest.mark.parametrize("name", RMCL_CASES)
def test_rmcl_step_and_loop_match_golden_bitwise(golden, name):
    A = get(golden, name + "_A")
    bitwise(ol.o_rmcl_onestep(A, A), get(golden, name + "_step_raw"), name + " one step")
    loop, iters, hist = ol.o_rmcl_iter(A, A, 6)
    assert iters == 6 and len(hist) == 6
    bitwise(loop, get(golden, name + "_iter6_raw"), name + " 6 iterations")
    bitwise(ol.o_make_ordered(loop), get(golden, name + "_iter6_sorted"), name + " 6 iterations sorted")


@pytest.mark.parametrize("name", CASES)
def test_pcsr_matches_golden(golden, name):
    A = get(golden, name + "_A")
    bp, rp, J, V = ol.o_pcsr_split(A, 2)
    assert np.array_equal(bp, golden[name + "_pcsr2_bp"])
    assert np.array_equal(rp, golden[name + "_pcsr2_rp"])
    assert np.array_equal(J, golden[name + "_pcsr2_J"])
    assert np.array_equal(V, golden[name + "_pcsr2_V"])


def test_threshold_and_inflation_known_answers(golden):
    o = ol.oracle()
    got = np.array([o.oracle_compute_threshold(a, m) for a, m in zip(golden["thr_avg"], golden["thr_max"])])
    assert np.array_equal(got.view(np.int64), golden["thr_out"].view(np.int64))
    # tests/util_test.cc:21-28: arrayInflationR2({4,3,-2,0}) = {16,9,4,0}
    cols, vals = np.arange(4, dtype=np.int32), golden["infl_in"].copy()
    assert list(vals * vals) == list(golden["infl_out"])


def test_rmcl_init_t2(golden):
    m = ol.o_rmcl_init(golden["t2_edges_r"], golden["t2_edges_c"], 3)
    bitwise(m, get(golden, "t2_A"), "rmclInit(t2.snap)")


def test_row_epilogue_edge_cases():
    # empty row: 0/0 average -> threshold collapses to max = 0, nothing kept (util.cc:4-9)
    c, v = ol.o_row_epilogue([], [])
    assert len(c) == 0
    # single entry is always kept and normalises to 1
    c, v = ol.o_row_epilogue([7], [0.3])
    assert list(c) == [7] and v[0] == 1.0
    # uniform row keeps everything (threshold = 0.9 * avg)
    c, v = ol.o_row_epilogue([3, 1, 2], [0.2, 0.2, 0.2])
    assert list(c) == [3, 1, 2] and np.allclose(v, 1 / 3)
    # small entries are pruned, order preserved: squares .25,.0001,.25 -> thresh 0.125
    c, v = ol.o_row_epilogue([5, 9, 1], [0.5, 0.01, 0.5])
    assert list(c) == [5, 1] and list(v) == [0.5, 0.5]
    # max far above the average drives the threshold negative -> floor 1e-7 keeps everything
    c, v = ol.o_row_epilogue([5, 9, 1, 4], [0.01, 0.9, 0.02, 0.5])
    assert list(c) == [5, 9, 1, 4]


def test_chaos_and_argmax_definitions():
    m = ol.M([0, 2, 3, 3], [0, 2, 1], [0.5, 0.5, 1.0], 3, 3)
    assert ol.o_chaos(m) == 0.5 - 0.5  # row0: .5 - (.25+.25) = 0; row1: 1 - 1 = 0
    assert list(ol.o_row_argmax(m)) == [0, 1, -1]  # tie -> smallest column; empty -> -1


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref/libref.so not present")
@pytest.mark.parametrize("seed,scale,ef,sym", [(1, 9, 8, True), (2, 10, 4, False), (3, 7, 16, True)])
def test_oracle_equals_reference_on_fresh_inputs(smf, seed, scale, ef, sym):
    A = ol.from_csr(smf.synth_rmat(scale, ef, seed, sym))
    bitwise(ol.o_spgemm(A, A), ol.r_spgemm(A, A, 3), "spgemm vs flops_omp_CSR_SpMM")
    bitwise(ol.o_rmcl_onestep(A, A), ol.r_rmcl_onestep(A, A, 2), "step vs static_omp_CSR_RMCL_OneStep")
    got, _, _ = ol.o_rmcl_iter(A, A, 4)
    bitwise(got, ol.r_rmcl_iter(A, A, 4, 4), "loop vs mtRmclIter(SOMP)")
    pre = ol.o_flops_prefix(A, A)
    assert np.array_equal(pre, ol.r_flops_prefix(A, A))
    for parts in (2, 5, 8):
        assert np.array_equal(ol.o_equal_partition64(pre, parts), ol.r_equal_partition64(pre, parts))


def test_topk_rule_of_the_checker():
    """Opt-in top-k (not in the reference): of the entries that pass the threshold the k largest
    stay, equal values ranked by ascending column; k = 0 leaves the reference's rule untouched."""
    cols = np.array([7, 3, 9, 1, 5, 8], dtype=np.int32)
    vals = np.sqrt(np.array([0.30, 0.20, 0.20, 0.20, 0.05, 0.05]))
    base = ol.o_row_epilogue(cols.copy(), vals.copy())
    try:
        ol.o_set_topk(3)
        c3, v3 = ol.o_row_epilogue(cols.copy(), vals.copy())
        # 0.30 and two of the three 0.20s: the ones in columns 1 and 3 (ascending column), storage order kept
        assert list(c3) == [7, 3, 1] and np.allclose(v3, np.array([0.3, 0.2, 0.2]) / 0.7)
        ol.o_set_topk(1)
        c1, v1 = ol.o_row_epilogue(cols.copy(), vals.copy())
        assert list(c1) == [7] and v1[0] == 1.0
        ol.o_set_topk(100)
        c9, v9 = ol.o_row_epilogue(cols.copy(), vals.copy())
        assert list(c9) == list(base[0]) and np.array_equal(v9, base[1])
    finally:
        ol.o_set_topk(0)
