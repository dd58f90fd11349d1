"""Parity tests proper: the CUDA path, called through the C-ABI, against the checker
(oracle/oracle.c, itself pinned to the reference) on the same seeded inputs and against the
committed golden vectors.  Bar (BASELINE.json north_star): rowPtr exact, per-row sorted colInd
exact, values within 1e-12 relative; cluster labels exact."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu
TOL = 1e-12  # relative, per entry (north_star)


def M_of(c):
    return ol.from_csr(c)


def gpu_spgemm(smf, A, B):
    """Through the host-buffer entry point b200_spgemm_csr (mirror of CSR::flops_spmm)."""
    C = A.flops_spmm(B)
    return M_of(C)


def want_spgemm(A, B):
    return ol.o_make_ordered(ol.o_spgemm(M_of(A), M_of(B)))


def random_csr(smf, rows, cols, density, seed, sort=True, empty_rows=0.0):
    rng = np.random.default_rng(seed)
    cnt = rng.binomial(cols, density, size=rows)
    cnt[rng.random(rows) < empty_rows] = 0
    rowPtr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
    colInd = np.empty(rowPtr[-1], dtype=np.int32)
    for i in range(rows):
        c = rng.choice(cols, size=cnt[i], replace=False)
        colInd[rowPtr[i]:rowPtr[i + 1]] = np.sort(c) if sort else c
    vals = rng.random(rowPtr[-1]) + 0.01
    return smf.CSR(vals, colInd, rowPtr, rows, cols)


GOLDEN = ["t2", "mtx4a", "mtx4b", "rmat8", "stencil543", "planted200"]


@pytest.mark.parametrize("name", GOLDEN)
def test_spgemm_golden(gpu, golden, name):
    r, c = golden[name + "_A_shape"]
    A = gpu.CSR(golden[name + "_A_V"], golden[name + "_A_J"], golden[name + "_A_I"], int(r), int(c))
    got = gpu_spgemm(gpu, A, A)
    want = ol.M(golden[name + "_AA_sorted_I"], golden[name + "_AA_sorted_J"], golden[name + "_AA_sorted_V"], int(r), int(c))
    ol.assert_same(got, want, TOL, name)


@pytest.mark.parametrize("name", [g for g in GOLDEN if g != "mtx4b"])
def test_rmcl_golden(gpu, golden, name):
    r, c = golden[name + "_A_shape"]
    A = gpu.CSR(golden[name + "_A_V"], golden[name + "_A_J"], golden[name + "_A_I"], int(r), int(c))
    step = M_of(A.staticOmpRmclOneStep(A).makeOrdered())
    ol.assert_same(step, ol.M(golden[name + "_step_sorted_I"], golden[name + "_step_sorted_J"],
                              golden[name + "_step_sorted_V"], int(r), int(c)), TOL, name + " step")
    Mt, iters, hist = gpu.gpuRmclIter(6, A, A)
    assert iters == 6
    ol.assert_same(M_of(Mt), ol.M(golden[name + "_iter6_sorted_I"], golden[name + "_iter6_sorted_J"],
                                  golden[name + "_iter6_sorted_V"], int(r), int(c)), TOL, name + " 6 iters")


SYNTH = [
    ("rmat10_sym", lambda s: s.synth_rmat(10, 16, 12345, True)),       # hub rows -> bitmap bin
    ("rmat12_dir", lambda s: s.synth_rmat(12, 16, 12345, False)),
    ("rmat13_sym_ef4", lambda s: s.synth_rmat(13, 4, 7, True)),
    ("stencil_9x8x7", lambda s: s.synth_stencil27(9, 8, 7)),           # all rows in warp bins
    ("stencil_1x1x40", lambda s: s.synth_stencil27(1, 1, 40)),
    ("planted_3000", lambda s: s.synth_planted(3000, 10, 16, 2, 12345)),
]


@pytest.mark.parametrize("name,make", SYNTH)
def test_spgemm_synthetic(gpu, name, make):
    A = make(gpu)
    got = gpu_spgemm(gpu, A, A)
    ol.assert_same(got, want_spgemm(A, A), TOL, name)


@pytest.mark.parametrize("name,make", SYNTH)
def test_rmcl_synthetic(gpu, name, make):
    A = make(gpu)
    want1 = ol.o_make_ordered(ol.o_rmcl_onestep(M_of(A), M_of(A)))
    step = A.staticOmpRmclOneStep(A)
    step.makeOrdered()
    ol.assert_same(M_of(step), want1, TOL, name + " one step")
    assert abs(step.chaos - ol.o_chaos(want1)) <= 1e-12
    want, it_w, hist_w = ol.o_rmcl_iter(M_of(A), M_of(A), 8)
    ol.o_make_ordered(want)
    Mt, iters, hist = gpu.gpuRmclIter(8, A, A)
    assert iters == it_w == 8
    ol.assert_same(M_of(Mt), want, TOL, name + " 8 iterations")
    assert np.allclose(hist, hist_w, rtol=0, atol=1e-12)
    # cluster labels: argmax per row, ties -> lowest column, bit-exact
    d = Mt.toGpuCSR()
    assert np.array_equal(d.row_argmax(), ol.o_row_argmax(want))
    d.deviceDispose()


def test_small_rows_are_bit_identical(gpu):
    """Warp-per-row bins reproduce the reference's accumulation order: values bitwise equal."""
    A = gpu.synth_stencil27(10, 9, 8)
    got, want = gpu_spgemm(gpu, A, A), want_spgemm(A, A)
    assert np.array_equal(got.V.view(np.int64), want.V.view(np.int64))


@pytest.mark.parametrize("make,chain_exact", [(lambda s: s.synth_stencil27(10, 9, 8), True),
                                              (lambda s: s.synth_planted(3000, 30, 6, 1, 1), False)])
def test_rmcl_hash_bins_are_bit_identical_in_reference_order(gpu, make, chain_exact):
    """Rows of <= 256 distinct columns go through the warp hash bins, which keep the
    reference's first-touch order and run its row math sequentially: the raw (unsorted) step
    output and 8 chained iterations equal the reference's bit for bit."""
    A = make(gpu)
    raw = ol.o_rmcl_onestep(M_of(A), M_of(A))
    assert np.diff(ol.o_spgemm(M_of(A), M_of(A)).I).max() <= 256, "test input must stay in the hash bins"
    got = M_of(A.staticOmpRmclOneStep(A))
    assert np.array_equal(got.I, raw.I) and np.array_equal(got.J, raw.J)
    assert np.array_equal(got.V.view(np.int64), raw.V.view(np.int64))
    want, _, _ = ol.o_rmcl_iter(M_of(A), M_of(A), 8)
    ol.o_make_ordered(want)
    Mt, _, _ = gpu.gpuRmclIter(8, A, A)
    assert np.array_equal(Mt.colInd, want.J)
    if chain_exact:   # every row of every iterate stays in the hash bins
        assert np.array_equal(Mt.values.view(np.int64), want.V.view(np.int64))
    else:             # later iterates have rows in the bitmap bin (fp64 RED: order not fixed)
        ol.assert_same(M_of(Mt), want, TOL, "8 iterations")


def test_rmcl_to_convergence(gpu):
    A = gpu.synth_planted(2000, 8, 16, 1, 3)
    want, it_w, hist_w = ol.o_rmcl_iter(M_of(A), M_of(A), 40, eps=1e-6)
    Mt, iters, hist = gpu.gpuRmclIter(40, A, A, eps=1e-6)
    assert iters == it_w and iters < 40
    ol.assert_same(M_of(Mt), ol.o_make_ordered(want), TOL, "converged Mt")
    assert np.array_equal(ol.o_row_argmax(M_of(Mt)), ol.o_row_argmax(want))


@pytest.mark.parametrize("name,make", SYNTH)
def test_wide_fallback_paths(gpu, name, make, b200_options):
    """B200_FORCE_WIDE=1 makes the library treat B as too wide for a shared-memory bitmap: the
    optimistic / full-size warp tables (with the device-side overflow retry) and the HBM-bitmap
    kernels then handle these inputs.  Same parity bar."""
    b200_options(B200_FORCE_WIDE=1)
    A = make(gpu)
    ol.assert_same(gpu_spgemm(gpu, A, A), want_spgemm(A, A), TOL, name + " wide")
    want1 = ol.o_make_ordered(ol.o_rmcl_onestep(M_of(A), M_of(A)))
    step = A.staticOmpRmclOneStep(A)
    step.makeOrdered()
    ol.assert_same(M_of(step), want1, TOL, name + " wide rMCL step")


HEAVY = [
    ("rmat12_dir", lambda s: s.synth_rmat(12, 16, 12345, False)),
    ("rmat13_sym", lambda s: s.synth_rmat(13, 16, 3, True)),           # hub rows with > 10^3 A entries
    ("planted_dense", lambda s: s.synth_planted(3000, 3, 60, 4, 7)),   # ~1000-column rows, ~16 K products
]
# the numeric pass of the heavy rows: the part kernel (default) and the on-chip range items with
# few wide / many narrow column ranges
HEAVY_MODES = [
    ("default", {}),                                   # part kernel: global fp64 RED
    ("on_chip", {"B200_ON_CHIP": 1}),                  # range items, one warp each, A-entry order
    ("deterministic", {"B200_DETERMINISTIC": 1}),
    ("ranges4", {"B200_ON_CHIP": 1, "B200_RANGES": 4}),   # few wide ranges: items that overflow -> RED
    ("ranges7", {"B200_ON_CHIP": 1, "B200_RANGES": 7}),
    ("ranges128", {"B200_ON_CHIP": 1, "B200_RANGES": 128}),
    ("ranges2_deterministic", {"B200_DETERMINISTIC": 1, "B200_RANGES": 2}),  # overflow -> column passes
]


@pytest.mark.parametrize("mode,opts", HEAVY_MODES)
@pytest.mark.parametrize("name,make", HEAVY)
def test_heavy_row_paths(gpu, name, make, mode, opts, b200_options):
    """Rows of more than 256 output columns: every way the numeric pass can take them gives the
    reference's structure exactly and its values within 1e-12; the same for one rMCL step."""
    b200_options(**opts)
    A = make(gpu)
    dA = A.toGpuCSR()
    dC, st = gpu.gpuSpMMWrapper(dA, dA, want_stats=True)
    got = M_of(dC.toCpuCSR())
    dC.deviceDispose()
    dA.deviceDispose()
    assert st["bins_rows"][5] > 0, "test input must have rows in the bitmap bin"
    assert (st["range_items"] > 0) == (mode != "default"), st["range_items"]
    ol.assert_same(got, want_spgemm(A, A), TOL, name + " " + mode)
    want1 = ol.o_make_ordered(ol.o_rmcl_onestep(M_of(A), M_of(A)))
    step = A.staticOmpRmclOneStep(A)
    step.makeOrdered()
    ol.assert_same(M_of(step), want1, TOL, name + " " + mode + " rMCL step")


@pytest.mark.parametrize("ranges", [128, 2])
@pytest.mark.parametrize("name,make", HEAVY)
def test_on_chip_items_are_bit_identical(gpu, name, make, ranges, b200_options):
    """An ordered on-chip item commits the B-row segments in A-entry order with separately rounded
    multiply and add (indexProcessCRowI's order): with B200_DETERMINISTIC every item runs that
    way and the whole product equals the reference's bit for bit — heavy rows included — and is
    the same on every run."""
    b200_options(B200_DETERMINISTIC=1, B200_RANGES=ranges)   # 2 ranges: they overflow a warp's pool
    A = make(gpu)
    got, want = gpu_spgemm(gpu, A, A), want_spgemm(A, A)
    assert np.array_equal(got.I, want.I) and np.array_equal(got.J, want.J)
    assert np.array_equal(got.V.view(np.int64), want.V.view(np.int64))
    again = gpu_spgemm(gpu, A, A)
    assert np.array_equal(again.V.view(np.int64), got.V.view(np.int64))


@pytest.mark.parametrize("name,make", [SYNTH[0], SYNTH[2], SYNTH[5], HEAVY[2]])
def test_rmcl_through_a_bounded_arena(gpu, name, make, b200_options):
    """The rMCL step never needs the unpruned product of the whole step in memory
    (static_omp_csr_kernel.cc:241-242 allocates it): with a small arena the rows run as
    consecutive tiles, each pruned on its own.  Same iterates as one pass and as the checker."""
    A = make(gpu)
    dG = A.toGpuCSR()
    one, ch1, st1 = gpu.gpuRmclOneStep(dG, dG, want_stats=True)
    assert st1["row_tiles"] <= 1
    b200_options(B200_ARENA_ENTRIES=max(1000, st1["products"] // 7))
    til, ch2, st2 = gpu.gpuRmclOneStep(dG, dG, want_stats=True)
    assert st2["row_tiles"] >= 5, st2["row_tiles"]
    assert st2["products"] == st1["products"] and st2["nnz_unpruned"] == st1["nnz_unpruned"]
    a, b = one.toCpuCSR(), til.toCpuCSR()
    one.deviceDispose(); til.deviceDispose(); dG.deviceDispose()
    want1 = ol.o_make_ordered(ol.o_rmcl_onestep(M_of(A), M_of(A)))
    assert np.array_equal(a.rowPtr, b.rowPtr)
    ol.assert_same(M_of(b.makeOrdered()), want1, TOL, name + " tiled step")
    assert abs(ch2 - ol.o_chaos(want1)) <= 1e-12 and abs(ch1 - ch2) <= 1e-12
    # the loop, every iteration tiled
    want, it_w, hist_w = ol.o_rmcl_iter(M_of(A), M_of(A), 6)
    ol.o_make_ordered(want)
    Mt, iters, hist = gpu.gpuRmclIter(6, A, A)
    ol.assert_same(M_of(Mt), want, TOL, name + " tiled loop")
    assert np.allclose(hist, hist_w, rtol=0, atol=1e-12)


def _check_row_blocks(gpu, A, nblocks, picks, what):
    """C = A*A computed ONCE on the GPU for the whole matrix (so the heavy-row kernels run with the
    geometry of the full problem), then the chosen flops-balanced row blocks against the checker
    on the same rows: rowPtr and sorted colInd exact, values within 1e-12 relative."""
    M = M_of(A)
    dA = A.toGpuCSR()
    pre = gpu.flops_prefix(dA, dA)
    ends = gpu.arrayEqualPartition64(pre, nblocks)
    dC, st = gpu.gpuSpMMWrapper(dA, dA, want_stats=True)
    assert st["products"] == int(pre[-1])
    worst = 0.0
    for b in picks:
        lo, hi = int(ends[b]), int(ends[b + 1])
        got = M_of(dC.toCpuCSR(lo, hi))
        blk = ol.M(M.I[lo:hi + 1], M.J, M.V, hi - lo, M.cols)
        want = ol.o_make_ordered(ol.o_spgemm(blk, M))
        assert np.array_equal(got.I, want.I), "%s block %d: rowPtr" % (what, b)
        assert np.array_equal(got.J, want.J), "%s block %d: colInd" % (what, b)
        rel = float(np.max(np.abs(got.V - want.V) / np.abs(want.V))) if want.nnz else 0.0
        assert rel <= TOL, "%s block %d: relative value error %.3e" % (what, b, rel)
        worst = max(worst, rel)
    dC.deviceDispose()
    dA.deviceDispose()
    return st, worst


def test_full_scale_rmat18_every_row_block(gpu):
    """R-MAT scale 18 (products 2.9e9, nnz(C) 1.28e9): the whole product, every row block."""
    A = gpu.synth_rmat(18, 16, 12345, False)
    st, worst = _check_row_blocks(gpu, A, 3, [0, 1, 2], "rmat18")
    assert st["nnz_out"] > 1_000_000_000 and st["part_kernel"] == 1


def test_headline_config_rmat20_hub_middle_tail_blocks(gpu):
    """BASELINE.json configs[1] at full size (R-MAT scale 20: products 2.09e10, nnz(C) 9.7e9 =
    116 GB in HBM, computed exactly as bench.py does): the hub block, a middle block and the
    tail block of 21 equal-products row blocks against the checker."""
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 160e9:
        pytest.skip("needs a 180 GB device")
    A = gpu.synth_rmat(20, 16, 12345, False)
    st, worst = _check_row_blocks(gpu, A, 21, [0, 10, 20], "rmat20")
    assert st["products"] == 20933223103 and st["nnz_out"] == 9700343247


def _squares_around(target):
    """doubles v whose correctly rounded squares are just below, at (when one exists) and just
    above `target`"""
    out = {}
    v0 = np.sqrt(target)
    v = np.nextafter(v0, 0)
    for _ in range(8):
        v = np.nextafter(v, 0)
    for _ in range(40):
        sq = v * v
        key = "below" if sq < target else ("at" if sq == target else "above")
        if key == "below" or key not in out:
            out[key] = v          # the largest v below, the first v at / above
        v = np.nextafter(v, 1)
    return out


@pytest.mark.parametrize("cnt", [200, 1000, 5000])
def test_rmcl_rows_with_entries_at_the_pruning_threshold(gpu, cnt):
    """Keep / drop decisions on the threshold itself (arrayThreshPruneNormalize keeps v >= t,
    nlibs/tools/util.cc:47-69) for hash-bin rows (200 columns) and bitmap-bin rows (1000, 5000):
    every row's threshold is the 1e-7 floor (computeThreshold, util.cc:4-9), and its inflated
    entries are the doubles just below, at and just above 1e-7.  Each product entry is
    0.5 v + 0.5 v — exact in any order — so the structure must equal the checker's exactly."""
    sq = _squares_around(1.0e-7)
    vals = [sq["below"], sq["above"]] + ([sq["at"]] if "at" in sq else [])
    assert sq["below"] ** 2 < 1e-7 <= sq["above"] ** 2
    rng = np.random.default_rng(cnt)
    rows_out = 40
    n = 2 * rows_out + cnt + 50
    # Mt rows 2k, 2k+1 (k < rows_out) hold the same `cnt` columns and values; Mgt row k = 0.5, 0.5 on them
    tI, tJ, tV = [0], [], []
    for k in range(rows_out):
        cols = np.sort(rng.choice(n, cnt, replace=False)).astype(np.int32)
        v = rng.choice(vals, cnt)
        if k % 4 == 1:
            v[:] = vals[k % len(vals)]           # a row of equal entries
        for _ in range(2):
            tJ.append(cols); tV.append(v); tI.append(tI[-1] + cnt)
    for _ in range(n - 2 * rows_out):
        tI.append(tI[-1])
    Mt = gpu.CSR(np.concatenate(tV), np.concatenate(tJ), np.array(tI, dtype=np.int32), n, n)
    gI = [0]
    gJ, gV = [], []
    for k in range(n):
        if k < rows_out:
            gJ += [2 * k, 2 * k + 1]; gV += [0.5, 0.5]
        gI.append(len(gJ))
    Mg = gpu.CSR(np.array(gV), np.array(gJ, dtype=np.int32), np.array(gI, dtype=np.int32), n, n)
    want = ol.o_make_ordered(ol.o_rmcl_onestep(M_of(Mg), M_of(Mt)))
    got = Mg.staticOmpRmclOneStep(Mt)
    got.makeOrdered()
    kept = np.diff(want.I)[:rows_out]
    assert kept.min() < cnt and kept.max() > 0, "the rows must actually prune something"
    ol.assert_same(M_of(got), want, TOL, "threshold rows, %d columns" % cnt)


def test_sharded_loop_two_ranks_nccl():
    """b200_rmcl_iter_sharded on 2 GPUs (NCCL all-gather of the pruned row blocks + chaos) against
    the checker, one process per GPU.  Needs two devices: skipped on a single-GPU box."""
    import os, subprocess, sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for kind, size, iters in (("planted", "20000", "6"), ("rmat", "12", "5")):
        out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                              "--master-addr", "127.0.0.1", "--master-port", "29731",
                              os.path.join(root, "tools", "run_sharded_rmcl.py"), kind, size, iters],
                             capture_output=True, text=True, timeout=600)
        assert out.returncode == 0 and "parity with the checker: OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]


def test_cost_prefix_is_products_plus_a_charge_per_heavy_row(gpu):
    """b200_cost_prefix (the footprint-style cost the multi-GPU cuts balance): products, plus the
    charge for every row with more than 512 products; charge 0 is the flops prefix itself."""
    A = gpu.synth_rmat(11, 16, 3, True)
    dA = A.toGpuCSR()
    flops = gpu.flops_prefix(dA, dA)
    assert np.array_equal(gpu.cost_prefix(dA, dA, 0), flops)
    per = np.diff(flops)
    want = np.concatenate([[0], np.cumsum(per + 1000 * (per > 512))])
    assert np.array_equal(gpu.cost_prefix(dA, dA, 1000), want) and (per > 512).sum() > 0
    dA.deviceDispose()


def test_device_partition_matches_host_partition(gpu):
    """The sharded loop finds its cut points on the device; same arithmetic as
    arrayEqualPartition64 (util.cc:123-135): a 1-rank and the tiled path exercise it; here the
    flops prefix from the device gives the same cuts through the host routine for 1..9 parts."""
    A = gpu.synth_rmat(12, 16, 5, True)
    dA = A.toGpuCSR()
    prefix = gpu.flops_prefix(dA, dA)
    dA.deviceDispose()
    assert np.array_equal(prefix, ol.o_flops_prefix(M_of(A), M_of(A)))
    for parts in range(1, 10):
        ends = gpu.arrayEqualPartition64(prefix, parts)
        assert np.array_equal(ends, ol.o_equal_partition64(prefix, parts))


def test_sharded_loop_single_rank(gpu):
    """b200_rmcl_iter_sharded with one rank (no collective): same iterates as the reference loop.
    The 2-rank NCCL path is exercised by tools/run_sharded_rmcl.py under torchrun."""
    A = gpu.synth_planted(4000, 20, 10, 2, 9)
    want, it_w, hist_w = ol.o_rmcl_iter(M_of(A), M_of(A), 6)
    ol.o_make_ordered(want)
    dG, dT = A.toGpuCSR(), A.toGpuCSR()
    iters, hist, ms = gpu.gpuRmclIterSharded(6, dG, dT)
    got = dT.toCpuCSR()
    dG.deviceDispose()
    dT.deviceDispose()
    assert iters == 6 and len(ms) == 6 and np.all(ms > 0)
    ol.assert_same(M_of(got), want, TOL, "sharded loop, 1 rank")
    assert np.allclose(hist, hist_w, rtol=0, atol=1e-12)


def test_rectangular_unsorted_and_empty_rows(gpu):
    A = random_csr(gpu, 300, 200, 0.03, 1, empty_rows=0.2)
    B = random_csr(gpu, 200, 500, 0.05, 2, sort=False, empty_rows=0.1)
    ol.assert_same(gpu_spgemm(gpu, A, B), want_spgemm(A, B), TOL, "rect")


def test_dense_rows_each_bin(gpu):
    """Rows sized to land in every symbolic / numeric bin, incl. the hash tables at capacity."""
    rng = np.random.default_rng(5)
    n = 5000
    B = random_csr(gpu, n, n, 0.02, 11)           # ~100 per row
    sizes = [0, 1, 2, 3, 31, 32, 33, 64, 65, 127, 128, 129, 600, 1500, 3000]
    rowPtr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    cols = np.concatenate([np.sort(rng.choice(n, size=s, replace=False)) for s in sizes]).astype(np.int32)
    A = gpu.CSR(rng.random(len(cols)) + 0.1, cols, rowPtr, len(sizes), n)
    ol.assert_same(gpu_spgemm(gpu, A, B), want_spgemm(A, B), TOL, "bins")


def test_exact_capacity_rows(gpu):
    """nnz(C row) exactly 64 / 256 / 1024 / 1025 (numeric bin boundaries)."""
    n = 4096
    targets = [63, 64, 65, 255, 256, 257, 1023, 1024, 1025, 4096]
    # B = identity; A row i has `t` entries -> C row has exactly t entries
    B = gpu.CSR(np.ones(n), np.arange(n, dtype=np.int32), np.arange(n + 1, dtype=np.int32), n, n)
    rng = np.random.default_rng(9)
    rowPtr = np.concatenate([[0], np.cumsum(targets)]).astype(np.int32)
    cols = np.concatenate([np.sort(rng.choice(n, size=t, replace=False)) for t in targets]).astype(np.int32)
    A = gpu.CSR(rng.random(len(cols)) + 0.5, cols, rowPtr, len(targets), n)
    got = gpu_spgemm(gpu, A, B)
    ol.assert_same(got, want_spgemm(A, B), TOL, "capacity")
    assert list(np.diff(got.I)) == targets


def test_empty_operands(gpu):
    Z = gpu.CSR(np.zeros(0), np.zeros(0, dtype=np.int32), np.zeros(6, dtype=np.int32), 5, 5)
    C = Z.flops_spmm(Z)
    assert C.nnz == 0 and list(C.rowPtr) == [0] * 6
    A = gpu.synth_rmat(5, 4, 1, True)
    C = A.flops_spmm(gpu.CSR(np.zeros(0), np.zeros(0, dtype=np.int32), np.zeros(A.rows + 1, dtype=np.int32), A.rows, 7))
    assert C.nnz == 0 and C.cols == 7


@pytest.mark.parametrize("n", [1_900_000, 2_500_000])
def test_wide_matrix_uses_hbm_bitmap(gpu, n):
    """More columns than one shared-memory bitmap holds: n = 1.9M runs the part-wise symbolic and
    numeric kernels (4 column parts), n = 2.5M the HBM-bitmap variant."""
    rng = np.random.default_rng(3)
    k = 400
    # B: k rows, each ~60 columns spread over n
    cnt = np.full(k, 60)
    rowPtr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
    cols = np.concatenate([np.sort(rng.choice(n, size=60, replace=False)) for _ in range(k)]).astype(np.int32)
    B = gpu.CSR(rng.random(len(cols)) + 0.1, cols, rowPtr, k, n)
    # A: 6 rows; two of them reference many B rows (P > 8192 -> bitmap bins)
    sizes = [3, 350, 40, 400, 0, 200]
    arp = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    acol = np.concatenate([np.sort(rng.choice(k, size=s, replace=False)) for s in sizes]).astype(np.int32)
    A = gpu.CSR(rng.random(len(acol)) + 0.1, acol, arp, len(sizes), k)
    ol.assert_same(gpu_spgemm(gpu, A, B), want_spgemm(A, B), TOL, "wide")


def test_unsorted_b_with_two_column_parts(gpu):
    """B has 600 K columns (two column parts) and UNSORTED rows, as an rMCL iterate in first-touch
    order has: the part-wise kernels then work on a column-sorted copy of B.  SpGEMM and one
    rMCL step against the checker."""
    n, k = 600_000, 2000
    rng = np.random.default_rng(21)
    per = 300
    cols = np.concatenate([rng.choice(n, size=per, replace=False) for _ in range(k)]).astype(np.int32)  # unsorted
    rowPtr = (np.arange(k + 1) * per).astype(np.int32)
    vals = rng.random(k * per) + 0.05
    vals /= np.repeat(np.add.reduceat(vals, rowPtr[:-1]), per)        # row-stochastic, like Mt
    B = gpu.CSR(vals, cols, rowPtr, k, n)
    sizes = [1, 3, 2, 40, 7, 200, 15, 1000, 60, 2000, 5, 400]
    arp = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    acol = np.concatenate([np.sort(rng.choice(k, size=sz, replace=False)) for sz in sizes]).astype(np.int32)
    aval = rng.random(len(acol)) + 0.1
    A = gpu.CSR(aval, acol, arp, len(sizes), k)
    ol.assert_same(gpu_spgemm(gpu, A, B), want_spgemm(A, B), TOL, "unsorted B, 2 parts")
    want1 = ol.o_make_ordered(ol.o_rmcl_onestep(M_of(A), M_of(B)))
    step = A.staticOmpRmclOneStep(B)
    step.makeOrdered()
    ol.assert_same(M_of(step), want1, TOL, "unsorted B, 2 parts, rMCL step")
    assert abs(step.chaos - ol.o_chaos(want1)) <= 1e-12


def test_row_blocks_concatenate_to_the_whole(gpu):
    """SURVEY.md §7 hard part 1: row-block calls with flops-balanced cut points."""
    A = gpu.synth_rmat(11, 8, 5, True)
    dA = A.toGpuCSR()
    pre = gpu.flops_prefix(dA, dA)
    assert np.array_equal(pre, ol.o_flops_prefix(M_of(A), M_of(A)))
    ends = gpu.arrayEqualPartition64(pre, 3)
    assert np.array_equal(ends, ol.o_equal_partition64(pre, 3))
    whole = want_spgemm(A, A)
    for b in range(3):
        lo, hi = int(ends[b]), int(ends[b + 1])
        dC = gpu.gpuSpMMWrapper(dA, dA, lo, hi)
        got = M_of(dC.toCpuCSR())
        dC.deviceDispose()
        assert np.array_equal(got.I, whole.I[lo:hi + 1] - whole.I[lo])
        s, e = whole.I[lo], whole.I[hi]
        assert np.array_equal(got.J, whole.J[s:e])
        assert np.all(np.abs(got.V - whole.V[s:e]) <= TOL * np.abs(whole.V[s:e]))
    dA.deviceDispose()


def test_properties_at_scale(gpu):
    """Size-independent properties on a graph too large for the checker to be quick:
    (1) every row of an rMCL step sums to 1 and is sorted; (2) SpGEMM row sums obey
    sum_j C[i,j] = sum_k A[i,k] * rowsum(B[k,:]) (linearity); (3) nnz(C) <= products."""
    A = gpu.synth_rmat(15, 16, 99, True)
    dA = A.toGpuCSR()
    dC, st = gpu.gpuSpMMWrapper(dA, dA, want_stats=True)
    assert st["nnz_out"] <= st["products"]
    C = dC.toCpuCSR()
    dC.deviceDispose()
    rowid = np.repeat(np.arange(C.rows), np.diff(C.rowPtr))
    assert np.all(np.diff(C.colInd)[np.diff(rowid) == 0] > 0), "ascending columns"
    rs_B = np.add.reduceat(A.values, A.rowPtr[:-1])
    want_rs = np.add.reduceat(A.values * rs_B[A.colInd], A.rowPtr[:-1])
    got_rs = np.zeros(C.rows)
    np.add.at(got_rs, rowid, C.values)
    assert np.allclose(got_rs, want_rs, rtol=1e-12, atol=0)
    dM, chaos = gpu.gpuRmclOneStep(dA, dA)
    dM.makeOrdered()
    Mt = dM.toCpuCSR()
    dM.deviceDispose()
    dA.deviceDispose()
    rid = np.repeat(np.arange(Mt.rows), np.diff(Mt.rowPtr))
    sums = np.zeros(Mt.rows)
    np.add.at(sums, rid, Mt.values)
    assert np.allclose(sums, 1.0, rtol=0, atol=1e-12)
    assert np.all(np.diff(Mt.colInd)[np.diff(rid) == 0] > 0)
    assert 0.0 <= chaos <= 1.0


def test_cpp_host_layer(gpu, tmp_path):
    """include/b200_nlibs.hpp (the C++ mirror of the reference's CSR / COO / PCSR / RMCL names)
    compiled with g++ and run like one of the reference's one-main tests: all lines `Same`."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "shim_test.x")
    lib = os.path.join(root, "sparse_matrix_with_flops_b200")
    subprocess.check_call(["g++", "-O1", "-std=c++11", "-I" + os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "shim_test.cc"), "-o", exe,
                           "-L" + lib, "-lb200spgemm", "-Wl,-rpath," + lib])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "Diffs" not in out.stdout, out.stdout + out.stderr
    assert out.stdout.count("Same") == 11


def test_baseline_config_c1_nrmcl_rmat16(gpu):
    """BASELINE.json configs[0]: rMCL (nrmcl, maxIters = 5) on symmetrised R-MAT scale 16, edge
    factor 16 — the reference's own CPU-runnable case — against the checker at full size:
    structure exact, values <= 1e-12 relative, chaos history equal, cluster labels exact."""
    A = gpu.synth_rmat(16, 16, 12345, True)
    want, it_w, hist_w = ol.o_rmcl_iter(M_of(A), M_of(A), 5)
    ol.o_make_ordered(want)
    Mt, iters, hist = gpu.gpuRmclIter(5, A, A)
    assert iters == it_w == 5
    ol.assert_same(M_of(Mt), want, TOL, "C1 nrmcl")
    assert np.allclose(hist, hist_w, rtol=0, atol=1e-12)
    d = Mt.toGpuCSR()
    assert np.array_equal(d.row_argmax(), ol.o_row_argmax(want))
    d.deviceDispose()


def _glue(blocks, rows, cols, gpu):
    """Row blocks (lo, hi, CSR with rowPtr[0] == 0), in order, back into one host CSR."""
    nxt, rp, J, V = 0, [np.zeros(1, dtype=np.int64)], [], []
    for lo, hi, blk in blocks:
        assert lo == nxt and hi >= lo and blk.rows == hi - lo and blk.rowPtr[0] == 0
        assert blk.nnz == blk.rowPtr[-1] == len(blk.colInd) == len(blk.values)
        rp.append(blk.rowPtr[1:].astype(np.int64) + rp[-1][-1])
        J.append(blk.colInd); V.append(blk.values)
        nxt = hi
    assert nxt == rows
    return gpu.CSR(np.concatenate(V) if V else np.zeros(0), np.concatenate(J) if J else np.zeros(0, np.int32),
                   np.concatenate(rp).astype(np.int32), rows, cols)


@pytest.mark.parametrize("case", ["rmat_directed", "rect_with_empty_rows", "one_heavy_row"])
def test_streamed_row_blocks_match_the_checker(gpu, case):
    """b200_spgemm_csr_stream: the product delivered as row blocks cut on the intermediate-product
    prefix (download of block b overlapped with the computation of block b+1) equals the
    checker's flops_omp_CSR_SpMM restatement: structure exact, values <= 1e-12 relative."""
    rng = np.random.default_rng(7)
    if case == "rmat_directed":
        A = B = gpu.synth_rmat(12, 16, 99, False)
        block = 300_000
    elif case == "rect_with_empty_rows":
        import scipy.sparse as sp
        a = sp.random(700, 300, 0.02, format="csr", random_state=3, dtype=np.float64)
        a = sp.vstack([a[:200], sp.csr_matrix((150, 300)), a[200:], sp.csr_matrix((40, 300))]).tocsr()
        b = sp.random(300, 5000, 0.01, format="csr", random_state=4, dtype=np.float64)
        A = gpu.CSR(a.data, a.indices, a.indptr, a.shape[0], a.shape[1])
        B = gpu.CSR(b.data, b.indices, b.indptr, b.shape[0], b.shape[1])
        block = 2_000
    else:
        import scipy.sparse as sp
        a = sp.random(400, 400, 0.01, format="lil", random_state=5, dtype=np.float64)
        a[123, :] = rng.random(400) + 0.1           # one row heavier than a whole block
        a = a.tocsr()
        A = B = gpu.CSR(a.data, a.indices, a.indptr, 400, 400)
        block = 500
    want = ol.o_spgemm(M_of(A), M_of(B))
    ol.o_make_ordered(want)
    got = []
    A.spmm_blocks(B, lambda lo, hi, blk: got.append((lo, hi, blk)) and 0, block)
    assert len(got) > 3
    ol.assert_same(M_of(_glue(got, A.rows, B.cols, gpu)), want, TOL, "streamed " + case)
    # default block size: a single block for these sizes, same answer
    one = []
    A.spmm_blocks(B, lambda lo, hi, blk: one.append((lo, hi, blk)) and 0)
    assert len(one) == 1
    ol.assert_same(M_of(_glue(one, A.rows, B.cols, gpu)), want, TOL, "streamed, one block " + case)


def test_streamed_row_blocks_stop_and_errors(gpu):
    """A callback that returns non-zero abandons the product with B200_ERR_CALLBACK (8); an
    exception in the Python callback is re-raised; the library stays usable."""
    A = gpu.synth_rmat(10, 8, 5, False)
    seen = []
    with pytest.raises(gpu._lib.B200Error) as ei:
        A.spmm_blocks(A, lambda lo, hi, blk: seen.append(lo) or len(seen) >= 2, 20_000)
    assert ei.value.code == 8 and len(seen) == 2
    with pytest.raises(ZeroDivisionError):
        A.spmm_blocks(A, lambda lo, hi, blk: 1 // 0, 20_000)
    want = ol.o_spgemm(M_of(A), M_of(A))
    ol.o_make_ordered(want)
    ol.assert_same(M_of(A.flops_spmm(A)), want, TOL, "after an abandoned stream")
    # a matrix without entries comes back as one empty block
    Z = gpu.CSR(np.zeros(0), np.zeros(0, np.int32), np.zeros(6, np.int32), 5, 5)
    got = []
    Z.spmm_blocks(Z, lambda lo, hi, blk: got.append((lo, hi, blk.nnz)) and 0, 10)
    assert got == [(0, 5, 0)]


def test_error_behaviour(gpu):
    """Error convention of the C ABI: a dimension mismatch (the reference asserts,
    nlibs/CSR.cc:183) and an out-of-range row block come back as error codes with a message, and
    leave the library usable."""
    import ctypes as C
    lib = gpu._lib.load()
    A = gpu.synth_rmat(6, 4, 1, True)                      # 64 x 64
    Bad = gpu.CSR(np.ones(3), np.array([0, 1, 2], dtype=np.int32), np.array([0, 1, 2, 3], dtype=np.int32), 3, 5)
    dA, dB = A.toGpuCSR(), Bad.toGpuCSR()
    h = gpu._lib.csr_t()
    rc = lib.b200_spgemm_device(dA.handle, dB.handle, C.byref(h), None)
    assert rc == 1 and b"dimension mismatch" in lib.b200_last_error()       # B200_ERR_BAD_ARG
    rc = lib.b200_spgemm_device_rows(dA.handle, dA.handle, 10, 5, C.byref(h), None)
    assert rc == 1 and b"row range" in lib.b200_last_error()
    with pytest.raises(AssertionError):
        A.flops_spmm(Bad)                                   # the Python mirror asserts like the reference
    # an array that is not a CSR is refused at upload with BAD_ARG, not turned into an
    # out-of-bounds read by the first kernel that follows an index (the context stays healthy)
    one = np.ones(3)
    for I, J in (([0, 1, 2, 3], [0, 1, 7]),      # column outside [0, cols)
                 ([0, 1, 2, 3], [0, -1, 2]),     # negative column
                 ([0, 2, 1, 3], [0, 1, 2]),      # row offsets decrease
                 ([1, 1, 2, 3], [0, 1, 2]),      # do not start at 0
                 ([0, 1, 2, 2], [0, 1, 2])):     # do not end at nnz
        bad = gpu.CSR(one, np.array(J, dtype=np.int32), np.array(I, dtype=np.int32), 3, 5, nnz=3)
        with pytest.raises(gpu._lib.B200Error) as ei:
            bad.toGpuCSR()
        assert ei.value.code == 1 and "bad CSR" in str(ei.value)
    ok = gpu.gpuSpMMWrapper(dA, dA)                         # still works afterwards
    assert ok.nnz > 0
    ok.deviceDispose(); dA.deviceDispose(); dB.deviceDispose()


def test_cpp_driver_same_and_diffs(gpu, tmp_path):
    """The C++ driver in the shape of nrmcl.cc, run as a user would: nrmcl_b200.x --input FILE
    -r B200 --maxIters 5 --expect FILE prints the reference driver's verdict line (nrmcl.cc:27-32):
    Same against the checker's Mt, Diffs against a perturbed one."""
    import os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "sparse_matrix_with_flops_b200", "nrmcl_b200.x")
    assert os.path.exists(exe), "built by sparse_matrix_with_flops_b200/csrc/Makefile"
    rng = np.random.default_rng(5)
    n, m = 600, 6000
    key = np.unique(rng.integers(0, n, m).astype(np.int64) * n + rng.integers(0, n, m))
    key = np.unique(np.concatenate([key, (key % n) * n + key // n]))          # symmetric, no repeated pairs
    er, ec = (key // n).astype(np.int32), (key % n).astype(np.int32)
    edges = tmp_path / "g.snap"
    with open(edges, "w") as f:
        f.write("# synthetic\n%d %d\n" % (n, len(er)))
        for a, b in zip(er, ec):
            f.write("%d %d\n" % (a, b))
    # readSNAPFile's default is the transposed read (row = to, col = from; nlibs/COO.cc:142-148)
    M0 = ol.o_rmcl_init(ec, er, n)
    want, _, _ = ol.o_rmcl_iter(M0, M0, 5)
    ol.o_make_ordered(want)

    def write(path, Mx, bump=None):
        with open(path, "w") as f:
            f.write("%d %d %d\n" % (Mx.rows, Mx.cols, Mx.nnz))
            rows = np.repeat(np.arange(Mx.rows), np.diff(Mx.I))
            for k, (r, c, v) in enumerate(zip(rows, Mx.J, Mx.V)):
                f.write("%d %d %.17g\n" % (r, c, v * (1 + 1e-9) if k == bump else v))
    good, bad, out = tmp_path / "want.txt", tmp_path / "bad.txt", tmp_path / "got.txt"
    write(good, want)
    write(bad, want, bump=want.nnz // 2)
    run = lambda exp: subprocess.run([exe, "--input", str(edges), "-r", "B200", "--maxIters", "5", "--expect", str(exp),
                                      "--output", str(out)], capture_output=True, text=True, timeout=300)
    r = run(good)
    assert r.returncode == 0 and r.stdout.strip().splitlines()[-1] == "Same", r.stdout[-1500:] + r.stderr[-1500:]
    assert "iters 5" in r.stdout and open(out).readline().split() == [str(n), str(n), str(want.nnz)]
    r = run(bad)
    assert r.returncode == 1 and r.stdout.strip().splitlines()[-1] == "Diffs", r.stdout[-1500:]


@pytest.mark.parametrize("k", [1, 3, 12, 100000])
@pytest.mark.parametrize("name,make", [SYNTH[0], SYNTH[5], HEAVY[2]])
def test_topk_pruning_matches_the_checker(gpu, name, make, k, b200_options):
    """Opt-in top-k (b200_set_topk): same rule as the checker's — the k largest entries above
    the threshold, ties by ascending column — in the hash bins and in the bitmap bin; a k larger
    than any row changes nothing.  Off again afterwards (parity runs use the threshold alone).
    A rule that ranks VALUES needs the same values on both sides: these graphs are full of
    exactly equal entries (1 / rowcount products), and a heavy row accumulated with fp64 RED can
    differ from the reference in the last bit, which would break such a tie the other way — so
    the heavy rows run in the bit-exact ordered mode here (B200_DETERMINISTIC)."""
    b200_options(B200_DETERMINISTIC=1)
    A = make(gpu)
    try:
        ol.o_set_topk(k)
        gpu.set_topk(k)
        want1 = ol.o_make_ordered(ol.o_rmcl_onestep(M_of(A), M_of(A)))
        step = A.staticOmpRmclOneStep(A)
        step.makeOrdered()
        assert np.diff(want1.I).max() <= k
        ol.assert_same(M_of(step), want1, TOL, "%s top-%d step" % (name, k))
        want, _, hist_w = ol.o_rmcl_iter(M_of(A), M_of(A), 5)
        ol.o_make_ordered(want)
        Mt, _, hist = gpu.gpuRmclIter(5, A, A)
        ol.assert_same(M_of(Mt), want, TOL, "%s top-%d loop" % (name, k))
    finally:
        ol.o_set_topk(0)
        gpu.set_topk(0)
    if k == 100000:
        plain = ol.o_make_ordered(ol.o_rmcl_onestep(M_of(A), M_of(A)))
        ol.assert_same(want1, plain, 0.0, "a k above every row length is the plain rule")


@pytest.mark.parametrize("c", [2, 3, 7])
def test_device_pcsr_column_striped_product(gpu, c):
    """PCSR on the device (nlibs/PCSR.cc:3-56): the stripes equal the checker's split exactly,
    and A x PCSR(B), stripe by stripe and glued back, equals the plain product
    (correctTests/pcsrTest.cc)."""
    A = random_csr(gpu, 700, 900, 0.02, 3)
    B = gpu.synth_rmat(10, 16, 4, True)          # 1024 columns, hub rows
    A = random_csr(gpu, 700, B.rows, 0.02, 3)
    dA, dB = A.toGpuCSR(), B.toGpuCSR()
    P = gpu.DevicePCSR(dB, c)
    bp, rp, J, V = ol.o_pcsr_split(M_of(B), c)
    assert P.stride == (B.cols + c - 1) // c and len(P.blocks) == c
    for b, blk in enumerate(P.blocks):
        h = blk.toCpuCSR()
        assert np.array_equal(h.rowPtr, rp[b]) and np.array_equal(h.colInd, J[bp[b]:bp[b + 1]])
        assert np.array_equal(h.values, V[bp[b]:bp[b + 1]])
    dC = P.leftMultiply(dA)
    got = M_of(dC.toCpuCSR())
    ol.assert_same(got, want_spgemm(A, B), TOL, "PCSR product, c=%d" % c)
    dC.deviceDispose(); P.dispose(); dA.deviceDispose(); dB.deviceDispose()


@pytest.mark.parametrize("name,make", [SYNTH[0], SYNTH[2], HEAVY[0], HEAVY[1]])
def test_rmcl_rows_sorted_on_chip(gpu, name, make, b200_options):
    """Mid-size rows of an rMCL step (1024 < products <= 8192) expanded, sorted by column and
    summed in A-entry order on chip (esc.cuh; the default in matrices wider than 1 M columns,
    forced here): same step and same loop as the checker, and the pre-prune entry count the
    library reports for them is exact although their arena slices are reserved by products."""
    A = make(gpu)
    dG = A.toGpuCSR()
    # plain SpGEMM: the same rows are finished on chip during the symbolic phase and moved into C
    b200_options(B200_ESC=1)
    dC, stc = gpu.gpuSpMMWrapper(dG, dG, want_stats=True)
    got = M_of(dC.toCpuCSR())
    dC.deviceDispose()
    assert stc["bins_rows"][7] + stc["bins_rows"][8] > 0
    wantC = want_spgemm(A, A)
    ol.assert_same(got, wantC, TOL, name + " SpGEMM, rows sorted on chip")
    assert stc["nnz_out"] == wantC.nnz == stc["nnz_unpruned"]
    b200_options(B200_ESC=0)
    ref_step, ch0, st0 = gpu.gpuRmclOneStep(dG, dG, want_stats=True)
    b200_options(B200_ESC=1)
    esc_step, ch1, st1 = gpu.gpuRmclOneStep(dG, dG, want_stats=True)
    assert st0["bins_rows"][7] + st0["bins_rows"][8] == 0 and st1["bins_rows"][7] + st1["bins_rows"][8] > 0
    assert st1["nnz_unpruned"] == st0["nnz_unpruned"] and st1["products"] == st0["products"]
    a, b = ref_step.toCpuCSR().makeOrdered(), esc_step.toCpuCSR().makeOrdered()
    ref_step.deviceDispose(); esc_step.deviceDispose(); dG.deviceDispose()
    want1 = ol.o_make_ordered(ol.o_rmcl_onestep(M_of(A), M_of(A)))
    ol.assert_same(M_of(b), want1, TOL, name + " step, rows sorted on chip")
    ol.assert_same(M_of(b), M_of(a), TOL, name + " both paths")
    assert abs(ch1 - ol.o_chaos(want1)) <= 1e-12
    want, _, hist_w = ol.o_rmcl_iter(M_of(A), M_of(A), 6)
    ol.o_make_ordered(want)
    Mt, _, hist = gpu.gpuRmclIter(6, A, A)
    ol.assert_same(M_of(Mt), want, TOL, name + " loop, rows sorted on chip")
    assert np.allclose(hist, hist_w, rtol=0, atol=1e-12)


@pytest.mark.parametrize("name,make", [("stencil27", lambda s: s.synth_stencil27(12, 11, 10)), SYNTH[0], SYNTH[2],
                                       HEAVY[0]])
def test_spgemm_rows_finished_in_the_symbolic_phase(gpu, name, make, b200_options):
    """Plain SpGEMM rows of 512..2048 products are tried as 128-entry numeric rows while the
    symbolic phase runs (k_num_warp_fused): those that fit are only copied into C afterwards,
    the others take the ordinary symbolic + numeric passes.  A warp row accumulates in A-entry
    order either way, so both settings give the same bits, and both match the checker."""
    A = make(gpu)
    dG = A.toGpuCSR()
    b200_options(B200_FUSE=0)
    dC0, st0 = gpu.gpuSpMMWrapper(dG, dG, want_stats=True)
    two_pass = dC0.toCpuCSR()
    dC0.deviceDispose()
    b200_options(B200_FUSE=1)
    dC1, st1 = gpu.gpuSpMMWrapper(dG, dG, want_stats=True)
    fused = dC1.toCpuCSR()
    dC1.deviceDispose()
    b200_options(B200_FUSE=2)   # the instance for operands with more than 2^31 entries (64-bit offsets)
    dC2, st2 = gpu.gpuSpMMWrapper(dG, dG, want_stats=True)
    wide = dC2.toCpuCSR()
    dC2.deviceDispose(); dG.deviceDispose()
    assert st2["bins_rows"][9] == st1["bins_rows"][9]
    assert np.array_equal(wide.rowPtr, fused.rowPtr) and np.array_equal(wide.colInd, fused.colInd)
    assert np.allclose(wide.values, fused.values, rtol=1e-12, atol=0)
    assert st0["bins_rows"][9] == 0
    if name == "stencil27":
        assert st1["bins_rows"][9] > 0 and st1["num_bin_nnzC"][9] > 0
    assert st1["nnz_out"] == st0["nnz_out"]
    assert np.array_equal(fused.rowPtr, two_pass.rowPtr) and np.array_equal(fused.colInd, two_pass.colInd)
    # rows of the warp bins are bit-identical; heavy rows (global fp64 RED) to the tolerance
    assert np.allclose(fused.values, two_pass.values, rtol=1e-12, atol=0)
    ol.assert_same(M_of(fused), want_spgemm(A, A), TOL, name + " rows finished in the symbolic phase")


def test_page_locked_block_cache(gpu):
    """b200_host_cache_pin(1): blocks released through b200_host_free are page-locked once and the
    next download lands in them with one DMA (no staging copy).  Same bytes as the default path;
    eviction, drop and a block that is not kept all unregister before the memory is freed."""
    import ctypes as C
    from sparse_matrix_with_flops_b200 import _lib
    lib = _lib.load()
    ip, dp = _lib.c_int_p, _lib.c_double_p
    A = gpu.synth_rmat(14, 16, 5, False)
    dA = A.toGpuCSR()
    dC = gpu.gpuSpMMWrapper(dA, dA)
    want = dC.toCpuCSR()
    assert want.nnz * 8 >= (64 << 20), "blocks below 64 MB are not cached"

    def download():
        IC, JC, Cv, nnz = ip(), ip(), dp(), C.c_int(0)
        _lib.check(lib.b200_csr_download_rows(dC.handle, 0, A.rows, C.byref(IC), C.byref(JC), C.byref(Cv),
                                              C.byref(nnz)))
        n = nnz.value
        got = (np.ctypeslib.as_array(IC, (A.rows + 1,)).copy(), np.ctypeslib.as_array(JC, (n,)).copy(),
               np.ctypeslib.as_array(Cv, (n,)).copy())
        for p in (IC, JC, Cv):
            lib.b200_host_free(C.cast(p, C.c_void_p))
        return got

    lib.b200_host_cache_drop()
    assert lib.b200_host_cache_pin(1) == 0
    try:
        first = download()      # fresh blocks: staged; released -> kept and page-locked
        second = download()     # the same blocks again: direct DMA
        third = download()
        for got in (first, second, third):
            assert np.array_equal(got[0], want.rowPtr) and np.array_equal(got[1], want.colInd)
            assert np.array_equal(got[2], want.values)
        b, n, l = C.c_longlong(), C.c_int(), C.c_longlong()
        lib.b200_host_cache_info(C.byref(b), C.byref(n), C.byref(l))
        assert n.value >= 1 and b.value >= want.nnz * 8
        hits, miss, direct, pinned = C.c_longlong(), C.c_longlong(), C.c_longlong(), C.c_int()
        lib.b200_host_cache_stats(C.byref(hits), C.byref(miss), C.byref(direct), C.byref(pinned))
        assert pinned.value >= 2 and direct.value >= 4, (hits.value, miss.value, direct.value, pinned.value)
    finally:
        lib.b200_host_cache_drop()
        lib.b200_host_cache_pin(0)
    plain = download()
    assert np.array_equal(plain[2], want.values)
    lib.b200_host_cache_drop()
    dC.deviceDispose(); dA.deviceDispose()
