"""Full-scale parity check against the checker on row blocks (run on a B200):
   python tools/validate_large.py rmat20 [nblocks_to_check]
The product C = A*A is computed once on the GPU for the whole matrix (so the bitmap-part /
team kernels run exactly as in bench.py); a few flops-balanced row blocks — the hub block, a
middle one and the tail — are downloaded and compared with oracle_spgemm on the same rows:
rowPtr and sorted colInd exact, values within 1e-12 relative."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as ol
import sparse_matrix_with_flops_b200 as smf

name = sys.argv[1] if len(sys.argv) > 1 else "rmat18"
ncheck = int(sys.argv[2]) if len(sys.argv) > 2 else 3
smf.init(0)
A = smf.synth_rmat(int(name[4:]), 16, 12345, False) if name.startswith("rmat") else smf.synth_stencil27(int(name[7:]), int(name[7:]), int(name[7:]))
M = ol.from_csr(A)
dA = A.toGpuCSR()
pre = smf.flops_prefix(dA, dA)
P = int(pre[-1])
nblk = max(ncheck, int(np.ceil(P / 1.0e9)))
ends = smf.arrayEqualPartition64(pre, nblk)
t0 = time.time()
dC, st = smf.gpuSpMMWrapper(dA, dA, want_stats=True)
print("%s: rows %d nnzA %d products %d nnzC %d, GPU %.1f ms, part kernel %d parts %d" % (
    name, A.rows, A.nnz, P, st["nnz_out"], st["ms_total"], st["part_kernel"], st["part_count"]), flush=True)
picks = sorted(set([0, nblk // 2, nblk - 1] if ncheck == 3 else np.linspace(0, nblk - 1, ncheck).astype(int).tolist()))
worst = 0.0
for b in picks:
    lo, hi = int(ends[b]), int(ends[b + 1])
    got = ol.from_csr(dC.toCpuCSR(lo, hi))
    blk = ol.M(M.I[lo:hi + 1], M.J, M.V, hi - lo, M.cols)
    t1 = time.time()
    want = ol.o_make_ordered(ol.o_spgemm(blk, M))
    assert np.array_equal(got.I, want.I), "rowPtr differs in block %d" % b
    assert np.array_equal(got.J, want.J), "colInd differs in block %d" % b
    rel = float(np.max(np.abs(got.V - want.V) / np.abs(want.V))) if want.nnz else 0.0
    worst = max(worst, rel)
    print("  block %d rows [%d,%d): nnz %d  structure exact, max rel value error %.3e (checker %.1f s)" % (
        b, lo, hi, want.nnz, rel, time.time() - t1), flush=True)
    assert rel <= 1e-12
dC.deviceDispose(); dA.deviceDispose()
print("PARITY OK on %d of %d row blocks, worst relative value error %.3e" % (len(picks), nblk, worst))
