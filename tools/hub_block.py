"""Development aid: time / profile the numeric part kernel on one row block of R-MAT scale 20
(default: the hub block, rows [0, 1538) = 1e9 products)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparse_matrix_with_flops_b200 as smf
lo, hi = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (0, 1538)
smf.init(0)
A = smf.synth_rmat(20, 16, 12345, False)
dA = A.toGpuCSR()
for rep in range(3):
    dC, st = smf.gpuSpMMWrapper(dA, dA, lo, hi, want_stats=True)
    dC.deviceDispose()
    print("rows [%d,%d) P=%d nnzC=%d total %.2f ms sym %.2f num %.2f  (num kernel %.2f ms -> %.1f Gproducts/s)" % (
        lo, hi, st["products"], st["nnz_out"], st["ms_total"], st["ms_symbolic"], st["ms_numeric"],
        st["ms_num_bin"][5], st["num_bin_products"][5] / st["ms_num_bin"][5] / 1e6), flush=True)
