"""Per-bin timing of the first rMCL iterations on the C4 graph (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import sparse_matrix_with_flops_b200 as smf
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
smf.init(0)
A = smf.synth_planted(n, k, 16, 2, 12345)
dG = A.toGpuCSR(); dT = A.toGpuCSR()
for it in range(int(sys.argv[3]) if len(sys.argv) > 3 else 3):
    dN, ch, st = smf.gpuRmclOneStep(dG, dT, want_stats=True)
    print("iter %d: total %.1f ms flops %.1f sym %.1f num %.1f other %.1f | tiles %d products %.3g unpruned %.3g kept %.3g parts %d" % (
        it, st["ms_total"], st["ms_flops"], st["ms_symbolic"], st["ms_numeric"], st["ms_other"], st["row_tiles"],
        st["products"], st["nnz_unpruned"], st["nnz_out"], st["part_count"]))
    print("   sym bins rows", st["sym_bin_rows"][:6], "ms", [round(x, 1) for x in st["ms_sym_bin"][:6]])
    print("   num bins rows", st["bins_rows"][:7], "ms", [round(x, 1) for x in st["ms_num_bin"][:7]], flush=True)
    dT.deviceDispose(); dT = dN
