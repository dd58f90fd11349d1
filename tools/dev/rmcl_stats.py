"""Per-phase / per-bin timing of the rMCL iterations of a bench workload (development aid).
usage: rmcl_stats.py <scale> [edge factor] [iterations]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import sparse_matrix_with_flops_b200 as smf
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ef = int(sys.argv[2]) if len(sys.argv) > 2 else 16
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
smf.init(0)
A = smf.synth_rmat(scale, ef, 12345, True)
dG = A.toGpuCSR()
for rep in range(2):
    dT = A.toGpuCSR()
    for it in range(iters):
        t0 = time.perf_counter()
        dN, ch, st = smf.gpuRmclOneStep(dG, dT, want_stats=True)
        wall = (time.perf_counter() - t0) * 1e3
        if rep == 1:
            print("iter %d: wall %.1f | total %.1f ms flops %.1f sym %.1f num %.1f other %.1f | tiles %d launches %d products %.3g unpruned %.3g kept %.3g" % (
                it, wall, st["ms_total"], st["ms_flops"], st["ms_symbolic"], st["ms_numeric"], st["ms_other"], st["row_tiles"], st["launches"],
                st["products"], st["nnz_unpruned"], st["nnz_out"]))
            print("   sym bins rows", st["sym_bin_rows"][:8], "ms", [round(x, 1) for x in st["ms_sym_bin"][:8]])
            print("   num bins rows", st["bins_rows"][:9], "ms", [round(x, 1) for x in st["ms_num_bin"][:9]], flush=True)
        dT.deviceDispose(); dT = dN
    dT.deviceDispose()
