"""Quick GPU probe: run a few SpGEMM / rMCL cases and print phase timings (development aid)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sparse_matrix_with_flops_b200 as smf

smf.init(0)
cases = sys.argv[1:] or ["rmat16s", "stencil64", "rmat18d"]
for c in cases:
    t0 = time.time()
    if c == "rmat16s": A = smf.synth_rmat(16, 16, 12345, True)
    elif c == "rmat18d": A = smf.synth_rmat(18, 16, 12345, False)
    elif c == "rmat20d": A = smf.synth_rmat(20, 16, 12345, False)
    elif c == "stencil64": A = smf.synth_stencil27(64, 64, 64)
    elif c == "stencil128": A = smf.synth_stencil27(128, 128, 128)
    elif c == "planted": A = smf.synth_planted(400000, 100, 16, 2, 12345)
    elif c == "planted1m": A = smf.synth_planted(1000000, 1000, 16, 2, 12345)
    elif c == "planted4m": A = smf.synth_planted(4000000, 1000, 16, 2, 12345)
    else: raise SystemExit(c)
    print(c, "rows", A.rows, "nnz", A.nnz, "gen %.1fs" % (time.time() - t0), flush=True)
    dA = A.toGpuCSR()
    for rep in range(3):
        dC, st = smf.gpuSpMMWrapper(dA, dA, want_stats=True)
        dC.deviceDispose()
        by = 12 * A.nnz + 4 * (A.rows + 1) + 12 * st["products"] + 8 * A.nnz + 12 * st["nnz_out"] + 4 * (A.rows + 1)
        print("  spgemm rep%d total %.2f ms (flops %.2f sym %.2f num %.2f other %.2f) P=%d nnzC=%d GF=%.1f algGB/s=%.0f bins=%s" % (
            rep, st["ms_total"], st["ms_flops"], st["ms_symbolic"], st["ms_numeric"], st["ms_other"],
            st["products"], st["nnz_out"], 2 * st["products"] / st["ms_total"] / 1e6, by / st["ms_total"] / 1e6,
            st["bins_rows"][:6]), flush=True)
        if rep == 2:
            for b in range(1, 6):
                print("    sym bin%d rows %8d P %12d  %8.2f ms | num bin%d rows %8d P %12d nnzC %11d %8.2f ms" % (
                    b, st["sym_bin_rows"][b], st["sym_bin_products"][b], st["ms_sym_bin"][b],
                    b, st["bins_rows"][b], st["num_bin_products"][b], st["num_bin_nnzC"][b], st["ms_num_bin"][b]), flush=True)
    if os.environ.get("DUMP_DIST"):
        dC = smf.gpuSpMMWrapper(dA, dA)
        import ctypes as C
        from sparse_matrix_with_flops_b200 import _lib
        pre = smf.flops_prefix(dA, dA)
        rp = C.c_void_p()
        _lib.load().b200_csr_device_ptrs(dC.handle, C.byref(rp), None, None)
        out = np.empty(A.rows + 1, dtype=np.int64)
        rt = C.CDLL("libcudart.so.12")
        rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        assert rt.cudaMemcpy(out.ctypes.data, rp.value, out.nbytes, 2) == 0
        os.makedirs("gpurun_out", exist_ok=True)
        np.savez_compressed("gpurun_out/dist_%s.npz" % c, P=np.diff(pre).astype(np.int64), nnzC=np.diff(out).astype(np.int32),
                            nnzA=np.diff(A.rowPtr).astype(np.int32))
        dC.deviceDispose()
    if c != "rmat20d":
        dM = dA
        for it in range(6):
            dN, chaos, st = smf.gpuRmclOneStep(dA, dM, want_stats=True)
            print("  rmcl it%d total %.2f ms (sym %.2f num %.2f) P=%d unpruned=%d nnz=%d chaos=%.4g" % (
                it, st["ms_total"], st["ms_symbolic"], st["ms_numeric"], st["products"], st["nnz_unpruned"], st["nnz_out"], chaos), flush=True)
            if dM is not dA: dM.deviceDispose()
            dM = dN
    dA.deviceDispose()
