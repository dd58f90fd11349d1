"""Small run for compute-sanitizer (development aid): exercises warp tables, the bitmap kernels
with teams, the wide fallbacks and one rMCL step on inputs small enough for a sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sparse_matrix_with_flops_b200 as smf
smf.init(0)
A = smf.synth_rmat(12, 16, 12345, True)
dA = A.toGpuCSR()
dC, st = smf.gpuSpMMWrapper(dA, dA, want_stats=True)
print("spgemm", st["products"], st["nnz_out"], st["bins_rows"][:7], "parts", st["part_kernel"], st["part_count"])
dC.deviceDispose()
dM, ch = smf.gpuRmclOneStep(dA, dA)
dM2, ch2 = smf.gpuRmclOneStep(dA, dM)     # B unsorted (first-touch rows)
print("rmcl", dM.nnz, ch, dM2.nnz, ch2)
dM.deviceDispose(); dM2.deviceDispose()
os.environ["B200_FORCE_WIDE"] = "1"
dC = smf.gpuSpMMWrapper(dA, dA)
dM, ch = smf.gpuRmclOneStep(dA, dA)
print("wide", dC.nnz, dM.nnz)
dC.deviceDispose(); dM.deviceDispose(); dA.deviceDispose()
S = smf.synth_stencil27(12, 12, 12)
dS = S.toGpuCSR()
dC = smf.gpuSpMMWrapper(dS, dS)
print("stencil", dC.nnz)
dC.deviceDispose(); dS.deviceDispose()
print("done")
