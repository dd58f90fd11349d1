import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle_lib as ol
import sparse_matrix_with_flops_b200 as smf
smf.init(0)
A = smf.synth_planted(3000, 30, 6, 1, 1)
M = ol.from_csr(A)
raw = ol.o_rmcl_onestep(M, M)
C = ol.o_spgemm(M, M)
got = ol.from_csr(A.staticOmpRmclOneStep(A))
print("I equal", np.array_equal(got.I, raw.I), "J equal", np.array_equal(got.J, raw.J))
if np.array_equal(got.I, raw.I):
    bad = np.nonzero((got.J != raw.J) | (got.V.view(np.int64) != raw.V.view(np.int64)))[0]
    print("bad entries", len(bad), "of", len(raw.J))
    rows = np.searchsorted(raw.I, bad, side="right") - 1
    ur = np.unique(rows)
    print("bad rows", len(ur), ur[:10])
    for r in ur[:3]:
        s, e = raw.I[r], raw.I[r + 1]
        print("row", r, "nnzC unpruned", C.I[r + 1] - C.I[r], "kept", e - s)
        print(" want J", raw.J[s:e][:12], "\n got  J", got.J[s:e][:12])
        print(" want V", raw.V[s:e][:6], "\n got  V", got.V[s:e][:6])
        print(" maxrel", np.max(np.abs(np.sort(got.V[s:e]) - np.sort(raw.V[s:e])) / np.sort(raw.V[s:e])))
else:
    d = np.nonzero(np.diff(got.I) != np.diff(raw.I))[0]
    print("rows with different counts", len(d), d[:10])
