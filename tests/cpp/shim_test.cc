// C++ test of the header-only host layer (include/b200_nlibs.hpp) in the style of the
// reference's one-main tests (correctTests/pcsrTest.cc:36-58, nrmcl.cc:12-37, tests/omp_spmm_test.cc):
// compute with the B200 path, compute a trusted answer (a dense product here: the matrices are
// tiny), makeOrdered() both, isEqual(), print Same / Diffs.  Exit code = number of Diffs.
#include <vector>
#include "b200_nlibs.hpp"
using namespace b200::nlibs;

static CSR dense_to_csr(const std::vector<double>& d, int rows, int cols) {
  CSR m;
  m.rows = rows; m.cols = cols;
  m.rowPtr = (int*)calloc(rows + 1, sizeof(int));
  std::vector<int> J; std::vector<double> V;
  for (int i = 0; i < rows; ++i) {
    for (int j = 0; j < cols; ++j) if (d[(size_t)i * cols + j] != 0.0) { J.push_back(j); V.push_back(d[(size_t)i * cols + j]); }
    m.rowPtr[i + 1] = (int)J.size();
  }
  m.nnz = (int)J.size();
  m.colInd = (int*)malloc((J.size() + 1) * sizeof(int));
  m.values = (double*)malloc((V.size() + 1) * sizeof(double));
  memcpy(m.colInd, J.data(), J.size() * sizeof(int));
  memcpy(m.values, V.data(), V.size() * sizeof(double));
  return m;
}
static std::vector<double> csr_to_dense(const CSR& m) {
  std::vector<double> d((size_t)m.rows * m.cols, 0.0);
  for (int i = 0; i < m.rows; ++i) for (int p = m.rowPtr[i]; p < m.rowPtr[i + 1]; ++p) d[(size_t)i * m.cols + m.colInd[p]] = m.values[p];
  return d;
}
static int report(const char* what, bool same) { printf("%s: %s\n", what, same ? "Same" : "Diffs"); return same ? 0 : 1; }

int main() {
  int diffs = 0;
  // a 40 x 40 matrix with a regular pattern plus a dense row (lands in the large-row bin)
  const int n = 40;
  std::vector<double> a((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) { a[(size_t)i * n + i] = 1.0 + i; a[(size_t)i * n + (i * 7 + 3) % n] = 0.5; a[(size_t)i * n + (i * 11 + 5) % n] += 0.25; }
  for (int j = 0; j < n; ++j) a[(size_t)5 * n + j] = 0.125 * (j + 1);
  CSR A = dense_to_csr(a, n, n);
  // SpGEMM through every mirrored entry point
  std::vector<double> want((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) for (int k = 0; k < n; ++k) if (a[(size_t)i * n + k] != 0.0) for (int j = 0; j < n; ++j) want[(size_t)i * n + j] += a[(size_t)i * n + k] * a[(size_t)k * n + j];
  CSR W = dense_to_csr(want, n, n);
  CSR C1 = A.flops_spmm(A), C2 = A.omp_spmm(A);
  int *IC, *JC, nnzC; QValue* Cv;
  flops_omp_CSR_SpMM(A.rowPtr, A.colInd, A.values, A.nnz, A.rowPtr, A.colInd, A.values, A.nnz, IC, JC, Cv, nnzC, n, n, n, 512);
  CSR C3(Cv, JC, IC, n, n, nnzC);
  C1.makeOrdered(); C2.makeOrdered(); C3.makeOrdered();
  diffs += report("flops_spmm vs dense product", C1.isEqual(W, 1e-12));
  diffs += report("omp_spmm vs flops_spmm", C2.isEqual(C1, 0.0));
  diffs += report("flops_omp_CSR_SpMM vs flops_spmm", C3.isEqual(C1, 0.0));
  {  // static_omp_CSR_SpMM with thread scratch, the SOMP variant's raw entry point (cpu_csr_kernel.h:91-94)
    int *I2, *J2, nnz2; QValue* V2;
    static_omp_CSR_SpMM(A.rowPtr, A.colInd, A.values, A.nnz, A.rowPtr, A.colInd, A.values, A.nnz, I2, J2, V2, nnz2,
                        n, n, n, (const thread_data_t*)NULL, 512);
    CSR C6(V2, J2, I2, n, n, nnz2);
    C6.makeOrdered();
    diffs += report("static_omp_CSR_SpMM vs flops_spmm", C6.isEqual(C1, 0.0));
    C6.dispose();
  }
  // device round trip + gpuSpMMWrapper
  CSR dA = A.toGpuCSR();
  CSR dC = gpuSpMMWrapper(dA, dA);
  CSR C4 = dC.toCpuCSR();
  diffs += report("gpuSpMMWrapper vs flops_spmm", C4.isEqual(C1, 0.0));
  dC.deviceDispose(); dA.deviceDispose();
  // the streamed row-block product, cut every 60 intermediate products, glued back together
  {
    struct Glue {
      std::vector<int> I, J; std::vector<double> V; int next;
      void operator()(int lo, int hi, CSR blk) {
        if (lo != next) I.assign(1, -1);               // blocks must arrive in row order
        next = hi;
        for (int i = 0; i < hi - lo; ++i) I.push_back((int)J.size() + blk.rowPtr[i + 1]);
        J.insert(J.end(), blk.colInd, blk.colInd + blk.nnz);
        V.insert(V.end(), blk.values, blk.values + blk.nnz);
        blk.dispose();
      }
    } g;
    g.I.assign(1, 0); g.next = 0;
    struct Ref { Glue* g; void operator()(int lo, int hi, CSR blk) { (*g)(lo, hi, blk); } } ref = {&g};
    A.spmmBlocks(A, ref, 60);
    bool ok = g.next == n && (int)g.I.size() == n + 1 && (int)g.J.size() == C1.nnz;
    if (ok) ok = CSR(g.V.data(), g.J.data(), g.I.data(), n, n, (int)g.J.size()).isEqual(C1, 0.0);
    diffs += report("spmmBlocks vs flops_spmm", ok);
  }
  // PCSR: column-striped product equals the plain one (correctTests/pcsrTest.cc)
  PCSR P(A, 3);
  CSR C5 = P.leftMultiply(A);
  C5.makeOrdered();
  diffs += report("PCSR(3) product vs plain", C5.isEqual(C1, 1e-12));
  diffs += report("PCSR nnz", P.nnz() == A.nnz);
  P.dispose();
  {  // the same container on device matrices: stripes cut, multiplied and glued on the device
    CSR dA2 = A.toGpuCSR();
    DevicePCSR DP(dA2, 3);
    CSR dC7 = DP.leftMultiply(dA2);
    CSR C7 = dC7.toCpuCSR();
    C7.makeOrdered();
    diffs += report("DevicePCSR(3) product vs plain", C7.isEqual(C1, 1e-12));
    dC7.deviceDispose(); DP.dispose(); dA2.deviceDispose();
  }
  // rMCL: rmclInit on a small ring-with-chords graph, 4 iterations; every row sums to 1 and the
  // loop equals 4 single steps
  std::vector<int> er, ec; std::vector<double> ev;
  for (int i = 0; i < n; ++i) { er.push_back(i); ec.push_back((i + 1) % n); er.push_back((i + 1) % n); ec.push_back(i); if (i % 5 == 0) { er.push_back(i); ec.push_back((i + 7) % n); er.push_back((i + 7) % n); ec.push_back(i); } }
  ev.assign(er.size(), 1.0);
  COO coo(ev.data(), ec.data(), er.data(), n, n, (int)er.size());
  CSR Mt = rmclInit(coo);
  CSR Mgt = Mt.deepCopy();
  CSR step = Mt.deepCopy();
  for (int it = 0; it < 4; ++it) { CSR nx = Mgt.staticOmpRmclOneStep(step, NULL, 512); step.dispose(); step = nx; }
  step.makeOrdered();
  gpuRmclIter(4, Mgt, Mt);
  diffs += report("gpuRmclIter(4) vs 4 x staticOmpRmclOneStep", Mt.isEqual(step, 1e-12));
  bool stochastic = true;
  for (int i = 0; i < n; ++i) { double s = 0; for (int p = Mt.rowPtr[i]; p < Mt.rowPtr[i + 1]; ++p) s += Mt.values[p]; stochastic &= fabs(s - 1.0) < 1e-12; }
  diffs += report("rows of Mt sum to 1", stochastic);
  (void)csr_to_dense;
  A.dispose(); W.dispose(); C1.dispose(); C2.dispose(); C3.dispose(); C4.dispose(); C5.dispose();
  Mt.dispose(); Mgt.dispose(); step.dispose(); coo.dispose();
  b200_finalize();
  return diffs;
}
