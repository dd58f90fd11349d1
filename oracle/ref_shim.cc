// ref_shim.cc — extern "C" doorway into the UNMODIFIED reference, for the checker only.
//
// TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header).  This file contains no algorithm: it
// includes the reference's own headers from /root/reference/nlibs and forwards to the
// reference's own functions, compiled from where they lie by oracle/Makefile into
// oracle/_ref/libref.so (git-ignored; never a copy of reference sources in this repo).
// It does not exist on the GPU box unless the prebuilt oracle/_ref/libref.so travelled there.
#include <omp.h>
#include <string.h>
#include "CSR.h"
#include "COO.h"
#include "PCSR.h"
#include "qrmcl.h"
#include "process_args.h"
#include "tools/util.h"
#include "tools/prefixSum.h"
#include "tools/prefixSum64.h"

// non-static but undeclared in qrmcl.h (nlibs/qrmcl.cc:8, :86)
void mtRmclIter(const int maxIter, const CSR Mgt, CSR& Mt, const int stride,
                const RunOptions runOptions);
void seqRmclIter(const int maxIter, const CSR Mgt, CSR& Mt);
// the thread_datas overload of flops_omp_CSR_SpMM (nlibs/flops_csr_kernel.cc:33) is not
// declared in cpu_csr_kernel.h
void flops_omp_CSR_SpMM(const int IA[], const int JA[], const QValue A[], const int nnzA,
                        const int IB[], const int JB[], const QValue B[], const int nnzB,
                        int*& IC, int*& JC, QValue*& C, int& nnzC, const int m, const int k,
                        const int n, const thread_data_t* thread_datas, const int stride);

static int nthreads_now() {
  int nt = 1;
#pragma omp parallel
#pragma omp master
  nt = omp_get_num_threads();
  return nt;
}

extern "C" {

int ref_num_threads() { return nthreads_now(); }

// variant: 0 sequential_CSR_SpMM, 1 omp_CSR_SpMM, 2 static_omp_CSR_SpMM, 3 flops_omp_CSR_SpMM
int ref_spgemm(int variant, const int* IA, const int* JA, const double* A, int nnzA,
               const int* IB, const int* JB, const double* B, int nnzB, int** IC, int** JC,
               double** C, int* nnzC, int m, int k, int n, int stride) {
  int *ic = NULL, *jc = NULL; double* c = NULL; int nn = 0;
  switch (variant) {
    case 0: sequential_CSR_SpMM(IA, JA, A, nnzA, IB, JB, B, nnzB, ic, jc, c, nn, m, k, n); break;
    case 1: omp_CSR_SpMM(IA, JA, A, nnzA, IB, JB, B, nnzB, ic, jc, c, nn, m, k, n, stride); break;
    case 2: static_omp_CSR_SpMM(IA, JA, A, nnzA, IB, JB, B, nnzB, ic, jc, c, nn, m, k, n, stride); break;
    case 3: flops_omp_CSR_SpMM(IA, JA, A, nnzA, IB, JB, B, nnzB, ic, jc, c, nn, m, k, n, stride); break;
    default: return 1;
  }
  *IC = ic; *JC = jc; *C = c; *nnzC = nn;
  return 0;
}

// Timed variant for the CPU baseline: thread scratch allocated outside the timed region, as
// perfTests/only-somp.cc:24-35 does.  Returns milliseconds of the best of `reps` runs.
double ref_spgemm_timed(int variant, const int* IA, const int* JA, const double* A, int nnzA,
                        const int* IB, const int* JB, const double* B, int nnzB, int m, int k,
                        int n, int stride, int reps, long long* nnzC_out) {
  const int nt = nthreads_now();
  thread_data_t* td = allocateThreadDatas(nt, n);
  double best = 1e300;
  for (int r = 0; r < reps; ++r) {
    int *ic = NULL, *jc = NULL; double* c = NULL; int nn = 0;
    double t0 = omp_get_wtime();
    if (variant == 1) omp_CSR_SpMM(IA, JA, A, nnzA, IB, JB, B, nnzB, ic, jc, c, nn, m, k, n, td, stride);
    else if (variant == 2) static_omp_CSR_SpMM(IA, JA, A, nnzA, IB, JB, B, nnzB, ic, jc, c, nn, m, k, n, td, stride);
    else {
      flops_omp_CSR_SpMM(IA, JA, A, nnzA, IB, JB, B, nnzB, ic, jc, c, nn, m, k, n, td, stride);
    }
    double ms = (omp_get_wtime() - t0) * 1e3;
    if (ms < best) best = ms;
    if (nnzC_out) *nnzC_out = nn;
    free(ic); free(jc); free(c);
  }
  freeThreadDatas(td, nt);
  return best;
}

// variant: 1 omp_CSR_RMCL_OneStep, 2 static_omp_CSR_RMCL_OneStep, 6 static_fair_CSR_RMCL_OneStep
int ref_rmcl_onestep(int variant, const int* IA, const int* JA, const double* A, int nnzA,
                     const int* IB, const int* JB, const double* B, int nnzB, int** IC, int** JC,
                     double** C, int* nnzC, int m, int k, int n, int stride) {
  int *ic = NULL, *jc = NULL; double* c = NULL; int nn = 0;
  const int nt = nthreads_now();
  if (variant == 6) {
    static_fair_CSR_RMCL_OneStep(IA, JA, A, nnzA, IB, JB, B, nnzB, ic, jc, c, nn, m, k, n, stride);
  } else {
    thread_data_t* td = allocateThreadDatas(nt, n);
    if (variant == 1) omp_CSR_RMCL_OneStep(IA, JA, A, nnzA, IB, JB, B, nnzB, ic, jc, c, nn, m, k, n, td, stride);
    else static_omp_CSR_RMCL_OneStep(IA, JA, A, nnzA, IB, JB, B, nnzB, ic, jc, c, nn, m, k, n, td, stride);
    freeThreadDatas(td, nt);
  }
  *IC = ic; *JC = jc; *C = c; *nnzC = nn;
  return 0;
}

// The reference's loop (nlibs/qrmcl.cc): runOption 0 = SEQ (seqRmclIter), 1 = OMP, 4 = SOMP,
// 6 = SFOMP (mtRmclIter).  Inputs are copied; the final Mt is returned malloc()'d.
// ms_out (may be NULL) gets the wall time of the loop.
int ref_rmcl_iter(int runOption, int maxIter, const int* IG, const int* JG, const double* G,
                  const int* IT, const int* JT, const double* T, int n, int** IM, int** JM,
                  double** M, int* nnzM, double* ms_out) {
  CSR Mgt(const_cast<double*>(G), const_cast<int*>(JG), const_cast<int*>(IG), n, n, IG[n]);
  CSR Tin(const_cast<double*>(T), const_cast<int*>(JT), const_cast<int*>(IT), n, n, IT[n]);
  CSR Mt = Tin.deepCopy();
  options.stats = false;
  options.stride = 512;
  double t0 = omp_get_wtime();
  if (runOption == 0) seqRmclIter(maxIter, Mgt, Mt);
  else mtRmclIter(maxIter, Mgt, Mt, 512, (RunOptions)runOption);
  if (ms_out) *ms_out = (omp_get_wtime() - t0) * 1e3;
  *IM = Mt.rowPtr; *JM = Mt.colInd; *M = Mt.values; *nnzM = Mt.nnz;
  return 0;
}

void ref_make_ordered(int* I, int* J, double* V, int rows, int cols) {
  CSR M(V, J, I, rows, cols, I[rows]);
  M.makeOrdered();
}

int ref_is_equal(int* I1, int* J1, double* V1, int* I2, int* J2, double* V2, int rows, int cols) {
  CSR X(V1, J1, I1, rows, cols, I1[rows]);
  CSR Y(V2, J2, I2, rows, cols, I2[rows]);
  return X.isEqual(Y) ? 1 : 0;
}

double ref_compute_threshold(double avg, double mx) { return computeThreshold(avg, mx); }

void ref_inflation_r2(const double* in, int count, double* out) { arrayInflationR2(in, count, out); }

void ref_max_sum(const double* v, int count, double* mx, double* sum) {
  std::pair<double, double> ms = arrayMaxSum(v, count);
  *mx = ms.first; *sum = ms.second;
}

double ref_thresh_prune_normalize(double thresh, const int* rind, const double* rval, int* count,
                                  int* ind, double* val) {
  return arrayThreshPruneNormalize(thresh, rind, rval, count, ind, val);
}

void ref_equal_partition64(long* prefix, int n, int nthreads, int* ends) {
  arrayEqualPartition64(prefix, n, nthreads, ends);
}

// dynamic_omp_CSR_flops must run inside a parallel region (flops_csr_kernel.cc:10-13)
void ref_flops_prefix(const int* IA, const int* JA, const int* IB, const int* JB, int m, int n,
                      long* rowFlops) {
#pragma omp parallel
  dynamic_omp_CSR_flops(IA, JA, IB, JB, m, n, rowFlops, 512);
}

// rmclInit (nlibs/qrmcl.cc:126-134) on an in-memory COO (COO.cc:24-35 constructor).
int ref_rmcl_init(const int* er, const int* ec, int nedges, int n, int** I, int** J, double** V,
                  int* nnz) {
  double* ones = (double*)malloc(sizeof(double) * (nedges + 1));
  for (int e = 0; e < nedges; ++e) ones[e] = 1.0;
  COO coo(ones, ec, er, n, n, nedges);
  free(ones);
  CSR M = rmclInit(coo);
  coo.dispose();
  *I = M.rowPtr; *J = M.colInd; *V = M.values; *nnz = M.nnz;
  return 0;
}

// PCSR(csr, c) (nlibs/PCSR.cc:3-56): returns copies of the c blocks laid out like
// oracle_pcsr_split's outputs.
void ref_pcsr_split(int* I, int* J, double* V, int rows, int cols, int c, int* blockPtr,
                    int* rowPtr, int* Jout, double* Vout) {
  CSR M(V, J, I, rows, cols, I[rows]);
  PCSR P(M, c);
  blockPtr[0] = 0;
  for (int b = 0; b < c; ++b) {
    blockPtr[b + 1] = blockPtr[b] + P.blocks[b].nnz;
    memcpy(rowPtr + (size_t)b * (rows + 1), P.blocks[b].rowPtr, sizeof(int) * (rows + 1));
    memcpy(Jout + blockPtr[b], P.blocks[b].colInd, sizeof(int) * P.blocks[b].nnz);
    memcpy(Vout + blockPtr[b], P.blocks[b].values, sizeof(double) * P.blocks[b].nnz);
  }
  P.dispose();
}

// COO::orderedAndDuplicatesRemoving (nlibs/COO.cc:237-266) on an in-memory COO: entries sorted
// by (row, col), the values of repeated pairs SUMMED, returns the new nnz (also the method's
// return value, in *ret).  Outputs are copies of the first new-nnz entries.
int ref_coo_dedup(const int* er, const int* ec, const double* ev, int nedges, int rows, int cols,
                  int* r_out, int* c_out, double* v_out, int* ret) {
  COO coo(ev, ec, er, rows, cols, nedges);   // (values, colIndex, rowIndex, ...) deep copy, COO.cc:24-35
  *ret = coo.orderedAndDuplicatesRemoving();
  const int nn = coo.nnz;
  memcpy(r_out, coo.cooRowIndex, sizeof(int) * nn);
  memcpy(c_out, coo.cooColIndex, sizeof(int) * nn);
  memcpy(v_out, coo.cooVal, sizeof(double) * nn);
  coo.dispose();
  return nn;
}

void ref_free(void* p) { free(p); }

}  // extern "C"
