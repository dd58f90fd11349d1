// nrmcl_b200.cc — C++ host driver over the C-ABI, in the shape of the reference's one-main
// drivers (nrmcl.cc:12-37 for rMCL, perfTests/only-somp.cc:24-37 for SpGEMM timing).  Host code
// only: every computation goes through include/b200_nlibs.hpp -> libb200spgemm.so.
//
//   nrmcl_b200.x --input FILE [--maxIters N] [--rmclOptions B200] [--eps E]     (nrmcl.cc's flags)
//                [--expect FILE]   ends with nrmcl.cc's verdict line: Same / Diffs (exit code 0 / 1)
//                [--output FILE]   writes the final Mt ("rows cols nnz", then "row col value" lines)
//   nrmcl_b200.x rmcl  <rmat|stencil|planted> <size> [maxIters] [eps]
//   nrmcl_b200.x spmm  <rmat|stencil|planted> <size> [reps]
//   nrmcl_b200.x spmm-host <rmat|stencil|planted> <size> [reps]   (host CSR in, host row blocks out)
//
// size: R-MAT scale / stencil grid edge / planted-partition vertices (1000-vertex blocks).
#include <chrono>
#include <set>
#include "b200_nlibs.hpp"
#include "b200_synth.h"   // harness-only generators (libb200synth.so)

using namespace b200::nlibs;

static double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

static CSR make_input(const char* kind, int size, bool symmetrise) {
  int rows = 0; long long nnz = 0;
  int *I = NULL, *J = NULL; double* V = NULL;
  int rc;
  if (!strcmp(kind, "rmat")) rc = b200_synth_rmat(size, 16, 12345ull, symmetrise, &rows, &I, &J, &V, &nnz);
  else if (!strcmp(kind, "stencil")) rc = b200_synth_stencil27(size, size, size, &rows, &I, &J, &V, &nnz);
  else if (!strcmp(kind, "planted")) rc = b200_synth_planted(size, std::max(1, size / 1000), 16, 2, 12345ull, &rows, &I, &J, &V, &nnz, NULL);
  else { fprintf(stderr, "unknown input kind %s\n", kind); exit(EXIT_FAILURE); }
  b200_check(rc, "b200_synth");
  return CSR(V, J, I, rows, rows, (int)nnz);
}

// nrmcl.cc:12-37 with the GPU path: RMCL(file), time, cluster count
static int run_file(int argc, char* argv[]) {
  Options options;
  if (process_args(argc, argv, options)) return 0;
  printf("input=%s maxIters=%d\n", options.inputFileName, options.maxIters);
  double t0 = now_ms();
  COO cooAt;
  cooAt.readSNAPFile(options.inputFileName);
  printf("rows=%d cols=%d nnz=%d\n", cooAt.rows, cooAt.cols, cooAt.nnz);
  int iters = 0;
  CSR Mt = RMCL(cooAt, options.maxIters, options.rmclOption, options.eps, &iters);
  cooAt.dispose();
  printf("time pass b200 rmcl total = %lf\n", now_ms() - t0);
  std::set<int> attractors;
  for (int i = 0; i < Mt.rows; ++i) {
    int best = -1; double bv = 0.0;
    for (int p = Mt.rowPtr[i]; p < Mt.rowPtr[i + 1]; ++p)
      if (best < 0 || Mt.values[p] > bv) { best = Mt.colInd[p]; bv = Mt.values[p]; }
    attractors.insert(best);
  }
  printf("iters %d final nnz %d clusters %zu\n", iters, Mt.nnz, attractors.size());
  if (options.outputFileName[0]) Mt.writeText(options.outputFileName);
  int ret = 0;
  if (options.expectFileName[0]) {
    // nrmcl.cc:25-32: both matrices ordered, compared, one verdict line.  The matrix to compare
    // with comes from a file here (the reference computes it with its SEQ path in the same
    // process; this library has no second path to compute it with)
    CSR want = CSR::readText(options.expectFileName);
    want.makeOrdered();
    Mt.makeOrdered();
    const bool isSame = Mt.isEqual(want);
    printf(isSame ? "Same\n" : "Diffs\n");
    ret = isSame ? 0 : 1;
    want.dispose();
  }
  Mt.dispose();
  b200_finalize();
  return ret;
}

int main(int argc, char* argv[]) {
  if (argc >= 2 && argv[1][0] == '-') return run_file(argc, argv);
  if (argc < 4) {
    fprintf(stderr, "usage: %s rmcl|spmm|spmm-host rmat|stencil|planted size [maxIters|reps] [eps]\n", argv[0]);
    return 2;
  }
  const bool rmcl = !strcmp(argv[1], "rmcl");
  const int size = atoi(argv[3]);
  const int count = argc > 4 ? atoi(argv[4]) : 5;   // maxIters default 5 (process_args.h:28)
  const double eps = argc > 5 ? atof(argv[5]) : 0.0;
  b200_ensure_init();
  CSR A = make_input(argv[2], size, rmcl);
  printf("input %s %d: rows %d nnz %d\n", argv[2], size, A.rows, A.nnz);
  if (rmcl) {
    CSR Mgt = A.deepCopy();
    std::vector<double> hist(std::max(1, count));
    int iters = 0;
    double t0 = now_ms();
    gpuRmclIter(count, Mgt, A, eps, &iters, hist.data());
    double ms = now_ms() - t0;
    std::set<int> attractors;
    for (int i = 0; i < A.rows; ++i) {
      int best = -1; double bv = 0.0;
      for (int p = A.rowPtr[i]; p < A.rowPtr[i + 1]; ++p)
        if (best < 0 || A.values[p] > bv) { best = A.colInd[p]; bv = A.values[p]; }
      attractors.insert(best);
    }
    printf("time pass b200 rmcl total = %lf ms, iters %d, %.3f iter/s, final nnz %d, chaos %.6g, clusters %zu\n",
           ms, iters, iters / (ms * 1e-3), A.nnz, iters ? hist[iters - 1] : 0.0, attractors.size());
    Mgt.dispose();
  } else if (!strcmp(argv[1], "spmm-host")) {
    // host buffers in, host row blocks out (CSR::flops_spmm's contract, streamed because nnz(C)
    // may exceed an int): every block is consumed and handed back, as a caller that writes the
    // product out block by block would do
    const long long products = A.spMMFlops(A);
    struct Sink {
      long long nnz; int blocks;
      void operator()(int, int, CSR blk) {
        nnz += blk.nnz; ++blocks;
        b200_host_free(blk.values); b200_host_free(blk.colInd); b200_host_free(blk.rowPtr);
      }
    };
    double best = 1e300;
    Sink last = {0, 0};
    for (int r = 0; r < count + 1; ++r) {
      Sink sink = {0, 0};
      struct Ref { Sink* s; void operator()(int lo, int hi, CSR blk) { (*s)(lo, hi, blk); } } ref = {&sink};
      double t0 = now_ms();
      A.spmmBlocks(A, ref);
      double ms = now_ms() - t0;
      if (r) best = std::min(best, ms);
      last = sink;
    }
    printf("b200 spmm-host best %lf ms, products %lld, nnz(C) %lld in %d row blocks, GFLOPS %.3f (host to host)\n",
           best, products, last.nnz, last.blocks, 2.0 * products / best / 1e6);
  } else {
    const long long products = A.spMMFlops(A);
    CSR dA = A.toGpuCSR();
    double best = 1e300;
    for (int r = 0; r < count + 1; ++r) {   // 1 warm-up + count reps, as perfTests/only-somp.cc does
      double t0 = now_ms();
      CSR dC = gpuSpMMWrapper(dA, dA);
      double ms = now_ms() - t0;
      if (r) best = std::min(best, ms);
      dC.deviceDispose();
    }
    printf("b200 spmm best %lf ms, products %lld, GFLOPS %.3f\n", best, products, 2.0 * products / best / 1e6);
    dA.deviceDispose();
  }
  A.dispose();
  b200_finalize();
  return 0;
}
