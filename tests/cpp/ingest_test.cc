// Host-only test of the C++ layer's ingest path (no GPU): COO::readSNAPFile -> rmclInit, printed
// as plain text for tests/test_capi_cpu.py to compare with the golden vectors of the reference.
#include "b200_nlibs.hpp"
using namespace b200::nlibs;
int main(int argc, char** argv) {
  if (argc < 3) return 2;
  COO coo;
  coo.readSNAPFile(argv[1], atoi(argv[2]) != 0);
  printf("coo %d %d %d\n", coo.rows, coo.cols, coo.nnz);
  const int before = coo.nnz;
  const int after = coo.orderedAndDuplicatesRemoving();   // returns the NEW nnz (nlibs/COO.cc:265)
  printf("removed %d\n", before - after);
  if (argc > 4) {   // dedup check: "r c v" triples on stdin, printed back after the call
    COO d;
    d.rows = atoi(argv[4]); d.cols = atoi(argv[5]);
    int cap = 1 << 16, n = 0;
    d.cooRowIndex = (int*)malloc(cap * sizeof(int)); d.cooColIndex = (int*)malloc(cap * sizeof(int));
    d.cooVal = (QValue*)malloc(cap * sizeof(QValue));
    while (n < cap && scanf("%d %d %lf", &d.cooRowIndex[n], &d.cooColIndex[n], &d.cooVal[n]) == 3) ++n;
    d.nnz = n;
    const int ret = d.orderedAndDuplicatesRemoving();
    printf("dedup %d %d\n", ret, d.nnz);
    for (int e = 0; e < d.nnz; ++e) printf("d %d %d %.17g\n", d.cooRowIndex[e], d.cooColIndex[e], d.cooVal[e]);
    d.dispose();
  }
  for (int e = 0; e < coo.nnz; ++e) printf("e %d %d %.17g\n", coo.cooRowIndex[e], coo.cooColIndex[e], coo.cooVal[e]);
  CSR M = rmclInit(coo);
  printf("csr %d %d %d\n", M.rows, M.cols, M.nnz);
  for (int i = 0; i <= M.rows; ++i) printf("p %d\n", M.rowPtr[i]);
  for (int p = 0; p < M.nnz; ++p) printf("v %d %.17g\n", M.colInd[p], M.values[p]);
  // PCSR: c column stripes (nlibs/PCSR.cc:3-56), printed block by block
  const int c = argc > 3 ? atoi(argv[3]) : 2;
  PCSR P(M, c);
  printf("pcsr %d %d\n", c, P.stride());
  for (int b = 0; b < c; ++b) {
    printf("blk %d %d\n", b, P.blocks[b].nnz);
    for (int i = 0; i <= M.rows; ++i) printf("bp %d %d\n", b, P.blocks[b].rowPtr[i]);
    for (int p = 0; p < P.blocks[b].nnz; ++p) printf("bv %d %d %.17g\n", b, P.blocks[b].colInd[p], P.blocks[b].values[p]);
  }
  P.dispose();
  Options o;
  const char* av[] = {"x", "-i", "some.snap", "--maxIters", "7", "-r", "SOMP", "--stride", "64", "-s"};
  process_args(10, (char**)av, o);
  printf("opts %s %d %d %d %d\n", o.inputFileName, o.maxIters, (int)o.rmclOption, o.stride, (int)o.stats);
  M.dispose(); coo.dispose();
  return 0;
}
