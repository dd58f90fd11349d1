// Synthetic graph generators for the SpGEMM / rMCL hot path (host side, C++17 + OpenMP).
//
// The reference has no generators: its drivers read SNAP / MatrixMarket files that are not in
// the repository (SURVEY.md §4 "Fixtures").  BASELINE.json names three synthetic families
// (R-MAT, 27-point stencil, planted partition); SURVEY.md §8(d) fixes their parameters.  All
// three end in the reference's `rmclInit` semantics (nlibs/qrmcl.cc:126-134): every vertex
// has a self-loop (COO::addSelfLoopIfNeeded, nlibs/COO.cc:160-188), rows are sorted by column
// (COO::makeOrdered, COO.cc:222-235) and every value is 1/rowcount
// (CSR::averAndNormRowQValue, nlibs/CSR.cc:88-95).  Edges are de-duplicated first because
// duplicate entries corrupt the reference (SURVEY.md §8c "Input hazards").
//
// Determinism: random draws are made in fixed chunks of 65536 edges, chunk c seeded with
// seed + c * golden-ratio constant, so the output does not depend on the thread count.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "b200_synth.h"

// return codes: the values of include/b200_spgemm.h (0 ok, 1 bad argument, 3 host allocation,
// 4 does not fit int32), repeated here so that this library does not depend on the product's
enum { B200_OK = 0, B200_ERR_BAD_ARG = 1, B200_ERR_HOST_ALLOC = 3, B200_ERR_INT32_OVERFLOW = 4 };

namespace {

// Build CSR (sorted, unique columns, self loops, values 1/rowcount) from an edge list.
// rows/cols of each edge in r[], c[]; consumes the vectors.
int edges_to_csr(int n, std::vector<int>& r, std::vector<int>& c, bool symmetrise,
                 int** IA_out, int** JA_out, double** A_out, long long* nnz_out) {
  const size_t E = r.size();
  // counting sort by row (capacity = in-degree bound incl. mirrored edges + self loop)
  std::vector<long long> start((size_t)n + 1, 0);
  for (size_t e = 0; e < E; ++e) {
    start[(size_t)r[e] + 1]++;
    if (symmetrise) start[(size_t)c[e] + 1]++;
  }
  for (int i = 0; i < n; ++i) start[(size_t)i + 1] += 1;  // room for the self loop
  for (int i = 0; i < n; ++i) start[(size_t)i + 1] += start[i];
  const long long cap = start[n];
  std::vector<int> cols((size_t)cap);
  std::vector<long long> fill(start.begin(), start.end() - 1);
  for (int i = 0; i < n; ++i) cols[(size_t)fill[i]++] = i;  // self loop
  for (size_t e = 0; e < E; ++e) {
    cols[(size_t)fill[r[e]]++] = c[e];
    if (symmetrise) cols[(size_t)fill[c[e]]++] = r[e];
  }
  std::vector<int>().swap(r);
  std::vector<int>().swap(c);
  std::vector<int> cnt((size_t)n);
#pragma omp parallel for schedule(dynamic, 1024)
  for (int i = 0; i < n; ++i) {
    int* b = cols.data() + start[i];
    int* e = cols.data() + start[(size_t)i + 1];
    std::sort(b, e);
    cnt[i] = (int)(std::unique(b, e) - b);
  }
  long long nnz = 0;
  for (int i = 0; i < n; ++i) nnz += cnt[i];
  if (nnz > 2147483647LL) return B200_ERR_INT32_OVERFLOW;
  int* IA = (int*)malloc(((size_t)n + 1) * sizeof(int));
  int* JA = (int*)malloc((size_t)nnz * sizeof(int));
  double* A = (double*)malloc((size_t)nnz * sizeof(double));
  if (!IA || !JA || !A) { free(IA); free(JA); free(A); return B200_ERR_HOST_ALLOC; }
  IA[0] = 0;
  for (int i = 0; i < n; ++i) IA[i + 1] = IA[i] + cnt[i];
#pragma omp parallel for schedule(dynamic, 1024)
  for (int i = 0; i < n; ++i) {
    const int k = cnt[i];
    const double v = 1.0 / k;  // CSR::averAndNormRowQValue, CSR.cc:88-95
    memcpy(JA + IA[i], cols.data() + start[i], (size_t)k * sizeof(int));
    for (int j = 0; j < k; ++j) A[IA[i] + j] = v;
  }
  *IA_out = IA; *JA_out = JA; *A_out = A; *nnz_out = nnz;
  return B200_OK;
}

}  // namespace

extern "C" {

int b200_synth_rmat(int scale, int edge_factor, unsigned long long seed, int symmetrise,
                    int* rows, int** IA, int** JA, double** A, long long* nnz) {
  if (scale < 1 || scale > 30 || edge_factor < 1) return B200_ERR_BAD_ARG;
  const int n = 1 << scale;
  const long long E = (long long)edge_factor * n;
  std::vector<int> r((size_t)E), c((size_t)E);
  // Graph500 parameters (a,b,c,d) = (.57,.19,.19,.05) as 16-bit fixed-point thresholds.
  const unsigned ta = (unsigned)(0.57 * 65536.0);
  const unsigned tb = ta + (unsigned)(0.19 * 65536.0);
  const unsigned tc = tb + (unsigned)(0.19 * 65536.0);
  const long long CH = 65536;
  const long long nch = (E + CH - 1) / CH;
#pragma omp parallel for schedule(dynamic, 4)
  for (long long ch = 0; ch < nch; ++ch) {
    std::mt19937_64 rng(seed + (unsigned long long)ch * 0x9E3779B97F4A7C15ULL);
    const long long lo = ch * CH, hi = std::min(E, lo + CH);
    for (long long e = lo; e < hi; ++e) {
      int u = 0, v = 0;
      unsigned long long w = 0;
      for (int bit = 0; bit < scale; ++bit) {
        if ((bit & 3) == 0) w = rng();
        const unsigned x = (unsigned)(w & 0xFFFF);
        w >>= 16;
        // one quadrant choice per bit: a=(0,0) b=(0,1) c=(1,0) d=(1,1)
        const int rb = x >= tb;               // c or d
        const int cb = (x >= ta && x < tb) || x >= tc;  // b or d
        u = (u << 1) | rb;
        v = (v << 1) | cb;
      }
      r[(size_t)e] = u;
      c[(size_t)e] = v;
    }
  }
  *rows = n;
  return edges_to_csr(n, r, c, symmetrise != 0, IA, JA, A, nnz);
}

int b200_synth_stencil27(int gx, int gy, int gz, int* rows, int** IA, int** JA, double** A,
                         long long* nnz) {
  if (gx < 1 || gy < 1 || gz < 1) return B200_ERR_BAD_ARG;
  const long long nn = (long long)gx * gy * gz;
  if (nn > 2147483647LL) return B200_ERR_INT32_OVERFLOW;
  const int n = (int)nn;
  // vertex id = (z*gy + y)*gx + x; clipped (non-periodic) 27-point neighbourhood incl. self
  std::vector<int> cnt((size_t)n);
#pragma omp parallel for schedule(static)
  for (int z = 0; z < gz; ++z)
    for (int y = 0; y < gy; ++y)
      for (int x = 0; x < gx; ++x) {
        const int nx = 1 + (x > 0) + (x < gx - 1);
        const int ny = 1 + (y > 0) + (y < gy - 1);
        const int nz = 1 + (z > 0) + (z < gz - 1);
        cnt[((size_t)z * gy + y) * gx + x] = nx * ny * nz;
      }
  long long tot = 0;
  for (int i = 0; i < n; ++i) tot += cnt[i];
  if (tot > 2147483647LL) return B200_ERR_INT32_OVERFLOW;
  int* ia = (int*)malloc(((size_t)n + 1) * sizeof(int));
  int* ja = (int*)malloc((size_t)tot * sizeof(int));
  double* a = (double*)malloc((size_t)tot * sizeof(double));
  if (!ia || !ja || !a) { free(ia); free(ja); free(a); return B200_ERR_HOST_ALLOC; }
  ia[0] = 0;
  for (int i = 0; i < n; ++i) ia[i + 1] = ia[i] + cnt[i];
#pragma omp parallel for schedule(static)
  for (int z = 0; z < gz; ++z)
    for (int y = 0; y < gy; ++y)
      for (int x = 0; x < gx; ++x) {
        const int i = (int)(((size_t)z * gy + y) * gx + x);
        int p = ia[i];
        const double v = 1.0 / cnt[i];
        for (int dz = -1; dz <= 1; ++dz) {
          if (z + dz < 0 || z + dz >= gz) continue;
          for (int dy = -1; dy <= 1; ++dy) {
            if (y + dy < 0 || y + dy >= gy) continue;
            for (int dx = -1; dx <= 1; ++dx) {
              if (x + dx < 0 || x + dx >= gx) continue;
              ja[p] = (int)(((size_t)(z + dz) * gy + (y + dy)) * gx + (x + dx));
              a[p] = v;
              ++p;
            }
          }
        }
      }
  *rows = n; *IA = ia; *JA = ja; *A = a; *nnz = tot;
  return B200_OK;
}

int b200_synth_planted(int n, int nblocks, int intra, int inter, unsigned long long seed,
                       int* rows, int** IA, int** JA, double** A, long long* nnz,
                       int** labels) {
  if (n < 1 || nblocks < 1 || nblocks > n || intra < 0 || inter < 0) return B200_ERR_BAD_ARG;
  const int bs = (n + nblocks - 1) / nblocks;  // block b = vertices [b*bs, min(n,(b+1)*bs))
  const int per = intra + inter;
  std::vector<int> r((size_t)n * per), c((size_t)n * per);
  const int CH = 65536;
  const int nch = (n + CH - 1) / CH;
#pragma omp parallel for schedule(dynamic, 1)
  for (int ch = 0; ch < nch; ++ch) {
    std::mt19937_64 rng(seed + (unsigned long long)ch * 0x9E3779B97F4A7C15ULL);
    const int lo = ch * CH, hi = std::min(n, lo + CH);
    for (int v = lo; v < hi; ++v) {
      const int b = v / bs;
      const int b0 = b * bs, b1 = std::min(n, b0 + bs);
      size_t p = (size_t)v * per;
      for (int k = 0; k < intra; ++k, ++p) {
        r[p] = v;
        c[p] = b0 + (int)(rng() % (unsigned long long)(b1 - b0));
      }
      for (int k = 0; k < inter; ++k, ++p) {
        r[p] = v;
        c[p] = (int)(rng() % (unsigned long long)n);
      }
    }
  }
  if (labels) {
    int* lab = (int*)malloc((size_t)n * sizeof(int));
    if (!lab) return B200_ERR_HOST_ALLOC;
    for (int v = 0; v < n; ++v) lab[v] = v / bs;
    *labels = lab;
  }
  *rows = n;
  return edges_to_csr(n, r, c, true, IA, JA, A, nnz);
}

// b200_host_free lives in capi.cu (it feeds the host block cache)

}  // extern "C"
