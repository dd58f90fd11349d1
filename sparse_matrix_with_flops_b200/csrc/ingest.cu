// ingest.cu — COO -> CSR on the device (the step immediately before the
// hot path).  Restates, for an edge list already in device memory:
//   COO::addSelfLoopIfNeeded        nlibs/COO.cc:160-188  (one (i,i,1.0) per vertex without a diagonal)
//   COO::makeOrdered / toCSR        nlibs/COO.cc:222-235  (entries ordered by (row, col), row offsets)
//   orderedAndDuplicatesRemoving    nlibs/COO.cc:237-266  (repeated (row, col) pairs: ONE entry, values added)
//   CSR::averAndNormRowQValue       nlibs/CSR.cc:88-95    (every entry of a row = 1 / rowcount)
// i.e. rmclInit (nlibs/qrmcl.cc:126-134) when self loops + normalisation are asked for.
//
// One stable radix sort of 64-bit (row << 32 | col) keys carrying the entry index; self-loop
// candidates are appended AFTER the real entries, so that among equal keys a real diagonal comes
// first and the candidate is dropped; a flag pass marks survivors and counts rows; two scans give
// positions and row offsets; one scatter writes columns and values.  All passes are streaming
// (HBM-bound, 12-24 B per entry and pass).
#include <cub/cub.cuh>
#include "common.cuh"

namespace b200 {
namespace {

struct IntAsI64 {
  __host__ __device__ __forceinline__ long long operator()(const int& x) const { return (long long)x; }
};

__global__ void k_coo_keys(const int* __restrict__ row, const int* __restrict__ col, long long nnz,
                           long long total, int rows, int cols, unsigned long long* __restrict__ key,
                           unsigned* __restrict__ idx, int* __restrict__ bad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int r, c;
  if (i < nnz) {
    r = row[i];
    c = col[i];
    if ((unsigned)r >= (unsigned)rows || (unsigned)c >= (unsigned)cols) { *bad = 1; r = 0; c = 0; }
  } else {
    r = c = (int)(i - nnz);  // self-loop candidate of vertex i - nnz
  }
  key[i] = ((unsigned long long)(unsigned)r << 32) | (unsigned)c;
  idx[i] = (unsigned)i;
}

// survivor flags + entries per row
__global__ void k_coo_keep(const unsigned long long* __restrict__ key, const unsigned* __restrict__ idx,
                           long long total, long long nnz, int dedup, int* __restrict__ keep,
                           int* __restrict__ rowcnt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const unsigned long long k = key[i];
  const bool same = i > 0 && key[i - 1] == k;
  const bool candidate = (long long)idx[i] >= nnz;
  const int kp = !(same && (dedup || candidate));
  keep[i] = kp;
  if (kp) atomicAdd(&rowcnt[(int)(k >> 32)], 1);
}

__global__ void k_coo_scatter(const unsigned long long* __restrict__ key, const unsigned* __restrict__ idx,
                              const int* __restrict__ keep, const long long* __restrict__ pos,
                              long long total, long long nnz, const double* __restrict__ val,
                              const int64_t* __restrict__ rowptr, int normalise, int dedup,
                              int* __restrict__ col_out, double* __restrict__ val_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total || !keep[i]) return;
  const unsigned long long k = key[i];
  const int r = (int)(k >> 32);
  const long long p = pos[i];
  col_out[p] = (int)(unsigned)(k & 0xffffffffull);
  double v;
  if (normalise) {
    v = 1.0 / (double)(rowptr[r + 1] - rowptr[r]);     // nlibs/CSR.cc:92: 1.0 / count
  } else {
    const long long e = (long long)idx[i];
    v = e < nnz ? (val ? val[e] : 1.0) : 1.0;           // nlibs/COO.cc:183: self loops carry 1.0
    // orderedAndDuplicatesRemoving (COO.cc:246-248): the entry that stays collects the values of
    // the repeated pairs behind it, in input order (the sort is stable; a self-loop candidate
    // behind a real diagonal adds nothing)
    if (dedup)
      for (long long q = i + 1; q < total && key[q] == k; ++q) {
        const long long eq = (long long)idx[q];
        if (eq < nnz) v += val ? val[eq] : 1.0;
      }
  }
  val_out[p] = v;
}

}  // namespace

// d_row / d_col / d_val: device arrays of nnz entries (d_val may be null: all ones)
int coo_build_device(const int* d_row, const int* d_col, const double* d_val, long long nnz, int rows,
                     int cols, int flags, DevCSR* out) {
  Ctx& c = ctx();
  cudaStream_t st = c.stream;
  const bool dedup = flags & 1, loops = flags & 2, normalise = flags & 4;
  if (loops && rows != cols) { set_error("self loops need a square matrix"); return B200_ERR_BAD_ARG; }
  const long long total = nnz + (loops ? rows : 0);
  if (total > 0xffffffffll) { set_error("edge list too long"); return B200_ERR_BAD_ARG; }
  rb_reset();
  *out = DevCSR();
  Temps T;
  DevCSR d;
  d.rows = rows; d.cols = cols;
  B200_CUDA(T.alloc(&d.rowptr, (size_t)rows + 1));
  int* d_rowcnt = nullptr;
  B200_CUDA(T.alloc(&d_rowcnt, (size_t)rows + 1));
  B200_CUDA(cudaMemsetAsync(d_rowcnt, 0, ((size_t)rows + 1) * sizeof(int), st));
  unsigned long long *key = nullptr, *key2 = nullptr;
  unsigned *idx = nullptr, *idx2 = nullptr;
  int *keep = nullptr, *bad = nullptr;
  long long* pos = nullptr;
  B200_CUDA(T.alloc(&key, (size_t)total));
  B200_CUDA(T.alloc(&key2, (size_t)total));
  B200_CUDA(T.alloc(&idx, (size_t)total));
  B200_CUDA(T.alloc(&idx2, (size_t)total));
  B200_CUDA(T.alloc(&keep, (size_t)total));
  B200_CUDA(T.alloc(&pos, (size_t)total + 1));
  B200_CUDA(T.alloc(&bad, 1));
  B200_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), st));
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (total > 0) {
    k_coo_keys<<<grid, 256, 0, st>>>(d_row, d_col, nnz, total, rows, cols, key, idx, bad);
    int end_bit = 33;
    while (end_bit < 64 && (1ll << (end_bit - 32)) < (long long)rows) ++end_bit;
    void* tmp = nullptr;
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, key, key2, idx, idx2, total, 0, end_bit, st);
    B200_CUDA(T.alloc((char**)&tmp, tb ? tb : 1));
    cub::DeviceRadixSort::SortPairs(tmp, tb, key, key2, idx, idx2, total, 0, end_bit, st);
    B200_CUDA(cudaGetLastError());
    k_coo_keep<<<grid, 256, 0, st>>>(key2, idx2, total, nnz, dedup ? 1 : 0, keep, d_rowcnt);
    cub::TransformInputIterator<long long, IntAsI64, const int*> it(keep, IntAsI64());
    tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, it, pos, total, st);
    void* tmp2 = nullptr;
    B200_CUDA(T.alloc((char**)&tmp2, tb ? tb : 1));
    cub::DeviceScan::ExclusiveSum(tmp2, tb, it, pos, total, st);
    B200_CUDA(cudaGetLastError());
  }
  {
    cub::TransformInputIterator<long long, IntAsI64, const int*> it(d_rowcnt, IntAsI64());
    size_t tb = 0;
    void* tmp = nullptr;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, it, (long long*)d.rowptr, rows + 1, st);
    B200_CUDA(T.alloc((char**)&tmp, tb ? tb : 1));
    cub::DeviceScan::ExclusiveSum(tmp, tb, it, (long long*)d.rowptr, rows + 1, st);
    B200_CUDA(cudaGetLastError());
  }
  int h_bad = 0;
  long long h_nnz = 0;
  B200_CUDA(d2h_small(&h_bad, bad, sizeof(int), st));
  B200_CUDA(d2h_small(&h_nnz, d.rowptr + rows, sizeof(long long), st));
  B200_CUDA(sync_fetch(st));
  if (h_bad) { set_error("edge list holds a row or column index outside the matrix"); return B200_ERR_BAD_ARG; }
  d.nnz = h_nnz;
  B200_CUDA(T.alloc(&d.col, (size_t)h_nnz));
  B200_CUDA(T.alloc(&d.val, (size_t)h_nnz));
  if (total > 0)
    k_coo_scatter<<<grid, 256, 0, st>>>(key2, idx2, keep, pos, total, nnz, d_val, d.rowptr,
                                       normalise ? 1 : 0, dedup ? 1 : 0, d.col, d.val);
  B200_CUDA(cudaGetLastError());
  B200_CUDA(sync_fetch(st));
  d.sorted_rows = dedup;   // strictly ascending by construction once repeated pairs are gone
  if (!dedup) {
    const int rc = check_sorted_device(&d);
    if (rc) return rc;
  }
  T.keep(d.rowptr); T.keep(d.col); T.keep(d.val);
  *out = d;
  return B200_OK;
}

}  // namespace b200
