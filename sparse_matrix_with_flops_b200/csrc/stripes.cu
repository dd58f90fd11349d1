// stripes.cu — column stripes of a device CSR and their row-wise re-assembly: the device form of
// the reference's PCSR container (nlibs/PCSR.h:5-101, PCSR.cc:3-56: c column stripes of width
// ceil(cols / c), block b holding local column = column - b * stride) and of the merge in
// PCSR::leftMultiply (correctTests/pcsrTest.cc:7-19: C = A x PCSR(B) stripe by stripe).
// C = A x B[:, lo:hi) is then b200_spgemm_device(A, stripe): column blocking of B bounds the
// accumulator width of a pass — a second route for B wider than one pass can index.
#include <cub/cub.cuh>
#include "common.cuh"

namespace b200 {
namespace {

constexpr unsigned FULLW = 0xffffffffu;

// entries of every row inside [lo, hi): one warp per row, order kept
__global__ void __launch_bounds__(256)
k_stripe_count(const int64_t* __restrict__ rp, const int* __restrict__ col, int m, int lo, int hi,
               long long* __restrict__ cnt) {
  const int row = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (row >= m) return;
  int c = 0;
  for (int64_t p = rp[row] + lane; p < rp[row + 1]; p += 32) { const int j = col[p]; c += (j >= lo && j < hi) ? 1 : 0; }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULLW, c, o);
  if (lane == 0) cnt[row] = c;
}

__global__ void __launch_bounds__(256)
k_stripe_fill(const int64_t* __restrict__ rp, const int* __restrict__ col, const double* __restrict__ val,
              int m, int lo, int hi, const int64_t* __restrict__ orp, int* __restrict__ ocol,
              double* __restrict__ oval) {
  const int row = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (row >= m) return;
  int64_t out = orp[row];
  const int64_t s = rp[row], e = rp[row + 1];
  for (int64_t p0 = s; p0 < e; p0 += 32) {
    const int64_t p = p0 + lane;
    const int j = p < e ? col[p] : -1;
    const bool in = j >= lo && j < hi;
    const unsigned msk = __ballot_sync(FULLW, in);
    if (in) {
      const int64_t o = out + __popc(msk & ((1u << lane) - 1u));
      ocol[o] = j - lo;               // local column of the stripe (PCSR.cc:38)
      oval[o] = val[p];
    }
    out += __popc(msk);
  }
}

// row i of the result = row i of block 0, then of block 1 (columns shifted by block 0's width), ...
__global__ void __launch_bounds__(256)
k_concat_cols_rows(const int64_t* __restrict__ rp, const int* __restrict__ col, const double* __restrict__ val,
                   int m, int shift, const int64_t* __restrict__ orp, const long long* __restrict__ before,
                   int* __restrict__ ocol, double* __restrict__ oval) {
  const int row = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (row >= m) return;
  const int64_t s = rp[row], e = rp[row + 1], o = orp[row] + before[row];
  for (int64_t p = s + lane; p < e; p += 32) { ocol[o + (p - s)] = col[p] + shift; oval[o + (p - s)] = val[p]; }
}

__global__ void __launch_bounds__(256)
k_add_lengths(const int64_t* __restrict__ rp, int m, long long* __restrict__ acc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) acc[i] += rp[i + 1] - rp[i];
}

int scan_to_rowptr(const long long* d_cnt, int m, int64_t* d_rp) {
  Ctx& c = ctx();
  Temps T;
  void* tmp = nullptr;
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, d_cnt, (long long*)d_rp, m + 1, c.stream);
  B200_CUDA(T.alloc((char**)&tmp, tb ? tb : 1));
  cub::DeviceScan::ExclusiveSum(tmp, tb, d_cnt, (long long*)d_rp, m + 1, c.stream);
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

}  // namespace

int column_stripe_device(const DevCSR& B, int lo, int hi, DevCSR* out) {
  Ctx& c = ctx();
  cudaStream_t st = c.stream;
  const int m = B.rows;
  *out = DevCSR();
  Temps T;
  DevCSR d;
  d.rows = m; d.cols = hi - lo;
  long long* d_cnt = nullptr;
  B200_CUDA(T.alloc(&d_cnt, (size_t)m + 1));
  B200_CUDA(cudaMemsetAsync(d_cnt, 0, ((size_t)m + 1) * sizeof(long long), st));
  B200_CUDA(T.alloc(&d.rowptr, (size_t)m + 1));
  const unsigned grid = (unsigned)(((long long)m * 32 + 255) / 256);
  if (m) k_stripe_count<<<grid, 256, 0, st>>>(B.rowptr, B.col, m, lo, hi, d_cnt);
  int rc = scan_to_rowptr(d_cnt, m, d.rowptr);
  if (rc) return rc;
  long long nnz = 0;
  B200_CUDA(cudaMemcpyAsync(&nnz, d.rowptr + m, sizeof(long long), cudaMemcpyDeviceToHost, st));
  B200_CUDA(cudaStreamSynchronize(st));
  d.nnz = nnz;
  B200_CUDA(T.alloc(&d.col, (size_t)nnz));
  B200_CUDA(T.alloc(&d.val, (size_t)nnz));
  if (m) k_stripe_fill<<<grid, 256, 0, st>>>(B.rowptr, B.col, B.val, m, lo, hi, d.rowptr, d.col, d.val);
  B200_CUDA(cudaGetLastError());
  d.sorted_rows = B.sorted_rows;
  T.keep(d.rowptr); T.keep(d.col); T.keep(d.val);
  *out = d;
  return B200_OK;
}

int concat_cols_device(const std::vector<DevCSR>& blocks, DevCSR* out) {
  Ctx& c = ctx();
  cudaStream_t st = c.stream;
  *out = DevCSR();
  const int m = blocks[0].rows;
  Temps T;
  DevCSR d;
  d.rows = m;
  long long *d_len = nullptr, *d_before = nullptr;
  B200_CUDA(T.alloc(&d_len, (size_t)m + 1));
  B200_CUDA(T.alloc(&d_before, (size_t)m + 1));
  B200_CUDA(cudaMemsetAsync(d_len, 0, ((size_t)m + 1) * sizeof(long long), st));
  bool sorted = true;
  for (const DevCSR& b : blocks) {
    if (m) k_add_lengths<<<(m + 255) / 256, 256, 0, st>>>(b.rowptr, m, d_len);
    d.cols += b.cols;
    d.nnz += b.nnz;
    sorted = sorted && b.sorted_rows;
  }
  B200_CUDA(T.alloc(&d.rowptr, (size_t)m + 1));
  int rc = scan_to_rowptr(d_len, m, d.rowptr);
  if (rc) return rc;
  B200_CUDA(T.alloc(&d.col, (size_t)d.nnz));
  B200_CUDA(T.alloc(&d.val, (size_t)d.nnz));
  B200_CUDA(cudaMemsetAsync(d_before, 0, ((size_t)m + 1) * sizeof(long long), st));
  int shift = 0;
  const unsigned grid = (unsigned)(((long long)m * 32 + 255) / 256);
  for (const DevCSR& b : blocks) {
    if (m) {
      k_concat_cols_rows<<<grid, 256, 0, st>>>(b.rowptr, b.col, b.val, m, shift, d.rowptr, d_before, d.col, d.val);
      k_add_lengths<<<(m + 255) / 256, 256, 0, st>>>(b.rowptr, m, d_before);
    }
    shift += b.cols;
  }
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaStreamSynchronize(st));
  d.sorted_rows = sorted;   // stripes are in ascending column order
  T.keep(d.rowptr); T.keep(d.col); T.keep(d.val);
  *out = d;
  return B200_OK;
}

}  // namespace b200
