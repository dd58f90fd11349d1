// tiles.cu — one rMCL step through a BOUNDED arena (row tiles), and device-side cut points.
//
// Reference behaviour being replaced: static_omp_CSR_RMCL_OneStep,
// nlibs/static_omp_csr_kernel.cc:208-284 — it allocates the UNPRUNED product of the whole step
// (`JC`, `C` at :241-242) and compacts afterwards (omp_matrix_relocation,
// nlibs/omp_csr_kernel.cc:201-236).  At BASELINE scale that buffer does not exist: the second
// iteration on the 4 M-vertex planted-partition graph has 7.6e9 unpruned entries (91 GB), R-MAT
// scale 22 more than HBM holds even sharded eight ways.  A row of the result depends on one row
// of A only, so the step is run over consecutive row TILES whose intermediate products — an
// upper bound of their unpruned entries, known from the flops analysis
// (dynamic_omp_CSR_flops, nlibs/flops_csr_kernel.cc:14-31) before anything is allocated — fit a
// fixed arena; each tile is pruned and compacted on its own and the pruned tiles are
// concatenated.  Cut points are found on the device (the prefix never travels to the host).
#include <algorithm>
#include <vector>
#include "common.cuh"

namespace b200 {
namespace {

// cut points of [0, m) such that every tile holds at most `cap` products (a single heavier row
// is a tile of its own): cuts[0] = 0 < cuts[1] < ... < cuts[*ncuts] = m; at most max_cuts tiles
// (the last one takes the rest — the caller sizes max_cuts so that this cannot happen).
__global__ void k_tile_cuts(const int64_t* __restrict__ prefix, int m, long long cap, int max_cuts,
                            int* __restrict__ cuts, int* __restrict__ ncuts) {
  if (threadIdx.x || blockIdx.x) return;
  int pos = 0, k = 0;
  cuts[0] = 0;
  while (pos < m && k + 1 < max_cuts) {
    const long long target = prefix[pos] + cap;
    int lo = pos + 1, hi = m;  // largest e in [pos+1, m] with prefix[e] <= target, at least pos+1
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (prefix[mid] <= target) lo = mid; else hi = mid - 1;
    }
    pos = lo;
    cuts[++k] = pos;
  }
  if (pos < m) cuts[++k] = m;
  *ncuts = k;
}

// arrayEqualPartition64 (nlibs/tools/util.cc:123-135) on a device-resident prefix: same
// arithmetic as b200_equal_partition64, one thread (the cut points depend on each other)
__global__ void k_equal_partition(const int64_t* __restrict__ prefix, int n, int nparts,
                                  int* __restrict__ ends) {
  if (threadIdx.x || blockIdx.x) return;
  const long long total = prefix[n];
  const long long chunk = (total + nparts - 1) / nparts;
  ends[0] = 0;
  int now = 0;
  for (int t = 0; t + 1 < nparts; ++t) {
    const long long target = min((long long)(t + 1) * chunk, total);
    int lo = now, hi = n + 1;  // upper_bound(prefix + now, prefix + n + 1, target)
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (prefix[mid] <= target) lo = mid + 1; else hi = mid;
    }
    int e = max(lo - 1, now + 1);
    e = min(e, n);
    ends[t + 1] = e;
    now = e;
  }
  ends[nparts] = n;
}

__global__ void k_rebase(const int64_t* __restrict__ in, int64_t* __restrict__ out, long long n,
                         long long add) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] - in[0] + add;
}

void release_csr(DevCSR& d) {
  dfree(d.rowptr); dfree(d.col); dfree(d.val);
  d = DevCSR();
}

}  // namespace

int equal_partition_device(const int64_t* d_prefix, int n, int nparts, int* h_ends) {
  Ctx& c = ctx();
  int* d_ends = nullptr;
  Temps T;
  B200_CUDA(T.alloc(&d_ends, (size_t)nparts + 1));
  k_equal_partition<<<1, 32, 0, c.stream>>>(d_prefix, n, nparts, d_ends);
  B200_CUDA(cudaMemcpyAsync(h_ends, d_ends, ((size_t)nparts + 1) * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  B200_CUDA(cudaStreamSynchronize(c.stream));
  return B200_OK;
}

int concat_rows_device(const std::vector<DevCSR>& blocks, int cols, DevCSR* out) {
  Ctx& c = ctx();
  long long rows = 0, nnz = 0;
  for (const DevCSR& b : blocks) { rows += b.rows; nnz += b.nnz; }
  DevCSR d;
  d.rows = (int)rows; d.cols = cols; d.nnz = nnz;
  Temps T;
  B200_CUDA(T.alloc(&d.rowptr, (size_t)rows + 1));
  B200_CUDA(T.alloc(&d.col, (size_t)nnz));
  B200_CUDA(T.alloc(&d.val, (size_t)nnz));
  if (blocks.empty()) B200_CUDA(cudaMemsetAsync(d.rowptr, 0, sizeof(int64_t), c.stream));
  long long r0 = 0, z0 = 0;
  for (const DevCSR& s : blocks) {
    k_rebase<<<(unsigned)((s.rows + 1 + 255) / 256), 256, 0, c.stream>>>(s.rowptr, d.rowptr + r0, s.rows + 1, z0);
    if (s.nnz) {
      B200_CUDA(cudaMemcpyAsync(d.col + z0, s.col, (size_t)s.nnz * sizeof(int), cudaMemcpyDeviceToDevice, c.stream));
      B200_CUDA(cudaMemcpyAsync(d.val + z0, s.val, (size_t)s.nnz * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
    }
    r0 += s.rows;
    z0 += s.nnz;
  }
  B200_CUDA(cudaGetLastError());
  T.keep(d.rowptr); T.keep(d.col); T.keep(d.val);
  *out = d;
  return B200_OK;
}

// products a tile may hold: the arena takes 12 bytes per unpruned entry, and a tile's products
// bound its unpruned entries.  B200_ARENA_ENTRIES overrides (tests force tiling on small inputs).
static long long arena_budget_entries() {
  Ctx& c = ctx();
  if (c.tun.arena_entries > 0) return c.tun.arena_entries;
  if (c.arena_budget == 0) {
    // blocks an earlier call freed (a 100 GB product, say) sit in the stream-ordered pool: give
    // them back first, or they would count as used
    cudaStreamSynchronize(c.stream);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, c.device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    // Give the pool its working set in ONE piece before the arena takes its share.  The step's
    // temporaries and row blocks (a few GB) are then carved from this block; without it the pool
    // grows in driver calls in the middle of the loop — 8 ms each as a rule, but one
    // cudaMallocAsync of 110 MB was measured at 702 ms next to a 53 GB arena (R-MAT scale 20,
    // B200_PROF=1), 11 % of the whole loop.  (The pool keeps what it gets: release threshold = max.)
    {
      const size_t reserve = std::min<size_t>(free_b / 10, (size_t)16 << 30);
      void* p = nullptr;
      if (reserve > 0 && cudaMallocAsync(&p, reserve, c.stream) == cudaSuccess) {
        cudaFreeAsync(p, c.stream);
        cudaStreamSynchronize(c.stream);
        free_b -= reserve;
      } else {
        cudaGetLastError();
      }
    }
    // 40 % of what is free now plus what the arena already holds; the operands, the result and
    // the pipeline's scratch need the rest
    c.arena_budget = (long long)(0.40 * (double)(free_b + c.arena_cap * 12)) / 12;
    if (c.arena_budget < (1 << 20)) c.arena_budget = 1 << 20;
  }
  return c.arena_budget;
}

int rmcl_step_device(const DevCSR& A, const DevCSR& B, int row_lo, int row_hi, DevCSR* C,
                     double* chaos, b200_stats* stats) {
  Ctx& c = ctx();
  cudaStream_t st = c.stream;
  const int m = row_hi - row_lo;
  *C = DevCSR();
  if (stats) memset(stats, 0, sizeof(*stats));
  const long long cap = arena_budget_entries();
  // ---- products prefix of the block and the tile cuts, on the device
  Temps T;
  int64_t* d_prefix = nullptr;
  int *d_cuts = nullptr, *d_ncuts = nullptr;
  B200_CUDA(T.alloc(&d_prefix, (size_t)m + 1));
  int rc = flops_prefix_device(A, B, row_lo, row_hi, d_prefix);
  if (rc) return rc;
  long long P = 0;
  B200_CUDA(cudaMemcpyAsync(&P, d_prefix + m, sizeof(long long), cudaMemcpyDeviceToHost, st));
  B200_CUDA(cudaStreamSynchronize(st));
  if (P <= cap || m <= 1) return run_pipeline(A, B, row_lo, row_hi, MODE_RMCL, C, chaos, stats);
  // every tile but a single heavy row holds more than cap / 2 products... no: a tile closes at the
  // last row that fits, so two consecutive tiles hold more than cap together: <= 2 P / cap + 2
  const int max_cuts = (int)std::min<long long>(2 * (P / cap) + 4, (long long)m + 1);
  B200_CUDA(T.alloc(&d_cuts, (size_t)max_cuts + 1));
  B200_CUDA(T.alloc(&d_ncuts, 1));
  k_tile_cuts<<<1, 32, 0, st>>>(d_prefix, m, cap, max_cuts, d_cuts, d_ncuts);
  std::vector<int> cuts((size_t)max_cuts + 1);
  int ncuts = 0;
  B200_CUDA(cudaMemcpyAsync(&ncuts, d_ncuts, sizeof(int), cudaMemcpyDeviceToHost, st));
  B200_CUDA(cudaMemcpyAsync(cuts.data(), d_cuts, ((size_t)max_cuts + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
  B200_CUDA(cudaStreamSynchronize(st));
  // ---- a column-sorted copy of B, made once for all tiles when the bitmap kernels will want one
  DevCSR Bs = B;
  if (!B.sorted_rows && B.nnz > 0) {
    rc = sorted_copy_of(B, &Bs);
    if (rc) return rc;
    T.adopt(Bs.col);
    T.adopt(Bs.val);
  }
  // ---- the tiles
  std::vector<DevCSR> blocks;
  blocks.reserve((size_t)ncuts);
  double ch_max = 0.0;
  b200_stats ts;
  for (int k = 0; k < ncuts && !rc; ++k) {
    DevCSR blk;
    double ch = 0.0;
    rc = run_pipeline(A, B, row_lo + cuts[k], row_lo + cuts[k + 1], MODE_RMCL, &blk, &ch,
                      stats ? &ts : nullptr, B.sorted_rows ? nullptr : &Bs);
    if (rc) { release_csr(blk); break; }
    blocks.push_back(blk);
    ch_max = std::max(ch_max, ch);
    if (stats) {
      stats->ms_total += ts.ms_total; stats->ms_flops += ts.ms_flops; stats->ms_symbolic += ts.ms_symbolic;
      stats->ms_numeric += ts.ms_numeric; stats->ms_other += ts.ms_other; stats->products += ts.products;
      stats->nnz_out += ts.nnz_out; stats->nnz_unpruned += ts.nnz_unpruned; stats->launches += ts.launches;
      stats->part_kernel = ts.part_kernel; stats->part_count = ts.part_count;
      for (int b = 0; b < 16; ++b) {
        stats->bins_rows[b] += ts.bins_rows[b];
        stats->ms_sym_bin[b] += ts.ms_sym_bin[b]; stats->ms_num_bin[b] += ts.ms_num_bin[b];
        stats->sym_bin_rows[b] += ts.sym_bin_rows[b]; stats->sym_bin_products[b] += ts.sym_bin_products[b];
        stats->sym_bin_nnzA[b] += ts.sym_bin_nnzA[b]; stats->num_bin_products[b] += ts.num_bin_products[b];
        stats->num_bin_nnzA[b] += ts.num_bin_nnzA[b]; stats->num_bin_nnzC[b] += ts.num_bin_nnzC[b];
      }
    }
  }
  if (!rc) rc = concat_rows_device(blocks, B.cols, C);
  for (DevCSR& b : blocks) release_csr(b);
  if (rc) return rc;
  C->sorted_rows = false;
  if (stats) stats->row_tiles = ncuts;
  if (chaos) *chaos = ch_max;
  B200_CUDA(cudaStreamSynchronize(st));
  return B200_OK;
}

}  // namespace b200
