// spgemm.cu — flops analysis, binning, symbolic and numeric Gustavson SpGEMM kernels and the
// fused rMCL epilogue, hand-written for sm_100a.
//
// Reference behaviour being replaced (paths relative to the reference root):
//   flops analysis   nlibs/flops_csr_kernel.cc:14-31      (dynamic_omp_CSR_flops)
//   symbolic row     nlibs/cpu_csr_kernel.h:234-262       (cRowiCount, dense bool map)
//   numeric row      nlibs/cpu_csr_kernel.h:134-188       (indexProcessCRowI, dense index)
//   rMCL epilogue    nlibs/static_omp_csr_kernel.cc:264-271 + nlibs/tools/util.cc:4-69
//   compaction       nlibs/omp_csr_kernel.cc:201-236      (omp_matrix_relocation)
//
// Design (DESIGN.md has the full account):
//  * rows are binned twice: for the symbolic pass by intermediate products P_i (an upper bound
//    of nnz(C_i)), for the numeric pass by the exact nnz(C_i);
//  * small rows: one warp per row, a warp-private hash table in shared memory that maps
//    column -> slot in FIRST-TOUCH order.  The warp walks the A entries of the row one at a
//    time and spreads its lanes over one B row, so every C entry accumulates its products in
//    ascending A-entry order with separately rounded multiply and add — the same order and
//    rounding as indexProcessCRowI, hence bit-identical values.  Because a B row has unique
//    columns, a step never holds two products of the same column: no atomics are needed,
//    insertion is a warp-synchronous write-then-verify;
//  * large rows (more distinct columns than a shared-memory table holds): one CTA per row, a
//    column bitmap in shared memory built with shared atomics in the symbolic pass and kept in
//    HBM; the numeric pass turns it into rank = popcount-prefix, so the output row is produced
//    directly in ascending column order and products are reduced with fp64 RED into the row's
//    final location (order of additions not fixed: values agree to rounding, not bitwise);
//  * output columns ascend in every row (in-shared-memory bitonic sort of packed
//    (column,slot) keys for the hash bins; by construction for the bitmap bin);
//  * rMCL: inflation, row max/sum, threshold, prune, normalise and the chaos term are fused
//    behind the numeric row while it is still on chip; pruned rows go to a bump-allocated
//    arena and are gathered into the final CSR after a scan of the kept counts, so the
//    unpruned product is never materialised in CSR form.
#include <cub/cub.cuh>
#include <algorithm>
#include <vector>
#include "common.cuh"

namespace b200 {

namespace {

constexpr int EMPTY = -1;
constexpr unsigned FULL = 0xffffffffu;

// symbolic bins (by products P_i)
constexpr int SB_NONE = 0;    // nnz known without hashing (P_i == 0 or a single A entry)
constexpr int SB_W256 = 1;    // P <= 128
constexpr int SB_W1K = 2;     // P <= 512
constexpr int SB_W4K = 3;     // P <= 2048
constexpr int SB_W16K = 4;    // P <= 8192
constexpr int SB_BITMAP = 5;  // larger
// numeric bins (by nnz(C_i))
constexpr int NB_NONE = 0;    // empty row
constexpr int NB_W64 = 1;
constexpr int NB_W256 = 2;
constexpr int NB_W1K = 3;
constexpr int NB_W2K = 4;
constexpr int NB_BITMAP = 5;

__host__ __device__ inline int sym_bin_of(long long P, int annz) {
  if (P == 0 || annz <= 1) return SB_NONE;
  if (P <= 128) return SB_W256;
  if (P <= 512) return SB_W1K;
  if (P <= 2048) return SB_W4K;
  if (P <= 8192) return SB_W16K;
  return SB_BITMAP;
}
__host__ __device__ inline int num_bin_of(int cnt) {
  if (cnt == 0) return NB_NONE;
  if (cnt <= 64) return NB_W64;
  if (cnt <= 256) return NB_W256;
  if (cnt <= 1024) return NB_W1K;
  if (cnt <= 2048) return NB_W2K;
  return NB_BITMAP;
}

// Fibonacci hashing: the slot is the TOP log2(H) bits of c * 2^32/phi.  (Taking low bits of a
// product would make columns that differ by a multiple of H collide — exactly what a stencil on
// a power-of-two grid produces.)
template <int H>
__device__ __forceinline__ unsigned hash_col(int c) {
  static_assert((H & (H - 1)) == 0 && H >= 2, "table size must be a power of two");
  return ((unsigned)c * 0x9E3779B1u) >> (32 - __builtin_ctz(H));
}
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}
__device__ __forceinline__ long long shfl64(long long v, int src) {
  int lo = __shfl_sync(FULL, (int)(v & 0xffffffffLL), src);
  int hi = __shfl_sync(FULL, (int)(v >> 32), src);
  return ((long long)hi << 32) | (unsigned)lo;
}
__device__ __forceinline__ double shfld(double v, int src) {
  return __longlong_as_double(shfl64(__double_as_longlong(v), src));
}
__device__ __forceinline__ double shfld_xor(double v, int m) {
  long long x = __double_as_longlong(v);
  int lo = __shfl_xor_sync(FULL, (int)(x & 0xffffffffLL), m);
  int hi = __shfl_xor_sync(FULL, (int)(x >> 32), m);
  return __longlong_as_double(((long long)hi << 32) | (unsigned)lo);
}
// fixed-shape warp reductions (lane-strided partials, then xor tree 16..1): deterministic
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, shfld_xor(v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, shfld_xor(v, o));
  return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// computeThreshold (nlibs/tools/util.cc:4-9) with the reference's operation order and no
// fused multiply-add: ((0.9*avg) * (1 - (2*(max-avg)))), floor 1e-7, cap at max.
__device__ __forceinline__ double compute_threshold(double avg, double mx) {
  double t = __dmul_rn(__dmul_rn(0.90, avg), __dsub_rn(1.0, __dmul_rn(2.0, __dsub_rn(mx, avg))));
  t = (t > 1.0e-7) ? t : 1.0e-7;
  t = (t > mx) ? mx : t;
  return t;
}

// ------------------------------------------------------------------------------------------
// flops analysis: P_i = sum_{j in A_i} nnz(B_j)  (flops_csr_kernel.cc:14-31)
// one thread per row; also emits the symbolic bin and, for rows that need no hashing, nnz(C_i).
__global__ void __launch_bounds__(256)
k_row_flops(const int64_t* __restrict__ Arp, const int* __restrict__ Acol,
            const int64_t* __restrict__ Brp, int row_lo, int m, long long* __restrict__ flops,
            unsigned char* __restrict__ sbin, int* __restrict__ rownnz) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int64_t a0 = Arp[row_lo + i], a1 = Arp[row_lo + i + 1];
  long long f = 0;
  for (int64_t p = a0; p < a1; ++p) {
    int j = __ldg(Acol + p);
    f += __ldg(Brp + j + 1) - __ldg(Brp + j);
  }
  flops[i] = f;
  int b = sym_bin_of(f, (int)(a1 - a0));
  sbin[i] = (unsigned char)b;
  if (b == SB_NONE) rownnz[i] = (int)f;  // 0, or the length of the single B row
}

__global__ void __launch_bounds__(256)
k_num_bins(const int* __restrict__ rownnz, int m, unsigned char* __restrict__ nbin) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) nbin[i] = (unsigned char)num_bin_of(rownnz[i]);
}

// histogram of bin ids (<= 16 bins)
__global__ void __launch_bounds__(256)
k_bin_hist(const unsigned char* __restrict__ bin, int m, int* __restrict__ hist) {
  __shared__ int s[16];
  if (threadIdx.x < 16) s[threadIdx.x] = 0;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
    atomicAdd(&s[bin[i]], 1);
  __syncthreads();
  if (threadIdx.x < 16 && s[threadIdx.x]) atomicAdd(&hist[threadIdx.x], s[threadIdx.x]);
}

// per-bin totals for b200_stats: agg[b*3+0] += a[i], +1 += b[i], +2 += c[i] for bin[i]==b
__global__ void __launch_bounds__(256)
k_bin_aggregate(const unsigned char* __restrict__ bin, int m, const int64_t* __restrict__ Arp,
                int row_lo, const long long* __restrict__ flops, const int* __restrict__ rownnz,
                unsigned long long* __restrict__ agg) {
  __shared__ unsigned long long s[48];
  if (threadIdx.x < 48) s[threadIdx.x] = 0ull;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    const int b = bin[i];
    atomicAdd(&s[b * 3 + 0], (unsigned long long)flops[i]);
    atomicAdd(&s[b * 3 + 1], (unsigned long long)(Arp[row_lo + i + 1] - Arp[row_lo + i]));
    if (rownnz) atomicAdd(&s[b * 3 + 2], (unsigned long long)rownnz[i]);
  }
  __syncthreads();
  if (threadIdx.x < 48 && s[threadIdx.x]) atomicAdd(&agg[threadIdx.x], s[threadIdx.x]);
}

// scatter row ids into per-bin lists (cursor[b] starts at the bin's offset)
__global__ void __launch_bounds__(256)
k_bin_scatter(const unsigned char* __restrict__ bin, int m, int* __restrict__ cursor,
              int* __restrict__ list) {
  __shared__ int s_cnt[16];
  __shared__ int s_base[16];
  if (threadIdx.x < 16) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int b = -1, pos = 0;
  if (i < m) { b = bin[i]; pos = atomicAdd(&s_cnt[b], 1); }
  __syncthreads();
  if (threadIdx.x < 16 && s_cnt[threadIdx.x])
    s_base[threadIdx.x] = atomicAdd(&cursor[threadIdx.x], s_cnt[threadIdx.x]);
  __syncthreads();
  if (i < m) list[s_base[b] + pos] = i;
}

// ------------------------------------------------------------------------------------------
// Warp-synchronous insertion into a warp-private open-addressing table (linear probing).
// All 32 lanes call it together; keys offered in one call are pairwise distinct (they come
// from one B row), so a lane can only lose a slot to a DIFFERENT key.  Returns true if the
// key was new; `h` is left at the key's slot.
template <int H>
__device__ __forceinline__ bool warp_find_or_insert(int* keys, int c, bool active, unsigned& h) {
  constexpr unsigned mask = H - 1;
  bool done = !active, isnew = false;
  h = hash_col<H>(c);
  while (true) {
    bool attempt = false;
    if (!done) {
      int k = keys[h];
      if (k == c) done = true;
      else if (k == EMPTY) { keys[h] = c; attempt = true; }
      else h = (h + 1) & mask;
    }
    __syncwarp();
    if (attempt) {
      if (keys[h] == c) { done = true; isnew = true; }
      else h = (h + 1) & mask;
    }
    if (__all_sync(FULL, done)) break;
  }
  return isnew;
}

// ------------------------------------------------------------------------------------------
// symbolic, one warp per row, H key slots per warp (cRowiCount, cpu_csr_kernel.h:234-262)
template <int H>
__global__ void __launch_bounds__(256)
k_sym_warp(const int* __restrict__ list, int count, int row_lo,
           const int64_t* __restrict__ Arp, const int* __restrict__ Acol,
           const int64_t* __restrict__ Brp, const int* __restrict__ Bcol,
           int* __restrict__ rownnz) {
  extern __shared__ int smem_i[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int idx = blockIdx.x * (blockDim.x >> 5) + warp;
  if (idx >= count) return;
  const int i = list[idx];
  int* keys = smem_i + warp * H;
  for (int k = lane; k < H; k += 32) keys[k] = EMPTY;
  __syncwarp();
  const int64_t a0 = Arp[row_lo + i], a1 = Arp[row_lo + i + 1];
  int cnt = 0;
  for (int64_t base = a0; base < a1; base += 32) {
    const int64_t p = base + lane;
    long long bs = 0, be = 0;
    if (p < a1) { int j = __ldg(Acol + p); bs = __ldg(Brp + j); be = __ldg(Brp + j + 1); }
    const int nn = (int)min((int64_t)32, a1 - base);
    for (int t = 0; t < nn; ++t) {
      const long long s = shfl64(bs, t), e = shfl64(be, t);
      for (long long q0 = s; q0 < e; q0 += 32) {
        const long long q = q0 + lane;
        const bool act = q < e;
        const int c = act ? __ldg(Bcol + q) : 0;
        unsigned h;
        cnt += warp_find_or_insert<H>(keys, c, act, h) ? 1 : 0;
      }
    }
  }
  cnt = warp_sum_int(cnt);
  if (lane == 0) rownnz[i] = cnt;
}

// ------------------------------------------------------------------------------------------
// in-shared-memory bitonic sort of n2 (power of two) packed keys by one warp
__device__ __forceinline__ void warp_bitonic_sort(unsigned long long* sb, int n2, int lane) {
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (n2 >> 1); t += 32) {
        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int hi = lo | j;
        const bool up = (lo & k) == 0;
        const unsigned long long a = sb[lo], b = sb[hi];
        if ((a > b) == up) { sb[lo] = b; sb[hi] = a; }
      }
      __syncwarp();
    }
  }
}

struct RmclOut {          // where a fused rMCL row goes
  int* arena_col;
  double* arena_val;
  unsigned long long* cursor;   // bump allocator over the arena
  long long* row_off;           // [m] arena offset of the row
  int* row_kept;                // [m] kept entries
  unsigned long long* chaos_bits;  // max over rows of (max - sum sq), as ordered bits
};

// numeric, one warp per row, CAP output entries per warp (indexProcessCRowI,
// cpu_csr_kernel.h:134-188) + sort (+ fused rMCL epilogue when RMCL).
// shared memory per warp (24*CAP bytes):
//   vals  fp64[CAP]      slot -> accumulated value (first-touch order)
//   keys  int[2*CAP]     hash keys; later reused as the u64[CAP] sort buffer
//   cols  int[CAP]       slot -> column           \ later reused together as fp64[CAP]:
//   slot  ushort[2*CAP]  hash slot -> output slot / squared values in sorted order
template <int CAP, bool RMCL>
__global__ void __launch_bounds__(256)
k_num_warp(const int* __restrict__ list, int count, int row_lo,
           const int64_t* __restrict__ Arp, const int* __restrict__ Acol,
           const double* __restrict__ Aval, const int64_t* __restrict__ Brp,
           const int* __restrict__ Bcol, const double* __restrict__ Bval,
           const int64_t* __restrict__ Crp, int* __restrict__ Ccol, double* __restrict__ Cval,
           RmclOut ro) {
  constexpr int H = 2 * CAP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int idx = blockIdx.x * (blockDim.x >> 5) + warp;
  if (idx >= count) return;
  const int i = list[idx];
  unsigned char* wbase = smem_raw + (size_t)warp * (24 * CAP);
  double* vals = (double*)wbase;
  int* keys = (int*)(wbase + 8 * CAP);
  int* cols = (int*)(wbase + 16 * CAP);
  unsigned short* slot = (unsigned short*)(wbase + 20 * CAP);
  for (int k = lane; k < H; k += 32) keys[k] = EMPTY;
  __syncwarp();

  const int64_t a0 = Arp[row_lo + i], a1 = Arp[row_lo + i + 1];
  int cnt = 0;
  for (int64_t base = a0; base < a1; base += 32) {
    const int64_t p = base + lane;
    long long bs = 0, be = 0;
    double av = 0.0;
    if (p < a1) {
      int j = __ldg(Acol + p);
      av = __ldg(Aval + p);
      bs = __ldg(Brp + j);
      be = __ldg(Brp + j + 1);
    }
    const int nn = (int)min((int64_t)32, a1 - base);
    for (int t = 0; t < nn; ++t) {
      const long long s = shfl64(bs, t), e = shfl64(be, t);
      const double a = shfld(av, t);
      for (long long q0 = s; q0 < e; q0 += 32) {
        const long long q = q0 + lane;
        const bool act = q < e;
        int c = 0;
        double prod = 0.0;
        if (act) { c = __ldg(Bcol + q); prod = __dmul_rn(a, __ldg(Bval + q)); }
        unsigned h;
        const bool isnew = warp_find_or_insert<H>(keys, c, act, h);
        const unsigned newmask = __ballot_sync(FULL, isnew);
        if (isnew) {
          const int sl = cnt + __popc(newmask & lanemask_lt());
          slot[h] = (unsigned short)sl;
          cols[sl] = c;
          vals[sl] = prod;
        } else if (act) {
          const int sl = slot[h];
          vals[sl] = __dadd_rn(vals[sl], prod);
        }
        cnt += __popc(newmask);
        __syncwarp();
      }
    }
  }

  if (!RMCL) {
    // ---- sort (column, slot) pairs ascending by column, then write the row
    unsigned long long* sb = (unsigned long long*)keys;
    int n2 = 1;
    while (n2 < cnt) n2 <<= 1;
    for (int k = lane; k < n2; k += 32)
      sb[k] = (k < cnt) ? (((unsigned long long)(unsigned)cols[k] << 32) | (unsigned)k) : ~0ull;
    __syncwarp();
    warp_bitonic_sort(sb, n2, lane);
    const int64_t ob = Crp[i];
    for (int k = lane; k < cnt; k += 32) {
      const unsigned long long e = sb[k];
      Ccol[ob + k] = (int)(e >> 32);
      Cval[ob + k] = vals[(unsigned)(e & 0xffffffffu)];
    }
    return;
  }

  // ---- fused rMCL epilogue in the reference's own order.  cols[]/vals[] hold the row in
  // FIRST-TOUCH order, exactly the layout static_omp_CSR_RMCL_OneStep works on
  // (nlibs/static_omp_csr_kernel.cc:256-271), so the sequential sums below reproduce
  // arrayMaxSum (util.cc:21-31) and arrayThreshPruneNormalize (util.cc:47-69) bit for bit.
  // Every lane runs the same serial chain on broadcast shared-memory reads (no divergence,
  // no shuffle); other warps hide its latency.
  for (int k = lane; k < cnt; k += 32) { const double v = vals[k]; vals[k] = __dmul_rn(v, v); }
  __syncwarp();
  double rmax = 0.0, rsum = 0.0;
#pragma unroll 4
  for (int k = 0; k < cnt; ++k) {
    const double v = vals[k];
    if (rmax < v) rmax = v;
    rsum = __dadd_rn(rsum, v);
  }
  const double thresh = compute_threshold(__ddiv_rn(rsum, (double)cnt), rmax);
  double ksum = 0.0;
  int kept = 0;
#pragma unroll 4
  for (int k = 0; k < cnt; ++k) {
    const double v = vals[k];
    const bool keep = v >= thresh;
    ksum = keep ? __dadd_rn(ksum, v) : ksum;
    kept += keep ? 1 : 0;
  }
  unsigned long long off = 0;
  if (lane == 0) off = atomicAdd(ro.cursor, (unsigned long long)kept);
  off = (unsigned long long)shfl64((long long)off, 0);
  double sq_p = 0.0;
  int written = 0;
  for (int k0 = 0; k0 < cnt; k0 += 32) {
    const int k = k0 + lane;
    const double v = (k < cnt) ? vals[k] : 0.0;
    const bool keep = (k < cnt) && (v >= thresh);
    const unsigned km = __ballot_sync(FULL, keep);
    if (keep) {
      const double w = __ddiv_rn(v, ksum);
      const long long o = (long long)off + written + __popc(km & lanemask_lt());
      ro.arena_col[o] = cols[k];
      ro.arena_val[o] = w;
      sq_p = __dadd_rn(sq_p, __dmul_rn(w, w));
    }
    written += __popc(km);
  }
  const double sq = warp_sum(sq_p);
  if (lane == 0) {
    ro.row_off[i] = (long long)off;
    ro.row_kept[i] = kept;
    double ch = (kept > 0) ? __dsub_rn(__ddiv_rn(rmax, ksum), sq) : 0.0;
    if (ch < 0.0) ch = 0.0;
    atomicMax(ro.chaos_bits, (unsigned long long)__double_as_longlong(ch));
  }
}

// ------------------------------------------------------------------------------------------
// block-wide helpers for the bitmap (large-row) kernels; blockDim.x == BT
template <int BT>
__device__ __forceinline__ int block_sum_int(int v, int* s_red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum_int(v);
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  int t = 0;
#pragma unroll
  for (int w = 0; w < BT / 32; ++w) t += s_red[w];
  return t;
}
// exclusive scan of one int per thread; returns the thread's exclusive prefix, *total = sum
template <int BT>
__device__ __forceinline__ int block_excl_scan(int v, int* s_red, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(FULL, inc, o);
    if (lane >= o) inc += y;
  }
  __syncthreads();
  if (lane == 31) s_red[warp] = inc;
  __syncthreads();
  int wbase = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < BT / 32; ++w) {
    int x = s_red[w];
    if (w < warp) wbase += x;
    tot += x;
  }
  *total = tot;
  return wbase + inc - v;
}
// deterministic block sum / max of doubles: per-thread partial -> warp xor tree -> ordered
// sum over warps
template <int BT>
__device__ __forceinline__ double block_sum_d(double v, double* s_redd) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) s_redd[warp] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < BT / 32; ++w) t = __dadd_rn(t, s_redd[w]);
  return t;
}
template <int BT>
__device__ __forceinline__ double block_max_d(double v, double* s_redd) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) s_redd[warp] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < BT / 32; ++w) t = fmax(t, s_redd[w]);
  return t;
}

// Walk the products of one A row with the whole CTA: each warp takes chunks of 32 A entries,
// loads (col, val, B row bounds) lane-parallel, then spreads its lanes over each B row.
template <typename F>
__device__ __forceinline__ void cta_for_each_product(int64_t a0, int64_t a1,
                                                     const int* __restrict__ Acol,
                                                     const double* __restrict__ Aval,
                                                     const int64_t* __restrict__ Brp,
                                                     const int* __restrict__ Bcol,
                                                     const double* __restrict__ Bval, F f) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int64_t base = a0 + (int64_t)warp * 32; base < a1; base += (int64_t)nwarps * 32) {
    const int64_t p = base + lane;
    long long bs = 0, be = 0;
    double av = 0.0;
    if (p < a1) {
      int j = __ldg(Acol + p);
      if (Aval) av = __ldg(Aval + p);
      bs = __ldg(Brp + j);
      be = __ldg(Brp + j + 1);
    }
    const int nn = (int)min((int64_t)32, a1 - base);
    for (int t = 0; t < nn; ++t) {
      const long long s = shfl64(bs, t), e = shfl64(be, t);
      const double a = shfld(av, t);
      for (long long q = s + lane; q < e; q += 32) {
        const int c = __ldg(Bcol + q);
        f(c, a, q);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// symbolic for large rows: CTA per row (persistent, dynamic row fetch), column bitmap of
// nw64 64-bit words either in shared memory (SMEM_BM) or in a per-CTA HBM scratch.  The
// bitmap is stored to `bm_store` (per listed row, nw64 words) when bm_store != nullptr.
template <int BT, bool SMEM_BM>
__global__ void __launch_bounds__(BT, 1)
k_sym_bitmap(const int* __restrict__ list, int count, int row_lo,
             const int64_t* __restrict__ Arp, const int* __restrict__ Acol,
             const int64_t* __restrict__ Brp, const int* __restrict__ Bcol, int nw64,
             unsigned long long* __restrict__ gscratch, unsigned long long* __restrict__ bm_store,
             int* __restrict__ rownnz, int* __restrict__ work_counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_red[BT / 32];
  __shared__ int s_idx;
  unsigned long long* bm =
      SMEM_BM ? (unsigned long long*)smem_raw : gscratch + (size_t)blockIdx.x * nw64;
  unsigned* bm32 = (unsigned*)bm;
  for (int w = threadIdx.x; w < nw64; w += BT) bm[w] = 0ull;
  while (true) {
    __syncthreads();
    if (threadIdx.x == 0) s_idx = atomicAdd(work_counter, 1);
    __syncthreads();
    const int idx = s_idx;
    if (idx >= count) break;
    const int i = list[idx];
    const int64_t a0 = Arp[row_lo + i], a1 = Arp[row_lo + i + 1];
    cta_for_each_product(a0, a1, Acol, (const double*)nullptr, Brp, Bcol, (const double*)nullptr,
                         [&](int c, double, long long) {
                           atomicOr(&bm32[c >> 5], 1u << (c & 31));
                         });
    __syncthreads();
    int cnt = 0;
    unsigned long long* dst = bm_store ? bm_store + (size_t)idx * nw64 : nullptr;
    for (int w = threadIdx.x; w < nw64; w += BT) {
      const unsigned long long x = bm[w];
      cnt += __popcll(x);
      if (dst) dst[w] = x;
      bm[w] = 0ull;
    }
    cnt = block_sum_int<BT>(cnt, s_red);
    if (threadIdx.x == 0) rownnz[i] = cnt;
  }
}

// numeric for large rows: rank(col) = prefix[col/64] + popc(bitmap word below col); products
// are reduced with fp64 RED into `acc` (the row's final slice of C.val for SpGEMM, a per-CTA
// scratch for rMCL, where the epilogue then runs over the scratch).
// bm_index[idx] >= 0: the row's bitmap was stored by the symbolic pass at that slot;
// < 0: rebuild it here.
template <int BT, bool SMEM_BM, bool RMCL>
__global__ void __launch_bounds__(BT, 1)
k_num_bitmap(const int* __restrict__ list, int count, int row_lo,
             const int64_t* __restrict__ Arp, const int* __restrict__ Acol,
             const double* __restrict__ Aval, const int64_t* __restrict__ Brp,
             const int* __restrict__ Bcol, const double* __restrict__ Bval, int nw64,
             unsigned long long* __restrict__ gscratch, const unsigned long long* __restrict__ bm_store,
             const int* __restrict__ bm_index, const int64_t* __restrict__ Crp,
             int* __restrict__ Ccol, double* __restrict__ Cval, int* __restrict__ scr_col,
             double* __restrict__ scr_val, long long scr_stride, RmclOut ro,
             int* __restrict__ work_counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_red[BT / 32];
  __shared__ double s_redd[BT / 32];
  __shared__ int s_idx;
  __shared__ unsigned long long s_off;
  // bitmap words then 32-bit exclusive popcount prefix per word
  unsigned long long* bm =
      SMEM_BM ? (unsigned long long*)smem_raw
              : gscratch + (size_t)blockIdx.x * ((size_t)nw64 + ((size_t)nw64 + 1) / 2);
  unsigned* pref = (unsigned*)(bm + nw64);
  unsigned* bm32 = (unsigned*)bm;
  const int chunk = (nw64 + BT - 1) / BT;  // contiguous words per thread
  while (true) {
    __syncthreads();
    if (threadIdx.x == 0) s_idx = atomicAdd(work_counter, 1);
    __syncthreads();
    const int idx = s_idx;
    if (idx >= count) break;
    const int i = list[idx];
    const int64_t a0 = Arp[row_lo + i], a1 = Arp[row_lo + i + 1];
    const int slotno = bm_index ? bm_index[idx] : -1;
    if (slotno >= 0) {
      const unsigned long long* src = bm_store + (size_t)slotno * nw64;
      for (int w = threadIdx.x; w < nw64; w += BT) bm[w] = src[w];
    } else {
      for (int w = threadIdx.x; w < nw64; w += BT) bm[w] = 0ull;
      __syncthreads();
      cta_for_each_product(a0, a1, Acol, (const double*)nullptr, Brp, Bcol,
                           (const double*)nullptr, [&](int c, double, long long) {
                             atomicOr(&bm32[c >> 5], 1u << (c & 31));
                           });
    }
    __syncthreads();
    // popcount prefix: thread t owns words [t*chunk, (t+1)*chunk)
    const int w0 = threadIdx.x * chunk, w1 = min(nw64, w0 + chunk);
    int local = 0;
    for (int w = w0; w < w1; ++w) local += __popcll(bm[w]);
    int total;
    int run = block_excl_scan<BT>(local, s_red, &total);
    for (int w = w0; w < w1; ++w) { pref[w] = (unsigned)run; run += __popcll(bm[w]); }
    const int cnt = total;
    double* acc = RMCL ? scr_val + (size_t)blockIdx.x * scr_stride : Cval + Crp[i];
    int* ocol = RMCL ? scr_col + (size_t)blockIdx.x * scr_stride : Ccol + Crp[i];
    for (int k = threadIdx.x; k < cnt; k += BT) acc[k] = 0.0;
    __syncthreads();
    // columns in ascending order straight from the bitmap
    {
      int pos = (w0 < nw64) ? (int)pref[w0] : 0;
      for (int w = w0; w < w1; ++w) {
        unsigned long long x = bm[w];
        while (x) {
          const int b = __ffsll((long long)x) - 1;
          x &= x - 1;
          ocol[pos++] = w * 64 + b;
        }
      }
    }
    cta_for_each_product(a0, a1, Acol, Aval, Brp, Bcol, Bval, [&](int c, double a, long long q) {
      const int w = c >> 6;
      const unsigned long long below = bm[w] & ((1ull << (c & 63)) - 1ull);
      const int rank = (int)pref[w] + __popcll(below);
      atomicAdd(acc + rank, __dmul_rn(a, __ldg(Bval + q)));
    });
    if (!RMCL) continue;
    __syncthreads();
    // ---- fused rMCL epilogue over the scratch row (ascending columns)
    double psum = 0.0, pmax = 0.0;
    for (int k = threadIdx.x; k < cnt; k += BT) {
      const double v = __ldcg(acc + k);  // written by RED at L2: bypass L1
      const double v2 = __dmul_rn(v, v);
      acc[k] = v2;
      psum = __dadd_rn(psum, v2);
      pmax = fmax(pmax, v2);
    }
    const double rsum = block_sum_d<BT>(psum, s_redd);
    const double rmax = block_max_d<BT>(pmax, s_redd);
    const double thresh = compute_threshold(__ddiv_rn(rsum, (double)cnt), rmax);
    double ksum_p = 0.0;
    int kept_p = 0;
    for (int k = threadIdx.x; k < cnt; k += BT) {
      const double v2 = acc[k];
      if (v2 >= thresh) { ksum_p = __dadd_rn(ksum_p, v2); ++kept_p; }
    }
    const double ksum = block_sum_d<BT>(ksum_p, s_redd);
    const int kept = block_sum_int<BT>(kept_p, s_red);
    if (threadIdx.x == 0) s_off = atomicAdd(ro.cursor, (unsigned long long)kept);
    __syncthreads();
    const long long off = (long long)s_off;
    double sq_p = 0.0;
    int written = 0;
    for (int k0 = 0; k0 < cnt; k0 += BT) {
      const int k = k0 + threadIdx.x;
      const bool keep = (k < cnt) && (acc[k] >= thresh);
      int tot;
      const int ex = block_excl_scan<BT>(keep ? 1 : 0, s_red, &tot);
      if (keep) {
        const double w = __ddiv_rn(acc[k], ksum);
        ro.arena_col[off + written + ex] = ocol[k];
        ro.arena_val[off + written + ex] = w;
        sq_p = __dadd_rn(sq_p, __dmul_rn(w, w));
      }
      written += tot;
    }
    const double sq = block_sum_d<BT>(sq_p, s_redd);
    if (threadIdx.x == 0) {
      ro.row_off[i] = off;
      ro.row_kept[i] = kept;
      double ch = (kept > 0) ? __dsub_rn(__ddiv_rn(rmax, ksum), sq) : 0.0;
      if (ch < 0.0) ch = 0.0;
      atomicMax(ro.chaos_bits, (unsigned long long)__double_as_longlong(ch));
    }
  }
}

// rMCL rows whose product is empty keep nothing
__global__ void __launch_bounds__(256)
k_rmcl_empty_rows(const int* __restrict__ list, int count, RmclOut ro) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < count) { ro.row_off[list[t]] = 0; ro.row_kept[list[t]] = 0; }
}

// gather pruned rows from the arena into the final CSR (omp_matrix_relocation,
// nlibs/omp_csr_kernel.cc:201-236); one warp per row
__global__ void __launch_bounds__(256)
k_gather_rows(int m, const long long* __restrict__ row_off, const int64_t* __restrict__ Crp,
              const int* __restrict__ arena_col, const double* __restrict__ arena_val,
              int* __restrict__ Ccol, double* __restrict__ Cval) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= m) return;
  const int64_t o = Crp[warp];
  const int k = (int)(Crp[warp + 1] - o);
  const long long s = row_off[warp];
  for (int t = lane; t < k; t += 32) {
    Ccol[o + t] = arena_col[s + t];
    Cval[o + t] = arena_val[s + t];
  }
}

struct IntToI64 {
  __host__ __device__ __forceinline__ long long operator()(const int& x) const { return (long long)x; }
};

// ---- host helpers -------------------------------------------------------------------------

struct Bins {
  int* d_list = nullptr;  // row ids grouped by bin
  int off[17] = {0};      // bin offsets into d_list
  int cnt[16] = {0};
};

// histogram + scatter; one blocking read of the 16 counts
int make_bins(const unsigned char* d_bin, int m, Bins* out, int* launches) {
  Ctx& c = ctx();
  int* d_hist = nullptr;
  B200_CUDA(dalloc(&d_hist, 32));
  B200_CUDA(cudaMemsetAsync(d_hist, 0, 32 * sizeof(int), c.stream));
  B200_CUDA(dalloc(&out->d_list, (size_t)m));
  if (m > 0) {
    int grid = std::min((m + 255) / 256, c.sm_count * 8);
    k_bin_hist<<<grid, 256, 0, c.stream>>>(d_bin, m, d_hist);
    ++*launches;
  }
  B200_CUDA(cudaMemcpyAsync(out->cnt, d_hist, 16 * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  B200_CUDA(cudaStreamSynchronize(c.stream));
  out->off[0] = 0;
  for (int b = 0; b < 16; ++b) out->off[b + 1] = out->off[b] + out->cnt[b];
  B200_CUDA(cudaMemcpyAsync(d_hist + 16, out->off, 16 * sizeof(int), cudaMemcpyHostToDevice, c.stream));
  if (m > 0) {
    k_bin_scatter<<<(m + 255) / 256, 256, 0, c.stream>>>(d_bin, m, d_hist + 16, out->d_list);
    ++*launches;
  }
  dfree(d_hist);
  return B200_OK;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return B200_OK;
}

constexpr int BT_BIG = 1024;

}  // namespace

// ------------------------------------------------------------------------------------------
// CSR::makeOrdered on the device (nlibs/CSR.cc:73-86): sort every row by column.  Runs once
// at the end of an rMCL loop — not on the per-iteration path — so it uses the CUB segmented
// sort as plain library plumbing.
int sort_rows_device(DevCSR* d) {
  Ctx& c = ctx();
  if (d->nnz == 0 || d->rows == 0) return B200_OK;
  int* col2 = nullptr;
  double* val2 = nullptr;
  B200_CUDA(dalloc(&col2, (size_t)d->nnz));
  B200_CUDA(dalloc(&val2, (size_t)d->nnz));
  void* tmp = nullptr;
  size_t tb = 0;
  cub::DeviceSegmentedSort::SortPairs(nullptr, tb, d->col, col2, d->val, val2, d->nnz, d->rows,
                                      d->rowptr, d->rowptr + 1, c.stream);
  B200_CUDA(cudaMallocAsync(&tmp, tb ? tb : 1, c.stream));
  cub::DeviceSegmentedSort::SortPairs(tmp, tb, d->col, col2, d->val, val2, d->nnz, d->rows,
                                      d->rowptr, d->rowptr + 1, c.stream);
  B200_CUDA(cudaGetLastError());
  cudaFreeAsync(tmp, c.stream);
  dfree(d->col);
  dfree(d->val);
  d->col = col2;
  d->val = val2;
  B200_CUDA(cudaStreamSynchronize(c.stream));
  return B200_OK;
}

// ------------------------------------------------------------------------------------------
int flops_prefix_device(const DevCSR& A, const DevCSR& B, int row_lo, int row_hi,
                        int64_t* d_prefix) {
  Ctx& c = ctx();
  const int m = row_hi - row_lo;
  long long* d_flops = nullptr;
  unsigned char* d_bin = nullptr;
  int* d_cnt = nullptr;
  B200_CUDA(dalloc(&d_flops, (size_t)m + 1));
  B200_CUDA(dalloc(&d_bin, (size_t)m));
  B200_CUDA(dalloc(&d_cnt, (size_t)m));
  B200_CUDA(cudaMemsetAsync(d_flops + m, 0, sizeof(long long), c.stream));
  if (m > 0)
    k_row_flops<<<(m + 255) / 256, 256, 0, c.stream>>>(A.rowptr, A.col, B.rowptr, row_lo, m,
                                                       d_flops, d_bin, d_cnt);
  void* tmp = nullptr;
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, d_flops, (long long*)d_prefix, m + 1, c.stream);
  B200_CUDA(cudaMallocAsync(&tmp, tb ? tb : 1, c.stream));
  cub::DeviceScan::ExclusiveSum(tmp, tb, d_flops, (long long*)d_prefix, m + 1, c.stream);
  B200_CUDA(cudaGetLastError());
  cudaFreeAsync(tmp, c.stream);
  dfree(d_flops); dfree(d_bin); dfree(d_cnt);
  return B200_OK;
}

// ------------------------------------------------------------------------------------------
int run_pipeline(const DevCSR& A, const DevCSR& B, int row_lo, int row_hi, Mode mode, DevCSR* C,
                 double* chaos, b200_stats* stats) {
  Ctx& c = ctx();
  cudaStream_t st = c.stream;
  const int m = row_hi - row_lo;
  const int n = B.cols;
  int launches = 0;
  if (stats) memset(stats, 0, sizeof(*stats));
  *C = DevCSR();
  C->rows = m;
  C->cols = n;
  B200_CUDA(cudaEventRecord(c.ev[0], st));
  // per-kernel brackets (only when the caller asked for stats)
  bool sym_timed[16] = {false}, num_timed[16] = {false};
  auto tick = [&](int slot) { if (stats) cudaEventRecord(c.kev[slot], st); };

  // ---- 1. flops analysis + symbolic binning
  long long* d_flops = nullptr;
  unsigned char* d_bin = nullptr;
  int* d_cnt = nullptr;
  long long* d_P = nullptr;
  B200_CUDA(dalloc(&d_flops, (size_t)m));
  B200_CUDA(dalloc(&d_bin, (size_t)m));
  B200_CUDA(dalloc(&d_cnt, (size_t)m + 1));
  B200_CUDA(dalloc(&d_P, 2));
  if (m > 0) {
    k_row_flops<<<(m + 255) / 256, 256, 0, st>>>(A.rowptr, A.col, B.rowptr, row_lo, m, d_flops,
                                                 d_bin, d_cnt);
    ++launches;
  }
  {
    void* tmp = nullptr;
    size_t tb = 0;
    cub::DeviceReduce::Sum(nullptr, tb, d_flops, d_P, m, st);
    B200_CUDA(cudaMallocAsync(&tmp, tb ? tb : 1, st));
    cub::DeviceReduce::Sum(tmp, tb, d_flops, d_P, m, st);
    cudaFreeAsync(tmp, st);
    ++launches;
  }
  Bins sb;
  int rc = make_bins(d_bin, m, &sb, &launches);
  if (rc) return rc;
  B200_CUDA(cudaEventRecord(c.ev[1], st));

  // ---- 2. symbolic per bin
  auto launch_sym_warp = [&](int bin, auto kernel, int H, int WPB) -> int {
    const int cntb = sb.cnt[bin];
    if (!cntb) return B200_OK;
    const size_t smem = (size_t)WPB * H * sizeof(int);
    int r = set_smem(kernel, smem);
    if (r) return r;
    tick(2 * bin);
    kernel<<<(cntb + WPB - 1) / WPB, WPB * 32, smem, st>>>(sb.d_list + sb.off[bin], cntb, row_lo,
                                                           A.rowptr, A.col, B.rowptr, B.col, d_cnt);
    tick(2 * bin + 1);
    sym_timed[bin] = true;
    ++launches;
    return B200_OK;
  };
  if ((rc = launch_sym_warp(SB_W256, k_sym_warp<256>, 256, 8))) return rc;
  if ((rc = launch_sym_warp(SB_W1K, k_sym_warp<1024>, 1024, 8))) return rc;
  if ((rc = launch_sym_warp(SB_W4K, k_sym_warp<4096>, 4096, 8))) return rc;
  if ((rc = launch_sym_warp(SB_W16K, k_sym_warp<16384>, 16384, 3))) return rc;

  // large rows: bitmap.  nw64 words of 64 columns.
  const int nw64 = (n + 63) / 64;
  const size_t bm_bytes = (size_t)nw64 * 8;
  const size_t bm_pref_bytes = bm_bytes + (((size_t)nw64 + 1) / 2) * 8;
  const size_t smem_cap = c.smem_optin - 1024;  // static shared of the kernels
  const bool sym_smem = bm_bytes <= smem_cap;
  const bool num_smem = bm_pref_bytes <= smem_cap;
  unsigned long long* d_bmstore = nullptr;  // stored bitmaps of the symbolic bitmap bin
  unsigned long long* d_gscr = nullptr;     // per-CTA bitmap(+prefix) scratch when not in smem
  int* d_work = nullptr;
  B200_CUDA(dalloc(&d_work, 4));
  B200_CUDA(cudaMemsetAsync(d_work, 0, 4 * sizeof(int), st));
  const int nbig = sb.cnt[SB_BITMAP];
  int big_grid = std::min(nbig, c.sm_count);
  bool store_bitmaps = false;
  if (nbig) {
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    // keep the bitmaps only if they use a modest share of what is free (C itself comes later)
    store_bitmaps = (double)nbig * (double)bm_bytes < 0.20 * (double)free_b;
    if (store_bitmaps) B200_CUDA(dalloc(&d_bmstore, (size_t)nbig * nw64));
    if (!sym_smem || !num_smem)
      B200_CUDA(dalloc(&d_gscr, (size_t)c.sm_count * (bm_pref_bytes / 8)));
    tick(2 * SB_BITMAP);
    sym_timed[SB_BITMAP] = true;
    if (sym_smem) {
      if ((rc = set_smem(k_sym_bitmap<BT_BIG, true>, bm_bytes))) return rc;
      k_sym_bitmap<BT_BIG, true><<<big_grid, BT_BIG, bm_bytes, st>>>(
          sb.d_list + sb.off[SB_BITMAP], nbig, row_lo, A.rowptr, A.col, B.rowptr, B.col, nw64,
          nullptr, d_bmstore, d_cnt, d_work + 0);
    } else {
      k_sym_bitmap<BT_BIG, false><<<big_grid, BT_BIG, 0, st>>>(
          sb.d_list + sb.off[SB_BITMAP], nbig, row_lo, A.rowptr, A.col, B.rowptr, B.col, nw64,
          d_gscr, d_bmstore, d_cnt, d_work + 0);
    }
    tick(2 * SB_BITMAP + 1);
    ++launches;
  }
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaEventRecord(c.ev[2], st));

  // ---- 3. numeric binning, row offsets
  unsigned char* d_nbin = nullptr;
  B200_CUDA(dalloc(&d_nbin, (size_t)m));
  if (m > 0) {
    k_num_bins<<<(m + 255) / 256, 256, 0, st>>>(d_cnt, m, d_nbin);
    ++launches;
  }
  B200_CUDA(cudaMemsetAsync(d_cnt + m, 0, sizeof(int), st));
  int64_t* d_urp = nullptr;  // offsets of the unpruned product (== C.rowptr for SpGEMM)
  B200_CUDA(dalloc(&d_urp, (size_t)m + 1));
  {
    cub::TransformInputIterator<long long, IntToI64, const int*> it(d_cnt, IntToI64());
    void* tmp = nullptr;
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, it, (long long*)d_urp, m + 1, st);
    B200_CUDA(cudaMallocAsync(&tmp, tb ? tb : 1, st));
    cub::DeviceScan::ExclusiveSum(tmp, tb, it, (long long*)d_urp, m + 1, st);
    cudaFreeAsync(tmp, st);
    ++launches;
  }
  long long h_tot[2] = {0, 0};  // unpruned nnz, products
  B200_CUDA(cudaMemcpyAsync(&h_tot[0], d_urp + m, sizeof(long long), cudaMemcpyDeviceToHost, st));
  B200_CUDA(cudaMemcpyAsync(&h_tot[1], d_P, sizeof(long long), cudaMemcpyDeviceToHost, st));
  Bins nb;
  rc = make_bins(d_nbin, m, &nb, &launches);  // synchronises the stream
  if (rc) return rc;
  const long long unpruned = h_tot[0];

  // map rows of the numeric bitmap bin to their stored bitmap (or -1)
  int* d_bmindex = nullptr;
  const int nbig_num = nb.cnt[NB_BITMAP];
  std::vector<int> h_bmindex;
  if (nbig_num) {
    // stored slot = position in the symbolic bitmap list; build inverse on the host (rare,
    // small lists) — rows not in the symbolic bitmap bin rebuild their bitmap in the kernel.
    std::vector<int> h_symlist(nbig), h_numlist(nbig_num);
    if (nbig)
      B200_CUDA(cudaMemcpyAsync(h_symlist.data(), sb.d_list + sb.off[SB_BITMAP], nbig * sizeof(int),
                                cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaMemcpyAsync(h_numlist.data(), nb.d_list + nb.off[NB_BITMAP],
                              nbig_num * sizeof(int), cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    h_bmindex.assign(nbig_num, -1);
    if (store_bitmaps && nbig) {
      std::vector<std::pair<int, int>> inv(nbig);
      for (int t = 0; t < nbig; ++t) inv[t] = {h_symlist[t], t};
      std::sort(inv.begin(), inv.end());
      for (int t = 0; t < nbig_num; ++t) {
        auto it = std::lower_bound(inv.begin(), inv.end(), std::make_pair(h_numlist[t], -1));
        if (it != inv.end() && it->first == h_numlist[t]) h_bmindex[t] = it->second;
      }
    }
    B200_CUDA(dalloc(&d_bmindex, (size_t)nbig_num));
    B200_CUDA(cudaMemcpyAsync(d_bmindex, h_bmindex.data(), nbig_num * sizeof(int),
                              cudaMemcpyHostToDevice, st));
  }
  B200_CUDA(cudaEventRecord(c.ev[3], st));

  // ---- 4. outputs
  RmclOut ro = {};
  int* d_arena_col = nullptr;
  double* d_arena_val = nullptr;
  unsigned long long* d_cursor = nullptr;  // [0] bump cursor, [1] chaos bits
  long long* d_rowoff = nullptr;
  int* d_kept = nullptr;
  int* d_scr_col = nullptr;
  double* d_scr_val = nullptr;
  long long scr_stride = 0;
  if (mode == MODE_SPGEMM) {
    C->rowptr = d_urp;
    C->nnz = unpruned;
    B200_CUDA(dalloc(&C->col, (size_t)unpruned));
    B200_CUDA(dalloc(&C->val, (size_t)unpruned));
  } else {
    B200_CUDA(dalloc(&d_arena_col, (size_t)unpruned));
    B200_CUDA(dalloc(&d_arena_val, (size_t)unpruned));
    B200_CUDA(dalloc(&d_cursor, 2));
    B200_CUDA(cudaMemsetAsync(d_cursor, 0, 2 * sizeof(unsigned long long), st));
    B200_CUDA(dalloc(&d_rowoff, (size_t)m));
    B200_CUDA(dalloc(&d_kept, (size_t)m + 1));
    ro.arena_col = d_arena_col;
    ro.arena_val = d_arena_val;
    ro.cursor = d_cursor;
    ro.chaos_bits = d_cursor + 1;
    ro.row_off = d_rowoff;
    ro.row_kept = d_kept;
    if (nbig_num) {
      scr_stride = n;  // a row has at most n distinct columns
      B200_CUDA(dalloc(&d_scr_col, (size_t)std::min(nbig_num, c.sm_count) * scr_stride));
      B200_CUDA(dalloc(&d_scr_val, (size_t)std::min(nbig_num, c.sm_count) * scr_stride));
    }
    if (nb.cnt[NB_NONE]) {
      k_rmcl_empty_rows<<<(nb.cnt[NB_NONE] + 255) / 256, 256, 0, st>>>(nb.d_list + nb.off[NB_NONE],
                                                                      nb.cnt[NB_NONE], ro);
      ++launches;
    }
  }

  // ---- 5. numeric per bin
  auto launch_num_warp = [&](int bin, auto kernel, int CAP, int WPB) -> int {
    const int cntb = nb.cnt[bin];
    if (!cntb) return B200_OK;
    const size_t smem = (size_t)WPB * 24 * CAP;
    int r = set_smem(kernel, smem);
    if (r) return r;
    tick(32 + 2 * bin);
    kernel<<<(cntb + WPB - 1) / WPB, WPB * 32, smem, st>>>(
        nb.d_list + nb.off[bin], cntb, row_lo, A.rowptr, A.col, A.val, B.rowptr, B.col, B.val,
        d_urp, C->col, C->val, ro);
    tick(32 + 2 * bin + 1);
    num_timed[bin] = true;
    ++launches;
    return B200_OK;
  };
  if (mode == MODE_SPGEMM) {
    if ((rc = launch_num_warp(NB_W64, k_num_warp<64, false>, 64, 8))) return rc;
    if ((rc = launch_num_warp(NB_W256, k_num_warp<256, false>, 256, 8))) return rc;
    if ((rc = launch_num_warp(NB_W1K, k_num_warp<1024, false>, 1024, 8))) return rc;
    if ((rc = launch_num_warp(NB_W2K, k_num_warp<2048, false>, 2048, 4))) return rc;
  } else {
    if ((rc = launch_num_warp(NB_W64, k_num_warp<64, true>, 64, 8))) return rc;
    if ((rc = launch_num_warp(NB_W256, k_num_warp<256, true>, 256, 8))) return rc;
    if ((rc = launch_num_warp(NB_W1K, k_num_warp<1024, true>, 1024, 8))) return rc;
    if ((rc = launch_num_warp(NB_W2K, k_num_warp<2048, true>, 2048, 4))) return rc;
  }
  if (nbig_num) {
    const int grid = std::min(nbig_num, c.sm_count);
    if (!num_smem && !d_gscr)
      B200_CUDA(dalloc(&d_gscr, (size_t)c.sm_count * (bm_pref_bytes / 8)));
    const int* lst = nb.d_list + nb.off[NB_BITMAP];
    tick(32 + 2 * NB_BITMAP);
    num_timed[NB_BITMAP] = true;
#define LAUNCH_NUM_BM(SM, RM)                                                                   \
  do {                                                                                          \
    if (SM && (rc = set_smem(k_num_bitmap<BT_BIG, SM, RM>, bm_pref_bytes))) return rc;          \
    k_num_bitmap<BT_BIG, SM, RM><<<grid, BT_BIG, SM ? bm_pref_bytes : 0, st>>>(                 \
        lst, nbig_num, row_lo, A.rowptr, A.col, A.val, B.rowptr, B.col, B.val, nw64, d_gscr,    \
        d_bmstore, d_bmindex, d_urp, C->col, C->val, d_scr_col, d_scr_val, scr_stride, ro,      \
        d_work + 1);                                                                            \
  } while (0)
    if (num_smem) {
      if (mode == MODE_SPGEMM) LAUNCH_NUM_BM(true, false); else LAUNCH_NUM_BM(true, true);
    } else {
      if (mode == MODE_SPGEMM) LAUNCH_NUM_BM(false, false); else LAUNCH_NUM_BM(false, true);
    }
#undef LAUNCH_NUM_BM
    tick(32 + 2 * NB_BITMAP + 1);
    ++launches;
  }
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaEventRecord(c.ev[4], st));

  // ---- 6. rMCL: scan kept counts, gather arena -> final CSR
  long long nnz_out = unpruned;
  if (mode == MODE_RMCL) {
    B200_CUDA(cudaMemsetAsync(d_kept + m, 0, sizeof(int), st));
    B200_CUDA(dalloc(&C->rowptr, (size_t)m + 1));
    cub::TransformInputIterator<long long, IntToI64, const int*> it(d_kept, IntToI64());
    void* tmp = nullptr;
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, it, (long long*)C->rowptr, m + 1, st);
    B200_CUDA(cudaMallocAsync(&tmp, tb ? tb : 1, st));
    cub::DeviceScan::ExclusiveSum(tmp, tb, it, (long long*)C->rowptr, m + 1, st);
    cudaFreeAsync(tmp, st);
    ++launches;
    unsigned long long h_cur[2] = {0, 0};
    B200_CUDA(cudaMemcpyAsync(h_cur, d_cursor, 2 * sizeof(unsigned long long),
                              cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    nnz_out = (long long)h_cur[0];
    if (chaos) {
      long long bits = (long long)h_cur[1];
      memcpy(chaos, &bits, sizeof(double));
    }
    C->nnz = nnz_out;
    B200_CUDA(dalloc(&C->col, (size_t)nnz_out));
    B200_CUDA(dalloc(&C->val, (size_t)nnz_out));
    if (m > 0) {
      const long long threads = (long long)m * 32;
      k_gather_rows<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(
          m, d_rowoff, C->rowptr, d_arena_col, d_arena_val, C->col, C->val);
      ++launches;
    }
    dfree(d_urp);
    dfree(d_arena_col); dfree(d_arena_val); dfree(d_cursor); dfree(d_rowoff); dfree(d_kept);
    dfree(d_scr_col); dfree(d_scr_val);
  }
  B200_CUDA(cudaEventRecord(c.ev[5], st));
  unsigned long long h_agg[96] = {0};
  if (stats && m > 0) {
    unsigned long long* d_agg = nullptr;
    B200_CUDA(dalloc(&d_agg, 96));
    B200_CUDA(cudaMemsetAsync(d_agg, 0, 96 * sizeof(unsigned long long), st));
    const int grid = std::min((m + 255) / 256, c.sm_count * 8);
    k_bin_aggregate<<<grid, 256, 0, st>>>(d_bin, m, A.rowptr, row_lo, d_flops, nullptr, d_agg);
    k_bin_aggregate<<<grid, 256, 0, st>>>(d_nbin, m, A.rowptr, row_lo, d_flops, d_cnt, d_agg + 48);
    B200_CUDA(cudaMemcpyAsync(h_agg, d_agg, sizeof h_agg, cudaMemcpyDeviceToHost, st));
    dfree(d_agg);
  }
  dfree(d_flops); dfree(d_bin); dfree(d_cnt); dfree(d_P); dfree(d_nbin);
  dfree(sb.d_list); dfree(nb.d_list); dfree(d_bmstore); dfree(d_gscr); dfree(d_work);
  dfree(d_bmindex);
  B200_CUDA(cudaStreamSynchronize(st));
  B200_CUDA(cudaGetLastError());
  if (stats) {
    float t01, t12, t23, t34, t45, t05;
    cudaEventElapsedTime(&t01, c.ev[0], c.ev[1]);
    cudaEventElapsedTime(&t12, c.ev[1], c.ev[2]);
    cudaEventElapsedTime(&t23, c.ev[2], c.ev[3]);
    cudaEventElapsedTime(&t34, c.ev[3], c.ev[4]);
    cudaEventElapsedTime(&t45, c.ev[4], c.ev[5]);
    cudaEventElapsedTime(&t05, c.ev[0], c.ev[5]);
    stats->ms_total = t05;
    stats->ms_flops = t01;
    stats->ms_symbolic = t12;
    stats->ms_numeric = t34;
    stats->ms_other = t23 + t45;
    stats->products = h_tot[1];
    stats->nnz_out = nnz_out;
    stats->nnz_unpruned = unpruned;
    stats->launches = launches;
    for (int b = 0; b < 16; ++b) {
      stats->bins_rows[b] = nb.cnt[b];
      stats->sym_bin_rows[b] = sb.cnt[b];
      stats->sym_bin_products[b] = (long long)h_agg[b * 3 + 0];
      stats->sym_bin_nnzA[b] = (long long)h_agg[b * 3 + 1];
      stats->num_bin_products[b] = (long long)h_agg[48 + b * 3 + 0];
      stats->num_bin_nnzA[b] = (long long)h_agg[48 + b * 3 + 1];
      stats->num_bin_nnzC[b] = (long long)h_agg[48 + b * 3 + 2];
      float t = 0;
      if (sym_timed[b]) { cudaEventElapsedTime(&t, c.kev[2 * b], c.kev[2 * b + 1]); stats->ms_sym_bin[b] = t; }
      if (num_timed[b]) { cudaEventElapsedTime(&t, c.kev[32 + 2 * b], c.kev[32 + 2 * b + 1]); stats->ms_num_bin[b] = t; }
    }
  }
  return B200_OK;
}

}  // namespace b200
