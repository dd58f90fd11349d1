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
//  * large rows: a column BITMAP in shared memory indexes the accumulator,
//    rank(c) = popcount prefix of the word + popcount of the bits below c, so the output row is
//    produced directly in ascending column order.  Symbolic: CTA per row, test-then-atomicOr.
//    Numeric (SpGEMM, sorted B): the columns of B are cut into parts of <= 8192 bitmap words,
//    the work unit is a (row, part) slot handled by 512-thread CTAs (two per SM), heavy slots
//    by a team of CTAs; the products are found with a flat, coalesced walk over the B-row
//    segments of the part (k_bsplit table) and reduced with fp64 RED into the row's final slice
//    of C.val (order of additions not fixed: values agree to rounding, not bitwise);
//  * rMCL: inflation, row max/sum, threshold, prune, normalise and the chaos term are fused
//    behind the numeric row while it is still on chip (hash bins, on-chip sort bin: only the kept
//    entries leave the chip) or run in place over the row's arena slice (bitmap bin: the part
//    kernel writes the unpruned row there first).  The arena is sized by the unpruned entries of
//    the rows of ONE call; tiles.cu bounds it by running a step as consecutive row tiles.  Kept
//    entries are gathered into the final CSR after a scan of the kept counts;
//  * mid-size rows of matrices wider than two column parts: expanded, sorted by column and
//    summed on chip (esc.cuh), no symbolic pass;
//  * opt-in (B200_ON_CHIP / B200_DETERMINISTIC): heavy rows accumulated on chip in A-entry order
//    (ranges.cuh) instead of with RED.
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
constexpr int SB_ESC2K = 6;   // sorted on chip (esc.cuh): <= 2048 products from <= 256 A entries
constexpr int SB_ESC8K = 7;   // ... <= 8192 products from <= 512 A entries.  No symbolic hashing: the
                              // row's products bound its entries
// numeric bins (by nnz(C_i))
constexpr int NB_NONE = 0;    // empty row
constexpr int NB_W64 = 1;
constexpr int NB_W256 = 2;
constexpr int NB_W1K = 3;
constexpr int NB_W2K = 4;
constexpr int NB_BITMAP = 5;
constexpr int NB_W128 = 6;   // nnz(C_i) in (64, 128]: e.g. every interior row of a 27-point stencil
constexpr int NB_ESC2K = 7;  // rMCL rows sorted on chip: <= 2048 products from <= 256 A entries
constexpr int NB_ESC8K = 8;  // ... <= 8192 products from <= 512 A entries
constexpr int NB_FUSED = 9;  // SpGEMM rows finished in the symbolic phase by k_num_warp_fused (<= 128 columns)
constexpr int FUSED_CAP = 128;
constexpr int ESC_MIN_P = 1024;  // below: the small warp tables (first-touch order, bit-identical row sums) serve well

// `big_from`: rows above this size go to the bitmap bin.  When the column bitmap of B fits in
// shared memory it is the cheapest accumulator index for every row past the small warp tables
// (its per-row cost is O(n/64) words to clear), so the cut is low (P > 512, nnz(C_i) > 256);
// otherwise the larger warp tables take the rows up to 8192 products / 2048 entries.
__host__ __device__ inline int sym_bin_of(long long P, int annz, long long big_from, long long esc_max = 0) {
  if (P == 0 || annz <= 1) return SB_NONE;
  if (P > ESC_MIN_P && P <= esc_max) {
    if (P <= 2048 && annz <= 256) return SB_ESC2K;
    if (annz <= 512) return SB_ESC8K;
  }
  if (P <= 128) return SB_W256;
  if (P <= 512) return SB_W1K;
  if (P > big_from) return SB_BITMAP;
  if (P <= 2048) return SB_W4K;
  return SB_W16K;
}
// `light_p`: with the bitmap available, a row of up to 2048 columns still goes to the 1024- or
// 2048-entry warp table when it has few products — the bitmap item costs ~20 us of fixed work per (row,
// part), the big warp table ~0.1 us per product at its 8 warps / SM: measured crossover on
// planted-partition rows ~2.5 K products.
__host__ __device__ inline int num_bin_of(int cnt, int big_from, long long P, long long light_p,
                                          long long light2k_p) {
  if (cnt == 0) return NB_NONE;
  if (cnt <= 64) return NB_W64;
  if (cnt <= 128) return NB_W128;
  if (cnt <= 256) return NB_W256;
  if (cnt <= 1024 && P <= light_p) return NB_W1K;
  if (cnt <= 2048 && P <= light2k_p) return NB_W2K;  // only with > 2 column parts (measured)
  if (cnt > big_from) return NB_BITMAP;
  if (cnt <= 1024) return NB_W1K;
  return NB_W2K;
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
__device__ __forceinline__ long long shfl64_xor(long long v, int m) {
  int lo = __shfl_xor_sync(FULL, (int)(v & 0xffffffffLL), m);
  int hi = __shfl_xor_sync(FULL, (int)(v >> 32), m);
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

// ---- L2 eviction-priority hints (createpolicy + .L2::cache_hint) -----------------------------
// The numeric pass of a heavy row keeps three kinds of lines in flight: the row's accumulators
// (zeroed, then hit by one fp64 RED per product: best kept in L2 until the row is done), the
// gathered B rows (re-used across rows only when they are hub rows) and pure streams (the
// column list being written, the stored bitmaps being read).  Measured on R-MAT scale 20:
// accumulators evict_last + streams evict_first is worth 6 % of the kernel; the kernel is bound
// by the SM-side fp64 RED issue rate, not by where the line lives (profiles/README.md).
enum L2Prio { L2_NORMAL = 0, L2_FIRST = 1, L2_LAST = 2 };
__device__ __forceinline__ unsigned long long l2_policy(int prio) {
  unsigned long long p;
  if (prio == L2_LAST) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  else if (prio == L2_FIRST) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ int ldg_hint(const int* a, unsigned long long pol) {
  int v;
  asm volatile("ld.global.nc.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol));
  return v;
}
__device__ __forceinline__ double ldg_hint(const double* a, unsigned long long pol) {
  double v;
  asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(a), "l"(pol));
  return v;
}
__device__ __forceinline__ unsigned long long ldg_hint(const unsigned long long* a, unsigned long long pol) {
  unsigned long long v;
  asm volatile("ld.global.nc.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(v) : "l"(a), "l"(pol));
  return v;
}
__device__ __forceinline__ ulonglong2 ldg_hint(const ulonglong2* a, unsigned long long pol) {
  ulonglong2 v;
  asm volatile("ld.global.nc.L2::cache_hint.v2.b64 {%0, %1}, [%2], %3;" : "=l"(v.x), "=l"(v.y) : "l"(a), "l"(pol));
  return v;
}
__device__ __forceinline__ void stg_hint(double* a, double v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(a), "d"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void stg_hint(int* a, int v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;" ::"l"(a), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void stg_hint(unsigned long long* a, unsigned long long v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.b64 [%0], %1, %2;" ::"l"(a), "l"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void red_add_hint(double* a, double v, unsigned long long pol) {
  asm volatile("red.global.add.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(a), "d"(v), "l"(pol) : "memory");
}
// drop a 128-byte line back to normal priority (after the row that owned it is complete)
__device__ __forceinline__ void l2_demote(const void* a) {
  asm volatile("applypriority.global.L2::evict_normal [%0], 128;" ::"l"(a) : "memory");
}
// packed per-kernel policy choices (host side: see l2_modes() in run_pipeline)
struct L2Modes { int acc, ocol, bgather, bmstore, demote; };

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
// Eight lanes per row (a hub row of 30 000 entries walked by one thread would be the whole
// kernel's critical path); also emits the symbolic bin and, for rows that need no hashing,
// nnz(C_i).
__global__ void __launch_bounds__(256)
k_row_flops(const int64_t* __restrict__ Arp, const int* __restrict__ Acol,
            const int64_t* __restrict__ Brp, int row_lo, int m, long long big_from,
            long long* __restrict__ flops, unsigned char* __restrict__ sbin,
            int* __restrict__ rownnz, long long esc_max = 0) {
  const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int i = (int)(gt >> 3), sub = (int)(gt & 7);
  const bool live = i < m;
  long long f = 0;
  int64_t a0 = 0, a1 = 0;
  if (live) {
    a0 = Arp[row_lo + i];
    a1 = Arp[row_lo + i + 1];
    for (int64_t p = a0 + sub; p < a1; p += 8) {
      const int j = __ldg(Acol + p);
      f += __ldg(Brp + j + 1) - __ldg(Brp + j);
    }
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) f += shfl64_xor(f, o);
  if (!live || sub != 0) return;
  flops[i] = f;
  const int b = sym_bin_of(f, (int)(a1 - a0), big_from, esc_max);
  sbin[i] = (unsigned char)b;
  // 0, or the length of the single B row; for a row sorted on chip its products BOUND its entries
  if (b == SB_NONE || b == SB_ESC2K || b == SB_ESC8K) rownnz[i] = (int)f;
}

__global__ void __launch_bounds__(256)
k_num_bins(const int* __restrict__ rownnz, const long long* __restrict__ flops, int m, int big_from,
           long long light_p, long long light2k_p, const unsigned char* __restrict__ sbin,
           const int64_t* __restrict__ Arp, int row_lo, const unsigned char* __restrict__ fused,
           unsigned char* __restrict__ nbin) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  if (fused && fused[i]) nbin[i] = (unsigned char)NB_FUSED;
  else if (sbin[i] == SB_ESC2K) nbin[i] = (unsigned char)NB_ESC2K;
  else if (sbin[i] == SB_ESC8K) nbin[i] = (unsigned char)NB_ESC8K;
  else nbin[i] = (unsigned char)num_bin_of(rownnz[i], big_from, flops[i], light_p, light2k_p);
}

// arena length of a row finished on chip in the symbolic phase (SpGEMM, esc.cuh): its products
__global__ void __launch_bounds__(256)
k_esc_len(const unsigned char* __restrict__ sbin, const long long* __restrict__ flops, int m,
          long long* __restrict__ len) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= m) len[i] = (i < m && (sbin[i] == 6 || sbin[i] == 7)) ? flops[i] : 0;
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
  const int lane = threadIdx.x & 31;
  // (grid-stride by whole warps; 32 consecutive rows usually share a bin: one set of atomics per
  // warp then — 16.7 M rows of one bin made this kernel 4 ms of same-address shared atomics)
  for (long long base = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < m;
       base += (long long)gridDim.x * blockDim.x) {
    const long long i = base + lane;
    const bool live = i < m;
    const int b = live ? (int)bin[i] : -1;
    unsigned long long v0 = 0, v1 = 0, v2 = 0;
    if (live) {
      v0 = (unsigned long long)flops[i];
      v1 = (unsigned long long)(Arp[row_lo + i + 1] - Arp[row_lo + i]);
      if (rownnz) v2 = (unsigned long long)rownnz[i];
    }
    const int b0 = __shfl_sync(FULL, b, 0);   // (lane 0 is live: base < m)
    if (__all_sync(FULL, !live || b == b0)) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        v0 += (unsigned long long)shfl64_xor((long long)v0, o);
        v1 += (unsigned long long)shfl64_xor((long long)v1, o);
        v2 += (unsigned long long)shfl64_xor((long long)v2, o);
      }
      if (lane == 0) {
        atomicAdd(&s[b0 * 3 + 0], v0);
        atomicAdd(&s[b0 * 3 + 1], v1);
        if (rownnz) atomicAdd(&s[b0 * 3 + 2], v2);
      }
    } else if (live) {
      atomicAdd(&s[b * 3 + 0], v0);
      atomicAdd(&s[b * 3 + 1], v1);
      if (rownnz) atomicAdd(&s[b * 3 + 2], v2);
    }
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
k_sym_warp(const int* __restrict__ list, int count, const int* __restrict__ count_dev,
           int row_lo, const int64_t* __restrict__ Arp, const int* __restrict__ Acol,
           const int64_t* __restrict__ Brp, const int* __restrict__ Bcol,
           int* __restrict__ rownnz, int limit, int* __restrict__ overflow_list,
           int* __restrict__ overflow_count) {
  // `limit` < H: OPTIMISTIC sizing.  A table sized for the products P_i (an upper bound of
  // nnz(C_i)) costs 16 KB per warp for a 27-point stencil row (P = 729, nnz = 125) and leaves
  // one 8-warp block per SM; instead the row is tried in a table a quarter of that size and
  // handed to the next size up (overflow_list) only if it really has more than `limit` columns.
  extern __shared__ int smem_i[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (count_dev) count = *count_dev;   // retry pass: the list length lives on the device
  int* keys = smem_i + warp * H;
  // (grid-stride: a retry pass is launched with a capped grid, its list is usually short)
  for (int idx = blockIdx.x * (blockDim.x >> 5) + warp; idx < count; idx += gridDim.x * (blockDim.x >> 5)) {
    const int i = list[idx];
    __syncwarp();
    for (int k = lane; k < H; k += 32) keys[k] = EMPTY;
    __syncwarp();
    const int64_t a0 = Arp[row_lo + i], a1 = Arp[row_lo + i + 1];
    int cnt = 0;  // warp-uniform
    bool over = false;
    for (int64_t base = a0; base < a1 && !over; base += 32) {
      const int64_t p = base + lane;
      long long bs = 0, be = 0;
      if (p < a1) { int j = __ldg(Acol + p); bs = __ldg(Brp + j); be = __ldg(Brp + j + 1); }
      const int nn = (int)min((int64_t)32, a1 - base);
      for (int t = 0; t < nn && !over; ++t) {
        const long long s = shfl64(bs, t), e = shfl64(be, t);
        for (long long q0 = s; q0 < e; q0 += 32) {
          const long long q = q0 + lane;
          const bool act = q < e;
          const int c = act ? __ldg(Bcol + q) : 0;
          unsigned h;
          const bool isnew = warp_find_or_insert<H>(keys, c, act, h);
          cnt += __popc(__ballot_sync(FULL, isnew));
          if (cnt > limit) { over = true; break; }   // the next step could fill the table
        }
      }
    }
    if (lane == 0) {
      if (over) overflow_list[atomicAdd(overflow_count, 1)] = i;
      else rownnz[i] = cnt;
    }
  }
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
  long long* row_off;           // [m] arena offset of the row
  int* row_kept;                // [m] kept entries
  unsigned long long* chaos_bits;  // max over rows of (max - sum sq), as ordered bits
  int topk;                        // > 0: keep at most this many entries per row (b200_set_topk)
};

// ---- top-k pruning (opt-in; not in the reference, whose rule is the threshold alone, SURVEY.md
// §8a; oracle/oracle.c states the same rule for the checker).  Of the entries that pass the
// threshold the k largest stay; equal values are ranked by ascending column.  The cut is found
// without sorting: the k-th largest value by bisection on the bit pattern of the (non-negative)
// values, then the last admitted column among its ties by bisection on the column.
// count(pred) sums over the row: warp-wide or block-wide, supplied by the caller.
template <typename Count>
__device__ __forceinline__ void topk_cut(Count count, double thresh, double rmax, int k, int ncols,
                                         double* tk_out, int* cstar_out) {
  unsigned long long lo = (unsigned long long)__double_as_longlong(thresh);
  unsigned long long hi = (unsigned long long)__double_as_longlong(rmax) + 1ull;
  while (hi - lo > 1ull) {   // count(v >= lo) >= k > count(v >= hi)
    const unsigned long long mid = lo + ((hi - lo) >> 1);
    const double m = __longlong_as_double((long long)mid);
    if (count([&](double v, int) { return v >= m; }) >= k) lo = mid; else hi = mid;
  }
  const double tk = __longlong_as_double((long long)lo);
  const int need = k - count([&](double v, int) { return v > tk; });  // ties to admit, >= 1
  int clo = -1, chi = ncols - 1;   // count(v == tk, col <= clo) < need <= count(v == tk, col <= chi)
  while (chi - clo > 1) {
    const int mid = clo + ((chi - clo) >> 1);
    if (count([&](double v, int c) { return v == tk && c <= mid; }) >= need) chi = mid; else clo = mid;
  }
  *tk_out = tk;
  *cstar_out = chi;
}

// numeric, one warp per row, CAP output entries per warp (indexProcessCRowI,
// cpu_csr_kernel.h:134-188) + sort (+ fused rMCL epilogue when RMCL).
// shared memory per warp (24*CAP bytes):
//   vals  fp64[CAP]      slot -> accumulated value (first-touch order)
//   keys  int[2*CAP]     hash keys; later reused as the u64[CAP] sort buffer
//   cols  int[CAP]       slot -> column           \ later reused together as fp64[CAP]:
//   slot  ushort[2*CAP]  hash slot -> output slot / squared values in sorted order
template <int CAP, bool RMCL>
__device__ __forceinline__ void num_warp_row(int i, unsigned char* wbase, int lane, int row_lo,
           const int64_t* __restrict__ Arp, const int* __restrict__ Acol,
           const double* __restrict__ Aval, const int64_t* __restrict__ Brp,
           const int* __restrict__ Bcol, const double* __restrict__ Bval,
           const int64_t* __restrict__ Crp, int* __restrict__ Ccol, double* __restrict__ Cval,
           const RmclOut& ro) {
  constexpr int H = 2 * CAP;
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
  // opt-in top-k: the cut (tk, cstar) narrows the keep rule to the k largest entries
  double tk = 0.0;
  int cstar = 0x7fffffff;
  bool cut = false;
  if (ro.topk > 0) {
    auto wcount = [&](auto pred) {
      int c = 0;
      for (int k = lane; k < cnt; k += 32) c += (vals[k] >= thresh && pred(vals[k], cols[k])) ? 1 : 0;
      return warp_sum_int(c);
    };
    if (wcount([](double, int) { return true; }) > ro.topk) {
      cut = true;
      topk_cut(wcount, thresh, rmax, ro.topk, 0x7fffffff, &tk, &cstar);
    }
  }
  auto keeps = [&](double v, int c) { return v >= thresh && (!cut || v > tk || (v == tk && c <= cstar)); };
  double ksum = 0.0;
  int kept = 0;
#pragma unroll 4
  for (int k = 0; k < cnt; ++k) {
    const double v = vals[k];
    const bool keep = keeps(v, cols[k]);
    ksum = keep ? __dadd_rn(ksum, v) : ksum;
    kept += keep ? 1 : 0;
  }
  // the row's slice of the arena is where its UNPRUNED product would start (Crp = offsets of
  // the unpruned product): kept entries go to the front of it
  const unsigned long long off = (unsigned long long)Crp[i];
  double sq_p = 0.0;
  int written = 0;
  for (int k0 = 0; k0 < cnt; k0 += 32) {
    const int k = k0 + lane;
    const double v = (k < cnt) ? vals[k] : 0.0;
    const bool keep = (k < cnt) && keeps(v, cols[k]);
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


// `work_counter` != nullptr: the warps draw rows from a shared counter instead of owning one
// row each — with the large tables one 8-warp block fills an SM, and a block of statically
// assigned rows lasts as long as its heaviest row.
template <int CAP, bool RMCL>
__global__ void __launch_bounds__(256)
k_num_warp(const int* __restrict__ list, int count, int row_lo,
           const int64_t* __restrict__ Arp, const int* __restrict__ Acol,
           const double* __restrict__ Aval, const int64_t* __restrict__ Brp,
           const int* __restrict__ Bcol, const double* __restrict__ Bval,
           const int64_t* __restrict__ Crp, int* __restrict__ Ccol, double* __restrict__ Cval,
           RmclOut ro, int* __restrict__ work_counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* wbase = smem_raw + (size_t)warp * (24 * CAP);
  if (!work_counter) {
    const int idx = blockIdx.x * (blockDim.x >> 5) + warp;
    if (idx < count)
      num_warp_row<CAP, RMCL>(list[idx], wbase, lane, row_lo, Arp, Acol, Aval, Brp, Bcol, Bval, Crp, Ccol, Cval, ro);
    return;
  }
  while (true) {
    int idx = 0;
    if (lane == 0) idx = atomicAdd(work_counter, 1);
    idx = __shfl_sync(FULL, idx, 0);
    if (idx >= count) break;
    num_warp_row<CAP, RMCL>(list[idx], wbase, lane, row_lo, Arp, Acol, Aval, Brp, Bcol, Bval, Crp, Ccol, Cval, ro);
    __syncwarp();   // the next row re-uses the tables
  }
}

// The numeric row built OPTIMISTICALLY in the symbolic phase (plain SpGEMM): for rows whose
// products would need one of the large tables but whose columns may well fit the smallest ones —
// every interior row of a 27-point stencil has 729 products and 125 columns — the row is
// accumulated at once in a 128-entry table.  If it fits, the sorted row waits in a fixed arena
// slot (list position x 128) for C's row offsets, its length goes into the row counts, and the
// row needs neither a symbolic pass nor a second numeric one; if not, the row is handed to the
// regular two-pass path (overflow_list).
//
// A leaner loop than k_num_warp's (which keeps first-touch order for the rMCL epilogue): the
// value sits AT the key's hash slot (no slot indirection), a new key is claimed with one
// shared-memory CAS, the (col, val) loads of the next step are issued before the current step is
// accumulated, and the row is sorted in registers (4 columns per lane, shuffles) with the values
// looked up by column afterwards.  Every entry still receives its products in ascending A-entry
// order with separately rounded multiply and add: the same bits as k_num_warp and the reference.
// shared memory per warp (28 * CAP bytes): vals fp64[2 CAP] | keys int[2 CAP] | cols int[CAP]
__device__ __forceinline__ void cmpex_u32(unsigned& a, unsigned& b, bool up) {
  const unsigned lo = min(a, b), hi = max(a, b);
  a = up ? lo : hi;
  b = up ? hi : lo;
}
// bitonic sort of 128 keys held 4 per lane: element e = 4 * lane + r, ascending
__device__ __forceinline__ void warp_sort128_regs(unsigned (&key)[4], int lane) {
#pragma unroll
  for (int k = 2; k <= 128; k <<= 1) {
    // direction of the length-k run element e belongs to: bit log2(k) of e (k = 128: ascending)
    const bool up = (k >= 4) ? ((lane & (k >> 2)) == 0) : true;
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 4) {
        const bool lower = (lane & (j >> 2)) == 0;
        const bool keepmin = lower == up;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const unsigned other = __shfl_xor_sync(FULL, key[r], j >> 2);
          key[r] = keepmin ? min(key[r], other) : max(key[r], other);
        }
      } else if (j == 2) {
        // k >= 4 here; (r, r + 2) pairs
        cmpex_u32(key[0], key[2], up);
        cmpex_u32(key[1], key[3], up);
      } else {  // j == 1: (0,1) and (2,3); for k == 2 the direction is bit 1 of e, i.e. of r
        cmpex_u32(key[0], key[1], k == 2 ? true : up);
        cmpex_u32(key[2], key[3], k == 2 ? false : up);
      }
    }
  }
}

// IDX32: B has fewer than 2^31 entries — its row offsets travel between lanes in one register
template <int CAP, bool IDX32>
__global__ void __launch_bounds__(256)
k_num_warp_fused(const int* __restrict__ list, int count, int row_lo,
                 const int64_t* __restrict__ Arp, const int* __restrict__ Acol,
                 const double* __restrict__ Aval, const int64_t* __restrict__ Brp,
                 const int* __restrict__ Bcol, const double* __restrict__ Bval,
                 long long arena_base, int* __restrict__ arena_col, double* __restrict__ arena_val,
                 int64_t* __restrict__ arena_off, int* __restrict__ rownnz,
                 unsigned char* __restrict__ fused, int* __restrict__ overflow_list,
                 int* __restrict__ overflow_count) {
  static_assert(CAP == 128, "the register sort holds 4 x 32 columns");
  constexpr int H = 2 * CAP;
  constexpr unsigned mask = H - 1;
  using idx_t = typename std::conditional<IDX32, unsigned, unsigned long long>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int idx = blockIdx.x * (blockDim.x >> 5) + warp;
  if (idx >= count) return;
  const int i = list[idx];
  unsigned char* wbase = smem_raw + (size_t)warp * (28 * CAP);
  double* vals = (double*)wbase;
  int* keys = (int*)(wbase + 16 * CAP);
  int* cols = (int*)(wbase + 24 * CAP);
#pragma unroll
  for (int k = 0; k < H / 32; ++k) keys[k * 32 + lane] = EMPTY;
  __syncwarp();
  const int64_t a0 = Arp[row_lo + i], a1 = Arp[row_lo + i + 1];
  int cnt = 0;
  bool over = false;

  // a step = 32 consecutive entries of one B row
  struct Step { idx_t s; int t, o0, len, c; double a, v; };
  idx_t bs = 0;
  int bl = 0;
  double av = 0.0;
  auto fetch = [&](Step& n, int t, int o0) {   // issue the (col, val) loads of step (t, o0)
    n.t = t;
    n.o0 = o0;
    if (IDX32) n.s = (idx_t)__shfl_sync(FULL, (unsigned)bs, t);
    else n.s = (idx_t)shfl64((long long)bs, t);
    n.len = __shfl_sync(FULL, bl, t);
    n.a = shfld(av, t);
    n.c = -1;   // -1 (EMPTY is never a column): this lane has no product in the step
    n.v = 0.0;
    if (o0 + lane < n.len) {
      const idx_t q = n.s + (idx_t)(o0 + lane);
      n.c = __ldg(Bcol + q);
      n.v = __ldg(Bval + q);
    }
  };
  // find or claim the slot of the step's column (the columns of a step are pairwise distinct, so
  // a lane can lose a slot only to a different column), accumulate there; false: row overflows
  auto process = [&](const Step& st) -> bool {
    const int c = st.c;
    const bool act = c >= 0;
    const double prod = __dmul_rn(st.a, st.v);
    unsigned h = hash_col<H>(c);
    int k = act ? keys[h] : c;
    int state = (k == c) ? 1 : 0;   // 0 pending, 1 present (or no product), 2 claimed now
    if (!__all_sync(FULL, state != 0)) {
      while (true) {
        if (state == 0) {
          if (k == EMPTY && atomicCAS(&keys[h], EMPTY, c) == EMPTY) {
            state = 2;
          } else {
            h = (h + 1) & mask;
            k = keys[h];
            if (k == c) state = 1;
          }
        }
        if (__all_sync(FULL, state != 0)) break;
      }
      const unsigned newmask = __ballot_sync(FULL, state == 2);
      const int nnew = __popc(newmask);
      // (2 CAP slots, at most CAP + 31 columns when this fires: the table is never full)
      if (cnt + nnew > CAP) return false;
      if (state == 2) {
        vals[h] = prod;
        cols[cnt + __popc(newmask & lanemask_lt())] = c;
      }
      cnt += nnew;
    }
    if (state == 1 && act) vals[h] = __dadd_rn(vals[h], prod);
    __syncwarp();
    return true;
  };

  for (int64_t abase = a0; abase < a1 && !over; abase += 32) {
    const int64_t p = abase + lane;
    bs = 0;
    bl = 0;
    av = 0.0;
    if (p < a1) {
      const int j = __ldg(Acol + p);
      av = __ldg(Aval + p);
      const long long b0 = __ldg(Brp + j);
      bs = (idx_t)b0;
      bl = (int)(__ldg(Brp + j + 1) - b0);
    }
    const int nn = (int)min((int64_t)32, a1 - abase);
    // two step records used alternately: the loads of the next step are in flight while the
    // current one is accumulated, and no register moves between them
    Step A, B;
    fetch(A, 0, 0);
    while (true) {
      int nt = A.t, no = A.o0 + 32;
      if (no >= A.len) { ++nt; no = 0; }
      const bool moreB = nt < nn;
      if (moreB) fetch(B, nt, no);
      if (!process(A)) { over = true; break; }
      if (!moreB) break;
      nt = B.t; no = B.o0 + 32;
      if (no >= B.len) { ++nt; no = 0; }
      const bool moreA = nt < nn;
      if (moreA) fetch(A, nt, no);
      if (!process(B)) { over = true; break; }
      if (!moreA) break;
    }
  }
  if (over) {
    if (lane == 0) overflow_list[atomicAdd(overflow_count, 1)] = i;
    return;
  }
  // ---- ascending columns: sort in registers, look the values up by column
  unsigned key[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int e = 4 * lane + r;
    key[r] = (e < cnt) ? (unsigned)cols[e] : 0xffffffffu;   // (columns are < 2^31: the pad sorts last)
  }
  warp_sort128_regs(key, lane);
  const long long ob = arena_base + (long long)idx * CAP;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int e = 4 * lane + r;
    if (e < cnt) {
      const int cc = (int)key[r];
      unsigned h = hash_col<H>(cc);
      while (keys[h] != cc) h = (h + 1) & mask;
      arena_col[ob + e] = cc;
      arena_val[ob + e] = vals[h];
    }
  }
  if (lane == 0) { arena_off[i] = ob; rownnz[i] = cnt; fused[i] = 1; }
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

// exclusive scan of one 64-bit value per thread; s_red64 needs BT/32 + 1 slots.  The warp
// totals are scanned by warp 0 with shuffles (not by every thread with a 32-step loop).
__device__ __forceinline__ long long shfl_up64(long long v, int o) {
  const int lo = __shfl_up_sync(FULL, (int)(v & 0xffffffffLL), o);
  const int hi = __shfl_up_sync(FULL, (int)(v >> 32), o);
  return ((long long)hi << 32) | (unsigned)lo;
}
template <int BT>
__device__ __forceinline__ long long block_excl_scan64(long long v, long long* s_red64,
                                                       long long* total) {
  static_assert(BT / 32 <= 32, "one warp scans the warp totals");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long y = shfl_up64(inc, o);
    if (lane >= o) inc += y;
  }
  __syncthreads();
  if (lane == 31) s_red64[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const long long t = (lane < BT / 32) ? s_red64[lane] : 0;
    long long ti = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long y = shfl_up64(ti, o);
      if (lane >= o) ti += y;
    }
    if (lane < BT / 32) s_red64[lane] = ti - t;
    if (lane == 31) s_red64[BT / 32] = ti;
  }
  __syncthreads();
  *total = s_red64[BT / 32];
  return s_red64[warp] + inc - v;
}

// Shared-memory work area of the flat product walk: one slot per A entry of the current batch.
template <int BT>
struct WalkSmem {
  long long end[BT];   // inclusive end of the entry's products in the batch's flat index space
  long long base[BT];  // B-row start minus the entry's first flat index: q = base + x
  double a[BT];        // A value of the entry
  long long red[BT / 32 + 2];  // + total, + pad: the bitmap behind this struct is 16-byte aligned
};
static_assert(sizeof(WalkSmem<1024>) % 16 == 0 && sizeof(WalkSmem<512>) % 16 == 0 &&
                  sizeof(WalkSmem<256>) % 16 == 0,
              "bitmap must stay 16-byte aligned");

// Walk the products of one A row with the whole CTA, FLAT: the products of a batch of up to BT
// A entries are numbered 0..T-1 in (A entry, position in its B row) order.  Every thread is busy
// whatever the shape of the row — 5 A entries with B rows of 1500 columns (the typical heavy
// R-MAT row) or 30 000 entries of mixed lengths — and consecutive lanes read consecutive
// elements of a B row.  Three steps:
//   walk_prepare   load the batch's A entries, scan their B-row lengths into shared memory;
//   walk_prefetch  ask L2 for the B-row segments (one line per 32 columns, two per 32 values);
//                  issued early in a row's life so the DRAM latency of the gather overlaps the
//                  bitmap / prefix / column-emission phases instead of stalling the products;
//   walk_run       each warp takes a contiguous slice of the flat index space (a multiple of 32
//                  long), its lanes consecutive indices; four products are in flight per thread.
template <int BT, bool WITH_VAL>
__device__ __forceinline__ long long walk_prepare(int64_t b0, int nb, const int* __restrict__ Acol,
                                                  const double* __restrict__ Aval,
                                                  const int64_t* __restrict__ Brp,
                                                  WalkSmem<BT>& ws) {
  long long len = 0, bs = 0;
  double a = 0.0;
  if ((int)threadIdx.x < nb) {
    const int j = __ldg(Acol + b0 + threadIdx.x);
    if (WITH_VAL) a = __ldg(Aval + b0 + threadIdx.x);
    bs = __ldg(Brp + j);
    len = __ldg(Brp + j + 1) - bs;
  }
  long long total;
  const long long ex = block_excl_scan64<BT>(len, ws.red, &total);
  ws.end[threadIdx.x] = ((int)threadIdx.x < nb) ? ex + len : total;  // pad: never matched
  ws.base[threadIdx.x] = bs - ex;
  ws.a[threadIdx.x] = a;
  __syncthreads();
  return total;
}

// slice of the flat index space owned by this warp, and the entry holding its first index
// (a team of CTAs can share one batch: member r of T takes the r-th T-th of the index space)
template <int BT>
__device__ __forceinline__ bool walk_slice(const WalkSmem<BT>& ws, int nb, long long total,
                                           long long* xb, long long* xe, int* e0, int r = 0,
                                           int T = 1) {
  constexpr int NWARPS = BT / 32;
  const int warp = threadIdx.x >> 5;
  long long glo = 0, ghi = total;
  if (T > 1) {
    const long long per_member = (((total + T - 1) / T) + 31) & ~31LL;
    glo = min(total, (long long)r * per_member);
    ghi = min(total, glo + per_member);
  }
  const long long per_warp = ((((ghi - glo) + NWARPS - 1) / NWARPS) + 31) & ~31LL;
  *xb = glo + (long long)warp * per_warp;
  *xe = min(ghi, *xb + per_warp);
  if (*xb >= *xe) return false;
  int lo = 0, hi = nb - 1;  // xb < total guarantees an answer
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (*xb < ws.end[mid]) hi = mid; else lo = mid + 1;
  }
  *e0 = lo;
  return true;
}

template <int BT, bool WITH_VAL>
__device__ __forceinline__ void walk_prefetch(const WalkSmem<BT>& ws, int nb, long long total,
                                              const int* __restrict__ Bcol,
                                              const double* __restrict__ Bval, int r = 0,
                                              int T = 1) {
  long long xb, xe;
  int e;
  if (!walk_slice<BT>(ws, nb, total, &xb, &xe, &e, r, T)) return;
  const int lane = threadIdx.x & 31;
  // lane l looks after the 32-index step starting at xb + 32*l (then +1024, ...): one line of
  // columns and two of values per step, addressed by the step's first element
  for (long long x = xb + 32LL * lane; x < xe; x += 1024) {
    while (x >= ws.end[e]) ++e;
    const long long q = ws.base[e] + x;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(Bcol + q));
    if (WITH_VAL) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(Bval + q));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(Bval + q + 16));
    }
  }
}

template <int BT, bool WITH_VAL, typename F>
__device__ __forceinline__ void walk_run(const WalkSmem<BT>& ws, int nb, long long total,
                                         const int* __restrict__ Bcol,
                                         const double* __restrict__ Bval, unsigned long long bpol,
                                         F f, int r = 0, int T = 1) {
  long long xb, xe;
  int e;
  if (!walk_slice<BT>(ws, nb, total, &xb, &xe, &e, r, T)) return;
  const int lane = threadIdx.x & 31;
  long long x = xb + lane;
  // eight products in flight while they all lie in one B row (long rows: the common case)
  for (; x + 224 < xe; x += 256) {
    while (x >= ws.end[e]) ++e;
    if (x + 224 >= ws.end[e]) break;
    const long long base = ws.base[e] + x;
    const double a = ws.a[e];
    int col[8];
    double bv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      col[u] = ldg_hint(Bcol + base + 32 * u, bpol);
      if (WITH_VAL) bv[u] = ldg_hint(Bval + base + 32 * u, bpol);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) f(col[u], WITH_VAL ? __dmul_rn(a, bv[u]) : 0.0);
  }
  for (; x + 96 < xe; x += 128) {
    long long q[4];
    double av[4];
    while (x >= ws.end[e]) ++e;
    if (x + 96 < ws.end[e]) {  // all four in the same B row
      const long long base = ws.base[e] + x;
      const double a = ws.a[e];
#pragma unroll
      for (int u = 0; u < 4; ++u) { q[u] = base + 32 * u; av[u] = a; }
    } else {
      int eu = e;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        while (x + 32 * u >= ws.end[eu]) ++eu;
        q[u] = ws.base[eu] + x + 32 * u;
        av[u] = ws.a[eu];
      }
    }
    int col[4];
    double bv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      col[u] = ldg_hint(Bcol + q[u], bpol);
      if (WITH_VAL) bv[u] = ldg_hint(Bval + q[u], bpol);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) f(col[u], WITH_VAL ? __dmul_rn(av[u], bv[u]) : 0.0);
  }
  for (; x < xe; x += 32) {
    while (x >= ws.end[e]) ++e;
    const long long q = ws.base[e] + x;
    const int col = ldg_hint(Bcol + q, bpol);
    f(col, WITH_VAL ? __dmul_rn(ws.a[e], ldg_hint(Bval + q, bpol)) : 0.0);
  }
}

// walk_run with DYNAMIC distribution: the warps of the CTA draw chunks of CH consecutive flat
// indices of [glo, ghi) from a shared counter (*next, set to glo by the caller before a
// barrier), so a warp that was served from L2 takes more chunks than one that waited for DRAM
// and the CTA reaches the closing barrier together.  One binary search per chunk.
template <int BT, int CH, bool WITH_VAL, typename F>
__device__ __forceinline__ void walk_run_dynamic(const WalkSmem<BT>& ws, int nb, long long glo,
                                                 long long ghi, long long* next,
                                                 const int* __restrict__ Bcol,
                                                 const double* __restrict__ Bval,
                                                 unsigned long long bpol, F f) {
  static_assert(CH % 128 == 0, "a chunk is a whole number of 4-deep warp steps");
  const int lane = threadIdx.x & 31;
  while (true) {
    long long xb = 0;
    if (lane == 0) xb = (long long)atomicAdd((unsigned long long*)next, (unsigned long long)CH);
    xb = shfl64(xb, 0);
    if (xb >= ghi) break;
    const long long xe = min(ghi, xb + CH);
    int e;
    {
      int lo = 0, hi = nb - 1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (xb < ws.end[mid]) hi = mid; else lo = mid + 1;
      }
      e = lo;
    }
    long long x = xb + lane;
    // eight products in flight while they all lie in one B row (long rows: the common case)
    for (; x + 224 < xe; x += 256) {
      while (x >= ws.end[e]) ++e;
      if (x + 224 >= ws.end[e]) break;
      const long long base = ws.base[e] + x;
      const double a = ws.a[e];
      int col[8];
      double bv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        col[u] = ldg_hint(Bcol + base + 32 * u, bpol);
        if (WITH_VAL) bv[u] = ldg_hint(Bval + base + 32 * u, bpol);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) f(col[u], WITH_VAL ? __dmul_rn(a, bv[u]) : 0.0);
    }
    for (; x + 96 < xe; x += 128) {
      long long q[4];
      double av[4];
      while (x >= ws.end[e]) ++e;
      if (x + 96 < ws.end[e]) {
        const long long base = ws.base[e] + x;
        const double a = ws.a[e];
#pragma unroll
        for (int u = 0; u < 4; ++u) { q[u] = base + 32 * u; av[u] = a; }
      } else {
        int eu = e;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          while (x + 32 * u >= ws.end[eu]) ++eu;
          q[u] = ws.base[eu] + x + 32 * u;
          av[u] = ws.a[eu];
        }
      }
      int col[4];
      double bv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        col[u] = ldg_hint(Bcol + q[u], bpol);
        if (WITH_VAL) bv[u] = ldg_hint(Bval + q[u], bpol);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) f(col[u], WITH_VAL ? __dmul_rn(av[u], bv[u]) : 0.0);
    }
    for (; x < xe; x += 32) {
      while (x >= ws.end[e]) ++e;
      const long long q = ws.base[e] + x;
      const int col = ldg_hint(Bcol + q, bpol);
      f(col, WITH_VAL ? __dmul_rn(ws.a[e], ldg_hint(Bval + q, bpol)) : 0.0);
    }
  }
}

// prepare + run over every batch of the row (no early prefetch)
template <int BT, bool WITH_VAL, typename F>
__device__ __forceinline__ void cta_products_flat(int64_t a0, int64_t a1,
                                                  const int* __restrict__ Acol,
                                                  const double* __restrict__ Aval,
                                                  const int64_t* __restrict__ Brp,
                                                  const int* __restrict__ Bcol,
                                                  const double* __restrict__ Bval,
                                                  WalkSmem<BT>& ws, unsigned long long bpol,
                                                  F f) {
  for (int64_t b0 = a0; b0 < a1; b0 += BT) {
    const int nb = (int)min((int64_t)BT, a1 - b0);
    const long long total = walk_prepare<BT, WITH_VAL>(b0, nb, Acol, Aval, Brp, ws);
    walk_run<BT, WITH_VAL>(ws, nb, total, Bcol, Bval, bpol, f);
    __syncthreads();
  }
}

// set bit c of a shared/global bitmap; the plain read first turns most of the work of a row
// with many repeated columns into loads (an atomic OR only for a bit not seen yet)
__device__ __forceinline__ void bitmap_set(unsigned* bm32, int c) {
  const unsigned bit = 1u << (c & 31);
  unsigned* w = bm32 + (c >> 5);
  if (!(*(volatile unsigned*)w & bit)) atomicOr(w, bit);
}

constexpr int PARTS_MAX = 16;   // column parts of the part-wise kernels (16 x 512 K columns = 8 M)
constexpr int PARTS_WHOLE = 4;  // ... of which the whole-row symbolic kernel can see at most 4
                                // (its bitmap holds <= 1.8 M columns)
// column emission: a bitmap word with at least this many columns is expanded by the whole warp
// (~14 instructions), a sparser one bit by bit by its own lane (~8 instructions per bit, for the
// whole warp); 16 measured best on R-MAT 20 (8: -3 %)
constexpr int DENSE_WORD = 16;

// ------------------------------------------------------------------------------------------
// symbolic for large rows: CTA per row (persistent, dynamic row fetch, heaviest rows first),
// column bitmap of nw64 64-bit words either in shared memory (SMEM_BM) or in a per-CTA HBM
// scratch.  The bitmaps of the first `store_rows` listed rows are kept in `bm_store` (slot =
// list position, recorded in bm_slot[row]) for the numeric pass; the others are rebuilt there.
template <int BT, bool SMEM_BM>
__global__ void __launch_bounds__(BT, 1)
k_sym_bitmap(const int* __restrict__ list, int count, int row_lo,
             const int64_t* __restrict__ Arp, const int* __restrict__ Acol,
             const int64_t* __restrict__ Brp, const int* __restrict__ Bcol,
             const long long* __restrict__ flops, int nw64,
             unsigned long long* __restrict__ gscratch, unsigned long long* __restrict__ bm_store,
             int store_rows, int* __restrict__ bm_slot, int* __restrict__ rownnz,
             int nparts, int wpp, int* __restrict__ partcnt, int* __restrict__ work_counter,
             L2Modes l2, const unsigned char* __restrict__ wr, int R, int* __restrict__ rcnt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_pc[BT / 32][PARTS_WHOLE];
  __shared__ int s_idx;
  // output columns per static column range (ranges.cuh), for the rows whose bitmap is stored
  __shared__ int s_rc[128];
  if (threadIdx.x < 128) s_rc[threadIdx.x] = 0;
  WalkSmem<BT>& ws = *reinterpret_cast<WalkSmem<BT>*>(smem_raw);
  unsigned long long* bm = SMEM_BM ? (unsigned long long*)(smem_raw + sizeof(WalkSmem<BT>))
                                   : gscratch + (size_t)blockIdx.x * nw64;
  unsigned* bm32 = (unsigned*)bm;
  const unsigned long long pol_b = l2_policy(l2.bgather), pol_bm = l2_policy(l2.bmstore);
  for (int w = threadIdx.x; w < nw64; w += BT) bm[w] = 0ull;
  while (true) {
    __syncthreads();
    if (threadIdx.x == 0) s_idx = atomicAdd(work_counter, 1);
    __syncthreads();
    const int idx = s_idx;
    if (idx >= count) break;
    const int i = list[idx];
    const int64_t a0 = Arp[row_lo + i], a1 = Arp[row_lo + i + 1];
    cta_products_flat<BT, false>(a0, a1, Acol, (const double*)nullptr, Brp, Bcol,
                                 (const double*)nullptr, ws, pol_b,
                                 [&](int c, double) { bitmap_set(bm32, c); });
    // count (per column part, see k_num_bitmap_part), store and clear in one sweep; wpp is a
    // multiple of BT, so the part of a sweep step is the same for the whole CTA
    int pc[PARTS_WHOLE] = {0, 0, 0, 0};
    unsigned long long* dst = (bm_store && idx < store_rows) ? bm_store + (size_t)idx * nw64 : nullptr;
    const bool want_rc = rcnt != nullptr && dst != nullptr;
    for (int w0 = 0, h = 0, hnext = wpp; w0 < nw64; w0 += BT) {
      const int w = w0 + threadIdx.x;
      if (w0 >= hnext) { ++h; hnext += wpp; }
      int c = 0;
      if (w < nw64) {
        const unsigned long long x = bm[w];
        c = __popcll(x);
#pragma unroll
        for (int k = 0; k < PARTS_WHOLE; ++k) pc[k] += (k == h) ? c : 0;
        if (dst) stg_hint(dst + w, x, pol_bm);
        bm[w] = 0ull;
      }
      // the 32 words of a warp are consecutive: they lie in one range, or in a few
      const int wf = w0 + (threadIdx.x & ~31);
      if (want_rc && wf < nw64) {
        const int rf = wr[wf], rl = wr[min(nw64 - 1, wf + 31)];
        const int myr = (rf == rl || w >= nw64) ? rf : (int)wr[w];
        for (int r = rf; r <= rl; ++r) {
          const int sr = warp_sum_int(myr == r ? c : 0);
          if ((threadIdx.x & 31) == 0 && sr) atomicAdd(&s_rc[r], sr);
        }
      }
    }
    {
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
      for (int k = 0; k < PARTS_WHOLE; ++k) pc[k] = warp_sum_int(pc[k]);
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < PARTS_WHOLE; ++k) s_pc[warp][k] = pc[k];
      }
      __syncthreads();
    }
    int cnt = 0;
    if (threadIdx.x == 0) {
      int tot[PARTS_WHOLE] = {0, 0, 0, 0};
      for (int wq = 0; wq < BT / 32; ++wq)
#pragma unroll
        for (int k = 0; k < PARTS_WHOLE; ++k) tot[k] += s_pc[wq][k];
#pragma unroll
      for (int k = 0; k < PARTS_WHOLE; ++k) cnt += tot[k];
      if (partcnt)
#pragma unroll
        for (int k = 0; k < PARTS_WHOLE; ++k) partcnt[(size_t)i * PARTS_MAX + k] = tot[k];
    }
    if (threadIdx.x == 0) {
      rownnz[i] = cnt;
      bm_slot[i] = dst ? idx : -1;
    }
    if (want_rc && (int)threadIdx.x < R) {  // (s_rc complete: the barrier after s_pc above)
      rcnt[(size_t)idx * R + threadIdx.x] = s_rc[threadIdx.x];
      s_rc[threadIdx.x] = 0;
    }
  }
}

// numeric for large rows: rank(col) = prefix[col/64] + popc(bitmap word below col); products
// are reduced with fp64 RED into `acc` (the row's final slice of C.val for SpGEMM, a per-CTA
// scratch for rMCL, where the epilogue then runs over the scratch).
// bm_slot[row] >= 0: the row's bitmap was stored by the symbolic pass at that slot;
// < 0: rebuild it here.
template <int BT, bool SMEM_BM, bool RMCL>
__global__ void __launch_bounds__(BT, 1)
k_num_bitmap(const int* __restrict__ list, int count, int row_lo,
             const int64_t* __restrict__ Arp, const int* __restrict__ Acol,
             const double* __restrict__ Aval, const int64_t* __restrict__ Brp,
             const int* __restrict__ Bcol, const double* __restrict__ Bval,
             const long long* __restrict__ flops, int nw64,
             unsigned long long* __restrict__ gscratch, const unsigned long long* __restrict__ bm_store,
             const int* __restrict__ bm_slot, const int64_t* __restrict__ Crp,
             int* __restrict__ Ccol, double* __restrict__ Cval, int* __restrict__ scr_col,
             double* __restrict__ scr_val, long long scr_stride, RmclOut ro,
             int* __restrict__ work_counter, unsigned long long* __restrict__ prof,
             L2Modes l2) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_red[BT / 32];
  __shared__ double s_redd[BT / 32];
  // developer diagnostics (B200_PROF=1): cycles per phase, summed over this CTA's rows
  long long pc[5] = {0, 0, 0, 0, 0};
  long long tprev = 0;
  auto lap = [&](int k) {
    if (prof && threadIdx.x == 0) { const long long t = clock64(); pc[k] += t - tprev; tprev = t; }
  };
  __shared__ int s_idx;
  // walk area, then bitmap words, then 32-bit exclusive popcount prefix per word
  WalkSmem<BT>& ws = *reinterpret_cast<WalkSmem<BT>*>(smem_raw);
  unsigned long long* bm =
      SMEM_BM ? (unsigned long long*)(smem_raw + sizeof(WalkSmem<BT>))
              : gscratch + (size_t)blockIdx.x * ((size_t)nw64 + ((size_t)nw64 + 1) / 2);
  unsigned* pref = (unsigned*)(bm + nw64);
  unsigned* bm32 = (unsigned*)bm;
  const unsigned long long pol_acc = l2_policy(l2.acc), pol_ocol = l2_policy(l2.ocol),
                           pol_b = l2_policy(l2.bgather), pol_bm = l2_policy(l2.bmstore);
  while (true) {
    __syncthreads();
    if (threadIdx.x == 0) s_idx = atomicAdd(work_counter, 1);
    __syncthreads();
    const int idx = s_idx;
    if (idx >= count) break;
    if (prof && threadIdx.x == 0) tprev = clock64();
    const int i = list[idx];
    const int64_t a0 = Arp[row_lo + i], a1 = Arp[row_lo + i + 1];
    const int slotno = bm_slot[i];
    // Stored bitmap: the walk area is free until the products, so the first batch of A entries
    // is prepared now and its B-row segments are requested from L2 ahead of time.
    const int nb0 = (int)min((int64_t)BT, a1 - a0);
    long long total0 = 0;
    const bool early = slotno >= 0 && nb0 > 0;
    if (early) {
      total0 = walk_prepare<BT, true>(a0, nb0, Acol, Aval, Brp, ws);
      walk_prefetch<BT, true>(ws, nb0, total0, Bcol, Bval);
    }
    if (slotno >= 0) {
      const ulonglong2* src = reinterpret_cast<const ulonglong2*>(bm_store + (size_t)slotno * nw64);
      ulonglong2* dst2 = reinterpret_cast<ulonglong2*>(bm);
      for (int w = threadIdx.x; w < (nw64 >> 1); w += BT) dst2[w] = ldg_hint(src + w, pol_bm);
    } else {
      for (int w = threadIdx.x; w < nw64; w += BT) bm[w] = 0ull;
      cta_products_flat<BT, false>(a0, a1, Acol, (const double*)nullptr, Brp, Bcol,
                                   (const double*)nullptr, ws, pol_b,
                                   [&](int c, double) { bitmap_set(bm32, c); });
    }
    __syncthreads();
    lap(0);
    // popcount prefix.  Warp wp owns a contiguous run of wpw words and walks it 32 words at a
    // time (lane = word: conflict-free shared-memory access), scanning the popcounts with
    // shuffles; the warp totals are then scanned across the CTA and added back.
    constexpr int NW = BT / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wpw = (((nw64 + NW - 1) / NW) + 31) & ~31;
    const int wbeg = warp * wpw, wend = min(nw64, wbeg + wpw);
    {
      int run = 0;
      for (int w = wbeg + lane; w < wbeg + wpw; w += 32) {
        const int c = (w < wend) ? __popcll(bm[w]) : 0;
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int y = __shfl_up_sync(FULL, inc, o);
          if (lane >= o) inc += y;
        }
        if (w < wend) pref[w] = (unsigned)(run + inc - c);
        run += __shfl_sync(FULL, inc, 31);
      }
      __syncthreads();
      if (lane == 0) s_red[warp] = run;
      __syncthreads();
    }
    int cnt = 0, wbase = 0;
#pragma unroll
    for (int k = 0; k < NW; ++k) {
      const int x = s_red[k];
      if (k < warp) wbase += x;
      cnt += x;
    }
    for (int w = wbeg + lane; w < wend; w += 32) pref[w] += (unsigned)wbase;
    __syncthreads();
    lap(1);
    double* acc = RMCL ? scr_val + (size_t)blockIdx.x * scr_stride : Cval + Crp[i];
    int* ocol = RMCL ? scr_col + (size_t)blockIdx.x * scr_stride : Ccol + Crp[i];
    // The accumulators zeroed, and the columns in ascending order straight from the bitmap.
    // Groups of 32 words are dealt round-robin to the warps (the dense low-column region of a
    // power-law row is shared by all of them), lane = word.  A dense word (>= DENSE_WORD columns) is
    // expanded by the whole warp — lane l tests bits l and l+32, the position is a popcount —
    // a sparse word by its own lane, bit by bit.  The positions of a group are consecutive
    // (pref[]), so either way a store instruction lands in a few adjacent sectors.
    for (int k = threadIdx.x; k < cnt; k += BT) stg_hint(acc + k, 0.0, pol_acc);
    {
      const int ngroups = (nw64 + 31) >> 5;
      const unsigned lt = lanemask_lt();
      for (int g = warp; g < ngroups; g += NW) {
        const int w = g * 32 + lane;
        unsigned long long x = (w < nw64) ? bm[w] : 0ull;
        int pos = (w < nw64) ? (int)pref[w] : 0;
        const bool is_dense = __popcll(x) >= DENSE_WORD;
        unsigned dense = __ballot_sync(FULL, is_dense);
        while (dense) {
          const int src = __ffs((int)dense) - 1;
          dense &= dense - 1;
          const unsigned long long xs = (unsigned long long)shfl64((long long)x, src);
          const int ps = __shfl_sync(FULL, pos, src);
          const unsigned lo32 = (unsigned)xs, hi32 = (unsigned)(xs >> 32);
          const int cb = (g * 32 + src) * 64;
          if ((lo32 >> lane) & 1u) stg_hint(ocol + ps + __popc(lo32 & lt), cb + lane, pol_ocol);
          if ((hi32 >> lane) & 1u)
            stg_hint(ocol + ps + __popc(lo32) + __popc(hi32 & lt), cb + 32 + lane, pol_ocol);
        }
        if (is_dense) x = 0ull;
        const int cb = w * 64;
        while (x) {
          const int b = __ffsll((long long)x) - 1;
          x &= x - 1;
          stg_hint(ocol + pos++, cb + b, pol_ocol);
        }
      }
    }
    lap(2);
    {
      auto accumulate = [&](int c, double prod) {
        const int w = c >> 6;
        const unsigned long long below = bm[w] & ((1ull << (c & 63)) - 1ull);
        const int rank = (int)pref[w] + __popcll(below);
        red_add_hint(acc + rank, prod, pol_acc);
      };
      __syncthreads();  // acc[] zeroed by every thread before the first RED lands
      for (int64_t b0 = a0; b0 < a1; b0 += BT) {
        const int nb = (int)min((int64_t)BT, a1 - b0);
        const long long total = (early && b0 == a0)
                                    ? total0
                                    : walk_prepare<BT, true>(b0, nb, Acol, Aval, Brp, ws);
        walk_run<BT, true>(ws, nb, total, Bcol, Bval, pol_b, accumulate);
        __syncthreads();
      }
    }
    if (l2.demote)  // the row is complete: its accumulator lines need no protection any more
      for (int k = threadIdx.x * 16; k < cnt; k += BT * 16) l2_demote(acc + k);
    lap(3);
    if (!RMCL) continue;
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
    // opt-in top-k (see topk_cut): block-wide counts
    double tk = 0.0;
    int cstar = 0x7fffffff;
    bool cut = false;
    if (ro.topk > 0) {
      auto bcount = [&](auto pred) {
        int c = 0;
        for (int k = threadIdx.x; k < cnt; k += BT) c += (acc[k] >= thresh && pred(acc[k], ocol[k])) ? 1 : 0;
        return block_sum_int<BT>(c, s_red);
      };
      if (bcount([](double, int) { return true; }) > ro.topk) {
        cut = true;
        topk_cut(bcount, thresh, rmax, ro.topk, 0x7fffffff, &tk, &cstar);
      }
    }
    auto keeps = [&](double v, int c) { return v >= thresh && (!cut || v > tk || (v == tk && c <= cstar)); };
    double ksum_p = 0.0;
    int kept_p = 0;
    for (int k = threadIdx.x; k < cnt; k += BT) {
      const double v2 = acc[k];
      if (keeps(v2, ocol[k])) { ksum_p = __dadd_rn(ksum_p, v2); ++kept_p; }
    }
    const double ksum = block_sum_d<BT>(ksum_p, s_redd);
    const int kept = block_sum_int<BT>(kept_p, s_red);
    const long long off = (long long)Crp[i];
    double sq_p = 0.0;
    int written = 0;
    for (int k0 = 0; k0 < cnt; k0 += BT) {
      const int k = k0 + threadIdx.x;
      const bool keep = (k < cnt) && keeps(acc[k], ocol[k]);
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
    lap(4);
  }
  if (prof && threadIdx.x == 0)
    for (int k = 0; k < 5; ++k) atomicAdd(prof + k, (unsigned long long)pc[k]);
}

// first position in the sorted run col[lo..hi) whose column is >= key
__device__ __forceinline__ long long lower_bound_col(const int* __restrict__ col, long long lo,
                                                     long long hi, int key) {
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (__ldg(col + mid) < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// offsets of the column-part boundaries inside every (sorted) B row: one binary search per row
// and boundary, once per call, instead of one per A entry and work item
__global__ void __launch_bounds__(256)
k_bsplit(const int64_t* __restrict__ Brp, const int* __restrict__ Bcol, int krows, int nparts,
         int wpp, int* __restrict__ bsplit) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)krows * (nparts - 1)) return;
  const int p = (int)(t / krows) + 1, j = (int)(t % krows);
  const long long s = Brp[j], e = Brp[j + 1];
  bsplit[t] = (int)(lower_bound_col(Bcol, s, e, p * wpp * 64) - s);
}

// walk_prepare restricted to one column part of every B row (rows sorted): the segment of a B
// row inside the part comes from the k_bsplit table, so the products outside the part are
// never read
template <int BT>
__device__ __forceinline__ long long walk_prepare_part(int64_t b0, int nb,
                                                       const int* __restrict__ Acol,
                                                       const double* __restrict__ Aval,
                                                       const int64_t* __restrict__ Brp,
                                                       const int* __restrict__ bsplit, int krows,
                                                       int h, int nparts, WalkSmem<BT>& ws) {
  long long len = 0, bs = 0;
  double a = 0.0;
  if ((int)threadIdx.x < nb) {
    const int j = __ldg(Acol + b0 + threadIdx.x);
    a = __ldg(Aval + b0 + threadIdx.x);
    // bsplit[(p-1)*krows + j]: offset inside B row j of its first column of part p (k_bsplit)
    const long long r0 = __ldg(Brp + j);
    const long long s = (h == 0) ? r0 : r0 + __ldg(bsplit + (size_t)(h - 1) * krows + j);
    const long long e = (h == nparts - 1) ? __ldg(Brp + j + 1) : r0 + __ldg(bsplit + (size_t)h * krows + j);
    bs = s;
    len = e - s;
  }
  long long total;
  const long long ex = block_excl_scan64<BT>(len, ws.red, &total);
  ws.end[threadIdx.x] = ((int)threadIdx.x < nb) ? ex + len : total;
  ws.base[threadIdx.x] = bs - ex;
  ws.a[threadIdx.x] = a;
  __syncthreads();
  return total;
}

// numeric for large rows, PART-WISE (SpGEMM with sorted B rows): the columns of B are cut into
// nparts parts of wpp bitmap words, a work item is (row, part), handled by a 512-thread CTA with
// the part's bitmap + popcount prefix in shared memory (<= 100 KB) — so TWO CTAs are resident
// per SM and the serial phases of one item (ticket, bitmap load, prefix, column emission)
// overlap the RED-bound product phase of the other.  Item (row, h) writes the slice of the
// output row that starts partcnt[row][0..h) entries in.
template <int BT, int MINB>
__global__ void __launch_bounds__(BT, MINB)
k_num_bitmap_part(const int* __restrict__ list, int count, int nparts, int wpp, int row_lo,
                  const int64_t* __restrict__ Arp, const int* __restrict__ Acol,
                  const double* __restrict__ Aval, const int64_t* __restrict__ Brp,
                  const int* __restrict__ Bcol, const double* __restrict__ Bval, int nw64,
                  const unsigned long long* __restrict__ bm_store,
                  const int* __restrict__ bm_slot, const int* __restrict__ partcnt,
                  const int64_t* __restrict__ Crp, int* __restrict__ Ccol,
                  double* __restrict__ Cval, const int* __restrict__ itemoff,
                  const int* __restrict__ ticket_slot, int* __restrict__ team_ready,
                  const int* __restrict__ bsplit, int krows,
                  int* __restrict__ work_counter, L2Modes l2) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_red[BT / 32];
  __shared__ int s_idx, s_slot;
  __shared__ long long s_next;
  constexpr int NW = BT / 32;
  WalkSmem<BT>& ws = *reinterpret_cast<WalkSmem<BT>*>(smem_raw);
  unsigned long long* bm = (unsigned long long*)(smem_raw + sizeof(WalkSmem<BT>));
  unsigned* pref = (unsigned*)(bm + wpp);
  unsigned* bm32 = (unsigned*)bm;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long pol_acc = l2_policy(l2.acc), pol_ocol = l2_policy(l2.ocol),
                           pol_b = l2_policy(l2.bgather), pol_bm = l2_policy(l2.bmstore);
  // Work items: (row, column part) slot q = listpos * nparts + h owns the tickets
  // [itemoff[q], itemoff[q+1]) — a TEAM of T CTAs for a heavy slot (T ~ products / 96K), one
  // CTA otherwise (tuned on R-MAT scale 20: 96K).  Teams bound how many distinct accumulator slices are in flight (296 CTAs
  // of hub rows would keep ~1 GB of accumulators live and every RED would miss L2) and make
  // item durations uniform.  Team members split the zeroing, the column emission and the
  // products; a member starts its REDs only after all members have zeroed their share
  // (team_ready counter; members hold consecutive tickets, and a CTA only ever waits for
  // holders of earlier-or-adjacent tickets that are already resident or will be fetched next,
  // never the other way round, so the wait cannot deadlock).
  const int nslots = count * nparts;
  const int tickets = itemoff[nslots];
  while (true) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const int t = atomicAdd(work_counter, 1);
      s_idx = t;
      if (t < tickets) s_slot = ticket_slot[t];  // slot q with itemoff[q] <= t < itemoff[q+1]
    }
    __syncthreads();
    const int t = s_idx;
    if (t >= tickets) break;
    const int q = s_slot;
    const int team_T = itemoff[q + 1] - itemoff[q], team_r = t - itemoff[q];
    const int i = list[q / nparts], h = q % nparts;
    const int* pc = partcnt + (size_t)i * PARTS_MAX;
    // pc[0] < 0: the row did not go through k_sym_bitmap (few products but many columns, or a
    // single A entry), so its per-part counts are not known: they are counted here
    const bool known = pc[0] >= 0;
    if ((known && pc[h] == 0) || h * wpp >= nw64) continue;
    const int64_t a0 = Arp[row_lo + i], a1 = Arp[row_lo + i + 1];
    const int slotno = bm_slot[i];
    // builds in bm[] the bitmap of column part hh of this row
    auto build_part = [&](int hh, bool reuse_batch0, long long total_b0) {
      const int wl = hh * wpp, nw = min(wpp, nw64 - wl);
      for (int w = threadIdx.x; w < nw; w += BT) bm[w] = 0ull;
      __syncthreads();
      for (int64_t b0 = a0; b0 < a1; b0 += BT) {
        const int nb = (int)min((int64_t)BT, a1 - b0);
        const long long total =
            (reuse_batch0 && b0 == a0)
                ? total_b0
                : walk_prepare_part<BT>(b0, nb, Acol, Aval, Brp, bsplit, krows, hh, nparts, ws);
        walk_run<BT, false>(ws, nb, total, Bcol, (const double*)nullptr, pol_b,
                            [&](int c, double) { bitmap_set(bm32, c - wl * 64); });
        __syncthreads();
      }
    };
    int base = 0;
    if (known) {
      for (int k = 0; k < h; ++k) base += pc[k];
    } else {
      for (int hh = 0; hh < h; ++hh) {
        build_part(hh, false, 0);
        const int nw = min(wpp, nw64 - hh * wpp);
        int c = 0;
        for (int w = threadIdx.x; w < nw; w += BT) c += __popcll(bm[w]);
        base += block_sum_int<BT>(c, s_red);
        __syncthreads();
      }
    }
    const int w_lo = h * wpp, nwp = min(wpp, nw64 - w_lo);
    const int c_lo = w_lo * 64;
    // How the team splits the products.  The A entries are walked in batches of BT; with at
    // least as many batches as members (hub rows: 30 000 entries = 59 batches) member r takes
    // whole batches r, r+T, ... — no member repeats another's batch preparation — otherwise
    // the members are dealt round-robin to the batches and share a batch's flat index space.
    const int nbat = (int)((a1 - a0 + BT - 1) / BT);
    int b_first, b_step, sub_r, sub_T;
    if (team_T >= nbat) {
      b_first = team_r % nbat;
      b_step = nbat;  // a single batch
      sub_r = team_r / nbat;
      sub_T = team_T / nbat + ((team_r % nbat) < (team_T % nbat) ? 1 : 0);
    } else {
      b_first = team_r;
      b_step = team_T;
      sub_r = 0;
      sub_T = 1;
    }
    // this member's first batch: segments located, B lines requested from L2 ahead of use
    const int64_t bf0 = a0 + (int64_t)b_first * BT;
    const int nb0 = (int)min((int64_t)BT, a1 - bf0);
    const long long total0 =
        walk_prepare_part<BT>(bf0, nb0, Acol, Aval, Brp, bsplit, krows, h, nparts, ws);
    walk_prefetch<BT, true>(ws, nb0, total0, Bcol, Bval, sub_r, sub_T);
    bool batch0_ready = true;
    // ---- the part's bitmap
    if (slotno >= 0) {
      const ulonglong2* src =
          reinterpret_cast<const ulonglong2*>(bm_store + (size_t)slotno * nw64 + w_lo);
      ulonglong2* dst2 = reinterpret_cast<ulonglong2*>(bm);
      for (int w = threadIdx.x; w < (nwp >> 1); w += BT) dst2[w] = ldg_hint(src + w, pol_bm);
    } else {
      build_part(h, b_first == 0, total0);
      batch0_ready = nbat == 1;  // the walk area still holds this member's first batch only then
    }
    __syncthreads();
    // ---- popcount prefix (lane = word), warp totals scanned across the CTA
    const int wpw = (((nwp + NW - 1) / NW) + 31) & ~31;
    const int wbeg = warp * wpw, wend = min(nwp, wbeg + wpw);
    {
      int run = 0;
      for (int w = wbeg + lane; w < wbeg + wpw; w += 32) {
        const int c = (w < wend) ? __popcll(bm[w]) : 0;
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int y = __shfl_up_sync(FULL, inc, o);
          if (lane >= o) inc += y;
        }
        if (w < wend) pref[w] = (unsigned)(run + inc - c);
        run += __shfl_sync(FULL, inc, 31);
      }
      __syncthreads();
      if (lane == 0) s_red[warp] = run;
      __syncthreads();
    }
    int wbase = 0, cnt = 0;
#pragma unroll
    for (int k = 0; k < NW; ++k) {
      const int x = s_red[k];
      wbase += (k < warp) ? x : 0;
      cnt += x;
    }
    for (int w = wbeg + lane; w < wend; w += 32) pref[w] += (unsigned)wbase;
    __syncthreads();
    if (cnt == 0) continue;
    double* acc = Cval + Crp[i] + base;
    int* ocol = Ccol + Crp[i] + base;
    // ---- this member's share of the accumulators zeroed and of the columns emitted
    {
      const int per = (cnt + team_T - 1) / team_T;
      const int k1 = min(cnt, (team_r + 1) * per);
      for (int k = team_r * per + threadIdx.x; k < k1; k += BT) stg_hint(acc + k, 0.0, pol_acc);
    }
    {
      const int ngroups = (nwp + 31) >> 5;
      const unsigned lt = lanemask_lt();
      for (int g = warp + NW * team_r; g < ngroups; g += NW * team_T) {
        const int w = g * 32 + lane;
        unsigned long long x = (w < nwp) ? bm[w] : 0ull;
        int pos = (w < nwp) ? (int)pref[w] : 0;
        const bool is_dense = __popcll(x) >= DENSE_WORD;
        unsigned dense = __ballot_sync(FULL, is_dense);
        while (dense) {
          const int src = __ffs((int)dense) - 1;
          dense &= dense - 1;
          const unsigned long long xs = (unsigned long long)shfl64((long long)x, src);
          const int ps = __shfl_sync(FULL, pos, src);
          const unsigned lo32 = (unsigned)xs, hi32 = (unsigned)(xs >> 32);
          const int cb = c_lo + (g * 32 + src) * 64;
          if ((lo32 >> lane) & 1u) stg_hint(ocol + ps + __popc(lo32 & lt), cb + lane, pol_ocol);
          if ((hi32 >> lane) & 1u)
            stg_hint(ocol + ps + __popc(lo32) + __popc(hi32 & lt), cb + 32 + lane, pol_ocol);
        }
        if (is_dense) x = 0ull;
        const int cb = c_lo + w * 64;
        while (x) {
          const int b = __ffsll((long long)x) - 1;
          x &= x - 1;
          stg_hint(ocol + pos++, cb + b, pol_ocol);
        }
      }
    }
    // ---- products of the part
    auto accumulate = [&](int c, double prod) {
      const int cc = c - c_lo;
      const int w = cc >> 6;
      const unsigned long long below = bm[w] & ((1ull << (cc & 63)) - 1ull);
      red_add_hint(acc + (int)pref[w] + __popcll(below), prod, pol_acc);
    };
    __syncthreads();  // acc[] zeroed by every thread of this CTA before its first RED lands
    if (team_T > 1) {   // ... and by every other member of the team
      if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(team_ready + q, 1);
        while (*(volatile int*)(team_ready + q) < team_T) __nanosleep(200);
        __threadfence();
      }
      __syncthreads();
    }
    for (int bi = b_first; bi < nbat; bi += b_step) {
      const int64_t b0 = a0 + (int64_t)bi * BT;
      const int nb = (int)min((int64_t)BT, a1 - b0);
      const long long total =
          (batch0_ready && bi == b_first)
              ? total0
              : walk_prepare_part<BT>(b0, nb, Acol, Aval, Brp, bsplit, krows, h, nparts, ws);
      // this member's share of the batch, drawn by its warps in chunks
      const long long per_member = (((total + sub_T - 1) / sub_T) + 31) & ~31LL;
      const long long glo = min(total, (long long)sub_r * per_member);
      const long long ghi = min(total, glo + per_member);
      if (threadIdx.x == 0) s_next = glo;
      __syncthreads();
      walk_run_dynamic<BT, 512, true>(ws, nb, glo, ghi, &s_next, Bcol, Bval, pol_b, accumulate);
      __syncthreads();
    }
  }
}

// symbolic for large rows, PART-WISE (same geometry as k_num_bitmap_part: 512 threads, the
// part's bitmap in shared memory, two CTAs per SM).  Item (row, part) counts the distinct
// columns of the row inside the part (partcnt), and keeps the part's bitmap in the store for
// the numeric pass when the row has a slot there.  rownnz = sum of the parts (k_sum_parts).
template <int BT, int MINB>
__global__ void __launch_bounds__(BT, MINB)
k_sym_bitmap_part(const int* __restrict__ list, int count, int nparts, int wpp, int row_lo,
                  const int64_t* __restrict__ Arp, const int* __restrict__ Acol,
                  const double* __restrict__ Aval, const int64_t* __restrict__ Brp,
                  const int* __restrict__ Bcol, int nw64,
                  unsigned long long* __restrict__ bm_store, int store_rows,
                  int* __restrict__ bm_slot, int* __restrict__ partcnt,
                  const int* __restrict__ bsplit, int krows, int* __restrict__ work_counter,
                  L2Modes l2) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_red[BT / 32];
  __shared__ int s_idx;
  __shared__ long long s_next;
  WalkSmem<BT>& ws = *reinterpret_cast<WalkSmem<BT>*>(smem_raw);
  unsigned long long* bm = (unsigned long long*)(smem_raw + sizeof(WalkSmem<BT>));
  unsigned* bm32 = (unsigned*)bm;
  const unsigned long long pol_b = l2_policy(l2.bgather), pol_bm = l2_policy(l2.bmstore);
  const int items = count * nparts;
  for (int w = threadIdx.x; w < wpp; w += BT) bm[w] = 0ull;
  while (true) {
    __syncthreads();
    if (threadIdx.x == 0) s_idx = atomicAdd(work_counter, 1);
    __syncthreads();
    const int t = s_idx;
    if (t >= items) break;
    const int idx = t / nparts, h = t % nparts;
    const int i = list[idx];
    const int w_lo = h * wpp, nwp = min(wpp, nw64 - w_lo);
    if (nwp <= 0) {
      if (threadIdx.x == 0) partcnt[(size_t)i * PARTS_MAX + h] = 0;
      continue;
    }
    const int c_lo = w_lo * 64;
    const int64_t a0 = Arp[row_lo + i], a1 = Arp[row_lo + i + 1];
    for (int64_t b0 = a0; b0 < a1; b0 += BT) {
      const int nb = (int)min((int64_t)BT, a1 - b0);
      const long long total =
          walk_prepare_part<BT>(b0, nb, Acol, Aval, Brp, bsplit, krows, h, nparts, ws);
      if (threadIdx.x == 0) s_next = 0;
      __syncthreads();
      walk_run_dynamic<BT, 512, false>(ws, nb, 0, total, &s_next, Bcol, (const double*)nullptr,
                                       pol_b, [&](int c, double) { bitmap_set(bm32, c - c_lo); });
      __syncthreads();
    }
    int cnt = 0;
    unsigned long long* dst =
        (bm_store && idx < store_rows) ? bm_store + (size_t)idx * nw64 + w_lo : nullptr;
    for (int w = threadIdx.x; w < nwp; w += BT) {
      const unsigned long long x = bm[w];
      cnt += __popcll(x);
      if (dst) stg_hint(dst + w, x, pol_bm);
      bm[w] = 0ull;
    }
    cnt = block_sum_int<BT>(cnt, s_red);
    if (threadIdx.x == 0) {
      partcnt[(size_t)i * PARTS_MAX + h] = cnt;
      if (h == 0) bm_slot[i] = dst ? idx : -1;
    }
  }
}

// rownnz[i] = sum over parts of partcnt[i][.] for the rows of a list
__global__ void __launch_bounds__(256)
k_sum_parts(const int* __restrict__ list, int count, int nparts, const int* __restrict__ partcnt,
            int* __restrict__ rownnz) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const int i = list[t];
  int s = 0;
  for (int k = 0; k < nparts; ++k) s += partcnt[(size_t)i * PARTS_MAX + k];
  rownnz[i] = s;
}

// ticket -> slot table (one load per work item instead of a binary search over itemoff)
__global__ void __launch_bounds__(256)
k_fill_tickets(const int* __restrict__ itemoff, int nslots, int* __restrict__ ticket_slot) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nslots) return;
  for (int t = itemoff[q]; t < itemoff[q + 1]; ++t) ticket_slot[t] = q;
}

// team size of every (row, part) slot of the part-wise numeric kernel: 0 for an empty part,
// else ceil(products of the row / nparts / team_products), at most team_max
__global__ void __launch_bounds__(256)
k_team_sizes(const int* __restrict__ list, int count, int nparts,
             const long long* __restrict__ flops, const int* __restrict__ partcnt,
             long long team_products, int team_max, int* __restrict__ tsize) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= count * nparts) return;
  const int i = list[q / nparts], h = q % nparts;
  const int* pc = partcnt + (size_t)i * PARTS_MAX;
  int T = 1;
  if (pc[0] >= 0) {
    if (pc[h] == 0) T = 0;
    else {
      int nonempty = 0;
      for (int k = 0; k < nparts; ++k) nonempty += pc[k] > 0;
      const long long share = flops[i] / max(1, nonempty);
      T = (int)min((long long)team_max, max(1LL, (share + team_products - 1) / team_products));
    }
  }
  tsize[q] = T;
}

// every row strictly ascending?  flag[0] is cleared by any violating pair.  With `validate`
// the CSR itself is checked too — row offsets start at 0, never decrease, end at nnz; columns lie
// in [0, cols) — and flag[1] is set by any violation (the reference trusts its inputs; here a bad
// index would otherwise become an out-of-bounds read in the flops kernel and a sticky CUDA error)
__global__ void __launch_bounds__(256)
k_check_sorted(const int64_t* __restrict__ rp, const int* __restrict__ col, int m, int cols,
               long long nnz, int validate, int* __restrict__ flag) {
  const int warp = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= m) return;
  int64_t s = rp[warp], e = rp[warp + 1];
  if (validate) {
    bool bad = s < 0 || e < s || e > nnz || (warp == 0 && s != 0) || (warp == m - 1 && e != nnz);
    if (bad) { if (lane == 0) flag[1] = 1; return; }
  }
  bool ok = true, inside = true;
  for (int64_t p = s + lane; p < e; p += 32) {
    const int c = col[p];
    if (p > s) ok &= col[p - 1] < c;
    inside &= (unsigned)c < (unsigned)cols;
  }
  if (!ok) flag[0] = 0;
  if (validate && !inside) flag[1] = 1;
}

// keys for ordering a bitmap-bin list heaviest first: key[t] = flops[list[t]]
__global__ void __launch_bounds__(256)
k_gather_keys(const int* __restrict__ list, int count, const long long* __restrict__ flops,
              long long* __restrict__ keys) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < count) keys[t] = flops[list[t]];
}

// rMCL rows whose product is empty keep nothing
__global__ void __launch_bounds__(256)
k_rmcl_empty_rows(const int* __restrict__ list, int count, RmclOut ro) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < count) { ro.row_off[list[t]] = 0; ro.row_kept[list[t]] = 0; }
}

// rMCL epilogue of the rows whose unpruned product was written into their arena slice by the
// part-wise numeric kernel (ascending columns): inflate, max / sum, threshold, prune,
// normalise, chaos term (nlibs/tools/util.cc:4-69), in place — the kept entries are compacted
// to the front of the slice.  One 256-thread CTA per row, persistent with dynamic fetch; sums
// are fixed-shape (thread-strided partials, warp xor tree, ordered sum over warps), so the
// result does not depend on scheduling.
template <int BT>
__global__ void __launch_bounds__(BT)
k_rmcl_epilogue_rows(const int* __restrict__ list, int count, const int64_t* __restrict__ Crp,
                     RmclOut ro, int* __restrict__ work_counter) {
  __shared__ int s_red[BT / 32];
  __shared__ double s_redd[BT / 32];
  __shared__ int s_idx;
  while (true) {
    __syncthreads();
    if (threadIdx.x == 0) s_idx = atomicAdd(work_counter, 1);
    __syncthreads();
    const int idx = s_idx;
    if (idx >= count) break;
    const int i = list[idx];
    const long long off = (long long)Crp[i];
    const int cnt = (int)(Crp[i + 1] - off);
    int* col = ro.arena_col + off;
    double* acc = ro.arena_val + off;
    double psum = 0.0, pmax = 0.0;
    for (int k = threadIdx.x; k < cnt; k += BT) {
      const double v = __ldcg(acc + k);  // written by RED at L2 in the previous kernel
      const double v2 = __dmul_rn(v, v);
      acc[k] = v2;
      psum = __dadd_rn(psum, v2);
      pmax = fmax(pmax, v2);
    }
    const double rsum = block_sum_d<BT>(psum, s_redd);
    const double rmax = block_max_d<BT>(pmax, s_redd);
    const double thresh = compute_threshold(__ddiv_rn(rsum, (double)cnt), rmax);
    // opt-in top-k (see topk_cut): block-wide counts
    double tk = 0.0;
    int cstar = 0x7fffffff;
    bool cut = false;
    if (ro.topk > 0) {
      auto bcount = [&](auto pred) {
        int c = 0;
        for (int k = threadIdx.x; k < cnt; k += BT) c += (acc[k] >= thresh && pred(acc[k], col[k])) ? 1 : 0;
        return block_sum_int<BT>(c, s_red);
      };
      if (bcount([](double, int) { return true; }) > ro.topk) {
        cut = true;
        topk_cut(bcount, thresh, rmax, ro.topk, 0x7fffffff, &tk, &cstar);
      }
    }
    auto keeps = [&](double v, int c) { return v >= thresh && (!cut || v > tk || (v == tk && c <= cstar)); };
    double ksum_p = 0.0;
    int kept_p = 0;
    for (int k = threadIdx.x; k < cnt; k += BT) {
      const double v2 = acc[k];
      if (keeps(v2, col[k])) { ksum_p = __dadd_rn(ksum_p, v2); ++kept_p; }
    }
    const double ksum = block_sum_d<BT>(ksum_p, s_redd);
    const int kept = block_sum_int<BT>(kept_p, s_red);
    double sq_p = 0.0;
    int written = 0;
    for (int k0 = 0; k0 < cnt; k0 += BT) {
      const int k = k0 + threadIdx.x;
      const double v2 = (k < cnt) ? acc[k] : 0.0;
      const int c = (k < cnt) ? col[k] : 0;
      const bool keep = (k < cnt) && keeps(v2, c);
      int tot;
      // (the scan's barriers separate this chunk's reads from the writes below, which land at
      // or before the positions just read: compaction only moves entries to the left)
      const int ex = block_excl_scan<BT>(keep ? 1 : 0, s_red, &tot);
      if (keep) {
        const double w = __ddiv_rn(v2, ksum);
        col[written + ex] = c;
        acc[written + ex] = w;
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

#include "ranges.cuh"
#include "esc.cuh"

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
  B200_CUDA(d2h_small(out->cnt, d_hist, 16 * sizeof(int), c.stream));
  B200_CUDA(sync_fetch(c.stream));
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

// ---- small read-backs through mapped pinned memory (common.cuh) -------------------------------
__global__ void k_readback(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst, int bytes) {
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) dst[i] = src[i];
}

void load_tunables(Tunables* t) {
  *t = Tunables();
  auto flag = [](const char* k) { const char* e = getenv(k); return e != nullptr && *e != 0; };
  t->force_wide = flag("B200_FORCE_WIDE");
  t->parts4 = flag("B200_PARTS4");
  t->no_parts = flag("B200_NO_PARTS");
  t->on_chip = flag("B200_ON_CHIP");
  t->prof = flag("B200_PROF");
  if (const char* e = getenv("B200_L2POL"))
    sscanf(e, "%d,%d,%d,%d,%d", &t->l2[0], &t->l2[1], &t->l2[2], &t->l2[3], &t->l2[4]);
  if (const char* e = getenv("B200_SYM_BIG_FROM")) t->sym_big_from = atoll(e);
  if (const char* e = getenv("B200_NUM_BIG_FROM")) t->num_big_from = atoi(e);
  if (const char* e = getenv("B200_LIGHT_P")) t->light_p = atoll(e);
  if (const char* e = getenv("B200_TEAM_P")) t->team_products = std::max(1LL, atoll(e));
  if (const char* e = getenv("B200_TEAM_MAX")) t->team_max = std::max(1, atoi(e));
  if (const char* e = getenv("B200_RANGES")) t->ranges = atoi(e);
  if (const char* e = getenv("B200_ARENA_ENTRIES")) t->arena_entries = atoll(e);
  if (const char* e = getenv("B200_ROW_CHARGE")) t->row_charge = std::max(0LL, atoll(e));
  t->deterministic = flag("B200_DETERMINISTIC");
  if (const char* e = getenv("B200_ESC")) t->esc = atoi(e) ? 1 : 0;
  if (const char* e = getenv("B200_FUSE")) t->fuse = atoi(e);   // 0 off, 1 on, 2 on with 64-bit B offsets (testing)
  if (t->deterministic) t->on_chip = true;
}

void rb_reset() {
  Ctx& c = ctx();
  c.rb_n = 0;
  c.rb_used = 0;
}

cudaError_t d2h_small(void* dst, const void* src, size_t bytes, cudaStream_t st) {
  Ctx& c = ctx();
  if (!c.rb_host) {
    void* h = nullptr;
    void* d = nullptr;
    if (cudaHostAlloc(&h, Ctx::RB_BYTES, cudaHostAllocMapped) == cudaSuccess &&
        cudaHostGetDevicePointer(&d, h, 0) == cudaSuccess) {
      c.rb_host = (unsigned char*)h;
      c.rb_dev = (unsigned char*)d;
    } else {
      cudaGetLastError();
      if (h) cudaFreeHost(h);
    }
  }
  const size_t need = (bytes + 15) & ~(size_t)15;
  if (!c.rb_host || c.rb_n == Ctx::RB_MAX || c.rb_used + need > (size_t)Ctx::RB_BYTES)
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st);   // still correct, may wait
  k_readback<<<1, 128, 0, st>>>((const unsigned char*)src, c.rb_dev + c.rb_used, (int)bytes);
  c.rb_pending[c.rb_n].dst = dst;
  c.rb_pending[c.rb_n].off = c.rb_used;
  c.rb_pending[c.rb_n].bytes = bytes;
  ++c.rb_n;
  c.rb_used += need;
  return cudaGetLastError();
}

cudaError_t sync_fetch(cudaStream_t st) {
  Ctx& c = ctx();
  const cudaError_t e = cudaStreamSynchronize(st);
  if (e == cudaSuccess)
    for (int i = 0; i < c.rb_n; ++i) memcpy(c.rb_pending[i].dst, c.rb_host + c.rb_pending[i].off, c.rb_pending[i].bytes);
  c.rb_n = 0;
  c.rb_used = 0;
  return e;
}

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
  d->sorted_rows = true;
  B200_CUDA(sync_fetch(c.stream));
  return B200_OK;
}

// a column-sorted copy of (col, val) of `d` (row offsets shared); caller frees col2 / val2
static int sorted_copy_device(const DevCSR& d, int** col2, double** val2) {
  Ctx& c = ctx();
  B200_CUDA(dalloc(col2, (size_t)d.nnz));
  B200_CUDA(dalloc(val2, (size_t)d.nnz));
  void* tmp = nullptr;
  size_t tb = 0;
  cub::DeviceSegmentedSort::SortPairs(nullptr, tb, d.col, *col2, d.val, *val2, d.nnz, d.rows,
                                      d.rowptr, d.rowptr + 1, c.stream);
  B200_CUDA(cudaMallocAsync(&tmp, tb ? tb : 1, c.stream));
  cub::DeviceSegmentedSort::SortPairs(tmp, tb, d.col, *col2, d.val, *val2, d.nnz, d.rows,
                                      d.rowptr, d.rowptr + 1, c.stream);
  B200_CUDA(cudaGetLastError());
  cudaFreeAsync(tmp, c.stream);
  return B200_OK;
}

int check_sorted_device(DevCSR* d, bool validate) {
  Ctx& c = ctx();
  d->sorted_rows = true;
  if (d->rows == 0) {
    if (validate && d->nnz != 0) { set_error("bad CSR: entries without rows"); return B200_ERR_BAD_ARG; }
    return B200_OK;
  }
  if (d->nnz == 0 && !validate) return B200_OK;
  Temps T;
  int* d_flag = nullptr;
  int h[2] = {1, 0};
  rb_reset();
  B200_CUDA(T.alloc(&d_flag, 2));
  B200_CUDA(cudaMemcpyAsync(d_flag, h, sizeof h, cudaMemcpyHostToDevice, c.stream));
  const long long threads = (long long)d->rows * 32;
  k_check_sorted<<<(unsigned)((threads + 255) / 256), 256, 0, c.stream>>>(d->rowptr, d->col, d->rows, d->cols,
                                                                         d->nnz, validate ? 1 : 0, d_flag);
  B200_CUDA(d2h_small(h, d_flag, sizeof h, c.stream));
  B200_CUDA(sync_fetch(c.stream));
  if (h[1]) {
    set_error("bad CSR: row offsets must start at 0, never decrease and end at nnz; columns must lie in [0, cols)");
    return B200_ERR_BAD_ARG;
  }
  d->sorted_rows = h[0] != 0;
  return B200_OK;
}

// ------------------------------------------------------------------------------------------
// cost of a row for the partition of a step among GPUs: its products, plus a fixed charge for
// every row heavy enough for the CTA-per-row kernels (their per-row set-up does not shrink with
// the row: measured on R-MAT scale 20, equal products alone leave the rank that holds the long
// tail 1.7x slower than the one that holds the hubs).  The reference's static variant balances a
// footprint of the same kind, (products + nnz(C_i) + 32 + nnz(A_i)) >> 1
// (nlibs/static_omp_csr_kernel.cc:28-62).
__global__ void __launch_bounds__(256)
k_row_cost(long long* __restrict__ flops, int m, long long row_charge, long long heavy_from) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) { const long long f = flops[i]; flops[i] = f + (f > heavy_from ? row_charge : 0); }
}

int flops_prefix_device(const DevCSR& A, const DevCSR& B, int row_lo, int row_hi,
                        int64_t* d_prefix, long long row_charge, long long* d_products) {
  Ctx& c = ctx();
  const int m = row_hi - row_lo;
  Temps T;
  long long* d_flops = nullptr;
  unsigned char* d_bin = nullptr;
  int* d_cnt = nullptr;
  B200_CUDA(T.alloc(&d_flops, (size_t)m + 1));
  B200_CUDA(T.alloc(&d_bin, (size_t)m));
  B200_CUDA(T.alloc(&d_cnt, (size_t)m));
  B200_CUDA(cudaMemsetAsync(d_flops + m, 0, sizeof(long long), c.stream));
  if (m > 0)
    k_row_flops<<<(unsigned)(((long long)m * 8 + 255) / 256), 256, 0, c.stream>>>(
        A.rowptr, A.col, B.rowptr, row_lo, m, 8192LL, d_flops, d_bin, d_cnt);
  void* tmp = nullptr;
  size_t tb = 0;
  if (d_products) {   // the plain product count, before any charge
    cub::DeviceReduce::Sum(nullptr, tb, d_flops, d_products, m, c.stream);
    B200_CUDA(T.alloc((char**)&tmp, tb ? tb : 1));
    cub::DeviceReduce::Sum(tmp, tb, d_flops, d_products, m, c.stream);
  }
  if (row_charge > 0 && m > 0) k_row_cost<<<(m + 255) / 256, 256, 0, c.stream>>>(d_flops, m, row_charge, 512);
  tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, d_flops, (long long*)d_prefix, m + 1, c.stream);
  void* tmp2 = nullptr;
  B200_CUDA(T.alloc((char**)&tmp2, tb ? tb : 1));
  cub::DeviceScan::ExclusiveSum(tmp2, tb, d_flops, (long long*)d_prefix, m + 1, c.stream);
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

// cudaMalloc for the context's long-lived buffers; on failure the stream-ordered pool (which
// keeps every freed block, e.g. a 100 GB result of an earlier call) is trimmed and the
// allocation retried once
static cudaError_t malloc_with_trim(void** p, size_t bytes) {
  cudaError_t e = cudaMalloc(p, bytes);
  if (e == cudaSuccess) return e;
  cudaGetLastError();
  Ctx& c = ctx();
  cudaStreamSynchronize(c.stream);
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, c.device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
  e = cudaMalloc(p, bytes);
  if (e != cudaSuccess) cudaGetLastError();
  return e;
}

// ------------------------------------------------------------------------------------------
int sorted_copy_of(const DevCSR& d, DevCSR* out) {
  *out = d;
  int* col2 = nullptr;
  double* val2 = nullptr;
  const int rc = sorted_copy_device(d, &col2, &val2);
  if (rc) { dfree(col2); dfree(val2); return rc; }
  out->col = col2;
  out->val = val2;
  out->sorted_rows = true;
  return B200_OK;
}

int run_pipeline(const DevCSR& A, const DevCSR& B, int row_lo, int row_hi, Mode mode, DevCSR* C,
                 double* chaos, b200_stats* stats, const DevCSR* Bsorted) {
  Ctx& c = ctx();
  cudaStream_t st = c.stream;
  const int m = row_hi - row_lo;
  const int n = B.cols;
  int launches = 0, range_items = 0;
  rb_reset();
  if (stats) memset(stats, 0, sizeof(*stats));
  *C = DevCSR();
  // on any early return the result is left empty (callers release it) and every temporary is
  // freed: `out` is destroyed after `T`
  struct OutGuard { DevCSR* C; bool ok; ~OutGuard() { if (!ok) *C = DevCSR(); } } out = {C, false};
  Temps T;
  C->rows = m;
  C->cols = n;
  B200_CUDA(cudaEventRecord(c.ev[0], st));
  // per-kernel brackets (only when the caller asked for stats)
  bool sym_timed[16] = {false}, num_timed[16] = {false};
  auto tick = [&](int slot) { if (stats) cudaEventRecord(c.kev[slot], st); };

  // large rows use a column bitmap of nw64 words of 64 columns; where it lives decides the
  // bin cut points (see sym_bin_of)
  const int nw64 = (((n + 63) / 64) + 1) & ~1;  // even: rows of the bitmap store stay 16-byte aligned
  const size_t bm_bytes = (size_t)nw64 * 8;
  const size_t bm_pref_bytes = bm_bytes + (((size_t)nw64 + 1) / 2) * 8;
  const size_t walk_bytes = sizeof(WalkSmem<BT_BIG>);
  const size_t smem_cap = c.smem_optin - 1024 - walk_bytes;  // static shared + the walk area
  // B200_FORCE_WIDE=1 (testing switch): behave as if B had too many columns for any
  // shared-memory bitmap, so small inputs exercise the big warp tables and the HBM bitmap
  const bool force_wide = c.tun.force_wide;
  const bool sym_smem = !force_wide && bm_bytes <= smem_cap;
  const bool num_smem = !force_wide && bm_pref_bytes <= smem_cap;
  // L2 eviction priorities of the bitmap kernels (B200_L2POL=acc,ocol,bgather,bmstore,demote
  // overrides them: a developer switch for A/B runs)
  const L2Modes l2m = {c.tun.l2[0], c.tun.l2[1], c.tun.l2[2], c.tun.l2[3], c.tun.l2[4]};
  // Part-wise numeric kernel (SpGEMM, sorted B rows): column parts of <= 8192 bitmap words so
  // that two 512-thread CTAs fit one SM.
  const bool parts4 = c.tun.parts4;  // developer switch: 4 x 256 threads / SM
  const int part_words_max = parts4 ? 4096 : 8192;
  const int part_ctas = parts4 ? 4 : 2;
  const size_t walk_part_bytes = parts4 ? sizeof(WalkSmem<256>) : sizeof(WalkSmem<512>);
  int nparts = (nw64 + part_words_max - 1) / part_words_max;
  int wpp = (((nw64 + nparts - 1) / nparts) + 1023) & ~1023;  // a multiple of the sweep step of k_sym_bitmap
  const size_t part_smem = walk_part_bytes + (size_t)wpp * 12;
  // (rMCL mode too: the part kernel then writes the unpruned row into its arena slice and
  // k_rmcl_epilogue_rows prunes it in place.  Unsorted B rows — an rMCL iterate in first-touch
  // order — need a column-sorted copy of B when there is more than one part.)
  const bool use_parts = nparts <= PARTS_MAX &&
                         part_ctas * (part_smem + 2048) <= c.smem_optin + 1024 &&
                         !c.tun.no_parts && !force_wide;
  if (!use_parts) { nparts = 1; wpp = nw64; }
  // symbolic cut: rows of up to 1024 products are cheaper in the (optimistically sized) warp
  // tables at 56 warps / SM than as one bitmap item each (a 27-point stencil row has 729)
  // (with more than two column parts a bitmap row costs one item per part it touches, and the
  // cut moves up: measured on a 4 M-column planted-partition graph, 8 parts: 785 -> 243 ms)
  long long sym_big_from = (sym_smem || use_parts) ? (nparts > 2 ? 4096 : 1024) : 8192;
  if (c.tun.sym_big_from >= 0) sym_big_from = c.tun.sym_big_from;  // developer switch
  int num_big_from = (num_smem || use_parts) ? 256 : 2048;
  if (c.tun.num_big_from >= 0) num_big_from = c.tun.num_big_from;  // developer switch

  // On-chip numeric pass of the heavy rows (ranges.cuh): needs the whole-row symbolic kernel
  // (it counts the output columns per static range) and the part kernel as its fallback.
  const int RNG = c.tun.ranges;
  const bool want_ranges = use_parts && sym_smem && RNG >= 2 && RNG <= RANGES_MAX && c.tun.on_chip;
  // four 4-warp CTAs per SM, every warp an independent worker with its own pool
  constexpr int UNIT_CTAS = 4;
  const int unit_pool = (int)((((c.smem_optin + 1024) / UNIT_CTAS - 1024) / UNIT_WARPS) & ~(size_t)15);
  const size_t unit_dyn = (size_t)unit_pool * UNIT_WARPS;
  const int item_pool = unit_pool - 16;

  // mid-size rows of an rMCL step are sorted on chip (esc.cuh), without a symbolic pass
  // — in very wide matrices, where a bitmap row costs one work item per 512 K-column part and the
  // large warp tables run at 4 warps per SM (measured on the 4 M-vertex planted-partition graph:
  // light iterate 873 -> 265 ms, heavy iterate 3136 -> 1342 ms); below ~1 M columns the bitmap
  // path is the faster one (R-MAT scale 18 loop: 215 ms against 237), so it stays.  B200_ESC=0 / 1
  // forces the choice.
  const bool use_esc = c.tun.esc < 0 ? nparts > 2 : c.tun.esc > 0;
  const long long esc_max = use_esc ? 8192 : 0;

  // arena of unpruned rows (rMCL) / of the rows finished on chip before C exists (SpGEMM, esc.cuh)
  auto ensure_arena = [&](size_t need_entries) -> int {
    const size_t unpruned = need_entries;
    if (c.arena_cap < (size_t)unpruned) {
      B200_CUDA(sync_fetch(st));
      // head-room: an rMCL loop alternates between a few sizes, and re-allocating tens of GB
      // costs ~1 s; grow by at least 2x the old capacity (falls back to the exact size below)
      const size_t old_cap = c.arena_cap;
      const size_t want = std::max((size_t)unpruned + (size_t)unpruned / 8 + 1, 2 * old_cap);
      if (c.arena_col) { cudaFree(c.arena_col); cudaFree(c.arena_val); }
      c.arena_col = nullptr; c.arena_val = nullptr; c.arena_cap = 0;
      size_t got = want;
      if (cudaMalloc((void**)&c.arena_col, got * sizeof(int)) != cudaSuccess ||
          cudaMalloc((void**)&c.arena_val, got * sizeof(double)) != cudaSuccess) {
        cudaGetLastError();
        if (c.arena_col) { cudaFree(c.arena_col); c.arena_col = nullptr; }
        got = (size_t)unpruned + 1;  // without the head-room, and with the pool trimmed if needed
        if (malloc_with_trim((void**)&c.arena_col, got * sizeof(int)) != cudaSuccess ||
            malloc_with_trim((void**)&c.arena_val, got * sizeof(double)) != cudaSuccess) {
          if (c.arena_col) { cudaFree(c.arena_col); c.arena_col = nullptr; }
          c.arena_val = nullptr;
          set_error("out of device memory for the rMCL arena (unpruned product of this row block); "
                    "shard the rows over more GPUs or call the *_rows entry points on smaller blocks");
          return B200_ERR_CUDA;
        }
      }
      c.arena_cap = got;
      if (c.tun.prof)
        fprintf(stderr, "[b200 prof] arena: %zu -> %zu entries (asked %zu)\n", old_cap, got, (size_t)unpruned);
    }
    return B200_OK;
  };

  // ---- 1. flops analysis + symbolic binning
  long long* d_flops = nullptr;
  unsigned char* d_bin = nullptr;
  int* d_cnt = nullptr;
  long long* d_P = nullptr;
  B200_CUDA(T.alloc(&d_flops, (size_t)m));
  B200_CUDA(T.alloc(&d_bin, (size_t)m));
  B200_CUDA(T.alloc(&d_cnt, (size_t)m + 1));
  B200_CUDA(T.alloc(&d_P, 2));
  if (m > 0) {
    k_row_flops<<<(unsigned)(((long long)m * 8 + 255) / 256), 256, 0, st>>>(
        A.rowptr, A.col, B.rowptr, row_lo, m, sym_big_from, d_flops, d_bin, d_cnt, esc_max);
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
  T.adopt(sb.d_list);
  if (rc) return rc;
  // heaviest rows first inside a bitmap bin: the persistent CTAs fetch rows dynamically, and a
  // hub row picked up last would be the tail of the kernel
  auto order_heavy_first = [&](int* list, int count) -> int {
    if (count < 2) return B200_OK;
    long long *k0 = nullptr, *k1 = nullptr;
    int* l1 = nullptr;
    B200_CUDA(T.alloc(&k0, (size_t)count));
    B200_CUDA(T.alloc(&k1, (size_t)count));
    B200_CUDA(T.alloc(&l1, (size_t)count));
    k_gather_keys<<<(count + 255) / 256, 256, 0, st>>>(list, count, d_flops, k0);
    void* tmp = nullptr;
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, tb, k0, k1, list, l1, count, 0, 64, st);
    B200_CUDA(cudaMallocAsync(&tmp, tb ? tb : 1, st));
    cub::DeviceRadixSort::SortPairsDescending(tmp, tb, k0, k1, list, l1, count, 0, 64, st);
    B200_CUDA(cudaMemcpyAsync(list, l1, (size_t)count * sizeof(int), cudaMemcpyDeviceToDevice, st));
    cudaFreeAsync(tmp, st);

    launches += 2;
    return B200_OK;
  };
  if ((rc = order_heavy_first(sb.d_list + sb.off[SB_BITMAP], sb.cnt[SB_BITMAP]))) return rc;
  B200_CUDA(cudaEventRecord(c.ev[1], st));

  // ---- 2. symbolic per bin
  // warp-table bins; `Hopt`: optimistic table size tried first (0 = none), rows that overflow it
  // are collected on the device and retried in the full-size table without a host round trip
  int* d_over = nullptr;  // [0] overflow count per retry pass, then the lists
  // `src_list` / `src_count`: the rows left over by an earlier pass (list and length on the device)
  // instead of the whole bin
  auto launch_sym_warp = [&](int bin, auto kernel_full, int H, int WPB, auto kernel_opt, int Hopt,
                             int slot, const int* src_list = nullptr, const int* src_count = nullptr) -> int {
    const int cntb = sb.cnt[bin];
    if (!cntb) return B200_OK;
    const int* lst = src_list ? src_list : sb.d_list + sb.off[bin];
    if (!src_list) tick(2 * bin);
    sym_timed[bin] = true;
    if (Hopt) {
      const int WO = 8;
      const size_t so = (size_t)WO * Hopt * sizeof(int);
      int r = set_smem(kernel_opt, so);
      if (r) return r;
      int* ocount = d_over + slot;
      int* olist = d_over + 8 + (size_t)slot * m;
      // (passes that work on a device-side list: a grid that covers the machine, rows by stride)
      const int retry_grid = c.sm_count * 8;
      kernel_opt<<<src_list ? std::min((cntb + WO - 1) / WO, retry_grid) : (cntb + WO - 1) / WO, WO * 32, so, st>>>(
                                                          lst, cntb, src_count, row_lo, A.rowptr, A.col,
                                                          B.rowptr, B.col, d_cnt, Hopt * 3 / 4 - 32,
                                                          olist, ocount);
      const size_t sf = (size_t)WPB * H * sizeof(int);
      if ((r = set_smem(kernel_full, sf))) return r;
      kernel_full<<<std::min((cntb + WPB - 1) / WPB, retry_grid), WPB * 32, sf, st>>>(olist, cntb, ocount, row_lo, A.rowptr,
                                                              A.col, B.rowptr, B.col, d_cnt, H,
                                                              olist, ocount);
      launches += 2;
    } else {
      const size_t smem = (size_t)WPB * H * sizeof(int);
      int r = set_smem(kernel_full, smem);
      if (r) return r;
      kernel_full<<<src_list ? std::min((cntb + WPB - 1) / WPB, c.sm_count * 8) : (cntb + WPB - 1) / WPB, WPB * 32, smem, st>>>(lst, cntb, src_count, row_lo, A.rowptr,
                                                                A.col, B.rowptr, B.col, d_cnt, H,
                                                                nullptr, nullptr);
      ++launches;
    }
    tick(2 * bin + 1);
    return B200_OK;
  };
  if (sb.cnt[SB_W4K] || sb.cnt[SB_W16K]) {
    B200_CUDA(T.alloc(&d_over, (size_t)8 + 3 * (size_t)std::max(m, 1)));
    B200_CUDA(cudaMemsetAsync(d_over, 0, 8 * sizeof(int), st));
  }

  // rows finished in the symbolic phase (plain SpGEMM): the finished row waits in the arena, its
  // length goes into the row counts like any symbolic result.  Two kinds:
  //  - rows sorted on chip (esc.cuh): arena slice = the row's products;
  //  - SB_W4K rows tried as 128-entry numeric rows (k_num_warp_fused): arena slot = list position x 128,
  //    after the slices of the first kind.
  int64_t* d_escoff = nullptr;
  int* d_escwork = nullptr;
  unsigned char* d_fused = nullptr;
  int esc_col_bits = 1;
  while (esc_col_bits < 32 && (1ll << esc_col_bits) < (long long)n) ++esc_col_bits;
#define LAUNCH_ESC(RM, LIST, COUNT, BTE, IPTE, SLOT, OFFS, ROUT, NNZ, TIMER)                    \
  if (COUNT) {                                                                                  \
    const size_t esm = sizeof(EscSmem<BTE, IPTE>);                                              \
    if ((rc = set_smem(k_esc_rmcl<BTE, IPTE, RM>, esm))) return rc;                             \
    const int per_sm = std::max(1, (int)std::min<size_t>(2048 / BTE, (c.smem_optin + 1024) / (esm + 1024))); \
    const int egrid = std::min((int)(COUNT), per_sm * c.sm_count);                              \
    tick(TIMER);                                                                                \
    k_esc_rmcl<BTE, IPTE, RM><<<egrid, BTE, esm, st>>>(LIST, COUNT, row_lo, A.rowptr, A.col, A.val, \
        B.rowptr, B.col, B.val, OFFS, esc_col_bits, ROUT, d_esc, d_escwork + SLOT, NNZ);         \
    tick(TIMER + 1);                                                                            \
    ++launches;                                                                                 \
  }
  unsigned long long* d_esc = nullptr;  // [0] true unpruned entries of the rows sorted on chip
  const bool esc_rows = sb.cnt[SB_ESC2K] || sb.cnt[SB_ESC8K];
  if (esc_rows) {
    B200_CUDA(T.alloc(&d_esc, 2));
    B200_CUDA(cudaMemsetAsync(d_esc, 0, 2 * sizeof(unsigned long long), st));
    B200_CUDA(T.alloc(&d_escwork, 2));
    B200_CUDA(cudaMemsetAsync(d_escwork, 0, 2 * sizeof(int), st));
  }
  bool fuse = mode == MODE_SPGEMM && c.tun.fuse != 0 && sb.cnt[SB_W4K] > 0;
  if (fuse) {  // the slots must fit next to C itself: otherwise the ordinary two-pass path
    const size_t slots = (size_t)sb.cnt[SB_W4K] * FUSED_CAP;
    if (c.arena_cap < slots) {   // (an arena that already holds them costs nothing)
      size_t free_b = 0, total_b = 0;
      cudaMemGetInfo(&free_b, &total_b);
      if (slots * 12 * 3 > free_b) fuse = false;     // slots + the rows again in C + head-room
    }
  }
  if (mode == MODE_SPGEMM && (esc_rows || fuse)) {
    long long h_need = 0;
    B200_CUDA(T.alloc(&d_escoff, (size_t)m + 1));
    if (esc_rows) {
      long long* d_len = nullptr;
      B200_CUDA(T.alloc(&d_len, (size_t)m + 1));
      k_esc_len<<<(m + 256) / 256, 256, 0, st>>>(d_bin, d_flops, m, d_len);
      void* tmp = nullptr;
      size_t tb = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, tb, d_len, (long long*)d_escoff, m + 1, st);
      B200_CUDA(T.alloc((char**)&tmp, tb ? tb : 1));
      cub::DeviceScan::ExclusiveSum(tmp, tb, d_len, (long long*)d_escoff, m + 1, st);
      B200_CUDA(d2h_small(&h_need, d_escoff + m, sizeof(long long), st));
      B200_CUDA(sync_fetch(st));
      launches += 2;
    }
    const size_t fused_slots = fuse ? (size_t)sb.cnt[SB_W4K] * FUSED_CAP : 0;
    if ((rc = ensure_arena((size_t)h_need + fused_slots))) return rc;
    RmclOut eo = {};
    eo.arena_col = c.arena_col;
    eo.arena_val = c.arena_val;
    sym_timed[SB_ESC2K] = sb.cnt[SB_ESC2K] > 0;
    sym_timed[SB_ESC8K] = sb.cnt[SB_ESC8K] > 0;
    LAUNCH_ESC(false, sb.d_list + sb.off[SB_ESC2K], sb.cnt[SB_ESC2K], 256, 8, 0, d_escoff, eo, d_cnt, 2 * SB_ESC2K)
    LAUNCH_ESC(false, sb.d_list + sb.off[SB_ESC8K], sb.cnt[SB_ESC8K], 512, 16, 1, d_escoff, eo, d_cnt, 2 * SB_ESC8K)
    if (fuse) {
      const int cntb = sb.cnt[SB_W4K];
      B200_CUDA(T.alloc(&d_fused, (size_t)m));
      B200_CUDA(cudaMemsetAsync(d_fused, 0, (size_t)m, st));
      const size_t smem = (size_t)8 * 28 * FUSED_CAP;
      const bool idx32 = B.nnz < (1LL << 31) && c.tun.fuse != 2;
      if ((rc = idx32 ? set_smem(k_num_warp_fused<FUSED_CAP, true>, smem)
                      : set_smem(k_num_warp_fused<FUSED_CAP, false>, smem))) return rc;
      int* fcount = d_over + 2;
      int* flist = d_over + 8 + 2 * (size_t)m;
      tick(2 * SB_W4K);
      (idx32 ? k_num_warp_fused<FUSED_CAP, true> : k_num_warp_fused<FUSED_CAP, false>)<<<(cntb + 7) / 8, 256, smem, st>>>(
          sb.d_list + sb.off[SB_W4K], cntb, row_lo, A.rowptr, A.col, A.val, B.rowptr, B.col, B.val,
          h_need, c.arena_col, c.arena_val, d_escoff, d_cnt, d_fused, flist, fcount);
      ++launches;
      // the rows with more than 128 columns: ordinary symbolic pass (and numeric pass later)
      if ((rc = launch_sym_warp(SB_W4K, k_sym_warp<4096>, 4096, 8, k_sym_warp<1024>, 1024, 0, flist, fcount)))
        return rc;
    }
  }
  if ((rc = launch_sym_warp(SB_W256, k_sym_warp<256>, 256, 8, k_sym_warp<256>, 0, 0))) return rc;
  if ((rc = launch_sym_warp(SB_W1K, k_sym_warp<1024>, 1024, 8, k_sym_warp<1024>, 0, 0))) return rc;
  if (!fuse && (rc = launch_sym_warp(SB_W4K, k_sym_warp<4096>, 4096, 8, k_sym_warp<1024>, 1024, 0))) return rc;
  if ((rc = launch_sym_warp(SB_W16K, k_sym_warp<16384>, 16384, 3, k_sym_warp<4096>, 4096, 1))) return rc;

  // large rows: bitmap
  unsigned long long* d_bmstore = nullptr;  // stored bitmaps of the symbolic bitmap bin
  unsigned long long* d_gscr = nullptr;     // per-CTA bitmap(+prefix) scratch when not in smem
  int* d_bmslot = nullptr;                  // [m] slot in d_bmstore, or -1
  int* d_work = nullptr;
  B200_CUDA(T.alloc(&d_work, 4));
  B200_CUDA(cudaMemsetAsync(d_work, 0, 4 * sizeof(int), st));
  B200_CUDA(T.alloc(&d_bmslot, (size_t)m));
  B200_CUDA(cudaMemsetAsync(d_bmslot, 0xff, (size_t)std::max(m, 1) * sizeof(int), st));
  int* d_rw = nullptr;            // range boundaries in bitmap words [RNG + 1] (ranges.cuh)
  unsigned char* d_wr = nullptr;  // bitmap word -> range
  unsigned* d_split = nullptr;    // [B.rows][RNG] offsets of the range boundaries inside a B row
  int* d_rcnt = nullptr;          // [store_rows][RNG] output columns per range
  int* d_bsplit = nullptr;   // column-part boundaries inside every B row (k_bsplit)
  int* d_itemoff = nullptr;  // ticket offsets of the (row, part) slots of k_num_bitmap_part
  int* d_partcnt = nullptr;  // [m][PARTS_MAX] columns of the row per column part
  if (use_parts) {
    B200_CUDA(T.alloc(&d_partcnt, (size_t)m * PARTS_MAX));
    B200_CUDA(cudaMemsetAsync(d_partcnt, 0xff, (size_t)std::max(m, 1) * PARTS_MAX * sizeof(int), st));
  }
  const int nbig = sb.cnt[SB_BITMAP];
  int big_grid = std::min(nbig, c.sm_count);
  int store_rows = 0;
  // the bitmap kernels read B through `Bs`: B itself, or a column-sorted copy of it when its
  // rows are unsorted and have to be cut at column-part boundaries
  DevCSR Bs = B;
  int* d_bs_col = nullptr;
  double* d_bs_val = nullptr;
  if (((use_parts && nparts > 1) || (want_ranges && nbig > 0)) && !B.sorted_rows && B.nnz > 0 && Bsorted) {
    Bs = *Bsorted;
  } else if (((use_parts && nparts > 1) || (want_ranges && nbig > 0)) && !B.sorted_rows && B.nnz > 0) {
    rc = sorted_copy_device(B, &d_bs_col, &d_bs_val);
    T.adopt(d_bs_col);
    T.adopt(d_bs_val);
    if (rc) return rc;
    Bs.col = d_bs_col;
    Bs.val = d_bs_val;
    Bs.sorted_rows = true;
    launches += 2;
  }
  if (nbig) {
    // keep the bitmaps of the heaviest rows while they use a modest share of the memory (C
    // itself comes later); the rest are rebuilt by the numeric kernel.  The store lives in the
    // context and is re-used by later calls.
    const size_t want_words = (size_t)nbig * nw64;
    // (cudaMemGetInfo costs ~20 ms of host time when the stream-ordered pool holds > 100 GB:
    // once the store has been sized against the budget it is not re-examined)
    if (c.bm_store_words < want_words && !c.bm_store_capped) {
      size_t free_b = 0, total_b = 0;
      cudaMemGetInfo(&free_b, &total_b);
      const size_t budget_words = (size_t)(0.15 * (double)(free_b + c.bm_store_words * 8)) / 8;
      const size_t new_words = std::min(want_words, budget_words);
      // grow only for a real gain: the budget moves a little from call to call with what else
      // is allocated, and re-allocating tens of GB costs ~15 ms
      if (new_words > c.bm_store_words + c.bm_store_words / 4) {
        B200_CUDA(sync_fetch(st));
        if (c.bm_store) cudaFree(c.bm_store);
        c.bm_store = nullptr;
        c.bm_store_words = 0;
        if (malloc_with_trim((void**)&c.bm_store, new_words * 8) == cudaSuccess) c.bm_store_words = new_words;
        // else no store: every bitmap is rebuilt
      }
      c.bm_store_capped = c.bm_store_words < want_words;
    }
    store_rows = (int)std::min<size_t>((size_t)nbig, c.bm_store_words / (size_t)nw64);
    d_bmstore = store_rows ? c.bm_store : nullptr;
    if (want_ranges && store_rows > 0 && Bs.nnz > 0) {
      // static column ranges of equal column mass, their boundaries inside every B row
      unsigned* d_hist = nullptr;
      B200_CUDA(T.alloc(&d_hist, (size_t)nw64));
      B200_CUDA(cudaMemsetAsync(d_hist, 0, (size_t)nw64 * sizeof(unsigned), st));
      B200_CUDA(T.alloc(&d_rw, (size_t)RNG + 1));
      B200_CUDA(T.alloc(&d_wr, (size_t)nw64));
      B200_CUDA(T.alloc(&d_split, (size_t)Bs.rows * RNG));
      B200_CUDA(T.alloc(&d_rcnt, (size_t)store_rows * RNG));
      const int hgrid = (int)std::min<long long>((Bs.nnz + 255) / 256, (long long)c.sm_count * 16);
      k_col_hist<<<hgrid, 256, 0, st>>>(Bs.col, Bs.nnz, d_hist);
      // a range is at most half a pool wide, so that any single range fits as a window
      const int wcap = std::max(1, item_pool / 16 / 2);
      k_make_ranges<<<1, 1024, 0, st>>>(d_hist, nw64, Bs.nnz, RNG, wcap, d_rw, d_wr);
      k_range_split<<<(unsigned)(((long long)Bs.rows * 32 + 255) / 256), 256, 0, st>>>(
          Bs.rowptr, Bs.col, Bs.rows, RNG, d_rw, d_split);
      launches += 3;
    }
    if (!sym_smem || !num_smem)
      B200_CUDA(T.alloc(&d_gscr, (size_t)c.sm_count * (bm_pref_bytes / 8)));
    tick(2 * SB_BITMAP);
    sym_timed[SB_BITMAP] = true;
    if (use_parts && !sym_smem) {  // (the whole-row kernel is a little faster when it fits)
      // column-part boundaries inside every B row (shared with the numeric pass)
      B200_CUDA(T.alloc(&d_bsplit, (size_t)std::max(1, nparts - 1) * B.rows));
      if (nparts > 1) {
        const long long nt = (long long)B.rows * (nparts - 1);
        k_bsplit<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(Bs.rowptr, Bs.col, Bs.rows, nparts, wpp, d_bsplit);
        ++launches;
      }
      const long long sitems = (long long)nbig * nparts;
      const int sgrid = (int)std::min<long long>(sitems, (long long)part_ctas * c.sm_count);
      const size_t ssm = walk_part_bytes + (size_t)wpp * 8;
#define LAUNCH_SYM_PART(BTP, MINB)                                                              \
  do {                                                                                          \
    if ((rc = set_smem(k_sym_bitmap_part<BTP, MINB>, ssm))) return rc;                          \
    k_sym_bitmap_part<BTP, MINB><<<sgrid, BTP, ssm, st>>>(                                      \
        sb.d_list + sb.off[SB_BITMAP], nbig, nparts, wpp, row_lo, A.rowptr, A.col, A.val,       \
        Bs.rowptr, Bs.col, nw64, d_bmstore, store_rows, d_bmslot, d_partcnt, d_bsplit, B.rows,  \
        d_work + 0, l2m);                                                                       \
  } while (0)
      if (parts4) LAUNCH_SYM_PART(256, 4); else LAUNCH_SYM_PART(512, 2);
#undef LAUNCH_SYM_PART
      k_sum_parts<<<(nbig + 255) / 256, 256, 0, st>>>(sb.d_list + sb.off[SB_BITMAP], nbig, nparts,
                                                      d_partcnt, d_cnt);
      ++launches;
    } else if (sym_smem) {
      if ((rc = set_smem(k_sym_bitmap<BT_BIG, true>, walk_bytes + bm_bytes))) return rc;
      k_sym_bitmap<BT_BIG, true><<<big_grid, BT_BIG, walk_bytes + bm_bytes, st>>>(
          sb.d_list + sb.off[SB_BITMAP], nbig, row_lo, A.rowptr, A.col, B.rowptr, B.col, d_flops,
          nw64, nullptr, d_bmstore, store_rows, d_bmslot, d_cnt, nparts, wpp, d_partcnt,
          d_work + 0, l2m, d_wr, RNG, d_rcnt);
    } else {
      if ((rc = set_smem(k_sym_bitmap<BT_BIG, false>, walk_bytes))) return rc;
      k_sym_bitmap<BT_BIG, false><<<big_grid, BT_BIG, walk_bytes, st>>>(
          sb.d_list + sb.off[SB_BITMAP], nbig, row_lo, A.rowptr, A.col, B.rowptr, B.col, d_flops,
          nw64, d_gscr, d_bmstore, store_rows, d_bmslot, d_cnt, nparts, wpp, d_partcnt,
          d_work + 0, l2m, nullptr, 0, nullptr);
    }
    tick(2 * SB_BITMAP + 1);
    ++launches;
  }
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaEventRecord(c.ev[2], st));

  // ---- 3. numeric binning, row offsets
  unsigned char* d_nbin = nullptr;
  B200_CUDA(T.alloc(&d_nbin, (size_t)m));
  if (m > 0) {
    const long long light_p = c.tun.light_p >= 0 ? c.tun.light_p : 2560LL * std::max(1, nparts / 2);
    k_num_bins<<<(m + 255) / 256, 256, 0, st>>>(d_cnt, d_flops, m, num_big_from, light_p,
                                                nparts > 2 ? light_p : 0, d_bin, A.rowptr, row_lo, d_fused, d_nbin);
    ++launches;
  }
  B200_CUDA(cudaMemsetAsync(d_cnt + m, 0, sizeof(int), st));
  int64_t* d_urp = nullptr;  // offsets of the unpruned product (== C.rowptr for SpGEMM)
  B200_CUDA(T.alloc(&d_urp, (size_t)m + 1));
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
  B200_CUDA(d2h_small(&h_tot[0], d_urp + m, sizeof(long long), st));
  B200_CUDA(d2h_small(&h_tot[1], d_P, sizeof(long long), st));
  Bins nb;
  rc = make_bins(d_nbin, m, &nb, &launches);  // synchronises the stream
  T.adopt(nb.d_list);
  if (rc) return rc;
  const long long unpruned = h_tot[0];
  const int nbig_num = nb.cnt[NB_BITMAP];
  if ((rc = order_heavy_first(nb.d_list + nb.off[NB_BITMAP], nbig_num))) return rc;
  B200_CUDA(cudaEventRecord(c.ev[3], st));

  // ---- 4. outputs
  RmclOut ro = {};
  int* d_arena_col = nullptr;
  double* d_arena_val = nullptr;
  unsigned long long* d_cursor = nullptr;  // [0] unused, [1] chaos bits
  long long* d_rowoff = nullptr;
  int* d_kept = nullptr;
  int* d_scr_col = nullptr;
  double* d_scr_val = nullptr;
  long long scr_stride = 0;
  if (mode == MODE_SPGEMM) {
    C->rowptr = d_urp;
    C->nnz = unpruned;
    C->sorted_rows = true;
    B200_CUDA(T.alloc(&C->col, (size_t)unpruned));
    B200_CUDA(T.alloc(&C->val, (size_t)unpruned));
  } else {
    if ((rc = ensure_arena((size_t)unpruned))) return rc;
    d_arena_col = c.arena_col;
    d_arena_val = c.arena_val;
    B200_CUDA(T.alloc(&d_cursor, 2));
    B200_CUDA(cudaMemsetAsync(d_cursor, 0, 2 * sizeof(unsigned long long), st));
    B200_CUDA(T.alloc(&d_rowoff, (size_t)m));
    B200_CUDA(T.alloc(&d_kept, (size_t)m + 1));
    ro.arena_col = d_arena_col;
    ro.arena_val = d_arena_val;
    ro.chaos_bits = d_cursor + 1;
    ro.row_off = d_rowoff;
    ro.row_kept = d_kept;
    ro.topk = c.topk;
    if (nbig_num && !use_parts) {
      scr_stride = n;  // a row has at most n distinct columns
      B200_CUDA(T.alloc(&d_scr_col, (size_t)std::min(nbig_num, c.sm_count) * scr_stride));
      B200_CUDA(T.alloc(&d_scr_val, (size_t)std::min(nbig_num, c.sm_count) * scr_stride));
    }
    if (nb.cnt[NB_NONE]) {
      k_rmcl_empty_rows<<<(nb.cnt[NB_NONE] + 255) / 256, 256, 0, st>>>(nb.d_list + nb.off[NB_NONE],
                                                                      nb.cnt[NB_NONE], ro);
      ++launches;
    }
  }

  // ---- 5. numeric per bin
  int* d_wnum = nullptr;   // row counters of the dynamically scheduled warp-table kernels
  B200_CUDA(T.alloc(&d_wnum, 2));
  B200_CUDA(cudaMemsetAsync(d_wnum, 0, 2 * sizeof(int), st));
  auto launch_num_warp = [&](int bin, auto kernel, int CAP, int WPB) -> int {
    const int cntb = nb.cnt[bin];
    if (!cntb) return B200_OK;
    const size_t smem = (size_t)WPB * 24 * CAP;
    int r = set_smem(kernel, smem);
    if (r) return r;
    // (large tables: one or two blocks per SM — a grid that covers the machine, rows drawn dynamically)
    const bool dynamic = CAP >= 1024;
    const int per_sm = std::max(1, (int)((c.smem_optin + 1024) / (smem + 1024)));
    const int grid = dynamic ? std::min((cntb + WPB - 1) / WPB, per_sm * c.sm_count) : (cntb + WPB - 1) / WPB;
    tick(32 + 2 * bin);
    kernel<<<grid, WPB * 32, smem, st>>>(
        nb.d_list + nb.off[bin], cntb, row_lo, A.rowptr, A.col, A.val, B.rowptr, B.col, B.val,
        d_urp, C->col, C->val, ro, dynamic ? d_wnum + (CAP >= 2048 ? 1 : 0) : nullptr);
    tick(32 + 2 * bin + 1);
    num_timed[bin] = true;
    ++launches;
    return B200_OK;
  };
  if (mode == MODE_SPGEMM) {
    for (int bin : {NB_ESC2K, NB_ESC8K, NB_FUSED})
      if (nb.cnt[bin]) {   // rows finished on chip during the symbolic phase: from the arena into C
        tick(32 + 2 * bin);
        k_esc_gather<<<(unsigned)(((long long)nb.cnt[bin] * 32 + 255) / 256), 256, 0, st>>>(
            nb.d_list + nb.off[bin], nb.cnt[bin], d_escoff, d_urp, c.arena_col, c.arena_val, C->col, C->val);
        tick(32 + 2 * bin + 1);
        num_timed[bin] = true;
        ++launches;
      }
    if ((rc = launch_num_warp(NB_W64, k_num_warp<64, false>, 64, 8))) return rc;
    if ((rc = launch_num_warp(NB_W128, k_num_warp<128, false>, 128, 8))) return rc;
    if ((rc = launch_num_warp(NB_W256, k_num_warp<256, false>, 256, 8))) return rc;
    if ((rc = launch_num_warp(NB_W1K, k_num_warp<1024, false>, 1024, 8))) return rc;
    if ((rc = launch_num_warp(NB_W2K, k_num_warp<2048, false>, 2048, 4))) return rc;
  } else {
    if ((rc = launch_num_warp(NB_W64, k_num_warp<64, true>, 64, 8))) return rc;
    if ((rc = launch_num_warp(NB_W128, k_num_warp<128, true>, 128, 8))) return rc;
    if ((rc = launch_num_warp(NB_W256, k_num_warp<256, true>, 256, 8))) return rc;
    if ((rc = launch_num_warp(NB_W1K, k_num_warp<1024, true>, 1024, 8))) return rc;
    if ((rc = launch_num_warp(NB_W2K, k_num_warp<2048, true>, 2048, 4))) return rc;
    num_timed[NB_ESC2K] = nb.cnt[NB_ESC2K] > 0;
    num_timed[NB_ESC8K] = nb.cnt[NB_ESC8K] > 0;
    LAUNCH_ESC(true, nb.d_list + nb.off[NB_ESC2K], nb.cnt[NB_ESC2K], 256, 8, 0, d_urp, ro, nullptr, 32 + 2 * NB_ESC2K)
    LAUNCH_ESC(true, nb.d_list + nb.off[NB_ESC8K], nb.cnt[NB_ESC8K], 512, 16, 1, d_urp, ro, nullptr, 32 + 2 * NB_ESC8K)
  }
#undef LAUNCH_ESC
  if (nbig_num) {
    const int grid = std::min(nbig_num, c.sm_count);
    if (!num_smem && !d_gscr)
      B200_CUDA(T.alloc(&d_gscr, (size_t)c.sm_count * (bm_pref_bytes / 8)));
    const int* lst = nb.d_list + nb.off[NB_BITMAP];
    unsigned long long* d_prof = nullptr;
    if (c.tun.prof) {
      B200_CUDA(T.alloc(&d_prof, 8));
      B200_CUDA(cudaMemsetAsync(d_prof, 0, 8 * sizeof(unsigned long long), st));
    }
    tick(32 + 2 * NB_BITMAP);
    num_timed[NB_BITMAP] = true;
#define LAUNCH_NUM_BM(SM, RM)                                                                   \
  do {                                                                                          \
    const size_t sm_ = walk_bytes + (SM ? bm_pref_bytes : 0);                                   \
    if ((rc = set_smem(k_num_bitmap<BT_BIG, SM, RM>, sm_))) return rc;                          \
    k_num_bitmap<BT_BIG, SM, RM><<<grid, BT_BIG, sm_, st>>>(                                    \
        lst, nbig_num, row_lo, A.rowptr, A.col, A.val, B.rowptr, B.col, B.val, d_flops, nw64,   \
        d_gscr, d_bmstore, d_bmslot, d_urp, C->col, C->val, d_scr_col, d_scr_val, scr_stride,   \
        ro, d_work + 1, d_prof, l2m);                                                           \
  } while (0)
    if (use_parts) {
      // SpGEMM: straight into C; rMCL: the unpruned row into its arena slice, pruned in place by
      // the epilogue kernel
      int* out_col = mode == MODE_SPGEMM ? C->col : d_arena_col;
      double* out_val = mode == MODE_SPGEMM ? C->val : d_arena_val;
      // rows of the part kernel: all of them, or those the range planner could not take
      const int* plist = lst;
      int pcount = nbig_num;
      if (d_rcnt) {
        // ---- range items (ranges.cuh): plan, then one persistent launch
        int *d_nitems = nullptr, *d_itoff = nullptr, *d_fb = nullptr, *d_fbcnt = nullptr;
        B200_CUDA(T.alloc(&d_nitems, (size_t)nbig_num + 1));
        B200_CUDA(T.alloc(&d_itoff, (size_t)nbig_num + 1));
        B200_CUDA(T.alloc(&d_fb, (size_t)nbig_num));
        B200_CUDA(T.alloc(&d_fbcnt, 1));
        B200_CUDA(cudaMemsetAsync(d_nitems + nbig_num, 0, sizeof(int), st));
        B200_CUDA(cudaMemsetAsync(d_fbcnt, 0, sizeof(int), st));
        const int pgridp = (nbig_num + 255) / 256;
        k_plan_items<<<pgridp, 256, 0, st>>>(lst, nbig_num, d_bmslot, d_rcnt, d_rw, RNG, item_pool,
                                             nw64, c.tun.deterministic ? 1 : 0, A.rowptr, row_lo, d_urp,
                                             d_nitems, nullptr, nullptr, d_fb, d_fbcnt);
        {
          void* tmp = nullptr;
          size_t tb = 0;
          cub::DeviceScan::ExclusiveSum(nullptr, tb, d_nitems, d_itoff, nbig_num + 1, st);
          B200_CUDA(cudaMallocAsync(&tmp, tb ? tb : 1, st));
          cub::DeviceScan::ExclusiveSum(tmp, tb, d_nitems, d_itoff, nbig_num + 1, st);
          cudaFreeAsync(tmp, st);
        }
        int h_items = 0, h_fb = 0;
        B200_CUDA(d2h_small(&h_items, d_itoff + nbig_num, sizeof(int), st));
        B200_CUDA(d2h_small(&h_fb, d_fbcnt, sizeof(int), st));
        B200_CUDA(sync_fetch(st));
        launches += 2;
        if (h_items > 0) {
          RangeItem* d_items = nullptr;
          B200_CUDA(T.alloc(&d_items, (size_t)h_items));
          k_plan_items<<<pgridp, 256, 0, st>>>(lst, nbig_num, d_bmslot, d_rcnt, d_rw, RNG, item_pool,
                                               nw64, c.tun.deterministic ? 1 : 0, A.rowptr, row_lo, d_urp,
                                               d_nitems, d_itoff, d_items, d_fb, d_fbcnt);
          if ((rc = set_smem(k_num_units<UNIT_WARPS>, unit_dyn))) return rc;
          const int igrid = std::min((h_items + UNIT_WARPS - 1) / UNIT_WARPS, UNIT_CTAS * c.sm_count);
          k_num_units<UNIT_WARPS><<<igrid, UNIT_WARPS * 32, unit_dyn, st>>>(
              d_items, h_items, RNG, unit_pool, A.col, A.val, Bs.rowptr, Bs.col, Bs.val, d_split,
              d_bmstore, out_col, out_val, d_work + 3, l2m);
          launches += 2;
          if (c.tun.prof) fprintf(stderr, "[b200 prof] k_num_units: %d items, %d fallback rows, pool %d B per warp\n",
                                  h_items, h_fb, unit_pool);
        }
        range_items = h_items;
        plist = d_fb;
        pcount = h_fb;
      }
      if (pcount > 0) {
      // team sizes -> ticket offsets of the (row, part) slots
      const int nslots = pcount * nparts;
      int* d_tsize = nullptr;
      int* d_ready = nullptr;
      B200_CUDA(T.alloc(&d_tsize, (size_t)nslots + 1));
      B200_CUDA(T.alloc(&d_itemoff, (size_t)nslots + 1));
      B200_CUDA(T.alloc(&d_ready, (size_t)nslots));
      B200_CUDA(cudaMemsetAsync(d_tsize + nslots, 0, sizeof(int), st));
      B200_CUDA(cudaMemsetAsync(d_ready, 0, (size_t)nslots * sizeof(int), st));
      if (!d_bsplit) {  // no row went through the part-wise symbolic kernel
        B200_CUDA(T.alloc(&d_bsplit, (size_t)std::max(1, nparts - 1) * B.rows));
        if (nparts > 1) {
          const long long nt = (long long)B.rows * (nparts - 1);
          k_bsplit<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(Bs.rowptr, Bs.col, Bs.rows, nparts, wpp, d_bsplit);
          ++launches;
        }
      }
      // a team must fit, with room to spare, among the CTAs that are resident at once (its
      // members wait for each other): at most half of one CTA per SM
      const int team_cap = std::max(1, c.sm_count / 2);
      const int team_max = std::min(team_cap, c.tun.team_max);
      k_team_sizes<<<(nslots + 255) / 256, 256, 0, st>>>(plist, pcount, nparts, d_flops, d_partcnt,
                                                         c.tun.team_products, team_max, d_tsize);
      {
        void* tmp = nullptr;
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, d_tsize, d_itemoff, nslots + 1, st);
        B200_CUDA(cudaMallocAsync(&tmp, tb ? tb : 1, st));
        cub::DeviceScan::ExclusiveSum(tmp, tb, d_tsize, d_itemoff, nslots + 1, st);
        cudaFreeAsync(tmp, st);
      }
      int* d_ticket_slot = nullptr;  // at most team_max tickets per slot, usually ~1.2 per slot
      int h_tickets = 0;
      B200_CUDA(d2h_small(&h_tickets, d_itemoff + nslots, sizeof(int), st));
      B200_CUDA(sync_fetch(st));
      B200_CUDA(T.alloc(&d_ticket_slot, (size_t)h_tickets + 1));
      k_fill_tickets<<<(nslots + 255) / 256, 256, 0, st>>>(d_itemoff, nslots, d_ticket_slot);
      launches += 3;
      const int pgrid = part_ctas * c.sm_count;  // all resident: team members wait for each other

#define LAUNCH_PART(BTP, MINB)                                                                  \
  do {                                                                                          \
    if ((rc = set_smem(k_num_bitmap_part<BTP, MINB>, part_smem))) return rc;                    \
    k_num_bitmap_part<BTP, MINB><<<pgrid, BTP, part_smem, st>>>(                                \
        plist, pcount, nparts, wpp, row_lo, A.rowptr, A.col, A.val, Bs.rowptr, Bs.col, Bs.val,  \
        nw64, d_bmstore, d_bmslot, d_partcnt, d_urp, out_col, out_val, d_itemoff,               \
        d_ticket_slot, d_ready, d_bsplit, B.rows, d_work + 1, l2m);                             \
  } while (0)
      if (parts4) LAUNCH_PART(256, 4); else LAUNCH_PART(512, 2);
#undef LAUNCH_PART
      }
      if (mode == MODE_RMCL) {
        const int egrid = std::min(nbig_num, c.sm_count * 8);
        k_rmcl_epilogue_rows<256><<<egrid, 256, 0, st>>>(lst, nbig_num, d_urp, ro, d_work + 2);
        ++launches;
      }

    } else if (num_smem) {
      if (mode == MODE_SPGEMM) LAUNCH_NUM_BM(true, false); else LAUNCH_NUM_BM(true, true);
    } else {
      if (mode == MODE_SPGEMM) LAUNCH_NUM_BM(false, false); else LAUNCH_NUM_BM(false, true);
    }
#undef LAUNCH_NUM_BM
    tick(32 + 2 * NB_BITMAP + 1);
    ++launches;
    if (d_prof) {
      unsigned long long h[8];
      B200_CUDA(d2h_small(h, d_prof, sizeof h, st));
      B200_CUDA(sync_fetch(st));
      fprintf(stderr, "[b200 prof] k_num_bitmap Mcycles/CTA: bitmap %.2f prefix %.2f emit %.2f products %.2f epilogue %.2f (grid %d)\n",
              h[0] / 1e6 / grid, h[1] / 1e6 / grid, h[2] / 1e6 / grid, h[3] / 1e6 / grid, h[4] / 1e6 / grid, grid);

    }
  }
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaEventRecord(c.ev[4], st));

  // ---- 6. rMCL: scan kept counts, gather arena -> final CSR
  long long nnz_out = unpruned;
  unsigned long long h_esc_true = 0;
  if (mode == MODE_RMCL) {
    B200_CUDA(cudaMemsetAsync(d_kept + m, 0, sizeof(int), st));
    B200_CUDA(T.alloc(&C->rowptr, (size_t)m + 1));
    cub::TransformInputIterator<long long, IntToI64, const int*> it(d_kept, IntToI64());
    void* tmp = nullptr;
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, it, (long long*)C->rowptr, m + 1, st);
    B200_CUDA(cudaMallocAsync(&tmp, tb ? tb : 1, st));
    cub::DeviceScan::ExclusiveSum(tmp, tb, it, (long long*)C->rowptr, m + 1, st);
    cudaFreeAsync(tmp, st);
    ++launches;
    unsigned long long h_cur[2] = {0, 0};
    long long h_kept_total = 0;
    if (d_esc) B200_CUDA(d2h_small(&h_esc_true, d_esc, sizeof(unsigned long long), st));
    B200_CUDA(d2h_small(h_cur, d_cursor, 2 * sizeof(unsigned long long), st));
    B200_CUDA(d2h_small(&h_kept_total, C->rowptr + m, sizeof(long long), st));
    B200_CUDA(sync_fetch(st));
    nnz_out = h_kept_total;
    if (chaos) {
      long long bits = (long long)h_cur[1];
      memcpy(chaos, &bits, sizeof(double));
    }
    C->nnz = nnz_out;
    B200_CUDA(T.alloc(&C->col, (size_t)nnz_out));
    B200_CUDA(T.alloc(&C->val, (size_t)nnz_out));
    if (m > 0) {
      const long long threads = (long long)m * 32;
      k_gather_rows<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(
          m, d_rowoff, C->rowptr, d_arena_col, d_arena_val, C->col, C->val);
      ++launches;
    }


  }
  B200_CUDA(cudaEventRecord(c.ev[5], st));
  unsigned long long h_agg[96] = {0};
  if (stats && m > 0) {
    unsigned long long* d_agg = nullptr;
    B200_CUDA(T.alloc(&d_agg, 96));
    B200_CUDA(cudaMemsetAsync(d_agg, 0, 96 * sizeof(unsigned long long), st));
    const int grid = std::min((m + 255) / 256, c.sm_count * 8);
    k_bin_aggregate<<<grid, 256, 0, st>>>(d_bin, m, A.rowptr, row_lo, d_flops, nullptr, d_agg);
    k_bin_aggregate<<<grid, 256, 0, st>>>(d_nbin, m, A.rowptr, row_lo, d_flops, d_cnt, d_agg + 48);
    B200_CUDA(d2h_small(h_agg, d_agg, sizeof h_agg, st));

  }



  B200_CUDA(sync_fetch(st));
  B200_CUDA(cudaGetLastError());
  T.keep(C->rowptr);
  T.keep(C->col);
  T.keep(C->val);
  out.ok = true;
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
    if (c.tun.prof)
      fprintf(stderr, "[b200 prof] rows %d..%d: flops %.2f symbolic %.2f bins+offsets %.2f numeric %.2f compaction %.2f ms\n",
              row_lo, row_hi, t01, t12, t23, t34, t45);
    stats->products = h_tot[1];
    stats->nnz_out = nnz_out;
    // (rows sorted on chip reserved their products as a bound: report what they really held)
    stats->nnz_unpruned = unpruned;
    if (d_esc && mode == MODE_RMCL) {
      const long long bound = (long long)h_agg[48 + NB_ESC2K * 3 + 0] + (long long)h_agg[48 + NB_ESC8K * 3 + 0];
      stats->nnz_unpruned = unpruned - bound + (long long)h_esc_true;
    }
    stats->launches = launches;
    stats->part_kernel = use_parts ? 1 : 0;
    stats->part_count = nparts;
    stats->range_items = range_items;
    stats->ranges = range_items ? RNG : 0;
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
