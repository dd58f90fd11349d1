// esc.cuh — expand / sort / compress for MID-SIZE rows of an rMCL step (included by spgemm.cu,
// inside its anonymous namespace).
//
// Reference behaviour being replaced: indexProcessCRowI + the row epilogue of
// static_omp_CSR_RMCL_OneStep (nlibs/cpu_csr_kernel.h:134-188, static_omp_csr_kernel.cc:264-271).
//
// Rows with a few hundred to a few thousand intermediate products in a VERY wide matrix (the
// planted-partition and R-MAT scale-22 iterates of BASELINE configs C4 / C5: ~1 - 8 K products,
// ~1 - 4 K distinct columns out of 4 M) fit neither accumulator of this library well: a warp's hash
// table for 2 K entries takes 48 KB (4 warps per SM), and the column bitmap costs one work item per
// 512 K-column part the row touches (8 items of ~20 us each).  Such a row is small enough to be
// SORTED on chip instead: one CTA expands its products into registers in (A entry, position in
// the B row) order, a stable block radix sort orders them by column (cub::BlockRadixSort over the
// column bits that the matrix width needs), and the entry that heads a run of equal columns adds
// the run up front to back — the products of a column in A-entry order with separately rounded
// multiply and add, i.e. the order and rounding of the reference: the unpruned values are
// bit-identical to indexProcessCRowI's.  The row is then inflated, thresholded, pruned and
// normalised where it lies (fixed-shape block sums, as in k_rmcl_epilogue_rows) and only the
// kept entries leave the chip.  No symbolic pass is needed for these rows: a row's products
// bound its unpruned entries, so its arena slice is reserved by the flops analysis alone.
// Works on unsorted B rows (no column cut is needed).

#ifndef ESC_RADIX_BITS
#define ESC_RADIX_BITS 6
#endif
template <int BT, int IPT>
struct EscSmem {
  // 6-bit digits: a 22-bit column (4 M columns) takes 4 ranking passes instead of 6
  typedef cub::BlockRadixSort<unsigned, BT, IPT, double, ESC_RADIX_BITS> Sort;
  union {
    typename Sort::TempStorage sort;
    struct { int col[BT * IPT]; double val[BT * IPT]; } row;   // the sorted products, then the reduced row
  } u;
  long long pre[BT + 1];   // exclusive prefix of the B-row lengths of the A entries (one batch: nnz(A_i) <= BT)
  long long bstart[BT];
  double aval[BT];
  int red[BT / 32 + 1];
  double redd[BT / 32 + 1];
};

// RMCL = false (plain SpGEMM): the same expand / sort / compress runs in the SYMBOLIC phase — it
// leaves the finished row (ascending columns) in the arena slice reserved by the row's products
// and its length in rownnz; after the row offsets of C are known, k_esc_gather moves it there.
template <int BT, int IPT, bool RMCL>
__global__ void __launch_bounds__(BT)
k_esc_rmcl(const int* __restrict__ list, int count, int row_lo, const int64_t* __restrict__ Arp,
           const int* __restrict__ Acol, const double* __restrict__ Aval,
           const int64_t* __restrict__ Brp, const int* __restrict__ Bcol,
           const double* __restrict__ Bval, const int64_t* __restrict__ Crp, int col_bits, RmclOut ro,
           unsigned long long* __restrict__ true_unpruned, int* __restrict__ work_counter,
           int* __restrict__ rownnz) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  EscSmem<BT, IPT>& sm = *reinterpret_cast<EscSmem<BT, IPT>*>(smem_raw);
  __shared__ int s_idx;
  constexpr int CAP = BT * IPT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  while (true) {
    __syncthreads();
    if (threadIdx.x == 0) s_idx = atomicAdd(work_counter, 1);
    __syncthreads();
    const int idx = s_idx;
    if (idx >= count) break;
    const int i = list[idx];
    const int64_t a0 = Arp[row_lo + i];
    const int nA = (int)(Arp[row_lo + i + 1] - a0);   // <= BT: the bin's rule (products <= CAP and nA <= BT)
    // ---- the A entries and the prefix of their B-row lengths
    long long len = 0;
    if ((int)threadIdx.x < nA) {
      const int j = __ldg(Acol + a0 + threadIdx.x);
      const long long bs = __ldg(Brp + j);
      len = __ldg(Brp + j + 1) - bs;
      sm.bstart[threadIdx.x] = bs;
      sm.aval[threadIdx.x] = __ldg(Aval + a0 + threadIdx.x);
    }
    long long total;
    {
      // exclusive scan of len over the block (warp scans + warp totals)
      long long inc = len;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long y = shfl_up64(inc, o);
        if (lane >= o) inc += y;
      }
      __shared__ long long s_wt[BT / 32];
      if (lane == 31) s_wt[warp] = inc;
      __syncthreads();
      long long base = 0, tot = 0;
#pragma unroll
      for (int w = 0; w < BT / 32; ++w) { const long long x = s_wt[w]; base += (w < warp) ? x : 0; tot += x; }
      sm.pre[threadIdx.x] = base + inc - len;
      if (threadIdx.x == 0) sm.pre[BT] = tot;
      total = tot;
      __syncthreads();
    }
    const int P = (int)total;   // <= CAP
    // ---- expand: thread t holds products [t*IPT, (t+1)*IPT) of the row, in reference order
    unsigned keys[IPT];
    double vals[IPT];
    {
      const int x0 = threadIdx.x * IPT;
      int e = 0;
      if (x0 < P) {
        int lo = 0, hi = nA - 1;  // last entry with pre[e] <= x0
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (sm.pre[mid] <= (long long)x0) lo = mid; else hi = mid - 1;
        }
        e = lo;
      }
#pragma unroll
      for (int k = 0; k < IPT; ++k) {
        const int x = x0 + k;
        keys[k] = 0xffffffffu;   // padding sorts to the end
        vals[k] = 0.0;
        if (x < P) {
          while (e + 1 < nA && sm.pre[e + 1] <= (long long)x) ++e;   // (skips empty B rows too)
          const long long q = sm.bstart[e] + ((long long)x - sm.pre[e]);
          keys[k] = (unsigned)__ldg(Bcol + q);
          vals[k] = __dmul_rn(sm.aval[e], __ldg(Bval + q));
        }
      }
    }
    __syncthreads();
    // ---- stable sort by column (equal columns keep the order above: A entry, then B position)
    EscSmem<BT, IPT>::Sort(sm.u.sort).Sort(keys, vals, 0, col_bits);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
      const int x = threadIdx.x * IPT + k;
      sm.u.row.col[x] = (int)keys[k];
      sm.u.row.val[x] = vals[k];
    }
    __syncthreads();
    // ---- compress: the head of a run of equal columns adds the run up, front to back
    int heads = 0;
    unsigned headmask = 0;
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
      const int x = threadIdx.x * IPT + k;
      const bool head = x < P && (x == 0 || sm.u.row.col[x - 1] != sm.u.row.col[x]);
      heads += head ? 1 : 0;
      headmask |= head ? (1u << k) : 0u;
    }
    int cnt;
    const int hbase = block_excl_scan<BT>(heads, sm.red, &cnt);   // barriers inside
    // (results are kept in registers until every run has been read, then written over the row)
    int ocol[IPT];
    double oval[IPT];
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
      ocol[k] = -1;
      oval[k] = 0.0;
      if ((headmask >> k) & 1u) {
        int x = threadIdx.x * IPT + k;
        const int c = sm.u.row.col[x];
        double s = sm.u.row.val[x];
        for (++x; x < P && sm.u.row.col[x] == c; ++x) s = __dadd_rn(s, sm.u.row.val[x]);
        ocol[k] = c;
        oval[k] = s;
      }
    }
    __syncthreads();
    {
      int h = 0;
#pragma unroll
      for (int k = 0; k < IPT; ++k)
        if (ocol[k] >= 0) { sm.u.row.col[hbase + h] = ocol[k]; sm.u.row.val[hbase + h] = oval[k]; ++h; }
    }
    __syncthreads();
    if (!RMCL) {
      const long long off = (long long)Crp[i];   // (here: the row's slice of the arena)
      for (int k = threadIdx.x; k < cnt; k += BT) {
        ro.arena_col[off + k] = sm.u.row.col[k];
        ro.arena_val[off + k] = sm.u.row.val[k];
      }
      if (threadIdx.x == 0) rownnz[i] = cnt;
      continue;
    }
    // ---- rMCL epilogue over the row in shared memory (ascending columns; nlibs/tools/util.cc:4-69)
    int* rcol = sm.u.row.col;
    double* acc = sm.u.row.val;
    double psum = 0.0, pmax = 0.0;
    for (int k = threadIdx.x; k < cnt; k += BT) {
      const double v2 = __dmul_rn(acc[k], acc[k]);
      acc[k] = v2;
      psum = __dadd_rn(psum, v2);
      pmax = fmax(pmax, v2);
    }
    const double rsum = block_sum_d<BT>(psum, sm.redd);
    const double rmax = block_max_d<BT>(pmax, sm.redd);
    const double thresh = compute_threshold(__ddiv_rn(rsum, (double)cnt), rmax);
    double tk = 0.0;
    int cstar = 0x7fffffff;
    bool cut = false;
    if (ro.topk > 0) {
      auto bcount = [&](auto pred) {
        int c = 0;
        for (int k = threadIdx.x; k < cnt; k += BT) c += (acc[k] >= thresh && pred(acc[k], rcol[k])) ? 1 : 0;
        return block_sum_int<BT>(c, sm.red);
      };
      if (bcount([](double, int) { return true; }) > ro.topk) {
        cut = true;
        topk_cut(bcount, thresh, rmax, ro.topk, 0x7fffffff, &tk, &cstar);
      }
    }
    auto keeps = [&](double v, int c) { return v >= thresh && (!cut || v > tk || (v == tk && c <= cstar)); };
    double ksum_p = 0.0;
    int kept_p = 0;
    for (int k = threadIdx.x; k < cnt; k += BT)
      if (keeps(acc[k], rcol[k])) { ksum_p = __dadd_rn(ksum_p, acc[k]); ++kept_p; }
    const double ksum = block_sum_d<BT>(ksum_p, sm.redd);
    const int kept = block_sum_int<BT>(kept_p, sm.red);
    const long long off = (long long)Crp[i];
    double sq_p = 0.0;
    int written = 0;
    for (int k0 = 0; k0 < cnt; k0 += BT) {
      const int k = k0 + threadIdx.x;
      const double v2 = (k < cnt) ? acc[k] : 0.0;
      const int c = (k < cnt) ? rcol[k] : 0;
      const bool keep = (k < cnt) && keeps(v2, c);
      int tot;
      const int ex = block_excl_scan<BT>(keep ? 1 : 0, sm.red, &tot);
      if (keep) {
        const double w = __ddiv_rn(v2, ksum);
        ro.arena_col[off + written + ex] = c;
        ro.arena_val[off + written + ex] = w;
        sq_p = __dadd_rn(sq_p, __dmul_rn(w, w));
      }
      written += tot;
    }
    const double sq = block_sum_d<BT>(sq_p, sm.redd);
    if (threadIdx.x == 0) {
      ro.row_off[i] = off;
      ro.row_kept[i] = kept;
      double ch = (kept > 0) ? __dsub_rn(__ddiv_rn(rmax, ksum), sq) : 0.0;
      if (ch < 0.0) ch = 0.0;
      atomicMax(ro.chaos_bits, (unsigned long long)__double_as_longlong(ch));
      atomicAdd(true_unpruned, (unsigned long long)cnt);
    }
  }
}

// SpGEMM: rows finished in the arena by k_esc_rmcl<.., false> move to their place in C
__global__ void __launch_bounds__(256)
k_esc_gather(const int* __restrict__ list, int count, const int64_t* __restrict__ escoff,
             const int64_t* __restrict__ Crp, const int* __restrict__ arena_col,
             const double* __restrict__ arena_val, int* __restrict__ Ccol, double* __restrict__ Cval) {
  const int w = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (w >= count) return;
  const int i = list[w];
  const int64_t s = escoff[i], o = Crp[i];
  const int k = (int)(Crp[i + 1] - o);
  for (int t = lane; t < k; t += 32) { Ccol[o + t] = arena_col[s + t]; Cval[o + t] = arena_val[s + t]; }
}
