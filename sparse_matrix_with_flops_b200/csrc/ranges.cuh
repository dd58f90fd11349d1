// ranges.cuh — numeric pass of the heavy rows with ON-CHIP accumulation (included by spgemm.cu,
// inside its anonymous namespace).
//
// Reference behaviour being replaced: indexProcessCRowI, nlibs/cpu_csr_kernel.h:134-188 — one
// dense accumulator per output row, products added in (A entry, B entry) order.
//
// The columns of B are cut ONCE PER CALL into R static ranges of (about) equal column mass
// (k_col_hist + k_make_ranges: every range holds ~1/R of the entries of B, so a hot low-index
// range is narrow and a cold one wide), and the offsets of the range boundaries inside every
// sorted B row are tabulated (k_range_split, row-major [row][R]: the device form of the
// reference's column-striped PCSR, nlibs/PCSR.cc:3-56, without copying B).  The symbolic pass
// counts the output columns of every heavy row per range (rcnt).  A planner then groups
// consecutive ranges of a row into work ITEMS whose accumulators (8 B per output column) and
// column index (16 B per 64 columns: {32 bitmap bits, 32-bit rank prefix} pairs, so one
// shared-memory load gives the rank of a column) fit the shared-memory pool of a CTA.
//
// An item is handled by one CTA (k_num_items):
//   table     the B-row segment [start, start+len) of every A entry inside the item's ranges
//             (two loads from the split table, no search), padded to whole warp slots;
//   window    the item's bitmap words from the store of the symbolic pass -> packed pairs;
//   products  a flat, coalesced walk over the segments, loads three steps ahead of use.  Every
//             lane looks up its rank, then the segments COMMIT IN A-ENTRY ORDER: a B row has
//             unique columns, so the lanes of one segment update distinct accumulators with a
//             plain shared-memory load / add / store (no atomics: shared atomics cost 2 cycles
//             per lane, a global RED 1.3), and a barrier separates consecutive segments.  Every
//             output entry therefore sums its products in ascending A-entry order with
//             separately rounded multiply and add — the order and rounding of the reference:
//             the values are BIT-IDENTICAL to indexProcessCRowI, and run-to-run reproducible;
//   flush     values and columns leave through shared memory as coalesced streams.
// Items whose segments are too short for that (hub rows: thousands of A entries, a dozen
// products each per range) or whose single range overflows the pool run in RED mode: same
// table / window / walk, accumulators in the row's final slice of C.val (zeroed first), one
// fp64 RED per product, order not fixed (values to rounding).

constexpr int RANGES_MAX = 64;
// how an item accumulates: on chip with the segments committed in A-entry order (bit-exact), in
// HBM with fp64 RED, or on chip with tag arbitration (order not fixed)
constexpr int ITEM_ORD = 0, ITEM_RED = 1, ITEM_TAG = 2;
constexpr int SLOTS_MAX = 2048;  // lane groups (32 or 8 lanes) of one table batch

// Everything the kernel needs to start an item, so that an item costs ONE dependent load (its
// descriptor, fetched while the previous item is being processed).
struct RangeItem {
  int desc;          // r0 | r1 << 8 | mode << 16
  int cnt;           // output columns inside the ranges
  int W;             // bitmap words of the window
  int c_lo;          // first column of the window
  long long a0;      // the row's A entries: [a0, a0 + nA)
  long long ob;      // first output position (in C.col / C.val, or the rMCL arena)
  long long bmoff;   // first window word inside the bitmap store
  int nA;
  int pad;
};
static_assert(sizeof(RangeItem) == 48, "descriptor is three 16-byte loads");

// entries of B per bitmap word (64 columns)
__global__ void __launch_bounds__(256)
k_col_hist(const int* __restrict__ Bcol, long long nnz, unsigned* __restrict__ hist) {
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < nnz;
       p += (long long)gridDim.x * blockDim.x)
    atomicAdd(hist + (__ldg(Bcol + p) >> 6), 1u);
}

// range boundaries in bitmap words: rw[0..R], rw[0] = 0, rw[R] = nw64; a range closes when it
// holds total/R entries or wcap words.  One CTA; wr[w] = range of word w.
__global__ void __launch_bounds__(1024)
k_make_ranges(const unsigned* __restrict__ hist, int nw64, long long total, int R, int wcap,
              int* __restrict__ rw, unsigned char* __restrict__ wr) {
  __shared__ unsigned long long s_part[1024];
  __shared__ int s_rw[RANGES_MAX + 1];
  const int per = (nw64 + 1023) / 1024;
  const int w0 = threadIdx.x * per, w1 = min(nw64, w0 + per);
  unsigned long long s = 0;
  for (int w = w0; w < w1; ++w) s += hist[w];
  s_part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    // walk over the 1024 partial sums; only a partial in which a range can close is walked
    // word by word (at most R - 1 + nw64 / wcap of them)
    const unsigned long long target = (unsigned long long)((total + R - 1) / R);
    int r = 0, start = 0;
    unsigned long long acc = 0;
    s_rw[0] = 0;
    for (int p = 0; p < 1024 && r < R - 1; ++p) {
      const int p0 = p * per, p1 = min(nw64, p0 + per);
      if (p0 >= p1) break;
      if (acc + s_part[p] < target && p1 - start < wcap) { acc += s_part[p]; continue; }
      for (int w = p0; w < p1 && r < R - 1; ++w) {
        acc += hist[w];
        if (acc >= target || w + 1 - start >= wcap) {
          s_rw[++r] = w + 1;
          start = w + 1;
          acc = 0;
        }
      }
    }
    while (r < R) s_rw[++r] = nw64;
  }
  __syncthreads();
  if (threadIdx.x <= R) rw[threadIdx.x] = s_rw[threadIdx.x];
  for (int w = threadIdx.x; w < nw64; w += 1024) {
    int lo = 0, hi = R - 1;  // last r with rw[r] <= w
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (s_rw[mid] <= w) lo = mid; else hi = mid - 1;
    }
    wr[w] = (unsigned char)lo;
  }
}

// split[j*R + r-1] = offset inside the sorted B row j of its first column >= 64*rw[r], r = 1..R-1
// (slot R-1 of a row is unused).  One warp per row: short rows by ballot, long ones by search.
__global__ void __launch_bounds__(256)
k_range_split(const int64_t* __restrict__ Brp, const int* __restrict__ Bcol, int krows, int R,
              const int* __restrict__ rw, unsigned* __restrict__ split) {
  __shared__ int s_b[RANGES_MAX + 1];
  if (threadIdx.x <= R) s_b[threadIdx.x] = rw[threadIdx.x] * 64;
  __syncthreads();
  const int j = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (j >= krows) return;
  const long long s = Brp[j], e = Brp[j + 1];
  const int len = (int)(e - s);
  unsigned* out = split + (size_t)j * R;
  if (len <= 32) {
    const int c = lane < len ? __ldg(Bcol + s + lane) : 0x7fffffff;
    for (int r0 = 1; r0 < R; r0 += 32) {
      unsigned mine = 0;
      for (int k = 0; k < 32 && r0 + k < R; ++k) {
        const unsigned below = __popc(__ballot_sync(FULL, c < s_b[r0 + k]));
        if (k == lane) mine = below;
      }
      if (r0 + lane < R) out[r0 + lane - 1] = mine;
    }
  } else {
    for (int r = 1 + lane; r < R; r += 32) {
      const int key = s_b[r];
      long long lo = s, hi = e;
      while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (__ldg(Bcol + mid) < key) lo = mid + 1; else hi = mid;
      }
      out[r - 1] = (unsigned)(lo - s);
    }
  }
}

// ---- planner: groups of consecutive ranges per row ------------------------------------------
// One thread per row of the numeric bitmap list.  pass 0 counts the items of the row
// (nitems[t]; 0 = the row is not planned: no stored bitmap -> it goes to the fallback list),
// pass 1 writes them at itemoff[t].
struct PlanCtx {
  const int* rw;
  int R, pool_bytes, chunk_min, nw64, short_mode;
  double seg_per_range;  // expected products of one A entry inside one range
  long long a0, ob, bmrow;
  int nA;
  RangeItem* out;
};
__device__ __forceinline__ int plan_row(const int* __restrict__ rc, const PlanCtx& pc) {
  int n = 0, off = 0;
  // a pending RED group: adjacent RED groups merge while their window fits the pool
  int pr0 = -1, pr1 = 0, pcnt = 0;
  auto write = [&](int a, int b, int cnt, int mode) {
    if (pc.out) {
      RangeItem it;
      it.desc = a | (b << 8) | (mode << 16);
      it.cnt = cnt;
      it.W = pc.rw[b] - pc.rw[a];
      it.c_lo = pc.rw[a] * 64;
      it.a0 = pc.a0;
      it.ob = pc.ob + off;
      it.bmoff = pc.bmrow + pc.rw[a];
      it.nA = pc.nA;
      it.pad = 0;
      pc.out[n] = it;
    }
    off += cnt;
    ++n;
  };
  auto flush_red = [&]() {
    if (pr0 >= 0) write(pr0, pr1, pcnt, ITEM_RED);
    pr0 = -1;
    pcnt = 0;
  };
  auto emit = [&](int a, int b, int cnt, bool forced_red) {
    // expected products of one A entry inside the group: long segments commit in order (a
    // barrier per segment pays above ~100 products), short ones use tag arbitration (or RED)
    const bool is_short = pc.seg_per_range * (b - a) < (double)pc.chunk_min;
    const int mode = forced_red ? ITEM_RED : (is_short ? pc.short_mode : ITEM_ORD);
    if (mode != ITEM_RED) { flush_red(); write(a, b, cnt, mode); return; }
    if (pr0 >= 0 && 16 * (pc.rw[b] - pc.rw[pr0]) > pc.pool_bytes) flush_red();
    if (pr0 < 0) pr0 = a;
    pr1 = b;
    pcnt += cnt;
  };
  // a group runs from its first to its last non-empty range (empty ranges at either end would
  // only widen the window)
  int r0 = 0, rend = 0, c = 0;
  for (int r = 0; r < pc.R; ++r) {
    const int cr = rc[r];
    if (cr == 0) continue;
    if (c > 0 && 10 * (c + cr) + 16 * (pc.rw[r + 1] - pc.rw[r0]) > pc.pool_bytes) {
      emit(r0, rend, c, false);
      c = 0;
    }
    if (c == 0) {
      r0 = r;
      if (10 * cr + 16 * (pc.rw[r + 1] - pc.rw[r]) > pc.pool_bytes) {
        emit(r, r + 1, cr, true);  // one range overflows the pool: its accumulators stay in HBM
        continue;
      }
    }
    c += cr;
    rend = r + 1;
  }
  if (c > 0) emit(r0, rend, c, false);
  flush_red();
  return n;
}

__global__ void __launch_bounds__(256)
k_plan_items(const int* __restrict__ list, int count, const int* __restrict__ bm_slot,
             const int* __restrict__ rcnt, const int* __restrict__ rw, int R, int pool_bytes,
             int chunk_min, int short_mode, int nw64, const long long* __restrict__ flops,
             const int64_t* __restrict__ Arp, int row_lo, const int64_t* __restrict__ Crp,
             int* __restrict__ nitems, const int* __restrict__ itemoff,
             RangeItem* __restrict__ items, int* __restrict__ fb_list, int* __restrict__ fb_count) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const int i = list[t];
  const int slot = bm_slot[i];
  if (slot < 0) {
    if (!items) { nitems[t] = 0; fb_list[atomicAdd(fb_count, 1)] = i; }
    return;
  }
  PlanCtx pc;
  pc.rw = rw; pc.R = R; pc.pool_bytes = pool_bytes; pc.chunk_min = chunk_min; pc.nw64 = nw64; pc.short_mode = short_mode;
  pc.a0 = Arp[row_lo + i];
  pc.nA = (int)(Arp[row_lo + i + 1] - pc.a0);
  pc.seg_per_range = (double)flops[i] / (double)max(1, pc.nA) / (double)R;
  pc.ob = Crp[i];
  pc.bmrow = (long long)slot * nw64;
  pc.out = items ? items + itemoff[t] : nullptr;
  const int n = plan_row(rcnt + (size_t)slot * R, pc);
  if (!items) nitems[t] = n;
}

// ---- the item kernel ----------------------------------------------------------------------
template <int BT>
struct ItemSmem {
  long long start[BT];  // first B entry of the segment
  double a[BT];         // A value of the entry
  int ws[BT + 1];       // first warp slot of the segment (exclusive scan of the slot counts)
  int len[BT];          // products of the segment
  unsigned short eslot[SLOTS_MAX];  // warp slot -> segment
  int red[BT / 32 + 2];
  int next[12];         // descriptor of this CTA's next item (fetched one item ahead)
  // staging ring of the product walk: every thread copies its own (column, value) of the steps
  // s+1 .. s+3 with cp.async while step s is being accumulated
  int rcol[3][BT];
  double rval[3][BT];
};

// Items are dealt round-robin to the CTAs (they are bounded in size and there are thousands per
// CTA, so a dynamic ticket buys nothing): the next item of a CTA is known, its descriptor is
// fetched and its window / A entries are requested from L2 while the current item runs.
template <int BT, int MINB>
__global__ void __launch_bounds__(BT, MINB)
k_num_items(const RangeItem* __restrict__ items, int nitems, int R,
            const int* __restrict__ Acol, const double* __restrict__ Aval,
            const int64_t* __restrict__ Brp, const int* __restrict__ Bcol,
            const double* __restrict__ Bval, const unsigned* __restrict__ split,
            const unsigned long long* __restrict__ bm_store, int* __restrict__ Ccol,
            double* __restrict__ Cval, L2Modes l2, unsigned long long* __restrict__ prof) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NW = BT / 32;
  // developer diagnostics (B200_PROF): cycles per phase, summed over this CTA's items
  long long pcyc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tsub = 0;
  auto sub = [&](int k) {
    if (prof && threadIdx.x == 0) { const long long t = clock64(); pcyc[k] += t - tsub; tsub = t; }
  };
  long long tprev = 0;
  auto lap = [&](int k) {
    if (prof && threadIdx.x == 0) { const long long t = clock64(); pcyc[k] += t - tprev; tprev = t; }
  };
  ItemSmem<BT>& sm = *reinterpret_cast<ItemSmem<BT>*>(smem_raw);
  unsigned char* pool = smem_raw + ((sizeof(ItemSmem<BT>) + 15) & ~(size_t)15);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long pol_acc = l2_policy(l2.acc), pol_out = l2_policy(L2_FIRST),
                           pol_b = l2_policy(l2.bgather), pol_bm = l2_policy(l2.bmstore);
  if ((int)blockIdx.x < nitems && threadIdx.x < 12)
    sm.next[threadIdx.x] = reinterpret_cast<const int*>(items + blockIdx.x)[threadIdx.x];
  for (int t = blockIdx.x; t < nitems; t += gridDim.x) {
    __syncthreads();
    if (prof && threadIdx.x == 0) tprev = clock64();
    RangeItem it;
    {
      int* d = reinterpret_cast<int*>(&it);
#pragma unroll
      for (int k = 0; k < 12; ++k) d[k] = sm.next[k];
    }
    __syncthreads();
    // the descriptor after this one: loaded now, parked in shared memory at the end of the item
    const int tn = t + gridDim.x;
    int nx = 0;
    if (tn < nitems && threadIdx.x < 12) nx = __ldg(reinterpret_cast<const int*>(items + tn) + threadIdx.x);
    const int r0 = it.desc & 255, r1 = (it.desc >> 8) & 255, mode = it.desc >> 16;
    const int cnt = it.cnt, W = it.W, c_lo = it.c_lo;
    const int64_t a0 = it.a0, a1 = it.a0 + it.nA;
    const int64_t ob = it.ob;
    const bool on_chip = mode != ITEM_RED;
    // lanes per slot of the flat walk: whole warps when segments commit in order, 8-lane groups
    // otherwise (a 13-product segment padded to 32 lanes would idle most of them)
    const int gsh = mode == ITEM_ORD ? 5 : 3;
    const int GPS = BT >> gsh;  // slots per step
    // pool: accumulators (on-chip modes), the packed window, the tags (tag mode)
    double* acc = reinterpret_cast<double*>(pool);
    const size_t acc_bytes = on_chip ? (((size_t)cnt * 8 + 15) & ~(size_t)15) : 0;
    uint2* pk = reinterpret_cast<uint2*>(pool + acc_bytes);
    unsigned short* tag = reinterpret_cast<unsigned short*>(pool + acc_bytes + (size_t)W * 16);
    // ---- window: bitmap words -> {bits of 32 columns, rank of the first of them}.  Pass 1:
    // every word fetched with independent loads (one round trip to HBM for the whole window);
    // pass 2: the popcount prefix from shared memory, warp w owning a contiguous run.
    {
      const unsigned long long* src = bm_store + it.bmoff;
      for (int w = threadIdx.x; w < W; w += BT) {
        const unsigned long long x = ldg_hint(src + w, pol_bm);
        pk[2 * w].x = (unsigned)x;
        pk[2 * w + 1].x = (unsigned)(x >> 32);
      }
      // accumulators zeroed under the shadow of those loads
      if (on_chip) {
        for (int k = threadIdx.x; k < cnt; k += BT) acc[k] = 0.0;
      } else {
        for (int k = threadIdx.x; k < cnt; k += BT) stg_hint(Cval + ob + k, 0.0, pol_acc);
      }
      __syncthreads();
      const int wpw = (((W + NW - 1) / NW) + 31) & ~31;
      const int wbeg = warp * wpw, wend = min(W, wbeg + wpw);
      int run = 0;
      for (int w = wbeg + lane; w < wbeg + wpw; w += 32) {
        unsigned lo32 = 0, hi32 = 0;
        if (w < wend) { lo32 = pk[2 * w].x; hi32 = pk[2 * w + 1].x; }
        const int cl = __popc(lo32), c = cl + __popc(hi32);
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int y = __shfl_up_sync(FULL, inc, o);
          if (lane >= o) inc += y;
        }
        if (w < wend) {
          const unsigned ex = (unsigned)(run + inc - c);
          pk[2 * w].y = ex;
          pk[2 * w + 1].y = ex + cl;
        }
        run += __shfl_sync(FULL, inc, 31);
      }
      if (lane == 0) sm.red[warp] = run;
      __syncthreads();
      int wbase = 0;
#pragma unroll
      for (int k = 0; k < NW; ++k) wbase += (k < warp) ? sm.red[k] : 0;
      if (wbase)
        for (int w = wbeg + lane; w < wend; w += 32) { pk[2 * w].y += wbase; pk[2 * w + 1].y += wbase; }
    }
    double* gacc = Cval + ob;
    lap(0);
    // ---- products: batches of up to BT A entries (fewer when their slots exceed the table)
    int e_prev = -1;  // last segment committed by this CTA (barriers separate segments)
    for (int64_t b0 = a0; b0 < a1;) {
      int nb = (int)min((int64_t)BT, a1 - b0);
      int myslots = 0;
      if ((int)threadIdx.x < nb) {
        const int j = __ldg(Acol + b0 + threadIdx.x);
        const long long rp = __ldg(Brp + j);
        const unsigned* sp = split + (size_t)j * R;
        const long long o0 = (r0 == 0) ? 0 : (long long)__ldg(sp + r0 - 1);
        const long long o1 = (r1 == R) ? (__ldg(Brp + j + 1) - rp) : (long long)__ldg(sp + r1 - 1);
        sm.start[threadIdx.x] = rp + o0;
        sm.len[threadIdx.x] = (int)(o1 - o0);
        sm.a[threadIdx.x] = __ldg(Aval + b0 + threadIdx.x);
        myslots = (int)((o1 - o0 + (1 << gsh) - 1) >> gsh);
      }
      int total;
      const int ex = block_excl_scan<BT>(myslots, sm.red, &total);  // barriers inside
      // entries whose slots fit the table; a single oversized segment is walked alone
      bool single = false;
      if (total > SLOTS_MAX) {
        const int fit = __syncthreads_count((int)threadIdx.x < nb && ex + myslots <= SLOTS_MAX);
        single = fit == 0;
        nb = single ? 1 : fit;  // (slot counts are non-negative: the fitting entries are a prefix)
      }
      if ((int)threadIdx.x < nb) sm.ws[threadIdx.x] = ex;
      if ((int)threadIdx.x == nb) sm.ws[nb] = ex;   // (thread nb exists when nb < BT)
      if (nb == BT && threadIdx.x == 0) sm.ws[BT] = total;
      __syncthreads();
      const int nslots = single ? (int)((sm.len[0] + (1 << gsh) - 1) >> gsh) : sm.ws[nb];
      if (!single)
        for (int sl = threadIdx.x; sl < nslots; sl += BT) {
          int lo = 0, hi = nb - 1;  // last e with ws[e] <= sl
          while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (sm.ws[mid] <= sl) lo = mid; else hi = mid - 1;
          }
          sm.eslot[sl] = (unsigned short)lo;
        }
      __syncthreads();
      const int nsteps = (nslots + GPS - 1) / GPS;
      lap(1);
      // a stage of the staging ring.  (Register prefetch does not pipeline here: ptxas puts the
      // loads of all stages on one scoreboard, so the first use of stage s waits for the loads
      // of stages s+1 and s+2 as well; cp.async groups are tracked per group.)
      const int gl = threadIdx.x >> gsh, gk = threadIdx.x & ((1 << gsh) - 1);
      auto issue = [&](int s, int stage, int& e) {
        const int sl = s * GPS + gl;
        e = -1;
        if (sl < nslots) {
          const int ee = single ? 0 : (int)sm.eslot[sl];
          const int k = ((sl - (single ? 0 : sm.ws[ee])) << gsh) + gk;
          if (k < sm.len[ee]) {
            const long long q = sm.start[ee] + k;
            e = ee;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(&sm.rcol[stage][threadIdx.x])), "l"(Bcol + q) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(&sm.rval[stage][threadIdx.x])), "l"(Bval + q) : "memory");
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
      auto consume = [&](int s, int stage, int e) {
        int rank = 0, c = -1;
        double prod = 0.0;
        if (prof && threadIdx.x == 0) tsub = clock64();
        asm volatile("cp.async.wait_group 2;" ::: "memory");  // this step's copies have landed
        if (e >= 0) {
          c = sm.rcol[stage][threadIdx.x];
          const double v = sm.rval[stage][threadIdx.x];
          const uint2 p = pk[(c - c_lo) >> 5];
          rank = (int)p.y + __popc(p.x & ((1u << (c & 31)) - 1u));
          prod = __dmul_rn(sm.a[e], v);
        }
        if (prof) { asm volatile("" ::"r"(rank), "d"(prod)); sub(4); }
        if (mode == ITEM_RED) {
          if (c >= 0) red_add_hint(gacc + rank, prod, pol_acc);
        } else if (mode == ITEM_TAG) {
          // tag arbitration: every pending lane writes its id over its accumulator's tag; after
          // the barrier the lane that reads its own id back owns the accumulator for this pass
          // (two lanes can never both read their own id in one pass: the tag only changes when a
          // loser re-writes it for the next pass, which can at worst make a winner retry).
          // Commits of consecutive passes are separated by the next pass's barrier.
          bool pending = c >= 0;
          while (true) {
            if (pending) tag[rank] = (unsigned short)threadIdx.x;
            if (!__syncthreads_or(pending)) break;
            if (pending && tag[rank] == (unsigned short)threadIdx.x) {
              acc[rank] = __dadd_rn(acc[rank], prod);
              pending = false;
            }
          }
        } else {
          // ordered commit: the distinct segments of this step, in order, one barrier each.
          // Lanes 0..NW-1 of every warp look at the step's slots (one per warp).
          const int sl = s * NW + lane;
          const bool valid = lane < NW && sl < nslots;
          const int sv = valid ? (single ? 0 : (int)sm.eslot[sl]) : -1;
          const int pv = __shfl_up_sync(FULL, sv, 1);
          const unsigned starts = __ballot_sync(FULL, valid && (lane == 0 || sv != pv));
          const int nrounds = __popc(starts);
          const int my_round = __popc(starts & ((2u << warp) - 1u)) - 1;
          const int e_first = __shfl_sync(FULL, sv, 0);
          const int e_last = __shfl_sync(FULL, sv, __popc(__ballot_sync(FULL, valid)) - 1);
          for (int r = 0; r < nrounds; ++r) {
            if (r > 0 || e_first != e_prev) __syncthreads();  // the previous segment is complete
            if (r == my_round && c >= 0) acc[rank] = __dadd_rn(acc[rank], prod);
          }
          e_prev = e_last;
        }
        sub(5);
      };
      int e0, e1, e2;
      issue(0, 0, e0);
      issue(1, 1, e1);
      issue(2, 2, e2);
      for (int s = 0; s < nsteps; s += 3) {
        consume(s, 0, e0);
        issue(s + 3, 0, e0);
        sub(6);
        if (s + 1 < nsteps) { consume(s + 1, 1, e1); issue(s + 4, 1, e1); }
        if (s + 2 < nsteps) { consume(s + 2, 2, e2); issue(s + 5, 2, e2); }
      }
      // (a segment walked alone may be longer than the table: all its slots belong to it)
      b0 += nb;
      e_prev = -1;  // segment numbers restart with the next batch: force a barrier first
      __syncthreads();
      lap(2);
    }
    // ---- the next item: descriptor parked, its window and A entries requested from L2
    if (tn < nitems) {
      if (threadIdx.x < 12) sm.next[threadIdx.x] = nx;
      if (warp == 0) {
        const int nW = __shfl_sync(FULL, nx, 2), nnA = __shfl_sync(FULL, nx, 10);
        const long long nbm = ((long long)__shfl_sync(FULL, nx, 9) << 32) | (unsigned)__shfl_sync(FULL, nx, 8);
        const long long na0 = ((long long)__shfl_sync(FULL, nx, 5) << 32) | (unsigned)__shfl_sync(FULL, nx, 4);
        for (int w = lane * 16; w < nW; w += 32 * 16)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(bm_store + nbm + w));
        for (int k = lane * 32; k < min(nnA, BT); k += 32 * 32) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(Acol + na0 + k));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(Aval + na0 + k));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(Aval + na0 + k + 16));
        }
      }
    }
    // ---- flush
    if (on_chip) {
      for (int k = threadIdx.x; k < cnt; k += BT) stg_hint(Cval + ob + k, acc[k], pol_out);
      __syncthreads();
      int* cols = reinterpret_cast<int*>(pool);  // the accumulators' space, now free
      for (int w = threadIdx.x; w < 2 * W; w += BT) {
        const uint2 p = pk[w];
        unsigned x = p.x;
        int pos = (int)p.y;
        const int cb = c_lo + w * 32;
        while (x) {
          const int b = __ffs((int)x) - 1;
          x &= x - 1;
          cols[pos++] = cb + b;
        }
      }
      __syncthreads();
      for (int k = threadIdx.x; k < cnt; k += BT) stg_hint(Ccol + ob + k, cols[k], pol_out);
    } else {
      for (int w = threadIdx.x; w < 2 * W; w += BT) {
        const uint2 p = pk[w];
        unsigned x = p.x;
        int pos = (int)p.y;
        const int cb = c_lo + w * 32;
        while (x) {
          const int b = __ffs((int)x) - 1;
          x &= x - 1;
          stg_hint(Ccol + ob + pos++, cb + b, pol_out);
        }
      }
    }
    lap(3);
  }
  if (prof && threadIdx.x == 0)
    for (int k = 0; k < 8; ++k) atomicAdd(prof + k, (unsigned long long)pcyc[k]);
}
