// ranges.cuh — numeric pass of the heavy rows with ON-CHIP, ORDERED accumulation (included by
// spgemm.cu, inside its anonymous namespace).  Opt-in (B200_ON_CHIP / B200_DETERMINISTIC): it
// reproduces the reference's values bit for bit and run to run, at 1.2 - 1.5x the time of the
// default part kernel (global fp64 RED) on R-MAT scale 20 — DESIGN.md §3b has the measurements.
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
// shared-memory load gives the rank of a column) fit the shared-memory pool of ONE WARP
// (k_num_units, below).  An item whose single range overflows the pool keeps its accumulators in
// the row's final slice of C.val (zeroed first) and uses one fp64 RED per product.

constexpr int RANGES_MAX = 128;
// how an item accumulates: on chip, B-row segments in A-entry order (bit-exact), or — when one
// range alone overflows the warp's pool — in HBM with fp64 RED (order not fixed); with
// B200_DETERMINISTIC such a range is instead walked once per pool-sized piece of its columns
constexpr int ITEM_ORD = 0, ITEM_RED = 1, ITEM_MULTI = 2;  // MULTI: on chip, in several column passes

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
  int R, pool_bytes, nw64, deterministic;
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
    const int mode = forced_red ? (pc.deterministic ? ITEM_MULTI : ITEM_RED) : ITEM_ORD;
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
    if (c > 0 && 8 * (c + cr) + 16 * (pc.rw[r + 1] - pc.rw[r0]) > pc.pool_bytes) {
      emit(r0, rend, c, false);
      c = 0;
    }
    if (c == 0) {
      r0 = r;
      if (8 * cr + 16 * (pc.rw[r + 1] - pc.rw[r]) > pc.pool_bytes) {
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
             int nw64, int deterministic,
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
  pc.rw = rw; pc.R = R; pc.pool_bytes = pool_bytes; pc.nw64 = nw64; pc.deterministic = deterministic;
  pc.a0 = Arp[row_lo + i];
  pc.nA = (int)(Arp[row_lo + i + 1] - pc.a0);
  pc.ob = Crp[i];
  pc.bmrow = (long long)slot * nw64;
  pc.out = items ? items + itemoff[t] : nullptr;
  const int n = plan_row(rcnt + (size_t)slot * R, pc);
  if (!items) nitems[t] = n;
}

// ---- the item kernel ----------------------------------------------------------------------
// One WARP per item, warp-synchronous from start to end: no block barrier, no atomics, no shared
// state between warps.  A 128-thread CTA is four independent workers with a quarter of the
// CTA's shared memory each; four CTAs per SM give 16 workers whose latencies overlap.
//   window   the item's bitmap words -> {32 bits, rank prefix} pairs in the warp's pool;
//   walk     the A entries 32 at a time (lane = entry: its B-row segment inside the item's ranges
//            comes from two adjacent loads of the split table), then the segments ONE AT A TIME,
//            lanes across a segment: a B row has unique columns, so the lanes update distinct
//            accumulators with a plain shared-memory load / add / store, and every output entry
//            sums its products in ascending A-entry order with separately rounded multiply and
//            add — indexProcessCRowI's order and rounding (cpu_csr_kernel.h:159-170): the values
//            are bit-identical to the reference's and reproducible.  Loads of four segment
//            chunks are issued together before the first of them is used;
//   flush    values, then the columns (expanded from the window into the same pool), as
//            coalesced streams.
constexpr int UNIT_WARPS = 4;     // workers per CTA
constexpr int UNIT_CHUNK = 4;     // items a worker draws per ticket
constexpr int UNIT_DEPTH = 8;     // segment chunks whose loads are in flight together

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 4)
k_num_units(const RangeItem* __restrict__ items, int nitems, int R, int pool_bytes,
            const int* __restrict__ Acol, const double* __restrict__ Aval,
            const int64_t* __restrict__ Brp, const int* __restrict__ Bcol,
            const double* __restrict__ Bval, const unsigned* __restrict__ split,
            const unsigned long long* __restrict__ bm_store, int* __restrict__ Ccol,
            double* __restrict__ Cval, int* __restrict__ work_counter, L2Modes l2) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* pool = smem_raw + (size_t)warp * pool_bytes;
  const unsigned long long pol_acc = l2_policy(l2.acc), pol_out = l2_policy(L2_FIRST),
                           pol_b = l2_policy(l2.bgather), pol_bm = l2_policy(l2.bmstore);
  // descriptor fields travel one per lane: lanes 0..11 hold the current item's, 16..27 the next's
  auto field = [&](int dv, int k) { return __shfl_sync(FULL, dv, k); };
  auto field64 = [&](int dv, int k) {
    return ((long long)__shfl_sync(FULL, dv, k + 1) << 32) | (unsigned)__shfl_sync(FULL, dv, k);
  };
  while (true) {
    int t0 = 0;
    if (lane == 0) t0 = atomicAdd(work_counter, UNIT_CHUNK);
    t0 = __shfl_sync(FULL, t0, 0);
    if (t0 >= nitems) break;
    const int t1 = min(nitems, t0 + UNIT_CHUNK);
    for (int t = t0; t < t1; ++t) {
      int dv = 0;
      {
        const int k = lane & 15, tt = t + (lane >> 4);
        if (k < 12 && tt < t1) dv = __ldg(reinterpret_cast<const int*>(items + tt) + k);
      }
      const int desc = field(dv, 0), cnt = field(dv, 1), W = field(dv, 2), c_lo = field(dv, 3);
      const long long a0 = field64(dv, 4), ob = field64(dv, 6), bmoff = field64(dv, 8);
      const int nA = field(dv, 10);
      if (t + 1 < t1) {
        // the next item of this worker: its window and its A entries requested from L2 now
        const int nW = field(dv, 16 + 2), nnA = field(dv, 16 + 10);
        const long long nbm = field64(dv, 16 + 8), na0 = field64(dv, 16 + 4);
        for (int w = lane * 16; w < nW; w += 32 * 16)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(bm_store + nbm + w));
        if (lane * 32 < nnA) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(Acol + na0 + lane * 32));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(Aval + na0 + lane * 32));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(Aval + na0 + lane * 32 + 16));
        }
      }
      const int r0 = desc & 255, r1 = (desc >> 8) & 255, mode = desc >> 16;
      const bool on_chip = mode != ITEM_RED;
      const unsigned long long* src = bm_store + bmoff;
      // MULTI: the item's columns are walked in pieces [sw0, sw1) of its window, each piece's
      // accumulators fitting the pool (16-word steps: 1024 columns can always be held)
      int sw0 = 0;
      long long ob_cur = ob;
      do {
      int sw1 = W, scnt = cnt;
      if (mode == ITEM_MULTI) {
        scnt = 0;
        sw1 = sw0;
        while (sw1 < W) {
          const int nwb = min(16, W - sw1);
          const int blk = warp_sum_int(lane < nwb ? __popcll(ldg_hint(src + sw1 + lane, pol_bm)) : 0);
          if (sw1 > sw0 && 8 * (scnt + blk) + 16 * (sw1 - sw0 + nwb) > pool_bytes - 16) break;
          scnt += blk;
          sw1 += nwb;
        }
      }
      const int SW = sw1 - sw0;                 // words of this piece
      const int sc_lo = c_lo + 64 * sw0, sc_hi = c_lo + 64 * sw1;
      double* acc = reinterpret_cast<double*>(pool);
      uint2* pk = reinterpret_cast<uint2*>(pool + (on_chip ? (((size_t)scnt * 8 + 15) & ~(size_t)15) : 0));
      double* gacc = Cval + ob_cur;
      // ---- window.  Pass 1: every word fetched with independent loads (one round trip for the
      // whole window), raw bits parked in the pool; the accumulators are zeroed in its shadow.
      {
#pragma unroll 4
        for (int w = lane; w < SW; w += 32) {
          const unsigned long long x = ldg_hint(src + sw0 + w, pol_bm);
          pk[2 * w].x = (unsigned)x;
          pk[2 * w + 1].x = (unsigned)(x >> 32);
        }
        if (on_chip) {
          for (int k = lane; k < scnt; k += 32) acc[k] = 0.0;
        } else {
          for (int k = lane; k < scnt; k += 32) stg_hint(gacc + k, 0.0, pol_acc);
        }
        __syncwarp();
        // pass 2: rank prefix, 32 words (64 packed entries) per step
        int run = 0;
        for (int w0 = 0; w0 < SW; w0 += 32) {
          const int w = w0 + lane;
          unsigned lo32 = 0, hi32 = 0;
          if (w < SW) { lo32 = pk[2 * w].x; hi32 = pk[2 * w + 1].x; }
          const int cl = __popc(lo32), c = cl + __popc(hi32);
          int inc = c;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += y;
          }
          if (w < SW) {
            const unsigned ex = (unsigned)(run + inc - c);
            pk[2 * w].y = ex;
            pk[2 * w + 1].y = ex + cl;
          }
          run += __shfl_sync(FULL, inc, 31);
        }
        __syncwarp();
      }
      // ---- walk
      for (long long b0 = 0; b0 < nA; b0 += 32) {
        long long seg_s = 0;
        int seg_l = 0;
        double seg_a = 0.0;
        if (b0 + lane < nA) {
          const int j = __ldg(Acol + a0 + b0 + lane);
          seg_a = __ldg(Aval + a0 + b0 + lane);
          const long long rp = __ldg(Brp + j);
          const unsigned* sp = split + (size_t)j * R;
          const long long o0 = (r0 == 0) ? 0 : (long long)__ldg(sp + r0 - 1);
          const long long o1 = (r1 == R) ? (__ldg(Brp + j + 1) - rp) : (long long)__ldg(sp + r1 - 1);
          seg_s = rp + o0;
          seg_l = (int)(o1 - o0);
        }
        unsigned todo = __ballot_sync(FULL, seg_l > 0);
        long long cur_s = 0;
        int rem = 0;
        double cur_a = 0.0;
        while (todo || rem) {
          // up to UNIT_DEPTH chunks (<= 32 consecutive products of one segment each), in order:
          // all their loads first, then all their rank look-ups, then the updates one by one
          int cc[UNIT_DEPTH], rk[UNIT_DEPTH];
          double cv[UNIT_DEPTH], ca[UNIT_DEPTH];
          unsigned fresh = 0;  // bit u: chunk u starts a new segment (its updates wait for the previous one's)
#pragma unroll
          for (int u = 0; u < UNIT_DEPTH; ++u) {
            cc[u] = -1;
            cv[u] = 0.0;
            ca[u] = 0.0;
            if (rem == 0 && todo) {
              const int e = __ffs((int)todo) - 1;
              todo &= todo - 1;
              cur_s = shfl64(seg_s, e);
              rem = __shfl_sync(FULL, seg_l, e);
              cur_a = shfld(seg_a, e);
              fresh |= 1u << u;
            }
            if (rem > 0) {
              if (lane < rem) {
                cc[u] = ldg_hint(Bcol + cur_s + lane, pol_b);
                cv[u] = ldg_hint(Bval + cur_s + lane, pol_b);
                // (a piece of a MULTI item sees the whole segment and takes its own columns)
                if (mode == ITEM_MULTI && (cc[u] < sc_lo || cc[u] >= sc_hi)) cc[u] = -1;
              }
              ca[u] = cur_a;
              const int n = min(rem, 32);
              cur_s += n;
              rem -= n;
            }
          }
#pragma unroll
          for (int u = 0; u < UNIT_DEPTH; ++u) {
            rk[u] = 0;
            if (cc[u] >= 0) {
              const uint2 p = pk[(cc[u] - sc_lo) >> 5];
              rk[u] = (int)p.y + __popc(p.x & ((1u << (cc[u] & 31)) - 1u));
              cv[u] = __dmul_rn(ca[u], cv[u]);
            }
          }
#pragma unroll
          for (int u = 0; u < UNIT_DEPTH; ++u) {
            // chunks of one segment touch distinct entries: only a new segment has to wait
            if ((fresh >> u) & 1u) __syncwarp();
            if (cc[u] >= 0) {
              if (on_chip) acc[rk[u]] = __dadd_rn(acc[rk[u]], cv[u]);
              else red_add_hint(gacc + rk[u], cv[u], pol_acc);
            }
          }
          __syncwarp();
        }
      }
      // ---- flush: values, then the columns through the pool
      if (on_chip)
        for (int k = lane; k < scnt; k += 32) stg_hint(Cval + ob_cur + k, acc[k], pol_out);
      __syncwarp();
      const bool stage = on_chip || (size_t)SW * 16 + (size_t)scnt * 4 <= (size_t)pool_bytes;
      int* cols = on_chip ? reinterpret_cast<int*>(pool) : reinterpret_cast<int*>(pool + (size_t)SW * 16);
      for (int w = lane; w < 2 * SW; w += 32) {
        const uint2 p = pk[w];
        unsigned x = p.x;
        int pos = (int)p.y;
        const int cb = sc_lo + w * 32;
        while (x) {
          const int b = __ffs((int)x) - 1;
          x &= x - 1;
          if (stage) cols[pos++] = cb + b;
          else stg_hint(Ccol + ob_cur + pos++, cb + b, pol_out);
        }
      }
      __syncwarp();
      if (stage)
        for (int k = lane; k < scnt; k += 32) stg_hint(Ccol + ob_cur + k, cols[k], pol_out);
      __syncwarp();
      ob_cur += scnt;
      sw0 = sw1;
      } while (sw0 < W);
    }
  }
}
