// capi.cu — the C-ABI (include/b200_spgemm.h): context, device CSR container, host-buffer and
// device-handle entry points.  No algorithm lives here; see spgemm.cu.
#include <limits.h>
#include <stdint.h>
#include <malloc.h>
#include <sys/mman.h>
#include <unistd.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <omp.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "common.cuh"

namespace b200 {

static thread_local std::string g_err;
void set_error(const std::string& s) { g_err = s; }
int fail_cuda(cudaError_t e, const char* what, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e),
           file, line, what);
  g_err = buf;
  return B200_ERR_CUDA;
}
Ctx& ctx() {
  static Ctx c;
  return c;
}

namespace {

__global__ void k_i32_to_i64(const int* __restrict__ in, int64_t* __restrict__ out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}
// out[i] = in[lo+i] - in[lo]
__global__ void k_i64_to_i32_rebased(const int64_t* __restrict__ in, int* __restrict__ out,
                                     long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int)(in[i] - in[0]);
}
// argmax per row, ties -> smallest column; one thread per row (rows of a converged Mt are tiny)
__global__ void k_row_argmax(const int64_t* __restrict__ rp, const int* __restrict__ col,
                             const double* __restrict__ val, int m, int* __restrict__ lab) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  int best = -1;
  double bv = 0.0;
  for (int64_t p = rp[i]; p < rp[i + 1]; ++p) {
    const double v = val[p];
    const int c = col[p];
    if (best < 0 || v > bv || (v == bv && c < best)) { best = c; bv = v; }
  }
  lab[i] = best;
}

// ---- staged host <-> device copies ------------------------------------------------------------
// The reference's containers are plain malloc() blocks (nlibs/CSR.h:323-327), i.e. pageable
// memory; a direct cudaMemcpy to or from it runs at a few GB/s (driver-side staging on one
// thread plus first-touch page faults).  Instead: DMA through two pinned buffers on a second
// stream while all host cores copy the previous chunk between the pinned buffer and the user's
// block (which also first-touches a fresh malloc block in parallel).
// Chunks of 32 MB: large enough for the full DMA rate, small enough that the library's own small
// read-backs (row totals, bin counts: the same copy engine serves them) wait at most ~1 ms behind
// a streamed download; with 128 MB chunks a row block computed next to a download took 200 ms
// instead of 30.
constexpr size_t PIN_BYTES = 32u << 20;

int ensure_staging() {
  Ctx& c = ctx();
  if (c.pin[0]) return B200_OK;
  B200_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
  for (int k = 0; k < 2; ++k) {
    B200_CUDA(cudaHostAlloc(&c.pin[k], PIN_BYTES, cudaHostAllocDefault));
    B200_CUDA(cudaEventCreateWithFlags(&c.pin_ev[k], cudaEventDisableTiming));
  }
  B200_CUDA(cudaEventCreateWithFlags(&c.xfer_ev, cudaEventDisableTiming));
  return B200_OK;
}

// Ask for transparent huge pages on the 2 MB-aligned interior of a large result block before it
// is first touched: a 10 GB block is 2.6 M page faults with 4 KB pages, 5 K with 2 MB pages.
void advise_huge(void* p, size_t bytes) {
  const uintptr_t two_mb = (uintptr_t)2 << 20;
  if (!p || bytes < (16u << 20)) return;
  const uintptr_t a = ((uintptr_t)p + two_mb - 1) & ~(two_mb - 1);
  const uintptr_t e = ((uintptr_t)p + bytes) & ~(two_mb - 1);
  if (e > a) madvise((void*)a, e - a, MADV_HUGEPAGE);
}

// Streaming copy: non-temporal 16-byte stores spare the destination the read-for-ownership of a
// regular store and keep 128 MB chunks out of the host caches.  Measured on the B200 box (16
// cores, tools/micro/d2h_rate.cu): pinned -> fresh malloc block 35 GB/s against 15 GB/s for
// memcpy, pinned -> touched block 57 against 36; the DMA itself runs at 56.6 GB/s.
void stream_copy(char* dst, const char* src, size_t bytes) {
#if defined(__SSE2__)
  const size_t head = std::min(bytes, (size_t)((16 - ((uintptr_t)dst & 15)) & 15));
  if (head) { memcpy(dst, src, head); dst += head; src += head; bytes -= head; }
  const size_t n16 = bytes / 16;
  __m128i* d = (__m128i*)dst;
  const __m128i* s = (const __m128i*)src;
  size_t i = 0;
  for (; i + 4 <= n16; i += 4) {
    const __m128i a = _mm_loadu_si128(s + i), b = _mm_loadu_si128(s + i + 1);
    const __m128i c = _mm_loadu_si128(s + i + 2), e = _mm_loadu_si128(s + i + 3);
    _mm_stream_si128(d + i, a); _mm_stream_si128(d + i + 1, b);
    _mm_stream_si128(d + i + 2, c); _mm_stream_si128(d + i + 3, e);
  }
  for (; i < n16; ++i) _mm_stream_si128(d + i, _mm_loadu_si128(s + i));
  if (bytes & 15) memcpy(dst + n16 * 16, src + n16 * 16, bytes & 15);
  _mm_sfence();
#else
  memcpy(dst, src, bytes);
#endif
}

// one contiguous piece per host thread; `spare` cores are left to other busy threads (the
// caller of the streamed product spins in the CUDA runtime while the download worker copies)
void parallel_copy(void* dst, const void* src, size_t bytes, int spare = 0) {
#pragma omp parallel num_threads(std::max(1, omp_get_max_threads() - spare))
  {
    const size_t nt = (size_t)omp_get_num_threads(), t = (size_t)omp_get_thread_num();
    const size_t piece = (((bytes + nt - 1) / nt) + 4095) & ~(size_t)4095;
    const size_t off = t * piece;
    if (off < bytes) stream_copy((char*)dst + off, (const char*)src + off, std::min(piece, bytes - off));
  }
}

// true if [dst, dst + bytes) lies inside a page-locked block of the host cache (defined below)
bool host_block_is_pinned(const void* dst, size_t bytes);

// device -> host block.  ordered == true: everything queued on the library stream before the
// call is waited for.  ordered == false (download worker): only the copy stream is touched; the
// submitting thread has already made the copy stream wait for the producer.
int d2h_staged(void* dst, const void* src, size_t bytes, bool ordered = true) {
  Ctx& c = ctx();
  if (!bytes) return B200_OK;
  if (bytes < (4u << 20)) {
    cudaStream_t st = ordered ? c.stream : c.copy_stream;
    B200_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    return B200_OK;
  }
  if (ordered) {
    int rc = ensure_staging();
    if (rc) return rc;
    B200_CUDA(cudaEventRecord(c.xfer_ev, c.stream));
    B200_CUDA(cudaStreamWaitEvent(c.copy_stream, c.xfer_ev, 0));
  }
  if (host_block_is_pinned(dst, bytes)) {
    // a block of the page-locked cache (b200_host_cache_pin): one DMA straight into it, no
    // staging buffer and no host copy
    B200_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c.copy_stream));
    B200_CUDA(cudaStreamSynchronize(c.copy_stream));
    return B200_OK;
  }
  const size_t nchunks = (bytes + PIN_BYTES - 1) / PIN_BYTES;
  for (size_t k = 0; k <= nchunks; ++k) {
    if (k < nchunks) {
      const size_t off = k * PIN_BYTES, len = std::min(PIN_BYTES, bytes - off);
      B200_CUDA(cudaMemcpyAsync(c.pin[k & 1], (const char*)src + off, len, cudaMemcpyDeviceToHost, c.copy_stream));
      B200_CUDA(cudaEventRecord(c.pin_ev[k & 1], c.copy_stream));
    }
    if (k > 0) {
      const size_t off = (k - 1) * PIN_BYTES, len = std::min(PIN_BYTES, bytes - off);
      B200_CUDA(cudaEventSynchronize(c.pin_ev[(k - 1) & 1]));
      parallel_copy((char*)dst + off, c.pin[(k - 1) & 1], len, ordered ? 0 : 1);
    }
  }
  return B200_OK;
}

// host block -> device; returns after the last DMA has been queued AND completed
int h2d_staged(void* dst, const void* src, size_t bytes) {
  Ctx& c = ctx();
  if (!bytes) return B200_OK;
  if (bytes < (4u << 20)) {
    B200_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c.stream));
    B200_CUDA(cudaStreamSynchronize(c.stream));
    return B200_OK;
  }
  int rc = ensure_staging();
  if (rc) return rc;
  B200_CUDA(cudaEventRecord(c.xfer_ev, c.stream));   // dst was allocated on the library stream
  B200_CUDA(cudaStreamWaitEvent(c.copy_stream, c.xfer_ev, 0));
  const size_t nchunks = (bytes + PIN_BYTES - 1) / PIN_BYTES;
  for (size_t k = 0; k < nchunks; ++k) {
    const size_t off = k * PIN_BYTES, len = std::min(PIN_BYTES, bytes - off);
    if (k >= 2) B200_CUDA(cudaEventSynchronize(c.pin_ev[k & 1]));  // buffer free again
    parallel_copy(c.pin[k & 1], (const char*)src + off, len);
    B200_CUDA(cudaMemcpyAsync((char*)dst + off, c.pin[k & 1], len, cudaMemcpyHostToDevice, c.copy_stream));
    B200_CUDA(cudaEventRecord(c.pin_ev[k & 1], c.copy_stream));
  }
  B200_CUDA(cudaStreamSynchronize(c.copy_stream));
  return B200_OK;
}

// ---- host block cache ---------------------------------------------------------------------------
// Result blocks are plain malloc() memory that the caller owns and may free().  A 10 GB block
// that comes fresh from the OS costs ~0.3 s of page faults on the B200 box before the first
// byte lands (tools/micro/d2h_rate.cu: 35 GB/s into fresh pages, 57 GB/s into touched ones, the
// DMA runs at 56.6).  A caller that hands blocks back through b200_host_free() instead of
// free() lets the library keep them (still malloc blocks, their capacity read back with
// malloc_usable_size) for the next download, up to B200_HOST_CACHE_GB (default: a quarter of the
// physical memory, at most 64 GB; 0 turns the cache off).
//
// Opt-in (b200_host_cache_pin(1) or B200_HOST_PIN=1): a block that enters the cache is
// page-locked (cudaHostRegister, once) and stays so while it circulates, so later downloads DMA
// straight into it — no pinned staging buffer, no host copy; with several ranks on one host the
// staging copies are what bounds the download (two host threads per rank at 8 ranks).  The price
// is the contract: in this mode blocks MUST come back through b200_host_free(), never free().
struct HostCache {
  std::mutex mu;
  std::vector<std::pair<size_t, void*>> blocks;   // (capacity, pointer)
  std::vector<std::pair<char*, size_t>> pinned;   // page-locked blocks, kept or handed out
  size_t total = 0, limit = 0;
  long hits = 0, misses = 0, direct = 0;
  bool limit_known = false;
  int pin = -1;                                    // -1: read B200_HOST_PIN on first use
  static constexpr size_t BIG = (size_t)64 << 20;

  bool pin_mode() {
    if (pin < 0) { const char* e = getenv("B200_HOST_PIN"); pin = (e && atoi(e) != 0) ? 1 : 0; }
    return pin > 0;
  }
  // (mu held)
  int find_pinned(const void* p) const {
    for (int i = 0; i < (int)pinned.size(); ++i) if ((const void*)pinned[i].first == p) return i;
    return -1;
  }
  void unpin_locked(void* p) {
    const int i = find_pinned(p);
    if (i < 0) return;
    if (ctx().ready) cudaSetDevice(ctx().device);   // (never create a context on another device)
    if (cudaHostUnregister(p) != cudaSuccess) cudaGetLastError();
    pinned.erase(pinned.begin() + i);
  }
  void pin_locked(void* p, size_t cap) {
    if (!pin_mode() || find_pinned(p) >= 0 || !ctx().ready) return;
    cudaSetDevice(ctx().device);
    if (cudaHostRegister(p, cap, cudaHostRegisterPortable) == cudaSuccess)
      pinned.push_back(std::make_pair((char*)p, cap));
    else
      cudaGetLastError();   // (locked-memory limit, ...): the block stays an ordinary one
  }
  bool covers(const void* dst, size_t bytes) {
    std::lock_guard<std::mutex> lk(mu);
    const char* d = (const char*)dst;
    for (auto& b : pinned)
      if (d >= b.first && d + bytes <= b.first + b.second) { ++direct; return true; }
    return false;
  }
  // a block that leaves the library's custody for good
  void release(void* p) {
    { std::lock_guard<std::mutex> lk(mu); unpin_locked(p); }
    free(p);
  }

  size_t cap_limit() {
    if (!limit_known) {
      limit_known = true;
      const char* e = getenv("B200_HOST_CACHE_GB");
      if (e) {
        limit = (size_t)(atof(e) * (double)((size_t)1 << 30));
      } else {
        const long pages = sysconf(_SC_PHYS_PAGES), psz = sysconf(_SC_PAGE_SIZE);
        const size_t phys = pages > 0 && psz > 0 ? (size_t)pages * (size_t)psz : 0;
        limit = std::min(phys / 4, (size_t)64 << 30);
      }
    }
    return limit;
  }
  // smallest kept block that holds `bytes` without wasting more than 3/4 of itself
  void* take(size_t bytes) {
    if (bytes < BIG) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    int best = -1;
    for (int i = 0; i < (int)blocks.size(); ++i)
      if (blocks[i].first >= bytes && blocks[i].first / 4 <= bytes &&
          (best < 0 || blocks[i].first < blocks[best].first)) best = i;
    if (best < 0) { ++misses; return nullptr; }
    ++hits;
    void* p = blocks[best].second;
    total -= blocks[best].first;
    blocks.erase(blocks.begin() + best);
    return p;
  }
  // true: kept; false: the caller frees it
  bool give(void* p) {
    const size_t cap = malloc_usable_size(p);
    std::lock_guard<std::mutex> lk(mu);
    if (cap < BIG || cap > cap_limit()) return false;
    while (total + cap > limit && !blocks.empty()) {     // make room: smallest blocks go first
      int s = 0;
      for (int i = 1; i < (int)blocks.size(); ++i) if (blocks[i].first < blocks[s].first) s = i;
      if (blocks[s].first >= cap) return false;            // everything kept is at least as useful
      total -= blocks[s].first;
      unpin_locked(blocks[s].second);
      free(blocks[s].second);
      blocks.erase(blocks.begin() + s);
    }
    if (total + cap > limit) return false;
    blocks.push_back(std::make_pair(cap, p));
    total += cap;
    pin_locked(p, cap);
    return true;
  }
  void drop_all() {
    std::lock_guard<std::mutex> lk(mu);
    for (auto& b : blocks) { unpin_locked(b.second); free(b.second); }
    blocks.clear();
    total = 0;
  }
};
HostCache& host_cache() { static HostCache* h = new HostCache; return *h; }
bool host_block_is_pinned(const void* dst, size_t bytes) { return host_cache().covers(dst, bytes); }

void* host_block_alloc(size_t bytes) {
  void* p = host_cache().take(bytes);
  if (p) return p;
  p = malloc(bytes);
  advise_huge(p, bytes);
  return p;
}

// A row block on its way to the host: the three malloc()'d arrays plus the device range that
// still has to be copied into JC / C.
struct HostBlock {
  int* I = nullptr; int* J = nullptr; double* V = nullptr;
  int64_t begin = 0, cnt = 0;
  void drop() {   // (J and V may be blocks of the cache, possibly page-locked: never plain free())
    free(I);
    for (void* p : {(void*)J, (void*)V})
      if (p && !host_cache().give(p)) host_cache().release(p);
    I = nullptr; J = nullptr; V = nullptr;
  }
};

// allocate the host block and bring the (rebased, 32-bit) row offsets over; library stream
int download_prepare(const DevCSR& d, int lo, int hi, HostBlock* hb) {
  Ctx& c = ctx();
  if (lo < 0 || hi > d.rows || lo > hi) { set_error("row range out of bounds"); return B200_ERR_BAD_ARG; }
  const int m = hi - lo;
  int64_t ends[2] = {0, 0};
  B200_CUDA(cudaMemcpyAsync(&ends[0], d.rowptr + lo, sizeof(int64_t), cudaMemcpyDeviceToHost, c.stream));
  B200_CUDA(cudaMemcpyAsync(&ends[1], d.rowptr + hi, sizeof(int64_t), cudaMemcpyDeviceToHost, c.stream));
  B200_CUDA(cudaStreamSynchronize(c.stream));
  const int64_t cnt = ends[1] - ends[0];
  if (cnt > INT_MAX) {
    set_error("row block holds more than INT_MAX entries; download a smaller block");
    return B200_ERR_INT32_OVERFLOW;
  }
  hb->begin = ends[0]; hb->cnt = cnt;
  hb->I = (int*)malloc(((size_t)m + 1) * sizeof(int));
  hb->J = (int*)host_block_alloc(((size_t)cnt + 1) * sizeof(int));
  hb->V = (double*)host_block_alloc(((size_t)cnt + 1) * sizeof(double));
  if (!hb->I || !hb->J || !hb->V) { hb->drop(); set_error("host malloc failed"); return B200_ERR_HOST_ALLOC; }
  int* d32 = nullptr;
  cudaError_t e = dalloc(&d32, (size_t)m + 1);
  if (e != cudaSuccess) { hb->drop(); B200_CUDA(e); }
  k_i64_to_i32_rebased<<<(unsigned)((m + 1 + 255) / 256), 256, 0, c.stream>>>(d.rowptr + lo, d32, m + 1);
  e = cudaMemcpyAsync(hb->I, d32, ((size_t)m + 1) * sizeof(int), cudaMemcpyDeviceToHost, c.stream);
  dfree(d32);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
  if (e != cudaSuccess) { hb->drop(); B200_CUDA(e); }
  return B200_OK;
}

// the bulk of the block: columns and values through the pinned pipeline
int download_bulk(const DevCSR& d, const HostBlock& hb, bool ordered) {
  if (!hb.cnt) return B200_OK;
  int rc = d2h_staged(hb.J, d.col + hb.begin, (size_t)hb.cnt * sizeof(int), ordered);
  if (!rc) rc = d2h_staged(hb.V, d.val + hb.begin, (size_t)hb.cnt * sizeof(double), ordered);
  return rc;
}

int download_rows(const DevCSR& d, int lo, int hi, int** I, int** J, double** V, int* nnz) {
  HostBlock hb;
  int rc = download_prepare(d, lo, hi, &hb);
  if (rc) return rc;
  rc = download_bulk(d, hb, true);
  if (rc) { hb.drop(); return rc; }
  *I = hb.I; *J = hb.J; *V = hb.V; *nnz = (int)hb.cnt;
  return B200_OK;
}

// ---- download worker ----------------------------------------------------------------------------
// One helper thread that runs download_bulk for the streamed product while the calling thread
// computes the next row block.  It touches the copy stream, the pinned buffers and their events
// only; the library stream, the device pool and the error string stay with the calling thread.
// The thread is detached and the object is never destroyed: a joinable thread (or a condition
// variable with a waiter) inside a static object blocks the process at exit.
struct Downloader {
  std::mutex mu;
  std::condition_variable cv;
  bool started = false, has_job = false, busy = false, quit = false, exited = false;
  DevCSR src;
  HostBlock hb;
  int rc = B200_OK;
  std::string err;

  void loop(int device) {
    cudaSetDevice(device);
    std::unique_lock<std::mutex> lk(mu);
    for (;;) {
      cv.wait(lk, [&] { return has_job || quit; });
      if (quit) break;
      has_job = false;
      lk.unlock();
      const int r = download_bulk(src, hb, false);
      const std::string e = r ? std::string(b200_last_error()) : std::string();
      lk.lock();
      rc = r; err = e; busy = false;
      cv.notify_all();
    }
    exited = true;
    cv.notify_all();
  }
  // calling thread: the block's producer has finished on the library stream (run_pipeline ends
  // with a synchronisation); make the copy stream wait for it anyway, then hand the job over
  int submit(const DevCSR& d, const HostBlock& block) {
    Ctx& c = ctx();
    int r = ensure_staging();
    if (r) return r;
    B200_CUDA(cudaEventRecord(c.xfer_ev, c.stream));
    B200_CUDA(cudaStreamWaitEvent(c.copy_stream, c.xfer_ev, 0));
    std::lock_guard<std::mutex> lk(mu);
    if (!started) { std::thread(&Downloader::loop, this, c.device).detach(); started = true; }
    src = d; hb = block; has_job = true; busy = true;
    cv.notify_all();
    return B200_OK;
  }
  int wait() {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return !busy; });
    if (rc) set_error(err);
    return rc;
  }
  // b200_finalize: the worker leaves before the copy stream and the pinned buffers go away
  void stop() {
    std::unique_lock<std::mutex> lk(mu);
    if (!started) return;
    cv.wait(lk, [&] { return !busy; });
    quit = true;
    cv.notify_all();
    cv.wait(lk, [&] { return exited; });
    started = quit = exited = has_job = false;
  }
};
Downloader& downloader() { static Downloader* d = new Downloader; return *d; }

int upload(const int* I, const int* J, const double* V, int rows, int cols, int nnz, DevCSR* out) {
  Ctx& c = ctx();
  if (rows < 0 || cols < 0 || nnz < 0 || !I || (nnz && (!J || !V))) {
    set_error("bad CSR arguments");
    return B200_ERR_BAD_ARG;
  }
  DevCSR d;
  d.rows = rows; d.cols = cols; d.nnz = nnz;
  Temps T;  // every array is freed on an error path; the three result arrays are kept at the end
  int* tmp = nullptr;
  B200_CUDA(T.alloc(&tmp, (size_t)rows + 1));
  B200_CUDA(T.alloc(&d.rowptr, (size_t)rows + 1));
  B200_CUDA(T.alloc(&d.col, (size_t)nnz));
  B200_CUDA(T.alloc(&d.val, (size_t)nnz));
  B200_CUDA(cudaMemcpyAsync(tmp, I, ((size_t)rows + 1) * sizeof(int), cudaMemcpyHostToDevice, c.stream));
  k_i32_to_i64<<<(unsigned)((rows + 1 + 255) / 256), 256, 0, c.stream>>>(tmp, d.rowptr, rows + 1);
  B200_CUDA(cudaStreamSynchronize(c.stream));
  if (nnz) {
    int rc = h2d_staged(d.col, J, (size_t)nnz * sizeof(int));
    if (!rc) rc = h2d_staged(d.val, V, (size_t)nnz * sizeof(double));
    if (rc) return rc;
  }
  // sorted rows? and: is it a CSR at all (a bad index must not become an out-of-bounds read)
  const int rc = check_sorted_device(&d, true);
  if (rc) return rc;
  T.keep(d.rowptr); T.keep(d.col); T.keep(d.val);
  *out = d;
  return B200_OK;
}

void release(DevCSR& d) {
  dfree(d.rowptr); dfree(d.col); dfree(d.val);
  d = DevCSR();
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

const char* b200_last_error(void) { return g_err.c_str(); }

void b200_host_free(void* p) {
  if (p && !host_cache().give(p)) host_cache().release(p);
}

int b200_host_cache_stats(long long* hits, long long* misses, long long* direct, int* pinned_blocks) {
  HostCache& h = host_cache();
  std::lock_guard<std::mutex> lk(h.mu);
  if (hits) *hits = h.hits;
  if (misses) *misses = h.misses;
  if (direct) *direct = h.direct;
  if (pinned_blocks) *pinned_blocks = (int)h.pinned.size();
  return B200_OK;
}

int b200_host_cache_pin(int on) {
  HostCache& h = host_cache();
  std::lock_guard<std::mutex> lk(h.mu);
  h.pin = on ? 1 : 0;
  return B200_OK;
}

int b200_host_cache_info(long long* bytes, int* blocks, long long* limit_bytes) {
  HostCache& h = host_cache();
  std::lock_guard<std::mutex> lk(h.mu);
  if (bytes) *bytes = (long long)h.total;
  if (blocks) *blocks = (int)h.blocks.size();
  if (limit_bytes) *limit_bytes = (long long)h.cap_limit();
  return B200_OK;
}

int b200_host_cache_drop(void) {
  host_cache().drop_all();
  return B200_OK;
}

int b200_init(int device) {
  Ctx& c = ctx();
  if (c.ready) {
    if (c.device == device) return B200_OK;
    b200_finalize();
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error(std::string("no CUDA device available (") + cudaGetErrorString(e) +
              "); this library has no CPU fallback");
    return B200_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= ndev) { set_error("device index out of range"); return B200_ERR_BAD_ARG; }
  B200_CUDA(cudaSetDevice(device));
  cudaDeviceProp p;
  B200_CUDA(cudaGetDeviceProperties(&p, device));
  c.device = device;
  c.sm_count = p.multiProcessorCount;
  c.smem_optin = p.sharedMemPerBlockOptin;
  c.hbm_bytes = p.totalGlobalMem;
  strncpy(c.name, p.name, sizeof(c.name) - 1);
  B200_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  for (auto& ev : c.ev) B200_CUDA(cudaEventCreate(&ev));
  for (auto& ev : c.kev) B200_CUDA(cudaEventCreate(&ev));
  // keep freed blocks in the stream-ordered pool: per-iteration cudaMalloc was a cost centre of
  // the reference's GPU loop (nlibs/gpus/gpu_csr_kernel.cu:258-259)
  cudaMemPool_t pool;
  B200_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  unsigned long long thr = ~0ull;
  B200_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
  load_tunables(&c.tun);
  c.ready = true;
  return B200_OK;
}

int b200_set_topk(int k) {
  if (k < 0) { set_error("top-k must be >= 0 (0 = off)"); return B200_ERR_BAD_ARG; }
  ctx().topk = k;
  return B200_OK;
}

int b200_options_reload(void) {
  load_tunables(&ctx().tun);
  return B200_OK;
}

int b200_finalize(void) {
  Ctx& c = ctx();
  if (!c.ready) return B200_OK;
  downloader().stop();
  host_cache().drop_all();
  cudaStreamSynchronize(c.stream);
  for (auto& ev : c.ev) { cudaEventDestroy(ev); ev = nullptr; }
  for (auto& ev : c.kev) { cudaEventDestroy(ev); ev = nullptr; }
  if (c.bm_store) { cudaFree(c.bm_store); c.bm_store = nullptr; c.bm_store_words = 0; }
  c.bm_store_capped = false;
  if (c.rb_host) { cudaFreeHost(c.rb_host); c.rb_host = nullptr; c.rb_dev = nullptr; }
  rb_reset();
  if (c.arena_col) { cudaFree(c.arena_col); cudaFree(c.arena_val); c.arena_col = nullptr; c.arena_val = nullptr; c.arena_cap = 0; }
  cudaStreamDestroy(c.stream);
  c.stream = nullptr;
  if (c.pin[0]) {
    for (int k = 0; k < 2; ++k) { cudaFreeHost(c.pin[k]); c.pin[k] = nullptr; cudaEventDestroy(c.pin_ev[k]); }
    cudaEventDestroy(c.xfer_ev);
    cudaStreamDestroy(c.copy_stream);
    c.copy_stream = nullptr;
  }
  c.ready = false;
  return B200_OK;
}

int b200_stream(void** stream) {
  B200_REQUIRE_INIT();
  if (!stream) { set_error("null argument"); return B200_ERR_BAD_ARG; }
  *stream = (void*)ctx().stream;
  return B200_OK;
}

int b200_device_info(int* sm_count, long long* hbm_bytes, char* name, int name_len) {
  B200_REQUIRE_INIT();
  Ctx& c = ctx();
  if (sm_count) *sm_count = c.sm_count;
  if (hbm_bytes) *hbm_bytes = (long long)c.hbm_bytes;
  if (name && name_len > 0) { strncpy(name, c.name, name_len - 1); name[name_len - 1] = 0; }
  return B200_OK;
}

// ---- device container ----------------------------------------------------------------------

int b200_csr_upload(const int* I, const int* J, const double* V, int rows, int cols, int nnz,
                    b200_csr_t* out) {
  B200_REQUIRE_INIT();
  if (!out) { set_error("null output handle"); return B200_ERR_BAD_ARG; }
  b200_csr* h = new b200_csr();
  int rc = upload(I, J, V, rows, cols, nnz, &h->d);
  if (rc) { delete h; return rc; }
  *out = h;
  return B200_OK;
}

int b200_coo_to_csr(const int* rowIndex, const int* colIndex, const double* val, long long nnz,
                    int rows, int cols, int flags, b200_csr_t* out) {
  B200_REQUIRE_INIT();
  if (!out || rows < 0 || cols < 0 || nnz < 0 || (nnz && (!rowIndex || !colIndex)) || (flags & ~7)) {
    set_error("bad COO arguments");
    return B200_ERR_BAD_ARG;
  }
  Temps T;
  int *d_row = nullptr, *d_col = nullptr;
  double* d_val = nullptr;
  B200_CUDA(T.alloc(&d_row, (size_t)nnz));
  B200_CUDA(T.alloc(&d_col, (size_t)nnz));
  int rc = h2d_staged(d_row, rowIndex, (size_t)nnz * sizeof(int));
  if (!rc) rc = h2d_staged(d_col, colIndex, (size_t)nnz * sizeof(int));
  if (!rc && val && !(flags & B200_COO_NORMALISE)) {   // normalised values do not depend on the input's
    B200_CUDA(T.alloc(&d_val, (size_t)nnz));
    rc = h2d_staged(d_val, val, (size_t)nnz * sizeof(double));
  }
  if (rc) return rc;
  b200_csr* h = new b200_csr();
  rc = coo_build_device(d_row, d_col, d_val, nnz, rows, cols, flags, &h->d);
  if (rc) { delete h; return rc; }
  *out = h;
  return B200_OK;
}

int b200_csr_info(b200_csr_t h, int* rows, int* cols, long long* nnz) {
  if (!h) { set_error("null handle"); return B200_ERR_BAD_ARG; }
  if (rows) *rows = h->d.rows;
  if (cols) *cols = h->d.cols;
  if (nnz) *nnz = h->d.nnz;
  return B200_OK;
}

int b200_csr_download(b200_csr_t h, int** I, int** J, double** V, int* nnz) {
  B200_REQUIRE_INIT();
  if (!h || !I || !J || !V || !nnz) { set_error("null argument"); return B200_ERR_BAD_ARG; }
  if (h->d.nnz > INT_MAX) {
    set_error("nnz exceeds INT_MAX: the reference's int CSR cannot hold it; use b200_csr_download_rows");
    return B200_ERR_INT32_OVERFLOW;
  }
  return download_rows(h->d, 0, h->d.rows, I, J, V, nnz);
}

int b200_csr_download_rows(b200_csr_t h, int row_lo, int row_hi, int** I, int** J, double** V,
                           int* nnz) {
  B200_REQUIRE_INIT();
  if (!h || !I || !J || !V || !nnz) { set_error("null argument"); return B200_ERR_BAD_ARG; }
  return download_rows(h->d, row_lo, row_hi, I, J, V, nnz);
}

int b200_csr_free(b200_csr_t h) {
  if (!h) return B200_OK;
  if (ctx().ready) { release(h->d); cudaStreamSynchronize(ctx().stream); }
  delete h;
  return B200_OK;
}

int b200_csr_sort_rows(b200_csr_t h) {
  B200_REQUIRE_INIT();
  if (!h) { set_error("null handle"); return B200_ERR_BAD_ARG; }
  return sort_rows_device(&h->d);
}

int b200_csr_device_ptrs(b200_csr_t h, void** rowptr64, void** colind32, void** values64) {
  if (!h) { set_error("null handle"); return B200_ERR_BAD_ARG; }
  if (rowptr64) *rowptr64 = h->d.rowptr;
  if (colind32) *colind32 = h->d.col;
  if (values64) *values64 = h->d.val;
  return B200_OK;
}

// ---- device entry points -------------------------------------------------------------------

static int check_mul(b200_csr_t A, b200_csr_t B, int lo, int hi) {
  if (!A || !B) { set_error("null handle"); return B200_ERR_BAD_ARG; }
  if (A->d.cols != B->d.rows) {  // assert(cols == B.rows), nlibs/CSR.cc:183
    set_error("dimension mismatch: A.cols != B.rows");
    return B200_ERR_BAD_ARG;
  }
  if (lo < 0 || hi > A->d.rows || lo > hi) { set_error("row range out of bounds"); return B200_ERR_BAD_ARG; }
  return B200_OK;
}

int b200_spgemm_device_rows(b200_csr_t A, b200_csr_t B, int row_lo, int row_hi, b200_csr_t* C,
                            b200_stats* stats) {
  B200_REQUIRE_INIT();
  int rc = check_mul(A, B, row_lo, row_hi);
  if (rc) return rc;
  if (!C) { set_error("null output handle"); return B200_ERR_BAD_ARG; }
  b200_csr* h = new b200_csr();
  rc = run_pipeline(A->d, B->d, row_lo, row_hi, MODE_SPGEMM, &h->d, nullptr, stats);
  if (rc) { release(h->d); delete h; return rc; }
  *C = h;
  return B200_OK;
}

int b200_spgemm_device(b200_csr_t A, b200_csr_t B, b200_csr_t* C, b200_stats* stats) {
  if (!A) { set_error("null handle"); return B200_ERR_BAD_ARG; }
  return b200_spgemm_device_rows(A, B, 0, A->d.rows, C, stats);
}

// Row blocks of A x B, block b on its way over PCIe while block b+1 is computed.
static int stream_blocks(const DevCSR& A, const DevCSR& B, long long block_products,
                         b200_block_fn fn, void* user) {
  Ctx& c = ctx();
  const int m = A.rows;
  if (block_products <= 0) block_products = 2000000000LL;
  std::vector<long long> prefix((size_t)m + 1, 0);
  {
    int64_t* d_prefix = nullptr;
    B200_CUDA(dalloc(&d_prefix, (size_t)m + 1));
    int rc = flops_prefix_device(A, B, 0, m, d_prefix);
    cudaError_t e = cudaSuccess;
    if (!rc) e = cudaMemcpyAsync(prefix.data(), d_prefix, ((size_t)m + 1) * sizeof(long long),
                                 cudaMemcpyDeviceToHost, c.stream);
    dfree(d_prefix);
    if (rc) return rc;
    B200_CUDA(e);
    B200_CUDA(cudaStreamSynchronize(c.stream));
  }
  Downloader& dl = downloader();
  const bool prof = ctx().tun.prof;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  struct Flight { bool on = false; DevCSR dev; HostBlock hb; int lo = 0, hi = 0; } fly, done;
  // wait for the block in flight and free its device copy; it becomes `done`
  auto land = [&]() -> int {
    if (!fly.on) return B200_OK;
    const int rc = dl.wait();
    release(fly.dev);
    if (rc) { fly.hb.drop(); fly.on = false; return rc; }
    done = fly; fly.on = false;
    return B200_OK;
  };
  // pass the landed block on; the callback owns the arrays from here
  auto deliver = [&]() -> int {
    if (!done.on) return B200_OK;
    done.on = false;
    if (fn(user, done.lo, done.hi, done.hb.I, done.hb.J, done.hb.V, (int)done.hb.cnt)) {
      set_error("block callback asked to stop");
      return B200_ERR_CALLBACK;
    }
    return B200_OK;
  };
  int lo = 0;
  do {
    // largest hi with prefix[hi] - prefix[lo] <= block_products, at least one row
    int hi = (int)(std::upper_bound(prefix.begin() + lo, prefix.end(), prefix[lo] + block_products) -
                   prefix.begin()) - 1;
    hi = std::min(m, std::max(hi, lo + 1));
    DevCSR dC;
    HostBlock hb;
    const double t0 = now();
    b200_stats st;
    memset(&st, 0, sizeof st);
    int rc = run_pipeline(A, B, lo, hi, MODE_SPGEMM, &dC, nullptr, prof ? &st : nullptr);
    const double t1 = now();
    // block b-1 has had the whole computation of block b to arrive.  Only when it has landed
    // are block b's row offsets fetched (their copy shares the device-to-host copy engine with
    // the download and would wait for it anyway), and block b's transfer starts before the
    // callback sees block b-1, so that the copy engine never waits for the callback.
    const int rc_land = land();
    const double t2 = now();
    if (!rc) rc = rc_land;
    if (!rc) rc = download_prepare(dC, 0, hi - lo, &hb);
    const double t3 = now();
    if (!rc) rc = dl.submit(dC, hb);
    if (rc) {
      release(dC); hb.drop();
      if (done.on) { done.hb.drop(); done.on = false; }
      return rc;
    }
    fly.on = true; fly.dev = dC; fly.hb = hb; fly.lo = lo; fly.hi = hi;
    rc = deliver();
    if (prof) {
      double kern = 0.0;
      for (int b = 0; b < 16; ++b) kern += st.ms_sym_bin[b] + st.ms_num_bin[b];
      fprintf(stderr, "[b200 stream] on-stream %.1f ms (flops %.1f symbolic %.1f numeric %.1f other %.1f; kernels alone %.1f)\n",
              st.ms_total, st.ms_flops, st.ms_symbolic, st.ms_numeric, st.ms_other, kern);
    }
    if (prof)
      fprintf(stderr, "[b200 stream] rows %d-%d nnz %lld: compute %.1f wait-for-previous %.1f prepare %.1f callback %.1f ms; host cache %ld hits %ld misses\n",
              lo, hi, (long long)hb.cnt, t1 - t0, t2 - t1, t3 - t2, now() - t3, host_cache().hits, host_cache().misses);
    if (rc) { land(); if (done.on) { done.hb.drop(); done.on = false; } return rc; }
    lo = hi;
  } while (lo < m);
  int rc = land();
  if (!rc) rc = deliver();
  return rc;
}

int b200_spgemm_device_stream(b200_csr_t A, b200_csr_t B, long long block_products,
                              b200_block_fn fn, void* user) {
  B200_REQUIRE_INIT();
  int rc = check_mul(A, B, 0, A ? A->d.rows : 0);
  if (rc) return rc;
  if (!fn) { set_error("null block callback"); return B200_ERR_BAD_ARG; }
  return stream_blocks(A->d, B->d, block_products, fn, user);
}

int b200_rmcl_step_device_rows(b200_csr_t Mgt, b200_csr_t Mt, int row_lo, int row_hi,
                               b200_csr_t* newMt, double* chaos, b200_stats* stats) {
  B200_REQUIRE_INIT();
  int rc = check_mul(Mgt, Mt, row_lo, row_hi);
  if (rc) return rc;
  if (!newMt) { set_error("null output handle"); return B200_ERR_BAD_ARG; }
  b200_csr* h = new b200_csr();
  double ch = 0.0;
  rc = rmcl_step_device(Mgt->d, Mt->d, row_lo, row_hi, &h->d, &ch, stats);
  if (rc) { release(h->d); delete h; return rc; }
  if (chaos) *chaos = ch;
  *newMt = h;
  return B200_OK;
}

int b200_rmcl_step_device(b200_csr_t Mgt, b200_csr_t Mt, b200_csr_t* newMt, double* chaos,
                          b200_stats* stats) {
  if (!Mgt) { set_error("null handle"); return B200_ERR_BAD_ARG; }
  return b200_rmcl_step_device_rows(Mgt, Mt, 0, Mgt->d.rows, newMt, chaos, stats);
}

int b200_flops_prefix(b200_csr_t A, b200_csr_t B, long long* prefix) {
  return b200_cost_prefix(A, B, 0, prefix);
}

int b200_cost_prefix(b200_csr_t A, b200_csr_t B, long long row_charge, long long* prefix) {
  B200_REQUIRE_INIT();
  int rc = check_mul(A, B, 0, A ? A->d.rows : 0);
  if (rc) return rc;
  if (!prefix) { set_error("null output"); return B200_ERR_BAD_ARG; }
  Ctx& c = ctx();
  const int m = A->d.rows;
  Temps T;
  int64_t* d_prefix = nullptr;
  B200_CUDA(T.alloc(&d_prefix, (size_t)m + 1));
  rc = flops_prefix_device(A->d, B->d, 0, m, d_prefix, row_charge < 0 ? c.tun.row_charge : row_charge);
  if (rc) return rc;
  B200_CUDA(cudaMemcpyAsync(prefix, d_prefix, ((size_t)m + 1) * sizeof(long long),
                            cudaMemcpyDeviceToHost, c.stream));
  B200_CUDA(cudaStreamSynchronize(c.stream));
  return B200_OK;
}

// arrayEqualPartition64 (nlibs/tools/util.cc:123-135); host arithmetic only
int b200_equal_partition64(const long long* prefix, int n, int nparts, int* ends) {
  if (!prefix || !ends || n < 0 || nparts < 1) { set_error("bad argument"); return B200_ERR_BAD_ARG; }
  const long long total = prefix[n];
  const long long chunk = (total + nparts - 1) / nparts;
  ends[0] = 0;
  int now = 0;
  for (int t = 0; t + 1 < nparts; ++t) {
    const long long target = std::min((long long)(t + 1) * chunk, total);
    const long long* up = std::upper_bound(prefix + now, prefix + n + 1, target);
    int e = std::max((int)(up - prefix - 1), now + 1);
    e = std::min(e, n);
    ends[t + 1] = e;
    now = e;
  }
  ends[nparts] = n;
  return B200_OK;
}

int b200_csr_row_argmax(b200_csr_t h, int* labels) {
  B200_REQUIRE_INIT();
  if (!h || !labels) { set_error("null argument"); return B200_ERR_BAD_ARG; }
  Ctx& c = ctx();
  const int m = h->d.rows;
  int* d_lab = nullptr;
  B200_CUDA(dalloc(&d_lab, (size_t)m));
  if (m) k_row_argmax<<<(m + 255) / 256, 256, 0, c.stream>>>(h->d.rowptr, h->d.col, h->d.val, m, d_lab);
  B200_CUDA(cudaMemcpyAsync(labels, d_lab, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  dfree(d_lab);
  B200_CUDA(cudaStreamSynchronize(c.stream));
  return B200_OK;
}

int b200_csr_concat_rows(const b200_csr_t* blocks, int nblocks, b200_csr_t* out) {
  B200_REQUIRE_INIT();
  if (!blocks || nblocks < 1 || !out) { set_error("bad argument"); return B200_ERR_BAD_ARG; }
  long long rows = 0;
  std::vector<DevCSR> v;
  for (int b = 0; b < nblocks; ++b) {
    if (!blocks[b] || blocks[b]->d.cols != blocks[0]->d.cols) { set_error("blocks disagree on cols"); return B200_ERR_BAD_ARG; }
    rows += blocks[b]->d.rows;
    v.push_back(blocks[b]->d);
  }
  if (rows > INT_MAX) { set_error("too many rows"); return B200_ERR_INT32_OVERFLOW; }
  DevCSR d;
  const int rc = concat_rows_device(v, blocks[0]->d.cols, &d);   // frees its arrays on an error path
  if (rc) return rc;
  B200_CUDA(cudaStreamSynchronize(ctx().stream));
  b200_csr* h = new b200_csr();
  h->d = d;
  *out = h;
  return B200_OK;
}

int b200_csr_column_stripe(b200_csr_t B, int col_lo, int col_hi, b200_csr_t* out) {
  B200_REQUIRE_INIT();
  if (!B || !out || col_lo < 0 || col_hi < col_lo || col_hi > B->d.cols) {
    set_error("bad column stripe");
    return B200_ERR_BAD_ARG;
  }
  b200_csr* h = new b200_csr();
  const int rc = column_stripe_device(B->d, col_lo, col_hi, &h->d);
  if (rc) { delete h; return rc; }
  *out = h;
  return B200_OK;
}

int b200_csr_concat_cols(const b200_csr_t* blocks, int nblocks, b200_csr_t* out) {
  B200_REQUIRE_INIT();
  if (!blocks || nblocks < 1 || !out) { set_error("bad argument"); return B200_ERR_BAD_ARG; }
  std::vector<DevCSR> v;
  long long cols = 0;
  for (int b = 0; b < nblocks; ++b) {
    if (!blocks[b] || blocks[b]->d.rows != blocks[0]->d.rows) { set_error("stripes disagree on rows"); return B200_ERR_BAD_ARG; }
    cols += blocks[b]->d.cols;
    v.push_back(blocks[b]->d);
  }
  if (cols > INT_MAX) { set_error("too many columns"); return B200_ERR_INT32_OVERFLOW; }
  b200_csr* h = new b200_csr();
  const int rc = concat_cols_device(v, &h->d);
  if (rc) { delete h; return rc; }
  *out = h;
  return B200_OK;
}

// ---- host-buffer entry points ----------------------------------------------------------------

static int host_mul(Mode mode, const int* IA, const int* JA, const double* A, int nnzA,
                    const int* IB, const int* JB, const double* B, int nnzB, int** IC, int** JC,
                    double** C, int* nnzC, int m, int k, int n, double* chaos) {
  B200_REQUIRE_INIT();
  if (!IC || !JC || !C || !nnzC) { set_error("null output argument"); return B200_ERR_BAD_ARG; }
  DevCSR dA, dB, dC;
  int rc = upload(IA, JA, A, m, k, nnzA, &dA);
  if (rc) return rc;
  rc = upload(IB, JB, B, k, n, nnzB, &dB);
  if (rc) { release(dA); return rc; }
  rc = mode == MODE_RMCL ? rmcl_step_device(dA, dB, 0, m, &dC, chaos, nullptr)
                         : run_pipeline(dA, dB, 0, m, mode, &dC, chaos, nullptr);
  if (!rc) {
    if (dC.nnz > INT_MAX) {
      set_error("nnz(C) exceeds INT_MAX: use the row-block device API");
      rc = B200_ERR_INT32_OVERFLOW;
    } else {
      rc = download_rows(dC, 0, m, IC, JC, C, nnzC);
    }
  }
  release(dA); release(dB); release(dC);
  cudaStreamSynchronize(ctx().stream);
  return rc;
}

int b200_spgemm_csr(const int* IA, const int* JA, const double* A, int nnzA, const int* IB,
                    const int* JB, const double* B, int nnzB, int** IC, int** JC, double** C,
                    int* nnzC, int m, int k, int n) {
  return host_mul(MODE_SPGEMM, IA, JA, A, nnzA, IB, JB, B, nnzB, IC, JC, C, nnzC, m, k, n, nullptr);
}

int b200_spgemm_csr_stream(const int* IA, const int* JA, const double* A, int nnzA, const int* IB,
                           const int* JB, const double* B, int nnzB, int m, int k, int n,
                           long long block_products, b200_block_fn fn, void* user) {
  B200_REQUIRE_INIT();
  if (!fn) { set_error("null block callback"); return B200_ERR_BAD_ARG; }
  DevCSR dA, dB;
  int rc = upload(IA, JA, A, m, k, nnzA, &dA);
  if (rc) return rc;
  // A x A with the very same arrays (the headline case): one device copy serves both sides
  const bool same = IB == IA && JB == JA && B == A && nnzB == nnzA && k == m && n == k;
  if (!same) {
    rc = upload(IB, JB, B, k, n, nnzB, &dB);
    if (rc) { release(dA); return rc; }
  }
  rc = stream_blocks(dA, same ? dA : dB, block_products, fn, user);
  release(dA);
  if (!same) release(dB);
  cudaStreamSynchronize(ctx().stream);
  return rc;
}

int b200_rmcl_onestep_csr(const int* IA, const int* JA, const double* A, int nnzA, const int* IB,
                          const int* JB, const double* B, int nnzB, int** IC, int** JC, double** C,
                          int* nnzC, int m, int k, int n, double* chaos) {
  double ch = 0.0;
  int rc = host_mul(MODE_RMCL, IA, JA, A, nnzA, IB, JB, B, nnzB, IC, JC, C, nnzC, m, k, n, &ch);
  if (!rc && chaos) *chaos = ch;
  return rc;
}

int b200_rmcl_iter(int maxIter, double eps, const int* IG, const int* JG, const double* G,
                   int nnzG, const int* IT, const int* JT, const double* T, int nnzT, int** IM,
                   int** JM, double** M, int* nnzM, int n, int* iters_done, double* chaos_hist) {
  B200_REQUIRE_INIT();
  if (!IM || !JM || !M || !nnzM || maxIter < 0) { set_error("bad argument"); return B200_ERR_BAD_ARG; }
  DevCSR dG, dT;
  int rc = upload(IG, JG, G, n, n, nnzG, &dG);
  if (rc) return rc;
  rc = upload(IT, JT, T, n, n, nnzT, &dT);
  if (rc) { release(dG); return rc; }
  int it = 0;
  double prev_ch = 0.0;
  for (; it < maxIter; ++it) {
    DevCSR dN;
    double ch = 0.0;
    rc = rmcl_step_device(dG, dT, 0, n, &dN, &ch, nullptr);
    if (rc) { release(dN); break; }
    release(dT);  // Mt.dispose(); Mt = newMt  (nlibs/qrmcl.cc:72-73)
    dT = dN;
    if (chaos_hist) chaos_hist[it] = ch;
    if (rmcl_converged(ch, prev_ch, it, eps)) { ++it; break; }
    prev_ch = ch;
  }
  // Mt.makeOrdered() — what the reference's drivers do before comparing (nrmcl.cc:25-26)
  if (!rc) rc = sort_rows_device(&dT);
  if (!rc) {
    if (dT.nnz > INT_MAX) { set_error("nnz exceeds INT_MAX"); rc = B200_ERR_INT32_OVERFLOW; }
    else rc = download_rows(dT, 0, n, IM, JM, M, nnzM);
  }
  if (iters_done) *iters_done = it;
  release(dG); release(dT);
  cudaStreamSynchronize(ctx().stream);
  return rc;
}

}  // extern "C"
