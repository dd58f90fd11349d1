// common.cuh — shared declarations of the B200 SpGEMM / rMCL library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <chrono>
#include <vector>
#include "b200_spgemm.h"

namespace b200 {

// Device-resident CSR (replaces the device mirror of struct CSR, nlibs/CSR.cc:342-379).
// Layout in HBM: rowptr int64[rows+1] (64-bit so that nnz(C) > 2^31 fits, SURVEY.md §7 hard
// part 1), col int32[nnz], val fp64[nnz]; three separate allocations from the stream-ordered
// pool so row blocks can be viewed without copies.
struct DevCSR {
  int64_t* rowptr = nullptr;
  int* col = nullptr;
  double* val = nullptr;
  int rows = 0, cols = 0;
  int64_t nnz = 0;
  // every row has strictly ascending columns (checked at upload, true by construction for an
  // SpGEMM result, false for an rMCL step in first-touch order): lets the numeric pass cut a B
  // row at a column boundary with a binary search
  bool sorted_rows = false;
};

}  // namespace b200
// the opaque handle of include/b200_spgemm.h
struct b200_csr {
  b200::DevCSR d;
};
namespace b200 {

// Developer switches.  Read from the environment ONCE, by b200_init and by
// b200_options_reload — never on the per-call path.
struct Tunables {
  bool force_wide = false;   // B200_FORCE_WIDE: treat B as too wide for any shared-memory bitmap
  bool parts4 = false;       // B200_PARTS4: 4 x 256-thread CTAs / SM geometry of the part kernel
  bool no_parts = false;     // B200_NO_PARTS: whole-row numeric bitmap kernel
  bool on_chip = false;      // B200_ON_CHIP: heavy rows through the range-item kernel (ranges.cuh);
                             // implied by B200_DETERMINISTIC.  Default off: measured slower than the
                             // part kernel (global RED) on R-MAT scale 20, see DESIGN.md §3b
  int esc = -1;              // B200_ESC=0/1: rMCL rows of 1024..8192 products sorted on chip (esc.cuh) never /
                             // always; default (-1): in matrices wider than two column parts (> 1 M columns)
  int fuse = -1;             // B200_FUSE=0/1: plain SpGEMM rows of 512..2048 products are first tried as 128-entry
                             // numeric rows in the symbolic phase (k_num_warp_fused) never / always; default: always
                             // (2: always, with the 64-bit B offsets of operands beyond 2^31 entries — testing)
  bool prof = false;         // B200_PROF: per-phase diagnostics on stderr
  int l2[5] = {2, 1, 0, 1, 0};  // B200_L2POL=acc,ocol,bgather,bmstore,demote (L2Prio values)
  long long sym_big_from = -1;  // B200_SYM_BIG_FROM, B200_NUM_BIG_FROM, B200_LIGHT_P: bin cuts (-1: default)
  int num_big_from = -1;
  long long light_p = -1;
  long long team_products = 98304;  // B200_TEAM_P, B200_TEAM_MAX: team sizing of the part kernel
  int team_max = 64;
  long long row_charge = 32768;  // B200_ROW_CHARGE: products-equivalent cost of a heavy row in the
                                 // partition of an rMCL step among GPUs (0: equal products, as the reference)
  long long arena_entries = 0;  // B200_ARENA_ENTRIES: products per row tile of an rMCL step (0: from free memory)
  int ranges = 128;     // B200_RANGES: static column ranges of the on-chip numeric pass (2..128)
  bool deterministic = false;  // B200_DETERMINISTIC: B200_ON_CHIP (heavy rows accumulate on chip in
                               // A-entry order: bit-identical to the reference, reproducible)
};
void load_tunables(Tunables* t);

struct Ctx {
  bool ready = false;
  Tunables tun;
  int device = -1;
  int sm_count = 0;
  size_t smem_optin = 0;
  size_t hbm_bytes = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[8] = {};
  cudaEvent_t kev[64] = {};  // per-kernel brackets: [2*b], [2*b+1] symbolic bin b; 32+ numeric
  char name[128] = {0};
  // host <-> device staging for the host-buffer entry points (capi.cu): two pinned buffers on a
  // second stream, so the DMA of chunk k overlaps the host-side copy of chunk k-1
  cudaStream_t copy_stream = nullptr;
  void* pin[2] = {nullptr, nullptr};
  cudaEvent_t pin_ev[2] = {};
  cudaEvent_t xfer_ev = nullptr;
  // bitmap store of the symbolic pass, kept between calls (re-allocating tens of GB from the
  // stream-ordered pool every call costs ~10 ms of remapping)
  unsigned long long* bm_store = nullptr;
  size_t bm_store_words = 0;
  bool bm_store_capped = false;  // the last sizing was limited by the memory budget: do not retry
  // arena of the fused rMCL step (unpruned rows before compaction), grow-only between calls:
  // its size changes every iteration and re-carving tens of GB out of the stream-ordered pool
  // stalled single iterations for up to seconds
  int* arena_col = nullptr;
  double* arena_val = nullptr;
  size_t arena_cap = 0;
  int topk = 0;                // rMCL: keep at most this many entries per row (b200_set_topk; 0 = off)
  long long arena_budget = 0;  // entries a row tile of an rMCL step may hold (tiles.cu); 0: not sized yet
  // small read-backs (row totals, bin counts): a kernel writes them into mapped pinned memory
  // instead of a cudaMemcpy, because the device-to-host copy engine may be busy for 100+ ms with
  // a streamed download (profiles/r1_stream_blocks_rmat20.txt)
  static constexpr int RB_BYTES = 8192, RB_MAX = 16;
  unsigned char* rb_host = nullptr;
  unsigned char* rb_dev = nullptr;
  size_t rb_used = 0;
  int rb_n = 0;
  struct { void* dst; size_t off, bytes; } rb_pending[RB_MAX];
};
Ctx& ctx();

// Queue a small device -> host read-back on `st`; the value is in *dst after sync_fetch(st).
// dst must stay valid until then; rb_reset() forgets queued read-backs (entry points call it, so
// that an error return between the two cannot leave a pointer to a dead stack frame behind).
cudaError_t d2h_small(void* dst, const void* src, size_t bytes, cudaStream_t st);
cudaError_t sync_fetch(cudaStream_t st);
void rb_reset();

void set_error(const std::string& s);
int fail_cuda(cudaError_t e, const char* what, const char* file, int line);

#define B200_CUDA(call)                                                        \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) return b200::fail_cuda(e__, #call, __FILE__, __LINE__); \
  } while (0)

#define B200_REQUIRE_INIT()                                                     \
  do {                                                                          \
    if (!b200::ctx().ready) {                                                   \
      b200::set_error("b200_init() has not been called (no CPU fallback exists)"); \
      return B200_ERR_NOT_INIT;                                                 \
    }                                                                           \
  } while (0)

// stream-ordered allocation helpers
template <typename T>
inline cudaError_t dalloc(T** p, size_t count) {
  if (count == 0) count = 1;
  if (!ctx().tun.prof) return cudaMallocAsync((void**)p, count * sizeof(T), ctx().stream);
  // B200_PROF: report allocations that stall the host (the pool had to get memory from the driver)
  const auto t0 = std::chrono::steady_clock::now();
  const cudaError_t e = cudaMallocAsync((void**)p, count * sizeof(T), ctx().stream);
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (ms > 1.0) fprintf(stderr, "[b200 prof] cudaMallocAsync of %.1f MB took %.1f ms\n", count * sizeof(T) / 1e6, ms);
  return e;
}
template <typename T>
inline void dfree(T* p) {
  if (p) cudaFreeAsync((void*)p, ctx().stream);
}

// Stopping rule of the rMCL loops (not in the reference, whose loop is fixed-count,
// nlibs/qrmcl.cc:31): with eps > 0 stop after the first iteration whose chaos is < eps or moved
// by less than eps since the previous one (an rMCL fixed point keeps fractional rows, so its
// chaos settles at a non-zero value).  oracle/oracle.c states the same rule.
inline bool rmcl_converged(double ch, double prev, int it, double eps) {
  if (!(eps > 0)) return false;
  const double d = ch > prev ? ch - prev : prev - ch;
  return ch < eps || (it > 0 && d < eps);
}

// Stream-ordered temporaries of one call: freed at scope exit, on the error paths too.  An
// array that becomes part of the result is taken out with keep().
struct Temps {
  std::vector<void*> v;
  template <typename T>
  cudaError_t alloc(T** p, size_t count) {
    const cudaError_t e = dalloc(p, count);
    if (e == cudaSuccess) v.push_back((void*)*p);
    return e;
  }
  void adopt(void* p) { if (p) v.push_back(p); }
  void keep(const void* p) {
    for (size_t k = 0; k < v.size(); ++k) if (v[k] == p) { v[k] = v.back(); v.pop_back(); return; }
  }
  ~Temps() { for (void* p : v) cudaFreeAsync(p, ctx().stream); }
};

// mode of the row pipeline
enum Mode { MODE_SPGEMM = 0, MODE_RMCL = 1 };

// COO -> CSR on the device (ingest.cu); flags: 1 drop repeated pairs, 2 add missing self loops,
// 4 values = 1 / rowcount
int coo_build_device(const int* d_row, const int* d_col, const double* d_val, long long nnz, int rows,
                     int cols, int flags, DevCSR* out);

// Core pipeline (spgemm.cu): C = A[row_lo:row_hi) x B, or the fused rMCL step.
// `Bsorted`: a column-sorted copy of B the caller already has (used when the bitmap kernels need
// one and B's rows are unsorted; nullptr: made on demand, per call).
int run_pipeline(const DevCSR& A, const DevCSR& B, int row_lo, int row_hi, Mode mode,
                 DevCSR* C, double* chaos, b200_stats* stats, const DevCSR* Bsorted = nullptr);
// One rMCL step newMt[row_lo:row_hi) through a bounded arena (tiles.cu): the entry point every
// rMCL caller uses.
int rmcl_step_device(const DevCSR& A, const DevCSR& B, int row_lo, int row_hi, DevCSR* C,
                     double* chaos, b200_stats* stats);
// (col, val) of `d` sorted by column inside every row; shares d's row offsets (spgemm.cu)
int sorted_copy_of(const DevCSR& d, DevCSR* out);
// concatenation of row blocks (tiles.cu); arrayEqualPartition64 on a device prefix (tiles.cu)
int concat_rows_device(const std::vector<DevCSR>& blocks, int cols, DevCSR* out);
int equal_partition_device(const int64_t* d_prefix, int n, int nparts, int* h_ends);
// column stripe B[:, lo:hi) with local column indices, and the row-wise re-assembly of stripes
// (stripes.cu: the device PCSR, nlibs/PCSR.cc:3-56)
int column_stripe_device(const DevCSR& B, int lo, int hi, DevCSR* out);
int concat_cols_device(const std::vector<DevCSR>& blocks, DevCSR* out);

// CSR::makeOrdered on the device (spgemm.cu)
int sort_rows_device(DevCSR* d);
// sets d->sorted_rows (spgemm.cu)
int check_sorted_device(DevCSR* d, bool validate = false);

// flops prefix on device (spgemm.cu)
// row_charge > 0: prefix of the partition COST (products + row_charge per heavy row) instead;
// d_products (device, may be null) then still receives the plain product total
int flops_prefix_device(const DevCSR& A, const DevCSR& B, int row_lo, int row_hi,
                        int64_t* d_prefix /* rows+1 */, long long row_charge = 0,
                        long long* d_products = nullptr);

}  // namespace b200
