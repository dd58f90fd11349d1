// comm.cu — multi-GPU rMCL: one process per GPU, NCCL over NVLink 5 / NVSwitch.
//
// Not in the reference (single process, single GPU; SURVEY.md §2.1 strategy table).  Sharding
// follows SURVEY.md §8(e): the reference computes newMt = Mgt x Mt (nlibs/qrmcl.cc:39 ->
// nlibs/CSR.cc:265-276), so A = Mgt is the row-partitioned operand and B = Mt the gathered
// one.  Every rank holds Mgt and the current Mt in full; each iteration
//   1. cuts the rows into nranks contiguous blocks of equal intermediate products
//      (arrayEqualPartition64 on the flops prefix, nlibs/tools/util.cc:123-135 with
//      nthreads -> nranks),
//   2. runs the fused rMCL row pipeline on its block (no collective on the data path),
//   3. all-gathers the PRUNED row blocks (ragged: one ncclBroadcast per rank inside a group)
//      and max-all-reduces the chaos scalar.
// A row's result does not depend on which rank computed it.  Rows of the hash bins are bitwise
// the same for every GPU count; rows of the bitmap bin are accumulated with fp64 RED in an order
// that is not fixed, so they agree to rounding (<= 1e-12 relative) and are not reproducible run
// to run — unless the library runs with B200_DETERMINISTIC=1 (ordered on-chip accumulation).
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>
#include <string>
#include <algorithm>
#include <vector>
#include "common.cuh"

namespace b200 {
namespace {
ncclComm_t g_comm = nullptr;
int g_rank = 0, g_nranks = 1;

// NCCL is bound at first use with dlopen("libnccl.so.2") instead of at link time: the host
// program may already have loaded a (newer) NCCL — torch bundles its own — and two different
// libnccl.so.2 in one process do not mix.  nccl.h supplies the types only.
struct NcclApi {
  void* handle = nullptr;
  decltype(&::ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&::ncclCommInitRank) CommInitRank = nullptr;
  decltype(&::ncclCommDestroy) CommDestroy = nullptr;
  decltype(&::ncclAllGather) AllGather = nullptr;
  decltype(&::ncclAllReduce) AllReduce = nullptr;
  decltype(&::ncclBroadcast) Broadcast = nullptr;
  decltype(&::ncclGroupStart) GroupStart = nullptr;
  decltype(&::ncclGroupEnd) GroupEnd = nullptr;
  decltype(&::ncclGetErrorString) GetErrorString = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.handle) return B200_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
  if (!h) { set_error(std::string("cannot load libnccl.so.2: ") + dlerror()); return B200_ERR_NCCL; }
#define B200_SYM(field, name)                                                     \
  g_nccl.field = (decltype(g_nccl.field))dlsym(h, name);                          \
  if (!g_nccl.field) { set_error("libnccl lacks " name); dlclose(h); return B200_ERR_NCCL; }
  B200_SYM(GetUniqueId, "ncclGetUniqueId")
  B200_SYM(CommInitRank, "ncclCommInitRank")
  B200_SYM(CommDestroy, "ncclCommDestroy")
  B200_SYM(AllGather, "ncclAllGather")
  B200_SYM(AllReduce, "ncclAllReduce")
  B200_SYM(Broadcast, "ncclBroadcast")
  B200_SYM(GroupStart, "ncclGroupStart")
  B200_SYM(GroupEnd, "ncclGroupEnd")
  B200_SYM(GetErrorString, "ncclGetErrorString")
#undef B200_SYM
  g_nccl.handle = h;
  return B200_OK;
}

int fail_nccl(ncclResult_t r, const char* what) {
  set_error(std::string("NCCL error in ") + what + ": " + g_nccl.GetErrorString(r));
  return B200_ERR_NCCL;
}
#define B200_NCCL(call)                                  \
  do {                                                   \
    ncclResult_t r__ = (g_nccl.call);                    \
    if (r__ != ncclSuccess) return fail_nccl(r__, #call); \
  } while (0)

__global__ void k_shift_rowptr(int64_t* rp, long long n, long long add) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) rp[i] += add;
}
}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

int b200_comm_unique_id(char id[128]) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId u;
  int lrc = load_nccl();
  if (lrc) return lrc;
  B200_NCCL(GetUniqueId(&u));
  memcpy(id, &u, 128);
  return B200_OK;
}

int b200_comm_init(int rank, int nranks, const char id[128]) {
  B200_REQUIRE_INIT();
  if (g_comm) b200_comm_destroy();
  int lrc = load_nccl();
  if (lrc) return lrc;
  ncclUniqueId u;
  memcpy(&u, id, 128);
  B200_NCCL(CommInitRank(&g_comm, nranks, u, rank));
  g_rank = rank;
  g_nranks = nranks;
  return B200_OK;
}

int b200_comm_destroy(void) {
  if (g_comm && g_nccl.CommDestroy) { g_nccl.CommDestroy(g_comm); g_comm = nullptr; }
  g_rank = 0; g_nranks = 1;
  return B200_OK;
}

int b200_rmcl_iter_sharded(int maxIter, double eps, b200_csr_t Mgt, b200_csr_t* Mt_io,
                           int* iters_done, double* chaos_hist, double* ms_per_iter) {
  return b200_rmcl_iter_sharded_stats(maxIter, eps, Mgt, Mt_io, iters_done, chaos_hist, ms_per_iter,
                                      nullptr);
}

int b200_rmcl_iter_sharded_stats(int maxIter, double eps, b200_csr_t Mgt, b200_csr_t* Mt_io,
                                 int* iters_done, double* chaos_hist, double* ms_per_iter,
                                 long long* counts_per_iter) {
  B200_REQUIRE_INIT();
  if (!Mgt || !Mt_io || !*Mt_io) { set_error("null handle"); return B200_ERR_BAD_ARG; }
  if (g_nranks > 1 && !g_comm) { set_error("b200_comm_init() not called"); return B200_ERR_NCCL; }
  Ctx& c = ctx();
  cudaStream_t st = c.stream;
  const int n = Mgt->d.rows;
  const int R = g_nranks, r = g_rank;
  // `cur` is the current iterate.  It starts as the caller's matrix and is replaced only when a
  // new iterate is complete, so whatever happens *Mt_io ends up holding a valid matrix (the last
  // completed iterate): every exit goes through the single hand-back below.
  DevCSR cur = (*Mt_io)->d;
  struct CsrGuard {  // frees a matrix under construction unless it has been handed on
    DevCSR d;
    ~CsrGuard() { dfree(d.rowptr); dfree(d.col); dfree(d.val); }
    DevCSR take() { DevCSR t = d; d = DevCSR(); return t; }
  };
  struct Events {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~Events() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
  } ev;
  int it = 0;
  auto body = [&]() -> int {
    Temps T;  // d_prefix, d_meta: freed on every path
    int64_t* d_prefix = nullptr;
    long long* d_meta = nullptr;  // per rank: {nnz of its block, error code, chaos bits, unpruned nnz, row tiles}
    long long* d_P = nullptr;
    constexpr int MW = 5;
    B200_CUDA(T.alloc(&d_P, 1));
    B200_CUDA(T.alloc(&d_prefix, (size_t)n + 1));
    B200_CUDA(T.alloc(&d_meta, (size_t)R * MW));
    B200_CUDA(cudaEventCreate(&ev.e0));
    B200_CUDA(cudaEventCreate(&ev.e1));
    std::vector<int> ends((size_t)R + 1);
    std::vector<long long> meta((size_t)R * MW), off((size_t)R + 1, 0);
    b200_stats bst;
    double prev_ch = 0.0;
    for (; it < maxIter; ++it) {
      B200_CUDA(cudaEventRecord(ev.e0, st));
      // 1. flops-balanced cut points, found on the device (identical on every rank: same
      //    inputs, same arithmetic as arrayEqualPartition64); 2. the local block.  A local
      //    failure is not returned yet: the peers are about to enter a collective, so the error
      //    code travels with the block sizes and every rank leaves the loop together.
      CsrGuard blk;
      double ch = 0.0;
      // (with more than one rank the cut balances products + a charge per heavy row, see
      // k_row_cost; one rank has nothing to balance)
      int lrc = flops_prefix_device(Mgt->d, cur, 0, n, d_prefix, R > 1 ? c.tun.row_charge : 0, d_P);
      if (!lrc) lrc = equal_partition_device(d_prefix, n, R, ends.data());
      long long P_total = 0;
      if (!lrc && counts_per_iter)
        B200_CUDA(cudaMemcpyAsync(&P_total, d_P, sizeof(long long), cudaMemcpyDeviceToHost, st));
      if (!lrc) lrc = rmcl_step_device(Mgt->d, cur, ends[r], ends[r + 1], &blk.d, &ch,
                                       counts_per_iter ? &bst : nullptr);
      long long unpruned = (counts_per_iter && !lrc) ? bst.nnz_unpruned : 0;
      long long tiles = (counts_per_iter && !lrc) ? std::max(1, bst.row_tiles) : 0;
      const long long launches = (counts_per_iter && !lrc) ? bst.launches : 0;
      // With several ranks every rank would otherwise sort the WHOLE gathered Mt again at the
      // start of the next step (the bitmap kernels cut B rows at column boundaries): each rank
      // orders its own block before the exchange instead — 1/R of that work.  (Ascending rows
      // change the first-touch order the next iteration's hash bins see, i.e. the order of its
      // row sums: values agree with a 1-rank run to rounding, not bitwise.)
      if (R > 1 && !lrc && blk.d.nnz > 0 && !blk.d.sorted_rows) lrc = sort_rows_device(&blk.d);
      CsrGuard next;
      if (R == 1) {
        if (lrc) return lrc;
        next.d = blk.take();
      } else {
        // 3. one all-gather of {block nnz, error code, chaos}, then the ragged blocks
        long long mine[MW] = {lrc ? 0 : blk.d.nnz, (long long)lrc, 0, unpruned, tiles};
        memcpy(&mine[2], &ch, sizeof(double));
        B200_CUDA(cudaMemcpyAsync(d_meta + MW * r, mine, sizeof mine, cudaMemcpyHostToDevice, st));
        B200_NCCL(AllGather(d_meta + MW * r, d_meta, MW, ncclInt64, g_comm, st));
        B200_CUDA(cudaMemcpyAsync(meta.data(), d_meta, (size_t)R * MW * sizeof(long long),
                                  cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaStreamSynchronize(st));
        unpruned = 0;
        tiles = 0;
        for (int q = 0; q < R; ++q) {
          if (meta[MW * q + 1] != 0) {
            if (!lrc) set_error("rank " + std::to_string(q) + " failed in its rMCL step (code " +
                                std::to_string(meta[MW * q + 1]) + "); the loop stops on every rank");
            return lrc ? lrc : (int)meta[MW * q + 1];
          }
          double chq;
          memcpy(&chq, &meta[MW * q + 2], sizeof(double));
          ch = std::max(ch, chq);
          off[q + 1] = off[q] + meta[MW * q];
          unpruned += meta[MW * q + 3];
          tiles = std::max(tiles, meta[MW * q + 4]);
        }
        next.d.rows = n; next.d.cols = cur.cols; next.d.nnz = off[R];
        next.d.sorted_rows = true;   // every rank sorted its block
        B200_CUDA(dalloc(&next.d.rowptr, (size_t)n + 1));
        B200_CUDA(dalloc(&next.d.col, (size_t)next.d.nnz));
        B200_CUDA(dalloc(&next.d.val, (size_t)next.d.nnz));
        // my block into place (row offsets shifted to global positions)
        const int myrows = ends[r + 1] - ends[r];
        B200_CUDA(cudaMemcpyAsync(next.d.rowptr + ends[r], blk.d.rowptr, ((size_t)myrows + 1) * sizeof(int64_t),
                                  cudaMemcpyDeviceToDevice, st));
        k_shift_rowptr<<<(unsigned)((myrows + 1 + 255) / 256), 256, 0, st>>>(next.d.rowptr + ends[r],
                                                                            myrows + 1, off[r]);
        if (blk.d.nnz) {
          B200_CUDA(cudaMemcpyAsync(next.d.col + off[r], blk.d.col, (size_t)blk.d.nnz * sizeof(int), cudaMemcpyDeviceToDevice, st));
          B200_CUDA(cudaMemcpyAsync(next.d.val + off[r], blk.d.val, (size_t)blk.d.nnz * sizeof(double), cudaMemcpyDeviceToDevice, st));
        }
        // (between GroupStart and GroupEnd a failed call must still close the group)
        ncclResult_t nres = g_nccl.GroupStart();
        for (int q = 0; q < R && nres == ncclSuccess; ++q) {
          const int rows_q = ends[q + 1] - ends[q];
          // row offsets: rows_q entries starting at ends[q]; the closing entry of block q is the
          // opening entry of block q+1 (or nnz for the last), so broadcast rows_q(+1 for the last)
          const size_t nrp = (size_t)rows_q + (q == R - 1 ? 1 : 0);
          const size_t cq = (size_t)meta[MW * q];
          if (nrp) nres = g_nccl.Broadcast(next.d.rowptr + ends[q], next.d.rowptr + ends[q], nrp, ncclInt64, q, g_comm, st);
          if (cq && nres == ncclSuccess)
            nres = g_nccl.Broadcast(next.d.col + off[q], next.d.col + off[q], cq, ncclInt32, q, g_comm, st);
          if (cq && nres == ncclSuccess)
            nres = g_nccl.Broadcast(next.d.val + off[q], next.d.val + off[q], cq, ncclDouble, q, g_comm, st);
        }
        const ncclResult_t gres = g_nccl.GroupEnd();
        if (nres != ncclSuccess) return fail_nccl(nres, "Broadcast (all-gather of the row blocks)");
        if (gres != ncclSuccess) return fail_nccl(gres, "GroupEnd");
      }
      B200_CUDA(cudaEventRecord(ev.e1, st));
      B200_CUDA(cudaStreamSynchronize(st));   // the new iterate is complete on this rank
      dfree(cur.rowptr); dfree(cur.col); dfree(cur.val);
      cur = next.take();
      if (ms_per_iter) { float ms = 0; cudaEventElapsedTime(&ms, ev.e0, ev.e1); ms_per_iter[it] = ms; }
      if (chaos_hist) chaos_hist[it] = ch;
      if (counts_per_iter) {
        counts_per_iter[5 * it + 0] = P_total;
        counts_per_iter[5 * it + 1] = cur.nnz;
        counts_per_iter[5 * it + 2] = unpruned;
        counts_per_iter[5 * it + 3] = tiles;
        counts_per_iter[5 * it + 4] = launches;
      }
      if (rmcl_converged(ch, prev_ch, it, eps)) { ++it; break; }
      prev_ch = ch;
    }
    return sort_rows_device(&cur);  // Mt.makeOrdered() (nrmcl.cc:25-26)
  };
  const int rc = body();
  (*Mt_io)->d = cur;
  if (iters_done) *iters_done = it;
  cudaStreamSynchronize(st);
  return rc;
}

}  // extern "C"
