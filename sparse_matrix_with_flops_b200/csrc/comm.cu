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
// Because a row's result does not depend on which rank computed it, the result is bit-identical
// for every GPU count.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>
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
  B200_REQUIRE_INIT();
  if (!Mgt || !Mt_io || !*Mt_io) { set_error("null handle"); return B200_ERR_BAD_ARG; }
  if (g_nranks > 1 && !g_comm) { set_error("b200_comm_init() not called"); return B200_ERR_NCCL; }
  Ctx& c = ctx();
  cudaStream_t st = c.stream;
  const int n = Mgt->d.rows;
  const int R = g_nranks, r = g_rank;
  DevCSR cur = (*Mt_io)->d;
  std::vector<long long> prefix((size_t)n + 1);
  std::vector<int> ends((size_t)R + 1);
  int64_t* d_prefix = nullptr;
  long long* d_meta = nullptr;   // per rank: nnz of its block
  double* d_chaos = nullptr;
  B200_CUDA(dalloc(&d_prefix, (size_t)n + 1));
  B200_CUDA(dalloc(&d_meta, (size_t)R));
  B200_CUDA(dalloc(&d_chaos, 1));
  cudaEvent_t e0, e1;
  B200_CUDA(cudaEventCreate(&e0));
  B200_CUDA(cudaEventCreate(&e1));
  int it = 0, rc = B200_OK;
  double prev_ch = 0.0;
  for (; it < maxIter; ++it) {
    B200_CUDA(cudaEventRecord(e0, st));
    // 1. flops-balanced cut points (identical on every rank: same inputs, same arithmetic)
    if ((rc = flops_prefix_device(Mgt->d, cur, 0, n, d_prefix))) break;
    B200_CUDA(cudaMemcpyAsync(prefix.data(), d_prefix, ((size_t)n + 1) * sizeof(long long),
                              cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    b200_equal_partition64(prefix.data(), n, R, ends.data());
    // 2. local block
    DevCSR blk;
    double ch = 0.0;
    if ((rc = run_pipeline(Mgt->d, cur, ends[r], ends[r + 1], MODE_RMCL, &blk, &ch, nullptr))) break;
    DevCSR next;
    if (R == 1) {
      next = blk;
    } else {
      // 3. exchange sizes, then the ragged blocks
      long long mine = blk.nnz;
      B200_CUDA(cudaMemcpyAsync(d_meta + r, &mine, sizeof(long long), cudaMemcpyHostToDevice, st));
      B200_NCCL(AllGather(d_meta + r, d_meta, 1, ncclInt64, g_comm, st));
      std::vector<long long> cnt((size_t)R), off((size_t)R + 1, 0);
      B200_CUDA(cudaMemcpyAsync(cnt.data(), d_meta, (size_t)R * sizeof(long long),
                                cudaMemcpyDeviceToHost, st));
      B200_CUDA(cudaMemcpyAsync(d_chaos, &ch, sizeof(double), cudaMemcpyHostToDevice, st));
      B200_NCCL(AllReduce(d_chaos, d_chaos, 1, ncclDouble, ncclMax, g_comm, st));
      B200_CUDA(cudaMemcpyAsync(&ch, d_chaos, sizeof(double), cudaMemcpyDeviceToHost, st));
      B200_CUDA(cudaStreamSynchronize(st));
      for (int q = 0; q < R; ++q) off[q + 1] = off[q] + cnt[q];
      next.rows = n; next.cols = cur.cols; next.nnz = off[R];
      B200_CUDA(dalloc(&next.rowptr, (size_t)n + 1));
      B200_CUDA(dalloc(&next.col, (size_t)next.nnz));
      B200_CUDA(dalloc(&next.val, (size_t)next.nnz));
      // my block into place (row offsets shifted to global positions)
      const int myrows = ends[r + 1] - ends[r];
      B200_CUDA(cudaMemcpyAsync(next.rowptr + ends[r], blk.rowptr, ((size_t)myrows + 1) * sizeof(int64_t),
                                cudaMemcpyDeviceToDevice, st));
      k_shift_rowptr<<<(unsigned)((myrows + 1 + 255) / 256), 256, 0, st>>>(next.rowptr + ends[r],
                                                                          myrows + 1, off[r]);
      if (blk.nnz) {
        B200_CUDA(cudaMemcpyAsync(next.col + off[r], blk.col, (size_t)blk.nnz * sizeof(int), cudaMemcpyDeviceToDevice, st));
        B200_CUDA(cudaMemcpyAsync(next.val + off[r], blk.val, (size_t)blk.nnz * sizeof(double), cudaMemcpyDeviceToDevice, st));
      }
      B200_NCCL(GroupStart());
      for (int q = 0; q < R; ++q) {
        const int rows_q = ends[q + 1] - ends[q];
        // row offsets: rows_q entries starting at ends[q]; the closing entry of block q is the
        // opening entry of block q+1 (or nnz for the last), so broadcast rows_q(+1 for the last)
        const size_t nrp = (size_t)rows_q + (q == R - 1 ? 1 : 0);
        if (nrp) B200_NCCL(Broadcast(next.rowptr + ends[q], next.rowptr + ends[q], nrp, ncclInt64, q, g_comm, st));
        if (cnt[q]) {
          B200_NCCL(Broadcast(next.col + off[q], next.col + off[q], (size_t)cnt[q], ncclInt32, q, g_comm, st));
          B200_NCCL(Broadcast(next.val + off[q], next.val + off[q], (size_t)cnt[q], ncclDouble, q, g_comm, st));
        }
      }
      B200_NCCL(GroupEnd());
      dfree(blk.rowptr); dfree(blk.col); dfree(blk.val);
    }
    dfree(cur.rowptr); dfree(cur.col); dfree(cur.val);
    cur = next;
    B200_CUDA(cudaEventRecord(e1, st));
    B200_CUDA(cudaStreamSynchronize(st));
    if (ms_per_iter) { float ms = 0; cudaEventElapsedTime(&ms, e0, e1); ms_per_iter[it] = ms; }
    if (chaos_hist) chaos_hist[it] = ch;
    if (rmcl_converged(ch, prev_ch, it, eps)) { ++it; break; }
    prev_ch = ch;
  }
  if (!rc) rc = sort_rows_device(&cur);  // Mt.makeOrdered() (nrmcl.cc:25-26)
  (*Mt_io)->d = cur;
  if (iters_done) *iters_done = it;
  dfree(d_prefix); dfree(d_meta); dfree(d_chaos);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaStreamSynchronize(st);
  return rc;
}

}  // extern "C"
