/* b200_spgemm.h — C-ABI of the B200-native SpGEMM / rMCL hot path.
 *
 * Drop-in boundary for ONE path of ankur-maximos/Sparse_Matrix_with_Flops: row-wise Gustavson
 * CSR x CSR multiplication with flops-based row binning, and the rMCL expansion / inflation /
 * prune / normalise loop it drives.  Every entry point cites the reference interface it
 * replaces (paths relative to the reference root).  Plain pointers and sizes only; no C++ or
 * torch types cross this boundary.  All functions return B200_OK (0) or an error code; the
 * text of the last error on the calling thread is available from b200_last_error().
 *
 * Conventions shared with the reference (nlibs/CSR.h:23-50): 0-based CSR, `int` indices,
 * `double` values (the reference's QValue built with -DQValue=double -DFDOUBLE), rows need not
 * be sorted on input, a row must not contain duplicate columns (nlibs/cpu_csr_kernel.h:151-158
 * relies on this as well).  Host output arrays are malloc() blocks owned by the caller and
 * released with free(), exactly like CSR::dispose() (nlibs/CSR.h:323-327).
 *
 * Output order.  SpGEMM: every row of C has ASCENDING column indices (the reference emits
 * first-touch order and sorts later in CSR::makeOrdered, nlibs/CSR.cc:73-86).  A single rMCL
 * step (b200_rmcl_onestep_csr, b200_rmcl_step_device*) keeps the REFERENCE's storage order —
 * first-touch order with pruned entries removed — for every row computed by a hash bin,
 * because the next iteration's row sums run in that order and only so do the iterates stay
 * bit-identical to the reference's; rows computed by the large-row (bitmap) bin come out
 * ascending.  The loops (b200_rmcl_iter, b200_rmcl_iter_sharded) sort the final matrix
 * (makeOrdered) before returning it; b200_csr_sort_rows does it on demand.
 */
#ifndef B200_SPGEMM_H_
#define B200_SPGEMM_H_

#ifdef __cplusplus
extern "C" {
#endif

enum {
  B200_OK = 0,
  B200_ERR_BAD_ARG = 1,
  B200_ERR_CUDA = 2,
  B200_ERR_HOST_ALLOC = 3,
  B200_ERR_INT32_OVERFLOW = 4, /* result does not fit the reference's int CSR (CSR.h:38) */
  B200_ERR_NO_DEVICE = 5,
  B200_ERR_NCCL = 6,
  B200_ERR_NOT_INIT = 7,
  B200_ERR_CALLBACK = 8 /* a b200_block_fn returned non-zero: the streamed product was abandoned */
};

/* Phase timings and counters of the last device call (all times in milliseconds, CUDA events
 * on the library's stream). */
typedef struct b200_stats {
  double ms_total;
  double ms_flops;     /* flops analysis + scan + binning   (flops_csr_kernel.cc:14-31,61) */
  double ms_symbolic;  /* nnz(C row) count                  (cpu_csr_kernel.h:234-262)     */
  double ms_numeric;   /* accumulate + sort (+ rMCL epilogue) (cpu_csr_kernel.h:134-188)   */
  double ms_other;     /* scans, allocation, compaction                                     */
  long long products;  /* intermediate products P = sum_i sum_{j in A_i} nnz(B_j)           */
  long long nnz_out;   /* nnz of the result                                                 */
  long long nnz_unpruned; /* rMCL only: nnz before pruning (== nnz_out for SpGEMM)          */
  int launches;        /* kernels of this library launched by the call                      */
  int bins_rows[16];   /* rows per numeric bin (diagnostic)                                 */
  /* Per-bin breakdown (bench.py's roofline leg): device time of each bin's symbolic and
   * numeric kernel (CUDA events on the library stream around the launch), and the rows,
   * intermediate products, nnz(A rows) and nnz(C rows) each bin processed.  Symbolic bins are
   * keyed by products per row, numeric bins by nnz(C row); see DESIGN.md for the cut points. */
  double ms_sym_bin[16];
  double ms_num_bin[16];
  long long sym_bin_rows[16], sym_bin_products[16], sym_bin_nnzA[16];
  long long num_bin_products[16], num_bin_nnzA[16], num_bin_nnzC[16];
  int part_kernel;     /* 1: the bitmap bin's numeric pass ran as k_num_bitmap_part */
  int part_count;      /* column parts it used */
  int range_items;     /* work items of the on-chip numeric pass (k_num_items); 0: not used */
  int ranges;          /* static column ranges it planned with */
  int row_tiles;       /* rMCL step: row tiles it ran through the bounded arena (0 or 1: one pass) */
} b200_stats;

/* ---- context ------------------------------------------------------------------------- */

/* Select CUDA device `device`, create the library stream and memory pool.  Fails loudly with
 * B200_ERR_NO_DEVICE when no CUDA device is present: there is no CPU fallback. */
int b200_init(int device);
int b200_finalize(void);
/* Developer switches (B200_* environment variables) are read once by b200_init;
 * this re-reads them (tests and A/B measurements change them between calls). */
int b200_options_reload(void);
/* Opt-in top-k pruning of the rMCL steps (every entry point below that runs one): of the entries
 * of a row that pass the reference's threshold (nlibs/tools/util.cc:4-9, 47-69) only the k
 * largest stay — equal values ranked by ascending column — and the row is normalised over them.
 * 0 (the default) switches it off: the reference has the threshold rule only,
 * and parity runs use that.  oracle/oracle.c states the same rule for the checker. */
int b200_set_topk(int k);
const char* b200_last_error(void);
/* The CUDA stream (a cudaStream_t) every kernel of this library is launched on, so that a
 * harness can bracket calls with its own CUDA events. */
int b200_stream(void** stream);
/* Launch configuration facts, for the harness: SM count and device name. */
int b200_device_info(int* sm_count, long long* hbm_bytes, char* name, int name_len);

/* ---- host-buffer entry points -------------------------------------------------------- */

/* C = A x B.  Replaces flops_omp_CSR_SpMM (nlibs/flops_csr_kernel.cc:122-142, declared at
 * nlibs/cpu_csr_kernel.h:95-98) and omp_CSR_SpMM (nlibs/omp_csr_kernel.cc:296-315,
 * cpu_csr_kernel.h:73-76): same argument meaning; IC/JC/C are malloc()'d here.
 * Returns B200_ERR_INT32_OVERFLOW if nnz(C) > INT_MAX (use the row-block device API). */
int b200_spgemm_csr(const int* IA, const int* JA, const double* A, int nnzA,
                    const int* IB, const int* JB, const double* B, int nnzB,
                    int** IC, int** JC, double** C, int* nnzC,
                    int m, int k, int n);

/* One rMCL iteration newMt = prune(inflate(A x B)).  Replaces static_omp_CSR_RMCL_OneStep
 * (nlibs/static_omp_csr_kernel.cc:208-284, cpu_csr_kernel.h:194-197) and
 * omp_CSR_RMCL_OneStep (nlibs/omp_csr_kernel.cc:154-198).  `chaos` (may be NULL) receives
 * max_i (max_j M[i,j] - sum_j M[i,j]^2) of the result (not in the reference). */
int b200_rmcl_onestep_csr(const int* IA, const int* JA, const double* A, int nnzA,
                          const int* IB, const int* JB, const double* B, int nnzB,
                          int** IC, int** JC, double** C, int* nnzC,
                          int m, int k, int n, double* chaos);

/* The rMCL loop.  Replaces gpuRmclIter (nlibs/gpus/gpu_csr_kernel.cu:281-312, declared at
 * nlibs/gpus/gpu_csr_kernel.h:5) and mtRmclIter (nlibs/qrmcl.cc:8-84): Mt <- Mgt x Mt with
 * inflation/prune/normalise, `maxIter` times; with eps > 0 it also stops after the first iteration whose chaos is
 * < eps or moved by less than eps since the previous iteration (the reference's
 * loop is fixed-count; oracle/oracle.c states the same rule for the checker).  The input
 * Mt arrays are borrowed; the final Mt is returned in malloc()'d IM/JM/M.  chaos_hist (may be
 * NULL) must have room for maxIter doubles. */
int b200_rmcl_iter(int maxIter, double eps,
                   const int* IG, const int* JG, const double* G, int nnzG,
                   const int* IT, const int* JT, const double* T, int nnzT,
                   int** IM, int** JM, double** M, int* nnzM,
                   int n, int* iters_done, double* chaos_hist);

/* C = A x B delivered as consecutive row blocks, for products larger than the reference's
 * `int` CSR can hold (nnz(A x A) of BASELINE.json's R-MAT scale 20 is 9.7e9).  Same operands
 * as flops_omp_CSR_SpMM (nlibs/cpu_csr_kernel.h:95-98); the row space is cut where the
 * intermediate-product prefix (dynamic_omp_CSR_flops, nlibs/flops_csr_kernel.cc:14-31) reaches
 * `block_products` (<= 0: 2e9, so that every block's nnz fits an int; a single heavier row is a
 * block of its own).  Each block is handed to `fn` as a malloc()'d int CSR with IC[0] = 0
 * (IC has row_hi-row_lo+1 entries) — `fn` owns the three arrays (b200_host_free / free) and
 * returns 0 to continue.  The download of block b runs on a copy stream and a helper thread
 * while block b+1 is being computed, so a step costs about max(compute, PCIe) instead of their
 * sum.  `fn` is called on the calling thread, in row order; it must not call back into the
 * library (a transfer is in flight), except for b200_host_free, which also lets the library
 * reuse the block's memory for a later block instead of faulting in fresh pages. */
typedef int (*b200_block_fn)(void* user, int row_lo, int row_hi, int* IC, int* JC, double* C,
                             int nnzC);
int b200_spgemm_csr_stream(const int* IA, const int* JA, const double* A, int nnzA,
                           const int* IB, const int* JB, const double* B, int nnzB,
                           int m, int k, int n, long long block_products,
                           b200_block_fn fn, void* user);

/* ---- device-resident CSR (replaces CSR::toGpuCSR / toCpuCSR / deviceDispose,
 *      nlibs/CSR.cc:342-379) ------------------------------------------------------------ */

typedef struct b200_csr* b200_csr_t;

int b200_csr_upload(const int* I, const int* J, const double* V, int rows, int cols, int nnz,
                    b200_csr_t* out);
int b200_csr_info(b200_csr_t h, int* rows, int* cols, long long* nnz);
/* Whole matrix to malloc()'d int CSR; B200_ERR_INT32_OVERFLOW if nnz > INT_MAX. */
int b200_csr_download(b200_csr_t h, int** I, int** J, double** V, int* nnz);
/* Rows [row_lo,row_hi) as a malloc()'d int CSR with I[0]=0 (row-block convention of
 * SURVEY.md §7 hard part 1, for results larger than the reference's int CSR can hold). */
int b200_csr_download_rows(b200_csr_t h, int row_lo, int row_hi, int** I, int** J, double** V,
                           int* nnz);
int b200_csr_free(b200_csr_t h);
/* Edge list (COO, host arrays of nnz entries; val may be NULL = all ones) to a device CSR, built
 * on the device.  Replaces COO::addSelfLoopIfNeeded (nlibs/COO.cc:160-188), COO::makeOrdered +
 * COO::toCSR (COO.cc:222-235), orderedAndDuplicatesRemoving (COO.cc:237-266) and
 * CSR::averAndNormRowQValue (nlibs/CSR.cc:88-95); B200_COO_SELF_LOOPS | B200_COO_NORMALISE is
 * rmclInit (nlibs/qrmcl.cc:126-134).  Entries come out ordered by (row, column); with
 * B200_COO_DEDUP repeated (row, column) pairs become ONE entry whose value is the sum of theirs,
 * added in input order, as orderedAndDuplicatesRemoving does (COO.cc:246-248; with
 * B200_COO_NORMALISE the value is 1 / rowcount anyway) — without the flag they all stay, which
 * the multiplication entry points do not accept (SURVEY.md §8c input hazards).
 * Indices outside the matrix are an error. */
#define B200_COO_DEDUP 1
#define B200_COO_SELF_LOOPS 2
#define B200_COO_NORMALISE 4
int b200_coo_to_csr(const int* rowIndex, const int* colIndex, const double* val, long long nnz,
                    int rows, int cols, int flags, b200_csr_t* out);
/* CSR::makeOrdered (nlibs/CSR.cc:73-86) on the device: sort every row by column, in place. */
int b200_csr_sort_rows(b200_csr_t h);
/* Device pointers of a handle (row offsets are 64-bit on the device). */
int b200_csr_device_ptrs(b200_csr_t h, void** rowptr64, void** colind32, void** values64);

/* C = A x B on the device.  Replaces gpuSpMMWrapper (nlibs/gpus/gpu_csr_kernel.cu:128-172,
 * gpu_csr_kernel.h:6).  stats may be NULL. */
int b200_spgemm_device(b200_csr_t A, b200_csr_t B, b200_csr_t* C, b200_stats* stats);
/* Rows [row_lo,row_hi) of A only: C has row_hi-row_lo rows. (Row-block / multi-GPU shard.) */
int b200_spgemm_device_rows(b200_csr_t A, b200_csr_t B, int row_lo, int row_hi, b200_csr_t* C,
                            b200_stats* stats);
/* b200_spgemm_csr_stream on uploaded operands (gpuSpMMWrapper + CSR::toCpuCSR per row block,
 * nlibs/gpus/gpu_csr_kernel.cu:128-172, nlibs/CSR.cc:356-371, pipelined). */
int b200_spgemm_device_stream(b200_csr_t A, b200_csr_t B, long long block_products,
                              b200_block_fn fn, void* user);
/* One rMCL iteration on the device.  Replaces gpuRmclOneStepWrapper
 * (nlibs/gpus/gpu_csr_kernel.cu:243-279). */
int b200_rmcl_step_device(b200_csr_t Mgt, b200_csr_t Mt, b200_csr_t* newMt, double* chaos,
                          b200_stats* stats);
int b200_rmcl_step_device_rows(b200_csr_t Mgt, b200_csr_t Mt, int row_lo, int row_hi,
                               b200_csr_t* newMt, double* chaos, b200_stats* stats);

/* Per-row intermediate-product counts as an exclusive 64-bit prefix sum, prefix[rows] = P.
 * Replaces dynamic_omp_CSR_flops (nlibs/flops_csr_kernel.cc:14-31).  `prefix` is a host
 * array of rows+1 long long. */
int b200_flops_prefix(b200_csr_t A, b200_csr_t B, long long* prefix);
/* The same prefix over a per-row COST: products plus `row_charge` for every row heavy enough for
 * the CTA-per-row kernels (more than 512 products) — the device analogue of the footprint
 * static_omp_CSR_SpMM balances, (products + nnz(C_i) + 32 + nnz(A_i)) >> 1
 * (nlibs/static_omp_csr_kernel.cc:28-62, dynamic_omp_CSR_IC_nnzC_footprints :68-95): with
 * b200_equal_partition64 it cuts a product into row blocks of equal TIME for several GPUs.
 * row_charge < 0: the library's default (32768, measured on R-MAT scale 20; B200_ROW_CHARGE). */
int b200_cost_prefix(b200_csr_t A, b200_csr_t B, long long row_charge, long long* prefix);
/* Equal-flops contiguous row cut points, ends[0..nparts].  Same arithmetic as
 * arrayEqualPartition64 (nlibs/tools/util.cc:123-135); host-only, no device needed. */
int b200_equal_partition64(const long long* prefix, int n, int nparts, int* ends);

/* cluster(i) = argmax_j M[i,j], ties -> smallest j (not in the reference).
 * labels is a host array of `rows` ints; an empty row gets -1. */
int b200_csr_row_argmax(b200_csr_t h, int* labels);
/* Concatenate row blocks (device) into one CSR: the all-gather's local assembly step. */
int b200_csr_concat_rows(const b200_csr_t* blocks, int nblocks, b200_csr_t* out);

/* Column stripe B[:, col_lo:col_hi) as a device CSR with LOCAL column indices (column - col_lo):
 * one block of the reference's PCSR (nlibs/PCSR.h:5-101, PCSR.cc:3-56; stripe b of c is
 * [b * stride, min(cols, (b + 1) * stride)) with stride = ceil(cols / c)).  A x stripe through
 * b200_spgemm_device is the column-striped product of PCSR::leftMultiply
 * (correctTests/pcsrTest.cc:7-19); b200_csr_concat_cols glues the per-stripe results back
 * together row by row (block b's columns shifted by the widths of the blocks before it). */
int b200_csr_column_stripe(b200_csr_t B, int col_lo, int col_hi, b200_csr_t* out);
int b200_csr_concat_cols(const b200_csr_t* blocks, int nblocks, b200_csr_t* out);

/* ---- multi-GPU (one process per GPU; NCCL over NVLink) --------------------------------
 * Not in the reference (single GPU, SURVEY.md §2.1 strategy table).  The 128-byte unique id
 * is created on rank 0 and shipped to the other ranks by the host program. */
int b200_comm_unique_id(char id[128]);
int b200_comm_init(int rank, int nranks, const char id[128]);
int b200_comm_destroy(void);
/* Sharded rMCL: every rank holds full Mgt and Mt handles; rank r computes rows
 * [ends[r],ends[r+1]) (flops-balanced, recomputed every iteration), then the pruned row
 * blocks are all-gathered and chaos is max-all-reduced.  On return *Mt_io is the full new Mt
 * on every rank (the old handle is freed). */
int b200_rmcl_iter_sharded(int maxIter, double eps, b200_csr_t Mgt, b200_csr_t* Mt_io,
                           int* iters_done, double* chaos_hist, double* ms_per_iter);
/* The same loop; counts_per_iter (may be NULL; room for 5 * maxIter values) receives per
 * iteration {intermediate products of the whole step, nnz of the new Mt, its unpruned nnz summed
 * over the ranks, most row tiles any rank ran the step in, kernels this rank launched} — what a
 * harness needs for the roofline of an iteration. */
int b200_rmcl_iter_sharded_stats(int maxIter, double eps, b200_csr_t Mgt, b200_csr_t* Mt_io,
                                 int* iters_done, double* chaos_hist, double* ms_per_iter,
                                 long long* counts_per_iter);

/* Releases a malloc()'d block returned by this library.  Large blocks (>= 64 MB) are kept for
 * the next download instead of going back to the OS, up to B200_HOST_CACHE_GB (environment;
 * default a quarter of the physical memory, at most 64 GB; 0 = never keep): the first touch of
 * a fresh 10 GB block costs more than copying into it.  free() on such blocks stays legal. */
void b200_host_free(void* p);
/* Bytes and blocks currently kept by b200_host_free, and the limit in force. */
int b200_host_cache_info(long long* bytes, int* blocks, long long* limit_bytes);
/* Gives every kept block back to the OS (b200_finalize does the same). */
int b200_host_cache_drop(void);
/* on != 0: blocks that enter the cache are page-locked (cudaHostRegister, once) and stay so while
 * they circulate; a later download into such a block is one DMA, without the staging buffer and
 * the host copy of the default path (B200_HOST_PIN=1 in the environment does the same).  In this
 * mode every block the library returned MUST be released with b200_host_free(); free() on a
 * page-locked block is an error.  Default: off (plain malloc blocks, free() legal). */
int b200_host_cache_pin(int on);
/* Counters since the process started: large-block requests served from the cache / not served,
 * downloads that went straight into a page-locked block, and page-locked blocks alive now. */
int b200_host_cache_stats(long long* hits, long long* misses, long long* direct, int* pinned_blocks);

#ifdef __cplusplus
}
#endif
#endif /* B200_SPGEMM_H_ */
