// Where does the time of a large device -> malloc'd-host download go?  (tools/micro, not product)
//   nvcc -O3 -Xcompiler -fopenmp,-mavx2 -o d2h_rate d2h_rate.cu && ./d2h_rate [GB]
// Prints: raw DMA rate into pinned memory, host copy rates pinned -> pageable (fresh / touched,
// memcpy / non-temporal), parallel first-touch rate, cudaHostRegister / Unregister cost of a
// block and the direct DMA rate into a registered block.
#include <cuda_runtime.h>
#include <immintrin.h>
#include <omp.h>
#include <sys/mman.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void advise_huge(void* p, size_t bytes) {
  const uintptr_t two_mb = (uintptr_t)2 << 20;
  const uintptr_t a = ((uintptr_t)p + two_mb - 1) & ~(two_mb - 1), e = ((uintptr_t)p + bytes) & ~(two_mb - 1);
  if (e > a) madvise((void*)a, e - a, MADV_HUGEPAGE);
}
static void par_memcpy(void* dst, const void* src, size_t bytes) {
#pragma omp parallel
  {
    const size_t nt = omp_get_num_threads(), t = omp_get_thread_num();
    const size_t piece = (((bytes + nt - 1) / nt) + 4095) & ~(size_t)4095, off = t * piece;
    if (off < bytes) memcpy((char*)dst + off, (const char*)src + off, bytes - off < piece ? bytes - off : piece);
  }
}
static void par_ntcopy(void* dst, const void* src, size_t bytes) {   // bytes multiple of 32, aligned
#pragma omp parallel
  {
    const size_t nt = omp_get_num_threads(), t = omp_get_thread_num();
    const size_t piece = (((bytes + nt - 1) / nt) + 4095) & ~(size_t)4095, off = t * piece;
    if (off < bytes) {
      const size_t len = bytes - off < piece ? bytes - off : piece;
      const __m256i* s = (const __m256i*)((const char*)src + off);
      __m256i* d = (__m256i*)((char*)dst + off);
      for (size_t i = 0; i < len / 32; ++i) _mm256_stream_si256(d + i, _mm256_load_si256(s + i));
      _mm_sfence();
    }
  }
}
static void par_touch(void* p, size_t bytes) {
#pragma omp parallel
  {
    const size_t nt = omp_get_num_threads(), t = omp_get_thread_num();
    const size_t piece = (((bytes + nt - 1) / nt) + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1), off = t * piece;
    for (size_t o = off; o < bytes && o < off + piece; o += 4096) ((volatile char*)p)[o] = 0;
  }
}
static void* fresh(size_t bytes) {
  void* p = aligned_alloc(2u << 20, bytes);
  advise_huge(p, bytes);
  return p;
}

int main(int argc, char** argv) {
  const size_t GB = (size_t)1 << 30;
  const size_t bytes = (argc > 1 ? (size_t)atoi(argv[1]) : 4) * GB;
  const size_t PIN = 128u << 20;
  printf("threads %d, block %.1f GB\n", omp_get_max_threads(), bytes / 1e9);
  if (system("cat /sys/kernel/mm/transparent_hugepage/enabled /sys/kernel/mm/transparent_hugepage/defrag; nproc; grep -E 'MemTotal|MemFree' /proc/meminfo; lscpu | grep -E 'Model name|Socket|NUMA node\\(s\\)'")) {}
  char* dev; CK(cudaMalloc(&dev, bytes)); CK(cudaMemset(dev, 1, bytes));
  char* pin[2]; for (int k = 0; k < 2; ++k) CK(cudaHostAlloc(&pin[k], PIN, cudaHostAllocDefault));
  cudaStream_t s; CK(cudaStreamCreate(&s));
  // 1. raw DMA into pinned chunks
  for (int rep = 0; rep < 2; ++rep) {
    double t0 = now();
    for (size_t off = 0; off < bytes; off += PIN) CK(cudaMemcpyAsync(pin[(off / PIN) & 1], dev + off, PIN, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    printf("dma -> pinned (128 MB chunks)       %6.1f GB/s\n", bytes / (now() - t0) / 1e9);
  }
  // 2. host copy pinned -> fresh / touched
  {
    char* d = (char*)fresh(bytes);
    double t0 = now();
    for (size_t off = 0; off < bytes; off += PIN) par_memcpy(d + off, pin[0], PIN);
    double t1 = now();
    for (size_t off = 0; off < bytes; off += PIN) par_memcpy(d + off, pin[0], PIN);
    double t2 = now();
    for (size_t off = 0; off < bytes; off += PIN) par_ntcopy(d + off, pin[0], PIN);
    double t3 = now();
    printf("memcpy pinned -> fresh pageable     %6.1f GB/s\n", bytes / (t1 - t0) / 1e9);
    printf("memcpy pinned -> touched pageable   %6.1f GB/s\n", bytes / (t2 - t1) / 1e9);
    printf("ntcopy pinned -> touched pageable   %6.1f GB/s\n", bytes / (t3 - t2) / 1e9);
    free(d);
    d = (char*)fresh(bytes);
    t0 = now();
    for (size_t off = 0; off < bytes; off += PIN) par_ntcopy(d + off, pin[0], PIN);
    printf("ntcopy pinned -> fresh pageable     %6.1f GB/s\n", bytes / (now() - t0) / 1e9);
    free(d);
  }
  // 3. first touch, 4. register / direct DMA / unregister
  {
    char* d = (char*)fresh(bytes);
    double t0 = now();
    par_touch(d, bytes);
    double t1 = now();
    CK(cudaHostRegister(d, bytes, cudaHostRegisterDefault));
    double t2 = now();
    CK(cudaMemcpyAsync(d, dev, bytes, cudaMemcpyDeviceToHost, s)); CK(cudaStreamSynchronize(s));
    double t3 = now();
    CK(cudaHostUnregister(d));
    double t4 = now();
    printf("parallel first touch                %6.1f GB/s\n", bytes / (t1 - t0) / 1e9);
    printf("cudaHostRegister (touched)          %6.1f GB/s  (%.3f s)\n", bytes / (t2 - t1) / 1e9, t2 - t1);
    printf("dma -> registered block             %6.1f GB/s\n", bytes / (t3 - t2) / 1e9);
    printf("cudaHostUnregister                  %6.1f GB/s  (%.3f s)\n", bytes / (t4 - t3) / 1e9, t4 - t3);
    printf("touch+register+dma+unregister       %6.1f GB/s\n", bytes / (t4 - t0) / 1e9);
    free(d);
    d = (char*)fresh(bytes);
    t0 = now();
    CK(cudaHostRegister(d, bytes, cudaHostRegisterDefault));
    t1 = now();
    printf("cudaHostRegister (fresh)            %6.1f GB/s  (%.3f s)\n", bytes / (t1 - t0) / 1e9, t1 - t0);
    CK(cudaHostUnregister(d));
    free(d);
  }
  // 5. chunked register pipeline: a helper thread touches + registers 512 MB pieces ahead of the DMA
  {
    const size_t PIECE = 512u << 20;
    char* d = (char*)fresh(bytes);
    const size_t np = bytes / PIECE;
    volatile int ready = 0;
    double t0 = now();
#pragma omp parallel num_threads(2)
    {
      if (omp_get_thread_num() == 0) {
        for (size_t k = 0; k < np; ++k) {
          char* q = d + k * PIECE;
          for (size_t o = 0; o < PIECE; o += 4096) ((volatile char*)q)[o] = 0;
          CK(cudaHostRegister(q, PIECE, cudaHostRegisterDefault));
          __sync_synchronize(); ready = (int)k + 1;
        }
      } else {
        for (size_t k = 0; k < np; ++k) {
          while (ready <= (int)k) { }
          CK(cudaMemcpyAsync(d + k * PIECE, dev + k * PIECE, PIECE, cudaMemcpyDeviceToHost, s));
        }
        CK(cudaStreamSynchronize(s));
      }
    }
    double t1 = now();
    for (size_t k = 0; k < np; ++k) CK(cudaHostUnregister(d + k * PIECE));
    double t2 = now();
    printf("pipelined touch+register | dma      %6.1f GB/s, unregister %.3f s => %6.1f GB/s total\n",
           bytes / (t1 - t0) / 1e9, t2 - t1, bytes / (t2 - t0) / 1e9);
    free(d);
  }
  return 0;
}
