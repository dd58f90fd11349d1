// Microbenchmark (development aid): SM-side throughput of scattered fp64 RED on B200.
// Each CTA owns a private region (L2-resident) and every lane adds to a pseudo-random slot.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(double* acc, int region, int iters, int spread) {
  double* a = acc + (size_t)blockIdx.x * region;
  unsigned s = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
  const int lane = threadIdx.x & 31;
  for (int it = 0; it < iters; ++it) {
    s = s * 1664525u + 1013904223u;
    // spread = 0: consecutive lanes hit consecutive slots around a random warp base (coalesced);
    // spread = k: lanes k slots apart (k*8 bytes): 4 -> one sector per lane
    unsigned wbase = __shfl_sync(0xffffffffu, s, 0) % (unsigned)(region - 32 * (spread + 1));
    unsigned idx = spread < 0 ? (s >> 8) % (unsigned)region : wbase + lane * (spread ? spread : 1);
    if (MODE == 0) atomicAdd(a + idx, 1.0);
    else if (MODE == 1) asm volatile("red.relaxed.cta.global.add.f64 [%0], %1;" ::"l"(a + idx), "d"(1.0) : "memory");
    else if (MODE == 2) atomicAdd((float*)(a + idx), 1.0f);
    else if (MODE == 3) a[idx] = 1.0;
    else if (MODE == 4) atomicAdd((unsigned long long*)(a + idx), 1ull);
  }
}
template <int MODE> void run(const char* name, double* d, int region, int iters, int spread, int sms) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms, 1024>>>(d, region, 10, spread);
  cudaEventRecord(e0);
  k<MODE><<<sms, 1024>>>(d, region, iters, spread);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)sms * 1024 * iters;
  printf("%-22s spread %3d region %8d: %8.3f ms  %7.2f Gop/s  %.3f lane-ops/cycle/SM (at 1.965 GHz)\n", name, spread, region, ms,
         ops / ms / 1e6, ops / (ms * 1e-3) / sms / 1.965e9);
}
int main() {
  int sms = 148; const int region_max = 1 << 20; double* d;
  cudaMalloc(&d, (size_t)sms * region_max * 8); cudaMemset(d, 0, (size_t)sms * region_max * 8);
  for (int region : {1 << 12, 1 << 15, 1 << 17, 1 << 20})
    for (int spread : {-1, 1, 4}) {
      run<0>("red.f64 gpu", d, region, 2000, spread, sms);
    }
  for (int spread : {-1, 1, 4}) {
    run<1>("red.f64 cta", d, 1 << 15, 2000, spread, sms);
    run<2>("red.f32 gpu", d, 1 << 15, 2000, spread, sms);
    run<4>("red.u64 gpu", d, 1 << 15, 2000, spread, sms);
    run<3>("st.f64", d, 1 << 15, 2000, spread, sms);
  }
  return 0;
}
