// Micro-benchmark: how many random 8-byte loads per second the chip sustains (a) as a function of
// the footprint they fall in (L2-resident ... HBM-resident) and (b) as a function of the shared
// memory the resident CTAs hold and of the load flavour.  This is the hardware ceiling of the Bloom
// stage of the enumeration kernels (one random 8-B word = one 32-B sector per probe).
// Every thread keeps 8 independent loads in flight.  nvcc -O3 -arch=sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}
template <int FLAVOUR>
__device__ __forceinline__ unsigned long long ld(const unsigned long long* p) {
  unsigned long long v;
  if (FLAVOUR == 0) return __ldg(p);                                                            // ld.global.nc
  if (FLAVOUR == 1) { asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p)); return v; }  // L2 only
  if (FLAVOUR == 2) { asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p)); return v; }
  asm volatile("ld.global.ca.u64 %0, [%1];" : "=l"(v) : "l"(p)); return v;
}
template <int FLAVOUR, int U>
__global__ void __launch_bounds__(256) probe(const unsigned long long* __restrict__ f, uint32_t nwords, uint64_t n, unsigned long long* sink) {
  extern __shared__ unsigned char dyn[];
  unsigned long long acc = 0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < n; t += U * stride) {
    unsigned long long v[U];
#pragma unroll
    for (int k = 0; k < U; k++) v[k] = ld<FLAVOUR>(f + __umulhi((uint32_t)mix(t + k * stride), nwords));
#pragma unroll
    for (int k = 0; k < U; k++) acc += v[k];
  }
  if (acc == 0x1234567) { *sink = acc; dyn[0] = 1; }
}
template <int FLAVOUR, int U>
float run(const unsigned long long* f, uint32_t nwords, uint64_t n, unsigned long long* sink, int ctas_per_sm, size_t smem) {
  cudaFuncSetAttribute(probe<FLAVOUR, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    probe<FLAVOUR, U><<<148 * ctas_per_sm, 256, smem>>>(f, nwords, n, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}
int main() {
  const uint64_t n = 2000000000ull;
  unsigned long long *f, *sink;
  const size_t max_bytes = 512ull << 20;
  cudaMalloc(&f, max_bytes); cudaMemset(f, 1, max_bytes); cudaMalloc(&sink, 8);
  const int mibs[] = {16, 48, 64, 80, 96, 128, 256, 512};
  for (int m : mibs) {
    const float ms = run<0, 8>(f, (uint32_t)(((size_t)m << 20) / 8), n, sink, 8, 0);
    printf("footprint %4d MiB  8 CTAs/SM x 256 thr x 8 loads  %6.1f G loads/s  %5.2f TB/s of 32-B sectors\n", m, n / ms / 1e6, n / ms / 1e6 * 32 / 1e3);
  }
  const uint32_t nw48 = (48u << 20) / 8;
  printf("-- 48 MiB footprint: resident threads, loads in flight per thread, shared memory held, load flavour\n");
  const int ctas[] = {2, 3, 4, 8};
  const size_t smems[] = {0, 16 << 10, 35 << 10, 51 << 10};
  for (int c : ctas)
    for (size_t s : smems) {
      if ((size_t)c * s > (200u << 10)) continue;
      printf("ctas/SM %d smem/CTA %3zu KB:", c, s >> 10);
      printf("  U2 nc %6.1f", n / run<0, 2>(f, nw48, n, sink, c, s) / 1e6);
      printf("  U4 nc %6.1f", n / run<0, 4>(f, nw48, n, sink, c, s) / 1e6);
      printf("  U8 nc %6.1f", n / run<0, 8>(f, nw48, n, sink, c, s) / 1e6);
      printf("  U2 cg %6.1f", n / run<1, 2>(f, nw48, n, sink, c, s) / 1e6);
      printf("  U4 cg %6.1f", n / run<1, 4>(f, nw48, n, sink, c, s) / 1e6);
      printf("  U4 noalloc %6.1f", n / run<2, 4>(f, nw48, n, sink, c, s) / 1e6);
      printf("  G loads/s\n");
    }
  printf("-- 48 MiB footprint, 3 CTAs/SM x 256 thr x 2 loads: explicit shared-memory carve-out (L1 = 256 KB - carve-out)\n");
  const int carve_kb[] = {0, 32, 64, 100, 132, 164, 196, 228};
  for (int kb : carve_kb) {
    cudaFuncSetAttribute(probe<0, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, kb * 100 / 228);
    cudaFuncSetAttribute(probe<1, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, kb * 100 / 228);
    printf("carve-out ~%3d KB:  nc %6.1f  cg %6.1f G loads/s\n", kb, n / run<0, 2>(f, nw48, n, sink, 3, 1024) / 1e6,
           n / run<1, 2>(f, nw48, n, sink, 3, 1024) / 1e6);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
