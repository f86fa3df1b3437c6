// Micro-benchmark: rates of the memory operations the table build is made of, at C3 geometry
// (10^8 keys, 2^28 x 16-B slots, 48 MiB + 150 MB filters).  nvcc -O3 -arch=sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}
struct Slot { unsigned long long hash, idx; };
template <int MODE>
__global__ void __launch_bounds__(256) k(uint64_t n, Slot* table, uint64_t mask, unsigned long long* f1, uint32_t nb1, unsigned long long* f2, uint32_t nb2, unsigned long long* sink, int sorted) {
  unsigned long long acc = 0;
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x) {
    // sorted: the key's top bits follow t (partitioned order), low bits random
    const uint64_t r = mix(t);
    const uint64_t h = sorted ? (((t * ((1ull << 62) / n)) & ~((1ull << 34) - 1)) | (r & ((1ull << 34) - 1))) : (r >> 2);
    const uint64_t slot = (h >> (62 - 28)) & mask;
    const uint32_t b1 = __umulhi((uint32_t)(h >> 30), nb1), b2 = __umulhi((uint32_t)(h >> 30), nb2);
    if (MODE == 0) acc += *reinterpret_cast<volatile unsigned long long*>(&table[slot].idx);           // random 8-B load
    if (MODE == 1) acc += atomicCAS(&table[slot].idx, ~0ull, h);                                        // CAS
    if (MODE == 2) { acc += atomicCAS(&table[slot].idx, ~0ull, h); table[slot].hash = h; }              // CAS + store
    if (MODE == 3) atomicOr(f1 + b1, 1ull << (r & 63));                                                 // RED 48 MiB
    if (MODE == 4) atomicOr(f2 + b2, 1ull << (r & 63));                                                 // RED 150 MB
    if (MODE == 5) { acc += atomicCAS(&table[slot].idx, ~0ull, h); table[slot].hash = h; atomicOr(f1 + b1, 1ull << (r & 63)); atomicOr(f2 + b2, 1ull << (r & 63)); }
    if (MODE == 6) table[slot].hash = h;                                                                // random 8-B store
    if (MODE == 7) { table[slot].idx = h; table[slot].hash = h; }                                       // random 16-B store (2x8)
  }
  if (acc == 0x1234567) *sink = acc;
}
int main() {
  const uint64_t n = 100000000ull, slots = 1ull << 28;
  const uint32_t nb1 = 48u << 17, nb2 = (uint32_t)(n * 12 / 64);
  Slot* table; unsigned long long *f1, *f2, *sink;
  cudaMalloc(&table, slots * sizeof(Slot)); cudaMalloc(&f1, (size_t)nb1 * 8); cudaMalloc(&f2, (size_t)nb2 * 8); cudaMalloc(&sink, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[] = {"load8", "cas", "cas+store", "red_f1_48MiB", "red_f2_150MB", "all(cas,store,2 red)", "store8", "store16"};
  for (int sorted = 0; sorted < 2; sorted++)
    for (int mode = 0; mode < 8; mode++) {
      float best = 1e9;
      for (int rep = 0; rep < 3; rep++) {
        cudaMemset(table, 0xff, slots * sizeof(Slot)); cudaMemset(f1, 0, (size_t)nb1 * 8); cudaMemset(f2, 0, (size_t)nb2 * 8);
        cudaEventRecord(e0);
        const int g = 148 * 16;
        switch (mode) {
          case 0: k<0><<<g, 256>>>(n, table, slots - 1, f1, nb1, f2, nb2, sink, sorted); break;
          case 1: k<1><<<g, 256>>>(n, table, slots - 1, f1, nb1, f2, nb2, sink, sorted); break;
          case 2: k<2><<<g, 256>>>(n, table, slots - 1, f1, nb1, f2, nb2, sink, sorted); break;
          case 3: k<3><<<g, 256>>>(n, table, slots - 1, f1, nb1, f2, nb2, sink, sorted); break;
          case 4: k<4><<<g, 256>>>(n, table, slots - 1, f1, nb1, f2, nb2, sink, sorted); break;
          case 5: k<5><<<g, 256>>>(n, table, slots - 1, f1, nb1, f2, nb2, sink, sorted); break;
          case 6: k<6><<<g, 256>>>(n, table, slots - 1, f1, nb1, f2, nb2, sink, sorted); break;
          case 7: k<7><<<g, 256>>>(n, table, slots - 1, f1, nb1, f2, nb2, sink, sorted); break;
        }
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      printf("%s %-22s %7.2f ms  %6.1f G ops/s\n", sorted ? "sorted  " : "unsorted", names[mode], best, n / best / 1e6);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
