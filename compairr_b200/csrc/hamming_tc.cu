// hamming_tc.cu — K5 on the 5th-generation tensor cores: d >= 3 Hamming comparison of one length
// bucket as a one-hot int8 GEMM (replaces process_trad + seq_diff, overlap.cc:286-359,
// util.cc:172-184, for buckets where it really is a dense contraction).
//
//   A (set-A tile)  128 sequences x K one-hot bytes, K = sigma * L rounded up to 32
//   B (set-B tile)  256 sequences x K one-hot bytes
//   D = A * B^T     128 x 256 int32 in TMEM = number of EQUAL positions of every pair
//   epilogue        tcgen05.ld -> registers, match <=> D >= L - d; rare matches -> score,
//                   matrix atomics, pair append.  D never leaves the SM.
//
// tcgen05.mma.cta_group::1.kind::i8, M = 128, N = 256, K = 32 per instruction, operands in shared
// memory in the canonical K-major no-swizzle ("interleave") layout: 8-row x 16-byte core matrices,
// row groups SBO = 128 B apart, 16-byte K chunks LBO = (rows / 8) * 128 B apart.  The one-hot
// tiles are written by the CTA's threads (generic proxy), so a fence.proxy.async precedes the MMAs.
// One thread issues the MMAs and commits them to an mbarrier; all four warps wait on it and read
// their own 32-lane quarter of the accumulator.  Double-buffered: while the MMAs of B tile t run,
// the warps build the one-hot image of tile t+1 — 2 x 256 TMEM columns, two B buffers in shared
// memory.
#include <cuda_runtime.h>

#include "device_utils.cuh"
#include "hamming_tc.cuh"

namespace cb {

constexpr int TC_THREADS = 256;  // warps 0-3: epilogue (one TMEM lane quarter each); warps 4-7: tile builders
constexpr uint32_t TC_IDESC = (2u << 4)              // D format: S32
                              | (0u << 7) | (0u << 10)  // A, B format: unsigned 8-bit
                              | ((TC_N >> 3) << 17)     // N
                              | ((TC_M >> 4) << 24);    // M

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// K-major, no swizzle: start address, LBO, SBO in 16-byte units; descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ull << 46);
}

__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(TC_IDESC), "r"(accumulate), "r"(0u)
      : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t phase) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}\n" ::"r"(mbar),
      "r"(phase)
      : "memory");
}

// byte offset of (row, k) inside a one-hot tile of `rows` rows
__device__ __forceinline__ uint32_t onehot_off(uint32_t row, uint32_t k, uint32_t rows) {
  return (k >> 4) * (rows * 16) + (row >> 3) * 128 + (row & 7) * 16 + (k & 15);
}

// Zero a tile and write the one-hot image of n sequences of a bucket into it; executed by
// `nthreads` threads numbered t0 = 0..nthreads-1 that share named barrier `bar`.  The residues come
// from the bucket's packed array (4 per word, word-major), so the loads coalesce across rows.
__device__ __forceinline__ void build_tile(uint8_t* tile, uint32_t rows, uint32_t kpad,
                                           const uint32_t* __restrict__ packed, uint64_t pack_off,
                                           uint32_t bucket_n, uint32_t pos0, uint32_t n, uint32_t len,
                                           uint32_t sigma, uint32_t t0, uint32_t nthreads, int bar) {
  uint4* p = reinterpret_cast<uint4*>(tile);
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (uint32_t i = t0; i < rows * kpad / 16; i += nthreads) p[i] = z;
  asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nthreads) : "memory");
  const uint32_t words = (len + 3) >> 2;
  for (uint32_t r = t0; r < n; r += nthreads) {
    const uint32_t* src = packed + pack_off + pos0 + r;
    for (uint32_t k = 0; k < words; k++) {
      const uint32_t w = __ldg(src + (uint64_t)k * bucket_n);
#pragma unroll
      for (uint32_t bb = 0; bb < 4; bb++) {
        const uint32_t pp = k * 4 + bb;
        if (pp < len) tile[onehot_off(r, pp * sigma + ((w >> (8 * bb)) & 0xff), rows)] = 1;
      }
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__global__ void __launch_bounds__(TC_THREADS, 1) hamming_tc_kernel(const __grid_constant__ TcLaunch P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t tid = threadIdx.x, warp = tid >> 5;
  const bool epilogue_warp = warp < 4;
  // carve: A tile | B tile 0 | B tile 1 | barriers | tmem pointer
  uint8_t* const tile_a = smem;
  uint8_t* const tile_b0 = tile_a + (size_t)TC_M * P.kmax;
  uint8_t* const tile_b1 = tile_b0 + (size_t)TC_N * P.kmax;
  uint64_t* const mbar = reinterpret_cast<uint64_t*>(tile_b1 + (size_t)TC_N * P.kmax);  // two barriers
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(mbar + 2);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(2u * TC_N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar + 1)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  uint32_t phase[2] = {0, 0};
  uint32_t nmatch = 0;

  for (uint32_t it = blockIdx.x; it < P.n_items; it += gridDim.x) {
    const TcItem I = P.items[it];
    const uint32_t kpad = I.kpad, ksteps = kpad >> 5;
    const int thr = (int)I.len - P.differences;
    const uint32_t n_tiles = (I.b_n + TC_N - 1) / TC_N;

    // item prologue, all 256 threads: A tile and the first B tile
    build_tile(tile_a, TC_M, kpad, P.a_packed, I.a_pack, I.a_bucket, I.a_pos, I.a_n, I.len, P.sigma, tid, TC_THREADS, 1);
    build_tile(tile_b0, TC_N, kpad, P.b_packed, I.b_pack, I.b_bucket, I.b_pos, min((uint32_t)TC_N, I.b_n), I.len,
               P.sigma, tid, TC_THREADS, 1);
    const uint32_t aseq = (epilogue_warp && tid < I.a_n) ? __ldg(P.a_order + I.a_start + tid) : 0xffffffffu;
    __syncthreads();

    // software pipeline over the B tiles: iteration t issues MMA[t], runs epilogue[t-1] on warps
    // 0-3 and builds B[t+1] on warps 4-7, all three concurrently
    for (uint32_t t = 0; t <= n_tiles; t++) {
      if (t < n_tiles && tid == 128) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t buf = t & 1;
        const uint32_t a_addr = smem_u32(tile_a), b_addr = smem_u32(buf ? tile_b1 : tile_b0);
        const uint32_t lbo_a = TC_M * 16, lbo_b = TC_N * 16;
        for (uint32_t s = 0; s < ksteps; s++)
          mma_i8(tmem_base + buf * TC_N, make_desc(a_addr + 2 * s * lbo_a, lbo_a, 128),
                 make_desc(b_addr + 2 * s * lbo_b, lbo_b, 128), s > 0);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         smem_u32(mbar + buf))
                     : "memory");
      }
      if (t >= 1) {  // MMA[t-1] complete: its accumulator half is readable, its B buffer reusable
        const uint32_t pb = (t - 1) & 1;
        mbar_wait(smem_u32(mbar + pb), phase[pb]);
        phase[pb] ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      if (epilogue_warp) {
        if (t >= 1) {
          const uint32_t pt = t - 1, pb = pt & 1;
          const uint32_t bn = min((uint32_t)TC_N, I.b_n - pt * TC_N);
          for (uint32_t c0 = 0; c0 < bn; c0 += 32) {  // warp-uniform trip count
            uint32_t v[32];
            const uint32_t taddr = tmem_base + pb * TC_N + ((warp * 32u) << 16) + c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
                  "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
                  "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
                  "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            // threshold test without branches: one bit per column, then a (rare) loop over set bits
            uint32_t hits = 0;
#pragma unroll
            for (int j = 0; j < 32; j++) hits |= ((int)v[j] >= thr ? 1u : 0u) << j;
            if (bn - c0 < 32) hits &= (1u << (bn - c0)) - 1;
            if (aseq == 0xffffffffu) hits = 0;
            while (hits) {
              const uint32_t j = __ffs(hits) - 1;
              hits &= hits - 1;
              const uint32_t hit = __ldg(P.b_order + I.b_start + (uint64_t)pt * TC_N + c0 + j);
              const SeqMeta am = ld_meta(P.a.meta + aseq);
              const SeqMeta bm = ld_meta(P.b.meta + hit);
              nmatch++;
              if (!P.no_matrix) {
                const uint64_t mrow = P.existence ? (uint64_t)aseq - P.a_first : am.rep;
                atomicAdd(P.matrix + mrow * P.n_cols + bm.rep, score_of(P.score, P.ignore_counts, am.count, bm.count));
              }
              if (P.want_pairs) {
                const unsigned long long at = atomicAdd(P.counters + CTR_PAIRS, 1ull);
                if (at < P.pairs_cap) {
                  PairOut po;
                  po.a = aseq + P.a.index_base;
                  po.b = hit + P.b.index_base;
                  P.pairs[at] = po;
                }
              }
            }
          }
        }
      } else if (t + 1 < n_tiles) {
        const uint32_t nt = t + 1;
        build_tile((nt & 1) ? tile_b1 : tile_b0, TC_N, kpad, P.b_packed, I.b_pack, I.b_bucket,
                   I.b_pos + nt * TC_N, min((uint32_t)TC_N, I.b_n - nt * TC_N), I.len, P.sigma, tid - 128, 128, 2);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nmatch += __shfl_xor_sync(FULL, nmatch, o);
  if ((tid & 31) == 0 && nmatch) atomicAdd(P.counters + CTR_MATCHES, (unsigned long long)nmatch);
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2u * TC_N));
}

size_t tc_smem_bytes(uint32_t kmax) {
  return (size_t)(TC_M + 2 * TC_N) * kmax + 16 + 16 + 1024;
}

int launch_hamming_tc(const TcLaunch& p, int sm_count, cudaStream_t st, const char** err) {
  const size_t smem = tc_smem_bytes(p.kmax);
  if (smem > 227 * 1024) {
    *err = "one-hot tiles do not fit shared memory";
    return -1;
  }
  if (cudaFuncSetAttribute(hamming_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
      cudaSuccess) {
    *err = "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed for the tensor-core kernel";
    return -1;
  }
  const unsigned grid = (unsigned)(p.n_items < (uint32_t)sm_count ? p.n_items : (uint32_t)sm_count);
  hamming_tc_kernel<<<grid, TC_THREADS, smem, st>>>(p);
  return 1;
}

}  // namespace cb
