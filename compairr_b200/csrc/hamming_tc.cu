// hamming_tc.cu — K5 on the 5th-generation tensor cores: d >= 3 Hamming comparison of one length
// bucket as an int8 GEMM (replaces process_trad + seq_diff, overlap.cc:286-359, util.cc:172-184,
// for buckets where it really is a dense contraction).
//
//   row image        NT: 4 one-hot bytes per position (D = number of equal positions, exact);
//                    AA: an 8-byte weighted code per position (equal residues contribute 8,
//                    different ones at most 6): D >= 8 (L - d) is NECESSARY for distance <= d, the
//                    few candidates are verified exactly.  2.5 x fewer MACs and bytes than one-hot.
//   work item        256 set-A rows (two 128-row A tiles, built once) x a run of 128-row set-B tiles
//   D = A * B^T      tcgen05.mma.cta_group::1.kind::i8, M = 128, N = 128, K = 32 per instruction,
//                    S32 accumulators in TMEM: 2 stages x 2 A tiles x 128 columns = all 512 columns
//   operands         shared memory, canonical K-major no-swizzle layout (8-row x 16-byte core
//                    matrices, row groups SBO = 128 B apart, 16-byte K chunks LBO = (rows / 8) * 128 B
//                    apart), written by the CTA's own threads from packed residues through a
//                    residue-pair LUT (no TMA: the tiles do not exist in memory); fence.proxy.async
//                    before the MMAs
//   warps (13)       0-7 epilogue (TMEM lane quarter x column half: tcgen05.ld, threshold test by a
//                    tree of 3-input maxima, rare candidates -> per-warp queue -> exact verify,
//                    score, matrix atomics, pair append; D never leaves the SM), 8-11 tile builders
//                    (thread = set-B row, next tile's words prefetched), 12 MMA issue
//   pipeline         mbarriers only, no CTA barrier inside an item: FULL_B[s] (128 builder
//                    arrivals) -> MMA; tcgen05.commit -> ACC_FULL[a] and B_FREE[s]; ACC_FREE[a]
//                    as soon as an accumulator stage is in registers; up to 4 B stages
#include <cuda_runtime.h>

#include "device_utils.cuh"
#include "hamming_tc.cuh"

namespace cb {

constexpr int TC_THREADS = 416;  // warps 0-7: epilogue; 8-11: tile builders; 12: MMA issue
constexpr uint32_t TC_LUT_STRIDE = TC_PACK_PAD + 1;  // residue-pair LUT: (r0, r1) -> 16-byte K chunk
constexpr uint32_t TC_IDESC = (2u << 4)              // D format: S32
                              | (0u << 7) | (0u << 10)  // A, B format: unsigned 8-bit
                              | ((TC_NB >> 3) << 17)    // N
                              | ((TC_ROWS >> 4) << 24); // M

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

// Write the image of one sequence (or an all-zero row) into row `row` of a 128-row tile: every
// 16-byte K chunk of the row is written exactly once with one 128-bit store, so tiles are never
// zeroed.  Residues come from the bucket's packed array (4 per word, word-major: coalesced across
// rows; the last word of a sequence is filled up with TC_PACK_PAD, whose image is zero); `first`
// holds packed words 0-3 of the row, requested by the caller ahead of time so their L2 latency is
// off the critical path; later groups of four are requested one group ahead.
//
//   AA (alphabet <= 20)   8 bytes per position, values {0,1,2}; a chunk = two positions, read
//                         ready-made from the residue-pair LUT in shared memory
//   NT (alphabet <= 4)    4 bytes per position, one-hot; a chunk = four positions = one packed word
template <bool AA>
__device__ __forceinline__ void build_row(uint8_t* tile, uint32_t row, bool valid, const uint32_t* __restrict__ src,
                                          uint32_t stride, uint32_t len, uint32_t kpad, const uint4* lut,
                                          const uint32_t (&first)[4]) {
  constexpr uint32_t PADW = TC_PACK_PAD * 0x01010101u;
  const uint32_t words = valid ? (len + 3) >> 2 : 0;
  const uint32_t nw = AA ? (kpad + 31) >> 5 : kpad >> 4;  // packed-word slots that cover the padded row
  uint8_t* const dst = tile + (row >> 3) * 128 + (row & 7) * 16;
  uint32_t cur[4], nxt[4];
#pragma unroll
  for (int j = 0; j < 4; j++) cur[j] = first[j];
  for (uint32_t g = 0; g * 4 < nw; g++) {
#pragma unroll
    for (uint32_t j = 0; j < 4; j++) {
      const uint32_t k = (g + 1) * 4 + j;
      nxt[j] = k < words ? __ldg(src + (uint64_t)k * stride) : PADW;
    }
#pragma unroll
    for (uint32_t j = 0; j < 4; j++) {
      const uint32_t k = g * 4 + j;
      if (k < nw) {  // warp-uniform
        const uint32_t w = cur[j];
        if (AA) {
          *reinterpret_cast<uint4*>(dst + (2 * k) * (TC_ROWS * 16)) =
              lut[(w & 0xff) * TC_LUT_STRIDE + ((w >> 8) & 0xff)];
          if (2 * k + 1 < (kpad >> 4))
            *reinterpret_cast<uint4*>(dst + (2 * k + 1) * (TC_ROWS * 16)) =
                lut[((w >> 16) & 0xff) * TC_LUT_STRIDE + (w >> 24)];
        } else {
          uint32_t c[4];
#pragma unroll
          for (uint32_t bb = 0; bb < 4; bb++) {
            const uint32_t r = (w >> (8 * bb)) & 0xff;
            c[bb] = r < 4 ? 1u << (r * 8) : 0u;
          }
          *reinterpret_cast<uint4*>(dst + k * (TC_ROWS * 16)) = make_uint4(c[0], c[1], c[2], c[3]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) cur[j] = nxt[j];
  }
}

// packed words 0-3 of one sequence of a bucket (all padding for a row past the end of the bucket)
__device__ __forceinline__ void request_row(uint32_t (&w)[4], const uint32_t* __restrict__ src, uint32_t stride,
                                            bool valid, uint32_t len) {
  const uint32_t words = (len + 3) >> 2;
#pragma unroll
  for (uint32_t j = 0; j < 4; j++)
    w[j] = (valid && j < words) ? __ldg(src + (uint64_t)j * stride) : TC_PACK_PAD * 0x01010101u;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// wait for the next phase of barrier `idx` of this role's sequence; `phases` keeps one parity bit
// per barrier (every completion of a barrier is awaited exactly once by each of its waiters)
__device__ __forceinline__ void mbar_wait_next(uint64_t* bars, uint32_t idx, uint32_t& phases) {
  mbar_wait(smem_u32(bars + idx), (phases >> idx) & 1);
  phases ^= 1u << idx;
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

#define TMEM_LD32(v, taddr)                                                                                     \
  asm volatile(                                                                                                 \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                 \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                 \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                 \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),         \
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),   \
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), \
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])  \
      : "r"(taddr))

// tcgen05.wait::ld, with the destination registers as operands so no use of them can be scheduled
// ahead of the wait
#define TMEM_WAIT32(v)                                                                                          \
  asm volatile("tcgen05.wait::ld.sync.aligned;"                                                                 \
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), \
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]),       \
                 "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]),     \
                 "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]),     \
                 "+r"(v[29]), "+r"(v[30]), "+r"(v[31])::"memory")

// largest of 32 accumulator columns, as a depth-4 tree of 3-input maxima (VIMNMX3): the common
// case is "no column of this row reaches the threshold"
__device__ __forceinline__ int max3(int a, int b, int c) { return max(a, max(b, c)); }
__device__ __forceinline__ int max32(const uint32_t (&v)[32]) {
  int t[11];
#pragma unroll
  for (int j = 0; j < 10; j++) t[j] = max3((int)v[3 * j], (int)v[3 * j + 1], (int)v[3 * j + 2]);
  t[10] = max((int)v[30], (int)v[31]);
  const int u0 = max3(t[0], t[1], t[2]), u1 = max3(t[3], t[4], t[5]), u2 = max3(t[6], t[7], t[8]);
  return max3(max3(u0, u1, u2), t[9], t[10]);
}

__device__ __forceinline__ uint32_t hit_mask(const uint32_t (&v)[32], int thr) {
  uint32_t hits = 0;
#pragma unroll
  for (int j = 0; j < 32; j++) hits |= ((int)v[j] >= thr ? 1u : 0u) << j;
  return hits;
}

// barrier indices: up to TC_BSTAGES shared-memory stages of B tiles, two accumulator stages
enum { BAR_FULL_B = 0, BAR_B_FREE = TC_BSTAGES, BAR_ACC_FULL = 2 * TC_BSTAGES, BAR_ACC_FREE = 2 * TC_BSTAGES + 2,
       BAR_COUNT = 2 * TC_BSTAGES + 4 };

template <bool AA>
__global__ void __launch_bounds__(TC_THREADS, 1) hamming_tc_kernel(const __grid_constant__ TcLaunch P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool builder_warp = warp >= 8 && warp < 12, mma_warp = warp == 12;  // warps 0-7: epilogue
  // carve: A tiles 0,1 | B tile stages | residue-pair LUT | candidate queues | barriers | tmem pointer | item slot
  const uint32_t tile_bytes = TC_ROWS * P.kmax;
  uint8_t* const tile_a = smem;
  uint8_t* const tile_b = smem + 2 * (size_t)tile_bytes;
  const uint32_t n_bs = P.b_stages;
  uint4* const lut = reinterpret_cast<uint4*>(smem + (2 + n_bs) * (size_t)tile_bytes);
  uint2* const queues = reinterpret_cast<uint2*>(lut + TC_LUT_STRIDE * TC_LUT_STRIDE);
  uint64_t* const bars = reinterpret_cast<uint64_t*>(queues + 8 * TC_QCAP);
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(bars + BAR_COUNT);
  uint32_t* const item_slot = tmem_slot + 1;
  uint2* const q = queues + (warp & 7) * TC_QCAP;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 32) {
    for (uint32_t s2 = 0; s2 < TC_BSTAGES; s2++) {
      mbar_init(bars + BAR_FULL_B + s2, 128);  // every builder thread arrives
      mbar_init(bars + BAR_B_FREE + s2, 1);    // tcgen05.commit
    }
    for (uint32_t s2 = 0; s2 < 2; s2++) {
      mbar_init(bars + BAR_ACC_FULL + s2, 1);  // tcgen05.commit
      mbar_init(bars + BAR_ACC_FREE + s2, 8);  // one lane of every epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (AA) {
    // residue code, 8 bytes: low digit r & 3 one-hot with weight 2 in bytes 0-3; high digit r >> 2
    // as weight 2 in byte 4 + (r >> 2), its fifth value (residues 16-19) as 1,1,1,1.  Equal
    // residues score 8, different ones 0, 2, 4 or 6; TC_PACK_PAD is the all-zero code.
    auto code = [](uint32_t r) {
      const uint32_t h = r >> 2;
      return r < 20 ? make_uint2(2u << ((r & 3) * 8), h < 4 ? 2u << (h * 8) : 0x01010101u) : make_uint2(0, 0);
    };
    for (uint32_t i = tid; i < TC_LUT_STRIDE * TC_LUT_STRIDE; i += TC_THREADS) {
      const uint2 c0 = code(i / TC_LUT_STRIDE), c1 = code(i % TC_LUT_STRIDE);
      lut[i] = make_uint4(c0.x, c0.y, c1.x, c1.y);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  uint32_t phases = 0;    // this thread's parity bit per barrier it waits on
  uint32_t acc_used = 0;  // MMA warp: accumulator stages written at least once
  uint32_t nmatch = 0;

  for (;;) {
    // all roles meet here: every MMA of the previous item has completed (the builders waited for
    // them), every accumulator has been read
    if (tid == 0) *item_slot = (uint32_t)atomicAdd(P.counters + CTR_WORK, 1ull);
    __syncthreads();
    const uint32_t it = *item_slot;
    if (it >= P.n_items) break;
    const TcItem I = P.items[it];
    const uint32_t kpad = I.kpad, ksteps = kpad >> 5;
    const int thr = (AA ? 8 : 1) * ((int)I.len - P.differences);
    const uint32_t n_tiles = (I.b_n + TC_NB - 1) / TC_NB;
    const uint32_t n_at = (I.a_n + TC_ROWS - 1) / TC_ROWS;  // 1 or 2 accumulator row blocks
    const uint32_t* const a_src = P.a_packed + I.a_pack + I.a_pos;
    const uint32_t* const b_src = P.b_packed + I.b_pack + I.b_pos;

    // item prologue: threads 0-255 write one set-A row each (made visible to the MMA warp by the
    // CTA barrier below); the builders start their pipeline with B tile 0
    uint32_t pre[4];
    if (tid < n_at * TC_ROWS) {
      const bool valid = tid < I.a_n;
      request_row(pre, a_src + tid, I.a_bucket, valid, I.len);
      build_row<AA>(tile_a + (size_t)(tid >> 7) * tile_bytes, tid & 127, valid, a_src + tid, I.a_bucket, I.len, kpad,
                    lut, pre);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    } else if (builder_warp) {
      request_row(pre, b_src + (tid - TC_MA), I.b_bucket, tid - TC_MA < I.b_n, I.len);
    }
    __syncthreads();

    if (mma_warp) {
      // ---- MMA issue: tile t needs its B tile built and its accumulator stage read out
      for (uint32_t t = 0, bs = 0; t < n_tiles; t++, bs = bs + 1 == n_bs ? 0 : bs + 1) {
        const uint32_t s2 = t & 1;
        mbar_wait_next(bars, BAR_FULL_B + bs, phases);
        if (acc_used & (1u << s2)) mbar_wait_next(bars, BAR_ACC_FREE + s2, phases);
        acc_used |= 1u << s2;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          const uint32_t b_addr = smem_u32(tile_b + (size_t)bs * tile_bytes);
          constexpr uint32_t lbo = TC_ROWS * 16;
          for (uint32_t at = 0; at < n_at; at++) {
            const uint32_t a_addr = smem_u32(tile_a + (size_t)at * tile_bytes);
            for (uint32_t s = 0; s < ksteps; s++)
              mma_i8(tmem_base + s2 * 256 + at * TC_NB, make_desc(a_addr + 2 * s * lbo, lbo, 128),
                     make_desc(b_addr + 2 * s * lbo, lbo, 128), s > 0);
          }
          mma_commit(bars + BAR_ACC_FULL + s2);
          mma_commit(bars + BAR_B_FREE + bs);
        }
        __syncwarp();
      }
    } else if (builder_warp) {
      // ---- tile builders: thread = row; tile t goes to stage t mod n_bs once MMA[t - n_bs] has
      // released it
      const uint32_t brow = tid - TC_MA;
      for (uint32_t t = 0, bs = 0; t < n_tiles; t++, bs = bs + 1 == n_bs ? 0 : bs + 1) {
        const uint32_t r = t * TC_NB + brow;
        if (t >= n_bs) mbar_wait_next(bars, BAR_B_FREE + bs, phases);
        uint32_t nxt[4];
        request_row(nxt, b_src + r + TC_NB, I.b_bucket, r + TC_NB < I.b_n, I.len);
        build_row<AA>(tile_b + (size_t)bs * tile_bytes, brow, r < I.b_n, b_src + r, I.b_bucket, I.len, kpad, lut, pre);
#pragma unroll
        for (int j = 0; j < 4; j++) pre[j] = nxt[j];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(bars + BAR_FULL_B + bs);
      }
      // the releases of the last n_bs tiles: every completion is awaited exactly once, and after
      // them no MMA of this item is still reading shared memory
      for (uint32_t t = n_tiles >= n_bs ? n_tiles - n_bs : 0; t < n_tiles; t++)
        mbar_wait_next(bars, BAR_B_FREE + t % n_bs, phases);
    } else {
      // ---- epilogue: warp w reads lane quarter w & 3 (set-A rows) and column half w >> 2 of each
      // accumulator block, 32 columns at a time into two alternating register sets; the stage is
      // handed back to the MMA warp as soon as its last chunk is in registers.  Rows past a_n and
      // columns past the tile's b_n are all-zero rows: dot product 0 < thr, no masking needed.
      uint32_t qn = 0;  // this warp's queued candidates (warp-uniform)
      // K5b: candidates of this warp, one per lane: exact residue compare (the AA code is a
      // filter; the NT one-hot dot product is already the exact number of equal positions), then
      // score, matrix atomics and pair append (overlap.cc:300-340)
      auto drain = [&]() {
        for (uint32_t b0 = 0; b0 < qn; b0 += 32) {
          if (b0 + lane < qn) {
            const uint2 e = q[b0 + lane];
            bool ok = true;
            if (AA) {
              const uint32_t words = (I.len + 3) >> 2;
              uint32_t mism = 0;
              for (uint32_t k = 0; k < words; k++)
                mism += __popc(__vcmpne4(__ldg(a_src + (uint64_t)k * I.a_bucket + e.x),
                                         __ldg(b_src + (uint64_t)k * I.b_bucket + e.y)) &
                               0x01010101u);
              ok = (int)mism <= P.differences;
            }
            if (ok) {
              const uint32_t aseq = __ldg(P.a_order + I.a_start + e.x);
              const uint32_t bseq = __ldg(P.b_order + I.b_start + e.y);
              const SeqMeta am = ld_meta(P.a.meta + aseq);
              const SeqMeta bm = ld_meta(P.b.meta + bseq);
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
                  po.b = bseq + P.b.index_base;
                  P.pairs[at] = po;
                }
              }
            }
          }
        }
        __syncwarp();
        qn = 0;
      };
      // append this lane's hit columns (bit mask over columns col0..col0+31 of row a_local), one
      // per lane per round, compacted by ballot
      auto push = [&](uint32_t hits, uint32_t a_local, uint32_t col0) {
        for (;;) {
          const bool has = hits != 0;
          const uint32_t bal = __ballot_sync(FULL, has);
          if (!bal) break;
          if (qn > TC_QCAP - 32) drain();
          if (has) {
            const uint32_t j = __ffs(hits) - 1;
            hits &= hits - 1;
            q[qn + __popc(bal & ((1u << lane) - 1))] = make_uint2(a_local, col0 + j);
          }
          qn += __popc(bal);
          __syncwarp();
        }
      };
      const uint32_t quarter = warp & 3, col = (warp >> 2) * 64;
      const uint32_t row0 = quarter * 32 + lane;
      for (uint32_t t = 0; t < n_tiles; t++) {
        const uint32_t s2 = t & 1;
        mbar_wait_next(bars, BAR_ACC_FULL + s2, phases);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tbase = tmem_base + s2 * 256 + col + ((quarter * 32u) << 16);
        const uint32_t bcol = t * TC_NB + col;
        auto release = [&]() {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(bars + BAR_ACC_FREE + s2);
        };
        auto test = [&](const uint32_t (&v)[32], uint32_t a_local, uint32_t col0) {
          if (__any_sync(FULL, max32(v) >= thr)) push(hit_mask(v, thr), a_local, col0);
        };
        uint32_t va[32], vb[32];
        TMEM_LD32(va, tbase);
        TMEM_WAIT32(va);
        TMEM_LD32(vb, tbase + 32);
        test(va, row0, bcol);
        TMEM_WAIT32(vb);
        if (n_at == 2) TMEM_LD32(va, tbase + TC_NB); else release();
        test(vb, row0, bcol + 32);
        if (n_at == 2) {
          TMEM_WAIT32(va);
          TMEM_LD32(vb, tbase + TC_NB + 32);
          test(va, TC_ROWS + row0, bcol);
          TMEM_WAIT32(vb);
          release();
          test(vb, TC_ROWS + row0, bcol + 32);
        }
        if (qn >= 32) drain();
      }
      drain();
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nmatch += __shfl_xor_sync(FULL, nmatch, o);
  if (lane == 0 && nmatch) atomicAdd(P.counters + CTR_MATCHES, (unsigned long long)nmatch);
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

size_t tc_smem_bytes(uint32_t kmax, uint32_t b_stages) {
  return (size_t)(2 + b_stages) * TC_ROWS * kmax + TC_LUT_STRIDE * TC_LUT_STRIDE * sizeof(uint4) + 8 * TC_QCAP * sizeof(uint2) +
         BAR_COUNT * 8 + 16;
}

uint32_t tc_cols_per_position(uint32_t sigma) { return sigma <= 4 ? 4 : sigma <= 20 ? 8 : 0; }

int launch_hamming_tc(TcLaunch p, int sm_count, cudaStream_t st, const char** err) {
  // as many B stages as fit next to the two A tiles (narrow rows leave room for a deeper ring)
  p.b_stages = TC_BSTAGES;
  while (p.b_stages > 2 && tc_smem_bytes(p.kmax, p.b_stages) > 227 * 1024) p.b_stages--;
  const size_t smem = tc_smem_bytes(p.kmax, p.b_stages);
  if (smem > 227 * 1024) {
    *err = "tiles do not fit shared memory";
    return -1;
  }
  auto kernel = p.aa ? hamming_tc_kernel<true> : hamming_tc_kernel<false>;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    *err = "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed for the tensor-core kernel";
    return -1;
  }
  if (cudaMemsetAsync(p.counters + CTR_WORK, 0, sizeof(unsigned long long), st) != cudaSuccess) {
    *err = "cudaMemsetAsync failed";
    return -1;
  }
  const unsigned grid = (unsigned)(p.n_items < (uint32_t)sm_count ? p.n_items : (uint32_t)sm_count);
  kernel<<<grid, TC_THREADS, smem, st>>>(p);
  return 1;
}

}  // namespace cb
