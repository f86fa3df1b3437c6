// kernels.cu — hand-written sm_100a kernels of the overlap hot path.
//
//   pack_meta_kernel   SoA upload -> 32-byte SeqMeta records (+ longest length)
//   hash_kernel        K1  batched Zobrist hashing                    (replaces zobrist.cc:74-88, db.cc:903-916)
//   build_kernel       K2  open-addressing insert (+ filter bits)     (replaces overlap.cc:63-128 insert part,
//                                                                      hashtable.h:48-77, bloompat.h:50-53)
//   build_tile_kernel  K2  the same insert for a whole set into an empty table: tiles of 4096 slots built in
//                          shared memory from keys sorted by tile, the table written once (+ tile_bounds_kernel)
//   dups_kernel        K2b exact-duplicate count                      (replaces overlap.cc:63-128 dup part, :579-605)
//   filter_kernel      K2  the four class filters of a large set, L2-sized word ranges at a time
//   identical_kernel   K3/K4 for d = 0: one thread per seed           (replaces overlap.cc:253-284 with variants.cc:260-268)
//   (d = 1, 2: the enumeration kernels and the table stage are in variant.cu; d >= 3: brute.cu, hamming_tc.cu)
//
// Integer and byte work: streaming passes (pack, hash), random read-modify-writes (table, filters)
// and dependent random loads (probe chains); nothing here is a GEMM, so no tensor cores (DESIGN.md
// has each kernel's bound).
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>

#include "kernels.cuh"
#include "device_utils.cuh"

namespace cb {

// ---------------------------------------------------------------------------------------------
// pack_meta
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ uint64_t col_load(const void* p, uint32_t w, uint64_t i) {
  switch (w) {
    case 1: return reinterpret_cast<const uint8_t*>(p)[i];
    case 2: return reinterpret_cast<const uint16_t*>(p)[i];
    case 4: return reinterpret_cast<const uint32_t*>(p)[i];
    default: return reinterpret_cast<const uint64_t*>(p)[i];
  }
}

// Lengths of any width -> u64 (input of the device prefix sum that turns lengths into offsets).
__global__ void __launch_bounds__(256) widen_kernel(const void* __restrict__ src, uint32_t w, uint64_t n,
                                                    uint64_t* __restrict__ dst) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    dst[i] = col_load(src, w, i);
}

void launch_widen(const void* src, uint32_t w, uint64_t n, uint64_t* dst, cudaStream_t st) {
  if (n == 0) return;
  const uint64_t blocks = (n + 255) / 256;
  widen_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>(src, w, n, dst);
}

// One chunk of the upload: columns of caller-chosen widths -> SeqMeta records.  starts[i] is the
// first residue of sequence i: either offsets[i] - off_sub (offsets mode) or scan[i] + res_add
// (lengths mode, scan = exclusive prefix sum of the chunk's lengths).
__global__ void __launch_bounds__(256)
pack_meta_kernel(PackCols k, uint64_t n, SeqRec* __restrict__ out, unsigned long long* counters) {
  uint32_t mymax = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t off, len;
    if (k.lengths) {
      off = k.starts[i] + k.res_add;
      len = col_load(k.lengths, k.len_w, i);
    } else {
      const uint64_t o0 = k.starts[i], o1 = k.starts[i + 1];
      off = o0 - k.off_sub;
      len = o1 - o0;
    }
    const uint32_t len32 = len > SEQ_MAX_LEN ? SEQ_MAX_LEN : (uint32_t)len;  // over-long: caught by the host
    SeqRec m;
    m.off_len = pack_off_len(off, len32);
    m.count = k.count ? col_load(k.count, k.count_w, i) : 1ull;
    m.v = k.v ? (uint32_t)col_load(k.v, k.v_w, i) : 0u;
    m.j = k.j ? (uint32_t)col_load(k.j, k.j_w, i) : 0u;
    m.rep = k.rep ? (uint32_t)col_load(k.rep, k.rep_w, i) : 0u;
    m.next = SEQ_NIL;
    out[i] = m;
    mymax = max(mymax, len32);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mymax = max(mymax, __shfl_xor_sync(FULL, mymax, o));
  if ((threadIdx.x & 31) == 0 && mymax) atomicMax(counters + CTR_MAXLEN, (unsigned long long)mymax);
}

void launch_pack_meta(const PackCols& k, uint64_t n, SeqRec* out, unsigned long long* counters,
                      cudaStream_t st) {
  if (n == 0) return;
  const uint64_t blocks = (n + 255) / 256;
  pack_meta_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>(k, n, out, counters);
}

// ---------------------------------------------------------------------------------------------
// K1: Zobrist hash.  One thread per sequence; the table (sigma x rows u64) is staged in shared
// memory when it fits, else read through L1.  Streaming: (L + 32) bytes in, 8 bytes out.
// ---------------------------------------------------------------------------------------------

template <bool ZSMEM>
__global__ void __launch_bounds__(256)
hash_kernel(const SeqRec* __restrict__ meta, const uint8_t* __restrict__ res, uint64_t n,
            const uint64_t* __restrict__ ztab, uint32_t zrows, uint32_t sigma, uint64_t seed,
            bool ignore_genes, uint64_t* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* zs = reinterpret_cast<uint64_t*>(smem_raw);
  if (ZSMEM) {
    for (uint32_t i = threadIdx.x; i < zrows * sigma; i += blockDim.x) zs[i] = ztab[i];
    __syncthreads();
  }
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const SeqMeta m = ld_meta(meta + i);
    uint64_t h = ignore_genes ? 0ull : vj_hash(seed, m.v, m.j);
    const uint32_t len = m.len <= zrows ? m.len : 0;  // longer than the table: rehashed later
    // The residues as aligned 64-bit words, one load ahead (a byte load per residue kept L1 at 84 %
    // of its peak and DRAM at 39 %).  Only words that hold a residue of this sequence are read.
    const uint64_t* w = reinterpret_cast<const uint64_t*>(res + (m.off & ~7ull));
    const uint32_t sh = (uint32_t)(m.off & 7);
    const uint32_t nw = len ? (sh + len + 7) >> 3 : 0;
    uint64_t cur = nw ? __ldg(w) : 0ull;
    uint32_t p = 0;
    for (uint32_t k = 0; k < nw; k++) {
      const uint64_t nxt = k + 1 < nw ? __ldg(w + k + 1) : 0ull;
      const uint32_t b0 = k == 0 ? sh : 0u;
      const uint32_t left = sh + len - 8 * k;  // bytes of the sequence from the start of this word on
      const uint32_t b1 = left < 8 ? left : 8u;
      uint64_t x = cur >> (8 * b0);
      for (uint32_t b = b0; b < b1; b++, p++) {
        const uint32_t r = (uint32_t)x & 0xffu;
        x >>= 8;
        h ^= ZSMEM ? zs[p * sigma + r] : __ldg(ztab + p * sigma + r);
      }
      cur = nxt;
    }
    out[i] = h;
  }
}

void launch_hash(const SeqRec* meta, const uint8_t* res, uint64_t n, const uint64_t* ztab,
                 uint32_t zrows, uint32_t sigma, uint64_t seed, bool ignore_genes, uint64_t* out,
                 cudaStream_t st) {
  if (n == 0) return;
  const uint64_t blocks = (n + 255) / 256;
  const unsigned grid = (unsigned)(blocks < 148 * 8 ? blocks : 148 * 8);
  const size_t zbytes = (size_t)zrows * sigma * sizeof(uint64_t);
  if (zbytes <= 40 * 1024) {
    hash_kernel<true><<<grid, 256, zbytes, st>>>(meta, res, n, ztab, zrows, sigma, seed,
                                                 ignore_genes, out);
  } else {
    hash_kernel<false><<<grid, 256, 0, st>>>(meta, res, n, ztab, zrows, sigma, seed,
                                              ignore_genes, out);
  }
}

// ---------------------------------------------------------------------------------------------
// K2: table + Bloom build.  One thread per set-B sequence.  Slot order inside a chain differs
// from the serial reference; harmless, every equal-hash slot of a chain is visited on probe.
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) table_clear_kernel(Slot* table, uint64_t slots) {
  const ulonglong2 e = make_ulonglong2(0ull, SLOT_EMPTY);
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < slots;
       i += (uint64_t)gridDim.x * blockDim.x)
    reinterpret_cast<ulonglong2*>(table)[i] = e;
}

void launch_table_clear(Slot* table, uint64_t slots, cudaStream_t st) {
  const uint64_t blocks = (slots + 255) / 256;
  table_clear_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>(table,
                                                                                        slots);
}

// Plain (coherent) loads for data other threads of the same kernel are publishing.
__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t* p) {
  return *reinterpret_cast<const volatile uint64_t*>(p);
}
__device__ __forceinline__ SeqMeta ld_meta_plain(const SeqRec* p) {
  const ulonglong2 lo = *reinterpret_cast<const ulonglong2*>(p);
  const uint4 hi = *(reinterpret_cast<const uint4*>(p) + 1);
  return unpack_rec(lo.x, lo.y, hi.x, hi.y, hi.z, hi.w);
}

// One thread per set-B sequence.  Identical (sequence, V, J) share ONE slot: the first arrival
// owns the slot (one CAS on the index word), later arrivals verify they really are the same
// sequence and push themselves onto the slot's occurrence list (atomicExch of the head,
// SeqRec.next = old head).  So probe chains never walk clusters of duplicates, the exact verify
// runs once per distinct sequence, and the filters hold distinct keys only.  The reference
// inserts every sequence into its own slot (overlap.cc:63-128); the set of (seed, hit) matches is
// the same.
// The index word carries the low 32 bits of the hash above the 32-bit head index, so a thread
// that meets an occupied slot can tell "possibly my sequence" from the one atomic word it read —
// no second word to wait for, no fence, no lock.  Slot.hash is written plainly: only the probe
// kernels (later launches) read it.
__global__ void __launch_bounds__(256)
build_kernel(SeqRec* meta, const uint8_t* __restrict__ res, const uint64_t* __restrict__ hash,
             const uint64_t* __restrict__ part_hash, const uint32_t* __restrict__ part_idx,
             uint64_t first, uint64_t n, bool ignore_genes, Slot* table, uint64_t mask,
             unsigned long long* bloom, uint32_t bloom_blocks, const uint32_t* __restrict__ sel,
             const unsigned long long* sel_n) {
  // sel / sel_n: insert only the sorted positions sel[0 .. *sel_n) — the keys the tiled build
  // (build_tile_kernel) could not place inside their tile.
  if (sel) n = *sel_n < n ? *sel_n : n;
  // part_hash/part_idx: the same keys sorted by their top hash bits (position t holds sequence
  // first + part_idx[t]); the grid then sweeps the table and the filters in address order.
  // Both loops have warp-uniform trip counts and the probe loop is voted: lanes that finish early
  // wait for their warp instead of running ahead — left to themselves the lanes of a warp drift
  // apart for good (measured: 8 of 32 lanes active on average) and every memory round trip is
  // paid four times over.
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t rounds = (n + stride - 1) / stride;
  for (uint64_t r = 0; r < rounds; r++) {
    const uint64_t k = r * stride + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    bool walking = k < n;
    const uint64_t t = walking && sel ? sel[k] : k;
    const uint64_t i = walking ? first + (part_idx ? part_idx[t] : t) : 0;
    const uint64_t h = walking ? (part_hash ? part_hash[t] * CB_HOME_INV : hash[i]) : 0;  // sorted keys are h * CB_HOME_MUL
    const uint32_t tag = slot_tag(h);
    const unsigned long long tagged = ((unsigned long long)tag << 32) | i;  // i < 2^32 - 1 (checked at upload)
    uint64_t slot = table_home(h, mask);
    SeqMeta me;
    bool have_me = false;
    while (__any_sync(FULL, walking)) {
      if (walking) {
        unsigned long long* idxp = reinterpret_cast<unsigned long long*>(&table[slot].idx);
        unsigned long long cur = ld_volatile_u64(&table[slot].idx);
        if (cur == SLOT_EMPTY) {
          cur = atomicCAS(idxp, SLOT_EMPTY, tagged);
          if (cur == SLOT_EMPTY) {  // we own the slot
            table[slot].hash = h;  // SeqRec.next is SEQ_NIL already (pack kernel / reset_next_kernel)
            if (bloom) {  // direct (pipelined-upload) path: the key's bits in the four class filters
#pragma unroll
              for (uint32_t cls = 0; cls < CB_CLASSES; cls++)
                atomicOr(bloom + pfilter_word(h, bloom_blocks, cls), pfilter_pattern(h, cls));
            }
            walking = false;
          }
        }
        if (walking && (uint32_t)(cur >> 32) == tag) {  // same hash tag: compare the sequences
          if (!have_me) {
            me = ld_meta_plain(meta + i);
            have_me = true;
          }
          const SeqMeta o = ld_meta_plain(meta + (uint32_t)cur);
          if (o.len == me.len && (ignore_genes || (o.v == me.v && o.j == me.j)) &&
              seq_equal(res, me.off, o.off, me.len)) {
            const unsigned long long old = atomicExch(idxp, tagged);
            meta[i].next = (uint32_t)old;
            walking = false;
          }
        }
        slot = (slot + 1) & mask;
      }
    }
  }
}

// Before a set is inserted a second time its occurrence links are reset — one streaming pass
// instead of a random store per slot owner inside the build kernel.
__global__ void __launch_bounds__(256) reset_next_kernel(SeqRec* meta, uint64_t n) {
  // only the records that carry a link are written: the rest of the pass is a read
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    if (meta[i].next != SEQ_NIL) meta[i].next = SEQ_NIL;
}

void launch_reset_next(SeqRec* meta, uint64_t n, cudaStream_t st) {
  if (n == 0) return;
  const uint64_t blocks = (n + 255) / 256;
  reset_next_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>(meta, n);
}

void launch_build(SeqRec* meta, const uint8_t* res, const uint64_t* hash, const uint64_t* part_hash,
                  const uint32_t* part_idx, uint64_t first, uint64_t n, bool ignore_genes, Slot* table,
                  uint64_t mask, unsigned long long* bloom, uint32_t bloom_blocks, cudaStream_t st) {
  if (n == 0) return;
  const uint64_t blocks = (n + 255) / 256;
  build_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>(
      meta, res, hash, part_hash, part_idx, first, n, ignore_genes, table, mask, bloom, bloom_blocks, nullptr, nullptr);
}

// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// K2, a whole set into an EMPTY table: the tiled build.  The keys arrive sorted by the tile
// (BUILD_TILE_SLOTS consecutive slots) their home slot lies in.  One CTA builds one tile in shared
// memory — the insert of build_kernel, CAS and occurrence lists on shared-memory words — and streams
// it out, empty slots included: the table is written once, front to back, and is neither cleared
// beforehand nor read.
// In two passes.  A key that meets a slot with its own hash tag is, nearly always, another
// occurrence of a sequence the tile already holds (a fifth of the keys of a repertoire collection)
// and has to be compared residue by residue before it is linked into the slot's list: three
// dependent DRAM round trips.  Done where it came up, a few lanes of every warp of every round
// waited for them in turn and the rest of the CTA at its barrier (first version, ncu: 44 % long-
// scoreboard + 26 % barrier stalls, 3.9 ms at 10^8 keys).  So pass 1 only claims empty slots and
// sets those keys aside; pass 2 inserts them all at once, every lane with its own round trips in
// flight.
// A key whose probe run leaves the tile goes to the spill list and is inserted by build_kernel
// (sel) once every tile is in memory; the linear-probing invariant holds — every slot from its
// home to the end of the tile was full when it left, and stays full.
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) tile_bounds_kernel(const uint64_t* __restrict__ key, uint64_t n, uint32_t shift,
                                                          uint32_t ntiles, uint32_t* __restrict__ tile_first) {
  // tile_first[g] = first sorted position whose key belongs to tile g or a later one; [ntiles] = n
  for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g <= ntiles; g += gridDim.x * blockDim.x) {
    uint64_t lo = g == ntiles ? n : 0, hi = n;
    const uint64_t want = (uint64_t)g << shift;
    while (lo < hi) {
      const uint64_t mid = (lo + hi) >> 1;
      if (key[mid] < want) lo = mid + 1; else hi = mid;
    }
    tile_first[g] = (uint32_t)lo;
  }
}

// Append t to the spill list for the lanes with p set: one atomicAdd per warp.  Every lane of the warp calls.
__device__ __forceinline__ void push_spill(bool p, uint32_t t, uint32_t* spill, unsigned long long* cursor) {
  const unsigned m = __ballot_sync(FULL, p);
  if (m == 0) return;
  const unsigned lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  unsigned long long base = 0;
  if (lane == leader) base = atomicAdd(cursor, (unsigned long long)__popc(m));
  base = __shfl_sync(FULL, base, leader);
  if (p) spill[base + __popc(m & ((1u << lane) - 1))] = t;
}

__global__ void __launch_bounds__(BUILD_TILE_THREADS, CB_TILE_CTAS)
build_tile_kernel(SeqRec* meta, const uint8_t* __restrict__ res, const uint64_t* __restrict__ part_key,
                  const uint32_t* __restrict__ part_idx, const uint32_t* __restrict__ tile_first, uint64_t first,
                  bool ignore_genes, Slot* table, uint32_t tbits, uint32_t ntiles, uint32_t* defer, uint32_t* spill,
                  unsigned long long* counters) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t n_defer;
  Slot* tile = reinterpret_cast<Slot*>(smem_raw);
  const ulonglong2 empty = make_ulonglong2(0ull, SLOT_EMPTY);
  // The key a thread inserts next and the bounds of the CTA's next tile are loaded one step ahead.
  uint32_t g = blockIdx.x;
  uint32_t t0 = g < ntiles ? tile_first[g] : 0, t1 = g < ntiles ? tile_first[g + 1] : 0;
  for (; g < ntiles; g += gridDim.x) {
    uint32_t t = t0 + threadIdx.x;
    uint64_t key_n = t < t1 ? __ldcs(part_key + t) : 0;  // h * CB_HOME_MUL: its top bits are the home slot
    uint32_t idx_n = t < t1 ? __ldcs(part_idx + t) : 0;
    const uint32_t g_n = g + gridDim.x;
    const uint32_t t0_n = g_n < ntiles ? tile_first[g_n] : 0, t1_n = g_n < ntiles ? tile_first[g_n + 1] : 0;
    for (uint32_t k = threadIdx.x; k < BUILD_TILE_SLOTS; k += BUILD_TILE_THREADS) reinterpret_cast<ulonglong2*>(tile)[k] = empty;
    if (threadIdx.x == 0) n_defer = 0;
    __syncthreads();

    // Pass 1, every key of the tile: claim the first empty slot of its probe run.  A key that meets its own
    // hash tag on the way is set aside (defer[t0 ...], the tile's own stretch of a scratch array).
    const uint32_t rounds = (t1 - t0 + BUILD_TILE_THREADS - 1) / BUILD_TILE_THREADS;
    for (uint32_t r = 0; r < rounds; r++, t += BUILD_TILE_THREADS) {  // warp-uniform trip counts, voted probe loops
      bool walking = t < t1;
      const uint64_t key = key_n;
      const uint64_t i = first + idx_n;
      if (t + BUILD_TILE_THREADS < t1) {
        key_n = __ldcs(part_key + t + BUILD_TILE_THREADS);
        idx_n = __ldcs(part_idx + t + BUILD_TILE_THREADS);
      }
      const uint64_t h = key * CB_HOME_INV;
      const uint32_t tag = slot_tag(h);
      const unsigned long long tagged = ((unsigned long long)tag << 32) | i;
      uint32_t slot = (uint32_t)(key >> (64 - tbits)) & (BUILD_TILE_SLOTS - 1);
      while (__any_sync(FULL, walking)) {
        const bool off = walking && slot >= BUILD_TILE_SLOTS;  // ran off the tile
        push_spill(off, t, spill, counters + CTR_SPILL);
        if (off) walking = false;
        if (walking) {
          unsigned long long* idxp = reinterpret_cast<unsigned long long*>(&tile[slot].idx);
          unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(idxp);
          if (cur == SLOT_EMPTY) {
            cur = atomicCAS(idxp, SLOT_EMPTY, tagged);
            if (cur == SLOT_EMPTY) {
              tile[slot].hash = h;
              walking = false;
            }
          }
          if (walking && (uint32_t)(cur >> 32) == tag) {
            defer[t0 + atomicAdd(&n_defer, 1u)] = t;
            // pass 2 will compare the two sequences: start their records on the way to L2 now
            asm volatile("prefetch.global.L2 [%0];" ::"l"(meta + i));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(meta + (uint32_t)cur));
            walking = false;
          }
          slot++;
        }
      }
    }
    __syncthreads();

    // Pass 2, the keys set aside, all at once: the full insert of build_kernel on the shared-memory tile.
    const uint32_t nd = n_defer;
    for (uint32_t k = threadIdx.x; k < (nd + BUILD_TILE_THREADS - 1) / BUILD_TILE_THREADS * BUILD_TILE_THREADS;
         k += BUILD_TILE_THREADS) {
      bool walking = k < nd;
      const uint32_t td = walking ? defer[t0 + k] : 0;
      const uint64_t key = walking ? part_key[td] : 0;
      const uint64_t i = walking ? first + part_idx[td] : 0;
      const uint64_t h = key * CB_HOME_INV;
      const uint32_t tag = slot_tag(h);
      const unsigned long long tagged = ((unsigned long long)tag << 32) | i;
      uint32_t slot = (uint32_t)(key >> (64 - tbits)) & (BUILD_TILE_SLOTS - 1);
      while (__any_sync(FULL, walking)) {
        const bool off = walking && slot >= BUILD_TILE_SLOTS;
        push_spill(off, td, spill, counters + CTR_SPILL);
        if (off) walking = false;
        if (walking) {
          unsigned long long* idxp = reinterpret_cast<unsigned long long*>(&tile[slot].idx);
          unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(idxp);
          if (cur == SLOT_EMPTY) {
            cur = atomicCAS(idxp, SLOT_EMPTY, tagged);
            if (cur == SLOT_EMPTY) {
              tile[slot].hash = h;
              walking = false;
            }
          }
          if (walking && (uint32_t)(cur >> 32) == tag) {  // same hash tag: compare the sequences
            const SeqMeta me = ld_meta_plain(meta + i), o = ld_meta_plain(meta + (uint32_t)cur);
            if (o.len == me.len && (ignore_genes || (o.v == me.v && o.j == me.j)) &&
                seq_equal(res, me.off, o.off, me.len)) {
              const unsigned long long old = atomicExch(idxp, tagged);
              meta[i].next = (uint32_t)old;
              walking = false;
            }
          }
          slot++;
        }
      }
    }
    __syncthreads();
    ulonglong2* out = reinterpret_cast<ulonglong2*>(table + (uint64_t)g * BUILD_TILE_SLOTS);
    for (uint32_t k = threadIdx.x; k < BUILD_TILE_SLOTS; k += BUILD_TILE_THREADS)
      __stcs(out + k, reinterpret_cast<const ulonglong2*>(tile)[k]);  // streamed: not to push the filter words out of L2
    __syncthreads();
    t0 = t0_n;
    t1 = t1_n;
  }
}

// part_key sorted on its top (tbits - BUILD_TILE_BITS) bits at least.  tile_first: ntiles + 1 words,
// defer, spill: n words each (all scratch).  counters[CTR_SPILL] must be zero on entry.
int launch_build_tiled(SeqRec* meta, const uint8_t* res, const uint64_t* part_key, const uint32_t* part_idx,
                       uint64_t first, uint64_t n, bool ignore_genes, Slot* table, uint32_t tbits, uint32_t* tile_first,
                       uint32_t* defer, uint32_t* spill, unsigned long long* counters, int sm_count, cudaStream_t st) {
  const uint32_t ntiles = 1u << (tbits - BUILD_TILE_BITS);
  const int smem = (int)(BUILD_TILE_SLOTS * sizeof(Slot));
  // on every launch: the attribute is per device, and one process may drive several (CLI --gpus N)
  cudaFuncSetAttribute(build_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  tile_bounds_kernel<<<(ntiles + 256) / 256, 256, 0, st>>>(part_key, n, 64 - (tbits - BUILD_TILE_BITS), ntiles, tile_first);
  const unsigned grid = (unsigned)std::min<uint64_t>(ntiles, (uint64_t)sm_count * CB_TILE_CTAS);
  build_tile_kernel<<<grid, BUILD_TILE_THREADS, smem, st>>>(meta, res, part_key, part_idx, tile_first, first, ignore_genes,
                                                           table, tbits, ntiles, defer, spill, counters);
  // the keys that left their tile (a fraction of a percent at the default load), by the direct insert
  build_kernel<<<(unsigned)sm_count * 4, 256, 0, st>>>(meta, res, nullptr, part_key, part_idx, first, n, ignore_genes, table,
                                                       (1ull << tbits) - 1, nullptr, 0, spill, counters + CTR_SPILL);
  return 3;
}

// The class filters of a large set, built apart from the table: one launch per (filter, range of its
// words), every launch a streaming pass over all hashes that sets the bits of the keys whose word
// falls into the range.  Random 8-byte REDs into a range that fits L2 run at L2 speed; the same
// REDs spread over four filters of 190 MiB each, interleaved with the table sweep, were DRAM
// sector read-modify-writes (measured inside build_kernel: +4 ms per filter at 10^8 keys).
// Duplicate keys set the same bits again: harmless.
template <int U>
__global__ void __launch_bounds__(256)
filter_kernel(const uint64_t* __restrict__ hash, uint64_t n, unsigned long long* bloom, uint32_t bloom_blocks,
              uint32_t cls, uint32_t w_lo, uint32_t w_hi) {
  // U independent streaming loads in flight per thread (evict-first: the hashes pass through L2
  // once, the word range stays)
  for (uint64_t base = (uint64_t)blockIdx.x * (256 * U); base < n; base += (uint64_t)gridDim.x * (256 * U)) {
    uint64_t h[U];
#pragma unroll
    for (int k = 0; k < U; k++) {
      const uint64_t i = base + (uint64_t)k * 256 + threadIdx.x;
      h[k] = i < n ? __ldcs(hash + i) : 0ull;
    }
#pragma unroll
    for (int k = 0; k < U; k++) {
      const uint32_t w = mulhi32(blind_field(h[k], cls), bloom_blocks);
      if (w >= w_lo && w < w_hi && base + (uint64_t)k * 256 + threadIdx.x < n)
        atomicOr(bloom + (uint64_t)cls * bloom_blocks + w, pfilter_pattern(h[k], cls));
    }
  }
}

int launch_filters(const uint64_t* hash, uint64_t n, unsigned long long* bloom, uint32_t bloom_blocks, int sm_count,
                   cudaStream_t st) {
  if (n == 0) return 0;
  // word ranges of at most ~48 MiB: resident in L2 beside the streamed hashes (whole build at 10^8 keys with
  // ranges of 24 / 48 / 64 / 96 / 200 MiB: 14.4 / 10.6 / 11.2 / 13.5 / 17.7 ms)
  // (COMPAIRR_B200_FILTER_PART_MIB: tuning knob for measurements)
  static const uint64_t part_mib = [] {
    const char* e = getenv("COMPAIRR_B200_FILTER_PART_MIB");
    const long v = e ? strtol(e, nullptr, 10) : 0;
    return (uint64_t)(v > 0 ? v : 48);
  }();
  const uint64_t bytes = (uint64_t)bloom_blocks * 8;
  uint32_t parts = (uint32_t)((bytes + (part_mib << 20) - 1) / (part_mib << 20));
  if (parts < 1) parts = 1;
  if (parts > 16) parts = 16;
  // four keys per thread per trip: 12.1 -> 11.2 ms for the whole build at 10^8 keys against one.  CTAs per SM: 8
  // (5, 4, 3, 2 tried to leave room for the table build running beside the passes: 10.9, 10.9, 11.2, 11.8 ms
  // against 10.6)
  constexpr int U = 4;
  const uint64_t blocks = (n + 256 * U - 1) / (256 * U);
  const unsigned grid = (unsigned)(blocks < (uint64_t)sm_count * 8 ? blocks : (uint64_t)sm_count * 8);
  for (uint32_t cls = 0; cls < CB_CLASSES; cls++)
    for (uint32_t k = 0; k < parts; k++) {
      const uint32_t lo = (uint32_t)((uint64_t)bloom_blocks * k / parts), hi = (uint32_t)((uint64_t)bloom_blocks * (k + 1) / parts);
      filter_kernel<U><<<grid, 256, 0, st>>>(hash, n, bloom, bloom_blocks, cls, lo, hi);
    }
  return (int)(CB_CLASSES * parts);
}

__global__ void __launch_bounds__(256) iota_kernel(uint32_t* p, uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    p[i] = (uint32_t)i;
}

// Input of the partition sort: key = h * CB_HOME_MUL (its top bits are the home slot), value = index.
__global__ void __launch_bounds__(256) partition_keys_kernel(const uint64_t* __restrict__ hash, uint64_t n,
                                                             uint64_t* __restrict__ key, uint32_t* __restrict__ idx) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    key[i] = hash[i] * CB_HOME_MUL;
    idx[i] = (uint32_t)i;
  }
}

void launch_partition_keys(const uint64_t* hash, uint64_t n, uint64_t* key, uint32_t* idx, cudaStream_t st) {
  if (n == 0) return;
  const uint64_t blocks = (n + 255) / 256;
  partition_keys_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>(hash, n, key, idx);
}

void launch_iota(uint32_t* p, uint64_t n, cudaStream_t st) {
  if (n == 0) return;
  const uint64_t blocks = (n + 255) / 256;
  iota_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>(p, n);
}

// Exact duplicates (overlap.cc:63-128, 865-873): sequence i is a duplicate iff another occurrence
// of the same (sequence, V, J) FURTHER DOWN its list has the same repertoire.  Summed over a
// group that is (members per repertoire - 1) per repertoire — the number the serial reference
// counts, independent of insertion order.
__global__ void __launch_bounds__(256) dups_kernel(DeviceSetView s, unsigned long long* counters) {
  uint32_t dups = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < s.n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint4 hi = __ldg(reinterpret_cast<const uint4*>(s.meta + i) + 1);  // v, j, rep, next
    const uint32_t rep = hi.z;
    uint32_t node = hi.w;
    while (node != SEQ_NIL) {
      const uint4 o = __ldg(reinterpret_cast<const uint4*>(s.meta + node) + 1);
      if (o.z == rep) {
        dups++;
        break;
      }
      node = o.w;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dups += __shfl_xor_sync(FULL, dups, o);
  if ((threadIdx.x & 31) == 0 && dups) atomicAdd(counters + CTR_DUPS, (unsigned long long)dups);
}

void launch_count_dups(DeviceSetView s, unsigned long long* counters, cudaStream_t st) {
  if (s.n == 0) return;
  const uint64_t blocks = (s.n + 255) / 256;
  dups_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>(s, counters);
}

// Deduplication (src/dedup.cc:62-137 process(), :27-59 report()): sequences with equal
// (repertoire, V, J unless -g, residues) form a group; the group is reported once, at its FIRST
// member in file order, with the summed count.  On the occurrence lists built by build_kernel every
// unordered pair of group members is seen exactly once — by whichever of the two sits higher up
// the list — so two atomicMin per pair leave lead[x] = smallest index of x's group.  The serial
// reference links each duplicate to the latest earlier one instead (next_seq[]), same groups.
__global__ void __launch_bounds__(256) dedup_lead_kernel(DeviceSetView s, uint32_t* lead) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < s.n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint4 hi = __ldg(reinterpret_cast<const uint4*>(s.meta + i) + 1);  // v, j, rep, next
    const uint32_t rep = hi.z;
    uint32_t node = hi.w, mine = (uint32_t)i;
    while (node != SEQ_NIL) {
      const uint4 o = __ldg(reinterpret_cast<const uint4*>(s.meta + node) + 1);
      if (o.z == rep) {
        atomicMin(lead + node, (uint32_t)i);
        mine = min(mine, node);
      }
      node = o.w;
    }
    if (mine != (uint32_t)i) atomicMin(lead + i, mine);
  }
}

// counts[lead[i]] += count of i (1 with -f, dedup.cc:34,39); counters[CTR_DUPS] += members merged away
__global__ void __launch_bounds__(256)
dedup_sum_kernel(DeviceSetView s, const uint32_t* __restrict__ lead, bool ignore_counts,
                 unsigned long long* sums, unsigned long long* counters) {
  uint32_t merged = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < s.n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t l = lead[i];
    const unsigned long long c = ignore_counts ? 1ull : (unsigned long long)s.meta[i].count;
    if (l == (uint32_t)i) {
      atomicAdd(sums + i, c);  // members may be adding to the same cell
    } else {
      atomicAdd(sums + l, c);
      merged++;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) merged += __shfl_xor_sync(FULL, merged, o);
  if ((threadIdx.x & 31) == 0 && merged) atomicAdd(counters + CTR_DUPS, (unsigned long long)merged);
}

void launch_dedup(DeviceSetView s, bool ignore_counts, uint32_t* lead, unsigned long long* sums,
                  unsigned long long* counters, cudaStream_t st) {
  if (s.n == 0) return;
  const uint64_t blocks = (s.n + 255) / 256;
  const unsigned grid = (unsigned)(blocks < 148 * 16 ? blocks : 148 * 16);
  iota_kernel<<<grid, 256, 0, st>>>(lead, s.n);
  dedup_lead_kernel<<<grid, 256, 0, st>>>(s, lead);
  dedup_sum_kernel<<<grid, 256, 0, st>>>(s, lead, ignore_counts, sums, counters);
}

// Bookkeeping: closed-form number of variants for a range of seeds (SURVEY section 8d "unit of work").
__global__ void __launch_bounds__(256)
count_probes_kernel(DeviceSetView a, uint64_t first, uint64_t count, uint32_t sigma, int d,
                    bool indels, unsigned long long* counters) {
  unsigned long long sum = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const SeqMeta m = ld_meta(a.meta + first + i);
    sum += probe_count(a.res + m.off, m.len, sigma, d, indels);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
  if ((threadIdx.x & 31) == 0 && sum) atomicAdd(counters + CTR_PROBES, sum);
}

void launch_count_probes(DeviceSetView a, uint64_t first, uint64_t count, uint32_t sigma, int d,
                         bool indels, unsigned long long* counters, cudaStream_t st) {
  if (count == 0) return;
  const uint64_t blocks = (count + 255) / 256;
  count_probes_kernel<<<(unsigned)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, st>>>(
      a, first, count, sigma, d, indels, counters);
}

// ---------------------------------------------------------------------------------------------
// K3/K4, d = 0: one thread per seed (one probe per seed: a hash join).
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256, 4) identical_kernel(const __grid_constant__ ProbeParams P) {
  extern __shared__ __align__(16) unsigned char tile_raw[];
  double* const tile = matrix_tile_begin(P, tile_raw);
  uint32_t nmatch = 0, npass = 0;
  const uint32_t lane = threadIdx.x & 31;
  // warp-uniform trip count: probe_chains() re-converges with warp-wide votes
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i0 = blockIdx.x * (uint64_t)blockDim.x + (threadIdx.x & ~31u); i0 < P.w_count;
       i0 += stride) {
    const bool in = i0 + lane < P.w_count;
    const uint64_t i = P.w_first + (in ? i0 + lane : 0);  // relative to a_first
    const uint64_t sidx = P.a_first + i;
    const uint64_t h = P.a.hash[sidx];
    bool walking = in;
    if (P.use_bloom && in) {
      walking = pfilter_test(P.bloom, P.bloom_blocks, h, 0);
    }
    npass += walking;
    nmatch += probe_chains(&P, walking, h, pack_var(VK_IDENTICAL, 0, 0, 0, 0), sidx, (uint32_t)i, tile);
  }
  matrix_tile_flush(P, tile);
  flush_counters(P, nmatch, P.count_bloom ? npass : 0);
}

int launch_variant_kernels(const ProbeParams& p, int sm_count, cudaStream_t st, const char** err);  // variant.cu

int launch_probe(const ProbeParams& p, int sm_count, cudaStream_t st, const char** err) {
  if (p.w_count == 0) return 0;
  if (p.differences == 0) {
    const uint64_t blocks = (p.w_count + 255) / 256;
    const size_t smem = (size_t)p.tile_cells * sizeof(double);
    if (smem > 48 * 1024) cudaFuncSetAttribute(identical_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const uint64_t cap = (uint64_t)sm_count * (smem ? (smem > 48 * 1024 ? 2 : 4) : 8);
    identical_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, smem, st>>>(p);
    return 1;
  }
  return launch_variant_kernels(p, sm_count, st, err);
}

}  // namespace cb
