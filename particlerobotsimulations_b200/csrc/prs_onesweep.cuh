/*
 * prs_onesweep.cuh — hand-written stable LSD radix sort of (cell key, robot index) pairs for
 * sm_100a ("onesweep": one histogram sweep over the keys, then ONE read + ONE write of the pairs
 * per 8-bit digit, tiles chained by decoupled look-back instead of a separate scan kernel).
 *
 * Replaces the reference's `thrust::sort_by_key` (particlebot_cuda.cu:377-382; CUB radix sort with
 * per-call cudaMalloc/cudaFree and a blocking sync).  Semantics that must hold bit-exactly:
 * ascending by 32-bit key, STABLE (equal keys keep their input order, i.e. ascending robot index
 * after calcHash).  Only ceil(key_bits/8) digits are processed: cell keys are < numCells.
 *
 * Per tile (512 threads x 8 keys): keys are ranked warp by warp with __match_any_sync (lanes
 * holding the same digit elect a leader that bumps a per-warp digit counter in shared memory),
 * per-warp counters are prefix-summed across warps by one thread per digit, the tile's digit
 * counts are published as AGGREGATE, the exclusive prefix over earlier tiles is fetched by
 * look-back (one thread per digit), the pairs are staged in shared memory in tile-sorted order
 * and written out in runs, so global stores are coalesced.
 *
 * HBM bytes per pair: 4 (histogram sweep) + 16 per digit pass.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace prs_sort {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int THREADS = 512;
constexpr int WARPS = THREADS / 32;
constexpr int ITEMS = 8;
constexpr int TILE = THREADS * ITEMS; /* 4096 pairs per tile, two tiles resident per SM: short per-warp dependency chains */
constexpr int MAX_PASSES = 4;
constexpr int HIST_THREADS = 256;

constexpr uint32_t FLAG_AGG = 1u << 30;
constexpr uint32_t FLAG_PREFIX = 2u << 30;
constexpr uint32_t VALUE_MASK = (1u << 30) - 1;

/* dynamic shared memory of k_onesweep */
struct __align__(16) Smem {
  uint32_t cnt[WARPS][RADIX];   /* per-warp digit counters -> exclusive offsets across warps */
  uint32_t tile_base[RADIX];    /* first slot of each digit inside the tile */
  uint32_t gbase[RADIX];        /* output position of slot 0 of each digit minus its tile slot */
  uint32_t keys[TILE];
  uint32_t vals[TILE];
  uint32_t warp_tot[WARPS];
  uint32_t tile;
};

/* Digit histograms of every pass in one sweep.  Cell keys of neighbouring robots share their
 * high digits, so a warp whose 32 keys agree on a digit adds 32 with one atomic; mixed warps use
 * plain shared-memory atomics (few-way conflicts). */
__global__ void __launch_bounds__(HIST_THREADS) k_histogram(const uint32_t *__restrict__ keys, uint32_t n,
                                                            uint32_t *__restrict__ ghist, int npass) {
  __shared__ uint32_t sh[MAX_PASSES][RADIX];
  for (int i = threadIdx.x; i < MAX_PASSES * RADIX; i += HIST_THREADS) (&sh[0][0])[i] = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t n_round = (n + 31u) & ~31u;
  for (uint32_t i = blockIdx.x * HIST_THREADS + threadIdx.x; i < n_round; i += gridDim.x * HIST_THREADS) {
    const bool valid = i < n;
    const uint32_t k = valid ? keys[i] : 0u;
    const uint32_t active = __ballot_sync(0xffffffffu, valid);
    const int src = __ffs(active) - 1;
    for (int p = 0; p < npass; p++) {
      const uint32_t d = (k >> (p * RADIX_BITS)) & (RADIX - 1);
      const uint32_t d0 = __shfl_sync(0xffffffffu, d, src);
      const bool uniform = __all_sync(0xffffffffu, !valid || d == d0);
      if (uniform) {
        if ((int)lane == src) atomicAdd(&sh[p][d0], (uint32_t)__popc(active));
      } else if (valid) {
        atomicAdd(&sh[p][d], 1u);
      }
    }
  }
  __syncthreads();
  for (int p = 0; p < npass; p++) {
    const uint32_t c = sh[p][threadIdx.x];
    if (c) atomicAdd(&ghist[p * RADIX + threadIdx.x], c);
  }
}

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, uint32_t lane, uint32_t *total) {
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (uint32_t)o) inc += t;
  }
  *total = __shfl_sync(0xffffffffu, inc, 31);
  return inc - v;
}

/* exclusive scan over the block of one value per thread (threads >= RADIX pass 0) */
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *s_warp_tot) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t tot;
  uint32_t ex = warp_excl_scan(v, lane, &tot);
  if (lane == 0) s_warp_tot[warp] = tot;
  __syncthreads();
  uint32_t add = 0;
#pragma unroll
  for (int w = 0; w < WARPS; w++) add += (w < (int)warp) ? s_warp_tot[w] : 0u;
  __syncthreads();
  return ex + add;
}

/* One digit pass.  vin == nullptr means "values are the input positions" (first pass after
 * calcHash, where index[i] = i), which saves reading 4 B per pair. */
__global__ void __launch_bounds__(THREADS, 2)
k_onesweep(const uint32_t *__restrict__ kin, const uint32_t *__restrict__ vin, uint32_t *__restrict__ kout,
           uint32_t *__restrict__ vout, uint32_t n, int shift, const uint32_t *__restrict__ ghist,
           volatile uint32_t *status, uint32_t *tile_counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem &S = *reinterpret_cast<Smem *>(smem_raw);

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) S.tile = atomicAdd(tile_counter, 1u);
  for (int i = tid; i < WARPS * RADIX; i += THREADS) (&S.cnt[0][0])[i] = 0;
  __syncthreads();
  const uint32_t tile = S.tile;
  const uint32_t tile_base = tile * (uint32_t)TILE;
  const uint32_t tile_valid = min((uint32_t)TILE, n - tile_base);
  const uint32_t wbase = tile_base + warp * (ITEMS * 32);

  uint32_t key[ITEMS];
  uint16_t rank[ITEMS];
#pragma unroll
  for (int t = 0; t < ITEMS; t++) {
    const uint32_t i = wbase + t * 32 + lane;
    key[t] = (i < n) ? kin[i] : 0xffffffffu; /* padding sorts to the very end of the tile */
  }
  /* stable ranks inside the warp's 512 keys: lanes holding the same digit are found with
   * match_any (one vote when the whole warp agrees), their leader bumps the warp's counter */
  const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int t = 0; t < ITEMS; t++) {
    const uint32_t d = (key[t] >> shift) & (RADIX - 1);
    const uint32_t d0 = __shfl_sync(0xffffffffu, d, 0);
    uint32_t m = 0xffffffffu;
    if (!__all_sync(0xffffffffu, d == d0)) m = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(m) - 1;
    uint32_t prev = 0;
    if ((int)lane == leader) {
      prev = S.cnt[warp][d];
      S.cnt[warp][d] = prev + __popc(m);
    }
    prev = __shfl_sync(0xffffffffu, prev, leader);
    rank[t] = (uint16_t)(prev + __popc(m & lt_mask));
    __syncwarp();
  }
  __syncthreads();

  /* thread d < 256 owns digit d: warp counts -> exclusive offsets across warps */
  uint32_t total = 0, count = 0;
  if (tid < RADIX) {
#pragma unroll
    for (int w = 0; w < WARPS; w++) {
      const uint32_t cw = S.cnt[w][tid];
      S.cnt[w][tid] = total;
      total += cw;
    }
    count = total;
    if (tid == RADIX - 1) count -= (uint32_t)TILE - tile_valid; /* padding is not data */
    if (tile != 0) status[(size_t)tile * RADIX + tid] = count | FLAG_AGG;
  }
  const uint32_t tbase = block_excl_scan(total, S.warp_tot);                               /* digit start inside the tile */
  const uint32_t gex = block_excl_scan(tid < RADIX ? ghist[tid] : 0u, S.warp_tot);         /* digit start in the output */

  /* Look-back, one thread per digit.  Up to a few hundred tiles are resident at once and finish
   * ranking together, so the walk back to the nearest published PREFIX is long; the
   * predecessors' words are therefore fetched LOOKBACK at a time (independent volatile loads in
   * flight together) and consumed in order, stopping at the first word not published yet. */
  if (tid < RADIX) {
    constexpr int LOOKBACK = 16;
    uint32_t excl = 0;
    if (tile != 0) {
      int64_t t = (int64_t)tile - 1;
      bool done = false;
      while (!done) {
        uint32_t v[LOOKBACK];
#pragma unroll
        for (int i = 0; i < LOOKBACK; i++) {
          const int64_t tt = t - i;
          v[i] = (tt >= 0) ? status[(size_t)tt * RADIX + tid] : (uint32_t)(2u << 30); /* before tile 0: PREFIX 0 */
        }
        int used = 0;
#pragma unroll
        for (int i = 0; i < LOOKBACK; i++) {
          const uint32_t f = v[i] & ~VALUE_MASK;
          if (done || used != i || f == 0) continue; /* consume strictly in order */
          excl += v[i] & VALUE_MASK;
          used = i + 1;
          if (f == (uint32_t)(2u << 30)) done = true;
        }
        t -= used;
      }
    }
    status[(size_t)tile * RADIX + tid] = (excl + count) | FLAG_PREFIX;
    S.tile_base[tid] = tbase;
    S.gbase[tid] = gex + excl - tbase;
  }
  __syncthreads();

#pragma unroll
  for (int t = 0; t < ITEMS; t++) {
    const uint32_t d = (key[t] >> shift) & (RADIX - 1);
    const uint32_t p = S.tile_base[d] + S.cnt[warp][d] + rank[t];
    const uint32_t i = wbase + t * 32 + lane;
    S.keys[p] = key[t];
    S.vals[p] = (i < n) ? (vin ? vin[i] : i) : 0u;
  }
  __syncthreads();
  for (uint32_t j = tid; j < tile_valid; j += THREADS) {
    const uint32_t k = S.keys[j];
    const uint32_t d = (k >> shift) & (RADIX - 1);
    const uint32_t o = S.gbase[d] + j;
    kout[o] = k;
    vout[o] = S.vals[j];
  }
}

/* Persistent scratch of the sort: ping-pong pair buffers, histograms, tile status.  Grown on
 * demand, never freed per call (the reference pays a cudaMalloc/cudaFree per sort). */
struct Workspace {
  uint32_t *keys[2] = {nullptr, nullptr};
  uint32_t *vals[2] = {nullptr, nullptr};
  uint32_t *meta = nullptr; /* [hist 4*256][counters 4][status passes*tiles*256] */
  size_t cap_pairs = 0, cap_meta = 0;
};

inline size_t meta_words(uint32_t n, int npass) {
  const size_t tiles = (n + TILE - 1) / TILE;
  return (size_t)MAX_PASSES * RADIX + MAX_PASSES + (size_t)npass * tiles * RADIX;
}

}  // namespace prs_sort
