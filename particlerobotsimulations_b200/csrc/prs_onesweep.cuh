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
 * Per tile (NT threads x 8 pairs; NT = 512 with several tiles resident per SM, or 1024):
 *   1. keys AND values are requested up front (the values are not needed before step 5, but
 *      their HBM latency would otherwise sit in the middle of the tile's critical path);
 *   2. keys are ranked warp by warp: the lanes holding the same digit find each other through a
 *      shared-memory word per (warp, digit) — atomicOr of the lane bit, read back — and the lowest
 *      lane bumps the warp's digit counter (stable: ranks follow lane order, then item order);
 *   3. per-warp counters are prefix-summed across warps, the tile's digit counts are published as
 *      AGGREGATE words;
 *   4. WIDE look-back: 32 or 64 predecessor words per digit are requested in one L2 round trip
 *      (2 or 4 lanes per digit, 16 words each) ...
 *   5. ... and while those loads are in flight the pairs are scattered into shared memory in
 *      tile-sorted order; then the predecessors' words are consumed (run of published words up to
 *      the first inclusive PREFIX; retry from the first unpublished one), the tile's PREFIX is
 *      published;
 *   6. the staged pairs are written out in digit runs, so global stores are coalesced.
 *
 * Tuning history (B200, %globaltimer stamps per phase, scripts/sort_timeline.py): MATCH.ANY and
 * 8-ballot matching made step 2 the longest phase (MATCH serialises over the distinct values of
 * a warp, ballots cost ~60 instructions per key); a 16-word look-back kept most in-flight tiles
 * in the aggregate-only state, so every tile walked back through all of them.
 *
 * HBM bytes per pair: 4 (histogram sweep) + 16 per digit pass.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace prs_sort {

/* optional per-tile phase stamps (nanoseconds, %globaltimer) for tuning: 8 words per tile */
__device__ __forceinline__ void stamp(unsigned long long *timeline, uint32_t tile, int phase) {
  if (timeline && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    timeline[(size_t)tile * 8 + phase] = t;
  }
}

/* Digit width and pairs per thread are template parameters of the pass kernel, chosen per sort by plan_passes():
 * 8-bit digits (256 bins) by default; 9-bit digits (512 bins) where they save a whole pass over the pairs — 26-bit cell
 * keys of the 8192^2 grid in (9, 9, 8) instead of four passes, 18-bit keys of the reference's 512^2 grid in (9, 9) instead
 * of three — a 512-bin pass costs ~10 % more than a 256-bin one, a pass saved is worth 100 %; 12 pairs per thread instead
 * of 8 for large inputs with 8-bit digits (fewer tiles: fewer look-backs and barriers per pair). */
constexpr int MAX_RADIX_BITS = 9;
constexpr int MAX_RADIX = 1 << MAX_RADIX_BITS;
constexpr int MAX_PASSES = 4;
constexpr int HIST_THREADS = 256;

constexpr uint32_t FLAG_AGG = 1u << 30;
constexpr uint32_t FLAG_PREFIX = 2u << 30;
constexpr uint32_t VALUE_MASK = (1u << 30) - 1;

template <int NT, int BITS, int IT>
struct Cfg {
  static_assert(BITS == 8 || BITS == 9, "supported digit widths");
  static constexpr int RADIX_BITS = BITS;
  static constexpr int RADIX = 1 << BITS;
  static constexpr int SCAN_WARPS = RADIX / 32;        /* warps whose threads hold one digit each in the block-wide scans */
  static constexpr int ITEMS = IT;
  static constexpr int THREADS = NT;
  static constexpr int WARPS = NT / 32;
  static constexpr int TILE = NT * ITEMS;
  static constexpr int GROUPS = NT / RADIX;            /* threads per digit in the cross-warp prefix: each sums a group of warps */
  static constexpr int WPG = WARPS / GROUPS;           /* warps per group: 8 (256 bins) or 16 (512 bins) */
  static constexpr int DIGITS_PER_WARP = RADIX / WARPS; /* look-back: digits owned by a warp */
  static constexpr int LANES_PER_DIGIT = 32 / DIGITS_PER_WARP;
  static_assert(GROUPS >= 1 && WPG * GROUPS == WARPS && (NT == 512 || NT == 1024), "supported tile shapes");
};

/* dynamic shared memory of k_onesweep (NT = 512, 8 bits, 8 items: 68 KB; NT = 1024: 133 KB) */
template <int NT, int BITS, int IT>
struct __align__(16) Smem {
  static constexpr int RADIX = Cfg<NT, BITS, IT>::RADIX;
  uint32_t cnt[Cfg<NT, BITS, IT>::WARPS][RADIX]; /* per-warp digit counters -> exclusive offsets across warps */
  uint32_t part[Cfg<NT, BITS, IT>::GROUPS][RADIX]; /* digit totals of each group of WPG warps */
  uint32_t tile_base[RADIX];          /* first slot of each digit inside the tile */
  uint32_t gbase[RADIX];              /* output position of slot 0 of each digit minus its tile slot */
  uint32_t count[RADIX];              /* this tile's digit counts (padding removed) */
  uint32_t mask[Cfg<NT, BITS, IT>::WARPS][RADIX]; /* (warp, digit) match words of the ranking loop */
  uint32_t keys[Cfg<NT, BITS, IT>::TILE];       /* staging of the tile in sorted order */
  uint32_t vals[Cfg<NT, BITS, IT>::TILE];
  uint32_t warp_tot[2][Cfg<NT, BITS, IT>::SCAN_WARPS];
  uint32_t tile;
};

/* Digit histograms of every pass in one sweep.  Cell keys of neighbouring robots share their
 * high digits, so a warp whose 32 keys agree on a digit adds 32 with one atomic; mixed warps use
 * plain shared-memory atomics (few-way conflicts). */
/* the digits of one sort: pass p takes `bits[p]` bits from bit `shift[p]` on (plan_passes) */
struct PassPlan {
  int npass;
  int shift[MAX_PASSES], bits[MAX_PASSES];
  int items; /* pairs per thread of the pass kernels */
};
/* 8-bit digits unless 9-bit ones save a pass; then as few 9-bit passes as cover the key */
inline PassPlan plan_passes(int key_bits, uint32_t n) {
  PassPlan P;
  const int n8 = (key_bits + 7) / 8, n9 = (key_bits + 8) / 9;
  P.npass = n9 < n8 ? n9 : n8;
  const int wide = n9 < n8 ? (key_bits - 8 * P.npass > 0 ? key_bits - 8 * P.npass : 0) : 0; /* passes that need the ninth bit */
  int at = 0;
  for (int p = 0; p < MAX_PASSES; p++) {
    P.shift[p] = at;
    P.bits[p] = (p < wide) ? 9 : 8;
    if (p < P.npass) at += P.bits[p];
  }
  P.items = (wide == 0 && n >= (1u << 22)) ? 12 : 8;
  return P;
}

__global__ void __launch_bounds__(HIST_THREADS) k_histogram(const uint32_t *__restrict__ keys, uint32_t n,
                                                            uint32_t *__restrict__ ghist, const PassPlan plan,
                                                            const uint32_t *__restrict__ n_dev) {
  if (n_dev) n = *n_dev;
  const int npass = plan.npass;
  __shared__ uint32_t sh[MAX_PASSES][MAX_RADIX];
  for (int i = threadIdx.x; i < MAX_PASSES * MAX_RADIX; i += HIST_THREADS) (&sh[0][0])[i] = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t n_round = (n + 31u) & ~31u;
  for (uint32_t i = blockIdx.x * HIST_THREADS + threadIdx.x; i < n_round; i += gridDim.x * HIST_THREADS) {
    const bool valid = i < n;
    const uint32_t k = valid ? keys[i] : 0u;
    const uint32_t active = __ballot_sync(0xffffffffu, valid);
    const int src = __ffs(active) - 1;
#pragma unroll
    for (int p = 0; p < MAX_PASSES; p++) { /* unrolled: the plan is read from the parameter bank with constant indices */
      if (p < npass) {
        const uint32_t d = (k >> plan.shift[p]) & ((1u << plan.bits[p]) - 1u);
        const uint32_t d0 = __shfl_sync(0xffffffffu, d, src);
        const bool uniform = __all_sync(0xffffffffu, !valid || d == d0);
        if (uniform) {
          if ((int)lane == src) atomicAdd(&sh[p][d0], (uint32_t)__popc(active));
        } else if (valid) {
          atomicAdd(&sh[p][d], 1u);
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int p = 0; p < MAX_PASSES; p++) {
    if (p < npass) {
      for (int d = threadIdx.x; d < (1 << plan.bits[p]); d += HIST_THREADS) {
        const uint32_t c = sh[p][d];
        if (c) atomicAdd(&ghist[p * MAX_RADIX + d], c);
      }
    }
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

/* exclusive scans over digits 0..RADIX-1 of TWO values held by threads 0..RADIX-1 (one set of barriers);
 * every thread of the block must call it */
template <int SCAN_WARPS>
__device__ __forceinline__ void excl_scan2(uint32_t &a, uint32_t &b, uint32_t (*s_tot)[SCAN_WARPS]) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t ta = 0, tb = 0, ea = 0, eb = 0;
  if (warp < SCAN_WARPS) {
    ea = warp_excl_scan(a, lane, &ta);
    eb = warp_excl_scan(b, lane, &tb);
    if (lane == 0) { s_tot[0][warp] = ta; s_tot[1][warp] = tb; }
  }
  __syncthreads();
  if (warp < SCAN_WARPS) {
#pragma unroll
    for (int w = 0; w < SCAN_WARPS; w++) {
      ea += (w < (int)warp) ? s_tot[0][w] : 0u;
      eb += (w < (int)warp) ? s_tot[1][w] : 0u;
    }
    a = ea;
    b = eb;
  }
}

/* One digit pass.  vin == nullptr means "values are the input positions" (first pass after
 * calcHash, where index[i] = i), which saves reading 4 B per pair. */
template <int NT, int BITS, int IT>
__global__ void __launch_bounds__(NT, (NT == 512) ? 2 : 1)
k_onesweep(const uint32_t *__restrict__ kin, const uint32_t *__restrict__ vin, uint32_t *__restrict__ kout,
           uint32_t *__restrict__ vout, uint32_t n, int shift, const uint32_t *__restrict__ ghist,
           volatile uint32_t *status, uint32_t *tile_counter, unsigned long long *timeline,
           const uint32_t *__restrict__ n_dev) {
  using C = Cfg<NT, BITS, IT>;
  if (n_dev) n = *n_dev; /* slab ranks: launched for the capacity, tiles past the real count exit */
  constexpr int THREADS = C::THREADS, WARPS = C::WARPS, TILE = C::TILE, GROUPS = C::GROUPS;
  constexpr int RADIX = C::RADIX, RADIX_BITS = C::RADIX_BITS, ITEMS = C::ITEMS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<NT, BITS, IT> &S = *reinterpret_cast<Smem<NT, BITS, IT> *>(smem_raw);

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) S.tile = atomicAdd(tile_counter, 1u);
  for (int i = tid; i < WARPS * RADIX; i += THREADS) { (&S.cnt[0][0])[i] = 0; (&S.mask[0][0])[i] = 0; }
  const uint32_t gh = (tid < RADIX) ? ghist[tid] : 0u; /* requested now, needed in step 3 */
  __syncthreads();
  const uint32_t tile = S.tile;
  stamp(timeline, tile, 0);
  const uint32_t tile_base = tile * (uint32_t)TILE;
  if (tile_base >= n) return; /* uniform for the block; only possible with a device-side count */
  const uint32_t tile_valid = min((uint32_t)TILE, n - tile_base);
  const uint32_t wbase = tile_base + warp * (ITEMS * 32);

  /* 1. keys and values */
  uint32_t key[ITEMS], val[ITEMS];
#pragma unroll
  for (int t = 0; t < ITEMS; t++) {
    const uint32_t i = wbase + t * 32 + lane;
    key[t] = (i < n) ? kin[i] : 0xffffffffu; /* padding sorts to the very end of the tile */
  }
#pragma unroll
  for (int t = 0; t < ITEMS; t++) {
    const uint32_t i = wbase + t * 32 + lane;
    val[t] = (i < n) ? (vin ? vin[i] : i) : 0u;
  }

  /* 2. EARLY COUNTS: per-warp digit counts with plain shared-memory atomics (order is irrelevant
   * for counting), so that the tile's aggregate can be published long before the stable ranking
   * is done — by the time the successors look back, it is there. */
#pragma unroll
  for (int t = 0; t < ITEMS; t++) {
    const uint32_t d = (key[t] >> shift) & (RADIX - 1);
    const uint32_t d0 = __shfl_sync(0xffffffffu, d, 0);
    if (__all_sync(0xffffffffu, d == d0)) {
      if (lane == 0) S.cnt[warp][d] += 32u; /* only this warp touches its counters */
    } else {
      atomicAdd(&S.cnt[warp][d], 1u);
    }
    __syncwarp();
  }
  __syncthreads();
  stamp(timeline, tile, 1);

  /* 3. warp counts -> exclusive offsets across the warps: thread (digit, group of WPG warps) */
  const uint32_t dg = tid & (RADIX - 1), grp = tid >> RADIX_BITS;
  {
    constexpr int WPG = C::WPG;
    uint32_t c[WPG], run = 0;
#pragma unroll
    for (int w = 0; w < WPG; w++) c[w] = S.cnt[grp * WPG + w][dg];
#pragma unroll
    for (int w = 0; w < WPG; w++) { const uint32_t cw = c[w]; c[w] = run; run += cw; }
    S.part[grp][dg] = run;
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (int g = 0; g < GROUPS; g++) {
      const uint32_t pg = S.part[g][dg];
      before += (g < (int)grp) ? pg : 0u;
      total += pg;
    }
    uint32_t tbase = 0, gex = 0;
    if (grp == 0) {
      uint32_t count = total;
      if (dg == ((0xffffffffu >> shift) & (RADIX - 1))) count -= (uint32_t)TILE - tile_valid; /* padding (key 0xffffffff: the last entries of the largest digit) is not data */
      S.count[dg] = count;
      if (tile != 0) status[(size_t)tile * RADIX + dg] = count | FLAG_AGG;
      tbase = total;
      gex = gh;
    }
    excl_scan2<C::SCAN_WARPS>(tbase, gex, S.warp_tot); /* digit start inside the tile / in the output */
    if (tid < RADIX) {
      S.tile_base[tid] = tbase;
      S.gbase[tid] = gex - tbase;
    }
    __syncthreads();
    /* cnt[w][d] becomes the first staging slot of warp w's keys with digit d */
    const uint32_t off = S.tile_base[dg] + before;
#pragma unroll
    for (int w = 0; w < WPG; w++) S.cnt[grp * WPG + w][dg] = off + c[w];
  }
  __syncthreads();
  stamp(timeline, tile, 2);

  /* 4. wide look-back, first trip requested here.  Warp w owns DIGITS_PER_WARP digits; lane =
   * (digit, part): part q fetches the 16 predecessors t-16q .. t-16q-15 of its digit. */
  constexpr int B = 16;
  constexpr int LPD = C::LANES_PER_DIGIT, DPW = C::DIGITS_PER_WARP;
  const uint32_t ld = warp * DPW + (lane & (DPW - 1)), lq = lane / DPW;
  uint32_t excl = 0;
  bool done = (tile == 0);
  int64_t t_next = (int64_t)tile - 1;
  uint32_t v[B];
  auto request = [&]() {
    const int64_t first = t_next - (int64_t)(B * lq);
#pragma unroll
    for (int i = 0; i < B; i++) {
      const int64_t tt = first - i;
      v[i] = (!done && tt >= 0) ? (uint32_t)status[(size_t)tt * RADIX + ld] : (uint32_t)(2u << 30); /* before tile 0: PREFIX 0 */
    }
  };
  request();

  /* 5. stable ranking straight into the staging area, while the predecessors' words are in
   * flight.  The lanes holding the same digit find each other through a shared-memory word per
   * (warp, digit): every lane ORs its lane bit in and reads the word back; the lowest lane (the
   * leader) clears it and advances the warp's slot cursor of that digit.  A warp whose 32 keys
   * share the digit — the common case for the high digits of cell keys — skips the exchange. */
  {
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int t = 0; t < ITEMS; t++) {
      const uint32_t d = (key[t] >> shift) & (RADIX - 1);
      const uint32_t d0 = __shfl_sync(0xffffffffu, d, 0);
      uint32_t m = 0xffffffffu;
      if (!__all_sync(0xffffffffu, d == d0)) {
        atomicOr(&S.mask[warp][d], 1u << lane);
        __syncwarp();
        m = *reinterpret_cast<volatile uint32_t *>(&S.mask[warp][d]);
        __syncwarp();
      }
      const int leader = __ffs(m) - 1;
      uint32_t slot = 0;
      if ((int)lane == leader) {
        S.mask[warp][d] = 0;
        slot = S.cnt[warp][d];
        S.cnt[warp][d] = slot + __popc(m);
      }
      slot = __shfl_sync(0xffffffffu, slot, leader) + __popc(m & lt_mask);
      S.keys[slot] = key[t];
      S.vals[slot] = val[t];
      __syncwarp();
    }
  }
  stamp(timeline, tile, 3);

  while (true) {
    uint32_t sum = 0, used = 0, st = 0; /* st: 0 ran through, 1 hit a PREFIX, 2 hit an unpublished word */
    uint32_t all_and = v[0], all_or = v[0];
#pragma unroll
    for (int i = 1; i < B; i++) { all_and &= v[i]; all_or |= v[i]; }
    if ((all_and >> 30) == 1u && (all_or >> 30) == 1u) { /* 16 aggregates: the common case */
#pragma unroll
      for (int i = 0; i < B; i++) sum += v[i];
      sum -= (uint32_t)B << 30;
      used = B;
    } else {
#pragma unroll
      for (int i = 0; i < B; i++) {
        const uint32_t f = v[i] >> 30;
        if (st == 0) {
          if (f == 0) st = 2;
          else { sum += v[i] & VALUE_MASK; used++; if (f == 2) st = 1; }
        }
      }
    }
    /* stitch the parts of each digit together in order */
    const uint32_t packed = used | (st << 8);
    uint32_t tot = 0, adv = 0, fin = 0;
#pragma unroll
    for (int q = 0; q < LPD; q++) {
      const uint32_t sq = __shfl_sync(0xffffffffu, sum, (lane & (DPW - 1)) + DPW * q);
      const uint32_t pq = __shfl_sync(0xffffffffu, packed, (lane & (DPW - 1)) + DPW * q);
      if (fin == 0) { tot += sq; adv += pq & 0xffu; fin = pq >> 8; }
    }
    if (!done) {
      excl += tot;
      t_next -= adv;
      if (fin == 1) done = true;
    }
    if (!__any_sync(0xffffffffu, !done)) break;
    request();
  }
  if (lq == 0) {
    status[(size_t)tile * RADIX + ld] = (excl + S.count[ld]) | FLAG_PREFIX;
    S.gbase[ld] += excl;
  }
  stamp(timeline, tile, 4);
  __syncthreads();
  stamp(timeline, tile, 5);

  /* 6. write out in digit runs */
  for (uint32_t j = tid; j < tile_valid; j += THREADS) {
    const uint32_t k = S.keys[j];
    const uint32_t d = (k >> shift) & (RADIX - 1);
    const uint32_t o = S.gbase[d] + j;
    kout[o] = k;
    vout[o] = S.vals[j];
  }
  stamp(timeline, tile, 6);
}

/* Persistent scratch of the sort: ping-pong pair buffers, histograms, tile status.  Grown on
 * demand, never freed per call (the reference pays a cudaMalloc/cudaFree per sort). */
struct Workspace {
  uint32_t *keys[2] = {nullptr, nullptr};
  uint32_t *vals[2] = {nullptr, nullptr};
  uint32_t *meta = nullptr; /* [hist 4*256][counters 4][status passes*tiles*256] */
  size_t cap_pairs = 0, cap_meta = 0;
};

/* meta: [histograms MAX_PASSES * MAX_RADIX][tile counters MAX_PASSES][status: passes * tiles * MAX_RADIX] */
inline size_t meta_words(uint32_t n, int npass, int tile_pairs) {
  const size_t tiles = (n + tile_pairs - 1) / tile_pairs;
  return (size_t)MAX_PASSES * MAX_RADIX + MAX_PASSES + (size_t)npass * tiles * MAX_RADIX;
}

}  // namespace prs_sort
