/*
 * prs_cellbin.cuh — cell binning: the fused step's route from robots to (sorted hash, sorted index,
 * cellStart, cellEnd) WITHOUT a multi-pass radix sort.  Included by prs_kernels.cu.
 *
 * What must come out (bit-exact): the pairs (hash, index) stably sorted by hash —
 * i.e. robots of one cell in ascending original index (thrust::sort_by_key, particlebot_cuda.cu:377) —
 * and the reference's cell table (cellStart = 0xffffffff for empty cells, cellEnd untouched for them,
 * kernel_impl.cuh:469-538).  The keys are cell numbers, so the sort is a counting sort whose
 * histogram IS the cell table:
 *
 *   K1  (k_control_integrate_hash<.., COUNT>)  hash h of each robot + arrival ticket within its cell:
 *        ticket = atomicAdd(&cellCount[h], 1)                   (order of arrival: arbitrary)
 *   scan (k_cell_tile_sums, k_cell_scan_tiles, k_cell_apply)  over the C cells:
 *        cellStart[c] = exclusive sum (or 0xffffffff), cellEnd[c] = start + count (non-empty cells
 *        only), cellCount[c] = 0 for the next step — this replaces the cudaMemset of the table
 *   scatter (k_cell_scatter)  robot i -> slot cellStart[h] + ticket: arrival-ordered index list
 *   K3  (k_reorder_binned)  slot k ranks its index among the few indices of its cell (ascending
 *        original index = the stable order), writes hash/index and gathers the packed sorted copy
 *
 * HBM bytes per robot: K1 +4 (ticket) ; scan 16*C/N ; scatter 8 + 4 (table lookup) + 8 ; K3 +8 —
 * against 4 + 16 per radix pass, and three dependent tile-chained passes fewer.  The general
 * onesweep sort (prs_onesweep.cuh) stays behind sortParticlebots / prs_sort_pairs, the per-call
 * path and the slab path, and is the fused path's route whenever binning is not admitted:
 * many more cells than robots, or crowded cells (the in-cell ranking is quadratic in the cell
 * population; the largest population is tracked on the device and reported to the host one step
 * late, see prs_fused_step).
 */
#pragma once
#include "prs_patchlist.cuh"

namespace prs_bin {

/* optional DENSE start table for collide's fresh-table steps: dense[c] = number of robots with a key below c for EVERY cell
 * (the reference's cellStart says 0xffffffff for an empty cell, so a stencil row's slot range needs the first and the last
 * non-empty of its five cells: 25 + 5 table words per robot; with the dense table it is dense[first cell], dense[last cell + 1]:
 * 10 words, one round trip).  `live` = per scan tile "a robot hashed into this tile or into one within `dil` tiles of it":
 * the tiles collide can look into this step; the others keep stale dense entries that nobody reads. */
struct DenseArgs {
  uint32_t *dense = nullptr, *live = nullptr;
  uint32_t dil = 0;
};

/* optional work list for k_collide_patch: patches of PATCH_W x PH cells that hold robots */
struct PatchListArgs {
  uint32_t *epoch_of = nullptr, *list = nullptr, *count = nullptr;
  uint32_t epoch = 0, log2_gx = 0, PH = 0;
};

/* Slab ranks (prs_slab.cuh): tile RANGE of the scan.  A slab owns whole grid rows, of which its robots may occupy a small
 * part (the first and the last slab of a swarm own every row down to / up to the grid's edge): `range` = {first tile, last
 * tile, violated} — the tiles (relative to the slab's first cell) between which the slab's sorted keys lay after the
 * previous sort; the scan processes those tiles widened by `dil` (a grid row of movement between two sorts + the reach of a
 * stencil) and skips the rest, whose table entries are empty and stay so.  Every ticket taken more than a grid row outside
 * the range (K1, the arrivals of the migration) sets `violated`, and the scan of that step processes every tile: always
 * correct, never dependent on the host. */
__device__ __forceinline__ uint32_t range_margin() { /* tiles a robot may move between two sorts without notice: one grid row */
  return max(1u, c_prm.p.gridSize.x / (uint32_t)(512 * 8));
}
__device__ __forceinline__ void range_check(uint32_t *range, uint32_t tile) {
  const uint32_t m = range_margin();
  if (tile + m < range[0] || tile > range[1] + m) atomicOr(range + 2, 1u);
}
struct RangeArgs {
  const uint32_t *range = nullptr;
  uint32_t dil = 0;
};
__device__ __forceinline__ bool range_skips(const RangeArgs &ra, uint32_t tile) {
  if (!ra.range) return false;
  const uint32_t lo = ra.range[0], hi = ra.range[1], violated = ra.range[2];
  if (violated) return false;
  if (lo > hi) return true; /* an empty slab */
  return tile + ra.dil < lo || tile > hi + ra.dil;
}

constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
constexpr uint32_t MAX_RANKED_CELL = 1024; /* populations above this are not ranked (error flag) */
/* tile marks ("a robot hashed into this scan tile in this step"): MARK_WAYS words per tile, the writer picks
 * one by its block number, so that the stores of a step spread over many L2 addresses (thousands of stores
 * to ONE word serialise in its L2 slice: measured +4 us on K1 at 2^20 robots) */
constexpr uint32_t MARK_WAYS = 32;
__device__ __forceinline__ bool tile_marked(const uint32_t *marks, uint32_t tile) {
  const uint32_t lane = threadIdx.x & 31u;
  return __any_sync(0xffffffffu, marks[tile * MARK_WAYS + lane] != 0u);
}

/* The scan over the C cells has no inter-block waiting (a single-pass chained scan was tried first:
 * with ~450 tiles in flight every tile spent most of its life waiting for its predecessors'
 * aggregates, 44 us for 4 M cells):
 *   k_cell_tile_sums   sums of tiles of 4096 cells
 *   k_cell_scan_tiles  one block: exclusive scan of the tile sums
 *   k_cell_apply       exclusive sums inside each tile + the tile's offset -> cellStart / cellEnd,
 *                      counters back to zero
 * scratch: [1] largest cell population, [2] error flag, [4..4+T) tile sums, [4+T..4+2T) tile offsets */
__device__ __forceinline__ void load_counts(const uint32_t *cellCount, uint32_t c0, uint32_t C, uint32_t (&cnt)[SCAN_ITEMS]) {
  if (c0 + SCAN_ITEMS <= C) {
    const uint4 a = *reinterpret_cast<const uint4 *>(cellCount + c0), b = *reinterpret_cast<const uint4 *>(cellCount + c0 + 4);
    cnt[0] = a.x; cnt[1] = a.y; cnt[2] = a.z; cnt[3] = a.w; cnt[4] = b.x; cnt[5] = b.y; cnt[6] = b.z; cnt[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) cnt[i] = (c0 + i < C) ? cellCount[c0 + i] : 0u;
  }
}

/* sums of the tiles of 4096 cells (no atomics: thousands of same-address atomics — a "last block
 * done" counter, per-robot tile sums from K1, a running maximum — each cost 10-100 us in L2) */
__global__ void __launch_bounds__(SCAN_THREADS)
k_cell_tile_sums(const uint32_t *__restrict__ cellCount, uint32_t C, uint32_t *scratch, const uint32_t *marks = nullptr,
                 const DenseArgs dn = DenseArgs(), const RangeArgs ra = RangeArgs()) {
  prs::pdl_sync();
  __shared__ uint32_t s_sum[SCAN_THREADS / 32];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  bool own = !range_skips(ra, blockIdx.x);
  if (marks && dn.live) {
    /* liveness for the dense table: this tile or one within dil tiles (cyclically: the hash wraps) was hashed into.  All mark
     * words are requested at once (one round trip), the first warp decides */
    uint32_t mine = marks[blockIdx.x * MARK_WAYS + lane], around = 0u;
    if (warp == 0) {
      for (uint32_t d = 1; d <= dn.dil; d++) {
        const uint32_t up = (blockIdx.x + d) % gridDim.x, down = (blockIdx.x + gridDim.x - d % gridDim.x) % gridDim.x;
        around |= marks[up * MARK_WAYS + lane] | marks[down * MARK_WAYS + lane];
      }
    }
    own = __any_sync(0xffffffffu, mine != 0u);
    if (warp == 0) {
      const bool alive = __any_sync(0xffffffffu, (mine | around) != 0u);
      if (lane == 0) dn.live[blockIdx.x] = alive ? 1u : 0u;
    }
  } else if (marks) {
    own = tile_marked(marks, blockIdx.x);
  }
  if (!own) { /* no robot hashed into this tile in this step: nothing to read */
    if (tid == 0) scratch[4 + blockIdx.x] = 0u;
    return;
  }
  uint32_t cnt[SCAN_ITEMS];
  load_counts(cellCount, blockIdx.x * SCAN_TILE + tid * SCAN_ITEMS, C, cnt);
  uint32_t sum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) sum += cnt[i];
  sum = __reduce_add_sync(0xffffffffu, sum);
  if (lane == 0) s_sum[warp] = sum;
  __syncthreads();
  if (tid == 0) {
    uint32_t total = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; w++) total += s_sum[w];
    scratch[4 + blockIdx.x] = total;
  }
}

/* exclusive scan of the T tile sums (one block) */
__global__ void __launch_bounds__(1024) k_cell_scan_tiles(uint32_t *scratch, uint32_t num_tiles) {
  prs::pdl_sync();
  __shared__ uint32_t s_sum[32];
  __shared__ uint32_t s_carry;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t *sums = scratch + 4, *offsets = scratch + 4 + num_tiles;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < num_tiles; base += 1024) {
    const uint32_t i = base + tid;
    const uint32_t v = (i < num_tiles) ? sums[i] : 0u;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= (uint32_t)o) inc += t;
    }
    if (lane == 31) s_sum[warp] = inc;
    __syncthreads();
    uint32_t before = s_carry, total = 0;
#pragma unroll
    for (int w = 0; w < 32; w++) {
      const uint32_t t = s_sum[w];
      before += (w < (int)warp) ? t : 0u;
      total += t;
    }
    if (i < num_tiles) offsets[i] = before + inc - v;
    __syncthreads();
    if (tid == 0) s_carry += total;
    __syncthreads();
  }
}

template <bool SELF_PREFIX>
__global__ void __launch_bounds__(SCAN_THREADS)
k_cell_apply(uint32_t *__restrict__ cellCount, uint32_t *__restrict__ cellStart, uint32_t *__restrict__ cellEnd, uint32_t C,
             uint32_t *scratch, uint32_t slot_offset, uint32_t *marks = nullptr, uint32_t *prev_marks = nullptr,
             const PatchListArgs pl = PatchListArgs(), const DenseArgs dn = DenseArgs(), const RangeArgs ra = RangeArgs()) {
  prs::pdl_sync();
  __shared__ uint32_t s_warp[SCAN_THREADS / 32];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t c0 = blockIdx.x * SCAN_TILE + tid * SCAN_ITEMS;
  if (range_skips(ra, blockIdx.x)) return; /* outside the slab's tile range: empty before, empty now */
  if (marks) {
    /* Tiles no robot hashed into (K1 marks the tile of every hash): their counters are all zero.  If the
     * tile was also empty in the previous step its cellStart words are 0xffffffff already and nothing is
     * touched; if it held robots then, only the empty markers are written.  (A world much larger than the
     * swarm — S1: 4 M cells, a third of the tiles occupied — otherwise pays for the whole table.) */
    /* with a dense table the tiles AROUND the hashed ones are processed too (their dense entries are read by collide) */
    const bool now = dn.live ? dn.live[blockIdx.x] != 0u : tile_marked(marks, blockIdx.x);
    const uint32_t before_ = prev_marks[blockIdx.x];
    __syncthreads(); /* every thread has read the words before they are rewritten */
    if (tid < MARK_WAYS) marks[blockIdx.x * MARK_WAYS + tid] = 0u;
    if (tid == 0) prev_marks[blockIdx.x] = now ? 1u : 0u;
    if (!now) {
      if (before_) {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++)
          if (c0 + i < C) cellStart[c0 + i] = 0xffffffffu;
      }
      return;
    }
  }
  /* offset of this tile: from k_cell_scan_tiles, or — few tiles (SELF_PREFIX) — summed here from the
   * tile sums, which saves the one-block scan kernel and its launch (5 us at 2^20 robots) */
  uint32_t tile_offset = slot_offset; /* slab ranks: slots start after the lower halo */
  if (SELF_PREFIX) {
    uint32_t part = 0;
    for (uint32_t t = tid; t < blockIdx.x; t += SCAN_THREADS) part += scratch[4 + t];
    part = __reduce_add_sync(0xffffffffu, part);
    __shared__ uint32_t s_part[SCAN_THREADS / 32];
    if (lane == 0) s_part[warp] = part;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; w++) tile_offset += s_part[w];
  } else {
    tile_offset += scratch[4 + gridDim.x + blockIdx.x];
  }
  uint32_t cnt[SCAN_ITEMS];
  load_counts(cellCount, c0, C, cnt);
  uint32_t sum = 0, mx = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) { sum += cnt[i]; mx = max(mx, cnt[i]); }
  /* my 8 cells are half a patch row: the patch holds robots if they do */
  if (pl.list) prs::patch_mark(sum != 0u, c0, pl.log2_gx, pl.PH, pl.epoch_of, pl.list, pl.count, pl.epoch);
  /* fullest cell (guard of the in-cell ranking): one atomic per warp at most, none once the running
   * maximum is reached — same-address traffic serialises in one L2 slice */
  mx = __reduce_max_sync(0xffffffffu, mx);
  if (lane == 0 && mx > 1u && mx > *reinterpret_cast<volatile const uint32_t *>(&scratch[1])) atomicMax(&scratch[1], mx);
  uint32_t inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (uint32_t)o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  uint32_t before = 0;
#pragma unroll
  for (int w = 0; w < SCAN_THREADS / 32; w++) before += (w < (int)warp) ? s_warp[w] : 0u;
  uint32_t run = tile_offset + before + (inc - sum);
  uint32_t st[SCAN_ITEMS], ds[SCAN_ITEMS];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    st[i] = cnt[i] ? run : 0xffffffffu;
    ds[i] = run;
    if (cnt[i] && c0 + i < C) cellEnd[c0 + i] = run + cnt[i]; /* empty cells keep their stale cellEnd (reference) */
    run += cnt[i];
  }
  if (dn.dense) {
    if (c0 + SCAN_ITEMS <= C) {
      *reinterpret_cast<uint4 *>(dn.dense + c0) = make_uint4(ds[0], ds[1], ds[2], ds[3]);
      *reinterpret_cast<uint4 *>(dn.dense + c0 + 4) = make_uint4(ds[4], ds[5], ds[6], ds[7]);
      if (c0 + SCAN_ITEMS == C) dn.dense[C] = run; /* one past the last cell: the robot count */
    } else {
#pragma unroll
      for (int i = 0; i < SCAN_ITEMS; i++)
        if (c0 + i <= C) dn.dense[c0 + i] = ds[i];
    }
  }
  if (c0 + SCAN_ITEMS <= C) {
    *reinterpret_cast<uint4 *>(cellStart + c0) = make_uint4(st[0], st[1], st[2], st[3]);
    *reinterpret_cast<uint4 *>(cellStart + c0 + 4) = make_uint4(st[4], st[5], st[6], st[7]);
    if (sum) { /* counters back to zero for the next step */
      *reinterpret_cast<uint4 *>(cellCount + c0) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4 *>(cellCount + c0 + 4) = make_uint4(0, 0, 0, 0);
    }
  } else {
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++)
      if (c0 + i < C) { cellStart[c0 + i] = st[i]; cellCount[c0 + i] = 0u; }
  }
}

/* robot i goes to slot cellStart[hash] + ticket of the arrival-ordered list */
__global__ void __launch_bounds__(256)
k_cell_scatter(const uint32_t *__restrict__ hash, const uint32_t *__restrict__ ticket, const uint32_t *__restrict__ cellStart,
               uint32_t *__restrict__ hash_by_slot, uint32_t *__restrict__ index_by_slot, uint32_t n) {
  prs::pdl_sync();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t h = hash[i];
  const uint32_t slot = __ldg(cellStart + h) + ticket[i];
  hash_by_slot[slot] = h;
  index_by_slot[slot] = i;
}

/* K3 of the binned route: stable order inside each cell + the packed sorted copy.
 * Slot k holds some robot `a` of cell h; its place in the stable order is cellStart[h] + (number
 * of robots of the cell with a smaller original index). */
__global__ void __launch_bounds__(256)
k_reorder_binned(const uint32_t *__restrict__ hash_by_slot, const uint32_t *__restrict__ index_by_slot,
                 const uint32_t *__restrict__ cellStart, const uint32_t *__restrict__ cellEnd, uint32_t *__restrict__ hash,
                 uint32_t *__restrict__ index, float4 *__restrict__ sortedPR, float2 *__restrict__ sortedVel,
                 const float2 *__restrict__ pos, const float2 *__restrict__ vel, const float *__restrict__ rad, uint32_t n,
                 uint32_t *scratch) {
  prs::pdl_sync();
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t h = hash_by_slot[k];
  const uint32_t a = index_by_slot[k];
  const float2 p = pos[a];
  const float2 v = vel[a];
  const float r = rad[a];
  const uint32_t s = __ldg(cellStart + h), e = __ldg(cellEnd + h);
  uint32_t below = 0;
  if (e - s <= MAX_RANKED_CELL) {
    for (uint32_t j = s; j < e; j++) below += (index_by_slot[j] < a) ? 1u : 0u;
  } else {
    atomicOr(&scratch[2], 1u); /* crowded beyond what the guard admits: the host fails loudly */
    below = k - s;
  }
  const uint32_t dst = s + below;
  hash[k] = h; /* every slot of the cell carries the cell's key */
  index[dst] = a;
  sortedPR[dst] = make_float4(p.x, p.y, r, __uint_as_float(a));
  sortedVel[dst] = v;
}

constexpr uint32_t SELF_PREFIX_MAX_TILES = 2048; /* up to 8 M cells: every apply block sums the tile sums before it */
inline size_t scan_scratch_words(uint32_t C) { return 4 + 2 * ((size_t)(C + SCAN_TILE - 1) / SCAN_TILE); }
constexpr int SCAN_TILE_LOG2 = 12;
static_assert((1 << SCAN_TILE_LOG2) == SCAN_TILE, "tile size");

}  // namespace prs_bin
