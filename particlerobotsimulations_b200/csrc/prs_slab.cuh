/*
 * prs_slab.cuh — slab (multi-GPU) engine: the fused step cut at the two points where ranks
 * exchange robots, with every count kept ON THE DEVICE (DESIGN.md "multi-GPU").
 * Included by prs_kernels.cu; the kernels are C++ (file-local), only the prs_slab_* / prs_ipc_* entry points are extern "C".
 *
 * A rank owns the grid rows [row_lo, row_hi).  Nothing in a step needs the host to know how many
 * robots it owns, how many left, or how long the halos are: kernels are launched for the slab's
 * CAPACITY and read the live counts from `counts` (uint32[16], PRS_SC_*); the neighbour exchange
 * moves FIXED-SIZE buffers whose first word is the record count.  So a step is a pure stream of
 * kernel launches and NCCL sends/receives — no host synchronisation, no device-to-host copy.
 *
 *   K1 -> [sort steps: migrate_pack -> exchange -> migrate_unpack -> sort (+ ties by global id)]
 *      -> gather -> halo_pack -> exchange -> halo_unpack -> cell_table -> collide
 *
 * The sort is the onesweep radix sort followed by an in-cell insertion sort by global id, or —
 * while the swarm is known to be sparse (same asynchronous guard as the fused single-GPU step) —
 * cell binning (prs_cellbin.cuh): tickets, a scan over the owned rows' cells that IS their cell
 * table, scatter, and the in-cell order by global id folded into the gather.  Same results.
 *
 * Migration record (23 words, structure-of-arrays with stride mig_cap after the count word):
 *   pos.xy vel.xy rad phase absForce_a absForce_r dead gid hash rng[12]
 * Halo record (7 words, stride halo_cap): sortedPR.xyzw sortedVel.xy hash
 * Errors that the host must hear about (capacity overflow, a robot crossing more than one slab)
 * set bits in counts[PRS_SC_ERR]; the host polls it asynchronously.
 */
#pragma once

#define SLAB_MIG_WORDS 23
#define SLAB_HALO_WORDS 7

/* -------------------------------------------------------------------------------------------- */
/* migration                                                                                      */
/* -------------------------------------------------------------------------------------------- */
__device__ __forceinline__ void slab_store_record(uint32_t *buf, uint32_t stride, uint32_t q, const prs_slab &s, uint32_t i) {
  uint32_t *o = buf + 1 + q;
  const float2 p = ((const float2 *)s.pos)[i], v = ((const float2 *)s.vel)[i];
  o[0 * stride] = __float_as_uint(p.x); o[1 * stride] = __float_as_uint(p.y);
  o[2 * stride] = __float_as_uint(v.x); o[3 * stride] = __float_as_uint(v.y);
  o[4 * stride] = __float_as_uint(s.rad[i]); o[5 * stride] = __float_as_uint(s.phase[i]);
  o[6 * stride] = __float_as_uint(s.absForce_a[i]); o[7 * stride] = __float_as_uint(s.absForce_r[i]);
  o[8 * stride] = (uint32_t)s.dead[i]; o[9 * stride] = s.gid[i]; o[10 * stride] = s.hash[i];
  const uint32_t *r = (const uint32_t *)s.rng + (size_t)i * 12;
#pragma unroll
  for (int w = 0; w < 12; w++) o[(11 + w) * stride] = r[w];
}
__device__ __forceinline__ void slab_load_record(const uint32_t *buf, uint32_t stride, uint32_t q, const prs_slab &s, uint32_t i) {
  const uint32_t *o = buf + 1 + q;
  ((float2 *)s.pos)[i] = make_float2(__uint_as_float(o[0 * stride]), __uint_as_float(o[1 * stride]));
  ((float2 *)s.vel)[i] = make_float2(__uint_as_float(o[2 * stride]), __uint_as_float(o[3 * stride]));
  s.rad[i] = __uint_as_float(o[4 * stride]); s.phase[i] = __uint_as_float(o[5 * stride]);
  s.absForce_a[i] = __uint_as_float(o[6 * stride]); s.absForce_r[i] = __uint_as_float(o[7 * stride]);
  s.dead[i] = (int)o[8 * stride]; s.gid[i] = o[9 * stride]; s.hash[i] = o[10 * stride];
  uint32_t *r = (uint32_t *)s.rng + (size_t)i * 12;
#pragma unroll
  for (int w = 0; w < 12; w++) r[w] = o[(11 + w) * stride];
}
__device__ __forceinline__ void slab_move_record(const prs_slab &s, uint32_t dst, uint32_t src) {
  ((float2 *)s.pos)[dst] = ((const float2 *)s.pos)[src];
  ((float2 *)s.vel)[dst] = ((const float2 *)s.vel)[src];
  s.rad[dst] = s.rad[src]; s.phase[dst] = s.phase[src];
  s.absForce_a[dst] = s.absForce_a[src]; s.absForce_r[dst] = s.absForce_r[src];
  s.dead[dst] = s.dead[src]; s.gid[dst] = s.gid[src]; s.hash[dst] = s.hash[src];
  const uint4 *a = (const uint4 *)((const uint32_t *)s.rng + (size_t)src * 12);
  uint4 *b = (uint4 *)((uint32_t *)s.rng + (size_t)dst * 12);
  b[0] = a[0]; b[1] = a[1]; b[2] = a[2];
}

/* resets the per-step counters (everything except n and the statistics) */
__global__ void k_slab_begin_step(uint32_t *counts, uint32_t *range) {
  prs::pdl_sync();
  const int i = threadIdx.x;
  if (i >= PRS_SC_NLO && i <= PRS_SC_KEEPERS) counts[i] = 0;
  if (i == 0 && range) range[2] = 0u; /* "a ticket was taken outside the scan's tile range": per step */
}

/* M1: robots whose new grid row left [row_lo, row_hi) are packed for the neighbour that owns it
 * and listed; `scratch[i]` = 1 marks a leaver */
__global__ void __launch_bounds__(256) k_slab_select(prs_slab s, uint32_t *__restrict__ send_dn, uint32_t *__restrict__ send_up,
                                                     uint32_t log2_gx) {
  prs::pdl_sync();
  const uint32_t n = s.counts[PRS_SC_N];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t row = s.hash[i] >> log2_gx;
  bool dn = row < s.row_lo, up = row >= s.row_hi;
  if (s.wrap && (dn || up)) {
    /* ring of slabs: the row lies on the side it is cyclically nearer to (a robot leaving the top row of the grid re-enters at
     * row 0, which belongs to the first slab: "up" from the last one) */
    const uint32_t gy_mask = c_prm.p.gridSize.y - 1u;
    const uint32_t d_up = (row - s.row_hi) & gy_mask, d_dn = (s.row_lo - 1u - row) & gy_mask;
    up = d_up <= d_dn;
    dn = !up;
  }
  s.scratch[i] = (dn || up) ? 1u : 0u;
  if (!(dn || up)) return;
  if ((dn && !s.has_dn) || (up && !s.has_up)) { atomicOr(&s.counts[PRS_SC_ERR], PRS_SLAB_ERR_LEFT_WORLD); s.scratch[i] = 0u; return; }
  const uint32_t q = atomicAdd(&s.counts[dn ? PRS_SC_MIGDN : PRS_SC_MIGUP], 1u);
  if (q >= s.mig_cap) { atomicOr(&s.counts[PRS_SC_ERR], PRS_SLAB_ERR_MIG_CAP); s.scratch[i] = 0u; return; }
  slab_store_record(dn ? send_dn : send_up, s.mig_cap, q, s, i);
  s.lists[atomicAdd(&s.counts[PRS_SC_LEAVERS], 1u)] = i;
}
/* M2: with L leavers the survivors must end up in [0, n-L).  Holes = leavers below n-L, movers =
 * survivors at or above it; both lists have the same length. */
__global__ void __launch_bounds__(256) k_slab_list_holes(prs_slab s, uint32_t *__restrict__ send_dn, uint32_t *__restrict__ send_up) {
  prs::pdl_sync();
  const uint32_t n = s.counts[PRS_SC_N], L = s.counts[PRS_SC_LEAVERS];
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q == 0) { /* the count words of the outgoing buffers */
    send_dn[0] = min(s.counts[PRS_SC_MIGDN], s.mig_cap);
    send_up[0] = min(s.counts[PRS_SC_MIGUP], s.mig_cap);
  }
  if (q >= L) return;
  const uint32_t new_n = n - L;
  uint32_t *holes = s.lists + 2 * s.mig_cap, *movers = s.lists + 4 * s.mig_cap;
  const uint32_t li = s.lists[q];
  if (li < new_n) holes[atomicAdd(&s.counts[PRS_SC_HOLES], 1u)] = li;
  const uint32_t j = new_n + q; /* the L slots of the tail */
  if (!s.scratch[j]) movers[atomicAdd(&s.counts[PRS_SC_KEEPERS], 1u)] = j;
}
/* M3: movers fill the holes; then the arrivals are appended and n is updated */
__global__ void __launch_bounds__(256) k_slab_fill_holes(prs_slab s, uint32_t *__restrict__ ticket) {
  prs::pdl_sync();
  const uint32_t H = s.counts[PRS_SC_HOLES];
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= H) return;
  const uint32_t dst = s.lists[2 * s.mig_cap + q], src = s.lists[4 * s.mig_cap + q];
  slab_move_record(s, dst, src);
  if (ticket) ticket[dst] = ticket[src]; /* the cell ticket K1 took travels with the robot */
}
__global__ void __launch_bounds__(256) k_slab_append(prs_slab s, const uint32_t *__restrict__ recv_dn, const uint32_t *__restrict__ recv_up,
                                                     uint32_t log2_gx, uint32_t *__restrict__ cellCount, uint32_t *__restrict__ ticket,
                                                     uint32_t *range, uint32_t range_tile0) {
  prs::pdl_sync();
  const uint32_t n = s.counts[PRS_SC_N], L = s.counts[PRS_SC_LEAVERS];
  const uint32_t c_dn = s.has_dn ? min(recv_dn[0], s.mig_cap) : 0u, c_up = s.has_up ? min(recv_up[0], s.mig_cap) : 0u;
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= c_dn + c_up) return;
  const uint32_t dst = n - L + q;
  if (dst >= s.cap) { atomicOr(&s.counts[PRS_SC_ERR], PRS_SLAB_ERR_CAPACITY); return; }
  if (q < c_dn) slab_load_record(recv_dn, s.mig_cap, q, s, dst);
  else slab_load_record(recv_up, s.mig_cap, q - c_dn, s, dst);
  const uint32_t h = s.hash[dst];
  const uint32_t row = h >> log2_gx; /* a robot may cross one slab per sort at most */
  const bool mine = row >= s.row_lo && row < s.row_hi;
  if (!mine) atomicOr(&s.counts[PRS_SC_ERR], PRS_SLAB_ERR_TWO_SLABS);
  if (ticket) ticket[dst] = mine ? atomicAdd(&cellCount[h], 1u) : 0xffffffffu; /* binned sort: the arrival's cell ticket */
  if (ticket && mine && range) prs_bin::range_check(range, h / prs_bin::SCAN_TILE - range_tile0);
}
__global__ void k_slab_commit_count(prs_slab s, const uint32_t *__restrict__ recv_dn, const uint32_t *__restrict__ recv_up) {
  prs::pdl_sync();
  const uint32_t c_dn = s.has_dn ? min(recv_dn[0], s.mig_cap) : 0u, c_up = s.has_up ? min(recv_up[0], s.mig_cap) : 0u;
  const uint32_t L = s.counts[PRS_SC_LEAVERS];
  uint32_t n = s.counts[PRS_SC_N] - L + c_dn + c_up;
  if (n > s.cap) n = s.cap;
  s.counts[PRS_SC_N] = n;
  s.counts[PRS_SC_STAT_MIG] += L;
}

/* -------------------------------------------------------------------------------------------- */
/* sorted view: gather, ties, halo, cell table                                                    */
/* -------------------------------------------------------------------------------------------- */
/* Slab ranks hold their robots in arbitrary local slots, but the reference's stable sort leaves the
 * robots of one cell in ascending ORIGINAL index.  After the local sort (ties by local slot) the
 * thread at each cell start insertion-sorts that cell's few entries by global id, so forces are
 * summed in exactly the single-GPU order (bit-equal results across any number of slabs). */
__global__ void __launch_bounds__(256)
k_fix_ties_by_gid(const uint32_t *__restrict__ hash, uint32_t *__restrict__ index, const uint32_t *__restrict__ gid,
                  const uint32_t *__restrict__ n_dev) {
  prs::pdl_sync();
  const uint32_t n = *n_dev;
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t h = hash[k];
  if (k > 0 && hash[k - 1] == h) return; /* not a cell start */
  uint32_t e = k + 1;
  while (e < n && hash[e] == h) e++;
  for (uint32_t a = k + 1; a < e; a++) {
    const uint32_t slot = index[a];
    const uint32_t g = gid[slot];
    uint32_t b = a;
    while (b > k && gid[index[b - 1]] > g) { index[b] = index[b - 1]; b--; }
    index[b] = slot;
  }
}
/* Between two sorts the table is frozen while robots move: collide looks up the stencil around a
 * robot's CURRENT cell, and this rank only holds the rows [row_lo - halo_rows, row_hi + halo_rows).  An owned
 * robot whose current row has left [row_lo - (halo_rows - 2), row_hi + (halo_rows - 2)) would silently miss
 * neighbours the single-GPU run sees: flag it (sticky), the host hears about it through check(). */
__device__ __forceinline__ void slab_drift_check(const prs_slab &s, float y) {
  const uint32_t gy = c_prm.p.gridSize.y;
  const uint32_t row = (uint32_t)((int)floorf((y - c_prm.p.worldOrigin.y) / c_prm.p.cellSize.y)) & (gy - 1u);
  const uint32_t slack = s.halo_rows >= 2u ? s.halo_rows - 2u : 0u;
  bool bad;
  if (s.wrap) { /* cyclic distance beyond the owned rows on either side */
    const uint32_t d_up = (row - s.row_hi) & (gy - 1u), d_dn = (s.row_lo - 1u - row) & (gy - 1u);
    const bool owned = row >= s.row_lo && row < s.row_hi;
    bad = !owned && min(d_up, d_dn) >= slack;
  } else {
    bad = (s.has_dn && row + slack < s.row_lo) || (s.has_up && row >= s.row_hi + slack);
  }
  if (bad) atomicOr(&s.counts[PRS_SC_ERR], PRS_SLAB_ERR_DRIFT);
}
/* packed sorted copy of the owned robots at [halo_cap, halo_cap + n); pr.w = GLOBAL id (the robot's identity: the
 * transported object is robot nCells - 1 wherever it lives); collide scatters its results through index_sorted */
__global__ void __launch_bounds__(256) k_slab_gather(prs_slab s) {
  prs::pdl_sync();
  const uint32_t n = s.counts[PRS_SC_N];
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t src = s.index_sorted[k];
  const float2 p = ((const float2 *)s.pos)[src];
  slab_drift_check(s, p.y);
  ((float4 *)s.sortedPR)[s.halo_cap + k] = make_float4(p.x, p.y, s.rad[src], __uint_as_float(s.gid[src]));
  ((float2 *)s.sortedVel)[s.halo_cap + k] = ((const float2 *)s.vel)[src];
}
/* first owned sorted slot whose key is >= bound (binary search over the device-side count) */
__device__ __forceinline__ uint32_t slab_lower_bound(const uint32_t *hash, uint32_t n, uint32_t key) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (hash[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}
/* The first / last halo_rows grid rows of the owned sorted range are contiguous slices; thread 0
 * finds them, then the block copies them into the outgoing buffers (count word first). */
__global__ void __launch_bounds__(256) k_slab_halo_pack(prs_slab s, uint32_t *__restrict__ send_dn, uint32_t *__restrict__ send_up,
                                                        uint32_t gx, uint32_t *range, uint32_t range_tile0) {
  prs::pdl_sync();
  const uint32_t n = s.counts[PRS_SC_N];
  const uint32_t *hs = s.hash_cat + s.halo_cap; /* owned keys, sorted */
  __shared__ uint32_t sh[2];
  if (threadIdx.x == 0) {
    uint32_t k_dn = 0, k_up = 0;
    if (s.has_dn) k_dn = slab_lower_bound(hs, n, min(s.row_lo + s.halo_rows, s.row_hi) * gx);
    if (s.has_up) k_up = n - slab_lower_bound(hs, n, (s.row_hi > s.row_lo + s.halo_rows ? s.row_hi - s.halo_rows : s.row_lo) * gx);
    if (k_dn > s.halo_cap || k_up > s.halo_cap) {
      if (blockIdx.x == 0) atomicOr(&s.counts[PRS_SC_ERR], PRS_SLAB_ERR_HALO_CAP);
      k_dn = min(k_dn, s.halo_cap); k_up = min(k_up, s.halo_cap);
    }
    sh[0] = k_dn; sh[1] = k_up;
    if (blockIdx.x == 0) {
      send_dn[0] = k_dn; send_up[0] = k_up; s.counts[PRS_SC_KDN] = k_dn; s.counts[PRS_SC_KUP] = k_up;
      if (range) { /* tiles between which the owned sorted keys lie: the range of the NEXT sort's scan (prs_cellbin.cuh) */
        range[0] = n ? hs[0] / prs_bin::SCAN_TILE - range_tile0 : 1u;
        range[1] = n ? hs[n - 1] / prs_bin::SCAN_TILE - range_tile0 : 0u;
      }
    }
  }
  __syncthreads();
  const uint32_t k_dn = sh[0], k_up = sh[1];
  const float4 *pr = (const float4 *)s.sortedPR + s.halo_cap;
  const float2 *sv = (const float2 *)s.sortedVel + s.halo_cap;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < k_dn + k_up; q += gridDim.x * blockDim.x) {
    const bool lower = q < k_dn;
    const uint32_t r = lower ? q : q - k_dn;              /* record number in its buffer */
    const uint32_t k = lower ? q : n - k_up + r;           /* owned sorted slot */
    uint32_t *o = (lower ? send_dn : send_up) + 1 + r;
    const float4 a = pr[k];
    const float2 v = sv[k];
    o[0 * s.halo_cap] = __float_as_uint(a.x); o[1 * s.halo_cap] = __float_as_uint(a.y);
    o[2 * s.halo_cap] = __float_as_uint(a.z); o[3 * s.halo_cap] = __float_as_uint(a.w);
    o[4 * s.halo_cap] = __float_as_uint(v.x); o[5 * s.halo_cap] = __float_as_uint(v.y);
    o[6 * s.halo_cap] = hs[k];
  }
}
/* arrivals go to the flanks: the lower neighbour's rows end at halo_cap, the upper neighbour's
 * start at halo_cap + n */
__global__ void __launch_bounds__(256) k_slab_halo_unpack(prs_slab s, const uint32_t *__restrict__ recv_dn, const uint32_t *__restrict__ recv_up) {
  prs::pdl_sync();
  const uint32_t n = s.counts[PRS_SC_N];
  const uint32_t n_lo = s.has_dn ? min(recv_dn[0], s.halo_cap) : 0u, n_hi = s.has_up ? min(recv_up[0], s.halo_cap) : 0u;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    s.counts[PRS_SC_NLO] = n_lo; s.counts[PRS_SC_NHI] = n_hi;
    s.counts[PRS_SC_STAT_HALO] += n_lo + n_hi;
  }
  float4 *pr = (float4 *)s.sortedPR;
  float2 *sv = (float2 *)s.sortedVel;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n_lo + n_hi; q += gridDim.x * blockDim.x) {
    const bool lower = q < n_lo;
    const uint32_t r = lower ? q : q - n_lo;
    const uint32_t k = lower ? s.halo_cap - n_lo + r : s.halo_cap + n + r;
    const uint32_t *o = (lower ? recv_dn : recv_up) + 1 + r;
    pr[k] = make_float4(__uint_as_float(o[0 * s.halo_cap]), __uint_as_float(o[1 * s.halo_cap]),
                        __uint_as_float(o[2 * s.halo_cap]), __uint_as_float(o[3 * s.halo_cap]));
    sv[k] = make_float2(__uint_as_float(o[4 * s.halo_cap]), __uint_as_float(o[5 * s.halo_cap]));
    s.hash_cat[k] = o[6 * s.halo_cap];
  }
}
/* cellStart/cellEnd (reference format) over the keys of [lower halo | owned | upper halo] */
__global__ void __launch_bounds__(256) k_slab_cell_table(prs_slab s) {
  prs::pdl_sync();
  const uint32_t n_lo = s.counts[PRS_SC_NLO], n_tot = n_lo + s.counts[PRS_SC_N] + s.counts[PRS_SC_NHI];
  const uint32_t slot0 = s.halo_cap - n_lo;
  const uint32_t *hash = s.hash_cat + slot0;
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_tot) return;
  const uint32_t h = hash[k];
  const uint32_t hp = (k > 0) ? hash[k - 1] : 0u;
  if (k == 0 || h != hp) {
    s.cellStart[h] = slot0 + k;
    if (k > 0) s.cellEnd[hp] = slot0 + k;
  }
  if (k == n_tot - 1) s.cellEnd[h] = slot0 + k + 1;
}
/* XORWOW states for robots that carry GLOBAL ids: subsequence = global id, so every robot's noise
 * stream is the one the single-GPU run (and the reference) gives it */
__global__ void __launch_bounds__(256) k_curand_setup_ids(curandState *__restrict__ st, const uint32_t *__restrict__ gid, uint32_t n) {
  prs::pdl_sync();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) curand_init(c_prm.p.seed, gid[i], 0, &st[i]);
}

/* ---- binned route of the slab sort (prs_cellbin.cuh): tickets, scatter, in-cell order by GLOBAL id ---- */
__global__ void __launch_bounds__(256) k_slab_tickets(prs_slab s, uint32_t *__restrict__ cellCount, uint32_t *__restrict__ ticket) {
  prs::pdl_sync();
  const uint32_t n = s.counts[PRS_SC_N];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  /* a robot that could not be sent away (migration buffer full / left the world: error bits are set already)
   * keeps a key outside the owned rows — it must not touch counters or slots that belong to nobody */
  const uint32_t h = s.hash[i];
  const uint32_t row = h / c_prm.p.gridSize.x;
  if (row < s.row_lo || row >= s.row_hi) { ticket[i] = 0xffffffffu; return; }
  ticket[i] = atomicAdd(&cellCount[h], 1u);
}
__global__ void __launch_bounds__(256)
k_slab_scatter(prs_slab s, const uint32_t *__restrict__ ticket, uint32_t *__restrict__ index_by_slot) {
  prs::pdl_sync();
  const uint32_t n = s.counts[PRS_SC_N];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t h = s.hash[i];
  if (ticket[i] == 0xffffffffu) return; /* stranded robot, see k_slab_tickets */
  const uint32_t slot = __ldg(s.cellStart + h) + ticket[i]; /* absolute slot of [halo | owned | halo] */
  s.hash_cat[slot] = h;
  index_by_slot[slot - s.halo_cap] = i;
}
/* the robots of one cell in ascending GLOBAL id (= the single-GPU stable order) + packed sorted copy */
__global__ void __launch_bounds__(256)
k_slab_gather_binned(prs_slab s, const uint32_t *__restrict__ index_by_slot, uint32_t *scratch) {
  prs::pdl_sync();
  const uint32_t n = s.counts[PRS_SC_N];
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t h = s.hash_cat[s.halo_cap + k];
  const uint32_t a = index_by_slot[k];
  const uint32_t g = s.gid[a];
  const float2 p = ((const float2 *)s.pos)[a];
  const float2 v = ((const float2 *)s.vel)[a];
  const float r = s.rad[a];
  const uint32_t c0 = __ldg(s.cellStart + h), c1 = __ldg(s.cellEnd + h);
  uint32_t below = 0;
  if (c1 - c0 <= prs_bin::MAX_RANKED_CELL) {
    for (uint32_t j = c0; j < c1; j++) below += (s.gid[index_by_slot[j - s.halo_cap]] < g) ? 1u : 0u;
  } else {
    atomicOr(&scratch[2], 1u);
    below = s.halo_cap + k - c0;
  }
  const uint32_t dst = c0 + below;
  s.index_sorted[dst - s.halo_cap] = a;
  ((float4 *)s.sortedPR)[dst] = make_float4(p.x, p.y, r, __uint_as_float(g));
  ((float2 *)s.sortedVel)[dst] = v;
}
/* fullest owned cell (guard of the binned route) from the sorted owned keys and the finished table */
__global__ void __launch_bounds__(256) k_slab_max_population(prs_slab s, uint32_t *__restrict__ out_max) {
  prs::pdl_sync();
  const uint32_t n = s.counts[PRS_SC_N];
  const uint32_t *hash = s.hash_cat + s.halo_cap;
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t m = 0;
  if (k < n) {
    const uint32_t h = hash[k];
    if (k == 0 || hash[k - 1] != h) m = s.cellEnd[h] - s.cellStart[h];
  }
  __shared__ uint32_t s_m[8];
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < 8; w++) m = max(m, s_m[w]);
    if (m > *reinterpret_cast<volatile uint32_t *>(out_max)) atomicMax(out_max, m);
  }
}
/* cell table entries of the two halo flanks only (the owned rows' entries come from the scan) */
__global__ void __launch_bounds__(256) k_slab_halo_table(prs_slab s) {
  prs::pdl_sync();
  const uint32_t n_lo = s.counts[PRS_SC_NLO], n = s.counts[PRS_SC_N], n_hi = s.counts[PRS_SC_NHI];
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_lo + n_hi) return;
  const bool lower = q < n_lo;
  const uint32_t first = lower ? s.halo_cap - n_lo : s.halo_cap + n;   /* first slot of the flank */
  const uint32_t len = lower ? n_lo : n_hi;
  const uint32_t r = lower ? q : q - n_lo;
  const uint32_t *hash = s.hash_cat + first;
  const uint32_t h = hash[r];
  const uint32_t hp = (r > 0) ? hash[r - 1] : 0u;
  if (r == 0 || h != hp) {
    s.cellStart[h] = first + r;
    if (r > 0) s.cellEnd[hp] = first + r;
  }
  if (r == len - 1) s.cellEnd[h] = first + r + 1;
}

/* -------------------------------------------------------------------------------------------- */
/* fused forms used by prs_slab_step (peer-to-peer exchange): fewer, fatter launches               */
/* -------------------------------------------------------------------------------------------- */
#define PRS_SC_PACK_DONE 13 /* counts[13]: blocks of the halo pack kernel that have finished (reset by the last one) */

__device__ __forceinline__ bool slab_spin(const volatile unsigned *flag, unsigned seq) {
  unsigned long long spins = 0;
  while ((int)(*flag - seq) < 0) {
    __nanosleep(64);
    if (++spins > (1ull << 24)) return false; /* never hang the GPU */
  }
  return true;
}

/* The rest of the migration after k_slab_select, in ONE block (a step moves a few dozen robots across a slab cut; six
 * launches — list_holes, signal, wait, fill_holes, append, commit — cost more than the work): count words of the outgoing
 * buffers and the flags that publish them; holes and movers; movers into holes; wait for the neighbours' records; arrivals
 * appended (with their cell tickets on the binned route); robot count committed. */
__global__ void __launch_bounds__(1024)
k_slab_mig_finish(prs_slab s, uint32_t *__restrict__ send_dn, uint32_t *__restrict__ send_up, unsigned *flag_out_dn,
                  unsigned *flag_out_up, const unsigned *flag_in_dn, const unsigned *flag_in_up, unsigned seq,
                  uint32_t *recv_dn, uint32_t *recv_up, uint32_t log2_gx, uint32_t *__restrict__ cellCount,
                  uint32_t *__restrict__ ticket, uint32_t *range, uint32_t range_tile0) {
  prs::pdl_sync();
  const uint32_t tid = threadIdx.x;
  const uint32_t n = s.counts[PRS_SC_N], L = s.counts[PRS_SC_LEAVERS];
  if (tid == 0) {
    send_dn[0] = min(s.counts[PRS_SC_MIGDN], s.mig_cap);
    send_up[0] = min(s.counts[PRS_SC_MIGUP], s.mig_cap);
    __threadfence_system(); /* the records (k_slab_select, complete) and the counts before the flags */
    if (flag_out_dn) *reinterpret_cast<volatile unsigned *>(flag_out_dn) = seq;
    if (flag_out_up) *reinterpret_cast<volatile unsigned *>(flag_out_up) = seq;
    __threadfence_system();
  }
  const uint32_t new_n = n - L;
  uint32_t *holes = s.lists + 2 * s.mig_cap, *movers = s.lists + 4 * s.mig_cap;
  for (uint32_t q = tid; q < L; q += blockDim.x) {
    const uint32_t li = s.lists[q];
    if (li < new_n) holes[atomicAdd(&s.counts[PRS_SC_HOLES], 1u)] = li;
    const uint32_t j = new_n + q; /* the L slots of the tail */
    if (!s.scratch[j]) movers[atomicAdd(&s.counts[PRS_SC_KEEPERS], 1u)] = j;
  }
  __syncthreads();
  const uint32_t H = *reinterpret_cast<volatile uint32_t *>(&s.counts[PRS_SC_HOLES]);
  for (uint32_t q = tid; q < H; q += blockDim.x) {
    const uint32_t dst = holes[q], src = movers[q];
    slab_move_record(s, dst, src);
    if (ticket) ticket[dst] = ticket[src]; /* the cell ticket K1 took travels with the robot */
  }
  /* the neighbours' records */
  if (tid < 2) {
    const unsigned *f = tid == 0 ? flag_in_dn : flag_in_up;
    if (f && !slab_spin(f, seq)) {
      atomicOr(&s.counts[PRS_SC_ERR], PRS_SLAB_ERR_PEER_TIMEOUT);
      *(tid == 0 ? recv_dn : recv_up) = 0u; /* nothing stale is consumed */
    }
    __threadfence_system();
  }
  __syncthreads();
  const uint32_t c_dn = s.has_dn ? min(*reinterpret_cast<volatile uint32_t *>(recv_dn), s.mig_cap) : 0u;
  const uint32_t c_up = s.has_up ? min(*reinterpret_cast<volatile uint32_t *>(recv_up), s.mig_cap) : 0u;
  for (uint32_t q = tid; q < c_dn + c_up; q += blockDim.x) {
    const uint32_t dst = new_n + q;
    if (dst >= s.cap) { atomicOr(&s.counts[PRS_SC_ERR], PRS_SLAB_ERR_CAPACITY); continue; }
    if (q < c_dn) slab_load_record(recv_dn, s.mig_cap, q, s, dst);
    else slab_load_record(recv_up, s.mig_cap, q - c_dn, s, dst);
    const uint32_t h = s.hash[dst];
    const uint32_t row = h >> log2_gx; /* a robot may cross one slab per sort at most */
    const bool mine = row >= s.row_lo && row < s.row_hi;
    if (!mine) atomicOr(&s.counts[PRS_SC_ERR], PRS_SLAB_ERR_TWO_SLABS);
    if (ticket) ticket[dst] = mine ? atomicAdd(&cellCount[h], 1u) : 0xffffffffu; /* binned sort: the arrival's cell ticket */
    if (ticket && mine && range) prs_bin::range_check(range, h / prs_bin::SCAN_TILE - range_tile0);
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t nn = new_n + c_dn + c_up;
    if (nn > s.cap) nn = s.cap;
    s.counts[PRS_SC_N] = nn;
    s.counts[PRS_SC_STAT_MIG] += L;
  }
}

/* k_slab_halo_pack whose LAST block publishes the exchange's sequence number in the neighbours' flag words */
__global__ void __launch_bounds__(256) k_slab_halo_pack_signal(prs_slab s, uint32_t *__restrict__ send_dn, uint32_t *__restrict__ send_up,
                                                               uint32_t gx, unsigned *flag_dn, unsigned *flag_up, unsigned seq, uint32_t *range,
                                                               uint32_t range_tile0) {
  prs::pdl_sync();
  const uint32_t n = s.counts[PRS_SC_N];
  const uint32_t *hs = s.hash_cat + s.halo_cap; /* owned keys, sorted */
  __shared__ uint32_t sh[2];
  if (threadIdx.x == 0) {
    uint32_t k_dn = 0, k_up = 0;
    if (s.has_dn) k_dn = slab_lower_bound(hs, n, min(s.row_lo + s.halo_rows, s.row_hi) * gx);
    if (s.has_up) k_up = n - slab_lower_bound(hs, n, (s.row_hi > s.row_lo + s.halo_rows ? s.row_hi - s.halo_rows : s.row_lo) * gx);
    if (k_dn > s.halo_cap || k_up > s.halo_cap) {
      if (blockIdx.x == 0) atomicOr(&s.counts[PRS_SC_ERR], PRS_SLAB_ERR_HALO_CAP);
      k_dn = min(k_dn, s.halo_cap); k_up = min(k_up, s.halo_cap);
    }
    sh[0] = k_dn; sh[1] = k_up;
    if (blockIdx.x == 0) {
      send_dn[0] = k_dn; send_up[0] = k_up; s.counts[PRS_SC_KDN] = k_dn; s.counts[PRS_SC_KUP] = k_up;
      if (range) { /* tiles between which the owned sorted keys lie: the range of the NEXT sort's scan (prs_cellbin.cuh) */
        range[0] = n ? hs[0] / prs_bin::SCAN_TILE - range_tile0 : 1u;
        range[1] = n ? hs[n - 1] / prs_bin::SCAN_TILE - range_tile0 : 0u;
      }
    }
  }
  __syncthreads();
  const uint32_t k_dn = sh[0], k_up = sh[1];
  const float4 *pr = (const float4 *)s.sortedPR + s.halo_cap;
  const float2 *sv = (const float2 *)s.sortedVel + s.halo_cap;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < k_dn + k_up; q += gridDim.x * blockDim.x) {
    const bool lower = q < k_dn;
    const uint32_t r = lower ? q : q - k_dn;              /* record number in its buffer */
    const uint32_t k = lower ? q : n - k_up + r;           /* owned sorted slot */
    uint32_t *o = (lower ? send_dn : send_up) + 1 + r;
    const float4 a = pr[k];
    const float2 v = sv[k];
    o[0 * s.halo_cap] = __float_as_uint(a.x); o[1 * s.halo_cap] = __float_as_uint(a.y);
    o[2 * s.halo_cap] = __float_as_uint(a.z); o[3 * s.halo_cap] = __float_as_uint(a.w);
    o[4 * s.halo_cap] = __float_as_uint(v.x); o[5 * s.halo_cap] = __float_as_uint(v.y);
    o[6 * s.halo_cap] = hs[k];
  }
  /* every thread's peer stores ordered at system scope, then the block checks in; the last block publishes */
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(&s.counts[PRS_SC_PACK_DONE], 1u);
    if (done == gridDim.x - 1) {
      s.counts[PRS_SC_PACK_DONE] = 0u;
      __threadfence_system();
      if (flag_dn) *reinterpret_cast<volatile unsigned *>(flag_dn) = seq;
      if (flag_up) *reinterpret_cast<volatile unsigned *>(flag_up) = seq;
      __threadfence_system();
    }
  }
}

/* up to four ranges of cellStart words to be set to "empty" (the halo rows around the slab, wrapped ones included) */
struct SlabClears { uint32_t *ptr[4]; uint32_t words[4]; };

/* k_slab_halo_unpack that waits for the neighbours' flags itself (every block polls: the grid is resident) and clears
 * the halo rows of the cell table on the way */
__global__ void __launch_bounds__(256) k_slab_halo_unpack_wait(prs_slab s, uint32_t *recv_dn, uint32_t *recv_up, const unsigned *flag_dn,
                                                               const unsigned *flag_up, unsigned seq, const SlabClears cl) {
  prs::pdl_sync();
#pragma unroll
  for (int r = 0; r < 4; r++)
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < cl.words[r]; w += gridDim.x * blockDim.x) cl.ptr[r][w] = 0xffffffffu;
  __shared__ uint32_t s_drop[2];
  if (threadIdx.x < 2) {
    const unsigned *f = threadIdx.x == 0 ? flag_dn : flag_up;
    bool ok = true;
    if (f && !slab_spin(f, seq)) {
      ok = false;
      atomicOr(&s.counts[PRS_SC_ERR], PRS_SLAB_ERR_PEER_TIMEOUT);
    }
    s_drop[threadIdx.x] = ok ? 0u : 1u;
    __threadfence_system();
  }
  __syncthreads();
  const uint32_t n = s.counts[PRS_SC_N];
  const uint32_t n_lo = (s.has_dn && !s_drop[0]) ? min(*reinterpret_cast<volatile uint32_t *>(recv_dn), s.halo_cap) : 0u;
  const uint32_t n_hi = (s.has_up && !s_drop[1]) ? min(*reinterpret_cast<volatile uint32_t *>(recv_up), s.halo_cap) : 0u;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    s.counts[PRS_SC_NLO] = n_lo; s.counts[PRS_SC_NHI] = n_hi;
    s.counts[PRS_SC_STAT_HALO] += n_lo + n_hi;
  }
  float4 *pr = (float4 *)s.sortedPR;
  float2 *sv = (float2 *)s.sortedVel;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n_lo + n_hi; q += gridDim.x * blockDim.x) {
    const bool lower = q < n_lo;
    const uint32_t r = lower ? q : q - n_lo;
    const uint32_t k = lower ? s.halo_cap - n_lo + r : s.halo_cap + n + r;
    const uint32_t *o = (lower ? recv_dn : recv_up) + 1 + r;
    pr[k] = make_float4(__uint_as_float(o[0 * s.halo_cap]), __uint_as_float(o[1 * s.halo_cap]),
                        __uint_as_float(o[2 * s.halo_cap]), __uint_as_float(o[3 * s.halo_cap]));
    sv[k] = make_float2(__uint_as_float(o[4 * s.halo_cap]), __uint_as_float(o[5 * s.halo_cap]));
    s.hash_cat[k] = o[6 * s.halo_cap];
  }
}

/* -------------------------------------------------------------------------------------------- */
/* C entry points                                                                                 */
/* -------------------------------------------------------------------------------------------- */
extern "C" {

static uint32_t slab_log2_gx() {
  uint32_t b = 0;
  while ((1u << b) < g_prs.h_prm.p.gridSize.x) b++;
  return b;
}
static void slab_check(const prs_slab *s) {
  if (!s->cap || !s->halo_cap || !s->mig_cap) { fprintf(stderr, "prs_slab: zero capacity\n"); exit(EXIT_FAILURE); }
}

/* tile range of the slab's scan (prs_cellbin.cuh RangeArgs): usable when the slab's first cell starts a scan tile.
 * slab_range_ptr: the device words, or nullptr when the feature is off / not applicable;
 * slab_range_valid: the words describe THIS slab (a halo pack kernel wrote them for this geometry) */
static uint32_t slab_range_tile0(const prs_slab *s) {
  return (uint32_t)(((size_t)s->row_lo * g_prs.h_prm.p.gridSize.x) / prs_bin::SCAN_TILE);
}
static uint32_t *slab_range_ptr(const prs_slab *s) {
  if (!g_prs.slab_scan_range || !g_prs.bin.range) return nullptr;
  if (((size_t)s->row_lo * g_prs.h_prm.p.gridSize.x) % prs_bin::SCAN_TILE) return nullptr;
  return g_prs.bin.range;
}
static bool slab_range_valid(const prs_slab *s) {
  const PrsBinState &B = g_prs.bin;
  return slab_range_ptr(s) && B.range_table == (const void *)s->cellStart && B.range_row_lo == s->row_lo && B.range_row_hi == s->row_hi;
}
static void slab_range_written(const prs_slab *s) {
  PrsBinState &B = g_prs.bin;
  B.range_table = s->cellStart; B.range_row_lo = s->row_lo; B.range_row_hi = s->row_hi;
}

size_t prs_slab_mig_words(unsigned mig_cap) { return 1 + (size_t)SLAB_MIG_WORDS * mig_cap; }
size_t prs_slab_halo_words(unsigned halo_cap) { return 1 + (size_t)SLAB_HALO_WORDS * halo_cap; }

void prs_slab_rng_setup(const prs_slab *s, unsigned n) {
  if (!n) return;
  PRS_LAUNCH(k_curand_setup_ids, div_up(n, 256), 256, 0, (curandState *)s->rng, s->gid, n);
}
/* K1.  On sort steps the route of the sort is decided HERE (binned while the swarm is known to be sparse, the
 * same asynchronous guard as the single-GPU step) because the binned route takes the cell tickets inside K1:
 * one pass over the robots less than ticketing after the migration.  Robots that leave keep no ticket, arrivals
 * take theirs when they are appended. */
void prs_slab_k1(const prs_slab *s, float time, float dt, int do_hash) {
  slab_check(s);
  const int run_controller = (g_prs.h_prm.p.control == LIGHT_WAVE && time >= 0) ? 1 : 0;
  StageScope t(PRS_STAGE_K1);
  PRS_LAUNCH_PDL(k_slab_begin_step, 1, 32, s->counts, g_prs.bin.range);
  g_prs.slab_tickets = false;
  if (do_hash) {
    PrsBinState &B = g_prs.bin;
    bin_poll_report();
    const unsigned gx = g_prs.h_prm.p.gridSize.x;
    const unsigned cells = (s->row_hi - s->row_lo) * gx; /* cells of the owned rows */
    g_prs.slab_binned = (B.mode == 2 || (B.mode == 0 && B.admitted)) && (unsigned long long)cells <= 16ull * s->cap;
    if (g_prs.slab_binned) {
      bin_ensure(s->cap, g_prs.h_prm.p.numCells);
      PRS_CUDA(cudaMemsetAsync(B.scratch, 0, 16, g_prs.stream));
      /* tickets outside the tile range of this step's scan are noticed where they are taken (only while the range is in use) */
      g_prs.slab_range_in_use = slab_range_valid(s);
      uint32_t *k1_range = g_prs.slab_range_in_use ? slab_range_ptr(s) : nullptr;
      const uintptr_t al = (uintptr_t)s->pos | (uintptr_t)s->vel | (((uintptr_t)s->rad | (uintptr_t)s->phase | (uintptr_t)s->absForce_a |
                            (uintptr_t)s->absForce_r | (uintptr_t)s->dead | (uintptr_t)s->hash | (uintptr_t)g_prs.sort_ws.vals[0]) << 1);
      if (g_prs.k1_x2 && (al & 15u) == 0) {
        PRS_LAUNCH_PDL(k_control_integrate_hash_x2, div_up(div_up(s->cap, 2), 256), 256, (float4 *)s->pos, (float4 *)s->vel, (float2 *)s->rad,
                       (const float2 *)s->phase, (const float2 *)s->absForce_a, (const float2 *)s->absForce_r, (const int2 *)s->dead,
                       (uint2 *)s->hash, (uint2 *)g_prs.sort_ws.vals[0], time, dt, run_controller, s->cap, B.cellCount, (uint32_t *)nullptr,
                       (const uint32_t *)(s->counts + PRS_SC_N), s->row_lo, s->row_hi, slab_log2_gx(), k1_range, slab_range_tile0(s));
      } else {
        PRS_LAUNCH_PDL((k_control_integrate_hash<true, true>), div_up(s->cap, 256), 256, (float2 *)s->pos, (float2 *)s->vel, s->rad,
                       s->phase, s->absForce_a, s->absForce_r, s->dead, s->hash, g_prs.sort_ws.vals[0], time, dt, run_controller, s->cap,
                       (const uint32_t *)(s->counts + PRS_SC_N), B.cellCount, (uint32_t *)nullptr, s->row_lo, s->row_hi, slab_log2_gx(),
                       k1_range, slab_range_tile0(s));
      }
      g_prs.slab_tickets = true;
    } else {
      PRS_LAUNCH_PDL((k_control_integrate_hash<true, false>), div_up(s->cap, 256), 256, (float2 *)s->pos, (float2 *)s->vel, s->rad,
                     s->phase, s->absForce_a, s->absForce_r, s->dead, s->hash, s->scratch, time, dt, run_controller, s->cap,
                     (const uint32_t *)(s->counts + PRS_SC_N), (uint32_t *)nullptr, (uint32_t *)nullptr, 0u, 0xffffffffu, 0u, (uint32_t *)nullptr, 0u);
    }
  } else {
    PRS_LAUNCH_PDL((k_control_integrate_hash<false, false>), div_up(s->cap, 256), 256, (float2 *)s->pos, (float2 *)s->vel, s->rad,
                   s->phase, s->absForce_a, s->absForce_r, s->dead, s->hash, s->scratch, time, dt, run_controller, s->cap,
                   (const uint32_t *)(s->counts + PRS_SC_N), (uint32_t *)nullptr, (uint32_t *)nullptr, 0u, 0xffffffffu, 0u, (uint32_t *)nullptr, 0u);
  }
}
void prs_slab_migrate_pack(const prs_slab *s, unsigned *send_dn, unsigned *send_up) {
  StageScope t(PRS_STAGE_EXCHANGE);
  PRS_LAUNCH_PDL(k_slab_select, div_up(s->cap, 256), 256, *s, send_dn, send_up, slab_log2_gx());
  PRS_LAUNCH_PDL(k_slab_list_holes, div_up(2 * s->mig_cap, 256), 256, *s, send_dn, send_up);
}
void prs_slab_migrate_unpack(const prs_slab *s, const unsigned *recv_dn, const unsigned *recv_up) {
  StageScope t(PRS_STAGE_EXCHANGE);
  uint32_t *ticket = g_prs.slab_tickets ? g_prs.sort_ws.vals[0] : nullptr;
  PRS_LAUNCH_PDL(k_slab_fill_holes, div_up(2 * s->mig_cap, 256), 256, *s, ticket);
  PRS_LAUNCH_PDL(k_slab_append, div_up(2 * s->mig_cap, 256), 256, *s, recv_dn, recv_up, slab_log2_gx(),
                 g_prs.slab_tickets ? g_prs.bin.cellCount : (uint32_t *)nullptr, ticket,
                 (g_prs.slab_tickets && g_prs.slab_range_in_use) ? slab_range_ptr(s) : (uint32_t *)nullptr, slab_range_tile0(s));
  PRS_LAUNCH_PDL(k_slab_commit_count, 1, 1, *s, recv_dn, recv_up);
}
/* (hash, local slot) of the owned robots sorted by hash into hash_cat[halo_cap ..] / index_sorted,
 * robots of one cell in ascending global id */
void prs_slab_sort(const prs_slab *s) {
  PrsBinState &B = g_prs.bin;
  const unsigned gx = g_prs.h_prm.p.gridSize.x;
  const unsigned cells = (s->row_hi - s->row_lo) * gx; /* cells of the owned rows */
  StageScope t(PRS_STAGE_SORT);
  if (g_prs.slab_binned) {
    /* tickets (taken by K1 and by the arrivals) -> scan of the owned rows' cells (= their cell table, slots
     * offset by the lower halo) -> scatter; the in-cell order by global id is part of the gather */
    prs_sort::Workspace &w = g_prs.sort_ws;
    const size_t c_lo = (size_t)s->row_lo * gx;
    const unsigned tiles = div_up(cells, prs_bin::SCAN_TILE);
    if (!g_prs.slab_tickets) { /* K1 ran without the route being known (callers of earlier builds) */
      bin_ensure(s->cap, g_prs.h_prm.p.numCells);
      PRS_CUDA(cudaMemsetAsync(B.scratch, 0, 16, g_prs.stream));
      PRS_LAUNCH_PDL(k_slab_tickets, div_up(s->cap, 256), 256, *s, B.cellCount, w.vals[0]);
    }
    /* dense start table over the owned rows' cells (same slot offset as the table): read by the collide of the interior rows */
    prs_bin::DenseArgs dn;
    if (g_prs.collide_dense) dn.dense = B.dense + c_lo;
    /* tiles outside the range the slab's robots occupy are skipped (empty before, empty now); the range was in use in K1 and
     * for the arrivals, so a ticket outside it has set the violation word and the kernels process everything */
    prs_bin::RangeArgs ra;
    if (g_prs.slab_tickets && g_prs.slab_range_in_use) {
      ra.range = slab_range_ptr(s);
      /* a row of movement between two sorts + the rows a stencil reaches (2, and the five-cell spill into the next) */
      ra.dil = (4u * gx + prs_bin::SCAN_TILE - 1u) / prs_bin::SCAN_TILE + 1u;
    }
    PRS_LAUNCH_PDL(prs_bin::k_cell_tile_sums, tiles, prs_bin::SCAN_THREADS, (const uint32_t *)(B.cellCount + c_lo), cells, B.scratch,
                   (const uint32_t *)nullptr, prs_bin::DenseArgs(), ra);
    if (tiles <= prs_bin::SELF_PREFIX_MAX_TILES) {
      PRS_LAUNCH_PDL(prs_bin::k_cell_apply<true>, tiles, prs_bin::SCAN_THREADS, B.cellCount + c_lo, s->cellStart + c_lo, s->cellEnd + c_lo,
                     cells, B.scratch, s->halo_cap, (uint32_t *)nullptr, (uint32_t *)nullptr, prs_bin::PatchListArgs(), dn, ra);
    } else {
      PRS_LAUNCH_PDL(prs_bin::k_cell_scan_tiles, 1, 1024, B.scratch, tiles);
      PRS_LAUNCH_PDL(prs_bin::k_cell_apply<false>, tiles, prs_bin::SCAN_THREADS, B.cellCount + c_lo, s->cellStart + c_lo, s->cellEnd + c_lo,
                     cells, B.scratch, s->halo_cap, (uint32_t *)nullptr, (uint32_t *)nullptr, prs_bin::PatchListArgs(), dn, ra);
    }
    PRS_LAUNCH_PDL(k_slab_scatter, div_up(s->cap, 256), 256, *s, (const uint32_t *)w.vals[0], w.vals[1]);
    g_prs.slab_table_fresh = true; /* consumed by this step's gather and cell_table */
    return;
  }
  g_prs.slab_sorted_onesweep = true;
  sort_pairs(s->hash, nullptr, s->hash_cat + s->halo_cap, s->index_sorted, s->cap, key_bits_of_grid(), true, s->counts + PRS_SC_N);
  PRS_LAUNCH_PDL(k_fix_ties_by_gid, div_up(s->cap, 256), 256, (const uint32_t *)(s->hash_cat + s->halo_cap), s->index_sorted,
                 (const uint32_t *)s->gid, (const uint32_t *)(s->counts + PRS_SC_N));
}
void prs_slab_gather(const prs_slab *s) {
  StageScope t(PRS_STAGE_REORDER);
  if (g_prs.slab_binned && g_prs.slab_table_fresh) {
    PRS_LAUNCH_PDL(k_slab_gather_binned, div_up(s->cap, 256), 256, *s, (const uint32_t *)g_prs.sort_ws.vals[1], g_prs.bin.scratch);
    bin_send_report(g_prs.bin.scratch + 1);
    return;
  }
  PRS_LAUNCH_PDL(k_slab_gather, div_up(s->cap, 256), 256, *s);
}
void prs_slab_halo_pack(const prs_slab *s, unsigned *send_dn, unsigned *send_up) {
  StageScope t(PRS_STAGE_EXCHANGE);
  PRS_LAUNCH_PDL(k_slab_halo_pack, min(div_up(2 * s->halo_cap, 256), 592u), 256, *s, send_dn, send_up, g_prs.h_prm.p.gridSize.x,
                 slab_range_ptr(s), slab_range_tile0(s));
  if (slab_range_ptr(s)) slab_range_written(s);
}
void prs_slab_halo_unpack(const prs_slab *s, const unsigned *recv_dn, const unsigned *recv_up) {
  StageScope t(PRS_STAGE_EXCHANGE);
  PRS_LAUNCH_PDL(k_slab_halo_unpack, min(div_up(2 * s->halo_cap, 256), 592u), 256, *s, recv_dn, recv_up);
}
/* cell table over [halo | owned | halo].  After a binned sort the owned rows' entries already exist
 * (the scan wrote them): only the halo rows are cleared and filled.  Otherwise (onesweep route, or
 * a step without sort: stale table) everything this rank can see is rebuilt.
 * slab_table_clears: which words of cellStart must be set to "empty" first; slab_table_fill: the kernel that writes
 * the entries (and, on the onesweep route, the report that can admit the binned route). */
static SlabClears slab_table_clears(const prs_slab *s) {
  SlabClears cl;
  for (int r = 0; r < 4; r++) { cl.ptr[r] = nullptr; cl.words[r] = 0u; }
  int k = 0;
  auto add = [&](size_t first_row, size_t rows) { if (rows) { cl.ptr[k] = s->cellStart + first_row * g_prs.h_prm.p.gridSize.x; cl.words[k] = (uint32_t)(rows * g_prs.h_prm.p.gridSize.x); k++; } };
  const unsigned gy = g_prs.h_prm.p.gridSize.y;
  const unsigned r0 = s->row_lo > s->halo_rows ? s->row_lo - s->halo_rows : 0u;
  const unsigned r1 = min(s->row_hi + s->halo_rows, gy);
  if (s->wrap) { /* ring of slabs: the halo rows that lie beyond the grid edge are the rows at its other end */
    if (s->row_lo < s->halo_rows) { const unsigned rows = min(s->halo_rows - s->row_lo, gy); add(gy - rows, rows); }
    if (s->row_hi + s->halo_rows > gy) add(0, min(s->row_hi + s->halo_rows - gy, gy));
  }
  if (g_prs.slab_binned && g_prs.slab_table_fresh) {
    if (s->row_lo > r0) add(r0, s->row_lo - r0);
    if (r1 > s->row_hi) add(s->row_hi, r1 - s->row_hi);
  } else {
    add(r0, r1 - r0);
  }
  return cl;
}
static void slab_table_fill(const prs_slab *s) {
  if (g_prs.slab_binned && g_prs.slab_table_fresh) {
    PRS_LAUNCH_PDL(k_slab_halo_table, div_up(2 * s->halo_cap, 256), 256, *s);
    g_prs.slab_table_fresh = false;
    return;
  }
  PRS_LAUNCH_PDL(k_slab_cell_table, div_up(s->cap + 2 * s->halo_cap, 256), 256, *s);
  if (g_prs.slab_sorted_onesweep && g_prs.bin.mode == 0 && !g_prs.bin.admitted) {
    /* report the fullest cell so that the binned route can be admitted */
    bin_ensure(s->cap, g_prs.h_prm.p.numCells);
    PRS_CUDA(cudaMemsetAsync(g_prs.bin.scratch, 0, 16, g_prs.stream));
    PRS_LAUNCH_PDL(k_slab_max_population, div_up(s->cap, 256), 256, *s, g_prs.bin.scratch + 1);
    bin_send_report(g_prs.bin.scratch + 1);
  }
  g_prs.slab_sorted_onesweep = false;
}
void prs_slab_cell_table(const prs_slab *s) {
  StageScope t(PRS_STAGE_REORDER);
  const SlabClears cl = slab_table_clears(s);
  for (int r = 0; r < 4; r++)
    if (cl.words[r]) PRS_CUDA(cudaMemsetAsync(cl.ptr[r], 0xff, (size_t)cl.words[r] * sizeof(unsigned), g_prs.stream));
  slab_table_fill(s);
}
/* collide for the owned sorted slots [halo_cap, halo_cap + n); results go to the local slots.
 * band 0: all of them; 1: the interior rows only (their stencils stay inside the owned rows: needs no halo);
 * 2 / 3: the first / last halo_rows rows (after the halo has been unpacked and tabled) */
static void slab_collide_band(const prs_slab *s, float dt, int band) {
  const bool need_fa = g_prs.h_prm.p.constrained_contraction != 0;
  prs::PackedLayout in{(const float4 *)s->sortedPR, (const float2 *)s->sortedVel};
  const unsigned span = (band >= 2) ? min(s->halo_cap, s->cap) : s->cap; /* an edge band is what goes out as a halo: <= halo_cap slots */
  /* interior rows right after a binned sort: their stencils stay inside the owned rows, whose dense start table the scan has written */
  const uint32_t *dense = (band == 1 && g_prs.collide_dense && g_prs.slab_binned && g_prs.slab_table_fresh) ? g_prs.bin.dense : nullptr;
  prs_launch_collide_t((float2 *)s->vel, s->absForce_a, s->absForce_r, in, s->cellStart, s->cellEnd, s->halo_cap + span, dt,
                       need_fa, s->halo_cap, s->counts + PRS_SC_N, band, s->index_sorted - s->halo_cap, dense,
                       dense ? (const uint32_t *)s->hash_cat : nullptr);
}
void prs_slab_collide(const prs_slab *s, float dt) {
  slab_check(s);
  StageScope t(PRS_STAGE_COLLIDE);
  slab_collide_band(s, dt, 0);
}
void prs_slab_min_light_distance(const prs_slab *s, float *d_min_d) {
  PRS_CUDA(cudaMemsetAsync(d_min_d, 0x7f, sizeof(float), g_prs.stream));
  const unsigned blocks = min(div_up(s->cap, 256 * 4), 148u * 8u);
  PRS_LAUNCH(k_min_light_distance, blocks, 256, 0, (const float2 *)s->pos, s->cap, (uint32_t *)d_min_d, s->counts + PRS_SC_N);
}
void prs_slab_update_phase(const prs_slab *s, float spacing, const float *d_min_d) {
  PRS_LAUNCH(k_update_phase, div_up(s->cap, 256), 256, 0, (const float2 *)s->pos, s->phase, spacing, 0.0f, d_min_d, s->cap,
             s->counts + PRS_SC_N);
}
void prs_slab_add_noise(const prs_slab *s, float std) {
  PRS_LAUNCH(k_add_normal_noise, div_up(s->cap, 256), 256, 0, (curandState *)s->rng, s->phase, std, s->cap, s->counts + PRS_SC_N);
}

/* -------------------------------------------------------------------------------------------- */
/* peer-to-peer exchange: the pack kernels write straight into the NEIGHBOUR's mailbox over       */
/* NVLink (peer-mapped memory), a flag word carries the sequence number, the consumer waits for   */
/* it on the device — no NCCL call, no host involvement, nothing between the producing kernel     */
/* and the neighbour's HBM.  Mailboxes are double-buffered by step parity: a rank can only be one */
/* exchange ahead of its neighbour (it needs the neighbour's data of the current exchange to go   */
/* on), so the buffer of exchange t+2 is free when t+2 is written.                                */
/* -------------------------------------------------------------------------------------------- */
__global__ void k_slab_signal(unsigned *remote_dn, unsigned *remote_up, unsigned seq) {
  prs::pdl_sync();
  /* the pack kernel that ran before this one in the stream has completed: order its peer writes
   * before the flag at system scope */
  __threadfence_system();
  if (remote_dn) *reinterpret_cast<volatile unsigned *>(remote_dn) = seq;
  if (remote_up) *reinterpret_cast<volatile unsigned *>(remote_up) = seq;
  __threadfence_system();
}
/* recv_dn / recv_up (optional): the local buffers the neighbours fill.  A neighbour that does not show up within
 * ~15 s sets the sticky PEER_TIMEOUT bit AND empties that buffer (count word 0), so that the unpack kernels do not
 * consume whatever an earlier exchange left there; the host hears about the bit one step later at most. */
__global__ void k_slab_wait(const unsigned *local_dn, const unsigned *local_up, unsigned seq, unsigned *counts,
                            unsigned *recv_dn, unsigned *recv_up) {
  prs::pdl_sync();
  const volatile unsigned *f = threadIdx.x == 0 ? local_dn : local_up;
  if (threadIdx.x < 2 && f) {
    unsigned long long spins = 0;
    while ((int)(*f - seq) < 0) {
      __nanosleep(64);
      if (++spins > (1ull << 24)) { /* never hang the GPU */
        atomicOr(&counts[PRS_SC_ERR], PRS_SLAB_ERR_PEER_TIMEOUT);
        unsigned *r = threadIdx.x == 0 ? recv_dn : recv_up;
        if (r) *r = 0u;
        break;
      }
    }
  }
  __threadfence_system();
}

size_t prs_ipc_handle_size(void) { return sizeof(cudaIpcMemHandle_t); }
/* device memory that can be mapped by peers: plain cudaMalloc, zeroed */
void *prs_slab_mailbox_alloc(size_t words) {
  void *p = nullptr;
  PRS_CUDA(cudaMalloc(&p, words * 4));
  PRS_CUDA(cudaMemset(p, 0, words * 4));
  return p;
}
void prs_slab_mailbox_free(void *p) { if (p) PRS_CUDA(cudaFree(p)); }
void prs_ipc_export(void *dev_ptr, void *handle_out) {
  PRS_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle_out, dev_ptr));
}
/* returns NULL (and clears the error) if the peer's memory cannot be mapped — the caller then
 * falls back to the NCCL exchange on every rank */
void *prs_ipc_open(const void *handle) {
  void *p = nullptr;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    fprintf(stderr, "prs_ipc_open: %s — peer-to-peer exchange not available\n", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return nullptr;
  }
  return p;
}
void prs_ipc_close(void *p) { if (p) PRS_CUDA(cudaIpcCloseMemHandle(p)); }
void prs_slab_signal(unsigned *remote_flag_dn, unsigned *remote_flag_up, unsigned seq) {
  StageScope t(PRS_STAGE_EXCHANGE);
  PRS_LAUNCH_PDL(k_slab_signal, 1, 1, remote_flag_dn, remote_flag_up, seq);
}
void prs_slab_wait(const prs_slab *s, const unsigned *local_flag_dn, const unsigned *local_flag_up, unsigned seq) {
  StageScope t(PRS_STAGE_EXCHANGE);
  PRS_LAUNCH_PDL(k_slab_wait, 1, 32, local_flag_dn, local_flag_up, seq, s->counts, (unsigned *)nullptr, (unsigned *)nullptr);
}
static void slab_wait_drop(const prs_slab *s, const unsigned *flag_dn, const unsigned *flag_up, unsigned seq, unsigned *recv_dn,
                           unsigned *recv_up) {
  StageScope t(PRS_STAGE_EXCHANGE);
  PRS_LAUNCH_PDL(k_slab_wait, 1, 32, flag_dn, flag_up, seq, s->counts, flag_dn ? recv_dn : (unsigned *)nullptr,
                 flag_up ? recv_up : (unsigned *)nullptr);
}

/* ---------------------------------------------------------------------------------------------------------------
 * One whole step of a slab rank, orchestrated here (Particlebot::update, particlebot.cpp:170-300, cut at the two
 * exchanges) — the caller only provides what needs the process group: the mapped mailboxes of the two neighbours
 * (set up once) and, every phase_update_interval, the MIN over all ranks of one float.
 *
 *   [phase gate] local min light distance -> allreduce_min callback -> phase offsets (+ noise)
 *   K1 (controller, integrate, hash + cell tickets on sort steps)
 *   [sort steps] leavers packed straight into the neighbours' mailboxes -> signal -> wait -> arrivals appended -> sort
 *   gather -> first / last halo_rows rows packed into the neighbours' mailboxes -> signal
 *   collide of the INTERIOR rows (their stencils need no halo): the neighbours' halos arrive underneath it
 *   wait -> halo unpack -> halo rows of the cell table -> collide of the two edge bands
 *
 * Mailbox layout (uint32 words), identical on every rank: for source s in {0: from the lower neighbour, 1: from the upper}
 * halo[2 parities][hw] then mig[2 parities][mw]; after both regions the flag words halo_seq[2], mig_seq[2].
 * Returns the sticky error bits seen so far (counts[PRS_SC_ERR], copied to pinned memory asynchronously: at most one
 * step late and never a synchronisation); a peer time-out makes the receiving side drop that exchange (count 0).
 * --------------------------------------------------------------------------------------------------------------- */
static inline unsigned *mb_buf(unsigned *base, const prs_slab_ctx *c, int src, int halo, unsigned parity) {
  const size_t region = 2 * (size_t)c->hw + 2 * (size_t)c->mw;
  return base + src * region + (halo ? parity * (size_t)c->hw : 2 * (size_t)c->hw + parity * (size_t)c->mw);
}
static inline unsigned *mb_flag(unsigned *base, const prs_slab_ctx *c, int src, int halo) {
  const size_t region = 2 * (size_t)c->hw + 2 * (size_t)c->mw;
  return base + 2 * region + (halo ? 0 : 2) + src;
}
static inline bool slab_gate(float time, float interval, float dt) { return time - interval * floorf(time / interval) < dt; }

size_t prs_slab_mailbox_words(unsigned mig_cap, unsigned halo_cap) {
  return 2 * (2 * prs_slab_halo_words(halo_cap) + 2 * prs_slab_mig_words(mig_cap)) + 16;
}

unsigned prs_slab_step(prs_slab_ctx *c, float dt, float sort_interval) {
  const prs_slab *s = &c->slab;
  const SimParams &P = g_prs.h_prm.p;
  const float time = c->time;
  const bool phase_step = P.control == LIGHT_WAVE && slab_gate(time, P.phase_update_interval, dt);
  const bool sort_step = slab_gate(time, sort_interval, dt) || !c->sorted_once;
  if (phase_step) {
    StageScope t(PRS_STAGE_PHASE);
    prs_slab_min_light_distance(s, c->d_min_d);
    if (c->allreduce_min) c->allreduce_min(c->d_min_d, c->user);
    prs_slab_update_phase(s, 2.0f * P.min_radius, c->d_min_d);
    if (P.phase_std) prs_slab_add_noise(s, P.phase_std);
  }
  prs_slab_k1(s, time, dt, sort_step ? 1 : 0);
  if (sort_step) {
    const unsigned q = ++c->seq_mig, par = q & 1u;
    unsigned *dn = c->peer_dn ? mb_buf(c->peer_dn, c, 1, 0, par) : c->scratch_mig[0]; /* I am the lower rank's UPPER neighbour */
    unsigned *up = c->peer_up ? mb_buf(c->peer_up, c, 0, 0, par) : c->scratch_mig[1];
    if (c->fused_exchange) {
      /* leavers packed straight into the neighbours' mailboxes, then ONE single-block kernel for the rest of the migration:
       * count words + flags out, holes filled, flags in, arrivals appended, count committed */
      StageScope t(PRS_STAGE_EXCHANGE);
      PRS_LAUNCH_PDL(k_slab_select, div_up(s->cap, 256), 256, *s, dn, up, slab_log2_gx());
      uint32_t *ticket = g_prs.slab_tickets ? g_prs.sort_ws.vals[0] : nullptr;
      PRS_LAUNCH_PDL(k_slab_mig_finish, 1, 1024, *s, dn, up, c->peer_dn ? mb_flag(c->peer_dn, c, 1, 0) : (unsigned *)nullptr,
                     c->peer_up ? mb_flag(c->peer_up, c, 0, 0) : (unsigned *)nullptr,
                     (const unsigned *)(c->peer_dn ? mb_flag(c->mailbox, c, 0, 0) : nullptr),
                     (const unsigned *)(c->peer_up ? mb_flag(c->mailbox, c, 1, 0) : nullptr), q, mb_buf(c->mailbox, c, 0, 0, par),
                     mb_buf(c->mailbox, c, 1, 0, par), slab_log2_gx(), g_prs.slab_tickets ? g_prs.bin.cellCount : (uint32_t *)nullptr, ticket,
                     (g_prs.slab_tickets && g_prs.slab_range_in_use) ? slab_range_ptr(s) : (uint32_t *)nullptr, slab_range_tile0(s));
    } else {
      prs_slab_migrate_pack(s, dn, up);
      prs_slab_signal(c->peer_dn ? mb_flag(c->peer_dn, c, 1, 0) : nullptr, c->peer_up ? mb_flag(c->peer_up, c, 0, 0) : nullptr, q);
      slab_wait_drop(s, c->peer_dn ? mb_flag(c->mailbox, c, 0, 0) : nullptr, c->peer_up ? mb_flag(c->mailbox, c, 1, 0) : nullptr, q,
                     mb_buf(c->mailbox, c, 0, 0, par), mb_buf(c->mailbox, c, 1, 0, par));
      prs_slab_migrate_unpack(s, mb_buf(c->mailbox, c, 0, 0, par), mb_buf(c->mailbox, c, 1, 0, par));
    }
    prs_slab_sort(s);
    c->sorted_once = 1;
  }
  k1_done(); /* owned state final in its slots: host-buffer steps may start copying positions / radii back */
  prs_slab_gather(s);
  {
    const unsigned q = ++c->seq_halo, par = q & 1u;
    unsigned *dn = c->peer_dn ? mb_buf(c->peer_dn, c, 1, 1, par) : c->scratch_halo[0];
    unsigned *up = c->peer_up ? mb_buf(c->peer_up, c, 0, 1, par) : c->scratch_halo[1];
    if (c->fused_exchange) { /* the last block of the pack kernel publishes the flags */
      StageScope t(PRS_STAGE_EXCHANGE);
      PRS_LAUNCH_PDL(k_slab_halo_pack_signal, min(div_up(2 * s->halo_cap, 256), 592u), 256, *s, dn, up, g_prs.h_prm.p.gridSize.x,
                     c->peer_dn ? mb_flag(c->peer_dn, c, 1, 1) : (unsigned *)nullptr,
                     c->peer_up ? mb_flag(c->peer_up, c, 0, 1) : (unsigned *)nullptr, q, slab_range_ptr(s), slab_range_tile0(s));
      if (slab_range_ptr(s)) slab_range_written(s);
    } else {
      prs_slab_halo_pack(s, dn, up);
      prs_slab_signal(c->peer_dn ? mb_flag(c->peer_dn, c, 1, 1) : nullptr, c->peer_up ? mb_flag(c->peer_up, c, 0, 1) : nullptr, q);
    }
    const bool split = c->overlap_exchange && (c->peer_dn || c->peer_up);
    if (split) {
      /* interior first: after a binned sort the owned rows' table entries exist already; otherwise (onesweep route,
       * steps without a sort) the table is built by cell_table below, which needs the halo — no overlap then */
      if (g_prs.slab_binned && g_prs.slab_table_fresh) {
        StageScope t(PRS_STAGE_COLLIDE);
        slab_collide_band(s, dt, 1);
      } else {
        c->split_fallbacks++;
      }
    }
    const bool interior_done = split && g_prs.slab_binned && g_prs.slab_table_fresh;
    if (c->fused_exchange) { /* one kernel waits for the flags, unpacks the halos and clears the halo rows of the table */
      {
        StageScope t(PRS_STAGE_EXCHANGE);
        const SlabClears cl = slab_table_clears(s);
        PRS_LAUNCH_PDL(k_slab_halo_unpack_wait, min(div_up(2 * s->halo_cap, 256), 592u), 256, *s, mb_buf(c->mailbox, c, 0, 1, par),
                       mb_buf(c->mailbox, c, 1, 1, par), (const unsigned *)(c->peer_dn ? mb_flag(c->mailbox, c, 0, 1) : nullptr),
                       (const unsigned *)(c->peer_up ? mb_flag(c->mailbox, c, 1, 1) : nullptr), q, cl);
      }
      StageScope t(PRS_STAGE_REORDER);
      slab_table_fill(s);
    } else {
      slab_wait_drop(s, c->peer_dn ? mb_flag(c->mailbox, c, 0, 1) : nullptr, c->peer_up ? mb_flag(c->mailbox, c, 1, 1) : nullptr, q,
                     mb_buf(c->mailbox, c, 0, 1, par), mb_buf(c->mailbox, c, 1, 1, par));
      prs_slab_halo_unpack(s, mb_buf(c->mailbox, c, 0, 1, par), mb_buf(c->mailbox, c, 1, 1, par));
      prs_slab_cell_table(s);
    }
    StageScope t(PRS_STAGE_COLLIDE);
    if (interior_done) {
      slab_collide_band(s, dt, 2);
      slab_collide_band(s, dt, 3);
    } else {
      slab_collide_band(s, dt, 0);
    }
  }
  c->time = time + dt;
  /* sticky error bits -> pinned host word, asynchronously */
  if (!c->h_err) {
    PRS_CUDA(cudaMallocHost((void **)&c->h_err, sizeof(unsigned)));
    *c->h_err = 0u;
  }
  PRS_CUDA(cudaMemcpyAsync(c->h_err, s->counts + PRS_SC_ERR, sizeof(unsigned), cudaMemcpyDeviceToHost, g_prs.stream));
  return *c->h_err;
}
void prs_slab_ctx_release(prs_slab_ctx *c) {
  if (c->h_err) { PRS_CUDA(cudaStreamSynchronize(g_prs.stream)); PRS_CUDA(cudaFreeHost(c->h_err)); c->h_err = nullptr; }
}

}  // extern "C"
