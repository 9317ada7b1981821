/*
 * prs_patchlist.cuh — the list of cell patches that hold robots (work list of k_collide_patch,
 * prs_collide_patch.cuh), filled by the kernels that build the cell table.
 */
#pragma once
#include <stdint.h>

namespace prs {

constexpr int PATCH_W = 16; /* cells per patch row; a patch is PATCH_W columns x PH rows (PH chosen per launch) */

/* "this patch holds robots": called by the kernels that build the cell table, once per non-empty
 * cell group; the first caller of a patch in this step (epoch) appends it to the list.  Warp-aggregated:
 * one atomicAdd on the counter per warp. */
__device__ __forceinline__ void patch_mark(bool nonempty, uint32_t cell, uint32_t log2_gx, uint32_t PH, uint32_t *patchEpoch,
                                           uint32_t *patchList, uint32_t *patchCount, uint32_t epoch) {
  const unsigned active = __activemask();
  uint32_t p = 0u;
  bool first = false;
  if (nonempty) {
    const uint32_t cx = cell & ((1u << log2_gx) - 1u), cy = cell >> log2_gx;
    p = (cy / PH) * ((1u << log2_gx) / PATCH_W) + cx / PATCH_W;
    first = atomicExch(&patchEpoch[p], epoch) != epoch;
  }
  const unsigned fm = __ballot_sync(active, first);
  if (fm) {
    const uint32_t lane = threadIdx.x & 31u;
    const int leader = __ffs(fm) - 1;
    uint32_t base = 0u;
    if ((int)lane == leader) base = atomicAdd(patchCount, (uint32_t)__popc(fm));
    base = __shfl_sync(active, base, leader);
    if (first) patchList[base + (uint32_t)__popc(fm & ((1u << lane) - 1u))] = p;
  }
}

}  // namespace prs
