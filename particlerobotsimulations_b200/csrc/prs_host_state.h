/*
 * prs_host_state.h — host-side state of libparticlebot_b200 (stream, parameter shadow, sort
 * scratch, launch counter) and the error macro.  One context per process, like the reference's
 * global `__constant__ params` (particlebot_kernel_impl.cuh:27).
 */
#pragma once
#include <cuda_runtime.h>
#include "prs_device.cuh"
#include "prs_onesweep.cuh"
#include <vector>

/* guard state of the binned (counting-sort) route of the fused step, see prs_fused_step */
struct PrsBinState {
  uint32_t *cellCount = nullptr, *scratch = nullptr;
  uint32_t *dense = nullptr, *live = nullptr; /* dense start table (numCells + 1 words) and per-tile liveness, see prs_cellbin.cuh */
  uint32_t *marks = nullptr;     /* per scan tile: a robot hashed into it this step / the previous step (prs_cellbin.cuh) */
  const void *marks_table = nullptr; /* the cellStart array the previous-step marks describe */
  unsigned marks_cells = 0, marks_generation = 0;
  size_t cap_cells = 0;
  int mode = 0;               /* 0 auto, 1 never, 2 always */
  bool admitted = false;      /* the swarm is known to be sparse enough */
  unsigned generation = 0;    /* bumped whenever positions are rewritten behind the library's back */
  uint32_t *range = nullptr;  /* slab ranks: {first tile, last tile, violated, -} of the scan (prs_cellbin.cuh: range_check) */
  const void *range_table = nullptr; /* the slab geometry those words were written for: cellStart, row_lo, row_hi */
  unsigned range_row_lo = 0, range_row_hi = 0;
  uint32_t *h_report = nullptr; /* pinned: [0] largest cell population, [1] error flag */
  cudaEvent_t report_event = nullptr;
  bool report_pending = false;
  unsigned report_generation = 0;
};

/* work list of k_collide_patch (prs_collide_patch.cuh): patches that hold robots, rebuilt at every sort */
struct PrsPatchState {
  uint32_t *epoch_of = nullptr, *list = nullptr, *count = nullptr; /* device: last epoch per patch, list, counter */
  size_t cap_cells = 0;
  unsigned epoch = 0;
  unsigned rows = 8;             /* patch height PH in cells */
  int num_sms = 0;
  bool smem_opt_in = false;
  uint32_t *stats = nullptr, *stats_buf = nullptr; /* tuning aid, see prs_patch_stats */
};

/* which cell table the fused step built last (steps without a sort reuse it: same keys, same table) */
struct PrsTableState {
  const void *cellStart = nullptr, *hash = nullptr;
  unsigned n = 0, numCells = 0, generation = 0;
};

struct PrsHostState {
  cudaStream_t stream = 0;            /* legacy default stream, like the reference */
  unsigned long long launches = 0;    /* kernels launched by this library */
  PrsDevParams h_prm = {};            /* shadow of the constant block */
  bool params_set = false;
  float world_half = 64.0f;           /* reference wall (kernel_impl.cuh:75-97) */
  unsigned fuse_gather_max = 65536;   /* steps without a sort: swarms up to this size run K1 and the gather as one kernel */
  /* host-buffer step (prs_sim_update_host): second stream for the device-to-host copies that may start as soon
   * as K1 has written positions and radii, and the event K1's completion is recorded in */
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t k1_event = nullptr;
  bool k1_event_armed = false;
  /* host-buffer step as a pipeline (prs_host_step_plan): where the next fused step takes its inputs from / sends positions
   * and radii back to; chunk_events: one event per chunk of K1 */
  struct HostPlan {
    bool active = false;
    const float *pos_in = nullptr, *vel_in = nullptr, *rad_in = nullptr;
    float *pos_out = nullptr, *rad_out = nullptr;
  } plan;
  std::vector<cudaEvent_t> chunk_events, upload_events;
  cudaStream_t upload_stream = nullptr;
  cudaEvent_t step_event = nullptr;
  unsigned plan_chunks = 0;           /* chunks of the pipelined host-buffer step (0: default 2) */
  int slab_scan_range = 1;            /* slab ranks: the scan skips the tiles outside the range the slab's robots occupy */
  int collide_dense = 1;              /* binned sort steps of plain swarms: collide reads the dense start table of the scan */
  int k1_x2 = 1;                      /* K1 of the fused binned step: two robots per thread, vector accesses */
  int pdl = 1;                        /* 1: the fused step's kernels are launched with programmatic dependent launch */
  int collide_tile = 0;               /* 1: sort steps of plain large swarms use k_collide_patch (TMA-staged patches, pairs evaluated
                                         once; bit-equal, measured slower: profiles/r2_collide_patch.md) */
  unsigned collide_warp_max = 16384;  /* swarms up to this size use the warp-per-robot collide kernel */
  prs_sort::Workspace sort_ws;
  PrsBinState bin;
  PrsTableState table;
  PrsPatchState patch;
  bool slab_sorted_onesweep = false;
  bool slab_binned = false, slab_table_fresh = false; /* slab engine: route of the last sort / its table not consumed yet */
  bool slab_range_in_use = false;     /* slab engine: this sort step's tickets are checked against the scan's tile range */
  bool slab_tickets = false;          /* slab engine: this sort step's K1 took the cell tickets (binned route) */
  int sort_threads = 0;               /* tile shape of k_onesweep: 512 / 1024 threads, 0 = by size */
  unsigned sort_tile_pairs = 0;       /* pairs per tile of the last sort (threads x pairs per thread) */
  unsigned long long *sort_timeline = nullptr; /* tuning aid, see prs_sort_set_timeline */
  /* optional per-stage CUDA-event timing of the fused step (bench.py's roofline numbers) */
  bool stage_timing = false;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  struct StageSpan { int stage; cudaEvent_t a, b; };
  std::vector<StageSpan> spans;
};
enum { PRS_STAGE_K1 = 0, PRS_STAGE_SORT = 1, PRS_STAGE_REORDER = 2, PRS_STAGE_COLLIDE = 3, PRS_STAGE_PHASE = 4, PRS_STAGE_EXCHANGE = 5, PRS_NUM_STAGES = 6 };
extern PrsHostState g_prs;

void prs_fail(const char *what, cudaError_t e, const char *file, int line);
#define PRS_CUDA(call)                                                 \
  do {                                                                 \
    cudaError_t e_ = (call);                                           \
    if (e_ != cudaSuccess) prs_fail(#call, e_, __FILE__, __LINE__);    \
  } while (0)
