/*
 * prs_cabi.h — C-ABI of libparticlebot_b200.so (hand-written sm_100a CUDA behind plain pointers).
 *
 * Part 1 re-declares, name for name and argument for argument, the `extern "C"` launch wrappers
 * the reference's host class binds (declared particlebot.cuh:15-121, defined
 * particlebot_cuda.cu:26-384).  A maintainer of the reference links particlebot.cpp against this
 * library INSTEAD of particlebot_cuda.o and nothing else changes.  Conventions
 * kept from the reference: every pointer is a DEVICE pointer owned by the caller unless it says
 * host; pos/vel arrays are float2-per-robot, rad/phase/absForce are float-per-robot, all in
 * ORIGINAL robot order; no return codes — a CUDA error prints to stderr and exit(EXIT_FAILURE)s
 * (include/helper_cuda.h:999-1031); work is issued on one stream (default: the legacy default
 * stream) and nothing synchronises the host except threadSync/copyArrayFromDevice/copyArrayToDevice.
 *
 * Part 2 is what the reference does not have: stream selection, a parametric world wall, the
 * fused whole-step path, device-side light-distance reduction, and the simulation-object API
 * (the C++ class Particlebot of include/prs_particlebot.hpp wrapped for C/ctypes callers).
 */
#ifndef PRS_CABI_H
#define PRS_CABI_H

#include <stddef.h>
#include "prs_simparams.h"

/* the library is built with -fvisibility=hidden: only what this header declares is exported */
#pragma GCC visibility push(default)
#ifdef __cplusplus
extern "C" {
#endif

struct cudaGraphicsResource;
struct curandStateXORWOW; /* 48-byte generator state of the toolkit's device API */

/* ------------------------------------------------------------------------------------------
 * Part 1 — the reference's entry points
 * ------------------------------------------------------------------------------------------ */
void cudaInit(int argc, char **argv);                 /* particlebot_cuda.cu:29  (-device=N honoured) */
void cudaGLInit(int argc, char **argv);               /* :43 headless: same as cudaInit */
void allocateArray(void **devPtr, size_t size);       /* :49 */
void freeArray(void *devPtr);                         /* :54 */
void threadSync(void);                                /* :59 */
void copyArrayToDevice(void *device, const void *host, int offset, int size);  /* :64 */
/* :95 — a non-null resource means "map the GL buffer first" in the reference; headless builds
 * have no GL buffers, so a non-null resource is an error (exit), never silently ignored. */
void copyArrayFromDevice(void *host, const void *device, struct cudaGraphicsResource **res, int size);
/* GL interop (:69-93): present so that the reference host code links; they abort when called
 * because this library is built without OpenGL (rendering is optional, north_star (5)). */
void registerGLBufferObject(unsigned vbo, struct cudaGraphicsResource **res);
void unregisterGLBufferObject(struct cudaGraphicsResource *res);
void *mapGLBufferObject(struct cudaGraphicsResource **res);
void unmapGLBufferObject(struct cudaGraphicsResource *res);

void setParameters(SimParams *hostParams);            /* :111 params + 7 obstacle arrays -> constant memory */

unsigned iDivUp(unsigned a, unsigned b);              /* :126 */
#ifdef __cplusplus
/* :132, :139 — exported by the reference's translation unit with C linkage although no header declares them */
void computeGridSize(unsigned n, unsigned blockSize, unsigned &numBlocks, unsigned &numThreads);
void computeGridSize2(unsigned n, unsigned blockSize, unsigned &numBlocks, unsigned &numThreads);
#endif

void integrateSystem(float *pos, float *vel, float *rad, float deltaTime, unsigned nCells,
                     float time);                     /* :145 + kernel_impl.cuh:53-103 */
void calcHash(unsigned *gridParticlebotHash, unsigned *gridParticlebotIndex, float *pos,
              int nCells);                            /* :162 + kernel_impl.cuh:446-465 */
void sortParticlebots(unsigned *dGridParticlebotHash, unsigned *dGridParticlebotIndex,
                      unsigned nCells);               /* :377 stable sort by key, in place */
void reorderDataAndFindCellStart(unsigned *cellStart, unsigned *cellEnd, float *sortedPos,
                                 float *sortedVel, float *sortedRad, unsigned *gridParticlebotHash,
                                 unsigned *gridParticlebotIndex, float *oldPos, float *oldVel,
                                 float *oldRad, unsigned nCells, unsigned numCells); /* :284 */
void collide(float *newVel, float *absForce_a, float *absForce_r, float *sortedPos,
             float *sortedVel, float *sortedRad, unsigned *gridParticlebotIndex,
             unsigned *cellStart, unsigned *cellEnd, unsigned nCells, unsigned numCells,
             float deltaTime);                        /* :331 + kernel_impl.cuh:541-831 */
void updateRad_light_wave(float *pos, float *absForce_a, float *absForce_r, float *rad,
                          float *phase, float time, float deltaTime, int *dead, int nCells); /* :180 */
void updatePhase(float *pos, float *phase, float spacing, float max_d, float min_d, int nCells); /* :210 */
void curand_setup(struct curandStateXORWOW *state, int N);                      /* :198 */
void add_normal_noise(struct curandStateXORWOW *state, float *val, float std, int N); /* :204 */
/* :241 centroid of pos[0..n) scaled by 1/n, y tagged +2000, stored at pos[n + ((int)(time/hist_int))%hist_steps]
 * (the render trail).  temppos/temppos1 are scratch of >= n float2 as in the reference. */
void calcCOG(float *pos, float *temppos, float *temppos1, int nCells, float time, int hist_steps,
             float hist_int);
/* :227 colours for the point-sprite renderer */
void updateCol(float *rad, float *col, int nCells, float *pos, float *phase, int *dead);

/* ------------------------------------------------------------------------------------------
 * Part 2 — B200 additions
 * ------------------------------------------------------------------------------------------ */
const char *prs_version(void);
void prs_set_stream(void *cuda_stream);   /* all later launches/copies go to this cudaStream_t */
void *prs_get_stream(void);
/* wall position of integrate; the reference hard-codes 64 (kernel_impl.cuh:75-97) */
void prs_set_world_half_extent(float half);
float prs_get_world_half_extent(void);
/* collide runs one WARP per robot for swarms of up to max_robots (latency-bound sizes; default 16384,
 * 0 = always one thread per robot).  Same bits either way. */
void prs_set_collide_warp_max(unsigned max_robots);
/* 1 = on sort steps of plain swarms (no transported object, absForce_a not wanted) collide runs as k_collide_patch
 * (csrc/prs_collide_patch.cuh): a block owns a patch of 16 x 8 cells, stages it and its 2-cell halo in shared memory
 * with 1-D TMA bulk copies (cp.async.bulk + mbarrier) and evaluates every pair of two patch robots ONCE, parking the
 * force for the partner so that each robot still adds its forces in the reference's order.  Same bits either way;
 * default 0 (measured slower than the thread-per-robot kernel, profiles/r2_collide_patch.md). */
void prs_set_collide_tile(int on);
/* height of a patch in cells (1..8, default 8; the width is 16): smaller for denser swarms so that a patch's robots fit one block */
void prs_set_patch_rows(unsigned rows);
/* tuning aid: counts {patches taken, patches handed to the per-robot slow lane, patches redone} while on; out (host, 3 words, may be NULL) */
void prs_patch_stats(int on, unsigned *out);
int prs_get_collide_tile(void);
/* 1 = the kernels of prs_fused_step are launched with programmatic dependent launch (each kernel's blocks
 * become resident while the previous kernel drains; griddepcontrol.wait orders the data).  Same results. */
void prs_set_pdl(int on);
/* 1 (default) = K1 of the fused binned step handles two robots per thread with 128-bit position / velocity and 64-bit
 * scalar accesses (swarms of >= 65536 robots with 16-byte aligned arrays); 0 = one robot per thread.  Same results. */
void prs_set_k1_x2(int on);
/* 1 (default) = on binned sort steps of plain swarms collide takes the slot range of a stencil row from the dense start table
 * the scan writes next to cellStart / cellEnd (two words per row instead of six); 0 = always from cellStart / cellEnd.  Same results. */
void prs_set_collide_dense(int on);
/* slab ranks, binned sort: 1 (default) the scan over the owned rows' cells skips the tiles outside the range the slab's
 * robots occupied after the previous sort, widened by a row of movement and the reach of a stencil; tickets taken outside
 * it make that step scan everything (device-side, csrc/prs_cellbin.cuh).  Same results. */
void prs_set_slab_scan_range(int on);
/* steps without a sort: swarms of up to max_robots run controller+integrate and the gather into the sorted
 * copy as ONE kernel (one launch less per step; default 65536, 0 = never).  Same bits. */
void prs_set_fuse_gather_max(unsigned max_robots);
int prs_get_pdl(void);
/* building blocks of the host-buffer step (Particlebot::updateHost / prs_sim_update_host): asynchronous copies
 * between PINNED host memory and the device around prs_fused_step; the device-to-host copies of positions and
 * radii wait only for K1 (controller + integrate) and run under the sort and collide kernels on a second stream */
void prs_h2d_async(void *device, const void *host, size_t bytes);
void prs_arm_k1_event(int on);
void prs_d2h_async(void *host, const void *device, size_t bytes, int after_k1);
void prs_host_step_sync(void);
/* The host-buffer step as a PIPELINE (binned sort steps of swarms of 2^18 robots and more): the next prs_fused_step takes
 * positions, velocities and radii of robots [0, n) from the given pinned host buffers in two chunks — upload of chunk c,
 * K1 on chunk c, and on the second stream the way back of chunk c's new positions and radii — so that the device-to-host
 * direction of the link is busy while the host-to-device direction still is.  A step on another route (first steps,
 * crowded swarms, steps without a sort, small swarms) uploads everything first and sends positions and radii back after K1
 * as prs_d2h_async(.., 1) would.  Either way positions and radii are on their way when prs_fused_step returns; the caller
 * adds the velocities (prs_d2h_async(vel, .., 0)) and prs_host_step_sync.  These uploads are the evolving state of the SAME
 * swarm: unlike prs_h2d_async / setArray they do not withdraw the binned route's admission (the device-side guard of the
 * in-cell ranking stays in force).  Cleared by the step that consumes it. */
void prs_host_step_plan(const float *pos_in, const float *vel_in, const float *rad_in, float *pos_out, float *rad_out);
void prs_set_plan_chunks(unsigned chunks); /* chunks of that pipeline (0 = default 2; 1 = no overlap of the two directions) */
/* number of kernels this library launched since the last reset (bench.py's gpu_launches) */
unsigned long long prs_launch_count(int reset);

/* per-stage CUDA-event timing of prs_fused_step on the launching stream: stages are
 * 0 controller+integrate(+hash), 1 sort, 2 reorder+cell table, 3 collide, 4 phase update.
 * prs_stage_times synchronises, writes summed ms and span counts (5 entries each) and clears. */
void prs_stage_timing(int enable);
void prs_stage_times(float *ms, unsigned *counts);

/* device-side replacement of the host loop particlebot.cpp:214-228: d_min_d[0] = min_i |light-p_i| */
void prs_min_light_distance(const float *pos, int n, float *d_min_d);
/* updatePhase reading min_d from device memory (no host round trip) */
void prs_update_phase_dev(const float *pos, float *phase, float spacing, const float *d_min_d, int n);
/* swarm centroid sum_i pos_i / n into d_out[0..1] (observable; deterministic two-stage tree) */
void prs_centroid(const float *pos, int n, float *d_scratch, float *d_out);

/* sort with caller-visible key width: keys must be < 2^key_bits; out-of-place when out_* != in_* */
void prs_sort_pairs(const unsigned *in_keys, const unsigned *in_vals, unsigned *out_keys,
                    unsigned *out_vals, unsigned n, int key_bits);

/* tuning aid: per-tile phase stamps of the sort kernels (8 x uint64 nanoseconds per tile and pass) */
void prs_sort_set_timeline(unsigned long long *device_buf);
unsigned prs_sort_tile_size(void); /* pairs per tile of the last sort */
/* the digits prs_sort_pairs uses for n pairs of key_bits-bit keys: returns the number of passes; out[0..3] = bits per
 * pass (8, or 9 where 9-bit digits save a whole pass), out[4] = pairs per thread, out[5] = 0 */
int prs_sort_plan(int key_bits, unsigned n, int *out);
void prs_sort_set_threads(int threads_per_tile); /* 512, 1024, or 0 = chosen by size (default) */

/* Fused whole step on the reference's buffers (same observable results as the call sequence
 * updateRad_light_wave -> integrateSystem -> [calcHash, sortParticlebots] ->
 * reorderDataAndFindCellStart -> collide of particlebot.cpp:238-296):
 *   do_sort != 0 on steps where the sort gate fires. */
typedef struct {
  float *pos, *vel, *rad, *phase, *absForce_a, *absForce_r; int *dead;   /* original order */
  unsigned *hash, *index, *cellStart, *cellEnd;                          /* grid tables */
  float *sortedPos, *sortedVel, *sortedRad;                              /* sorted copies */
  unsigned nCells, numCells;
  /* optional: float4 {x, y, radius, original-index bits} per sorted slot.  When non-null the fused
   * step fills THIS instead of sortedPos/sortedRad (prs_unpack_sorted converts on demand) and,
   * unless constrained_contraction is set, does not produce absForce_a (nothing reads it then). */
  float *sortedPR;
} prs_step_buffers;
void prs_fused_step(const prs_step_buffers *b, float time, float deltaTime, int do_sort);
/* The fused step sorts by cell BINNING (counting sort whose histogram scan is the cell table,
 * csrc/prs_cellbin.cuh) while the swarm is known to be sparse (<= 64 robots in the fullest cell,
 * tracked asynchronously) and numCells <= 16 * nCells; otherwise by the onesweep radix sort.  Both
 * give identical results.  Callers that rewrite positions through their own copies call
 * prs_bin_invalidate(); prs_bin_set_mode: 0 auto (default), 1 onesweep only, 2 binning only. */
void prs_bin_invalidate(void);
void prs_bin_set_mode(int mode);
int prs_bin_active(void);

/* ---- slab (multi-GPU) engine: the fused step cut where ranks exchange robots, every count kept
 * on the device (csrc/prs_slab.cuh; particlerobotsimulations_b200/multigpu.py drives it;
 * DESIGN.md "multi-GPU").  A rank owns grid rows [row_lo, row_hi). ---- */
enum { /* words of prs_slab.counts (device uint32[16]) */
  PRS_SC_N = 0,        /* robots owned (local slots [0, n)) */
  PRS_SC_NLO = 1, PRS_SC_NHI = 2,     /* halo robots received from the lower / upper neighbour */
  PRS_SC_KDN = 3, PRS_SC_KUP = 4,     /* halo robots sent down / up */
  PRS_SC_MIGDN = 5, PRS_SC_MIGUP = 6, /* robots migrating down / up this step */
  PRS_SC_LEAVERS = 7, PRS_SC_HOLES = 8, PRS_SC_KEEPERS = 9, /* compaction scratch */
  PRS_SC_ERR = 10,     /* sticky error bits, PRS_SLAB_ERR_* */
  PRS_SC_STAT_MIG = 11, PRS_SC_STAT_HALO = 12 /* running totals */
};
enum {
  PRS_SLAB_ERR_MIG_CAP = 1,    /* more migrants than mig_cap in one step */
  PRS_SLAB_ERR_HALO_CAP = 2,   /* a halo longer than halo_cap */
  PRS_SLAB_ERR_CAPACITY = 4,   /* more robots than cap */
  PRS_SLAB_ERR_TWO_SLABS = 8,  /* a robot crossed more than one slab between two sorts */
  PRS_SLAB_ERR_LEFT_WORLD = 16, /* a robot left the rows of the first / last slab (hash wrap-around) */
  PRS_SLAB_ERR_PEER_TIMEOUT = 32, /* peer-to-peer exchange: the neighbour's data did not arrive */
  PRS_SLAB_ERR_DRIFT = 64      /* between two sorts an owned robot drifted so far that its 5x5 stencil leaves the rows
                                  this rank holds (owned + halo_rows): raise halo_rows or sort more often */
};
typedef struct {
  /* owned robots, local slots [0, cap) */
  float *pos, *vel, *rad, *phase, *absForce_a, *absForce_r;
  int *dead;
  unsigned *gid;       /* global robot id (original index of the single-GPU run) */
  void *rng;           /* curandStateXORWOW[cap], seeded by global id */
  unsigned *hash;      /* cell key per local slot */
  unsigned *scratch;   /* uint32[cap] */
  /* sorted view, [halo_cap | cap | halo_cap] slots: lower halo ends at halo_cap, owned start there */
  float *sortedPR, *sortedVel;
  unsigned *hash_cat, *index_sorted;
  unsigned *cellStart, *cellEnd;    /* whole-grid tables (only this rank's rows are used) */
  unsigned *counts;    /* device uint32[16], PRS_SC_* */
  unsigned *lists;     /* device uint32[6 * mig_cap] */
  unsigned cap, halo_cap, mig_cap;
  unsigned row_lo, row_hi, halo_rows;
  int has_dn, has_up;  /* neighbours below (rank - 1) / above (rank + 1) exist */
  int wrap;            /* 1: the slabs form a RING in the row index (the cell hash wraps around the grid): the first
                          slab's lower neighbour is the last one and vice versa; rows are compared cyclically */
} prs_slab;
/* exchange buffers are uint32 arrays: word 0 = record count, then structure-of-arrays records */
size_t prs_slab_mig_words(unsigned mig_cap);
size_t prs_slab_halo_words(unsigned halo_cap);
void prs_slab_rng_setup(const prs_slab *s, unsigned n);
void prs_slab_k1(const prs_slab *s, float time, float dt, int do_hash);
void prs_slab_migrate_pack(const prs_slab *s, unsigned *send_dn, unsigned *send_up);
void prs_slab_migrate_unpack(const prs_slab *s, const unsigned *recv_dn, const unsigned *recv_up);
void prs_slab_sort(const prs_slab *s);
void prs_slab_gather(const prs_slab *s);
void prs_slab_halo_pack(const prs_slab *s, unsigned *send_dn, unsigned *send_up);
void prs_slab_halo_unpack(const prs_slab *s, const unsigned *recv_dn, const unsigned *recv_up);
void prs_slab_cell_table(const prs_slab *s);
void prs_slab_collide(const prs_slab *s, float dt);
void prs_slab_min_light_distance(const prs_slab *s, float *d_min_d);
void prs_slab_update_phase(const prs_slab *s, float spacing, const float *d_min_d);
void prs_slab_add_noise(const prs_slab *s, float std);
/* peer-to-peer exchange (multigpu.py exchange="p2p"): the send buffers handed to the pack calls are
 * the NEIGHBOUR's mailbox, mapped through CUDA IPC; prs_slab_signal publishes the sequence number
 * in the neighbours' flag words after the pack kernel, prs_slab_wait holds the stream until both
 * neighbours' flags have reached it (device-side spin, bounded) */
size_t prs_ipc_handle_size(void);
void *prs_slab_mailbox_alloc(size_t words);
void prs_slab_mailbox_free(void *p);
void prs_ipc_export(void *dev_ptr, void *handle_out);
void *prs_ipc_open(const void *handle);
void prs_ipc_close(void *p);
void prs_slab_signal(unsigned *remote_flag_dn, unsigned *remote_flag_up, unsigned seq);
void prs_slab_wait(const prs_slab *s, const unsigned *local_flag_dn, const unsigned *local_flag_up, unsigned seq);

/* One whole step of a slab rank with the peer-to-peer exchange, orchestrated by the library (what Particlebot::update,
 * particlebot.cpp:170-300, is for one GPU).  The caller sets the context up once — the slab, its own mailbox
 * (prs_slab_mailbox_alloc(prs_slab_mailbox_words(..))) and the two neighbours' mailboxes mapped through CUDA IPC, NULL
 * where there is no neighbour — and supplies the one collective the path has: allreduce_min(dev, user) must replace
 * dev[0] (device memory, one float) by its minimum over all ranks, ordered on the library's stream; it is called on the
 * steps where the phase gate fires (every phase_update_interval of simulated time).  overlap_exchange != 0: collide of
 * the interior rows is launched before the wait for the neighbours' halos, the two edge bands after it.
 * Returns the sticky error bits (PRS_SLAB_ERR_*) seen so far, at most one step late, without synchronising. */
typedef struct {
  prs_slab slab;
  unsigned *mailbox, *peer_dn, *peer_up;   /* own mailbox; the lower / upper neighbour's, mapped (NULL: none) */
  unsigned mw, hw;                         /* prs_slab_mig_words(mig_cap), prs_slab_halo_words(halo_cap) */
  unsigned *scratch_mig[2], *scratch_halo[2]; /* local send buffers for the sides without a neighbour (mw / hw words) */
  float *d_min_d;                          /* device, >= 1 float */
  void (*allreduce_min)(float *dev, void *user);
  void *user;
  int overlap_exchange;
  /* state, owned by the library after the first call */
  float time;
  int sorted_once;
  unsigned seq_halo, seq_mig, split_fallbacks;
  unsigned *h_err;
  /* 1: the exchange kernels of the step in their fused forms — the rest of the migration after the leavers are packed as
   * ONE single-block kernel, the flags published by the last block of the halo pack kernel, the wait for the neighbours
   * folded into the unpack kernel together with the clearing of the table's halo rows (15 launches per sort step instead
   * of 25); 0: one kernel per operation.  Same results. */
  int fused_exchange;
} prs_slab_ctx;
size_t prs_slab_mailbox_words(unsigned mig_cap, unsigned halo_cap);
unsigned prs_slab_step(prs_slab_ctx *c, float dt, float sort_interval);
void prs_slab_ctx_release(prs_slab_ctx *c);

/* synthetic hex block placed by a kernel (the same bits as Particlebot::initHexBlock's host loop): n = nx * ny robots on a hex
 * lattice centred on the origin, jitter from a counter hash of (seed, robot), velocity 0, radius min_radius, phase 0, alive */
void prs_init_hex_block(float *pos, float *vel, float *rad, float *phase, int *dead, unsigned n, unsigned nx, unsigned ny,
                        float pitch, float jitter, unsigned seed, float min_radius);
void prs_unpack_sorted(const float *sortedPR, float *sortedPos, float *sortedRad, unsigned n);
/* self-test: number of operand pairs for which the shared-reciprocal division used by collide
 * differs from __fdiv_rn (must be 0) */
unsigned long long prs_selftest_div(const float *d_x, const float *d_d, unsigned n);

/* ---- headless frames and video (csrc/prs_frame.cuh, csrc/prs_video.cpp) ----
 * prs_render_frame replaces, without OpenGL, what one displayed frame of the reference consists of: the scene of
 * main.cpp:366-466 (floor, light marker, obstacles), the point-sprite pass of render.cpp:53-125 / shaders.cpp:40-86 (one
 * flat disc per entry of the position buffer in the colour updateCol gave it; entries with y > 1000 are the centroid
 * trail, kernel_impl.cuh:343) and PostprocessKernel's read-back (postprocess.cu:32-55: rows top-down, bytes B, G, R).
 * The view is the reference's camera looking straight down: world_per_pixel = 2 camera_y tan(30 deg) / height
 * (prs_view_from_camera).  d_bgr: device, width*height*3 bytes; d_keys: device scratch, 2*width*height words;
 * pos / rad / col: device, n_points entries (robots, then the trail; col is float4 per entry). */
typedef struct {
  unsigned width, height;
  float center_x, center_y;
  float world_per_pixel;
  float light_radius;
} prs_view;
void prs_view_from_camera(prs_view *v, unsigned width, unsigned height, float camera_y, float light_radius);
void prs_render_frame(unsigned char *d_bgr, unsigned *d_keys, const prs_view *view, const float *pos, const float *rad,
                      const float *col, unsigned n_points);
/* Video file of such frames — what cv::VideoWriter does in postprocess.cu:101-118 (20 frames per second, one frame every
 * VIDEO_INTERVAL displayed frames), as an uncompressed AVI (BI_RGB, 24 bit) that needs no codec library.  Frames are host
 * memory, top-down B, G, R rows as prs_render_frame produces them.  A RIFF file ends at 4 GiB: prs_video_write returns -1
 * once the next frame would not fit (the file stays valid).  No GPU involved. */
typedef struct prs_video prs_video;
prs_video *prs_video_open(const char *path, unsigned width, unsigned height, double fps);
int prs_video_write(prs_video *v, const unsigned char *bgr_top_down);
int prs_video_close(prs_video *v); /* patches the header counts; returns the number of frames written, -1 on I/O error */

/* ---- simulation object (class Particlebot, include/prs_particlebot.hpp) for C callers ---- */
typedef struct prs_sim prs_sim;
/* fills *p with the defaults of main.cpp:833-911 and the derived grid of :932-939; extras receive
 * timestep, sort_interval, dump_interval (floats).  Obstacle arrays are owned by the library. */
typedef struct {
  float timestep, sort_interval, dump_interval;
  float camera_x, camera_y, light_radius;
  int display_interval, video_interval;
  char csv_filename[300], video_filename[300];
  /* extension keys of this repository's parser (the reference skips unknown names together with their value line, so a
   * cfg that uses them still loads there): init_config = random | grid | hex | line | hexblock; hexblock_nx, hexblock_ny,
   * hexblock_pitch, hexblock_jitter (fraction of max_radius), hexblock_seed; world_half (wall of integrate, reference: 64),
   * grid_dim (cells per axis, power of two, reference: 512) */
  int init_hexblock;             /* 1: place the swarm with initHexBlock instead of reset() */
  unsigned hexblock_nx, hexblock_ny, hexblock_seed;
  float hexblock_pitch, hexblock_jitter;
  float world_half;              /* 0 = the reference's 64 */
  unsigned grid_dim;             /* 0 = the reference's 512 */
} prs_run_options;
void prs_params_defaults(SimParams *p, prs_run_options *opt);
/* parses a .cfg file with the reference's grammar and quirks (main.cpp:594-816, 913-928) on top
 * of *p / *opt, then derives cellSize/gridSize/numCells/worldOrigin.  Returns 0, or -1 if the
 * file cannot be opened (the reference silently runs with defaults then). */
int prs_params_load_cfg(const char *path, SimParams *p, prs_run_options *opt);
void prs_params_derive_grid(SimParams *p);
/* synthetic worlds of SURVEY.md §8d: grid_dim cells per axis (power of two), world half extent */
void prs_params_set_world(SimParams *p, unsigned grid_dim, float world_half);

/* ---- slab runs without Python: one PROCESS per GPU, forked by the caller's process (csrc/prs_multi.cpp) ----
 * What `ParticleBot <cfg> --gpus N` runs (the north_star's "host side is C++"): the calling process forks
 * `gpus` ranks BEFORE any CUDA call; each rank picks its device, builds its slab of the swarm (initial state: the cfg's
 * placement computed identically on every rank, or the init_config = hexblock generator for its own lattice rows; slab
 * boundaries by equal robot count), maps its two neighbours' mailboxes through CUDA IPC and then calls prs_slab_step per
 * step.  The control plane is a block of process-shared memory: a barrier, the IPC handles, the MIN of one float every
 * phase_update_interval, and — at dump times only — the swarm gathered in robot order so that rank 0 writes the reference's
 * CSV (particlebot.cpp:303-367) byte for byte as one GPU would.  No NCCL, no MPI, no interpreter on the path.
 * Returns 0, or 1 if a rank failed (message on stderr). */
typedef struct {
  int gpus;                 /* number of ranks */
  int oversubscribe;        /* 1: rank r runs on device r % (devices present) — several ranks per GPU (tests on one GPU) */
  long steps;               /* number of steps; < 0: until time > max_time */
  int csv;                  /* rank 0 writes opt->csv_filename */
  int quiet;
  const char *final_state;  /* optional file: rank 0 writes nCells, then pos, vel, rad, phase of the whole swarm in robot order */
  int overlap_exchange;     /* prs_slab_ctx.overlap_exchange */
  int fused_exchange;       /* prs_slab_ctx.fused_exchange */
} prs_multi_options;
int prs_multi_run(const SimParams *p, const prs_run_options *opt, const prs_multi_options *m);
/* position of robot i of the synthetic hex block (the generator of Particlebot::initHexBlock / prs_init_hex_block) */
void prs_hex_block_position(unsigned long long i, unsigned nx, unsigned ny, float pitch, float jitter, unsigned seed, float *xy);

/* backend: 0 = fused native path, 1 = native per-call path (the reference's call sequence on this
 * library's entry points), 2 = per-call path on an external library with the reference's ABI
 * (path given; used to drive oracle/_ref/libprs_refcuda.so from tests and bench.py) */
prs_sim *prs_sim_create(const SimParams *p, float world_half, int backend, const char *ext_lib);
void prs_sim_destroy(prs_sim *s);
void prs_sim_srand(prs_sim *s, unsigned seed);            /* main.cpp:929 */
void prs_sim_reset(prs_sim *s);                           /* Particlebot::reset */
void prs_sim_init_hex(prs_sim *s, unsigned nx, unsigned ny, float pitch, float jitter, unsigned seed); /* synthetic swarms */
int prs_sim_update(prs_sim *s, float dt, float sort_interval); /* Particlebot::update; returns 1 when time > max_time */
/* One step with the state held by the HOST: positions, velocities and radii (original robot order, the layouts of
 * getArray/setArray) go up from the in buffers, Particlebot::update runs, the new values come back in the out
 * buffers (may alias the in buffers).  Same results as setArray x3, update, getArray x3; the copies are
 * asynchronous and the downloads of positions and radii overlap the sort and collide kernels (pinned buffers). */
int prs_sim_update_host(prs_sim *s, const float *pos_in, const float *vel_in, const float *rad_in, float *pos_out,
                        float *vel_out, float *rad_out, float dt, float sort_interval);
float prs_sim_time(const prs_sim *s);
void prs_sim_sync(prs_sim *s);
/* which: ParticlebotArray values, plus 100 absForce_a, 101 absForce_r, 102 hash, 103 index,
 * 104 cellStart, 105 cellEnd, 106 sortedPos, 107 sortedVel, 108 sortedRad, 109 rng state */
void *prs_sim_device_ptr(prs_sim *s, int which);
void prs_sim_get(prs_sim *s, int which, void *host, size_t bytes);
void prs_sim_set(prs_sim *s, int which, const void *host, size_t offset_bytes, size_t bytes);
/* dumpParticlebot (particlebot.cpp:303-367): CSV row when the dump gate fires */
void prs_sim_dump(prs_sim *s, void *FILE_ptr, float dump_interval, unsigned testing);
/* loadFromFile (particlebot.cpp:369-411): time, positions, velocities, radii from the LAST row of a testing=1 CSV */
void prs_sim_load(prs_sim *s, void *FILE_ptr);
/* full binary checkpoint (Particlebot::saveCheckpoint / loadCheckpoint): a restored simulation of the same
 * shape continues bit for bit, on either sort cadence.  0 on success, -1 on I/O error or shape mismatch. */
int prs_sim_checkpoint_save(prs_sim *s, const char *path);
int prs_sim_checkpoint_load(prs_sim *s, const char *path);
/* Particlebot::renderFrame: colours (updateCol) + prs_render_frame + copy to host memory owned by the object (valid until
 * the next call); width*height*3 bytes, top-down B, G, R */
const unsigned char *prs_sim_render_frame(prs_sim *s, const prs_view *view);

#ifdef __cplusplus
}
#endif
#pragma GCC visibility pop
#endif /* PRS_CABI_H */
