/*
 * prs_kernels.cu — hand-written sm_100a kernels of the per-timestep particle-robot update and the
 * C-ABI launch wrappers that carry the reference's names (include/prs_cabi.h part 1).
 *
 * Every kernel cites the reference code whose RESULT it reproduces; the implementation is new:
 *   - one constant block instead of eight symbols, one stream, no per-call allocation, no host sync;
 *   - calcHash / integrate / controller are streaming kernels (and exist fused, k_control_integrate_hash);
 *   - sort is the onesweep radix sort of prs_onesweep.cuh;
 *   - collide (prs_collide.cuh) keeps the operation order of the reference — bit-identical results — in
 *     three kernels: thread per robot through L1/L2, thread per robot with TMA-staged shared-memory
 *     windows, warp per robot for small swarms.
 *
 * Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo   (NO --use_fast_math: hashes
 * need IEEE divides, the reference is built without it, Makefile:78-84).
 */
#include <cuda_runtime.h>
#ifdef PRS_WITH_GL
#include <cuda_gl_interop.h> /* needs the platform's <GL/gl.h> */
#endif
#include <curand_kernel.h>
#include <math.h>
#include <algorithm>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "prs_cabi.h"
#include "prs_device.cuh"
#include "prs_onesweep.cuh"
#include "prs_cellbin.cuh"
#include "prs_host_state.h"

#include "prs_collide.cuh"
#include "prs_collide_patch.cuh"

using namespace prs;

/* ------------------------------------------------------------------------------------------
 * host-side state of the library (one simulation context per process, like the reference's
 * global __constant__ params)
 * ------------------------------------------------------------------------------------------ */
PrsHostState g_prs;

void prs_fail(const char *what, cudaError_t e, const char *file, int line) {
  /* error convention of the reference: message on stderr, exit(EXIT_FAILURE)
   * (include/helper_cuda.h:999-1031) */
  fprintf(stderr, "%s(%i) : CUDA error %d (%s) in %s\n", file, line, (int)e, cudaGetErrorString(e), what);
  exit(EXIT_FAILURE);
}

#define PRS_LAUNCH(kernel, grid, block, smem, ...)                                   \
  do {                                                                               \
    kernel<<<(grid), (block), (smem), g_prs.stream>>>(__VA_ARGS__);                  \
    g_prs.launches++;                                                                \
    cudaError_t e_ = cudaGetLastError();                                             \
    if (e_ != cudaSuccess) prs_fail(#kernel, e_, __FILE__, __LINE__);                \
  } while (0)

/* launch with the programmatic-dependent-launch attribute (see prs::pdl_sync): the kernel's blocks may become
 * resident while the previous kernel of the stream drains.  Only kernels that start with pdl_sync() may be
 * launched this way. */
#define PRS_LAUNCH_PDL(kernel, grid, block, ...)                                       \
  do {                                                                               \
    cudaLaunchConfig_t cfg_ = {};                                                    \
    cfg_.gridDim = dim3(grid);                                                       \
    cfg_.blockDim = dim3(block);                                                     \
    cfg_.stream = g_prs.stream;                                                      \
    cudaLaunchAttribute at_[1];                                                      \
    at_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                  \
    at_[0].val.programmaticStreamSerializationAllowed = g_prs.pdl ? 1 : 0;           \
    cfg_.attrs = at_;                                                                \
    cfg_.numAttrs = 1;                                                               \
    cudaError_t e_ = cudaLaunchKernelEx(&cfg_, kernel, __VA_ARGS__);                 \
    g_prs.launches++;                                                                \
    if (e_ != cudaSuccess) prs_fail(#kernel, e_, __FILE__, __LINE__);                \
  } while (0)

static inline unsigned div_up(unsigned a, unsigned b) { return (a + b - 1) / b; }

/* stage timing: an event pair per stage and step on the launching stream, read after a sync */
static cudaEvent_t next_event() {
  if (g_prs.ev_used == g_prs.ev_pool.size()) {
    cudaEvent_t e;
    PRS_CUDA(cudaEventCreate(&e));
    g_prs.ev_pool.push_back(e);
  }
  return g_prs.ev_pool[g_prs.ev_used++];
}
struct StageScope {
  int stage;
  cudaEvent_t a;
  explicit StageScope(int s) : stage(s), a(0) {
    if (g_prs.stage_timing) { a = next_event(); PRS_CUDA(cudaEventRecord(a, g_prs.stream)); }
  }
  ~StageScope() {
    if (g_prs.stage_timing) {
      cudaEvent_t b = next_event();
      PRS_CUDA(cudaEventRecord(b, g_prs.stream));
      g_prs.spans.push_back({stage, a, b});
    }
  }
};

/* ------------------------------------------------------------------------------------------
 * kernels
 * ------------------------------------------------------------------------------------------ */

/* cell hash per robot + identity index  (result of calcHashD, kernel_impl.cuh:446-465) */
__global__ void __launch_bounds__(256) k_calc_hash(uint32_t *__restrict__ hash, uint32_t *__restrict__ index,
                                                   const float2 *__restrict__ pos, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float2 p = pos[i];
  const int2 g = cell_of(p.x, p.y);
  hash[i] = cell_hash(g.x, g.y);
  index[i] = i;
}

/* cell table + gather into sorted order (result of reorderDataAndFindCellStartD, :469-538).
 * The previous key comes from a second (L1-resident) load instead of a shared-memory halo. */
__global__ void __launch_bounds__(256)
k_reorder(uint32_t *__restrict__ cellStart, uint32_t *__restrict__ cellEnd, float2 *__restrict__ sortedPos,
          float2 *__restrict__ sortedVel, float *__restrict__ sortedRad, const uint32_t *__restrict__ hash,
          const uint32_t *__restrict__ index, const float2 *__restrict__ pos, const float2 *__restrict__ vel,
          const float *__restrict__ rad, uint32_t n) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t h = hash[k];
  const uint32_t src = index[k];
  const uint32_t hp = (k > 0) ? hash[k - 1] : 0u;
  const float2 p = pos[src];
  const float2 v = vel[src];
  const float r = rad[src];
  if (k == 0 || h != hp) {
    cellStart[h] = k;
    if (k > 0) cellEnd[hp] = k;
  }
  if (k == n - 1) cellEnd[h] = k + 1;
  sortedRad[k] = r;
  sortedPos[k] = p;
  sortedVel[k] = v;
}

/* Fused-path variant of the gather: same cell tables, but the sorted copy is packed for the
 * collide kernel's 128-bit neighbour loads — float4 {x, y, radius, original index bits} plus the
 * float2 velocity (24 B written per robot instead of 20, one LDG.128 per neighbour instead of
 * three loads; north_star (1)). */
__global__ void __launch_bounds__(256)
k_reorder_packed(uint32_t *__restrict__ cellStart, uint32_t *__restrict__ cellEnd, float4 *__restrict__ sortedPR,
                 float2 *__restrict__ sortedVel, const uint32_t *__restrict__ hash, const uint32_t *__restrict__ index,
                 const float2 *__restrict__ pos, const float2 *__restrict__ vel, const float *__restrict__ rad,
                 uint32_t n, const prs_bin::PatchListArgs pl = prs_bin::PatchListArgs()) {
  prs::pdl_sync();
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t h = hash[k];
  const uint32_t src = index[k];
  const uint32_t hp = (k > 0) ? hash[k - 1] : 0u;
  const float2 p = pos[src];
  const float2 v = vel[src];
  const float r = rad[src];
  if (k == 0 || h != hp) {
    cellStart[h] = k;
    if (k > 0) cellEnd[hp] = k;
  }
  if (pl.list) prs::patch_mark(k == 0 || h != hp, h, pl.log2_gx, pl.PH, pl.epoch_of, pl.list, pl.count, pl.epoch);
  if (k == n - 1) cellEnd[h] = k + 1;
  sortedPR[k] = make_float4(p.x, p.y, r, __uint_as_float(src));
  sortedVel[k] = v;
}
/* Steps WITHOUT a sort (reference cadence: hash and sort only every sort_interval): hash and index are
 * unchanged, so the cell table the reference rebuilds every step (cudaMemset + the same compares,
 * particlebot_cuda.cu:301-309) comes out identical — only the sorted copy has to be refreshed. */
__global__ void __launch_bounds__(256)
k_gather_packed(float4 *__restrict__ sortedPR, float2 *__restrict__ sortedVel, const uint32_t *__restrict__ index,
                const float2 *__restrict__ pos, const float2 *__restrict__ vel, const float *__restrict__ rad, uint32_t n) {
  prs::pdl_sync();
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t src = index[k];
  const float2 p = pos[src];
  sortedPR[k] = make_float4(p.x, p.y, rad[src], __uint_as_float(src));
  sortedVel[k] = vel[src];
}

/* packed -> the reference's sortedPos / sortedRad arrays (only when a caller asks for them) */
__global__ void __launch_bounds__(256) k_unpack_sorted(const float4 *__restrict__ pr, float2 *__restrict__ sortedPos,
                                                       float *__restrict__ sortedRad, uint32_t n) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const float4 q = pr[k];
  sortedPos[k] = make_float2(q.x, q.y);
  sortedRad[k] = q.z;
}

/* bit-compare of the shared-reciprocal division against __fdiv_rn (self-test, see prs_collide.cuh) */
__global__ void k_selftest_div(const float *__restrict__ x, const float *__restrict__ d, uint32_t n,
                               unsigned long long *__restrict__ mismatches) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float r1 = prs::rcp_refined(d[i]);
  const float q = prs::div_shared(x[i], d[i], r1);
  const float w = __fdiv_rn(x[i], d[i]);
  /* x = +-0 may come out as +0: collide normalises zero signs right after (tempforce += f) */
  const bool both_zero = q == 0.0f && w == 0.0f;
  if (__float_as_uint(q) != __float_as_uint(w) && !both_zero) atomicAdd(mismatches, 1ull);
  /* d doubles as a sqrt / gap^2 operand */
  const float s = prs::sqrt_fast_path(d[i]);
  if (__float_as_uint(s) != __float_as_uint(__fsqrt_rn(d[i]))) atomicAdd(mismatches + 1, 1ull);
  const float g = prs::powf2_fast_path(d[i]);
  if (__float_as_uint(g) != __float_as_uint(__powf(d[i], 2.0f))) atomicAdd(mismatches + 2, 1ull);
  /* x / sqrt(d) with the reciprocal seeded by rsqrt(d) instead of MUFU.RCP (collide's unit vector) */
  float y;
  const float sq = prs::sqrt_fast_path(d[i], &y);
  const float r1s = fmaf(y, fmaf(y, -sq, 1.0f), y);
  const float qs = prs::div_shared(x[i], sq, r1s);
  const float ws = __fdiv_rn(x[i], sq);
  if (__float_as_uint(qs) != __float_as_uint(ws) && !(qs == 0.0f && ws == 0.0f)) atomicAdd(mismatches + 3, 1ull);
}

/* Euler step + wall bounce for one robot (result of integrate_functor, :53-103) */
__device__ __forceinline__ void integrate_one(float2 &pos, float2 &vel, float rad, float dt) {
  const float W = c_prm.world_half;
  const float bd = c_prm.p.boundaryDamping;
  pos.x += vel.x * dt;
  pos.y += vel.y * dt;
  if (pos.x > W - rad) { pos.x = W - rad; vel.x *= bd; }
  if (pos.x < -W + rad) { pos.x = -W + rad; vel.x *= bd; }
  if (pos.y > W - rad) { pos.y = W - rad; vel.y *= bd; }
  if (pos.y < -W + rad) { pos.y = -W + rad; vel.y *= bd; }
}

__global__ void __launch_bounds__(256) k_integrate(float2 *__restrict__ pos, float2 *__restrict__ vel,
                                                   const float *__restrict__ rad, float dt, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float2 p = pos[i], v = vel[i];
  integrate_one(p, v, rad[i], dt);
  pos[i] = p;
  vel[i] = v;
}

/* radius-phase controller for one robot (result of updateRad_light_wave, :124-181); returns the
 * new radius, or the old one when the robot holds its radius */
template <bool CC>
__device__ __forceinline__ float controller_one(float rad, float phase, float fr, float fa, float time, float dt) {
  const SimParams &P = c_prm.p;
  if (phase > 10000000.0f) return rad;
  const float period = (P.Nx + 1) * P.rise_period;
  float t1 = time + phase;
  if (t1 < 0) t1 = t1 + 100 * (P.Nx + 1) * P.rise_period;
  if (t1 >= period) t1 = t1 - period * floorf(t1 / period);
  if (t1 >= 2 * P.rise_period) return rad;
  float target;
  if (t1 <= P.rise_period)
    target = P.min_radius + (P.max_radius - P.min_radius) / P.rise_period * t1;
  else
    target = P.max_radius + (P.min_radius - P.max_radius) / P.rise_period * (t1 - P.rise_period);
  const float dr1 = target - rad;
  float dr = 0;
  const float max_speed = 0.1f;
  float torque = dr1 * P.constraint * rad / max_speed / P.max_radius / dt;
  torque = fminf(torque, P.constraint);
  if (dr1 > 0) {
    if (torque / rad > fr) dr = max_speed * P.max_radius / P.constraint * (torque / rad - fr) * dt;
  } else {
    if (CC) {
      if (-P.constraint_contraction * dr1 > fa * rad)
        dr = (P.constraint_contraction * dr1 + fa * rad) / (P.constraint_contraction);
      dr = fmaxf(dr, -P.max_radius * dt);
    } else {
      dr = dr1;
    }
  }
  dr = rad + dr;
  if (dr > P.max_radius) dr = P.max_radius;
  if (dr < P.min_radius) dr = P.min_radius;
  return dr;
}

__global__ void __launch_bounds__(256)
k_update_rad(const float *__restrict__ fa, const float *__restrict__ fr, float *__restrict__ rad,
             const float *__restrict__ phase, float time, float dt, const int *__restrict__ dead, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (dead[i]) return;
  const float r = rad[i];
  float rn;
  if (c_prm.p.constrained_contraction) rn = controller_one<true>(r, phase[i], fr[i], fa[i], time, dt);
  else rn = controller_one<false>(r, phase[i], fr[i], 0.0f, time, dt);
  if (rn != r) rad[i] = rn;
}

/* Fused K1: controller -> integrate -> (hash) with one read and one write of each robot
 * (north_star (4); SURVEY.md §8d K1: 60 B per robot on sort steps).  COUNT: the robot also takes
 * its arrival ticket in its cell (cell binning, prs_cellbin.cuh) — index[i] then holds the ticket. */
template <bool DO_SORT, bool COUNT = false>
__global__ void __launch_bounds__(256)
k_control_integrate_hash(float2 *__restrict__ pos, float2 *__restrict__ vel, float *__restrict__ rad,
                         const float *__restrict__ phase, const float *__restrict__ fa, const float *__restrict__ fr,
                         const int *__restrict__ dead, uint32_t *__restrict__ hash, uint32_t *__restrict__ index,
                         float time, float dt, int run_controller, uint32_t n, const uint32_t *__restrict__ n_dev,
                         uint32_t *__restrict__ cellCount = nullptr, uint32_t *__restrict__ tileMark = nullptr,
                         uint32_t row_lo = 0u, uint32_t row_hi = 0xffffffffu, uint32_t log2_gx = 0u,
                         uint32_t *range = nullptr, uint32_t range_tile0 = 0u) {
  prs::pdl_sync();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = *n_dev; /* slab ranks keep their robot count on the device */
  if (i >= n) return;
  float2 p = pos[i];
  float2 v = vel[i];
  float r = rad[i];
  if (run_controller) {
    /* phase and the |force| sums are requested together with the dead flag, not after it has arrived: the
     * kernel is one chain of memory round trips at 2^20 robots (reading them for a dead robot is harmless) */
    const int is_dead = dead[i];
    const float ph = phase[i], f_r = fr[i];
    const float f_a = c_prm.p.constrained_contraction ? fa[i] : 0.0f;
    if (!is_dead) {
      float rn;
      if (c_prm.p.constrained_contraction) rn = controller_one<true>(r, ph, f_r, f_a, time, dt);
      else rn = controller_one<false>(r, ph, f_r, 0.0f, time, dt);
      if (rn != r) { rad[i] = rn; r = rn; }
    }
  }
  const float2 v0 = v;
  integrate_one(p, v, r, dt);
  pos[i] = p;
  if (v.x != v0.x || v.y != v0.y) vel[i] = v;
  if (DO_SORT) {
    const int2 g = cell_of(p.x, p.y);
    const uint32_t h = cell_hash(g.x, g.y);
    hash[i] = h;
    if (COUNT) {
      /* slab ranks: a robot whose new row left [row_lo, row_hi) migrates and takes its ticket where it arrives */
      const uint32_t row = h >> log2_gx;
      const bool mine = row >= row_lo && row < row_hi;
      index[i] = mine ? atomicAdd(&cellCount[h], 1u) : 0xffffffffu;
      if (range && mine) prs_bin::range_check(range, h / prs_bin::SCAN_TILE - range_tile0);
    } else {
      index[i] = i;
    }
    if (COUNT && tileMark) {
      /* the scan skips tiles nobody marked: the first lane of a run of equal tiles stores, into the way of its block */
      const uint32_t tile = h / prs_bin::SCAN_TILE;
      const unsigned act = __activemask();
      const uint32_t left = __shfl_up_sync(act, tile, 1);
      if ((threadIdx.x & 31u) == 0 || left != tile) tileMark[tile * prs_bin::MARK_WAYS + (blockIdx.x % prs_bin::MARK_WAYS)] = 1u;
    }
  }
}

/* K1 of the fused binned step with TWO robots per thread and vector accesses (north_star (1): 128-bit coalesced loads and
 * stores of positions and velocities, 64-bit ones of radius / phase / |force| / dead flag / key / ticket): the kernel is a
 * chain of memory round trips at 2^20 robots, and a thread that keeps two robots' loads in flight walks that chain half as
 * often.  Per-robot arithmetic is controller_one / integrate_one / cell_of as above — same bits.  Needs 16-byte aligned
 * pos / vel and 8-byte aligned scalar arrays (the launcher checks); an odd last robot is handled by the scalar tail. */
__global__ void __launch_bounds__(256)
k_control_integrate_hash_x2(float4 *__restrict__ pos, float4 *__restrict__ vel, float2 *__restrict__ rad,
                            const float2 *__restrict__ phase, const float2 *__restrict__ fa, const float2 *__restrict__ fr,
                            const int2 *__restrict__ dead, uint2 *__restrict__ hash, uint2 *__restrict__ ticket, float time, float dt,
                            int run_controller, uint32_t n, uint32_t *__restrict__ cellCount, uint32_t *__restrict__ tileMark,
                            const uint32_t *__restrict__ n_dev, uint32_t row_lo, uint32_t row_hi, uint32_t log2_gx,
                            uint32_t *range, uint32_t range_tile0) {
  prs::pdl_sync();
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t i0 = 2u * t;
  if (n_dev) n = *n_dev; /* slab ranks keep their robot count on the device */
  if (i0 >= n) return;
  /* slab ranks: a robot whose new row left [row_lo, row_hi) migrates and takes its ticket where it arrives; a ticket outside
   * the scan's tile range of this step widens the scan to every tile (prs_bin::range_check) */
  auto take = [&](uint32_t h) {
    const uint32_t row = h >> log2_gx;
    if (!(row >= row_lo && row < row_hi)) return 0xffffffffu;
    if (range) prs_bin::range_check(range, h / prs_bin::SCAN_TILE - range_tile0);
    return atomicAdd(&cellCount[h], 1u);
  };
  const bool cc = c_prm.p.constrained_contraction != 0;
  if (i0 + 1u >= n) { /* odd tail: one robot, scalar accesses */
    float2 p = reinterpret_cast<float2 *>(pos)[i0], v = reinterpret_cast<float2 *>(vel)[i0];
    float r = reinterpret_cast<float *>(rad)[i0];
    if (run_controller && !reinterpret_cast<const int *>(dead)[i0]) {
      const float ph = reinterpret_cast<const float *>(phase)[i0], f_r = reinterpret_cast<const float *>(fr)[i0];
      const float rn = cc ? controller_one<true>(r, ph, f_r, reinterpret_cast<const float *>(fa)[i0], time, dt)
                          : controller_one<false>(r, ph, f_r, 0.0f, time, dt);
      if (rn != r) { reinterpret_cast<float *>(rad)[i0] = rn; r = rn; }
    }
    const float2 v0 = v;
    integrate_one(p, v, r, dt);
    reinterpret_cast<float2 *>(pos)[i0] = p;
    if (v.x != v0.x || v.y != v0.y) reinterpret_cast<float2 *>(vel)[i0] = v;
    const int2 g = cell_of(p.x, p.y);
    const uint32_t h = cell_hash(g.x, g.y);
    reinterpret_cast<uint32_t *>(hash)[i0] = h;
    reinterpret_cast<uint32_t *>(ticket)[i0] = take(h);
    if (tileMark) tileMark[(h / prs_bin::SCAN_TILE) * prs_bin::MARK_WAYS + (blockIdx.x % prs_bin::MARK_WAYS)] = 1u;
    return;
  }
  const float4 P = pos[t], V = vel[t];
  float2 R = rad[t];
  float2 p0 = make_float2(P.x, P.y), p1 = make_float2(P.z, P.w), v0 = make_float2(V.x, V.y), v1 = make_float2(V.z, V.w);
  if (run_controller) {
    const int2 D = dead[t];
    const float2 PH = phase[t], FR = fr[t];
    const float2 FA = cc ? fa[t] : make_float2(0.0f, 0.0f);
    const float2 R_in = R;
    if (!D.x) R.x = cc ? controller_one<true>(R.x, PH.x, FR.x, FA.x, time, dt) : controller_one<false>(R.x, PH.x, FR.x, 0.0f, time, dt);
    if (!D.y) R.y = cc ? controller_one<true>(R.y, PH.y, FR.y, FA.y, time, dt) : controller_one<false>(R.y, PH.y, FR.y, 0.0f, time, dt);
    if (R.x != R_in.x || R.y != R_in.y) rad[t] = R;
  }
  integrate_one(p0, v0, R.x, dt);
  integrate_one(p1, v1, R.y, dt);
  pos[t] = make_float4(p0.x, p0.y, p1.x, p1.y);
  if (v0.x != V.x || v0.y != V.y || v1.x != V.z || v1.y != V.w) vel[t] = make_float4(v0.x, v0.y, v1.x, v1.y);
  const int2 g0 = cell_of(p0.x, p0.y), g1 = cell_of(p1.x, p1.y);
  const uint32_t h0 = cell_hash(g0.x, g0.y), h1 = cell_hash(g1.x, g1.y);
  hash[t] = make_uint2(h0, h1);
  const uint32_t k0 = take(h0), k1 = take(h1);
  ticket[t] = make_uint2(k0, k1);
  if (tileMark) {
    const uint32_t tile0 = h0 / prs_bin::SCAN_TILE, tile1 = h1 / prs_bin::SCAN_TILE;
    const unsigned act = __activemask();
    const uint32_t left = __shfl_up_sync(act, tile1, 1);
    const uint32_t way = blockIdx.x % prs_bin::MARK_WAYS;
    if ((threadIdx.x & 31u) == 0 || left != tile0) tileMark[tile0 * prs_bin::MARK_WAYS + way] = 1u;
    if (tile1 != tile0) tileMark[tile1 * prs_bin::MARK_WAYS + way] = 1u;
  }
}

/* Steps WITHOUT a sort, small swarms: K1 and the gather in ONE launch.  Thread k owns sorted slot k, runs
 * controller + integrate for the robot in that slot (every robot sits in exactly one slot; same arithmetic
 * as k_control_integrate_hash) and writes both its state and its packed sorted record.  One launch less per
 * step, which is what a 300-robot step is made of. */
__global__ void __launch_bounds__(256)
k_control_integrate_gather(float2 *__restrict__ pos, float2 *__restrict__ vel, float *__restrict__ rad,
                           const float *__restrict__ phase, const float *__restrict__ fa, const float *__restrict__ fr,
                           const int *__restrict__ dead, const uint32_t *__restrict__ index, float4 *__restrict__ sortedPR,
                           float2 *__restrict__ sortedVel, float time, float dt, int run_controller, uint32_t n) {
  prs::pdl_sync();
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t i = index[k];
  float2 p = pos[i];
  float2 v = vel[i];
  float r = rad[i];
  if (run_controller) {
    /* phase and the |force| sums are requested together with the dead flag, not after it has arrived: the
     * kernel is one chain of memory round trips at 2^20 robots (reading them for a dead robot is harmless) */
    const int is_dead = dead[i];
    const float ph = phase[i], f_r = fr[i];
    const float f_a = c_prm.p.constrained_contraction ? fa[i] : 0.0f;
    if (!is_dead) {
      float rn;
      if (c_prm.p.constrained_contraction) rn = controller_one<true>(r, ph, f_r, f_a, time, dt);
      else rn = controller_one<false>(r, ph, f_r, 0.0f, time, dt);
      if (rn != r) { rad[i] = rn; r = rn; }
    }
  }
  const float2 v0 = v;
  integrate_one(p, v, r, dt);
  pos[i] = p;
  if (v.x != v0.x || v.y != v0.y) vel[i] = v;
  sortedPR[k] = make_float4(p.x, p.y, r, __uint_as_float(i));
  sortedVel[k] = v;
}

/* largest cell population of a sorted key array (guard of the binned route, see prs_fused_step) */
__global__ void __launch_bounds__(256)
k_max_population(const uint32_t *__restrict__ hash, const uint32_t *__restrict__ cellStart, const uint32_t *__restrict__ cellEnd,
                 uint32_t n, uint32_t *__restrict__ out_max) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t m = 0;
  if (k < n) {
    const uint32_t h = hash[k];
    if (k == 0 || hash[k - 1] != h) m = cellEnd[h] - cellStart[h];
  }
  __shared__ uint32_t s_m[8];
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < 8; w++) m = max(m, s_m[w]);
    /* one atomic per block at most, none once the running maximum is reached */
    if (m > *reinterpret_cast<volatile uint32_t *>(out_max)) atomicMax(out_max, m);
  }
}

/* ---- light shadow tests (results of checkIntersectionLine/Circle/checkIntersection, :184-262) ---- */
__device__ int seg_hit(float x0, float y0, float x1, float y1, float x3, float y3, float x4, float y4) {
  if (fabsf((x4 - x3) / (x1 - x0)) == fabsf((y4 - y3) / (y1 - y0))) return 0;
  float t, t1;
  if (fabsf(y4 - y3) > 0) {
    t = (x3 - x0 - (y3 - y0) * (x3 - x4) / (y3 - y4)) * ((y3 - y4) / ((x1 - x0) * (y3 - y4) - (y1 - y0) * (x3 - x4)));
    if (t <= 0 || t >= 1) return 0;
    t1 = (y3 - y0 - t * (y1 - y0)) / (y3 - y4);
    if (t1 <= 0 || t1 >= 1) return 0;
  } else if (fabsf(x4 - x3) > 0) {
    t = (y3 - y0 - (x3 - x0) * (y3 - y4) / (x3 - x4)) * ((x3 - x4) / ((y1 - y0) * (x3 - x4) - (x1 - x0) * (y3 - y4)));
    if (t <= 0 || t >= 1) return 0;
    t1 = (x3 - x0 - t * (x1 - x0)) / (x3 - x4);
    if (t1 <= 0 || t1 >= 1) return 0;
  } else {
    return 0;
  }
  return 1;
}
__device__ int disc_hit(float lx, float ly, float px, float py, float ox, float oy, float orad) {
  const float C1 = powf(lx, 2) + powf(ly, 2), C2 = powf(px, 2) + powf(py, 2), C3 = powf(ox, 2) + powf(oy, 2);
  const float C4 = lx * px + ly * py, C5 = lx * ox + ly * oy, C6 = px * ox + py * oy;
  const float A = C1 + C2 - 2 * C4;
  const float B = -2 * C1 + 2 * C4 + 2 * C5 - 2 * C6;
  const float C = C1 + C3 - 2 * C5 - powf(orad, 2);
  const float D = powf(B, 2) - 4 * A * C;
  if (D >= 0) {
    const float R1 = (-B + powf(D, 0.5f)) / 2 / A, R2 = (-B - powf(D, 0.5f)) / 2 / A;
    if (R1 > 0 && R1 < 1) return 1;
    if (R2 > 0 && R2 < 1) return 1;
  }
  return 0;
}
__device__ int in_shadow(float px, float py) {
  const float lx = c_prm.p.light_x, ly = c_prm.p.light_y;
  for (int i = 0; i < c_prm.p.n_cir_obstacles; i++)
    if (disc_hit(lx, ly, px, py, c_prm.x_cir[i], c_prm.y_cir[i], c_prm.r_cir[i])) return 1;
  for (int i = 0; i < c_prm.p.nobstacles; i++) {
    const float x1 = c_prm.x1obs[i], x2 = c_prm.x2obs[i], y1 = c_prm.y1obs[i], y2 = c_prm.y2obs[i];
    if (seg_hit(lx, ly, px, py, x1, y1, x1, y2)) return 1;
    if (seg_hit(lx, ly, px, py, x1, y2, x2, y2)) return 1;
    if (seg_hit(lx, ly, px, py, x2, y2, x2, y1)) return 1;
    if (seg_hit(lx, ly, px, py, x2, y1, x1, y1)) return 1;
  }
  return 0;
}

/* light-distance phase offsets (result of updatePhase, :264-290); min_d from a scalar or from
 * device memory (d_min_d != nullptr) */
__global__ void __launch_bounds__(256) k_update_phase(const float2 *__restrict__ pos, float *__restrict__ phase,
                                                      float spacing, float min_d_host, const float *__restrict__ d_min_d,
                                                      uint32_t n, const uint32_t *__restrict__ n_dev) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = *n_dev;
  if (i >= n) return;
  const float min_d = d_min_d ? d_min_d[0] : min_d_host;
  const float2 p = pos[i];
  const float dist = norm2(mk(p.x, p.y) - mk(c_prm.p.light_x, c_prm.p.light_y));
  int visible = 1;
  if (c_prm.p.light_shadow) {
    if (in_shadow(p.x, p.y)) visible = 0;
  }
  if (!visible) {
    if (c_prm.p.light_shadow == 1) phase[i] = -(c_prm.p.Nx - 1) * c_prm.p.rise_period;
    if (c_prm.p.light_shadow == 2) phase[i] = 9999999999.0f;
  } else {
    phase[i] = (min_d - dist) / (spacing)*c_prm.p.rise_period;
  }
}

/* min_i |light - p_i| on the device.  The reference does this on the host with glibc powf
 * (particlebot.cpp:214-228); the _rn intrinsics below evaluate the same correctly rounded
 * square / sum / square root (no FMA), and positive floats order like their bit patterns. */
__global__ void __launch_bounds__(256) k_min_light_distance(const float2 *__restrict__ pos, uint32_t n,
                                                            uint32_t *__restrict__ out_bits, const uint32_t *__restrict__ n_dev) {
  if (n_dev) n = *n_dev;
  float m = __int_as_float(0x7f7f7f7f);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float2 p = pos[i];
    const float dx = c_prm.p.light_x - p.x, dy = c_prm.p.light_y - p.y;
    const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    m = fminf(m, d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMin(out_bits, (uint32_t)__float_as_int(m));
}

/* XORWOW per robot, the toolkit's own device API so the stream is the reference's bit for bit
 * (results of curand_setup_kernel / add_normal_noise_kernel, :36-51) */
__global__ void __launch_bounds__(256) k_curand_setup(curandState *__restrict__ st, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) curand_init(c_prm.p.seed, i, 0, &st[i]);
}
__global__ void __launch_bounds__(256) k_add_normal_noise(curandState *__restrict__ st, float *__restrict__ val,
                                                          float std, uint32_t n, const uint32_t *__restrict__ n_dev) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = *n_dev;
  if (i >= n) return;
  const float noise = std * curand_normal(st + i);
  val[i] += noise;
}

/* one level of the reference's 64-wide centroid tree (calcCOG/calcCOG1, :295-349): block b sums
 * in[b*64 .. b*64+63] pairwise (+32, +16, ... +1); the last level scales and tags y with +2000 */
__global__ void __launch_bounds__(64) k_cog_level(const float2 *__restrict__ in, float2 *__restrict__ out, int n,
                                                  int last, float mul) {
  __shared__ float2 s[64];
  const int tid = threadIdx.x;
  const int i = blockIdx.x * 64 + tid;
  float2 a = make_float2(0.0f, 0.0f);
  if (i < n) { const float2 t = in[i]; a.x += t.x; a.y += t.y; }
  s[tid] = a;
  __syncthreads();
  for (int o = 32; o > 0; o >>= 1) {
    if (tid < o) {
      const float2 b = s[tid + o];
      s[tid].x += b.x;
      s[tid].y += b.y;
    }
    __syncthreads();
  }
  if (tid == 0) {
    float2 r = s[0];
    if (last) { r.x *= mul; r.y *= mul; r.y = r.y + 2000.0f; }
    out[blockIdx.x] = r;
  }
}

/* swarm centroid as an observable: double accumulation, fixed two-stage shape (deterministic) */
__global__ void __launch_bounds__(256) k_centroid_partial(const float2 *__restrict__ pos, uint32_t n, double2 *__restrict__ part) {
  __shared__ double sx[256], sy[256];
  double ax = 0, ay = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float2 p = pos[i];
    ax += p.x; ay += p.y;
  }
  sx[threadIdx.x] = ax; sy[threadIdx.x] = ay;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) { sx[threadIdx.x] += sx[threadIdx.x + o]; sy[threadIdx.x] += sy[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = make_double2(sx[0], sy[0]);
}
__global__ void k_centroid_final(const double2 *__restrict__ part, int nb, uint32_t n, float *__restrict__ out) {
  double ax = 0, ay = 0;
  for (int b = 0; b < nb; b++) { ax += part[b].x; ay += part[b].y; }
  out[0] = (float)(ax / (double)n);
  out[1] = (float)(ay / (double)n);
}

/* Synthetic hex block ON THE DEVICE: robot i of an nx-wide hex lattice, every coordinate
 * jittered by a counter hash of (seed, i) — the arithmetic of Particlebot::initHexBlock's host loop operation for operation
 * (plain IEEE mul / add, no contraction: identical bits), plus the state reset() gives a fresh swarm (zero velocity, minimum
 * radius, zero phase, alive).  A 2^26-robot swarm is placed in a millisecond instead of a host loop and 1.3 GB of uploads. */
__global__ void __launch_bounds__(256)
k_init_hex_block(float2 *__restrict__ pos, float2 *__restrict__ vel, float *__restrict__ rad, float *__restrict__ phase,
                 int *__restrict__ dead, unsigned long long n, uint32_t nx, float pitch, float row, float x0, float y0, float jitter,
                 uint32_t seed, float min_radius) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t iy = (uint32_t)(i / nx), ix = (uint32_t)(i % nx);
  unsigned long long z = ((unsigned long long)seed << 32) ^ i;
  z += 0x9e3779b97f4a7c15ull; z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; z = z ^ (z >> 31);
  const float jx = __fmul_rn(__fsub_rn(__fdiv_rn((float)(uint32_t)(z & 0xffffffull), 8388608.0f), 1.0f), jitter);
  const float jy = __fmul_rn(__fsub_rn(__fdiv_rn((float)(uint32_t)((z >> 24) & 0xffffffull), 8388608.0f), 1.0f), jitter);
  const float odd = (iy & 1u) ? __fmul_rn(0.5f, pitch) : 0.0f;
  const float x = __fadd_rn(__fadd_rn(__fadd_rn(x0, __fmul_rn((float)ix, pitch)), odd), jx);
  const float y = __fadd_rn(__fadd_rn(y0, __fmul_rn((float)iy, row)), jy);
  pos[i] = make_float2(x, y);
  vel[i] = make_float2(0.0f, 0.0f);
  rad[i] = min_radius;
  phase[i] = 0.0f;
  dead[i] = 0;
}

/* radius -> colour for the optional renderer (result of updateCol_k, :401-443; the shadow
 * darkening halves the HSL lightness) */
__device__ float hue_channel(float p, float q, float t) {
  if (t < 0) t += 1;
  if (t > 1) t -= 1;
  if (t < 1.0 / 6.0) return p + (q - p) * 6.0 * t;
  if (t < 1.0 / 2.0) return q;
  if (t < 2.0 / 3.0) return p + (q - p) * (2.0 / 3.0 - t) * 6.0;
  return p;
}
__global__ void __launch_bounds__(256) k_update_col(const float *__restrict__ rad, float4 *__restrict__ col,
                                                    const float2 *__restrict__ pos, const int *__restrict__ dead, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float r = rad[i];
  float4 cc = col[i];
  if (dead[i]) {
    cc.x = 0.0f; cc.y = 0.0f; cc.z = 0.0f;
  } else {
    const float lo = c_prm.p.min_radius, hi = c_prm.p.max_radius;
    cc.x = 30.0f / 255.0f;
    cc.y = (20.0f + (200.0f - 20.0f) * powf(hi - r, 2.0f) / powf(hi - lo, 2.0f)) / 255.0f;
    cc.z = (30.0f + (210.0f - 30.0f) * powf(r - lo, 0.5f) / powf(hi - lo, 0.5f)) / 255.0f;
    if (c_prm.p.display_shadow) {
      const float2 p = pos[i];
      if (in_shadow(p.x, p.y)) {
        const float mx = fmaxf(fmaxf(cc.x, cc.y), cc.z), mn = fminf(fminf(cc.x, cc.y), cc.z);
        float h, s, l = (mx + mn) / 2;
        if (mx == mn) {
          h = s = 0;
        } else {
          const float d = mx - mn;
          s = l > 0.5 ? d / (2.0 - mx - mn) : d / (mx + mn);
          if (mx == cc.x) h = (cc.y - cc.z) / d + (cc.y < cc.z ? 6.0 : 0.0);
          else if (mx == cc.y) h = (cc.z - cc.x) / d + 2.0;
          else h = (cc.x - cc.y) / d + 4.0;
          h /= 6.0;
        }
        l = l / 2.0;
        if (s == 0) {
          cc.x = cc.y = cc.z = l;
        } else {
          const float q = l < 0.5 ? l * (1.0 + s) : l + s - l * s;
          const float pp = 2.0 * l - q;
          cc.x = hue_channel(pp, q, h + 1.0 / 3.0);
          cc.y = hue_channel(pp, q, h);
          cc.z = hue_channel(pp, q, h - 1.0 / 3.0);
        }
      }
    }
  }
  col[i] = cc;
}

/* ------------------------------------------------------------------------------------------
 * sort driver
 * ------------------------------------------------------------------------------------------ */
static void ensure_sort_workspace(uint32_t n, int npass, int tile_pairs) {
  prs_sort::Workspace &w = g_prs.sort_ws;
  if (w.cap_pairs < n) {
    for (int b = 0; b < 2; b++) {
      if (w.keys[b]) PRS_CUDA(cudaFree(w.keys[b]));
      if (w.vals[b]) PRS_CUDA(cudaFree(w.vals[b]));
      PRS_CUDA(cudaMalloc(&w.keys[b], (size_t)n * 4));
      PRS_CUDA(cudaMalloc(&w.vals[b], (size_t)n * 4));
    }
    w.cap_pairs = n;
  }
  const size_t need = prs_sort::meta_words(n, npass, tile_pairs);
  if (w.cap_meta < need) {
    if (w.meta) PRS_CUDA(cudaFree(w.meta));
    PRS_CUDA(cudaMalloc(&w.meta, need * 4));
    w.cap_meta = need;
  }
}

template <int NT, int BITS, int IT>
static void launch_onesweep(unsigned tiles, const uint32_t *src_k, const uint32_t *src_v, uint32_t *dst_k, uint32_t *dst_v,
                            uint32_t n, int shift, const uint32_t *ghist, uint32_t *status, uint32_t *counter,
                            unsigned long long *timeline, const uint32_t *n_dev) {
  using S = prs_sort::Smem<NT, BITS, IT>;
  static bool smem_opt_in = false;
  if (!smem_opt_in) {
    PRS_CUDA(cudaFuncSetAttribute(prs_sort::k_onesweep<NT, BITS, IT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(S)));
    smem_opt_in = true;
  }
  PRS_LAUNCH((prs_sort::k_onesweep<NT, BITS, IT>), tiles, NT, sizeof(S), src_k, src_v, dst_k, dst_v, n, shift, ghist, status, counter,
             timeline, n_dev);
}

/* stable sort of n pairs by the low key_bits of the key; result in out_* (may alias in_*).
 * vals_are_iota: in_vals[i] == i is known (skips reading them in the first pass). */
static void sort_pairs(const uint32_t *in_k, const uint32_t *in_v, uint32_t *out_k, uint32_t *out_v, uint32_t n,
                       int key_bits, bool vals_are_iota, const uint32_t *n_dev = nullptr) {
  /* n_dev != nullptr: the pair count lives on the device (slab ranks), n is its upper bound —
   * grids and scratch are sized for n, tiles past the real count exit at once */
  using namespace prs_sort;
  if (n == 0) return;
  if (key_bits < 1) key_bits = 1;
  if (key_bits > 32) key_bits = 32;
  /* digits of this sort: 8 bits, or 9 where that saves a pass; 8 or 12 pairs per thread (prs_onesweep.cuh) */
  const PassPlan plan = plan_passes(key_bits, n);
  const int npass = plan.npass;
  /* tile shape: 1024 threads while all tiles fit one wave (one per SM), else 512 with two tiles resident per SM;
   * prs_sort_set_threads() pins one of them.  12 pairs per thread only exist for 512 threads. */
  int nt = g_prs.sort_threads ? g_prs.sort_threads : ((n <= 148u * 8192u) ? 1024 : 512);
  const int items = (nt == 512) ? plan.items : 8;
  const int tile_pairs = nt * items;
  g_prs.sort_tile_pairs = (unsigned)tile_pairs;
  ensure_sort_workspace(n, npass, tile_pairs);
  Workspace &w = g_prs.sort_ws;
  const uint32_t tiles = div_up(n, tile_pairs);
  uint32_t *ghist = w.meta;
  uint32_t *counters = w.meta + MAX_PASSES * MAX_RADIX;
  uint32_t *status = counters + MAX_PASSES;
  PRS_CUDA(cudaMemsetAsync(w.meta, 0, meta_words(n, npass, tile_pairs) * 4, g_prs.stream));
  const unsigned hist_blocks = min(div_up(n, HIST_THREADS * 8), 148u * 8u);
  PRS_LAUNCH(k_histogram, hist_blocks, HIST_THREADS, 0, in_k, n, ghist, plan, n_dev);
  const uint32_t *src_k = in_k, *src_v = vals_are_iota ? nullptr : in_v;
  /* a single pass (key_bits <= 8) would scatter into the buffer its other tiles still read when out_* aliases
   * in_*: it then lands in scratch and is copied back */
  const bool bounce = npass == 1 && (out_k == in_k || (in_v && out_v == in_v));
  for (int p = 0; p < npass; p++) {
    /* ping-pong through the two scratch pairs so that the LAST pass lands in out_* */
    uint32_t *dst_k, *dst_v;
    if (p == npass - 1 && !bounce) { dst_k = out_k; dst_v = out_v; }
    else { dst_k = w.keys[p & 1]; dst_v = w.vals[p & 1]; }
    unsigned long long *tl = g_prs.sort_timeline ? g_prs.sort_timeline + (size_t)p * tiles * 8 : nullptr;
    const uint32_t *gh = ghist + p * MAX_RADIX;
    uint32_t *st = status + (size_t)p * tiles * MAX_RADIX;
    const int shift = plan.shift[p];
    const bool wide = plan.bits[p] == 9;
    if (nt == 512) {
      if (wide) launch_onesweep<512, 9, 8>(tiles, src_k, src_v, dst_k, dst_v, n, shift, gh, st, counters + p, tl, n_dev);
      else if (items == 12) launch_onesweep<512, 8, 12>(tiles, src_k, src_v, dst_k, dst_v, n, shift, gh, st, counters + p, tl, n_dev);
      else launch_onesweep<512, 8, 8>(tiles, src_k, src_v, dst_k, dst_v, n, shift, gh, st, counters + p, tl, n_dev);
    } else {
      if (wide) launch_onesweep<1024, 9, 8>(tiles, src_k, src_v, dst_k, dst_v, n, shift, gh, st, counters + p, tl, n_dev);
      else launch_onesweep<1024, 8, 8>(tiles, src_k, src_v, dst_k, dst_v, n, shift, gh, st, counters + p, tl, n_dev);
    }
    src_k = dst_k;
    src_v = dst_v;
  }
  if (bounce) {
    PRS_CUDA(cudaMemcpyAsync(out_k, w.keys[0], (size_t)n * 4, cudaMemcpyDeviceToDevice, g_prs.stream));
    PRS_CUDA(cudaMemcpyAsync(out_v, w.vals[0], (size_t)n * 4, cudaMemcpyDeviceToDevice, g_prs.stream));
  }
}

static int key_bits_of_grid() {
  /* cell keys are < numCells (power of two); before setParameters assume full 32-bit keys */
  if (!g_prs.params_set || g_prs.h_prm.p.numCells == 0) return 32;
  int b = 0;
  while ((1ull << b) < (unsigned long long)g_prs.h_prm.p.numCells) b++;
  return b < 1 ? 1 : b;
}

/* ------------------------------------------------------------------------------------------
 * C-ABI part 1: the reference's entry points
 * ------------------------------------------------------------------------------------------ */
extern "C" {

void cudaInit(int argc, char **argv) {
  /* -device=N picks the device (helper_cuda.h:1246-1283); otherwise the current device stays */
  int dev = -1;
  for (int i = 1; i < argc; i++) {
    const char *a = argv[i];
    while (*a == '-') a++;
    if (strncmp(a, "device=", 7) == 0) dev = atoi(a + 7);
  }
  int count = 0;
  PRS_CUDA(cudaGetDeviceCount(&count));
  if (count == 0) { printf("No CUDA Capable devices found, exiting...\n"); exit(EXIT_SUCCESS); }
  if (dev >= 0) PRS_CUDA(cudaSetDevice(dev));
}
void cudaGLInit(int argc, char **argv) { cudaInit(argc, argv); }

void allocateArray(void **devPtr, size_t size) {
  /* zero-filled: the reference's host class reads absForce_a / absForce_r (first controller step, `0 * absForce_r[i]` in
   * the first collide) and its grid arrays before anything wrote them and relies on a fresh process handing out zeroed
   * device memory; a long-lived process does not, so the assumption is made true here */
  PRS_CUDA(cudaMalloc(devPtr, size));
  if (size) PRS_CUDA(cudaMemsetAsync(*devPtr, 0, size, g_prs.stream));
}
void freeArray(void *devPtr) {
  g_prs.table.cellStart = nullptr; /* a recycled address must not look like the cached table */
  g_prs.bin.marks_table = nullptr;
  PRS_CUDA(cudaFree(devPtr));
}
void threadSync(void) { PRS_CUDA(cudaDeviceSynchronize()); }

void copyArrayToDevice(void *device, const void *host, int offset, int size) {
  /* state may change behind the fused step's back: the binned route re-earns its admission */
  g_prs.bin.admitted = false;
  g_prs.bin.generation++;
  PRS_CUDA(cudaMemcpyAsync((char *)device + offset, host, (size_t)size, cudaMemcpyHostToDevice, g_prs.stream));
  PRS_CUDA(cudaStreamSynchronize(g_prs.stream));
}
[[maybe_unused]] static void no_gl(const char *fn) {
  fprintf(stderr, "%s: libparticlebot_b200 is built headless (no OpenGL interop); rendering is optional\n", fn);
  exit(EXIT_FAILURE);
}
void copyArrayFromDevice(void *host, const void *device, struct cudaGraphicsResource **res, int size) {
#ifdef PRS_WITH_GL
  if (res) device = mapGLBufferObject(res); /* particlebot_cuda.cu:97-100: the source is the mapped buffer object */
#else
  if (res) no_gl("copyArrayFromDevice(mapped VBO)");
#endif
  PRS_CUDA(cudaMemcpyAsync(host, device, (size_t)size, cudaMemcpyDeviceToHost, g_prs.stream));
  PRS_CUDA(cudaStreamSynchronize(g_prs.stream));
#ifdef PRS_WITH_GL
  if (res) unmapGLBufferObject(*res);
#endif
}
#ifdef PRS_WITH_GL
/* OpenGL build (-DPRS_WITH_GL): buffer objects of a GL host — the reference's own main.cpp / render.cpp /
 * particlebot.cpp — are registered with and mapped through the CUDA graphics API, what particlebot_cuda.cu:69-93 does, on
 * this library's stream.  A GL context must be current on the calling thread (cudaGLInit, main.cpp:348). */
void registerGLBufferObject(unsigned vbo, struct cudaGraphicsResource **res) {
  PRS_CUDA(cudaGraphicsGLRegisterBuffer(res, vbo, cudaGraphicsMapFlagsNone));
}
void unregisterGLBufferObject(struct cudaGraphicsResource *res) { PRS_CUDA(cudaGraphicsUnregisterResource(res)); }
void *mapGLBufferObject(struct cudaGraphicsResource **res) {
  void *ptr = nullptr;
  size_t bytes = 0;
  PRS_CUDA(cudaGraphicsMapResources(1, res, g_prs.stream));
  PRS_CUDA(cudaGraphicsResourceGetMappedPointer(&ptr, &bytes, *res));
  return ptr;
}
void unmapGLBufferObject(struct cudaGraphicsResource *res) { PRS_CUDA(cudaGraphicsUnmapResources(1, &res, g_prs.stream)); }
#else
void registerGLBufferObject(unsigned, struct cudaGraphicsResource **) { no_gl("registerGLBufferObject"); }
void unregisterGLBufferObject(struct cudaGraphicsResource *) { no_gl("unregisterGLBufferObject"); }
void *mapGLBufferObject(struct cudaGraphicsResource **) { no_gl("mapGLBufferObject"); return nullptr; }
void unmapGLBufferObject(struct cudaGraphicsResource *) { no_gl("unmapGLBufferObject"); }
#endif

void setParameters(SimParams *hp) {
  PrsDevParams &d = g_prs.h_prm;
  memset(&d, 0, sizeof(d));
  d.p = *hp;
  if (hp->nobstacles > PRS_MAX_OBSTACLES || hp->n_cir_obstacles > PRS_MAX_OBSTACLES || hp->nobstacles < 0 ||
      hp->n_cir_obstacles < 0) {
    /* the reference overruns its 10-entry constant arrays here; refuse instead */
    fprintf(stderr, "setParameters: more than %d obstacles of one kind\n", PRS_MAX_OBSTACLES);
    exit(EXIT_FAILURE);
  }
  for (int i = 0; i < hp->nobstacles; i++) {
    d.x1obs[i] = hp->x1obs[i]; d.x2obs[i] = hp->x2obs[i]; d.y1obs[i] = hp->y1obs[i]; d.y2obs[i] = hp->y2obs[i];
  }
  for (int i = 0; i < hp->n_cir_obstacles; i++) {
    d.x_cir[i] = hp->x_cir_obs[i]; d.y_cir[i] = hp->y_cir_obs[i]; d.r_cir[i] = hp->r_cir_obs[i];
  }
  d.world_half = g_prs.world_half;
  g_prs.params_set = true;
  PRS_CUDA(cudaMemcpyToSymbolAsync(c_prm, &d, sizeof(d), 0, cudaMemcpyHostToDevice, g_prs.stream));
  PRS_LAUNCH(prs::k_collide_constants, 1, 1, 0); /* per-parameter-set constants of collide (prs_collide.cuh) */
  PRS_CUDA(cudaStreamSynchronize(g_prs.stream));
}

unsigned iDivUp(unsigned a, unsigned b) { return (a % b != 0) ? (a / b + 1) : (a / b); }
void computeGridSize(unsigned n, unsigned blockSize, unsigned &numBlocks, unsigned &numThreads) {
  numThreads = blockSize < n ? blockSize : n;
  numBlocks = iDivUp(n, numThreads);
}
void computeGridSize2(unsigned n, unsigned blockSize, unsigned &numBlocks, unsigned &numThreads) {
  numThreads = blockSize;
  numBlocks = iDivUp(n, numThreads);
}

void integrateSystem(float *pos, float *vel, float *rad, float deltaTime, unsigned nCells, float /*time*/) {
  if (!nCells) return;
  PRS_LAUNCH(k_integrate, div_up(nCells, 256), 256, 0, (float2 *)pos, (float2 *)vel, rad, deltaTime, nCells);
}

void calcHash(unsigned *hash, unsigned *index, float *pos, int nCells) {
  g_prs.table.cellStart = nullptr;
  if (nCells <= 0) return;
  PRS_LAUNCH(k_calc_hash, div_up(nCells, 256), 256, 0, hash, index, (const float2 *)pos, (uint32_t)nCells);
}

void sortParticlebots(unsigned *hash, unsigned *index, unsigned nCells) {
  g_prs.table.cellStart = nullptr; /* the fused step's cached table no longer describes these keys */
  sort_pairs(hash, index, hash, index, nCells, key_bits_of_grid(), false);
}

void reorderDataAndFindCellStart(unsigned *cellStart, unsigned *cellEnd, float *sortedPos, float *sortedVel,
                                 float *sortedRad, unsigned *hash, unsigned *index, float *oldPos, float *oldVel,
                                 float *oldRad, unsigned nCells, unsigned numCells) {
  g_prs.table.cellStart = nullptr;
  g_prs.bin.marks_table = nullptr;
  PRS_CUDA(cudaMemsetAsync(cellStart, 0xff, (size_t)numCells * sizeof(unsigned), g_prs.stream));
  if (!nCells) return;
  PRS_LAUNCH(k_reorder, div_up(nCells, 256), 256, 0, cellStart, cellEnd, (float2 *)sortedPos, (float2 *)sortedVel,
             sortedRad, hash, index, (const float2 *)oldPos, (const float2 *)oldVel, oldRad, nCells);
}

void collide(float *newVel, float *absForce_a, float *absForce_r, float *sortedPos, float *sortedVel,
             float *sortedRad, unsigned *index, unsigned *cellStart, unsigned *cellEnd, unsigned nCells,
             unsigned /*numCells*/, float deltaTime) {
  if (!nCells) return;
  prs_launch_collide((float2 *)newVel, absForce_a, absForce_r, (const float2 *)sortedPos, (const float2 *)sortedVel,
                     sortedRad, index, cellStart, cellEnd, nCells, deltaTime);
}

void updateRad_light_wave(float * /*pos*/, float *absForce_a, float *absForce_r, float *rad, float *phase, float time,
                          float deltaTime, int *dead, int nCells) {
  if (nCells <= 0) return;
  PRS_LAUNCH(k_update_rad, div_up(nCells, 256), 256, 0, absForce_a, absForce_r, rad, phase, time, deltaTime, dead,
             (uint32_t)nCells);
}

void updatePhase(float *pos, float *phase, float spacing, float /*max_d*/, float min_d, int nCells) {
  if (nCells <= 0) return;
  PRS_LAUNCH(k_update_phase, div_up(nCells, 256), 256, 0, (const float2 *)pos, phase, spacing, min_d,
             (const float *)nullptr, (uint32_t)nCells, (const uint32_t *)nullptr);
}

void curand_setup(struct curandStateXORWOW *state, int N) {
  if (N <= 0) return;
  PRS_LAUNCH(k_curand_setup, div_up(N, 256), 256, 0, (curandState *)state, (uint32_t)N);
}
void add_normal_noise(struct curandStateXORWOW *state, float *val, float std, int N) {
  if (N <= 0) return;
  PRS_LAUNCH(k_add_normal_noise, div_up(N, 256), 256, 0, (curandState *)state, val, std, (uint32_t)N, (const uint32_t *)nullptr);
}

void calcCOG(float *pos, float *temppos, float *temppos1, int nCells, float time, int hist_steps, float hist_int) {
  if (nCells <= 0) return;
  const int ind = ((int)(time / hist_int)) % hist_steps;
  const float mul = 1.0f / float(nCells);
  /* levels of 64: in -> a -> b -> a ... ; the last level writes the scaled, tagged centroid */
  const float2 *in = (const float2 *)pos;
  float2 *bufs[2] = {(float2 *)temppos, (float2 *)temppos1};
  int n = nCells, which = 0;
  while (true) {
    const int nb = (n + 63) / 64;
    const int last = (nb == 1);
    float2 *out = last ? (float2 *)temppos1 : bufs[which];
    if (last && in == (const float2 *)temppos1) {
      /* never read and write temppos1 in the same level */
      PRS_CUDA(cudaMemcpyAsync(temppos, temppos1, (size_t)n * sizeof(float2), cudaMemcpyDeviceToDevice, g_prs.stream));
      in = (const float2 *)temppos;
    }
    PRS_LAUNCH(k_cog_level, nb, 64, 0, in, out, n, last, mul);
    if (last) break;
    in = out;
    which ^= 1;
    n = nb;
  }
  PRS_CUDA(cudaMemcpyAsync(pos + 2 * (size_t)(ind + nCells), temppos1, 2 * sizeof(float), cudaMemcpyDeviceToDevice,
                           g_prs.stream));
}

void updateCol(float *rad, float *col, int nCells, float *pos, float * /*phase*/, int *dead) {
  if (nCells <= 0) return;
  PRS_LAUNCH(k_update_col, div_up(nCells, 256), 256, 0, rad, (float4 *)col, (const float2 *)pos, dead, (uint32_t)nCells);
}

/* ------------------------------------------------------------------------------------------
 * C-ABI part 2: additions
 * ------------------------------------------------------------------------------------------ */
const char *prs_version(void) { return "particlebot-b200 0.1 (sm_100a)"; }
void prs_set_stream(void *s) { g_prs.stream = (cudaStream_t)s; }
void *prs_get_stream(void) { return (void *)g_prs.stream; }

void prs_set_world_half_extent(float half) {
  g_prs.world_half = half;
  g_prs.h_prm.world_half = half;
  PRS_CUDA(cudaMemcpyToSymbolAsync(c_prm, &half, sizeof(float), offsetof(PrsDevParams, world_half),
                                   cudaMemcpyHostToDevice, g_prs.stream));
  PRS_CUDA(cudaStreamSynchronize(g_prs.stream));
}
float prs_get_world_half_extent(void) { return g_prs.world_half; }
/* swarms of up to max_robots use the warp-per-robot collide kernel (0 = never); default 16384 */
void prs_set_collide_warp_max(unsigned max_robots) { g_prs.collide_warp_max = max_robots; }
void prs_set_collide_tile(int on) { g_prs.collide_tile = on ? 1 : 0; }
void prs_set_patch_rows(unsigned rows) { g_prs.patch.rows = rows < 1 ? 1 : (rows > (unsigned)prs::PATCH_HMAX ? (unsigned)prs::PATCH_HMAX : rows); }
/* tuning aid: on != 0 starts counting {patches taken by the patch kernel, patches handed to the per-robot slow lane,
 * patches redone because a pair left the admitted operand ranges}; returns the counts so far in out[3] and clears them */
void prs_patch_stats(int on, unsigned *out) {
  PrsPatchState &T = g_prs.patch;
  if (!T.stats_buf) {
    PRS_CUDA(cudaMalloc(&T.stats_buf, 4 * 4));
    PRS_CUDA(cudaMemsetAsync(T.stats_buf, 0, 4 * 4, g_prs.stream));
  }
  if (out) {
    PRS_CUDA(cudaMemcpyAsync(out, T.stats_buf, 3 * 4, cudaMemcpyDeviceToHost, g_prs.stream));
    PRS_CUDA(cudaStreamSynchronize(g_prs.stream));
  }
  PRS_CUDA(cudaMemsetAsync(T.stats_buf, 0, 4 * 4, g_prs.stream));
  T.stats = on ? T.stats_buf : nullptr;
}
void prs_set_pdl(int on) { g_prs.pdl = on ? 1 : 0; }
/* 1 (default): K1 of the fused binned step handles two robots per thread with 128 / 64-bit accesses; 0: one robot per thread */
void prs_set_k1_x2(int on) { g_prs.k1_x2 = on ? 1 : 0; }
/* 1 (default): on binned sort steps of plain swarms collide finds its stencil rows in the dense start table the scan writes (10 table
 * words per robot instead of 30); 0: always through cellStart / cellEnd */
void prs_set_collide_dense(int on) { g_prs.collide_dense = on ? 1 : 0; }
void prs_set_slab_scan_range(int on) { g_prs.slab_scan_range = on ? 1 : 0; g_prs.bin.range_table = nullptr; }

/* ---- host-buffer step (Particlebot::updateHost): asynchronous copies around prs_fused_step ----
 * prs_h2d_async: pinned host -> device on the launching stream (counts as an upload: the binned route re-earns
 * its admission).  prs_arm_k1_event(1): the next fused step records an event once K1 (controller + integrate)
 * is launched — positions and radii are final from there on.  prs_d2h_async(.., after_k1 = 1): device -> host on
 * a second stream that waits for that event only, so the copy runs under the sort and collide kernels;
 * after_k1 = 0: on the launching stream (after everything).  prs_host_step_sync waits for both streams. */
void prs_h2d_async(void *device, const void *host, size_t bytes) {
  g_prs.bin.admitted = false;
  g_prs.bin.generation++;
  PRS_CUDA(cudaMemcpyAsync(device, host, bytes, cudaMemcpyHostToDevice, g_prs.stream));
}
void prs_arm_k1_event(int on) {
  if (on && !g_prs.k1_event) {
    PRS_CUDA(cudaEventCreateWithFlags(&g_prs.k1_event, cudaEventDisableTiming));
    PRS_CUDA(cudaStreamCreateWithFlags(&g_prs.copy_stream, cudaStreamNonBlocking));
  }
  g_prs.k1_event_armed = on != 0;
}
void prs_d2h_async(void *host, const void *device, size_t bytes, int after_k1) {
  if (after_k1 && g_prs.k1_event) {
    PRS_CUDA(cudaStreamWaitEvent(g_prs.copy_stream, g_prs.k1_event, 0));
    PRS_CUDA(cudaMemcpyAsync(host, device, bytes, cudaMemcpyDeviceToHost, g_prs.copy_stream));
  } else {
    PRS_CUDA(cudaMemcpyAsync(host, device, bytes, cudaMemcpyDeviceToHost, g_prs.stream));
  }
}
void prs_host_step_plan(const float *pos_in, const float *vel_in, const float *rad_in, float *pos_out, float *rad_out) {
  if (!g_prs.copy_stream) PRS_CUDA(cudaStreamCreateWithFlags(&g_prs.copy_stream, cudaStreamNonBlocking));
  g_prs.plan.active = true;
  g_prs.plan.pos_in = pos_in; g_prs.plan.vel_in = vel_in; g_prs.plan.rad_in = rad_in;
  g_prs.plan.pos_out = pos_out; g_prs.plan.rad_out = rad_out;
}
void prs_set_plan_chunks(unsigned chunks) { g_prs.plan_chunks = chunks; }
void prs_host_step_sync(void) {
  g_prs.plan.active = false; /* a plan whose step never reached prs_fused_step must not outlive the host-buffer step */
  if (g_prs.copy_stream) PRS_CUDA(cudaStreamSynchronize(g_prs.copy_stream));
  PRS_CUDA(cudaStreamSynchronize(g_prs.stream));
}
void prs_set_fuse_gather_max(unsigned max_robots) { g_prs.fuse_gather_max = max_robots; }
int prs_get_pdl(void) { return g_prs.pdl; }
int prs_get_collide_tile(void) { return g_prs.collide_tile; }
unsigned long long prs_launch_count(int reset) {
  const unsigned long long v = g_prs.launches;
  if (reset) g_prs.launches = 0;
  return v;
}

void prs_min_light_distance(const float *pos, int n, float *d_min_d) {
  PRS_CUDA(cudaMemsetAsync(d_min_d, 0x7f, sizeof(float), g_prs.stream));
  if (n <= 0) return;
  const unsigned blocks = min(div_up((unsigned)n, 256 * 4), 148u * 8u);
  PRS_LAUNCH(k_min_light_distance, blocks, 256, 0, (const float2 *)pos, (uint32_t)n, (uint32_t *)d_min_d, (const uint32_t *)nullptr);
}
void prs_update_phase_dev(const float *pos, float *phase, float spacing, const float *d_min_d, int n) {
  if (n <= 0) return;
  PRS_LAUNCH(k_update_phase, div_up(n, 256), 256, 0, (const float2 *)pos, phase, spacing, 0.0f, d_min_d, (uint32_t)n, (const uint32_t *)nullptr);
}
void prs_centroid(const float *pos, int n, float *d_scratch, float *d_out) {
  if (n <= 0) return;
  const int nb = (int)min(div_up((unsigned)n, 256 * 8), 256u);
  PRS_LAUNCH(k_centroid_partial, nb, 256, 0, (const float2 *)pos, (uint32_t)n, (double2 *)d_scratch);
  PRS_LAUNCH(k_centroid_final, 1, 1, 0, (const double2 *)d_scratch, nb, (uint32_t)n, d_out);
}

/* tuning aid: when set, the next sorts write 8 %globaltimer stamps per tile and pass into buf
 * (device memory, passes * tiles * 8 words); nullptr switches it off */
void prs_sort_set_timeline(unsigned long long *buf) { g_prs.sort_timeline = buf; }
int prs_sort_plan(int key_bits, unsigned n, int *out) {
  if (key_bits < 1) key_bits = 1;
  if (key_bits > 32) key_bits = 32;
  const prs_sort::PassPlan plan = prs_sort::plan_passes(key_bits, n);
  for (int p = 0; p < 4; p++) out[p] = p < plan.npass ? plan.bits[p] : 0;
  out[4] = plan.items;
  out[5] = 0;
  return plan.npass;
}
unsigned prs_sort_tile_size(void) { return g_prs.sort_tile_pairs ? g_prs.sort_tile_pairs : 4096u; } /* pairs per tile of the last sort */
/* tile shape of the sort: 512 (several tiles per SM) or 1024 threads x 8 pairs */
void prs_sort_set_threads(int nt) { g_prs.sort_threads = (nt == 1024) ? 1024 : (nt == 512 ? 512 : 0); }

void prs_sort_pairs(const unsigned *in_keys, const unsigned *in_vals, unsigned *out_keys, unsigned *out_vals,
                    unsigned n, int key_bits) {
  sort_pairs(in_keys, in_vals, out_keys, out_vals, n, key_bits, false);
}

/* ---- guard of the binned route ----
 * The in-cell ranking of prs_cellbin.cuh is quadratic in the cell population, so the route is only
 * taken while the swarm is known to be sparse enough.  The largest population of every sort step
 * (binned or not) is copied to pinned host memory asynchronously and looked at when it has
 * arrived — the host never waits for it; anything that rewrites positions behind the library's
 * back (setArray, reset, ...) calls prs_bin_invalidate().  Both routes give identical results. */
static void bin_poll_report() {
  PrsBinState &B = g_prs.bin;
  if (!B.report_pending || cudaEventQuery(B.report_event) != cudaSuccess) return;
  B.report_pending = false;
  if (B.h_report[1]) {
    /* a cell outgrew the in-cell ranking between two reports (in practice: a swarm that blew up to NaN,
     * every NaN position hashes to cell 0).  Those cells were filed in arrival order for the steps in
     * between; from here on the onesweep route is taken, as the pop > 64 test below would also decide. */
    static bool warned = false;
    if (!warned) fprintf(stderr, "prs_fused_step: a cell held more than %u robots; cell binning switched off\n", prs_bin::MAX_RANKED_CELL);
    warned = true;
    B.admitted = false;
  }
  const uint32_t pop = B.h_report[0];
  if (B.admitted) { if (pop > 64u) B.admitted = false; }
  else if (B.report_generation == B.generation && pop <= 48u) B.admitted = true;
}
static void bin_send_report(const uint32_t *d_two_words) {
  PrsBinState &B = g_prs.bin;
  if (B.report_pending) return; /* one report in flight at a time */
  if (!B.h_report) {
    PRS_CUDA(cudaMallocHost(&B.h_report, 2 * sizeof(uint32_t)));
    PRS_CUDA(cudaEventCreateWithFlags(&B.report_event, cudaEventDisableTiming));
  }
  PRS_CUDA(cudaMemcpyAsync(B.h_report, d_two_words, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, g_prs.stream));
  PRS_CUDA(cudaEventRecord(B.report_event, g_prs.stream));
  B.report_pending = true;
  B.report_generation = B.generation;
}
static void bin_ensure(uint32_t n, uint32_t C) {
  PrsBinState &B = g_prs.bin;
  if (B.cap_cells < C) {
    if (B.cellCount) PRS_CUDA(cudaFree(B.cellCount));
    if (B.scratch) PRS_CUDA(cudaFree(B.scratch));
    PRS_CUDA(cudaMalloc(&B.cellCount, (size_t)C * 4));
    PRS_CUDA(cudaMemsetAsync(B.cellCount, 0, (size_t)C * 4, g_prs.stream));
    PRS_CUDA(cudaMalloc(&B.scratch, prs_bin::scan_scratch_words(C) * 4));
    PRS_CUDA(cudaMemsetAsync(B.scratch, 0, prs_bin::scan_scratch_words(C) * 4, g_prs.stream));
    if (B.marks) PRS_CUDA(cudaFree(B.marks));
    if (!B.range) {
      PRS_CUDA(cudaMalloc(&B.range, 4 * 4));
      PRS_CUDA(cudaMemsetAsync(B.range, 0, 4 * 4, g_prs.stream));
    }
    B.range_table = nullptr;
    const size_t tiles = (C + prs_bin::SCAN_TILE - 1) / prs_bin::SCAN_TILE;
    /* [0, tiles * MARK_WAYS) marks of this step, then one word per tile for the previous step */
    PRS_CUDA(cudaMalloc(&B.marks, (tiles * prs_bin::MARK_WAYS + tiles) * 4));
    PRS_CUDA(cudaMemsetAsync(B.marks, 0, (tiles * prs_bin::MARK_WAYS + tiles) * 4, g_prs.stream));
    B.marks_table = nullptr;
    /* dense start table (collide's fresh-table steps) + per-tile liveness */
    if (B.dense) { PRS_CUDA(cudaFree(B.dense)); PRS_CUDA(cudaFree(B.live)); }
    PRS_CUDA(cudaMalloc(&B.dense, ((size_t)C + 4) * 4));
    PRS_CUDA(cudaMalloc(&B.live, tiles * 4));
    PRS_CUDA(cudaMemsetAsync(B.dense, 0, ((size_t)C + 4) * 4, g_prs.stream));
    PRS_CUDA(cudaMemsetAsync(B.live, 0, tiles * 4, g_prs.stream));
    B.cap_cells = C;
  }
  ensure_sort_workspace(n, 1, 4096);
}
void prs_bin_invalidate(void) {
  g_prs.bin.marks_table = nullptr;
  g_prs.bin.admitted = false;
  g_prs.bin.generation++;
}
void prs_bin_set_mode(int mode) { g_prs.bin.mode = mode; } /* 0 auto (default), 1 never (always onesweep), 2 always */
int prs_bin_active(void) { return g_prs.bin.admitted ? 1 : 0; }

/* ---- work list + launch of k_collide_patch (prs_collide_patch.cuh) ---- */
static bool patch_eligible(uint32_t n, bool need_fa, bool fresh_table) {
  const SimParams &p = g_prs.h_prm.p;
  const unsigned gx = p.gridSize.x, gy = p.gridSize.y;
  return g_prs.collide_tile && fresh_table && p.nDead != -1 && !need_fa && n > g_prs.collide_warp_max && gx >= 32 && gy >= 16 &&
         (gx & (gx - 1)) == 0 && (gy & (gy - 1)) == 0 && gx * gy == p.numCells;
}
/* new epoch: the table kernels of this step append the patches that hold robots */
static prs_bin::PatchListArgs patch_begin_step() {
  PrsPatchState &T = g_prs.patch;
  const SimParams &p = g_prs.h_prm.p;
  const size_t need = (size_t)p.numCells / prs::PATCH_W; /* patches of one row at most */
  if (T.cap_cells < need) {
    if (T.epoch_of) { PRS_CUDA(cudaFree(T.epoch_of)); PRS_CUDA(cudaFree(T.list)); }
    PRS_CUDA(cudaMalloc(&T.epoch_of, need * 4));
    PRS_CUDA(cudaMalloc(&T.list, need * 4));
    PRS_CUDA(cudaMemsetAsync(T.epoch_of, 0, need * 4, g_prs.stream));
    T.cap_cells = need;
  }
  if (!T.count) {
    PRS_CUDA(cudaMalloc(&T.count, 2 * 4));
    PRS_CUDA(cudaMemsetAsync(T.count, 0, 2 * 4, g_prs.stream));
  }
  T.epoch++;
  if (T.epoch == 0) { /* wrapped: forget every stamp */
    PRS_CUDA(cudaMemsetAsync(T.epoch_of, 0, T.cap_cells * 4, g_prs.stream));
    T.epoch = 1;
  }
  prs_bin::PatchListArgs a;
  a.epoch_of = T.epoch_of; a.list = T.list; a.count = T.count + (T.epoch & 1u); a.epoch = T.epoch;
  a.PH = T.rows;
  while ((1u << a.log2_gx) < p.gridSize.x) a.log2_gx++;
  return a;
}
static void launch_collide_patch(float2 *newVel, float *fr, const float4 *pr, const float2 *svel, const uint32_t *cellStart,
                                 const uint32_t *cellEnd, float dt) {
  PrsPatchState &T = g_prs.patch;
  if (!T.smem_opt_in) {
    PRS_CUDA(cudaFuncSetAttribute(prs::k_collide_patch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(prs::PatchSmem)));
    PRS_CUDA(cudaFuncSetAttribute(prs::k_collide_patch, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int dev = 0;
    PRS_CUDA(cudaGetDevice(&dev));
    PRS_CUDA(cudaDeviceGetAttribute(&T.num_sms, cudaDevAttrMultiProcessorCount, dev));
    T.smem_opt_in = true;
  }
  const SimParams &p = g_prs.h_prm.p;
  const unsigned long long patches = (unsigned long long)(p.gridSize.x / prs::PATCH_W) * ((p.gridSize.y + T.rows - 1) / T.rows);
  const unsigned grid = (unsigned)std::min<unsigned long long>(patches, 2ull * (unsigned)T.num_sms);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(prs::PATCH_NT);
  cfg.dynamicSmemBytes = sizeof(prs::PatchSmem);
  cfg.stream = g_prs.stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = g_prs.pdl ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  const uint32_t *count = T.count + (T.epoch & 1u);
  uint32_t *next = T.count + ((T.epoch + 1u) & 1u);
  const cudaError_t e = cudaLaunchKernelEx(&cfg, prs::k_collide_patch, newVel, fr, pr, svel, cellStart, cellEnd,
                                           (const uint32_t *)T.list, count, next, (uint32_t)T.rows, dt, T.stats);
  g_prs.launches++;
  if (e != cudaSuccess) prs_fail("k_collide_patch", e, __FILE__, __LINE__);
}

static void plan_download(const prs_step_buffers *b, uint32_t first, uint32_t count, size_t chunk);
static const prs_step_buffers *g_plan_after_k1 = nullptr; /* planned host-buffer step on a route without the pipeline */
static inline void k1_done() {
  if (g_prs.k1_event_armed) PRS_CUDA(cudaEventRecord(g_prs.k1_event, g_prs.stream));
  if (g_plan_after_k1) {
    plan_download(g_plan_after_k1, 0u, g_plan_after_k1->nCells, 0);
    g_plan_after_k1 = nullptr;
  }
}

/* host-buffer step (prs_host_step_plan) on a route that is not pipelined: everything up first ... */
static void plan_upload_all(const prs_step_buffers *b, uint32_t n) {
  const PrsHostState::HostPlan &H = g_prs.plan;
  PRS_CUDA(cudaMemcpyAsync(b->pos, H.pos_in, (size_t)n * 8, cudaMemcpyHostToDevice, g_prs.stream));
  PRS_CUDA(cudaMemcpyAsync(b->vel, H.vel_in, (size_t)n * 8, cudaMemcpyHostToDevice, g_prs.stream));
  PRS_CUDA(cudaMemcpyAsync(b->rad, H.rad_in, (size_t)n * 4, cudaMemcpyHostToDevice, g_prs.stream));
}
/* ... and positions and radii of robots [first, first + count) back on the second stream once K1 has written them */
static void plan_download(const prs_step_buffers *b, uint32_t first, uint32_t count, size_t chunk) {
  PrsHostState &G = g_prs;
  while (G.chunk_events.size() <= chunk) {
    cudaEvent_t e;
    PRS_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    G.chunk_events.push_back(e);
  }
  PRS_CUDA(cudaEventRecord(G.chunk_events[chunk], G.stream));
  PRS_CUDA(cudaStreamWaitEvent(G.copy_stream, G.chunk_events[chunk], 0));
  PRS_CUDA(cudaMemcpyAsync(G.plan.pos_out + 2 * (size_t)first, b->pos + 2 * (size_t)first, (size_t)count * 8, cudaMemcpyDeviceToHost, G.copy_stream));
  PRS_CUDA(cudaMemcpyAsync(G.plan.rad_out + first, b->rad + first, (size_t)count * 4, cudaMemcpyDeviceToHost, G.copy_stream));
}

void prs_fused_step(const prs_step_buffers *b, float time, float dt, int do_sort) {
  const uint32_t n = b->nCells;
  if (!n) { g_prs.plan.active = false; return; }
  /* the plan is this step's: whatever route the step takes, it is consumed here */
  struct PlanScope { ~PlanScope() { g_prs.plan.active = false; g_plan_after_k1 = nullptr; } } plan_scope;
  const bool planned = g_prs.plan.active;
  const int run_controller = (g_prs.h_prm.p.control == LIGHT_WAVE && time >= 0) ? 1 : 0;
  const bool need_fa = g_prs.h_prm.p.constrained_contraction != 0;
  PrsBinState &B = g_prs.bin;
  bool binned = false;
  if (do_sort && b->sortedPR) {
    bin_poll_report();
    const bool shape_ok = (unsigned long long)b->numCells <= 16ull * n && n < (1u << 30);
    binned = shape_ok && (B.mode == 2 || (B.mode == 0 && B.admitted));
  }
  const bool patch = patch_eligible(n, need_fa, do_sort != 0) && b->sortedPR;
  const prs_bin::PatchListArgs pl = patch ? patch_begin_step() : prs_bin::PatchListArgs();
  const bool pipelined = planned && binned && n >= (1u << 18);
  if (planned && !pipelined) {
    plan_upload_all(b, n);
    g_plan_after_k1 = b;
  }
  if (binned) {
    /* K1 + tickets -> scan (= cell table) -> scatter -> in-cell order + gather */
    bin_ensure(n, b->numCells);
    prs_sort::Workspace &w = g_prs.sort_ws;
    uint32_t *ticket = w.vals[0], *hash_by_slot = w.keys[0], *index_by_slot = w.vals[1];
    const unsigned tiles = div_up(b->numCells, prs_bin::SCAN_TILE);
    uint32_t *marks = B.marks, *prev_marks = B.marks + (size_t)tiles * prs_bin::MARK_WAYS;
    if (B.marks_table != (const void *)b->cellStart || B.marks_cells != b->numCells || B.marks_generation != B.generation) {
      /* this table was last written by somebody else (other route, other buffers, an upload): every tile
       * counts as "held robots before", i.e. gets its empty markers rewritten */
      PRS_CUDA(cudaMemsetAsync(prev_marks, 0xff, (size_t)tiles * 4, g_prs.stream));
      B.marks_table = b->cellStart; B.marks_cells = b->numCells; B.marks_generation = B.generation;
    }
    {
      StageScope t(PRS_STAGE_K1);
      PRS_CUDA(cudaMemsetAsync(B.scratch, 0, 16, g_prs.stream));
      const uintptr_t al = (uintptr_t)b->pos | (uintptr_t)b->vel | (((uintptr_t)b->rad | (uintptr_t)b->phase | (uintptr_t)b->absForce_a |
                            (uintptr_t)b->absForce_r | (uintptr_t)b->dead | (uintptr_t)b->hash | (uintptr_t)ticket) << 1);
      const bool x2 = g_prs.k1_x2 && (al & 15u) == 0 && n >= 65536u;
      /* robots [first, first + count): the whole swarm, or one chunk of the pipelined host-buffer step (multiples of 1024
       * robots: the vector accesses of the x2 kernel stay aligned) */
      auto k1_range = [&](uint32_t first, uint32_t count) {
        if (x2) {
          PRS_LAUNCH_PDL(k_control_integrate_hash_x2, div_up(div_up(count, 2), 256), 256, (float4 *)(b->pos + 2 * (size_t)first),
                         (float4 *)(b->vel + 2 * (size_t)first), (float2 *)(b->rad + first), (const float2 *)(b->phase + first),
                         (const float2 *)(b->absForce_a + first), (const float2 *)(b->absForce_r + first), (const int2 *)(b->dead + first),
                         (uint2 *)(b->hash + first), (uint2 *)(ticket + first), time, dt, run_controller, count, B.cellCount, marks,
                         (const uint32_t *)nullptr, 0u, 0xffffffffu, 0u, (uint32_t *)nullptr, 0u);
        } else {
          PRS_LAUNCH_PDL((k_control_integrate_hash<true, true>), div_up(count, 256), 256, (float2 *)(b->pos + 2 * (size_t)first),
                         (float2 *)(b->vel + 2 * (size_t)first), b->rad + first, b->phase + first, b->absForce_a + first,
                         b->absForce_r + first, b->dead + first, b->hash + first, ticket + first, time, dt, run_controller, count,
                         (const uint32_t *)nullptr, B.cellCount, marks, 0u, 0xffffffffu, 0u, (uint32_t *)nullptr, 0u);
        }
      };
      if (pipelined) {
        /* three queues: the uploads run back to back on their own stream (a copy queue that alternates with kernels pays a
         * hand-over per alternation), K1 of chunk c waits for its three uploads, the way back of chunk c waits for K1 */
        const PrsHostState::HostPlan &H = g_prs.plan;
        PrsHostState &G = g_prs;
        const uint32_t nchunks = G.plan_chunks ? G.plan_chunks : 2u; /* measured at 2^20 robots: 1: 0.816, 2: 0.799, 4: 0.867, 8: 0.98 ms */
        const uint32_t chunk = std::max<uint32_t>(1u << 16, ((n / nchunks) + 1023u) & ~1023u);
        if (!G.upload_stream) PRS_CUDA(cudaStreamCreateWithFlags(&G.upload_stream, cudaStreamNonBlocking));
        if (!G.step_event) PRS_CUDA(cudaEventCreateWithFlags(&G.step_event, cudaEventDisableTiming));
        /* the uploads overwrite what the previous step's kernels (and copies) still read: start after them */
        PRS_CUDA(cudaEventRecord(G.step_event, G.stream));
        PRS_CUDA(cudaStreamWaitEvent(G.upload_stream, G.step_event, 0));
        size_t c = 0;
        for (uint32_t first = 0; first < n; first += chunk, c++) {
          const uint32_t count = std::min<uint32_t>(chunk, n - first);
          PRS_CUDA(cudaMemcpyAsync(b->pos + 2 * (size_t)first, H.pos_in + 2 * (size_t)first, (size_t)count * 8, cudaMemcpyHostToDevice, G.upload_stream));
          PRS_CUDA(cudaMemcpyAsync(b->vel + 2 * (size_t)first, H.vel_in + 2 * (size_t)first, (size_t)count * 8, cudaMemcpyHostToDevice, G.upload_stream));
          PRS_CUDA(cudaMemcpyAsync(b->rad + first, H.rad_in + first, (size_t)count * 4, cudaMemcpyHostToDevice, G.upload_stream));
          while (G.upload_events.size() <= c) {
            cudaEvent_t e;
            PRS_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            G.upload_events.push_back(e);
          }
          PRS_CUDA(cudaEventRecord(G.upload_events[c], G.upload_stream));
          PRS_CUDA(cudaStreamWaitEvent(G.stream, G.upload_events[c], 0));
          k1_range(first, count);
          plan_download(b, first, count, c);
        }
      } else {
        k1_range(0u, n);
      }
      k1_done();
    }
    /* dense start table for collide (plain swarms, thread-per-robot kernel): the scan writes it for the tiles collide can reach */
    const bool use_dense = g_prs.collide_dense && !patch && !need_fa && g_prs.h_prm.p.nDead != -1 && n > g_prs.collide_warp_max;
    prs_bin::DenseArgs dn;
    if (use_dense) {
      dn.dense = B.dense; dn.live = B.live;
      dn.dil = (2u * g_prs.h_prm.p.gridSize.x + 3u + prs_bin::SCAN_TILE - 1u) / prs_bin::SCAN_TILE;
    }
    {
      StageScope t(PRS_STAGE_SORT);
      PRS_LAUNCH_PDL(prs_bin::k_cell_tile_sums, tiles, prs_bin::SCAN_THREADS, B.cellCount, b->numCells, B.scratch, (const uint32_t *)marks, dn, prs_bin::RangeArgs());
      if (tiles <= prs_bin::SELF_PREFIX_MAX_TILES) {
        PRS_LAUNCH_PDL(prs_bin::k_cell_apply<true>, tiles, prs_bin::SCAN_THREADS, B.cellCount, b->cellStart, b->cellEnd, b->numCells,
                       B.scratch, 0u, marks, prev_marks, pl, dn, prs_bin::RangeArgs());
      } else {
        PRS_LAUNCH_PDL(prs_bin::k_cell_scan_tiles, 1, 1024, B.scratch, tiles);
        PRS_LAUNCH_PDL(prs_bin::k_cell_apply<false>, tiles, prs_bin::SCAN_THREADS, B.cellCount, b->cellStart, b->cellEnd, b->numCells,
                       B.scratch, 0u, marks, prev_marks, pl, dn, prs_bin::RangeArgs());
      }
      PRS_LAUNCH_PDL(prs_bin::k_cell_scatter, div_up(n, 256), 256, b->hash, ticket, b->cellStart, hash_by_slot, index_by_slot, n);
    }
    {
      StageScope t(PRS_STAGE_REORDER);
      PRS_LAUNCH_PDL(prs_bin::k_reorder_binned, div_up(n, 256), 256, hash_by_slot, index_by_slot, b->cellStart, b->cellEnd,
                     b->hash, b->index, (float4 *)b->sortedPR, (float2 *)b->sortedVel, (const float2 *)b->pos,
                     (const float2 *)b->vel, b->rad, n, B.scratch);
    }
    g_prs.table.cellStart = b->cellStart; g_prs.table.hash = b->hash; g_prs.table.n = n;
    g_prs.table.numCells = b->numCells; g_prs.table.generation = B.generation;
    {
      StageScope t(PRS_STAGE_COLLIDE);
      prs::PackedLayout in{(const float4 *)b->sortedPR, (const float2 *)b->sortedVel};
      if (patch) launch_collide_patch((float2 *)b->vel, b->absForce_r, in.pr, in.vel, b->cellStart, b->cellEnd, dt);
      else prs_launch_collide_t((float2 *)b->vel, b->absForce_a, b->absForce_r, in, b->cellStart, b->cellEnd, n, dt, need_fa, 0u,
                                (const uint32_t *)nullptr, 0, (const uint32_t *)nullptr, use_dense ? (const uint32_t *)B.dense : nullptr,
                                use_dense ? (const uint32_t *)b->hash : nullptr);
    }
    /* fullest cell of this sort -> pinned host memory (nobody touches the two words before the next step's
     * memset); issued after collide so that the kernels of the step stay adjacent in the stream */
    bin_send_report(B.scratch + 1);
    return;
  }
  if (!do_sort && b->sortedPR && n <= g_prs.fuse_gather_max) {
    const PrsTableState &T = g_prs.table;
    if (T.cellStart == b->cellStart && T.hash == b->hash && T.n == n && T.numCells == b->numCells && T.generation == B.generation) {
      /* the table of the last sort is current (nothing was rewritten since): K1 + gather in one launch, then collide */
      {
        StageScope t(PRS_STAGE_K1);
        PRS_LAUNCH_PDL(k_control_integrate_gather, div_up(n, 256), 256, (float2 *)b->pos, (float2 *)b->vel, b->rad, b->phase,
                       b->absForce_a, b->absForce_r, b->dead, b->index, (float4 *)b->sortedPR, (float2 *)b->sortedVel, time, dt,
                       run_controller, n);
        k1_done();
      }
      StageScope t(PRS_STAGE_COLLIDE);
      prs::PackedLayout in{(const float4 *)b->sortedPR, (const float2 *)b->sortedVel};
      prs_launch_collide_t((float2 *)b->vel, b->absForce_a, b->absForce_r, in, b->cellStart, b->cellEnd, n, dt, need_fa);
      return;
    }
  }
  if (do_sort) {
    {
      StageScope t(PRS_STAGE_K1);
      PRS_LAUNCH(k_control_integrate_hash<true>, div_up(n, 256), 256, 0, (float2 *)b->pos, (float2 *)b->vel, b->rad,
                 b->phase, b->absForce_a, b->absForce_r, b->dead, b->hash, b->index, time, dt, run_controller, n, (const uint32_t *)nullptr);
      k1_done();
    }
    StageScope t(PRS_STAGE_SORT);
    sort_pairs(b->hash, b->index, b->hash, b->index, n, key_bits_of_grid(), true);
  } else {
    StageScope t(PRS_STAGE_K1);
    PRS_LAUNCH_PDL((k_control_integrate_hash<false, false>), div_up(n, 256), 256, (float2 *)b->pos, (float2 *)b->vel, b->rad,
                   b->phase, b->absForce_a, b->absForce_r, b->dead, b->hash, b->index, time, dt, run_controller, n,
                   (const uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, 0u, 0xffffffffu, 0u, (uint32_t *)nullptr, 0u);
    k1_done();
  }
  if (b->sortedPR) {
    bool report = false;
    {
      StageScope t(PRS_STAGE_REORDER);
      PrsTableState &T = g_prs.table;
      const bool table_current = !do_sort && T.cellStart == b->cellStart && T.hash == b->hash && T.n == n &&
                                 T.numCells == b->numCells && T.generation == B.generation;
      if (table_current) {
        PRS_LAUNCH_PDL(k_gather_packed, div_up(n, 256), 256, (float4 *)b->sortedPR, (float2 *)b->sortedVel, b->index,
                       (const float2 *)b->pos, (const float2 *)b->vel, b->rad, n);
      } else {
        B.marks_table = nullptr; /* the table is rebuilt from scratch here: the binned route's tile marks no longer describe it */
        PRS_CUDA(cudaMemsetAsync(b->cellStart, 0xff, (size_t)b->numCells * sizeof(unsigned), g_prs.stream));
        PRS_LAUNCH(k_reorder_packed, div_up(n, 256), 256, 0, b->cellStart, b->cellEnd, (float4 *)b->sortedPR,
                   (float2 *)b->sortedVel, b->hash, b->index, (const float2 *)b->pos, (const float2 *)b->vel, b->rad, n, pl);
        T.cellStart = b->cellStart; T.hash = b->hash; T.n = n; T.numCells = b->numCells; T.generation = B.generation;
      }
      if (do_sort && B.mode == 0 && !B.admitted && (unsigned long long)b->numCells <= 16ull * n) {
        /* not on the binned route: report the largest cell population so that it can be taken */
        bin_ensure(n, b->numCells);
        PRS_CUDA(cudaMemsetAsync(B.scratch, 0, 16, g_prs.stream));
        PRS_LAUNCH(k_max_population, div_up(n, 256), 256, 0, b->hash, b->cellStart, b->cellEnd, n, B.scratch + 1);
        report = true;
      }
    }
    {
      StageScope t(PRS_STAGE_COLLIDE);
      prs::PackedLayout in{(const float4 *)b->sortedPR, (const float2 *)b->sortedVel};
      if (patch) launch_collide_patch((float2 *)b->vel, b->absForce_r, in.pr, in.vel, b->cellStart, b->cellEnd, dt);
      else prs_launch_collide_t((float2 *)b->vel, b->absForce_a, b->absForce_r, in, b->cellStart, b->cellEnd, n, dt, need_fa);
    }
    /* the 8-byte report goes out AFTER collide: a device-to-host copy queued before it would sit behind whatever
     * the copy engine is busy with (the host-buffer step downloads 12 MB right then) and hold collide back */
    if (report) bin_send_report(B.scratch + 1);
    return;
  }
  {
    StageScope t(PRS_STAGE_REORDER);
    reorderDataAndFindCellStart(b->cellStart, b->cellEnd, b->sortedPos, b->sortedVel, b->sortedRad, b->hash, b->index,
                                b->pos, b->vel, b->rad, n, b->numCells);
  }
  StageScope t(PRS_STAGE_COLLIDE);
  prs_launch_collide((float2 *)b->vel, b->absForce_a, b->absForce_r, (const float2 *)b->sortedPos,
                     (const float2 *)b->sortedVel, b->sortedRad, b->index, b->cellStart, b->cellEnd, n, dt, need_fa);
}

}  // extern "C"

#include "prs_slab.cuh"
#include "prs_frame.cuh"

extern "C" {

/* headless frame (prs_frame.cuh): the reference's straight-down camera as a scale of the floor plane */
void prs_view_from_camera(prs_view *v, unsigned width, unsigned height, float camera_y, float light_radius) {
  v->width = width;
  v->height = height;
  v->center_x = 0.0f;
  v->center_y = 0.0f;
  v->world_per_pixel = 2.0f * camera_y * 0.57735026918962576f / (float)height; /* gluPerspective(60, ..), main.cpp:519 */
  v->light_radius = light_radius;
}
void prs_render_frame(unsigned char *d_bgr, unsigned *d_keys, const prs_view *view, const float *pos, const float *rad,
                      const float *col, unsigned n_points) {
  if (!view->width || !view->height || !(view->world_per_pixel > 0.0f)) {
    fprintf(stderr, "prs_render_frame: empty view\n");
    exit(EXIT_FAILURE);
  }
  prs::FrameView v;
  v.width = view->width; v.height = view->height;
  v.center_x = view->center_x; v.center_y = view->center_y;
  v.world_per_pixel = view->world_per_pixel; v.light_radius = view->light_radius;
  const size_t npix = (size_t)v.width * v.height;
  PRS_CUDA(cudaMemsetAsync(d_keys, 0xff, 2 * npix * sizeof(uint32_t), g_prs.stream));
  if (n_points) PRS_LAUNCH(prs::k_frame_splat, div_up(n_points, 256), 256, 0, d_keys, v, (const float2 *)pos, rad, n_points);
  PRS_LAUNCH(prs::k_frame_resolve, div_up((unsigned)npix, 256), 256, 0, d_bgr, (const uint32_t *)d_keys, v, (const float4 *)col);
}

void prs_init_hex_block(float *pos, float *vel, float *rad, float *phase, int *dead, unsigned n, unsigned nx, unsigned ny, float pitch,
                        float jitter, unsigned seed, float min_radius) {
  if (!n) return;
  (void)ny;
  const float row = pitch * 0.8660254037844386f; /* the host generator's constants, same expressions */
  const float x0 = -0.5f * ((float)(nx - 1) * pitch + 0.5f * pitch), y0 = -0.5f * (float)(ny - 1) * row;
  prs_bin_invalidate(); /* positions are rewritten behind the fused step's back */
  PRS_LAUNCH(k_init_hex_block, div_up(n, 256), 256, 0, (float2 *)pos, (float2 *)vel, rad, phase, dead, (unsigned long long)n, nx, pitch, row,
             x0, y0, jitter, seed, min_radius);
}

void prs_unpack_sorted(const float *sortedPR, float *sortedPos, float *sortedRad, unsigned n) {
  if (!n) return;
  PRS_LAUNCH(k_unpack_sorted, div_up(n, 256), 256, 0, (const float4 *)sortedPR, (float2 *)sortedPos, sortedRad, n);
}

unsigned long long prs_selftest_div(const float *d_x, const float *d_d, unsigned n) {
  unsigned long long *dm, h[4] = {0, 0, 0, 0};
  PRS_CUDA(cudaMalloc(&dm, sizeof(h)));
  PRS_CUDA(cudaMemsetAsync(dm, 0, sizeof(h), g_prs.stream));
  PRS_LAUNCH(k_selftest_div, div_up(n, 256), 256, 0, d_x, d_d, n, dm);
  PRS_CUDA(cudaMemcpyAsync(h, dm, sizeof(h), cudaMemcpyDeviceToHost, g_prs.stream));
  PRS_CUDA(cudaStreamSynchronize(g_prs.stream));
  PRS_CUDA(cudaFree(dm));
  if (h[0] || h[1] || h[2] || h[3])
    fprintf(stderr, "prs_selftest_div: div %llu sqrt %llu powf2 %llu div-by-sqrt %llu mismatches\n", h[0], h[1], h[2], h[3]);
  return h[0] + h[1] + h[2] + h[3];
}

void prs_stage_timing(int enable) {
  PRS_CUDA(cudaStreamSynchronize(g_prs.stream));
  g_prs.stage_timing = enable != 0;
  g_prs.spans.clear();
  g_prs.ev_used = 0;
}
/* sums the recorded spans per stage (ms) and their counts, then clears them */
void prs_stage_times(float *ms, unsigned *counts) {
  PRS_CUDA(cudaStreamSynchronize(g_prs.stream));
  for (int s = 0; s < PRS_NUM_STAGES; s++) { ms[s] = 0.0f; counts[s] = 0; }
  for (const auto &sp : g_prs.spans) {
    float t = 0.0f;
    PRS_CUDA(cudaEventElapsedTime(&t, sp.a, sp.b));
    ms[sp.stage] += t;
    counts[sp.stage]++;
  }
  g_prs.spans.clear();
  g_prs.ev_used = 0;
}

}  // extern "C"
