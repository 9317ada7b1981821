/*
 * prs_device.cuh — device-side parameter block and the small fp32 helpers shared by the kernels.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "prs_simparams.h"

/* What setParameters uploads (reference: `__constant__ SimParams params` + seven obstacle arrays,
 * particlebot_kernel_impl.cuh:27-34).  One block instead of eight symbols; the host pointers
 * embedded in `p` are meaningless on the device and never dereferenced there. */
struct PrsDevParams {
  SimParams p;
  float x1obs[PRS_MAX_OBSTACLES], x2obs[PRS_MAX_OBSTACLES], y1obs[PRS_MAX_OBSTACLES], y2obs[PRS_MAX_OBSTACLES];
  float x_cir[PRS_MAX_OBSTACLES], y_cir[PRS_MAX_OBSTACLES], r_cir[PRS_MAX_OBSTACLES];
  float world_half; /* wall of integrate; 64 in the reference (kernel_impl.cuh:75-97) */
};

/* the library is ONE CUDA translation unit (prs_kernels.cu); the constant block is defined here */
__constant__ PrsDevParams c_prm;

namespace prs {

/* Programmatic dependent launch (sm_90+): first statement of every kernel of the fused step.  When the
 * kernel was launched with cudaLaunchAttributeProgrammaticStreamSerialization its blocks may become resident
 * while the previous kernel of the stream drains; `wait` holds them until that kernel has completed and its
 * writes are visible, `launch_dependents` lets the NEXT kernel's blocks do the same behind this one.  Without
 * the attribute both are no-ops. */
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

struct v2 { float x, y; };
__device__ __forceinline__ v2 mk(float x, float y) { v2 r; r.x = x; r.y = y; return r; }
__device__ __forceinline__ v2 operator+(v2 a, v2 b) { return mk(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ v2 operator-(v2 a, v2 b) { return mk(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ v2 operator-(v2 a) { return mk(-a.x, -a.y); }
__device__ __forceinline__ v2 operator*(float s, v2 a) { return mk(a.x * s, a.y * s); }
__device__ __forceinline__ v2 operator*(v2 a, float s) { return mk(a.x * s, a.y * s); }
__device__ __forceinline__ v2 operator/(v2 a, float s) { return mk(a.x / s, a.y / s); } /* two IEEE divides */
__device__ __forceinline__ void operator+=(v2 &a, v2 b) { a.x += b.x; a.y += b.y; }
__device__ __forceinline__ void operator-=(v2 &a, v2 b) { a.x -= b.x; a.y -= b.y; }
__device__ __forceinline__ float dot2(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ float norm2(v2 a) { return sqrtf(dot2(a, a)); }

/* cell coordinates of a position: floor((p - origin) / cell) per axis with a true IEEE divide
 * (hashes must be bit-exact; reference calcGridPos kernel_impl.cuh:106-112) */
__device__ __forceinline__ int2 cell_of(float x, float y) {
  int2 g;
  g.x = (int)floorf((x - c_prm.p.worldOrigin.x) / c_prm.p.cellSize.x);
  g.y = (int)floorf((y - c_prm.p.worldOrigin.y) / c_prm.p.cellSize.y);
  return g;
}
/* wrap to the power-of-two grid and linearise row-major (reference calcGridHash :115-120) */
__device__ __forceinline__ uint32_t cell_hash(int gx, int gy) {
  const uint32_t mx = c_prm.p.gridSize.x - 1u, my = c_prm.p.gridSize.y - 1u;
  return ((uint32_t)gy & my) * c_prm.p.gridSize.x + ((uint32_t)gx & mx);
}

}  // namespace prs
