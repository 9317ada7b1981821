/*
 * prs_host_state.h — host-side state of libparticlebot_b200 (stream, parameter shadow, sort
 * scratch, launch counter) and the error macro.  One context per process, like the reference's
 * global `__constant__ params` (particlebot_kernel_impl.cuh:27).
 */
#pragma once
#include <cuda_runtime.h>
#include "prs_device.cuh"
#include "prs_onesweep.cuh"

struct PrsHostState {
  cudaStream_t stream = 0;            /* legacy default stream, like the reference */
  unsigned long long launches = 0;    /* kernels launched by this library */
  PrsDevParams h_prm = {};            /* shadow of the constant block */
  bool params_set = false;
  float world_half = 64.0f;           /* reference wall (kernel_impl.cuh:75-97) */
  int collide_mode = 0;               /* 0 exact, 1 fast */
  prs_sort::Workspace sort_ws;
};
extern PrsHostState g_prs;

void prs_fail(const char *what, cudaError_t e, const char *file, int line);
#define PRS_CUDA(call)                                                 \
  do {                                                                 \
    cudaError_t e_ = (call);                                           \
    if (e_ != cudaSuccess) prs_fail(#call, e_, __FILE__, __LINE__);    \
  } while (0)
