/*
 * prs_simparams.h — parameter block of the particle-robot update, binary-compatible with the
 * reference's `SimParams` (reference: particlebot_kernel.cuh:58-120; enums :30-55).
 *
 * The reference passes this struct by pointer to `setParameters` (particlebot_cuda.cu:111-123),
 * which uploads it to constant memory together with the seven obstacle arrays the embedded HOST
 * pointers refer to (capacity 10 each, particlebot_kernel_impl.cuh:28-34).  A drop-in must keep
 * field order, types and padding: sizeof == 256, alignment 8 (offsets checked below; the same
 * offsets are listed in SURVEY.md §8b).
 *
 * Plain C so that the oracle (gcc), the CUDA library (nvcc) and ctypes mirrors agree.
 */
#ifndef PRS_SIMPARAMS_H
#define PRS_SIMPARAMS_H

#include <stddef.h>

#ifdef __CUDACC__
#include <vector_types.h>
#else
/* host-only translation units (oracle, cfg parser) get layout-identical stand-ins */
#ifndef __VECTOR_TYPES_H__
typedef struct { unsigned int x, y; } prs_uint2_;
typedef struct { int x, y; } prs_int2_;
typedef struct __attribute__((aligned(8))) { float x, y; } prs_float2_;
#define uint2 prs_uint2_
#define int2 prs_int2_
#define float2 prs_float2_
#define PRS_OWN_VECTOR_TYPES 1
#endif
#endif

typedef unsigned int uint;

#define PRS_MAX_OBSTACLES 10 /* capacity of each constant obstacle array, kernel_impl.cuh:28-34 */

/* initial-placement selector (reference enum ParticlebotConfig; only CONFIG_RANDOM is reachable
 * from a cfg file, main.cpp:794-809 / :900) */
enum ParticlebotConfig {
  CONFIG_RANDOM, CONFIG_GRID, CONFIG_BLOB, CONFIG_BLOB_UPLEFT, CONFIG_HEX, CONFIG_LINE,
  CONFIG_LIGHTTEST_7, _NUM_CONFIGS
};
/* array selector of Particlebot::getArray/setArray (particlebot.cpp:803-867) */
enum ParticlebotArray { POSITION, VELOCITY, RADII, PHASE, FREQUENCY, DEAD };
enum ParticlebotControl { LIGHT_WAVE };

struct SimParams {
  uint2 gridSize;              /*   0 power-of-two cells per axis (hash wraps with &(size-1)) */
  uint numCells;               /*   8 gridSize.x*gridSize.y */
  float2 worldOrigin;          /*  16 */
  float2 cellSize;             /*  24 */
  uint nCells;                 /*  32 number of ROBOTS (reference naming) */
  int nDead;                   /*  36 -1 => last robot is the transported object */
  uint maxParticlebotsPerCell; /*  40 never set */
  float gravity;               /*  44 */
  float spring;                /*  48 */
  float damping;               /*  52 */
  float shear;                 /*  56 */
  float attraction;            /*  60 */
  float boundaryDamping;       /*  64 */
  float friction;              /*  68 */
  float massFactor;            /*  72 */
  float frictionFactor;        /*  76 */
  float radFactor;             /*  80 */
  float attractionFactor;      /*  84 */
  float constraint;            /*  88 */
  float constraint_contraction;/*  92 */
  int centroid_steps;          /*  96 */
  float centroid_int;          /* 100 */
  float centroid_radius;       /* 104 */
  float light_x;               /* 108 */
  float light_y;               /* 112 */
  float phase_update_interval; /* 116 */
  enum ParticlebotControl control; /* 120 */
  enum ParticlebotConfig config;   /* 124 */
  float min_radius;            /* 128 */
  float max_radius;            /* 132 */
  float rise_period;           /* 136 */
  float freq;                  /* 140 */
  int nobstacles;              /* 144 rectangular walls */
  float *x1obs;                /* 152 HOST pointers, read during setParameters */
  float *x2obs;                /* 160 */
  float *y1obs;                /* 168 */
  float *y2obs;                /* 176 */
  int n_cir_obstacles;         /* 184 disc obstacles */
  float *x_cir_obs;            /* 192 */
  float *y_cir_obs;            /* 200 */
  float *r_cir_obs;            /* 208 */
  int Nx;                      /* 216 */
  float phase_std;             /* 220 */
  unsigned seed;               /* 224 */
  uint light_shadow;           /* 228 */
  uint testing;                /* 232 */
  uint constrained_contraction;/* 236 */
  uint display_shadow;         /* 240 */
  float time_to_dead;          /* 244 */
  float max_time;              /* 248 */
};

#ifndef __cplusplus
typedef struct SimParams SimParams;
#endif

#ifdef __cplusplus
static_assert(sizeof(SimParams) == 256, "SimParams must stay 256 bytes");
static_assert(offsetof(SimParams, worldOrigin) == 16 && offsetof(SimParams, nCells) == 32 &&
              offsetof(SimParams, control) == 120 && offsetof(SimParams, nobstacles) == 144 &&
              offsetof(SimParams, x1obs) == 152 && offsetof(SimParams, n_cir_obstacles) == 184 &&
              offsetof(SimParams, x_cir_obs) == 192 && offsetof(SimParams, Nx) == 216 &&
              offsetof(SimParams, seed) == 224 && offsetof(SimParams, max_time) == 248,
              "SimParams offsets must match the reference layout");
#endif

#ifdef PRS_OWN_VECTOR_TYPES
#undef uint2
#undef int2
#undef float2
#endif

#endif /* PRS_SIMPARAMS_H */
