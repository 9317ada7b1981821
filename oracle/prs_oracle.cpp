/*
 * prs_oracle.cpp — CPU oracle of the particle-robot update.  TEST INFRASTRUCTURE ONLY (see
 * prs_oracle.h for who may load it and for the parity-pinning status).
 *
 * Every function restates the algorithm of the reference file:line it cites; nothing here is
 * shared with the product's CUDA code.  All arithmetic is IEEE fp32 evaluated operation by
 * operation (build: -O2 -ffp-contract=off, no -ffast-math), fp32 `time` accumulator, glibc
 * rand()/powf where the reference's HOST code uses them, XORWOW restated from the toolkit header
 * curand_kernel.h (CUDA 12.9) with the skip-ahead matrices of curand_precalc.h.
 */
#include "prs_oracle.h"

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

/* skip-ahead matrices of XORWOW: third-party data (CUDA toolkit 12.9, curand_precalc.h); the
 * *_host copy is a plain array once the CUDA qualifiers are defined away. */
#define __device__
#define __constant__
#include <curand_precalc.h>
#undef __device__
#undef __constant__

static int g_threads = 1;

extern "C" void prso_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
extern "C" int prso_get_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

#define PRSO_PAR_FOR _Pragma("omp parallel for schedule(static) num_threads(g_threads)")
#define PRSO_PAR_FOR_DYN _Pragma("omp parallel for schedule(dynamic, 256) num_threads(g_threads)")

struct f2 { float x, y; };
static inline float len2(float x, float y) { return sqrtf(x * x + y * y); } /* helper_math.h:1244,1287 */

/* ------------------------------------------------------------------------------------------
 * glibc rand(): additive feedback generator TYPE_3 (degree 31, separation 3), the stream the
 * reference consumes for placement and the dead draw (particlebot.cpp:29,187,647,677,684;
 * main.cpp:929).  Restated so that several simulations can hold independent streams.
 * ------------------------------------------------------------------------------------------ */
extern "C" void prso_srand(prso_glibc_rand *g, unsigned seed) {
  if (seed == 0) seed = 1;
  g->r[0] = (int)seed;
  for (int i = 1; i < 31; i++) {
    long hi = g->r[i - 1] / 127773, lo = g->r[i - 1] % 127773;
    long word = 16807 * lo - 2836 * hi;
    if (word < 0) word += 2147483647;
    g->r[i] = (int)word;
  }
  g->f = 3;
  g->b = 0;
  for (int i = 0; i < 310; i++) (void)prso_rand(g);
}
extern "C" int prso_rand(prso_glibc_rand *g) {
  uint32_t v = (uint32_t)g->r[g->f] + (uint32_t)g->r[g->b];
  g->r[g->f] = (int)v;
  g->f = (g->f + 1) % 31;
  g->b = (g->b + 1) % 31;
  return (int)(v >> 1);
}

/* ------------------------------------------------------------------------------------------
 * A1 hash  (particlebot_kernel_impl.cuh:106-120 calcGridPos/calcGridHash, :446-465 calcHashD)
 * ------------------------------------------------------------------------------------------ */
static inline void grid_pos(const SimParams *p, float x, float y, int *gx, int *gy) {
  *gx = (int)floorf((x - p->worldOrigin.x) / p->cellSize.x);
  *gy = (int)floorf((y - p->worldOrigin.y) / p->cellSize.y);
}
static inline unsigned grid_hash(const SimParams *p, int gx, int gy) {
  gx &= (int)(p->gridSize.x - 1);
  gy &= (int)(p->gridSize.y - 1);
  return (unsigned)gy * p->gridSize.x + (unsigned)gx;
}
extern "C" void prso_calc_hash(const SimParams *p, const float *pos, unsigned *hash, unsigned *index,
                               int n) {
  PRSO_PAR_FOR
  for (int i = 0; i < n; i++) {
    int gx, gy;
    grid_pos(p, pos[2 * i], pos[2 * i + 1], &gx, &gy);
    hash[i] = grid_hash(p, gx, gy);
    index[i] = (unsigned)i;
  }
}

/* stable sort of (hash,index) by hash: the semantics of thrust::sort_by_key at
 * particlebot_cuda.cu:377-382 (LSD radix sort => stable => ties keep ascending index). */
extern "C" void prso_sort_pairs(unsigned *hash, unsigned *index, int n) {
  std::vector<unsigned> h2(n), i2(n);
  unsigned *src_h = hash, *src_i = index, *dst_h = h2.data(), *dst_i = i2.data();
  for (int pass = 0; pass < 4; pass++) {
    size_t cnt[257] = {0};
    int sh = 8 * pass;
    for (int k = 0; k < n; k++) cnt[((src_h[k] >> sh) & 255) + 1]++;
    for (int d = 0; d < 256; d++) cnt[d + 1] += cnt[d];
    for (int k = 0; k < n; k++) {
      size_t o = cnt[(src_h[k] >> sh) & 255]++;
      dst_h[o] = src_h[k];
      dst_i[o] = src_i[k];
    }
    std::swap(src_h, dst_h);
    std::swap(src_i, dst_i);
  }
  /* 4 passes: result is back in the caller's arrays */
}

/* A2 cell table + gather (kernel_impl.cuh:469-538; memset particlebot_cuda.cu:301).  cellEnd is
 * NOT cleared: entries of cells that emptied keep their previous value. */
extern "C" void prso_reorder_find_cell_start(const SimParams *, unsigned *cellStart,
                                             unsigned *cellEnd, float *sortedPos, float *sortedVel,
                                             float *sortedRad, const unsigned *hash,
                                             const unsigned *index, const float *pos,
                                             const float *vel, const float *rad, int n,
                                             unsigned numCells) {
  memset(cellStart, 0xff, (size_t)numCells * sizeof(unsigned));
  PRSO_PAR_FOR
  for (int k = 0; k < n; k++) {
    unsigned h = hash[k];
    if (k == 0 || h != hash[k - 1]) {
      cellStart[h] = (unsigned)k;
      if (k > 0) cellEnd[hash[k - 1]] = (unsigned)k;
    }
    if (k == n - 1) cellEnd[h] = (unsigned)k + 1;
    unsigned s = index[k];
    sortedPos[2 * k] = pos[2 * s];
    sortedPos[2 * k + 1] = pos[2 * s + 1];
    sortedVel[2 * k] = vel[2 * s];
    sortedVel[2 * k + 1] = vel[2 * s + 1];
    sortedRad[k] = rad[s];
  }
}

/* A3 pair force (kernel_impl.cuh:541-594 collideSpheres).  __powf(x,2) of the device code is
 * restated as the exact square (the device value is an ex2(2*lg2 x) approximation, Q7). */
static inline void pair_force(const SimParams *p, f2 pa, f2 pb, f2 va, f2 vb, float ra, float rb,
                              float attraction, f2 *force, float *forcea, float *forcer) {
  float rx = pb.x - pa.x, ry = pb.y - pa.y;
  /* The ONE place where the oracle follows nvcc's FMA contraction (PTX of the reference build:
   * mul ry*ry; fma rx*rx + that; sqrt.rn): the model is discontinuous at dist == ra+rb (contact
   * spring ~0 vs attraction 2.5) and the aggregation placement puts pairs exactly at touching
   * distance, so the regime decision must see the device's bits of `dist`. */
  float dist = sqrtf(fmaf(rx, rx, ry * ry));
  float cd = ra + rb;
  float tx = 0.0f, ty = 0.0f;
  if (dist < cd) {
    float nx = rx / dist, ny = ry / dist;
    float rvx = vb.x - va.x, rvy = vb.y - va.y;
    float dn = rvx * nx + rvy * ny;
    float tvx = rvx - dn * nx, tvy = rvy - dn * ny;
    float s = -p->spring * (cd - dist);
    tx += s * nx;            ty += s * ny;
    tx += p->damping * rvx;  ty += p->damping * rvy;
    tx += p->shear * tvx;    ty += p->shear * tvy;
    force->x += tx;          force->y += ty;
    *forcer += len2(tx, ty);
  } else {
    const float int1 = 0.0009f, int2 = 0.0019f, min_attr = 2.5f;
    float gap = dist - cd;
    float ux = rx / dist, uy = ry / dist;
    if (gap < int1) {
      tx += min_attr * ux;   ty += min_attr * uy;
    } else if (gap < int2) {
      float m = min_attr + (attraction / (int2 * int2) - min_attr) / (int2 - int1) * (gap - int1);
      tx += m * ux;          ty += m * uy;
    } else {
      float g2 = gap * gap;
      tx += attraction * ux / g2;  ty += attraction * uy / g2;
    }
    force->x += tx;          force->y += ty;
    *forcea += len2(tx, ty);
  }
}

/* A3-A6 collide (kernel_impl.cuh:597-653 collideCell, :657-831 collideD) */
extern "C" void prso_collide(const SimParams *p, float *newVel, float *absForce_a, float *absForce_r,
                             const float *sortedPos, const float *sortedVel, const float *sortedRad,
                             const unsigned *index, const unsigned *cellStart,
                             const unsigned *cellEnd, int n, float dt) {
  const bool object_mode = (p->nDead == -1);
  const unsigned obj = p->nCells - 1;
  PRSO_PAR_FOR_DYN
  for (int k = 0; k < n; k++) {
    f2 pos = {sortedPos[2 * k], sortedPos[2 * k + 1]};
    f2 vel = {sortedVel[2 * k], sortedVel[2 * k + 1]};
    float rad = sortedRad[k];
    int gx, gy;
    grid_pos(p, pos.x, pos.y, &gx, &gy);
    f2 force = {0.0f, 0.0f};
    unsigned orig = index[k];
    float fa = 0.0f;
    float fr = 0.0f * absForce_r[orig]; /* :688 — a NaN there sticks (Q6) */
    float att1 = (object_mode && orig == obj) ? p->attractionFactor : 1.0f;
    for (int y = -2; y <= 2; y++) {
      for (int x = -2; x <= 2; x++) {
        unsigned h = grid_hash(p, gx + x, gy + y);
        unsigned s = cellStart[h];
        if (s == 0xffffffffu) continue;
        unsigned e = cellEnd[h];
        for (unsigned j = s; j < e; j++) {
          if (j == (unsigned)k) continue;
          float att2 = (object_mode && index[j] == obj) ? p->attractionFactor : 1.0f;
          f2 p2 = {sortedPos[2 * j], sortedPos[2 * j + 1]};
          f2 v2 = {sortedVel[2 * j], sortedVel[2 * j + 1]};
          pair_force(p, pos, p2, vel, v2, rad, sortedRad[j], p->attraction * att2 * att1, &force,
                     &fa, &fr);
        }
      }
    }
    /* A4 disc obstacles (:703-728); device powf(x,2)/powf(x,.5f) -> glibc powf */
    for (int i = 0; i < p->n_cir_obstacles; i++) {
      float ox = p->x_cir_obs[i], oy = p->y_cir_obs[i], orad = p->r_cir_obs[i];
      float d2 = powf(pos.x - ox, 2.0f) + powf(pos.y - oy, 2.0f);
      if (d2 < powf(rad + orad, 2.0f)) {
        float dx = -pos.x + ox, dy = -pos.y + oy;
        float l = len2(dx, dy);
        dx = dx / l;  dy = dy / l;
        float rvx = -vel.x, rvy = -vel.y;
        float dn = rvx * dx + rvy * dy;
        float tvx = rvx - dn * dx, tvy = rvy - dn * dy;
        float s = 2.0f * p->spring * (rad + orad - powf(d2, 0.5f));
        float tx = 0.0f, ty = 0.0f;
        tx += s * (-dx);          ty += s * (-dy);
        tx += p->damping * rvx;   ty += p->damping * rvy;
        tx += p->shear * tvx;     ty += p->shear * tvy;
        force.x += tx;            force.y += ty;
        fr += len2(tx, ty);
      }
    }
    /* A5 rectangular obstacles (:729-798) */
    for (int i = 0; i < p->nobstacles; i++) {
      float x1 = p->x1obs[i], x2 = p->x2obs[i], y1 = p->y1obs[i], y2 = p->y2obs[i];
      int hit = 0;
      float dx = 0.0f, dy = 0.0f, ov = 0.0f;
      auto corner = [&](float cx, float cy) {
        float ex = pos.x - cx, ey = pos.y - cy;
        float l = len2(ex, ey);
        dx = -ex / l;  dy = -ey / l;
        hit = 1;
        ov = rad - powf(powf(pos.x - cx, 2.0f) + powf(pos.y - cy, 2.0f), 0.5f);
      };
      auto in_corner = [&](float cx, float cy) {
        return powf(pos.x - cx, 2.0f) + powf(pos.y - cy, 2.0f) < powf(rad, 2.0f);
      };
      if (pos.y > y1 && pos.y < y2) {
        if (pos.x > x1 - rad && pos.x < x2 - rad) { hit = 1; dx = 1.0f; dy = 0.0f; ov = pos.x - x1 + rad; }
        if (pos.x < x2 + rad && pos.x > x1 + rad) { hit = 1; dx = -1.0f; dy = 0.0f; ov = -pos.x + x2 + rad; }
      } else if (pos.x > x1 && pos.x < x2) {
        if (pos.y > y1 - rad && pos.y < y2 - rad) { hit = 1; dx = 0.0f; dy = 1.0f; ov = pos.y - y1 + rad; }
        if (pos.y < y2 + rad && pos.y > y1 + rad) { hit = 1; dx = 0.0f; dy = -1.0f; ov = -pos.y + y2 + rad; }
      } else if (in_corner(x2, y2)) corner(x2, y2);
      else if (in_corner(x1, y2)) corner(x1, y2);
      else if (in_corner(x1, y1)) corner(x1, y1);
      else if (in_corner(x2, y1)) corner(x2, y1);
      if (hit) {
        float rvx = -vel.x, rvy = -vel.y;
        float dn = rvx * dx + rvy * dy;
        float tvx = rvx - dn * dx, tvy = rvy - dn * dy;
        float s = -2.0f * p->spring * ov;
        float tx = 0.0f, ty = 0.0f;
        tx += s * dx;             ty += s * dy;
        tx += p->damping * rvx;   ty += p->damping * rvy;
        tx += p->shear * tvx;     ty += p->shear * tvy;
        force.x += tx;            force.y += ty;
        fr += len2(tx, ty);
      }
    }
    /* A6 friction + velocity (:801-830) */
    float friction = p->friction, gravity = p->gravity;
    const bool is_obj = object_mode && orig == obj;
    if (is_obj) { friction *= p->frictionFactor; gravity *= p->massFactor; }
    if (len2(vel.x, vel.y) < 0.000001f && len2(force.x, force.y) < (2.0f * friction * gravity)) {
      force.x = 0.0f;  force.y = 0.0f;
    }
    if (is_obj) {
      vel.x = vel.x + force.x / p->massFactor * dt;
      vel.y = vel.y + force.y / p->massFactor * dt;
    } else {
      vel.x = vel.x + force.x * dt;
      vel.y = vel.y + force.y * dt;
    }
    float kin = friction * gravity * dt;
    float vl = len2(vel.x, vel.y);
    if (vl < kin) {
      vel.x = 0.0f;  vel.y = 0.0f;
    } else {
      vel.x -= kin * (vel.x / vl);
      vel.y -= kin * (vel.y / vl);
    }
    newVel[2 * orig] = vel.x;
    newVel[2 * orig + 1] = vel.y;
    absForce_a[orig] = fa;
    absForce_r[orig] = fr;
  }
}

/* A7 integrate (kernel_impl.cuh:53-103).  The reference hard-codes the wall at 64; world_half
 * makes it a parameter for the synthetic swarms. */
extern "C" void prso_integrate(const SimParams *p, float *pos, float *vel, const float *rad, float dt,
                               int n, float world_half) {
  PRSO_PAR_FOR
  for (int i = 0; i < n; i++) {
    float x = pos[2 * i], y = pos[2 * i + 1], vx = vel[2 * i], vy = vel[2 * i + 1], r = rad[i];
    x += vx * dt;
    y += vy * dt;
    if (x > world_half - r) { x = world_half - r; vx *= p->boundaryDamping; }
    if (x < -world_half + r) { x = -world_half + r; vx *= p->boundaryDamping; }
    if (y > world_half - r) { y = world_half - r; vy *= p->boundaryDamping; }
    if (y < -world_half + r) { y = -world_half + r; vy *= p->boundaryDamping; }
    pos[2 * i] = x;  pos[2 * i + 1] = y;
    vel[2 * i] = vx; vel[2 * i + 1] = vy;
  }
}

/* A8 radius-phase controller (kernel_impl.cuh:124-181) */
extern "C" void prso_update_rad(const SimParams *p, const float *absForce_a, const float *absForce_r,
                                float *rad, const float *phase, float time, float dt,
                                const int *dead, int n) {
  PRSO_PAR_FOR
  for (int i = 0; i < n; i++) {
    if (dead[i]) continue;
    if (phase[i] > 10000000.0f) continue;
    float period = (float)(p->Nx + 1) * p->rise_period;
    float t1 = time + phase[i];
    if (t1 < 0) t1 = t1 + (float)(100 * (p->Nx + 1)) * p->rise_period;
    if (t1 >= period) t1 = t1 - period * floorf(t1 / period);
    if (t1 >= 2 * p->rise_period) continue;
    float target;
    if (t1 <= p->rise_period)
      target = p->min_radius + (p->max_radius - p->min_radius) / p->rise_period * t1;
    else
      target = p->max_radius + (p->min_radius - p->max_radius) / p->rise_period * (t1 - p->rise_period);
    float r = rad[i];
    float dr1 = target - r;
    float dr = 0;
    const float max_speed = 0.1f;
    float torque = dr1 * p->constraint * r / max_speed / p->max_radius / dt;
    torque = fminf(torque, p->constraint);
    if (dr1 > 0) {
      if (torque / r > absForce_r[i])
        dr = max_speed * p->max_radius / p->constraint * (torque / r - absForce_r[i]) * dt;
    } else {
      if (p->constrained_contraction) {
        if (-p->constraint_contraction * dr1 > absForce_a[i] * r)
          dr = (p->constraint_contraction * dr1 + absForce_a[i] * r) / (p->constraint_contraction);
        dr = fmaxf(dr, -p->max_radius * dt);
      } else {
        dr = dr1;
      }
    }
    dr = r + dr;
    if (dr > p->max_radius) dr = p->max_radius;
    if (dr < p->min_radius) dr = p->min_radius;
    rad[i] = dr;
  }
}

/* A9 shadow tests (kernel_impl.cuh:184-209 segment, :211-236 disc, :238-262 driver) */
static int hit_segment(float x0, float y0, float x1, float y1, float x3, float y3, float x4, float y4) {
  if (fabsf((x4 - x3) / (x1 - x0)) == fabsf((y4 - y3) / (y1 - y0))) return 0;
  float t, t1;
  if (fabsf(y4 - y3) > 0) {
    t = (x3 - x0 - (y3 - y0) * (x3 - x4) / (y3 - y4)) *
        ((y3 - y4) / ((x1 - x0) * (y3 - y4) - (y1 - y0) * (x3 - x4)));
    if (t <= 0 || t >= 1) return 0;
    t1 = (y3 - y0 - t * (y1 - y0)) / (y3 - y4);
    if (t1 <= 0 || t1 >= 1) return 0;
  } else if (fabsf(x4 - x3) > 0) {
    t = (y3 - y0 - (x3 - x0) * (y3 - y4) / (x3 - x4)) *
        ((x3 - x4) / ((y1 - y0) * (x3 - x4) - (x1 - x0) * (y3 - y4)));
    if (t <= 0 || t >= 1) return 0;
    t1 = (x3 - x0 - t * (x1 - x0)) / (x3 - x4);
    if (t1 <= 0 || t1 >= 1) return 0;
  } else {
    return 0;
  }
  return 1;
}
static int hit_disc(float lx, float ly, float px, float py, float ox, float oy, float orad) {
  float C1 = powf(lx, 2) + powf(ly, 2), C2 = powf(px, 2) + powf(py, 2), C3 = powf(ox, 2) + powf(oy, 2);
  float C4 = lx * px + ly * py, C5 = lx * ox + ly * oy, C6 = px * ox + py * oy;
  float A = C1 + C2 - 2 * C4;
  float B = -2 * C1 + 2 * C4 + 2 * C5 - 2 * C6;
  float C = C1 + C3 - 2 * C5 - powf(orad, 2);
  float D = powf(B, 2) - 4 * A * C;
  if (D >= 0) {
    float R1 = (-B + powf(D, 0.5f)) / 2 / A, R2 = (-B - powf(D, 0.5f)) / 2 / A;
    if (R1 > 0 && R1 < 1) return 1;
    if (R2 > 0 && R2 < 1) return 1;
  }
  return 0;
}
static int in_shadow(const SimParams *p, float px, float py) {
  for (int i = 0; i < p->n_cir_obstacles; i++)
    if (hit_disc(p->light_x, p->light_y, px, py, p->x_cir_obs[i], p->y_cir_obs[i], p->r_cir_obs[i]))
      return 1;
  for (int i = 0; i < p->nobstacles; i++) {
    float x1 = p->x1obs[i], x2 = p->x2obs[i], y1 = p->y1obs[i], y2 = p->y2obs[i];
    if (hit_segment(p->light_x, p->light_y, px, py, x1, y1, x1, y2)) return 1; /* left */
    if (hit_segment(p->light_x, p->light_y, px, py, x1, y2, x2, y2)) return 1; /* top */
    if (hit_segment(p->light_x, p->light_y, px, py, x2, y2, x2, y1)) return 1; /* right */
    if (hit_segment(p->light_x, p->light_y, px, py, x2, y1, x1, y1)) return 1; /* bottom */
  }
  return 0;
}

/* A9 phase offsets (kernel_impl.cuh:264-290) */
extern "C" void prso_update_phase(const SimParams *p, const float *pos, float *phase, float spacing,
                                  float min_d, int n) {
  PRSO_PAR_FOR
  for (int i = 0; i < n; i++) {
    float px = pos[2 * i], py = pos[2 * i + 1];
    float dist = len2(px - p->light_x, py - p->light_y);
    int visible = 1;
    if (p->light_shadow && in_shadow(p, px, py)) visible = 0;
    if (!visible) {
      if (p->light_shadow == 1) phase[i] = -(float)(p->Nx - 1) * p->rise_period;
      if (p->light_shadow == 2) phase[i] = 9999999999.0f;
    } else {
      phase[i] = (min_d - dist) / (spacing)*p->rise_period;
    }
  }
}

/* host min distance to the light (particlebot.cpp:214-228), glibc powf like the reference */
extern "C" float prso_min_light_distance(const SimParams *p, const float *pos, int n) {
  float min_d = 0;
  for (int i = 0; i < n; i++) {
    float d = powf(powf(p->light_x - pos[2 * i], 2) + powf(p->light_y - pos[2 * i + 1], 2), 0.5f);
    if (i == 0) min_d = d;
    else min_d = (min_d < d ? min_d : d);
  }
  return min_d;
}

/* ------------------------------------------------------------------------------------------
 * XORWOW (toolkit curand_kernel.h: _curand_init_inplace :800-826, _skipahead_sequence_inplace
 * :721-736, __curand_matvec_inplace :316-334, curand() :863-874; curand_normal.h:70-87,313-326).
 * The reference seeds one generator per robot with curand_init(seed, i, 0) (kernel_impl.cuh:36-41)
 * and adds std*N(0,1) to each phase (:43-51).  Integer state is bit-exact; the Box-Muller floats
 * use host logf/sinf/cosf where the device uses logf/__sincosf (differences ~1e-7 abs).
 * ------------------------------------------------------------------------------------------ */
static void xorwow_matvec5(unsigned *v, const unsigned *m) {
  unsigned r[5] = {0, 0, 0, 0, 0};
  for (int i = 0; i < 5; i++)
    for (int j = 0; j < 32; j++)
      if (v[i] & (1u << j))
        for (int k = 0; k < 5; k++) r[k] ^= m[5 * (i * 32 + j) + k];
  for (int i = 0; i < 5; i++) v[i] = r[i];
}
static unsigned xorwow_next(prso_rng_state *s) {
  unsigned t = (s->v[0] ^ (s->v[0] >> 2));
  s->v[0] = s->v[1]; s->v[1] = s->v[2]; s->v[2] = s->v[3]; s->v[3] = s->v[4];
  s->v[4] = (s->v[4] ^ (s->v[4] << 4)) ^ (t ^ (t << 1));
  s->d += 362437;
  return s->v[4] + s->d;
}
extern "C" void prso_curand_setup(prso_rng_state *st, unsigned seed32, int n) {
  PRSO_PAR_FOR
  for (int i = 0; i < n; i++) {
    prso_rng_state *s = &st[i];
    unsigned long long seed = seed32; /* params.seed is `unsigned`, widened at the call */
    unsigned s0 = ((unsigned)seed) ^ 0xaad26b49u, s1 = (unsigned)(seed >> 32) ^ 0xf7dcefddu;
    unsigned t0 = 1099087573u * s0, t1 = 2591861531u * s1;
    s->d = 6615241u + t1 + t0;
    s->v[0] = 123456789u + t0; s->v[1] = 362436069u ^ t0; s->v[2] = 521288629u + t1;
    s->v[3] = 88675123u ^ t1;  s->v[4] = 5783321u + t0;
    unsigned long long x = (unsigned long long)i; /* subsequence = robot id */
    int matrix_num = 0;
    while (x) {
      for (unsigned t = 0; t < (x & PRECALC_BLOCK_MASK); t++)
        xorwow_matvec5(s->v, precalc_xorwow_matrix_host[matrix_num]);
      x >>= PRECALC_BLOCK_SIZE;
      matrix_num++;
    }
    s->boxmuller_flag = 0; s->boxmuller_flag_double = 0;
    s->boxmuller_extra = 0.f; s->pad_ = 0; s->boxmuller_extra_double = 0.;
  }
}
extern "C" void prso_add_normal_noise(prso_rng_state *st, float *val, float std, int n) {
  PRSO_PAR_FOR
  for (int i = 0; i < n; i++) {
    prso_rng_state *s = &st[i];
    float z;
    if (s->boxmuller_flag != 1) {
      unsigned x = xorwow_next(s), y = xorwow_next(s);
      float u = x * 2.3283064e-10f + (2.3283064e-10f / 2);
      float v = y * (2.3283064e-10f * 6.2831855f) + ((2.3283064e-10f * 6.2831855f) / 2);
      float r = sqrtf(-2.0f * logf(u));
      z = sinf(v) * r;
      s->boxmuller_extra = cosf(v) * r;
      s->boxmuller_flag = 1;
    } else {
      s->boxmuller_flag = 0;
      z = s->boxmuller_extra;
    }
    val[i] += std * z;
  }
}

/* ------------------------------------------------------------------------------------------
 * Simulation object: buffers of Particlebot::_initialize (particlebot.cpp:77-166; force
 * accumulators zeroed = the reference's de-facto behaviour, Q6), reset and update.
 * ------------------------------------------------------------------------------------------ */
struct PrsOracle {
  SimParams p;
  float world_half;
  float ox[7][PRS_MAX_OBSTACLES];
  int n;
  float time;
  prso_glibc_rand rng;
  std::vector<float> pos, vel, rad, phase, fa, fr, spos, svel, srad;
  std::vector<int> dead;
  std::vector<unsigned> hash, index, cellStart, cellEnd;
  std::vector<prso_rng_state> state;
};

extern "C" int prso_gate(float time, float interval, float dt) {
  return time - interval * floorf(time / interval) < dt;
}

extern "C" PrsOracle *prso_create(const SimParams *p, float world_half) {
  PrsOracle *o = new PrsOracle();
  o->p = *p;
  o->world_half = world_half;
  float *const *src[7] = {&p->x1obs, &p->x2obs, &p->y1obs, &p->y2obs, &p->x_cir_obs, &p->y_cir_obs, &p->r_cir_obs};
  float **dst[7] = {&o->p.x1obs, &o->p.x2obs, &o->p.y1obs, &o->p.y2obs, &o->p.x_cir_obs, &o->p.y_cir_obs, &o->p.r_cir_obs};
  for (int a = 0; a < 7; a++) {
    int cnt = a < 4 ? p->nobstacles : p->n_cir_obstacles;
    for (int i = 0; i < PRS_MAX_OBSTACLES; i++) o->ox[a][i] = (i < cnt && *src[a]) ? (*src[a])[i] : 0.0f;
    *dst[a] = o->ox[a];
  }
  int n = o->n = (int)p->nCells;
  o->time = 0;
  o->pos.assign(2 * (size_t)n, 0); o->vel.assign(2 * (size_t)n, 0); o->rad.assign(n, 0);
  o->phase.assign(n, 0); o->fa.assign(n, 0); o->fr.assign(n, 0);
  o->spos.assign(2 * (size_t)n, 0); o->svel.assign(2 * (size_t)n, 0); o->srad.assign(n, 0);
  o->dead.assign(n, 0); o->hash.assign(n, 0); o->index.assign(n, 0);
  o->cellStart.assign(p->numCells, 0); o->cellEnd.assign(p->numCells, 0);
  o->state.resize(n);
  prso_srand(&o->rng, 1);
  prso_curand_setup(o->state.data(), p->seed, n); /* particlebot.cpp:165 */
  return o;
}
extern "C" void prso_destroy(PrsOracle *o) { delete o; }
extern "C" void prso_srand_sim(PrsOracle *o, unsigned seed) { prso_srand(&o->rng, seed); }
extern "C" float prso_time(const PrsOracle *o) { return o->time; }
extern "C" void *prso_array(PrsOracle *o, int which) {
  switch (which) {
    case 0: return o->pos.data();   case 1: return o->vel.data();   case 2: return o->rad.data();
    case 3: return o->phase.data(); case 4: return o->fa.data();    case 5: return o->fr.data();
    case 6: return o->dead.data();  case 7: return o->hash.data();  case 8: return o->index.data();
    case 9: return o->cellStart.data(); case 10: return o->cellEnd.data();
    case 11: return o->spos.data(); case 12: return o->svel.data(); case 13: return o->srad.data();
    case 14: return o->state.data();
  }
  return nullptr;
}

static inline float host_length(float x, float y) { /* particlebot.cpp:32-34 */
  return powf(powf(x, 2.0f) + powf(y, 2.0f), 0.5f);
}

/* Particlebot::reset, CONFIG_RANDOM branch (particlebot.cpp:612-748) + radii/dead/phase init
 * (:775-800).  Sequential "random aggregation": every new disc is attached to a random already
 * placed one and pivoted in 10-degree steps until it would overlap. */
extern "C" void prso_reset(PrsOracle *o) {
  const SimParams &P = o->p;
  const int n = o->n;
  float *hPos = o->pos.data(), *hVel = o->vel.data();
  o->time = 0;
  const int GX = (int)P.gridSize.x, GY = (int)P.gridSize.y;
  std::vector<std::vector<int>> cells((size_t)GX * GY);
  auto cell_of = [&](float x, float y, int *cx, int *cy) {
    *cx = ((int)floorf((x - P.worldOrigin.x) / P.cellSize.x)) & (GX - 1);
    *cy = ((int)floorf((y - P.worldOrigin.y) / P.cellSize.y)) & (GY - 1);
  };
  /* the reference indexes its occupancy grid with unwrapped xg-1..xg+1 (undefined at the grid
   * edge); wrapping is the only defined completion and never triggers for the example seeds */
  auto bucket = [&](int cx, int cy) -> std::vector<int> & {
    return cells[(size_t)(cx & (GX - 1)) * GY + (cy & (GY - 1))];
  };
  const float PI_F = 3.141592654f;
  int p = 0, v = 0, xg, yg, xgs, ygs;
  unsigned placed = 0, start_ind = 0, max_fail = 200, fails = 0;
  if (n > 0) {
    hPos[p++] = 5.0f; hPos[p++] = 0.0f; hVel[v++] = 0.0f; hVel[v++] = 0.0f;
    cell_of(0.0f, 0.0f, &xg, &yg); /* :635-637 registers disc 0 under the cell of the ORIGIN */
    bucket(xg, yg).push_back(0);
  }
  float x = 0, y = 0, theta = 0, r = 0, old_theta = 0, min_x = 9999999.0f;
  float increment_theta = (float)(2 * PI_F / 360.0 * 10.0);
  for (int i = 1; i < n; i++) {
    if (i == 2) { /* :646-672 third disc sits perpendicular to the first pair */
      int j = prso_rand(&o->rng) % 2;
      float dx = hPos[2] - hPos[0], dy = hPos[3] - hPos[1];
      float l = host_length(dx, dy);
      dy = dy / l; dx = dx / l;
      float ex, ey;
      if (j) { ex = dy; ey = -dx; } else { ex = -dy; ey = dx; }
      x = (hPos[2] + hPos[0]) / 2.0f + ex * P.min_radius;
      y = (hPos[3] + hPos[1]) / 2.0f + ey * P.min_radius;
      if (x < min_x) min_x = x;
      hPos[p++] = x; hPos[p++] = y; hVel[v++] = 0.0f; hVel[v++] = 0.0f;
      cell_of(hPos[2 * i], hPos[2 * i + 1], &xg, &yg);
      bucket(xg, yg).push_back(i);
      continue;
    }
    placed = 0;
    r = P.min_radius;
    while (!placed) {
      start_ind = (unsigned)prso_rand(&o->rng) % (unsigned)i;
      placed = 1;
      if (fails == max_fail) { fails = 0; r += P.min_radius; }
      theta = 2 * (prso_rand(&o->rng) / (float)RAND_MAX) * PI_F;
      x = hPos[2 * start_ind] + 2 * r * cosf(theta);
      y = hPos[2 * start_ind + 1] + 2 * r * sinf(theta);
      cell_of(x, y, &xgs, &ygs);
      for (xg = xgs - 1; (xg <= xgs + 1) & placed; xg++)
        for (yg = ygs - 1; (yg <= ygs + 1) & placed; yg++) {
          std::vector<int> &b = bucket(xg, yg);
          for (size_t t = 0; t < b.size() && placed; t++)
            if (host_length(x - hPos[2 * b[t]], y - hPos[2 * b[t] + 1]) < 2 * 1.0 * P.min_radius) {
              placed = 0;
              fails++;
              break;
            }
        }
      if (!placed) continue;
      old_theta = theta;
      int flag = 0;
      while (theta - old_theta < 2 * PI_F) {
        theta += increment_theta;
        x = hPos[2 * start_ind] + 2 * r * cosf(theta);
        y = hPos[2 * start_ind + 1] + 2 * r * sinf(theta);
        cell_of(x, y, &xgs, &ygs);
        for (xg = xgs - 1; xg <= xgs + 1; xg++)
          for (yg = ygs - 1; yg <= ygs + 1; yg++) {
            std::vector<int> &b = bucket(xg, yg);
            for (size_t t = 0; t < b.size(); t++)
              if (host_length(x - hPos[2 * b[t]], y - hPos[2 * b[t] + 1]) < 2 * 1.0 * P.min_radius) {
                flag = 1;
                break; /* leaves only the innermost loop, as in the reference */
              }
          }
        if (flag) { theta -= increment_theta; break; }
      }
      x = hPos[2 * start_ind] + 2 * r * cosf(theta);
      y = hPos[2 * start_ind + 1] + 2 * r * sinf(theta);
    }
    if (x < min_x) min_x = x;
    if (P.nDead == -1 && i == n - 1) { /* :731-735 the object starts left of the swarm */
      x = min_x - 1 * P.min_radius * P.radFactor - 2 * P.min_radius;
      y = 0;
    }
    hPos[p++] = x; hPos[p++] = y;
    cell_of(x, y, &xg, &yg);
    bucket(xg, yg).push_back(i);
    hVel[v++] = 0.0f; hVel[v++] = 0.0f;
  }
  for (int i = 0; i < n; i++) { /* :784-791 */
    o->rad[i] = P.min_radius;
    o->dead[i] = 0;
    if (P.nDead == -1 && i == n - 1) { o->rad[i] = P.min_radius * P.radFactor; o->dead[i] = 1; }
    o->phase[i] = 0;
  }
  std::fill(o->fa.begin(), o->fa.end(), 0.0f);
  std::fill(o->fr.begin(), o->fr.end(), 0.0f);
}

/* Particlebot::update (particlebot.cpp:170-300), headless: calcCOG (:207-209) and updateCol
 * (:254) only feed the renderer and are omitted; the max_time exit (:174-176) is the caller's. */
extern "C" void prso_update(PrsOracle *o, float dt, float sort_interval) {
  SimParams &P = o->p;
  const int n = o->n;
  float time = o->time;
  if (time >= P.time_to_dead && time < P.time_to_dead + dt) { /* :178-194 dead draw */
    std::vector<int> inds(n);
    for (int i = 0; i < n; i++) inds[i] = i;
    int count = 0;
    while (count < P.nDead) {
      int i = (int)((unsigned long)prso_rand(&o->rng) % inds.size());
      o->dead[inds[i]] = 1;
      inds.erase(inds.begin() + i);
      count++;
    }
  }
  if (P.control == LIGHT_WAVE) {
    if (prso_gate(time, P.phase_update_interval, dt)) { /* :212-237 */
      float min_d = prso_min_light_distance(&P, o->pos.data(), n);
      float spacing = 2.0f * P.min_radius;
      prso_update_phase(&P, o->pos.data(), o->phase.data(), spacing, min_d, n);
      if (P.phase_std) prso_add_normal_noise(o->state.data(), o->phase.data(), P.phase_std, n);
    }
    if (time >= 0)
      prso_update_rad(&P, o->fa.data(), o->fr.data(), o->rad.data(), o->phase.data(), time, dt,
                      o->dead.data(), n);
  }
  prso_integrate(&P, o->pos.data(), o->vel.data(), o->rad.data(), dt, n, o->world_half);
  if (prso_gate(time, sort_interval, dt)) { /* :256-268 */
    prso_calc_hash(&P, o->pos.data(), o->hash.data(), o->index.data(), n);
    prso_sort_pairs(o->hash.data(), o->index.data(), n);
  }
  prso_reorder_find_cell_start(&P, o->cellStart.data(), o->cellEnd.data(), o->spos.data(),
                               o->svel.data(), o->srad.data(), o->hash.data(), o->index.data(),
                               o->pos.data(), o->vel.data(), o->rad.data(), n, P.numCells);
  prso_collide(&P, o->vel.data(), o->fa.data(), o->fr.data(), o->spos.data(), o->svel.data(),
               o->srad.data(), o->index.data(), o->cellStart.data(), o->cellEnd.data(), n, dt);
  o->time = time + dt;
}
