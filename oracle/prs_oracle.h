/*
 * prs_oracle.h — C API of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  The oracle is a CPU restatement (C++/OpenMP, IEEE fp32, no
 * fast-math, no FMA contraction) of the reference's per-timestep particle-robot update.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * it; the product library never links or calls it.
 *
 * Parity status: the reference ships NO tests, golden vectors or CPU path, so
 * the oracle cannot be pinned against reference-owned vectors.  It is pinned instead against the
 * reference's own kernels compiled verbatim for sm_100a (oracle/_ref/libprs_refcuda.so, built by
 * oracle/Makefile from the sources where they lie in /root/reference) and run on a B200:
 * tests/golden/ holds outputs of that run with the generating script.  Integer outputs (hash,
 * sort order, cell tables) must be bit-equal; fp32 outputs agree to the tolerances written in
 * tests/ (the device code uses FMA contraction and approximate __powf, which a CPU cannot
 * reproduce bit-for-bit — SURVEY.md Q7).
 */
#ifndef PRS_ORACLE_H
#define PRS_ORACLE_H

#include "prs_simparams.h"

#ifdef __cplusplus
extern "C" {
#endif

/* XORWOW generator state, same 48-byte layout as the toolkit's curandStateXORWOW
 * (curand_kernel.h) so device states can be compared word for word. */
typedef struct {
  unsigned int d, v[5];
  int boxmuller_flag;
  int boxmuller_flag_double;
  float boxmuller_extra;
  int pad_;
  double boxmuller_extra_double;
} prso_rng_state;

typedef struct PrsOracle PrsOracle;

void prso_set_threads(int n);      /* OpenMP threads used by the loops over robots (default 1) */
int prso_get_max_threads(void);

/* ---- glibc rand() restatement (TYPE_3 additive feedback), independent of the process-global
 * stream; tests check it against the real srand()/rand(). */
typedef struct { int r[34]; int f, b; } prso_glibc_rand;
void prso_srand(prso_glibc_rand *g, unsigned seed);
int prso_rand(prso_glibc_rand *g);

/* ---- per-kernel restatements on raw arrays (cited in prs_oracle.cpp) ---- */
void prso_calc_hash(const SimParams *p, const float *pos, unsigned *hash, unsigned *index, int n);
void prso_sort_pairs(unsigned *hash, unsigned *index, int n); /* stable by key */
void prso_reorder_find_cell_start(const SimParams *p, unsigned *cellStart, unsigned *cellEnd,
                                  float *sortedPos, float *sortedVel, float *sortedRad,
                                  const unsigned *hash, const unsigned *index, const float *pos,
                                  const float *vel, const float *rad, int n, unsigned numCells);
void prso_collide(const SimParams *p, float *newVel, float *absForce_a, float *absForce_r,
                  const float *sortedPos, const float *sortedVel, const float *sortedRad,
                  const unsigned *index, const unsigned *cellStart, const unsigned *cellEnd, int n,
                  float dt);
void prso_integrate(const SimParams *p, float *pos, float *vel, const float *rad, float dt, int n,
                    float world_half);
void prso_update_rad(const SimParams *p, const float *absForce_a, const float *absForce_r,
                     float *rad, const float *phase, float time, float dt, const int *dead, int n);
void prso_update_phase(const SimParams *p, const float *pos, float *phase, float spacing,
                       float min_d, int n);
float prso_min_light_distance(const SimParams *p, const float *pos, int n);
void prso_curand_setup(prso_rng_state *st, unsigned seed, int n);
void prso_add_normal_noise(prso_rng_state *st, float *val, float std, int n);

/* ---- whole simulation object: Particlebot::{ctor,reset,update} ---- */
PrsOracle *prso_create(const SimParams *p, float world_half);
void prso_destroy(PrsOracle *o);
void prso_srand_sim(PrsOracle *o, unsigned seed);   /* main.cpp:929 */
void prso_reset(PrsOracle *o);                      /* particlebot.cpp:485-801 */
void prso_update(PrsOracle *o, float dt, float sort_interval); /* particlebot.cpp:170-300 */
float prso_time(const PrsOracle *o);
/* which: 0 pos(2n) 1 vel(2n) 2 rad 3 phase 4 absForce_a 5 absForce_r 6 dead(int) 7 hash 8 index
 * 9 cellStart 10 cellEnd 11 sortedPos 12 sortedVel 13 sortedRad 14 rng state */
void *prso_array(PrsOracle *o, int which);
/* gate(T): time - T*floor(time/T) < dt in fp32 (particlebot.cpp:207,212,256) */
int prso_gate(float time, float interval, float dt);

#ifdef __cplusplus
}
#endif
#endif
