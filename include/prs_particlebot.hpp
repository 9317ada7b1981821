/*
 * prs_particlebot.hpp — headless `class Particlebot`, the simulation object of the reference
 * (particlebot.h:14-140, particlebot.cpp) on top of the C-ABI of prs_cabi.h.
 *
 * Same public surface as the reference class (constructor from SimParams, update, reset,
 * getArray/setArray, dumpParticlebot, loadFromFile, world getters).  What differs:
 *   - no OpenGL: positions/radii/colours live in plain device buffers of the sizes the
 *     reference gives its VBOs (N + centroid_steps + 1 entries), the VBO getters return 0 and the
 *     getCuda*VBO getters return the device pointers;
 *   - update() does not exit(0) at max_time, it returns true and the caller stops;
 *   - the kernels are reached through a table of function pointers with the reference's exact
 *     signatures (PrsBackend), so the same host logic drives (a) this library's fused path,
 *     (b) this library's per-call entry points in the reference's call order, or (c) ANY library
 *     exporting the reference ABI — tests and bench.py use (c) to run the reference's own
 *     kernels (oracle/_ref/libprs_refcuda.so) under identical host logic;
 *   - glibc rand() is restated per object (PrsRand) so that several simulations in one process
 *     have independent, reproducible streams (srand(seed) == Particlebot::srand(seed)).
 */
#ifndef PRS_PARTICLEBOT_HPP
#define PRS_PARTICLEBOT_HPP

#include <stdio.h>
#include "prs_cabi.h"

#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
struct PrsBackend {
  void *dl_handle;
  void (*allocateArray)(void **, size_t);
  void (*freeArray)(void *);
  void (*threadSync)(void);
  void (*copyArrayToDevice)(void *, const void *, int, int);
  void (*copyArrayFromDevice)(void *, const void *, struct cudaGraphicsResource **, int);
  void (*setParameters)(SimParams *);
  void (*integrateSystem)(float *, float *, float *, float, unsigned, float);
  void (*calcHash)(unsigned *, unsigned *, float *, int);
  void (*sortParticlebots)(unsigned *, unsigned *, unsigned);
  void (*reorderDataAndFindCellStart)(unsigned *, unsigned *, float *, float *, float *, unsigned *, unsigned *,
                                      float *, float *, float *, unsigned, unsigned);
  void (*collide)(float *, float *, float *, float *, float *, float *, unsigned *, unsigned *, unsigned *, unsigned,
                  unsigned, float);
  void (*updateRad_light_wave)(float *, float *, float *, float *, float *, float, float, int *, int);
  void (*updatePhase)(float *, float *, float, float, float, int);
  void (*curand_setup)(struct curandStateXORWOW *, int);
  void (*add_normal_noise)(struct curandStateXORWOW *, float *, float, int);
  void (*calcCOG)(float *, float *, float *, int, float, int, float);
};

/* glibc rand()/srand() (TYPE_3 additive feedback) as an object */
class PrsRand {
 public:
  PrsRand() { seed(1); }
  void seed(unsigned s);
  int next();
  void getState(int out[33]) const { for (int i = 0; i < 31; i++) out[i] = r_[i]; out[31] = f_; out[32] = b_; }
  void setState(const int in[33]) { for (int i = 0; i < 31; i++) r_[i] = in[i]; f_ = in[31]; b_ = in[32]; }
 private:
  int r_[31];
  int f_, b_;
};

enum PrsBackendKind { PRS_BACKEND_FUSED = 0, PRS_BACKEND_PERCALL = 1, PRS_BACKEND_EXTERNAL = 2 };

class Particlebot {
 public:
  Particlebot(SimParams params, float world_half = 64.0f, int backend = PRS_BACKEND_FUSED,
              const char *external_library = 0);
  ~Particlebot();

  bool update(float deltaTime, float sort_interval); /* true once time > max_time */
  /* update() for a caller that keeps the state on the host (see prs_sim_update_host in prs_cabi.h) */
  bool updateHost(const float *pos_in, const float *vel_in, const float *rad_in, float *pos_out, float *vel_out,
                  float *rad_out, float deltaTime, float sort_interval);
  void reset();
  void srand(unsigned seed) { rng_.seed(seed); }
  const PrsRand &randStream() const { return rng_; } /* where the glibc stream stands (the slab launcher continues it on every rank) */
  /* synthetic swarms: nx*ny hex lattice centred on the origin */
  void initHexBlock(unsigned nx, unsigned ny, float pitch, float jitter, unsigned seed);

  float *getArray(ParticlebotArray array);
  void setArray(ParticlebotArray array, const float *data, int start, int count);

  unsigned int getCurrentReadBuffer() const { return 0; }
  unsigned int getColorBuffer() const { return 0; }
  unsigned int getRadBuffer() const { return 0; }
  void *getCudaPosVBO() const { return (void *)dPos; }
  void *getCudaColorVBO() const { return (void *)dCol; } /* allocated by the first renderFrame */
  void *getCudaRadVBO() const { return (void *)dRad; }

  void dumpParticlebot(unsigned start, unsigned count, FILE *fp, float dump_interval, unsigned testing, float light_x,
                       float light_y);
  void loadFromFile(unsigned start, unsigned count, FILE *fp, float dump_interval);
  /* Full checkpoint (the reference's loadFromFile restores only time, positions, velocities
   * and radii from 6-digit CSV text).  Binary image of everything the next update() depends on: time, the
   * host rand() stream, positions (with the centroid trail), velocities, radii, phases, both |force| sums,
   * dead flags, the cuRAND states of the phase noise, and the frozen sort order (hash, index) that steps
   * between two sorts walk.  A restored simulation continues bit for bit.  Return 0, or -1 on I/O or
   * shape mismatch (nothing is modified then). */
  int saveCheckpoint(FILE *fp);
  int loadCheckpoint(FILE *fp);

  void getWorldOrigin(float *xy) const { xy[0] = params.worldOrigin.x; xy[1] = params.worldOrigin.y; }
  void getCellSize(float *xy) const { xy[0] = params.cellSize.x; xy[1] = params.cellSize.y; }

  /* additions */
  /* One displayed frame without OpenGL: colours of the current state (updateCol, particlebot.cpp:254; the
   * trail entries keep the reference's red, :125-135), the scene and the discs by prs_render_frame; returns host memory
   * owned by the object (valid until the next call): width * height * 3 bytes, top-down B, G, R rows. */
  const unsigned char *renderFrame(const prs_view &view);
  float getTime() const { return time; }
  void sync();
  void *devicePtr(int which);
  size_t arrayBytes(int which) const;
  const SimParams &getParams() const { return params; }

 protected:
  void _initialize();
  void _finalize();
  void uploadInitialState();

  bool sorted_once_ = false;
  float *hPos, *hVel, *hRad;
  int *hDead;
  float *hphase, *hfreq;

  float *dPos, *dVel, *dRad, *dAbsForce_a, *dAbsForce_r;
  int *dDead;
  struct curandStateXORWOW *dState;
  float *dfreq, *dphase;
  float *dSortedPos, *tempPos1, *tempPos2, *dSortedVel, *dSortedRad;
  unsigned *dGridParticleHash, *dGridParticleIndex, *dCellStart, *dCellEnd;
  float *dMinD; /* device scalar for the fused path's light-distance reduction */
  float *dSortedPR; /* fused path: packed float4 sorted copy (see prs_step_buffers) */
  /* headless frames: colour buffer (N + centroid_steps + 1 float4, the reference's colorVBO), device image + key planes,
   * pinned host image; all allocated by the first renderFrame */
  float *dCol = 0;
  unsigned char *dFrame = 0, *hFrame = 0;
  unsigned *dFrameKeys = 0;
  size_t framePixels_ = 0;

  float time;
  SimParams params;
  float obstacles_[7][PRS_MAX_OBSTACLES];
  float world_half_;
  int backend_kind_;
  PrsBackend be_;
  PrsRand rng_;
  unsigned configSizeX_;
};
#pragma GCC visibility pop

#endif
