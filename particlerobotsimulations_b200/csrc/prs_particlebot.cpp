/*
 * prs_particlebot.cpp — host logic of the headless `class Particlebot` (include/prs_particlebot.hpp)
 * and its C wrappers (prs_sim_* of include/prs_cabi.h).  Host-only C++; every device operation
 * goes through the reference-shaped C-ABI (PrsBackend) or this library's fused entry points.
 *
 * Results reproduced (reference particlebot.cpp): _initialize :77-166, update :170-300 (gate
 * arithmetic in fp32, dead draw, host light-distance loop in the per-call modes), reset :485-801
 * (CONFIG_RANDOM aggregation placement + CONFIG_GRID/HEX/LINE generators), dumpParticlebot
 * :303-367 (byte-compatible CSV), loadFromFile :369-411, getArray/setArray :803-867.
 */
#include "prs_particlebot.hpp"

#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

/* ------------------------------------------------------------------------------------------ */
void PrsRand::seed(unsigned s) {
  /* glibc srandom_r for TYPE_3: r[0] = seed, r[i] = 16807 * r[i-1] mod (2^31 - 1) by Schrage's
   * method, front pointer 3 ahead of the rear pointer, 310 values discarded */
  if (s == 0) s = 1;
  r_[0] = (int)s;
  long prev = r_[0];
  for (int i = 1; i < 31; i++) {
    long q = prev / 127773, rem = prev % 127773;
    long w = 16807 * rem - 2836 * q;
    if (w < 0) w += 2147483647;
    r_[i] = (int)w;
    prev = w;
  }
  f_ = 3;
  b_ = 0;
  for (int i = 0; i < 310; i++) next();
}
int PrsRand::next() {
  const uint32_t sum = (uint32_t)r_[f_] + (uint32_t)r_[b_];
  r_[f_] = (int)sum;
  if (++f_ == 31) f_ = 0;
  if (++b_ == 31) b_ = 0;
  return (int)(sum >> 1);
}

/* ------------------------------------------------------------------------------------------ */
static void native_backend(PrsBackend *b) {
  b->dl_handle = 0;
  b->allocateArray = allocateArray;
  b->freeArray = freeArray;
  b->threadSync = threadSync;
  b->copyArrayToDevice = copyArrayToDevice;
  b->copyArrayFromDevice = copyArrayFromDevice;
  b->setParameters = setParameters;
  b->integrateSystem = integrateSystem;
  b->calcHash = calcHash;
  b->sortParticlebots = sortParticlebots;
  b->reorderDataAndFindCellStart = reorderDataAndFindCellStart;
  b->collide = collide;
  b->updateRad_light_wave = updateRad_light_wave;
  b->updatePhase = updatePhase;
  b->curand_setup = curand_setup;
  b->add_normal_noise = add_normal_noise;
  b->calcCOG = calcCOG;
}

template <class F>
static void bind(void *h, const char *name, F *out) {
  void *s = dlsym(h, name);
  if (!s) {
    fprintf(stderr, "Particlebot: external backend lacks symbol %s\n", name);
    exit(EXIT_FAILURE);
  }
  *out = (F)s;
}

static void external_backend(PrsBackend *b, const char *path) {
  void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!h) {
    fprintf(stderr, "Particlebot: cannot load external backend %s: %s\n", path ? path : "(null)", dlerror());
    exit(EXIT_FAILURE);
  }
  b->dl_handle = h;
  bind(h, "allocateArray", &b->allocateArray);
  bind(h, "freeArray", &b->freeArray);
  bind(h, "threadSync", &b->threadSync);
  bind(h, "copyArrayToDevice", &b->copyArrayToDevice);
  bind(h, "copyArrayFromDevice", &b->copyArrayFromDevice);
  bind(h, "setParameters", &b->setParameters);
  bind(h, "integrateSystem", &b->integrateSystem);
  bind(h, "calcHash", &b->calcHash);
  bind(h, "sortParticlebots", &b->sortParticlebots);
  bind(h, "reorderDataAndFindCellStart", &b->reorderDataAndFindCellStart);
  bind(h, "collide", &b->collide);
  bind(h, "updateRad_light_wave", &b->updateRad_light_wave);
  bind(h, "updatePhase", &b->updatePhase);
  bind(h, "curand_setup", &b->curand_setup);
  bind(h, "add_normal_noise", &b->add_normal_noise);
  bind(h, "calcCOG", &b->calcCOG);
}

/* ------------------------------------------------------------------------------------------ */
Particlebot::Particlebot(SimParams simparams, float world_half, int backend, const char *external_library)
    : hPos(0), hVel(0), dPos(0), dVel(0), time(0) {
  params = simparams;
  /* keep private copies of the obstacle lists: the caller's arrays may go away */
  float *src[7] = {simparams.x1obs, simparams.x2obs, simparams.y1obs, simparams.y2obs,
                   simparams.x_cir_obs, simparams.y_cir_obs, simparams.r_cir_obs};
  float **dst[7] = {&params.x1obs, &params.x2obs, &params.y1obs, &params.y2obs,
                    &params.x_cir_obs, &params.y_cir_obs, &params.r_cir_obs};
  for (int a = 0; a < 7; a++) {
    const int cnt = a < 4 ? simparams.nobstacles : simparams.n_cir_obstacles;
    for (int i = 0; i < PRS_MAX_OBSTACLES; i++) obstacles_[a][i] = (src[a] && i < cnt) ? src[a][i] : 0.0f;
    *dst[a] = obstacles_[a];
  }
  world_half_ = world_half;
  backend_kind_ = backend;
  configSizeX_ = 0;
  if (backend == PRS_BACKEND_EXTERNAL) external_backend(&be_, external_library);
  else native_backend(&be_);
  _initialize();
}

Particlebot::~Particlebot() { _finalize(); }

void Particlebot::_initialize() {
  const size_t n = params.nCells, trail = (size_t)params.centroid_steps + 1;
  hPos = new float[(n + trail) * 2]();
  hVel = new float[n * 2 + 2]();
  hRad = new float[n + trail]();
  hDead = new int[n + 1]();
  hphase = new float[n + 1]();
  hfreq = new float[n + 1]();

  const size_t vec = sizeof(float) * 2 * n;
  be_.allocateArray((void **)&dPos, vec + sizeof(float) * 2 * trail);
  be_.allocateArray((void **)&dRad, sizeof(float) * (n + trail));
  be_.allocateArray((void **)&dVel, vec);
  be_.allocateArray((void **)&dSortedPos, vec);
  be_.allocateArray((void **)&tempPos1, vec + 64);
  be_.allocateArray((void **)&tempPos2, vec + 64);
  be_.allocateArray((void **)&dSortedVel, vec);
  be_.allocateArray((void **)&dSortedRad, sizeof(float) * n);
  be_.allocateArray((void **)&dphase, sizeof(float) * n);
  be_.allocateArray((void **)&dfreq, sizeof(float) * n);
  be_.allocateArray((void **)&dAbsForce_a, sizeof(float) * n);
  be_.allocateArray((void **)&dAbsForce_r, sizeof(float) * n);
  be_.allocateArray((void **)&dGridParticleHash, n * sizeof(unsigned));
  be_.allocateArray((void **)&dGridParticleIndex, n * sizeof(unsigned));
  be_.allocateArray((void **)&dCellStart, params.numCells * sizeof(unsigned));
  be_.allocateArray((void **)&dCellEnd, params.numCells * sizeof(unsigned));
  be_.allocateArray((void **)&dDead, sizeof(int) * n);
  be_.allocateArray((void **)&dState, (size_t)48 * n);
  be_.allocateArray((void **)&dMinD, 64);
  dSortedPR = 0;
  if (backend_kind_ == PRS_BACKEND_FUSED) be_.allocateArray((void **)&dSortedPR, sizeof(float) * 4 * n + 16);

  /* The reference reads absForce_a/r (and the cell tables) before anything wrote them and relies
   * on fresh allocations being zero; make that explicit. */
  {
    std::vector<char> zeros(std::max<size_t>(sizeof(float) * n, 4), 0);
    const int nb = (int)(sizeof(float) * n);
    if (nb) {
      be_.copyArrayToDevice(dAbsForce_a, zeros.data(), 0, nb);
      be_.copyArrayToDevice(dAbsForce_r, zeros.data(), 0, nb);
      be_.copyArrayToDevice(dGridParticleHash, zeros.data(), 0, nb);
      be_.copyArrayToDevice(dGridParticleIndex, zeros.data(), 0, nb);
    }
    std::vector<char> zc((size_t)params.numCells * sizeof(unsigned), 0);
    be_.copyArrayToDevice(dCellStart, zc.data(), 0, (int)zc.size());
    be_.copyArrayToDevice(dCellEnd, zc.data(), 0, (int)zc.size());
  }

  if (backend_kind_ != PRS_BACKEND_EXTERNAL) prs_set_world_half_extent(world_half_);
  be_.setParameters(&params);
  be_.curand_setup(dState, (int)params.nCells);
}

void Particlebot::_finalize() {
  be_.threadSync();
  delete[] hPos; delete[] hVel; delete[] hRad; delete[] hDead; delete[] hphase; delete[] hfreq;
  void *bufs[] = {dPos, dRad, dVel, dSortedPos, tempPos1, tempPos2, dSortedVel, dSortedRad, dphase, dfreq,
                  dAbsForce_a, dAbsForce_r, dGridParticleHash, dGridParticleIndex, dCellStart, dCellEnd, dDead,
                  dState, dMinD};
  for (void *b : bufs) be_.freeArray(b);
  if (dSortedPR) be_.freeArray(dSortedPR);
  if (dCol) be_.freeArray(dCol);
  if (dFrame) be_.freeArray(dFrame);
  if (dFrameKeys) be_.freeArray(dFrameKeys);
  delete[] hFrame;
  if (be_.dl_handle) dlclose(be_.dl_handle);
}

/* ------------------------------------------------------------------------------------------
 * headless frame (the reference's display(): main.cpp:352-476 -> renderer->display() -> Postprocess())
 * ------------------------------------------------------------------------------------------ */
const unsigned char *Particlebot::renderFrame(const prs_view &view) {
  const size_t n = params.nCells, trail = (size_t)params.centroid_steps + 1;
  if (!dCol) {
    /* colorVBO of the reference (particlebot.cpp:110-135): white robots until updateCol runs, a red trail */
    be_.allocateArray((void **)&dCol, (n + trail) * 4 * sizeof(float));
    std::vector<float> c((n + trail) * 4, 1.0f);
    for (size_t i = n; i < n + trail; i++) { c[4 * i + 1] = 0.0f; c[4 * i + 2] = 0.0f; c[4 * i + 3] = (i + 1 < n + trail) ? 0.8f : 1.0f; }
    size_t done = 0;
    const size_t bytes = c.size() * sizeof(float);
    while (done < bytes) {
      const size_t chunk = std::min<size_t>(bytes - done, (size_t)1 << 30);
      be_.copyArrayToDevice((char *)dCol + done, (const char *)c.data() + done, 0, (int)chunk);
      done += chunk;
    }
  }
  if (backend_kind_ == PRS_BACKEND_EXTERNAL && !framePixels_) {
    /* the step kernels are somebody else's: the frame kernels read THIS library's parameter block */
    setParameters(&params);
    prs_set_world_half_extent(world_half_);
  }
  const size_t npix = (size_t)view.width * view.height;
  if (npix != framePixels_) {
    if (dFrame) { be_.freeArray(dFrame); be_.freeArray(dFrameKeys); delete[] hFrame; }
    be_.allocateArray((void **)&dFrame, npix * 3);
    be_.allocateArray((void **)&dFrameKeys, npix * 2 * sizeof(unsigned));
    hFrame = new unsigned char[npix * 3];
    framePixels_ = npix;
  }
  if (n) updateCol(dRad, dCol, (int)n, dPos, dphase, dDead);
  prs_render_frame(dFrame, dFrameKeys, &view, dPos, dRad, dCol, (unsigned)(n + params.centroid_steps));
  size_t done = 0;
  while (done < npix * 3) {
    const size_t chunk = std::min<size_t>(npix * 3 - done, (size_t)1 << 30);
    copyArrayFromDevice(hFrame + done, dFrame + done, 0, (int)chunk);
    done += chunk;
  }
  return hFrame;
}

void Particlebot::sync() { be_.threadSync(); }

static inline bool gate(float time, float interval, float dt) {
  /* fp32 on purpose: `time` is an fp32 accumulator and the gates drift with it */
  return time - interval * floorf(time / interval) < dt;
}

bool Particlebot::update(float deltaTime, float sort_interval) {
  if (time > params.max_time) return true;
  const int n = (int)params.nCells;

  if (time >= params.time_to_dead && time < params.time_to_dead + deltaTime) {
    /* dead-cell draw: nDead distinct robots, rand() % remaining with erase */
    std::vector<int> alive(n);
    for (int i = 0; i < n; i++) alive[i] = i;
    for (int drawn = 0; drawn < params.nDead; drawn++) {
      const size_t pick = (size_t)((unsigned long)rng_.next() % alive.size());
      hDead[alive[pick]] = 1;
      alive.erase(alive.begin() + pick);
    }
    if (n) be_.copyArrayToDevice(dDead, hDead, 0, n * (int)sizeof(int));
  }

  /* centroid into the render trail pos[n + k] (particlebot.cpp:207-209): every centroid_int of simulated time */
  if (n && gate(time, params.centroid_int, deltaTime))
    be_.calcCOG(dPos, tempPos1, tempPos2, n, time, params.centroid_steps, params.centroid_int);

  const bool phase_step = params.control == LIGHT_WAVE && gate(time, params.phase_update_interval, deltaTime);
  /* the first update always hashes and sorts: at time 0 the gate fires anyway; after loadFromFile (time > 0)
   * the reference would build its table from never-written hash/index arrays */
  const bool sort_step = gate(time, sort_interval, deltaTime) || !sorted_once_;
  sorted_once_ = true;
  const float spacing = 2.0f * params.min_radius;

  if (backend_kind_ == PRS_BACKEND_FUSED) {
    if (phase_step) {
      prs_min_light_distance(dPos, n, dMinD);
      prs_update_phase_dev(dPos, dphase, spacing, dMinD, n);
      if (params.phase_std) add_normal_noise(dState, dphase, params.phase_std, n);
    }
    prs_step_buffers b;
    b.pos = dPos; b.vel = dVel; b.rad = dRad; b.phase = dphase; b.absForce_a = dAbsForce_a; b.absForce_r = dAbsForce_r;
    b.dead = dDead; b.hash = dGridParticleHash; b.index = dGridParticleIndex; b.cellStart = dCellStart;
    b.cellEnd = dCellEnd; b.sortedPos = dSortedPos; b.sortedVel = dSortedVel; b.sortedRad = dSortedRad;
    b.nCells = params.nCells; b.numCells = params.numCells; b.sortedPR = dSortedPR;
    prs_fused_step(&b, time, deltaTime, sort_step ? 1 : 0);
  } else {
    /* the reference's own call sequence */
    if (params.control == LIGHT_WAVE) {
      if (phase_step) {
        be_.copyArrayFromDevice(hPos, dPos, 0, (int)(sizeof(float) * 2 * n));
        float min_d = 0, max_d = 0;
        for (int i = 0; i < n; i++) {
          const float d = powf(powf(params.light_x - hPos[i * 2], 2) + powf(params.light_y - hPos[i * 2 + 1], 2), 0.5f);
          if (i == 0) { max_d = d; min_d = d; }
          else { min_d = (min_d < d ? min_d : d); max_d = (max_d > d ? max_d : d); }
        }
        be_.updatePhase(dPos, dphase, spacing, max_d, min_d, n);
        if (params.phase_std) be_.add_normal_noise(dState, dphase, params.phase_std, n);
      }
      if (time >= 0) be_.updateRad_light_wave(dPos, dAbsForce_a, dAbsForce_r, dRad, dphase, time, deltaTime, dDead, n);
    }
    be_.integrateSystem(dPos, dVel, dRad, deltaTime, params.nCells, time);
    if (sort_step) {
      be_.calcHash(dGridParticleHash, dGridParticleIndex, dPos, n);
      be_.sortParticlebots(dGridParticleHash, dGridParticleIndex, params.nCells);
    }
    be_.reorderDataAndFindCellStart(dCellStart, dCellEnd, dSortedPos, dSortedVel, dSortedRad, dGridParticleHash,
                                    dGridParticleIndex, dPos, dVel, dRad, params.nCells, params.numCells);
    be_.collide(dVel, dAbsForce_a, dAbsForce_r, dSortedPos, dSortedVel, dSortedRad, dGridParticleIndex, dCellStart,
                dCellEnd, params.nCells, params.numCells, deltaTime);
  }
  time = time + deltaTime;
  return false;
}

bool Particlebot::updateHost(const float *pos_in, const float *vel_in, const float *rad_in, float *pos_out,
                             float *vel_out, float *rad_out, float deltaTime, float sort_interval) {
  const size_t n = params.nCells;
  if (backend_kind_ != PRS_BACKEND_FUSED) {
    /* per-call backends: the reference's own blocking copies */
    setArray(POSITION, pos_in, 0, (int)n);
    setArray(VELOCITY, vel_in, 0, (int)n);
    setArray(RADII, rad_in, 0, (int)n);
    const bool done = update(deltaTime, sort_interval);
    be_.copyArrayFromDevice(pos_out, dPos, 0, (int)(n * 8));
    be_.copyArrayFromDevice(vel_out, dVel, 0, (int)(n * 8));
    be_.copyArrayFromDevice(rad_out, dRad, 0, (int)(n * 4));
    return done;
  }
  /* Steps on which update() itself reads the state before the fused step (phase update: light-distance reduction over the
   * positions; centroid trail; dead-cell draw) or that end the run: everything up first.  All other steps hand the buffers
   * to the fused step, which pipelines upload, K1 and the way back of positions and radii where its route allows
   * (prs_host_step_plan). */
  const bool reads_first = (params.control == LIGHT_WAVE && gate(time, params.phase_update_interval, deltaTime)) ||
                           gate(time, params.centroid_int, deltaTime) ||
                           (time >= params.time_to_dead && time < params.time_to_dead + deltaTime) || time > params.max_time;
  if (reads_first) {
    prs_h2d_async(dPos, pos_in, n * 8);
    prs_h2d_async(dVel, vel_in, n * 8);
    prs_h2d_async(dRad, rad_in, n * 4);
    prs_arm_k1_event(1);
    const bool done = update(deltaTime, sort_interval);
    prs_arm_k1_event(0);
    /* positions and radii are final once K1 ran: their way back overlaps sort, reorder and collide */
    prs_d2h_async(pos_out, dPos, n * 8, done ? 0 : 1);
    prs_d2h_async(rad_out, dRad, n * 4, done ? 0 : 1);
    prs_d2h_async(vel_out, dVel, n * 8, 0);
    prs_host_step_sync();
    return done;
  }
  prs_host_step_plan(pos_in, vel_in, rad_in, pos_out, rad_out);
  const bool done = update(deltaTime, sort_interval);
  prs_d2h_async(vel_out, dVel, n * 8, 0);
  prs_host_step_sync();
  return done;
}

/* ------------------------------------------------------------------------------------------
 * initial state
 * ------------------------------------------------------------------------------------------ */
static inline float host_norm(float x, float y) { return powf(powf(x, 2.0f) + powf(y, 2.0f), 0.5f); }

namespace {
/* occupancy grid used by the aggregation placement: robots bucketed by cell; the reference
 * indexes it with unwrapped neighbours (undefined at the grid edge) — wrapped here */
struct Occupancy {
  int gx, gy;
  const SimParams *P;
  std::vector<std::vector<int>> cells;
  explicit Occupancy(const SimParams *p) : gx((int)p->gridSize.x), gy((int)p->gridSize.y), P(p), cells((size_t)gx * gy) {}
  void locate(float x, float y, int *cx, int *cy) const {
    *cx = ((int)floorf((x - P->worldOrigin.x) / P->cellSize.x)) & (gx - 1);
    *cy = ((int)floorf((y - P->worldOrigin.y) / P->cellSize.y)) & (gy - 1);
  }
  std::vector<int> &at(int cx, int cy) { return cells[(size_t)(cx & (gx - 1)) * gy + (cy & (gy - 1))]; }
  void insert(float x, float y, int id) { int cx, cy; locate(x, y, &cx, &cy); at(cx, cy).push_back(id); }
  /* does a disc of radius min_radius at (x,y) overlap any registered disc in the 3x3 cells? */
  bool overlaps(float x, float y, const float *pos) {
    int cx, cy;
    locate(x, y, &cx, &cy);
    for (int ax = cx - 1; ax <= cx + 1; ax++)
      for (int ay = cy - 1; ay <= cy + 1; ay++)
        for (int id : at(ax, ay))
          if (host_norm(x - pos[2 * id], y - pos[2 * id + 1]) < 2 * 1.0 * P->min_radius) return true;
    return false;
  }
};
}  // namespace

void Particlebot::reset() {
  time = 0;
  const int n = (int)params.nCells;
  const float PI_F = 3.141592654f;
  switch (params.config) {
    case CONFIG_GRID:
    case CONFIG_LINE: {
      /* initGrid (:413-436): one row along x, jittered; y = 0 */
      const float jitter = params.config == CONFIG_GRID ? params.max_radius * 0.01f : 0.0f;
      const unsigned sx = params.config == CONFIG_GRID ? (unsigned)ceilf(powf((float)n, 0.5f)) : (unsigned)n;
      const unsigned sy = params.config == CONFIG_GRID ? sx : 1u;
      const float spacing = params.min_radius * 2.0f;
      const float xs = sx * spacing / 2.0f;
      configSizeX_ = sx;
      for (unsigned y = 0; y < sy; y++)
        for (unsigned x = 0; x < sx; x++) {
          const unsigned i = y * sx + x;
          if (i < (unsigned)n) {
            hPos[i * 2] = (spacing * x) + params.min_radius - xs + ((rng_.next() / (float)RAND_MAX) * 2.0f - 1.0f) * jitter;
            hPos[i * 2 + 1] = 0;
            hVel[i * 2] = hVel[i * 2 + 1] = 0.0f;
          }
        }
    } break;
    case CONFIG_HEX: {
      /* initHexGrid (:438-481): concentric hexagonal rings around the origin */
      const float spacing = params.min_radius * 2.0f;
      const float h = powf(3, 0.5f) * 0.5f;
      const float dir[7][2] = {{1.0f, 0.0f}, {0.5f, h}, {-0.5f, h}, {-1.0f, 0.0f}, {-0.5f, -h}, {0.5f, -h}, {1.0f, 0.0f}};
      int i = 0;
      if (n > 0) { hPos[0] = hPos[1] = 0.0f; hVel[0] = hVel[1] = 0.0f; i = 1; }
      int ring = 1;
      while (i < n) {
        for (int k = 0; k < 6 && i < n; k++)
          for (int j = 0; j < ring && i < n; j++) {
            hPos[i * 2] = dir[k][0] * (ring - j) * spacing + dir[k + 1][0] * spacing * j;
            hPos[i * 2 + 1] = dir[k][1] * (ring - j) * spacing + dir[k + 1][1] * spacing * j;
            hVel[i * 2] = hVel[i * 2 + 1] = 0.0f;
            i++;
          }
        ring++;
      }
      configSizeX_ = ring * 2;
    } break;
    case CONFIG_RANDOM:
    default: {
      /* random aggregation (:612-748).  The fixed 10-robot test layouts of the other enum values
       * are unreachable from a cfg (main.cpp:794-809) and map to this branch. */
      Occupancy occ(&params);
      configSizeX_ = (unsigned)ceilf(powf((float)n, 1.0f / 2.0f));
      const float step = (float)(2 * PI_F / 360.0 * 10.0); /* 10 degree pivot increment */
      const unsigned max_fail = 200;
      unsigned fails = 0;
      float leftmost = 9999999.0f;
      if (n > 0) {
        hPos[0] = 5.0f; hPos[1] = 0.0f; hVel[0] = hVel[1] = 0.0f;
        occ.insert(0.0f, 0.0f, 0); /* the reference files disc 0 under the ORIGIN's cell (:635-637) */
      }
      for (int i = 1; i < n; i++) {
        float x, y;
        if (i == 2) {
          /* third disc: perpendicular to the first pair, side chosen by one rand() */
          const int side = rng_.next() % 2;
          float ax = hPos[2] - hPos[0], ay = hPos[3] - hPos[1];
          const float l = host_norm(ax, ay);
          ay = ay / l; ax = ax / l;
          const float px = side ? ay : -ay, py = side ? -ax : ax;
          x = (hPos[2] + hPos[0]) / 2.0f + px * params.min_radius;
          y = (hPos[3] + hPos[1]) / 2.0f + py * params.min_radius;
          if (x < leftmost) leftmost = x;
        } else {
          float r = params.min_radius;
          while (true) {
            const unsigned anchor = (unsigned)rng_.next() % (unsigned)i;
            if (fails == max_fail) { fails = 0; r += params.min_radius; }
            float theta = 2 * (rng_.next() / (float)RAND_MAX) * PI_F;
            x = hPos[2 * anchor] + 2 * r * cosf(theta);
            y = hPos[2 * anchor + 1] + 2 * r * sinf(theta);
            if (occ.overlaps(x, y, hPos)) { fails++; continue; }
            /* free spot: pivot around the anchor until the next step would overlap */
            const float theta0 = theta;
            while (theta - theta0 < 2 * PI_F) {
              theta += step;
              x = hPos[2 * anchor] + 2 * r * cosf(theta);
              y = hPos[2 * anchor + 1] + 2 * r * sinf(theta);
              if (occ.overlaps(x, y, hPos)) { theta -= step; break; }
            }
            x = hPos[2 * anchor] + 2 * r * cosf(theta);
            y = hPos[2 * anchor + 1] + 2 * r * sinf(theta);
            break;
          }
          if (x < leftmost) leftmost = x;
          if (params.nDead == -1 && i == n - 1) { /* the transported object starts left of the swarm */
            x = leftmost - 1 * params.min_radius * params.radFactor - 2 * params.min_radius;
            y = 0;
          }
        }
        hPos[2 * i] = x; hPos[2 * i + 1] = y;
        hVel[2 * i] = hVel[2 * i + 1] = 0.0f;
        occ.insert(x, y, i);
      }
    } break;
  }
  if (!params.Nx) params.Nx = (int)configSizeX_; /* never reaches the device, as in the reference (:772) */
  uploadInitialState();
}

extern "C" void prs_hex_block_position(unsigned long long i, unsigned nx, unsigned ny, float pitch, float jitter, unsigned seed, float *xy) {
  const float row = pitch * 0.8660254037844386f;
  const float x0 = -0.5f * ((float)(nx - 1) * pitch + 0.5f * pitch), y0 = -0.5f * (float)(ny - 1) * row;
  auto mix = [](uint64_t z) {
    z += 0x9e3779b97f4a7c15ull; z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31);
  };
  const unsigned iy = (unsigned)(i / nx), ix = (unsigned)(i % nx);
  const uint64_t hsh = mix(((uint64_t)seed << 32) ^ (uint64_t)i);
  const float jx = ((float)(hsh & 0xffffff) / 8388608.0f - 1.0f) * jitter;
  const float jy = ((float)((hsh >> 24) & 0xffffff) / 8388608.0f - 1.0f) * jitter;
  xy[0] = x0 + (float)ix * pitch + ((iy & 1u) ? 0.5f * pitch : 0.0f) + jx;
  xy[1] = y0 + (float)iy * row + jy;
}

void Particlebot::initHexBlock(unsigned nx, unsigned ny, float pitch, float jitter, unsigned seed) {
  /* nx*ny robots on a hex lattice (row pitch = pitch*sqrt(3)/2, odd rows
   * shifted by pitch/2), centred on the origin, each coordinate jittered uniformly in +-jitter by
   * a counter hash of (seed, robot id) so that any subset can be generated independently. */
  time = 0;
  const size_t n = params.nCells;
  if (backend_kind_ != PRS_BACKEND_EXTERNAL) {
    /* own library: the lattice is generated by a kernel (same bits as the loop below, which stays for other libraries and
     * as the cross-check of the tests); only the render trail and the frequency array come from the host */
    const size_t trail = (size_t)params.centroid_steps + 1;
    for (int i = 0; i < params.centroid_steps; i++) { hRad[n + i] = params.centroid_radius; hPos[(n + i) * 2] = -5000.0f; hPos[(n + i) * 2 + 1] = 0.0f; }
    hRad[n + params.centroid_steps] = 0.0f;
    hPos[(n + params.centroid_steps) * 2] = hPos[(n + params.centroid_steps) * 2 + 1] = 0.0f;
    memset(hDead, 0, n * sizeof(int));
    prs_init_hex_block(dPos, dVel, dRad, dphase, dDead, (unsigned)n, nx, ny, pitch, jitter, seed, params.min_radius);
    be_.copyArrayToDevice(dPos, hPos + 2 * n, (int)(n * 2 * sizeof(float)), (int)(trail * 2 * sizeof(float)));
    be_.copyArrayToDevice(dRad, hRad + n, (int)(n * sizeof(float)), (int)(trail * sizeof(float)));
    if (n) setArray(FREQUENCY, hfreq, 0, (int)n);
    if (params.nDead == -1 && n) { /* the last robot is the transported object (reset(), particlebot.cpp:784-791) */
      const float r_obj = params.min_radius * params.radFactor;
      const int one = 1;
      hDead[n - 1] = 1;
      be_.copyArrayToDevice(dRad, &r_obj, (int)((n - 1) * sizeof(float)), (int)sizeof(float));
      be_.copyArrayToDevice(dDead, &one, (int)((n - 1) * sizeof(int)), (int)sizeof(int));
    }
    return;
  }
  for (size_t i = 0; i < n; i++) {
    prs_hex_block_position(i, nx, ny, pitch, jitter, seed, hPos + 2 * i);
    hVel[2 * i] = hVel[2 * i + 1] = 0.0f;
  }
  (void)ny;
  uploadInitialState();
}

void Particlebot::uploadInitialState() {
  /* radii, dead flags, phases of reset() (:775-800) and the uploads */
  const size_t n = params.nCells, trail = (size_t)params.centroid_steps + 1;
  for (int i = 0; i < params.centroid_steps; i++) {
    hRad[n + i] = params.centroid_radius;
    hPos[(n + i) * 2] = -5000.0f;
  }
  hRad[n + params.centroid_steps] = 0.0f;
  for (size_t i = 0; i < n; i++) {
    hRad[i] = params.min_radius;
    if (params.nDead == -1 && i == n - 1) { hRad[i] = params.min_radius * params.radFactor; hDead[i] = 1; }
    hphase[i] = 0;
  }
  if (n) be_.copyArrayToDevice(dDead, hDead, 0, (int)(n * sizeof(int)));
  setArray(RADII, hRad, 0, (int)(n + trail));
  setArray(PHASE, hphase, 0, (int)n);
  setArray(FREQUENCY, hfreq, 0, (int)n);
  setArray(POSITION, hPos, 0, (int)(n + trail));
  setArray(VELOCITY, hVel, 0, (int)n);
}

/* ------------------------------------------------------------------------------------------
 * array access, CSV
 * ------------------------------------------------------------------------------------------ */
float *Particlebot::getArray(ParticlebotArray array) {
  const int n = (int)params.nCells;
  switch (array) {
    default:
    case POSITION: be_.copyArrayFromDevice(hPos, dPos, 0, n * 2 * (int)sizeof(float)); return hPos;
    case VELOCITY: be_.copyArrayFromDevice(hVel, dVel, 0, n * 2 * (int)sizeof(float)); return hVel;
    case RADII: be_.copyArrayFromDevice(hRad, dRad, 0, n * (int)sizeof(float)); return hRad;
    case PHASE: be_.copyArrayFromDevice(hphase, dphase, 0, n * (int)sizeof(float)); return hphase;
    case FREQUENCY: be_.copyArrayFromDevice(hfreq, dfreq, 0, n * (int)sizeof(float)); return hfreq;
    case DEAD: be_.copyArrayFromDevice(hDead, dDead, 0, n * (int)sizeof(int)); return (float *)hDead;
  }
}

void Particlebot::setArray(ParticlebotArray array, const float *data, int start, int count) {
  if (count <= 0) return;
  const int f = (int)sizeof(float);
  switch (array) {
    default:
    case POSITION: be_.copyArrayToDevice(dPos, data, start * 2 * f, count * 2 * f); break;
    case VELOCITY: be_.copyArrayToDevice(dVel, data, start * 2 * f, count * 2 * f); break;
    case PHASE: be_.copyArrayToDevice(dphase, data, start * f, count * f); break;
    case FREQUENCY: be_.copyArrayToDevice(dfreq, data, start * f, count * f); break;
    case RADII: be_.copyArrayToDevice(dRad, data, start * f, count * f); break;
    case DEAD: be_.copyArrayToDevice(dDead, data, start * f, count * f); break;
  }
}

void Particlebot::dumpParticlebot(unsigned start, unsigned count, FILE *fp, float dump_interval, unsigned testing,
                                  float light_x, float light_y) {
  if (time - dump_interval * floorf(time / dump_interval) > 0.01f) return;
  be_.copyArrayFromDevice(hPos, dPos, 0, (int)(sizeof(float) * 2 * count));
  be_.copyArrayFromDevice(hVel, dVel, 0, (int)(sizeof(float) * 2 * count));
  be_.copyArrayFromDevice(hRad, dRad, 0, (int)(sizeof(float) * count));
  if (time == 0) {
    fprintf(fp, "Seed, %u\n", params.seed);
    fprintf(fp, "Time,");
    if (testing) {
      for (unsigned i = start; i < start + count; i++) fprintf(fp, "Particlebot_%d_xpos, Particlebot_%d_ypos,", i, i);
      for (unsigned i = start; i < start + count; i++) fprintf(fp, "Particlebot_%d_xvel, Particlebot_%d_yvel,", i, i);
      for (unsigned i = start; i < start + count; i++) fprintf(fp, "Particlebot_%d_rad,", i);
    }
    fprintf(fp, "Centroid X, Centroid Y, Distance");
    fprintf(fp, "\n");
  }
  fprintf(fp, "%f,", time);
  if (testing) {
    for (unsigned i = start; i < start + count; i++) fprintf(fp, "%f, %f,", hPos[i * 2 + 0], hPos[i * 2 + 1]);
    for (unsigned i = start; i < start + count; i++) fprintf(fp, "%f, %f,", hVel[i * 2 + 0], hVel[i * 2 + 1]);
    for (unsigned i = start; i < start + count; i++) fprintf(fp, "%f,", hRad[i]);
  }
  float sumX = 0.0f, sumY = 0.0f;
  for (unsigned i = start; i < start + count; i++) { sumX += hPos[i * 2 + 0]; sumY += hPos[i * 2 + 1]; }
  const float cx = sumX / (float)count, cy = sumY / (float)count;
  fprintf(fp, "%f, %f, %f,", cx, cy, powf(powf(cx - light_x, 2.0f) + powf(cy - light_y, 2.0f), 0.5f));
  fprintf(fp, "\n");
  printf("%f %f %f \n", time, cx, cy);
}

void Particlebot::loadFromFile(unsigned start, unsigned count, FILE *fp, float /*dump_interval*/) {
  /* resume from the LAST complete row of a testing=1 CSV: time, positions, velocities, radii */
  fseek(fp, 0, SEEK_SET);
  long last_line = 0, after_newline = 0, pos = 0;
  for (int c = fgetc(fp); c != EOF; c = fgetc(fp)) {
    pos++;
    if (c == '\n') { last_line = after_newline; after_newline = pos; }
  }
  fseek(fp, last_line, SEEK_SET);
  if (fscanf(fp, "%f,", &time) != 1) return;
  for (unsigned i = start; i < start + count; i++) if (fscanf(fp, "%f, %f,", &hPos[i * 2 + 0], &hPos[i * 2 + 1]) != 2) return;
  for (unsigned i = start; i < start + count; i++) if (fscanf(fp, "%f, %f,", &hVel[i * 2 + 0], &hVel[i * 2 + 1]) != 2) return;
  for (unsigned i = start; i < start + count; i++) if (fscanf(fp, "%f,", &hRad[i]) != 1) return;
  setArray(RADII, hRad, 0, (int)params.nCells);
  setArray(POSITION, hPos, 0, (int)params.nCells);
  setArray(VELOCITY, hVel, 0, (int)params.nCells);
  printf("Time = %f\n", time);
}

/* ------------------------------------------------------------------------------------------
 * checkpoint
 * ------------------------------------------------------------------------------------------ */
namespace {
struct CheckpointHeader {
  char magic[8];            /* "PRSCKPT1" */
  uint32_t version, nCells, numCells, trail;
  float time;
  uint32_t sorted_once;
  int32_t rand_state[33];
  uint32_t seed, reserved[5];
};
struct CheckpointArray { int which; size_t bytes; };
}  // namespace

int Particlebot::saveCheckpoint(FILE *fp) {
  const size_t n = params.nCells, trail = (size_t)params.centroid_steps + 1;
  CheckpointHeader h;
  memset(&h, 0, sizeof(h));
  memcpy(h.magic, "PRSCKPT1", 8);
  h.version = 1; h.nCells = (uint32_t)n; h.numCells = params.numCells; h.trail = (uint32_t)trail;
  h.time = time; h.sorted_once = sorted_once_ ? 1u : 0u; h.seed = (uint32_t)params.seed;
  rng_.getState(h.rand_state);
  if (fwrite(&h, sizeof(h), 1, fp) != 1) return -1;
  be_.threadSync();
  const struct { void *dev; size_t bytes; } arrays[] = {
      {dPos, (n + trail) * 8}, {dVel, n * 8}, {dRad, (n + trail) * 4}, {dphase, n * 4}, {dAbsForce_a, n * 4},
      {dAbsForce_r, n * 4}, {dDead, n * 4}, {dState, n * 48}, {dGridParticleHash, n * 4}, {dGridParticleIndex, n * 4}};
  std::vector<char> buf;
  for (const auto &a : arrays) {
    buf.resize(a.bytes);
    for (size_t done = 0; done < a.bytes;) { /* the reference ABI counts bytes in int */
      const size_t chunk = std::min<size_t>(a.bytes - done, (size_t)1 << 30);
      be_.copyArrayFromDevice(buf.data() + done, (const char *)a.dev + done, 0, (int)chunk);
      done += chunk;
    }
    if (a.bytes && fwrite(buf.data(), 1, a.bytes, fp) != a.bytes) return -1;
  }
  return fflush(fp) == 0 ? 0 : -1;
}

int Particlebot::loadCheckpoint(FILE *fp) {
  const size_t n = params.nCells, trail = (size_t)params.centroid_steps + 1;
  CheckpointHeader h;
  if (fread(&h, sizeof(h), 1, fp) != 1) return -1;
  if (memcmp(h.magic, "PRSCKPT1", 8) != 0 || h.version != 1 || h.nCells != n || h.numCells != params.numCells || h.trail != trail) {
    fprintf(stderr, "loadCheckpoint: not a checkpoint of this swarm (robots %u/%zu, cells %u/%u)\n", h.nCells, n, h.numCells, params.numCells);
    return -1;
  }
  const struct { void *dev; size_t bytes; } arrays[] = {
      {dPos, (n + trail) * 8}, {dVel, n * 8}, {dRad, (n + trail) * 4}, {dphase, n * 4}, {dAbsForce_a, n * 4},
      {dAbsForce_r, n * 4}, {dDead, n * 4}, {dState, n * 48}, {dGridParticleHash, n * 4}, {dGridParticleIndex, n * 4}};
  /* read everything before touching the simulation */
  std::vector<std::vector<char>> data;
  for (const auto &a : arrays) {
    data.emplace_back(a.bytes);
    if (a.bytes && fread(data.back().data(), 1, a.bytes, fp) != a.bytes) return -1;
  }
  for (size_t k = 0; k < data.size(); k++) {
    const size_t bytes = arrays[k].bytes;
    for (size_t done = 0; done < bytes;) {
      const size_t chunk = std::min<size_t>(bytes - done, (size_t)1 << 30);
      be_.copyArrayToDevice((char *)arrays[k].dev + done, data[k].data() + done, 0, (int)chunk);
      done += chunk;
    }
  }
  memcpy(hPos, data[0].data(), n * 8);
  memcpy(hVel, data[1].data(), n * 8);
  memcpy(hRad, data[2].data(), n * 4);
  memcpy(hphase, data[3].data(), n * 4);
  memcpy(hDead, data[6].data(), n * 4);
  time = h.time;
  sorted_once_ = h.sorted_once != 0; /* the frozen order came along: the next update only sorts if its gate says so */
  rng_.setState(h.rand_state);
  return 0;
}

void *Particlebot::devicePtr(int which) {
  switch (which) {
    case POSITION: return dPos;      case VELOCITY: return dVel;   case RADII: return dRad;
    case PHASE: return dphase;       case FREQUENCY: return dfreq; case DEAD: return dDead;
    case 100: return dAbsForce_a;    case 101: return dAbsForce_r; case 102: return dGridParticleHash;
    case 103: return dGridParticleIndex; case 104: return dCellStart; case 105: return dCellEnd;
    case 106: case 108:
      /* the fused path keeps the sorted copy packed; materialise the reference's arrays on demand */
      if (dSortedPR) prs_unpack_sorted(dSortedPR, dSortedPos, dSortedRad, params.nCells);
      return which == 106 ? (void *)dSortedPos : (void *)dSortedRad;
    case 107: return dSortedVel;
    case 109: return dState;
  }
  return 0;
}
size_t Particlebot::arrayBytes(int which) const {
  const size_t n = params.nCells;
  switch (which) {
    case POSITION: case VELOCITY: case 106: case 107: return n * 8;
    case 104: case 105: return (size_t)params.numCells * 4;
    case 109: return n * 48;
    default: return n * 4;
  }
}

/* ------------------------------------------------------------------------------------------
 * C wrappers
 * ------------------------------------------------------------------------------------------ */
struct prs_sim { Particlebot *bot; };

extern "C" {
prs_sim *prs_sim_create(const SimParams *p, float world_half, int backend, const char *ext_lib) {
  prs_sim *s = new prs_sim;
  s->bot = new Particlebot(*p, world_half, backend, ext_lib);
  return s;
}
void prs_sim_destroy(prs_sim *s) { if (s) { delete s->bot; delete s; } }
void prs_sim_srand(prs_sim *s, unsigned seed) { s->bot->srand(seed); }
void prs_sim_reset(prs_sim *s) { s->bot->reset(); }
void prs_sim_init_hex(prs_sim *s, unsigned nx, unsigned ny, float pitch, float jitter, unsigned seed) {
  s->bot->initHexBlock(nx, ny, pitch, jitter, seed);
}
int prs_sim_update(prs_sim *s, float dt, float sort_interval) { return s->bot->update(dt, sort_interval) ? 1 : 0; }
int prs_sim_update_host(prs_sim *s, const float *pos_in, const float *vel_in, const float *rad_in, float *pos_out,
                        float *vel_out, float *rad_out, float dt, float sort_interval) {
  return s->bot->updateHost(pos_in, vel_in, rad_in, pos_out, vel_out, rad_out, dt, sort_interval) ? 1 : 0;
}
float prs_sim_time(const prs_sim *s) { return s->bot->getTime(); }
void prs_sim_sync(prs_sim *s) { s->bot->sync(); }
void *prs_sim_device_ptr(prs_sim *s, int which) { return s->bot->devicePtr(which); }
void prs_sim_get(prs_sim *s, int which, void *host, size_t bytes) {
  void *d = s->bot->devicePtr(which);
  if (!d) { fprintf(stderr, "prs_sim_get: unknown array %d\n", which); exit(EXIT_FAILURE); }
  s->bot->sync();
  /* chunked so that the reference ABI's int byte counts never overflow */
  size_t done = 0;
  while (done < bytes) {
    const size_t chunk = std::min<size_t>(bytes - done, (size_t)1 << 30);
    copyArrayFromDevice((char *)host + done, (const char *)d + done, 0, (int)chunk);
    done += chunk;
  }
}
void prs_sim_set(prs_sim *s, int which, const void *host, size_t offset_bytes, size_t bytes) {
  void *d = s->bot->devicePtr(which);
  if (!d) { fprintf(stderr, "prs_sim_set: unknown array %d\n", which); exit(EXIT_FAILURE); }
  size_t done = 0;
  while (done < bytes) {
    const size_t chunk = std::min<size_t>(bytes - done, (size_t)1 << 30);
    copyArrayToDevice((char *)d + offset_bytes + done, (const char *)host + done, 0, (int)chunk);
    done += chunk;
  }
}
void prs_sim_dump(prs_sim *s, void *fp, float dump_interval, unsigned testing) {
  const SimParams &P = s->bot->getParams();
  s->bot->dumpParticlebot(0, P.nCells, (FILE *)fp, dump_interval, testing, P.light_x, P.light_y);
}
int prs_sim_checkpoint_save(prs_sim *s, const char *path) {
  FILE *fp = fopen(path, "wb");
  if (!fp) return -1;
  const int rc = s->bot->saveCheckpoint(fp);
  return fclose(fp) == 0 ? rc : -1;
}
int prs_sim_checkpoint_load(prs_sim *s, const char *path) {
  FILE *fp = fopen(path, "rb");
  if (!fp) return -1;
  const int rc = s->bot->loadCheckpoint(fp);
  fclose(fp);
  return rc;
}
const unsigned char *prs_sim_render_frame(prs_sim *s, const prs_view *view) { return s->bot->renderFrame(*view); }
void prs_sim_load(prs_sim *s, void *fp) {
  const SimParams &P = s->bot->getParams();
  s->bot->loadFromFile(0, P.nCells, (FILE *)fp, 0.0f);
}
}
