/*
 * refhost_shim.cpp — the reference's OWN host class (particlebot.cpp: constructor, reset(), update(), setArray /
 * getArray), compiled verbatim into oracle/_ref/libprs_refhost.so, plus a small C interface to drive it headless.
 * TEST INFRASTRUCTURE ONLY: it pins the host logic of this repository (update order, fp32 gates, dead-cell draw,
 * CONFIG_RANDOM placement: csrc/prs_particlebot.cpp, oracle/prs_oracle.cpp) against the reference itself
 * (particlebot.cpp:170-300, 485-801).  No reference source is copied: the file is textually included from the
 * directory given with -I.  OpenGL is replaced by oracle/gl_stub (buffer objects = device allocations).
 */
#include "particlebot.cpp"

#include <cuda_runtime.h>

namespace {
/* the class keeps its device pointers protected: a derived probe reads them */
class RefProbe : public Particlebot {
 public:
  explicit RefProbe(SimParams p) : Particlebot(p) {}
  unsigned n() const { return params.nCells; }
  float simTime() const { return time; }
  /* which: 0 pos, 1 vel, 2 rad, 3 phase, 5 dead, 100 absForce_a, 101 absForce_r, 102 hash, 103 index, 104 cellStart, 105 cellEnd */
  const void *device(int which, size_t *bytes) {
    const size_t n = params.nCells;
    switch (which) {
      case 0: *bytes = n * 8; return mapGLBufferObject(&cuda_posvbo_resource);
      case 1: *bytes = n * 8; return dVel;
      case 2: *bytes = n * 4; return mapGLBufferObject(&cuda_radvbo_resource);
      case 3: *bytes = n * 4; return dphase;
      case 5: *bytes = n * 4; return dDead;
      case 100: *bytes = n * 4; return dAbsForce_a;
      case 101: *bytes = n * 4; return dAbsForce_r;
      case 102: *bytes = n * 4; return dGridParticleHash;
      case 103: *bytes = n * 4; return dGridParticleIndex;
      case 104: *bytes = (size_t)params.numCells * 4; return dCellStart;
      case 105: *bytes = (size_t)params.numCells * 4; return dCellEnd;
    }
    *bytes = 0;
    return nullptr;
  }
};
}  // namespace

extern "C" {
const char *prsref_identity() { return "reference host class (particlebot.cpp) + kernels (particlebot_cuda.cu), compiled verbatim, headless GL stand-in"; }
/* main.cpp:929-939, 306-310: srand(seed), then the object, then reset() */
void *prsref_create(const SimParams *p, unsigned seed) {
  srand(seed);
  return new RefProbe(*p);
}
void prsref_destroy(void *h) { delete (RefProbe *)h; }
void prsref_reset(void *h) { ((RefProbe *)h)->reset(); }
void prsref_update(void *h, float dt, float sort_interval) { ((RefProbe *)h)->update(dt, sort_interval); }
float prsref_time(void *h) { return ((RefProbe *)h)->simTime(); }
int prsref_get(void *h, int which, void *host, size_t max_bytes) {
  size_t bytes = 0;
  const void *d = ((RefProbe *)h)->device(which, &bytes);
  if (!d || bytes > max_bytes) return -1;
  if (cudaDeviceSynchronize() != cudaSuccess) return -2;
  return cudaMemcpy(host, d, bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? (int)0 : -3;
}
/* setArray of the reference (POSITION 0, VELOCITY 1, RADII 2, PHASE 3), counts in robots */
void prsref_set(void *h, int which, const float *data, int start, int count) { ((RefProbe *)h)->setArray((ParticlebotArray)which, data, start, count); }
}
