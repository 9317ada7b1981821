/*
 * prs_multi.cpp — slab runs driven from C++: one process per GPU, no interpreter, no NCCL (prs_multi_run, prs_cabi.h).
 *
 * The step of a slab rank is the library's (prs_slab_step, csrc/prs_slab.cuh); what a rank needs around it is small:
 *   - its part of the initial state.  Either the cfg's placement — every rank runs the SAME host code (Particlebot::reset,
 *     reference particlebot.cpp:485-801, on the glibc stream seeded as main.cpp:929 does) and keeps the robots whose grid
 *     row it owns — or the synthetic hex block, of which it generates only the lattice rows around its slab;
 *   - slab boundaries by equal robot count: cuts on the cumulative per-row histogram;
 *   - its two neighbours' mailboxes, mapped through CUDA IPC (handles passed through shared memory);
 *   - the one collective of the path: MIN of the light distance over the ranks on phase-update steps;
 *   - the dead-cell draw (particlebot.cpp:178-194): the same global ids on every rank, each marks the ones it holds;
 *   - at dump times, the swarm gathered in robot order so that rank 0 writes the reference's CSV row
 *     (particlebot.cpp:303-367: per-robot columns when testing, centroid summed in robot order) byte for byte.
 * Ranks are forked before any CUDA call and talk through one anonymous shared mapping (a process-shared barrier and a few
 * slots); the parent only waits for them and ends the others if one fails.
 */
#include <errno.h>
#include <math.h>
#include <pthread.h>
#include <signal.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <unordered_set>
#include <vector>

#include <cuda_runtime.h>

#include "prs_particlebot.hpp"

namespace {

constexpr int MAX_RANKS = 64;
constexpr unsigned HALO_ROWS = 3; /* the 2-row stencil + 1 guard row for drift between sorts */

struct Shared {
  pthread_barrier_t bar;
  unsigned char ipc[MAX_RANKS][64];
  float fmin_[MAX_RANKS];
  unsigned err[MAX_RANKS];
  unsigned long long count[MAX_RANKS];
  int ipc_ok[MAX_RANKS];
};

void die(int rank, const char *what) {
  fprintf(stderr, "prs_multi_run: rank %d: %s\n", rank, what);
  fflush(stderr);
  _exit(1);
}
void cu(cudaError_t e, int rank, const char *what) {
  if (e != cudaSuccess) {
    fprintf(stderr, "prs_multi_run: rank %d: %s: %s\n", rank, what, cudaGetErrorString(e));
    fflush(stderr);
    _exit(1);
  }
}

template <class T>
T *dalloc(size_t count) { /* zero-filled device array */
  void *p = nullptr;
  allocateArray(&p, std::max<size_t>(count, 1) * sizeof(T));
  return (T *)p;
}
void upload(void *dev, const void *host, size_t bytes) {
  size_t done = 0;
  while (done < bytes) { /* the reference ABI counts bytes in int */
    const size_t chunk = std::min<size_t>(bytes - done, (size_t)1 << 30);
    copyArrayToDevice((char *)dev + done, (const char *)host + done, 0, (int)chunk);
    done += chunk;
  }
}
void download(void *host, const void *dev, size_t bytes) {
  size_t done = 0;
  while (done < bytes) {
    const size_t chunk = std::min<size_t>(bytes - done, (size_t)1 << 30);
    copyArrayFromDevice((char *)host + done, (const char *)dev + done, 0, (int)chunk);
    done += chunk;
  }
}

inline bool gate(float time, float interval, float dt) { return time - interval * floorf(time / interval) < dt; }

/* grid row of a y coordinate: floor((y - origin.y) / cell.y) in fp32, wrapped to the power-of-two grid, as the device hash
 * computes it (calcGridPos / calcGridHash, kernel_impl.cuh:106-120) */
inline long grid_row(float y, const SimParams &p) {
  return (long)((unsigned)(int)floorf((y - p.worldOrigin.y) / p.cellSize.y) & (p.gridSize.y - 1u));
}

struct Rank {
  int rank, world;
  Shared *sh;
  float *d_min_d;
  void barrier() { pthread_barrier_wait(&sh->bar); }
};

/* the collective of the path: dev[0] = min over the ranks, ordered on the library's stream */
void allreduce_min_cb(float *dev, void *user) {
  Rank *r = (Rank *)user;
  float v;
  download(&v, dev, sizeof(float)); /* synchronises the stream */
  r->sh->fmin_[r->rank] = v;
  r->barrier();
  for (int k = 0; k < r->world; k++) v = fminf(v, r->sh->fmin_[k]);
  r->barrier();
  upload(dev, &v, sizeof(float));
}

int rank_main(int rank, int world, Shared *sh, float *g_state, SimParams p, const prs_run_options &opt, const prs_multi_options &m) {
  int ndev = 0;
  cu(cudaGetDeviceCount(&ndev), rank, "cudaGetDeviceCount");
  if (ndev == 0) die(rank, "no CUDA device");
  if (!m.oversubscribe && world > ndev) die(rank, "more ranks than devices (pass oversubscribe to share devices)");
  cu(cudaSetDevice(rank % ndev), rank, "cudaSetDevice");
  const float half = opt.world_half > 0.0f ? opt.world_half : 64.0f;
  const size_t n_total = p.nCells;
  const unsigned GY = p.gridSize.y;
  if (world > (int)GY) die(rank, "more ranks than grid rows");

  /* ---- initial state of the robots this rank may own, and the slab boundaries ---- */
  std::vector<float> pos, rad, phase;
  std::vector<int> dead;
  std::vector<unsigned> gid;
  std::vector<unsigned> rows(world + 1, 0);
  PrsRand stream; /* the glibc stream after the placement: the dead draw continues it */
  size_t row_max = 0; /* robots in the fullest grid row of the initial state (the same number on every rank) */
  if (opt.init_hexblock) {
    const unsigned nx = opt.hexblock_nx, ny = opt.hexblock_ny;
    if ((size_t)nx * ny != n_total) die(rank, "hexblock_nx * hexblock_ny != nCells");
    const float pitch = opt.hexblock_pitch, jitter = opt.hexblock_jitter * p.max_radius;
    const float row = pitch * 0.8660254037844386f, y0 = -0.5f * (float)(ny - 1) * row;
    /* lattice rows split evenly; each cut mapped to the grid row it falls in (jitter << cell) */
    for (int b = 1; b < world; b++) {
      const long r = grid_row(y0 + (float)(((size_t)ny * b) / world) * row, p);
      rows[b] = (unsigned)std::min<long>(std::max<long>(r, (long)rows[b - 1] + 1), (long)GY - (world - b));
    }
    rows[world] = GY;
    const size_t iy_lo = (size_t)std::max<long>((long)(((size_t)ny * rank) / world) - 4, 0);
    const size_t iy_hi = std::min<size_t>(((size_t)ny * (rank + 1)) / world + 4, ny);
    for (size_t i = iy_lo * nx; i < iy_hi * nx; i++) {
      float xy[2];
      prs_hex_block_position(i, nx, ny, pitch, jitter, opt.hexblock_seed, xy);
      const long r = grid_row(xy[1], p);
      if (r < (long)rows[rank] || r >= (long)rows[rank + 1]) continue;
      pos.push_back(xy[0]); pos.push_back(xy[1]);
      gid.push_back((unsigned)i);
      const bool object = p.nDead == -1 && i + 1 == n_total; /* reset(), particlebot.cpp:784-791 */
      rad.push_back(object ? p.min_radius * p.radFactor : p.min_radius);
      phase.push_back(0.0f);
      dead.push_back(object ? 1 : 0);
    }
    stream.seed(p.seed); /* main.cpp:929; a generated swarm draws nothing from the stream */
    row_max = (size_t)((double)nx * p.cellSize.y / row) + 1; /* lattice rows per grid row x robots per lattice row */
  } else {
    /* the cfg's placement: the same host code and the same stream on every rank */
    std::vector<float> hp(2 * n_total), hr(n_total), hph(n_total);
    std::vector<int> hd(n_total);
    {
      Particlebot bot(p, half, PRS_BACKEND_FUSED);
      bot.srand(p.seed);
      bot.reset();
      memcpy(hp.data(), bot.getArray(POSITION), n_total * 8);
      memcpy(hr.data(), bot.getArray(RADII), n_total * 4);
      memcpy(hph.data(), bot.getArray(PHASE), n_total * 4);
      memcpy(hd.data(), bot.getArray(DEAD), n_total * 4);
      stream = bot.randStream();
    }
    /* slab boundaries by equal robot count: cuts on the cumulative histogram of robots per grid row */
    std::vector<size_t> hist(GY, 0);
    for (size_t i = 0; i < n_total; i++) hist[(size_t)grid_row(hp[2 * i + 1], p)]++;
    row_max = *std::max_element(hist.begin(), hist.end());
    size_t cum = 0;
    unsigned row_i = 0;
    for (int b = 1; b < world; b++) {
      const double target = (double)n_total * b / world;
      while (row_i < GY && (double)(cum + hist[row_i]) < target) cum += hist[row_i++];
      const long cut = (long)row_i + 1; /* first row AFTER the b/world quantile */
      rows[b] = (unsigned)std::min<long>(std::max<long>(cut, (long)rows[b - 1] + 1), (long)GY - (world - b));
    }
    rows[world] = GY;
    for (size_t i = 0; i < n_total; i++) {
      const long r = grid_row(hp[2 * i + 1], p);
      if (r < (long)rows[rank] || r >= (long)rows[rank + 1]) continue;
      pos.push_back(hp[2 * i]); pos.push_back(hp[2 * i + 1]);
      gid.push_back((unsigned)i);
      rad.push_back(hr[i]); phase.push_back(hph[i]); dead.push_back(hd[i]);
    }
  }
  const unsigned n_own = (unsigned)gid.size();

  /* ---- the slab: capacities are the same on every rank (the mailbox layout is shared) ---- */
  prs_set_world_half_extent(half);
  setParameters(&p);
  const size_t n_expected = n_total / world + 1;
  const unsigned cap = (unsigned)(n_expected * 5 / 4) + 65536u; /* head room for unequal slabs and migration */
  /* a halo is halo_rows grid rows, a step's migrants a fraction of one: sized from the fullest grid row (1.6x head room) */
  const unsigned halo_cap = (unsigned)std::min<size_t>((size_t)(1.6 * HALO_ROWS * (double)row_max) + 4096u, (size_t)cap);
  const unsigned mig_cap = std::max<unsigned>(4096u, (unsigned)std::min<size_t>(row_max / 2, (size_t)cap));
  if (n_own > cap) die(rank, "slab capacity exceeded by the initial state");
  const size_t ncat = (size_t)cap + 2 * (size_t)halo_cap;
  prs_slab s;
  memset(&s, 0, sizeof(s));
  s.pos = dalloc<float>(2 * (size_t)cap); s.vel = dalloc<float>(2 * (size_t)cap);
  s.rad = dalloc<float>(cap); s.phase = dalloc<float>(cap);
  s.absForce_a = dalloc<float>(cap); s.absForce_r = dalloc<float>(cap);
  s.dead = dalloc<int>(cap); s.gid = dalloc<unsigned>(cap);
  s.rng = dalloc<unsigned>(12 * (size_t)cap);
  s.hash = dalloc<unsigned>(cap); s.scratch = dalloc<unsigned>(cap);
  s.sortedPR = dalloc<float>(4 * ncat); s.sortedVel = dalloc<float>(2 * ncat);
  s.hash_cat = dalloc<unsigned>(ncat); s.index_sorted = dalloc<unsigned>(cap);
  s.cellStart = dalloc<unsigned>(p.numCells); s.cellEnd = dalloc<unsigned>(p.numCells);
  s.counts = dalloc<unsigned>(16); s.lists = dalloc<unsigned>(6 * (size_t)mig_cap);
  cu(cudaMemset(s.cellStart, 0xff, (size_t)p.numCells * 4), rank, "cudaMemset");
  s.cap = cap; s.halo_cap = halo_cap; s.mig_cap = mig_cap;
  s.row_lo = rows[rank]; s.row_hi = rows[rank + 1]; s.halo_rows = HALO_ROWS;
  /* the cell hash wraps around the grid: where the grid does not cover the world the slabs form a ring */
  const bool ring = world > 1 && (float)GY * p.cellSize.y < 2.0f * half;
  s.has_dn = (rank > 0 || ring) ? 1 : 0;
  s.has_up = (rank < world - 1 || ring) ? 1 : 0;
  s.wrap = ring ? 1 : 0;
  if (n_own) {
    upload(s.pos, pos.data(), (size_t)n_own * 8);
    upload(s.rad, rad.data(), (size_t)n_own * 4);
    upload(s.phase, phase.data(), (size_t)n_own * 4);
    upload(s.dead, dead.data(), (size_t)n_own * 4);
    upload(s.gid, gid.data(), (size_t)n_own * 4);
  }
  upload(s.counts + PRS_SC_N, &n_own, 4);
  prs_slab_rng_setup(&s, n_own);

  /* ---- mailboxes ---- */
  Rank me{rank, world, sh, dalloc<float>(16)};
  prs_slab_ctx c;
  memset(&c, 0, sizeof(c));
  c.slab = s;
  c.mw = (unsigned)prs_slab_mig_words(mig_cap);
  c.hw = (unsigned)prs_slab_halo_words(halo_cap);
  c.mailbox = (unsigned *)prs_slab_mailbox_alloc(prs_slab_mailbox_words(mig_cap, halo_cap));
  for (int i = 0; i < 2; i++) { c.scratch_mig[i] = dalloc<unsigned>(c.mw); c.scratch_halo[i] = dalloc<unsigned>(c.hw); }
  if (prs_ipc_handle_size() > sizeof(sh->ipc[0])) die(rank, "IPC handle larger than its slot");
  prs_ipc_export(c.mailbox, sh->ipc[rank]);
  me.barrier();
  const int nb_dn = s.has_dn ? (rank + world - 1) % world : -1, nb_up = s.has_up ? (rank + 1) % world : -1;
  void *peer_dn = nullptr, *peer_up = nullptr;
  int ok = 1;
  if (nb_dn >= 0) { peer_dn = prs_ipc_open(sh->ipc[nb_dn]); ok &= peer_dn != nullptr; }
  if (nb_up >= 0) {
    peer_up = (nb_up == nb_dn) ? peer_dn : prs_ipc_open(sh->ipc[nb_up]); /* a ring of two: one neighbour, mapped once */
    ok &= peer_up != nullptr;
  }
  sh->ipc_ok[rank] = ok;
  me.barrier();
  for (int k = 0; k < world; k++) if (!sh->ipc_ok[k]) die(rank, "a neighbour's mailbox cannot be mapped (CUDA IPC / peer access unavailable)");
  c.peer_dn = (unsigned *)peer_dn;
  c.peer_up = (unsigned *)peer_up;
  c.d_min_d = me.d_min_d;
  c.allreduce_min = allreduce_min_cb;
  c.user = &me;
  c.overlap_exchange = m.overlap_exchange;
  c.fused_exchange = m.fused_exchange;

  /* ---- the run ---- */
  FILE *fp = nullptr;
  if (rank == 0) {
    fp = m.csv ? fopen(opt.csv_filename, "w+") : fopen("/dev/null", "w");
    if (!fp) die(rank, "cannot open the csv file");
    if (m.quiet && !freopen("/dev/null", "w", stdout)) die(rank, "freopen");
  }
  std::vector<unsigned> h_gid;
  std::vector<float> h_a, h_b, h_c;
  float *g_pos = g_state, *g_vel = g_state + 2 * n_total, *g_rad = g_state + 4 * n_total, *g_phase = g_state + 5 * n_total;
  auto owned = [&]() { unsigned n; download(&n, s.counts + PRS_SC_N, 4); return n; };
  /* every rank files its robots in the shared arrays by global id; all = velocities, radii and phases too */
  auto gather = [&](bool all) {
    const unsigned n = owned();
    h_gid.resize(n); h_a.resize(2 * (size_t)n);
    download(h_gid.data(), s.gid, (size_t)n * 4);
    download(h_a.data(), s.pos, (size_t)n * 8);
    for (unsigned k = 0; k < n; k++) { g_pos[2 * (size_t)h_gid[k]] = h_a[2 * k]; g_pos[2 * (size_t)h_gid[k] + 1] = h_a[2 * k + 1]; }
    if (all) {
      h_b.resize(2 * (size_t)n); h_c.resize(n);
      download(h_b.data(), s.vel, (size_t)n * 8);
      for (unsigned k = 0; k < n; k++) { g_vel[2 * (size_t)h_gid[k]] = h_b[2 * k]; g_vel[2 * (size_t)h_gid[k] + 1] = h_b[2 * k + 1]; }
      download(h_c.data(), s.rad, (size_t)n * 4);
      for (unsigned k = 0; k < n; k++) g_rad[h_gid[k]] = h_c[k];
      download(h_c.data(), s.phase, (size_t)n * 4);
      for (unsigned k = 0; k < n; k++) g_phase[h_gid[k]] = h_c[k];
    }
    sh->count[rank] = n;
    me.barrier();
  };
  const float dt = opt.timestep;
  const bool want_dump = m.csv || !m.quiet; /* the CSV row, or its echo on stdout (particlebot.cpp:366); neither: no gather */
  long steps = 0;
  const auto t0 = std::chrono::steady_clock::now();
  auto t_steady = t0;
  long steady_from = -1;
  while (m.steps < 0 || steps < m.steps) {
    const float time = c.time;
    if (steps == 20) { /* steady state: the first sorts (onesweep route, scratch allocations) are behind */
      threadSync();
      me.barrier();
      t_steady = std::chrono::steady_clock::now();
      steady_from = steps;
    }
    if (time > p.max_time) break; /* Particlebot::update: the reference exit(0)s here */
    /* dumpParticlebot (particlebot.cpp:303-367) on rank 0 from the gathered swarm */
    if (want_dump && !(time - opt.dump_interval * floorf(time / opt.dump_interval) > 0.01f)) {
      gather(p.testing != 0);
      if (rank == 0) {
        const unsigned count = (unsigned)n_total;
        if (time == 0) {
          fprintf(fp, "Seed, %u\n", p.seed);
          fprintf(fp, "Time,");
          if (p.testing) {
            for (unsigned i = 0; i < count; i++) fprintf(fp, "Particlebot_%d_xpos, Particlebot_%d_ypos,", i, i);
            for (unsigned i = 0; i < count; i++) fprintf(fp, "Particlebot_%d_xvel, Particlebot_%d_yvel,", i, i);
            for (unsigned i = 0; i < count; i++) fprintf(fp, "Particlebot_%d_rad,", i);
          }
          fprintf(fp, "Centroid X, Centroid Y, Distance");
          fprintf(fp, "\n");
        }
        fprintf(fp, "%f,", time);
        if (p.testing) {
          for (unsigned i = 0; i < count; i++) fprintf(fp, "%f, %f,", g_pos[i * 2 + 0], g_pos[i * 2 + 1]);
          for (unsigned i = 0; i < count; i++) fprintf(fp, "%f, %f,", g_vel[i * 2 + 0], g_vel[i * 2 + 1]);
          for (unsigned i = 0; i < count; i++) fprintf(fp, "%f,", g_rad[i]);
        }
        float sumX = 0.0f, sumY = 0.0f;
        for (unsigned i = 0; i < count; i++) { sumX += g_pos[i * 2 + 0]; sumY += g_pos[i * 2 + 1]; }
        const float cx = sumX / (float)count, cy = sumY / (float)count;
        fprintf(fp, "%f, %f, %f,", cx, cy, powf(powf(cx - p.light_x, 2.0f) + powf(cy - p.light_y, 2.0f), 0.5f));
        fprintf(fp, "\n");
        printf("%f %f %f \n", time, cx, cy);
      }
      me.barrier();
    }
    if (time >= p.time_to_dead && time < p.time_to_dead + dt && p.nDead > 0) {
      /* nDead distinct robots, rand() % remaining with erase — the same global ids on every rank */
      std::vector<int> alive(n_total);
      for (size_t i = 0; i < n_total; i++) alive[i] = (int)i;
      std::unordered_set<unsigned> ids;
      for (int drawn = 0; drawn < p.nDead && !alive.empty(); drawn++) {
        const size_t pick = (size_t)((unsigned long)stream.next() % alive.size());
        ids.insert((unsigned)alive[pick]);
        alive.erase(alive.begin() + pick);
      }
      const unsigned n = owned();
      h_gid.resize(n);
      std::vector<int> h_dead(n);
      download(h_gid.data(), s.gid, (size_t)n * 4);
      download(h_dead.data(), s.dead, (size_t)n * 4);
      for (unsigned k = 0; k < n; k++) if (ids.count(h_gid[k])) h_dead[k] = 1;
      if (n) upload(s.dead, h_dead.data(), (size_t)n * 4);
    }
    const unsigned err = prs_slab_step(&c, dt, opt.sort_interval);
    if (err) {
      fprintf(stderr, "prs_multi_run: rank %d: slab error bits 0x%x at step %ld (PRS_SLAB_ERR_* in prs_cabi.h)\n", rank, err, steps);
      _exit(1);
    }
    steps++;
  }
  threadSync();
  unsigned err_final = 0;
  download(&err_final, s.counts + PRS_SC_ERR, 4);
  if (err_final) {
    fprintf(stderr, "prs_multi_run: rank %d: slab error bits 0x%x (PRS_SLAB_ERR_* in prs_cabi.h)\n", rank, err_final);
    _exit(1);
  }
  me.barrier();
  const auto t_end = std::chrono::steady_clock::now();
  const double sec = std::chrono::duration<double>(t_end - t0).count();
  if (m.final_state) {
    gather(true);
    if (rank == 0) {
      FILE *out = fopen(m.final_state, "wb");
      if (!out) die(rank, "cannot open the final-state file");
      const unsigned long long n64 = n_total;
      unsigned long long total = 0;
      for (int k = 0; k < world; k++) total += sh->count[k];
      if (total != n64) die(rank, "robots were lost: the ranks' counts do not add up to nCells");
      fwrite(&n64, 8, 1, out);
      fwrite(g_state, 4, 6 * n_total, out);
      fclose(out);
    }
    me.barrier();
  }
  if (rank == 0) {
    fclose(fp);
    fprintf(stderr, "ParticleBot: %ld steps, %zu robots on %d ranks, %.3f s, %.1f steps/s, %.3e particle-steps/s\n", steps, n_total,
            world, sec, steps / sec, (double)steps * n_total / sec);
    if (steady_from >= 0 && steps > steady_from) {
      const double s2 = std::chrono::duration<double>(t_end - t_steady).count();
      fprintf(stderr, "ParticleBot: steps %ld..%ld: %.3f ms/step, %.3e particle-steps/s\n", steady_from, steps,
              1e3 * s2 / (steps - steady_from), (double)(steps - steady_from) * n_total / s2);
    }
  }
  prs_slab_ctx_release(&c);
  me.barrier(); /* nobody unmaps a mailbox a neighbour may still write to */
  if (peer_dn) prs_ipc_close(peer_dn);
  if (peer_up && peer_up != peer_dn) prs_ipc_close(peer_up);
  me.barrier();
  return 0;
}

}  // namespace

extern "C" int prs_multi_run(const SimParams *p, const prs_run_options *opt, const prs_multi_options *m) {
  const int world = m->gpus;
  if (world < 1 || world > MAX_RANKS) { fprintf(stderr, "prs_multi_run: 1..%d ranks\n", MAX_RANKS); return 1; }
  const size_t n_total = p->nCells;
  /* control block + the gathered swarm (pos 2, vel 2, rad 1, phase 1 floats per robot), shared by the ranks */
  const size_t bytes = sizeof(Shared) + 64 + 6 * n_total * sizeof(float);
  void *map = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  if (map == MAP_FAILED) { perror("prs_multi_run: mmap"); return 1; }
  Shared *sh = (Shared *)map;
  memset(sh, 0, sizeof(Shared));
  float *g_state = (float *)((char *)map + ((sizeof(Shared) + 63) & ~(size_t)63));
  pthread_barrierattr_t at;
  pthread_barrierattr_init(&at);
  pthread_barrierattr_setpshared(&at, PTHREAD_PROCESS_SHARED);
  pthread_barrier_init(&sh->bar, &at, (unsigned)world);
  fflush(stdout);
  fflush(stderr);
  std::vector<pid_t> kids;
  for (int r = 0; r < world; r++) {
    const pid_t pid = fork(); /* before any CUDA call of this process: every rank creates its own context */
    if (pid < 0) { perror("prs_multi_run: fork"); for (pid_t k : kids) kill(k, SIGKILL); return 1; }
    if (pid == 0) {
      const int rc = rank_main(r, world, sh, g_state, *p, *opt, *m);
      fflush(stdout);
      fflush(stderr);
      _exit(rc);
    }
    kids.push_back(pid);
  }
  int failed = 0;
  size_t left = kids.size();
  while (left) {
    int status = 0;
    const pid_t pid = wait(&status);
    if (pid < 0) { if (errno == EINTR) continue; break; }
    left--;
    const bool ok = WIFEXITED(status) && WEXITSTATUS(status) == 0;
    if (!ok && !failed) {
      failed = 1; /* a rank died: the others would wait at a barrier for ever */
      for (pid_t k : kids) if (k != pid) kill(k, SIGKILL);
    }
  }
  munmap(map, bytes);
  return failed;
}
