/*
 * prs_config.cpp — defaults, .cfg grammar and derived grid of the reference's front end
 * (main.cpp:833-911 defaults, :913-928 file loop, :594-816 setParam, :932-939 grid derivation),
 * without the GLUT/GL parts.  Host-only C++.
 *
 * The grammar is reproduced WITH its quirks because existing .cfg files depend on them:
 *   - the file is a sequence of (name line, value line) pairs; a candidate name line shorter than
 *     4 characters or starting with '#' is skipped WITHOUT consuming a value line (so `Nx` can
 *     never be set, main.cpp:924);
 *   - names are matched by strncmp prefix in a fixed order, first match wins: a line
 *     "constraint_contraction" or "constrained_contraction" hits the earlier 10-character test
 *     "constraint"/... — precisely: "constraint" (10) swallows "constraint_contraction";
 *     "constrained_contraction" does not start with "constraint" and is reachable;
 *   - "nobstacles" and "time_to_dead" are compared with a length that includes the terminator,
 *     i.e. they must match the whole line exactly (main.cpp:601, :749);
 *   - "x_cir_obs"/"y_cir_obs"/"r_cir_obs" are matched on 5 characters;
 *   - centroid_int and phase_update_interval are parsed with strtol (integers) into floats;
 *   - `config` compares the NAME against CONFIG_* and therefore never changes anything;
 *   - unknown names still consume their value line;
 *   - phase_std's default 0.3*rise_period is fixed before the file is read.
 *
 * Extension keys (not in the reference, which skips an unknown name together with its value line — a cfg that uses them
 * still loads there and runs its default CONFIG_RANDOM placement): init_config, hexblock_nx / _ny / _pitch / _jitter /
 * _seed, world_half, grid_dim.  They make the placement selector and the synthetic worlds of SURVEY.md §8d reachable
 * from a cfg file (the reference's own `config` key can never change anything, see above).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <fstream>
#include <string>

#include "prs_cabi.h"

static float *obstacle_array() { return (float *)calloc(PRS_MAX_OBSTACLES, sizeof(float)); }

extern "C" void prs_params_derive_grid(SimParams *p) {
  /* main.cpp:932-939 — the comparison and products are evaluated in double as written there */
  if (p->nDead == -1 && p->max_radius * 0.5 * p->radFactor > 2 * p->max_radius)
    p->cellSize.x = p->cellSize.y = p->max_radius * 0.5 * p->radFactor + 4 * p->max_radius;
  else
    p->cellSize.x = p->cellSize.y = p->max_radius * 2;
  p->gridSize.x = p->gridSize.y = 512;
  p->numCells = p->gridSize.x * p->gridSize.y;
  p->worldOrigin.x = -64.0f;
  p->worldOrigin.y = -64.0f;
}

extern "C" void prs_params_set_world(SimParams *p, unsigned grid_dim, float world_half) {
  p->gridSize.x = p->gridSize.y = grid_dim;
  p->numCells = grid_dim * grid_dim;
  p->worldOrigin.x = -world_half;
  p->worldOrigin.y = -world_half;
}

extern "C" void prs_params_defaults(SimParams *p, prs_run_options *o) {
  memset(p, 0, sizeof(*p));
  p->nobstacles = 0;
  p->x1obs = obstacle_array(); p->x2obs = obstacle_array(); p->y1obs = obstacle_array(); p->y2obs = obstacle_array();
  p->n_cir_obstacles = 0;
  p->x_cir_obs = obstacle_array(); p->y_cir_obs = obstacle_array(); p->r_cir_obs = obstacle_array();
  p->min_radius = 0.0775;
  p->max_radius = 0.1175;
  p->centroid_int = 10;
  p->centroid_radius = 0.05f;
  p->centroid_steps = 24000;
  p->testing = 0;
  p->friction = 0.4;
  p->spring = 1000.0f;
  p->damping = 10.0f;
  p->shear = 40.0f;
  p->constraint = 0.5f;
  p->constrained_contraction = 0;
  p->constraint_contraction = 10.0f;
  p->attraction = 3.0f * 0.000015884f;
  p->boundaryDamping = -1.0f;
  p->gravity = 9.81 * 0.566f; /* double product, then narrowed (main.cpp:866) */
  p->nCells = 501;
  p->nDead = -1;
  p->radFactor = 2.0;
  p->massFactor = 1.0;
  p->frictionFactor = 1.0;
  p->attractionFactor = 0.0f;
  p->time_to_dead = 0;
  p->max_time = 6400.0;
  p->seed = (unsigned)time(NULL);
  p->light_x = -5.0;
  p->light_y = 0;
  p->light_shadow = 0;
  p->rise_period = 2;
  p->phase_std = 0.3f * p->rise_period;
  p->config = CONFIG_RANDOM;
  p->display_shadow = 0;
  p->phase_update_interval = 12;
  p->control = LIGHT_WAVE;
  p->Nx = 5;
  p->freq = 0.5f / 25;
  if (o) {
    memset(o, 0, sizeof(*o));
    o->timestep = 0.01f;
    o->sort_interval = 180.0f;
    o->dump_interval = 60.0f;
    o->camera_y = 10;
    o->camera_x = 0;
    o->light_radius = 0.25f;
    o->display_interval = 600;
    o->video_interval = 1;
    snprintf(o->csv_filename, sizeof(o->csv_filename), "particle_bot_output_data.csv");
    snprintf(o->video_filename, sizeof(o->video_filename), "particle_bot_output_video.avi");
    o->hexblock_jitter = 0.01f;
  }
  prs_params_derive_grid(p);
}

static bool starts(const std::string &s, const char *lit, size_t n) { return strncmp(s.c_str(), lit, n) == 0; }

static void parse_list(std::string value, float *dst, int count) {
  /* space separated floats on one line, std::stof + substr like main.cpp:612-676 */
  std::string::size_type used = 0;
  for (int i = 0; i < count && i < PRS_MAX_OBSTACLES; i++) {
    if (i) value = value.substr(used);
    try {
      dst[i] = std::stof(value, &used);
    } catch (...) {
      fprintf(stderr, "cfg: obstacle list shorter than its count\n");
      exit(EXIT_FAILURE); /* the reference dies with an uncaught exception here */
    }
  }
}

static void set_param(const std::string &name, const std::string &value, SimParams *p, prs_run_options *o) {
  const char *v = value.c_str();
  auto F = [&]() { return strtof(v, NULL); };
  auto L = [&]() { return strtol(v, NULL, 10); };
  if (starts(name, "camera_y", 8)) o->camera_y = F();
  else if (starts(name, "camera_x", 8)) o->camera_x = F();
  else if (starts(name, "nobstacles", 11)) p->nobstacles = (int)L();
  else if (starts(name, "x1obs", 5)) parse_list(value, p->x1obs, p->nobstacles);
  else if (starts(name, "x2obs", 5)) parse_list(value, p->x2obs, p->nobstacles);
  else if (starts(name, "y1obs", 5)) parse_list(value, p->y1obs, p->nobstacles);
  else if (starts(name, "y2obs", 5)) parse_list(value, p->y2obs, p->nobstacles);
  else if (starts(name, "n_cir_obstacles", 15)) p->n_cir_obstacles = (int)L();
  else if (starts(name, "x_cir_obs", 5)) parse_list(value, p->x_cir_obs, p->n_cir_obstacles);
  else if (starts(name, "y_cir_obs", 5)) parse_list(value, p->y_cir_obs, p->n_cir_obstacles);
  else if (starts(name, "r_cir_obs", 5)) parse_list(value, p->r_cir_obs, p->n_cir_obstacles);
  else if (starts(name, "min_radius", 10)) p->min_radius = F();
  else if (starts(name, "max_radius", 10)) p->max_radius = F();
  else if (starts(name, "centroid_int", 12)) p->centroid_int = L();
  else if (starts(name, "centroid_radius", 15)) p->centroid_radius = F();
  else if (starts(name, "centroid_steps", 14)) p->centroid_steps = (int)L();
  else if (starts(name, "radFactor", 9)) p->radFactor = F();
  else if (starts(name, "massFactor", 10)) p->massFactor = F();
  else if (starts(name, "frictionFactor", 14)) p->frictionFactor = F();
  else if (starts(name, "attractionFactor", 16)) p->attractionFactor = F();
  else if (starts(name, "dump_interval", 13)) o->dump_interval = F();
  else if (starts(name, "sort_interval", 13)) o->sort_interval = F();
  else if (starts(name, "testing", 7)) p->testing = (unsigned)L();
  else if (starts(name, "friction", 8)) p->friction = F();
  else if (starts(name, "spring", 6)) p->spring = F();
  else if (starts(name, "damping", 7)) p->damping = F();
  else if (starts(name, "shear", 5)) p->shear = F();
  else if (starts(name, "constraint", 10)) p->constraint = F(); /* also swallows constraint_contraction */
  else if (starts(name, "constrained_contraction", 23)) p->constrained_contraction = (unsigned)L();
  else if (starts(name, "constraint_contraction", 22)) p->constraint_contraction = F(); /* unreachable */
  else if (starts(name, "attraction", 10)) p->attraction = F();
  else if (starts(name, "boundaryDamping", 15)) p->boundaryDamping = F();
  else if (starts(name, "gravity", 7)) p->gravity = F();
  else if (starts(name, "nCells", 6)) p->nCells = (unsigned)L();
  else if (starts(name, "nDead", 5)) p->nDead = (int)L();
  else if (starts(name, "time_to_dead", 14)) p->time_to_dead = F();
  else if (starts(name, "max_time", 8)) p->max_time = F();
  else if (starts(name, "seed", 4)) p->seed = (unsigned)L();
  else if (starts(name, "light_radius", 12)) o->light_radius = F();
  else if (starts(name, "light_x", 7)) p->light_x = F();
  else if (starts(name, "light_y", 7)) p->light_y = F();
  else if (starts(name, "timestep", 8)) o->timestep = F();
  else if (starts(name, "light_shadow", 12)) p->light_shadow = (unsigned)L();
  else if (starts(name, "csv_filename", 12)) snprintf(o->csv_filename, sizeof(o->csv_filename), "%s", v);
  else if (starts(name, "video_filename", 14)) snprintf(o->video_filename, sizeof(o->video_filename), "%s", v);
  else if (starts(name, "rise_period", 11)) p->rise_period = F();
  else if (starts(name, "phase_std", 9)) p->phase_std = F();
  else if (starts(name, "display_shadow", 14)) p->display_shadow = (unsigned)L();
  else if (starts(name, "phase_update_interval", 21)) p->phase_update_interval = L();
  else if (starts(name, "Nx", 2)) p->Nx = (int)L(); /* unreachable: 2-character names are filtered out */
  else if (starts(name, "config", 6)) { /* no-op in the reference: it tests the name against CONFIG_* */ }
  /* ---- extension keys ---- */
  else if (starts(name, "init_config", 11)) {
    o->init_hexblock = 0;
    if (starts(value, "random", 6)) p->config = CONFIG_RANDOM;
    else if (starts(value, "grid", 4)) p->config = CONFIG_GRID;
    else if (starts(value, "hexblock", 8)) o->init_hexblock = 1;
    else if (starts(value, "hex", 3)) p->config = CONFIG_HEX;
    else if (starts(value, "line", 4)) p->config = CONFIG_LINE;
    else { fprintf(stderr, "cfg: unknown init_config '%s' (random, grid, hex, line, hexblock)\n", v); exit(EXIT_FAILURE); }
  }
  else if (starts(name, "hexblock_nx", 11)) o->hexblock_nx = (unsigned)L();
  else if (starts(name, "hexblock_ny", 11)) o->hexblock_ny = (unsigned)L();
  else if (starts(name, "hexblock_pitch", 14)) o->hexblock_pitch = F();
  else if (starts(name, "hexblock_jitter", 15)) o->hexblock_jitter = F();
  else if (starts(name, "hexblock_seed", 13)) o->hexblock_seed = (unsigned)L();
  else if (starts(name, "world_half", 10)) o->world_half = F();
  else if (starts(name, "grid_dim", 8)) o->grid_dim = (unsigned)L();
  else if (starts(name, "DISPLAY_INTERVAL", 16)) o->display_interval = (int)L();
  else if (starts(name, "VIDEO_INTERVAL", 14)) o->video_interval = (int)L();
}

extern "C" int prs_params_load_cfg(const char *path, SimParams *p, prs_run_options *o) {
  prs_run_options scratch;
  if (!o) { o = &scratch; memset(o, 0, sizeof(*o)); }
  std::ifstream f(path);
  const bool opened = f.is_open();
  std::string name, value;
  while (opened && std::getline(f, name)) {
    if (!(name.length() < 4 || strncmp(name.c_str(), "#", 1) == 0))
      if (std::getline(f, value)) set_param(name, value, p, o);
  }
  prs_params_derive_grid(p);
  /* extension: synthetic world and hex block */
  if (o->grid_dim || o->world_half > 0.0f) {
    const unsigned g = o->grid_dim ? o->grid_dim : 512u;
    if (g & (g - 1)) { fprintf(stderr, "cfg: grid_dim must be a power of two\n"); exit(EXIT_FAILURE); }
    prs_params_set_world(p, g, o->world_half > 0.0f ? o->world_half : 64.0f);
  }
  if (o->init_hexblock) {
    if (!o->hexblock_nx || !o->hexblock_ny) { fprintf(stderr, "cfg: init_config hexblock needs hexblock_nx and hexblock_ny\n"); exit(EXIT_FAILURE); }
    p->nCells = o->hexblock_nx * o->hexblock_ny;
    if (!(o->hexblock_pitch > 0.0f)) o->hexblock_pitch = 2.0f * p->min_radius;
    if (!o->hexblock_seed) o->hexblock_seed = p->seed;
  }
  return opened ? 0 : -1;
}
