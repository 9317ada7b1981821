/*
 * prs_main.cpp — headless runner `ParticleBot <cfg> [--steps N] [--backend fused|percall|ext:<lib>]
 * [--no-csv] [--quiet] [--save-checkpoint FILE] [--resume-checkpoint FILE] [--video [FILE]] [--video-size WxH] [--frame-ppm FILE]
 * [--gpus N [--oversubscribe] [--no-overlap]] [--final-state FILE]`:
 * the reference's main() (main.cpp:823-967) without GLUT/GL/OpenCV.  The GLUT display callback that drives the reference
 * (dumpParticlebot, then update, main.cpp:360-361) becomes a plain loop.  Rendering is optional (north_star (5)): with
 * --video a frame is drawn every DISPLAY_INTERVAL steps by the headless frame kernels (Particlebot::renderFrame) and every
 * VIDEO_INTERVAL-th of them goes to the cfg's video_filename (or FILE) at 20 frames per second, as the reference's display()
 * + PostprocessCUDA do (main.cpp:366, 469; postprocess.cu:113-116) — an uncompressed AVI instead of XVID.  The reference's
 * window is 1920x1080 (main.cpp:65).  --frame-ppm writes the last frame as a binary PPM.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>

#include "prs_particlebot.hpp"

int main(int argc, char **argv) {
  SimParams params;
  prs_run_options opt;
  prs_params_defaults(&params, &opt);
  const char *cfg = "example.cfg";
  long max_steps = -1;
  int backend = PRS_BACKEND_FUSED;
  const char *ext = 0;
  bool csv = true, quiet = false;
  const char *ck_out = 0, *ck_in = 0;
  bool video = false;
  const char *video_path = 0, *ppm_path = 0;
  unsigned vw = 1920, vh = 1080;
  int gpus = 1, oversubscribe = 0, overlap = 1, fused_exchange = 1;
  const char *final_state = 0;
  int positional = 0;
  for (int i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "--steps") && i + 1 < argc) max_steps = atol(argv[++i]);
    else if (!strcmp(argv[i], "--backend") && i + 1 < argc) {
      const char *b = argv[++i];
      if (!strcmp(b, "fused")) backend = PRS_BACKEND_FUSED;
      else if (!strcmp(b, "percall")) backend = PRS_BACKEND_PERCALL;
      else if (!strncmp(b, "ext:", 4)) { backend = PRS_BACKEND_EXTERNAL; ext = b + 4; }
      else { fprintf(stderr, "unknown backend %s\n", b); return 2; }
    } else if (!strcmp(argv[i], "--save-checkpoint") && i + 1 < argc) ck_out = argv[++i];
    else if (!strcmp(argv[i], "--resume-checkpoint") && i + 1 < argc) ck_in = argv[++i];
    else if (!strcmp(argv[i], "--video")) {
      video = true;
      if (i + 1 < argc && argv[i + 1][0] != '-' && strstr(argv[i + 1], ".avi")) video_path = argv[++i];
    } else if (!strcmp(argv[i], "--video-size") && i + 1 < argc) {
      if (sscanf(argv[++i], "%ux%u", &vw, &vh) != 2 || !vw || !vh) { fprintf(stderr, "--video-size WxH\n"); return 2; }
    } else if (!strcmp(argv[i], "--frame-ppm") && i + 1 < argc) ppm_path = argv[++i];
    else if (!strcmp(argv[i], "--gpus") && i + 1 < argc) gpus = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--oversubscribe")) oversubscribe = 1;
    else if (!strcmp(argv[i], "--no-overlap")) overlap = 0;
    else if (!strcmp(argv[i], "--no-fused-exchange")) fused_exchange = 0;
    else if (!strcmp(argv[i], "--final-state") && i + 1 < argc) final_state = argv[++i];
    else if (!strcmp(argv[i], "--no-csv")) csv = false;
    else if (!strcmp(argv[i], "--quiet")) quiet = true;
    else if (!strcmp(argv[i], "--headless")) { /* default */ }
    else if (argv[i][0] != '-' && positional++ == 0) cfg = argv[i];
  }
  if (prs_params_load_cfg(cfg, &params, &opt) != 0)
    fprintf(stderr, "warning: cannot open %s, running with defaults (as the reference does)\n", cfg);

  if (gpus > 1) {
    /* slab decomposition, one process per GPU: the ranks are forked before this process touches CUDA (prs_multi.cpp) */
    if (backend != PRS_BACKEND_FUSED || ck_in || ck_out || video || ppm_path) {
      fprintf(stderr, "--gpus N runs the fused slab engine: no --backend / checkpoint / video options\n");
      return 2;
    }
    prs_multi_options mo;
    mo.gpus = gpus; mo.oversubscribe = oversubscribe; mo.steps = max_steps; mo.csv = csv ? 1 : 0; mo.quiet = quiet ? 1 : 0;
    mo.final_state = final_state; mo.overlap_exchange = overlap; mo.fused_exchange = fused_exchange;
    return prs_multi_run(&params, &opt, &mo);
  }

  cudaInit(argc, argv);
  FILE *fp = csv ? fopen(opt.csv_filename, "w+") : fopen("/dev/null", "w");
  if (!fp) { perror(opt.csv_filename); return 1; }
  if (quiet) { if (!freopen("/dev/null", "w", stdout)) return 1; }

  Particlebot bot(params, opt.world_half > 0.0f ? opt.world_half : 64.0f, backend, ext);
  bot.srand(params.seed); /* main.cpp:929 */
  if (opt.init_hexblock) /* extension key init_config = hexblock: the synthetic swarms of SURVEY.md §8d, placed on the device */
    bot.initHexBlock(opt.hexblock_nx, opt.hexblock_ny, opt.hexblock_pitch, opt.hexblock_jitter * params.max_radius, opt.hexblock_seed);
  else
    bot.reset();
  if (ck_in) {
    FILE *ck = fopen(ck_in, "rb");
    if (!ck || bot.loadCheckpoint(ck) != 0) { fprintf(stderr, "cannot restore %s\n", ck_in); return 1; }
    fclose(ck);
  }

  prs_view view;
  prs_view_from_camera(&view, vw, vh, opt.camera_y, opt.light_radius);
  prs_video *vid = 0;
  if (video) {
    vid = prs_video_open(video_path ? video_path : opt.video_filename, vw, vh, 20.0); /* FPS of postprocess.cu:24 */
    if (!vid) { perror(video_path ? video_path : opt.video_filename); return 1; }
  }
  const int display_every = opt.display_interval > 0 ? opt.display_interval : 1;
  const int video_every = opt.video_interval > 0 ? opt.video_interval : 1;
  long displayed = 0, frames_written = 0;

  const auto t0 = std::chrono::steady_clock::now();
  long steps = 0;
  while (max_steps < 0 || steps < max_steps) {
    bot.dumpParticlebot(0, params.nCells, fp, opt.dump_interval, params.testing, params.light_x, params.light_y);
    if (bot.update(opt.timestep, opt.sort_interval)) break; /* time > max_time: the reference exit(0)s */
    if (vid && steps % display_every == 0) { /* display(): frameCount % DISPLAY_INTERVAL == 0, then i % interval == 0 */
      if (displayed % video_every == 0) {
        if (prs_video_write(vid, bot.renderFrame(view)) != 0) { fprintf(stderr, "video file is full (4 GiB): no more frames\n"); prs_video_close(vid); vid = 0; }
        else frames_written++;
      }
      displayed++;
    }
    steps++;
  }
  bot.sync();
  if (vid) prs_video_close(vid);
  if (ppm_path) {
    FILE *pf = fopen(ppm_path, "wb");
    if (!pf) { perror(ppm_path); return 1; }
    const unsigned char *f = bot.renderFrame(view);
    fprintf(pf, "P6\n%u %u\n255\n", vw, vh);
    for (size_t i = 0; i < (size_t)vw * vh; i++) { const unsigned char rgb[3] = {f[3 * i + 2], f[3 * i + 1], f[3 * i]}; fwrite(rgb, 1, 3, pf); }
    fclose(pf);
  }
  if (ck_out) {
    FILE *ck = fopen(ck_out, "wb");
    if (!ck || bot.saveCheckpoint(ck) != 0) { fprintf(stderr, "cannot write %s\n", ck_out); return 1; }
    fclose(ck);
  }
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  fclose(fp);
  if (final_state) { /* nCells, then pos, vel, rad, phase in robot order: the format of prs_multi_run's final_state */
    FILE *out = fopen(final_state, "wb");
    if (!out) { perror(final_state); return 1; }
    const unsigned long long n64 = params.nCells;
    fwrite(&n64, 8, 1, out);
    fwrite(bot.getArray(POSITION), 4, 2 * (size_t)n64, out);
    fwrite(bot.getArray(VELOCITY), 4, 2 * (size_t)n64, out);
    fwrite(bot.getArray(RADII), 4, (size_t)n64, out);
    fwrite(bot.getArray(PHASE), 4, (size_t)n64, out);
    fclose(out);
  }
  fprintf(stderr, "ParticleBot: %ld steps, %u robots, %.3f s, %.1f steps/s, %.3e particle-steps/s\n", steps,
          params.nCells, sec, steps / sec, (double)steps * params.nCells / sec);
  if (video) fprintf(stderr, "ParticleBot: %ld video frames (%ux%u)\n", frames_written, vw, vh);
  return 0;
}
