/*
 * prs_frame.cuh — headless frame of the swarm.  Included by prs_kernels.cu.
 *
 * What the reference draws per displayed frame, reproduced without OpenGL:
 *   scene        main.cpp:366-466   grey clear colour 0.25, white floor square of the world, yellow light marker of radius
 *                                   light_radius, dark grey disc and box obstacles (lifted 0.01 / 0.02 above the floor:
 *                                   they cover the robots), then the robots
 *   robots       render.cpp:53-125, shaders.cpp:40-86: one point sprite per entry of the position buffer, a flat disc of
 *                                   the robot's radius in its colour (the diffuse term is overwritten, shaders.cpp:84-85);
 *                                   entries with y > 1000 are the centroid trail (calcCOG1 adds 2000, kernel_impl.cuh:343),
 *                                   lifted by their radius: drawn over everything
 *   read-back    postprocess.cu:32-55 PostprocessKernel: the rendered texture flipped to top-down rows and reordered to
 *                                   B, G, R bytes, the layout of the cv::Mat frame the video writer takes
 * The reference's camera sits at (camera_x, camera_y, 0) above the floor and looks at the origin with a 60 degree vertical
 * field of view (main.cpp:377-379, :519); for camera_x = 0 (the default and every shipped cfg) that is an exact uniform
 * scale of the floor plane, world_per_pixel = 2 camera_y tan(30 deg) / height, screen right = +x, screen up = +y (the
 * shader's x = -pos.x and the camera's right vector -x cancel).  prs_view carries that scale; oblique cameras are not
 * reproduced.
 *
 * Two kernels.  k_frame_splat: one thread per entry of the position buffer walks the pixels of its bounding box and files
 * its index with atomicMin into one of two key planes (robots / trail) — equal depth under GL_LESS lets the FIRST drawn
 * sprite win, i.e. the lowest index.  k_frame_resolve: one thread per pixel composes floor, robots, light, obstacles and
 * trail in depth order and writes B, G, R.  All coverage tests are single fp32 operations in a fixed order (no contraction),
 * so a host restatement produces the same bytes (tests/test_frame_gpu.py).
 */
#pragma once
#include "prs_device.cuh"

namespace prs {

struct FrameView {
  uint32_t width, height;
  float center_x, center_y;   /* world point at the image centre */
  float world_per_pixel;
  float light_radius;
};

/* world coordinate of a pixel centre: ((u + 0.5) - W/2) * s + cx ; rows go DOWN in the image, y goes up */
__device__ __forceinline__ float frame_world_x(const FrameView &v, uint32_t u) {
  return __fadd_rn(__fmul_rn(__fsub_rn(__fadd_rn((float)u, 0.5f), 0.5f * (float)v.width), v.world_per_pixel), v.center_x);
}
__device__ __forceinline__ float frame_world_y(const FrameView &v, uint32_t row) {
  return __fadd_rn(__fmul_rn(__fsub_rn(0.5f * (float)v.height, __fadd_rn((float)row, 0.5f)), v.world_per_pixel), v.center_y);
}
__device__ __forceinline__ bool frame_in_disc(float wx, float wy, float cx, float cy, float r) {
  const float dx = __fsub_rn(wx, cx), dy = __fsub_rn(wy, cy);
  return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(r, r);
}

/* keys: [0, W*H) robots, [W*H, 2*W*H) trail; both preset to 0xffffffff */
__global__ void __launch_bounds__(256) k_frame_splat(uint32_t *__restrict__ keys, const FrameView v, const float2 *__restrict__ pos,
                                                     const float *__restrict__ rad, uint32_t n_points) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_points) return;
  float2 p = pos[i];
  const float r = rad[i];
  uint32_t plane = 0;
  if (p.y > 1000.0f) { p.y = __fsub_rn(p.y, 2000.0f); plane = 1; } /* centroid trail marker (shaders.cpp:48-51) */
  if (!(r > 0.0f) || !(fabsf(p.x) < 1e30f) || !(fabsf(p.y) < 1e30f)) return;
  /* conservative pixel bounding box (one pixel of slack on every side; the exact test decides) */
  const float inv = 1.0f / v.world_per_pixel;
  const float uc = (p.x - v.center_x) * inv + 0.5f * (float)v.width, rc = 0.5f * (float)v.height - (p.y - v.center_y) * inv;
  const float rp = r * inv + 1.5f;
  const float u0f = floorf(uc - rp), u1f = ceilf(uc + rp), r0f = floorf(rc - rp), r1f = ceilf(rc + rp);
  if (u1f < 0.0f || r1f < 0.0f || u0f >= (float)v.width || r0f >= (float)v.height) return;
  const uint32_t u0 = (uint32_t)fmaxf(u0f, 0.0f), u1 = (uint32_t)fminf(u1f, (float)(v.width - 1));
  const uint32_t r0 = (uint32_t)fmaxf(r0f, 0.0f), r1 = (uint32_t)fminf(r1f, (float)(v.height - 1));
  uint32_t *plane_keys = keys + (size_t)plane * v.width * v.height;
  for (uint32_t row = r0; row <= r1; row++) {
    const float wy = frame_world_y(v, row);
    for (uint32_t u = u0; u <= u1; u++)
      if (frame_in_disc(frame_world_x(v, u), wy, p.x, p.y, r)) atomicMin(plane_keys + (size_t)row * v.width + u, i);
  }
}

__device__ __forceinline__ unsigned char frame_byte(float c) { /* GL's float -> unorm8 conversion */
  return (unsigned char)__float2int_rn(fminf(fmaxf(c, 0.0f), 1.0f) * 255.0f);
}

__global__ void __launch_bounds__(256) k_frame_resolve(unsigned char *__restrict__ bgr, const uint32_t *__restrict__ keys,
                                                       const FrameView v, const float4 *__restrict__ col) {
  const uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t npix = v.width * v.height;
  if (pix >= npix) return;
  const uint32_t row = pix / v.width, u = pix - row * v.width;
  const float wx = frame_world_x(v, u), wy = frame_world_y(v, row);
  const SimParams &P = c_prm.p;
  float cr = 0.25f, cg = 0.25f, cb = 0.25f;                       /* glClearColor (main.cpp:345) */
  const float W = c_prm.world_half;
  if (fabsf(wx) <= W && fabsf(wy) <= W) { cr = 1.0f; cg = 1.0f; cb = 1.0f; } /* floor polygon (main.cpp:389-396) */
  const uint32_t k = keys[pix];
  if (k != 0xffffffffu) { const float4 c = col[k]; cr = c.x; cg = c.y; cb = c.z; }
  if (frame_in_disc(wx, wy, P.light_x, P.light_y, v.light_radius)) { cr = 0.8f; cg = 0.8f; cb = 0.0f; } /* main.cpp:400-405 */
  bool obstacle = false;
  for (int i = 0; i < P.n_cir_obstacles; i++) obstacle |= frame_in_disc(wx, wy, c_prm.x_cir[i], c_prm.y_cir[i], c_prm.r_cir[i]);
  for (int i = 0; i < P.nobstacles; i++)
    obstacle |= wx >= fminf(c_prm.x1obs[i], c_prm.x2obs[i]) && wx <= fmaxf(c_prm.x1obs[i], c_prm.x2obs[i]) &&
                wy >= fminf(c_prm.y1obs[i], c_prm.y2obs[i]) && wy <= fmaxf(c_prm.y1obs[i], c_prm.y2obs[i]);
  if (obstacle) { cr = 0.2f; cg = 0.2f; cb = 0.2f; }                 /* main.cpp:408-462 */
  const uint32_t kt = keys[npix + pix];
  if (kt != 0xffffffffu) { const float4 c = col[kt]; cr = c.x; cg = c.y; cb = c.z; }
  unsigned char *o = bgr + (size_t)pix * 3;
  o[0] = frame_byte(cb);
  o[1] = frame_byte(cg);
  o[2] = frame_byte(cr);
}

}  // namespace prs
