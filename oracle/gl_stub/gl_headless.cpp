/*
 * gl_headless.cpp — buffer objects and CUDA-GL interop without OpenGL.  TEST INFRASTRUCTURE ONLY: part of
 * oracle/_ref/libprs_refhost.so, the reference's OWN host class (particlebot.cpp) and kernels compiled verbatim.
 *
 * A buffer object is a device allocation.  glMapBuffer(GL_WRITE_ONLY) hands out a zeroed host staging copy that
 * glUnmapBuffer uploads; glBufferSubData is a host-to-device copy.  The interop calls the reference's wrappers make
 * (particlebot_cuda.cu:69-93: cudaGraphicsGLRegisterBuffer, MapResources, ResourceGetMappedPointer, UnmapResources,
 * UnregisterResource) are renamed by the Makefile to the prs_glstub_* functions below: the "resource" is the buffer
 * record, mapping returns its device pointer.
 *
 * cudaMalloc is renamed too (prs_glstub_MallocZeroed): the reference reads absForce_a / absForce_r in its first
 * updateRad_light_wave and `0 * absForce_r[i]` in its first collide before anything wrote them (particlebot.cpp:
 * 148-160, 238; kernel_impl.cuh:688), i.e. it relies on a fresh process handing out zeroed device memory (SURVEY.md
 * Q6).  A test process that has used the device before hands out dirty memory, so the stand-in zeroes every allocation:
 * the harness makes the reference's assumption true instead of testing against garbage.
 */
#undef cudaMalloc /* the Makefile renames it for the reference's sources; this file needs the real one */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include <GL/glew.h>

namespace {
struct Buffer {
  void *dev = nullptr;
  size_t size = 0;
  void *staging = nullptr;
  bool live = false;
};
std::vector<Buffer> g_buffers(1); /* id 0 = "no buffer" */
GLuint g_bound = 0;

void check(cudaError_t e, const char *what) {
  if (e != cudaSuccess) {
    fprintf(stderr, "gl_headless: %s: %s\n", what, cudaGetErrorString(e));
    exit(EXIT_FAILURE);
  }
}
Buffer &bound(const char *fn) {
  if (g_bound == 0 || g_bound >= g_buffers.size() || !g_buffers[g_bound].live) {
    fprintf(stderr, "gl_headless: %s without a bound buffer\n", fn);
    exit(EXIT_FAILURE);
  }
  return g_buffers[g_bound];
}
}  // namespace

extern "C" {

void glGenBuffers(GLsizei n, GLuint *buffers) {
  for (GLsizei i = 0; i < n; i++) {
    g_buffers.push_back(Buffer());
    g_buffers.back().live = true;
    buffers[i] = (GLuint)(g_buffers.size() - 1);
  }
}
void glDeleteBuffers(GLsizei n, const GLuint *buffers) {
  for (GLsizei i = 0; i < n; i++) {
    const GLuint id = buffers[i];
    if (id == 0 || id >= g_buffers.size() || !g_buffers[id].live) continue;
    if (g_buffers[id].dev) check(cudaFree(g_buffers[id].dev), "cudaFree");
    free(g_buffers[id].staging);
    g_buffers[id] = Buffer();
  }
}
void glBindBuffer(GLenum, GLuint buffer) { g_bound = buffer; }
void glBufferData(GLenum, GLsizeiptr size, const void *data, GLenum) {
  Buffer &b = bound("glBufferData");
  if (b.dev) check(cudaFree(b.dev), "cudaFree");
  b.size = (size_t)size;
  check(cudaMalloc(&b.dev, b.size ? b.size : 1), "cudaMalloc");
  if (data) check(cudaMemcpy(b.dev, data, b.size, cudaMemcpyHostToDevice), "cudaMemcpy");
  else check(cudaMemset(b.dev, 0, b.size), "cudaMemset");
}
void glBufferSubData(GLenum, GLintptr offset, GLsizeiptr size, const void *data) {
  Buffer &b = bound("glBufferSubData");
  if ((size_t)offset + (size_t)size > b.size) { fprintf(stderr, "gl_headless: glBufferSubData out of range\n"); exit(EXIT_FAILURE); }
  check(cudaMemcpy((char *)b.dev + offset, data, (size_t)size, cudaMemcpyHostToDevice), "cudaMemcpy");
}
void *glMapBuffer(GLenum, GLenum) {
  Buffer &b = bound("glMapBuffer");
  free(b.staging);
  b.staging = calloc(b.size ? b.size : 1, 1);
  return b.staging;
}
GLboolean glUnmapBuffer(GLenum) {
  Buffer &b = bound("glUnmapBuffer");
  if (b.staging) {
    check(cudaMemcpy(b.dev, b.staging, b.size, cudaMemcpyHostToDevice), "cudaMemcpy");
    free(b.staging);
    b.staging = nullptr;
  }
  return 1;
}

cudaError_t prs_glstub_MallocZeroed(void **p, size_t size) {
  const cudaError_t e = cudaMalloc(p, size);
  if (e != cudaSuccess) return e;
  return cudaMemset(*p, 0, size);
}

/* ---- the CUDA-GL interop entry points, under the names the Makefile's -D renames give them ---- */
cudaError_t prs_glstub_GraphicsGLRegisterBuffer(struct cudaGraphicsResource **resource, GLuint buffer, unsigned int) {
  if (buffer == 0 || buffer >= g_buffers.size() || !g_buffers[buffer].live) return cudaErrorInvalidValue;
  *resource = (struct cudaGraphicsResource *)(size_t)buffer;
  return cudaSuccess;
}
cudaError_t prs_glstub_GraphicsUnregisterResource(struct cudaGraphicsResource *) { return cudaSuccess; }
cudaError_t prs_glstub_GraphicsMapResources(int, struct cudaGraphicsResource **, cudaStream_t) { return cudaSuccess; }
cudaError_t prs_glstub_GraphicsUnmapResources(int, struct cudaGraphicsResource **, cudaStream_t) { return cudaSuccess; }
cudaError_t prs_glstub_GraphicsResourceGetMappedPointer(void **devPtr, size_t *size, struct cudaGraphicsResource *resource) {
  const size_t id = (size_t)resource;
  if (id == 0 || id >= g_buffers.size() || !g_buffers[id].live) return cudaErrorInvalidValue;
  *devPtr = g_buffers[id].dev;
  if (size) *size = g_buffers[id].size;
  return cudaSuccess;
}

}  // extern "C"
