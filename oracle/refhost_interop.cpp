/*
 * refhost_interop.cpp — the six OpenGL-interop entry points of the reference's particlebot.cuh (register / unregister /
 * map / unmapGLBufferObject, declared particlebot.cuh:30-37, defined particlebot_cuda.cu:69-93) on top of the headless
 * buffer-object stand-in (gl_stub/gl_headless.cpp).  TEST INFRASTRUCTURE ONLY, part of oracle/_ref/libprs_dropin.so:
 * the reference's OWN host class linked against libparticlebot_b200.so INSTEAD of its particlebot_cuda.o — the drop-in
 * claim of INTEGRATION.md §1, executed.  The product library is built without OpenGL and aborts in these six functions;
 * a GL build keeps the reference's own versions of them (they do not touch the kernels), which is what this file stands for.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

extern "C" {
cudaError_t prs_glstub_GraphicsGLRegisterBuffer(struct cudaGraphicsResource **resource, unsigned int buffer, unsigned int flags);
cudaError_t prs_glstub_GraphicsResourceGetMappedPointer(void **devPtr, size_t *size, struct cudaGraphicsResource *resource);

static void must(cudaError_t e, const char *what) {
  if (e != cudaSuccess) { fprintf(stderr, "refhost_interop: %s failed\n", what); exit(EXIT_FAILURE); }
}
void registerGLBufferObject(unsigned int vbo, struct cudaGraphicsResource **res) { must(prs_glstub_GraphicsGLRegisterBuffer(res, vbo, 0), "register"); }
void unregisterGLBufferObject(struct cudaGraphicsResource *) {}
void *mapGLBufferObject(struct cudaGraphicsResource **res) {
  void *p = nullptr;
  size_t bytes = 0;
  must(prs_glstub_GraphicsResourceGetMappedPointer(&p, &bytes, *res), "map");
  return p;
}
void unmapGLBufferObject(struct cudaGraphicsResource *) {}
}
