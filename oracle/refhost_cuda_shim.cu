/*
 * refhost_cuda_shim.cu — the reference's CUDA translation unit (particlebot_cuda.cu + particlebot_kernel_impl.cuh),
 * verbatim, for oracle/_ref/libprs_refhost.so.  TEST INFRASTRUCTURE ONLY.  Same as refcuda_shim.cu except that the
 * Makefile renames the five CUDA-GL interop runtime calls to the headless stand-ins of gl_stub/gl_headless.cpp, so
 * that register / map / unmapGLBufferObject work on plain device allocations.  No reference source is copied.
 */
#include "particlebot_cuda.cu"
