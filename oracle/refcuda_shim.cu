/*
 * refcuda_shim.cu — builds the REFERENCE's own kernels and extern "C" launch wrappers, verbatim,
 * into oracle/_ref/libprs_refcuda.so for sm_100a.  TEST INFRASTRUCTURE ONLY.
 *
 * No reference source is copied into this repository: the translation unit below textually
 * includes particlebot_cuda.cu (which itself includes particlebot_kernel_impl.cuh) from the
 * directory given with -I (oracle/Makefile passes /root/reference).  OpenGL is replaced by the
 * few declarations in oracle/gl_stub/GL; the GL-interop wrappers the file defines still link
 * (cudart provides them) but are never called by the headless driver.
 *
 * The only additions are the two helpers at the bottom, which the reference does not have:
 * the world wall is hard-coded to +-64 there, so there is nothing to set.
 */
#include "particlebot_cuda.cu"

extern "C" {
/* identifies the library to the tests */
const char *prs_refcuda_identity() { return "reference kernels (particlebot_cuda.cu) compiled verbatim for sm_100a"; }
int prs_refcuda_sizeof_simparams() { return (int)sizeof(SimParams); }
}
