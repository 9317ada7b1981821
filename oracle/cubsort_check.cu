/*
 * cubsort_check.cu — cub::DeviceRadixSort::SortPairs behind two C functions.  TEST INFRASTRUCTURE ONLY
 * (oracle/_build/libprs_cubsort.so): the cross-check of the hand-written onesweep sort (north_star (2): "CUB only as
 * a cross-check") and the timing partner of `bench.py --sort-only`.  This is the sort the reference reaches through
 * thrust::sort_by_key (particlebot_cuda.cu:377-382), minus its per-call temporary allocation.
 */
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

extern "C" {
/* bytes of temporary storage for n pairs */
size_t prs_cub_sort_temp_bytes(unsigned n, int begin_bit, int end_bit) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr,
                                  (uint32_t *)nullptr, (int)n, begin_bit, end_bit, (cudaStream_t)0);
  return bytes;
}
/* stable sort of (key, value) pairs by bits [begin_bit, end_bit) of the key; returns the cudaError_t */
int prs_cub_sort_pairs(void *temp, size_t temp_bytes, const unsigned *keys_in, unsigned *keys_out, const unsigned *vals_in,
                       unsigned *vals_out, unsigned n, int begin_bit, int end_bit, void *stream) {
  return (int)cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, vals_out, (int)n, begin_bit, end_bit,
                                              (cudaStream_t)stream);
}
}
