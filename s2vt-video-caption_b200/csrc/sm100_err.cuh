// Per-translation-unit device error flag for the tensor-core kernels (no relocatable device code is used, so each
// .cu that includes this header gets its own flag and exports a host reader that s2vt_device_error_flag() polls).
#pragma once
#include <cuda_runtime.h>

namespace s2vt {
static __device__ int g_sm100_error = 0;   // set when an mbarrier wait times out (never expected)

static inline int read_sm100_error_flag() {
  int v = -1;
  if (cudaMemcpyFromSymbol(&v, g_sm100_error, sizeof(int)) != cudaSuccess) return -3;
  return v;
}
static inline int clear_sm100_error_flag() {
  const int z = 0;
  return cudaMemcpyToSymbol(g_sm100_error, &z, sizeof(int)) == cudaSuccess ? 0 : -3;
}
}  // namespace s2vt
