// Shared helpers for libs2vt_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/s2vt_b200.h"

namespace s2vt {

// ---- error plumbing: every extern "C" entry returns int and records a message --------------------
char* err_buf();
int fail(const char* fmt, ...);
void count_launch(int n = 1);

#define S2VT_CHECK_CUDA(expr)                                                              \
  do {                                                                                     \
    cudaError_t e__ = (expr);                                                              \
    if (e__ != cudaSuccess)                                                                \
      return s2vt::fail("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
  } while (0)

#define S2VT_CHECK_LAUNCH()                                                                \
  do {                                                                                     \
    cudaError_t e__ = cudaGetLastError();                                                  \
    if (e__ != cudaSuccess)                                                                \
      return s2vt::fail("%s:%d kernel launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
    s2vt::count_launch();                                                                  \
  } while (0)

#define S2VT_REQUIRE(cond, ...)                                                            \
  do {                                                                                     \
    if (!(cond)) return s2vt::fail(__VA_ARGS__);                                           \
  } while (0)

struct RowMap {
  int inner;
  long long so, si;
  __host__ __device__ __forceinline__ long long operator()(long long m) const {
    return (inner == 1) ? m * so : (m / inner) * so + (m % inner) * si;
  }
};
static inline RowMap to_rowmap(const s2vt_rowmap& r) { return RowMap{r.inner, (long long)r.stride_outer, (long long)r.stride_inner}; }

__device__ __forceinline__ float sigmoidf_exact(float x) { return 1.0f / (1.0f + expf(-x)); }

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int bulk_cta_cap();          // s2vt_set_bulk_cta_cap of the calling thread (0 = none); defined in misc.cu

}  // namespace s2vt
