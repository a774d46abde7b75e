// Stream-ordered 32-bit memory operations (driver API through the runtime's entry-point lookup, no link against libcuda):
// the host side of the wave-front coupling between two persistent sweeps (see s2vt_lstm_fwd_bf16_sync in the header).
#include "common.cuh"
#include <cuda.h>

namespace s2vt {

typedef CUresult (*StreamValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

static StreamValue32Fn lookup(const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) return (StreamValue32Fn)p;
  return nullptr;
}

}  // namespace s2vt

using namespace s2vt;

extern "C" int s2vt_stream_wait_value32(void* stream, const unsigned int* addr, unsigned int value) {
  static StreamValue32Fn fn = lookup("cuStreamWaitValue32");
  S2VT_REQUIRE(fn, "s2vt_stream_wait_value32: cuStreamWaitValue32 is not available from this driver");
  S2VT_REQUIRE(addr && (reinterpret_cast<uintptr_t>(addr) & 3) == 0, "s2vt_stream_wait_value32: addr must be a 4-byte aligned device address");
  const CUresult r = fn((CUstream)stream, (CUdeviceptr)(uintptr_t)addr, value, CU_STREAM_WAIT_VALUE_GEQ);
  S2VT_REQUIRE(r == CUDA_SUCCESS, "cuStreamWaitValue32 failed with CUresult %d", (int)r);
  return 0;
}

extern "C" int s2vt_stream_write_value32(void* stream, unsigned int* addr, unsigned int value) {
  static StreamValue32Fn fn = lookup("cuStreamWriteValue32");
  S2VT_REQUIRE(fn, "s2vt_stream_write_value32: cuStreamWriteValue32 is not available from this driver");
  S2VT_REQUIRE(addr && (reinterpret_cast<uintptr_t>(addr) & 3) == 0, "s2vt_stream_write_value32: addr must be a 4-byte aligned device address");
  const CUresult r = fn((CUstream)stream, (CUdeviceptr)(uintptr_t)addr, value, CU_STREAM_WRITE_VALUE_DEFAULT);
  S2VT_REQUIRE(r == CUDA_SUCCESS, "cuStreamWriteValue32 failed with CUresult %d", (int)r);
  return 0;
}
