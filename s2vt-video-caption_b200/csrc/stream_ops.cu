// Stream-ordered 32-bit memory operations (driver API through the runtime's entry-point lookup, no link against libcuda):
// the host side of the wave-front coupling between two persistent sweeps (see s2vt_lstm_fwd_bf16_sync in the header).
#include "common.cuh"
#include <cuda.h>
#include <vector>

namespace s2vt {

typedef CUresult (*StreamValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

static StreamValue32Fn lookup(const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) return (StreamValue32Fn)p;
  return nullptr;
}

__global__ void timestamp_kernel(unsigned long long* slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *slot = t;
}

}  // namespace s2vt

using namespace s2vt;

extern "C" int s2vt_timestamp(void* stream, unsigned long long* slot) {
  S2VT_REQUIRE(slot, "s2vt_timestamp: null pointer");
  timestamp_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(slot);
  S2VT_REQUIRE(cudaGetLastError() == cudaSuccess, "s2vt_timestamp: launch failed");
  return 0;
}

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

// ---- executable graphs that honour per-node priorities.  A stream-captured kernel node records the priority of the stream it was
// captured on (or its cudaLaunchAttributePriority), but an executable graph only uses those priorities when it is instantiated with
// cudaGraphInstantiateFlagUseNodePriority -- without it every node runs at the priority of the launch stream and the wave front's
// coupling products queue behind bulk CTAs.
extern "C" int s2vt_graph_instantiate(void* graph, int use_node_priority, void** exec_out) {
  S2VT_REQUIRE(graph && exec_out, "s2vt_graph_instantiate: null pointer");
  cudaGraphExec_t exec = nullptr;
  const unsigned long long flags = use_node_priority ? cudaGraphInstantiateFlagUseNodePriority : 0ull;
  S2VT_CHECK_CUDA(cudaGraphInstantiateWithFlags(&exec, (cudaGraph_t)graph, flags));
  *exec_out = (void*)exec;
  return 0;
}

extern "C" int s2vt_graph_launch(void* exec, void* stream) {
  S2VT_REQUIRE(exec, "s2vt_graph_launch: null graph");
  S2VT_CHECK_CUDA(cudaGraphLaunch((cudaGraphExec_t)exec, (cudaStream_t)stream));
  return 0;
}

extern "C" int s2vt_graph_exec_destroy(void* exec) {
  if (exec) S2VT_CHECK_CUDA(cudaGraphExecDestroy((cudaGraphExec_t)exec));
  return 0;
}

extern "C" int s2vt_graph_kernel_priorities(void* graph, int* prio_out, int max_nodes, int* n_out) {
  S2VT_REQUIRE(graph && n_out, "s2vt_graph_kernel_priorities: null pointer");
  size_t n = 0;
  S2VT_CHECK_CUDA(cudaGraphGetNodes((cudaGraph_t)graph, nullptr, &n));
  std::vector<cudaGraphNode_t> nodes(n);
  if (n) S2VT_CHECK_CUDA(cudaGraphGetNodes((cudaGraph_t)graph, nodes.data(), &n));
  int k = 0;
  for (size_t i = 0; i < n; ++i) {
    cudaGraphNodeType type;
    S2VT_CHECK_CUDA(cudaGraphNodeGetType(nodes[i], &type));
    if (type != cudaGraphNodeTypeKernel) continue;
    cudaKernelNodeAttrValue val;
    S2VT_CHECK_CUDA(cudaGraphKernelNodeGetAttribute(nodes[i], cudaKernelNodeAttributePriority, &val));
    if (prio_out && k < max_nodes) prio_out[k] = val.priority;
    ++k;
  }
  *n_out = k;
  return 0;
}
