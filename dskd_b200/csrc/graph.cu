// CUDA-graph launch policy for the captured loss step: latency-bound kernels first.
//
// The step is one HBM-bound kernel with thousands of CTAs (DSG-FD) next to chains of tiny kernels (mask build, BCDD,
// the prototype all-reduce).  Inside a graph the hardware dispatches CTAs in launch order, so every small kernel that
// becomes ready after the big one has been launched waits until the big kernel's last CTA has been DISPATCHED -- the
// BCDD tail and the NCCL all-reduce end up serialised behind 90 % of the streaming kernel instead of overlapping it.
// Stream priorities would fix that, but a graph only honours priorities when it is instantiated with
// cudaGraphInstantiateFlagUseNodePriority, which PyTorch does not pass.  These entry points take the cudaGraph_t that
// `torch.cuda.CUDAGraph(keep_graph=True).raw_cuda_graph()` exposes, give every kernel node with fewer than
// `big_grid_ctas` CTAs the highest priority and the others the lowest, and instantiate / launch it themselves.
#include <vector>

#include "common.cuh"

using namespace dskd;

extern "C" int dskd_graph_instantiate_prioritized(void* graph, int32_t big_grid_ctas, void** exec_out,
                                                  int32_t* num_small, int32_t* num_big) {
  DSKD_REQUIRE(graph != nullptr && exec_out != nullptr && big_grid_ctas > 0, "dskd_graph_instantiate_prioritized: bad arguments");
  cudaGraph_t g = static_cast<cudaGraph_t>(graph);
  size_t n = 0;
  DSKD_CUDA_OK(cudaGraphGetNodes(g, nullptr, &n));
  std::vector<cudaGraphNode_t> nodes(n);
  if (n) DSKD_CUDA_OK(cudaGraphGetNodes(g, nodes.data(), &n));
  int least = 0, greatest = 0;  // numerically: greatest priority is the LOWEST number
  DSKD_CUDA_OK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
  int small = 0, big = 0;
  for (size_t i = 0; i < n; ++i) {
    cudaGraphNodeType type;
    DSKD_CUDA_OK(cudaGraphNodeGetType(nodes[i], &type));
    if (type != cudaGraphNodeTypeKernel) continue;
    cudaKernelNodeParams kp;
    memset(&kp, 0, sizeof(kp));
    long long ctas = 1;  // kernels launched through the driver API (NCCL) may not expose runtime params: treat as small
    if (cudaGraphKernelNodeGetParams(nodes[i], &kp) == cudaSuccess) ctas = (long long)kp.gridDim.x * kp.gridDim.y * kp.gridDim.z;
    else (void)cudaGetLastError();
    const bool is_big = ctas >= big_grid_ctas;
    cudaLaunchAttributeValue v;
    memset(&v, 0, sizeof(v));
    v.priority = is_big ? least : greatest;
    DSKD_CUDA_OK(cudaGraphKernelNodeSetAttribute(nodes[i], cudaLaunchAttributePriority, &v));
    (is_big ? big : small)++;
  }
  cudaGraphExec_t exec = nullptr;
  DSKD_CUDA_OK(cudaGraphInstantiateWithFlags(&exec, g, cudaGraphInstantiateFlagUseNodePriority));
  *exec_out = exec;
  if (num_small) *num_small = small;
  if (num_big) *num_big = big;
  return DSKD_OK;
}

extern "C" int dskd_graph_launch(void* exec, void* stream) {
  DSKD_REQUIRE(exec != nullptr, "dskd_graph_launch: null executable graph");
  DSKD_CUDA_OK(cudaGraphLaunch(static_cast<cudaGraphExec_t>(exec), as_stream(stream)));
  return DSKD_OK;
}

extern "C" int dskd_graph_exec_destroy(void* exec) {
  if (exec != nullptr) DSKD_CUDA_OK(cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(exec)));
  return DSKD_OK;
}
