// The one exchange step of the hot path: the sum over ranks of the BCDD prototype table
// ([2, classes, C + 1] fp32 = 164 KB; reference: rank-local, gfl_deformable_detr_head_il.py:531-551) as ONE kernel over
// NVLink / NVSwitch peer memory instead of an NCCL all-reduce.  A 164 KB all-reduce is pure latency: NCCL needs
// 20-30 us per call between `dskd_bcdd_prototypes` and `dskd_bcdd_loss_and_grad`, which at 4 images per GPU is a third
// of the step.  Here every rank
//   1. publishes its table in its own "symmetric" buffer (mapped into every peer by CUDA IPC) and raises a flag in every
//      peer's buffer with one remote store each,
//   2. waits until the flags of all peers have arrived in its own buffer,
//   3. reads the peers' tables over NVLink (ld.relaxed.sys) and adds them up in rank order -- every rank gets the
//      bit-identical sum.
// Two table slots alternate by call parity: a rank can only overwrite a slot two calls later, after it has seen the
// flags of the call in between, which its peers raise after they finished reading.  The wait is bounded (about twenty
// seconds of SM clocks): a peer that never arrives makes the kernel trap -- a loud CUDA error instead of a hung GPU.
#include <cuda.h>

#include "common.cuh"

namespace dskd {

constexpr int kPeerMaxWorld = DSKD_PEER_MAX_WORLD;
constexpr int kPeerThreads = 256;
constexpr int kPeerMaxCtas = 64;
// control words at the start of a symmetric buffer (uint32): [0 .. 15] flag of rank r (written by rank r), then
constexpr int kCtlCalls = 16;    // completed calls of this rank
constexpr int kCtlPub = 17;      // CTAs that have published their share (ticket counter, reset by the last one)
constexpr int kCtlDone = 18;     // CTAs that have finished (ticket counter, reset by the last one)
constexpr int kCtlError = 19;    // 1: a peer's flag did not arrive in time (set right before the trap)

struct PeerParams {
  float* bufs[kPeerMaxWorld];  // every rank's symmetric buffer as mapped into this process (own one included)
  float* table;                // in: this rank's sums; out: the sum over ranks
  int64_t numel4;              // float4 elements of the table (the caller pads the table to a multiple of 4 floats)
  int64_t slot_floats;         // floats per slot
  int world, rank;
  long long timeout_cycles;
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_relaxed_sys_f4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}

__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_kernel(const __grid_constant__ PeerParams p) {
  __shared__ bool last_s;
  unsigned* ctl = reinterpret_cast<unsigned*>(p.bufs[p.rank]);
  const int tid = threadIdx.x;
  // `calls` is bumped by the last CTA of a launch to FINISH, i.e. after every CTA of the launch has read it here
  const unsigned seq = ld_acquire_sys(ctl + kCtlCalls) + 1u;
  const int64_t slot_off = DSKD_PEER_CTRL_FLOATS + (int64_t)(seq & 1u) * p.slot_floats;
  const int64_t stride = (int64_t)gridDim.x * kPeerThreads;

  // ---- 1. publish this rank's table, then (last CTA) raise this rank's flag everywhere
  {
    float4* mine = reinterpret_cast<float4*>(p.bufs[p.rank] + slot_off);
    const float4* src = reinterpret_cast<const float4*>(p.table);
    for (int64_t i = blockIdx.x * (int64_t)kPeerThreads + tid; i < p.numel4; i += stride) mine[i] = src[i];
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence_system();
    last_s = atomicAdd(ctl + kCtlPub, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last_s) {
    if (tid == 0) ctl[kCtlPub] = 0u;
    if (tid < p.world) {
      __threadfence_system();
      st_release_sys(reinterpret_cast<unsigned*>(p.bufs[tid]) + p.rank, seq);
    }
  }

  // ---- 2. wait for every rank's flag in THIS rank's buffer
  if (tid < p.world) {
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(ctl + tid) - seq) < 0) {
      if (clock64() - t0 > p.timeout_cycles) {
        ctl[kCtlError] = 1u;
        __threadfence_system();
        __trap();
      }
      __nanosleep(64);
    }
  }
  __syncthreads();

  // ---- 3. sum over ranks in rank order
  {
    float4* dst = reinterpret_cast<float4*>(p.table);
    for (int64_t i = blockIdx.x * (int64_t)kPeerThreads + tid; i < p.numel4; i += stride) {
      // every peer's value is requested before the first one is used: one NVLink round trip, not `world` of them
      float4 v[kPeerMaxWorld];
#pragma unroll
      for (int r = 0; r < kPeerMaxWorld; ++r)
        if (r < p.world) v[r] = ld_relaxed_sys_f4(p.bufs[r] + slot_off + 4 * i);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < kPeerMaxWorld; ++r)
        if (r < p.world) { acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w; }
      dst[i] = acc;
    }
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(ctl + kCtlDone, 1u) == gridDim.x - 1) {
      ctl[kCtlDone] = 0u;
      __threadfence();
      st_release_sys(ctl + kCtlCalls, seq);
    }
  }
}

}  // namespace dskd

using namespace dskd;

extern "C" int dskd_ipc_export(const void* d_ptr, void* handle_out, int64_t* offset_out) {
  DSKD_REQUIRE(d_ptr != nullptr && handle_out != nullptr && offset_out != nullptr, "dskd_ipc_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == DSKD_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
  // the driver entry point is resolved at run time: the library must load on a machine without libcuda (the CPU checks)
  typedef CUresult (*GetRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  DSKD_REQUIRE(cudaGetDriverEntryPoint("cuMemGetAddressRange", &sym, cudaEnableDefault, &qres) == cudaSuccess && sym != nullptr,
               "dskd_ipc_export: cuMemGetAddressRange is not available from this driver");
  CUdeviceptr base = 0;
  size_t size = 0;
  const CUresult r = reinterpret_cast<GetRangeFn>(sym)(&base, &size, reinterpret_cast<CUdeviceptr>(d_ptr));
  DSKD_REQUIRE(r == CUDA_SUCCESS, "dskd_ipc_export: cuMemGetAddressRange failed (%d)", (int)r);
  cudaIpcMemHandle_t h;
  DSKD_CUDA_OK(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
  memcpy(handle_out, &h, sizeof(h));
  *offset_out = (int64_t)(reinterpret_cast<CUdeviceptr>(d_ptr) - base);
  return DSKD_OK;
}

extern "C" int dskd_ipc_open(const void* handle, void** base_out) {
  DSKD_REQUIRE(handle != nullptr && base_out != nullptr, "dskd_ipc_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  DSKD_CUDA_OK(cudaIpcOpenMemHandle(base_out, h, cudaIpcMemLazyEnablePeerAccess));
  return DSKD_OK;
}

extern "C" int dskd_ipc_close(void* base) {
  if (base != nullptr) DSKD_CUDA_OK(cudaIpcCloseMemHandle(base));
  return DSKD_OK;
}

extern "C" int64_t dskd_peer_buffer_floats(int64_t table_floats) {
  if (table_floats <= 0) return -1;
  const int64_t slot = (table_floats + 3) / 4 * 4;
  return DSKD_PEER_CTRL_FLOATS + 2 * slot;
}

extern "C" int dskd_peer_allreduce(float* d_table, int64_t table_floats, void* const* bufs, int32_t world, int32_t rank,
                                   void* stream) {
  DSKD_REQUIRE(d_table != nullptr && bufs != nullptr, "dskd_peer_allreduce: null pointer");
  DSKD_REQUIRE(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "dskd_peer_allreduce: bad rank %d / world %d",
               rank, world);
  DSKD_REQUIRE(table_floats > 0 && table_floats % 4 == 0 && aligned16(d_table),
               "dskd_peer_allreduce: the table must be a multiple of 4 floats and 16-byte aligned");
  PeerParams p;
  memset(&p, 0, sizeof(p));
  for (int r = 0; r < world; ++r) {
    DSKD_REQUIRE(bufs[r] != nullptr && aligned16(bufs[r]), "dskd_peer_allreduce: buffer of rank %d is null or unaligned", r);
    p.bufs[r] = static_cast<float*>(bufs[r]);
  }
  p.table = d_table;
  p.numel4 = table_floats / 4;
  p.slot_floats = table_floats;
  p.world = world;
  p.rank = rank;
  p.timeout_cycles = 40000000000ll;  // about twenty seconds at 1.9 GHz
  const int ctas = (int)std::min<int64_t>(kPeerMaxCtas, ceil_div(p.numel4, kPeerThreads));
  peer_allreduce_kernel<<<ctas, kPeerThreads, 0, as_stream(stream)>>>(p);
  DSKD_LAUNCH_OK("peer_allreduce_kernel");
  return DSKD_OK;
}
