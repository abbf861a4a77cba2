// ABI plumbing: version, thread-local error string, device check, tiny utility kernels.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace dskd {
static thread_local char g_error[512] = "";
static std::atomic<uint64_t> g_launches{0};  // statistics only: kernels launched by this library

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

__global__ void scale_inplace_kernel(float* __restrict__ x, int64_t n, const float* __restrict__ factor) {
  const float f = __ldg(factor);
  if (f == 1.0f) return;  // common case (loss summed into the objective): nothing to do, no host sync
  const int64_t n4 = n >> 2;
  float4* x4 = reinterpret_cast<float4*>(x);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = x4[i];
    v.x *= f; v.y *= f; v.z *= f; v.w *= f;
    x4[i] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) x[(n4 << 2) + threadIdx.x] *= f;
}

__global__ void f64_to_f32_kernel(const double* __restrict__ in, float* __restrict__ out, int n, float scale) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += in[i];
    out[0] = (float)(s * (double)scale);
  }
}
}  // namespace dskd

using namespace dskd;

extern "C" int dskd_abi_version(void) { return DSKD_ABI_VERSION; }
extern "C" const char* dskd_last_error(void) { return g_error; }
extern "C" uint64_t dskd_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int64_t dskd_struct_size(int32_t which) {
  switch (which) {
    case 0: return (int64_t)sizeof(DskdLevel);
    case 1: return (int64_t)sizeof(DskdDsgfdMseArgs);
    case 2: return (int64_t)sizeof(DskdDsgfdKlArgs);
    case 3: return (int64_t)sizeof(DskdDsgfdStepArgs);
    case 4: return (int64_t)sizeof(DskdQmemArgs);
    default: return -1;
  }
}

extern "C" int dskd_check_device(void) {
  int dev = 0;
  DSKD_CUDA_OK(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  DSKD_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  DSKD_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10 || minor != 0) {
    set_error("dskd_b200 is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
    return DSKD_EUNSUPPORTED_ARCH;
  }
  return DSKD_OK;
}

extern "C" int dskd_scale_inplace(float* d_x, int64_t n, const float* d_factor, void* stream) {
  DSKD_REQUIRE(d_x && d_factor && n >= 0, "dskd_scale_inplace: null pointer or negative size");
  DSKD_REQUIRE(aligned16(d_x), "dskd_scale_inplace: d_x must be 16-byte aligned");
  if (n == 0) return DSKD_OK;
  const int block = 256;
  const int grid = (int)std::min<int64_t>(ceil_div(n, 4 * block), (int64_t)kNumSMs * 8);
  scale_inplace_kernel<<<grid, block, 0, as_stream(stream)>>>(d_x, n, d_factor);
  DSKD_LAUNCH_OK("scale_inplace_kernel");
  return DSKD_OK;
}

extern "C" int dskd_f64_to_f32(const double* d_in, float* d_out, int32_t n, float scale, void* stream) {
  DSKD_REQUIRE(d_in && d_out && n > 0, "dskd_f64_to_f32: bad arguments");
  f64_to_f32_kernel<<<1, 32, 0, as_stream(stream)>>>(d_in, d_out, n, scale);
  DSKD_LAUNCH_OK("f64_to_f32_kernel");
  return DSKD_OK;
}
