// common.cuh - shared host/device helpers for libdpi_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/dpi_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libdpi_b200 is written for sm_100a (B200) only"
#endif

namespace dpi {

// ---------------------------------------------------------------- host-side error plumbing
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define DPI_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      dpi::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return DPI_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define DPI_REQUIRE(cond, ...)                                                                 \
  do {                                                                                         \
    if (!(cond)) {                                                                             \
      dpi::set_error(__VA_ARGS__);                                                             \
      return DPI_ERR_INVALID;                                                                  \
    }                                                                                          \
  } while (0)

#define DPI_LAUNCH_CHECK()                                                                     \
  do {                                                                                         \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess) {                                                                   \
      dpi::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return DPI_ERR_CUDA;                                                                     \
    }                                                                                          \
    dpi::count_launch();                                                                       \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    target = dev;
  }
  ~DeviceGuard() {
    if (prev >= 0 && prev != target) cudaSetDevice(prev);
  }
  int target = -1;
};

inline int device_of_pointer(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return -1; }
  if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged) return -1;
  return a.device;
}

// Live-handle registry: every entry point that takes a dpi_flow_t / dpi_eht_t rejects a pointer that was never returned by
// its create function or has already been destroyed (a handle duplicated by copy.deepcopy / pickle must not be freed twice
// or used after its owner is gone).
void handle_register(const void* h, int kind);     // kind: 1 flow, 2 eht
bool handle_unregister(const void* h, int kind);   // false: not live (double destroy)
bool handle_live(const void* h, int kind);

int sm_count(int device);
int current_sm_count();  // SM count of the calling thread's current device (cached)

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__
constexpr float kLreluSlope = 0.01f;  // realnvpfc_model.py:177,184
constexpr float kBnEps = 1e-2f;       // realnvpfc_model.py:178,185
constexpr float kBnMomentum = 0.1f;   // nn.BatchNorm1d default

__device__ __forceinline__ float ldcg(const float* p) { return __ldcg(p); }
__device__ __forceinline__ double ldcg(const double* p) { return __ldcg(p); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum over the 16 lanes that share (lane >> 4)
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum, deterministic order; `scratch` holds >= 32 T. All threads get the result.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  T r = (lane < nw) ? scratch[lane] : T(0);
  r = warp_sum(r);
  return r;
}

__device__ __forceinline__ float softplus_f(float x) {
  // torch.nn.Softplus(beta=1, threshold=20)
  return x > 20.f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float lrelu_f(float v) { return v > 0.f ? v : kLreluSlope * v; }
#endif

}  // namespace dpi
