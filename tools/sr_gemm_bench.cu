// sr_gemm_bench.cu - the streamed micro-GEMM of flow_sr.cu in isolation: 16 CTAs x 512 threads, each CTA multiplies its 4
// samples with EVERY weight of a C1-sized flow (per stage 512x128, 128x128, 128x1024, k-major) straight from L2.
// Variants of the inner loop; prints cycles per stage and B/clk/SM.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int kT = 512;
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 ldw4(const float* p) {
  float4 v;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ldc4(const float* p) {   // default caching (L1 allocate)
  float4 v;
  asm("ld.global.ca.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void pf_l1(const float* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void fma16(float (&acc)[4][4], const float4 w, const float4 x) {
  const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    acc[f][0] = fmaf(wv[f], x.x, acc[f][0]); acc[f][1] = fmaf(wv[f], x.y, acc[f][1]);
    acc[f][2] = fmaf(wv[f], x.z, acc[f][2]); acc[f][3] = fmaf(wv[f], x.w, acc[f][3]);
  }
}

// V = 0: batches of U loads then U x 16 FMA; V = 1: two batches in flight (explicit double buffer of U loads each)
template <int V, int U, int pw>
__device__ __forceinline__ void gemm_stream(const float* __restrict__ W, int N, int K, const float* X, float* red) {
  const int JT = N >> 2, KG = max(1, kT / JT), kper = (K + KG - 1) / KG;
  const int t = threadIdx.x;
  if (t >= JT * KG) return;
  const int kg = t / JT, jt = t - kg * JT;
  const int k0 = min(K, kg * kper), k1 = min(K, k0 + kper);
  float acc[4][4];
#pragma unroll
  for (int f = 0; f < 4; ++f)
#pragma unroll
    for (int s = 0; s < 4; ++s) acc[f][s] = 0.f;
  const float* wp = W + (size_t)k0 * pw + 4 * jt;
  const float* xp = X + 4 * k0;
  int n = k1 - k0;
  if (V == 0) {
    while (n >= U) {
      float4 w[U];
#pragma unroll
      for (int u = 0; u < U; ++u) w[u] = ldw4(wp + (size_t)u * pw);
#pragma unroll
      for (int u = 0; u < U; ++u) fma16(acc, w[u], ld4(xp + 4 * u));
      wp += (size_t)U * pw; xp += 4 * U; n -= U;
    }
  } else if (V == 1) {
    float4 wa[U], wb[U];
    if (n >= U) {
#pragma unroll
      for (int u = 0; u < U; ++u) wa[u] = ldw4(wp + (size_t)u * pw);
    }
    while (n >= 2 * U) {
#pragma unroll
      for (int u = 0; u < U; ++u) wb[u] = ldw4(wp + (size_t)(U + u) * pw);
#pragma unroll
      for (int u = 0; u < U; ++u) fma16(acc, wa[u], ld4(xp + 4 * u));
      wp += (size_t)U * pw; xp += 4 * U; n -= U;
      if (n >= 2 * U) {
#pragma unroll
        for (int u = 0; u < U; ++u) wa[u] = ldw4(wp + (size_t)(U + u) * pw);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) fma16(acc, wb[u], ld4(xp + 4 * u));
      wp += (size_t)U * pw; xp += 4 * U; n -= U;
      if (n >= U && n < 2 * U) {   // odd number of batches: the last one is in wa already? no - reload
      }
    }
    if (n >= U) {   // one batch left: it is in wa only if the loop never ran or reloaded it
#pragma unroll
      for (int u = 0; u < U; ++u) wa[u] = ldw4(wp + (size_t)u * pw);
#pragma unroll
      for (int u = 0; u < U; ++u) fma16(acc, wa[u], ld4(xp + 4 * u));
      wp += (size_t)U * pw; xp += 4 * U; n -= U;
    }
  }
  if (V == 2) {
    // ring of U slots, prefetch distance U: a slot is re-armed right after it is consumed; no predicates in the main loop
    float4 w[U];
    if (n >= U) {
#pragma unroll
      for (int u = 0; u < U; ++u) w[u] = ldw4(wp + (size_t)u * pw);
      wp += (size_t)U * pw;
      while (n >= 2 * U) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          fma16(acc, w[u], ld4(xp + 4 * u));
          w[u] = ldw4(wp + (size_t)u * pw);
        }
        wp += (size_t)U * pw; xp += 4 * U; n -= U;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) fma16(acc, w[u], ld4(xp + 4 * u));
      xp += 4 * U; n -= U;
    }
  }
  if (V == 5 || V == 6 || V == 7 || V == 8) {   // diagnosis: 5 = no LDS (x constant), 6 = no LDG (w synthetic), 7 = loads only (1 FMA per load)
    while (n >= 8) {
      float4 w[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) w[u] = (V == 6 || V == 8) ? make_float4(acc[0][0], acc[1][1], acc[2][2], acc[3][3]) : ldw4(wp + (size_t)u * pw);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float4 x = (V == 5 || V == 8) ? make_float4(1.f, 2.f, 3.f, 4.f) : ld4(xp + 4 * u);
        if (V == 7) { acc[0][0] = fmaf(w[u].x, x.x, acc[0][0]); acc[1][1] = fmaf(w[u].y, x.y, acc[1][1]); acc[2][2] = fmaf(w[u].z, x.z, acc[2][2]); acc[3][3] = fmaf(w[u].w, x.w, acc[3][3]); }
        else fma16(acc, w[u], x);
      }
      wp += (size_t)8 * pw; xp += 32; n -= 8;
    }
  }
  if (V == 4) {
    // batches of 8 demand loads that hit L1; software prefetch into L1, U rows ahead (no registers held by loads in flight)
    for (int u = 0; u < U && u < n; ++u) pf_l1(wp + (size_t)u * pw);
    while (n >= 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (U + u < n) pf_l1(wp + (size_t)(U + u) * pw);
      float4 w[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) w[u] = ldc4(wp + (size_t)u * pw);
#pragma unroll
      for (int u = 0; u < 8; ++u) fma16(acc, w[u], ld4(xp + 4 * u));
      wp += (size_t)8 * pw; xp += 32; n -= 8;
    }
  }
  if (V == 3) {
    // explicit double buffer of 8 + 8 loads: 8 to 16 loads (128 to 256 B) per thread in flight at any time
    float4 wa[8], wb[8];
    if (n >= 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) wa[u] = ldw4(wp + (size_t)u * pw);
      wp += (size_t)8 * pw;
#pragma unroll 1
      while (n >= 16) {
#pragma unroll
        for (int u = 0; u < 8; ++u) wb[u] = ldw4(wp + (size_t)u * pw);
        wp += (size_t)8 * pw;
#pragma unroll
        for (int u = 0; u < 8; ++u) fma16(acc, wa[u], ld4(xp + 4 * u));
        xp += 32;
        if (n >= 24) {
#pragma unroll
          for (int u = 0; u < 8; ++u) wa[u] = ldw4(wp + (size_t)u * pw);
          wp += (size_t)8 * pw;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) fma16(acc, wb[u], ld4(xp + 4 * u));
        xp += 32;
        n -= 16;
      }
      if (n >= 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) fma16(acc, wa[u], ld4(xp + 4 * u));
        xp += 32; n -= 8;
      }
    }
  }
  for (; n > 0; --n) { fma16(acc, ldw4(wp), ld4(xp)); wp += pw; xp += 4; }
  float* rp = red + ((size_t)kg * N + 4 * jt) * 4;
#pragma unroll
  for (int f = 0; f < 4; ++f) st4(rp + 4 * f, make_float4(acc[f][0], acc[f][1], acc[f][2], acc[f][3]));
}

template <int V, int U>
__global__ void __launch_bounds__(kT, 1) bench(const float* __restrict__ wt, int S, int Da, int H, int Dout, unsigned long long* out, float* sink) {
  extern __shared__ __align__(16) float sm[];
  float* X = sm;                 // [1024][4]
  float* red = sm + 4096;        // 8192
  for (int i = threadIdx.x; i < 4096; i += kT) X[i] = 0.001f * i;
  __syncthreads();
  const size_t stage = (size_t)Da * H + (size_t)H * H + (size_t)H * Dout;
  const long long t0 = clock64();
  for (int e = 0; e < S; ++e) {
    const float* w1 = wt + e * stage;
    gemm_stream<V, U, 128>(w1, H, Da, X, red);
    __syncthreads();
    gemm_stream<V, U, 128>(w1 + (size_t)Da * H, H, H, X, red);
    __syncthreads();
    gemm_stream<V, U, 1024>(w1 + (size_t)Da * H + (size_t)H * H, Dout, H, X, red);
    __syncthreads();
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (red[threadIdx.x] == 123.456f) sink[0] = 1.f;
}

// steady-state rate: one long GEMM (N = 128, K = 16 * 3328: every warp streams 3328 rows of 512 B)
template <int V, int U>
__global__ void __launch_bounds__(kT, 1) bench_long(const float* __restrict__ wt, unsigned long long* out, float* sink) {
  extern __shared__ __align__(16) float sm[];
  float* X = sm;
  float* red = sm + 4096;
  for (int i = threadIdx.x; i < 4096; i += kT) X[i] = 0.001f * i;
  __syncthreads();
  const long long t0 = clock64();
  for (int rep = 0; rep < 13; ++rep) {   // 13 x 16 warps x 256 rows x 512 B = 27 MB; X has 1024 rows -> K per group 256 <= 1024 / ... use 64 rows of X cyclically
    gemm_stream<V, U, 128>(wt + (size_t)rep * 16 * 256 * 128, 128, 16 * 256, X, red);
    __syncthreads();
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (red[threadIdx.x] == 123.456f) sink[0] = 1.f;
}
template <int V, int U>
void run_long(const float* wt, unsigned long long* out, float* sink, const char* name) {
  cudaFuncSetAttribute((const void*)bench_long<V, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  bench_long<V, U><<<16, kT, 100 * 1024>>>(wt, out, sink);
  cudaDeviceSynchronize();
  bench_long<V, U><<<16, kT, 100 * 1024>>>(wt, out, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("failed: %s\n", cudaGetErrorString(e)); exit(1); }
  unsigned long long cyc; cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
  const double wk = 13.0 * 16 * 256;   // warp-k steps per warp... per CTA: 13 * 4096 rows of 512 B
  printf("%-28s long: %5.1f cycles per row of 512 B per warp-set (ideal 4 FMA-bound), %5.1f B/clk/SM\n", name, (double)cyc / (13.0 * 256), 13.0 * 4096 * 512 / (double)cyc);
}

template <int V, int U>
void run(const float* wt, unsigned long long* out, float* sink, const char* name) {
  const int S = 32, Da = 512, H = 128, Dout = 1024;
  cudaFuncSetAttribute((const void*)bench<V, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  float best = 1e30f; unsigned long long cyc = 0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<V, U><<<16, kT, 49152>>>(wt, S, Da, H, Dout, out, sink);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("failed: %s\n", cudaGetErrorString(e)); exit(1); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) { best = ms; cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost); }
  }
  const double bytes = 32.0 * (512 * 128 + 128 * 128 + 128 * 1024) * 4;
  printf("%-28s %7.1f us per pass, %6.0f cycles per stage, %5.1f B/clk/SM\n", name, best * 1e3, (double)cyc / 32, bytes / (double)cyc);
}

int main() {
  const size_t total = 32ull * (512 * 128 + 128 * 128 + 128 * 1024) * 4;
  float* wt; unsigned long long* out; float* sink;
  cudaMalloc(&wt, total); cudaMemset(wt, 0, total);
  cudaMalloc(&out, 64 * 8); cudaMalloc(&sink, 16);

  run<0, 8>(wt, out, sink, "batch 8");

  run<2, 8>(wt, out, sink, "ring 8");
  run<2, 16>(wt, out, sink, "ring 16");
  run<3, 8>(wt, out, sink, "double buffer 8 + 8");
  run<5, 8>(wt, out, sink, "batch 8, no LDS");
  run<6, 8>(wt, out, sink, "batch 8, no LDG");
  run<7, 8>(wt, out, sink, "batch 8, 4 FMA per load");
  run<8, 8>(wt, out, sink, "batch 8, FMA only");
  run_long<0, 8>(wt, out, sink, "batch 8");
  run_long<5, 8>(wt, out, sink, "batch 8, no LDS");
  run_long<6, 8>(wt, out, sink, "batch 8, no LDG");
  run_long<8, 8>(wt, out, sink, "batch 8, FMA only");
  run_long<7, 8>(wt, out, sink, "loads only");
  run_long<3, 8>(wt, out, sink, "double buffer 8 + 8");
  return 0;
}
