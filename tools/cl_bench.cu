// cl_bench.cu - micro-benchmark of the exchange primitives a 16-CTA cluster can use for the reduce-scatter /
// all-gather hops of the cluster-resident flow pass (dpi_b200/csrc/flow_cl.cu). Every round is an all-to-all of
// `bytes` per (source, destination) pair followed by a read of everything received:
//   mode 0: cp.async.bulk shared::cta -> shared::cluster (DSMEM), completion on the receiver's mbarrier (complete_tx)
//   mode 1: st.shared::cluster.v4 by all threads + one remote mbarrier.arrive.release.cluster per source CTA
//   mode 2: st.global + barrier.cluster (release / acquire) + ld.global.cg   (what flow_cl.cu did first)
//   mode 3: barrier.cluster alone
//   mode 4: multicast: st.global of this CTA's slice + cp.async.bulk.multicast::cluster global -> every CTA's smem
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/cl_bench tools/cl_bench.cu ; run: tools/cl_bench
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int kCL = 16, kT = 256;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned cta_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned mapa(unsigned laddr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(laddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = smem_u32(bar);
  unsigned done = 0;
  long long t0 = clock64();
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (!done && clock64() - t0 > 400000000LL) return false;
  }
  return true;
}

__global__ void __launch_bounds__(kT, 1) bench_kernel(int mode, int iters, int bytes, float* gbuf, unsigned long long* out, int* err) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ __align__(8) unsigned long long bars[2];
  const int t = threadIdx.x, c = (int)cta_rank();
  float* send = reinterpret_cast<float*>(sm);                           // [16][bytes]
  float* recv = reinterpret_cast<float*>(sm + kCL * bytes);             // [2][16][bytes]
  const int fl = bytes / 4;                                             // floats per pair
  if (t == 0) {
    mbar_init(bars, mode == 1 ? kCL + 1 : 1);
    mbar_init(bars + 1, mode == 1 ? kCL + 1 : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = t; i < kCL * fl; i += kT) send[i] = (float)(c * 1000 + i);
  __syncthreads();
  cluster_sync_all();
  float acc = 0.f;
  bool ok = true;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int par = it & 1;
    float* rb = recv + (size_t)par * kCL * fl;
    if (mode == 0) {
      // refresh a little of the send buffer (generic proxy), make it visible to the async proxy, push, wait, read
      for (int i = t; i < kCL * fl; i += kT) send[i] += 1.f;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (t == 0) mbar_expect_tx(bars + par, kCL * bytes);
      if (t < kCL) {
        const unsigned dst = mapa(smem_u32(rb + (size_t)c * fl), t), bar = mapa(smem_u32(bars + par), t);
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                     "r"(smem_u32(send + (size_t)t * fl)), "r"(bytes), "r"(bar) : "memory");
      }
      ok = mbar_wait(bars + (it & 1), (it >> 1) & 1) && ok;
      for (int i = t; i < kCL * fl; i += kT) acc += rb[i];
      // the source of a bulk copy may only be overwritten once the copy has read it
      if (t < kCL) { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
      __syncthreads();
    } else if (mode == 1) {
      for (int i = t; i < kCL * fl / 4; i += kT) {
        const int d = i / (fl / 4), o = i % (fl / 4);
        const unsigned dst = mapa(smem_u32(rb + (size_t)c * fl + o * 4), d);
        const float v = (float)(it + i);
        asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(v), "f"(v), "f"(v), "f"(v) : "memory");
      }
      __syncthreads();
      if (t < kCL) {
        const unsigned bar = mapa(smem_u32(bars + par), t);
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
      }
      if (t == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bars + par)) : "memory");
      ok = mbar_wait(bars + (it & 1), (it >> 1) & 1) && ok;
      for (int i = t; i < kCL * fl; i += kT) acc += rb[i];
      __syncthreads();
    } else if (mode == 2) {
      float* g = gbuf + (size_t)par * kCL * kCL * fl;
      for (int i = t; i < kCL * fl / 4; i += kT)
        __stcg(reinterpret_cast<float4*>(g + (size_t)c * kCL * fl) + i, make_float4((float)it, 1.f, 2.f, 3.f));
      cluster_sync_all();
      for (int i = t; i < kCL * fl / 4; i += kT) {
        const int s = i / (fl / 4), o = i % (fl / 4);
        const float4 v = __ldcg(reinterpret_cast<const float4*>(g + ((size_t)s * kCL + c) * fl) + o);
        acc += v.x + v.w;
      }
    } else if (mode == 3) {
      cluster_sync_all();
    } else {
      float* g = gbuf + ((size_t)par * kCL + c) * fl;
      for (int i = t; i < fl / 4; i += kT) __stcg(reinterpret_cast<float4*>(g) + i, make_float4((float)it, 1.f, 2.f, 3.f));
      asm volatile("fence.proxy.async.global;" ::: "memory");
      __syncthreads();
      if (t == 0) {
        mbar_expect_tx(bars + par, kCL * bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
                         smem_u32(rb + (size_t)c * fl)), "l"(g), "r"(bytes), "r"(smem_u32(bars + par)), "h"((unsigned short)0xffff) : "memory");
      }
      ok = mbar_wait(bars + (it & 1), (it >> 1) & 1) && ok;
      for (int i = t; i < kCL * fl; i += kT) acc += rb[i];
      __syncthreads();
    }
  }
  const long long t1 = clock64();
  cluster_sync_all();
  if (t == 0) { out[c] = (unsigned long long)(t1 - t0); if (!ok) atomicExch(err, 1); }
  if (acc == 12345.678f) out[kCL] = 1;
}

int main() {
  float* gbuf; unsigned long long* out; int* err;
  cudaMalloc(&gbuf, 2ull * kCL * kCL * 4096 * 4);
  cudaMalloc(&out, 32 * 8);
  cudaMalloc(&err, 4);
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  const char* names[5] = {"bulk smem->dsmem + complete_tx", "st.shared::cluster + remote arrive", "st.global + cluster barrier + ld.global",
                          "cluster barrier alone", "st.global + multicast bulk g->smem"};
  for (int mode = 0; mode < 5; ++mode)
    for (int bytes : {256, 2048, 4096}) {
      if (mode == 3 && bytes != 256) continue;
      const int iters = 200;
      cudaMemset(err, 0, 4);
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(kCL); cfg.blockDim = dim3(kT); cfg.dynamicSmemBytes = 3 * kCL * bytes + 1024;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = kCL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      for (int rep = 0; rep < 2; ++rep) {
        cudaError_t e = cudaLaunchKernelEx(&cfg, bench_kernel, mode, iters, bytes, gbuf, out, err);
        if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
      }
      unsigned long long h[kCL]; int herr;
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost);
      unsigned long long mx = 0;
      for (int i = 0; i < kCL; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("mode %d (%s) %5d B per pair (%3d KB in+out per CTA): %8.0f cycles per round%s\n", mode, names[mode], bytes,
             kCL * bytes / 1024, (double)mx / iters, herr ? "  [TIMEOUT]" : "");
    }
  return 0;
}
