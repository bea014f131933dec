// stream_bench.cu - how fast can ONE SM (of 16 working CTAs) ingest a weight stream from L2 into a shared-memory ring?
// Decides the design of the sample-resident small-batch pass (each CTA needs EVERY weight of the flow once per pass).
//   mode 0: 16 independent CTAs (no cluster), each streams the whole buffer with cp.async.bulk
//   mode 1: one cluster of 16, each CTA streams the whole buffer itself
//   mode 2: one cluster of 16, each CTA fetches 1/16 of every chunk and multicasts it to all 16 (slot reuse gated by
//           remote arrivals from all 16 consumers)
// Consumers (16 warps) read every chunk once with LDS.128 (like the FFMA micro-GEMM would) before releasing the slot.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/stream_bench tools/stream_bench.cu
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cuda_runtime.h>

constexpr int kCL = 16, kT = 576;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(unsigned long long* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_arrive_remote(unsigned long long* b, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(b)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(r) : "memory");
}
template <int TEST>
__device__ __forceinline__ void mbar_wait_t(unsigned long long* b, unsigned parity) {
  unsigned done = 0;
  if (TEST) {
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
  } else {
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
  }
}
#define mbar_wait(b, p) mbar_wait_t<TEST>(b, p)
__device__ __forceinline__ void mbar_wait_cluster(unsigned long long* b, unsigned parity) {
  unsigned done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(void* dst, const void* src, unsigned bytes, unsigned long long* bar, unsigned short mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}

template <int TEST>
__global__ void __launch_bounds__(kT, 1) stream_kernel(int mode, int chunk_bytes, int slots, int n_chunks, int work, int split, const float* __restrict__ src,
                                                       unsigned long long* out, float* sink) {
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ __align__(8) unsigned long long full[16], empty[16], done[16];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  unsigned rank = 0;
  if (mode == 1 || mode == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (t == 0) {
    for (int i = 0; i < slots; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, mode == 2 ? 16 : 1); mbar_init(done + i, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (mode == 1 || mode == 2) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  const long long t0 = clock64();
  float acc = 0.f;
  long long w_full = 0, w_empty = 0;
  if (mode == 3) {
    // engine only: one thread keeps `slots` copies in flight and reuses a slot as soon as its copy has landed
    if (t == 0)
      for (int c = 0; c < n_chunks + slots; ++c) {
        const int s = c % slots;
        if (c >= slots) mbar_wait(full + s, ((c / slots) - 1) & 1);
        if (c < n_chunks) {
          mbar_expect(full + s, chunk_bytes);
          bulk_g2s(ring + (size_t)s * chunk_bytes, reinterpret_cast<const char*>(src) + (size_t)c * chunk_bytes, chunk_bytes, full + s);
        }
      }
  } else if (warp == 16) {
    if (mode == 2) {
      if (lane == 0)
        for (int c = 0; c < n_chunks; ++c) {
          const int s = c % slots;
          if (c >= slots) mbar_wait_cluster(empty + s, ((c / slots) - 1) & 1);
          mbar_expect(full + s, chunk_bytes);
          const char* g = reinterpret_cast<const char*>(src) + (size_t)c * chunk_bytes;
          const int sl = chunk_bytes / kCL;
          bulk_g2s_mc(ring + (size_t)s * chunk_bytes + rank * sl, g + rank * sl, sl, full + s, 0xffff);
        }
    } else {
      // `split` lanes each own a sub-range of every chunk
      for (int c = 0; c < n_chunks; ++c) {
        const int s = c % slots;
        if (c >= slots && lane == 0) { const long long a0 = clock64(); mbar_wait(empty + s, ((c / slots) - 1) & 1); w_empty += clock64() - a0; }
        if (lane == 0) mbar_expect(full + s, chunk_bytes);
        __syncwarp();
        if (lane < split) {
          const int sl = chunk_bytes / split;
          const char* g = reinterpret_cast<const char*>(src) + (size_t)c * chunk_bytes + lane * sl;
          bulk_g2s(ring + (size_t)s * chunk_bytes + lane * sl, g, sl, full + s);
        }
      }
    }
  } else if (warp == 17) {
    if (mode == 2 && lane < kCL) {      // relay: this CTA has consumed slot s -> tell every issuer
      for (int c = 0; c < n_chunks; ++c) {
        const int s = c % slots;
        mbar_wait(done + s, (c / slots) & 1);
        mbar_arrive_remote(empty + s, lane);
      }
    }
  } else {
    for (int c = 0; c < n_chunks; ++c) {
      const int s = c % slots;
      if (t == 0) { const long long a0 = clock64(); mbar_wait(full + s, (c / slots) & 1); w_full += clock64() - a0; }
      asm volatile("bar.sync 1, 512;" ::: "memory");
      const float4* p = reinterpret_cast<const float4*>(ring + (size_t)s * chunk_bytes);
      const int n4 = chunk_bytes / 16;
      for (int rep = 0; rep < work; ++rep)
        for (int i = t; i < n4; i += 512) { const float4 v = p[i]; acc = fmaf(v.x, v.y, acc); acc = fmaf(v.z, v.w, acc); }
      asm volatile("bar.sync 2, 512;" ::: "memory");
      if (t == 0) {
        if (mode == 2) mbar_arrive(done + s);
        else mbar_arrive(empty + s);
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (t == 0) { out[blockIdx.x] = (unsigned long long)(t1 - t0); out[16 + blockIdx.x] = (unsigned long long)w_full; }
  if (t == 512) out[32 + blockIdx.x] = (unsigned long long)w_empty;
  if (acc == 123.456f) sink[0] = acc;
  if (mode == 1 || mode == 2) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
}

int main(int argc, char** argv) {
  const int ring_kb = argc > 1 ? atoi(argv[1]) : 192;
  const size_t total = 27648 * 1024;   // about one pass of C1's parameters
  float* src; unsigned long long* out; float* sink;
  cudaMalloc(&src, total); cudaMemset(src, 0, total);
  cudaMalloc(&out, 64 * 8); cudaMalloc(&sink, 16);
  cudaFuncSetAttribute((const void*)stream_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute((const void*)stream_kernel<0>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaFuncSetAttribute((const void*)stream_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute((const void*)stream_kernel<1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  const int chunk_opts[3] = {16384, 32768, 65536};
  const int split = argc > 2 ? atoi(argv[2]) : 1;
  const int tw = argc > 3 ? atoi(argv[3]) : 0;
  for (int mode = 0; mode < 1; mode += 2)
    for (int ci = 0; ci < 3; ++ci)
      for (int work = 0; work <= 1; ++work) {
        const int chunk = chunk_opts[ci];
        const int slots = std::min(16, (ring_kb * 1024) / chunk);
        const int n_chunks = (int)(total / chunk);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(kCL); cfg.blockDim = dim3(kT); cfg.dynamicSmemBytes = (size_t)slots * chunk;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (mode == 1 || mode == 2) ? kCL : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        float best = 1e30f; unsigned long long cyc = 0, wf = 0, we = 0;
        for (int rep = 0; rep < 4; ++rep) {
          cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
          cudaEventRecord(e0);
          cudaError_t e = cudaLaunchKernelEx(&cfg, tw ? stream_kernel<1> : stream_kernel<0>, mode, chunk, slots, n_chunks, work, split, (const float*)src, out, sink);
          cudaEventRecord(e1);
          if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
          e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          if (ms < best) { best = ms; cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&wf, out + 16, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&we, out + 32, 8, cudaMemcpyDeviceToHost); }
        }
        printf("mode %d chunk %5d B x %2d slots, LDS passes %d: %7.1f us per 27 MB pass, %5.1f B/clk/SM, %6.1f GB/s/SM | per chunk: total %5.0f, consumer waits full %5.0f, producer waits empty %5.0f cycles\n", mode, chunk, slots, work,
               best * 1e3, (double)total / (double)cyc, total / (best * 1e-3) / 1e9, (double)cyc / n_chunks, (double)wf / n_chunks, (double)we / n_chunks);
      }
  return 0;
}
