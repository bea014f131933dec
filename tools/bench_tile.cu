// Micro-benchmark + correctness check of the tile GEMM cores (mma.sync and tcgen05) on one tile per CTA.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bench_tile tools/bench_tile.cu && ./tools/bench_tile
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../dpi_b200/csrc/tile_cp.cuh"
using namespace dpi;

// A(m,k), B(n,k) as the operand kinds define them; out[cta][n_local][m] (the epilogue ownership)
template <int BN, int AK, int BK_, bool TC>
__global__ void __launch_bounds__(256) k(const float* Asrc, const int* map, const float* Bsrc, float* out, int M, int K,
                                         int N, int lda, int ldb, long long* cyc) {
  extern __shared__ __align__(16) float smem[];
  __shared__ unsigned slot;
  TcState tc;
  tc.opbuf = reinterpret_cast<unsigned char*>(smem) + kTileSmemBytes;
  tc.bars = reinterpret_cast<unsigned long long*>(tc.opbuf + 2 * kOpBufBytes);
  tc.tmem = 0;
  if (TC) {
    if (threadIdx.x < 32) tmem_alloc(&slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    tc.tmem = slot;
  }
  OperandA A{Asrc, lda, map, 0, M, K, true};
  OperandB B{Bsrc, ldb, nullptr, (int)blockIdx.x * BN, N, K, true, 0, 0};
  float acc[4][BN / 16];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < BN / 16; ++j) acc[i][j] = 0.f;
  long long t0 = clock64();
  if (TC) tile_gemm_tc<BN, AK, BK_>(acc, A, B, 0, K, smem, tc);
  else tile_gemm<BN, AK, BK_>(acc, A, B, 0, K, smem);
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  for (int j = 0; j < BN / 16; ++j)
    for (int i = 0; i < 4; ++i) out[((size_t)blockIdx.x * BN + ty + 16 * j) * 64 + tx * 4 + i] = acc[i][j];
  if (TC) {
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tc.tmem);
  }
}

template <int BN, int AK, int BK_, bool TC>
void run(int K, int ctas, const char* name) {
  const int M = 64, N = BN * ctas, R = 512;  // R = rows of the feature-major source for A_MCONTIG
  // host data
  std::vector<float> hA, hB;
  std::vector<int> hmap(K);
  int lda, ldb;
  if (AK == A_MCONTIG) { lda = M; hA.resize((size_t)R * M); for (int i = 0; i < K; ++i) hmap[i] = (i * 37 + 11) % R; }
  else { lda = K; hA.resize((size_t)M * K); }
  if (BK_ == B_KCONTIG) { ldb = K; hB.resize((size_t)N * K); } else { ldb = N; hB.resize((size_t)K * N); }
  srand(1);
  for (auto& v : hA) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
  for (auto& v : hB) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
  auto Aat = [&](int m, int kk) { return AK == A_MCONTIG ? hA[(size_t)hmap[kk] * M + m] : hA[(size_t)m * K + kk]; };
  auto Bat = [&](int n, int kk) { return BK_ == B_KCONTIG ? hB[(size_t)n * K + kk] : hB[(size_t)kk * N + n]; };
  float *dA, *dB, *dout;
  int* dmap;
  long long* cyc;
  cudaMalloc(&dA, hA.size() * 4); cudaMalloc(&dB, hB.size() * 4); cudaMalloc(&dout, (size_t)N * 64 * 4);
  cudaMalloc(&dmap, K * 4); cudaMalloc(&cyc, ctas * 8);
  cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dmap, hmap.data(), K * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k<BN, AK, BK_, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes);
  double best = 1e30;
  for (int rep = 0; rep < 4; ++rep) {
    k<BN, AK, BK_, TC><<<ctas, 256, kTcSmemBytes>>>(dA, AK == A_MCONTIG ? dmap : nullptr, dB, dout, M, K, N, lda, ldb, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
    std::vector<long long> c(ctas);
    cudaMemcpy(c.data(), cyc, ctas * 8, cudaMemcpyDeviceToHost);
    double mx = 0;
    for (auto v : c) mx = v > mx ? v : mx;
    if (rep > 0 && mx < best) best = mx;
  }
  std::vector<float> hout((size_t)N * 64);
  cudaMemcpy(hout.data(), dout, hout.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int n = 0; n < N; ++n)
    for (int m = 0; m < M; ++m) {
      double s = 0;
      for (int kk = 0; kk < K; ++kk) s += (double)Aat(m, kk) * (double)Bat(n, kk);
      maxerr = fmax(maxerr, fabs(s - hout[(size_t)n * 64 + m]));
      maxref = fmax(maxref, fabs(s));
    }
  printf("%-34s BN=%2d K=%4d ctas=%3d : %8.0f cyc (%6.1f / k-tile)  rel.err %.2e %s\n", name, BN, K, ctas, best,
         best / (K / 32), maxerr / maxref, maxerr / maxref < 2e-5 ? "OK" : "MISMATCH");
  cudaFree(dA); cudaFree(dB); cudaFree(dout); cudaFree(dmap); cudaFree(cyc);
}

int main() {
  for (int K : {64, 512}) {
    run<16, A_MCONTIG, B_KCONTIG, false>(K, 8, "mma  A=feat-major B=K-contig");
    run<16, A_MCONTIG, B_KCONTIG, true>(K, 8, "tc   A=feat-major B=K-contig");
    run<64, A_MCONTIG, B_KCONTIG, false>(K, 8, "mma  A=feat-major B=K-contig");
    run<64, A_MCONTIG, B_KCONTIG, true>(K, 8, "tc   A=feat-major B=K-contig");
    run<32, A_MCONTIG, B_NCONTIG, false>(K, 4, "mma  A=feat-major B=N-contig");
    run<32, A_MCONTIG, B_NCONTIG, true>(K, 4, "tc   A=feat-major B=N-contig");
    run<64, A_KCONTIG, B_KCONTIG, false>(K, 4, "mma  A=K-contig    B=K-contig");
    run<64, A_KCONTIG, B_KCONTIG, true>(K, 4, "tc   A=K-contig    B=K-contig");
  }
  run<64, A_MCONTIG, B_KCONTIG, true>(2048, 128, "tc   A=feat-major B=K-contig");
  run<16, A_MCONTIG, B_NCONTIG, true>(1024, 8, "tc   A=feat-major B=N-contig");
  printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
