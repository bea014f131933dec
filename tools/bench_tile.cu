// Micro-benchmark of the cp.async tile GEMM core: cycles per k-tile as a function of pipeline length.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o bench_tile tools/bench_tile.cu && ./bench_tile
#include <cstdio>
#include <vector>
#include "../dpi_b200/csrc/tile_cp.cuh"
using namespace dpi;

template <int BN, bool MAP>
__global__ void __launch_bounds__(256) k(const float* src, const int* idx, const float* W, float* out, int M, int K, int N,
                                         long long* cyc, bool vec) {
  extern __shared__ __align__(16) float smem[];
  OperandA A{src, M, MAP ? idx : nullptr, 0, M, K, vec};
  OperandB B{W, K, nullptr, (int)blockIdx.x * BN, N, K, vec, 0, 0};
  float acc[4][BN / 16] = {};
  long long t0 = clock64();
  tile_gemm<BN, A_MCONTIG, B_KCONTIG>(acc, A, B, 0, K, smem);
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  float s = 0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < BN / 16; ++j) s += acc[i][j];
  out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <int BN, bool MAP>
void run(int K, int ctas, bool flush, bool vec) {
  const int M = 64, N = BN * ctas, D = 4096;
  float *src, *W, *out, *big;
  int* idx;
  long long* cyc;
  cudaMalloc(&src, sizeof(float) * D * M);
  cudaMalloc(&W, sizeof(float) * (size_t)N * K);
  cudaMalloc(&out, sizeof(float) * 256 * ctas);
  cudaMalloc(&idx, sizeof(int) * K);
  cudaMalloc(&cyc, sizeof(long long) * ctas);
  cudaMalloc(&big, 256 << 20);
  cudaMemset(src, 0, sizeof(float) * D * M);
  cudaMemset(W, 0, sizeof(float) * (size_t)N * K);
  std::vector<int> h(K);
  for (int i = 0; i < K; ++i) h[i] = (i * 977) % D;
  cudaMemcpy(idx, h.data(), sizeof(int) * K, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k<BN, MAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmemBytes);
  double best = 1e30;
  for (int rep = 0; rep < 5; ++rep) {
    if (flush) cudaMemset(big, rep, 256 << 20);
    cudaDeviceSynchronize();
    k<BN, MAP><<<ctas, 256, kTileSmemBytes>>>(src, idx, W, out, M, K, N, cyc, vec);
    cudaDeviceSynchronize();
    std::vector<long long> c(ctas);
    cudaMemcpy(c.data(), cyc, sizeof(long long) * ctas, cudaMemcpyDeviceToHost);
    double mx = 0;
    for (auto v : c) mx = v > mx ? v : mx;
    if (rep > 0 && mx < best) best = mx;
  }
  printf("BN=%d map=%d K=%5d ktiles=%3d ctas=%3d flush=%d vec=%d : %8.0f cycles  (%6.1f per k-tile)\n", BN, (int)MAP, K, K / 32,
         ctas, (int)flush, (int)vec, best, best / (K / 32));
  cudaFree(src); cudaFree(W); cudaFree(out); cudaFree(idx); cudaFree(cyc); cudaFree(big);
}

int main() {
  cudaError_t e = cudaGetLastError();
  for (int K : {64, 128, 512, 2048}) {
    run<16, false>(K, 8, false, true);
    run<16, true>(K, 8, false, true);
    run<16, true>(K, 8, true, true);
    run<64, false>(K, 8, false, true);
    run<64, false>(K, 128, true, true);
  }
  run<16, false>(512, 8, false, false);
  e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
