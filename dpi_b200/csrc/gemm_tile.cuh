// gemm_tile.cuh - fp32 SIMT tile GEMM core shared by the flow program and the DFT kernels.
// 256 threads as a 16 x 16 grid; each thread owns rows ty*4+i (i<4) and columns j*16+tx (j<BN/16).
#pragma once
#include "common.cuh"

namespace dpi {

constexpr int kBM = 64;        // tile rows
constexpr int kBK = 32;        // k-tile
constexpr int kThreads = 256;

// ----------------------------------------------------------------------------- tile GEMM core
// acc[i][j] (+)= sum_k A(m = ty*4+i, k) * B(n = j*16+tx, k),  k in [kbeg, kend)
// Loaders: init() once per tile, ktile(k0) once per k-tile, operator()(row, k) -> value
// (0 outside the valid range). kContig selects the thread -> element mapping so that global
// reads are coalesced along the loader's contiguous axis.
template <int BN, class LA, class LB>
__device__ __forceinline__ void gemm_tile(float (&acc)[4][BN / 16], LA& la, LB& lb, int kbeg, int kend,
                                          float* __restrict__ sA, float* __restrict__ sB) {
  constexpr int TN = BN / 16;
  constexpr int NA = kBM * kBK / kThreads;
  constexpr int NB = BN * kBK / kThreads;
  constexpr int LD = kBK + 1;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float ra0[NA], rb0[NB], ra1[NA], rb1[NB];   // two k-tiles in flight while a third is computed
  la.init();
  lb.init();

  auto fetch = [&](int k0, float (&ra)[NA], float (&rb)[NB]) {
    la.ktile(k0);
    lb.ktile(k0);
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int e = tid + i * kThreads;
      const int kk = LA::kContig ? (e % kBK) : (e / kBM);
      const int mm = LA::kContig ? (e / kBK) : (e % kBM);
      const int k = k0 + kk;
      ra[i] = (k < kend) ? la(mm, k) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int e = tid + i * kThreads;
      const int kk = LB::kContig ? (e % kBK) : (e / BN);
      const int nn = LB::kContig ? (e / kBK) : (e % BN);
      const int k = k0 + kk;
      rb[i] = (k < kend) ? lb(nn, k) : 0.f;
    }
  };
  auto stash = [&](const float (&ra)[NA], const float (&rb)[NB]) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int e = tid + i * kThreads;
      const int kk = LA::kContig ? (e % kBK) : (e / kBM);
      const int mm = LA::kContig ? (e / kBK) : (e % kBM);
      sA[mm * LD + kk] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int e = tid + i * kThreads;
      const int kk = LB::kContig ? (e % kBK) : (e / BN);
      const int nn = LB::kContig ? (e / kBK) : (e % BN);
      sB[nn * LD + kk] = rb[i];
    }
  };
  auto compute = [&]() {
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      float a[4], b[TN];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[(ty * 4 + i) * LD + kk];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = sB[(j * 16 + tx) * LD + kk];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  };

  if (kbeg < kend) fetch(kbeg, ra0, rb0);
  if (kbeg + kBK < kend) fetch(kbeg + kBK, ra1, rb1);
  for (int k0 = kbeg; k0 < kend; k0 += 2 * kBK) {
    __syncthreads();  // previous tile's readers are done
    stash(ra0, rb0);
    __syncthreads();
    if (k0 + 2 * kBK < kend) fetch(k0 + 2 * kBK, ra0, rb0);
    compute();
    if (k0 + kBK >= kend) break;
    __syncthreads();
    stash(ra1, rb1);
    __syncthreads();
    if (k0 + 3 * kBK < kend) fetch(k0 + 3 * kBK, ra1, rb1);
    compute();
  }
  __syncthreads();
}

// Deterministic split-K: every split writes its partial tile, the last CTA to arrive sums all
// splits in index order. Returns true in the CTA that now holds the full sum in `acc`.
template <int BN>
__device__ __forceinline__ bool splitk_reduce(float (&acc)[4][BN / 16], float* part, int ks, int splits,
                                              unsigned* counter, int* s_flag) {
  constexpr int TN = BN / 16;
  if (splits == 1) return true;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float* mine = part + (size_t)ks * (kBM * BN);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) __stcg(mine + (ty * 4 + i) * BN + j * 16 + tx, acc[i][j]);
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    unsigned old = atomicAdd(counter, 1u);
    int last = (old == (unsigned)(splits - 1));
    if (last) *counter = 0;  // everyone has arrived; re-arm for the next use
    *s_flag = last;
  }
  __syncthreads();
  if (!*s_flag) return false;
  __threadfence();
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  // four splits' partials in flight at a time; summation order stays s = 0, 1, 2, ...
  for (int s0 = 0; s0 < splits; s0 += 4) {
    float v[4][4][TN];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float* src = part + (size_t)(s0 + u) * (kBM * BN);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j)
          v[u][i][j] = (s0 + u < splits) ? ldcg(src + (ty * 4 + i) * BN + j * 16 + tx) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] += v[u][i][j];
  }
  return true;
}

// Column totals over the 64 tile rows: each thread contributes v[j] for column j*16+tx; every
// thread receives the total of its own columns. Fixed summation order (ty = 0..15).
template <int TN>
__device__ __forceinline__ void col_total(const float (&v)[TN], float (&tot)[TN], float* s_red, int BN) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < TN; ++j) s_red[ty * BN + j * 16 + tx] = v[j];
  __syncthreads();
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 16; ++r) s += s_red[r * BN + j * 16 + tx];
    tot[j] = s;
  }
}

// ----------------------------------------------------------------------------- generic loaders
// B operand, K-contiguous weight rows: value = W[row(nn)*K + k]
struct LdWeightK {
  static constexpr bool kContig = true;
  const float* W; int K; int n0, N;
  // optional row remap for the output layer: local column -> (half*Db + j0 + jj)
  int pair_bj, pair_db;  // pair_bj == 0: plain
  __device__ __forceinline__ void init() {}
  __device__ __forceinline__ void ktile(int) {}
  __device__ __forceinline__ float operator()(int nn, int k) const {
    int row;
    if (pair_bj) {
      const int half = nn / pair_bj, jj = nn - half * pair_bj;
      if (n0 + jj >= pair_db) return 0.f;
      row = half * pair_db + n0 + jj;
    } else {
      row = n0 + nn;
      if (row >= N) return 0.f;
    }
    return k < K ? __ldg(W + (size_t)row * K + k) : 0.f;
  }
};

// A operand of dgrad: K-contiguous upstream gradient rows, value = g[m*ld + k]
struct LdGradRow {
  static constexpr bool kContig = true;
  const float* g; int ld; int m0, M, K;
  __device__ __forceinline__ void init() {}
  __device__ __forceinline__ void ktile(int) {}
  __device__ __forceinline__ float operator()(int mm, int k) const {
    const int m = m0 + mm;
    return (m < M && k < K) ? ldcg(g + (size_t)m * ld + k) : 0.f;
  }
};

// B operand of dgrad: W[k*ldw + n] (n contiguous)
struct LdWeightN {
  static constexpr bool kContig = false;
  const float* W; int ldw; int n0, N, K;
  __device__ __forceinline__ void init() {}
  __device__ __forceinline__ void ktile(int) {}
  __device__ __forceinline__ float operator()(int nn, int k) const {
    const int n = n0 + nn;
    return (n < N && k < K) ? __ldg(W + (size_t)k * ldw + n) : 0.f;
  }
};


}  // namespace dpi
