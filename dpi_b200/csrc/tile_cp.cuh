// tile_cp.cuh - fp32 SIMT tile GEMM core for FEATURE-MAJOR activations ([features][batch]) with a
// cp.async multi-stage shared-memory pipeline.
//
// The k-loop runs on the legacy tensor path: mma.sync m16n8k8 TF32 with the 3xTF32 split
// (x = hi + lo, a_lo*b_hi + a_hi*b_lo + a_hi*b_hi accumulated in fp32), which keeps fp32-level accuracy
// (the 1e-4 parity bar) while cutting the shared-memory operand traffic of an FFMA micro-tile by ~4x -
// these 64 x {16,32,64} tiles were LDS-bound, not FMA-bound (tools/bench_tile.cu). 8 warps = 4 (rows of
// 16) x 2 (column halves). The accumulator fragments are then re-laid through shared memory so that the
// epilogues see: thread (tx, ty) owns rows m = tx*4 + i (i < 4, contiguous -> float4) and columns
// n = ty + 16*j (j < BN/16). In the flow the GEMM-M axis is the batch (or, for weight gradients, the
// contiguous axis of the weight matrix), so every global access of an activation or a gradient is a run
// of consecutive floats and the fixed permutations of the flow are ROW gathers.
//
// Operand sources (all plain copies -> cp.async, no register staging):
//   A_MCONTIG : A(m, k) = base + row(k)*ld + m      (feature-major activation rows; row(k) optional map)
//   A_KCONTIG : A(m, k) = base + row(m)*ld + k      (weight-gradient: rows of an activation matrix)
//   B_KCONTIG : B(n, k) = base + row(n)*ld + k      (nn.Linear weight rows / activation rows)
//   B_NCONTIG : B(n, k) = base + k*ld + n           (dgrad: weight columns)
#pragma once
#include "common.cuh"

namespace dpi {

constexpr int kTM = 64;        // tile rows
constexpr int kTK = 32;        // k-tile
constexpr int kNT = 256;       // threads
#ifndef DPI_KSTAGES
#define DPI_KSTAGES 4
#endif
constexpr int kStages = DPI_KSTAGES;     // cp.async ring depth
constexpr int kStageFloatsA = 64 * 36;
constexpr int kStageFloatsB = 64 * 36;
constexpr int kStageFloats = kStageFloatsA + kStageFloatsB;
constexpr int kTileSmemBytes = kStages * kStageFloats * 4;   // 73,728 B dynamic shared memory

__device__ __forceinline__ void cp_async4(float* dst, const float* src, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int n = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async16(float* dst, const float* src, int valid_bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(valid_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

enum { A_MCONTIG = 0, A_KCONTIG = 1 };
enum { B_KCONTIG = 0, B_NCONTIG = 1 };

struct OperandA {
  const float* base; int ld;     // leading dimension in floats
  const int* map;                // optional row map (nullptr = identity)
  int m0, M;                     // first GEMM row of this tile, number of valid GEMM rows
  int K;                         // valid k extent
  bool vec;                      // 16-byte copies allowed (ld % 4 == 0, base 16-byte aligned, m0 % 4 == 0)
};
struct OperandB {
  const float* base; int ld;
  const int* map;                // optional row map for B_KCONTIG
  int n0, N;                     // first GEMM column, number of valid GEMM columns
  int K;
  bool vec;
  int pair_half, pair_db;        // B_KCONTIG only: output-layer pairing, local n -> (n / half) * db + n0 + n % half
};

// ---- per-stage issue of the asynchronous copies -------------------------------------------------
// Row indices a thread needs for one k-tile of an A_MCONTIG operand (2 chunks) - fetched one pipeline
// step ahead of the copies so the dependent index load never stalls the cp.async issue.
__device__ __forceinline__ void fetch_rows(const OperandA& a, int k0, int kend, int (&rows)[2]) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int k = k0 + ((threadIdx.x + i * kNT) >> 4);
    rows[i] = (k < kend) ? (a.map ? __ldg(a.map + k) : k) : -1;
  }
}
// Row indices for an A_KCONTIG operand: fixed for the whole tile (2 GEMM rows per thread).
__device__ __forceinline__ void fetch_rows_k(const OperandA& a, int (&rows)[2]) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int m = a.m0 + ((threadIdx.x + i * kNT) >> 3);
    rows[i] = (m < a.M) ? (a.map ? __ldg(a.map + m) : m) : -1;
  }
}

constexpr int kLdaM = 72;   // A_MCONTIG smem row stride (floats): [kk][72], bank-conflict-free fragment reads
constexpr int kLdK = 36;    // K-contiguous smem row stride (floats): [row][36]

template <int AK>
__device__ __forceinline__ void issue_A(float* sA, const OperandA& a, int k0, int kend, const int (&rows)[2],
                                        const int (&rowsk)[2]) {
  const int tid = threadIdx.x;
  if (AK == A_MCONTIG) {  // smem [kk][72]
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = tid + i * kNT;          // 32 rows x 16 chunks of 4 floats
      const int kk = c >> 4, m4 = (c & 15) * 4;
      const int m = a.m0 + m4;
      float* dst = sA + kk * kLdaM + m4;
      const bool krow = rows[i] >= 0;
      const float* src = a.base + (size_t)(krow ? rows[i] : 0) * a.ld + m;
      if (a.vec) {
        const int left = krow ? (a.M - m) : 0;
        cp_async16(dst, left > 0 ? src : a.base, left >= 4 ? 16 : (left > 0 ? left * 4 : 0));
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool ok = krow && (m + u < a.M);
          cp_async4(dst + u, ok ? src + u : a.base, ok);
        }
      }
    }
  } else {  // A_KCONTIG: smem [mm][36], 64 rows x 8 chunks of 4 k
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = tid + i * kNT;
      const int mm = c >> 3, k4 = (c & 7) * 4;
      const int k = k0 + k4;
      float* dst = sA + mm * kLdK + k4;
      const bool okm = rowsk[i] >= 0;
      const float* src = a.base + (size_t)(okm ? rowsk[i] : 0) * a.ld + k;
      if (a.vec) {
        const int left = okm ? (kend - k) : 0;
        cp_async16(dst, left > 0 ? src : a.base, left >= 4 ? 16 : (left > 0 ? left * 4 : 0));
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool ok = okm && (k + u < kend);
          cp_async4(dst + u, ok ? src + u : a.base, ok);
        }
      }
    }
  }
}

template <int BN, int BK_>
__device__ __forceinline__ void issue_B(float* sB, const OperandB& b, int k0, int kend) {
  const int tid = threadIdx.x;
  if (BK_ == B_KCONTIG) {  // smem [nn][36]
    constexpr int CH = BN * 8;  // chunks of 4 k
#pragma unroll
    for (int i = 0; i < (CH + kNT - 1) / kNT; ++i) {
      const int c = tid + i * kNT;
      if (c < CH) {
        const int nn = c >> 3, k4 = (c & 7) * 4;
        const int k = k0 + k4;
        int row; bool okn;
        if (b.pair_half) {
          const int half = nn / b.pair_half, jj = nn - half * b.pair_half;
          okn = b.n0 + jj < b.pair_db;
          row = half * b.pair_db + b.n0 + jj;
        } else {
          okn = b.n0 + nn < b.N;
          row = b.n0 + nn;
          if (okn && b.map) row = __ldg(b.map + row);
        }
        float* dst = sB + nn * 36 + k4;
        const float* src = b.base + (size_t)(okn ? row : 0) * b.ld + k;
        if (b.vec) {
          const int left = okn ? (kend - k) : 0;
          cp_async16(dst, left > 0 ? src : b.base, left >= 4 ? 16 : (left > 0 ? left * 4 : 0));
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const bool ok = okn && (k + u < kend);
            cp_async4(dst + u, ok ? src + u : b.base, ok);
          }
        }
      }
    }
  } else {  // B_NCONTIG: smem [kk][BN + 8]
    constexpr int LDN = BN + 8, CPR = BN / 4, CH = 32 * CPR;
#pragma unroll
    for (int i = 0; i < (CH + kNT - 1) / kNT; ++i) {
      const int c = tid + i * kNT;
      if (c < CH) {
        const int kk = c / CPR, n4 = (c % CPR) * 4;
        const int k = k0 + kk, n = b.n0 + n4;
        float* dst = sB + kk * LDN + n4;
        const bool krow = k < kend;
        const float* src = b.base + (size_t)(krow ? k : 0) * b.ld + n;
        if (b.vec) {
          const int left = krow ? (b.N - n) : 0;
          cp_async16(dst, left > 0 ? src : b.base, left >= 4 ? 16 : (left > 0 ? left * 4 : 0));
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const bool ok = krow && (n + u < b.N);
            cp_async4(dst + u, ok ? src + u : b.base, ok);
          }
        }
      }
    }
  }
}

__device__ __forceinline__ void split_tf32(float x, unsigned& hi, unsigned& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// acc[i][j] = sum_{k in [kbeg, kend)} A(m0 + tx*4 + i, k) * B(n0 + ty + 16 j, k)   (acc is overwritten)
template <int BN, int AK, int BK_>
__device__ __forceinline__ void tile_gemm(float (&acc)[4][BN / 16], const OperandA& a, const OperandB& b, int kbeg,
                                          int kend, float* smem) {
  constexpr int TN = BN / 16;
  constexpr int NB = BN / 16;            // 8-column blocks per warp (warp owns BN/2 columns)
  constexpr int LDN = BN + 8;
  constexpr int LDC = 68;                // accumulator staging [n][68]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int wm = (warp & 3) * 16, wn = (warp >> 2) * (BN / 2);
  const int nkt = (kend - kbeg + kTK - 1) / kTK;
  int rowsk[2] = {0, 0};
  int rows[kStages][2];
  if (AK == A_KCONTIG) {
    fetch_rows_k(a, rowsk);
#pragma unroll
    for (int s = 0; s < kStages; ++s) rows[s][0] = rows[s][1] = 0;
  } else {
#pragma unroll
    for (int s = 0; s < kStages; ++s) fetch_rows(a, kbeg + s * kTK, kend, rows[s]);  // all in flight together
  }
  float c[NB][4];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb)
#pragma unroll
    for (int q = 0; q < 4; ++q) c[nb][q] = 0.f;
  __syncthreads();  // the ring may still be read by the previous tile
#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    if (s < nkt) {
      issue_A<AK>(smem + s * kStageFloats, a, kbeg + s * kTK, kend, rows[s], rowsk);
      issue_B<BN, BK_>(smem + s * kStageFloats + kStageFloatsA, b, kbeg + s * kTK, kend);
    }
    cp_async_commit();
  }
  // rows[kStages-1] already holds the indices of tile kStages-1; later tiles rotate through r_use
  int r_use[2] = {rows[kStages - 1][0], rows[kStages - 1][1]};
  for (int t = 0; t < nkt; ++t) {
    cp_async_wait<kStages - 2>();
    __syncthreads();
    const int tn = t + kStages - 1;
    int r_next[2] = {-1, -1};
    if (AK == A_MCONTIG && tn + 1 < nkt) fetch_rows(a, kbeg + (tn + 1) * kTK, kend, r_next);
    if (tn < nkt) {
      float* st = smem + (tn % kStages) * kStageFloats;
      issue_A<AK>(st, a, kbeg + tn * kTK, kend, r_use, rowsk);
      issue_B<BN, BK_>(st + kStageFloatsA, b, kbeg + tn * kTK, kend);
    }
    cp_async_commit();
    r_use[0] = r_next[0]; r_use[1] = r_next[1];
    const float* sA = smem + (t % kStages) * kStageFloats;
    const float* sB = sA + kStageFloatsA;
#pragma unroll
    for (int ks = 0; ks < kTK / 8; ++ks) {
      float af[4];
      if (AK == A_MCONTIG) {
        af[0] = sA[(ks * 8 + t4) * kLdaM + wm + g];
        af[1] = sA[(ks * 8 + t4) * kLdaM + wm + g + 8];
        af[2] = sA[(ks * 8 + t4 + 4) * kLdaM + wm + g];
        af[3] = sA[(ks * 8 + t4 + 4) * kLdaM + wm + g + 8];
      } else {
        af[0] = sA[(wm + g) * kLdK + ks * 8 + t4];
        af[1] = sA[(wm + g + 8) * kLdK + ks * 8 + t4];
        af[2] = sA[(wm + g) * kLdK + ks * 8 + t4 + 4];
        af[3] = sA[(wm + g + 8) * kLdK + ks * 8 + t4 + 4];
      }
      unsigned ah[4], al[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) split_tf32(af[q], ah[q], al[q]);
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        const int n = wn + nb * 8 + g;
        float bf[2];
        if (BK_ == B_KCONTIG) {
          bf[0] = sB[n * kLdK + ks * 8 + t4];
          bf[1] = sB[n * kLdK + ks * 8 + t4 + 4];
        } else {
          bf[0] = sB[(ks * 8 + t4) * LDN + n];
          bf[1] = sB[(ks * 8 + t4 + 4) * LDN + n];
        }
        unsigned bh[2], bl[2];
        split_tf32(bf[0], bh[0], bl[0]);
        split_tf32(bf[1], bh[1], bl[1]);
        mma_tf32(c[nb], al, bh);   // small terms first
        mma_tf32(c[nb], ah, bl);
        mma_tf32(c[nb], ah, bh);
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  // fragments -> [n][68] staging -> the epilogue's (4 consecutive rows) x (columns ty + 16 j) ownership
  float* sC = smem;
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
    const int n = wn + nb * 8 + 2 * t4;
    sC[n * LDC + wm + g] = c[nb][0];
    sC[(n + 1) * LDC + wm + g] = c[nb][1];
    sC[n * LDC + wm + g + 8] = c[nb][2];
    sC[(n + 1) * LDC + wm + g + 8] = c[nb][3];
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const float4 v = *reinterpret_cast<const float4*>(sC + (ty + 16 * j) * LDC + tx * 4);
    acc[0][j] = v.x; acc[1][j] = v.y; acc[2][j] = v.z; acc[3][j] = v.w;
  }
  __syncthreads();
}

// ---- split-K: partial tiles stored feature-major [n_local][64], last arriver sums in split order ----
template <int BN>
__device__ __forceinline__ bool splitk_reduce_fm(float (&acc)[4][BN / 16], float* part, int ks, int splits,
                                                 unsigned* counter, int* s_flag) {
  constexpr int TN = BN / 16;
  if (splits == 1) return true;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float* mine = part + (size_t)ks * (kTM * BN);
#pragma unroll
  for (int j = 0; j < TN; ++j)
    __stcg(reinterpret_cast<float4*>(mine + (ty + 16 * j) * 64 + tx * 4),
           make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]));
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(counter), "r"(1u) : "memory");
    const int last = (old == (unsigned)(splits - 1));
    if (last) *counter = 0;  // everyone has arrived; re-arm for the next use
    *s_flag = last;
  }
  __syncthreads();
  if (!*s_flag) return false;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  for (int s0 = 0; s0 < splits; s0 += 4) {  // four splits in flight; summation order stays 0, 1, 2, ...
    float4 v[4][TN];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < TN; ++j)
        v[u][j] = (s0 + u < splits)
                      ? __ldcg(reinterpret_cast<const float4*>(part + (size_t)(s0 + u) * (kTM * BN) + (ty + 16 * j) * 64 + tx * 4))
                      : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        acc[0][j] += v[u][j].x; acc[1][j] += v[u][j].y; acc[2][j] += v[u][j].z; acc[3][j] += v[u][j].w;
      }
  }
  return true;
}

// sum over the 64 tile rows of a per-thread partial (already summed over the thread's 4 rows): the 16
// lanes that share `ty` hold the 16 row groups of the same columns
__device__ __forceinline__ float col_sum16(float v) { return half_warp_sum(v); }

// 4 consecutive floats at p (row-contiguous); vector store when allowed, bounded by n_valid
__device__ __forceinline__ void store4(float* p, const float (&v)[4], int n_valid, bool vec) {
  if (vec && n_valid >= 4) {
    __stcg(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (u < n_valid) __stcg(p + u, v[u]);
  }
}
__device__ __forceinline__ void load4(const float* p, float (&v)[4], int n_valid, bool vec) {
  if (vec && n_valid >= 4) {
    const float4 t = __ldcg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = (u < n_valid) ? __ldcg(p + u) : 0.f;
  }
}

}  // namespace dpi
