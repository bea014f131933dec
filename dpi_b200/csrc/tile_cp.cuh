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

constexpr int kLdaM = 68;   // A_MCONTIG raw smem row stride (floats): [kk][68], conflict-free 16-byte reads
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
  } else {  // B_NCONTIG: smem [kk][BN + 4]
    constexpr int LDN = BN + 4, CPR = BN / 4, CH = 32 * CPR;
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
  constexpr int LDN = BN + 4;
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

// =================================================================================================
// tcgen05 tile core: the same tile GEMM on the 5th-generation tensor cores.
//   D[128 x BN] (TMEM, fp32) += A[128 x 8] * B[BN x 8]  per tcgen05.mma kind::tf32 (cta_group::1).
// Only rows 0..63 of A are real (the batch / weight tile is 64 rows); rows 64..127 alias whatever
// follows the A tile in shared memory and produce accumulator rows that are never read.
// fp32 accuracy comes from the 3xTF32 split: after a raw k-tile has landed in the cp.async ring, all
// threads convert it into hi / lo operand tiles in the canonical no-swizzle UMMA layouts (core matrix
// = 8 x 16 bytes; K-major: ((8,n),2):((16B,SBO),LBO), MN-major: ((4,n),(8,k)):((4B,SBO),(16B,LBO)),
// cf. cute/atom/mma_traits_sm100.hpp) and one thread issues a_lo*b_hi, a_hi*b_lo, a_hi*b_hi.
// Operand tiles are double-buffered; tcgen05.commit -> mbarrier tells when a buffer may be rewritten.
// =================================================================================================
constexpr int kOpTileBytes = 8192;                       // one 64 x 32 fp32 operand tile
constexpr int kOpBufBytes = 4 * kOpTileBytes;            // A_hi, A_lo, B_hi, B_lo
constexpr int kTcSmemBytes = kTileSmemBytes + 2 * kOpBufBytes + 64;   // ring + 2 operand buffers + mbarriers
constexpr int kTmemCols = 64;

struct TcState {           // lives in registers / smem of the CTA for the whole kernel
  unsigned tmem;           // TMEM base address of the accumulator (64 columns x 128 lanes)
  unsigned char* opbuf;    // generic pointer to the operand double buffer (1024-byte aligned)
  unsigned long long* bars;// two mbarriers (shared memory)
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ unsigned long long umma_desc(unsigned saddr, unsigned lbo_bytes, unsigned sbo_bytes) {
  return (unsigned long long)((saddr >> 4) & 0x3FFFu) | ((unsigned long long)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((unsigned long long)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);   // version 1, no swizzle
}
// instruction descriptor: D fp32, A/B tf32, M = 128, N = BN
__device__ __forceinline__ unsigned umma_idesc(int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)a_mn_major << 15) | ((unsigned)b_mn_major << 16) |
         ((unsigned)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc,
                                          unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc),
      "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = smem_u32(bar);
  unsigned done = 0;
  long long t0 = clock64();
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (!done && clock64() - t0 > 2000000000LL) break;   // bounded: a protocol bug must not hang the GPU
  }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(unsigned* slot_smem) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "n"(kTmemCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned taddr) {       // the same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kTmemCols) : "memory");
}

__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
  unsigned h, l;
  split_tf32(v.x, h, l); hi.x = __uint_as_float(h); lo.x = __uint_as_float(l);
  split_tf32(v.y, h, l); hi.y = __uint_as_float(h); lo.y = __uint_as_float(l);
  split_tf32(v.z, h, l); hi.z = __uint_as_float(h); lo.z = __uint_as_float(l);
  split_tf32(v.w, h, l); hi.w = __uint_as_float(h); lo.w = __uint_as_float(l);
}

// raw ring stage -> canonical hi / lo operand tiles. A quarter-warp (8 lanes) always writes one
// contiguous 128-byte core matrix, and reads 8 different raw rows whose strides (68 / 36 / BN+4
// floats) put them in disjoint bank groups: no shared-memory conflicts on either side.
template <int BN, int AK, int BK_>
__device__ __forceinline__ void tc_convert(const float* sA, const float* sB, unsigned char* ob) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r8 = lane & 7, q = lane >> 3;           // row-in-core-matrix, quarter
  float4 hi, lo;
  // ---- A: 64 (m) x 32 (k) = 8 core-matrix rows x 64 groups; 512 chunks, 2 per thread
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int grp = warp * 8 + q * 2 + i;           // 0..63
    if (AK == A_MCONTIG) {                          // raw [kk][68]; canonical MN-major [kg][mg][k%8][4 m]
      const int kg = grp >> 4, mg = grp & 15;
      const float4 v = *reinterpret_cast<const float4*>(sA + (kg * 8 + r8) * kLdaM + mg * 4);
      split4(v, hi, lo);
      const int off = kg * 2048 + mg * 128 + r8 * 16;
      *reinterpret_cast<float4*>(ob + off) = hi;
      *reinterpret_cast<float4*>(ob + kOpTileBytes + off) = lo;
    } else {                                        // raw [mm][36]; canonical K-major [mg8][kc][m%8][4 k]
      const int mg = grp >> 3, kc = grp & 7;
      const float4 v = *reinterpret_cast<const float4*>(sA + (mg * 8 + r8) * kLdK + kc * 4);
      split4(v, hi, lo);
      const int off = mg * 1024 + kc * 128 + r8 * 16;
      *reinterpret_cast<float4*>(ob + off) = hi;
      *reinterpret_cast<float4*>(ob + kOpTileBytes + off) = lo;
    }
  }
  // ---- B: BN (n) x 32 (k): BN groups of 8 chunks
  unsigned char* obB = ob + 2 * kOpTileBytes;
#pragma unroll
  for (int i = 0; i < (BN + 31) / 32; ++i) {
    const int grp = warp * 4 + q + 32 * i;          // 0..BN-1 valid
    if (grp < BN) {
      if (BK_ == B_KCONTIG) {                       // raw [nn][36]; canonical K-major [ng8][kc][n%8][4 k]
        const int ng = grp >> 3, kc = grp & 7;
        const float4 v = *reinterpret_cast<const float4*>(sB + (ng * 8 + r8) * kLdK + kc * 4);
        split4(v, hi, lo);
        const int off = ng * 1024 + kc * 128 + r8 * 16;
        *reinterpret_cast<float4*>(obB + off) = hi;
        *reinterpret_cast<float4*>(obB + kOpTileBytes + off) = lo;
      } else {                                      // raw [kk][BN+4]; canonical MN-major [kg][ng4][k%8][4 n]
        constexpr int NG = BN / 4;
        const int kg = grp / NG, ng = grp % NG;
        const float4 v = *reinterpret_cast<const float4*>(sB + (kg * 8 + r8) * (BN + 4) + ng * 4);
        split4(v, hi, lo);
        const int off = kg * (NG * 128) + ng * 128 + r8 * 16;
        *reinterpret_cast<float4*>(obB + off) = hi;
        *reinterpret_cast<float4*>(obB + kOpTileBytes + off) = lo;
      }
    }
  }
}

// acc[i][j] = sum_{k in [kbeg, kend)} A(m0 + tx*4 + i, k) * B(n0 + ty + 16 j, k)   (acc is overwritten)
template <int BN, int AK, int BK_>
__device__ __forceinline__ void tile_gemm_tc(float (&acc)[4][BN / 16], const OperandA& a, const OperandB& b, int kbeg,
                                             int kend, float* smem, const TcState& tc) {
  constexpr int TN = BN / 16;
  constexpr int LDC = 68;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nkt = (kend - kbeg + kTK - 1) / kTK;
  int rowsk[2] = {0, 0};
  int rows[kStages][2];
  if (AK == A_KCONTIG) {
    fetch_rows_k(a, rowsk);
#pragma unroll
    for (int s = 0; s < kStages; ++s) rows[s][0] = rows[s][1] = 0;
  } else {
#pragma unroll
    for (int s = 0; s < kStages; ++s) fetch_rows(a, kbeg + s * kTK, kend, rows[s]);
  }
  if (threadIdx.x == 0) {
    mbar_init(tc.bars, 1);
    mbar_init(tc.bars + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();  // ring and operand buffers are free (previous tile finished), barriers initialised
#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    if (s < nkt) {
      issue_A<AK>(smem + s * kStageFloats, a, kbeg + s * kTK, kend, rows[s], rowsk);
      issue_B<BN, BK_>(smem + s * kStageFloats + kStageFloatsA, b, kbeg + s * kTK, kend);
    }
    cp_async_commit();
  }
  int r_use[2] = {rows[kStages - 1][0], rows[kStages - 1][1]};
  const unsigned idesc = umma_idesc(BN, AK == A_MCONTIG ? 1 : 0, BK_ == B_NCONTIG ? 1 : 0);
  for (int t = 0; t < nkt; ++t) {
    cp_async_wait<kStages - 2>();
    __syncthreads();                     // raw tile t visible to everyone; ring stage of tile t-1 is free
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
    const int bsel = t & 1;
    unsigned char* ob = tc.opbuf + bsel * kOpBufBytes;
    if (t >= 2) mbar_wait(tc.bars + bsel, ((t >> 1) - 1) & 1);   // the MMAs that read this buffer are done
    const float* sA = smem + (t % kStages) * kStageFloats;
    tc_convert<BN, AK, BK_>(sA, sA + kStageFloatsA, ob);
    fence_async_smem();                  // generic-proxy writes -> visible to the tensor core (async proxy)
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      const unsigned a_hi = smem_u32(ob), a_lo = a_hi + kOpTileBytes, b_hi = a_hi + 2 * kOpTileBytes,
                     b_lo = a_hi + 3 * kOpTileBytes;
#pragma unroll
      for (int ks = 0; ks < kTK / 8; ++ks) {
        // per k-step (8 k = 32 bytes): MN-major operands advance by one 8-k group, K-major by two 16-byte chunks
        const unsigned a_off = (AK == A_MCONTIG) ? ks * 2048 : ks * 256;
        const unsigned a_lbo = (AK == A_MCONTIG) ? 2048 : 128, a_sbo = (AK == A_MCONTIG) ? 128 : 1024;
        const unsigned b_off = (BK_ == B_KCONTIG) ? ks * 256 : ks * (BN / 4) * 128;
        const unsigned b_lbo = (BK_ == B_KCONTIG) ? 128 : (BN / 4) * 128, b_sbo = (BK_ == B_KCONTIG) ? 1024 : 128;
        const unsigned long long dah = umma_desc(a_hi + a_off, a_lbo, a_sbo), dal = umma_desc(a_lo + a_off, a_lbo, a_sbo);
        const unsigned long long dbh = umma_desc(b_hi + b_off, b_lbo, b_sbo), dbl = umma_desc(b_lo + b_off, b_lbo, b_sbo);
        umma_tf32(tc.tmem, dal, dbh, idesc, (t | ks) ? 1u : 0u);   // small terms first
        umma_tf32(tc.tmem, dah, dbl, idesc, 1u);
        umma_tf32(tc.tmem, dah, dbh, idesc, 1u);
      }
      umma_commit(tc.bars + bsel);
    }
  }
  cp_async_wait<0>();
  if (nkt > 0) mbar_wait(tc.bars + ((nkt - 1) & 1), ((nkt - 1) >> 1) & 1);   // all MMAs complete (commits are ordered)
  tc_fence_after();
  __syncthreads();
  // TMEM (lane = tile row, column = tile column) -> [n][68] staging -> the epilogue's ownership
  float* sC = smem;
  if (warp < 2 && nkt > 0) {
    const int m = warp * 32 + lane;
#pragma unroll
    for (int c0 = 0; c0 < BN; c0 += 8) {
      unsigned v[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(tc.tmem + ((unsigned)(warp * 32) << 16) + (unsigned)c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int u = 0; u < 8; ++u) sC[(c0 + u) * LDC + m] = __uint_as_float(v[u]);
    }
    tc_fence_before();
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    if (nkt > 0) {
      const float4 v = *reinterpret_cast<const float4*>(sC + (ty + 16 * j) * LDC + tx * 4);
      acc[0][j] = v.x; acc[1][j] = v.y; acc[2][j] = v.z; acc[3][j] = v.w;
    } else {
      acc[0][j] = acc[1][j] = acc[2][j] = acc[3][j] = 0.f;
    }
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
