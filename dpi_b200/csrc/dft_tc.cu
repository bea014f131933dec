// dft_tc.cu - tcgen05 GEMM for the dense DFT and its adjoint (see dft_tc.cuh).
#include <algorithm>

#include "dft_tc.cuh"
#include "tile_cp.cuh"   // UMMA descriptors, mbarrier / tcgen05 wrappers (umma_desc, umma_idesc, umma_tf32, ...)

namespace dpi {
namespace tc {
namespace {

__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Canonical K-major no-swizzle position (in floats) of element (r, kk) of an [R x 32] fp32 tile:
// 8-row group stride 1024 B (SBO), 4-k chunk stride 128 B (LBO), row stride 16 B inside a core matrix.
__host__ __device__ __forceinline__ int canon(int r, int kk) { return (r >> 3) * 256 + (kk >> 2) * 32 + (r & 7) * 4 + (kk & 3); }

// dst: [r_tiles][k_tiles][hi, lo][R x 32] <- src[row * s_r + k * s_k], zero outside [0, rows) x [0, K)
template <int R>
__global__ void pack_kernel(const float* __restrict__ src, long long s_r, long long s_k, int rows, int K, int r_tiles,
                            int k_tiles, float* __restrict__ dst) {
  const long long total = (long long)r_tiles * k_tiles * (R * kTK);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int within = (int)(e % (R * kTK));
    const long long tile = e / (R * kTK);
    const int kt = (int)(tile % k_tiles), rt = (int)(tile / k_tiles);
    // decode the canonical position so that consecutive threads write consecutive floats
    const int rg = within >> 8, rem = within & 255, kc = rem >> 5, rr = (rem & 31) >> 2, k4 = rem & 3;
    const int row = rt * R + rg * 8 + rr, k = kt * kTK + kc * 4 + k4;
    const float x = (row < rows && k < K) ? src[(long long)row * s_r + (long long)k * s_k] : 0.f;
    unsigned hi, lo;
    split_tf32(x, hi, lo);
    float* t = dst + tile * (2 * R * kTK);
    t[within] = __uint_as_float(hi);
    t[R * kTK + within] = __uint_as_float(lo);
  }
}

__global__ void __launch_bounds__(128) dft_tc_kernel(Plan p, const float* __restrict__ Apack, const float* __restrict__ Bpack,
                                                     float* __restrict__ part) {
  extern __shared__ unsigned char raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) unsigned long long bars[2 * kStagesTc + 1];
  __shared__ unsigned tmem_slot;
  unsigned long long* full = bars;
  unsigned long long* empty = bars + kStagesTc;
  unsigned long long* accum = bars + 2 * kStagesTc;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int t = blockIdx.x;
  const int sp = t % p.splits; t /= p.splits;
  const int nt = t % p.n_tiles, mt = t / p.n_tiles;
  const int kt0 = sp * p.kt_per_split, kt1 = min(p.k_tiles, kt0 + p.kt_per_split), nk = max(0, kt1 - kt0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStagesTc; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    mbar_init(accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem = tmem_slot;
  if (warp == 0 && lane == 0) {
    // ---- producer: two bulk copies per stage (A_hi|A_lo = 32 KB, B_hi|B_lo = 16 KB)
    for (int i = 0; i < nk; ++i) {
      const int s = i % kStagesTc;
      if (i >= kStagesTc) mbar_wait(empty + s, ((i / kStagesTc) - 1) & 1);
      mbar_expect_tx(full + s, kStageBytes);
      unsigned char* st = sm + s * kStageBytes;
      bulk_g2s(st, Apack + ((size_t)mt * p.k_tiles + kt0 + i) * (2 * kATileFloats), 2 * kATileFloats * 4, full + s);
      bulk_g2s(st + 2 * kATileFloats * 4, Bpack + ((size_t)nt * p.k_tiles + kt0 + i) * (2 * kBTileFloats),
               2 * kBTileFloats * 4, full + s);
    }
  } else if (warp == 1 && lane == 0) {
    // ---- MMA issuer: 4 k-steps x (a_lo b_hi + a_hi b_lo + a_hi b_hi) per stage, small terms first
    const unsigned idesc = umma_idesc(kTN, 0, 0);
    for (int i = 0; i < nk; ++i) {
      const int s = i % kStagesTc;
      mbar_wait(full + s, (i / kStagesTc) & 1);
      tc_fence_after();
      const unsigned a_hi = smem_u32(sm + s * kStageBytes), a_lo = a_hi + kATileFloats * 4;
      const unsigned b_hi = a_hi + 2 * kATileFloats * 4, b_lo = b_hi + kBTileFloats * 4;
#pragma unroll
      for (int ks = 0; ks < kTK / 8; ++ks) {
        const unsigned off = ks * 256;      // 8 k = two 16-byte chunks
        const unsigned long long dah = umma_desc(a_hi + off, 128, 1024), dal = umma_desc(a_lo + off, 128, 1024);
        const unsigned long long dbh = umma_desc(b_hi + off, 128, 1024), dbl = umma_desc(b_lo + off, 128, 1024);
        umma_tf32(tmem, dal, dbh, idesc, (i | ks) ? 1u : 0u);
        umma_tf32(tmem, dah, dbl, idesc, 1u);
        umma_tf32(tmem, dah, dbh, idesc, 1u);
      }
      umma_commit(empty + s);               // stage reusable once these MMAs have read it
    }
    umma_commit(accum);                     // accumulator complete
  }
  // ---- epilogue: TMEM (lane = tile row, column = tile column) -> partial tile [128][64] in global memory
  if (nk > 0) {
    mbar_wait(accum, 0);
    tc_fence_after();
  }
  __syncwarp();
  float* dst = part + (((size_t)sp * p.m_tiles + mt) * p.n_tiles + nt) * (size_t)(kTM * kTN) + (size_t)(warp * 32 + lane) * kTN;
#pragma unroll
  for (int c0 = 0; c0 < kTN; c0 += 16) {
    unsigned v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) v[u] = 0u;
    if (nk > 0) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(tmem + ((unsigned)(warp * 32) << 16) + (unsigned)c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
#pragma unroll
    for (int u = 0; u < 16; u += 4)
      *reinterpret_cast<float4*>(dst + c0 + u) = make_float4(__uint_as_float(v[u]), __uint_as_float(v[u + 1]),
                                                             __uint_as_float(v[u + 2]), __uint_as_float(v[u + 3]));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem);
}

// out <- sum over the K splits (in split order) of the partial tiles
__global__ void reduce_kernel(Plan p, const float* __restrict__ part, float* __restrict__ out, int mode) {
  const long long total = (long long)p.M * p.N;
  const size_t tile = (size_t)kTM * kTN;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(e / p.N), n = (int)(e % p.N);
    const size_t base = ((size_t)(m / kTM) * p.n_tiles + n / kTN) * tile + (size_t)(m % kTM) * kTN + n % kTN;
    const size_t sstride = (size_t)p.m_tiles * p.n_tiles * tile;
    float acc = 0.f;
    for (int s = 0; s < p.splits; ++s) acc += __ldcg(part + base + s * sstride);
    if (mode == 0) out[(size_t)m * p.N + (size_t)(n & 1) * (p.N / 2) + (n >> 1)] = acc;
    else out[(size_t)m * p.N + n] = acc;
  }
}

}  // namespace

Plan make_plan(int M, int N, int K, int n_sm) {
  Plan p{};
  p.M = M; p.N = N; p.K = K;
  p.m_tiles = (M + kTM - 1) / kTM;
  p.n_tiles = (N + kTN - 1) / kTN;
  p.k_tiles = (K + kTK - 1) / kTK;
  const int mn = p.m_tiles * p.n_tiles;
  int want = std::max(1, n_sm / std::max(1, mn));
  want = std::min(std::min(want, std::max(1, p.k_tiles / 2)), 16);
  p.kt_per_split = (p.k_tiles + want - 1) / want;
  p.splits = (p.k_tiles + p.kt_per_split - 1) / p.kt_per_split;
  return p;
}
size_t a_pack_floats(const Plan& p) { return (size_t)p.m_tiles * p.k_tiles * 2 * kATileFloats; }
size_t b_pack_floats(int N, int K) { return (size_t)((N + kTN - 1) / kTN) * ((K + kTK - 1) / kTK) * 2 * kBTileFloats; }
size_t part_floats(const Plan& p) { return (size_t)p.splits * p.m_tiles * p.n_tiles * kTM * kTN; }

int init() {
  DPI_CUDA(cudaFuncSetAttribute((const void*)dft_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  return DPI_OK;
}

int pack_b(const float* src, long long s_n, long long s_k, int N, int K, float* dst, cudaStream_t st) {
  const int n_tiles = (N + kTN - 1) / kTN, k_tiles = (K + kTK - 1) / kTK;
  const long long total = (long long)n_tiles * k_tiles * kBTileFloats;
  pack_kernel<kTN><<<(unsigned)std::min<long long>((total + 255) / 256, 4096), 256, 0, st>>>(src, s_n, s_k, N, K, n_tiles,
                                                                                             k_tiles, dst);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int gemm(const Plan& p, const float* A, int lda, const float* Bpack, float* a_pack, float* part, float* out, int mode,
         cudaStream_t st) {
  const long long total = (long long)p.m_tiles * p.k_tiles * kATileFloats;
  pack_kernel<kTM><<<(unsigned)std::min<long long>((total + 255) / 256, 8192), 256, 0, st>>>(A, lda, 1, p.M, p.K, p.m_tiles,
                                                                                             p.k_tiles, a_pack);
  DPI_LAUNCH_CHECK();
  dft_tc_kernel<<<p.m_tiles * p.n_tiles * p.splits, 128, kSmemBytes, st>>>(p, a_pack, Bpack, part);
  DPI_LAUNCH_CHECK();
  const long long out_n = (long long)p.M * p.N;
  reduce_kernel<<<(unsigned)std::min<long long>((out_n + 255) / 256, 8192), 256, 0, st>>>(p, part, out, mode);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

}  // namespace tc
}  // namespace dpi
