// flow_tc.cu - tcgen05 GEMMs of the coupling nets (see flow_tc.cuh).
#include <algorithm>
#include <cstdlib>

#include "flow_tc.cuh"
#include "tile_cp.cuh"   // umma_desc / umma_idesc / umma_tf32 / mbarrier wrappers, split_tf32

namespace dpi {
namespace ftc {
namespace {

__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc_n(unsigned* slot_smem) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_n(unsigned taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// ------------------------------------------------------------------------------------------- pack
// One CTA per R x 32 operand tile: coalesced reads in whichever direction the source is contiguous, transposed
// through shared memory, TF32 hi / lo split, float4 stores of the canonical K-major tile (position of element
// (r, kk) in floats: (r >> 3) * 256 + (kk >> 2) * 32 + (r & 7) * 4 + (kk & 3)).
template <int R>
__global__ void __launch_bounds__(256) pack_kernel(Pack p, int k_tiles, float* __restrict__ dst) {
  __shared__ float s[kK][R + 1];
  const int kt = blockIdx.x % k_tiles, rt = blockIdx.x / k_tiles;
  const int r0 = rt * R, k0 = kt * kK;
  const int tid = threadIdx.x;
  if (p.k_contig) {
    const int lane = tid & 31, w = tid >> 5;
    const int k = k0 + lane;
#pragma unroll 8
    for (int r = w; r < R; r += 8) {
      const int row = r0 + r;
      float v = 0.f;
      if (row < p.rows && k < p.K) {
        const int rr = p.map ? __ldg(p.map + row) : row;
        v = __ldg(p.src + (long long)rr * p.ld + k);
      }
      s[lane][r] = v;
    }
  } else {
#pragma unroll 8
    for (int i = tid; i < kK * R; i += 256) {
      const int kk = i / R, r = i % R;
      const int row = r0 + r, k = k0 + kk;
      float v = 0.f;
      if (row < p.rows && k < p.K) {
        const int kr = p.map ? __ldg(p.map + k) : k;
        v = __ldg(p.src + (long long)kr * p.ld + row);
      }
      s[kk][r] = v;
    }
  }
  __syncthreads();
  float* t = dst + (size_t)blockIdx.x * (2 * R * kK);
  for (int q = tid; q < R * kK / 4; q += 256) {
    const int within = q * 4;
    const int rg = within >> 8, rem = within & 255, kc = rem >> 5, rr = (rem & 31) >> 2;
    const int r = rg * 8 + rr, kk = kc * 4;
    unsigned h[4], l[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) split_tf32(s[kk + u][r], h[u], l[u]);
    *reinterpret_cast<float4*>(t + within) =
        make_float4(__uint_as_float(h[0]), __uint_as_float(h[1]), __uint_as_float(h[2]), __uint_as_float(h[3]));
    *reinterpret_cast<float4*>(t + R * kK + within) =
        make_float4(__uint_as_float(l[0]), __uint_as_float(l[1]), __uint_as_float(l[2]), __uint_as_float(l[3]));
  }
}

// ------------------------------------------------------------------------------------------- GEMM
// 128 threads. Lane 0 of warp 0: producer (two bulk copies per stage); lane 0 of warp 1: MMA issuer; all four
// warps: TMEM -> partial tile. transposed != 0 stores the tile as [n][m] (feature-major consumers read it
// coalesced along samples), else as [m][n].
template <int TN>
__global__ void __launch_bounds__(128) gemm_kernel(Gemm g, const float* __restrict__ Apack, const float* __restrict__ Bpack,
                                                   float* __restrict__ part, int transposed) {
  constexpr int kATile = kM * kK, kBTile = TN * kK;
  constexpr unsigned kStageBytes = (2 * kATile + 2 * kBTile) * 4;
  extern __shared__ unsigned char raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) unsigned long long bars[2 * kMaxStages + 1];
  __shared__ unsigned tmem_slot;
  const int S = g.stages;
  unsigned long long* full = bars;
  unsigned long long* empty = bars + kMaxStages;
  unsigned long long* accum = bars + 2 * kMaxStages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int t = blockIdx.x;
  const int sp = t % g.splits; t /= g.splits;
  const int nt = t % g.n_tiles, mt = t / g.n_tiles;
  const int kt0 = sp * g.kt_per_split, kt1 = min(g.k_tiles, kt0 + g.kt_per_split), nk = max(0, kt1 - kt0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    mbar_init(accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc_n<TN>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem = tmem_slot;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < nk; ++i) {
      const int s = i % S;
      if (i >= S) mbar_wait(empty + s, ((i / S) - 1) & 1);
      mbar_expect_tx(full + s, kStageBytes);
      unsigned char* st = sm + (size_t)s * kStageBytes;
      bulk_g2s(st, Apack + ((size_t)mt * g.k_tiles + kt0 + i) * (2 * kATile), 2 * kATile * 4, full + s);
      bulk_g2s(st + 2 * kATile * 4, Bpack + ((size_t)nt * g.k_tiles + kt0 + i) * (2 * kBTile), 2 * kBTile * 4, full + s);
    }
  } else if (warp == 1 && lane == 0) {
    const unsigned idesc = umma_idesc(TN, 0, 0);
    for (int i = 0; i < nk; ++i) {
      const int s = i % S;
      mbar_wait(full + s, (i / S) & 1);
      tc_fence_after();
      const unsigned a_hi = smem_u32(sm + (size_t)s * kStageBytes), a_lo = a_hi + kATile * 4;
      const unsigned b_hi = a_hi + 2 * kATile * 4, b_lo = b_hi + kBTile * 4;
#pragma unroll
      for (int ks = 0; ks < kK / 8; ++ks) {
        const unsigned off = ks * 256;      // 8 k = two 16-byte chunks
        const unsigned long long dah = umma_desc(a_hi + off, 128, 1024), dal = umma_desc(a_lo + off, 128, 1024);
        const unsigned long long dbh = umma_desc(b_hi + off, 128, 1024), dbl = umma_desc(b_lo + off, 128, 1024);
        umma_tf32(tmem, dal, dbh, idesc, (i | ks) ? 1u : 0u);   // small terms first
        umma_tf32(tmem, dah, dbl, idesc, 1u);
        umma_tf32(tmem, dah, dbh, idesc, 1u);
      }
      umma_commit(empty + s);
    }
    umma_commit(accum);
  }
  if (nk > 0) {
    mbar_wait(accum, 0);
    tc_fence_after();
  }
  __syncwarp();
  const int row = warp * 32 + lane;
  float* dst = part + (((size_t)sp * g.m_tiles + mt) * g.n_tiles + nt) * (size_t)(kM * TN);
#pragma unroll
  for (int c0 = 0; c0 < TN; c0 += 16) {
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
    if (transposed) {
#pragma unroll
      for (int u = 0; u < 16; ++u) __stcg(dst + (size_t)(c0 + u) * kM + row, __uint_as_float(v[u]));
    } else {
#pragma unroll
      for (int u = 0; u < 16; u += 4)
        __stcg(reinterpret_cast<float4*>(dst + (size_t)row * TN + c0 + u),
               make_float4(__uint_as_float(v[u]), __uint_as_float(v[u + 1]), __uint_as_float(v[u + 2]),
                           __uint_as_float(v[u + 3])));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc_n<TN>(tmem);
}

// ------------------------------------------------------------------------------------------- epilogue
// One CTA per output row: a feature row n (all samples m; transposed partial tiles, coalesced along m) for the
// feature-major modes, a gradient row for EPI_ROWMAJOR. Split partials are summed in split order.
__global__ void __launch_bounds__(256) epi_kernel(Gemm g, const float* __restrict__ part, Epi e) {
  __shared__ float red[32];
  const size_t tile = (size_t)kM * g.TN;
  const size_t sstride = (size_t)g.m_tiles * g.n_tiles * tile;
  const int tid = threadIdx.x;
  if (e.mode == EPI_ROWMAJOR) {
    const int m = blockIdx.x;
    const float* row = part + (size_t)(m / kM) * g.n_tiles * tile + (size_t)(m % kM) * g.TN;
    for (int n = tid; n < g.N; n += 256) {
      const float* p = row + (size_t)(n / g.TN) * tile + n % g.TN;
      float c = 0.f;
      for (int s = 0; s < g.splits; ++s) c += __ldcg(p + s * sstride);
      __stcg(e.out + (long long)m * e.ldo + n, c);
    }
    return;
  }
  const int n = blockIdx.x;
  const float* row = part + (size_t)(n / g.TN) * tile + (size_t)(n % g.TN) * kM;
  auto fetch = [&](int m) {
    const float* p = row + (size_t)(m >> 7) * g.n_tiles * tile + (m & 127);
    float c = 0.f;
    for (int s = 0; s < g.splits; ++s) c += __ldcg(p + s * sstride);
    return c;
  };
  constexpr int kPer = kEpiMaxM / 256;
  if (e.mode == EPI_T_RAW) {
    for (int m = tid; m < g.M; m += 256) __stcg(e.out + (long long)n * e.ldo + m, fetch(m));
  } else if (e.mode == EPI_T_SCATTER) {
    const float sc = expf(__ldg(e.lsi));
    const int r = __ldg(e.rowmap + n);
    for (int m = tid; m < g.M; m += 256)
      __stcg(e.out + (long long)r * e.ldo + m, fmaf(__ldcg(e.add + (long long)n * e.ldo + m), sc, fetch(m)));
  } else if (e.mode == EPI_T_HID) {
    const float bias = __ldg(e.bias + n);
    if (e.bn_mode == 1) {   // BatchNorm1d in training mode (realnvpfc_model.py:178,185): biased batch variance, eps 1e-2
      float v[kPer];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < kPer; ++i) {
        const int m = tid + 256 * i;
        v[i] = 0.f;
        if (m < g.M) { v[i] = lrelu_f(fetch(m) + bias); s += v[i]; }
      }
      const float invB = 1.0f / (float)g.M;
      const float mean = block_sum(s, red) * invB;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < kPer; ++i)
        if (tid + 256 * i < g.M) { const float d = v[i] - mean; q = fmaf(d, d, q); }
      const float var = block_sum(q, red) * invB;
      const float rstd = 1.0f / sqrtf(var + kBnEps);
      const float ga = __ldg(e.gamma + n) * rstd, be = __ldg(e.beta + n);
#pragma unroll
      for (int i = 0; i < kPer; ++i) {
        const int m = tid + 256 * i;
        if (m < g.M) {
          __stcg(e.out + (long long)n * e.ldo + m, v[i]);
          __stcg(e.out2 + (long long)n * e.ldo + m, fmaf(v[i] - mean, ga, be));
        }
      }
      if (tid == 0) {
        e.mean[n] = mean;
        e.rstd[n] = rstd;
        e.rmean[n] = (1.f - kBnMomentum) * e.rmean[n] + kBnMomentum * mean;
        e.rvar[n] = (1.f - kBnMomentum) * e.rvar[n] + kBnMomentum * (var * (float)g.M / (float)(g.M - 1));
      }
    } else {
      float ga = 1.f, be = 0.f;
      if (e.bn_mode == 2) {   // eval: running statistics
        const float rs = 1.0f / sqrtf(e.rvar[n] + kBnEps);
        ga = __ldg(e.gamma + n) * rs;
        be = __ldg(e.beta + n) - e.rmean[n] * ga;
      }
      for (int m = tid; m < g.M; m += 256) {
        const float v = lrelu_f(fetch(m) + bias);
        __stcg(e.out + (long long)n * e.ldo + m, v);
        if (e.out2) __stcg(e.out2 + (long long)n * e.ldo + m, e.bn_mode == 2 ? fmaf(v, ga, be) : v);
      }
    }
  } else {   // EPI_T_BNBWD
    float gv[kPer], l[kPer];
    float mean = 0.f, rstd = 1.f, gam = 1.f, s1 = 0.f, s2 = 0.f;
    if (e.bn_mode) { mean = e.mean[n]; rstd = e.rstd[n]; gam = __ldg(e.gamma + n); }
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int m = tid + 256 * i;
      gv[i] = 0.f; l[i] = 0.f;
      if (m < g.M) {
        gv[i] = fetch(m);
        l[i] = __ldcg(e.lsaved + (long long)n * e.ldo + m);
        s1 += gv[i];
        s2 = fmaf(gv[i], (l[i] - mean) * rstd, s2);
      }
    }
    float gbeta = 0.f, ggamma = 0.f;
    if (e.bn_mode) { gbeta = block_sum(s1, red); ggamma = block_sum(s2, red); }
    const float invB = 1.0f / (float)g.M;
    float s3 = 0.f;
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int m = tid + 256 * i;
      if (m < g.M) {
        float gl = gv[i];
        if (e.bn_mode) gl = gam * rstd * (gv[i] - gbeta * invB - (l[i] - mean) * rstd * ggamma * invB);
        const float v = gl * (l[i] > 0.f ? 1.f : kLreluSlope);
        __stcg(e.out + (long long)n * e.ldo + m, v);
        s3 += v;
      }
    }
    s3 = block_sum(s3, red);
    if (tid == 0) {
      if (e.bn_mode) { e.g_gamma[n] = ggamma; e.g_beta[n] = gbeta; }
      e.g_bias[n] = s3;
    }
  }
}

int g_force_tn = 0;

template <int TN>
constexpr int stages_for() { return TN == 64 ? 4 : TN == 128 ? 3 : 2; }
template <int TN>
constexpr int smem_for() { return stages_for<TN>() * (2 * kM * kK + 2 * TN * kK) * 4 + 1024; }

}  // namespace

// Column tile: the GEMM is fed from L2 (hi + lo copies of both operands), so the widest tile that still fills the
// GPU (with split-K over at least 4 k-tiles per split) wins; DPI_B200_TC_TN forces 64 / 128 / 256.
Gemm make_gemm(int M, int N, int K, int n_sm) {
  Gemm g{};
  g.M = M; g.N = N; g.K = K;
  g.m_tiles = (M + kM - 1) / kM;
  g.k_tiles = (K + kK - 1) / kK;
  const int max_splits = std::min(16, std::max(1, g.k_tiles / 4));
  const int cands[3] = {256, 128, 64};
  g.TN = 64;
  for (int c = 0; c < 3; ++c) {
    const int tn = cands[c];
    if (g_force_tn) { g.TN = g_force_tn; break; }
    if (tn > 64 && N < tn) continue;
    g.TN = tn;
    if ((long long)g.m_tiles * ((N + tn - 1) / tn) * max_splits >= (3LL * n_sm) / 4) break;
  }
  g.n_tiles = (N + g.TN - 1) / g.TN;
  const int mn = g.m_tiles * g.n_tiles;
  int want = std::max(1, (n_sm + mn - 1) / std::max(1, mn));
  if (mn >= (3 * n_sm) / 4) want = 1;
  want = std::min(want, max_splits);
  g.kt_per_split = (g.k_tiles + want - 1) / want;
  g.splits = (g.k_tiles + g.kt_per_split - 1) / g.kt_per_split;
  g.stages = g.TN == 64 ? stages_for<64>() : g.TN == 128 ? stages_for<128>() : stages_for<256>();
  return g;
}

int init() {
  static int done = 0;
  if (done) return DPI_OK;
  DPI_CUDA(cudaFuncSetAttribute((const void*)gemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_for<64>()));
  DPI_CUDA(cudaFuncSetAttribute((const void*)gemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_for<128>()));
  DPI_CUDA(cudaFuncSetAttribute((const void*)gemm_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_for<256>()));
  done = 1;
  return DPI_OK;
}

void configure() {
  const char* e = getenv("DPI_B200_TC_TN");
  const int v = e ? atoi(e) : 0;
  g_force_tn = (v == 64 || v == 128 || v == 256) ? v : 0;
}

int pack(const Pack& p, int R, float* dst, cudaStream_t st) {
  const int r_tiles = (p.rows + R - 1) / R, k_tiles = (p.K + kK - 1) / kK;
  if (R == 128) pack_kernel<128><<<r_tiles * k_tiles, 256, 0, st>>>(p, k_tiles, dst);
  else if (R == 256) pack_kernel<256><<<r_tiles * k_tiles, 256, 0, st>>>(p, k_tiles, dst);
  else if (R == 64) pack_kernel<64><<<r_tiles * k_tiles, 256, 0, st>>>(p, k_tiles, dst);
  else { set_error("ftc::pack: unsupported tile rows %d", R); return DPI_ERR_INVALID; }
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int run(const Gemm& g, const float* a_pack, const float* b_pack, float* part, const Epi& e, cudaStream_t st) {
  const int grid = g.m_tiles * g.n_tiles * g.splits, tr = e.mode != EPI_ROWMAJOR;
  if (g.TN == 64) gemm_kernel<64><<<grid, 128, smem_for<64>(), st>>>(g, a_pack, b_pack, part, tr);
  else if (g.TN == 128) gemm_kernel<128><<<grid, 128, smem_for<128>(), st>>>(g, a_pack, b_pack, part, tr);
  else gemm_kernel<256><<<grid, 128, smem_for<256>(), st>>>(g, a_pack, b_pack, part, tr);
  DPI_LAUNCH_CHECK();
  if ((e.mode == EPI_T_BNBWD || (e.mode == EPI_T_HID && e.bn_mode == 1)) && g.M > kEpiMaxM) {
    set_error("ftc::run: fused BatchNorm epilogue supports at most %d samples (got %d)", kEpiMaxM, g.M);
    return DPI_ERR_UNSUPPORTED;
  }
  epi_kernel<<<e.mode == EPI_ROWMAJOR ? g.M : g.N, 256, 0, st>>>(g, part, e);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

}  // namespace ftc
}  // namespace dpi
