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
// One CTA per R x KT operand tile: coalesced reads in whichever direction the source is contiguous, transposed
// through shared memory, TF32 hi / lo split, float4 stores of the canonical K-major tile (position of element
// (r, kk) in floats: (r >> 3) * 8 KT + (kk >> 2) * 32 + (r & 7) * 4 + (kk & 3)).
template <int R, int KT>
__device__ __forceinline__ void pack_tile(const Pack& p, int k_tiles, int vec, float* __restrict__ dst, int block,
                                          float (*s)[R + 1]) {
  const int kt = block % k_tiles, rt = block / k_tiles;
  const int r0 = rt * R, k0 = kt * KT;
  const int tid = threadIdx.x;
  constexpr int NI = R * KT / 1024;   // float4 items per thread
  static_assert(NI >= 1, "tile too small for 256 threads");
  if (vec) {
    // every load of the tile is issued before the first shared-memory store: the CTA is one latency, not NI of them
    float4 v[NI];
    if (p.k_contig) {
#pragma unroll
      for (int u = 0; u < NI; ++u) {
        const int i = tid + 256 * u, r = i / (KT / 4), k4 = (i % (KT / 4)) * 4;
        const int row = r0 + r, k = k0 + k4;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < p.rows && k < p.K) {
          const int rr = p.map ? __ldg(p.map + row) : row;
          v[u] = __ldg(reinterpret_cast<const float4*>(p.src + (long long)rr * p.ld + k));
        }
      }
#pragma unroll
      for (int u = 0; u < NI; ++u) {
        const int i = tid + 256 * u, r = i / (KT / 4), k4 = (i % (KT / 4)) * 4;
        s[k4][r] = v[u].x; s[k4 + 1][r] = v[u].y; s[k4 + 2][r] = v[u].z; s[k4 + 3][r] = v[u].w;
      }
    } else {
#pragma unroll
      for (int u = 0; u < NI; ++u) {
        const int i = tid + 256 * u, kk = i / (R / 4), r4 = (i % (R / 4)) * 4;
        const int row = r0 + r4, k = k0 + kk;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < p.rows && k < p.K) {
          const int kr = p.map ? __ldg(p.map + k) : k;
          v[u] = __ldg(reinterpret_cast<const float4*>(p.src + (long long)kr * p.ld + row));
        }
      }
#pragma unroll
      for (int u = 0; u < NI; ++u) {
        const int i = tid + 256 * u, kk = i / (R / 4), r4 = (i % (R / 4)) * 4;
        s[kk][r4] = v[u].x; s[kk][r4 + 1] = v[u].y; s[kk][r4 + 2] = v[u].z; s[kk][r4 + 3] = v[u].w;
      }
    }
  } else if (p.k_contig) {
#pragma unroll 8
    for (int i = tid; i < KT * R; i += 256) {
      const int r = i / KT, kk = i % KT;
      const int row = r0 + r, k = k0 + kk;
      float v = 0.f;
      if (row < p.rows && k < p.K) {
        if (p.im_h > 0) {   // weight-gradient operand: row = (channel, tap) of the 3 x 3 window, k = position
          const int c = row / 9, t9 = row - 9 * c, dy = t9 / 3 - 1, dx = t9 - 3 * (t9 / 3) - 1;
          const int hw = p.im_h * p.im_w, b = k / hw, rem = k - b * hw, y = rem / p.im_w, x = rem - y * p.im_w;
          const int yy = y + dy, xx = x + dx;
          v = (yy >= 0 && yy < p.im_h && xx >= 0 && xx < p.im_w)
                  ? __ldg(p.src + (long long)c * p.ld + (long long)b * hw + yy * p.im_w + xx) : p.pad;
        } else {
          const int rr = p.map ? __ldg(p.map + row) : row;
          v = __ldg(p.src + (long long)rr * p.ld + k);
        }
      }
      s[kk][r] = v;
    }
  } else {
#pragma unroll 8
    for (int i = tid; i < KT * R; i += 256) {
      const int kk = i / R, r = i % R;
      const int row = r0 + r, k = k0 + kk;
      float v = 0.f;
      if (row < p.rows && k < p.K) {
        const int kr = p.map ? __ldg(p.map + k) : k;
        if (p.im_h > 0) {   // 3 x 3 window gather with padding
          const int c = k / 9, t9 = k - 9 * c, dy = t9 / 3 - 1, dx = t9 - 3 * (t9 / 3) - 1;
          const int hw = p.im_h * p.im_w, b = row / hw, rem = row - b * hw, y = rem / p.im_w, x = rem - y * p.im_w;
          const int yy = y + dy, xx = x + dx;
          v = (yy >= 0 && yy < p.im_h && xx >= 0 && xx < p.im_w)
                  ? __ldg(p.src + (long long)c * p.ld + (long long)b * hw + yy * p.im_w + xx) : p.pad;
        } else {
          v = __ldg(p.src + (long long)kr * p.ld + row);
        }
      }
      s[kk][r] = v;
    }
  }
  __syncthreads();
  float* t = dst + (size_t)block * (2 * R * KT);
#pragma unroll
  for (int u = 0; u < NI; ++u) {
    const int within = (tid + 256 * u) * 4;
    const int rg = within / (8 * KT), rem = within % (8 * KT), kc = rem >> 5, rr = (rem & 31) >> 2;
    const int r = rg * 8 + rr, kk = kc * 4;
    unsigned h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) split_tf32(s[kk + q][r], h[q], l[q]);
    __stcg(reinterpret_cast<float4*>(t + within),
           make_float4(__uint_as_float(h[0]), __uint_as_float(h[1]), __uint_as_float(h[2]), __uint_as_float(h[3])));
    __stcg(reinterpret_cast<float4*>(t + R * KT + within),
           make_float4(__uint_as_float(l[0]), __uint_as_float(l[1]), __uint_as_float(l[2]), __uint_as_float(l[3])));
  }
}

// Both operands of one GEMM in one launch: blocks [0, na) pack A (RA-row tiles), the rest pack B (RB-row tiles).
template <int RA, int RB, int KT>
__global__ void __launch_bounds__(256) pack2_kernel(Pack pa, int vec_a, float* __restrict__ dst_a, int na, Pack pb, int vec_b,
                                                    float* __restrict__ dst_b) {
  constexpr int RM = RB > RA ? RB : RA;
  __shared__ float s[KT * (RM + 1)];
  if ((int)blockIdx.x < na)
    pack_tile<RA, KT>(pa, (pa.K + KT - 1) / KT, vec_a, dst_a, blockIdx.x, reinterpret_cast<float (*)[RA + 1]>(s));
  else
    pack_tile<RB, KT>(pb, (pb.K + KT - 1) / KT, vec_b, dst_b, blockIdx.x - na, reinterpret_cast<float (*)[RB + 1]>(s));
}

// Weights of all stages in one launch. Every stage has the same three layer shapes, so the stage and layer of a CTA
// follow from its index; the tile shape of a layer is the B-tile of the GEMM that consumes it.
__global__ void __launch_bounds__(256) packw_kernel(WeightPack w, const float* __restrict__ P, float* __restrict__ dst) {
  __shared__ float s[32 * 257];
  const int per_stage = w.tiles[0] + w.tiles[1] + w.tiles[2];
  const int stage = blockIdx.x / per_stage;
  int t = blockIdx.x - stage * per_stage, l = 0;
  if (t >= w.tiles[0]) { t -= w.tiles[0]; l = 1; }
  if (l == 1 && t >= w.tiles[1]) { t -= w.tiles[1]; l = 2; }
  Pack p{};
  p.src = P + __ldg(w.src_off + stage * 3 + l);
  p.ld = w.ld[l]; p.k_contig = w.k_contig[l]; p.map = nullptr; p.rows = w.rows[l]; p.K = w.K[l];
  const int vec = ((reinterpret_cast<uintptr_t>(p.src) & 15) == 0 && (p.ld & 3) == 0 && ((p.k_contig ? p.K : p.rows) & 3) == 0) ? 1 : 0;
  float* d = dst + (size_t)stage * w.stage_floats + w.dst_off[l];
  const int TN = w.TN[l], KT = w.KT[l], kt = (p.K + KT - 1) / KT;
  if (TN == 64 && KT == 32) pack_tile<64, 32>(p, kt, vec, d, t, reinterpret_cast<float (*)[65]>(s));
  else if (TN == 128 && KT == 32) pack_tile<128, 32>(p, kt, vec, d, t, reinterpret_cast<float (*)[129]>(s));
  else if (TN == 256 && KT == 32) pack_tile<256, 32>(p, kt, vec, d, t, reinterpret_cast<float (*)[257]>(s));
  else pack_tile<256, 16>(p, kt, vec, d, t, reinterpret_cast<float (*)[257]>(s));
}

// ------------------------------------------------------------------------------------------- GEMM
// 128 threads. Lane 0 of warp 0: producer (two bulk copies per stage); lane 0 of warp 1: MMA issuer; all four
// warps: TMEM -> partial tile. The CTA tile is (MH x 128) x TN: MH accumulators of 128 lanes x TN columns side by
// side in TMEM share every B tile (MH = 2, TN = 256: 256 x 256 per CTA, the whole TMEM, 2/3 of the operand bytes
// per MAC of a 128 x 256 tile). transposed != 0 stores partial tiles as [n][m] (feature-major consumers read them
// coalesced along samples), else as [m][n].
template <int TN, int KT, int MH>
__global__ void __launch_bounds__(128) gemm_kernel(Gemm g, const float* __restrict__ Apack, const float* __restrict__ Bpack,
                                                   float* __restrict__ part, int transposed) {
  constexpr int kATile = MH * kM * KT, kBTile = TN * KT;
  constexpr unsigned kStageBytes = (2 * kATile + 2 * kBTile) * 4;
  constexpr unsigned kSbo = 32 * KT;           // bytes between 8-row groups of a canonical tile
  constexpr int kCols = MH * TN;
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
  const int nt = t % g.n_tiles, mt = t / g.n_tiles;          // mt counts MH x 128-row tiles
  const int kt0 = sp * g.kt_per_split, kt1 = min(g.k_tiles, kt0 + g.kt_per_split), nk = max(0, kt1 - kt0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    mbar_init(accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc_n<kCols>(&tmem_slot);
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
      for (int ks = 0; ks < KT / 8; ++ks) {
        const unsigned off = ks * 256;      // 8 k = two 16-byte chunks
        const unsigned long long dbh = umma_desc(b_hi + off, 128, kSbo), dbl = umma_desc(b_lo + off, 128, kSbo);
#pragma unroll
        for (int h = 0; h < MH; ++h) {
          const unsigned ah = h * (kM * KT * 4);   // 128 rows further down the A tile
          const unsigned long long dah = umma_desc(a_hi + ah + off, 128, kSbo), dal = umma_desc(a_lo + ah + off, 128, kSbo);
          umma_tf32(tmem + h * TN, dal, dbh, idesc, (i | ks) ? 1u : 0u);   // small terms first
          umma_tf32(tmem + h * TN, dah, dbl, idesc, 1u);
          umma_tf32(tmem + h * TN, dah, dbh, idesc, 1u);
        }
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
#pragma unroll 1
  for (int h = 0; h < MH; ++h) {
    float* dst = part + (((size_t)sp * g.m_tiles + mt * MH + h) * g.n_tiles + nt) * (size_t)(kM * TN);
#pragma unroll 2
    for (int c0 = 0; c0 < TN; c0 += 16) {
      unsigned v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = 0u;
      if (nk > 0) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(tmem + ((unsigned)(warp * 32) << 16) + (unsigned)(h * TN + c0)));
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc_n<kCols>(tmem);
}

// ------------------------------------------------------------------------------------------- epilogue
// One CTA per output row: a feature row n (all samples m; transposed partial tiles, coalesced along m) for the
// feature-major modes, a gradient row for EPI_ROWMAJOR. Split partials are summed in split order. V = 4: 16-byte
// accesses (sample count and row stride multiples of 4), V = 1: scalar fallback.
template <int V>
__device__ __forceinline__ void ldv(const float* p, float (&x)[V]) {
  if constexpr (V == 4) {
    const float4 t = __ldcg(reinterpret_cast<const float4*>(p));
    x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
  } else {
    x[0] = __ldcg(p);
  }
}
template <int V>
__device__ __forceinline__ void stv(float* p, const float (&x)[V]) {
  if constexpr (V == 4) __stcg(reinterpret_cast<float4*>(p), make_float4(x[0], x[1], x[2], x[3]));
  else __stcg(p, x[0]);
}
// sum over the K splits of V consecutive elements of one partial-tile row
template <int V>
__device__ __forceinline__ void fetchv(const float* p, int splits, size_t sstride, float (&c)[V]) {
#pragma unroll
  for (int q = 0; q < V; ++q) c[q] = 0.f;
#pragma unroll 4
  for (int s = 0; s < splits; ++s) {
    float t[V];
    ldv<V>(p + s * sstride, t);
#pragma unroll
    for (int q = 0; q < V; ++q) c[q] += t[q];
  }
}

// Emit V consecutive samples m.. of feature n as elements (r = m, k = n) of the packed A operand of the next GEMM.
template <int V>
__device__ __forceinline__ void emit_tp(const Epi& e, int n, int m, const float (&x)[V]) {
  const int R = e.tp_R, KT = e.tp_KT;
  float* t = e.tp + ((size_t)(m / R) * e.tp_ktiles + n / KT) * (size_t)(2 * R * KT);
  const int base = ((m % R) >> 3) * (8 * KT) + ((n % KT) >> 2) * 32 + (n & 3);
#pragma unroll
  for (int q = 0; q < V; ++q) {   // V == 4: m is a multiple of 4, so the four samples stay inside one 8-row group
    unsigned hi, lo;
    split_tf32(x[q], hi, lo);
    const int o = base + ((m + q) & 7) * 4;
    t[o] = __uint_as_float(hi);
    t[R * KT + o] = __uint_as_float(lo);
  }
}

template <int MODE, int V, int PER>
__global__ void __launch_bounds__(256) epi_kernel(Gemm g, const float* __restrict__ part, Epi e) {
  __shared__ float red[32];
  const size_t tile = (size_t)kM * g.TN;
  const size_t sstride = (size_t)g.m_tiles * g.n_tiles * tile;
  const int tid = threadIdx.x;
  if constexpr (MODE == EPI_ROWMAJOR) {
    const int m = blockIdx.x;
    const float* row = part + (size_t)(m / kM) * g.n_tiles * tile + (size_t)(m % kM) * g.TN;
    for (int n = tid * V; n < g.N; n += 256 * V) {
      float c[V];
      fetchv<V>(row + (size_t)(n / g.TN) * tile + n % g.TN, g.splits, sstride, c);
      stv<V>(e.out + (long long)m * e.ldo + n, c);
    }
    return;
  }
  const int n = blockIdx.x;
  const float* row = part + (size_t)(n / g.TN) * tile + (size_t)(n % g.TN) * kM;
  auto src = [&](int m) { return row + (size_t)(m >> 7) * g.n_tiles * tile + (m & 127); };
  const long long orow = (long long)n * e.ldo;
  if constexpr (MODE == EPI_T_RAW) {
    for (int m = tid * V; m < g.M; m += 256 * V) {
      float c[V];
      fetchv<V>(src(m), g.splits, sstride, c);
      stv<V>(e.out + orow + m, c);
    }
  } else if constexpr (MODE == EPI_T_BIAS) {
    const float b = e.bias ? __ldg(e.bias + n) : 0.f;
    const float sc = e.scale ? expf(3.f * __ldg(e.scale + n)) : 1.f;
    for (int m = tid * V; m < g.M; m += 256 * V) {
      float c[V];
      fetchv<V>(src(m), g.splits, sstride, c);
#pragma unroll
      for (int q = 0; q < V; ++q) c[q] = (c[q] + b) * sc;
      stv<V>(e.out + orow + m, c);
    }
  } else if constexpr (MODE == EPI_T_SCATTER) {
    const float sc = expf(__ldg(e.lsi));
    const int r = __ldg(e.rowmap + n);
    for (int m = tid * V; m < g.M; m += 256 * V) {
      float c[V], ad[V];
      ldv<V>(e.add + orow + m, ad);
      fetchv<V>(src(m), g.splits, sstride, c);
#pragma unroll
      for (int q = 0; q < V; ++q) c[q] = fmaf(ad[q], sc, c[q]);
      stv<V>(e.out + (long long)r * e.ldo + m, c);
    }
  } else if constexpr (MODE == EPI_T_AFFINE) {
    const float mul = __ldg(e.gamma + n), sub = __ldg(e.beta + n);
    for (int m = tid * V; m < g.M; m += 256 * V) {
      float c[V];
      fetchv<V>(src(m), g.splits, sstride, c);
#pragma unroll
      for (int q = 0; q < V; ++q) c[q] = fmaf(c[q], mul, -sub);
      stv<V>(e.out + orow + m, c);
    }
  } else if constexpr (MODE == EPI_T_HID) {
    const float bias = __ldg(e.bias + n);
    const float slope = e.slope != 0.f ? e.slope : kLreluSlope, bn_eps = e.eps != 0.f ? e.eps : kBnEps;
    auto lrelu = [slope](float v) { return v > 0.f ? v : slope * v; };
    if (e.bn_mode == 1) {   // BatchNorm1d in training mode (realnvpfc_model.py:178,185): biased batch variance, eps 1e-2
      float v[PER][V];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int m = (tid + 256 * i) * V;
#pragma unroll
        for (int q = 0; q < V; ++q) v[i][q] = 0.f;
        if (m < g.M) {
          fetchv<V>(src(m), g.splits, sstride, v[i]);
#pragma unroll
          for (int q = 0; q < V; ++q) { v[i][q] = lrelu(v[i][q] + bias); s += v[i][q]; }
        }
      }
      const float invB = 1.0f / (float)g.M;
      const float mean = block_sum(s, red) * invB;
      float qq = 0.f;
#pragma unroll
      for (int i = 0; i < PER; ++i)
        if ((tid + 256 * i) * V < g.M) {
#pragma unroll
          for (int q = 0; q < V; ++q) { const float d = v[i][q] - mean; qq = fmaf(d, d, qq); }
        }
      const float var = block_sum(qq, red) * invB;
      const float rstd = 1.0f / sqrtf(var + bn_eps);
      const float ga = __ldg(e.gamma + n) * rstd, be = __ldg(e.beta + n);
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int m = (tid + 256 * i) * V;
        if (m < g.M) {
          float nv[V];
#pragma unroll
          for (int q = 0; q < V; ++q) nv[q] = fmaf(v[i][q] - mean, ga, be);
          stv<V>(e.out + orow + m, v[i]);
          stv<V>(e.out2 + orow + m, nv);
          if (e.tp) emit_tp<V>(e, n, m, nv);
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
        const float rs = 1.0f / sqrtf(e.rvar[n] + bn_eps);
        ga = __ldg(e.gamma + n) * rs;
        be = __ldg(e.beta + n) - e.rmean[n] * ga;
      }
      for (int m = tid * V; m < g.M; m += 256 * V) {
        float c[V], nv[V];
        fetchv<V>(src(m), g.splits, sstride, c);
#pragma unroll
        for (int q = 0; q < V; ++q) { c[q] = lrelu(c[q] + bias); nv[q] = e.bn_mode == 2 ? fmaf(c[q], ga, be) : c[q]; }
        stv<V>(e.out + orow + m, c);
        if (e.out2) stv<V>(e.out2 + orow + m, nv);
        if (e.tp) emit_tp<V>(e, n, m, nv);
      }
    }
  } else {   // EPI_T_BNBWD
    float gv[PER][V], l[PER][V];
    float mean = 0.f, rstd = 1.f, gam = 1.f, s1 = 0.f, s2 = 0.f;
    if (e.bn_mode) { mean = e.mean[n]; rstd = e.rstd[n]; gam = __ldg(e.gamma + n); }
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int m = (tid + 256 * i) * V;
#pragma unroll
      for (int q = 0; q < V; ++q) { gv[i][q] = 0.f; l[i][q] = 0.f; }
      if (m < g.M) {
        ldv<V>(e.lsaved + orow + m, l[i]);
        fetchv<V>(src(m), g.splits, sstride, gv[i]);
#pragma unroll
        for (int q = 0; q < V; ++q) { s1 += gv[i][q]; s2 = fmaf(gv[i][q], (l[i][q] - mean) * rstd, s2); }
      }
    }
    float gbeta = 0.f, ggamma = 0.f;
    if (e.bn_mode) { gbeta = block_sum(s1, red); ggamma = block_sum(s2, red); }
    const float invB = 1.0f / (float)g.M;
    float s3 = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int m = (tid + 256 * i) * V;
      if (m < g.M) {
        float o[V];
#pragma unroll
        for (int q = 0; q < V; ++q) {
          float gl = gv[i][q];
          if (e.bn_mode) gl = gam * rstd * (gl - gbeta * invB - (l[i][q] - mean) * rstd * ggamma * invB);
          o[q] = gl * (l[i][q] > 0.f ? 1.f : kLreluSlope);
          s3 += o[q];
        }
        stv<V>(e.out + orow + m, o);
        if (e.tp) emit_tp<V>(e, n, m, o);
      }
    }
    s3 = block_sum(s3, red);
    if (tid == 0) {
      if (e.bn_mode) { e.g_gamma[n] = ggamma; e.g_beta[n] = gbeta; }
      e.g_bias[n] = s3;
    }
  }
}

// The coupling tail on the raw ZeroFC result. CTA = 4 column lanes x 64 sample items: kTailCols coupling columns of
// 64 * V samples; the four lanes' logdet partial sums are combined in lane order (deterministic).
template <int V>
__global__ void __launch_bounds__(256) epi_tail_kernel(Gemm g, const float* __restrict__ part, Epi e, int m_blocks) {
  __shared__ float sld[4][64 * V];
  const size_t tile = (size_t)kM * g.TN;
  const size_t sstride = (size_t)g.m_tiles * g.n_tiles * tile;
  const int chunk = blockIdx.x / m_blocks, mb = blockIdx.x % m_blocks;
  const int cl = threadIdx.x >> 6, mi = threadIdx.x & 63;
  const int m = (mb * 64 + mi) * V;
  const bool okm = m < g.M;
  float qa = 1.f, qc = 0.f;
  if (e.post_mode == 0) { qa = expf(__ldg(e.post_lsi)); qc = -__ldg(e.post_loc); }
  else if (e.post_mode == 1) { const float inv = 1.0f / expf(__ldg(e.post_lsi)); qa = inv; qc = __ldg(e.post_loc) * inv; }
  float ld[V];
#pragma unroll
  for (int q = 0; q < V; ++q) ld[q] = 0.f;
  auto src = [&](int n) {
    return part + (size_t)(n / g.TN) * tile + (size_t)(n % g.TN) * kM + (size_t)(m >> 7) * g.n_tiles * tile + (m & 127);
  };
  if (okm) {
#pragma unroll 2
    for (int q4 = 0; q4 < kTailCols / 4; ++q4) {
      const int j = chunk * kTailCols + cl * (kTailCols / 4) + q4;
      if (j >= e.Db) break;
      const float e_ls = expf(3.f * __ldg(e.scale + j)), e_t = expf(3.f * __ldg(e.scale + e.Db + j));
      const float b_ls = __ldg(e.b3 + j), b_t = __ldg(e.b3 + e.Db + j);
      const int gb = __ldg(e.gmap + e.Da + j), ga = __ldg(e.gmap + j);
      float bv[V], av[V], cls[V], ct[V];
      ldv<V>(e.yprev + (long long)gb * e.ldo + m, bv);
      ldv<V>(e.yprev + (long long)ga * e.ldo + m, av);
      fetchv<V>(src(j), g.splits, sstride, cls);
      fetchv<V>(src(e.Db + j), g.splits, sstride, ct);
      float ols[V], ot[V], yb[V], ya[V];
#pragma unroll
      for (int q = 0; q < V; ++q) {
        ols[q] = (cls[q] + b_ls) * e_ls;
        ot[q] = (ct[q] + b_t) * e_t;
        const float ls = tanhf(ols[q]);
        float bp;
        if (e.dir == 0) { bp = bv[q] / expf(ls) - ot[q]; ld[q] -= ls; }
        else            { bp = (bv[q] + ot[q]) * expf(ls); ld[q] += ls; }
        yb[q] = fmaf(bp, qa, qc);
        ya[q] = fmaf(av[q], qa, qc);
      }
      stv<V>(e.out + (long long)j * e.ldo + m, ols);
      stv<V>(e.out + (long long)(e.Db + j) * e.ldo + m, ot);
      stv<V>(e.y + (long long)(e.Da + j) * e.ldo + m, yb);
      stv<V>(e.y + (long long)j * e.ldo + m, ya);
    }
    if (e.Da > e.Db && chunk == 0 && cl == 0) {   // odd D: the a-half has one more row (chunk(2,1) semantics, :241)
      const int col = e.Da - 1;
      float av[V];
      ldv<V>(e.yprev + (long long)__ldg(e.gmap + col) * e.ldo + m, av);
#pragma unroll
      for (int q = 0; q < V; ++q) av[q] = fmaf(av[q], qa, qc);
      stv<V>(e.y + (long long)col * e.ldo + m, av);
    }
  }
#pragma unroll
  for (int q = 0; q < V; ++q) sld[cl][mi * V + q] = ld[q];
  __syncthreads();
  if (cl == 0 && okm) {
    float o[V];
#pragma unroll
    for (int q = 0; q < V; ++q) o[q] = ((sld[0][mi * V + q] + sld[1][mi * V + q]) + sld[2][mi * V + q]) + sld[3][mi * V + q];
    stv<V>(e.ldp + (long long)chunk * e.ldo + m, o);
  }
}

int g_force_tn = 0, g_force_kt = 0, g_force_mh = 0;

constexpr int kSmemBudget = 200 * 1024;
inline int stage_bytes(int TN, int KT, int MH) { return (2 * MH * kM * KT + 2 * TN * KT) * 4; }
inline int stages_of(int TN, int KT, int MH) { return std::max(2, std::min(kMaxStages, kSmemBudget / stage_bytes(TN, KT, MH))); }
inline int smem_of(const Gemm& g) { return g.stages * stage_bytes(g.TN, g.KT, g.MH) + 1024; }

// the instantiated (TN, KT, MH) configurations
#define DPI_TC_CONFIGS(X) X(64, 32, 1) X(128, 32, 1) X(256, 32, 1) X(256, 16, 1) X(256, 16, 2)

void finish_plan(Gemm& g, int n_sm) {
  g.k_tiles = (g.K + g.KT - 1) / g.KT;
  g.m_tiles = ((g.M + kM - 1) / kM + g.MH - 1) / g.MH * g.MH;      // 128-row tiles, a multiple of MH
  g.n_tiles = (g.N + g.TN - 1) / g.TN;
  // one CTA per SM (the stage ring takes the whole shared memory): split K so the grid is as close to one full
  // wave as possible without spilling into a second, nearly empty one; at least 128 k per split
  const int max_splits = std::min(16, std::max(1, g.k_tiles * g.KT / 128));
  const int mn = (g.m_tiles / g.MH) * g.n_tiles;
  int want = std::max(1, n_sm / std::max(1, mn));
  want = std::min(want, max_splits);
  g.kt_per_split = (g.k_tiles + want - 1) / want;
  g.splits = (g.k_tiles + g.kt_per_split - 1) / g.kt_per_split;
  g.stages = stages_of(g.TN, g.KT, g.MH);
}

}  // namespace

// Tile choice: the GEMM is fed from L2 (hi + lo copies of both operands), so the widest column tile that still
// fills the GPU (with split-K) wins: 128 x 256 with k-tile 16 (4 stages; measured faster than k-tile 32 x 2 stages),
// else 128 x 128, else 128 x 64. The 256 x 256 CTA tile (two accumulators = the whole TMEM, 2/3 of the operand
// bytes per MAC) measured slower on the C4 shapes - fewer CTAs, more K splits, a 256 KB unoverlapped drain per CTA -
// and is only selectable: DPI_B200_TC_TN / _KT / _MH force a configuration.
Gemm make_gemm(int M, int N, int K, int n_sm) {
  Gemm g{};
  g.M = M; g.N = N; g.K = K;
  const int cand[3][3] = {{256, 16, 1}, {128, 32, 1}, {64, 32, 1}};
  for (int c = 0; c < 3; ++c) {
    g.TN = cand[c][0]; g.KT = cand[c][1]; g.MH = cand[c][2];
    if (g_force_tn) {
      g.TN = g_force_tn;
      g.KT = g_force_kt ? g_force_kt : 32;
      g.MH = g_force_mh ? g_force_mh : 1;
      if (g.TN != 256 && (g.KT != 32 || g.MH != 1)) { g.KT = 32; g.MH = 1; }
      if (g.MH == 2) g.KT = 16;
      finish_plan(g, n_sm);
      return g;
    }
    if (g.TN > 64 && N < g.TN) continue;
    if (g.MH == 2 && M < 256) continue;
    finish_plan(g, n_sm);
    if ((long long)(g.m_tiles / g.MH) * g.n_tiles * g.splits >= (3LL * n_sm) / 4) break;
  }
  return g;
}

int init() {
  // the opt-in to > 48 KB of dynamic shared memory is per device: remember which devices have it
  static std::atomic<unsigned long long> done{0};
  int dev = 0;
  DPI_CUDA(cudaGetDevice(&dev));
  const unsigned long long bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_acquire) & bit) return DPI_OK;
#define X(tn_, kt_, mh_)                                                                                               \
  DPI_CUDA(cudaFuncSetAttribute((const void*)gemm_kernel<tn_, kt_, mh_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                stages_of(tn_, kt_, mh_) * stage_bytes(tn_, kt_, mh_) + 1024));
  DPI_TC_CONFIGS(X)
#undef X
  done.fetch_or(bit, std::memory_order_release);
  return DPI_OK;
}

void configure() {
  auto geti = [](const char* n) { const char* e = getenv(n); return e ? atoi(e) : 0; };
  const int tn = geti("DPI_B200_TC_TN"), kt = geti("DPI_B200_TC_KT"), mh = geti("DPI_B200_TC_MH");
  g_force_tn = (tn == 64 || tn == 128 || tn == 256) ? tn : 0;
  g_force_kt = (kt == 16 || kt == 32) ? kt : 0;
  g_force_mh = (mh == 1 || mh == 2) ? mh : 0;
}

static int pack_vec(const Pack& p) {
  if (p.im_h > 0) return 0;   // window gathers are scalar
  // 16-byte loads need an aligned base, a row stride that keeps rows aligned, and the contiguous extent a multiple of 4
  return ((reinterpret_cast<uintptr_t>(p.src) & 15) == 0 && (p.ld & 3) == 0 && ((p.k_contig ? p.K : p.rows) & 3) == 0) ? 1 : 0;
}

int pack2(const Gemm& g, const Pack& pa, float* dst_a, const Pack& pb, float* dst_b, cudaStream_t st) {
  const int RA = g.MH * kM, RB = g.TN;
  const int na = ((pa.rows + RA - 1) / RA) * g.k_tiles, nb = ((pb.rows + RB - 1) / RB) * g.k_tiles;
  const int va = pack_vec(pa), vb = pack_vec(pb);
  bool ok = false;
#define X(tn_, kt_, mh_)                                                                                     \
  if (!ok && g.TN == tn_ && g.KT == kt_ && g.MH == mh_) {                                                    \
    pack2_kernel<mh_ * kM, tn_, kt_><<<na + nb, 256, 0, st>>>(pa, va, dst_a, na, pb, vb, dst_b);             \
    ok = true;                                                                                               \
  }
  DPI_TC_CONFIGS(X)
#undef X
  if (!ok) { set_error("ftc::pack2: unsupported configuration TN=%d KT=%d MH=%d", g.TN, g.KT, g.MH); return DPI_ERR_INVALID; }
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int pack_weights(const WeightPack& w, const float* P, float* dst, cudaStream_t st) {
  for (int l = 0; l < 3; ++l) {
    const bool ok = (w.TN[l] == 64 && w.KT[l] == 32) || (w.TN[l] == 128 && w.KT[l] == 32) || (w.TN[l] == 256 && (w.KT[l] == 32 || w.KT[l] == 16));
    if (!ok) { set_error("ftc::pack_weights: unsupported tile %d x %d", w.TN[l], w.KT[l]); return DPI_ERR_INVALID; }
  }
  const int per_stage = w.tiles[0] + w.tiles[1] + w.tiles[2];
  packw_kernel<<<w.S * per_stage, 256, 0, st>>>(w, P, dst);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int run(const Gemm& g, const float* a_pack, const float* b_pack, float* part, const Epi& e, cudaStream_t st) {
  const int grid = (g.m_tiles / g.MH) * g.n_tiles * g.splits, tr = e.mode != EPI_ROWMAJOR;
  {
    bool ok = false;
#define X(tn_, kt_, mh_)                                                                                     \
    if (!ok && g.TN == tn_ && g.KT == kt_ && g.MH == mh_) {                                                  \
      gemm_kernel<tn_, kt_, mh_><<<grid, 128, smem_of(g), st>>>(g, a_pack, b_pack, part, tr);                \
      ok = true;                                                                                             \
    }
    DPI_TC_CONFIGS(X)
#undef X
    if (!ok) { set_error("ftc::run: unsupported configuration TN=%d KT=%d MH=%d", g.TN, g.KT, g.MH); return DPI_ERR_INVALID; }
  }
  DPI_LAUNCH_CHECK();
  const bool fused = e.mode == EPI_T_BNBWD || (e.mode == EPI_T_HID && e.bn_mode == 1);
  if (fused && g.M > kEpiMaxM) {
    set_error("ftc::run: fused BatchNorm epilogue supports at most %d samples (got %d)", kEpiMaxM, g.M);
    return DPI_ERR_UNSUPPORTED;
  }
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const int ext = e.mode == EPI_ROWMAJOR ? g.N : g.M;   // the contiguous extent of an output row
  const bool vec = (ext & 3) == 0 && (e.ldo & 3) == 0 && al(e.out) && al(e.out2) && al(e.add) && al(e.lsaved) && al(e.yprev) &&
                   al(e.y) && al(e.ldp) && al(part);
#define DPI_EPI(MODE, PER)                                                                       \
  do {                                                                                           \
    if (vec) epi_kernel<MODE, 4, PER><<<e.mode == EPI_ROWMAJOR ? g.M : g.N, 256, 0, st>>>(g, part, e); \
    else epi_kernel<MODE, 1, 4 * PER><<<e.mode == EPI_ROWMAJOR ? g.M : g.N, 256, 0, st>>>(g, part, e); \
  } while (0)
  switch (e.mode) {
    case EPI_T_RAW: DPI_EPI(EPI_T_RAW, 1); break;
    case EPI_T_SCATTER: DPI_EPI(EPI_T_SCATTER, 1); break;
    case EPI_T_BIAS: DPI_EPI(EPI_T_BIAS, 1); break;
    case EPI_T_AFFINE: DPI_EPI(EPI_T_AFFINE, 1); break;
    case EPI_ROWMAJOR: DPI_EPI(EPI_ROWMAJOR, 1); break;
    case EPI_T_HID:
      if (g.M <= 1024) DPI_EPI(EPI_T_HID, 1); else DPI_EPI(EPI_T_HID, 4);
      break;
    case EPI_T_BNBWD:
      if (g.M <= 1024) DPI_EPI(EPI_T_BNBWD, 1); else DPI_EPI(EPI_T_BNBWD, 4);
      break;
    case EPI_T_TAIL: {
      const int chunks = (e.Db + kTailCols - 1) / kTailCols;
      if (vec) { const int mbk = (g.M + 255) / 256; epi_tail_kernel<4><<<chunks * mbk, 256, 0, st>>>(g, part, e, mbk); }
      else { const int mbk = (g.M + 63) / 64; epi_tail_kernel<1><<<chunks * mbk, 256, 0, st>>>(g, part, e, mbk); }
      break;
    }
    default:
      set_error("ftc::run: unknown epilogue mode %d", e.mode);
      return DPI_ERR_INVALID;
  }
#undef DPI_EPI
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

}  // namespace ftc
}  // namespace dpi
