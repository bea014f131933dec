// flow_cl.cu - RealNVP.reverse (and its backward) for the reference's primary shape as ONE thread-block cluster of
// 16 CTAs whose coupling-net GEMMs run on tcgen05 tensor cores.
//
// Math: generative_model/realnvpfc_model.py AffineCoupling.reverse :240-249, net :171-187, ZeroFC :137-141,
// ActNorm.reverse :49-66, Flow.reverse :457-474, RealNVP.reverse :594-603; backward per SURVEY.md Appendix A.
// Same feature-major workspace as flow_small.cu / flow_kernels.cu (X^T[feature][sample], permutations folded into
// row index maps), so either backward can consume this forward's saved activations.
//
// Why: at batch <= 64 a coupling stage is a chain of dependent GEMMs of 1-8 MFLOP. flow_small.cu spreads every GEMM
// over the whole GPU and pays a grid-wide barrier plus L2 round trips per hop (~5 us x 228 hops). Here the chain
// stays inside one cluster:
//   * swap-AB tcgen05.mma (kind::tf32, 3xTF32 split, M = 128 weight rows on the TMEM lanes, N = 64 samples on the
//     columns) - a hop's GEMM is 0.05-0.8 us on 16 SMs;
//   * net.0 (K = 512) and net.3 (K = 128) are split over K across the 16 CTAs; the partial accumulators are
//     reduce-scattered through L2 so that CTA c finalises hidden features 8c..8c+7 for ALL samples - BatchNorm batch
//     statistics are then reductions over 16 lanes of one CTA, no cross-CTA statistics exchange;
//   * net.6 (N = 1024) is split over output rows: CTA c owns coupling columns 32c..32c+31 (both the log-scale and
//     the shift row of each) and applies the tanh-exp coupling, ActNorm and the logdet partial as its epilogue;
//   * hops hand over through L2 with barrier.cluster (hardware, ~0.2 us) instead of a software grid barrier;
//   * weights never touch the critical path: a pre-pass splits them into TF32 hi / lo parts and tiles them per
//     (stage, CTA) in the canonical K-major UMMA layout; one cp.async.bulk per layer brings the NEXT stage's slice
//     into shared memory as soon as the current stage's MMAs have released the buffer.
// All reductions have a fixed order (bit-reproducible), there are no float atomics.
#include <algorithm>

#include "flow_cl.cuh"
#include "flow_host.cuh"

namespace dpi {
namespace fcl {
namespace {

constexpr int kT = kThreadsCl;
constexpr int kTmemColsCl = 256;        // three 64-column accumulators (power of two >= 192)

// ----------------------------------------------------------------------------------------------- small PTX wrappers
__device__ __forceinline__ unsigned cta_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// release / acquire at cluster scope: every global write of every CTA before the barrier is visible after it
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// bounded wait: a protocol bug or a lost copy sets the sticky error flag (ctrl[2]) instead of hanging the GPU; the
// final phase turns the flag into NaN outputs so that the caller cannot miss it
__device__ __forceinline__ void mbar_wait_flag(unsigned long long* bar, unsigned parity, unsigned* ctrl) {
  const unsigned a = smem_u32(bar);
  unsigned done = 0;
  long long t0 = 0;
  bool timed = false;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (!done) {
      if (!timed) { t0 = clock64(); timed = true; }
      else if (clock64() - t0 > 2000000000LL) { atomicExch(ctrl + 2, 1u); break; }
    }
  }
}
__device__ __forceinline__ void tmem_alloc_cl(unsigned* slot_smem) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "n"(kTmemColsCl) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cl(unsigned taddr) {      // the same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kTmemColsCl) : "memory");
}
// 32 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(unsigned taddr, float (&v)[32]) {
  unsigned r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
      "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// D[128 x 64] (+)= A[128 x 8 ksteps] * B[64 x 8 ksteps]^T, 3xTF32: a_lo b_hi + a_hi b_lo + a_hi b_hi (small terms first).
// Operand tiles are canonical K-major: 16-byte k-chunks 128 B apart (LBO), 8-row groups sbo bytes apart.
__device__ __forceinline__ void issue_gemm(unsigned tmem_d, unsigned a_hi, unsigned a_lo, unsigned a_sbo, unsigned b_hi,
                                           unsigned b_lo, unsigned b_sbo, int ksteps, unsigned idesc) {
#pragma unroll 1
  for (int ks = 0; ks < ksteps; ++ks) {
    const unsigned off = ks * 256;          // 8 k = two 16-byte chunks
    const unsigned long long dah = umma_desc(a_hi + off, 128, a_sbo), dal = umma_desc(a_lo + off, 128, a_sbo);
    const unsigned long long dbh = umma_desc(b_hi + off, 128, b_sbo), dbl = umma_desc(b_lo + off, 128, b_sbo);
    umma_tf32(tmem_d, dal, dbh, idesc, ks ? 1u : 0u);
    umma_tf32(tmem_d, dah, dbl, idesc, 1u);
    umma_tf32(tmem_d, dah, dbh, idesc, 1u);
  }
}

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void stcg4(float* p, float4 v) { __stcg(reinterpret_cast<float4*>(p), v); }

// ----------------------------------------------------------------------------------------------- weight pre-pass
// dst[(e * 16 + c) * kPackFloats ...] <- this (stage, CTA)'s three operand tiles, TF32 hi then lo.
//   forward : W1[h][32c + k] (128 x 32), W2[h][8c + k] (128 x 8), W3[row(r)][k] (64 x 128)
//   backward: W1^T: [kl][h] = W1[h][32c + kl] (32 x 128), W2^T: [h1][kl] = W2[8c + kl][h1] (128 x 8),
//             W3^T: [h][jl] = W3[row(jl)][h] (128 x 64)
// row(r) = 32c + r for r < 32 (log-scale rows), Db + 32c + r - 32 otherwise (shift rows)  (ZeroFC output chunk(2,1), :244)
__global__ void __launch_bounds__(256) pack_kernel(const StageTab* __restrict__ stb, const float* __restrict__ P, int S, int bwd,
                                                   float* __restrict__ dst) {
  constexpr int Da = kDc / 2, Db = kDc / 2, H = kHc;
  constexpr int C1 = kW1Floats / 8, C2 = kW2Floats / 8, C3 = kW3Floats / 8;   // 16-byte chunks of the hi parts
  constexpr int CT = C1 + C2 + C3;
  const long long total = (long long)S * kCL * CT;
  for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (long long)gridDim.x * blockDim.x) {
    const int q0 = (int)(id % CT);
    const int sc = (int)(id / CT), c = sc % kCL, e = sc / kCL;
    const StageTab& st = stb[e];
    float* blk = dst + (size_t)sc * kPackFloats;
    int q, KS, layer;
    float* tile;
    int tile_floats;
    if (q0 < C1) { q = q0; layer = 1; tile = blk; tile_floats = kW1Floats / 2; }
    else if (q0 < C1 + C2) { q = q0 - C1; layer = 2; tile = blk + kW1Floats; tile_floats = kW2Floats / 2; }
    else { q = q0 - C1 - C2; layer = 3; tile = blk + kW1Floats + kW2Floats; tile_floats = kW3Floats / 2; }
    if (!bwd) KS = layer == 1 ? kKS1 : layer == 2 ? kHS : H;
    else KS = layer == 1 ? H : layer == 2 ? kHS : 2 * kJB;
    const int kcs = KS >> 2;                        // chunks per row
    const int r8 = q & 7, kc = (q >> 3) % kcs, rg = (q >> 3) / kcs;
    const int row = rg * 8 + r8, k = kc * 4;
    float4 v;
    if (!bwd) {
      const float* src;
      if (layer == 1) src = P + st.p[P_W1] + (size_t)row * Da + c * kKS1 + k;
      else if (layer == 2) src = P + st.p[P_W2] + (size_t)row * H + c * kHS + k;
      else {
        const int j3 = row < kJB ? c * kJB + row : Db + c * kJB + row - kJB;
        src = P + st.p[P_W3] + (size_t)j3 * H + k;
      }
      v = __ldg(reinterpret_cast<const float4*>(src));
    } else {
      float t[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kk = k + i;
        if (layer == 1) t[i] = __ldg(P + st.p[P_W1] + (size_t)kk * Da + c * kKS1 + row);          // [kl = row][h = kk]
        else if (layer == 2) t[i] = __ldg(P + st.p[P_W2] + (size_t)(c * kHS + kk) * H + row);     // [h1 = row][kl = kk]
        else {
          const int j3 = kk < kJB ? c * kJB + kk : Db + c * kJB + kk - kJB;
          t[i] = __ldg(P + st.p[P_W3] + (size_t)j3 * H + row);                                      // [h = row][jl = kk]
        }
      }
      v = make_float4(t[0], t[1], t[2], t[3]);
    }
    float4 hi, lo;
    split4(v, hi, lo);
    *reinterpret_cast<float4*>(tile + (size_t)q * 4) = hi;
    *reinterpret_cast<float4*>(tile + tile_floats + (size_t)q * 4) = lo;
  }
}

// ----------------------------------------------------------------------------------------------- shared memory
struct SmemFwd {
  // byte offsets from the 1024-aligned dynamic base
  static constexpr int oW1 = 0;                                  // 32 KB
  static constexpr int oW2 = oW1 + kW1Floats * 4;                // 8 KB
  static constexpr int oW3 = oW2 + kW2Floats * 4;                // 64 KB (descriptors span 128 rows: the tile after it is read too)
  static constexpr int oN2 = oW3 + kW3Floats * 4;                // [64 x 128] hi | lo = 64 KB
  static constexpr int oAC = oN2 + 2 * kNS * kHc * 4;            // [64 x 32] hi | lo = 16 KB
  static constexpr int oN1 = oAC + 2 * kNS * kKS1 * 4;           // [64 x 8] hi | lo = 4 KB
  static constexpr int oEnd = oN1 + 2 * kNS * kHS * 4;
  static constexpr int oTT = oAC;                                // tail exchange tile [64][68] (17 KB) aliases AC + N1
  static constexpr int bytes = oEnd + 1024;                      // + alignment slack
};
static_assert(64 * 68 * 4 <= 2 * kNS * kKS1 * 4 + 2 * kNS * kHS * 4, "tail tile must fit the region it aliases");

#define CL_STAMP(slot)                                                                      \
  do {                                                                                      \
    if (a.tbuf && blockIdx.x == 0 && threadIdx.x == 0) a.tbuf[(size_t)e * 16 + (slot)] = (unsigned long long)clock64(); \
  } while (0)

// reduce-scatter epilogue of a hidden layer: sum the 16 K-split partials of hidden feature 8c + fl in CTA order, bias,
// LeakyReLU, BatchNorm (batch statistics over the 16 lanes that share the feature / running statistics / none)
struct HidParams { float bias, gamma, beta, rmean, rvar; };

__device__ __forceinline__ void hidden_epilogue(const KArgs& a, const StageTab& st, int e, int which, int c, const float* part,
                                                const HidParams& hp, float4& nrm, bool& valid) {
  const int t = threadIdx.x, fl = t >> 4, q = t & 15, h = c * kHS + fl;
  const int LB = a.ldb;
  const float* pp = part + (size_t)h * kNS + 4 * q;
  float4 v[kCL];
#pragma unroll
  for (int s = 0; s < kCL; ++s) v[s] = ldcg4(pp + (size_t)s * (kHc * kNS));
  float4 u = v[0];
#pragma unroll
  for (int s = 1; s < kCL; ++s) { u.x += v[s].x; u.y += v[s].y; u.z += v[s].z; u.w += v[s].w; }
  valid = 4 * q < a.B;
  const float l0 = lrelu_f(u.x + hp.bias), l1 = lrelu_f(u.y + hp.bias), l2 = lrelu_f(u.z + hp.bias), l3 = lrelu_f(u.w + hp.bias);
  float* Lr = a.W + (which == 1 ? a.oL1T : a.oL2T) + ((size_t)e * a.H + h) * LB + 4 * q;
  if (valid) stcg4(Lr, make_float4(l0, l1, l2, l3));
  float n0 = l0, n1 = l1, n2 = l2, n3 = l3;
  if (a.has_bn) {
    if (a.train) {
      const float invB = 1.0f / (float)a.B;
      const float mean = half_warp_sum(valid ? (l0 + l1) + (l2 + l3) : 0.f) * invB;
      const float d0 = l0 - mean, d1 = l1 - mean, d2 = l2 - mean, d3 = l3 - mean;
      const float var = half_warp_sum(valid ? (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3) : 0.f) * invB;
      const float rstd = 1.0f / sqrtf(var + kBnEps);
      const float ga = hp.gamma * rstd;
      n0 = fmaf(d0, ga, hp.beta); n1 = fmaf(d1, ga, hp.beta); n2 = fmaf(d2, ga, hp.beta); n3 = fmaf(d3, ga, hp.beta);
      if (q == 0) {
        float* stt = a.W + a.oST + ((size_t)e * 4 + (which - 1) * 2) * a.H;
        __stcg(stt + h, mean);
        __stcg(stt + a.H + h, rstd);
        a.R[st.r[(which - 1) * 2] + h] = (1.f - kBnMomentum) * hp.rmean + kBnMomentum * mean;
        a.R[st.r[(which - 1) * 2 + 1] + h] = (1.f - kBnMomentum) * hp.rvar + kBnMomentum * (var * (float)a.B / (float)(a.B - 1));
      }
    } else {
      const float rs = 1.0f / sqrtf(hp.rvar + kBnEps);
      const float ga = hp.gamma * rs, be = hp.beta - hp.rmean * ga;
      n0 = fmaf(l0, ga, be); n1 = fmaf(l1, ga, be); n2 = fmaf(l2, ga, be); n3 = fmaf(l3, ga, be);
    }
  }
  nrm = make_float4(n0, n1, n2, n3);
  float* Nr = a.W + (which == 1 ? a.oN1T : a.oN2T) + ((size_t)e * a.H + h) * LB + 4 * q;
  if (valid) stcg4(Nr, nrm);
}

// this CTA's K-split partial accumulator (TMEM lane = hidden feature, column = sample) -> part[c][feature][sample]
__device__ __forceinline__ void store_partial(unsigned tmem_d, float* part, int c) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = (warp & 3) * 32 + lane, c0 = (warp >> 2) * 32;
  float v[32];
  tmem_ld32(tmem_d + ((unsigned)((warp & 3) * 32) << 16) + (unsigned)c0, v);
  float* dst = part + ((size_t)c * kHc + m) * kNS + c0;
#pragma unroll
  for (int i = 0; i < 32; i += 4) stcg4(dst + i, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
}

}  // namespace

// =================================================================================================== forward kernel
__global__ void __launch_bounds__(kT, 1) flow_cl_fwd_kernel(const __grid_constant__ KArgs a, const float* __restrict__ pack,
                                                             float* __restrict__ scratch) {
  extern __shared__ unsigned char raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) unsigned long long bars[4];     // W1, W2, W3 landed; MMA group complete
  __shared__ unsigned tmem_slot;
  __shared__ int sIdx[2][64];                              // gather rows of this CTA's a-slice (32) and b-slice (32)
  __shared__ float sLd[64];                                // running logdet partial of this CTA
  __shared__ float sRed[8 * 64];
  __shared__ float sT[32 * 33];

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int c = (int)cta_rank();
  const int LB = a.ldb, B = a.B, S = a.S;
  constexpr int D = kDc, Da = kDc / 2, Db = kDc / 2, H = kHc;
  float* part1 = scratch;
  float* part2 = scratch + (size_t)kCL * kHc * kNS;
  float* ldrows = scratch + kPartFloats;

  unsigned long long* barW1 = bars, *barW2 = bars + 1, *barW3 = bars + 2, *barM = bars + 3;
  if (t == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(bars + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc_cl(&tmem_slot);
  if (t < 64) sLd[t] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem = tmem_slot;
  const unsigned tD1 = tmem, tD2 = tmem + 64, tD3 = tmem + 128;
  const unsigned sW1 = smem_u32(sm + SmemFwd::oW1), sW2 = smem_u32(sm + SmemFwd::oW2), sW3 = smem_u32(sm + SmemFwd::oW3);
  const unsigned sN2 = smem_u32(sm + SmemFwd::oN2), sAC = smem_u32(sm + SmemFwd::oAC), sN1 = smem_u32(sm + SmemFwd::oN1);
  float* AC = reinterpret_cast<float*>(sm + SmemFwd::oAC);
  float* N1C = reinterpret_cast<float*>(sm + SmemFwd::oN1);
  float* N2C = reinterpret_cast<float*>(sm + SmemFwd::oN2);
  float* TT = reinterpret_cast<float*>(sm + SmemFwd::oTT);
  const unsigned idesc = umma_idesc(kNS, 0, 0);
  unsigned mph = 0;                                        // parity of the next MMA-complete wait

  // weights of stage 0 (nothing depends on them yet: the copies fly while the input is transposed)
  if (t == 0) {
    const float* blk = pack + (size_t)c * kPackFloats;
    mbar_expect_tx(barW1, kW1Floats * 4); bulk_g2s(sm + SmemFwd::oW1, blk, kW1Floats * 4, barW1);
    mbar_expect_tx(barW2, kW2Floats * 4); bulk_g2s(sm + SmemFwd::oW2, blk + kW1Floats, kW2Floats * 4, barW2);
    mbar_expect_tx(barW3, kW3Floats * 4); bulk_g2s(sm + SmemFwd::oW3, blk + kW1Floats + kW2Floats, kW3Floats * 4, barW3);
  }
  if (t < 64) sIdx[0][t] = __ldg(a.gmaps + (t < 32 ? c * kKS1 + t : Da + c * kJB + (t - 32)));

  // ---- z [B][D] -> ZT [D][LB]: this CTA's 64 features, 32 x 32 tiles through shared memory
  {
    const int lx = t & 31, ly = t >> 5;
    for (int tile = 0; tile < 4; ++tile) {
      const int f0 = c * 64 + (tile & 1) * 32, m0 = (tile >> 1) * 32;
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = m0 + ly + 8 * i;
        sT[(ly + 8 * i) * 33 + lx] = m < B ? __ldg(a.in + (size_t)m * D + f0 + lx) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int f = f0 + ly + 8 * i, m = m0 + lx;
        if (m < B) __stcg(a.W + a.oZT + (size_t)f * LB + m, sT[lx * 33 + ly + 8 * i]);
      }
    }
  }
  cluster_sync_all();

#pragma unroll 1
  for (int e = 0; e < S; ++e) {
    const StageTab& st = a.st[e];
    const float* yprev = e == 0 ? a.W + a.oZT : a.W + a.oYT + (size_t)(e - 1) * D * LB;
    float* y = a.W + a.oYT + (size_t)e * D * LB;
    const int* idx = sIdx[e & 1];
    CL_STAMP(0);
    // ---- per-stage scalars (parameters: independent of the chain, their latency hides behind the gather)
    const float qa = expf(__ldg(a.P + st.lsi)), qc = -__ldg(a.P + st.loc);
    HidParams hp1{}, hp2{};
    if (t < 128) {
      const int h = c * kHS + (t >> 4);
      hp1.bias = __ldg(a.P + st.p[P_B1] + h); hp2.bias = __ldg(a.P + st.p[P_B2] + h);
      if (a.has_bn) {
        hp1.gamma = __ldg(a.P + st.p[P_G1] + h); hp1.beta = __ldg(a.P + st.p[P_BE1] + h);
        hp2.gamma = __ldg(a.P + st.p[P_G2] + h); hp2.beta = __ldg(a.P + st.p[P_BE2] + h);
        hp1.rmean = a.R[st.r[0] + h]; hp1.rvar = a.R[st.r[1] + h];
        hp2.rmean = a.R[st.r[2] + h]; hp2.rvar = a.R[st.r[3] + h];
      }
    }
    float b3v = 0.f, e3v = 1.f;
    const int trow = (warp & 3) * 32 + lane;               // TMEM lane of this thread; rows 0..63 of D3 are real
    if (trow < 2 * kJB) {
      const int j3 = trow < kJB ? c * kJB + trow : Db + c * kJB + trow - kJB;
      b3v = __ldg(a.P + st.p[P_B3] + j3);
      e3v = expf(3.f * __ldg(a.P + st.p[P_SCALE] + j3));
    }
    if (e + 1 < S && t < 64)
      sIdx[(e + 1) & 1][t] = __ldg(a.gmaps + (size_t)(e + 1) * D + (t < 32 ? c * kKS1 + t : Da + c * kJB + (t - 32)));

    // ---- gather: a-slice -> canonical B operand of net.0 (hi / lo) + pass-through rows of y; b-slice -> registers
    float4 bv0, bv1;
    {
      const int p = t >> 3, s0 = (t & 7) * 8;
      const float* br = yprev + (size_t)idx[32 + p] * LB + s0;
      bv0 = s0 < B ? ldcg4(br) : make_float4(0.f, 0.f, 0.f, 0.f);
      bv1 = s0 + 4 < B ? ldcg4(br + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    {
      const int n = t & 63;
      float v[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int kc = (t >> 6) + 4 * i;
#pragma unroll
        for (int u = 0; u < 4; ++u) v[i][u] = n < B ? __ldcg(yprev + (size_t)idx[4 * kc + u] * LB + n) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int kc = (t >> 6) + 4 * i;
        if (n < B) {
#pragma unroll
          for (int u = 0; u < 4; ++u) __stcg(y + (size_t)(c * kKS1 + 4 * kc + u) * LB + n, fmaf(v[i][u], qa, qc));
        }
        float4 hi, lo;
        split4(make_float4(v[i][0], v[i][1], v[i][2], v[i][3]), hi, lo);
        const int off = (n >> 3) * 256 + kc * 32 + (n & 7) * 4;      // [64 x 32] tile: SBO 1024 B, LBO 128 B
        *reinterpret_cast<float4*>(AC + off) = hi;
        *reinterpret_cast<float4*>(AC + kNS * kKS1 + off) = lo;
      }
    }
    fence_async_smem();
    __syncthreads();
    CL_STAMP(1);
    // ---- net.0: D1[h][n] = sum_k W1[h][32c + k] a[k][n]  (this CTA's K slice)
    if (t == 0) {
      mbar_wait_flag(barW1, e & 1, a.ctrl);
      tc_fence_after();
      issue_gemm(tD1, sW1, sW1 + kW1Floats * 2, kKS1 * 32, sAC, sAC + kNS * kKS1 * 4, kKS1 * 32, kKS1 / 8, idesc);
      umma_commit(barM);
    }
    mbar_wait_flag(barM, mph, a.ctrl); mph ^= 1;
    tc_fence_after();
    if (t == 0 && e + 1 < S) {           // W1 buffer is free: next stage's slice
      mbar_expect_tx(barW1, kW1Floats * 4);
      bulk_g2s(sm + SmemFwd::oW1, pack + ((size_t)(e + 1) * kCL + c) * kPackFloats, kW1Floats * 4, barW1);
    }
    CL_STAMP(2);
    store_partial(tD1, part1, c);
    tc_fence_before();
    cluster_sync_all();
    CL_STAMP(3);
    if (t < 128) {
      float4 nr; bool valid;
      hidden_epilogue(a, st, e, 1, c, part1, hp1, nr, valid);
      // canonical B operand of net.3: [64 samples x 8 k], k = this CTA's hidden features (SBO 256 B, LBO 128 B)
      const int fl = t >> 4, q = t & 15;
      const float nv[4] = {nr.x, nr.y, nr.z, nr.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n = 4 * q + i;
        unsigned hi, lo;
        split_tf32(valid ? nv[i] : 0.f, hi, lo);
        const int off = (n >> 3) * 64 + (fl >> 2) * 32 + (n & 7) * 4 + (fl & 3);
        N1C[off] = __uint_as_float(hi);
        N1C[kNS * kHS + off] = __uint_as_float(lo);
      }
    }
    fence_async_smem();
    __syncthreads();
    CL_STAMP(4);
    // ---- net.3: D2[h2][n] = sum_k W2[h2][8c + k] n1[8c + k][n]
    if (t == 0) {
      mbar_wait_flag(barW2, e & 1, a.ctrl);
      tc_fence_after();
      issue_gemm(tD2, sW2, sW2 + kW2Floats * 2, kHS * 32, sN1, sN1 + kNS * kHS * 4, kHS * 32, kHS / 8, idesc);
      umma_commit(barM);
    }
    mbar_wait_flag(barM, mph, a.ctrl); mph ^= 1;
    tc_fence_after();
    if (t == 0 && e + 1 < S) {
      mbar_expect_tx(barW2, kW2Floats * 4);
      bulk_g2s(sm + SmemFwd::oW2, pack + ((size_t)(e + 1) * kCL + c) * kPackFloats + kW1Floats, kW2Floats * 4, barW2);
    }
    store_partial(tD2, part2, c);
    tc_fence_before();
    cluster_sync_all();
    CL_STAMP(5);
    if (t < 128) {
      float4 nr; bool valid;
      hidden_epilogue(a, st, e, 2, c, part2, hp2, nr, valid);      // n2 rows 8c..8c+7 -> N2T (global)
    }
    cluster_sync_all();
    CL_STAMP(6);
    // ---- all of n2 [128][B] (written by the 16 CTAs) -> canonical B operand of net.6: [64 samples x 128 k]
    {
      const float* N2 = a.W + a.oN2T + (size_t)e * H * LB;
      const int n = t & 63;
      float v[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int kc = (t >> 6) + 4 * i;
#pragma unroll
        for (int u = 0; u < 4; ++u) v[i][u] = n < B ? __ldcg(N2 + (size_t)(4 * kc + u) * LB + n) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int kc = (t >> 6) + 4 * i;
        float4 hi, lo;
        split4(make_float4(v[i][0], v[i][1], v[i][2], v[i][3]), hi, lo);
        const int off = (n >> 3) * 1024 + kc * 32 + (n & 7) * 4;     // SBO 4096 B, LBO 128 B
        *reinterpret_cast<float4*>(N2C + off) = hi;
        *reinterpret_cast<float4*>(N2C + kNS * kHc + off) = lo;
      }
    }
    fence_async_smem();
    __syncthreads();
    CL_STAMP(7);
    // ---- net.6.fc: D3[r][n] = sum_k W3[row(r)][k] n2[k][n], rows 0..63 real
    if (t == 0) {
      mbar_wait_flag(barW3, e & 1, a.ctrl);
      tc_fence_after();
      issue_gemm(tD3, sW3, sW3 + kW3Floats * 2, kHc * 32, sN2, sN2 + kNS * kHc * 4, kHc * 32, kHc / 8, idesc);
      umma_commit(barM);
    }
    mbar_wait_flag(barM, mph, a.ctrl); mph ^= 1;
    tc_fence_after();
    if (t == 0 && e + 1 < S) {
      mbar_expect_tx(barW3, kW3Floats * 4);
      bulk_g2s(sm + SmemFwd::oW3, pack + ((size_t)(e + 1) * kCL + c) * kPackFloats + kW1Floats + kW2Floats, kW3Floats * 4, barW3);
    }
    CL_STAMP(8);
    // ---- tail, step 1: ZeroFC bias + exp(3 scale) (:137-141) on this thread's output row -> exchange tile
    if (trow < 2 * kJB) {
      const int c0 = (warp >> 2) * 32;
      float v[32];
      tmem_ld32(tD3 + ((unsigned)((warp & 3) * 32) << 16) + (unsigned)c0, v);
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        *reinterpret_cast<float4*>(TT + trow * 68 + c0 + i) =
            make_float4((v[i] + b3v) * e3v, (v[i + 1] + b3v) * e3v, (v[i + 2] + b3v) * e3v, (v[i + 3] + b3v) * e3v);
    }
    tc_fence_before();
    __syncthreads();
    // ---- tail, step 2: coupling (:244-249) + ActNorm.reverse (:66) + logdet partial; thread = (column p, 8 samples)
    {
      const int p = t >> 3, s0 = (t & 7) * 8;
      const int col = c * kJB + p;
      const float4 ol0 = *reinterpret_cast<const float4*>(TT + p * 68 + s0), ol1 = *reinterpret_cast<const float4*>(TT + p * 68 + s0 + 4);
      const float4 ot0 = *reinterpret_cast<const float4*>(TT + (kJB + p) * 68 + s0), ot1 = *reinterpret_cast<const float4*>(TT + (kJB + p) * 68 + s0 + 4);
      const float ols[8] = {ol0.x, ol0.y, ol0.z, ol0.w, ol1.x, ol1.y, ol1.z, ol1.w};
      const float ots[8] = {ot0.x, ot0.y, ot0.z, ot0.w, ot1.x, ot1.y, ot1.z, ot1.w};
      const float bvs[8] = {bv0.x, bv0.y, bv0.z, bv0.w, bv1.x, bv1.y, bv1.z, bv1.w};
      float yb[8], dl[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float ls = tanhf(ols[i]);
        const float bp = bvs[i] / expf(ls) - ots[i];
        yb[i] = fmaf(bp, qa, qc);
        dl[i] = s0 + i < B ? -ls : 0.f;
      }
      float* O = a.W + a.oOT + (size_t)e * (2 * Db) * LB;
      if (s0 < B) {
        stcg4(O + (size_t)col * LB + s0, ol0);
        stcg4(O + (size_t)(Db + col) * LB + s0, ot0);
        stcg4(y + (size_t)(Da + col) * LB + s0, make_float4(yb[0], yb[1], yb[2], yb[3]));
      }
      if (s0 + 4 < B) {
        stcg4(O + (size_t)col * LB + s0 + 4, ol1);
        stcg4(O + (size_t)(Db + col) * LB + s0 + 4, ot1);
        stcg4(y + (size_t)(Da + col) * LB + s0 + 4, make_float4(yb[4], yb[5], yb[6], yb[7]));
      }
      // per-sample sum over this CTA's 32 columns: 4 columns per warp (lanes 8 apart), then the 8 warps in order
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        dl[i] += __shfl_xor_sync(0xffffffffu, dl[i], 8);
        dl[i] += __shfl_xor_sync(0xffffffffu, dl[i], 16);
      }
      if (lane < 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) sRed[warp * 64 + lane * 8 + i] = dl[i];
      }
    }
    __syncthreads();
    if (t < 64) {
      float sum = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) sum += sRed[r * 64 + t];
      sLd[t] += sum;
    }
    CL_STAMP(9);
    cluster_sync_all();      // y of this stage is complete and visible to the gathers of the next one
  }

  // ---- logdet: per-CTA partial rows -> CTA 0 sums them in CTA order  (+ D * sum_e log_scale_inv_e, :60-64)
  if (t < 64) __stcg(ldrows + (size_t)c * kNS + t, sLd[t]);
  // ---- YT[S - 1] [D][LB] -> out [B][D]: this CTA's 64 features
  {
    const float* ylast = a.W + a.oYT + (size_t)(S - 1) * D * LB;
    const int lx = t & 31, ly = t >> 5;
    for (int tile = 0; tile < 4; ++tile) {
      const int f0 = c * 64 + (tile & 1) * 32, m0 = (tile >> 1) * 32;
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int f = f0 + ly + 8 * i, m = m0 + lx;
        sT[(ly + 8 * i) * 33 + lx] = m < B ? __ldcg(ylast + (size_t)f * LB + m) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = m0 + ly + 8 * i;
        if (m < B) a.out[(size_t)m * D + f0 + lx] = sT[lx * 33 + ly + 8 * i];
      }
    }
  }
  cluster_sync_all();
  if (c == 0 && t < 64) {
    float lsum = 0.f;
    for (int e = 0; e < S; ++e) lsum += __ldg(a.P + a.st[e].lsi);
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < kCL; ++r) acc += __ldcg(ldrows + (size_t)r * kNS + t);
    float res = acc + (float)D * lsum;
    if (*reinterpret_cast<volatile unsigned*>(a.ctrl + 2)) res = __int_as_float(0x7fc00000);   // a wait timed out: poison
    if (t < B) a.logdet[t] = res;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc_cl(tmem);
}

// =================================================================================================== host side
namespace {
int g_query_device = -1, g_query_ok = 0;
}

int has_backward() { return 0; }

bool shape_ok(int D, int H, int B) { return D == kDc && H == kHc && B >= 4 && B <= kNS && (B & 3) == 0; }

int query(int device) {
  if (g_query_device == device) return g_query_ok;
  g_query_device = device;
  g_query_ok = 0;
  if (cudaFuncSetAttribute((const void*)flow_cl_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemFwd::bytes) != cudaSuccess ||
      cudaFuncSetAttribute((const void*)flow_cl_fwd_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kCL); cfg.blockDim = dim3(kT); cfg.dynamicSmemBytes = SmemFwd::bytes;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kCL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int nc = 0;
  if (cudaOccupancyMaxActiveClusters(&nc, (const void*)flow_cl_fwd_kernel, &cfg) != cudaSuccess || nc < 1) {
    cudaGetLastError();
    return 0;
  }
  g_query_ok = 1;
  return 1;
}

static int launch_pack(const KArgs& a, int S, int bwd, float* dst, cudaStream_t st) {
  const long long total = (long long)S * kCL * (kPackFloats / 8);
  const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 8);
  pack_kernel<<<grid, 256, 0, st>>>(a.st, a.P, S, bwd, dst);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int forward(const KArgs& a, const Plan& p, int S, cudaStream_t st) {
  float* packf = a.W + p.oPackF;
  float* scratch = a.W + p.oScratch;
  DPI_CUDA(cudaMemsetAsync(a.ctrl, 0, kCtrlWords * sizeof(unsigned), st));
  int rc = launch_pack(a, S, 0, packf, st);
  if (rc) return rc;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kCL); cfg.blockDim = dim3(kT); cfg.dynamicSmemBytes = SmemFwd::bytes; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kCL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  const float* packc = packf;
  DPI_CUDA(cudaLaunchKernelEx(&cfg, flow_cl_fwd_kernel, a, packc, scratch));
  count_launch();
  return DPI_OK;
}

int backward(const KArgs&, const Plan&, int, cudaStream_t) {
  set_error("flow_cl: backward kernel not built");
  return DPI_ERR_UNSUPPORTED;
}

}  // namespace fcl
}  // namespace dpi
