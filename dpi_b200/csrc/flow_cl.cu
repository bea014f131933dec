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
// stays inside one cluster and no hop waits on a barrier at all:
//   * swap-AB tcgen05.mma (kind::tf32, 3xTF32 split, M = 128 weight rows on the TMEM lanes, N = 64 samples on the
//     columns) - a hop's GEMM is 0.05-0.8 us on 16 SMs;
//   * net.0 (K = 512) and net.3 (K = 128) are split over K across the 16 CTAs; the partial accumulators are
//     reduce-scattered so that CTA c finalises hidden features 8c..8c+7 for ALL samples - BatchNorm batch statistics
//     are then reductions over 16 lanes of one CTA, no cross-CTA statistics exchange;
//   * net.6 (N = 1024) is split over output rows: CTA c owns coupling columns 32c..32c+31 (both the log-scale and
//     the shift row of each) and applies the tanh-exp coupling, ActNorm and the logdet partial as its epilogue; its
//     K loop consumes the 16 slices of the normalised hidden activations as they arrive from their owners;
//   * every hand-off is a PUSH: the producer stores to L2, fences the async proxy and issues cp.async.bulk with
//     .multicast::cluster straight into the consumers' shared memory, completing on the consumer's mbarrier
//     (complete_tx). Measured on B200 (tools/cl_bench.cu): 1.1-1.5 k cycles per all-to-all round against 4.6 k for
//     st.global + barrier.cluster + ld.global and 3.4 k for shared::cta -> shared::cluster bulk copies. The fixed
//     permutation between stages is part of the push: each output row goes to the CTA that consumes it next;
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

constexpr int kT = kThreadsCl;      // 16 worker warps + 1 control warp
constexpr int kTmemColsCl = 256;        // three 64-column accumulators (power of two >= 192)
constexpr int kTilePitch = 68;          // row pitch (floats) of the shared-memory exchange tiles: conflict-free float4 rows

// ----------------------------------------------------------------------------------------------- small PTX wrappers
__device__ __forceinline__ unsigned cta_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
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
// global -> the same shared-memory offset of every CTA in `mask`, completing on the mbarrier at the same offset there
__device__ __forceinline__ void bulk_g2s_mc(void* dst, const void* src, unsigned bytes, unsigned long long* bar, unsigned mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"((unsigned short)mask)
      : "memory");
}
__device__ __forceinline__ unsigned mapa_u32(unsigned laddr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(laddr), "r"(rank));
  return r;
}
// 16 bytes from registers straight into a peer CTA's shared memory; the peer's mbarrier counts them (complete_tx):
// no staging, no proxy fence, no bulk-copy issue slot (each cp.async.bulk costs ~65 cycles of a per-SM serial resource)
__device__ __forceinline__ void st_async4(unsigned raddr, float4 v, unsigned rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr),
               "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)), "r"(rbar)
               : "memory");
}
// generic-proxy global writes -> visible to the bulk-copy engine
__device__ __forceinline__ void fence_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// bounded wait: a protocol bug or a lost copy sets the sticky error flag (ctrl[2]) instead of hanging the GPU; the
// final phase turns the flag into NaN outputs so that the caller cannot miss it. One copy in the binary (the stage
// loop has to stay inside the instruction cache), polled by ONE lane per warp: mbarrier.try_wait suspends the whole
// warp, so a lane that has work to do (the MMA issuer) must never share a warp with polling lanes.
__device__ __noinline__ void mbar_wait_flag(unsigned long long* bar, unsigned parity, unsigned* ctrl) {
  const unsigned a = smem_u32(bar);
  unsigned done = 0;
  for (int it = 0; it < (1 << 24); ++it) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (done) return;
  }
  atomicExch(ctrl + 2, 1u);
}
__device__ __forceinline__ void warp_wait(unsigned long long* bar, unsigned parity, unsigned* ctrl) {
  if ((threadIdx.x & 31) == 0) mbar_wait_flag(bar, parity, ctrl);
  __syncwarp();
}
__device__ __noinline__ float exp_cl(float x) { return expf(x); }
__device__ __noinline__ float tanh_cl(float x) { return tanhf(x); }
__device__ __noinline__ void split4_cl(const float4& v, float4& hi, float4& lo) { split4(v, hi, lo); }
__device__ __forceinline__ void tmem_alloc_cl(unsigned* slot_smem) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "n"(kTmemColsCl) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cl(unsigned taddr) {      // the same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kTmemColsCl) : "memory");
}
// 16 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(unsigned taddr, float (&v)[16]) {
  unsigned r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// one K = 8 step of D[128 x 64] (+)= A * B^T in 3xTF32: a_lo b_hi + a_hi b_lo + a_hi b_hi (small terms first).
// Operand tiles are canonical K-major: 16-byte k-chunks 128 B apart (LBO), 8-row groups sbo bytes apart.
__device__ __noinline__ void issue_kstep(unsigned tmem_d, unsigned a_hi, unsigned a_lo, unsigned a_sbo, unsigned b_hi,
                                            unsigned b_lo, unsigned b_sbo, unsigned idesc, unsigned accumulate) {
  const unsigned long long dah = umma_desc(a_hi, 128, a_sbo), dal = umma_desc(a_lo, 128, a_sbo);
  const unsigned long long dbh = umma_desc(b_hi, 128, b_sbo), dbl = umma_desc(b_lo, 128, b_sbo);
  umma_tf32(tmem_d, dal, dbh, idesc, accumulate);
  umma_tf32(tmem_d, dah, dbl, idesc, 1u);
  umma_tf32(tmem_d, dah, dbh, idesc, 1u);
}
__device__ __forceinline__ void issue_gemm(unsigned tmem_d, unsigned a_hi, unsigned a_lo, unsigned a_sbo, unsigned b_hi,
                                           unsigned b_lo, unsigned b_sbo, int ksteps, unsigned idesc) {
#pragma unroll 1
  for (int ks = 0; ks < ksteps; ++ks)      // 8 k = two 16-byte chunks = 256 B further along the tile
    issue_kstep(tmem_d, a_hi + ks * 256, a_lo + ks * 256, a_sbo, b_hi + ks * 256, b_lo + ks * 256, b_sbo, idesc, ks ? 1u : 0u);
}

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void stcg4(float* p, float4 v) { __stcg(reinterpret_cast<float4*>(p), v); }

// ----------------------------------------------------------------------------------------------- weight pre-pass
// dst[(e * 16 + c) * kPackFloats ...] <- this (stage, CTA)'s three operand tile sets, each TF32 hi then lo:
//   forward : W1[h][32c + k]   one tile  128 x 32        W2[h][8c + k]   one tile 128 x 8
//             W3[row(r)][k]    16 k-tiles of 64 x 8 (the K loop of net.6 consumes one tile per arriving slice)
//   backward: W1^T [kl][h] = W1[h][32c + kl]   16 k-tiles of 32 x 8       W2^T [h1][kl] = W2[8c + kl][h1]   one tile 128 x 8
//             W3^T [h][jl] = W3[row(jl)][h]    one tile 128 x 64
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
    int q, layer, R, KS;       // rows and k extent of one tile of the layer's tile set
    float* tile;
    int hi_floats;
    if (q0 < C1) { q = q0; layer = 1; tile = blk; hi_floats = kW1Floats / 2; }
    else if (q0 < C1 + C2) { q = q0 - C1; layer = 2; tile = blk + kW1Floats; hi_floats = kW2Floats / 2; }
    else { q = q0 - C1 - C2; layer = 3; tile = blk + kW1Floats + kW2Floats; hi_floats = kW3Floats / 2; }
    if (!bwd) { R = layer == 3 ? 2 * kJB : H; KS = layer == 1 ? kKS1 : kHS; }
    else { R = layer == 1 ? kKS1 : H; KS = layer == 3 ? 2 * kJB : kHS; }
    const int per_tile = R * KS / 4, kcs = KS >> 2;
    const int kt = q / per_tile, qq = q - kt * per_tile;
    const int r8 = qq & 7, kc = (qq >> 3) % kcs, rg = (qq >> 3) / kcs;
    const int row = rg * 8 + r8, k = kt * KS + kc * 4;
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
    *reinterpret_cast<float4*>(tile + hi_floats + (size_t)q * 4) = lo;
  }
}

// ----------------------------------------------------------------------------------------------- shared memory
// byte offsets from the 1024-aligned dynamic base. Regions are reused along a stage; every reuse is ordered by the
// data dependences of the chain itself (a peer can only send hop X after it has received what this CTA sends after
// it stopped using the region), see the hazard notes at the use sites.
struct SmemFwd {
  static constexpr int oW1 = 0;                                  // 32 KB   W1 slice (hi | lo)
  static constexpr int oW2 = oW1 + kW1Floats * 4;                // 8 KB    W2 slice
  static constexpr int oW3 = oW2 + kW2Floats * 4;                // 64 KB   W3 slice: 16 hi k-tiles | 16 lo k-tiles (a descriptor
                                                                 //         is an M = 64 operand)
  static constexpr int oN2 = oW3 + kW3Floats * 4;                // 64 KB   n2 operand: 16 slices x [64 x 8] (hi | lo), pushed by
                                                                 //         their owners; its first 32 KB double as the
                                                                 //         reduce-scatter buffer of net.0
  static constexpr int oZ = oN2 + kCL * 2 * kNS * kHS * 4;       // 32 KB   a-slice operand (16 KB) -> reduce-scatter buffer of
                                                                 //         net.3 (32 KB) -> tail exchange tile (17 KB)
  static constexpr int oXR = oZ + kCL * kHS * kNS * 4;           // 16 KB   stage input rows pushed by their producers: 32 a + 32 b
  static constexpr int oN1 = oXR + 64 * kNS * 4;                 // 4 KB    n1 operand [64 x 8] (hi | lo)
  static constexpr int oEnd = oN1 + 2 * kNS * kHS * 4;
  static constexpr int bytes = oEnd + 128;                       // + alignment slack (no-swizzle operands need 16 B)
};
static_assert(64 * kTilePitch * 4 <= kCL * kHS * kNS * 4, "tail tile must fit the region it aliases");
static_assert(SmemFwd::bytes <= 222 * 1024, "dynamic shared memory budget");

// debug stamps: every CTA records %globaltimer (ns, common to all SMs) at 16 points per stage: tbuf[(cta * 32 + e) * 16 + slot]
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define CL_STAMP(slot)                                                                      \
  do {                                                                                      \
    if (a.tbuf && threadIdx.x == 0 && e < 32) a.tbuf[((size_t)c * 32 + e) * 16 + (slot)] = gtime(); \
  } while (0)

// mbarrier slots of the forward kernel
enum { B_W1 = 0, B_W2, B_W3, B_M1, B_M2, B_M3, B_X, B_R1, B_R2, B_AC, B_N1, B_AG = 16, B_COUNT = 32 };

constexpr int kWorkers = 512;                 // 16 worker warps; warp 16 is the control warp (MMA issue, weight prefetch)
__device__ __forceinline__ void bar_workers() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// per-stage scalars of one thread, loaded one stage ahead (parameters: independent of the chain)
struct StageRegs {
  float lsi, loc;                      // ActNorm of the stage
  float b1, g1, be1, rm1, rv1;         // hidden feature 8c + warp (warps 0..7): net.0 bias, BatchNorm 1
  float b2, g2, be2, rm2, rv2;
  float b3, sc3;                       // ZeroFC row of this thread's TMEM lane (lanes 0..63)
  int push_a, push_b;                  // consumer (cta * 64 + slot) of output rows 32c + p and Da + 32c + p, p = t >> 4
};
__device__ __forceinline__ void load_stage_regs(StageRegs& r, const KArgs& a, int e, int c, int t) {
  constexpr int Da = kDc / 2, Db = kDc / 2;
  const StageTab& st = a.st[e];
  const int warp = t >> 5, lane = t & 31;
  r.lsi = __ldg(a.P + st.lsi); r.loc = __ldg(a.P + st.loc);
  r.b1 = r.g1 = r.be1 = r.rm1 = r.rv1 = r.b2 = r.g2 = r.be2 = r.rm2 = r.rv2 = 0.f;
  if (warp < 8) {
    const int h = c * kHS + warp;
    r.b1 = __ldg(a.P + st.p[P_B1] + h); r.b2 = __ldg(a.P + st.p[P_B2] + h);
    if (a.has_bn) {
      r.g1 = __ldg(a.P + st.p[P_G1] + h); r.be1 = __ldg(a.P + st.p[P_BE1] + h);
      r.g2 = __ldg(a.P + st.p[P_G2] + h); r.be2 = __ldg(a.P + st.p[P_BE2] + h);
      r.rm1 = a.R[st.r[0] + h]; r.rv1 = a.R[st.r[1] + h];
      r.rm2 = a.R[st.r[2] + h]; r.rv2 = a.R[st.r[3] + h];
    }
  }
  r.b3 = 0.f; r.sc3 = 0.f;
  const int trow = lane < 16 ? (warp & 3) * 16 + lane : 2 * kJB;      // M = 64 accumulator: row r in TMEM lane r % 16 + 32 (r / 16)
  if (trow < 2 * kJB) {
    const int j3 = trow < kJB ? c * kJB + trow : Db + c * kJB + trow - kJB;
    r.b3 = __ldg(a.P + st.p[P_B3] + j3);
    r.sc3 = __ldg(a.P + st.p[P_SCALE] + j3);
  }
  r.push_a = r.push_b = 0;
  if (e + 1 < a.S) {
    r.push_a = __ldg(a.pmaps + (size_t)e * kDc + c * kKS1 + (t >> 4));
    r.push_b = __ldg(a.pmaps + (size_t)e * kDc + Da + c * kJB + (t >> 4));
  }
}

// Reduce-scatter epilogue of a hidden layer (warps 0..7: warp = hidden feature 8c + fl, lane = sample quad q + 16 x
// source half): sum the 16 K-split partials in CTA order, bias, LeakyReLU, BatchNorm (batch statistics over the 16
// lanes that share the feature / running statistics / none). rs: [16 sources][8 features][64 samples].
__device__ __forceinline__ float4 hidden_epilogue(const KArgs& a, const StageTab& st, int e, int which, int c, const float* rs,
                                                  float bias, float gamma, float beta, float rmean, float rvar, bool& valid) {
  const int fl = threadIdx.x >> 5, lane = threadIdx.x & 31, q = lane & 15, half = lane >> 4, h = c * kHS + fl;
  const int LB = a.ldb;
  const float* pp = rs + (size_t)(half * 8) * (kHS * kNS) + fl * kNS + 4 * q;
  float4 v[8];
#pragma unroll
  for (int s = 0; s < 8; ++s) v[s] = *reinterpret_cast<const float4*>(pp + (size_t)s * (kHS * kNS));
  float4 u = v[0];
#pragma unroll
  for (int s = 1; s < 8; ++s) { u.x += v[s].x; u.y += v[s].y; u.z += v[s].z; u.w += v[s].w; }
  {   // sources 0..7 + sources 8..15 (both halves end up with the same value)
    const float ox = __shfl_xor_sync(0xffffffffu, u.x, 16), oy = __shfl_xor_sync(0xffffffffu, u.y, 16);
    const float oz = __shfl_xor_sync(0xffffffffu, u.z, 16), ow = __shfl_xor_sync(0xffffffffu, u.w, 16);
    if (half == 0) { u.x += ox; u.y += oy; u.z += oz; u.w += ow; }
    else { u.x = ox + u.x; u.y = oy + u.y; u.z = oz + u.z; u.w = ow + u.w; }
  }
  valid = 4 * q < a.B;
  const float l0 = lrelu_f(u.x + bias), l1 = lrelu_f(u.y + bias), l2 = lrelu_f(u.z + bias), l3 = lrelu_f(u.w + bias);
  const bool writer = valid && half == 0;
  const size_t row = ((size_t)e * kHc + h) * LB + 4 * q;
  if (writer) stcg4(a.W + (which == 1 ? a.oL1T : a.oL2T) + row, make_float4(l0, l1, l2, l3));
  float n0 = l0, n1 = l1, n2 = l2, n3 = l3;
  if (a.has_bn) {
    if (a.train) {
      const float invB = 1.0f / (float)a.B;
      const float mean = half_warp_sum(valid ? (l0 + l1) + (l2 + l3) : 0.f) * invB;
      const float d0 = l0 - mean, d1 = l1 - mean, d2 = l2 - mean, d3 = l3 - mean;
      const float var = half_warp_sum(valid ? (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3) : 0.f) * invB;
      const float rstd = 1.0f / sqrtf(var + kBnEps);
      const float ga = gamma * rstd;
      n0 = fmaf(d0, ga, beta); n1 = fmaf(d1, ga, beta); n2 = fmaf(d2, ga, beta); n3 = fmaf(d3, ga, beta);
      if (lane == 0) {
        float* stt = a.W + a.oST + ((size_t)e * 4 + (which - 1) * 2) * kHc;
        __stcg(stt + h, mean);
        __stcg(stt + kHc + h, rstd);
        a.R[st.r[(which - 1) * 2] + h] = (1.f - kBnMomentum) * rmean + kBnMomentum * mean;
        a.R[st.r[(which - 1) * 2 + 1] + h] = (1.f - kBnMomentum) * rvar + kBnMomentum * (var * (float)a.B / (float)(a.B - 1));
      }
    } else {
      const float rs_ = 1.0f / sqrtf(rvar + kBnEps);
      const float ga = gamma * rs_, be = beta - rmean * ga;
      n0 = fmaf(l0, ga, be); n1 = fmaf(l1, ga, be); n2 = fmaf(l2, ga, be); n3 = fmaf(l3, ga, be);
    }
  }
  const float4 nrm = make_float4(n0, n1, n2, n3);
  if (writer) stcg4(a.W + (which == 1 ? a.oN1T : a.oN2T) + row, nrm);
  return nrm;
}

// this CTA's K-split partial accumulator (TMEM lane = hidden feature, column = sample): rows 8d..8d+7 go straight from
// registers into slot c of CTA d's reduce-scatter buffer (st.async through distributed shared memory)
__device__ __forceinline__ void send_partial(unsigned tmem_d, int c, float* rsbuf, unsigned long long* bar,
                                             unsigned long long* stamp = nullptr) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = (warp & 3) * 32 + lane, c0 = (warp >> 2) * 16;
  float v[16];
  tmem_ld16(tmem_d + ((unsigned)((warp & 3) * 32) << 16) + (unsigned)c0, v);
  if (stamp) stamp[0] = gtime();
  // 4 x 4 transpose of 16-byte chunks inside every group of four lanes (rows 4g..4g+3): afterwards lane k of a group
  // holds chunk k of each of the group's rows, so that one st.async writes 64 contiguous bytes of ONE row per group
  // instead of 32 scattered 16-byte pieces (measured: 2.4 k cycles to issue the scattered form, DSMEM request-bound)
  float4 ch[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) ch[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  auto xchg = [](float4 x, int mask) {
    return make_float4(__shfl_xor_sync(0xffffffffu, x.x, mask), __shfl_xor_sync(0xffffffffu, x.y, mask),
                       __shfl_xor_sync(0xffffffffu, x.z, mask), __shfl_xor_sync(0xffffffffu, x.w, mask));
  };
  {   // bit 0: lanes with b0 = 0 keep even chunks; slot j holds (row r ^ (j & 1), chunk (j & ~1) | b0)
    const int b0 = lane & 1;
    const float4 s0 = b0 ? ch[0] : ch[1], s1 = b0 ? ch[2] : ch[3];
    const float4 r0 = xchg(s0, 1), r1 = xchg(s1, 1);
    if (b0) { ch[0] = r0; ch[2] = r1; } else { ch[1] = r0; ch[3] = r1; }
    // now: b0 = 0: ch[0] = (own, c0), ch[1] = (peer, c0), ch[2] = (own, c2), ch[3] = (peer, c2)
    //      b0 = 1: ch[0] = (peer, c1), ch[1] = (own, c1), ch[2] = (peer, c3), ch[3] = (own, c3)
  }
  {   // bit 1: lanes with b1 = 0 keep the low chunk pair
    const int b1 = (lane >> 1) & 1;
    const float4 s0 = b1 ? ch[0] : ch[2], s1 = b1 ? ch[1] : ch[3];
    const float4 r0 = xchg(s0, 2), r1 = xchg(s1, 2);
    if (b1) { ch[0] = r0; ch[1] = r1; } else { ch[2] = r0; ch[3] = r1; }
  }
  // lane L = (b1, b0) of a group now holds chunk k = L of all four rows, and slot j is row j of the group (step 1 moved
  // the slots whose low bit differs from b0, step 2 those whose high bit differs from b1: row = L ^ (j ^ L) = j)
  const int k = lane & 3, g4 = m & ~3;
  const unsigned d = (unsigned)m >> 3;
  const unsigned rbar = mapa_u32(smem_u32(bar), d);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int row = g4 + j;
    const unsigned dst = mapa_u32(smem_u32(rsbuf + (size_t)c * (kHS * kNS) + (row & 7) * kNS + c0 + 4 * k), d);
    st_async4(dst, ch[j], rbar);
  }
  if (stamp) stamp[1] = gtime();
}

// inline bounded wait of the control thread (its loop must stay tight: it is the only serial instruction stream)
__device__ __forceinline__ void ctl_wait(unsigned long long* bar, unsigned parity, unsigned* ctrl) {
  const unsigned ad = smem_u32(bar);
  unsigned done = 0;
#pragma unroll 1
  for (int it = 0; it < (1 << 24) && !done; ++it)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(ad), "r"(parity) : "memory");
  if (!done) atomicExch(ctrl + 2, 1u);
}
// instruction descriptor: D fp32, A / B tf32 K-major, M x N (tile_cp.cuh's umma_idesc fixes M = 128)
__device__ __forceinline__ unsigned idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(m >> 4) << 24);
}
__device__ __forceinline__ void mma3(unsigned tmem_d, unsigned long long ah, unsigned long long al, unsigned long long bh,
                                     unsigned long long bl, unsigned idesc, unsigned accumulate) {
  umma_tf32(tmem_d, al, bh, idesc, accumulate);      // small terms first
  umma_tf32(tmem_d, ah, bl, idesc, 1u);
  umma_tf32(tmem_d, ah, bh, idesc, 1u);
}

}  // namespace

// =================================================================================================== forward kernel
__global__ void __launch_bounds__(kT, 1) flow_cl_fwd_kernel(const __grid_constant__ KArgs a, const float* __restrict__ pack,
                                                             float* __restrict__ scratch) {
  extern __shared__ unsigned char raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 127) & ~(uintptr_t)127);
  __shared__ __align__(8) unsigned long long bars[B_COUNT];
  __shared__ unsigned tmem_slot;
  __shared__ float sLd[64];                                // running logdet partial of this CTA
  __shared__ __align__(16) float sRed[16 * 64];

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int c = (int)cta_rank();
  const int LB = a.ldb, B = a.B, S = a.S;
  constexpr int D = kDc, Da = kDc / 2, Db = kDc / 2;
  float* agbuf = scratch + kPartFloats;
  float* ldrows = agbuf + kAgFloats;

  if (t == 0) {
    for (int i = 0; i < B_COUNT; ++i) mbar_init(bars + i, i == B_AC ? 16 : (i == B_N1 ? 8 : 1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc_cl(&tmem_slot);
  if (t < 64) sLd[t] = 0.f;
  float* AC = reinterpret_cast<float*>(sm + SmemFwd::oZ);
  float* RS1 = reinterpret_cast<float*>(sm + SmemFwd::oN2);
  float* RS2 = reinterpret_cast<float*>(sm + SmemFwd::oZ);
  float* TT = reinterpret_cast<float*>(sm + SmemFwd::oZ);
  float* XR = reinterpret_cast<float*>(sm + SmemFwd::oXR);
  float* N1C = reinterpret_cast<float*>(sm + SmemFwd::oN1);
  float* N2C = reinterpret_cast<float*>(sm + SmemFwd::oN2);
  // rows are pushed LB floats at a time: the padding samples of every row must read as zero
  for (int i = t; i < 64 * kNS; i += kT) XR[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem = tmem_slot;
  const unsigned tD1 = tmem, tD2 = tmem + 64, tD3 = tmem + 128;

  // ---- stage 0 input rows straight from z [B][D] (row slot s <- feature gmap[j'(s)]), and z^T for the backward pass
  if (t < kWorkers) {
    const int n = t & 63;
#pragma unroll 1
    for (int s = t >> 6; s < 64; s += kWorkers / 64) {
      const int f = __ldg(a.gmaps + (s < 32 ? c * kKS1 + s : Da + c * kJB + (s - 32)));
      if (n < B) {
        const float v = __ldg(a.in + (size_t)n * D + f);
        XR[s * kNS + n] = v;
        __stcg(a.W + a.oZT + (size_t)f * LB + n, v);
      }
    }
  }
  // every CTA's barriers are initialised (and its input rows in place) before any peer can push into it
  cluster_sync_all();

  if (warp == kWorkers / 32) {
    // ================================================================ control warp: one thread issues every MMA
    if (lane == 0) {
      const unsigned sW1 = smem_u32(sm + SmemFwd::oW1), sW2 = smem_u32(sm + SmemFwd::oW2), sW3 = smem_u32(sm + SmemFwd::oW3);
      const unsigned sN2 = smem_u32(N2C), sAC = smem_u32(AC), sN1 = smem_u32(N1C);
      const unsigned idesc = idesc_tf32(128, kNS);
      // net.6 has 64 output rows per CTA: M = 64 halves the A-operand fetch (these K = 8 MMAs are bound by the tensor core's
      // shared-memory operand bandwidth, ~64 B/clk measured: (M + N) * 32 B per instruction)
      const unsigned idesc64 = idesc_tf32(64, kNS);
      // operand descriptors (canonical K-major, LBO 128 B); a K = 8 step is 256 B = 16 descriptor units further on
      const unsigned long long dW1h = umma_desc(sW1, 128, kKS1 * 32), dW1l = umma_desc(sW1 + kW1Floats * 2, 128, kKS1 * 32);
      const unsigned long long dACh = umma_desc(sAC, 128, kKS1 * 32), dACl = umma_desc(sAC + kNS * kKS1 * 4, 128, kKS1 * 32);
      const unsigned long long dW2h = umma_desc(sW2, 128, kHS * 32), dW2l = umma_desc(sW2 + kW2Floats * 2, 128, kHS * 32);
      const unsigned long long dN1h = umma_desc(sN1, 128, kHS * 32), dN1l = umma_desc(sN1 + kNS * kHS * 4, 128, kHS * 32);
      const unsigned long long dW3h = umma_desc(sW3, 128, kHS * 32), dW3l = umma_desc(sW3 + kW3Floats * 2, 128, kHS * 32);
      const unsigned long long dN2h = umma_desc(sN2, 128, kHS * 32), dN2l = umma_desc(sN2 + kNS * kHS * 4, 128, kHS * 32);
      {
        const float* blk = pack + (size_t)c * kPackFloats;
        mbar_expect_tx(bars + B_W1, kW1Floats * 4); bulk_g2s(sm + SmemFwd::oW1, blk, kW1Floats * 4, bars + B_W1);
        mbar_expect_tx(bars + B_W2, kW2Floats * 4); bulk_g2s(sm + SmemFwd::oW2, blk + kW1Floats, kW2Floats * 4, bars + B_W2);
        mbar_expect_tx(bars + B_W3, kW3Floats * 4); bulk_g2s(sm + SmemFwd::oW3, blk + kW1Floats + kW2Floats, kW3Floats * 4, bars + B_W3);
      }
#pragma unroll 1
      for (int e = 0; e < S; ++e) {
        const unsigned par = e & 1;
        const float* nxt = pack + ((size_t)(e + 1) * kCL + c) * kPackFloats;
        // ---- net.0: D1[h][n] = sum_k W1[h][32c + k] a[k][n]  (this CTA's K slice)
        ctl_wait(bars + B_W1, par, a.ctrl);
        ctl_wait(bars + B_AC, par, a.ctrl);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < kKS1 / 8; ++ks) mma3(tD1, dW1h + 16 * ks, dW1l + 16 * ks, dACh + 16 * ks, dACl + 16 * ks, idesc, ks ? 1u : 0u);
        umma_commit(bars + B_M1);
        ctl_wait(bars + B_M1, par, a.ctrl);
        if (e + 1 < S) { mbar_expect_tx(bars + B_W1, kW1Floats * 4); bulk_g2s(sm + SmemFwd::oW1, nxt, kW1Floats * 4, bars + B_W1); }
        // ---- net.3: D2[h2][n] = sum_k W2[h2][8c + k] n1[8c + k][n]
        ctl_wait(bars + B_W2, par, a.ctrl);
        ctl_wait(bars + B_N1, par, a.ctrl);
        tc_fence_after();
        mma3(tD2, dW2h, dW2l, dN1h, dN1l, idesc, 0u);
        umma_commit(bars + B_M2);
        ctl_wait(bars + B_M2, par, a.ctrl);
        if (e + 1 < S) { mbar_expect_tx(bars + B_W2, kW2Floats * 4); bulk_g2s(sm + SmemFwd::oW2, nxt + kW1Floats, kW2Floats * 4, bars + B_W2); }
        // ---- net.6.fc: D3[r][n] = sum_k W3[row(r)][k] n2[k][n], rows 0..63 real; k-step s as soon as slice s has landed
        ctl_wait(bars + B_W3, par, a.ctrl);
#pragma unroll 4
        for (int s = 0; s < kCL; ++s) {
          ctl_wait(bars + B_AG + s, par, a.ctrl);
          tc_fence_after();
          // W3 k-tile s: 64 x 8 floats = 2 KB = 128 units; n2 slice s: (hi | lo) = 4 KB = 256 units
          mma3(tD3, dW3h + 128ull * s, dW3l + 128ull * s, dN2h + 256ull * s, dN2l + 256ull * s, idesc64, s ? 1u : 0u);
        }
        umma_commit(bars + B_M3);
        ctl_wait(bars + B_M3, par, a.ctrl);
        if (e + 1 < S) {
          mbar_expect_tx(bars + B_W3, kW3Floats * 4);
          bulk_g2s(sm + SmemFwd::oW3, nxt + kW1Floats + kW2Floats, kW3Floats * 4, bars + B_W3);
        }
      }
    }
  } else {
    // ================================================================ worker warps
    StageRegs cur, nx;
    load_stage_regs(nx, a, 0, c, t);
    const int trow = lane < 16 ? (warp & 3) * 16 + lane : 2 * kJB;   // output row of net.6 held by this thread's TMEM lane
#pragma unroll 1
    for (int e = 0; e < S; ++e) {
      const StageTab& st = a.st[e];
      float* y = a.W + a.oYT + (size_t)e * D * LB;
      const unsigned par = e & 1;
      CL_STAMP(0);
      cur = nx;
      if (e + 1 < S) load_stage_regs(nx, a, e + 1, c, t);  // in flight for a whole stage
      const float qa = expf(cur.lsi), qc = -cur.loc;
      const float e3v = expf(3.f * cur.sc3);
      // arm this stage's exchange barriers, one thread each (a push that lands first just leaves the tx-count negative)
      if (t == 32) mbar_expect_tx(bars + B_R1, kCL * kHS * kNS * 4);
      else if (t == 64) mbar_expect_tx(bars + B_R2, kCL * kHS * kNS * 4);
      else if (t >= 96 && t < 96 + kCL) mbar_expect_tx(bars + B_AG + (t - 96), 2 * kNS * kHS * 4);
      // ---- input rows (pushed by the previous stage's owners) -> canonical B operand of net.0 (hi / lo) + the
      //      pass-through rows of y; this thread's four b values -> registers
      if (e > 0) warp_wait(bars + B_X, (e - 1) & 1, a.ctrl);
      CL_STAMP(10);
      float4 bv, ya;
      {
        const int p = t >> 4, s0 = (t & 15) * 4;
        bv = *reinterpret_cast<const float4*>(XR + (32 + p) * kNS + s0);
        const float4 av = *reinterpret_cast<const float4*>(XR + p * kNS + s0);
        // pass-through half: ActNorm.reverse only (:66); stored now, pushed to its consumer with the b rows at the tail
        ya = make_float4(fmaf(av.x, qa, qc), fmaf(av.y, qa, qc), fmaf(av.z, qa, qc), fmaf(av.w, qa, qc));
        if (s0 < B) stcg4(y + (size_t)(c * kKS1 + p) * LB + s0, ya);
        else ya = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      {
        const int n = t & 63, kc = t >> 6;                 // one 16-byte chunk (sample n, features 4kc..4kc+3) per thread
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = XR[(4 * kc + u) * kNS + n];
        float4 hi, lo;
        split4(make_float4(v[0], v[1], v[2], v[3]), hi, lo);
        const int off = (n >> 3) * 256 + kc * 32 + (n & 7) * 4;      // [64 x 32] tile: SBO 1024 B, LBO 128 B
        *reinterpret_cast<float4*>(AC + off) = hi;
        *reinterpret_cast<float4*>(AC + kNS * kKS1 + off) = lo;
      }
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + B_AC);
      CL_STAMP(1);
      // ---- net.0 partial -> reduce-scatter. RS1 aliases the n2 operand: its last readers were this CTA's net.6 MMAs of
      //      the previous stage (complete), and a peer's partial can only arrive after that peer received this CTA's rows
      warp_wait(bars + B_M1, par, a.ctrl);
      tc_fence_after();
      CL_STAMP(2);
      send_partial(tD1, c, RS1, bars + B_R1);
      if (t == 0 && e + 1 < S) mbar_expect_tx(bars + B_X, 64 * kNS * 4);     // the next stage's rows (full 64-sample rows)
      warp_wait(bars + B_R1, par, a.ctrl);
      CL_STAMP(3);
      if (warp < 8) {
        bool valid;
        const float4 nr = hidden_epilogue(a, st, e, 1, c, RS1, cur.b1, cur.g1, cur.be1, cur.rm1, cur.rv1, valid);
        CL_STAMP(15);
        // canonical B operand of net.3: [64 samples x 8 k], k = this CTA's hidden features (SBO 256 B, LBO 128 B)
        if (lane < 16) {
          const int fl = warp, q = lane;
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
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + B_N1);
      }
      CL_STAMP(4);
      // ---- net.3 partial -> reduce-scatter. RS2 aliases the a-slice operand: net.0's MMAs are complete, and a peer's
      //      net.3 partial can only arrive after it received this CTA's net.0 partial, sent after them
      warp_wait(bars + B_M2, par, a.ctrl);
      tc_fence_after();
      send_partial(tD2, c, RS2, bars + B_R2);
      warp_wait(bars + B_R2, par, a.ctrl);
      CL_STAMP(5);
      if (warp < 8) {
        bool valid;
        const float4 nr = hidden_epilogue(a, st, e, 2, c, RS2, cur.b2, cur.g2, cur.be2, cur.rm2, cur.rv2, valid);
        CL_STAMP(11);
        // n2 rows 8c..8c+7 -> one [64 x 8] operand tile (hi | lo) in L2: slice c of net.6's K loop on every CTA
        if (lane < 16) {
          const int fl = warp, q = lane;
          const float nv[4] = {nr.x, nr.y, nr.z, nr.w};
          float* dst = agbuf + (size_t)c * (2 * kNS * kHS);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int n = 4 * q + i;
            unsigned hi, lo;
            split_tf32(valid ? nv[i] : 0.f, hi, lo);
            const int off = (n >> 3) * 64 + (fl >> 2) * 32 + (n & 7) * 4 + (fl & 3);
            __stcg(dst + off, __uint_as_float(hi));
            __stcg(dst + kNS * kHS + off, __uint_as_float(lo));
          }
          CL_STAMP(12);
          fence_async_global();
        }
      }
      bar_workers();
      CL_STAMP(13);
      if (t == 0) bulk_g2s_mc(N2C + (size_t)c * (2 * kNS * kHS), agbuf + (size_t)c * (2 * kNS * kHS), 2 * kNS * kHS * 4, bars + B_AG + c, 0xffffu);
      CL_STAMP(6);
      // ---- net.6 accumulator -> tail
      warp_wait(bars + B_M3, par, a.ctrl);
      tc_fence_after();
      CL_STAMP(7);
      // step 1: ZeroFC bias + exp(3 scale) (:137-141) on this thread's output row -> exchange tile (aliases RS2: consumed)
      {
        const int c0 = (warp >> 2) * 16;
        float v[16];
        tmem_ld16(tD3 + ((unsigned)((warp & 3) * 32) << 16) + (unsigned)c0, v);     // warp-collective: lanes >= 16 idle
        if (trow < 2 * kJB) {
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(TT + trow * kTilePitch + c0 + i) =
                make_float4((v[i] + cur.b3) * e3v, (v[i + 1] + cur.b3) * e3v, (v[i + 2] + cur.b3) * e3v, (v[i + 3] + cur.b3) * e3v);
        }
      }
      tc_fence_before();
      bar_workers();
      CL_STAMP(8);
      // step 2: coupling (:244-249) + ActNorm.reverse (:66) + logdet partial; thread = (column p, 4 samples)
      {
        const int p = t >> 4, s0 = (t & 15) * 4;
        const int col = c * kJB + p;
        const float4 ol = *reinterpret_cast<const float4*>(TT + p * kTilePitch + s0);
        const float4 ot = *reinterpret_cast<const float4*>(TT + (kJB + p) * kTilePitch + s0);
        const float ols[4] = {ol.x, ol.y, ol.z, ol.w}, ots[4] = {ot.x, ot.y, ot.z, ot.w}, bvs[4] = {bv.x, bv.y, bv.z, bv.w};
        float yb[4], dl[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float ls = tanhf(ols[i]);
          const float bp = bvs[i] / expf(ls) - ots[i];
          yb[i] = fmaf(bp, qa, qc);
          dl[i] = s0 + i < B ? -ls : 0.f;
        }
        float4 yb4 = make_float4(yb[0], yb[1], yb[2], yb[3]);
        if (s0 < B) {
          float* O = a.W + a.oOT + (size_t)e * (2 * Db) * LB;
          stcg4(O + (size_t)col * LB + s0, ol);
          stcg4(O + (size_t)(Db + col) * LB + s0, ot);
          stcg4(y + (size_t)(Da + col) * LB + s0, yb4);
        } else {
          yb4 = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // push this thread's quads of output rows 32c + p (pass-through) and Da + 32c + p (coupled) to the CTAs that
        // consume them in the next stage: the inter-stage permutation (:599, :459) is this routing
        if (e + 1 < S) {
          const unsigned da = (unsigned)cur.push_a >> 6, db = (unsigned)cur.push_b >> 6;
          st_async4(mapa_u32(smem_u32(XR + (cur.push_a & 63) * kNS + s0), da), ya, mapa_u32(smem_u32(bars + B_X), da));
          st_async4(mapa_u32(smem_u32(XR + (cur.push_b & 63) * kNS + s0), db), yb4, mapa_u32(smem_u32(bars + B_X), db));
        }
        // per-sample sum over this CTA's 32 columns: 2 columns per warp (lanes 16 apart), then the 16 warps in order
#pragma unroll
        for (int i = 0; i < 4; ++i) dl[i] += __shfl_xor_sync(0xffffffffu, dl[i], 16);
        if (lane < 16) *reinterpret_cast<float4*>(sRed + warp * 64 + lane * 4) = make_float4(dl[0], dl[1], dl[2], dl[3]);
      }
      bar_workers();
      if (t < 64) {
        float sum = 0.f;
#pragma unroll
        for (int r = 0; r < 16; ++r) sum += sRed[r * 64 + t];
        sLd[t] += sum;
      }
      CL_STAMP(9);
    }
  }

  // ---- logdet: per-CTA partial rows -> CTA 0 sums them in CTA order  (+ D * sum_e log_scale_inv_e, :60-64)
  __syncthreads();
  if (t < 64) __stcg(ldrows + (size_t)c * kNS + t, sLd[t]);
  cluster_sync_all();
  // ---- YT[S - 1] [D][LB] -> out [B][D]: this CTA's 64 features, 32 x 32 tiles through shared memory
  {
    float* sT = TT;       // 32 x 33 floats of the (now idle) exchange tile
    const float* ylast = a.W + a.oYT + (size_t)(S - 1) * D * LB;
    const int lx = t & 31, ly = t >> 5;    // 16 rows per pass (worker warps)
    for (int tile = 0; tile < 4; ++tile) {
      const int f0 = c * 64 + (tile & 1) * 32, m0 = (tile >> 1) * 32;
      __syncthreads();
      if (t < kWorkers) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int f = f0 + ly + 16 * i, m = m0 + lx;
          sT[(ly + 16 * i) * 33 + lx] = m < B ? __ldcg(ylast + (size_t)f * LB + m) : 0.f;
        }
      }
      __syncthreads();
      if (t < kWorkers) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int m = m0 + ly + 16 * i;
          if (m < B) a.out[(size_t)m * D + f0 + lx] = sT[lx * 33 + ly + 16 * i];
        }
      }
    }
  }
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
  cluster_sync_all();      // no CTA leaves while a peer's copy into it could still be in flight
}

// =================================================================================================== host side
namespace {
int g_query_device = -1, g_query_ok = 0;

cudaLaunchConfig_t cluster_config(cudaLaunchAttribute* at, int smem, cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kCL); cfg.blockDim = dim3(kT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kCL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cfg;
}
}  // namespace

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
  cudaLaunchAttribute at[1];
  cudaLaunchConfig_t cfg = cluster_config(at, SmemFwd::bytes, nullptr);
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
  cudaLaunchAttribute at[1];
  cudaLaunchConfig_t cfg = cluster_config(at, SmemFwd::bytes, st);
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
