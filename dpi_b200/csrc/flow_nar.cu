// flow_nar.cu - RealNVP.reverse and its backward for NARROW flows as sample-resident persistent kernels.
//
// Math: generative_model/realnvpfc_model.py AffineCoupling.reverse :240-249, net :171-187, ZeroFC :137-141,
// ActNorm.reverse :49-66, Flow.reverse :457-474, RealNVP.reverse :594-603; backward per SURVEY.md Appendix A.
//
// Why: for the alpha-DPI flow (DPIx_interferometry.py:250: 18 dimensions, hidden 144, batch 2048) and the 2-D toy
// posterior a coupling stage is a few hundred kFLOP per sample slab - the tcgen05 GEMM path spends its time in ~28
// launches per stage (1200 per step, launch-bound), and its tensor-core accumulation misses the fp32 bar on the
// ndim-2 net. Here every CTA keeps a slab of samples in shared memory through ALL stages of the pass:
//   * the fixed permutations between stages are row re-indexing inside the CTA (all features of a sample are local);
//   * the three Linear layers are fp32 FFMA micro-GEMMs (4 x 4 register tiles) against the stage's weights in shared
//     memory (k-major copies from a transposing pre-pass, so a thread tile's weights are one conflict-free LDS.128);
//     the next stage's operands arrive by cp.async.bulk (one elected thread, mbarrier complete_tx) while the current
//     stage computes;
//   * the ONLY cross-CTA traffic is the BatchNorm batch statistics (per layer one (mean, M2) pair per feature forward,
//     two sums per feature backward): DSMEM reduction inside a cluster of 8 -> one slot per cluster in L2 -> grid
//     barrier -> every CTA combines the slots in cluster order (deterministic, bit-reproducible). The forward
//     statistics are merged as (count, mean, M2) triples (Chan et al.), so fp32 is enough and nothing cancels;
//   * weight / bias / scale / ActNorm gradients are NOT formed in the chain: the backward chain stores the per-stage
//     gradient rows and ONE batched kernel (nar_pgrad_kernel) reduces them over the batch afterwards (no barriers).
// Eval mode (running statistics) and batch_norm=False need no cross-CTA exchange at all: posterior sampling runs the
// slabs independently. Same feature-major workspace as the other paths (flow.cuh): the forward pass is
// interchangeable with them and its saved activations feed either backward.
#include <algorithm>

#include "flow_host.cuh"
#include "flow_nar.cuh"

namespace dpi {
namespace nar {

struct NArgs {
  KArgs k;
  Plan p;
};

namespace {

constexpr int kT = kThreadsN;

// ------------------------------------------------------------------------------- synchronisation / copies
__device__ __forceinline__ unsigned cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem(const float* p, unsigned rank) {
  const unsigned l = (unsigned)__cvta_generic_to_shared(p);
  unsigned r;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(l), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(r) : "memory");
  return v;
}
// Grid barrier on a monotonic counter (ctrl[0], zeroed by the host before the launch; cooperative launch guarantees
// co-residency). Bounded spin: a protocol bug sets the sticky flag ctrl[2], which the final phase turns into NaN.
__device__ __forceinline__ void grid_barrier(unsigned* ctrl, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctrl), "r"(1u) : "memory");
    unsigned cur;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cur) : "l"(ctrl) : "memory");
    if (cur < target) {
      const long long t0 = clock64();
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cur) : "l"(ctrl) : "memory");
        if (cur < target && clock64() - t0 > 4000000000LL) {
          atomicExch(ctrl + 2, 1u);
          break;
        }
      } while (cur < target);
    }
  }
  __syncthreads();
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void stcg4(float* p, float4 v) { __stcg(reinterpret_cast<float4*>(p), v); }

// (row, column) iteration over an [R][cols] index space without integer division in the loop: item = threadIdx.x + j * 256
struct RowQuad {
  int r, q, dr, dq, tiles;
  __device__ __forceinline__ explicit RowQuad(int cols) {
    tiles = cols;
    r = threadIdx.x / tiles; q = threadIdx.x - r * tiles;
    dr = kT / tiles; dq = kT - dr * tiles;
  }
  __device__ __forceinline__ void next() {
    r += dr; q += dq;
    if (q >= tiles) { q -= tiles; ++r; }
  }
};

// ------------------------------------------------------------------------------- per-stage operands in shared memory
// One stage block mirrors the flat parameter layout so that each piece is ONE bulk copy:
//   [first weight tile | b1 g1 be1 (stride Hp) | b2 g2 be2 | scale, W3 [Dout][H], b3 | 4 statistics vectors | idx | lsi, loc]
// forward: first tile = W1^T [Da4][H] (pre-pass), statistics = running mean / var of both BatchNorms (stride Hp);
// backward: first tile = W1 [H][Da] as stored, statistics = the batch statistics the forward pass saved (stride H).
struct Blk {
  float *W1, *v1, *v2, *scale, *W3, *b3, *stat;
  int* idx;
  float* act;   // [0] = log_scale_inv, [1] = loc
};
__device__ __forceinline__ Blk blk_at(float* base, const Plan& p, int Dout, int H) {
  Blk b;
  b.W1 = base + p.bW1; b.v1 = base + p.bV1; b.v2 = base + p.bV2;
  b.scale = base + p.bSC;
  const int DoutP = (Dout + 31) & ~31, W3P = (Dout * H + 31) & ~31;
  b.W3 = b.scale + DoutP;
  b.b3 = b.W3 + W3P;
  b.stat = base + p.bST;
  b.idx = reinterpret_cast<int*>(base + p.bIdx);
  b.act = base + p.bAct;
  return b;
}
// elected thread: bulk copies of stage e's block; threads 0..D: the (unaligned) index row and the ActNorm scalars
__device__ __forceinline__ void issue_block(const KArgs& a, const Plan& p, const StageTab* stb, const Blk& b, int e, bool bwd,
                                            unsigned long long* bar) {
  const StageTab& st = stb[e];
  const int H = a.H, Dout = a.Dout, D = a.D, Da = a.Da, t = threadIdx.x;
  if (t == 0) {
    const unsigned w1b = bwd ? (unsigned)(((H * Da + 31) & ~31) * 4) : (unsigned)(p.Da4 * H * 4);
    const unsigned vb = (unsigned)((a.has_bn ? 3 : 1) * p.Hp * 4);
    const unsigned scb = (unsigned)((2 * ((Dout + 31) & ~31) + ((Dout * H + 31) & ~31)) * 4);
    const bool stats = a.has_bn && (bwd ? a.train != 0 : a.R != nullptr);
    const unsigned stb_ = stats ? (unsigned)((bwd ? 4 * H : 4 * p.Hp) * 4) : 0u;
    mbar_expect_tx(bar, w1b + 2 * vb + scb + stb_);
    if (bwd) bulk_g2s(b.W1, a.P + st.p[P_W1], w1b, bar);
    else bulk_g2s(b.W1, a.W + p.oWT + (size_t)e * ((size_t)(p.Da4 + H) * H), w1b, bar);
    bulk_g2s(b.v1, a.P + st.p[P_B1], vb, bar);
    bulk_g2s(b.v2, a.P + st.p[P_B2], vb, bar);
    bulk_g2s(b.scale, a.P + st.p[P_SCALE], scb, bar);
    if (stats) {
      if (bwd) bulk_g2s(b.stat, a.W + a.oST + (size_t)e * 4 * H, stb_, bar);
      else bulk_g2s(b.stat, a.R + st.r[0], stb_, bar);
    }
    cp_async4(b.act, a.P + st.lsi, true);
    cp_async4(b.act + 1, a.P + st.loc, true);
  }
  for (int i = t; i < D; i += kT)
    cp_async4(reinterpret_cast<float*>(b.idx + i), reinterpret_cast<const float*>(a.gmaps + (size_t)e * D + i), true);
}
__device__ __forceinline__ void issue_w2(float* W2, const float* src, int H, unsigned long long* bar) {
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, (unsigned)(H * H * 4));
    bulk_g2s(W2, src, (unsigned)(H * H * 4), bar);
  }
}

// ------------------------------------------------------------------------------- micro-GEMMs (fp32 FFMA)
// out[n][s] = sum_k WT[k][n] * X(k)[s]  (+ bias[n], LeakyReLU) for n < N (multiple of 4), s < SP; WT rows are n-contiguous
// (pitch pw). X(k) = Xb + row(k) * SP with row(k) = idx ? idx[k] : k. Thread tile 4 features x 4 samples: per k one
// LDS.128 of the weights (consecutive lanes -> consecutive 16 B: conflict-free) and one of the activations, 16 FMA.
// The same routine is the data gradient through a Linear stored [n][k]: out[k][s] = sum_n W[n][k] G[n][s] (WT := W).
__device__ __forceinline__ void gemm_kn(float* out, const float* WT, int pw, const float* Xb, const int* idx, int N, int K, int SP,
                                        const float* bias, bool lrelu) {
  const int tiles_s = SP >> 2, tiles_n = N >> 2;
  for (int tile = threadIdx.x; tile < tiles_n * tiles_s; tile += kT) {
    const int tn = tile / tiles_s, ts = tile - tn * tiles_s;
    float acc[4][4];
#pragma unroll
    for (int f = 0; f < 4; ++f)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[f][i] = 0.f;
    const float* wp = WT + 4 * tn;
    const float* xp = Xb + 4 * ts;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const float4 w = ld4(wp + k * pw);
      const float4 x = ld4(xp + (idx ? idx[k] : k) * SP);
      const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        acc[f][0] = fmaf(wv[f], x.x, acc[f][0]);
        acc[f][1] = fmaf(wv[f], x.y, acc[f][1]);
        acc[f][2] = fmaf(wv[f], x.z, acc[f][2]);
        acc[f][3] = fmaf(wv[f], x.w, acc[f][3]);
      }
    }
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      const int n = 4 * tn + f;
      const float b = bias ? bias[n] : 0.f;
      float4 r = make_float4(acc[f][0] + b, acc[f][1] + b, acc[f][2] + b, acc[f][3] + b);
      if (lrelu) r = make_float4(lrelu_f(r.x), lrelu_f(r.y), lrelu_f(r.z), lrelu_f(r.w));
      st4(out + n * SP + 4 * ts, r);
    }
  }
}
// Few outputs, long reduction (N <= 32 rows, K ~ H): out[n][s] = sum_k Wel(n, k) * X[k][s], one output row quad per
// group of KQ adjacent lanes that split K and combine with shuffles (fixed order). Wel(n, k) = W[n * sn + k * sk].
__device__ __forceinline__ void gemm_small(float* out, const float* W, int sn, int sk, const float* X, int N, int K, int SP,
                                           int kq_log) {
  const int tiles_s = SP >> 2, KQ = 1 << kq_log;
  const int items = N * tiles_s;
  const int rounds = (items * KQ + kT - 1) / kT;
  for (int r = 0; r < rounds; ++r) {
    const int id = threadIdx.x + r * kT;
    const int item = id >> kq_log, kq = id & (KQ - 1);
    const bool ok = item < items;
    const int n = ok ? item / tiles_s : 0, ts = ok ? item - n * tiles_s : 0;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) {
      const float* wp = W + n * sn;
      const float* xp = X + 4 * ts;
#pragma unroll 4
      for (int k = kq; k < K; k += KQ) {
        const float w = wp[k * sk];
        const float4 x = ld4(xp + k * SP);
        acc.x = fmaf(w, x.x, acc.x); acc.y = fmaf(w, x.y, acc.y); acc.z = fmaf(w, x.z, acc.z); acc.w = fmaf(w, x.w, acc.w);
      }
    }
    for (int o = 1; o < KQ; o <<= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
      acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
    }
    if (ok && kq == 0) st4(out + n * SP + 4 * ts, acc);
  }
}

// ------------------------------------------------------------------------------- cross-CTA reductions
// Every CTA contributes part[0..H) and part[kMaxH..kMaxH + H) (shared memory, fp32); on return tot holds the combination
// over all CTAs in a fixed order: the 8 CTAs of a cluster in rank order (DSMEM, by the cluster's rank-0 CTA), then the
// clusters in order (one slot per cluster in L2, double-buffered by `parity`). welford: part = [mean | M2] of `cnt`
// samples per CTA, combined as (count, mean, M2) triples; otherwise two plain sums.
__device__ __forceinline__ void merge(float& n, float& mean, float& m2, float nb, float mb, float m2b) {
  if (nb > 0.f) {
    const float nn = n + nb, d = mb - mean;
    mean += d * (nb / nn);
    m2 += m2b + d * d * (n * nb / nn);
    n = nn;
  }
}
__device__ __forceinline__ void cross_reduce(const NArgs& A, float* part, float* tot, int H, bool welford, float cnt, int parity,
                                             unsigned& bar) {
  const KArgs& a = A.k;
  const int ncl = A.p.ncl, SW = 3 * kMaxH;                 // slot stride (floats)
  float* slots = a.W + A.p.oSlots + (size_t)parity * ncl * SW;
  const int cluster = blockIdx.x / kClusterN;
  float* cnts = part + 2 * kMaxH;                          // per-CTA sample count (Welford)
  if (threadIdx.x == 0) cnts[0] = cnt;
  cluster_sync_all();
  if (cluster_rank() == 0) {
    for (int h = threadIdx.x; h < H; h += kT) {
      float v0[kClusterN], v1[kClusterN], cn[kClusterN];
#pragma unroll
      for (int r = 0; r < kClusterN; ++r) {
        v0[r] = ld_dsmem(part + h, r);
        v1[r] = ld_dsmem(part + kMaxH + h, r);
        cn[r] = welford ? ld_dsmem(cnts, r) : 0.f;
      }
      float n = 0.f, s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int r = 0; r < kClusterN; ++r) {
        if (welford) merge(n, s0, s1, cn[r], v0[r], v1[r]);
        else { s0 += v0[r]; s1 += v1[r]; }
      }
      float* sl = slots + (size_t)cluster * SW;
      __stcg(sl + h, s0);
      __stcg(sl + kMaxH + h, s1);
      __stcg(sl + 2 * kMaxH + h, n);
    }
  }
  grid_barrier(a.ctrl, (++bar) * gridDim.x);
  for (int h = threadIdx.x; h < H; h += kT) {
    float v0[kMaxClusters], v1[kMaxClusters], cn[kMaxClusters];   // every slot load in flight before the ordered combine
#pragma unroll
    for (int c = 0; c < kMaxClusters; ++c) {
      const bool on = c < ncl;
      v0[c] = on ? __ldcg(slots + (size_t)c * SW + h) : 0.f;
      v1[c] = on ? __ldcg(slots + (size_t)c * SW + kMaxH + h) : 0.f;
      cn[c] = (on && welford) ? __ldcg(slots + (size_t)c * SW + 2 * kMaxH + h) : 0.f;
    }
    float n = 0.f, s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxClusters; ++c) {
      if (welford) merge(n, s0, s1, cn[c], v0[c], v1[c]);
      else { s0 += v0[c]; s1 += v1[c]; }
    }
    tot[h] = s0;
    tot[kMaxH + h] = s1;
  }
  __syncthreads();
}

#define NAR_STAMP(slot)                                                                     \
  do {                                                                                      \
    if (a.tbuf && blockIdx.x == 0 && threadIdx.x == 0 && e < 64) a.tbuf[(size_t)e * 16 + (slot)] = (unsigned long long)clock64(); \
  } while (0)

}  // namespace

// =================================================================================================== weight pre-pass
// wt[e] = [W1^T: Da4 x H (rows >= Da zero) | W2^T: H x H] for every stage: the forward GEMMs read weights k-major
__global__ void __launch_bounds__(256) nar_pack_kernel(const StageTab* __restrict__ stb, const float* __restrict__ P, int S, int Da,
                                                       int Da4, int H, float* __restrict__ wt) {
  const int per = (Da4 + H) * H;
  const long long total = (long long)S * per;
  for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(id / per), r = (int)(id - (long long)e * per);
    const int k = r / H, h = r - k * H;
    const StageTab& st = stb[e];
    float v;
    if (k < Da4) v = k < Da ? __ldg(P + st.p[P_W1] + (size_t)h * Da + k) : 0.f;
    else v = __ldg(P + st.p[P_W2] + (size_t)h * H + (k - Da4));
    wt[id] = v;
  }
}

// =================================================================================================== forward
__global__ void __launch_bounds__(kT, 1) nar_fwd_kernel(const __grid_constant__ NArgs A) {
  extern __shared__ __align__(16) float sm[];
  __shared__ __align__(8) unsigned long long bars[3];     // stage blocks 0 / 1, W2
  const KArgs& a = A.k;
  const Plan& p = A.p;
  const int t = threadIdx.x;
  const int D = a.D, Da = a.Da, Db = a.Db, Dout = a.Dout, H = a.H, S = a.S, B = a.B, LB = a.ldb, SP = p.SPB;
  const int s_base = blockIdx.x * SP;
  const int nv = max(0, min(SP, B - s_base));           // valid samples of this slab
  float* Xb[2] = {sm + p.oX0, sm + p.oX1};
  float* U1 = sm + p.oU1;
  float* U2 = sm + p.oU2;
  float* W2 = sm + p.oW2;
  float* sLd = sm + p.oLd;
  float* part = sm + p.oPart;
  float* tot = sm + p.oTot;
  StageTab* stb = reinterpret_cast<StageTab*>(sm + p.oSt);
  const bool train_bn = a.has_bn && a.train;
  const float* wt = a.W + p.oWT;
  const size_t wt_stage = (size_t)(p.Da4 + H) * H;
  unsigned bar = 0;

  // ---- one-off: stage table -> shared memory, barriers, stage-0 operands in flight, input slab z [B][D] -> X [D][SP]
  {
    const int4* src = reinterpret_cast<const int4*>(a.st);
    int4* dst = reinterpret_cast<int4*>(stb);
    for (int i = t; i < S * (int)(sizeof(StageTab) / 16); i += kT) dst[i] = __ldg(src + i);
  }
  if (t == 0) {
    for (int i = 0; i < 3; ++i) mbar_init(bars + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = t; i < SP; i += kT) sLd[i] = 0.f;
  __syncthreads();
  issue_block(a, p, stb, blk_at(sm + p.oBlk0, p, Dout, H), 0, false, bars + 0);
  issue_w2(W2, wt + (size_t)p.Da4 * H, H, bars + 2);
  cp_async_commit();
  for (int i = t; i < D * SP; i += kT) {
    const int s = i / D, f = i - s * D;
    Xb[0][f * SP + s] = s < nv ? __ldg(a.in + (size_t)(s_base + s) * D + f) : 0.f;
  }
  __syncthreads();
  for (RowQuad it(SP >> 2); it.r < D; it.next())          // z^T for the backward pass (stage 0 gathers from it)
    if (4 * it.q < nv) stcg4(a.W + a.oZT + (size_t)it.r * LB + s_base + 4 * it.q, ld4(Xb[0] + it.r * SP + 4 * it.q));

#pragma unroll 1
  for (int e = 0; e < S; ++e) {
    const StageTab& st = stb[e];
    const Blk sv = blk_at(sm + ((e & 1) ? p.oBlk1 : p.oBlk0), p, Dout, H);
    float* X = Xb[e & 1];
    float* Xn = Xb[(e + 1) & 1];
    NAR_STAMP(0);
    cp_async_wait<0>();
    mbar_wait(bars + (e & 1), (e >> 1) & 1);
    __syncthreads();                                      // this stage's operands landed; previous stage fully done
    NAR_STAMP(1);
    if (e + 1 < S) {                                      // next stage's block -> the other buffer
      issue_block(a, p, stb, blk_at(sm + ((e & 1) ? p.oBlk0 : p.oBlk1), p, Dout, H), e + 1, false, bars + ((e + 1) & 1));
      cp_async_commit();
    }
    const float qa = expf(sv.act[0]), qc = -sv.act[1];
    NAR_STAMP(2);
    // ---- net.0 + LeakyReLU
    gemm_kn(U1, sv.W1, H, X, sv.idx, H, Da, SP, sv.v1, true);
    __syncthreads();
    NAR_STAMP(3);
    for (int which = 1; which <= 2; ++which) {
      float* U = which == 1 ? U1 : U2;
      const float* vec = which == 1 ? sv.v1 : sv.v2;          // [bias | gamma | beta], stride Hp
      const float* rs = sv.stat + (which - 1) * 2 * p.Hp;     // [running mean | running var], stride Hp
      float* Lg = a.W + (which == 1 ? a.oL1T : a.oL2T) + (size_t)e * H * LB + s_base;
      float* Ng = a.W + (which == 1 ? a.oN1T : a.oN2T) + (size_t)e * H * LB + s_base;
      if (which == 2) {
        // ---- net.3 + LeakyReLU: U2 = lrelu(W2 n1 + b2); the W2 buffer is free afterwards -> next stage's W2^T
        NAR_STAMP(5);
        mbar_wait(bars + 2, e & 1);
        gemm_kn(U2, W2, H, U1, nullptr, H, H, SP, sv.v2, true);
        __syncthreads();
        NAR_STAMP(6);
        if (e + 1 < S) issue_w2(W2, wt + (size_t)(e + 1) * wt_stage + (size_t)p.Da4 * H, H, bars + 2);
        NAR_STAMP(7);
      }
      if (train_bn) {
        if (t < H) {                                       // this slab's (mean, M2) of feature t
          float s1 = 0.f;
          for (int s = 0; s < nv; ++s) s1 += U[t * SP + s];
          const float m = nv > 0 ? s1 / (float)nv : 0.f;
          float m2 = 0.f;
          for (int s = 0; s < nv; ++s) { const float d = U[t * SP + s] - m; m2 = fmaf(d, d, m2); }
          part[t] = m; part[kMaxH + t] = m2;
        }
        cross_reduce(A, part, tot, H, true, (float)nv, (2 * e + which - 1) & 1, bar);
      }
      // saved LeakyReLU outputs; BatchNorm in place (batch statistics: biased variance, eps 1e-2; eval: running
      // statistics; none); saved normalised activations
      float* stt = a.W + a.oST + ((size_t)e * 4 + (which - 1) * 2) * H;
      for (RowQuad it(SP >> 2); it.r < H; it.next()) {
        const int h = it.r, q = it.q;
        float4 v = ld4(U + h * SP + 4 * q);
        const bool live = 4 * q < nv;
        if (live) stcg4(Lg + (size_t)h * LB + 4 * q, v);
        if (a.has_bn) {
          float mean, rstd;
          if (train_bn) {
            mean = tot[h];
            const float var = tot[kMaxH + h] / (float)B;
            rstd = 1.0f / sqrtf(var + kBnEps);
            if (q == 0 && blockIdx.x == 0) {
              __stcg(stt + h, mean);
              __stcg(stt + H + h, rstd);
              a.R[st.r[(which - 1) * 2] + h] = (1.f - kBnMomentum) * rs[h] + kBnMomentum * mean;
              a.R[st.r[(which - 1) * 2 + 1] + h] = (1.f - kBnMomentum) * rs[p.Hp + h] + kBnMomentum * (var * (float)B / (float)(B - 1));
            }
          } else {
            mean = rs[h];
            rstd = 1.0f / sqrtf(rs[p.Hp + h] + kBnEps);
          }
          const float ga = vec[p.Hp + h] * rstd, be = vec[2 * p.Hp + h];
          v = make_float4(fmaf(v.x - mean, ga, be), fmaf(v.y - mean, ga, be), fmaf(v.z - mean, ga, be), fmaf(v.w - mean, ga, be));
          st4(U + h * SP + 4 * q, v);
        }
        if (live) stcg4(Ng + (size_t)h * LB + 4 * q, v);
      }
      __syncthreads();
    }
    NAR_STAMP(8);
    // ---- net.6.fc: raw ZeroFC outputs -> U1 rows [0, Dout) (U1 = n1 is dead), K = H split over 4 lanes
    gemm_small(U1, sv.W3, H, 1, U2, Dout, H, SP, 2);
    __syncthreads();
    NAR_STAMP(9);
    // ---- ZeroFC bias / exp(3 scale), tanh-exp coupling (:244-249), ActNorm.reverse (:66), logdet terms; one element per
    //      thread and step (consecutive lanes = consecutive samples)
    {
      float* Og = a.W + a.oOT + (size_t)e * Dout * LB + s_base;
      float* Yg = a.W + a.oYT + (size_t)e * D * LB + s_base;
      float* dl = U1 + Dout * SP;                          // [Db][SP] per-column logdet terms (scratch rows of U1)
      for (RowQuad it(SP); it.r < Db; it.next()) {
        const int j = it.r, s = it.q;
        const float ol = (U1[j * SP + s] + sv.b3[j]) * expf(3.f * sv.scale[j]);
        const float ot = (U1[(Db + j) * SP + s] + sv.b3[Db + j]) * expf(3.f * sv.scale[Db + j]);
        const float bv = X[sv.idx[Da + j] * SP + s];
        const float ls = tanhf(ol);
        const float yb = fmaf(bv / expf(ls) - ot, qa, qc);
        Xn[(Da + j) * SP + s] = yb;
        dl[j * SP + s] = -ls;
        if (s < nv) {
          __stcg(Og + (size_t)j * LB + s, ol);
          __stcg(Og + (size_t)(Db + j) * LB + s, ot);
          __stcg(Yg + (size_t)(Da + j) * LB + s, yb);
        }
      }
      for (RowQuad it(SP >> 2); it.r < Da; it.next()) {    // pass-through half (incl. the extra row of an odd D, :241)
        const int k = it.r, q = it.q;
        const float4 v = ld4(X + sv.idx[k] * SP + 4 * q);
        const float4 y = make_float4(fmaf(v.x, qa, qc), fmaf(v.y, qa, qc), fmaf(v.z, qa, qc), fmaf(v.w, qa, qc));
        st4(Xn + k * SP + 4 * q, y);
        if (4 * q < nv) stcg4(Yg + (size_t)k * LB + 4 * q, y);
      }
      __syncthreads();
      if (t < SP) {
        float sacc = 0.f;
        for (int j = 0; j < Db; ++j) sacc += dl[j * SP + t];
        sLd[t] += sacc;
      }
    }
    NAR_STAMP(10);
  }
  __syncthreads();
  // ---- outputs: x [B][D] sample-major, logdet = sum of the coupling terms + D * sum_e log_scale_inv_e (:60-64)
  {
    const float* X = Xb[S & 1];
    for (int i = t; i < D * SP; i += kT) {
      const int s = i / D, f = i - s * D;
      if (s < nv) a.out[(size_t)(s_base + s) * D + f] = X[f * SP + s];
    }
    if (t < nv) {
      float lsum = 0.f;
      for (int e = 0; e < S; ++e) lsum += __ldg(a.P + stb[e].lsi);
      float res = sLd[t] + (float)D * lsum;
      if (__ldcg(a.ctrl + 2)) res = __int_as_float(0x7fc00000);      // a barrier timed out: poison
      a.logdet[s_base + t] = res;
    }
  }
  cluster_sync_all();      // no CTA leaves while its cluster leader may still read its shared memory
}

// =================================================================================================== backward chain
// Gradient state GX [D][SP] (rows in the order of the current stage's output) walks the stages backwards; per stage the
// rows g_pre / p2 / p1 (gradients at the outputs of net.6.fc / net.3 / net.0, after the LeakyReLU and BatchNorm
// backward) are stored for nar_pgrad_kernel, and the BatchNorm gamma / beta gradients (the two batch sums the BatchNorm
// backward needs anyway) are written directly.
__global__ void __launch_bounds__(kT, 1) nar_bwd_kernel(const __grid_constant__ NArgs A) {
  extern __shared__ __align__(16) float sm[];
  __shared__ __align__(8) unsigned long long bars[3];     // stage blocks 0 / 1, W2
  __shared__ float sRed[2 * (kT / 32)];
  const KArgs& a = A.k;
  const Plan& p = A.p;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int D = a.D, Da = a.Da, Db = a.Db, Dout = a.Dout, H = a.H, S = a.S, B = a.B, LB = a.ldb, SP = p.SPB;
  const int s_base = blockIdx.x * SP;
  const int nv = max(0, min(SP, B - s_base));
  const int tq = SP >> 2;
  float* Gb[2] = {sm + p.oX0, sm + p.oX1};
  float* U1 = sm + p.oU1;
  float* U2 = sm + p.oU2;
  float* W2 = sm + p.oW2;
  float* sGld = sm + p.oLd;
  float* part = sm + p.oPart;
  float* tot = sm + p.oTot;
  float* GP = sm + p.oGP;
  float* Act[2] = {sm + p.oAct0, sm + p.oAct1};
  StageTab* stb = reinterpret_cast<StageTab*>(sm + p.oSt);
  const float invB = 1.0f / (float)B;
  unsigned bar = 0;

  {
    const int4* src = reinterpret_cast<const int4*>(a.st);
    int4* dst = reinterpret_cast<int4*>(stb);
    for (int i = t; i < S * (int)(sizeof(StageTab) / 16); i += kT) dst[i] = __ldg(src + i);
  }
  if (t == 0) {
    for (int i = 0; i < 3; ++i) mbar_init(bars + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = t; i < SP; i += kT) sGld[i] = i < nv ? __ldg(a.g_ld + s_base + i) : 0.f;
  __syncthreads();
  // saved activations of stage e for this slab -> shared memory rows [y (D) | o (Dout) | b_prev (Db) | l1 (H) | l2 (H)]
  auto issue_act = [&](float* dst, int e) {
    const float* yprev = e == 0 ? a.W + a.oZT : a.W + a.oYT + (size_t)(e - 1) * D * LB;
    for (RowQuad it(tq); it.r < p.actRows; it.next()) {
      const int r = it.r;
      const float* src;
      if (r < D) src = a.W + a.oYT + ((size_t)e * D + r) * LB;
      else if (r < D + Dout) src = a.W + a.oOT + ((size_t)e * Dout + (r - D)) * LB;
      else if (r < D + Dout + Db) src = yprev + (size_t)__ldg(a.gmaps + (size_t)e * D + Da + (r - D - Dout)) * LB;
      else if (r < D + Dout + Db + H) src = a.W + a.oL1T + ((size_t)e * H + (r - D - Dout - Db)) * LB;
      else src = a.W + a.oL2T + ((size_t)e * H + (r - D - Dout - Db - H)) * LB;
      cp_async16(dst + r * SP + 4 * it.q, src + s_base + 4 * it.q, 4 * it.q < nv ? 16 : 0);
    }
  };
  issue_block(a, p, stb, blk_at(sm + p.oBlk0, p, Dout, H), S - 1, true, bars + 0);
  issue_w2(W2, a.P + stb[S - 1].p[P_W2], H, bars + 2);
  issue_act(Act[0], S - 1);
  cp_async_commit();
  // gradient wrt the flow output, g_out [B][D] sample-major -> GX [D][SP]
  for (int i = t; i < D * SP; i += kT) {
    const int s = i / D, f = i - s * D;
    Gb[0][f * SP + s] = s < nv ? __ldg(a.g_out + (size_t)(s_base + s) * D + f) : 0.f;
  }
  __syncthreads();

#pragma unroll 1
  for (int u = 0; u < S; ++u) {
    const int e = S - 1 - u;
    const StageTab& st = stb[e];
    const Blk sv = blk_at(sm + ((u & 1) ? p.oBlk1 : p.oBlk0), p, Dout, H);
    float* GX = Gb[u & 1];
    float* GXn = Gb[(u + 1) & 1];
    const float* act = Act[u & 1];
    const float* Ys = act;
    const float* Os = act + D * SP;
    const float* BPs = Os + Dout * SP;
    const float* L1s = BPs + Db * SP;
    const float* L2s = L1s + H * SP;
    cp_async_wait<0>();
    mbar_wait(bars + (u & 1), (u >> 1) & 1);
    __syncthreads();
    if (u + 1 < S) {
      issue_block(a, p, stb, blk_at(sm + ((u & 1) ? p.oBlk0 : p.oBlk1), p, Dout, H), e - 1, true, bars + ((u + 1) & 1));
      issue_act(Act[(u + 1) & 1], e - 1);
      cp_async_commit();
    }
    const float e_lsi = expf(sv.act[0]), loc = sv.act[1];
    // ---- ActNorm.reverse + coupling tail, elementwise: g_pre, the b-half input gradient, ActNorm sums
    float S1 = 0.f, S2 = 0.f;
    for (RowQuad it(SP); it.r < Db; it.next()) {
      const int j = it.r, s = it.q;
      const float gb = GX[(Da + j) * SP + s];
      S1 += gb;
      S2 = fmaf(gb, Ys[(Da + j) * SP + s] + loc, S2);
      const float gbp = gb * e_lsi;
      const float ols = Os[j * SP + s];
      const float ls = tanhf(ols), ei = expf(-ls);
      const float g_t = -gbp;
      const float g_ls = -gbp * BPs[j * SP + s] * ei - sGld[s];
      const float g_ols = g_ls * (1.f - ls * ls);
      const float v_ls = g_ols * expf(3.f * sv.scale[j]), v_t = g_t * expf(3.f * sv.scale[Db + j]);
      GP[j * SP + s] = v_ls;
      GP[(Db + j) * SP + s] = v_t;
      GXn[sv.idx[Da + j] * SP + s] = gbp * ei;
      if (s < nv) {
        __stcg(a.W + p.oGPRE + ((size_t)e * Dout + j) * LB + s_base + s, v_ls);
        __stcg(a.W + p.oGPRE + ((size_t)e * Dout + Db + j) * LB + s_base + s, v_t);
      }
    }
    for (RowQuad it(SP); it.r < Da; it.next()) {
      const float ga = GX[it.r * SP + it.q];
      S1 += ga;
      S2 = fmaf(ga, Ys[it.r * SP + it.q] + loc, S2);
    }
    S1 = warp_sum(S1); S2 = warp_sum(S2);
    if (lane == 0) { sRed[warp] = S1; sRed[kT / 32 + warp] = S2; }
    __syncthreads();
    if (t == 0) {
      float t1 = 0.f, t2 = 0.f;
      for (int w = 0; w < kT / 32; ++w) { t1 += sRed[w]; t2 += sRed[kT / 32 + w]; }
      float* ap = a.W + p.oActPart + ((size_t)e * gridDim.x + blockIdx.x) * 2;
      __stcg(ap, t1);
      __stcg(ap + 1, t2);
    }
    // ---- through net.6.fc: g_n2[h][s] = sum_j W3[j][h] g_pre[j][s]
    gemm_kn(U2, sv.W3, H, GP, nullptr, H, Dout, SP, nullptr, false);
    __syncthreads();
    for (int which = 2; which >= 1; --which) {
      float* U = which == 2 ? U2 : U1;
      const float* Ls = which == 2 ? L2s : L1s;
      const float* vec = which == 2 ? sv.v2 : sv.v1;               // [bias | gamma | beta], stride Hp
      const float* mean = sv.stat + (which - 1) * 2 * H;            // saved batch statistics [mean | rstd], stride H
      const float* rstd = mean + H;
      if (which == 1) {
        // ---- through net.3: g_n1[k][s] = sum_n W2[n][k] p2[n][s]; W2 is free afterwards -> next stage's
        mbar_wait(bars + 2, u & 1);
        gemm_kn(U1, W2, H, U2, nullptr, H, H, SP, nullptr, false);
        __syncthreads();
        if (u + 1 < S) issue_w2(W2, a.P + stb[e - 1].p[P_W2], H, bars + 2);
      }
      if (a.has_bn) {
        if (t < H) {                                       // this slab's sums of g and g * xhat for feature t
          float s0 = 0.f, s1 = 0.f;
          const float m = mean[t], r = rstd[t];
          for (int s = 0; s < nv; ++s) {
            const float g = U[t * SP + s];
            s0 += g;
            s1 = fmaf(g, (Ls[t * SP + s] - m) * r, s1);
          }
          part[t] = s0; part[kMaxH + t] = s1;
        }
        cross_reduce(A, part, tot, H, false, 0.f, (2 * u + (2 - which)) & 1, bar);
        if (blockIdx.x == 0 && t < H) {                   // BatchNorm weight / bias gradients are exactly these sums
          a.G[st.p[which == 2 ? P_G2 : P_G1] + t] = tot[kMaxH + t];
          a.G[st.p[which == 2 ? P_BE2 : P_BE1] + t] = tot[t];
        }
      }
      // BatchNorm + LeakyReLU backward in place; rows kept for the weight gradients
      float* Pg = a.W + (which == 2 ? p.oGP2 : p.oGP1) + (size_t)e * H * LB + s_base;
      for (RowQuad it(tq); it.r < H; it.next()) {
        const int h = it.r, q = it.q;
        const float4 g = ld4(U + h * SP + 4 * q), l = ld4(Ls + h * SP + 4 * q);
        const float gv[4] = {g.x, g.y, g.z, g.w}, lv[4] = {l.x, l.y, l.z, l.w};
        float o[4];
        if (a.has_bn) {
          const float m = mean[h], r = rstd[h], gr = vec[p.Hp + h] * r;
          const float gbeta = tot[h] * invB, ggamma = tot[kMaxH + h] * invB;
#pragma unroll
          for (int i = 0; i < 4; ++i) o[i] = gr * (gv[i] - gbeta - (lv[i] - m) * r * ggamma);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) o[i] = gv[i];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = (4 * q + i < nv) ? o[i] * (lv[i] > 0.f ? 1.f : kLreluSlope) : 0.f;
        const float4 r4 = make_float4(o[0], o[1], o[2], o[3]);
        st4(U + h * SP + 4 * q, r4);
        if (4 * q < nv) stcg4(Pg + (size_t)h * LB + 4 * q, r4);
      }
      __syncthreads();
    }
    // ---- through net.0 (+ the direct path of the pass-through half through ActNorm): input gradient of the a rows
    gemm_small(U2, sv.W1, 1, Da, U1, Da, H, SP, 2);       // W1 stored [h][k]: element (k, h) at k + h * Da
    __syncthreads();
    for (RowQuad it(tq); it.r < Da; it.next()) {
      const int k = it.r, q = it.q;
      const float4 n = ld4(U2 + k * SP + 4 * q), g = ld4(GX + k * SP + 4 * q);
      st4(GXn + sv.idx[k] * SP + 4 * q, make_float4(fmaf(g.x, e_lsi, n.x), fmaf(g.y, e_lsi, n.y), fmaf(g.z, e_lsi, n.z), fmaf(g.w, e_lsi, n.w)));
    }
  }
  __syncthreads();
  if (a.g_in) {
    const float* GX = Gb[S & 1];
    for (int i = t; i < D * SP; i += kT) {
      const int s = i / D, f = i - s * D;
      if (s < nv) a.g_in[(size_t)(s_base + s) * D + f] = GX[f * SP + s];
    }
  }
  cluster_sync_all();
}

// =================================================================================================== parameter gradients
// One batched pass over the rows the chain stored: for every stage e
//   gW3[j][h] = sum_s g_pre[j][s] n2[h][s]    gW2[n][k] = sum_s p2[n][s] n1[k][s]    gW1[h][k] = sum_s p1[h][s] a[k][s]
//   gb3 = sum_s g_pre   gb2 = sum_s p2   gb1 = sum_s p1   gscale[j] = 3 sum_s g_pre[j][s] o[j][s] / exp(3 scale_j)
//   g_loc = -sum S1     g_lsi = sum S2 + D sum_s g_ld[s]
// Job = (stage, tile of one weight gradient | a block of row sums); no cross-CTA dependencies.

// Tile shapes follow the matrix: rows / columns per tile from {16, 32, 48, 64} with the least padding (hidden 144 -> 48,
// 18 output rows -> 32, 9 input columns -> 16), 4 x 4 outputs per thread.
struct PTiles {
  int ti2, tj2, ti3, tj3, ti1, tj1;     // tile rows / columns of gW2 [H][H], gW3 [Dout][H], gW1 [H][Da]
  int n2, n3, n1;                       // tiles per stage of each
};
__host__ __device__ inline int pg_tile_dim(int n) {
  int best = 64, waste = ((n + 63) / 64) * 64 - n;
  for (int t = 48; t >= 16; t -= 16) {
    const int w = ((n + t - 1) / t) * t - n;
    if (w < waste) { waste = w; best = t; }
  }
  return best;
}

__global__ void __launch_bounds__(256) nar_pgrad_kernel(const __grid_constant__ NArgs A, int jobs_per_stage, const PTiles pt) {
  __shared__ __align__(16) float Gs[32][68];
  __shared__ __align__(16) float As[32][68];
  const KArgs& a = A.k;
  const Plan& p = A.p;
  const int D = a.D, Da = a.Da, Dout = a.Dout, H = a.H, B = a.B, LB = a.ldb;
  const int e = blockIdx.x / jobs_per_stage;
  int job = blockIdx.x - e * jobs_per_stage;
  const StageTab& st = a.st[e];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const float* yprev = e == 0 ? a.W + a.oZT : a.W + a.oYT + (size_t)(e - 1) * D * LB;
  // ---- decode the job
  const float* Grows; const float* Arows; int I, J, TI, TJ; float* out; const int* amap = nullptr;
  if (job < pt.n2) { Grows = a.W + p.oGP2 + (size_t)e * H * LB; Arows = a.W + a.oN1T + (size_t)e * H * LB; I = H; J = H; TI = pt.ti2; TJ = pt.tj2; out = a.G + st.p[P_W2]; }
  else if ((job -= pt.n2) < pt.n3) { Grows = a.W + p.oGPRE + (size_t)e * Dout * LB; Arows = a.W + a.oN2T + (size_t)e * H * LB; I = Dout; J = H; TI = pt.ti3; TJ = pt.tj3; out = a.G + st.p[P_W3]; }
  else if ((job -= pt.n3) < pt.n1) { Grows = a.W + p.oGP1 + (size_t)e * H * LB; Arows = yprev; amap = a.gmaps + (size_t)e * D; I = H; J = Da; TI = pt.ti1; TJ = pt.tj1; out = a.G + st.p[P_W1]; }
  else {
    // ---- row sums: biases, ZeroFC scale, ActNorm; one warp per row, lanes stride the batch
    job -= pt.n1;                                          // 0..3: which quarter of the rows
    const int rows = 2 * H + 2 * Dout;                     // [b1: H | b2: H | b3: Dout | scale: Dout]
    for (int r = job * 8 + warp; r < rows; r += 32) {
      const float* g; const float* o = nullptr; float* dst; float mul = 1.f;
      if (r < H) { g = a.W + p.oGP1 + ((size_t)e * H + r) * LB; dst = a.G + st.p[P_B1] + r; }
      else if (r < 2 * H) { g = a.W + p.oGP2 + ((size_t)e * H + r - H) * LB; dst = a.G + st.p[P_B2] + (r - H); }
      else if (r < 2 * H + Dout) { g = a.W + p.oGPRE + ((size_t)e * Dout + r - 2 * H) * LB; dst = a.G + st.p[P_B3] + (r - 2 * H); }
      else {
        const int j = r - 2 * H - Dout;
        g = a.W + p.oGPRE + ((size_t)e * Dout + j) * LB; o = a.W + a.oOT + ((size_t)e * Dout + j) * LB;
        dst = a.G + st.p[P_SCALE] + j; mul = 3.f / expf(3.f * __ldg(a.P + st.p[P_SCALE] + j));
      }
      float acc = 0.f;
      for (int s = lane; s < B; s += 32) acc = o ? fmaf(__ldcg(g + s), __ldcg(o + s), acc) : acc + __ldcg(g + s);
      acc = warp_sum(acc);
      if (lane == 0) *dst = acc * mul;
    }
    if (job == 0 && warp == 0) {
      float s1 = 0.f, s2 = 0.f, gl = 0.f;
      const float* ap = a.W + p.oActPart + (size_t)e * p.G * 2;
      for (int c = lane; c < p.G; c += 32) { s1 += __ldcg(ap + 2 * c); s2 += __ldcg(ap + 2 * c + 1); }
      for (int s = lane; s < B; s += 32) gl += __ldg(a.g_ld + s);
      s1 = warp_sum(s1); s2 = warp_sum(s2); gl = warp_sum(gl);
      if (lane == 0) {
        a.G[st.loc] = -s1 + (__ldcg(a.ctrl + 2) ? __int_as_float(0x7fc00000) : 0.f);     // timed-out barrier -> NaN
        a.G[st.lsi] = s2 + (float)D * gl;
      }
    }
    return;
  }
  // ---- TI x TJ tile of C[i][j] = sum_s G[i][s] Arow(j)[s]
  const int tjn = (J + TJ - 1) / TJ;
  const int ti = job / tjn, tj = job - ti * tjn;
  const int i0 = ti * TI, j0 = tj * TJ;
  const int tcn = TJ >> 2;                                 // thread tile columns
  const bool on = t < (TI >> 2) * tcn;
  const int tr = on ? t / tcn : 0, tc = on ? t - tr * tcn : 0;   // thread tile: rows 4 tr.., columns 4 tc..
  float acc[4][4];
#pragma unroll
  for (int x = 0; x < 4; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) acc[x][y] = 0.f;
  for (int s0 = 0; s0 < B; s0 += 32) {
    __syncthreads();
    for (int id = t; id < (TI + TJ) * 8; id += 256) {      // (TI + TJ) rows x 8 sample quads
      const int rr = id >> 3, q = id & 7;
      const bool isg = rr < TI;
      const int r = isg ? rr : rr - TI;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (s0 + 4 * q < B) {
        if (isg) {
          if (i0 + r < I) v = __ldcg(reinterpret_cast<const float4*>(Grows + (size_t)(i0 + r) * LB + s0 + 4 * q));
        } else if (j0 + r < J) {
          const int row = amap ? __ldg(amap + j0 + r) : j0 + r;
          v = __ldcg(reinterpret_cast<const float4*>(Arows + (size_t)row * LB + s0 + 4 * q));
        }
      }
      float (*dst)[68] = isg ? Gs : As;
      dst[4 * q + 0][r] = v.x; dst[4 * q + 1][r] = v.y; dst[4 * q + 2][r] = v.z; dst[4 * q + 3][r] = v.w;
    }
    __syncthreads();
    if (on) {
#pragma unroll 8
      for (int s = 0; s < 32; ++s) {
        const float4 g = ld4(&Gs[s][4 * tr]), v = ld4(&As[s][4 * tc]);
        const float gv[4] = {g.x, g.y, g.z, g.w}, vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
          for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(gv[x], vv[y], acc[x][y]);
      }
    }
  }
  if (on) {
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y) {
        const int i = i0 + 4 * tr + x, j = j0 + 4 * tc + y;
        if (i < I && j < J) out[(size_t)i * J + j] = acc[x][y];
      }
  }
}

// =================================================================================================== host side
namespace {
int g_dev = -1, g_max_clusters = 0;
constexpr int kSmemMax = 220 * 1024;

cudaLaunchConfig_t make_cfg(cudaLaunchAttribute* at, int G, int smem, cudaStream_t st, bool coop) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(G); cfg.blockDim = dim3(kT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kClusterN; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  int na = 1;
  if (coop) { at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1; na = 2; }
  cfg.attrs = at; cfg.numAttrs = na;
  return cfg;
}
inline int pad32i(int n) { return (n + 31) & ~31; }
}  // namespace

// DPI_B200_NAR_COOP=0: plain (non-cooperative) launches, for Nsight Compute, which cannot profile cooperative cluster
// launches; the grid is sized to be co-resident either way.
static bool coop_allowed() {
  static const bool off = []() { const char* e = getenv("DPI_B200_NAR_COOP"); return e && e[0] == '0'; }();
  return !off;
}

int has_backward() {
  static const bool off = []() { const char* e = getenv("DPI_B200_NAR_BWD"); return e && e[0] == '0'; }();
  return off ? 0 : 1;
}

int query(int device) {
  if (g_dev == device) return g_max_clusters;
  g_dev = device;
  g_max_clusters = 0;
  if (cudaFuncSetAttribute((const void*)nar_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax) != cudaSuccess ||
      cudaFuncSetAttribute((const void*)nar_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchAttribute at[2];
  cudaLaunchConfig_t cfg = make_cfg(at, kClusterN * 16, kSmemMax, nullptr, false);
  int nc = 0;
  if (cudaOccupancyMaxActiveClusters(&nc, (const void*)nar_fwd_kernel, &cfg) != cudaSuccess || nc < 1) {
    cudaGetLastError();
    return 0;
  }
  g_max_clusters = std::min(nc, kMaxClusters);
  return g_max_clusters;
}

namespace {
// block and shared-memory layout for slabs of SPB samples on G CTAs
void fill_layout(Plan& p, int G, int SPB, int D, int H, int S) {
  const int Da = D - D / 2, Db = D / 2, Dout = 2 * Db;
  p.G = G; p.ncl = G / kClusterN; p.SPB = SPB;
  p.Da4 = (Da + 3) & ~3;
  p.Hp = pad32i(H);
  int b = 0;
  auto btake = [&](int n) { int r = b; b += (n + 3) & ~3; return r; };
  p.bW1 = btake(std::max(p.Da4 * H, pad32i(H * Da)));
  p.bV1 = btake(3 * p.Hp); p.bV2 = btake(3 * p.Hp);
  p.bSC = btake(2 * pad32i(Dout) + pad32i(Dout * H));
  p.bST = btake(4 * p.Hp);
  p.bIdx = btake(D); p.bAct = btake(4);
  p.blockFloats = b;
  int o = 0;
  auto take = [&](int n) { int r = o; o += (n + 3) & ~3; return r; };
  p.oSt = take(S * (int)(sizeof(StageTab) / 4));
  p.oBlk0 = take(p.blockFloats); p.oBlk1 = take(p.blockFloats);
  p.oW2 = take(H * H);
  p.oX0 = take(D * SPB); p.oX1 = take(D * SPB);
  const int urows = std::max(H, Dout + Db + 4);
  p.oU1 = take(urows * SPB); p.oU2 = take(urows * SPB);
  p.oLd = take(SPB);
  p.oPart = take(2 * kMaxH + 4);
  p.oTot = take(2 * kMaxH);
  p.smem_fwd = o * 4;
  // backward: g_pre rows and the double-buffered slab of saved activations [y | o | b_prev | l1 | l2]
  p.oGP = take((Dout + 4) * SPB);
  p.actRows = D + Dout + Db + 2 * H;
  p.oAct0 = take(p.actRows * SPB); p.oAct1 = take(p.actRows * SPB);
  p.smem_bwd = o * 4;
}
bool shape_ok(int D, int H, int B, int S) { return !(D > kMaxD || D < 2 || H > kMaxH || (H & 3) || B < 8 || (B & 3) || S > 128); }
}  // namespace

void make_plan(Plan& p, int max_clusters, int D, int H, int B, int S, int has_bn) {
  p = Plan{};
  (void)has_bn;
  if (max_clusters < 1 || !shape_ok(D, H, B, S)) return;
  const int cap = kClusterN * max_clusters;
  int SPB = ((B + cap - 1) / cap + 3) & ~3;
  SPB = std::max(SPB, 8);
  if (SPB > kMaxSPB) return;
  const int G = ((B + SPB - 1) / SPB + kClusterN - 1) / kClusterN * kClusterN;
  fill_layout(p, G, SPB, D, H, S);
  p.ok = (p.smem_bwd <= kSmemMax) ? 1 : 0;
}

// Forward-only plan for passes that need no cross-CTA exchange (eval mode: running statistics): the slabs are independent,
// so the grid may exceed the co-resident CTAs - any batch, 32 samples per CTA (posterior sampling at 8192,
// DPIx_interferometry.py:433). p.ok then means "the forward fits"; there is no backward for it.
void make_plan_eval(Plan& p, int avail, int D, int H, int B, int S) {
  p = Plan{};
  if (avail < 1 || !shape_ok(D, H, B, S)) return;
  const int SPB = 32;
  const int G = ((B + SPB - 1) / SPB + kClusterN - 1) / kClusterN * kClusterN;
  fill_layout(p, G, SPB, D, H, S);
  p.ok = (p.smem_fwd <= kSmemMax) ? 1 : 0;
}

long long workspace_floats_eval(const Plan& p, int D, int H, int S) {
  (void)D;
  return (long long)S * (p.Da4 + H) * H + 256;
}
void place_workspace_eval(Plan& p, long long base) { p.oWT = base; }

long long workspace_floats(const Plan& p, int D, int H, int B, int S) {
  const long long Dout = 2 * (D / 2);
  return (long long)S * (Dout + 2 * H) * B + 2LL * p.ncl * (3 * kMaxH) + 2LL * S * p.G + (long long)S * (p.Da4 + H) * H + 256;
}
void place_workspace(Plan& p, long long base, int D, int H, int B, int S) {
  const long long Dout = 2 * (D / 2);
  p.oGPRE = base; base += (long long)S * Dout * B;
  p.oGP2 = base; base += (long long)S * H * B;
  p.oGP1 = base; base += (long long)S * H * B;
  base = (base + 31) / 32 * 32;
  p.oSlots = base; base += 2LL * p.ncl * (3 * kMaxH);
  p.oActPart = base; base += 2LL * S * p.G;
  base = (base + 31) / 32 * 32;
  p.oWT = base;
}

int forward(const KArgs& a, const Plan& p, cudaStream_t st) {
  NArgs A{};
  A.k = a;
  A.p = p;
  DPI_CUDA(cudaMemsetAsync(a.ctrl, 0, kCtrlWords * sizeof(unsigned), st));
  {
    const long long total = (long long)a.S * (p.Da4 + a.H) * a.H;
    nar_pack_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, 148 * 8), 256, 0, st>>>(a.st, a.P, a.S, a.Da, p.Da4, a.H,
                                                                                                 a.W + p.oWT);
    DPI_LAUNCH_CHECK();
  }
  cudaLaunchAttribute at[2];
  // batch statistics need every CTA co-resident (grid barrier): cooperative launch; eval / no-BatchNorm passes do not
  const bool coop = a.has_bn && a.train && coop_allowed();
  cudaLaunchConfig_t cfg = make_cfg(at, p.G, p.smem_fwd, st, coop);
  DPI_CUDA(cudaLaunchKernelEx(&cfg, nar_fwd_kernel, A));
  count_launch();
  return DPI_OK;
}

int backward(const KArgs& a, const Plan& p, cudaStream_t st) {
  NArgs A{};
  A.k = a;
  A.p = p;
  DPI_CUDA(cudaMemsetAsync(a.ctrl, 0, kCtrlWords * sizeof(unsigned), st));
  cudaLaunchAttribute at[2];
  cudaLaunchConfig_t cfg = make_cfg(at, p.G, p.smem_bwd, st, a.has_bn != 0 && coop_allowed());
  DPI_CUDA(cudaLaunchKernelEx(&cfg, nar_bwd_kernel, A));
  count_launch();
  return param_grads(a, p, st);
}

// Batched weight / bias / scale / ActNorm gradients from the rows a backward chain stored (p.oGPRE, oGP2, oGP1, oActPart
// with p.G partial pairs per stage); also used by the sample-resident small-batch chain of flow_sr.cu.
int param_grads(const KArgs& a, const Plan& p, cudaStream_t st) {
  NArgs A{};
  A.k = a;
  A.p = p;
  PTiles pt{};
  pt.ti2 = pt.tj2 = pg_tile_dim(a.H);
  pt.ti3 = pg_tile_dim(a.Dout); pt.tj3 = pg_tile_dim(a.H);
  pt.ti1 = pg_tile_dim(a.H); pt.tj1 = pg_tile_dim(a.Da);
  auto cnt = [](int n, int t) { return (n + t - 1) / t; };
  pt.n2 = cnt(a.H, pt.ti2) * cnt(a.H, pt.tj2);
  pt.n3 = cnt(a.Dout, pt.ti3) * cnt(a.H, pt.tj3);
  pt.n1 = cnt(a.H, pt.ti1) * cnt(a.Da, pt.tj1);
  const int jobs = pt.n2 + pt.n3 + pt.n1 + 4;
  nar_pgrad_kernel<<<a.S * jobs, 256, 0, st>>>(A, jobs, pt);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

}  // namespace nar
}  // namespace dpi
