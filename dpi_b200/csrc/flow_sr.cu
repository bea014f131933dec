// flow_sr.cu - RealNVP.reverse and its backward for batch <= 64 of a mid-sized flow as SAMPLE-resident kernels.
//
// Math: generative_model/realnvpfc_model.py AffineCoupling.reverse :240-249, net :171-187, ZeroFC :137-141,
// ActNorm.reverse :49-66, Flow.reverse :457-474, RealNVP.reverse :594-603; backward per SURVEY.md Appendix A.
//
// Why: at the reference's default configuration (RealNVP(1024, 16), hidden 128, 64 samples) a coupling stage is three
// dependent GEMMs of a few MFLOP. Splitting their OUTPUT FEATURES over the GPU (flow_small.cu, 120 CTAs) makes every
// layer a grid-wide hand-off: 228 hops of ~5 us per step (grid barrier + L2 round trip), 1.65 % of the HBM roofline.
// Splitting them over a cluster with tensor cores (flow_cl.cu) still needs four all-to-all exchanges per stage and
// measured no faster (profiles/cl_fwd_r02.md). This file splits the SAMPLES instead:
//   * one cluster of 16 CTAs, 4 samples per CTA; a CTA carries its samples through all stages in shared memory, so the
//     permutations and the three layers of a stage need no exchange at all;
//   * every CTA therefore reads every weight once per pass - straight from L2 into registers (ld.global.nc.v4, each
//     weight is used by exactly one thread: staging it in shared memory would only add traffic). Measured on B200
//     (tools/ldg_bench.cu): 118 B/clk per SM with 16 or 32 SMs streaming, i.e. the 27.6 MB of C1's parameters in 130 us
//     per pass, against 501 us for the hop-bound kernel; weights are laid out k-major (forward: a transposing pre-pass;
//     backward: the stored [out][in] layout already is) so a warp reads 512 contiguous bytes per k and a thread tile
//     (4 output features x 4 samples) is one LDG.128 + one broadcast LDS.128 per 16 FMA; K is split over the warps and
//     combined through shared memory in a fixed order (bit-reproducible);
//   * the ONLY cross-CTA traffic is the BatchNorm batch statistics: each CTA publishes a (mean, M2) pair (forward) or two
//     sums (backward) per feature in its own shared memory, one barrier.cluster, and every CTA combines the 16 partials in
//     rank order through distributed shared memory - 0.7 us, twice per stage;
//   * weight / bias / scale / ActNorm gradients are not formed in the chain: it stores the per-stage gradient rows and the
//     batched kernel of flow_nar.cu (nar::param_grads, all SMs, no dependencies) reduces them over the batch.
// Same feature-major workspace as the other paths (flow.cuh): the saved activations are interchangeable with them.
#include <algorithm>
#include <cstdlib>

#include "flow_host.cuh"
#include "flow_sr.cuh"

namespace dpi {
namespace sr {

struct SArgs {
  KArgs k;
  Plan p;
};

namespace {

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
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void stcg4(float* p, float4 v) { __stcg(reinterpret_cast<float4*>(p), v); }
// streamed weights: read once per CTA and pass, never reused by this SM
__device__ __forceinline__ float4 ldw4(const float* p) {
  float4 v;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// ------------------------------------------------------------------------------- streamed micro-GEMM
// out[n][s] = sum_k W[k * pw + n] * X[k][s] for n < N (multiple of 4), s < 4; W in global memory (k-major, rows
// n-contiguous), X [K][4] in shared memory. N / 4 thread tiles of 4 features x 4 samples, K split over kT / (N / 4)
// thread groups; every group leaves its partial tile in red[group][n][s] (at most 16 * kT floats).
struct Split {
  int JT, KG, kper;
};
__device__ __forceinline__ Split split_of(int N, int K) {
  Split s;
  s.JT = N >> 2;
  s.KG = max(1, kT / s.JT);
  s.kper = (K + s.KG - 1) / s.KG;
  return s;
}
__device__ __forceinline__ void fma16(float (&acc)[4][4], const float4 w, const float4 x) {
  const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    acc[f][0] = fmaf(wv[f], x.x, acc[f][0]);
    acc[f][1] = fmaf(wv[f], x.y, acc[f][1]);
    acc[f][2] = fmaf(wv[f], x.z, acc[f][2]);
    acc[f][3] = fmaf(wv[f], x.w, acc[f][3]);
  }
}
// PW: compile-time row pitch (the loads of a batch become one base register + immediates), 0 = runtime pitch
template <int PW>
__device__ __noinline__ void gemm_stream_t(const float* __restrict__ W, int pw_rt, int N, int K, const float* X, float* red) {
  const int pw = PW ? PW : pw_rt;
  const Split sp = split_of(N, K);
  const int t = threadIdx.x;
  if (t >= sp.JT * sp.KG) return;
  const int kg = t / sp.JT, jt = t - kg * sp.JT;
  const int k0 = min(K, kg * sp.kper), k1 = min(K, k0 + sp.kper);
  float acc[4][4];
#pragma unroll
  for (int f = 0; f < 4; ++f)
#pragma unroll
    for (int s = 0; s < 4; ++s) acc[f][s] = 0.f;
  const float* wp = W + (size_t)k0 * pw + 4 * jt;
  const float* xp = X + 4 * k0;
  int n = k1 - k0;
  // batches of 8 independent 16-byte loads per thread (64 KB in flight per SM), then 8 x 16 FMA. Measured in isolation
  // (tools/sr_gemm_bench.cu): deeper prefetch (register double buffers, L1 prefetch) does not help - LDG.128 and FFMA
  // issue add up (~129 cycles per 8 KB of weights per SM), the loop is not latency-bound.
  while (n >= 8) {
    float4 w[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) w[u] = ldw4(wp + (size_t)u * pw);
#pragma unroll
    for (int u = 0; u < 8; ++u) fma16(acc, w[u], ld4(xp + 4 * u));
    wp += (size_t)8 * pw; xp += 32; n -= 8;
  }
  for (; n > 0; --n) {
    fma16(acc, ldw4(wp), ld4(xp));
    wp += pw; xp += 4;
  }
  float* rp = red + ((size_t)kg * N + 4 * jt) * 4;
#pragma unroll
  for (int f = 0; f < 4; ++f) st4(rp + 4 * f, make_float4(acc[f][0], acc[f][1], acc[f][2], acc[f][3]));
}
__device__ __forceinline__ void gemm_stream(const float* __restrict__ W, int pw, int N, int K, const float* X, float* red) {
  if (pw == 128) gemm_stream_t<128>(W, pw, N, K, X, red);
  else if (pw == 512) gemm_stream_t<512>(W, pw, N, K, X, red);
  else if (pw == 1024) gemm_stream_t<1024>(W, pw, N, K, X, red);
  else gemm_stream_t<0>(W, pw, N, K, X, red);
}
// element o = n * 4 + s of the combined result (groups in order)
__device__ __forceinline__ float red_sum(const float* red, int N, int KG, int o) {
  float v = 0.f;
  for (int g = 0; g < KG; ++g) v += red[(size_t)g * N * 4 + o];
  return v;
}

__device__ __forceinline__ float quad_sum(float v) {       // over the 4 lanes (samples) of a feature, fixed order
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

#define SR_STAMP(slot)                                                                                   \
  do {                                                                                                   \
    if (a.tbuf && cta == 0 && threadIdx.x == 0 && e < 64) a.tbuf[(size_t)e * 16 + (slot)] = (unsigned long long)clock64(); \
  } while (0)

__device__ __forceinline__ void load_stage_table(StageTab* stb, const StageTab* src, int S) {
  const int4* s4 = reinterpret_cast<const int4*>(src);
  int4* d4 = reinterpret_cast<int4*>(stb);
  for (int i = threadIdx.x; i < S * (int)(sizeof(StageTab) / 16); i += kT) d4[i] = __ldg(s4 + i);
}

}  // namespace

// =================================================================================================== weight pre-pass
// wt[e] = [W1^T: Da x H | W2^T: H x H | W3^T: H x Dout]; 32 x 32 tiles through shared memory (coalesced both ways)
__global__ void __launch_bounds__(256) sr_pack_kernel(const StageTab* __restrict__ stb, const float* __restrict__ P, int Da, int H, int Dout,
                                                      long long wt_stage, float* __restrict__ wt, int t1, int t2, int t3) {
  __shared__ float tile[32][33];
  const int per = t1 + t2 + t3;
  const int e = blockIdx.x / per;
  int job = blockIdx.x - e * per;
  const StageTab& st = stb[e];
  // source [R][C] row-major -> destination [C][R]
  const float* src; float* dst; int R, C;
  if (job < t1) { src = P + st.p[P_W1]; dst = wt + (size_t)e * wt_stage; R = H; C = Da; }
  else if ((job -= t1) < t2) { src = P + st.p[P_W2]; dst = wt + (size_t)e * wt_stage + (size_t)Da * H; R = H; C = H; }
  else { job -= t2; src = P + st.p[P_W3]; dst = wt + (size_t)e * wt_stage + (size_t)Da * H + (size_t)H * H; R = Dout; C = H; }
  const int tc = (C + 31) / 32;
  const int r0 = (job / tc) * 32, c0 = (job % tc) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8)
    tile[i][tx] = (r0 + i < R && c0 + tx < C) ? __ldg(src + (size_t)(r0 + i) * C + c0 + tx) : 0.f;
  __syncthreads();
  for (int i = ty; i < 32; i += 8)
    if (c0 + i < C && r0 + tx < R) dst[(size_t)(c0 + i) * R + r0 + tx] = tile[tx][i];
}

// =================================================================================================== forward
// Per-stage block of small parameter vectors in shared memory (cp.async, one stage ahead, double-buffered):
//   [b1 | gamma1 | beta1 | b2 | gamma2 | beta2 (H each) | scale | b3 (Dout each) | running mean1, var1, mean2, var2 (H each) | lsi, loc]
struct FwdBlk {
  int b1, g1, be1, b2, g2, be2, sc, b3, rs, act, floats;
  __host__ __device__ FwdBlk(int H, int Dout)
      : b1(0), g1(H), be1(2 * H), b2(3 * H), g2(4 * H), be2(5 * H), sc(6 * H), b3(6 * H + Dout), rs(6 * H + 2 * Dout),
        act(10 * H + 2 * Dout), floats(10 * H + 2 * Dout + 4) {}
};
// backward: [scale (Dout) | gamma1 | gamma2 (H each) | saved mean1, rstd1, mean2, rstd2 (H each) | lsi, loc]
struct BwdBlk {
  int sc, g1, g2, st, act, floats;
  __host__ __device__ BwdBlk(int H, int Dout) : sc(0), g1(Dout), g2(Dout + H), st(Dout + 2 * H), act(Dout + 6 * H), floats(Dout + 6 * H + 4) {}
};
namespace {
// n floats (multiple of 4, both sides 16-byte aligned) by the threads [t0, t0 + n / 4) of the CTA
__device__ __forceinline__ void copy_vec(float* dst, const float* src, int n, int t0) {
  const int i = (int)threadIdx.x - t0;
  if (i >= 0 && 4 * i < n) cp_async16(dst + 4 * i, src + 4 * i, 16);
}
}  // namespace

__global__ void __launch_bounds__(kT, 1) sr_fwd_kernel(const __grid_constant__ SArgs A) {
  extern __shared__ __align__(16) float sm[];
  const KArgs& a = A.k;
  const Plan& p = A.p;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int D = a.D, Da = a.Da, Db = a.Db, Dout = a.Dout, H = a.H, S = a.S, B = a.B, LB = a.ldb;
  const int cta = (int)cluster_rank();
  const int s_base = cta * kSP;
  const bool live = s_base < B;                           // B % 4 == 0: a slab is all samples or all padding
  float* Xb[2] = {sm + p.oX0, sm + p.oX1};
  float* XA = sm + p.oXA;
  float* U1 = sm + p.oU1;
  float* U2 = sm + p.oU2;
  float* red = sm + p.oRed;
  float* part = sm + p.oPart;                             // [2 parities][2 * kMaxH]
  float* esc = sm + p.oEsc;                               // exp(3 scale) of the current stage
  int* idxb[2] = {reinterpret_cast<int*>(sm + p.oIdx), reinterpret_cast<int*>(sm + p.oIdx) + D};
  float* misc = sm + p.oMisc;
  StageTab* stb = reinterpret_cast<StageTab*>(sm + p.oSt);
  const FwdBlk fb(H, Dout);
  float* pbk[2] = {sm + p.oPB, sm + p.oPB + ((fb.floats + 3) & ~3)};
  const bool train_bn = a.has_bn && a.train;
  const float* wt = a.W + p.oWT;
  const int s_me = t & 3, jh = t >> 2;                    // element-wise steps: sample, first coupling column / hidden feature
  float ld_acc = 0.f;                                     // this thread's share of -sum log_s over all stages

  // stage e's index row and small vectors -> buffer e & 1 (thread ranges chosen so that one stage is ~2 copies per thread)
  auto issue_stage = [&](int e) {
    const StageTab& st = stb[e];
    float* pb = pbk[e & 1];
    int* ix = idxb[e & 1];
    for (int i = t; i < D; i += kT) cp_async4(reinterpret_cast<float*>(ix + i), reinterpret_cast<const float*>(a.gmaps + (size_t)e * D + i), true);
    const int hq = H >> 2;
    copy_vec(pb + fb.b1, a.P + st.p[P_B1], H, 0);
    copy_vec(pb + fb.b2, a.P + st.p[P_B2], H, hq);
    if (a.has_bn) {
      copy_vec(pb + fb.g1, a.P + st.p[P_G1], H, 2 * hq);
      copy_vec(pb + fb.be1, a.P + st.p[P_BE1], H, 3 * hq);
      copy_vec(pb + fb.g2, a.P + st.p[P_G2], H, 4 * hq);
      copy_vec(pb + fb.be2, a.P + st.p[P_BE2], H, 5 * hq);
      for (int k = 0; k < 4; ++k) copy_vec(pb + fb.rs + k * H, a.R + st.r[k], H, (6 + k) * hq);
    }
    copy_vec(pb + fb.sc, a.P + st.p[P_SCALE], Dout, 0);
    copy_vec(pb + fb.b3, a.P + st.p[P_B3], Dout, Dout >> 2);
    if (t == kT - 1) { cp_async4(pb + fb.act, a.P + st.lsi, true); cp_async4(pb + fb.act + 1, a.P + st.loc, true); }
  };

  load_stage_table(stb, a.st, S);
  __syncthreads();
  issue_stage(0);
  cp_async_commit();
  // input slab z [B][D] (sample-major) -> X [D][4]; z^T for the backward pass and the parameter gradients
  for (int i = t; i < D * kSP; i += kT) {
    const int s = i / D, f = i - s * D;
    Xb[0][f * kSP + s] = live ? __ldg(a.in + (size_t)(s_base + s) * D + f) : 0.f;
  }
  __syncthreads();
  if (live)
    for (int f = t; f < D; f += kT) stcg4(a.W + a.oZT + (size_t)f * LB + s_base, ld4(Xb[0] + 4 * f));

#pragma unroll 1
  for (int e = 0; e < S; ++e) {
    const StageTab& st = stb[e];
    float* X = Xb[e & 1];
    float* Xn = Xb[(e + 1) & 1];
    const int* idx = idxb[e & 1];
    const float* pb = pbk[e & 1];
    SR_STAMP(0);
    cp_async_wait<0>();
    __syncthreads();                                      // this stage's index row / vectors landed; previous stage fully written
    SR_STAMP(1);
    if (e + 1 < S) {
      issue_stage(e + 1);
      cp_async_commit();
    }
    // ---- gathered a-half of the stage input; exp(3 scale) once per stage
    for (int k = t; k < Da; k += kT) st4(XA + 4 * k, ld4(X + 4 * idx[k]));
    for (int j = t; j < Dout; j += kT) esc[j] = expf(3.f * pb[fb.sc + j]);
    __syncthreads();
    SR_STAMP(2);
    const float* w1t = wt + (size_t)e * p.wt_stage;
    const float* w2t = w1t + (size_t)Da * H;
    const float* w3t = w2t + (size_t)H * H;
#pragma unroll
    for (int which = 1; which <= 2; ++which) {
      float* U = which == 1 ? U1 : U2;
      // ---- net.0 / net.3 + LeakyReLU
      if (which == 1) gemm_stream(w1t, H, H, Da, XA, red);
      else gemm_stream(w2t, H, H, H, U1, red);
      __syncthreads();
      SR_STAMP(which == 1 ? 3 : 6);
      // ---- bias + LeakyReLU, saved; BatchNorm (batch statistics over the cluster / running statistics / none), saved:
      //      thread = (feature jh, sample s_me); the 4 lanes of a feature share its statistics through shuffles
      const bool on = t < 4 * H;
      const int jf = on ? jh : 0;
      float v = lrelu_f(red_sum(red, H, split_of(H, which == 1 ? Da : H).KG, on ? t : 0) + pb[(which == 1 ? fb.b1 : fb.b2) + jf]);
      if (on && live) __stcg(a.W + (which == 1 ? a.oL1T : a.oL2T) + ((size_t)e * H + jf) * LB + s_base + s_me, v);
      if (a.has_bn) {
        const float rm = pb[fb.rs + (which - 1) * 2 * H + jf], rv = pb[fb.rs + ((which - 1) * 2 + 1) * H + jf];
        float mean, rstd;
        if (train_bn) {
          float* pp = part + ((2 * e + which - 1) & 1) * (2 * kMaxH);
          const float m = 0.25f * quad_sum(v);
          const float d = v - m;
          const float q = quad_sum(d * d);
          if (on && s_me == 0) { pp[jf] = m; pp[kMaxH + jf] = q; }
          cluster_sync_all();                             // every CTA's (mean, M2) of its 4 samples is in its shared memory
          SR_STAMP(which == 1 ? 4 : 7);
          // every live slab holds kSP samples: mean of the partial means, M2 = sum M2_r + kSP sum (m_r - mean)^2 (exact
          // identity for equal counts; nothing cancels). Lane s of a feature reads ranks s, s + 4, s + 8, s + 12.
          const int nl = B / kSP;
          float pm[4], p2[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const unsigned r = (unsigned)(s_me + 4 * i);
            pm[i] = ld_dsmem(pp + jf, r);
            p2[i] = ld_dsmem(pp + kMaxH + jf, r);
          }
          float ps = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) ps += (s_me + 4 * i < nl) ? pm[i] : 0.f;
          mean = quad_sum(ps) / (float)nl;
          float p2s = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float dd = pm[i] - mean;
            p2s += (s_me + 4 * i < nl) ? fmaf((float)kSP * dd, dd, p2[i]) : 0.f;
          }
          const float var = quad_sum(p2s) / (float)B;     // biased (normalisation), unbiased for the running estimate
          rstd = 1.0f / sqrtf(var + kBnEps);
          if (cta == 0 && on && s_me == 0) {
            float* stt = a.W + a.oST + ((size_t)e * 4 + (which - 1) * 2) * H;
            __stcg(stt + jf, mean);
            __stcg(stt + H + jf, rstd);
            a.R[st.r[(which - 1) * 2] + jf] = (1.f - kBnMomentum) * rm + kBnMomentum * mean;
            a.R[st.r[(which - 1) * 2 + 1] + jf] = (1.f - kBnMomentum) * rv + kBnMomentum * (var * (float)B / (float)(B - 1));
          }
        } else {
          mean = rm;
          rstd = 1.0f / sqrtf(rv + kBnEps);
        }
        v = fmaf(v - mean, pb[(which == 1 ? fb.g1 : fb.g2) + jf] * rstd, pb[(which == 1 ? fb.be1 : fb.be2) + jf]);
      }
      if (on) {
        U[t] = v;
        if (live) __stcg(a.W + (which == 1 ? a.oN1T : a.oN2T) + ((size_t)e * H + jf) * LB + s_base + s_me, v);
      }
      __syncthreads();
      SR_STAMP(which == 1 ? 5 : 8);
    }
    // ---- net.6.fc, then (per coupling column, from the combined partials): ZeroFC bias / exp(3 scale), tanh-exp coupling
    //      (:244-249), ActNorm.reverse (:66), logdet term; thread = (sample t & 3, columns jh + 128 i)
    gemm_stream(w3t, Dout, Dout, H, U2, red);
    __syncthreads();
    SR_STAMP(9);
    {
      const int KG3 = split_of(Dout, H).KG;
      const float qa = expf(pb[fb.act]), qc = -pb[fb.act + 1];
      float* Og = a.W + a.oOT + (size_t)e * Dout * LB + s_base + s_me;
      float* Yg = a.W + a.oYT + (size_t)e * D * LB + s_base + s_me;
      for (int j = jh; j < Db; j += kT / 4) {
        const float ol = (red_sum(red, Dout, KG3, j * 4 + s_me) + pb[fb.b3 + j]) * esc[j];
        const float ot = (red_sum(red, Dout, KG3, (Db + j) * 4 + s_me) + pb[fb.b3 + Db + j]) * esc[Db + j];
        const float bv = X[idx[Da + j] * kSP + s_me];
        const float ls = tanhf(ol);
        const float yb = fmaf(bv / expf(ls) - ot, qa, qc);
        Xn[(Da + j) * kSP + s_me] = yb;
        ld_acc -= ls;
        if (live) {
          __stcg(Og + (size_t)j * LB, ol);
          __stcg(Og + (size_t)(Db + j) * LB, ot);
          __stcg(Yg + (size_t)(Da + j) * LB, yb);
        }
      }
      for (int k = t; k < Da; k += kT) {                  // pass-through half (incl. the extra row of an odd split, :241)
        const float4 v = ld4(XA + 4 * k);
        const float4 y = make_float4(fmaf(v.x, qa, qc), fmaf(v.y, qa, qc), fmaf(v.z, qa, qc), fmaf(v.w, qa, qc));
        st4(Xn + 4 * k, y);
        if (live) stcg4(a.W + a.oYT + ((size_t)e * D + k) * LB + s_base, y);
      }
    }
    SR_STAMP(10);
  }
  __syncthreads();
  // ---- outputs: x [B][D] sample-major; logdet = sum of the coupling terms + D * sum_e log_scale_inv_e (:60-64)
  {
    const float* X = Xb[S & 1];
    if (live)
      for (int i = t; i < D * kSP; i += kT) {
        const int s = i / D, f = i - s * D;
        a.out[(size_t)(s_base + s) * D + f] = X[f * kSP + s];
      }
    float v = ld_acc;                                     // lanes with equal (lane & 3) hold the same sample
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    if (lane < 4) misc[warp * 4 + lane] = v;
    __syncthreads();
    if (t < kSP && live) {
      float tot = 0.f;
      for (int w = 0; w < kT / 32; ++w) tot += misc[w * 4 + t];
      float lsum = 0.f;
      for (int e = 0; e < S; ++e) lsum += __ldg(a.P + stb[e].lsi);
      a.logdet[s_base + t] = tot + (float)D * lsum;
    }
  }
  cluster_sync_all();      // no CTA leaves while a peer may still read its shared memory
}

// =================================================================================================== backward chain
// Gradient state GX [D][4] (rows in the order of the current stage's output) walks the stages backwards; per stage the
// rows g_pre / p2 / p1 (gradients at the outputs of net.6.fc / net.3 / net.0, after the LeakyReLU and BatchNorm
// backward) are stored for nar::param_grads; the BatchNorm gamma / beta gradients are the two batch sums the BatchNorm
// backward needs anyway and are written directly.
__global__ void __launch_bounds__(kT, 1) sr_bwd_kernel(const __grid_constant__ SArgs A) {
  extern __shared__ __align__(16) float sm[];
  __shared__ float sRed[2 * (kT / 32)];
  const KArgs& a = A.k;
  const Plan& p = A.p;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int D = a.D, Da = a.Da, Db = a.Db, Dout = a.Dout, H = a.H, S = a.S, B = a.B, LB = a.ldb;
  const int cta = (int)cluster_rank();
  const int s_base = cta * kSP;
  const bool live = s_base < B;
  float* Gb[2] = {sm + p.oX0, sm + p.oX1};
  float* U1 = sm + p.oU1;
  float* U2 = sm + p.oU2;
  float* red = sm + p.oRed;
  float* part = sm + p.oPart;
  float* esc = sm + p.oEsc;
  int* idx3 = reinterpret_cast<int*>(sm + p.oIdx);       // ring of 3 index rows
  float* GP = sm + p.oGP;
  float* Act[2] = {sm + p.oAct0, sm + p.oAct1};
  StageTab* stb = reinterpret_cast<StageTab*>(sm + p.oSt);
  const BwdBlk bb(H, Dout);
  float* pbk[2] = {sm + p.oPB, sm + p.oPB + ((bb.floats + 3) & ~3)};
  const float invB = 1.0f / (float)B;
  const int s_me = t & 3, jh = t >> 2;
  const float g_ld_me = live ? __ldg(a.g_ld + s_base + s_me) : 0.f;

  auto issue_idx = [&](int slot, int e) {
    for (int i = t; i < D; i += kT)
      cp_async4(reinterpret_cast<float*>(idx3 + slot * D + i), reinterpret_cast<const float*>(a.gmaps + (size_t)e * D + i), true);
  };
  // saved activations of stage e for this slab -> rows [y (D) | log-s rows of o (Db) | b_prev (Db) | l1 (H) | l2 (H)] (its index
  // row must already be in shared memory), and the stage's small vectors
  auto issue_stage = [&](int buf, int e, const int* gm) {
    float* dst = Act[buf];
    const float* yprev = e == 0 ? a.W + a.oZT : a.W + a.oYT + (size_t)(e - 1) * D * LB;
    for (int r = t; r < p.actRows; r += kT) {
      const float* src;
      if (r < D) src = a.W + a.oYT + ((size_t)e * D + r) * LB;
      else if (r < D + Db) src = a.W + a.oOT + ((size_t)e * Dout + (r - D)) * LB;
      else if (r < D + 2 * Db) src = yprev + (size_t)gm[Da + (r - D - Db)] * LB;
      else if (r < D + 2 * Db + H) src = a.W + a.oL1T + ((size_t)e * H + (r - D - 2 * Db)) * LB;
      else src = a.W + a.oL2T + ((size_t)e * H + (r - D - 2 * Db - H)) * LB;
      cp_async16(dst + r * kSP, live ? src + s_base : src, live ? 16 : 0);
    }
    const StageTab& st = stb[e];
    float* pb = pbk[buf];
    const int hq = H >> 2;
    copy_vec(pb + bb.sc, a.P + st.p[P_SCALE], Dout, 0);
    if (a.has_bn) {
      copy_vec(pb + bb.g1, a.P + st.p[P_G1], H, Dout >> 2);
      copy_vec(pb + bb.g2, a.P + st.p[P_G2], H, (Dout >> 2) + hq);
      copy_vec(pb + bb.st, a.W + a.oST + (size_t)e * 4 * H, 4 * H, 0);
    }
    if (t == kT - 1) { cp_async4(pb + bb.act, a.P + st.lsi, true); cp_async4(pb + bb.act + 1, a.P + st.loc, true); }
  };

  load_stage_table(stb, a.st, S);
  issue_idx(0, S - 1);
  if (S > 1) issue_idx(1, S - 2);
  cp_async_commit();
  // gradient wrt the flow output, g_out [B][D] sample-major -> GX [D][4]
  for (int i = t; i < D * kSP; i += kT) {
    const int s = i / D, f = i - s * D;
    Gb[0][f * kSP + s] = live ? __ldg(a.g_out + (size_t)(s_base + s) * D + f) : 0.f;
  }
  cp_async_wait<0>();
  __syncthreads();
  issue_stage(0, S - 1, idx3);
  cp_async_commit();

#pragma unroll 1
  for (int u = 0; u < S; ++u) {
    const int e = S - 1 - u;
    const StageTab& st = stb[e];
    float* GX = Gb[u & 1];
    float* GXn = Gb[(u + 1) & 1];
    const float* act = Act[u & 1];
    const float* Ys = act;
    const float* Os = act + D * kSP;
    const float* BPs = Os + Db * kSP;
    const float* L1s = BPs + Db * kSP;
    const float* L2s = L1s + H * kSP;
    const int* idx = idx3 + (u % 3) * D;
    const float* pb = pbk[u & 1];
    SR_STAMP(0);
    cp_async_wait<0>();
    __syncthreads();
    SR_STAMP(1);
    if (u + 1 < S) {                                      // stage e - 1: activations / vectors; stage e - 2: index row
      issue_stage((u + 1) & 1, e - 1, idx3 + ((u + 1) % 3) * D);
      if (u + 2 < S) issue_idx((u + 2) % 3, e - 2);
      cp_async_commit();
    }
    for (int j = t; j < Dout; j += kT) esc[j] = expf(3.f * pb[bb.sc + j]);
    const float loc = pb[bb.act + 1];
    const float e_lsi = expf(pb[bb.act]);
    __syncthreads();
    SR_STAMP(2);
    // ---- ActNorm.reverse + coupling tail, elementwise: g_pre, the b-half input gradient, ActNorm sums
    float S1 = 0.f, S2 = 0.f;
    {
      float* Gg = a.W + p.pg.oGPRE + (size_t)e * Dout * LB + s_base + s_me;
      for (int j = jh; j < Db; j += kT / 4) {
        const float gb = GX[(Da + j) * kSP + s_me];
        S1 += gb;
        S2 = fmaf(gb, Ys[(Da + j) * kSP + s_me] + loc, S2);
        const float gbp = gb * e_lsi;
        const float ls = tanhf(Os[j * kSP + s_me]), ei = expf(-ls);
        const float g_ls = -gbp * BPs[j * kSP + s_me] * ei - g_ld_me;
        const float v_ls = g_ls * (1.f - ls * ls) * esc[j];
        const float v_t = -gbp * esc[Db + j];
        GP[j * kSP + s_me] = v_ls;
        GP[(Db + j) * kSP + s_me] = v_t;
        GXn[idx[Da + j] * kSP + s_me] = gbp * ei;
        if (live) {
          __stcg(Gg + (size_t)j * LB, v_ls);
          __stcg(Gg + (size_t)(Db + j) * LB, v_t);
        }
      }
      for (int k = t; k < Da; k += kT) {
        const float4 g = ld4(GX + 4 * k), y = ld4(Ys + 4 * k);
        S1 += (g.x + g.y) + (g.z + g.w);
        S2 = fmaf(g.x, y.x + loc, fmaf(g.y, y.y + loc, fmaf(g.z, y.z + loc, fmaf(g.w, y.w + loc, S2))));
      }
      S1 = warp_sum(S1); S2 = warp_sum(S2);
      if (lane == 0) { sRed[warp] = S1; sRed[kT / 32 + warp] = S2; }
    }
    __syncthreads();
    if (t == 0) {
      float t1 = 0.f, t2 = 0.f;
      for (int w = 0; w < kT / 32; ++w) { t1 += sRed[w]; t2 += sRed[kT / 32 + w]; }
      float* ap = a.W + p.pg.oActPart + ((size_t)e * kCL + cta) * 2;
      __stcg(ap, live ? t1 : 0.f);
      __stcg(ap + 1, live ? t2 : 0.f);
    }
    // ---- through net.6.fc: g_n2[h][s] = sum_j W3[j][h] g_pre[j][s]  (W3 stored [Dout][H]: k-major as it is)
    SR_STAMP(3);
    gemm_stream(a.P + st.p[P_W3], H, H, Dout, GP, red);
    __syncthreads();
    SR_STAMP(4);
#pragma unroll
    for (int which = 2; which >= 1; --which) {
      float* U = which == 2 ? U2 : U1;
      const float* Ls = which == 2 ? L2s : L1s;
      if (which == 1) {
        // ---- through net.3: g_n1[k][s] = sum_n W2[n][k] p2[n][s]
        gemm_stream(a.P + st.p[P_W2], H, H, H, U2, red);
        __syncthreads();
        SR_STAMP(6);
      }
      // ---- BatchNorm + LeakyReLU backward: thread = (feature jh, sample s_me); rows kept for the weight gradients
      const bool on = t < 4 * H;
      const int jf = on ? jh : 0;
      float g = red_sum(red, H, split_of(H, which == 2 ? Dout : H).KG, on ? t : 0);
      const float l = Ls[jf * kSP + s_me];
      if (a.has_bn) {
        float* pp = part + ((2 * u + (2 - which)) & 1) * (2 * kMaxH);
        const float m = pb[bb.st + (which - 1) * 2 * H + jf], r = pb[bb.st + ((which - 1) * 2 + 1) * H + jf];
        const float xh = (l - m) * r;
        const float a0 = quad_sum(live ? g : 0.f), a1 = quad_sum(live ? g * xh : 0.f);
        if (on && s_me == 0) { pp[jf] = a0; pp[kMaxH + jf] = a1; }
        cluster_sync_all();
        float q0 = 0.f, q1 = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {                     // lane s of a feature adds ranks s, s + 4, s + 8, s + 12
          const unsigned rk = (unsigned)(s_me + 4 * i);
          q0 += ld_dsmem(pp + jf, rk);
          q1 += ld_dsmem(pp + kMaxH + jf, rk);
        }
        const float s0 = quad_sum(q0), s1 = quad_sum(q1);
        if (cta == 0 && on && s_me == 0) {                // BatchNorm weight / bias gradients are exactly these sums
          a.G[st.p[which == 2 ? P_G2 : P_G1] + jf] = s1;
          a.G[st.p[which == 2 ? P_BE2 : P_BE1] + jf] = s0;
        }
        g = pb[(which == 2 ? bb.g2 : bb.g1) + jf] * r * (g - s0 * invB - xh * s1 * invB);
      }
      g = live ? g * (l > 0.f ? 1.f : kLreluSlope) : 0.f;
      if (on) {
        U[t] = g;
        if (live) __stcg(a.W + (which == 2 ? p.pg.oGP2 : p.pg.oGP1) + ((size_t)e * H + jf) * LB + s_base + s_me, g);
      }
      __syncthreads();
      SR_STAMP(which == 2 ? 5 : 7);
    }
    // ---- through net.0 (W1 stored [H][Da]: k-major), plus the direct path of the pass-through half through ActNorm
    gemm_stream(a.P + st.p[P_W1], Da, Da, H, U1, red);
    __syncthreads();
    SR_STAMP(8);
    {
      const int KG1 = split_of(Da, H).KG;
      for (int o = t; o < Da * kSP; o += kT)
        GXn[idx[o >> 2] * kSP + (o & 3)] = fmaf(GX[o], e_lsi, red_sum(red, Da, KG1, o));
    }
    SR_STAMP(9);
  }
  __syncthreads();
  if (a.g_in && live) {
    const float* GX = Gb[S & 1];
    for (int i = t; i < D * kSP; i += kT) {
      const int s = i / D, f = i - s * D;
      a.g_in[(size_t)(s_base + s) * D + f] = GX[f * kSP + s];
    }
  }
  cluster_sync_all();
}

// =================================================================================================== host side
namespace {
int g_dev = -1, g_ok = 0;
constexpr int kSmemMax = 216 * 1024;

cudaLaunchConfig_t make_cfg(cudaLaunchAttribute* at, int smem, cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kCL); cfg.blockDim = dim3(kT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kCL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cfg;
}
}  // namespace

int query(int device) {
  if (g_dev == device) return g_ok;
  g_dev = device;
  g_ok = 0;
  const void* ks[2] = {(const void*)sr_fwd_kernel, (const void*)sr_bwd_kernel};
  for (const void* k : ks)
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax) != cudaSuccess ||
        cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
  cudaLaunchAttribute at[1];
  cudaLaunchConfig_t cfg = make_cfg(at, kSmemMax, nullptr);
  int nc = 0;
  if (cudaOccupancyMaxActiveClusters(&nc, (const void*)sr_bwd_kernel, &cfg) != cudaSuccess || nc < 1) {
    cudaGetLastError();
    return 0;
  }
  g_ok = 1;
  return 1;
}

void make_plan(Plan& p, int avail, int D, int H, int B, int S) {
  p = Plan{};
  const int Da = D - D / 2, Db = D / 2, Dout = 2 * Db;
  if (!avail || D > kMaxD || (D & 7) || H > kMaxH || (H & 3) || B < 4 || B > kCL * kSP || (B & 3) || S > 128) return;
  if ((long long)Da * H + (long long)H * H + (long long)Dout * H > kMaxStageFloats) return;
  if (Db > 4 * (kT / 4)) return;                            // coupling columns per thread (4)
  int o = 0;
  auto take = [&](int n) { int r = o; o += (n + 3) & ~3; return r; };
  p.oSt = take(S * (int)(sizeof(StageTab) / 4));
  p.oX0 = take(D * kSP); p.oX1 = take(D * kSP);
  p.oXA = take(Da * kSP);
  p.oU1 = take(H * kSP); p.oU2 = take(H * kSP);
  p.oRed = take(16 * kT);
  p.oPart = take(4 * kMaxH);
  p.oIdx = take(3 * D);
  p.oMisc = take(kT / 32 * 4);
  p.oEsc = take(Dout);
  p.oPB = take(2 * ((FwdBlk(H, Dout).floats + 3) & ~3));   // the backward block is smaller
  p.smem_fwd = o * 4;
  p.oGP = take(Dout * kSP);
  p.actRows = D + 2 * Db + 2 * H;
  p.oAct0 = take(p.actRows * kSP); p.oAct1 = take(p.actRows * kSP);
  p.smem_bwd = o * 4;
  p.wt_stage = (long long)Da * H + (long long)H * H + (long long)H * Dout;
  p.pg = nar::Plan{};
  p.pg.G = kCL;
  p.ok = p.smem_bwd <= kSmemMax ? 1 : 0;
}

long long workspace_floats(const Plan& p, int D, int H, int B, int S) {
  const long long Dout = 2 * (D / 2), LB = (B + 3) / 4 * 4;
  return (long long)S * (Dout + 2 * H) * LB + 2LL * S * kCL + (long long)S * p.wt_stage + 256;
}
void place_workspace(Plan& p, long long base, int D, int H, int B, int S) {
  const long long Dout = 2 * (D / 2), LB = (B + 3) / 4 * 4;
  p.pg.oGPRE = base; base += (long long)S * Dout * LB;
  p.pg.oGP2 = base; base += (long long)S * H * LB;
  p.pg.oGP1 = base; base += (long long)S * H * LB;
  base = (base + 31) / 32 * 32;
  p.pg.oActPart = base; base += 2LL * S * kCL;
  base = (base + 31) / 32 * 32;
  p.oWT = base;
}

int forward(const KArgs& a, const Plan& p, cudaStream_t st) {
  SArgs A{};
  A.k = a;
  A.p = p;
  DPI_CUDA(cudaMemsetAsync(a.ctrl, 0, kCtrlWords * sizeof(unsigned), st));
  {
    const int th = (a.H + 31) / 32;
    const int t1 = th * ((a.Da + 31) / 32), t2 = th * th, t3 = ((a.Dout + 31) / 32) * th;
    sr_pack_kernel<<<a.S * (t1 + t2 + t3), 256, 0, st>>>(a.st, a.P, a.Da, a.H, a.Dout, p.wt_stage, a.W + p.oWT, t1, t2, t3);
    DPI_LAUNCH_CHECK();
  }
  cudaLaunchAttribute at[1];
  cudaLaunchConfig_t cfg = make_cfg(at, p.smem_fwd, st);
  DPI_CUDA(cudaLaunchKernelEx(&cfg, sr_fwd_kernel, A));
  count_launch();
  return DPI_OK;
}

int backward(const KArgs& a, const Plan& p, cudaStream_t st) {
  SArgs A{};
  A.k = a;
  A.p = p;
  DPI_CUDA(cudaMemsetAsync(a.ctrl, 0, kCtrlWords * sizeof(unsigned), st));
  cudaLaunchAttribute at[1];
  cudaLaunchConfig_t cfg = make_cfg(at, p.smem_bwd, st);
  DPI_CUDA(cudaLaunchKernelEx(&cfg, sr_bwd_kernel, A));
  count_launch();
  return nar::param_grads(a, p.pg, st);
}

}  // namespace sr
}  // namespace dpi
