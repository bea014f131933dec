// flow_small.cu - RealNVP pass for batch <= 64 as one persistent kernel of thread-block clusters.
//
// Math: generative_model/realnvpfc_model.py AffineCoupling.reverse :240-249, .forward :222-231, net :171-187,
// ZeroFC :137-141, ActNorm.reverse :49-66, .forward :30-47, Flow.reverse :457-474, RealNVP.reverse :594-603;
// backward per SURVEY.md Appendix A. Same feature-major workspace as the tiled phase program
// (flow_kernels.cu): X^T[feature][sample], permutations folded into row index maps.
//
// Why a second kernel: at batch <= 64 a coupling stage is three (forward) / four (backward) dependent
// hops of a few MFLOP each. The tiled program spends 5-9 us per hop in pipeline fill and per-k-tile
// synchronisation (profiles/phase_timeline_r01.txt). Here a hop is
//   wait(grid barrier) -> ONE cp.async batch of the activation rows this CTA needs (L2 hits) -> fp32 FFMA
//   micro-GEMM (4 samples x 4 features per thread) -> k-part reduction in shared memory -> optional split-K
//   reduction over the 8 CTAs of a cluster through distributed shared memory -> fused epilogue (bias,
//   LeakyReLU, BatchNorm batch statistics / tanh-exp coupling, ActNorm, logdet / BatchNorm backward) ->
//   arrive(grid barrier)
// and everything that does not depend on the previous hop runs in the barrier's shadow: the next hop's
// weight slice, row indices and per-feature parameters are cp.async'ed into shared memory, and the weight
// gradients of the current stage are computed there (they are only consumed by the optimiser).
// All reductions have a fixed order: results are bit-reproducible, there are no float atomics.
#include <algorithm>

#include "flow_host.cuh"

namespace dpi {

struct SArgs {
  KArgs k;
  SmallPlan sp;
  int bwd;
};

namespace {

constexpr int kT = 256;     // threads per CTA
constexpr int kMB = 64;     // samples per shared-memory row (batch padded with zeros)
constexpr int kLdG = 68;    // row stride of the weight-gradient operand tiles (conflict-free LDS.128)

enum Kind { K_TIN = 0, K_F1, K_F2, K_F3, K_FFIN, K_B0, K_DG2, K_DG1, K_DGIN, K_BFIN };

struct Sm {
  float *X, *W, *E, *P, *R, *Pm;
  float *X2, *W2;         // second halves of the K-chunk ring (== X / W when every hop is a single chunk)
  int *idx, *idx2;
  float* red;             // 8 * 64 + 128 floats (static)
  float* ldacc;           // this CTA's running logdet partial over the stages of the pass (64 floats, static)
  const StageTab* st;     // stage table (staged into shared memory when it fits)
  int phase;              // current phase (debug stamps)
};

// ------------------------------------------------------------------------------- synchronisation
__device__ __forceinline__ unsigned cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ float ld_dsmem(const float* p, unsigned rank) {
  const unsigned l = (unsigned)__cvta_generic_to_shared(p);
  unsigned r;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(l), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(r) : "memory");
  return v;
}

__device__ __forceinline__ unsigned ld_flag(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Split grid barrier on a monotonic counter (zeroed by the host before the launch). arrive publishes
// every global write of the CTA (bar.sync, then a gpu-scope release by thread 0); wait spins with acquire
// loads, bounded (4e9 cycles, about 2 s) so a protocol bug sets the sticky error flag ctrl[2] instead of hanging the
// GPU; the final phase of the pass turns the flag into NaN outputs (logdet / ActNorm gradients) so it cannot be missed.
__device__ __forceinline__ void grid_arrive(unsigned* ctrl) {
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctrl), "r"(1u) : "memory");
}
__device__ __forceinline__ void grid_wait(unsigned* ctrl, unsigned target) {
  if (threadIdx.x == 0) {
    // (polling with relaxed loads and one trailing fence.acq_rel.gpu measured 0.45 us SLOWER per barrier)
    unsigned cur;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cur) : "l"(ctrl) : "memory");
    if (cur < target) {
      const long long t0 = clock64();
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cur) : "l"(ctrl) : "memory");
        if (cur < target && (clock64() - t0 > 4000000000LL || ld_flag(ctrl + 2))) {
          atomicExch(ctrl + 2, 1u);   // sticky: every later wait of every CTA gives up at once
          break;
        }
      } while (cur < target);
    }
  }
  __syncthreads();
}

// Accurate expf / tanhf exist once in the binary (instruction-cache footprint, see the loaders' note).
__device__ __noinline__ float exp_nf(float x) { return expf(x); }
__device__ __noinline__ float tanh_nf(float x) { return tanhf(x); }

// ------------------------------------------------------------------------------- addressing (as flow_kernels.cu Ctx)
struct Cx {
  const KArgs& a;
  int e;
  const StageTab* stb;
  const StageTab& st;
  const int* gmap;
  __device__ Cx(const KArgs& a_, const StageTab* stb_, int e_)
      : a(a_), e(e_), stb(stb_), st(stb_[a_.dir ? a_.S - 1 - e_ : e_]),
        gmap(a_.gmaps + ((size_t)a_.dir * a_.S + e_) * a_.D) {}
  __device__ __forceinline__ const float* yprev() const { return e == 0 ? a.W + a.oZT : a.W + a.oYT + (size_t)(e - 1) * a.D * a.ldb; }
  __device__ __forceinline__ float* y() const { return a.W + a.oYT + (size_t)e * a.D * a.ldb; }
  __device__ __forceinline__ float* L(int which) const { return a.W + (which == 1 ? a.oL1T : a.oL2T) + (size_t)e * a.H * a.ldb; }
  __device__ __forceinline__ float* Nrm(int which) const { return a.W + (which == 1 ? a.oN1T : a.oN2T) + (size_t)e * a.H * a.ldb; }
  __device__ __forceinline__ float* O() const { return a.W + a.oOT + (size_t)e * a.Dout * a.ldb; }
  __device__ __forceinline__ float* stats(int which, int what) const { return a.W + a.oST + ((size_t)e * 4 + (which - 1) * 2 + what) * a.H; }
  __device__ __forceinline__ const float* P(int id) const { return a.P + st.p[id]; }
  __device__ __forceinline__ float* G(int id) const { return a.G + st.p[id]; }
  __device__ __forceinline__ const float* gy() const { return a.W + a.oGYT + (size_t)(e & 1) * a.D * a.ldb; }
  __device__ __forceinline__ float* gyprev() const { return a.W + a.oGYT + (size_t)((e - 1) & 1) * a.D * a.ldb; }
  // affine applied to the coupling output before it is stored (reverse: this stage's ActNorm.reverse :66;
  // forward: the NEXT stage's ActNorm.forward :45 - scalar, so it commutes with the permutation in between)
  __device__ __forceinline__ void post_affine(float& qa, float& qc) const {
    if (a.dir == 0) {
      qa = exp_nf(__ldg(a.P + st.lsi));
      qc = -__ldg(a.P + st.loc);
    } else if (e + 1 < a.S) {
      const StageTab& nx = stb[a.S - 2 - e];
      const float inv = 1.0f / exp_nf(__ldg(a.P + nx.lsi));
      qa = inv;
      qc = __ldg(a.P + nx.loc) * inv;
    } else {
      qa = 1.f; qc = 0.f;
    }
  }
};

// ------------------------------------------------------------------------------- asynchronous loaders
// The kernel is instruction-fetch sensitive (a phase runs once, then ~100 other phases run before it runs
// again): every helper below exists ONCE in the binary and all GEMM phases share them, so that the hot loop of
// a pass stays resident in the instruction cache. Loops are deliberately not unrolled and avoid integer
// division (tile extents are powers of two).
// All rows of the workspace have a stride of LB = round_up4(B) floats and every weight row a length that
// is a multiple of 4 (the planner guarantees it), so every copy is a 16-byte cp.async.
//
// dst[r * ld + 0..63] <- row (B valid samples, zero-filled to 64), r < nrows.
//   mode 0: row = r0 + r; mode 1: row = idx[r]; mode 2: pair: half = r >> np_log, j = r & (np - 1), row = idx[half * 64 + j]
//   valid while r0 + r (mode 2: r0 + j) < limit; invalid rows are zero-filled.
__device__ __forceinline__ void load_rows(float* dst, int ld, const float* base, const int* idx, int mode, int r0, int limit,
                                          int np_log, int LB, int B, int nrows) {
  const int m4 = (threadIdx.x & 15) * 4;
  const int left = B - m4;
  const int bytes = left >= 4 ? 16 : (left > 0 ? left * 4 : 0);
#pragma unroll 1
  for (int r = threadIdx.x >> 4; r < nrows; r += 16) {
    bool ok;
    int row;
    if (mode == 2) {
      const int half = r >> np_log, j = r & ((1 << np_log) - 1);
      ok = r0 + j < limit;
      row = ok ? idx[half * 64 + j] : 0;
    } else {
      ok = r0 + r < limit;
      row = ok ? (mode == 1 ? idx[r] : r0 + r) : 0;
    }
    const bool v = ok && bytes > 0;
    cp_async16(dst + r * ld + m4, v ? base + (size_t)row * LB + m4 : base, v ? bytes : 0);
  }
}
// K-contiguous weight rows (one warp per row): W_s[r * ldw + kk] <- Wg[row(r) * K + kb + kk], kk < knp, zero beyond
// klimit. row(r) = n0 + r (< N), or with np_log >= 0 the output-layer pairing (r >> np_log) * N + n0 + (r & (np-1)).
__device__ __forceinline__ void load_w_k(float* W_s, int ldw, const float* Wg, int K, int nrows, int kb, int knp, int klimit,
                                         int n0, int N, int np_log) {
  const int lane = threadIdx.x & 31;
#pragma unroll 1
  for (int r = threadIdx.x >> 5; r < nrows; r += 8) {
    int row;
    bool ok;
    if (np_log >= 0) {
      const int half = r >> np_log, j = n0 + (r & ((1 << np_log) - 1));
      ok = j < N;
      row = half * N + j;
    } else {
      ok = n0 + r < N;
      row = n0 + r;
    }
    const float* src = Wg + (size_t)(ok ? row : 0) * K + kb;
#pragma unroll 1
    for (int k4 = lane * 4; k4 < knp; k4 += 128) {
      const int left = ok ? klimit - (kb + k4) : 0;
      cp_async16(W_s + r * ldw + k4, left > 0 ? src + k4 : Wg, left >= 4 ? 16 : (left > 0 ? left * 4 : 0));
    }
  }
}
// N-contiguous weights: W_s[kk * ldw + nn] <- Wg[(kb + kk) * ldg + n0 + nn], nn < NT = 4 << cpr_log
__device__ __forceinline__ void load_w_n(float* W_s, int ldw, const float* Wg, int ldg, int kb, int knp, int klimit, int n0,
                                         int cpr_log, int N) {
#pragma unroll 1
  for (int c = threadIdx.x; c < (knp << cpr_log); c += kT) {
    const int kk = c >> cpr_log, n4 = (c & ((1 << cpr_log) - 1)) * 4;
    const int k = kb + kk;
    const int left = k < klimit ? N - (n0 + n4) : 0;
    const float* src = left > 0 ? Wg + (size_t)k * ldg + n0 + n4 : Wg;
    cp_async16(W_s + kk * ldw + n4, src, left >= 4 ? 16 : (left > 0 ? left * 4 : 0));
  }
}

// ------------------------------------------------------------------------------- micro-GEMMs
// acc[f][i] += sum_k W(4 fq + f, k) * X_s[k][4 sq + i] over this thread's k-part of the chunk [0, knp).
// WK: W_s[n][ldw] (K-contiguous), 4 k x 4 features per step: 8 LDS.128 per 64 FMA.
// !WK: W_s[k][ldw] (N-contiguous), one k per step: 2 LDS.128 per 16 FMA.
template <bool WK>
__device__ __forceinline__ void act_mac(float (&acc)[4][4], const float* X_s, const float* W_s, int ldw, int knp, int sq,
                                        int fq, int kp, int KP) {
  if (WK) {
#pragma unroll 4
    for (int k4 = 4 * kp; k4 < knp; k4 += 4 * KP) {
      float4 x[4], w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) x[u] = *reinterpret_cast<const float4*>(X_s + (k4 + u) * kMB + 4 * sq);
#pragma unroll
      for (int f = 0; f < 4; ++f) w[f] = *reinterpret_cast<const float4*>(W_s + (4 * fq + f) * ldw + k4);
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        const float wv[4] = {w[f].x, w[f].y, w[f].z, w[f].w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          acc[f][0] = fmaf(wv[u], x[u].x, acc[f][0]);
          acc[f][1] = fmaf(wv[u], x[u].y, acc[f][1]);
          acc[f][2] = fmaf(wv[u], x[u].z, acc[f][2]);
          acc[f][3] = fmaf(wv[u], x[u].w, acc[f][3]);
        }
      }
    }
  } else {
#pragma unroll 4
    for (int k = kp; k < knp; k += KP) {
      const float4 x = *reinterpret_cast<const float4*>(X_s + k * kMB + 4 * sq);
      const float4 w = *reinterpret_cast<const float4*>(W_s + k * ldw + 4 * fq);
      const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        acc[f][0] = fmaf(wv[f], x.x, acc[f][0]);
        acc[f][1] = fmaf(wv[f], x.y, acc[f][1]);
        acc[f][2] = fmaf(wv[f], x.z, acc[f][2]);
        acc[f][3] = fmaf(wv[f], x.w, acc[f][3]);
      }
    }
  }
}

// ------------------------------------------------------------------------------- phase descriptors
// Every GEMM phase (F1, F2, F3, DG2, DG1, DGIN) is the same code driven by one of these.
struct PD {
  int kind, e, which;
  // decomposition (copied from the plan)
  int N, K, NT, nt_log, n_tiles, KC, KS, nq_log, kpn;
  bool split, wk;
  // operands
  const float* Xbase; const int* xmap;      // activation rows (k axis) + optional row map
  const float* Wg; int w_ld;                // weights; w_ld: row length when N-contiguous
  int np_log, pair_db;                      // F3 row pairing (np_log < 0: none)
  // per-tile parameter vectors -> s.Pm[64 * i + t] = pm[i][n0 + t], t < vcnt  (nullptr: unused)
  const float* pm[5]; int pm_lim, vcnt;
  // per-tile index vectors -> s.idx2[64 * i + t] = i2[i][n0 + t]
  const int* i2[2]; int i2_lim;
  // extra activation rows -> s.E: emode 0 none, 1 rows n0 + r of Ebase, 2 pair rows through idx2, 3 this rank's
  // finalised features n0 + rank + nr * i of Ebase
  const float* Ebase; int emode, elim;
  // epilogue destinations (meaning per kind, see make_desc) and per-phase scalars
  float* o[6];
  float qa, qc, e_lsi;
};

__device__ __forceinline__ void plan_into(PD& d, const ActPlan& pl) {
  d.N = pl.N; d.K = pl.K; d.NT = pl.NT; d.nt_log = pl.nt_log; d.n_tiles = pl.n_tiles; d.KC = pl.KC; d.KS = pl.KS;
  d.nq_log = pl.nq_log; d.kpn = pl.kp; d.split = pl.ksplit != 0;
}

__device__ __noinline__ PD make_desc(const SArgs& A, const StageTab* stb, int kind, int e) {
  const KArgs& a = A.k;
  PD d;
  d.kind = kind; d.e = e; d.which = 0;
  d.xmap = nullptr; d.w_ld = 0; d.np_log = -1; d.pair_db = 0;
#pragma unroll
  for (int i = 0; i < 5; ++i) d.pm[i] = nullptr;
  d.i2[0] = d.i2[1] = nullptr;
  d.pm_lim = a.H; d.i2_lim = 0; d.Ebase = nullptr; d.emode = 0; d.elim = 0;
  d.N = d.K = d.NT = d.nt_log = d.n_tiles = d.KC = d.KS = d.nq_log = d.kpn = 0; d.split = false; d.wk = true;
  d.Xbase = nullptr; d.Wg = nullptr; d.vcnt = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) d.o[i] = nullptr;
  d.qa = 1.f; d.qc = 0.f; d.e_lsi = 1.f;
  if (kind < K_F1 || kind == K_FFIN || kind == K_B0 || kind == K_BFIN) return d;
  Cx c(a, stb, e);
  switch (kind) {
    case K_F1:
    case K_F2: {
      const int which = kind == K_F1 ? 1 : 2;
      d.which = which;
      plan_into(d, which == 1 ? A.sp.f1 : A.sp.f2);
      d.wk = true;
      d.Xbase = which == 1 ? c.yprev() : c.Nrm(1);
      d.xmap = which == 1 ? c.gmap : nullptr;
      d.Wg = c.P(which == 1 ? P_W1 : P_W2);
      d.vcnt = d.NT;
      d.pm[0] = c.P(which == 1 ? P_B1 : P_B2);
      if (a.has_bn) {
        d.pm[1] = c.P(which == 1 ? P_G1 : P_G2);
        d.pm[2] = c.P(which == 1 ? P_BE1 : P_BE2);
        d.pm[3] = a.R + c.st.r[(which - 1) * 2];
        d.pm[4] = a.R + c.st.r[(which - 1) * 2 + 1];
        d.o[2] = c.stats(which, 0); d.o[3] = c.stats(which, 1);
        d.o[4] = a.R + c.st.r[(which - 1) * 2]; d.o[5] = a.R + c.st.r[(which - 1) * 2 + 1];
      }
      d.o[0] = c.L(which); d.o[1] = c.Nrm(which);     // saved LeakyReLU output / normalised activation rows
      break;
    }
    case K_F3: {
      plan_into(d, A.sp.f3);
      d.wk = true;
      d.Xbase = c.Nrm(2);
      d.Wg = c.P(P_W3);
      d.np_log = d.nt_log - 1; d.pair_db = a.Db;
      d.vcnt = d.NT >> 1;
      d.pm[0] = c.P(P_SCALE); d.pm[1] = c.P(P_SCALE) + a.Db; d.pm[2] = c.P(P_B3); d.pm[3] = c.P(P_B3) + a.Db;
      d.pm_lim = a.Db;
      d.i2[0] = c.gmap; d.i2[1] = c.gmap + a.Da; d.i2_lim = a.Db;
      d.Ebase = c.yprev(); d.emode = 2; d.elim = a.Db;
      d.o[0] = c.O(); d.o[1] = c.y(); d.o[2] = a.W + a.oLDP;
      c.post_affine(d.qa, d.qc);
      break;
    }
    case K_DG2:
    case K_DG1: {
      const int which = kind == K_DG2 ? 2 : 1;
      d.which = which;
      plan_into(d, which == 2 ? A.sp.dg2 : A.sp.dg1);
      d.wk = false;
      d.Xbase = a.W + (which == 2 ? a.oGPRET : a.oGP2T);
      d.Wg = c.P(which == 2 ? P_W3 : P_W2); d.w_ld = a.H;
      d.vcnt = d.NT;
      if (a.has_bn) {
        d.pm[0] = c.stats(which, 0); d.pm[1] = c.stats(which, 1); d.pm[2] = c.P(which == 1 ? P_G1 : P_G2);
      }
      d.Ebase = c.L(which); d.emode = 3; d.elim = a.H;
      d.o[0] = a.W + (which == 2 ? a.oGP2T : a.oGP1T);
      if (a.has_bn) { d.o[1] = c.G(which == 1 ? P_G1 : P_G2); d.o[2] = c.G(which == 1 ? P_BE1 : P_BE2); }
      d.o[3] = c.G(which == 1 ? P_B1 : P_B2);
      break;
    }
    default: {   // K_DGIN
      plan_into(d, A.sp.dgin);
      d.wk = false;
      d.Xbase = a.W + a.oGP1T;
      d.Wg = c.P(P_W1); d.w_ld = a.Da;
      d.vcnt = d.NT;
      d.i2[0] = c.gmap; d.i2_lim = a.Da;
      d.Ebase = c.gy(); d.emode = 1; d.elim = a.Da;
      d.o[0] = c.gyprev();
      d.e_lsi = exp_nf(__ldg(a.P + c.st.lsi));
      break;
    }
  }
  return d;
}

// weight slice of one k-chunk of `tile` -> Wdst; with_idx: also the row indices of this rank's whole K range
// [k_lo, k_hi) -> s.idx (chunks index it at kb - k_lo), so that later chunks can be issued without waiting for them
__device__ __forceinline__ void issue_weights(const Sm& s, const PD& d, float* Wdst, int tile, int kb, int knp, int k_hi,
                                              bool with_idx, int k_lo) {
  if (d.wk) {
    if (d.np_log >= 0) load_w_k(Wdst, d.KS + 4, d.Wg, d.K, d.NT, kb, knp, k_hi, tile << d.np_log, d.pair_db, d.np_log);
    else load_w_k(Wdst, d.KS + 4, d.Wg, d.K, d.NT, kb, knp, k_hi, tile << d.nt_log, d.N, -1);
  } else {
    load_w_n(Wdst, d.NT + 4, d.Wg, d.w_ld, kb, knp, k_hi, tile << d.nt_log, d.nt_log - 2, d.N);
  }
  if (with_idx && d.xmap) {
    const int cnt = (k_hi - k_lo + 3) & ~3;
    for (int t = threadIdx.x; t < cnt; t += kT) {
      const bool ok = k_lo + t < k_hi;
      cp_async4(reinterpret_cast<float*>(s.idx + t), reinterpret_cast<const float*>(ok ? d.xmap + k_lo + t : d.xmap), ok);
    }
  }
}
// per-tile parameter / index vectors
__device__ __forceinline__ void issue_params(const Sm& s, const PD& d, int tile) {
  const int t = threadIdx.x & 63, which = threadIdx.x >> 6;      // 4 vectors per pass of the CTA
  const int n0 = tile * d.vcnt;
  if (t < d.vcnt) {
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      if ((i & 3) == which && d.pm[i]) {
        const bool ok = n0 + t < d.pm_lim;
        cp_async4(s.Pm + 64 * i + t, ok ? d.pm[i] + n0 + t : d.pm[i], ok);
      }
    }
    if (which >= 2 && d.i2[which - 2]) {
      const int* src = d.i2[which - 2];
      const bool ok = n0 + t < d.i2_lim;
      cp_async4(reinterpret_cast<float*>(s.idx2 + 64 * (which - 2) + t), reinterpret_cast<const float*>(ok ? src + n0 + t : src), ok);
    }
  }
}

#define STAMP(i)                                                                        \
  do {                                                                                  \
    if (a.tbuf && blockIdx.x == 0 && threadIdx.x == 0) {                                \
      a.tbuf[1024 + 8 * s.phase + (i)] = (unsigned long long)clock64();                 \
    }                                                                                   \
  } while (0)

// sum over the nr (<= 8) ranks of the cluster, in rank order, of R_s[nl][lane], R_s[nl][lane + 32]
__device__ __forceinline__ void combine(const Sm& s, int nl, int lane, int nr, float& s0, float& s1) {
  const float* p = s.R + nl * kMB + lane;
  if (nr == 1) {
    s0 = p[0]; s1 = p[32];
  } else {
    float v0[8], v1[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      v0[r] = r < nr ? ld_dsmem(p, r) : 0.f;
      v1[r] = r < nr ? ld_dsmem(p + 32, r) : 0.f;
    }
    s0 = 0.f; s1 = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) { s0 += v0[r]; s1 += v1[r]; }
  }
}

// ------------------------------------------------------------------------------- the generic GEMM phase
// prefetch: everything of the phase's first tile that does not depend on the previous phase
__device__ __forceinline__ void prefetch_gemm(const SArgs& A, const Sm& s, const PD& d) {
  if (d.kind >= K_F1 && d.kind != K_FFIN && d.kind != K_B0 && d.kind != K_BFIN) {
    const int unit = d.split ? blockIdx.x / A.sp.cl : blockIdx.x;
    const int rank = d.split ? (int)cluster_rank() : 0;
    if (unit < d.n_tiles) {
      const int k_lo = rank * d.KC, k_hi = min(d.K, k_lo + d.KC);
      const int kn = max(0, min(d.KS, k_hi - k_lo)), knp = (kn + 3) & ~3;
      issue_weights(s, d, s.W, unit, k_lo, knp, k_hi, true, k_lo);
      issue_params(s, d, unit);
    }
  }
  cp_async_commit();
}

__device__ __forceinline__ void run_gemm(const SArgs& A, const Sm& s, const PD& d, bool& cl_pending) {
  const KArgs& a = A.k;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int LB = a.ldb, NT = d.NT;
  const float invB = 1.0f / (float)a.B;
  const int unit = d.split ? blockIdx.x / A.sp.cl : blockIdx.x;
  const int units = d.split ? A.sp.G : A.sp.NC;
  const int rank = d.split ? (int)cluster_rank() : 0;
  const int nr = d.split ? A.sp.cl : 1;
  const int sq = threadIdx.x & 15, g = threadIdx.x >> 4, fq = g & ((1 << d.nq_log) - 1), kp = g >> d.nq_log;
  const int ldw = d.wk ? d.KS + 4 : NT + 4;
  const int k_lo = rank * d.KC, k_hi = min(d.K, k_lo + d.KC);
  const float qa = d.qa, qc = d.qc, e_lsi = d.e_lsi;   // per-phase scalars of the epilogues
  bool pf = true;
#pragma unroll 1
  for (int tile = unit; tile < d.n_tiles; tile += units) {
    const int n0 = tile * d.vcnt;          // first feature (F3: first coupling column) of the tile
    if (!pf) {
      __syncthreads();
      issue_params(s, d, tile);            // joins the weight batch committed below
    }
    // ---- operands -> shared memory, micro-GEMM over this rank's K range (in chunks of KS rows)
    float acc[4][4];
#pragma unroll
    for (int f = 0; f < 4; ++f)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[f][i] = 0.f;
    // K range of this rank in chunks of KS rows through a two-buffer ring (s.X / s.X2, s.W / s.W2): chunk c + 1 is
    // in flight while chunk c is multiplied. Single-chunk hops (C1) take exactly one pass.
    if (!pf) {
      __syncthreads();     // earlier readers of W / idx / X are done
      const int kn0 = max(0, min(d.KS, k_hi - k_lo));
      issue_weights(s, d, s.W, tile, k_lo, (kn0 + 3) & ~3, k_hi, true, k_lo);
      cp_async_commit();
    }
    cp_async_wait<0>();
    __syncthreads();       // first weight chunk + all row indices + parameters landed
    STAMP(1);
    {
      const int kn0 = max(0, min(d.KS, k_hi - k_lo));
      load_rows(s.X, kMB, d.Xbase, s.idx, d.xmap ? 1 : 0, k_lo, k_hi, 0, LB, a.B, (kn0 + 3) & ~3);
      if (d.emode) {
        if (d.emode == 3) {      // saved LeakyReLU outputs of the features this rank finalises
          const int nslots = max(0, (NT - rank + nr - 1) / nr);
          const int r = threadIdx.x >> 4;
          const int m4 = (threadIdx.x & 15) * 4, left = a.B - m4;
          const int bytes = left >= 4 ? 16 : (left > 0 ? left * 4 : 0);
#pragma unroll 1
          for (int i = r; i < nslots; i += 16) {
            const int n = n0 + rank + nr * i;
            const bool v = n < d.elim && bytes > 0;
            cp_async16(s.E + i * kMB + m4, v ? d.Ebase + (size_t)n * LB + m4 : d.Ebase, v ? bytes : 0);
          }
        } else {
          load_rows(s.E, kMB, d.Ebase, s.idx2, d.emode == 2 ? 2 : 0, n0, d.elim, d.np_log, LB, a.B, NT);
        }
      }
      cp_async_commit();
    }
    int ch = 0;
    const bool ring = s.X2 != s.X;       // false: shared memory only allowed one buffer -> chunks strictly in turn
#pragma unroll 1
    for (int kb = k_lo; ch == 0 || kb < k_hi; kb += d.KS, ++ch) {
      const int kn = max(0, min(d.KS, k_hi - kb)), knp = (kn + 3) & ~3;
      const int kb2 = kb + d.KS;
      float* Xc = (ch & 1) ? s.X2 : s.X;
      float* Wc = (ch & 1) ? s.W2 : s.W;
      if (!ring) {
        if (ch > 0) {
          __syncthreads();
          issue_weights(s, d, s.W, tile, kb, knp, k_hi, false, k_lo);
          load_rows(s.X, kMB, d.Xbase, s.idx + (kb - k_lo), d.xmap ? 1 : 0, kb, k_hi, 0, LB, a.B, knp);
          cp_async_commit();
        }
        cp_async_wait<0>();
      } else if (kb2 < k_hi) {      // next chunk -> the other buffers (free since the barrier that ended chunk ch - 1)
        const int kn2 = min(d.KS, k_hi - kb2), knp2 = (kn2 + 3) & ~3;
        issue_weights(s, d, (ch & 1) ? s.W : s.W2, tile, kb2, knp2, k_hi, false, k_lo);
        load_rows((ch & 1) ? s.X : s.X2, kMB, d.Xbase, s.idx + (kb2 - k_lo), d.xmap ? 1 : 0, kb2, k_hi, 0, LB, a.B, knp2);
        cp_async_commit();
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      if (ch == 0) STAMP(2);
      if (d.wk) act_mac<true>(acc, Xc, Wc, ldw, knp, sq, fq, kp, d.kpn);
      else act_mac<false>(acc, Xc, Wc, ldw, knp, sq, fq, kp, d.kpn);
      if (ring && kb2 < k_hi) __syncthreads();   // everyone is done with this chunk's buffers before they are refilled
    }
    pf = false;
    // ---- k-part partials -> P_s[kp][n][m] -> R_s[n][m]
#pragma unroll
    for (int f = 0; f < 4; ++f)
      *reinterpret_cast<float4*>(s.P + ((size_t)(kp * NT + 4 * fq + f)) * kMB + 4 * sq) =
          make_float4(acc[f][0], acc[f][1], acc[f][2], acc[f][3]);
    if (cl_pending) {      // the peers have finished reading the previous contents of R_s
      cluster_wait();
      cl_pending = false;
    }
    __syncthreads();
    STAMP(3);
#pragma unroll 1
    for (int i = threadIdx.x; i < NT * 16; i += kT) {
      const int n = i >> 4, m4 = (i & 15) * 4;
      const float* pp = s.P + (size_t)n * kMB + m4;
      const int pstride = NT * kMB;
      float4 t = *reinterpret_cast<const float4*>(pp);
      if (d.kpn >= 4) {     // four parts in flight, summed in part order
#pragma unroll 1
        for (int q = 0; q < d.kpn; q += 4) {
          const float4 v0 = q ? *reinterpret_cast<const float4*>(pp + (size_t)q * pstride) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 v1 = *reinterpret_cast<const float4*>(pp + (size_t)(q + 1) * pstride);
          const float4 v2 = *reinterpret_cast<const float4*>(pp + (size_t)(q + 2) * pstride);
          const float4 v3 = *reinterpret_cast<const float4*>(pp + (size_t)(q + 3) * pstride);
          t.x = ((t.x + v0.x) + v1.x) + (v2.x + v3.x); t.y = ((t.y + v0.y) + v1.y) + (v2.y + v3.y);
          t.z = ((t.z + v0.z) + v1.z) + (v2.z + v3.z); t.w = ((t.w + v0.w) + v1.w) + (v2.w + v3.w);
        }
      } else if (d.kpn == 2) {
        const float4 v1 = *reinterpret_cast<const float4*>(pp + pstride);
        t.x += v1.x; t.y += v1.y; t.z += v1.z; t.w += v1.w;
      }
      *reinterpret_cast<float4*>(s.R + n * kMB + m4) = t;
    }
    if (d.split) {
      cluster_arrive();
      cluster_wait();
    } else {
      __syncthreads();
    }
    STAMP(4);
    // ---- fused epilogue
    const bool ok0 = lane < a.B, ok1 = lane + 32 < a.B;
    if (d.kind == K_F1 || d.kind == K_F2) {
      // net.0-2 / net.3-5: bias + LeakyReLU(0.01) + BatchNorm1d(eps 1e-2)  (:171-187); one warp per feature
      const int which = d.which;
#pragma unroll 1
      for (int i = warp;; i += 8) {
        const int nl = rank + nr * i;
        if (nl >= NT) break;
        float s0, s1;
        combine(s, nl, lane, nr, s0, s1);
        const int n = n0 + nl;
        if (n >= a.H) continue;
        const float bj = s.Pm[nl];
        const float v0 = lrelu_f(s0 + bj), v1 = lrelu_f(s1 + bj);
        float* Lr = d.o[0] + (size_t)n * LB;
        if (ok0) __stcg(Lr + lane, v0);
        if (ok1) __stcg(Lr + lane + 32, v1);
        float n0v = v0, n1v = v1;
        if (a.has_bn) {
          const float gam = s.Pm[64 + nl], bet = s.Pm[128 + nl], rm = s.Pm[192 + nl], rv = s.Pm[256 + nl];
          if (a.train) {
            const float mean = warp_sum((ok0 ? v0 : 0.f) + (ok1 ? v1 : 0.f)) * invB;
            const float d0 = v0 - mean, d1 = v1 - mean;
            const float var = warp_sum((ok0 ? d0 * d0 : 0.f) + (ok1 ? d1 * d1 : 0.f)) * invB;
            const float rstd = 1.0f / sqrtf(var + kBnEps);
            const float ga = gam * rstd;
            n0v = fmaf(d0, ga, bet);
            n1v = fmaf(d1, ga, bet);
            if (lane == 0) {
              __stcg(d.o[2] + n, mean);
              __stcg(d.o[3] + n, rstd);
              d.o[4][n] = (1.f - kBnMomentum) * rm + kBnMomentum * mean;
              d.o[5][n] = (1.f - kBnMomentum) * rv + kBnMomentum * (var * (float)a.B / (float)(a.B - 1));
            }
          } else {   // eval: running statistics
            const float rs = 1.0f / sqrtf(rv + kBnEps);
            const float ga = gam * rs;
            const float be = bet - rm * ga;
            n0v = fmaf(v0, ga, be);
            n1v = fmaf(v1, ga, be);
          }
        }
        float* Nr = d.o[1] + (size_t)n * LB;
        if (ok0) __stcg(Nr + lane, n0v);
        if (ok1) __stcg(Nr + lane + 32, n1v);
      }
    } else if (d.kind == K_F3) {
      // ZeroFC (:137-141) + affine coupling (:244-249 / :226-231) + ActNorm (:66 / :45) + logdet partial
      const int NP = NT >> 1;
      float ld0 = 0.f, ld1 = 0.f;
#pragma unroll 1
      for (int p = warp; p < NP; p += 8) {
        const int col = n0 + p;
        if (col >= a.Db) continue;
        const float e_ls = exp_nf(3.f * s.Pm[p]), e_t = exp_nf(3.f * s.Pm[64 + p]);
        const float b_ls = s.Pm[128 + p], b_t = s.Pm[192 + p];
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
          const int m = lane + 32 * hh;
          const float ols = (s.R[p * kMB + m] + b_ls) * e_ls;
          const float ot = (s.R[(NP + p) * kMB + m] + b_t) * e_t;
          const float av = s.E[p * kMB + m], bv = s.E[(NP + p) * kMB + m];
          const float ls = tanh_nf(ols);
          float bp, dl;
          if (a.dir == 0) { bp = __fdividef(bv, exp_nf(ls)) - ot; dl = -ls; }
          else            { bp = (bv + ot) * exp_nf(ls); dl = ls; }
          if (m < a.B) {
            __stcg(d.o[0] + (size_t)col * LB + m, ols);
            __stcg(d.o[0] + (size_t)(a.Db + col) * LB + m, ot);
            __stcg(d.o[1] + (size_t)(a.Da + col) * LB + m, fmaf(bp, qa, qc));
            __stcg(d.o[1] + (size_t)col * LB + m, fmaf(av, qa, qc));
            if (hh == 0) ld0 += dl; else ld1 += dl;
          }
        }
      }
      // odd D: the a-half has one more row than the b-half (chunk(2,1) semantics, :241)
      if (a.Da > a.Db && tile == 0 && warp == 0) {
        const int col = a.Da - 1;
        const float* src = d.Ebase + (size_t)__ldg(d.i2[0] + col) * LB;
        for (int m = lane; m < a.B; m += 32) __stcg(d.o[1] + (size_t)col * LB + m, fmaf(__ldcg(src + m), qa, qc));
      }
      // per-sample logdet partial of this tile: the 8 warps' sums in fixed order
      s.red[warp * 64 + lane] = ld0;
      s.red[warp * 64 + lane + 32] = ld1;
      __syncthreads();
      if (threadIdx.x < 64) {
        float t = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) t += s.red[r * 64 + threadIdx.x];
        // running per-CTA logdet partial over all stages (fixed order: stage by stage, tile by tile); one row
        // per CTA is summed by the final phase instead of one row per (stage, tile)
        float* row = d.o[2] + (size_t)blockIdx.x * LB + threadIdx.x;
        const float run = (a.persistent ? s.ldacc[threadIdx.x] : __ldcg(row)) + t;
        s.ldacc[threadIdx.x] = run;
        if ((int)threadIdx.x < a.B) __stcg(row, run);
      }
      __syncthreads();
    } else if (d.kind == K_DG2 || d.kind == K_DG1) {
      // dgrad through net.6.fc.weight / net.3.weight + BatchNorm and LeakyReLU backward of hidden layer `which`
      float* dstbuf = d.o[0];
#pragma unroll 1
      for (int i = warp;; i += 8) {
        const int nl = rank + nr * i;
        if (nl >= NT) break;
        float g0, g1;
        combine(s, nl, lane, nr, g0, g1);
        const int n = n0 + nl;
        if (n >= a.H) continue;
        if (!ok0) g0 = 0.f;
        if (!ok1) g1 = 0.f;
        const float l0 = s.E[i * kMB + lane], l1 = s.E[i * kMB + lane + 32];
        float gl0 = g0, gl1 = g1, gbeta = 0.f, ggamma = 0.f;
        if (a.has_bn) {
          const float mean = s.Pm[nl], rstd = s.Pm[64 + nl], gam = s.Pm[128 + nl];
          const float xh0 = (l0 - mean) * rstd, xh1 = (l1 - mean) * rstd;
          gbeta = warp_sum(g0 + g1);
          ggamma = warp_sum(fmaf(g0, xh0, g1 * xh1));
          gl0 = gam * rstd * (g0 - gbeta * invB - xh0 * ggamma * invB);
          gl1 = gam * rstd * (g1 - gbeta * invB - xh1 * ggamma * invB);
        }
        const float p0 = ok0 ? gl0 * (l0 > 0.f ? 1.f : kLreluSlope) : 0.f;
        const float p1 = ok1 ? gl1 * (l1 > 0.f ? 1.f : kLreluSlope) : 0.f;
        const float gbias = warp_sum(p0 + p1);
        float* dr = dstbuf + (size_t)n * LB;
        if (ok0) __stcg(dr + lane, p0);
        if (ok1) __stcg(dr + lane + 32, p1);
        if (lane == 0) {
          if (a.has_bn) {
            d.o[1][n] = ggamma;
            d.o[2][n] = gbeta;
          }
          d.o[3][n] = gbias;
        }
      }
    } else {
      // K_DGIN: dgrad through net.0.weight + the direct path through the a-half and ActNorm, written through the
      // fused permutation into the rows of the gradient wrt the previous stage's output
#pragma unroll 1
      for (int nl = warp; nl < NT; nl += 8) {
        const int n = n0 + nl;
        if (n >= a.Da) continue;
        float* dr = d.o[0] + (size_t)s.idx2[nl] * LB;
        if (ok0) __stcg(dr + lane, fmaf(s.E[nl * kMB + lane], e_lsi, s.R[nl * kMB + lane]));
        if (ok1) __stcg(dr + lane + 32, fmaf(s.E[nl * kMB + lane + 32], e_lsi, s.R[nl * kMB + lane + 32]));
      }
    }
    if (d.split) {
      cluster_arrive();
      cl_pending = true;
    }
  }
}

// ------------------------------------------------------------------------------- transposes / finals (cold)
// 32 x 32 tiles through shared memory. mode:
//   0 pass input : in [B][D] -> ZT [D][LB] (forward direction: stage-0 ActNorm applied)
//   1 backward input: g_out [B][D] -> GYT[(S-1)&1] [D][LB]
//   2 reverse output: YT[S-1] [D][LB] -> out [B][D]
//   3 forward output: YT[S-1][gf[j]][m] -> out [B][D] (trailing permutation :590 after :446)
//   4 backward output: GYT[1] [D][LB] -> g_in [B][D]
__device__ __noinline__ void run_transpose(const KArgs& a, const StageTab* stb, int mode, float* sT, int NC) {
  const bool to_fm = mode <= 1;      // sample-major [B][D] -> feature-major [D][LB]
  const int LB = a.ldb;
  const float* src; float* dst; const int* map = nullptr;
  float fa = 1.f, fc = 0.f;
  if (mode == 0) {
    src = a.in; dst = a.W + a.oZT;
    if (a.dir == 1) {
      const StageTab& s0 = stb[a.S - 1];
      const float inv = 1.0f / expf(__ldg(a.P + s0.lsi));
      fa = inv; fc = __ldg(a.P + s0.loc) * inv;
    }
  } else if (mode == 1) {
    src = a.g_out; dst = a.W + a.oGYT + (size_t)((a.S - 1) & 1) * a.D * LB;
  } else if (mode == 2 || mode == 3) {
    src = a.W + a.oYT + (size_t)(a.S - 1) * a.D * LB; dst = a.out;
    if (mode == 3) map = a.gmaps + (size_t)2 * a.S * a.D;
  } else {
    src = a.W + a.oGYT + (size_t)1 * a.D * LB; dst = a.g_in;
  }
  const int R = to_fm ? a.B : a.D, C = to_fm ? a.D : a.B;
  const int lds = to_fm ? a.D : LB, ldd = to_fm ? LB : a.D;
  const int tiles_c = (C + 31) / 32, tiles_r = (R + 31) / 32;
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  for (int t = blockIdx.x; t < tiles_c * tiles_r; t += NC) {
    const int tr = t / tiles_c, tc = t % tiles_c;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = tr * 32 + ly + 8 * i, cc = tc * 32 + lx;
      float v = 0.f;
      if (r < R && cc < C) {
        const int rr = map ? __ldg(map + r) : r;
        v = __ldcg(src + (size_t)rr * lds + cc);
      }
      sT[(ly + 8 * i) * 33 + lx] = fmaf(v, fa, fc);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int cc = tc * 32 + ly + 8 * i, r = tr * 32 + lx;
      if (r < R && cc < C) __stcg(dst + (size_t)cc * ldd + r, sT[lx * 33 + ly + 8 * i]);
    }
  }
}

// logdet[m] = sum over stages / tiles of the coupling partials + sum_s (+-) D * log_scale_inv_s  (CTA 0)
__device__ __noinline__ void run_logdet_final(const SArgs& A, const Sm& s) {
  const KArgs& a = A.k;
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int n = A.sp.NC;              // one running partial per CTA
  const int LB = a.ldb;
  const float* ldp = a.W + a.oLDP;
  float lsum = 0.f;
  for (int e = 0; e < a.S; ++e) lsum += __ldg(a.P + s.st[e].lsi);
  for (int mb = 0; mb < a.B; mb += 32) {
    const int m = mb + lane;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (m < a.B) {
      const int per = (n + 7) / 8, lo = slice * per, hi = min(n, lo + per);
      int i = lo;
      for (; i + 8 <= hi; i += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = ldcg(ldp + (size_t)(i + u) * LB + m);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u & 3] += v[u];
      }
      for (; i < hi; ++i) acc[0] += ldcg(ldp + (size_t)i * LB + m);
    }
    __syncthreads();
    s.red[slice * 32 + lane] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    __syncthreads();
    if (slice == 0 && m < a.B) {
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) t += s.red[r * 32 + lane];
      // a grid barrier that timed out (sticky flag ctrl[2]) must not pass silently: the outputs become NaN
      a.logdet[m] = t + (a.dir == 0 ? (float)a.D * lsum : -(float)a.D * lsum) + (__ldcg(a.ctrl + 2) ? __int_as_float(0x7fc00000) : 0.f);
    }
  }
}

// ------------------------------------------------------------------------------- backward-only phases
// Elementwise backward of ActNorm.reverse and the coupling tail, one warp per coupling column, a lane
// owns samples lane and lane + 32. Produces g_pre^T, the gradients of net.6.scale /
// net.6.fc.bias, the b-half rows of the gradient wrt the stage input and ActNorm partial sums.
__device__ __forceinline__ void run_b0(const SArgs& A, const Sm& s, int e) {
  const KArgs& a = A.k;
  Cx c(a, s.st, e);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const float lsi = __ldg(a.P + c.st.lsi), loc = __ldg(a.P + c.st.loc);
  const float e_lsi = exp_nf(lsi);
  const size_t LB = a.ldb;
#pragma unroll 1
  for (int t = blockIdx.x; t < A.sp.red_tiles; t += A.sp.NC) {
    const int col = t * 8 + w;
    float S1 = 0.f, S2 = 0.f;
    if (col < a.Db) {
      const float E_ls = exp_nf(3.f * __ldg(c.P(P_SCALE) + col));
      const float E_t = exp_nf(3.f * __ldg(c.P(P_SCALE) + a.Db + col));
      const int gcol = __ldg(c.gmap + a.Da + col);
      const float* gya = c.gy() + (size_t)col * LB;
      const float* gyb = c.gy() + (size_t)(a.Da + col) * LB;
      const float* ya = c.y() + (size_t)col * LB;
      const float* yb = c.y() + (size_t)(a.Da + col) * LB;
      const float* ols = c.O() + (size_t)col * LB;
      const float* ot = c.O() + (size_t)(a.Db + col) * LB;
      const float* bp = c.yprev() + (size_t)gcol * LB;
      float* gp_ls = a.W + a.oGPRET + (size_t)col * LB;
      float* gp_t = a.W + a.oGPRET + (size_t)(a.Db + col) * LB;
      float* dst = c.gyprev() + (size_t)gcol * LB;
      float v_ga[2], v_gb[2], v_ya[2], v_yb[2], v_ols[2], v_ot[2], v_b[2], v_gld[2];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {          // all loads of both samples in flight together
        const int m = lane + 32 * hh;
        const bool ok = m < a.B;
        v_ga[hh] = ok ? __ldcg(gya + m) : 0.f; v_gb[hh] = ok ? __ldcg(gyb + m) : 0.f;
        v_ya[hh] = ok ? __ldcg(ya + m) : 0.f;  v_yb[hh] = ok ? __ldcg(yb + m) : 0.f;
        v_ols[hh] = ok ? __ldcg(ols + m) : 0.f; v_ot[hh] = ok ? __ldcg(ot + m) : 0.f;
        v_b[hh] = ok ? __ldcg(bp + m) : 0.f;   v_gld[hh] = ok ? __ldcg(a.g_ld + m) : 0.f;
      }
      float sc_ls = 0.f, sc_t = 0.f, b3_ls = 0.f, b3_t = 0.f;
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        const int m = lane + 32 * hh;
        if (m < a.B) {
          const float ga_ = v_ga[hh], gb_ = v_gb[hh];
          S1 += ga_ + gb_;
          S2 += ga_ * (v_ya[hh] + loc) + gb_ * (v_yb[hh] + loc);
          const float gbp = gb_ * e_lsi;
          const float o_ls = v_ols[hh], o_t = v_ot[hh];
          const float ls = tanh_nf(o_ls);
          const float ei = exp_nf(-ls);
          const float g_t = -gbp;
          const float g_ls = -gbp * v_b[hh] * ei - v_gld[hh];
          const float g_ols = g_ls * (1.f - ls * ls);
          const float v_ls = g_ols * E_ls, v_t = g_t * E_t;
          __stcg(gp_ls + m, v_ls);
          __stcg(gp_t + m, v_t);
          sc_ls += 3.f * g_ols * o_ls;
          sc_t += 3.f * g_t * o_t;
          b3_ls += v_ls;
          b3_t += v_t;
          __stcg(dst + m, gbp * ei);
        }
      }
      sc_ls = warp_sum(sc_ls); sc_t = warp_sum(sc_t); b3_ls = warp_sum(b3_ls); b3_t = warp_sum(b3_t);
      if (lane == 0) {
        c.G(P_SCALE)[col] = sc_ls; c.G(P_SCALE)[a.Db + col] = sc_t;
        c.G(P_B3)[col] = b3_ls;    c.G(P_B3)[a.Db + col] = b3_t;
      }
    }
    if (a.Da > a.Db && t == 0) {   // odd D: the last a-row contributes to the ActNorm sums
      const float* gq = c.gy() + (size_t)(a.Da - 1) * LB;
      const float* yv = c.y() + (size_t)(a.Da - 1) * LB;
      for (int m = threadIdx.x; m < a.B; m += kT) {
        const float gv = __ldcg(gq + m);
        S1 += gv;
        S2 += gv * (__ldcg(yv + m) + loc);
      }
    }
    const float t1 = block_sum(S1, s.red);
    const float t2 = block_sum(S2, s.red + 64);
    if (threadIdx.x == 0) {
      float* red = a.W + a.oRED + ((size_t)e * A.sp.red_tiles + t) * 2;
      __stcg(red, t1);
      __stcg(red + 1, t2);
    }
    __syncthreads();
  }
}

// Weight gradients gW[i][j] = sum_m g^T[i][m] * act^T[j][m]  (which = 3: fc.weight, 2: net.3, 1: net.0), computed
// in the shadow of the grid barrier. Tile TI x TJ, thread tile 4 x 4 (rows 4 iq + a, columns jq + (TJ/4) b).
template <int TI, int TJ, int MS>
__device__ __noinline__ void wgrad_tiles(const SArgs& A, const Sm& s, int e, int which, int busy) {
  const KArgs& a = A.k;
  Cx c(a, s.st, e);
  const int I = which == 3 ? a.Dout : a.H;
  const int J = which == 1 ? a.Da : a.H;
  const float* act = which == 1 ? c.yprev() : c.Nrm(which - 1);
  const int* amap = which == 1 ? c.gmap : nullptr;
  const float* gsrc = a.W + (which == 3 ? a.oGPRET : which == 2 ? a.oGP2T : a.oGP1T);
  float* gw = c.G(which == 3 ? P_W3 : which == 2 ? P_W2 : P_W1);
  const int tiles_j = (J + TJ - 1) / TJ, tiles_i = (I + TI - 1) / TI;
  constexpr int NJQ = TJ / 4;
  float* G_s = s.X;
  float* A_s = s.X + TI * kLdG;
  float* red = s.X + (TI + TJ) * kLdG;      // MS == 4: [4][16][64] floats
  int* jidx = reinterpret_cast<int*>(red);  // which == 1: gathered row indices (dead before `red` is written)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int jq, iq, mp;
  if (MS == 4) { jq = lane & 15; iq = (lane >> 4) + 2 * (warp & 1); mp = warp >> 1; }
  else { jq = lane; iq = warp; mp = 0; }
  // The critical part of a phase deals its tiles from CTA 0 upwards and keeps `busy` CTAs occupied. If the other
  // CTAs can absorb all weight-gradient tiles with at most three each they take all of them (they run concurrently
  // with the critical part); otherwise the tiles are dealt over all CTAs from the last one downwards.
  const int n_tiles = tiles_i * tiles_j, idle = A.sp.NC - busy;
  int t0, tstep;
  if (idle > 0 && n_tiles <= 3 * idle) {
    t0 = (int)blockIdx.x >= busy ? (int)blockIdx.x - busy : n_tiles;
    tstep = idle;
  } else {
    t0 = A.sp.NC - 1 - (int)blockIdx.x;
    tstep = A.sp.NC;
  }
#pragma unroll 1
  for (int t = t0; t < n_tiles; t += tstep) {
    const int ti = t / tiles_j, tj = t - ti * tiles_j;
    const int i0 = ti * TI, j0 = tj * TJ;
    __syncthreads();   // previous users of the X region are done
    if (amap) {
      for (int r = threadIdx.x; r < TJ; r += kT) jidx[r] = j0 + r < J ? __ldg(amap + j0 + r) : 0;
      __syncthreads();
    }
    load_rows(G_s, kLdG, gsrc, nullptr, 0, i0, I, 0, a.ldb, a.B, TI);
    load_rows(A_s, kLdG, act, jidx, amap ? 1 : 0, j0, J, 0, a.ldb, a.B, TJ);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y) acc[x][y] = 0.f;
    constexpr int MW = kMB / MS;
#pragma unroll 1
    for (int m = mp * MW; m < (mp + 1) * MW; m += 4) {
      float4 g[4], v[4];
#pragma unroll
      for (int x = 0; x < 4; ++x) g[x] = *reinterpret_cast<const float4*>(G_s + (4 * iq + x) * kLdG + m);
#pragma unroll
      for (int y = 0; y < 4; ++y) v[y] = *reinterpret_cast<const float4*>(A_s + (jq + NJQ * y) * kLdG + m);
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) {
          acc[x][y] = fmaf(g[x].x, v[y].x, acc[x][y]);
          acc[x][y] = fmaf(g[x].y, v[y].y, acc[x][y]);
          acc[x][y] = fmaf(g[x].z, v[y].z, acc[x][y]);
          acc[x][y] = fmaf(g[x].w, v[y].w, acc[x][y]);
        }
    }
    if (MS == 4) {
      const int ij = iq * 16 + jq;
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) red[((mp * 16) + x * 4 + y) * 64 + ij] = acc[x][y];
      __syncthreads();
#pragma unroll 1
      for (int o = threadIdx.x; o < 1024; o += kT) {
        const int xy = o >> 6, ij2 = o & 63;
        const float t4 = ((red[(0 * 16 + xy) * 64 + ij2] + red[(1 * 16 + xy) * 64 + ij2]) +
                          red[(2 * 16 + xy) * 64 + ij2]) + red[(3 * 16 + xy) * 64 + ij2];
        const int i = i0 + 4 * (ij2 >> 4) + (xy >> 2), j = j0 + (ij2 & 15) + NJQ * (xy & 3);
        if (i < I && j < J) gw[(size_t)i * J + j] = t4;
      }
    } else {
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) {
          const int i = i0 + 4 * iq + x, j = j0 + jq + NJQ * y;
          if (i < I && j < J) gw[(size_t)i * J + j] = acc[x][y];
        }
    }
  }
}

// ActNorm.reverse gradients: g_loc = -sum g_y, g_lsi = sum g_y * (y + loc) + D * sum_m g_ld[m]  (CTA 0)
__device__ __noinline__ void run_bwd_final(const SArgs& A, const Sm& s) {
  const KArgs& a = A.k;
  float sacc = 0.f;
  for (int m = threadIdx.x; m < a.B; m += kT) sacc += ldcg(a.g_ld + m);
  const float gld = block_sum(sacc, s.red);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int e = a.fin_lo + w; e < a.fin_hi; e += kT / 32) {      // the stages this launch covered
    const float* red = a.W + a.oRED + (size_t)e * A.sp.red_tiles * 2;
    float s1 = 0.f, s2 = 0.f;
    for (int t = lane; t < A.sp.red_tiles; t += 32) { s1 += ldcg(red + 2 * t); s2 += ldcg(red + 2 * t + 1); }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
      const StageTab& st = s.st[e];
      const float bad = (__ldcg(a.ctrl + 2) ? __int_as_float(0x7fc00000) : 0.f);   // timed-out grid barrier -> NaN gradients
      a.G[st.loc] = -s1 + bad;
      a.G[st.lsi] = s2 + (float)a.D * gld + bad;
    }
  }
}

// ------------------------------------------------------------------------------- the phase sequence
__device__ __forceinline__ void decode(const SArgs& A, int p, int& kind, int& e) {
  const int S = A.k.S;
  if (!A.bwd) {
    if (p == 0) { kind = K_TIN; e = 0; }
    else if (p <= 3 * S) { e = (p - 1) / 3; kind = K_F1 + (p - 1) % 3; }
    else { kind = K_FFIN; e = S - 1; }
  } else {
    if (p == A.k.bfin_phase) { kind = K_BFIN; e = 0; }       // last phase of the launch (stage-range launches end early)
    else if (p == 0) { kind = K_TIN; e = S - 1; }
    else if (p <= 4 * S) { e = S - 1 - ((p - 1) >> 2); kind = K_B0 + ((p - 1) & 3); }
    else { kind = K_BFIN; e = 0; }
  }
}

}  // namespace

__global__ void __launch_bounds__(kT, 1) flow_small_kernel(const __grid_constant__ SArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float sRed[8 * 64 + 128];
  __shared__ float sLd[64];
  __shared__ __align__(16) StageTab sSt[64];
  Sm s;
  s.X = reinterpret_cast<float*>(smem_raw + A.sp.oX);
  s.W = reinterpret_cast<float*>(smem_raw + A.sp.oW);
  s.X2 = reinterpret_cast<float*>(smem_raw + A.sp.oX2);
  s.W2 = reinterpret_cast<float*>(smem_raw + A.sp.oW2);
  s.E = reinterpret_cast<float*>(smem_raw + A.sp.oE);
  s.P = reinterpret_cast<float*>(smem_raw + A.sp.oP);
  s.R = reinterpret_cast<float*>(smem_raw + A.sp.oR);
  s.Pm = reinterpret_cast<float*>(smem_raw + A.sp.oPm);
  s.idx = reinterpret_cast<int*>(smem_raw + A.sp.oIdx);
  s.idx2 = reinterpret_cast<int*>(smem_raw + A.sp.oIdx2);
  s.red = sRed;
  s.ldacc = sLd;
  if (threadIdx.x < 64) sLd[threadIdx.x] = 0.f;
  s.st = A.k.st;
  s.phase = A.k.phase_begin;
  if (A.k.S <= 64) {
    const int4* src = reinterpret_cast<const int4*>(A.k.st);
    int4* dst = reinterpret_cast<int4*>(sSt);
    for (int i = threadIdx.x; i < A.k.S * (int)(sizeof(StageTab) / 16); i += kT) dst[i] = __ldg(src + i);
    s.st = sSt;
  }
  __syncthreads();
  const KArgs& a = A.k;
  bool cl_pending = false;
  unsigned bar = 0;
  const bool persistent = a.persistent != 0;
  int kind = K_TIN, e = 0;
  // Phase descriptors depend only on (kind, stage): all of them are built once, in parallel, into shared memory
  // so that the per-phase critical path only reads them.
  PD* pd = reinterpret_cast<PD*>(smem_raw + A.sp.oPD);
  for (int p = a.phase_begin + (int)threadIdx.x; p < a.phase_end; p += kT) {
    int k2, e2;
    decode(A, p, k2, e2);
    pd[p - a.phase_begin] = make_desc(A, s.st, k2, e2);
  }
  __syncthreads();
  // iteration p = phase_begin - 1 is a virtual phase that only prefetches for the first one, so that
  // prefetch_gemm has a single call site (instruction-cache footprint)
#pragma unroll 1
  for (int p = a.phase_begin - 1; p < a.phase_end; ++p) {
    const bool real = p >= a.phase_begin;
    if (real) {
      const PD& d = pd[p - a.phase_begin];
      s.phase = p;
      if (persistent && p > a.phase_begin) grid_wait(a.ctrl, (++bar) * gridDim.x);
      STAMP(0);
      if (a.tbuf && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        a.tbuf[2 * p] = t;
      }
      kind = d.kind; e = d.e;
      if (kind == K_B0) {
        run_b0(A, s, e);
      } else if (kind == K_TIN) {
        if (!A.bwd && threadIdx.x < 64) __stcg(a.W + a.oLDP + (size_t)blockIdx.x * a.ldb + threadIdx.x, 0.f);
        run_transpose(a, s.st, A.bwd ? 1 : 0, s.X, A.sp.NC);
      } else if (kind == K_FFIN) {
        if (blockIdx.x == 0) run_logdet_final(A, s);
        __syncthreads();
        run_transpose(a, s.st, a.dir == 0 ? 2 : 3, s.X, A.sp.NC);
      } else if (kind == K_BFIN) {
        if (blockIdx.x == 0) run_bwd_final(A, s);
        __syncthreads();
        if (a.g_in) run_transpose(a, s.st, 4, s.X, A.sp.NC);
      } else {
        run_gemm(A, s, d, cl_pending);
      }
      STAMP(5);
      if (a.tbuf && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        a.tbuf[2 * p + 1] = t;
      }
    }
    if (p + 1 < a.phase_end) {
      if (real) {
        if (persistent) grid_arrive(a.ctrl); else __syncthreads();
      }
      prefetch_gemm(A, s, pd[p + 1 - a.phase_begin]);   // next hop's weights / indices / parameters -> shared memory
    }
    if (real && (kind == K_F1 || kind == K_B0)) {
      // L2 prefetch of the parameter block of the stage that runs next (parameters of a stage are contiguous in the
      // flat buffer, net.0.weight .. net.6.fc.bias): its weights are then L2 hits when the shadow prefetch of the
      // coming phases copies them into shared memory. One 128-byte line per thread and step.
      const int en = kind == K_F1 ? e + 1 : e - 1;
      if (en >= 0 && en < a.S) {
        const StageTab& sn = s.st[a.dir ? a.S - 1 - en : en];
        const char* lo = reinterpret_cast<const char*>(a.P + sn.p[P_W1]);
        const long long bytes = (sn.p[P_B3] + a.Dout - sn.p[P_W1]) * 4;
        for (long long off = ((long long)blockIdx.x * kT + threadIdx.x) * 128; off < bytes; off += (long long)gridDim.x * kT * 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(lo + off));
      }
    }
    if (real) {
      STAMP(6);
      // shadow of the barrier: weight gradients of this stage (inputs complete since the previous barrier)
      // net.6.fc.weight with DG2, net.3.weight with DG1; net.0.weight of stage e + 1 rides with B0 of stage e
      // (B0 keeps only Db / 8 CTAs busy; DGIN keeps all of them busy), the last stage of the launch with its own DGIN
      // (stage 0, or the lower end of a stage range: every gradient of the range is final when its launch ends)
      int which = 0, we = e;
      if (kind == K_DG2) which = 3;
      else if (kind == K_DG1) which = 2;
      else if (kind == K_B0 && e + 1 < a.fin_hi) { which = 1; we = e + 1; }
      else if (kind == K_DGIN && e == a.fin_lo) which = 1;
      if (which) {
        // CTAs occupied by this phase's critical part
        const PD& dc = pd[p - a.phase_begin];
        int busy;
        if (kind == K_B0) busy = min(A.sp.red_tiles, A.sp.NC);
        else busy = dc.split ? min(dc.n_tiles, A.sp.G) * A.sp.cl : min(dc.n_tiles, A.sp.NC);
        if (A.sp.wgv == 0) wgrad_tiles<16, 64, 4>(A, s, we, which, busy);
        else wgrad_tiles<32, 128, 1>(A, s, we, which, busy);
      }
      STAMP(7);
    }
  }
  cp_async_wait<0>();
  if (cl_pending) cluster_wait();
  if (A.sp.cl > 1) {   // no CTA of a cluster may exit while a peer can still read its shared memory
    cluster_arrive();
    cluster_wait();
  }
}

// =================================================================================================
// host side
// =================================================================================================
namespace {

inline int up4(int x) { return (x + 3) & ~3; }
constexpr int kSmallSmemMax = 214 * 1024;   // dynamic shared memory ceiling (232,448 B per block minus ~14 KB static)

ActPlan plan_act(int N, int K, bool split, int G, int NC, int cl, bool wk, int wmax) {
  ActPlan p{};
  p.N = N; p.K = K;
  p.ksplit = split ? 1 : 0;
  const int units = split ? G : NC;
  int nt = 4;
  while (nt < 64 && (N + nt - 1) / nt > units) nt *= 2;
  p.NT = nt;
  p.n_tiles = (N + nt - 1) / nt;
  const int nr = split ? cl : 1;
  p.KC = up4((K + nr - 1) / nr);
  p.KS = std::min(p.KC, 512);
  auto wbytes = [&](int ks) { return wk ? nt * (ks + 4) * 4 : ks * (nt + 4) * 4; };
  while (p.KS > 4 && wbytes(p.KS) > wmax) p.KS = up4(p.KS / 2);
  p.nq = nt / 4;
  p.kp = 16 / p.nq;
  p.nt_log = 0;
  while ((1 << p.nt_log) < nt) ++p.nt_log;
  p.nq_log = p.nt_log - 2;
  return p;
}

}  // namespace

int small_query(dpi_flow_s* h) {
  h->small_clusters = 0;
  h->small_cl = 1;
  const char* env = getenv("DPI_B200_SMALL");
  h->small_mode = (env && env[0] == '0') ? 0 : 1;
  const char* ks = getenv("DPI_B200_SMALL_KSPLIT_MIN");
  h->small_ksplit_min = ks ? std::max(8, atoi(ks)) : 512;
  if (!h->small_mode) return DPI_OK;
  const int max_smem = kSmallSmemMax;
  if (cudaFuncSetAttribute((const void*)flow_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) != cudaSuccess) {
    cudaGetLastError();
    h->small_mode = 0;
    return DPI_OK;
  }
  const char* nocl = getenv("DPI_B200_SMALL_NOCLUSTER");
  for (int i = 0; i < 3; ++i) h->small_G[i] = 0;
  if (!(nocl && nocl[0] == '1')) {
    const char* want = getenv("DPI_B200_SMALL_CL");    // force one cluster size (8, 4 or 2)
    for (int i = 0; i < 3; ++i) {
      const int cl = 8 >> i;
      if (want && atoi(want) != cl) continue;
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(cl * (h->n_sm / cl)); cfg.blockDim = dim3(kT); cfg.dynamicSmemBytes = max_smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, (const void*)flow_small_kernel, &cfg) == cudaSuccess && nc >= 2) {
        h->small_G[i] = std::min(nc, h->n_sm / cl);
        h->small_cl = 8;      // "cluster launch available"
      } else {
        cudaGetLastError();
      }
    }
  }
  h->small_clusters = h->n_sm;   // without clusters: plain cooperative launch, one CTA per SM
  return DPI_OK;
}

namespace {
// FMA on the busiest CTA of a hop (rounds x features per tile x K extent per rank), the quantity the hop's
// critical path scales with
long long hop_cost(const ActPlan& p, int units) {
  const int rounds = (p.n_tiles + units - 1) / units;
  return (long long)rounds * p.NT * p.KC;
}
void plan_hops(const dpi_flow_s* h, SmallPlan& sp, int cl, int G, int wmax) {
  const int NC = cl > 1 ? cl * G : G;
  sp.NC = NC; sp.G = cl > 1 ? G : NC; sp.cl = cl;
  const int H = h->H, Da = h->Da, Db = h->Db, Dout = h->Dout;
  auto split = [&](int K) { return cl > 1 && K >= h->small_ksplit_min; };
  sp.f1 = plan_act(H, Da, split(Da), sp.G, NC, cl, true, wmax);
  sp.f2 = plan_act(H, H, split(H), sp.G, NC, cl, true, wmax);
  {  // output layer: tiles of NP column pairs -> 2 NP weight rows, never split (N large, K = H small)
    int np = 2;
    while (np < 32 && (Db + np - 1) / np > NC) np *= 2;
    ActPlan p = plan_act(2 * np, H, false, sp.G, NC, cl, true, wmax);
    p.NT = 2 * np; p.nq = p.NT / 4; p.kp = 16 / p.nq;
    p.nt_log = 0;
    while ((1 << p.nt_log) < p.NT) ++p.nt_log;
    p.nq_log = p.nt_log - 2;
    p.N = Db; p.n_tiles = (Db + np - 1) / np;
    p.KS = std::min(p.KC, 512);
    while (p.KS > 4 && p.NT * (p.KS + 4) * 4 > wmax) p.KS = up4(p.KS / 2);
    sp.f3 = p;
  }
  sp.dg2 = plan_act(H, Dout, split(Dout), sp.G, NC, cl, false, wmax);
  sp.dg1 = plan_act(H, H, split(H), sp.G, NC, cl, false, wmax);
  sp.dgin = plan_act(Da, H, false, sp.G, NC, cl, false, wmax);
}
long long plan_cost(const SmallPlan& sp) {
  auto u = [&](const ActPlan& p) { return p.ksplit ? sp.G : sp.NC; };
  return hop_cost(sp.f1, u(sp.f1)) + hop_cost(sp.f2, u(sp.f2)) + hop_cost(sp.f3, u(sp.f3)) + hop_cost(sp.dg2, u(sp.dg2)) +
         hop_cost(sp.dg1, u(sp.dg1)) + hop_cost(sp.dgin, u(sp.dgin));
}
}  // namespace

namespace {
// One pass's plan. `pref`: cluster-size preference order as indices into small_G (0: 8 CTAs, 1: 4, 2: 2).
void small_plan_pass_w(const dpi_flow_s* h, const BatchPlan& bp, SmallPlan& sp, const int (&pref)[3], int wmax, bool allow_ring) {
  sp = SmallPlan{};
  if (!h->small_mode || bp.B > kMB || h->H > 512 || (h->H & 3) || (h->Da & 3)) return;
  const int H = h->H, Dout = h->Dout, Db = h->Db;
  long long best = -1;
  for (int q = 0; q < 3 && best < 0; ++q) {
    const int i = pref[q];
    if (h->small_G[i] <= 0) continue;
    plan_hops(h, sp, 8 >> i, h->small_G[i], wmax);
    best = plan_cost(sp);
  }
  if (best < 0) plan_hops(h, sp, 1, h->small_clusters, wmax);   // no cluster launch: single CTAs, K never split
  if (getenv("DPI_B200_SMALL_VERBOSE")) {
    const ActPlan* all[6] = {&sp.f1, &sp.f2, &sp.f3, &sp.dg2, &sp.dg1, &sp.dgin};
    const char* nm[6] = {"f1", "f2", "f3", "dg2", "dg1", "dgin"};
    fprintf(stderr, "dpi_b200 small plan: B=%d cluster=%d G=%d NC=%d (available: 8x%d 4x%d 2x%d) cost=%lld\n", bp.B, sp.cl,
            sp.G, sp.NC, h->small_G[0], h->small_G[1], h->small_G[2], best);
    for (int i = 0; i < 6; ++i)
      fprintf(stderr, "   %-4s N=%d K=%d NT=%d tiles=%d split=%d KC=%d KS=%d\n", nm[i], all[i]->N, all[i]->K, all[i]->NT,
              all[i]->n_tiles, all[i]->ksplit, all[i]->KC, all[i]->KS);
  }
  const int NC = sp.NC;
  const long long wg3 = ((long long)Dout + 15) / 16 * ((H + 63) / 64);
  sp.wgv = wg3 > 4LL * NC ? 1 : 0;
  if (const char* e = getenv("DPI_B200_SMALL_WGV")) sp.wgv = atoi(e) ? 1 : 0;   // A/B switch for the tile variant
  sp.ldp_tiles = std::max(sp.f3.n_tiles, (NC + h->S - 1) / h->S);   // workspace rows: one logdet partial per CTA
  sp.red_tiles = (Db + 7) / 8;
  const ActPlan* all[6] = {&sp.f1, &sp.f2, &sp.f3, &sp.dg2, &sp.dg1, &sp.dgin};
  int xb = 0, wb = 0, ib = 0;
  bool multi = false;
  int xb_gemm = 0;
  for (int i = 0; i < 6; ++i) {
    const ActPlan& p = *all[i];
    const bool wk = i < 3;
    xb = std::max(xb, p.KS * kMB * 4);
    xb_gemm = xb;
    wb = std::max(wb, wk ? p.NT * (p.KS + 4) * 4 : p.KS * (p.NT + 4) * 4);
    ib = std::max(ib, (p.KC + 4) * 4);
    if (p.KS < p.KC && allow_ring) multi = true;
  }
  const int wg_bytes = sp.wgv == 0 ? (16 + 64) * kLdG * 4 + 16 * 1024 : (32 + 128) * kLdG * 4 + 1024;
  xb = std::max(xb, std::max(wg_bytes, 33 * 32 * 4));
  auto al = [](int x) { return (x + 127) & ~127; };
  int o = 0;
  sp.oX = o; o += al(xb);
  sp.oW = o; o += al(wb);
  sp.oX2 = sp.oX; sp.oW2 = sp.oW;
  if (multi) {            // second halves of the K-chunk ring
    sp.oX2 = o; o += al(xb_gemm);
    sp.oW2 = o; o += al(wb);
  }
  sp.oE = o; o += 64 * kMB * 4;
  sp.oP = o; o += 16 * 4 * kMB * 4;
  sp.oR = o; o += 64 * kMB * 4;
  sp.oPm = o; o += 5 * 64 * 4;
  sp.oIdx = o; o += al(ib);
  sp.oIdx2 = o; o += 128 * 4;
  const int n_ph = 4 * h->S + 2;                       // backward pass has the most phases
  sp.oPD = o; o += al(n_ph * (int)sizeof(PD));
  sp.smem_bytes = o;
  sp.ok = o <= kSmallSmemMax ? 1 : 0;
}
// weight slices up to 70 KB keep the K loop of the MRI-sized layers at two chunks; smaller budget when that overflows
void small_plan_pass(const dpi_flow_s* h, const BatchPlan& bp, SmallPlan& sp, const int (&pref)[3]) {
  small_plan_pass_w(h, bp, sp, pref, 70 * 1024, true);
  if (!sp.ok) small_plan_pass_w(h, bp, sp, pref, 36 * 1024, true);
  if (!sp.ok) small_plan_pass_w(h, bp, sp, pref, 36 * 1024, false);
}
}  // namespace

// Cluster size per pass (DPI_B200_SMALL_CL forces one size for both). Measured on C1 / C2 (tools/cl_sweep.sh): the
// backward pass is fastest with 15 clusters of 8 - the CTAs a hop leaves idle absorb the weight-gradient tiles -
// while the forward pass, which has no such side work, is fastest with 33 clusters of 4 (more CTAs busy per hop).
void small_plan(const dpi_flow_s* h, BatchPlan& bp) {
  // ... as long as the hops are latency-bound (C1: 14 MFMA per stage); at MRI sizes (C2: 218 MFMA per stage) the
  // hops are FFMA-bound and 8 x 15 wins for both passes.
  const long long stage_fma = 64LL * ((long long)h->Da * h->H + (long long)h->H * h->H + (long long)h->H * h->Dout);
  const int pref_bwd[3] = {0, 1, 2}, pref_fwd4[3] = {1, 0, 2};
  small_plan_pass(h, bp, bp.small, pref_bwd);
  small_plan_pass(h, bp, bp.small_fwd, stage_fma < 50000000LL ? pref_fwd4 : pref_bwd);
  if (!bp.small.ok || !bp.small_fwd.ok) { bp.small.ok = 0; bp.small_fwd.ok = 0; }
  // the logdet partial rows (one per forward CTA) live in the workspace sized from bp.small
  bp.small.ldp_tiles = std::max(bp.small.ldp_tiles, bp.small_fwd.ldp_tiles);
}

int small_launch(dpi_flow_s* h, const BatchPlan& bp, KArgs& a, int bwd, cudaStream_t stream, int lo, int hi) {
  const SmallPlan& plan = bwd ? bp.small : bp.small_fwd;
  SArgs A{};
  A.k = a;
  A.sp = plan;
  A.bwd = bwd;
  A.k.phases = nullptr;
  A.k.use_tc = 0;
  A.k.tbuf = nullptr;
  A.k.tiles_m = 1;
  A.k.ldp_tiles = plan.ldp_tiles;
  A.k.red_tiles = plan.red_tiles;
  const int n_phases = bwd ? 4 * h->S + 2 : 3 * h->S + 2;
  // Backward over the stages [lo, hi) only (hi - 1 first): phases [p0, p1) of the full sequence, then the finalising phase
  // for exactly these stages (the input-gradient transpose belongs to the range that ends at stage 0). State between
  // ranges lives in the workspace, as between the launches of the one-launch-per-phase mode.
  const bool ranged = bwd && hi >= 0 && !(lo == 0 && hi == h->S);
  int p0 = 0, p1 = n_phases;
  A.k.fin_lo = 0; A.k.fin_hi = h->S; A.k.bfin_phase = bwd ? 4 * h->S + 1 : -1;
  if (ranged) {
    p0 = hi == h->S ? 0 : 1 + 4 * (h->S - hi);
    p1 = 1 + 4 * (h->S - lo) + 1;
    A.k.fin_lo = lo; A.k.fin_hi = hi; A.k.bfin_phase = p1 - 1;
    if (lo > 0) A.k.g_in = nullptr;
  }
  if (h->d_tbuf && n_phases <= 4096 && !ranged) {
    A.k.tbuf = h->d_tbuf;
    h->tbuf_phases = n_phases;
    h->last_ops.assign(n_phases, 0);
    for (int p = 0; p < n_phases; ++p) {
      int kind;
      if (!bwd) kind = p == 0 ? K_TIN : (p <= 3 * h->S ? K_F1 + (p - 1) % 3 : K_FFIN);
      else kind = p == 0 ? K_TIN : (p <= 4 * h->S ? K_B0 + (p - 1) % 4 : K_BFIN);
      h->last_ops[p] = (20 + kind) * 100;
    }
  }
  DPI_CUDA(cudaMemsetAsync(A.k.ctrl, 0, kCtrlWords * sizeof(unsigned), stream));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(plan.NC);
  cfg.blockDim = dim3(kT);
  cfg.dynamicSmemBytes = plan.smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (plan.cl > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = plan.cl; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
    ++na;
  }
  // DPI_B200_SMALL_COOP=0: plain (non-cooperative) launch of the persistent kernel. The grid never exceeds the
  // co-resident capacity, so this is safe on an otherwise idle GPU; it exists because Nsight Compute cannot
  // launch a cooperative kernel with a cluster dimension ("LaunchFailed").
  static const bool coop = []() { const char* e = getenv("DPI_B200_SMALL_COOP"); return !(e && e[0] == '0'); }();
  if (h->persistent && coop) {
    at[na].id = cudaLaunchAttributeCooperative;
    at[na].val.cooperative = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  if (h->persistent) {
    A.k.phase_begin = p0; A.k.phase_end = p1; A.k.persistent = 1;
    DPI_CUDA(cudaLaunchKernelEx(&cfg, flow_small_kernel, A));
    count_launch();
  } else {
    A.k.persistent = 0;
    for (int p = p0; p < p1; ++p) {
      A.k.phase_begin = p; A.k.phase_end = p + 1;
      DPI_CUDA(cudaLaunchKernelEx(&cfg, flow_small_kernel, A));
      count_launch();
    }
  }
  return DPI_OK;
}

}  // namespace dpi
