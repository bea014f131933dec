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
  int *idx, *idx2;
  float* red;             // 8 * 64 + 128 floats (static)
  const StageTab* st;     // stage table (staged into shared memory when it fits)
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
// loads, bounded (~2 s) so a protocol bug sets the error flag ctrl[2] instead of hanging the GPU.
__device__ __forceinline__ void grid_arrive(unsigned* ctrl) {
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctrl), "r"(1u) : "memory");
}
__device__ __forceinline__ void grid_wait(unsigned* ctrl, unsigned target) {
  if (threadIdx.x == 0) {
    unsigned cur;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cur) : "l"(ctrl) : "memory");
    if (cur < target) {
      const long long t0 = clock64();
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cur) : "l"(ctrl) : "memory");
        if (cur < target && (clock64() - t0 > 400000000LL || ld_flag(ctrl + 2))) {
          atomicExch(ctrl + 2, 1u);   // sticky: every later wait of every CTA gives up at once
          break;
        }
      } while (cur < target);
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------- addressing (as flow_kernels.cu Ctx)
struct Cx {
  const KArgs& a;
  int e;
  const StageTab* stb;
  const StageTab& st;
  const int* gmap;
  bool vec;
  __device__ Cx(const KArgs& a_, const StageTab* stb_, int e_)
      : a(a_), e(e_), stb(stb_), st(stb_[a_.dir ? a_.S - 1 - e_ : e_]),
        gmap(a_.gmaps + ((size_t)a_.dir * a_.S + e_) * a_.D), vec((a_.B & 3) == 0) {}
  __device__ __forceinline__ const float* yprev() const { return e == 0 ? a.W + a.oZT : a.W + a.oYT + (size_t)(e - 1) * a.D * a.B; }
  __device__ __forceinline__ float* y() const { return a.W + a.oYT + (size_t)e * a.D * a.B; }
  __device__ __forceinline__ float* L(int which) const { return a.W + (which == 1 ? a.oL1T : a.oL2T) + (size_t)e * a.H * a.B; }
  __device__ __forceinline__ float* Nrm(int which) const { return a.W + (which == 1 ? a.oN1T : a.oN2T) + (size_t)e * a.H * a.B; }
  __device__ __forceinline__ float* O() const { return a.W + a.oOT + (size_t)e * a.Dout * a.B; }
  __device__ __forceinline__ float* stats(int which, int what) const { return a.W + a.oST + ((size_t)e * 4 + (which - 1) * 2 + what) * a.H; }
  __device__ __forceinline__ const float* P(int id) const { return a.P + st.p[id]; }
  __device__ __forceinline__ float* G(int id) const { return a.G + st.p[id]; }
  __device__ __forceinline__ const float* gy() const { return a.W + a.oGYT + (size_t)(e & 1) * a.D * a.B; }
  __device__ __forceinline__ float* gyprev() const { return a.W + a.oGYT + (size_t)((e - 1) & 1) * a.D * a.B; }
  // affine applied to the coupling output before it is stored (reverse: this stage's ActNorm.reverse :66;
  // forward: the NEXT stage's ActNorm.forward :45 - scalar, so it commutes with the permutation in between)
  __device__ __forceinline__ void post_affine(float& qa, float& qc) const {
    if (a.dir == 0) {
      qa = expf(__ldg(a.P + st.lsi));
      qc = -__ldg(a.P + st.loc);
    } else if (e + 1 < a.S) {
      const StageTab& nx = stb[a.S - 2 - e];
      const float inv = 1.0f / expf(__ldg(a.P + nx.lsi));
      qa = inv;
      qc = __ldg(a.P + nx.loc) * inv;
    } else {
      qa = 1.f; qc = 0.f;
    }
  }
};

// ------------------------------------------------------------------------------- asynchronous loaders
// dst[r * ld + 0..63] <- row rowfn(r) of a feature-major matrix (B valid samples, zero-filled to 64);
// rowfn(r) < 0 gives a zero row.
template <typename RowFn>
__device__ __forceinline__ void load_rows(float* dst, int ld, const float* base, int B, bool vec, int nrows, RowFn rowfn) {
  for (int c = threadIdx.x; c < nrows * 16; c += kT) {
    const int r = c >> 4, m4 = (c & 15) * 4;
    const long long row = rowfn(r);
    float* d = dst + r * ld + m4;
    const float* s = base + (size_t)(row < 0 ? 0 : row) * B + m4;
    const int left = row < 0 ? 0 : B - m4;
    if (vec) {
      cp_async16(d, left > 0 ? s : base, left >= 4 ? 16 : (left > 0 ? left * 4 : 0));
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool ok = u < left;
        cp_async4(d + u, ok ? s + u : base, ok);
      }
    }
  }
}
// K-contiguous weight rows: W_s[r * ldw + kk] <- Wg[rowfn(r) * K + kb + kk], kk < knp, zero beyond klimit
template <typename RowFn>
__device__ __forceinline__ void load_w_k(float* W_s, int ldw, const float* Wg, int K, int nrows, int kb, int knp, int klimit,
                                         RowFn rowfn) {
  const int cpr = knp >> 2;
  const bool vec = (K & 3) == 0;
  for (int c = threadIdx.x; c < nrows * cpr; c += kT) {
    const int r = c / cpr, k4 = (c - r * cpr) * 4;
    const long long row = rowfn(r);
    const int k = kb + k4;
    const int left = row < 0 ? 0 : klimit - k;
    float* d = W_s + r * ldw + k4;
    const float* s = Wg + (size_t)(row < 0 ? 0 : row) * K + k;
    if (vec) {
      cp_async16(d, left > 0 ? s : Wg, left >= 4 ? 16 : (left > 0 ? left * 4 : 0));
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool ok = u < left;
        cp_async4(d + u, ok ? s + u : Wg, ok);
      }
    }
  }
}
// N-contiguous weights: W_s[kk * ldw + nn] <- Wg[(kb + kk) * ldg + n0 + nn], nn < NT, zero beyond N / klimit
__device__ __forceinline__ void load_w_n(float* W_s, int ldw, const float* Wg, int ldg, int kb, int knp, int klimit, int n0,
                                         int NT, int N) {
  const int cpr = NT >> 2;
  const bool vec = (ldg & 3) == 0;
  for (int c = threadIdx.x; c < knp * cpr; c += kT) {
    const int kk = c / cpr, n4 = (c - kk * cpr) * 4;
    const int k = kb + kk;
    const int left = k < klimit ? N - (n0 + n4) : 0;
    float* d = W_s + kk * ldw + n4;
    const float* s = Wg + (size_t)(k < klimit ? k : 0) * ldg + n0 + n4;
    if (vec) {
      cp_async16(d, left > 0 ? s : Wg, left >= 4 ? 16 : (left > 0 ? left * 4 : 0));
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool ok = u < left;
        cp_async4(d + u, ok ? s + u : Wg, ok);
      }
    }
  }
}
// dst[t] <- src[n0 + t] for t < cnt (zero where n0 + t >= N)
__device__ __forceinline__ void load_vec(float* dst, const float* src, int n0, int cnt, int N) {
  for (int t = threadIdx.x; t < cnt; t += kT) {
    const bool ok = n0 + t < N;
    cp_async4(dst + t, ok ? src + n0 + t : src, ok);
  }
}
__device__ __forceinline__ void load_ivec(int* dst, const int* src, int n0, int cnt, int N) {
  for (int t = threadIdx.x; t < cnt; t += kT) {
    const bool ok = n0 + t < N;
    cp_async4(reinterpret_cast<float*>(dst + t), reinterpret_cast<const float*>(ok ? src + n0 + t : src), ok);
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

// One activation GEMM of a phase.
struct ActDesc {
  const ActPlan* pl;
  const float* Xbase;   // feature-major activation matrix the k rows come from
  const int* xmap;      // optional row map (k -> row), nullptr = identity
  const float* Wg;      // weight matrix
  int wk;               // 1: rows n are K-contiguous (row length = K); 0: rows k are N-contiguous (row length w_ld)
  int w_ld;
  int pair_np, pair_db; // F3: local row r -> (r / np) * db + j0 + r % np, valid while j0 + r % np < db
};

__device__ __forceinline__ void issue_weights(const KArgs& a, const Sm& s, const ActDesc& d, int tile, int kb, int knp, int k_hi) {
  const ActPlan& pl = *d.pl;
  if (d.wk) {
    const int ldw = pl.KS + 4;
    if (d.pair_np) {
      const int np = d.pair_np, db = d.pair_db, j0 = tile * np;
      load_w_k(s.W, ldw, d.Wg, pl.K, pl.NT, kb, knp, k_hi, [=](int r) -> long long {
        const int half = r / np, j = j0 + (r - half * np);
        return j < db ? (long long)half * db + j : -1;
      });
    } else {
      const int n0 = tile * pl.NT, N = pl.N;
      load_w_k(s.W, ldw, d.Wg, pl.K, pl.NT, kb, knp, k_hi, [=](int r) -> long long { return n0 + r < N ? n0 + r : -1; });
    }
  } else {
    load_w_n(s.W, pl.NT + 4, d.Wg, d.w_ld, kb, knp, k_hi, tile * pl.NT, pl.NT, pl.N);
  }
  if (d.xmap) load_ivec(s.idx, d.xmap, kb, knp, k_hi);
}

// Issues (does not wait for) the first chunk of weights and row indices of `tile` for this rank.
__device__ __forceinline__ void prefetch_act(const KArgs& a, const Sm& s, const ActDesc& d, int tile, int rank) {
  const ActPlan& pl = *d.pl;
  const int k_lo = rank * pl.KC, k_hi = min(pl.K, k_lo + pl.KC);
  const int kn = max(0, min(pl.KS, k_hi - k_lo)), knp = (kn + 3) & ~3;
  issue_weights(a, s, d, tile, k_lo, knp, k_hi);
}

// Computes the tile into s.R ([NT][64], this rank's partial when split). On return s.R is visible to
// the CTA (and, when split, to the whole cluster). `extra` issues further activation loads that ride
// in the same cp.async batch as the first chunk.
template <typename Extra>
__device__ __forceinline__ void act_tile(const KArgs& a, const Sm& s, const ActDesc& d, int tile, int rank, bool split,
                                         bool prefetched, bool& cl_pending, Extra extra) {
  const ActPlan& pl = *d.pl;
  const int NT = pl.NT;
  const int k_lo = rank * pl.KC, k_hi = min(pl.K, k_lo + pl.KC);
  const int sq = threadIdx.x & 15, g = threadIdx.x >> 4, fq = g % pl.nq, kp = g / pl.nq;
  const int ldw = d.wk ? pl.KS + 4 : NT + 4;
  const bool vec = (a.B & 3) == 0;
  float acc[4][4];
#pragma unroll
  for (int f = 0; f < 4; ++f)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[f][i] = 0.f;
  bool first = true;
  for (int kb = k_lo; first || kb < k_hi; kb += pl.KS) {
    const int kn = max(0, min(pl.KS, k_hi - kb)), knp = (kn + 3) & ~3;
    if (!(prefetched && first)) {
      __syncthreads();   // earlier readers of W / idx / X are done
      issue_weights(a, s, d, tile, kb, knp, k_hi);
      cp_async_commit();
    }
    cp_async_wait<0>();
    __syncthreads();     // weights + indices landed (and X of the previous chunk is free)
    const int* idx = s.idx;
    const bool mapped = d.xmap != nullptr;
    load_rows(s.X, kMB, d.Xbase, a.B, vec, knp, [=](int r) -> long long {
      const int k = kb + r;
      return k < k_hi ? (mapped ? idx[r] : k) : -1;
    });
    if (first) extra();
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    if (d.wk) act_mac<true>(acc, s.X, s.W, ldw, knp, sq, fq, kp, pl.kp);
    else act_mac<false>(acc, s.X, s.W, ldw, knp, sq, fq, kp, pl.kp);
    first = false;
  }
  // k-part partials -> P_s[kp][n][m]
#pragma unroll
  for (int f = 0; f < 4; ++f)
    *reinterpret_cast<float4*>(s.P + ((size_t)(kp * NT + 4 * fq + f)) * kMB + 4 * sq) =
        make_float4(acc[f][0], acc[f][1], acc[f][2], acc[f][3]);
  if (cl_pending) {      // the peers have finished reading the previous contents of R_s
    cluster_wait();
    cl_pending = false;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NT * 16; i += kT) {
    const int n = i >> 4, m4 = (i & 15) * 4;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = 0; q < pl.kp; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(s.P + ((size_t)(q * NT + n)) * kMB + m4);
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    *reinterpret_cast<float4*>(s.R + n * kMB + m4) = t;
  }
  if (split) {
    cluster_arrive();
    cluster_wait();
  } else {
    __syncthreads();
  }
}

// sum over the ranks of the cluster (fixed order) of R_s[nl][lane], R_s[nl][lane + 32]
__device__ __forceinline__ void combine(const Sm& s, int nl, int lane, int nr, float& s0, float& s1) {
  const float* p = s.R + nl * kMB + lane;
  if (nr == 1) {
    s0 = p[0]; s1 = p[32];
  } else {
    float v0[8], v1[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) { v0[r] = ld_dsmem(p, r); v1[r] = ld_dsmem(p + 32, r); }
    s0 = 0.f; s1 = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) { s0 += v0[r]; s1 += v1[r]; }
  }
}

// ------------------------------------------------------------------------------- phases: forward pass
struct Who {            // which unit of a phase this CTA is
  int unit, units, rank, nr;
  bool split;
};
__device__ __forceinline__ Who who(const SArgs& A, const ActPlan& pl) {
  Who w;
  w.split = pl.ksplit != 0;
  if (w.split) {
    w.unit = blockIdx.x / A.sp.cl; w.units = A.sp.G; w.rank = (int)cluster_rank(); w.nr = A.sp.cl;
  } else {
    w.unit = blockIdx.x; w.units = A.sp.NC; w.rank = 0; w.nr = 1;
  }
  return w;
}

__device__ __forceinline__ ActDesc desc_hid_fwd(const SArgs& A, const Cx& c, int which) {
  ActDesc d;
  d.pl = which == 1 ? &A.sp.f1 : &A.sp.f2;
  d.Xbase = which == 1 ? c.yprev() : c.Nrm(1);
  d.xmap = which == 1 ? c.gmap : nullptr;
  d.Wg = c.P(which == 1 ? P_W1 : P_W2);
  d.wk = 1; d.w_ld = 0; d.pair_np = 0; d.pair_db = 0;
  return d;
}
// per-feature parameters of a hidden-layer tile: Pm[0] bias, [1] gamma, [2] beta, [3] running_mean, [4] running_var
__device__ __forceinline__ void issue_pm_hid_fwd(const KArgs& a, const Sm& s, const Cx& c, int which, int n0, int NT) {
  load_vec(s.Pm, c.P(which == 1 ? P_B1 : P_B2), n0, NT, a.H);
  if (a.has_bn) {
    load_vec(s.Pm + 64, c.P(which == 1 ? P_G1 : P_G2), n0, NT, a.H);
    load_vec(s.Pm + 128, c.P(which == 1 ? P_BE1 : P_BE2), n0, NT, a.H);
    load_vec(s.Pm + 192, a.R + c.st.r[(which - 1) * 2], n0, NT, a.H);
    load_vec(s.Pm + 256, a.R + c.st.r[(which - 1) * 2 + 1], n0, NT, a.H);
  }
}

// net.0-2 / net.3-5: Linear + LeakyReLU(0.01) + BatchNorm1d(eps 1e-2)  (:171-187)
__device__ void run_hid_fwd(const SArgs& A, const Sm& s, int e, int which, bool prefetched, bool& cl_pending) {
  const KArgs& a = A.k;
  Cx c(a, s.st, e);
  const ActDesc d = desc_hid_fwd(A, c, which);
  const ActPlan& pl = *d.pl;
  const Who w = who(A, pl);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float invB = 1.0f / (float)a.B;
  bool pf = prefetched;
  for (int tile = w.unit; tile < pl.n_tiles; tile += w.units) {
    const int n0 = tile * pl.NT;
    if (!pf) {
      __syncthreads();
      issue_pm_hid_fwd(a, s, c, which, n0, pl.NT);   // joins the weight batch committed inside act_tile
    }
    act_tile(a, s, d, tile, w.rank, w.split, pf, cl_pending, [] {});
    pf = false;
    for (int i = warp;; i += 8) {
      const int nl = w.rank + w.nr * i;
      if (nl >= pl.NT) break;
      float s0, s1;
      combine(s, nl, lane, w.nr, s0, s1);
      const int n = n0 + nl;
      if (n >= a.H) continue;
      const bool ok0 = lane < a.B, ok1 = lane + 32 < a.B;
      const float bj = s.Pm[nl];
      const float v0 = lrelu_f(s0 + bj), v1 = lrelu_f(s1 + bj);
      float* Lr = c.L(which) + (size_t)n * a.B;
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
            __stcg(c.stats(which, 0) + n, mean);
            __stcg(c.stats(which, 1) + n, rstd);
            a.R[c.st.r[(which - 1) * 2] + n] = (1.f - kBnMomentum) * rm + kBnMomentum * mean;
            a.R[c.st.r[(which - 1) * 2 + 1] + n] =
                (1.f - kBnMomentum) * rv + kBnMomentum * (var * (float)a.B / (float)(a.B - 1));
          }
        } else {   // eval: running statistics
          const float rs = 1.0f / sqrtf(rv + kBnEps);
          const float ga = gam * rs;
          const float be = bet - rm * ga;
          n0v = fmaf(v0, ga, be);
          n1v = fmaf(v1, ga, be);
        }
      }
      float* Nr = c.Nrm(which) + (size_t)n * a.B;
      if (ok0) __stcg(Nr + lane, n0v);
      if (ok1) __stcg(Nr + lane + 32, n1v);
    }
    if (w.split) {
      cluster_arrive();
      cl_pending = true;
    }
  }
}
__device__ void prefetch_hid_fwd(const SArgs& A, const Sm& s, int e, int which) {
  const KArgs& a = A.k;
  Cx c(a, s.st, e);
  const ActDesc d = desc_hid_fwd(A, c, which);
  const Who w = who(A, *d.pl);
  if (w.unit >= d.pl->n_tiles) return;
  prefetch_act(a, s, d, w.unit, w.rank);
  issue_pm_hid_fwd(a, s, c, which, w.unit * d.pl->NT, d.pl->NT);
}

// ZeroFC (:137-141) + affine coupling (:244-249 / :226-231) + ActNorm (:66 / :45) + logdet partial.
// A tile = NP coupling columns j: rows j (log_s) and Db + j (t) of net.6.fc.weight.
__device__ __forceinline__ ActDesc desc_out(const SArgs& A, const Cx& c) {
  ActDesc d;
  d.pl = &A.sp.f3;
  d.Xbase = c.Nrm(2); d.xmap = nullptr;
  d.Wg = c.P(P_W3); d.wk = 1; d.w_ld = 0;
  d.pair_np = A.sp.f3.NT / 2; d.pair_db = A.k.Db;
  return d;
}
__device__ __forceinline__ void issue_pm_out(const KArgs& a, const Sm& s, const Cx& c, int j0, int NP) {
  load_vec(s.Pm, c.P(P_SCALE), j0, NP, a.Db);
  load_vec(s.Pm + 64, c.P(P_SCALE) + a.Db, j0, NP, a.Db);
  load_vec(s.Pm + 128, c.P(P_B3), j0, NP, a.Db);
  load_vec(s.Pm + 192, c.P(P_B3) + a.Db, j0, NP, a.Db);
  load_ivec(s.idx2, c.gmap, j0, NP, a.Db);                 // a-half source rows
  load_ivec(s.idx2 + 64, c.gmap + a.Da, j0, NP, a.Db);     // b-half source rows
}
__device__ void run_out(const SArgs& A, const Sm& s, int e, bool prefetched, bool& cl_pending) {
  const KArgs& a = A.k;
  Cx c(a, s.st, e);
  const ActDesc d = desc_out(A, c);
  const ActPlan& pl = *d.pl;
  const int NP = pl.NT / 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* yp = c.yprev();
  float qa, qc;
  c.post_affine(qa, qc);
  bool pf = prefetched;
  for (int tile = blockIdx.x; tile < pl.n_tiles; tile += A.sp.NC) {
    const int j0 = tile * NP;
    if (!pf) {
      __syncthreads();
      issue_pm_out(a, s, c, j0, NP);
    }
    const int* idx2 = s.idx2;
    float* E = s.E;
    const int B = a.B, Db = a.Db;
    const bool vec = c.vec;
    act_tile(a, s, d, tile, 0, false, pf, cl_pending, [=] {
      load_rows(E, kMB, yp, B, vec, 2 * NP, [=](int r) -> long long {
        const int half = r / NP, j = j0 + (r - half * NP);
        return j < Db ? idx2[half * 64 + (r - half * NP)] : -1;
      });
    });
    pf = false;
    float ld0 = 0.f, ld1 = 0.f;
    for (int p = warp; p < NP; p += 8) {
      const int col = j0 + p;
      if (col >= a.Db) continue;
      const float e_ls = expf(3.f * s.Pm[p]), e_t = expf(3.f * s.Pm[64 + p]);
      const float b_ls = s.Pm[128 + p], b_t = s.Pm[192 + p];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int m = lane + 32 * hh;
        const float ols = (s.R[p * kMB + m] + b_ls) * e_ls;
        const float ot = (s.R[(NP + p) * kMB + m] + b_t) * e_t;
        const float av = s.E[p * kMB + m], bv = s.E[(NP + p) * kMB + m];
        const float ls = tanhf(ols);
        float bp, dl;
        if (a.dir == 0) { bp = bv / expf(ls) - ot; dl = -ls; }
        else            { bp = (bv + ot) * expf(ls); dl = ls; }
        if (m < a.B) {
          __stcg(c.O() + (size_t)col * a.B + m, ols);
          __stcg(c.O() + (size_t)(a.Db + col) * a.B + m, ot);
          __stcg(c.y() + (size_t)(a.Da + col) * a.B + m, fmaf(bp, qa, qc));
          __stcg(c.y() + (size_t)col * a.B + m, fmaf(av, qa, qc));
          if (hh == 0) ld0 += dl; else ld1 += dl;
        }
      }
    }
    // odd D: the a-half has one more row than the b-half (chunk(2,1) semantics, :241)
    if (a.Da > a.Db && tile == 0 && warp == 0) {
      const int col = a.Da - 1;
      const float* src = yp + (size_t)__ldg(c.gmap + col) * a.B;
      for (int m = lane; m < a.B; m += 32) __stcg(c.y() + (size_t)col * a.B + m, fmaf(__ldcg(src + m), qa, qc));
    }
    // per-sample logdet partial of this tile: the 8 warps' sums in fixed order
    s.red[warp * 64 + lane] = ld0;
    s.red[warp * 64 + lane + 32] = ld1;
    __syncthreads();
    if (threadIdx.x < 64) {
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) t += s.red[r * 64 + threadIdx.x];
      if ((int)threadIdx.x < a.B) __stcg(a.W + a.oLDP + ((size_t)e * A.sp.ldp_tiles + tile) * a.B + threadIdx.x, t);
    }
    __syncthreads();
  }
}
__device__ void prefetch_out(const SArgs& A, const Sm& s, int e) {
  const KArgs& a = A.k;
  Cx c(a, s.st, e);
  const ActDesc d = desc_out(A, c);
  if ((int)blockIdx.x >= d.pl->n_tiles) return;
  prefetch_act(a, s, d, blockIdx.x, 0);
  issue_pm_out(a, s, c, blockIdx.x * (d.pl->NT / 2), d.pl->NT / 2);
}

// ------------------------------------------------------------------------------- transposes / finals
// 32 x 32 tiles through shared memory. mode:
//   0 pass input : in [B][D] -> ZT [D][B] (forward direction: stage-0 ActNorm applied)
//   1 backward input: g_out [B][D] -> GYT[(S-1)&1] [D][B]
//   2 reverse output: YT[S-1] [D][B] -> out [B][D]
//   3 forward output: YT[S-1][gf[j]][m] -> out [B][D] (trailing permutation :590 after :446)
//   4 backward output: GYT[1] [D][B] -> g_in [B][D]
__device__ void run_transpose(const KArgs& a, const StageTab* stb, int mode, float* sT, int NC) {
  const bool to_fm = mode <= 1;
  const int R = to_fm ? a.B : a.D, C = to_fm ? a.D : a.B;
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
    src = a.g_out; dst = a.W + a.oGYT + (size_t)((a.S - 1) & 1) * a.D * a.B;
  } else if (mode == 2 || mode == 3) {
    src = a.W + a.oYT + (size_t)(a.S - 1) * a.D * a.B; dst = a.out;
    if (mode == 3) map = a.gmaps + (size_t)2 * a.S * a.D;
  } else {
    src = a.W + a.oGYT + (size_t)1 * a.D * a.B; dst = a.g_in;
  }
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
        v = __ldcg(src + (size_t)rr * C + cc);
      }
      sT[(ly + 8 * i) * 33 + lx] = fmaf(v, fa, fc);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int cc = tc * 32 + ly + 8 * i, r = tr * 32 + lx;
      if (r < R && cc < C) __stcg(dst + (size_t)cc * R + r, sT[lx * 33 + ly + 8 * i]);
    }
  }
}

// logdet[m] = sum over stages / tiles of the coupling partials + sum_s (+-) D * log_scale_inv_s  (CTA 0)
__device__ void run_logdet_final(const SArgs& A, const Sm& s) {
  const KArgs& a = A.k;
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int n = a.S * A.sp.ldp_tiles;
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
        for (int u = 0; u < 8; ++u) v[u] = ldcg(ldp + (size_t)(i + u) * a.B + m);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u & 3] += v[u];
      }
      for (; i < hi; ++i) acc[0] += ldcg(ldp + (size_t)i * a.B + m);
    }
    __syncthreads();
    s.red[slice * 32 + lane] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    __syncthreads();
    if (slice == 0 && m < a.B) {
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) t += s.red[r * 32 + lane];
      a.logdet[m] = t + (a.dir == 0 ? (float)a.D * lsum : -(float)a.D * lsum);
    }
  }
}

// ------------------------------------------------------------------------------- phases: backward pass
// Elementwise backward of ActNorm.reverse and the coupling tail, one warp per coupling column, a lane
// owns samples lane and lane + 32. Produces g_pre^T, the gradients of net.6.scale /
// net.6.fc.bias, the b-half rows of the gradient wrt the stage input and ActNorm partial sums.
__device__ void run_b0(const SArgs& A, const Sm& s, int e) {
  const KArgs& a = A.k;
  Cx c(a, s.st, e);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const float lsi = __ldg(a.P + c.st.lsi), loc = __ldg(a.P + c.st.loc);
  const float e_lsi = expf(lsi);
  const size_t B = a.B;
  for (int t = blockIdx.x; t < A.sp.red_tiles; t += A.sp.NC) {
    const int col = t * 8 + w;
    float S1 = 0.f, S2 = 0.f;
    if (col < a.Db) {
      const float E_ls = expf(3.f * __ldg(c.P(P_SCALE) + col));
      const float E_t = expf(3.f * __ldg(c.P(P_SCALE) + a.Db + col));
      const int gcol = __ldg(c.gmap + a.Da + col);
      const float* gya = c.gy() + (size_t)col * B;
      const float* gyb = c.gy() + (size_t)(a.Da + col) * B;
      const float* ya = c.y() + (size_t)col * B;
      const float* yb = c.y() + (size_t)(a.Da + col) * B;
      const float* ols = c.O() + (size_t)col * B;
      const float* ot = c.O() + (size_t)(a.Db + col) * B;
      const float* bp = c.yprev() + (size_t)gcol * B;
      float* gp_ls = a.W + a.oGPRET + (size_t)col * B;
      float* gp_t = a.W + a.oGPRET + (size_t)(a.Db + col) * B;
      float* dst = c.gyprev() + (size_t)gcol * B;
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
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int m = lane + 32 * hh;
        if (m < a.B) {
          const float ga_ = v_ga[hh], gb_ = v_gb[hh];
          S1 += ga_ + gb_;
          S2 += ga_ * (v_ya[hh] + loc) + gb_ * (v_yb[hh] + loc);
          const float gbp = gb_ * e_lsi;
          const float o_ls = v_ols[hh], o_t = v_ot[hh];
          const float ls = tanhf(o_ls);
          const float ei = 1.0f / expf(ls);
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
      const float* gq = c.gy() + (size_t)(a.Da - 1) * B;
      const float* yv = c.y() + (size_t)(a.Da - 1) * B;
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

// dgrad through net.6.fc.weight (which = 2: k = output row o, source g_pre^T) or net.3.weight (which = 1:
// k = h', source g_p2^T), with the BatchNorm + LeakyReLU backward of hidden layer `which` fused.
__device__ __forceinline__ ActDesc desc_dg_hid(const SArgs& A, const Cx& c, int which) {
  ActDesc d;
  d.pl = which == 2 ? &A.sp.dg2 : &A.sp.dg1;
  d.Xbase = A.k.W + (which == 2 ? A.k.oGPRET : A.k.oGP2T); d.xmap = nullptr;
  d.Wg = c.P(which == 2 ? P_W3 : P_W2); d.wk = 0; d.w_ld = A.k.H;
  d.pair_np = 0; d.pair_db = 0;
  return d;
}
__device__ __forceinline__ void issue_pm_dg_hid(const KArgs& a, const Sm& s, const Cx& c, int which, int n0, int NT) {
  if (a.has_bn) {
    load_vec(s.Pm, c.stats(which, 0), n0, NT, a.H);
    load_vec(s.Pm + 64, c.stats(which, 1), n0, NT, a.H);
    load_vec(s.Pm + 128, c.P(which == 1 ? P_G1 : P_G2), n0, NT, a.H);
  }
}
__device__ void run_dg_hid(const SArgs& A, const Sm& s, int e, int which, bool prefetched, bool& cl_pending) {
  const KArgs& a = A.k;
  Cx c(a, s.st, e);
  const ActDesc d = desc_dg_hid(A, c, which);
  const ActPlan& pl = *d.pl;
  const Who w = who(A, pl);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float invB = 1.0f / (float)a.B;
  float* dstbuf = a.W + (which == 2 ? a.oGP2T : a.oGP1T);
  bool pf = prefetched;
  for (int tile = w.unit; tile < pl.n_tiles; tile += w.units) {
    const int n0 = tile * pl.NT;
    if (!pf) {
      __syncthreads();
      issue_pm_dg_hid(a, s, c, which, n0, pl.NT);
    }
    float* E = s.E;
    const float* Lb = c.L(which);
    const int B = a.B, H = a.H, NT = pl.NT, rank = w.rank, nr = w.nr;
    const bool vec = c.vec;
    const int nslots = (NT - rank + nr - 1) / nr;     // features this rank finalises: n0 + rank + nr * i
    act_tile(a, s, d, tile, w.rank, w.split, pf, cl_pending, [=] {
      load_rows(E, kMB, Lb, B, vec, nslots > 0 ? nslots : 0, [=](int i) -> long long {
        const int n = n0 + rank + nr * i;
        return n < H ? n : -1;
      });
    });
    pf = false;
    for (int i = warp;; i += 8) {
      const int nl = w.rank + w.nr * i;
      if (nl >= pl.NT) break;
      float g0, g1;
      combine(s, nl, lane, w.nr, g0, g1);
      const int n = n0 + nl;
      if (n >= a.H) continue;
      const bool ok0 = lane < a.B, ok1 = lane + 32 < a.B;
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
      float* dr = dstbuf + (size_t)n * a.B;
      if (ok0) __stcg(dr + lane, p0);
      if (ok1) __stcg(dr + lane + 32, p1);
      if (lane == 0) {
        if (a.has_bn) {
          c.G(which == 1 ? P_G1 : P_G2)[n] = ggamma;
          c.G(which == 1 ? P_BE1 : P_BE2)[n] = gbeta;
        }
        c.G(which == 1 ? P_B1 : P_B2)[n] = gbias;
      }
    }
    if (w.split) {
      cluster_arrive();
      cl_pending = true;
    }
  }
}
__device__ void prefetch_dg_hid(const SArgs& A, const Sm& s, int e, int which) {
  const KArgs& a = A.k;
  Cx c(a, s.st, e);
  const ActDesc d = desc_dg_hid(A, c, which);
  const Who w = who(A, *d.pl);
  if (w.unit >= d.pl->n_tiles) return;
  prefetch_act(a, s, d, w.unit, w.rank);
  issue_pm_dg_hid(a, s, c, which, w.unit * d.pl->NT, d.pl->NT);
}

// dgrad through net.0.weight plus the direct path through the a-half and ActNorm, written through the
// fused permutation into the rows of the gradient wrt the previous stage's output.
__device__ __forceinline__ ActDesc desc_dg_in(const SArgs& A, const Cx& c) {
  ActDesc d;
  d.pl = &A.sp.dgin;
  d.Xbase = A.k.W + A.k.oGP1T; d.xmap = nullptr;
  d.Wg = c.P(P_W1); d.wk = 0; d.w_ld = A.k.Da;
  d.pair_np = 0; d.pair_db = 0;
  return d;
}
__device__ void run_dg_in(const SArgs& A, const Sm& s, int e, bool prefetched, bool& cl_pending) {
  const KArgs& a = A.k;
  Cx c(a, s.st, e);
  const ActDesc d = desc_dg_in(A, c);
  const ActPlan& pl = *d.pl;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float e_lsi = expf(__ldg(a.P + c.st.lsi));
  bool pf = prefetched;
  for (int tile = blockIdx.x; tile < pl.n_tiles; tile += A.sp.NC) {
    const int n0 = tile * pl.NT;
    if (!pf) {
      __syncthreads();
      load_ivec(s.idx2, c.gmap, n0, pl.NT, a.Da);
    }
    float* E = s.E;
    const float* gyb = c.gy();
    const int B = a.B, Da = a.Da, NT = pl.NT;
    const bool vec = c.vec;
    act_tile(a, s, d, tile, 0, false, pf, cl_pending, [=] {
      load_rows(E, kMB, gyb, B, vec, NT, [=](int r) -> long long { return n0 + r < Da ? n0 + r : -1; });
    });
    pf = false;
    for (int nl = warp; nl < pl.NT; nl += 8) {
      const int n = n0 + nl;
      if (n >= a.Da) continue;
      float* dr = c.gyprev() + (size_t)s.idx2[nl] * a.B;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int m = lane + 32 * hh;
        if (m < a.B) __stcg(dr + m, fmaf(s.E[nl * kMB + m], e_lsi, s.R[nl * kMB + m]));
      }
    }
  }
}
__device__ void prefetch_dg_in(const SArgs& A, const Sm& s, int e) {
  const KArgs& a = A.k;
  Cx c(a, s.st, e);
  const ActDesc d = desc_dg_in(A, c);
  if ((int)blockIdx.x >= d.pl->n_tiles) return;
  prefetch_act(a, s, d, blockIdx.x, 0);
  load_ivec(s.idx2, c.gmap, blockIdx.x * d.pl->NT, d.pl->NT, a.Da);
}

// Weight gradients gW[i][j] = sum_m g^T[i][m] * act^T[j][m]  (which = 3: fc.weight, 2: net.3, 1: net.0), computed
// in the shadow of the grid barrier. Tile TI x TJ, thread tile 4 x 4 (rows 4 iq + a, columns jq + (TJ/4) b).
template <int TI, int TJ, int MS>
__device__ __forceinline__ void wgrad_tiles(const SArgs& A, const Sm& s, int e, int which) {
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
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int jq, iq, mp;
  if (MS == 4) { jq = lane & 15; iq = (lane >> 4) + 2 * (warp & 1); mp = warp >> 1; }
  else { jq = lane; iq = warp; mp = 0; }
  const bool vec = c.vec;
  for (int t = blockIdx.x; t < tiles_i * tiles_j; t += A.sp.NC) {
    const int ti = t / tiles_j, tj = t - ti * tiles_j;
    const int i0 = ti * TI, j0 = tj * TJ;
    __syncthreads();   // previous users of the X region are done
    load_rows(G_s, kLdG, gsrc, a.B, vec, TI, [=](int r) -> long long { return i0 + r < I ? i0 + r : -1; });
    load_rows(A_s, kLdG, act, a.B, vec, TJ, [=](int r) -> long long {
      const int j = j0 + r;
      return j < J ? (amap ? __ldg(amap + j) : j) : -1;
    });
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y) acc[x][y] = 0.f;
    constexpr int MW = kMB / MS;
#pragma unroll 2
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
__device__ void run_wgrad(const SArgs& A, const Sm& s, int e, int which) {
  if (A.sp.wgv == 0) wgrad_tiles<16, 64, 4>(A, s, e, which);
  else wgrad_tiles<32, 128, 1>(A, s, e, which);
}

// ActNorm.reverse gradients: g_loc = -sum g_y, g_lsi = sum g_y * (y + loc) + D * sum_m g_ld[m]  (CTA 0)
__device__ void run_bwd_final(const SArgs& A, const Sm& s) {
  const KArgs& a = A.k;
  float sacc = 0.f;
  for (int m = threadIdx.x; m < a.B; m += kT) sacc += ldcg(a.g_ld + m);
  const float gld = block_sum(sacc, s.red);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int e = w; e < a.S; e += kT / 32) {
    const float* red = a.W + a.oRED + (size_t)e * A.sp.red_tiles * 2;
    float s1 = 0.f, s2 = 0.f;
    for (int t = lane; t < A.sp.red_tiles; t += 32) { s1 += ldcg(red + 2 * t); s2 += ldcg(red + 2 * t + 1); }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
      const StageTab& st = s.st[e];
      a.G[st.loc] = -s1;
      a.G[st.lsi] = s2 + (float)a.D * gld;
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
    if (p == 0) { kind = K_TIN; e = S - 1; }
    else if (p <= 4 * S) { e = S - 1 - (p - 1) / 4; kind = K_B0 + (p - 1) % 4; }
    else { kind = K_BFIN; e = 0; }
  }
}

__device__ __forceinline__ void prefetch_phase(const SArgs& A, const Sm& s, int kind, int e) {
  switch (kind) {
    case K_F1: prefetch_hid_fwd(A, s, e, 1); break;
    case K_F2: prefetch_hid_fwd(A, s, e, 2); break;
    case K_F3: prefetch_out(A, s, e); break;
    case K_DG2: prefetch_dg_hid(A, s, e, 2); break;
    case K_DG1: prefetch_dg_hid(A, s, e, 1); break;
    case K_DGIN: prefetch_dg_in(A, s, e); break;
    default: break;
  }
  cp_async_commit();
}

}  // namespace

__global__ void __launch_bounds__(kT, 1) flow_small_kernel(const __grid_constant__ SArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float sRed[8 * 64 + 128];
  __shared__ __align__(16) StageTab sSt[64];
  Sm s;
  s.X = reinterpret_cast<float*>(smem_raw + A.sp.oX);
  s.W = reinterpret_cast<float*>(smem_raw + A.sp.oW);
  s.E = reinterpret_cast<float*>(smem_raw + A.sp.oE);
  s.P = reinterpret_cast<float*>(smem_raw + A.sp.oP);
  s.R = reinterpret_cast<float*>(smem_raw + A.sp.oR);
  s.Pm = reinterpret_cast<float*>(smem_raw + A.sp.oPm);
  s.idx = reinterpret_cast<int*>(smem_raw + A.sp.oIdx);
  s.idx2 = reinterpret_cast<int*>(smem_raw + A.sp.oIdx2);
  s.red = sRed;
  s.st = A.k.st;
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
  int kind, e;
  decode(A, a.phase_begin, kind, e);
  prefetch_phase(A, s, kind, e);
  for (int p = a.phase_begin; p < a.phase_end; ++p) {
    decode(A, p, kind, e);
    if (persistent && p > a.phase_begin) grid_wait(a.ctrl, (++bar) * gridDim.x);
    if (a.tbuf && blockIdx.x == 0 && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      a.tbuf[2 * p] = t;
    }
    switch (kind) {
      case K_TIN: run_transpose(a, s.st, A.bwd ? 1 : 0, s.X, A.sp.NC); break;
      case K_F1: run_hid_fwd(A, s, e, 1, true, cl_pending); break;
      case K_F2: run_hid_fwd(A, s, e, 2, true, cl_pending); break;
      case K_F3: run_out(A, s, e, true, cl_pending); break;
      case K_FFIN:
        if (blockIdx.x == 0) run_logdet_final(A, s);
        __syncthreads();
        run_transpose(a, s.st, a.dir == 0 ? 2 : 3, s.X, A.sp.NC);
        break;
      case K_B0: run_b0(A, s, e); break;
      case K_DG2: run_dg_hid(A, s, e, 2, true, cl_pending); break;
      case K_DG1: run_dg_hid(A, s, e, 1, true, cl_pending); break;
      case K_DGIN: run_dg_in(A, s, e, true, cl_pending); break;
      case K_BFIN:
        if (blockIdx.x == 0) run_bwd_final(A, s);
        __syncthreads();
        if (a.g_in) run_transpose(a, s.st, 4, s.X, A.sp.NC);
        break;
    }
    if (a.tbuf && blockIdx.x == 0 && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      a.tbuf[2 * p + 1] = t;
    }
    if (p + 1 < a.phase_end) {
      if (persistent) grid_arrive(a.ctrl); else __syncthreads();
      int k2, e2;
      decode(A, p + 1, k2, e2);
      prefetch_phase(A, s, k2, e2);        // next hop's weights / indices / parameters -> shared memory
    }
    // shadow of the barrier: weight gradients of this stage (inputs complete since the previous barrier)
    if (kind == K_DG2) run_wgrad(A, s, e, 3);
    else if (kind == K_DG1) run_wgrad(A, s, e, 2);
    else if (kind == K_DGIN) run_wgrad(A, s, e, 1);
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

ActPlan plan_act(int N, int K, bool split, int G, int NC, int cl, bool wk, int max_nt_log) {
  ActPlan p{};
  p.N = N; p.K = K;
  p.ksplit = split ? 1 : 0;
  const int units = split ? G : NC;
  int nt = 4;
  while (nt < 64 && (N + nt - 1) / nt > units) nt *= 2;
  (void)max_nt_log;
  p.NT = nt;
  p.n_tiles = (N + nt - 1) / nt;
  const int nr = split ? cl : 1;
  p.KC = up4((K + nr - 1) / nr);
  p.KS = std::min(p.KC, 512);
  const int wmax = 36 * 1024;
  auto wbytes = [&](int ks) { return wk ? nt * (ks + 4) * 4 : ks * (nt + 4) * 4; };
  while (p.KS > 4 && wbytes(p.KS) > wmax) p.KS = up4(p.KS / 2);
  p.nq = nt / 4;
  p.kp = 16 / p.nq;
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
  const int max_smem = 200 * 1024;
  if (cudaFuncSetAttribute((const void*)flow_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) != cudaSuccess) {
    cudaGetLastError();
    h->small_mode = 0;
    return DPI_OK;
  }
  const char* nocl = getenv("DPI_B200_SMALL_NOCLUSTER");
  if (!(nocl && nocl[0] == '1')) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(8 * 18); cfg.blockDim = dim3(kT); cfg.dynamicSmemBytes = max_smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 8; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, (const void*)flow_small_kernel, &cfg) == cudaSuccess && nc >= 4) {
      h->small_clusters = nc;
      h->small_cl = 8;
    } else {
      cudaGetLastError();
    }
  }
  if (h->small_cl == 1) h->small_clusters = h->n_sm;   // plain cooperative launch, one CTA per SM
  return DPI_OK;
}

void small_plan(const dpi_flow_s* h, BatchPlan& bp) {
  SmallPlan& sp = bp.small;
  sp = SmallPlan{};
  if (!h->small_mode || bp.B > kMB || h->H > 512 || h->S < 1) return;
  const int cl = h->small_cl, G = h->small_clusters, NC = cl == 8 ? 8 * G : G;
  sp.NC = NC; sp.G = cl == 8 ? G : NC; sp.cl = cl;
  const int H = h->H, Da = h->Da, Db = h->Db, Dout = h->Dout;
  auto split = [&](int K) { return cl == 8 && K >= h->small_ksplit_min; };
  sp.f1 = plan_act(H, Da, split(Da), sp.G, NC, cl, true, 0);
  sp.f2 = plan_act(H, H, split(H), sp.G, NC, cl, true, 0);
  {  // output layer: tiles of NP column pairs -> 2 NP weight rows, never split (N large, K = H small)
    int np = 2;
    while (np < 32 && (Db + np - 1) / np > NC) np *= 2;
    ActPlan p = plan_act(2 * np, H, false, sp.G, NC, cl, true, 0);
    p.NT = 2 * np; p.nq = p.NT / 4; p.kp = 16 / p.nq;
    p.N = Db; p.n_tiles = (Db + np - 1) / np;
    p.KS = std::min(p.KC, 512);
    while (p.KS > 4 && p.NT * (p.KS + 4) * 4 > 36 * 1024) p.KS = up4(p.KS / 2);
    sp.f3 = p;
  }
  sp.dg2 = plan_act(H, Dout, split(Dout), sp.G, NC, cl, false, 0);
  sp.dg1 = plan_act(H, H, split(H), sp.G, NC, cl, false, 0);
  sp.dgin = plan_act(Da, H, false, sp.G, NC, cl, false, 0);
  const long long wg3 = ((long long)Dout + 15) / 16 * ((H + 63) / 64);
  sp.wgv = wg3 > 4LL * NC ? 1 : 0;
  sp.ldp_tiles = sp.f3.n_tiles;
  sp.red_tiles = (Db + 7) / 8;
  const ActPlan* all[6] = {&sp.f1, &sp.f2, &sp.f3, &sp.dg2, &sp.dg1, &sp.dgin};
  int xb = 0, wb = 0, ib = 0;
  for (int i = 0; i < 6; ++i) {
    const ActPlan& p = *all[i];
    const bool wk = i < 3;
    xb = std::max(xb, p.KS * kMB * 4);
    wb = std::max(wb, wk ? p.NT * (p.KS + 4) * 4 : p.KS * (p.NT + 4) * 4);
    ib = std::max(ib, p.KS * 4);
  }
  const int wg_bytes = sp.wgv == 0 ? (16 + 64) * kLdG * 4 + 16 * 1024 : (32 + 128) * kLdG * 4;
  xb = std::max(xb, std::max(wg_bytes, 33 * 32 * 4));
  auto al = [](int x) { return (x + 127) & ~127; };
  int o = 0;
  sp.oX = o; o += al(xb);
  sp.oW = o; o += al(wb);
  sp.oE = o; o += 64 * kMB * 4;
  sp.oP = o; o += 16 * 4 * kMB * 4;
  sp.oR = o; o += 64 * kMB * 4;
  sp.oPm = o; o += 5 * 64 * 4;
  sp.oIdx = o; o += al(ib);
  sp.oIdx2 = o; o += 128 * 4;
  sp.smem_bytes = o;
  sp.ok = o <= 200 * 1024 ? 1 : 0;
}

int small_launch(dpi_flow_s* h, const BatchPlan& bp, KArgs& a, int bwd, cudaStream_t stream) {
  SArgs A{};
  A.k = a;
  A.sp = bp.small;
  A.bwd = bwd;
  A.k.phases = nullptr;
  A.k.use_tc = 0;
  A.k.tbuf = nullptr;
  A.k.tiles_m = 1;
  A.k.ldp_tiles = bp.small.ldp_tiles;
  A.k.red_tiles = bp.small.red_tiles;
  const int n_phases = bwd ? 4 * h->S + 2 : 3 * h->S + 2;
  if (h->d_tbuf && n_phases <= 4096) {
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
  cfg.gridDim = dim3(bp.small.NC);
  cfg.blockDim = dim3(kT);
  cfg.dynamicSmemBytes = bp.small.smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (bp.small.cl > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = bp.small.cl; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
    ++na;
  }
  if (h->persistent) {
    at[na].id = cudaLaunchAttributeCooperative;
    at[na].val.cooperative = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  if (h->persistent) {
    A.k.phase_begin = 0; A.k.phase_end = n_phases; A.k.persistent = 1;
    DPI_CUDA(cudaLaunchKernelEx(&cfg, flow_small_kernel, A));
    count_launch();
  } else {
    A.k.persistent = 0;
    for (int p = 0; p < n_phases; ++p) {
      A.k.phase_begin = p; A.k.phase_end = p + 1;
      DPI_CUDA(cudaLaunchKernelEx(&cfg, flow_small_kernel, A));
      count_launch();
    }
  }
  return DPI_OK;
}

}  // namespace dpi
