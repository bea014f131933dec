// flow_kernels.cu - device code of the RealNVP phase program (fp32 SIMT path, feature-major layout).
//
// Math follows generative_model/realnvpfc_model.py:
//   AffineCoupling.reverse :240-249, .forward :222-231, net :171-187, ZeroFC :137-141,
//   ActNorm.reverse :49-66, .forward :30-47, Flow.reverse :457-474, RealNVP.reverse :594-603.
//
// Data layout in HBM: every activation / gradient of the flow is stored FEATURE-MAJOR,
// X^T[feature][sample]. The fixed permutations of the flow (:443,446,459,462,590,599) then never
// move data: a stage reads the previous stage's output ROWS through a fused index map (coalesced
// 256-byte row reads instead of 4-byte gathers) and writes its own rows contiguously. BatchNorm
// statistics are reductions along a row. Only the flow's input and output are sample-major
// ([batch, ndim], the reference's layout); two transposes per pass convert.
#include "flow.cuh"

namespace dpi {

// ----------------------------------------------------------------------------- grid barrier
// Monotonic-counter barrier over all CTAs of a cooperative launch: ctrl[0] counts arrivals over the
// whole launch (zeroed by the host before the launch), barrier number `idx` completes when it reaches
// (idx + 1) * nblocks. Thread 0 publishes the CTA's writes with a release-add (cumulative over the
// preceding bar.sync) and polls with acquire loads. The spin is bounded (4e9 cycles, about 2 s) so a protocol
// bug cannot hang the GPU; ctrl[2] is the sticky error flag, which the final phases turn into NaN outputs.
__device__ __forceinline__ void grid_sync(unsigned* ctrl, unsigned nblocks, unsigned idx) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned target = (idx + 1u) * nblocks;
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

// ----------------------------------------------------------------------------- addressing helpers
struct Ctx {
  const KArgs& a;
  int e;                 // execution position
  const StageTab& st;
  const int* gmap;       // fused row map of this stage's input
  bool vec;              // 16-byte accesses along the sample axis are legal
  __device__ Ctx(const KArgs& a_, int e_)
      : a(a_), e(e_), st(a_.st[a_.dir ? a_.S - 1 - e_ : e_]),
        gmap(a_.gmaps + ((size_t)a_.dir * a_.S + e_) * a_.D), vec((a_.B & 3) == 0) {}
  __device__ __forceinline__ const float* yprev() const {
    return e == 0 ? a.W + a.oZT : a.W + a.oYT + (size_t)(e - 1) * a.D * a.B;
  }
  __device__ __forceinline__ float* y() const { return a.W + a.oYT + (size_t)e * a.D * a.B; }
  __device__ __forceinline__ float* L(int which) const {
    return a.W + (which == 1 ? a.oL1T : a.oL2T) + (size_t)e * a.H * a.B;
  }
  __device__ __forceinline__ float* Nrm(int which) const {
    return a.W + (which == 1 ? a.oN1T : a.oN2T) + (size_t)e * a.H * a.B;
  }
  __device__ __forceinline__ float* O() const { return a.W + a.oOT + (size_t)e * a.Dout * a.B; }
  __device__ __forceinline__ float* stats(int which, int what) const {  // what: 0 mean, 1 rstd
    return a.W + a.oST + ((size_t)e * 4 + (which - 1) * 2 + what) * a.H;
  }
  __device__ __forceinline__ const float* P(int id) const { return a.P + st.p[id]; }
  __device__ __forceinline__ float* G(int id) const { return a.G + st.p[id]; }
  __device__ __forceinline__ const float* gy() const { return a.W + a.oGYT + (size_t)(e & 1) * a.D * a.B; }
  __device__ __forceinline__ float* gyprev() const { return a.W + a.oGYT + (size_t)((e - 1) & 1) * a.D * a.B; }
  // Affine applied to this stage's coupling output before it is stored:
  //   reverse: this stage's ActNorm.reverse, y = w * exp(lsi) - loc                              (:66)
  //   forward: the NEXT stage's ActNorm.forward, v = (w + loc') / exp(lsi')                      (:45)
  //            (scalar and elementwise, so it commutes with the permutation in between)
  __device__ __forceinline__ void post_affine(float& qa, float& qc) const {
    if (a.dir == 0) {
      qa = expf(__ldg(a.P + st.lsi));
      qc = -__ldg(a.P + st.loc);
    } else if (e + 1 < a.S) {
      const StageTab& nx = a.st[a.S - 2 - e];
      const float inv = 1.0f / expf(__ldg(a.P + nx.lsi));
      qa = inv;
      qc = __ldg(a.P + nx.loc) * inv;
    } else {
      qa = 1.f; qc = 0.f;
    }
  }
};

__device__ __forceinline__ void decode_tile(const SubOp& so, int t, int& tm, int& tn, int& ks) {
  ks = t % so.splits;
  const int mn = t / so.splits;
  tn = mn % so.tiles_n;
  tm = mn / so.tiles_n;
}

template <int TN>
__device__ __forceinline__ void zero_acc(float (&acc)[4][TN]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
}

// ----------------------------------------------------------------------------- transposes
// 32 x 32 tiles through shared memory. mode (so.which):
//   0 forward pass input : in [B][D]           -> ZT [D][B]   (forward direction: stage-0 ActNorm applied)
//   1 backward pass input: g_out [B][D]        -> GYT[(S-1)&1] [D][B]
//   2 reverse output     : YT[S-1] [D][B]      -> out [B][D]
//   3 forward output     : YT[S-1][gf[j]][m]   -> out [B][D]  (trailing permutation :590 after :446)
//   4 backward output    : GYT[1] [D][B]       -> g_in [B][D]
__device__ __noinline__ void op_transpose(const KArgs& a, const SubOp& so, int t, float* sT) {
  const int mode = so.which;
  const bool to_fm = mode <= 1;            // sample-major -> feature-major
  const int R = to_fm ? a.B : a.D;         // source rows
  const int C = to_fm ? a.D : a.B;         // source cols
  const float* src; float* dst; const int* map = nullptr;
  float fa = 1.f, fc = 0.f;
  if (mode == 0) {
    src = a.in; dst = a.W + a.oZT;
    if (a.dir == 1) {  // first ActNorm.forward of RealNVP.forward
      const StageTab& s0 = a.st[a.S - 1];
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
  const int tiles_c = (C + 31) / 32;
  const int tr = t / tiles_c, tc = t % tiles_c;
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;  // 32 x 8
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = tr * 32 + ly + 8 * i, c = tc * 32 + lx;
    float v = 0.f;
    if (r < R && c < C) {
      const int rr = map ? __ldg(map + r) : r;
      v = __ldcg(src + (size_t)rr * C + c);
    }
    sT[(ly + 8 * i) * 33 + lx] = fmaf(v, fa, fc);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = tc * 32 + ly + 8 * i, r = tr * 32 + lx;   // dst is [C][R]
    if (r < R && c < C) __stcg(dst + (size_t)c * R + r, sT[lx * 33 + ly + 8 * i]);
  }
}

// ----------------------------------------------------------------------------- forward ops
// net.0 / net.3: Linear + LeakyReLU(0.01) (+ BatchNorm1d batch statistics, eps 1e-2) :171-187.
// Output rows (features) n, columns (samples) m: L^T (saved for backward) and the normalised N^T.
template <int BN>
__device__ __noinline__ void op_fwd_hid(const KArgs& a, int e, const SubOp& so, int t, float* smem, int* sFlag,
                                        const TcState& tc) {
  constexpr int TN = BN / 16;
  Ctx c(a, e);
  int tm, tn, ks;
  decode_tile(so, t, tm, tn, ks);
  const int which = so.which;
  const int K = which == 1 ? a.Da : a.H;
  OperandA A{which == 1 ? c.yprev() : c.Nrm(1), a.B, which == 1 ? c.gmap : nullptr, tm * kTM, a.B, K, c.vec};
  OperandB B{c.P(which == 1 ? P_W1 : P_W2), K, nullptr, tn * BN, a.H, K, (K & 3) == 0, 0, 0};
  float acc[4][TN];
  zero_acc<TN>(acc);
  const int kbeg = ks * so.kchunk, kend = min(K, kbeg + so.kchunk);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  // the epilogue's per-column parameters do not depend on the GEMM: issue their loads first so the
  // latency is hidden behind the pipeline instead of being paid after it
  float pb[TN], pg[TN], pbe[TN], prm[TN], prv[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int n = tn * BN + ty + 16 * j;
    const bool okn = n < a.H;
    pb[j] = okn ? __ldg(c.P(which == 1 ? P_B1 : P_B2) + n) : 0.f;
    pg[j] = (okn && a.has_bn) ? __ldg(c.P(which == 1 ? P_G1 : P_G2) + n) : 0.f;
    pbe[j] = (okn && a.has_bn) ? __ldg(c.P(which == 1 ? P_BE1 : P_BE2) + n) : 0.f;
    prm[j] = (okn && a.has_bn) ? a.R[c.st.r[(which - 1) * 2] + n] : 0.f;
    prv[j] = (okn && a.has_bn) ? a.R[c.st.r[(which - 1) * 2 + 1] + n] : 1.f;
  }
  if (a.use_tc) tile_gemm_tc<BN, A_MCONTIG, B_KCONTIG>(acc, A, B, kbeg, kend, smem, tc);
  else tile_gemm<BN, A_MCONTIG, B_KCONTIG>(acc, A, B, kbeg, kend, smem);
  float* part = a.W + a.oPART + so.part_off + (size_t)(tm * so.tiles_n + tn) * so.splits * (kTM * BN);
  if (!splitk_reduce_fm<BN>(acc, part, ks, so.splits, a.ctrl + so.ctr_off + tm * so.tiles_n + tn, sFlag)) return;

  const int m = tm * kTM + tx * 4;
  const int nvalid = a.B - m;  // rows of this thread that exist (may be <= 0)
  const bool fused = a.has_bn && a.train && a.tiles_m == 1;
  const float invB = 1.0f / (float)a.B;
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int n = tn * BN + ty + 16 * j;
    const bool okn = n < a.H;
    const float bj = pb[j];
    float v[4], s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[i] = lrelu_f(acc[i][j] + bj);
      if (i < nvalid) s += v[i];
    }
    if (okn && nvalid > 0) store4(c.L(which) + (size_t)n * a.B + m, v, nvalid, c.vec);
    float nv[4];
    if (!a.has_bn) {
#pragma unroll
      for (int i = 0; i < 4; ++i) nv[i] = v[i];
    } else if (fused) {
      const float mean = col_sum16(s) * invB;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i < nvalid) { const float d = v[i] - mean; q = fmaf(d, d, q); }
      const float var = col_sum16(q) * invB;
      const float rstd = 1.0f / sqrtf(var + kBnEps);
      const float ga = pg[j] * rstd;
      const float be = pbe[j];
#pragma unroll
      for (int i = 0; i < 4; ++i) nv[i] = fmaf(v[i] - mean, ga, be);
      if (tx == 0 && okn) {
        __stcg(c.stats(which, 0) + n, mean);
        __stcg(c.stats(which, 1) + n, rstd);
        a.R[c.st.r[(which - 1) * 2] + n] = (1.f - kBnMomentum) * prm[j] + kBnMomentum * mean;
        a.R[c.st.r[(which - 1) * 2 + 1] + n] =
            (1.f - kBnMomentum) * prv[j] + kBnMomentum * (var * (float)a.B / (float)(a.B - 1));
      }
    } else if (!a.train) {  // eval: running statistics
      const float rs = 1.0f / sqrtf(prv[j] + kBnEps);
      const float ga = pg[j] * rs;
      const float be = pbe[j] - prm[j] * ga;
#pragma unroll
      for (int i = 0; i < 4; ++i) nv[i] = fmaf(v[i], ga, be);
    } else {
      continue;  // large batch: statistics + normalisation in OP_BN_STATS
    }
    if (okn && nvalid > 0) store4(c.Nrm(which) + (size_t)n * a.B + m, nv, nvalid, c.vec);
  }
}

// Batch statistics + normalisation when the batch spans several row tiles: one warp per feature row.
__device__ __noinline__ void op_bn_stats(const KArgs& a, int e, const SubOp& so, int t) {
  Ctx c(a, e);
  const int which = so.which;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n = t * 8 + w;
  if (n >= a.H) return;
  const float* row = c.L(which) + (size_t)n * a.B;
  float s = 0.f;
  for (int m = lane; m < a.B; m += 32) s += __ldcg(row + m);
  const float mean = warp_sum(s) / (float)a.B;
  float q = 0.f;
  for (int m = lane; m < a.B; m += 32) { const float d = __ldcg(row + m) - mean; q = fmaf(d, d, q); }
  const float var = warp_sum(q) / (float)a.B;
  const float rstd = 1.0f / sqrtf(var + kBnEps);
  const float ga = __ldg(c.P(which == 1 ? P_G1 : P_G2) + n) * rstd;
  const float be = __ldg(c.P(which == 1 ? P_BE1 : P_BE2) + n);
  float* out = c.Nrm(which) + (size_t)n * a.B;
  for (int m = lane; m < a.B; m += 32) __stcg(out + m, fmaf(__ldcg(row + m) - mean, ga, be));
  if (lane == 0) {
    __stcg(c.stats(which, 0) + n, mean);
    __stcg(c.stats(which, 1) + n, rstd);
    float* rm = a.R + c.st.r[(which - 1) * 2] + n;
    float* rv = a.R + c.st.r[(which - 1) * 2 + 1] + n;
    *rm = (1.f - kBnMomentum) * *rm + kBnMomentum * mean;
    *rv = (1.f - kBnMomentum) * *rv + kBnMomentum * (var * (float)a.B / (float)(a.B - 1));
  }
}

// ZeroFC (:137-141) + affine coupling (:244-249 / :226-231) + ActNorm (:66 / :45) + logdet partial.
// The tile covers BJ = BN/2 coupling columns j and computes both log_s (row j of fc.weight) and
// t (row Db + j) so the coupling is applied in registers.
template <int BN>
__device__ __noinline__ void op_fwd_out(const KArgs& a, int e, const SubOp& so, int t, float* smem, float* sRed,
                                        const TcState& tc) {
  constexpr int TN = BN / 16, BJ = BN / 2, TJ = TN / 2;
  Ctx c(a, e);
  const int tn = t % so.tiles_n, tm = t / so.tiles_n;
  OperandA A{c.Nrm(2), a.B, nullptr, tm * kTM, a.B, a.H, c.vec};
  OperandB B{c.P(P_W3), a.H, nullptr, tn * BJ, a.Dout, a.H, (a.H & 3) == 0, BJ, a.Db};
  float acc[4][TN];
  zero_acc<TN>(acc);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m = tm * kTM + tx * 4;
  const int nvalid = a.B - m;
  const float* yp = c.yprev();
  // everything the epilogue needs besides the GEMM result is independent of it: load it first
  float qa, qc;
  c.post_affine(qa, qc);
  float p_els[TJ], p_et[TJ], p_bls[TJ], p_bt[TJ], p_bv[TJ][4], p_av[TJ][4];
#pragma unroll
  for (int jp = 0; jp < TJ; ++jp) {
    const int col = tn * BJ + ty + 16 * jp;
    const bool ok = col < a.Db && nvalid > 0;
    p_els[jp] = ok ? __ldg(c.P(P_SCALE) + col) : 0.f;
    p_et[jp] = ok ? __ldg(c.P(P_SCALE) + a.Db + col) : 0.f;
    p_bls[jp] = ok ? __ldg(c.P(P_B3) + col) : 0.f;
    p_bt[jp] = ok ? __ldg(c.P(P_B3) + a.Db + col) : 0.f;
    const int gb = ok ? __ldg(c.gmap + a.Da + col) : 0, ga = ok ? __ldg(c.gmap + col) : 0;
    load4(yp + (size_t)gb * a.B + m, p_bv[jp], ok ? nvalid : 0, c.vec);
    load4(yp + (size_t)ga * a.B + m, p_av[jp], ok ? nvalid : 0, c.vec);
  }
  if (a.use_tc) tile_gemm_tc<BN, A_MCONTIG, B_KCONTIG>(acc, A, B, 0, a.H, smem, tc);
  else tile_gemm<BN, A_MCONTIG, B_KCONTIG>(acc, A, B, 0, a.H, smem);

  float ld[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int jp = 0; jp < TJ; ++jp) {
    const int col = tn * BJ + ty + 16 * jp;
    if (col >= a.Db || nvalid <= 0) continue;
    const float e_ls = expf(3.f * p_els[jp]);
    const float e_t = expf(3.f * p_et[jp]);
    const float b_ls = p_bls[jp], b_t = p_bt[jp];
    const float (&bv)[4] = p_bv[jp];
    const float (&av)[4] = p_av[jp];
    float ols[4], ot[4], yb[4], ya[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ols[i] = (acc[i][jp] + b_ls) * e_ls;
      ot[i] = (acc[i][jp + TJ] + b_t) * e_t;
      const float ls = tanhf(ols[i]);
      float bp;
      if (a.dir == 0) { bp = bv[i] / expf(ls) - ot[i]; if (i < nvalid) ld[i] -= ls; }
      else            { bp = (bv[i] + ot[i]) * expf(ls); if (i < nvalid) ld[i] += ls; }
      yb[i] = fmaf(bp, qa, qc);
      ya[i] = fmaf(av[i], qa, qc);
    }
    store4(c.O() + (size_t)col * a.B + m, ols, nvalid, c.vec);
    store4(c.O() + (size_t)(a.Db + col) * a.B + m, ot, nvalid, c.vec);
    store4(c.y() + (size_t)(a.Da + col) * a.B + m, yb, nvalid, c.vec);
    store4(c.y() + (size_t)col * a.B + m, ya, nvalid, c.vec);
  }
  // odd D: the a-half has one more row than the b-half (chunk(2,1) semantics, :241)
  if (a.Da > a.Db && tn == 0 && ty == 0 && nvalid > 0) {
    const int col = a.Da - 1;
    float av[4], ya[4];
    load4(yp + (size_t)__ldg(c.gmap + col) * a.B + m, av, nvalid, c.vec);
#pragma unroll
    for (int i = 0; i < 4; ++i) ya[i] = fmaf(av[i], qa, qc);
    store4(c.y() + (size_t)col * a.B + m, ya, nvalid, c.vec);
  }
  // per-sample logdet partial of this column tile: sum over the 16 column groups in fixed order
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) sRed[ty * 64 + tx * 4 + i] = ld[i];
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 16; ++r) s += sRed[r * 64 + threadIdx.x];
    const int mm = tm * kTM + threadIdx.x;
    if (mm < a.B) __stcg(a.W + a.oLDP + ((size_t)e * a.ldp_tiles + tn) * a.B + mm, s);
  }
}

// Tensor-core path (flow_tc.cu): the ZeroFC GEMM leaves its raw output in O^T; this strip op applies what
// op_fwd_out's epilogue applies (ZeroFC bias and exp(3 scale) :137-141, tanh/exp coupling :244-249 / :226-231,
// ActNorm :66 / :45) with one thread per sample over a chunk of so.kchunk coupling columns (every access is a
// coalesced row segment) and writes that chunk's logdet partial.
__device__ __noinline__ void op_tail(const KArgs& a, int e, const SubOp& so, int t) {
  Ctx c(a, e);
  const int chunks = so.tiles_n, chunk = t % chunks, mb = t / chunks;
  const int m = mb * kThreads + threadIdx.x;
  if (m >= a.B) return;
  float qa, qc;
  c.post_affine(qa, qc);
  const float* yp = c.yprev();
  float* O = c.O();
  float* y = c.y();
  const size_t B = a.B;
  const int j0 = chunk * so.kchunk, j1 = min(a.Db, j0 + so.kchunk);
  float ld = 0.f;
#pragma unroll 2
  for (int j = j0; j < j1; ++j) {
    const float e_ls = expf(3.f * __ldg(c.P(P_SCALE) + j)), e_t = expf(3.f * __ldg(c.P(P_SCALE) + a.Db + j));
    const float b_ls = __ldg(c.P(P_B3) + j), b_t = __ldg(c.P(P_B3) + a.Db + j);
    const int gb = __ldg(c.gmap + a.Da + j), ga = __ldg(c.gmap + j);
    const float bv = __ldcg(yp + (size_t)gb * B + m), av = __ldcg(yp + (size_t)ga * B + m);
    const float ols = (__ldcg(O + (size_t)j * B + m) + b_ls) * e_ls;
    const float ot = (__ldcg(O + (size_t)(a.Db + j) * B + m) + b_t) * e_t;
    const float ls = tanhf(ols);
    float bp;
    if (a.dir == 0) { bp = bv / expf(ls) - ot; ld -= ls; }
    else            { bp = (bv + ot) * expf(ls); ld += ls; }
    __stcg(O + (size_t)j * B + m, ols);
    __stcg(O + (size_t)(a.Db + j) * B + m, ot);
    __stcg(y + (size_t)(a.Da + j) * B + m, fmaf(bp, qa, qc));
    __stcg(y + (size_t)j * B + m, fmaf(av, qa, qc));
  }
  if (a.Da > a.Db && chunk == 0) {   // odd D: the a-half has one more row (chunk(2,1) semantics, :241)
    const int col = a.Da - 1;
    __stcg(y + (size_t)col * B + m, fmaf(__ldcg(yp + (size_t)__ldg(c.gmap + col) * B + m), qa, qc));
  }
  __stcg(a.W + a.oLDP + ((size_t)e * a.ldp_tiles + chunk) * B + m, ld);
}

// logdet[m] = sum over stages/tiles of the coupling partials + sum_s (+-) D * log_scale_inv_s.
// One CTA per 32 samples: lane = sample (coalesced), 8 warps each sum a fixed slice of the S*tiles
// partials with independent loads in flight; slices are combined in a fixed order (deterministic).
__device__ __noinline__ void op_logdet_final(const KArgs& a, int t, float* sRed) {
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int m = t * 32 + lane;
  const int n = a.S * a.ldp_tiles;
  const float* ldp = a.W + a.oLDP;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (m < a.B) {
    const int per = (n + 7) / 8, lo = slice * per, hi = min(n, lo + per);
    int i = lo;
    for (; i + 16 <= hi; i += 16) {   // 16 independent loads in flight, fixed summation order
      float v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = ldcg(ldp + (size_t)(i + u) * a.B + m);
#pragma unroll
      for (int u = 0; u < 16; ++u) acc[u & 3] += v[u];
    }
    for (; i < hi; ++i) acc[0] += ldcg(ldp + (size_t)i * a.B + m);
  }
  __syncthreads();
  sRed[slice * 32 + lane] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  __syncthreads();
  if (slice == 0 && m < a.B) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += sRed[r * 32 + lane];
    float lsum = 0.f;
    for (int e = 0; e < a.S; ++e) lsum += __ldg(a.P + a.st[e].lsi);
    // a grid barrier that timed out (sticky flag ctrl[2]) must not pass silently: the outputs become NaN
    a.logdet[m] = s + (a.dir == 0 ? (float)a.D * lsum : -(float)a.D * lsum) + (__ldcg(a.ctrl + 2) ? __int_as_float(0x7fc00000) : 0.f);
  }
}

// ----------------------------------------------------------------------------- backward ops
// Elementwise backward of ActNorm.reverse and the coupling tail, one warp per coupling column
// (SURVEY Appendix A). Produces g_pre^T (gradient wrt ZeroFC's linear output), the gradients of
// net.6.scale / net.6.fc.bias, the b-half rows of the gradient wrt the stage input, and per-tile
// partials of the ActNorm gradients.
__device__ __noinline__ void op_b0(const KArgs& a, int e, const SubOp& so, int t, float* sRed) {
  Ctx c(a, e);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // so.which == 1 (narrow flows, few coupling columns): the whole CTA works on ONE column, samples split over all
  // 256 threads; else one warp per column.
  const bool wide = so.which == 1;
  const int col = wide ? t : t * 8 + w;
  const int m0 = wide ? (int)threadIdx.x : lane, mstep = wide ? kThreads : 32;
  const float lsi = __ldg(a.P + c.st.lsi), loc = __ldg(a.P + c.st.loc);
  const float e_lsi = expf(lsi);
  const size_t B = a.B;
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
    float sc_ls = 0.f, sc_t = 0.f, b3_ls = 0.f, b3_t = 0.f;
    const float* gld = a.g_ld;
    auto elem = [&](float ga_, float gb_, float ya_, float yb_, float o_ls, float o_t, float b, float gldv, float& v_ls,
                    float& v_t, float& d) {
      S1 += ga_ + gb_;
      S2 += ga_ * (ya_ + loc) + gb_ * (yb_ + loc);
      const float gbp = gb_ * e_lsi;
      const float ls = tanhf(o_ls);
      const float ei = 1.0f / expf(ls);
      const float g_t = -gbp;
      const float g_ls = -gbp * b * ei - gldv;
      const float g_ols = g_ls * (1.f - ls * ls);
      v_ls = g_ols * E_ls;
      v_t = g_t * E_t;
      sc_ls += 3.f * g_ols * o_ls;
      sc_t += 3.f * g_t * o_t;
      b3_ls += v_ls;
      b3_t += v_t;
      d = gbp * ei;
    };
    if (c.vec && (reinterpret_cast<uintptr_t>(gld) & 15) == 0) {   // 16-byte accesses: eight independent loads in flight per lane
      for (int m = m0 * 4; m < a.B; m += mstep * 4) {
        const float4 A_ = __ldcg(reinterpret_cast<const float4*>(gya + m)), B_ = __ldcg(reinterpret_cast<const float4*>(gyb + m));
        const float4 YA = __ldcg(reinterpret_cast<const float4*>(ya + m)), YB = __ldcg(reinterpret_cast<const float4*>(yb + m));
        const float4 OL = __ldcg(reinterpret_cast<const float4*>(ols + m)), OT = __ldcg(reinterpret_cast<const float4*>(ot + m));
        const float4 BP = __ldcg(reinterpret_cast<const float4*>(bp + m)), GL = __ldcg(reinterpret_cast<const float4*>(gld + m));
        float4 vl, vt, dd;
        elem(A_.x, B_.x, YA.x, YB.x, OL.x, OT.x, BP.x, GL.x, vl.x, vt.x, dd.x);
        elem(A_.y, B_.y, YA.y, YB.y, OL.y, OT.y, BP.y, GL.y, vl.y, vt.y, dd.y);
        elem(A_.z, B_.z, YA.z, YB.z, OL.z, OT.z, BP.z, GL.z, vl.z, vt.z, dd.z);
        elem(A_.w, B_.w, YA.w, YB.w, OL.w, OT.w, BP.w, GL.w, vl.w, vt.w, dd.w);
        __stcg(reinterpret_cast<float4*>(gp_ls + m), vl);
        __stcg(reinterpret_cast<float4*>(gp_t + m), vt);
        __stcg(reinterpret_cast<float4*>(dst + m), dd);
      }
    } else {
      for (int m = m0; m < a.B; m += mstep) {
        float v_ls, v_t, d;
        elem(__ldcg(gya + m), __ldcg(gyb + m), __ldcg(ya + m), __ldcg(yb + m), __ldcg(ols + m), __ldcg(ot + m), __ldcg(bp + m),
             __ldcg(gld + m), v_ls, v_t, d);
        __stcg(gp_ls + m, v_ls);
        __stcg(gp_t + m, v_t);
        __stcg(dst + m, d);
      }
    }
    if (wide) {   // col < Db for every thread of the CTA (the grid is exactly Db CTAs)
      sc_ls = block_sum(sc_ls, sRed + 128); sc_t = block_sum(sc_t, sRed + 192);
      b3_ls = block_sum(b3_ls, sRed + 256); b3_t = block_sum(b3_t, sRed + 320);
    } else {
      sc_ls = warp_sum(sc_ls); sc_t = warp_sum(sc_t); b3_ls = warp_sum(b3_ls); b3_t = warp_sum(b3_t);
    }
    if (wide ? threadIdx.x == 0 : lane == 0) {
      c.G(P_SCALE)[col] = sc_ls; c.G(P_SCALE)[a.Db + col] = sc_t;
      c.G(P_B3)[col] = b3_ls;    c.G(P_B3)[a.Db + col] = b3_t;
    }
  }
  if (a.Da > a.Db && t == 0) {  // odd D: last a-row contributes to the ActNorm sums
    const float* g = c.gy() + (size_t)(a.Da - 1) * B;
    const float* yv = c.y() + (size_t)(a.Da - 1) * B;
    for (int m = threadIdx.x; m < a.B; m += kThreads) {
      const float gv = __ldcg(g + m);
      S1 += gv;
      S2 += gv * (__ldcg(yv + m) + loc);
    }
  }
  const float t1 = block_sum(S1, sRed);
  const float t2 = block_sum(S2, sRed + 64);
  if (threadIdx.x == 0) {
    float* red = a.W + a.oRED + ((size_t)e * a.red_tiles + t) * 2;
    __stcg(red, t1);
    __stcg(red + 1, t2);
  }
}

// dgrad through fc.weight (which=2: K = Dout, source g_pre^T) or net.3.weight (which=1: K = H, source
// g_p2^T); result = gradient wrt the BatchNorm output of hidden layer `which`, with the BatchNorm +
// LeakyReLU backward fused when this CTA owns complete feature rows (batch <= 64).
template <int BN>
__device__ __noinline__ void op_dgrad_hid(const KArgs& a, int e, const SubOp& so, int t, float* smem, int* sFlag,
                                          const TcState& tc) {
  constexpr int TN = BN / 16;
  Ctx c(a, e);
  int tm, tn, ks;
  decode_tile(so, t, tm, tn, ks);
  const int which = so.which;
  const int K = which == 2 ? a.Dout : a.H;
  OperandA A{a.W + (which == 2 ? a.oGPRET : a.oGP2T), a.B, nullptr, tm * kTM, a.B, K, c.vec};
  OperandB B{c.P(which == 2 ? P_W3 : P_W2), a.H, nullptr, tn * BN, a.H, K, (a.H & 3) == 0, 0, 0};
  float acc[4][TN];
  zero_acc<TN>(acc);
  const int kbeg = ks * so.kchunk, kend = min(K, kbeg + so.kchunk);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m = tm * kTM + tx * 4;
  const int nvalid = a.B - m;
  const bool fused = a.tiles_m == 1;
  // epilogue operands that do not depend on the GEMM: load before it
  float p_l[TN][4], p_mean[TN], p_rstd[TN], p_gam[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int n = tn * BN + ty + 16 * j;
    const bool okn = n < a.H;
    p_mean[j] = 0.f; p_rstd[j] = 1.f; p_gam[j] = 1.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) p_l[j][i] = 0.f;
    if (fused) {
      if (okn && nvalid > 0) load4(c.L(which) + (size_t)n * a.B + m, p_l[j], nvalid, c.vec);
      if (okn && a.has_bn) {
        p_mean[j] = __ldcg(c.stats(which, 0) + n); p_rstd[j] = __ldcg(c.stats(which, 1) + n);
        p_gam[j] = __ldg(c.P(which == 1 ? P_G1 : P_G2) + n);
      }
    }
  }
  if (a.use_tc) tile_gemm_tc<BN, A_MCONTIG, B_NCONTIG>(acc, A, B, kbeg, kend, smem, tc);
  else tile_gemm<BN, A_MCONTIG, B_NCONTIG>(acc, A, B, kbeg, kend, smem);
  float* part = a.W + a.oPART + so.part_off + (size_t)(tm * so.tiles_n + tn) * so.splits * (kTM * BN);
  if (!splitk_reduce_fm<BN>(acc, part, ks, so.splits, a.ctrl + so.ctr_off + tm * so.tiles_n + tn, sFlag)) return;

  const float invB = 1.0f / (float)a.B;
  float* dstbuf = a.W + (fused ? (which == 2 ? a.oGP2T : a.oGP1T) : a.oGNT);
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int n = tn * BN + ty + 16 * j;
    const bool okn = n < a.H;
    float g[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) g[i] = (i < nvalid && okn) ? acc[i][j] : 0.f;
    if (fused) {
      const float (&l)[4] = p_l[j];
      const float gam = p_gam[j], rstd = p_rstd[j], mean = p_mean[j];
      float gbeta = 0.f, ggamma = 0.f;
      float xh[4] = {0.f, 0.f, 0.f, 0.f};
      if (a.has_bn) {
        float p1 = 0.f, p2 = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          xh[i] = (l[i] - mean) * rstd;
          p1 += g[i];
          p2 = fmaf(g[i], xh[i], p2);
        }
        gbeta = col_sum16(p1);
        ggamma = col_sum16(p2);
      }
      float p3 = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float gl = g[i];
        if (a.has_bn) gl = gam * rstd * (g[i] - gbeta * invB - xh[i] * ggamma * invB);
        g[i] = (i < nvalid && okn) ? gl * (l[i] > 0.f ? 1.f : kLreluSlope) : 0.f;
        p3 += g[i];
      }
      const float gbias = col_sum16(p3);
      if (tx == 0 && okn) {
        if (a.has_bn) {
          c.G(which == 1 ? P_G1 : P_G2)[n] = ggamma;
          c.G(which == 1 ? P_BE1 : P_BE2)[n] = gbeta;
        }
        c.G(which == 1 ? P_B1 : P_B2)[n] = gbias;
      }
    }
    if (okn && nvalid > 0) store4(dstbuf + (size_t)n * a.B + m, g, nvalid, c.vec);
  }
}

// Large-batch BatchNorm + LeakyReLU backward: one warp per feature row (stats pass, then apply pass).
__device__ __noinline__ void op_bnbwd(const KArgs& a, int e, const SubOp& so, int t) {
  Ctx c(a, e);
  const int which = so.which;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n = t * 8 + w;
  if (n >= a.H) return;
  const float* gn = a.W + a.oGNT + (size_t)n * a.B;
  const float* lr = c.L(which) + (size_t)n * a.B;
  float* gp = a.W + (which == 2 ? a.oGP2T : a.oGP1T) + (size_t)n * a.B;
  float mean = 0.f, rstd = 1.f, gam = 1.f, gbeta = 0.f, ggamma = 0.f;
  if (a.has_bn) {
    mean = __ldcg(c.stats(which, 0) + n); rstd = __ldcg(c.stats(which, 1) + n);
    gam = __ldg(c.P(which == 1 ? P_G1 : P_G2) + n);
    float s1 = 0.f, s2 = 0.f;
    for (int m = lane; m < a.B; m += 32) {
      const float g = __ldcg(gn + m);
      s1 += g;
      s2 = fmaf(g, (__ldcg(lr + m) - mean) * rstd, s2);
    }
    gbeta = warp_sum(s1);
    ggamma = warp_sum(s2);
  }
  const float invB = 1.0f / (float)a.B;
  float s3 = 0.f;
  for (int m = lane; m < a.B; m += 32) {
    const float g = __ldcg(gn + m), l = __ldcg(lr + m);
    float gl = g;
    if (a.has_bn) gl = gam * rstd * (g - gbeta * invB - (l - mean) * rstd * ggamma * invB);
    const float v = gl * (l > 0.f ? 1.f : kLreluSlope);
    __stcg(gp + m, v);
    s3 += v;
  }
  s3 = warp_sum(s3);
  if (lane == 0) {
    if (a.has_bn) {
      c.G(which == 1 ? P_G1 : P_G2)[n] = ggamma;
      c.G(which == 1 ? P_BE1 : P_BE2)[n] = gbeta;
    }
    c.G(which == 1 ? P_B1 : P_B2)[n] = s3;
  }
}

// Weight gradients gW[i][j] = sum_m g^T[i][m] * act^T[j][m]  (which = 3: fc.weight, 2: net.3, 1: net.0).
// GEMM rows = j (the contiguous axis of gW -> float4 stores), GEMM columns = i, GEMM k = sample.
template <int BN>
__device__ __noinline__ void op_wgrad(const KArgs& a, int e, const SubOp& so, int t, float* smem, const TcState& tc) {
  constexpr int TN = BN / 16;
  Ctx c(a, e);
  const int tn = t % so.tiles_n, tm = t / so.tiles_n;
  const int which = so.which;
  const int I = which == 3 ? a.Dout : a.H;
  const int J = which == 1 ? a.Da : a.H;
  const float* act = which == 1 ? c.yprev() : c.Nrm(which - 1);
  const float* gsrc = a.W + (which == 3 ? a.oGPRET : which == 2 ? a.oGP2T : a.oGP1T);
  OperandA A{act, a.B, which == 1 ? c.gmap : nullptr, tm * kTM, J, a.B, c.vec};
  OperandB B{gsrc, a.B, nullptr, tn * BN, I, a.B, c.vec, 0, 0};
  float acc[4][TN];
  zero_acc<TN>(acc);
  if (a.use_tc) tile_gemm_tc<BN, A_KCONTIG, B_KCONTIG>(acc, A, B, 0, a.B, smem, tc);
  else tile_gemm<BN, A_KCONTIG, B_KCONTIG>(acc, A, B, 0, a.B, smem);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int j0 = tm * kTM + tx * 4;
  const int nvalid = J - j0;
  if (nvalid <= 0) return;
  float* gw = c.G(which == 3 ? P_W3 : which == 2 ? P_W2 : P_W1);
  const bool vecw = (J & 3) == 0;
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int i = tn * BN + ty + 16 * j;
    if (i >= I) continue;
    const float v[4] = {acc[0][j], acc[1][j], acc[2][j], acc[3][j]};
    float* p = gw + (size_t)i * J + j0;
    if (vecw && nvalid >= 4) {
      *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (u < nvalid) p[u] = v[u];
    }
  }
}

// dgrad through net.0.weight, plus the direct path through the a-half and ActNorm, written through
// the fused permutation into the rows of the gradient wrt the previous stage's output.
template <int BN>
__device__ __noinline__ void op_dgrad_in(const KArgs& a, int e, const SubOp& so, int t, float* smem, const TcState& tc) {
  constexpr int TN = BN / 16;
  Ctx c(a, e);
  const int tn = t % so.tiles_n, tm = t / so.tiles_n;
  OperandA A{a.W + a.oGP1T, a.B, nullptr, tm * kTM, a.B, a.H, c.vec};
  OperandB B{c.P(P_W1), a.Da, nullptr, tn * BN, a.Da, a.H, (a.Da & 3) == 0, 0, 0};
  float acc[4][TN];
  zero_acc<TN>(acc);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m = tm * kTM + tx * 4;
  const int nvalid = a.B - m;
  const float e_lsi = expf(__ldg(a.P + c.st.lsi));
  float p_gy[TN][4];
  int p_row[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) {   // independent of the GEMM: in flight while it runs
    const int n = tn * BN + ty + 16 * j;
    const bool ok = n < a.Da && nvalid > 0;
    p_row[j] = ok ? __ldg(c.gmap + n) : 0;
    load4(c.gy() + (size_t)(ok ? n : 0) * a.B + m, p_gy[j], ok ? nvalid : 0, c.vec);
  }
  if (a.use_tc) tile_gemm_tc<BN, A_MCONTIG, B_NCONTIG>(acc, A, B, 0, a.H, smem, tc);
  else tile_gemm<BN, A_MCONTIG, B_NCONTIG>(acc, A, B, 0, a.H, smem);
  if (nvalid <= 0) return;
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int n = tn * BN + ty + 16 * j;
    if (n >= a.Da) continue;
    float gu[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) gu[i] = fmaf(p_gy[j][i], e_lsi, acc[i][j]);
    store4(c.gyprev() + (size_t)p_row[j] * a.B + m, gu, nvalid, c.vec);
  }
}

// ActNorm.reverse gradients: g_loc = -sum g_y, g_lsi = sum g_y * (y + loc) + D * sum_m g_ld[m].
// One warp per stage sums that stage's per-tile partials (lane-strided, then a fixed shuffle tree).
__device__ __noinline__ void op_bwd_final(const KArgs& a, float* sRed) {
  float s = 0.f;
  for (int m = threadIdx.x; m < a.B; m += kThreads) s += ldcg(a.g_ld + m);
  const float gld = block_sum(s, sRed);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int e = a.fin_lo + w; e < a.fin_hi; e += kThreads / 32) {
    const float* red = a.W + a.oRED + (size_t)e * a.red_tiles * 2;
    float s1 = 0.f, s2 = 0.f;
    for (int t = lane; t < a.red_tiles; t += 32) { s1 += ldcg(red + 2 * t); s2 += ldcg(red + 2 * t + 1); }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
      const StageTab& st = a.st[e];
      const float bad = (__ldcg(a.ctrl + 2) ? __int_as_float(0x7fc00000) : 0.f);   // timed-out grid barrier -> NaN gradients
      a.G[st.loc] = -s1 + bad;
      a.G[st.lsi] = s2 + (float)a.D * gld + bad;
    }
  }
}

// ----------------------------------------------------------------------------- dispatcher
template <int BN>
__device__ __forceinline__ void run_gemm_op(const KArgs& a, int e, const SubOp& so, int t, float* smem, int* sFlag,
                                            const TcState& tc) {
  switch (so.op) {
    case OP_FWD_HID: op_fwd_hid<BN>(a, e, so, t, smem, sFlag, tc); break;
    case OP_DGRAD_HID: op_dgrad_hid<BN>(a, e, so, t, smem, sFlag, tc); break;
    case OP_WGRAD: op_wgrad<BN>(a, e, so, t, smem, tc); break;
    case OP_DGRAD_IN: op_dgrad_in<BN>(a, e, so, t, smem, tc); break;
    default: break;
  }
}

constexpr int kSmemPhases = 136;   // programs up to this many phases are staged in shared memory
constexpr int kSmemStages = 64;

__global__ void __launch_bounds__(kThreads) flow_program_kernel(KArgs a) {
  extern __shared__ __align__(16) float smem[];   // kTileSmemBytes: cp.async ring (also transpose scratch)
  __shared__ float sRed[1024 + 128];
  __shared__ int sFlag;
  __shared__ __align__(16) Phase sPh[kSmemPhases];
  __shared__ __align__(16) StageTab sSt[kSmemStages];
  __shared__ unsigned sTmemSlot;
  TcState tc;
  tc.opbuf = reinterpret_cast<unsigned char*>(smem) + kTileSmemBytes;
  tc.bars = reinterpret_cast<unsigned long long*>(tc.opbuf + 2 * kOpBufBytes);
  tc.tmem = 0;
  if (a.use_tc) {   // TMEM accumulator for the tcgen05 tile core: allocated once per CTA by warp 0
    if (threadIdx.x < 32) tmem_alloc(&sTmemSlot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    tc.tmem = sTmemSlot;
  }
  const int np = a.phase_end - a.phase_begin;
  const bool ph_smem = np <= kSmemPhases;
  if (ph_smem) {
    const int4* src = reinterpret_cast<const int4*>(a.phases + a.phase_begin);
    int4* dst = reinterpret_cast<int4*>(sPh);
    for (int i = threadIdx.x; i < np * (int)(sizeof(Phase) / 16); i += kThreads) dst[i] = __ldg(src + i);
  }
  if (a.S <= kSmemStages) {
    const int4* src = reinterpret_cast<const int4*>(a.st);
    int4* dst = reinterpret_cast<int4*>(sSt);
    for (int i = threadIdx.x; i < a.S * (int)(sizeof(StageTab) / 16); i += kThreads) dst[i] = __ldg(src + i);
    a.st = sSt;
  }
  __syncthreads();
  unsigned bar = 0;
  for (int p = a.phase_begin; p < a.phase_end; ++p) {
    const Phase& ph = ph_smem ? sPh[p - a.phase_begin] : a.phases[p];
    if (a.tbuf && blockIdx.x == 0 && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      a.tbuf[2 * p] = t;
    }
    const int n0 = ph.sub[0].ntiles;
    const int total = n0 + (ph.nsub > 1 ? ph.sub[1].ntiles : 0);
    for (int tt = blockIdx.x; tt < total; tt += gridDim.x) {
      const SubOp& so = tt < n0 ? ph.sub[0] : ph.sub[1];
      const int t = tt < n0 ? tt : tt - n0;
      switch (so.op) {
        case OP_FWD_OUT:
          if (so.bn == 64) op_fwd_out<64>(a, ph.stage, so, t, smem, sRed, tc);
          else op_fwd_out<32>(a, ph.stage, so, t, smem, sRed, tc);
          break;
        case OP_TRANSPOSE: op_transpose(a, so, t, smem); break;
        case OP_BN_STATS: op_bn_stats(a, ph.stage, so, t); break;
        case OP_LOGDET_FINAL: op_logdet_final(a, t, sRed); break;
        case OP_B0: op_b0(a, ph.stage, so, t, sRed); break;
        case OP_BNBWD: op_bnbwd(a, ph.stage, so, t); break;
        case OP_BWD_FINAL: op_bwd_final(a, sRed); break;
        case OP_TAIL: op_tail(a, ph.stage, so, t); break;
        default:
          if (so.bn == 64) run_gemm_op<64>(a, ph.stage, so, t, smem, &sFlag, tc);
          else if (so.bn == 32) run_gemm_op<32>(a, ph.stage, so, t, smem, &sFlag, tc);
          else run_gemm_op<16>(a, ph.stage, so, t, smem, &sFlag, tc);
          break;
      }
      __syncthreads();
    }
    if (a.tbuf && blockIdx.x == 0 && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      a.tbuf[2 * p + 1] = t;
    }
    if (a.persistent && p + 1 < a.phase_end) grid_sync(a.ctrl, gridDim.x, bar++);
  }
  if (a.use_tc) {
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tc.tmem);
  }
}

// One phase of strip ops only (no GEMM tiles): the launches between the tcgen05 GEMMs of the tensor-core path.
// Small static shared memory and a register cap, so several CTAs are resident per SM (these ops are HBM-bound).
__global__ void __launch_bounds__(kThreads, 3) flow_strip_kernel(KArgs a) {
  __shared__ float sRed[1024 + 128];
  __shared__ float sT[32 * 33];
  const Phase& ph = a.phases[a.phase_begin];
  const int n0 = ph.sub[0].ntiles;
  const int total = n0 + (ph.nsub > 1 ? ph.sub[1].ntiles : 0);
  for (int tt = blockIdx.x; tt < total; tt += gridDim.x) {
    const SubOp& so = tt < n0 ? ph.sub[0] : ph.sub[1];
    const int t = tt < n0 ? tt : tt - n0;
    switch (so.op) {
      case OP_TRANSPOSE: op_transpose(a, so, t, sT); break;
      case OP_BN_STATS: op_bn_stats(a, ph.stage, so, t); break;
      case OP_LOGDET_FINAL: op_logdet_final(a, t, sRed); break;
      case OP_B0: op_b0(a, ph.stage, so, t, sRed); break;
      case OP_BNBWD: op_bnbwd(a, ph.stage, so, t); break;
      case OP_BWD_FINAL: op_bwd_final(a, sRed); break;
      case OP_TAIL: op_tail(a, ph.stage, so, t); break;
      default: break;
    }
    __syncthreads();
  }
}

void* flow_kernel_ptr() { return (void*)flow_program_kernel; }
void* flow_strip_kernel_ptr() { return (void*)flow_strip_kernel; }

}  // namespace dpi
