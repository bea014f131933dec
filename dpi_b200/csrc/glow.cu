// glow.cu - the non-GEMM pieces of Glow.reverse (generative_model/glow_model.py), forward only, on feature-major
// image stacks [channel][b * h * w] (position = b * h * w + y * w + x):
//   * BatchNorm2d over a channel row (:263,266; batch statistics + running-statistics update, or running statistics)
//   * InvConv2dLU.calc_weight (:217-224) and its inverse (:229), one CTA, Gauss-Jordan with partial pivoting
//   * AffineCoupling.reverse tail (:299-311): in_b = out_b / exp(tanh(log_s0)) - t, logdet -= sum tanh
//   * the split prior of Block.reverse (:432-443): z = mean + exp(log_sd) * eps, logdet += sum log_sd
//   * unsqueeze (:453-459) and the scalar logdet terms of InvConv2dLU / ActNorm (:231, :141-143)
// The convolutions are conv_tc.cu. Per-sample logdet sums have a single writer per sample and a fixed order.
#include "common.cuh"

namespace dpi {
namespace {

constexpr int kT = 256;
constexpr int kMaxC = 64;   // channels of an invertible 1x1 convolution handled in shared memory

// one CTA per channel row
__global__ void __launch_bounds__(kT) bn_rows_kernel(int M, long long ld, const float* __restrict__ x, const float* __restrict__ gamma,
                                                    const float* __restrict__ beta, float* __restrict__ rmean,
                                                    float* __restrict__ rvar, int train, float eps, float* __restrict__ out) {
  __shared__ float red[32];
  const int n = blockIdx.x;
  const float* row = x + (long long)n * ld;
  float* orow = out + (long long)n * ld;
  float mean, var;
  if (train) {
    float s = 0.f;
    for (int m = threadIdx.x; m < M; m += kT) s += __ldcg(row + m);
    mean = block_sum(s, red) / (float)M;
    float q = 0.f;
    for (int m = threadIdx.x; m < M; m += kT) { const float d = __ldcg(row + m) - mean; q = fmaf(d, d, q); }
    var = block_sum(q, red) / (float)M;
    if (threadIdx.x == 0) {
      rmean[n] = (1.f - kBnMomentum) * rmean[n] + kBnMomentum * mean;
      rvar[n] = (1.f - kBnMomentum) * rvar[n] + kBnMomentum * (var * (float)M / (float)(M - 1));
    }
  } else {
    mean = rmean[n];
    var = rvar[n];
  }
  const float ga = __ldg(gamma + n) / sqrtf(var + eps), be = __ldg(beta + n);
  for (int m = threadIdx.x; m < M; m += kT) __stcg(orow + m, fmaf(__ldcg(row + m) - mean, ga, be));
}

// W = P (L * l_mask + I) (U * u_mask + diag(sign * exp(w_s))), then W^-1 (row-major [C][C]); one CTA
__global__ void __launch_bounds__(kT) lu_inverse_kernel(int C, const float* __restrict__ w_p, const float* __restrict__ w_l,
                                                       const float* __restrict__ w_s, const float* __restrict__ w_u,
                                                       const float* __restrict__ l_mask, const float* __restrict__ u_mask,
                                                       const float* __restrict__ s_sign, const float* __restrict__ l_eye,
                                                       float* __restrict__ winv, int invert) {
  __shared__ float Wm[kMaxC][2 * kMaxC + 1];   // [W | I] -> [I | W^-1]
  __shared__ float fac[kMaxC];
  __shared__ int piv;
  const int tid = threadIdx.x;
  auto Lm = [&](int r, int c) { return w_l[r * C + c] * l_mask[r * C + c] + l_eye[r * C + c]; };
  auto Um = [&](int r, int c) { return w_u[r * C + c] * u_mask[r * C + c] + (r == c ? s_sign[r] * expf(w_s[r]) : 0.f); };
  for (int i = tid; i < C * C; i += kT) {      // T = P L'   (parked in the right half)
    const int r = i / C, c = i - r * C;
    float s = 0.f;
    for (int k = 0; k < C; ++k) s = fmaf(w_p[r * C + k], Lm(k, c), s);
    Wm[r][kMaxC + c] = s;
  }
  __syncthreads();
  for (int i = tid; i < C * C; i += kT) {      // W = T U'
    const int r = i / C, c = i - r * C;
    float s = 0.f;
    for (int k = 0; k < C; ++k) s = fmaf(Wm[r][kMaxC + k], Um(k, c), s);
    Wm[r][c] = s;
  }
  __syncthreads();
  if (!invert) {                               // InvConv2dLU.forward uses W itself
    for (int i = tid; i < C * C; i += kT) winv[i] = Wm[i / C][i % C];
    return;
  }
  // from here the right half starts at column C (not kMaxC): move nothing, just overwrite with the identity
  for (int i = tid; i < C * C; i += kT) {
    const int r = i / C, c = i - r * C;
    Wm[r][C + c] = r == c ? 1.f : 0.f;
  }
  __syncthreads();
  for (int col = 0; col < C; ++col) {
    if (tid == 0) {
      int best = col;
      float bv = fabsf(Wm[col][col]);
      for (int r = col + 1; r < C; ++r)
        if (fabsf(Wm[r][col]) > bv) { bv = fabsf(Wm[r][col]); best = r; }
      piv = best;
    }
    __syncthreads();
    const int p = piv;
    if (p != col)
      for (int c = tid; c < 2 * C; c += kT) { const float t = Wm[col][c]; Wm[col][c] = Wm[p][c]; Wm[p][c] = t; }
    __syncthreads();
    const float inv = 1.0f / Wm[col][col];
    __syncthreads();
    for (int c = tid; c < 2 * C; c += kT) Wm[col][c] *= inv;
    for (int r = tid; r < C; r += kT) fac[r] = r == col ? 0.f : Wm[r][col];   // factors read before any row changes
    __syncthreads();
    for (int i = tid; i < C * 2 * C; i += kT) {
      const int r = i / (2 * C), c = i - r * 2 * C;
      Wm[r][c] = fmaf(-fac[r], Wm[col][c], Wm[r][c]);
    }
    __syncthreads();
  }
  for (int i = tid; i < C * C; i += kT) {
    const int r = i / C, c = i - r * C;
    winv[i] = Wm[r][C + c];
  }
}

// one CTA per sample: x rows [C/2, C) <- out_b / exp(tanh(log_s0)) - t; logdet[b] -= sum tanh
__global__ void __launch_bounds__(kT) affine_reverse_kernel(int hw, int C, const float* __restrict__ net, long long ldn,
                                                           float* __restrict__ x, long long ldx, float* __restrict__ logdet) {
  __shared__ float red[32];
  const int b = blockIdx.x, half = C / 2;
  float ld = 0.f;
  for (int p = threadIdx.x; p < hw; p += kT) {
    const long long m = (long long)b * hw + p;
    for (int j = 0; j < half; ++j) {
      const float ls = tanhf(__ldcg(net + (long long)j * ldn + m));
      const float t = __ldcg(net + (long long)(half + j) * ldn + m);
      float* xb = x + (long long)(half + j) * ldx + m;
      *xb = *xb / expf(ls) - t;
      ld -= ls;
    }
  }
  const float tot = block_sum(ld, red);
  if (threadIdx.x == 0) logdet[b] += tot;
}

// one CTA per sample: z = mean + exp(log_sd) * eps; logdet[b] += sum log_sd
__global__ void __launch_bounds__(kT) prior_sample_kernel(int hw, int C2, const float* __restrict__ ml, long long ldm,
                                                         const float* __restrict__ eps, long long lde, float* __restrict__ z,
                                                         long long ldz, float* __restrict__ logdet) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  float ld = 0.f;
  for (int p = threadIdx.x; p < hw; p += kT) {
    const long long m = (long long)b * hw + p;
    for (int j = 0; j < C2; ++j) {
      const float mean = __ldcg(ml + (long long)j * ldm + m), lsd = __ldcg(ml + (long long)(C2 + j) * ldm + m);
      z[(long long)j * ldz + m] = fmaf(expf(lsd), __ldcg(eps + (long long)j * lde + m), mean);
      ld += lsd;
    }
  }
  const float tot = block_sum(ld, red);
  if (threadIdx.x == 0) logdet[b] += tot;
}

// out[c][b, 2y + dy, 2x + dx] = in[4c + 2dy + dx][b, y, x]
__global__ void unsqueeze_kernel(int batch, int h, int w, int C, const float* __restrict__ in, float* __restrict__ out) {
  const long long Mi = (long long)batch * h * w, Mo = 4 * Mi, total = (long long)(C / 4) * Mo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i / Mo);
    const long long r = i - (long long)c * Mo;
    const int b = (int)(r / (4LL * h * w));
    const int q = (int)(r - (long long)b * 4 * h * w), Y = q / (2 * w), X = q - Y * 2 * w;
    const int y = Y >> 1, dy = Y & 1, x = X >> 1, dx = X & 1;
    out[i] = __ldcg(in + (long long)(4 * c + 2 * dy + dx) * Mi + (long long)b * h * w + y * w + x);
  }
}

// logdet[b] += hw * (a_ws * sum w_s + a_si * sum log|scale_inv|)
__global__ void __launch_bounds__(kT) logdet_scalar_kernel(int batch, int hw, int C, const float* __restrict__ w_s, float a_ws,
                                                          const float* __restrict__ scale_inv, float a_si,
                                                          float* __restrict__ logdet) {
  __shared__ float red[32];
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += kT) {
    if (w_s) s += a_ws * w_s[c];
    if (scale_inv) s += a_si * logf(fabsf(scale_inv[c]));
  }
  const float tot = (float)hw * block_sum(s, red);
  for (int b = threadIdx.x; b < batch; b += kT) logdet[b] += tot;
}

// one CTA per sample: x rows [C/2, C) <- (in_b + t) * exp(tanh(log_s0)); logdet[b] += sum tanh   (:277-297)
__global__ void __launch_bounds__(kT) affine_forward_kernel(int hw, int C, const float* __restrict__ net, long long ldn,
                                                           float* __restrict__ x, long long ldx, float* __restrict__ logdet) {
  __shared__ float red[32];
  const int b = blockIdx.x, half = C / 2;
  float ld = 0.f;
  for (int p = threadIdx.x; p < hw; p += kT) {
    const long long m = (long long)b * hw + p;
    for (int j = 0; j < half; ++j) {
      const float ls = tanhf(__ldcg(net + (long long)j * ldn + m));
      const float t = __ldcg(net + (long long)(half + j) * ldn + m);
      float* xb = x + (long long)(half + j) * ldx + m;
      *xb = (*xb + t) * expf(ls);
      ld += ls;
    }
  }
  const float tot = block_sum(ld, red);
  if (threadIdx.x == 0) logdet[b] += tot;
}

// one CTA per sample: z_eps = (z - mean) * exp(-log_sd); logdet[b] -= sum log_sd; log_p[b] += sum(-0.5 log 2 pi - 0.5 z_eps^2)
__global__ void __launch_bounds__(kT) prior_forward_kernel(int hw, int C2, const float* __restrict__ ml, long long ldm,
                                                          const float* __restrict__ z, long long ldz, float* __restrict__ z_eps,
                                                          long long lde, float* __restrict__ logdet, float* __restrict__ log_p) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  float ld = 0.f, lp = 0.f;
  for (int p = threadIdx.x; p < hw; p += kT) {
    const long long m = (long long)b * hw + p;
    for (int j = 0; j < C2; ++j) {
      const float mean = __ldcg(ml + (long long)j * ldm + m), lsd = __ldcg(ml + (long long)(C2 + j) * ldm + m);
      const float e = (__ldcg(z + (long long)j * ldz + m) - mean) * expf(-lsd);
      z_eps[(long long)j * lde + m] = e;
      ld += lsd;
      lp += -0.918938533204672742f - 0.5f * e * e;
    }
  }
  const float t1 = block_sum(ld, red);
  const float t2 = block_sum(lp, red);
  if (threadIdx.x == 0) { logdet[b] -= t1; log_p[b] += t2; }
}

// out[4c + 2dy + dx][b, y, x] = in[c][b, 2y + dy, 2x + dx]   (h, w: OUTPUT height / width)
__global__ void squeeze_kernel(int batch, int h, int w, int C, const float* __restrict__ in, float* __restrict__ out) {
  const long long Mo = (long long)batch * h * w, total = (long long)4 * C * Mo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i / Mo);
    const long long r = i - (long long)co * Mo;
    const int b = (int)(r / (h * w)), q = (int)(r - (long long)b * h * w), y = q / w, x = q - y * w;
    const int c = co >> 2, dy = (co >> 1) & 1, dx = co & 1;
    out[i] = __ldcg(in + (long long)c * 4 * Mo + (long long)b * 4 * h * w + (2 * y + dy) * (2 * w) + 2 * x + dx);
  }
}

// out[c][m] = (x[c][m] + add[c]) / div[c]   (ActNorm.forward :113-130)
__global__ void row_affine_kernel(int C, long long M, const float* __restrict__ x, const float* __restrict__ add,
                                  const float* __restrict__ div, float* __restrict__ out) {
  const long long total = (long long)C * M;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i / M);
    out[i] = (1.0f / __ldg(div + c)) * (__ldcg(x + i) + __ldg(add + c));
  }
}


// ================================================================================= training backward (sampling direction)
// Formulas: oracle/glow_backward_spec.py (each checked against autograd through the oracle, which is bit-identical to the
// reference). All stacks are feature-major [channel][b * hw + p]; per-channel reductions are one CTA per row with a
// fixed-order block reduction (bit-reproducible).

// ActNorm.reverse x = v * si - loc (glow_model.py:132-148), given y = x: v = (y + loc) / si.
// g_v = g * si; g_loc = -sum g; g_si = sum g v + hw_gld / si   (hw_gld = h w sum_b g_logdet[b])
__global__ void __launch_bounds__(kT) actnorm_reverse_bwd_kernel(long long M, const float* __restrict__ y, const float* __restrict__ loc,
                                                                const float* __restrict__ si, const float* __restrict__ g,
                                                                const float* __restrict__ hw_gld, float* __restrict__ g_v,
                                                                float* __restrict__ g_loc, float* __restrict__ g_si) {
  __shared__ float red[32];
  const int c = blockIdx.x;
  const float l = loc[c], s = si[c], inv = 1.0f / s;
  float a0 = 0.f, a1 = 0.f;
  for (long long m = threadIdx.x; m < M; m += kT) {
    const float gv = __ldcg(g + (long long)c * M + m);
    a0 += gv;
    a1 = fmaf(gv, (__ldcg(y + (long long)c * M + m) + l) * inv, a1);
    g_v[(long long)c * M + m] = gv * s;
  }
  const float t0 = block_sum(a0, red);
  const float t1 = block_sum(a1, red);
  if (threadIdx.x == 0) { g_loc[c] = -t0; g_si[c] = t1 + hw_gld[0] * inv; }
}

// AffineCoupling.reverse tail (:299-320): x_b = b / exp(tanh(ls0)) - t, logdet -= sum tanh(ls0); out = [ls0 ; t], xm = x_b.
// b exp(-ls) = x_b + t.  g_out[j] = (-g_b (x_b + t) - g_ld[b]) (1 - ls^2); g_out[half + j] = -g_b; g rows of the b half become
// g_b exp(-ls) in place (the a half is left to the caller).
__global__ void affine_reverse_bwd_kernel(int batch, int hw, int C, const float* __restrict__ net, long long ldn,
                                          const float* __restrict__ xm, long long ldx, float* __restrict__ g, long long ldg,
                                          const float* __restrict__ g_ld, float* __restrict__ g_out, long long ldo) {
  const int half = C / 2;
  const long long M = (long long)batch * hw, total = (long long)half * M;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i / M);
    const long long m = i - (long long)j * M;
    const float ls = tanhf(__ldcg(net + (long long)j * ldn + m)), t = __ldcg(net + (long long)(half + j) * ldn + m);
    const float gb = g[(long long)(half + j) * ldg + m];
    const float xb = __ldcg(xm + (long long)(half + j) * ldx + m);
    g_out[(long long)j * ldo + m] = (-gb * (xb + t) - __ldg(g_ld + m / hw)) * (1.f - ls * ls);
    g_out[(long long)(half + j) * ldo + m] = -gb;
    g[(long long)(half + j) * ldg + m] = gb * expf(-ls);
  }
}

// ZeroConv2d epilogue out = zc * exp(3 scale) (:246-251): g_zc = g_out e3 (in place); g_scale = 3 sum g_out out; g_bias = sum g_zc
__global__ void __launch_bounds__(kT) zeroconv_bwd_kernel(long long M, float* __restrict__ g, const float* __restrict__ out,
                                                         const float* __restrict__ scale, float* __restrict__ g_scale,
                                                         float* __restrict__ g_bias) {
  __shared__ float red[32];
  const int c = blockIdx.x;
  const float e3 = expf(3.f * scale[c]);
  float a0 = 0.f, a1 = 0.f;
  for (long long m = threadIdx.x; m < M; m += kT) {
    const float gv = g[(long long)c * M + m];
    a0 = fmaf(gv, __ldcg(out + (long long)c * M + m), a0);
    const float gz = gv * e3;
    a1 += gz;
    g[(long long)c * M + m] = gz;
  }
  const float t0 = block_sum(a0, red);
  const float t1 = block_sum(a1, red);
  if (threadIdx.x == 0) { g_scale[c] = 3.f * t0; g_bias[c] = t1; }
}

// Conv -> LeakyReLU -> BatchNorm2d (training): l = LeakyReLU output (the BatchNorm input). In place on g:
// g <- LeakyReLU'(l) * gamma rstd (g - mean(g) - xh mean(g xh)); g_gamma = sum g xh; g_beta = sum g; g_bias = sum of the result
__global__ void __launch_bounds__(kT) bn_lrelu_bwd_kernel(long long M, float* __restrict__ g, const float* __restrict__ l,
                                                         const float* __restrict__ gamma, float eps, float slope,
                                                         float* __restrict__ g_gamma, float* __restrict__ g_beta,
                                                         float* __restrict__ g_bias) {
  __shared__ float red[32];
  const int c = blockIdx.x;
  const float* lr = l + (long long)c * M;
  float* gr = g + (long long)c * M;
  float a = 0.f;
  for (long long m = threadIdx.x; m < M; m += kT) a += __ldcg(lr + m);
  const float mean = block_sum(a, red) / (float)M;
  a = 0.f;
  for (long long m = threadIdx.x; m < M; m += kT) { const float d = __ldcg(lr + m) - mean; a = fmaf(d, d, a); }
  const float rstd = 1.0f / sqrtf(block_sum(a, red) / (float)M + eps);
  float a0 = 0.f, a1 = 0.f;
  for (long long m = threadIdx.x; m < M; m += kT) {
    const float gv = gr[m];
    a0 += gv;
    a1 = fmaf(gv, (__ldcg(lr + m) - mean) * rstd, a1);
  }
  const float sb = block_sum(a0, red);
  const float sg = block_sum(a1, red);
  const float k = gamma[c] * rstd, mb = sb / (float)M, mg = sg / (float)M;
  float ab = 0.f;
  for (long long m = threadIdx.x; m < M; m += kT) {
    const float lv = __ldcg(lr + m);
    const float o = k * (gr[m] - mb - (lv - mean) * rstd * mg) * (lv > 0.f ? 1.f : slope);
    gr[m] = o;
    ab += o;
  }
  const float tb = block_sum(ab, red);
  if (threadIdx.x == 0) { g_gamma[c] = sg; g_beta[c] = sb; g_bias[c] = tb; }
}

// out[c] = coef * sum_m a[c][m]
__global__ void __launch_bounds__(kT) row_sum_kernel(long long M, const float* __restrict__ a, float coef, float* __restrict__ out) {
  __shared__ float red[32];
  const int c = blockIdx.x;
  float s = 0.f;
  for (long long m = threadIdx.x; m < M; m += kT) s += __ldcg(a + (long long)c * M + m);
  const float t = block_sum(s, red);
  if (threadIdx.x == 0) out[c] = coef * t;
}

// InvConv2dLU.reverse (:215-234): x = V u, V = W^-1, W = P Lm Um, Lm = w_l l_mask + I, Um = w_u u_mask + diag(d), d = s_sign exp(w_s).
// Given G_V = sum_m g[c][m] u[c'][m]: G_W = -V^T G_V V^T; G_Lm = P^T G_W Um^T; G_Um = (P Lm)^T G_W;
// g_wl = G_Lm l_mask; g_wu = G_Um u_mask; g_ws = diag(G_Um) d - hw_gld. One CTA, C <= kMaxC.
__global__ void __launch_bounds__(kT) lu_bwd_kernel(int C, const float* __restrict__ w_p, const float* __restrict__ w_l,
                                                   const float* __restrict__ w_s, const float* __restrict__ w_u,
                                                   const float* __restrict__ l_mask, const float* __restrict__ u_mask,
                                                   const float* __restrict__ s_sign, const float* __restrict__ l_eye,
                                                   const float* __restrict__ V, const float* __restrict__ GV,
                                                   const float* __restrict__ hw_gld, float* __restrict__ scratch,
                                                   float* __restrict__ g_wl, float* __restrict__ g_wu, float* __restrict__ g_ws) {
  // scratch: 4 C x C matrices in global memory (C <= 64: 64 KB, L2-resident): T1, GW, PL, T2
  float* T1 = scratch;
  float* GW = scratch + (size_t)C * C;
  float* PL = GW + (size_t)C * C;
  const int tid = threadIdx.x, n = C * C;
  auto Lm = [&](int r, int c) { return w_l[r * C + c] * l_mask[r * C + c] + l_eye[r * C + c]; };
  auto Um = [&](int r, int c) { return w_u[r * C + c] * u_mask[r * C + c] + (r == c ? s_sign[r] * expf(w_s[r]) : 0.f); };
  for (int i = tid; i < n; i += kT) {          // T1 = V^T G_V ; PL = P Lm
    const int r = i / C, c = i - r * C;
    float s = 0.f, q = 0.f;
    for (int k = 0; k < C; ++k) { s = fmaf(V[k * C + r], GV[k * C + c], s); q = fmaf(w_p[r * C + k], Lm(k, c), q); }
    T1[i] = s; PL[i] = q;
  }
  __syncthreads();
  for (int i = tid; i < n; i += kT) {          // G_W = -T1 V^T
    const int r = i / C, c = i - r * C;
    float s = 0.f;
    for (int k = 0; k < C; ++k) s = fmaf(T1[r * C + k], V[c * C + k], s);
    GW[i] = -s;
  }
  __syncthreads();
  for (int i = tid; i < n; i += kT) {          // T1 <- P^T G_W
    const int r = i / C, c = i - r * C;
    float s = 0.f;
    for (int k = 0; k < C; ++k) s = fmaf(w_p[k * C + r], GW[k * C + c], s);
    T1[i] = s;
  }
  __syncthreads();
  for (int i = tid; i < n; i += kT) {          // G_Lm = T1 Um^T ; G_Um = PL^T G_W
    const int r = i / C, c = i - r * C;
    float s = 0.f, q = 0.f;
    for (int k = 0; k < C; ++k) { s = fmaf(T1[r * C + k], Um(c, k), s); q = fmaf(PL[k * C + r], GW[k * C + c], q); }
    g_wl[i] = s * l_mask[i];
    g_wu[i] = q * u_mask[i];
    if (r == c) g_ws[r] = q * s_sign[r] * expf(w_s[r]) - hw_gld[0];
  }
}

// Split prior of Block.reverse (:436-452): z = mean + exp(lsd) eps, logdet += sum lsd; ml = [mean ; lsd].
// g_ml = [g_z ; g_z exp(lsd) eps + g_ld[b]]; g_eps = g_z exp(lsd)
__global__ void prior_sample_bwd_kernel(int batch, int hw, int C2, const float* __restrict__ ml, long long ldm,
                                        const float* __restrict__ eps, long long lde, const float* __restrict__ g_z, long long ldz,
                                        const float* __restrict__ g_ld, float* __restrict__ g_ml, long long ldo,
                                        float* __restrict__ g_eps, long long ldge) {
  const long long M = (long long)batch * hw, total = (long long)C2 * M;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i / M);
    const long long m = i - (long long)j * M;
    const float e = expf(__ldcg(ml + (long long)(C2 + j) * ldm + m)), gz = __ldcg(g_z + (long long)j * ldz + m);
    g_ml[(long long)j * ldo + m] = gz;
    g_ml[(long long)(C2 + j) * ldo + m] = fmaf(gz * e, __ldcg(eps + (long long)j * lde + m), __ldg(g_ld + m / hw));
    g_eps[(long long)j * ldge + m] = gz * e;
  }
}

// out[0] = coef * sum_b v[b]
__global__ void __launch_bounds__(kT) scaled_sum_kernel(int n, const float* __restrict__ v, float coef, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += kT) s += v[i];
  const float t = block_sum(s, red);
  if (threadIdx.x == 0) out[0] = coef * t;
}

}  // namespace
}  // namespace dpi

using namespace dpi;

extern "C" {

int dpi_glow_bn_rows(int channels, int64_t positions, int64_t ld, const float* x, const float* gamma, const float* beta,
                     float* running_mean, float* running_var, int train, float eps, float* out, dpi_stream_t stream) {
  DPI_REQUIRE(x && gamma && beta && running_mean && running_var && out, "dpi_glow_bn_rows: null argument");
  DPI_REQUIRE(channels >= 1 && positions >= 1 && ld >= positions && positions < (1LL << 31), "dpi_glow_bn_rows: bad shape");
  DPI_REQUIRE(!(train && positions < 2), "dpi_glow_bn_rows: Expected more than 1 value per channel when training");
  bn_rows_kernel<<<channels, kT, 0, (cudaStream_t)stream>>>((int)positions, ld, x, gamma, beta, running_mean, running_var,
                                                           train ? 1 : 0, eps, out);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int dpi_glow_lu_inverse(int channels, const float* w_p, const float* w_l, const float* w_s, const float* w_u,
                        const float* l_mask, const float* u_mask, const float* s_sign, const float* l_eye, float* winv,
                        dpi_stream_t stream) {
  DPI_REQUIRE(w_p && w_l && w_s && w_u && l_mask && u_mask && s_sign && l_eye && winv, "dpi_glow_lu_inverse: null argument");
  DPI_REQUIRE(channels >= 1 && channels <= kMaxC, "dpi_glow_lu_inverse: 1 <= channels <= %d (got %d)", kMaxC, channels);
  lu_inverse_kernel<<<1, kT, 0, (cudaStream_t)stream>>>(channels, w_p, w_l, w_s, w_u, l_mask, u_mask, s_sign, l_eye, winv, 1);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int dpi_glow_affine_reverse(int batch, int hw, int channels, const float* net_out, int64_t ld_net, float* x, int64_t ldx,
                            float* logdet, dpi_stream_t stream) {
  DPI_REQUIRE(net_out && x && logdet, "dpi_glow_affine_reverse: null argument");
  DPI_REQUIRE(batch >= 1 && hw >= 1 && channels >= 2 && (channels & 1) == 0, "dpi_glow_affine_reverse: bad shape");
  affine_reverse_kernel<<<batch, kT, 0, (cudaStream_t)stream>>>(hw, channels, net_out, ld_net, x, ldx, logdet);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int dpi_glow_prior_sample(int batch, int hw, int channels, const float* mean_logsd, int64_t ld_ml, const float* eps,
                          int64_t ld_eps, float* z, int64_t ld_z, float* logdet, dpi_stream_t stream) {
  DPI_REQUIRE(mean_logsd && eps && z && logdet, "dpi_glow_prior_sample: null argument");
  DPI_REQUIRE(batch >= 1 && hw >= 1 && channels >= 1, "dpi_glow_prior_sample: bad shape");
  prior_sample_kernel<<<batch, kT, 0, (cudaStream_t)stream>>>(hw, channels, mean_logsd, ld_ml, eps, ld_eps, z, ld_z, logdet);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int dpi_glow_unsqueeze(int batch, int h, int w, int channels, const float* x, float* out, dpi_stream_t stream) {
  DPI_REQUIRE(x && out, "dpi_glow_unsqueeze: null argument");
  DPI_REQUIRE(batch >= 1 && h >= 1 && w >= 1 && channels >= 4 && (channels & 3) == 0, "dpi_glow_unsqueeze: bad shape");
  const long long total = (long long)channels * batch * h * w;
  unsqueeze_kernel<<<(unsigned)((total + 255) / 256 > 65535 ? 65535 : (total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      batch, h, w, channels, x, out);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int dpi_glow_logdet_scalar(int batch, int hw, int channels, const float* w_s, float coef_ws, const float* scale_inv,
                           float coef_si, float* logdet, dpi_stream_t stream) {
  DPI_REQUIRE(logdet && batch >= 1 && hw >= 1 && channels >= 1, "dpi_glow_logdet_scalar: bad argument");
  logdet_scalar_kernel<<<1, kT, 0, (cudaStream_t)stream>>>(batch, hw, channels, w_s, coef_ws, scale_inv, coef_si, logdet);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int dpi_glow_lu_weight(int channels, const float* w_p, const float* w_l, const float* w_s, const float* w_u,
                       const float* l_mask, const float* u_mask, const float* s_sign, const float* l_eye, float* weight,
                       dpi_stream_t stream) {
  DPI_REQUIRE(w_p && w_l && w_s && w_u && l_mask && u_mask && s_sign && l_eye && weight, "dpi_glow_lu_weight: null argument");
  DPI_REQUIRE(channels >= 1 && channels <= kMaxC, "dpi_glow_lu_weight: 1 <= channels <= %d (got %d)", kMaxC, channels);
  lu_inverse_kernel<<<1, kT, 0, (cudaStream_t)stream>>>(channels, w_p, w_l, w_s, w_u, l_mask, u_mask, s_sign, l_eye, weight, 0);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int dpi_glow_affine_forward(int batch, int hw, int channels, const float* net_out, int64_t ld_net, float* x, int64_t ldx,
                            float* logdet, dpi_stream_t stream) {
  DPI_REQUIRE(net_out && x && logdet, "dpi_glow_affine_forward: null argument");
  DPI_REQUIRE(batch >= 1 && hw >= 1 && channels >= 2 && (channels & 1) == 0, "dpi_glow_affine_forward: bad shape");
  affine_forward_kernel<<<batch, kT, 0, (cudaStream_t)stream>>>(hw, channels, net_out, ld_net, x, ldx, logdet);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int dpi_glow_prior_forward(int batch, int hw, int channels, const float* mean_logsd, int64_t ld_ml, const float* z,
                           int64_t ld_z, float* z_eps, int64_t ld_eps, float* logdet, float* log_p, dpi_stream_t stream) {
  DPI_REQUIRE(mean_logsd && z && z_eps && logdet && log_p, "dpi_glow_prior_forward: null argument");
  DPI_REQUIRE(batch >= 1 && hw >= 1 && channels >= 1, "dpi_glow_prior_forward: bad shape");
  prior_forward_kernel<<<batch, kT, 0, (cudaStream_t)stream>>>(hw, channels, mean_logsd, ld_ml, z, ld_z, z_eps, ld_eps, logdet,
                                                              log_p);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int dpi_glow_squeeze(int batch, int h_out, int w_out, int channels_in, const float* x, float* out, dpi_stream_t stream) {
  DPI_REQUIRE(x && out, "dpi_glow_squeeze: null argument");
  DPI_REQUIRE(batch >= 1 && h_out >= 1 && w_out >= 1 && channels_in >= 1, "dpi_glow_squeeze: bad shape");
  const long long total = (long long)4 * channels_in * batch * h_out * w_out;
  squeeze_kernel<<<(unsigned)((total + 255) / 256 > 65535 ? 65535 : (total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      batch, h_out, w_out, channels_in, x, out);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int dpi_glow_row_affine(int channels, int64_t positions, const float* x, const float* add, const float* div, float* out,
                        dpi_stream_t stream) {
  DPI_REQUIRE(x && add && div && out && channels >= 1 && positions >= 1, "dpi_glow_row_affine: bad argument");
  const long long total = (long long)channels * positions;
  row_affine_kernel<<<(unsigned)((total + 255) / 256 > 65535 ? 65535 : (total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      channels, positions, x, add, div, out);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}


int dpi_glow_actnorm_reverse_bwd(int channels, int64_t positions, const float* y, const float* loc, const float* scale_inv,
                                 const float* g, const float* hw_gld, float* g_v, float* g_loc, float* g_scale_inv,
                                 dpi_stream_t stream) {
  DPI_REQUIRE(channels >= 1 && positions >= 1 && y && loc && scale_inv && g && hw_gld && g_v && g_loc && g_scale_inv,
              "dpi_glow_actnorm_reverse_bwd: bad argument");
  actnorm_reverse_bwd_kernel<<<channels, kT, 0, (cudaStream_t)stream>>>(positions, y, loc, scale_inv, g, hw_gld, g_v, g_loc, g_scale_inv);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int dpi_glow_affine_reverse_bwd(int batch, int hw, int channels, const float* net_out, int64_t ld_net, const float* x_mid,
                                int64_t ldx, float* g, int64_t ldg, const float* g_logdet, float* g_out, int64_t ldo,
                                dpi_stream_t stream) {
  DPI_REQUIRE(batch >= 1 && hw >= 1 && channels >= 2 && !(channels & 1) && net_out && x_mid && g && g_logdet && g_out,
              "dpi_glow_affine_reverse_bwd: bad argument");
  const long long total = (long long)(channels / 2) * batch * hw;
  affine_reverse_bwd_kernel<<<(unsigned)std::min<long long>((total + kT - 1) / kT, 148 * 16), kT, 0, (cudaStream_t)stream>>>(
      batch, hw, channels, net_out, ld_net, x_mid, ldx, g, ldg, g_logdet, g_out, ldo);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int dpi_glow_zeroconv_bwd(int channels, int64_t positions, float* g, const float* out, const float* scale, float* g_scale,
                          float* g_bias, dpi_stream_t stream) {
  DPI_REQUIRE(channels >= 1 && positions >= 1 && g && out && scale && g_scale && g_bias, "dpi_glow_zeroconv_bwd: bad argument");
  zeroconv_bwd_kernel<<<channels, kT, 0, (cudaStream_t)stream>>>(positions, g, out, scale, g_scale, g_bias);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int dpi_glow_bn_lrelu_bwd(int channels, int64_t positions, float* g, const float* lrelu_out, const float* gamma, float eps,
                          float slope, float* g_gamma, float* g_beta, float* g_bias, dpi_stream_t stream) {
  DPI_REQUIRE(channels >= 1 && positions >= 2 && g && lrelu_out && gamma && g_gamma && g_beta && g_bias,
              "dpi_glow_bn_lrelu_bwd: bad argument");
  bn_lrelu_bwd_kernel<<<channels, kT, 0, (cudaStream_t)stream>>>(positions, g, lrelu_out, gamma, eps, slope, g_gamma, g_beta, g_bias);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int dpi_glow_row_sum(int channels, int64_t positions, const float* a, float coef, float* out, dpi_stream_t stream) {
  DPI_REQUIRE(channels >= 1 && positions >= 1 && a && out, "dpi_glow_row_sum: bad argument");
  row_sum_kernel<<<channels, kT, 0, (cudaStream_t)stream>>>(positions, a, coef, out);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int dpi_glow_lu_bwd(int channels, const float* w_p, const float* w_l, const float* w_s, const float* w_u, const float* l_mask,
                    const float* u_mask, const float* s_sign, const float* l_eye, const float* w_inv, const float* g_winv,
                    const float* hw_gld, float* scratch, float* g_wl, float* g_wu, float* g_ws, dpi_stream_t stream) {
  DPI_REQUIRE(channels >= 1 && channels <= kMaxC, "dpi_glow_lu_bwd: channels out of range");
  DPI_REQUIRE(w_p && w_l && w_s && w_u && l_mask && u_mask && s_sign && l_eye && w_inv && g_winv && hw_gld && scratch && g_wl &&
                  g_wu && g_ws, "dpi_glow_lu_bwd: null argument");
  lu_bwd_kernel<<<1, kT, 0, (cudaStream_t)stream>>>(channels, w_p, w_l, w_s, w_u, l_mask, u_mask, s_sign, l_eye, w_inv, g_winv,
                                                     hw_gld, scratch, g_wl, g_wu, g_ws);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int dpi_glow_prior_sample_bwd(int batch, int hw, int channels, const float* mean_logsd, int64_t ld_ml, const float* eps,
                              int64_t ld_eps, const float* g_z, int64_t ld_gz, const float* g_logdet, float* g_ml, int64_t ld_gml,
                              float* g_eps, int64_t ld_geps, dpi_stream_t stream) {
  DPI_REQUIRE(batch >= 1 && hw >= 1 && channels >= 1 && mean_logsd && eps && g_z && g_logdet && g_ml && g_eps,
              "dpi_glow_prior_sample_bwd: bad argument");
  const long long total = (long long)channels * batch * hw;
  prior_sample_bwd_kernel<<<(unsigned)std::min<long long>((total + kT - 1) / kT, 148 * 16), kT, 0, (cudaStream_t)stream>>>(
      batch, hw, channels, mean_logsd, ld_ml, eps, ld_eps, g_z, ld_gz, g_logdet, g_ml, ld_gml, g_eps, ld_geps);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int dpi_glow_scaled_sum(int n, const float* v, float coef, float* out, dpi_stream_t stream) {
  DPI_REQUIRE(n >= 1 && v && out, "dpi_glow_scaled_sum: bad argument");
  scaled_sum_kernel<<<1, kT, 0, (cudaStream_t)stream>>>(n, v, coef, out);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

}  // extern "C"
