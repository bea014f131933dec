// image.cu - positivity transform (DPI_interferometry.py:205-211, DPI_MRI.py:174-179) and the
// image priors (interferometry_helpers.py:349-387, MRI_helpers.py:28-38). One CTA per image,
// every prior in one pass, deterministic block reductions.
#include "common.cuh"

namespace dpi {
namespace {

constexpr int kT = 256;

__global__ void __launch_bounds__(kT) positivity_fwd_kernel(int P, const float* __restrict__ x,
                                                            const float* __restrict__ log_scale,
                                                            float* __restrict__ img, float* __restrict__ det) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  const float ls = __ldg(log_scale), es = expf(ls);
  float s = 0.f;
  for (int p = threadIdx.x; p < P; p += kT) {
    const float v = x[(size_t)b * P + p];
    const float sp = softplus_f(v);
    img[(size_t)b * P + p] = sp * es;
    s += v - sp;
  }
  const float tot = block_sum(s, red);
  if (threadIdx.x == 0) det[b] = tot + ls * (float)P;
}

// g_x = (ga + gb) * exp(ls) * sigmoid(x) + g_det[b] * (1 - sigmoid(x))
__global__ void __launch_bounds__(kT) positivity_bwd_kernel(int P, const float* __restrict__ x,
                                                            const float* __restrict__ log_scale,
                                                            const float* __restrict__ img,
                                                            const float* __restrict__ ga, const float* __restrict__ gb,
                                                            const float* __restrict__ g_det, float* __restrict__ g_x,
                                                            float* __restrict__ g_ls_partial) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  const float es = expf(__ldg(log_scale));
  const float gd = g_det ? g_det[b] : 0.f;
  float s = 0.f;
  for (int p = threadIdx.x; p < P; p += kT) {
    const size_t i = (size_t)b * P + p;
    float g = ga ? ga[i] : 0.f;
    if (gb) g += gb[i];
    const float sg = sigmoid_f(x[i]);
    g_x[i] = g * es * sg + gd * (1.f - sg);
    s = fmaf(g, img[i], s);
  }
  const float tot = block_sum(s, red);
  if (threadIdx.x == 0 && g_ls_partial) g_ls_partial[b] = tot + (float)P * gd;
}

__device__ __forceinline__ float sgn(float v) { return (v > 0.f) - (v < 0.f); }

// values[b, 6] and optionally g_img[b, :] = sum_t gw[b, t] * d value_t / d img.
// If x != nullptr the image is first produced from x (softplus * exp(ls)) and det[b] is written.
__global__ void __launch_bounds__(kT) image_priors_kernel(int n, const float* __restrict__ x,
                                                          const float* __restrict__ log_scale,
                                                          float* __restrict__ img_io, float* __restrict__ det,
                                                          const float* __restrict__ prior, float flux, float center,
                                                          float* __restrict__ values, const float* __restrict__ gw,
                                                          float* __restrict__ g_img) {
  extern __shared__ float simg[];  // n*n
  __shared__ float red[32];
  const int b = blockIdx.x, P = n * n;
  float* img = img_io + (size_t)b * P;
  if (x) {
    const float ls = __ldg(log_scale), es = expf(ls);
    float s = 0.f;
    for (int p = threadIdx.x; p < P; p += kT) {
      const float v = x[(size_t)b * P + p];
      const float sp = softplus_f(v);
      const float im = sp * es;
      simg[p] = im;
      img[p] = im;
      s += v - sp;
    }
    const float tot = block_sum(s, red);
    if (threadIdx.x == 0) det[b] = tot + ls * (float)P;
  } else {
    for (int p = threadIdx.x; p < P; p += kT) simg[p] = img[p];
  }
  __syncthreads();

  float s_abs = 0.f, s_tsv = 0.f, s_tv = 0.f, s_x = 0.f, s_xc = 0.f, s_yr = 0.f, s_mem = 0.f;
  for (int p = threadIdx.x; p < P; p += kT) {
    const int r = p / n, c = p - r * n;
    const float v = simg[p];
    s_abs += fabsf(v);
    s_x += v;
    s_xc = fmaf(v, (float)c, s_xc);
    s_yr = fmaf(v, (float)r, s_yr);
    if (r + 1 < n) { const float d = simg[p + n] - v; s_tsv = fmaf(d, d, s_tsv); s_tv += fabsf(d); }
    if (c + 1 < n) { const float d = simg[p + 1] - v; s_tsv = fmaf(d, d, s_tsv); s_tv += fabsf(d); }
    if (prior) s_mem = fmaf(v, logf(v + 1e-12f) - logf(__ldg(prior + p) + 1e-12f), s_mem);
  }
  s_abs = block_sum(s_abs, red);
  s_tsv = block_sum(s_tsv, red);
  s_tv = block_sum(s_tv, red);
  s_x = block_sum(s_x, red);
  s_xc = block_sum(s_xc, red);
  s_yr = block_sum(s_yr, red);
  s_mem = block_sum(s_mem, red);
  const float invP = 1.0f / (float)P, invE = 1.0f / (float)((n - 1) * n);
  // Loss_center: mean(img*X)/mean(img); the 1/P cancels
  const float xc = s_xc / s_x, yc = s_yr / s_x;
  if (threadIdx.x == 0) {
    float* v = values + (size_t)b * DPI_PRIOR_COUNT;
    v[DPI_PRIOR_L1] = s_abs * invP;
    v[DPI_PRIOR_TSV] = s_tsv * invE;
    v[DPI_PRIOR_TV] = s_tv * invE;
    v[DPI_PRIOR_FLUX] = (s_x - flux) * (s_x - flux);
    v[DPI_PRIOR_CENTER] = 0.5f * ((xc - center) * (xc - center) + (yc - center) * (yc - center));
    v[DPI_PRIOR_MEM] = s_mem * invP;
  }
  if (!g_img) return;
  const float* w = gw + (size_t)b * DPI_PRIOR_COUNT;
  const float w_l1 = w[DPI_PRIOR_L1] * invP, w_tsv = w[DPI_PRIOR_TSV] * 2.f * invE, w_tv = w[DPI_PRIOR_TV] * invE;
  const float g_flux = w[DPI_PRIOR_FLUX] * 2.f * (s_x - flux);
  const float w_c = w[DPI_PRIOR_CENTER] / s_x, w_mem = w[DPI_PRIOR_MEM] * invP;
  const bool use_c = w[DPI_PRIOR_CENTER] != 0.f, use_mem = (w[DPI_PRIOR_MEM] != 0.f) && prior;
  for (int p = threadIdx.x; p < P; p += kT) {
    const int r = p / n, c = p - r * n;
    const float v = simg[p];
    float g = w_l1 * sgn(v) + g_flux;
    float dq = 0.f, da = 0.f;
    if (r > 0) { const float d = v - simg[p - n]; dq += d; da += sgn(d); }
    if (r + 1 < n) { const float d = simg[p + n] - v; dq -= d; da -= sgn(d); }
    if (c > 0) { const float d = v - simg[p - 1]; dq += d; da += sgn(d); }
    if (c + 1 < n) { const float d = simg[p + 1] - v; dq -= d; da -= sgn(d); }
    g = fmaf(w_tsv, dq, g);
    g = fmaf(w_tv, da, g);
    if (use_c) g = fmaf(w_c, (xc - center) * ((float)c - xc) + (yc - center) * ((float)r - yc), g);
    if (use_mem) g = fmaf(w_mem, logf(v + 1e-12f) - logf(__ldg(prior + p) + 1e-12f) + v / (v + 1e-12f), g);
    g_img[(size_t)b * P + p] = g;
  }
}

__global__ void reduce_sum_kernel(int n, const float* __restrict__ in, float* __restrict__ out,
                                  double* __restrict__ sq_out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)in[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    const float f = (float)s;
    if (out) out[0] = f;
    if (sq_out) sq_out[0] = (double)f * (double)f;
  }
}

struct TermArgs {
  const double* data64; const float* data32; const float* priors; const float* logdet_flow; const float* det;
  float w64, w32, w_logdet; float wp[DPI_PRIOR_COUNT];
};

// Loss assembly of DPI_interferometry.py:224-229 / DPI_MRI.py:186-193 (means over the batch).
// out: [0] loss, [1] mean data64 term, [2] mean data32 term, [3] mean weighted prior,
//      [4] mean total logdet, [5..10] mean of each prior value
__global__ void __launch_bounds__(256) step_terms_kernel(int B, TermArgs a, float* __restrict__ out) {
  __shared__ double red[32];
  double s64 = 0, s32 = 0, sld = 0, sp[DPI_PRIOR_COUNT] = {0, 0, 0, 0, 0, 0};
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    if (a.data64) s64 += a.data64[b];
    if (a.data32) s32 += (double)a.data32[b];
    sld += (double)a.logdet_flow[b] + (double)a.det[b];
    if (a.priors)
      for (int t = 0; t < DPI_PRIOR_COUNT; ++t) sp[t] += (double)a.priors[(size_t)b * DPI_PRIOR_COUNT + t];
  }
  s64 = block_sum(s64, red); s32 = block_sum(s32, red); sld = block_sum(sld, red);
  for (int t = 0; t < DPI_PRIOR_COUNT; ++t) sp[t] = block_sum(sp[t], red);
  if (threadIdx.x == 0) {
    const double inv = 1.0 / B;
    double prior = 0;
    for (int t = 0; t < DPI_PRIOR_COUNT; ++t) {
      if (a.wp[t] != 0.f) prior += (double)a.wp[t] * sp[t] * inv;
      out[5 + t] = (float)(sp[t] * inv);
    }
    out[1] = (float)(s64 * inv); out[2] = (float)(s32 * inv); out[3] = (float)prior; out[4] = (float)(sld * inv);
    out[0] = (float)((double)a.w64 * s64 * inv + (double)a.w32 * s32 * inv + prior - (double)a.w_logdet * sld * inv);
  }
}

// y[r][i] = s[r] * x[r][i] (y may be NULL) and dot[r] = coef * sum_i x[r][i] * b[r][i] (dot may be NULL); one CTA per row,
// fixed-order reduction. Used by the flux_flag=True form of the crescent renderer (geometric_model.py:373-376).
__global__ void __launch_bounds__(256) row_scale_dot_kernel(int n, const float* __restrict__ x, const float* __restrict__ s,
                                                           float* __restrict__ y, const float* __restrict__ b, float coef,
                                                           float* __restrict__ dot) {
  __shared__ float red[8];
  const int r = blockIdx.x, t = threadIdx.x;
  const float sc = s ? s[r] : 1.f;
  float acc = 0.f;
  for (int i = t; i < n; i += 256) {
    const float v = x[(size_t)r * n + i];
    if (y) y[(size_t)r * n + i] = sc * v;
    if (dot) acc = fmaf(v, b[(size_t)r * n + i], acc);
  }
  if (dot) {
    acc = warp_sum(acc);
    if ((t & 31) == 0) red[t >> 5] = acc;
    __syncthreads();
    if (t == 0) {
      float tot = 0.f;
      for (int w = 0; w < 8; ++w) tot += red[w];
      dot[r] = coef * tot;
    }
  }
}

// interferometry_helpers.py torch_complex_mul :30-34: out[o][:, i] = x[o][:, i] * y[:, i] as complex numbers stored as
// (real row, imaginary row); conj = 1 multiplies with conj(y) (the gradient wrt x).
__global__ void complex_mul_kernel(long long outer, int n, int conj, const float* __restrict__ x, const float* __restrict__ y,
                                   float* __restrict__ out) {
  const long long total = outer * n;
  for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (long long)gridDim.x * blockDim.x) {
    const long long o = id / n;
    const int i = (int)(id - o * n);
    const float xr = x[(o * 2) * n + i], xi = x[(o * 2 + 1) * n + i];
    const float yr = __ldg(y + i), yi = conj ? -__ldg(y + n + i) : __ldg(y + n + i);
    out[(o * 2) * n + i] = xr * yr - xi * yi;
    out[(o * 2 + 1) * n + i] = xr * yi + xi * yr;
  }
}

}  // namespace
}  // namespace dpi

using namespace dpi;

extern "C" {

int dpi_positivity_forward(int batch, int npix, const float* x, const float* log_scale, float* img, float* det,
                           dpi_stream_t stream) {
  DPI_REQUIRE(batch >= 1 && npix >= 1 && x && log_scale && img && det, "dpi_positivity_forward: bad argument");
  positivity_fwd_kernel<<<batch, kT, 0, (cudaStream_t)stream>>>(npix * npix, x, log_scale, img, det);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int dpi_positivity_backward(int batch, int npix, const float* x, const float* log_scale, const float* img,
                            const float* g_img, const float* g_img2, const float* g_det, float* g_x,
                            float* g_log_scale_partial, dpi_stream_t stream) {
  DPI_REQUIRE(batch >= 1 && npix >= 1 && x && log_scale && img && g_x, "dpi_positivity_backward: bad argument");
  positivity_bwd_kernel<<<batch, kT, 0, (cudaStream_t)stream>>>(npix * npix, x, log_scale, img, g_img, g_img2, g_det,
                                                               g_x, g_log_scale_partial);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

static int launch_priors(int batch, int npix, const float* x, const float* log_scale, float* img, float* det,
                         const float* prior_img, float flux, float center, float* values, const float* gw,
                         float* g_img, cudaStream_t stream) {
  const size_t smem = (size_t)npix * npix * sizeof(float);
  DPI_REQUIRE(smem <= 200 * 1024, "image priors: npix=%d too large for one CTA", npix);
  if (smem > 48 * 1024)
    DPI_CUDA(cudaFuncSetAttribute(image_priors_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  image_priors_kernel<<<batch, kT, smem, stream>>>(npix, x, log_scale, img, det, prior_img, flux, center, values, gw,
                                                   g_img);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int dpi_image_priors(int batch, int npix, const float* img, const float* prior_img, float flux, float center,
                     float* values, const float* gw, float* g_img, dpi_stream_t stream) {
  DPI_REQUIRE(batch >= 1 && npix >= 2 && img && values, "dpi_image_priors: bad argument");
  DPI_REQUIRE(!g_img || gw, "dpi_image_priors: g_img requested without weights");
  return launch_priors(batch, npix, nullptr, nullptr, const_cast<float*>(img), nullptr, prior_img, flux, center, values,
                       gw, g_img, (cudaStream_t)stream);
}

int dpi_image_head(int batch, int npix, const float* x, const float* log_scale, float* img, float* det,
                   const float* prior_img, float flux, float center, float* values, const float* gw, float* g_img,
                   dpi_stream_t stream) {
  DPI_REQUIRE(batch >= 1 && npix >= 2 && x && log_scale && img && det && values, "dpi_image_head: bad argument");
  DPI_REQUIRE(!g_img || gw, "dpi_image_head: g_img requested without weights");
  return launch_priors(batch, npix, x, log_scale, img, det, prior_img, flux, center, values, gw, g_img,
                       (cudaStream_t)stream);
}

int dpi_step_terms(int batch, const double* data64, float w64, const float* data32, float w32, const float* priors,
                   const float* prior_weights, const float* logdet_flow, const float* det, float w_logdet, float* out,
                   dpi_stream_t stream) {
  DPI_REQUIRE(batch >= 1 && logdet_flow && det && out, "dpi_step_terms: bad argument");
  TermArgs a{data64, data32, priors, logdet_flow, det, w64, w32, w_logdet, {0, 0, 0, 0, 0, 0}};
  if (prior_weights)
    for (int t = 0; t < DPI_PRIOR_COUNT; ++t) a.wp[t] = prior_weights[t];
  step_terms_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(batch, a, out);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int dpi_reduce_sum(int n, const float* in, float* out, double* sq_out, dpi_stream_t stream) {
  DPI_REQUIRE(n >= 1 && in, "dpi_reduce_sum: bad argument");
  reduce_sum_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(n, in, out, sq_out);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int dpi_complex_mul(long long outer, int n, int conj, const float* x, const float* y, float* out, dpi_stream_t stream) {
  DPI_REQUIRE(outer >= 1 && n >= 1 && x && y && out, "dpi_complex_mul: bad argument");
  const long long total = outer * n;
  complex_mul_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(outer, n, conj, x, y, out);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int dpi_row_scale_dot(int rows, int n, const float* x, const float* s, float* y, const float* b, float coef, float* dot,
                      dpi_stream_t stream) {
  DPI_REQUIRE(rows >= 1 && n >= 1 && x && (y || dot) && (!dot || b), "dpi_row_scale_dot: bad argument");
  row_scale_dot_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(n, x, s, y, b, coef, dot);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

}  // extern "C"
