// mri.cu - MRI measurement operator: orthonormal 2-D DFT of a real image (DPI_MRI.py
// fft2c_torch :70-74, identical to numpy fft2(norm="ortho") at :63-67), masked k-space chi^2
// (MRI_helpers.py Loss_kspace_diff2 :23-26 as assembled at DPI_MRI.py:181-182) and its adjoint.
// One CTA per image; the image and its spectrum stay in shared memory. The transform is two
// passes of 1-D DFTs with a shared twiddle table (exact for any npix, not only powers of two).
#include "common.cuh"

namespace dpi {
namespace {

constexpr int kT = 256;

// smem: tw[n] (float2), buf0[n*n] (float2), buf1[n*n] (float2)
// Y[r][kx] = sum_x X[r][x] w^(kx x);  Z[ky][kx] = sum_r Y[r][kx] w^(ky r);  w = exp(-+2 pi i / n)
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// dst[q][c] = sum_j src[j][c] * tw[(q*j) % n]   (transform along the row index), conj selects sign
__device__ void dft_cols(int n, const float2* __restrict__ src, float2* __restrict__ dst,
                         const float2* __restrict__ tw, bool conj) {
  for (int o = threadIdx.x; o < n * n; o += kT) {
    const int q = o / n, c = o - q * n;
    float sr = 0.f, si = 0.f;
    int ph = 0;
    for (int j = 0; j < n; ++j) {
      float2 w = tw[ph];
      if (conj) w.y = -w.y;
      const float2 v = src[j * n + c];
      sr = fmaf(v.x, w.x, sr); sr = fmaf(-v.y, w.y, sr);
      si = fmaf(v.x, w.y, si); si = fmaf(v.y, w.x, si);
      ph += q; if (ph >= n) ph -= n;
    }
    dst[o] = make_float2(sr, si);
  }
}
// dst[r][q] = sum_j src[r][j] * tw[(q*j) % n]   (transform along the column index)
__device__ void dft_rows(int n, const float2* __restrict__ src, float2* __restrict__ dst,
                         const float2* __restrict__ tw, bool conj) {
  for (int o = threadIdx.x; o < n * n; o += kT) {
    const int r = o / n, q = o - r * n;
    float sr = 0.f, si = 0.f;
    int ph = 0;
    for (int j = 0; j < n; ++j) {
      float2 w = tw[ph];
      if (conj) w.y = -w.y;
      const float2 v = src[r * n + j];
      sr = fmaf(v.x, w.x, sr); sr = fmaf(-v.y, w.y, sr);
      si = fmaf(v.x, w.y, si); si = fmaf(v.y, w.x, si);
      ph += q; if (ph >= n) ph -= n;
    }
    dst[o] = make_float2(sr, si);
  }
}

__device__ void make_twiddles(int n, float2* tw) {
  for (int j = threadIdx.x; j < n; j += kT) {
    float s, c;
    sincospif(-2.0f * (float)j / (float)n, &s, &c);  // exp(-2 pi i j / n)
    tw[j] = make_float2(c, s);
  }
}

// mode 0: kspace out only (fft2c forward). mode 1: adjoint only (g_kspace in -> g_img out).
// mode 2: fused data loss (+ gradient if g_img != nullptr).
__global__ void __launch_bounds__(kT) mri_kernel(int mode, int n, const float* __restrict__ img,
                                                 float* __restrict__ kspace, const float* __restrict__ g_kspace,
                                                 const float* __restrict__ mask, const float* __restrict__ ktrue,
                                                 float inv_sigma2_mm, float* __restrict__ loss,
                                                 const float* __restrict__ gw, float* __restrict__ g_img) {
  extern __shared__ float2 smem2[];
  __shared__ float red[32];
  float2* tw = smem2;
  float2* b0 = tw + n;
  float2* b1 = b0 + n * n;
  const int b = blockIdx.x, P = n * n;
  const float scale = 1.0f / (float)n;  // ortho: 1/sqrt(n*n)
  make_twiddles(n, tw);
  if (mode != 1) {
    for (int p = threadIdx.x; p < P; p += kT) b0[p] = make_float2(img[(size_t)b * P + p], 0.f);
    __syncthreads();
    dft_rows(n, b0, b1, tw, false);
    __syncthreads();
    dft_cols(n, b1, b0, tw, false);
    __syncthreads();
    if (mode == 0) {
      for (int p = threadIdx.x; p < P; p += kT) {
        kspace[((size_t)b * P + p) * 2] = b0[p].x * scale;
        kspace[((size_t)b * P + p) * 2 + 1] = b0[p].y * scale;
      }
      return;
    }
    // residual e = Z * mask - true; loss = sum e^2 / (2 P) / sigma^2 / mean_mask
    float s = 0.f;
    const float coef = inv_sigma2_mm / (float)(2 * P);
    const float gwb = (g_img && gw) ? gw[b] : 0.f;
    for (int p = threadIdx.x; p < P; p += kT) {
      const float mr = mask[2 * p], mi = mask[2 * p + 1];
      const float er = b0[p].x * scale * mr - ktrue[2 * p];
      const float ei = b0[p].y * scale * mi - ktrue[2 * p + 1];
      s = fmaf(er, er, s);
      s = fmaf(ei, ei, s);
      b0[p] = make_float2(gwb * 2.f * er * mr * coef, gwb * 2.f * ei * mi * coef);  // dL/dZ
    }
    const float tot = block_sum(s, red);
    if (threadIdx.x == 0) loss[b] = tot * coef;
    if (!g_img) return;
    __syncthreads();
  } else {
    for (int p = threadIdx.x; p < P; p += kT)
      b0[p] = make_float2(g_kspace[((size_t)b * P + p) * 2], g_kspace[((size_t)b * P + p) * 2 + 1]);
    __syncthreads();
  }
  // adjoint: g_img[r][x] = Re( sum_ky sum_kx conj(w^(ky r + kx x)) G[ky][kx] ) / n
  dft_cols(n, b0, b1, tw, true);
  __syncthreads();
  for (int o = threadIdx.x; o < P; o += kT) {
    const int r = o / n, x = o - r * n;
    float sr = 0.f;
    int ph = 0;
    for (int j = 0; j < n; ++j) {
      const float2 w = tw[ph];  // conj(w) = (w.x, -w.y); Re(v * conj(w)) = v.x w.x + v.y w.y
      const float2 v = b1[r * n + j];
      sr = fmaf(v.x, w.x, sr);
      sr = fmaf(v.y, w.y, sr);
      ph += x; if (ph >= n) ph -= n;
    }
    g_img[(size_t)b * P + o] = sr * scale;
  }
}

int launch(int mode, int batch, int npix, const float* img, float* kspace, const float* g_kspace, const float* mask,
           const float* ktrue, float inv_sigma2_mm, float* loss, const float* gw, float* g_img, cudaStream_t st) {
  DPI_REQUIRE(batch >= 1 && npix >= 2 && npix <= 128, "mri: npix=%d unsupported (2..128)", npix);
  const size_t smem = ((size_t)npix + 2 * (size_t)npix * npix) * sizeof(float2);
  if (smem > 48 * 1024)
    DPI_CUDA(cudaFuncSetAttribute(mri_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mri_kernel<<<batch, kT, smem, st>>>(mode, npix, img, kspace, g_kspace, mask, ktrue, inv_sigma2_mm, loss, gw, g_img);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

// MRI_helpers.py:18-26: loss[b] = mean_i |y_pred[b][i] - y_true[i]|^p / sigma^p, p = 1 (Loss_kspace_diff) or 2
// (Loss_kspace_diff2); with g_loss also g_pred[b][i] = g_loss[b] p |d|^(p-1) sign(d) / (n sigma^p). One CTA per sample,
// fixed-order block reduction.
__global__ void __launch_bounds__(256) kspace_loss_kernel(int power, int n, float inv_sigma_p, const float* __restrict__ y_true,
                                                          const float* __restrict__ y_pred, float* __restrict__ loss,
                                                          const float* __restrict__ g_loss, float* __restrict__ g_pred) {
  __shared__ float red[8];
  const int b = blockIdx.x, t = threadIdx.x;
  const float* yp = y_pred + (size_t)b * n;
  const float inv_n = 1.0f / (float)n;
  const float gs = g_loss ? g_loss[b] * inv_sigma_p * inv_n : 0.f;
  float acc = 0.f;
  for (int i = t; i < n; i += 256) {
    const float d = yp[i] - __ldg(y_true + i);
    acc += power == 2 ? d * d : fabsf(d);
    if (g_pred) g_pred[(size_t)b * n + i] = power == 2 ? 2.f * d * gs : (d > 0.f ? gs : (d < 0.f ? -gs : 0.f));
  }
  acc = warp_sum(acc);
  if ((t & 31) == 0) red[t >> 5] = acc;
  __syncthreads();
  if (t == 0 && loss) {
    float tot = 0.f;
    for (int w = 0; w < 8; ++w) tot += red[w];
    loss[b] = tot * inv_n * inv_sigma_p;
  }
}

}  // namespace
}  // namespace dpi

using namespace dpi;

extern "C" {

int dpi_fft2c_forward(int batch, int npix, const float* img, float* kspace, dpi_stream_t stream) {
  DPI_REQUIRE(img && kspace, "dpi_fft2c_forward: null argument");
  return launch(0, batch, npix, img, kspace, nullptr, nullptr, nullptr, 0.f, nullptr, nullptr, nullptr,
                (cudaStream_t)stream);
}

int dpi_fft2c_backward(int batch, int npix, const float* g_kspace, float* g_img, dpi_stream_t stream) {
  DPI_REQUIRE(g_kspace && g_img, "dpi_fft2c_backward: null argument");
  return launch(1, batch, npix, nullptr, nullptr, g_kspace, nullptr, nullptr, 0.f, nullptr, nullptr, g_img,
                (cudaStream_t)stream);
}

int dpi_mri_data_loss(int batch, int npix, const float* img, const float* mask, const float* kspace_true, float sigma,
                      float mean_mask, float* loss, const float* gw, float* g_img, dpi_stream_t stream) {
  DPI_REQUIRE(img && mask && kspace_true && loss && sigma > 0.f && mean_mask > 0.f, "dpi_mri_data_loss: bad argument");
  DPI_REQUIRE(!g_img || gw, "dpi_mri_data_loss: g_img requested without weights");
  return launch(2, batch, npix, img, nullptr, nullptr, mask, kspace_true, 1.0f / (sigma * sigma) / mean_mask, loss, gw,
                g_img, (cudaStream_t)stream);
}

int dpi_kspace_loss(int power, int batch, int n, float sigma, const float* y_true, const float* y_pred, float* loss,
                    const float* g_loss, float* g_pred, dpi_stream_t stream) {
  DPI_REQUIRE((power == 1 || power == 2) && batch >= 1 && n >= 1 && sigma > 0.f && y_true && y_pred, "dpi_kspace_loss: bad argument");
  DPI_REQUIRE(loss || (g_loss && g_pred), "dpi_kspace_loss: nothing to compute");
  DPI_REQUIRE(!g_pred || g_loss, "dpi_kspace_loss: g_pred requested without g_loss");
  const float isp = power == 2 ? 1.0f / (sigma * sigma) : 1.0f / sigma;
  kspace_loss_kernel<<<batch, 256, 0, (cudaStream_t)stream>>>(power, n, isp, y_true, y_pred, loss, g_loss, g_pred);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

}  // extern "C"
