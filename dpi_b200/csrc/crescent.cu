// crescent.cu - alpha-DPI (DPIx) pieces of the hot path:
//   * SimpleCrescentNuisanceFloor_Param2Img.forward with flux_flag=False (geometric_model.py:275-381) and its
//     analytic backward: one CTA per image, the image never leaves the SM between its five normalising reductions;
//   * sigmoid squashing + det_sigmoid (DPIx_interferometry.py:339,344) and its backward;
//   * the alpha-divergence objective (DPIx_interferometry.py:350-380): per-sample objective, self-normalised
//     importance weights softmax_i(-(1 - alpha) loss_i) over the WHOLE batch (detached), the weights' gradient
//     factors for the data term and for logdet, and the monitoring value loss_orig.
// All reductions are fixed-order (deterministic).
#include "common.cuh"

namespace dpi {
namespace {

constexpr int kCT = 256;
constexpr int kMaxG = 4;                 // nuisance Gaussians supported
constexpr float kEpsN = 1e-4f;           // geometric_model.py:263

// ranges of the constructor defaults used by DPIx_interferometry.py:228-229 (geometric_model.py:238-240)
struct Feat {
  float r, sg, s, eta, fl, cf;
  float x0[kMaxG], y0[kMaxG], a[kMaxG], sx[kMaxG], sy[kMaxG], ct[kMaxG], st[kMaxG];
  // per-image reciprocals of what every pixel would otherwise divide by (a division is ~8 instructions, and the pixel
  // loops run 3 (forward) / 5 (backward) times over the image)
  float i2s2, irs, isx2[kMaxG], isy2[kMaxG], gn[kMaxG];
  float ce, se;   // cos / sin of eta: cos(theta - eta) = (gx ce + gy se) / r_grid without atan2 + sincos per pixel
};
struct FeatScale {   // d feature / d parameter
  float r, sg, s, eta, fl, cf, sh, a, sxy, th;
};

__device__ __forceinline__ FeatScale feat_scale(float fov) {
  const float hf = 0.5f * fov;
  FeatScale k;
  k.r = (40.0f - 10.0f) / hf; k.sg = (40.0f - 1.0f) / hf; k.s = 0.99f - 1e-3f;
  k.eta = (float)(181.0 / 180.0 * 3.14159265358979323846 * 2.0);
  k.fl = 1.0f; k.cf = 2.0f - 1e-3f; k.sh = 400.0f / hf; k.a = 2.0f - 1e-3f; k.sxy = 99.0f / hf;
  k.th = (float)(181.0 / 180.0 * 0.5 * 3.14159265358979323846);
  return k;
}
__device__ __forceinline__ Feat features(const float* p, int g, float fov) {   // compute_features :275-321
  const float hf = 0.5f * fov;
  const FeatScale k = feat_scale(fov);
  Feat f;
  f.r = 10.0f / hf + p[0] * k.r;
  f.sg = 1.0f / hf + p[1] * k.sg;
  f.s = 1e-3f + p[2] * k.s;
  f.eta = (float)(181.0 / 180.0 * 3.14159265358979323846) * (2.0f * p[3] - 1.0f);
  f.fl = 0.0f + 1.0f * p[4 + 6 * g];
  f.cf = 1e-3f + k.cf * p[5 + 6 * g];
  for (int q = 0; q < kMaxG; ++q) {
    if (q < g) {
      f.x0[q] = -200.0f / hf + p[4 + 6 * q] * k.sh;
      f.y0[q] = -200.0f / hf + p[5 + 6 * q] * k.sh;
      f.a[q] = 1e-3f + p[6 + 6 * q] * k.a;
      f.sx[q] = 1.0f / hf + p[7 + 6 * q] * k.sxy;
      f.sy[q] = 1.0f / hf + p[8 + 6 * q] * k.sxy;
      const float th = k.th * p[9 + 6 * q];
      f.ct[q] = cosf(th); f.st[q] = sinf(th);
    } else {
      f.x0[q] = f.y0[q] = f.a[q] = 0.f; f.sx[q] = f.sy[q] = 1.f; f.ct[q] = 1.f; f.st[q] = 0.f;
    }
    f.isx2[q] = 1.0f / (f.sx[q] * f.sx[q]);
    f.isy2[q] = 1.0f / (f.sy[q] * f.sy[q]);
    f.gn[q] = 1.0f / (6.28318530717958648f * f.sx[q] * f.sy[q]);
  }
  f.ce = cosf(f.eta); f.se = sinf(f.eta);
  f.i2s2 = 0.5f / (f.sg * f.sg);
  f.irs = 1.0f / (1.41421356237309515f * f.sg);
  return f;
}

// grids :265-272: xs = arange(-1 + 1/npix, 1, 2/npix); grid_y, grid_x = meshgrid(-xs, xs)
__device__ __forceinline__ void grid_at(int pix, int npix, float& gx, float& gy) {
  const int row = pix / npix, col = pix - row * npix;
  const double gap = 1.0 / npix;
  gx = (float)(-1.0 + gap + col * 2.0 * gap);
  gy = -(float)(-1.0 + gap + row * 2.0 * gap);
}

struct Pix {
  float gx, gy, rg, th, ring, cs, sn, c, dk;
  float n[kMaxG], xr[kMaxG], yr[kMaxG];
};
__device__ __forceinline__ Pix pixel(int pix, int npix, const Feat& f, int g) {   // forward :341-366, unnormalised parts
  Pix q;
  grid_at(pix, npix, q.gx, q.gy);
  q.rg = sqrtf(q.gx * q.gx + q.gy * q.gy);
  q.th = 0.f;   // grid_theta (:272) only enters through cos / sin (theta - eta)
  const float irg = q.rg > 0.f ? 1.0f / q.rg : 0.f;   // odd npix: the centre pixel sits on the origin (atan2(0, 0) = 0)
  const float d = q.rg - f.r;
  q.ring = expf(-d * d * f.i2s2);
  q.cs = q.rg > 0.f ? (q.gx * f.ce + q.gy * f.se) * irg : f.ce;
  q.sn = q.rg > 0.f ? (q.gy * f.ce - q.gx * f.se) * irg : -f.se;
  q.c = (1.0f + f.s * q.cs) * q.ring;
  q.dk = 0.5f * (1.0f + erff((f.r - q.rg) * f.irs));
  for (int k = 0; k < kMaxG; ++k) {
    if (k < g) {
      const float xc = q.gx - f.x0[k], yc = q.gy - f.y0[k];
      q.xr[k] = xc * f.ct[k] + yc * f.st[k];
      q.yr[k] = -xc * f.st[k] + yc * f.ct[k];
      const float delta = 0.5f * (q.xr[k] * q.xr[k] * f.isx2[k] + q.yr[k] * q.yr[k] * f.isy2[k]);
      q.n[k] = f.gn[k] * expf(-delta);
    } else {
      q.n[k] = 0.f; q.xr[k] = q.yr[k] = 0.f;
    }
  }
  return q;
}

// The transcendental part of a pixel (ring, cos / sin of the azimuth offset, erf disk, Gaussian bumps) does not change
// between the normalising passes of one image: with CACHED the first pass stores it in dynamic shared memory
// ([4 + g][P] floats, each thread re-reads only what it wrote) and the later passes rebuild the Pix from it with the
// same arithmetic - bit-identical to recomputing, one evaluation of atan2 / erf / exp instead of three (five backward).
template <bool CACHED>
__device__ __forceinline__ Pix get_pixel(int pix, int npix, int P, const Feat& f, int g, float* cache, bool first) {
  if (!CACHED || first) {
    const Pix q = pixel(pix, npix, f, g);
    if (CACHED) {
      cache[pix] = q.ring; cache[P + pix] = q.cs; cache[2 * P + pix] = q.sn; cache[3 * P + pix] = q.dk;
      for (int k = 0; k < kMaxG; ++k)
        if (k < g) cache[(4 + k) * P + pix] = q.n[k];
    }
    return q;
  }
  Pix q;
  grid_at(pix, npix, q.gx, q.gy);
  q.rg = sqrtf(q.gx * q.gx + q.gy * q.gy);
  q.th = 0.f;   // only enters through cs / sn
  q.ring = cache[pix]; q.cs = cache[P + pix]; q.sn = cache[2 * P + pix]; q.dk = cache[3 * P + pix];
  q.c = (1.0f + f.s * q.cs) * q.ring;
  for (int k = 0; k < kMaxG; ++k) {
    if (k < g) {
      const float xc = q.gx - f.x0[k], yc = q.gy - f.y0[k];
      q.xr[k] = xc * f.ct[k] + yc * f.st[k];
      q.yr[k] = -xc * f.st[k] + yc * f.ct[k];
      q.n[k] = cache[(4 + k) * P + pix];
    } else {
      q.n[k] = 0.f; q.xr[k] = q.yr[k] = 0.f;
    }
  }
  return q;
}

// sums v[0..K) over the CTA in a fixed order; every thread receives the totals
template <int K>
__device__ __forceinline__ void block_sum_multi(float (&v)[K], float* sm /* >= 8 * K floats */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < K; ++k) sm[w * K + k] = v[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    float t = 0.f;
#pragma unroll
    for (int ww = 0; ww < kCT / 32; ++ww) t += sm[ww * K + k];
    v[k] = t;
  }
}

struct Norms { float C, Dk, T, N[kMaxG], iC, iDk, iT, iN[kMaxG]; };

// total image before the final normalisation at one pixel
__device__ __forceinline__ float tot_at(const Pix& q, const Feat& f, const Norms& nm, int g) {
  float t = f.cf * ((1.0f - f.fl) * (q.c * nm.iC) + f.fl * (q.dk * nm.iDk));
  for (int k = 0; k < kMaxG; ++k)
    if (k < g) t += f.a[k] * (q.n[k] * nm.iN[k]);
  return t;
}

template <bool CACHED>
__device__ __forceinline__ Norms compute_norms(int npix, const Feat& f, int g, float* sm, float* cache) {
  const int P = npix * npix;
  float v[2 + kMaxG];
#pragma unroll
  for (int k = 0; k < 2 + kMaxG; ++k) v[k] = 0.f;
  for (int pix = threadIdx.x; pix < P; pix += kCT) {
    const Pix q = get_pixel<CACHED>(pix, npix, P, f, g, cache, true);
    v[0] += q.c; v[1] += q.dk;
#pragma unroll
    for (int k = 0; k < kMaxG; ++k) v[2 + k] += q.n[k];
  }
  block_sum_multi<2 + kMaxG>(v, sm);
  Norms nm;
  nm.C = v[0] + kEpsN; nm.Dk = v[1] + kEpsN;
  nm.iC = 1.0f / nm.C; nm.iDk = 1.0f / nm.Dk;
#pragma unroll
  for (int k = 0; k < kMaxG; ++k) { nm.N[k] = v[2 + k] + kEpsN; nm.iN[k] = 1.0f / nm.N[k]; }
  float t[1] = {0.f};
  for (int pix = threadIdx.x; pix < P; pix += kCT) t[0] += tot_at(get_pixel<CACHED>(pix, npix, P, f, g, cache, false), f, nm, g);
  block_sum_multi<1>(t, sm);
  nm.T = t[0] + kEpsN;
  nm.iT = 1.0f / nm.T;
  return nm;
}

template <bool CACHED>
__global__ void __launch_bounds__(kCT) crescent_fwd_kernel(int npix, int g, float fov, const float* __restrict__ params,
                                                           float* __restrict__ img) {
  extern __shared__ float cache[];
  __shared__ float sm[8 * (2 + kMaxG)];
  __shared__ float sp[6 + 6 * kMaxG];
  const int b = blockIdx.x, np = 6 + 6 * g, P = npix * npix;
  if (threadIdx.x < np) sp[threadIdx.x] = params[(size_t)b * np + threadIdx.x];
  __syncthreads();
  const Feat f = features(sp, g, fov);
  const Norms nm = compute_norms<CACHED>(npix, f, g, sm, cache);
  float* out = img + (size_t)b * P;
  for (int pix = threadIdx.x; pix < P; pix += kCT)
    out[pix] = tot_at(get_pixel<CACHED>(pix, npix, P, f, g, cache, false), f, nm, g) * nm.iT;
}

// g_params[b, :] = gscale[b] * sum_p g_img[b, p] * d img[b, p] / d params[b, :]
template <bool CACHED>
__global__ void __launch_bounds__(kCT) crescent_bwd_kernel(int npix, int g, float fov, const float* __restrict__ params,
                                                           const float* __restrict__ g_img, const float* __restrict__ gscale,
                                                           float* __restrict__ g_params) {
  extern __shared__ float cache[];
  constexpr int KS = 4 + 5 * kMaxG;
  __shared__ float sm[8 * KS];
  __shared__ float sp[6 + 6 * kMaxG];
  const int b = blockIdx.x, np = 6 + 6 * g, P = npix * npix;
  if (threadIdx.x < np) sp[threadIdx.x] = params[(size_t)b * np + threadIdx.x];
  __syncthreads();
  const Feat f = features(sp, g, fov);
  const Norms nm = compute_norms<CACHED>(npix, f, g, sm, cache);
  const float* G = g_img + (size_t)b * P;
  const float gs = gscale ? gscale[b] : 1.0f;
  // img = tot / T: u = d L / d tot = (G - sum(G * img)) / T
  float gi[1] = {0.f};
  for (int pix = threadIdx.x; pix < P; pix += kCT) gi[0] += G[pix] * (tot_at(get_pixel<CACHED>(pix, npix, P, f, g, cache, false), f, nm, g) * nm.iT);
  block_sum_multi<1>(gi, sm);
  const float GI = gi[0];
  // second-level sums: sum u c, sum u dk, sum u n_k
  float u2[2 + kMaxG];
#pragma unroll
  for (int k = 0; k < 2 + kMaxG; ++k) u2[k] = 0.f;
  for (int pix = threadIdx.x; pix < P; pix += kCT) {
    const Pix q = get_pixel<CACHED>(pix, npix, P, f, g, cache, false);
    const float u = gs * (G[pix] - GI) * nm.iT;
    u2[0] += u * q.c; u2[1] += u * q.dk;
#pragma unroll
    for (int k = 0; k < kMaxG; ++k) u2[2 + k] += u * q.n[k];
  }
  block_sum_multi<2 + kMaxG>(u2, sm);
  const float Uc = u2[0], Ud = u2[1];
  // per-feature pixel sums
  float s[KS];
#pragma unroll
  for (int k = 0; k < KS; ++k) s[k] = 0.f;
  const float kc = f.cf * (1.0f - f.fl) * nm.iC, kd = f.cf * f.fl * nm.iDk;
  const float inv_sg = 1.0f / f.sg;
  const float UcC = Uc * nm.iC, UdD = Ud * nm.iDk;
  float vnk[kMaxG], unk[kMaxG], isx[kMaxG], isy[kMaxG];
#pragma unroll
  for (int k = 0; k < kMaxG; ++k) {
    vnk[k] = f.a[k] * nm.iN[k]; unk[k] = u2[2 + k] * nm.iN[k];
    isx[k] = 1.0f / f.sx[k]; isy[k] = 1.0f / f.sy[k];
  }
  for (int pix = threadIdx.x; pix < P; pix += kCT) {
    const Pix q = get_pixel<CACHED>(pix, npix, P, f, g, cache, false);
    const float u = gs * (G[pix] - GI) * nm.iT;
    const float vc = kc * (u - UcC);               // d L / d c (unnormalised crescent)
    const float vd = kd * (u - UdD);               // d L / d disk
    const float d = q.rg - f.r;
    const float Sx = 1.0f + f.s * q.cs;
    const float z = (f.r - q.rg) * inv_sg * 0.707106781186547524f;
    s[0] += vc * Sx * q.ring * d * inv_sg * inv_sg + vd * q.ring * inv_sg * 0.398942280401432678f;                      // r
    s[1] += vc * Sx * q.ring * d * d * inv_sg * inv_sg * inv_sg - vd * q.ring * z * inv_sg * 0.564189583547756287f;    // sigma
    s[2] += vc * q.ring * q.cs;                                                                                         // s
    s[3] += vc * q.ring * f.s * q.sn;                                                                                   // eta
#pragma unroll
    for (int k = 0; k < kMaxG; ++k) {
      if (k < g) {
        const float vn = vnk[k] * (u - unk[k]);                           // d L / d n_k (unnormalised Gaussian)
        const float n = q.n[k], xr = q.xr[k], yr = q.yr[k];
        const float ix2 = f.isx2[k], iy2 = f.isy2[k];
        const float ddx = xr * ix2, ddy = yr * iy2;                       // d delta / d xr, d delta / d yr
        s[4 + 5 * k + 0] += vn * n * (ddx * f.ct[k] - ddy * f.st[k]);     // x0:  -n * (ddx * (-ct) + ddy * st)
        s[4 + 5 * k + 1] += vn * n * (ddx * f.st[k] + ddy * f.ct[k]);     // y0:  -n * (ddx * (-st) + ddy * (-ct))
        s[4 + 5 * k + 2] += vn * n * (xr * xr * ix2 - 1.0f) * isx[k];     // sigma_x
        s[4 + 5 * k + 3] += vn * n * (yr * yr * iy2 - 1.0f) * isy[k];     // sigma_y
        s[4 + 5 * k + 4] += -vn * n * (ddx * yr - ddy * xr);              // theta: d xr = yr, d yr = -xr
      }
    }
  }
  block_sum_multi<KS>(s, sm);
  if (threadIdx.x == 0) {
    const FeatScale k = feat_scale(fov);
    float* o = g_params + (size_t)b * np;
    o[0] = s[0] * k.r; o[1] = s[1] * k.sg; o[2] = s[2] * k.s; o[3] = s[3] * k.eta;
    for (int q = 0; q < g; ++q) {
      o[4 + 6 * q] = s[4 + 5 * q] * k.sh;
      o[5 + 6 * q] = s[4 + 5 * q + 1] * k.sh;
      o[6 + 6 * q] = (u2[2 + q] * nm.iN[q]) * k.a;
      o[7 + 6 * q] = s[4 + 5 * q + 2] * k.sxy;
      o[8 + 6 * q] = s[4 + 5 * q + 3] * k.sxy;
      o[9 + 6 * q] = s[4 + 5 * q + 4] * k.th;
    }
    o[4 + 6 * g] = f.cf * (UdD - UcC) * k.fl;                                      // floor
    o[5 + 6 * g] = ((1.0f - f.fl) * UcC + f.fl * UdD) * k.cf;                      // crescent flux
  }
}

// params = sigmoid(ps); det[b] = sum(-ps - 2 softplus(-ps))  (DPIx_interferometry.py:339,344)
__global__ void sigmoid_fwd_kernel(int batch, int n, const float* __restrict__ ps, float* __restrict__ params,
                                   float* __restrict__ det) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= batch) return;
  float acc = 0.f;
  for (int j = lane; j < n; j += 32) {
    const float x = ps[(size_t)b * n + j];
    params[(size_t)b * n + j] = sigmoid_f(x);
    acc += -x - 2.0f * softplus_f(-x);
  }
  acc = warp_sum(acc);
  if (lane == 0) det[b] = acc;
}
// g_ps = g_params * s (1 - s) + g_det[b] * (1 - 2 s)
__global__ void sigmoid_bwd_kernel(int batch, int n, const float* __restrict__ ps, const float* __restrict__ g_params,
                                   const float* __restrict__ g_det, float* __restrict__ g_ps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * n) return;
  const float sg = sigmoid_f(ps[i]);
  const float gd = g_det ? g_det[i / n] : 0.f;
  g_ps[i] = (g_params ? g_params[i] : 0.f) * sg * (1.0f - sg) + gd * (1.0f - 2.0f * sg);
}

// The alpha-divergence objective over the whole batch, one CTA (DPIx_interferometry.py:350-380).
//   loss_data_i = 0.5 (w_ca l_ca_i + w_cp l_cp_i) / sf;  logprob_i = -(logdet_flow_i + det_i) - 0.5 |z_i|^2
//   loss_i = dw * loss_data_i + logprob_i;  weights w_i = 1/B (alpha == 1) or softmax_i(-(1 - alpha) loss_i), detached
//   loss = sum_i w_i sf loss_i;  gscale_i = 0.5 dw w_i (times d(w_ca l_ca + w_cp l_cp)/d img);  g_logdet_i = -w_i sf
// mode 0: everything from this batch alone. Data parallel (softmax over the global batch): mode 1 writes
// stats[0] = max_i t_i (t_i = -(1 - alpha) loss_i); mode 2 reads the global max from stats[0] and writes
// stats[1] = sum_i exp(t_i - max); mode 3 reads the global (max, sum) and finishes.
__global__ void __launch_bounds__(1024) alpha_objective_kernel(int B, int np, const float* __restrict__ z,
                                                               const double* __restrict__ l_cp, const float* __restrict__ l_ca,
                                                               const float* __restrict__ logdet_flow, const float* __restrict__ det,
                                                               float w_cp, float w_ca, float sf, float dw, float alpha, int mode,
                                                               int global_batch, float* __restrict__ loss_i_out,
                                                               float* __restrict__ gscale, float* __restrict__ g_logdet,
                                                               float* __restrict__ terms, float* __restrict__ stats) {
  __shared__ double sred[32];
  __shared__ float sbc[2];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
  auto bsum = [&](double v) -> double {
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sred[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int k = 0; k < nw; ++k) t += sred[k];
    return t;
  };
  auto bmax = [&](float v) -> float {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) sred[w] = (double)v;
    __syncthreads();
    float t = -INFINITY;
    for (int k = 0; k < nw; ++k) t = fmaxf(t, (float)sred[k]);
    return t;
  };
  auto sample = [&](int i, float& li, float& lo, float& ld) {
    float zz = 0.f;
    for (int j = 0; j < np; ++j) { const float v = z[(size_t)i * np + j]; zz = fmaf(v, v, zz); }
    ld = logdet_flow[i] + det[i];
    const float data = 0.5f * (w_ca * l_ca[i] + w_cp * (float)l_cp[i]) / sf;
    const float logprob = -ld - 0.5f * zz;
    li = dw * data + logprob;
    lo = data + logprob;
  };
  const float k1 = -(1.0f - alpha);
  const bool kl = alpha == 1.0f;
  float mx = -INFINITY, mxo = -INFINITY;
  for (int i = tid; i < B; i += blockDim.x) {
    float li, lo, ld;
    sample(i, li, lo, ld);
    if (loss_i_out) loss_i_out[i] = li;
    mx = fmaxf(mx, k1 * li); mxo = fmaxf(mxo, k1 * lo);
  }
  float M = kl ? 0.f : bmax(mx);
  const float Mo = kl ? 0.f : bmax(mxo);
  if (mode == 1) { if (tid == 0) stats[0] = M; return; }
  if (mode >= 2) M = stats[0];
  double zs = 0.0, zo = 0.0, s_cp = 0.0, s_ca = 0.0, s_ld = 0.0, s_li = 0.0, s_lo = 0.0;
  for (int i = tid; i < B; i += blockDim.x) {
    float li, lo, ld;
    sample(i, li, lo, ld);
    if (!kl) { zs += (double)expf(k1 * li - M); zo += (double)expf(k1 * lo - Mo); }
    s_cp += l_cp[i]; s_ca += (double)l_ca[i]; s_ld += (double)ld; s_li += (double)li; s_lo += (double)lo;
  }
  double Z = kl ? 1.0 : bsum(zs);
  if (mode == 2) { if (tid == 0) stats[1] = (float)Z; return; }
  if (mode == 3) Z = (double)stats[1];
  const double Zo = kl ? 1.0 : bsum(zo);
  s_cp = bsum(s_cp); s_ca = bsum(s_ca); s_ld = bsum(s_ld); s_li = bsum(s_li); s_lo = bsum(s_lo);
  double wl = 0.0;
  for (int i = tid; i < B; i += blockDim.x) {
    float li, lo, ld;
    sample(i, li, lo, ld);
    const float wi = kl ? 1.0f / (float)global_batch : (float)((double)expf(k1 * li - M) / Z);
    gscale[i] = 0.5f * dw * wi;
    g_logdet[i] = -wi * sf;
    wl += (double)wi * (double)sf * (double)li;
  }
  wl = bsum(wl);
  if (tid == 0) {
    terms[0] = (float)wl;                                                             // loss
    terms[1] = kl ? (float)(sf * s_lo / B)                                            // loss_orig (this batch)
                  : (float)(sf * ((double)Mo + log(Zo / B)) / ((double)alpha - 1.0));
    terms[2] = (float)(s_cp / B); terms[3] = (float)(s_ca / B); terms[4] = (float)(s_ld / B);
    terms[5] = (float)(s_li / B);
  }
  (void)sbc;
}

// dynamic shared memory of the cached variants: [4 + g][npix^2] floats; 0 = does not fit two CTAs per SM worth keeping
size_t cache_bytes(int npix, int g) {
  const size_t b = (size_t)(4 + g) * npix * npix * sizeof(float);
  return b <= 110 * 1024 ? b : 0;
}
bool cache_ready() {   // opt in to > 48 KB of dynamic shared memory, once per device
  static std::atomic<unsigned long long> done{0}, bad{0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  const unsigned long long bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_acquire) & bit) return !(bad.load(std::memory_order_acquire) & bit);
  const bool ok =
      cudaFuncSetAttribute((const void*)crescent_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024) == cudaSuccess &&
      cudaFuncSetAttribute((const void*)crescent_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024) == cudaSuccess;
  if (!ok) { cudaGetLastError(); bad.fetch_or(bit, std::memory_order_release); }
  done.fetch_or(bit, std::memory_order_release);
  return ok;
}

}  // namespace
}  // namespace dpi

using namespace dpi;

extern "C" {

int dpi_crescent_forward(int batch, int npix, int n_gaussian, float fov, const float* params, float* img,
                         dpi_stream_t stream) {
  DPI_REQUIRE(batch >= 1 && npix >= 2 && params && img, "dpi_crescent_forward: bad argument");
  DPI_REQUIRE(n_gaussian >= 0 && n_gaussian <= kMaxG, "dpi_crescent_forward: n_gaussian must be in [0, %d]", kMaxG);
  const size_t cache = cache_bytes(npix, n_gaussian);
  if (cache && cache_ready())
    crescent_fwd_kernel<true><<<batch, kCT, cache, (cudaStream_t)stream>>>(npix, n_gaussian, fov, params, img);
  else
    crescent_fwd_kernel<false><<<batch, kCT, 0, (cudaStream_t)stream>>>(npix, n_gaussian, fov, params, img);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int dpi_crescent_backward(int batch, int npix, int n_gaussian, float fov, const float* params, const float* g_img,
                          const float* gscale, float* g_params, dpi_stream_t stream) {
  DPI_REQUIRE(batch >= 1 && npix >= 2 && params && g_img && g_params, "dpi_crescent_backward: bad argument");
  DPI_REQUIRE(n_gaussian >= 0 && n_gaussian <= kMaxG, "dpi_crescent_backward: n_gaussian must be in [0, %d]", kMaxG);
  const size_t cache = cache_bytes(npix, n_gaussian);
  if (cache && cache_ready())
    crescent_bwd_kernel<true><<<batch, kCT, cache, (cudaStream_t)stream>>>(npix, n_gaussian, fov, params, g_img, gscale, g_params);
  else
    crescent_bwd_kernel<false><<<batch, kCT, 0, (cudaStream_t)stream>>>(npix, n_gaussian, fov, params, g_img, gscale, g_params);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int dpi_sigmoid_forward(int batch, int n, const float* ps, float* params, float* det, dpi_stream_t stream) {
  DPI_REQUIRE(batch >= 1 && n >= 1 && ps && params && det, "dpi_sigmoid_forward: bad argument");
  sigmoid_fwd_kernel<<<(batch + 7) / 8, 256, 0, (cudaStream_t)stream>>>(batch, n, ps, params, det);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int dpi_sigmoid_backward(int batch, int n, const float* ps, const float* g_params, const float* g_det, float* g_ps,
                         dpi_stream_t stream) {
  DPI_REQUIRE(batch >= 1 && n >= 1 && ps && g_ps, "dpi_sigmoid_backward: bad argument");
  const int total = batch * n;
  sigmoid_bwd_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(batch, n, ps, g_params, g_det, g_ps);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int dpi_alpha_objective(int batch, int nparams, const float* z, const double* loss_cp, const float* loss_ca,
                        const float* logdet_flow, const float* det, float w_cp, float w_ca, float scale_factor,
                        float data_weight, float alpha, int mode, int global_batch, float* loss_i, float* gscale,
                        float* g_logdet, float* terms, float* stats, dpi_stream_t stream) {
  DPI_REQUIRE(batch >= 1 && nparams >= 1 && z && loss_cp && loss_ca && logdet_flow && det, "dpi_alpha_objective: null input");
  DPI_REQUIRE(mode >= 0 && mode <= 3, "dpi_alpha_objective: mode must be 0..3");
  DPI_REQUIRE(mode == 0 || stats, "dpi_alpha_objective: stats required for modes 1-3");
  DPI_REQUIRE((mode == 1 || mode == 2) || (gscale && g_logdet && terms), "dpi_alpha_objective: null output");
  alpha_objective_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(batch, nparams, z, loss_cp, loss_ca, logdet_flow, det, w_cp,
                                                               w_ca, scale_factor, data_weight, alpha, mode,
                                                               global_batch > 0 ? global_batch : batch, loss_i, gscale,
                                                               g_logdet, terms, stats);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

}  // extern "C"
