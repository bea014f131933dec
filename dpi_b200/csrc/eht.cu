// eht.cu - VLBI measurement operator and chi^2 terms.
//   interferometry_helpers.py: torch_complex_matmul :36-39, eht_observation_pytorch.func :257-291,
//   Loss_angle_diff :299-307, Loss_logca_diff2 :318-323, Loss_vis_diff :326-331,
//   Loss_logamp_diff :334-339, Loss_visamp_diff :342-347.
// Layout: the DFT matrix stays in the reference's [npix^2, n_vis, 2] layout, i.e. a row-major
// [P, 2*n_vis] matrix with re/im interleaved, so image -> visibilities is ONE GEMM with
// N = 2*n_vis. The closure gathers run per sample in one CTA; their backward uses host-built
// transposed (CSR) tables so it is a deterministic gather, not an atomic scatter.
#include <algorithm>
#include <memory>
#include <vector>

#include "dft_tc.cuh"
#include "gemm_tile.cuh"

struct dpi_eht_s {
  int device = 0, npix = 0, P = 0, n_vis = 0, n_cp = 0, n_ca = 0, n_sm = 0;
  float* d_F = nullptr;            // [P, 2*n_vis]
  // tcgen05 DFT (dft_tc.cuh): F split into TF32 hi / lo and tiled as K-major UMMA operands, for the forward
  // GEMM (B[n][k] = F[k][n]) and for the adjoint (B[n][k] = F[n][k])
  int use_tc = 1;
  float* d_Bf = nullptr; float* d_Bb = nullptr;
  int* d_cp_ind = nullptr;         // [3, n_cp]
  double* d_cp_sign = nullptr;     // [3, n_cp]
  int* d_ca_ind = nullptr;         // [4, n_ca]
  // transposed tables: for visibility k, entries [ptr[k], ptr[k+1])
  int* d_cp_ptr = nullptr; int* d_cp_adj = nullptr; double* d_cp_adj_sign = nullptr;
  int* d_ca_ptr = nullptr; int* d_ca_adj = nullptr; float* d_ca_adj_sign = nullptr;
};

namespace dpi {
namespace {

constexpr int kCtrl = 8192;  // uint32 words at the start of the workspace (split-K counters)

struct DftPlan { int tiles_m, tiles_n, splits, kchunk; };

DftPlan plan_dft(int M, int N, int K, int n_sm) {
  DftPlan p{};
  p.tiles_m = (M + kBM - 1) / kBM;
  p.tiles_n = (N + 63) / 64;
  const int mn = p.tiles_m * p.tiles_n;
  int want = std::max(1, n_sm / std::max(1, mn));
  want = std::min(std::min(want, std::max(1, K / (2 * kBK))), 16);
  int chunk = (K + want - 1) / want;
  chunk = (chunk + kBK - 1) / kBK * kBK;
  p.kchunk = chunk;
  p.splits = (K + chunk - 1) / chunk;
  return p;
}

// C[M, N] = A[M, K] * Bm, where Bm is either [K, N] row-major (NN) or [N, K] row-major (NT).
// mode 0: store C planar as vis[m, c, k] with n = 2k + c (forward DFT); mode 1: plain row-major.
template <bool NT>
__global__ void __launch_bounds__(kThreads) dft_gemm_kernel(int M, int N, int K, const float* __restrict__ A,
                                                            const float* __restrict__ Bm, float* __restrict__ C,
                                                            int mode, DftPlan pl, float* part, unsigned* ctr) {
  __shared__ float sA[kBM * (kBK + 1)];
  __shared__ float sB[64 * (kBK + 1)];
  __shared__ int sFlag;
  const int t = blockIdx.x;
  const int ks = t % pl.splits, mn = t / pl.splits, tn = mn % pl.tiles_n, tm = mn / pl.tiles_n;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  LdGradRow la{A, K, tm * kBM, M, K};
  const int kbeg = ks * pl.kchunk, kend = min(K, kbeg + pl.kchunk);
  if (NT) {
    LdWeightK lb{Bm, K, tn * 64, N, 0, 0};
    gemm_tile<64>(acc, la, lb, kbeg, kend, sA, sB);
  } else {
    LdWeightN lb{Bm, N, tn * 64, N, K};
    gemm_tile<64>(acc, la, lb, kbeg, kend, sA, sB);
  }
  if (!splitk_reduce<64>(acc, part + (size_t)mn * pl.splits * (kBM * 64), ks, pl.splits, ctr + mn, &sFlag)) return;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = tm * kBM + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = tn * 64 + j * 16 + tx;
      if (n >= N) continue;
      if (mode == 0) C[(size_t)m * N + (size_t)(n & 1) * (N / 2) + (n >> 1)] = acc[i][j];
      else C[(size_t)m * N + n] = acc[i][j];
    }
  }
}

constexpr double kPi = 3.141592653589793;
constexpr float kEps = 1e-16f;  // interferometry_helpers.py:236

// One CTA per sample: amplitudes, closure phases (fp64, degrees), log closure amplitudes (:268-289).
__global__ void __launch_bounds__(256) closure_fwd_kernel(int n_vis, int n_cp, int n_ca, const float* __restrict__ vis,
                                                          const int* __restrict__ cp_ind,
                                                          const double* __restrict__ cp_sign,
                                                          const int* __restrict__ ca_ind, float* __restrict__ visamp,
                                                          double* __restrict__ cphase, float* __restrict__ logcamp) {
  const int b = blockIdx.x;
  const float* re = vis + (size_t)b * 2 * n_vis;
  const float* im = re + n_vis;
  if (visamp)
    for (int k = threadIdx.x; k < n_vis; k += blockDim.x)
      visamp[(size_t)b * n_vis + k] = sqrtf(re[k] * re[k] + im[k] * im[k] + kEps);
  if (cphase)
    for (int t = threadIdx.x; t < n_cp; t += blockDim.x) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int k = cp_ind[q * n_cp + t];
        s += cp_sign[q * n_cp + t] * (double)atan2f(im[k], re[k]);
      }
      cphase[(size_t)b * n_cp + t] = s * 180.0 / kPi;
    }
  if (logcamp)
    for (int t = threadIdx.x; t < n_ca; t += blockDim.x) {
      float a[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = ca_ind[q * n_ca + t];
        a[q] = sqrtf(re[k] * re[k] + im[k] * im[k] + kEps);
      }
      logcamp[(size_t)b * n_ca + t] = logf(a[0]) + logf(a[1]) - logf(a[2]) - logf(a[3]);
    }
}

struct ClosureLoss {  // fused chi^2 configuration (all device pointers)
  const float* cphase_true; const float* sigma_cp; const float* logcamp_true; const float* sigma_ca;
  float w_cp, w_ca;     // d(total loss)/d(loss_cp[b]) and d(total loss)/d(loss_ca[b])
  double* loss_cp; float* loss_ca;
};

// Backward of closure_fwd for one sample per CTA. Upstream gradients are either read from the
// caller (API mode) or produced in place from the chi^2 terms (FUSED: Loss_angle_diff +
// Loss_logca_diff2 with their weights). Output: interleaved g_vis[b, k, {re, im}].
template <bool FUSED>
__global__ void __launch_bounds__(256) closure_bwd_kernel(int n_vis, int n_cp, int n_ca, const float* __restrict__ vis,
                                                          const int* __restrict__ cp_ind,
                                                          const double* __restrict__ cp_sign,
                                                          const int* __restrict__ ca_ind,
                                                          const int* __restrict__ cp_ptr, const int* __restrict__ cp_adj,
                                                          const double* __restrict__ cp_adj_sign,
                                                          const int* __restrict__ ca_ptr, const int* __restrict__ ca_adj,
                                                          const float* __restrict__ ca_adj_sign,
                                                          const float* __restrict__ g_vis_in,
                                                          const float* __restrict__ g_visamp,
                                                          const double* __restrict__ g_cphase,
                                                          const float* __restrict__ g_logcamp, ClosureLoss cl,
                                                          float* __restrict__ g_vis_il) {
  extern __shared__ double sdyn[];
  double* s_gcp = sdyn;                    // [n_cp] dL/dcphase (per degree)
  float* s_gca = (float*)(sdyn + n_cp);    // [n_ca] dL/dlogcamp
  __shared__ double redd[32];
  __shared__ float redf[32];
  const int b = blockIdx.x;
  const float* re = vis + (size_t)b * 2 * n_vis;
  const float* im = re + n_vis;
  if (FUSED) {
    double lcp = 0.0;
    for (int t = threadIdx.x; t < n_cp; t += blockDim.x) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int k = cp_ind[q * n_cp + t];
        s += cp_sign[q * n_cp + t] * (double)atan2f(im[k], re[k]);
      }
      const double pred = s * 180.0 / kPi;
      // Loss_angle_diff: y_true*pi/180 and (sigma*pi/180)**2 are formed in fp32, the rest in fp64
      const float at = cl.cphase_true[t] * 3.14159274101257324f / 180.f;
      const float sg = cl.sigma_cp[t] * 3.14159274101257324f / 180.f;
      const double den = (double)(sg * sg);
      const double u = (double)at - pred * kPi / 180.0;
      lcp += (1.0 - cos(u)) / den;
      s_gcp[t] = (double)cl.w_cp * (2.0 / n_cp) * (-sin(u)) * (kPi / 180.0) / den;
    }
    float lca = 0.f;
    for (int t = threadIdx.x; t < n_ca; t += blockDim.x) {
      float a[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = ca_ind[q * n_ca + t];
        a[q] = sqrtf(re[k] * re[k] + im[k] * im[k] + kEps);
      }
      const float pred = logf(a[0]) + logf(a[1]) - logf(a[2]) - logf(a[3]);
      const float d = cl.logcamp_true[t] - pred, s2 = cl.sigma_ca[t] * cl.sigma_ca[t];
      lca += d * d / s2;
      s_gca[t] = cl.w_ca * (-2.f * d / s2) / (float)n_ca;
    }
    const double tcp = block_sum(lcp, redd);
    const float tca = block_sum(lca, redf);
    if (threadIdx.x == 0) {
      cl.loss_cp[b] = 2.0 * tcp / n_cp;
      cl.loss_ca[b] = tca / (float)n_ca;
    }
  } else {
    for (int t = threadIdx.x; t < n_cp; t += blockDim.x) s_gcp[t] = g_cphase ? g_cphase[(size_t)b * n_cp + t] : 0.0;
    for (int t = threadIdx.x; t < n_ca; t += blockDim.x) s_gca[t] = g_logcamp ? g_logcamp[(size_t)b * n_ca + t] : 0.f;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < n_vis; k += blockDim.x) {
    const float r = re[k], i = im[k];
    const float m2 = r * r + i * i;
    double gang = 0.0;
    for (int p = cp_ptr[k]; p < cp_ptr[k + 1]; ++p) gang += cp_adj_sign[p] * s_gcp[cp_adj[p]];
    const float ga = (float)(gang * 180.0 / kPi);
    float gl = 0.f;
    for (int p = ca_ptr[k]; p < ca_ptr[k + 1]; ++p) gl += ca_adj_sign[p] * s_gca[ca_adj[p]];
    float gr = 0.f, gi = 0.f;
    if (ga != 0.f) { gr += ga * (-i / m2); gi += ga * (r / m2); }
    const float a2 = m2 + kEps;
    gr += gl * r / a2;
    gi += gl * i / a2;
    if (g_visamp) {
      const float gv = g_visamp[(size_t)b * n_vis + k] / sqrtf(a2);
      gr += gv * r;
      gi += gv * i;
    }
    if (g_vis_in) {
      gr += g_vis_in[(size_t)b * 2 * n_vis + k];
      gi += g_vis_in[(size_t)b * 2 * n_vis + n_vis + k];
    }
    g_vis_il[((size_t)b * n_vis + k) * 2] = gr;
    g_vis_il[((size_t)b * n_vis + k) * 2 + 1] = gi;
  }
}

// chi^2 terms, one CTA per sample.
__global__ void __launch_bounds__(256) chi2_kernel(int mode, int n, const float* __restrict__ y_true,
                                                   const void* __restrict__ y_pred_, const float* __restrict__ sigma,
                                                   void* __restrict__ loss_, const void* __restrict__ g_loss_,
                                                   void* __restrict__ g_pred_) {
  __shared__ double redd[32];
  const int b = blockIdx.x;
  double acc = 0.0;
  if (mode == DPI_CHI2_ANGLE) {
    const double* yp = (const double*)y_pred_ + (size_t)b * n;
    const double gl = g_loss_ ? ((const double*)g_loss_)[b] : 0.0;
    double* gp = g_pred_ ? (double*)g_pred_ + (size_t)b * n : nullptr;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
      const float at = y_true[t] * 3.14159274101257324f / 180.f;
      const float sg = sigma[t] * 3.14159274101257324f / 180.f;
      const double den = (double)(sg * sg);
      const double u = (double)at - yp[t] * kPi / 180.0;
      acc += (1.0 - cos(u)) / den;
      if (gp) gp[t] = gl * (2.0 / n) * (-sin(u)) * (kPi / 180.0) / den;
    }
    acc = block_sum(acc, redd);
    if (threadIdx.x == 0) ((double*)loss_)[b] = 2.0 * acc / n;
    return;
  }
  const float gl = g_loss_ ? ((const float*)g_loss_)[b] : 0.f;
  float s = 0.f;
  if (mode == DPI_CHI2_VIS) {
    const float* yp = (const float*)y_pred_ + (size_t)b * 2 * n;
    float* gp = g_pred_ ? (float*)g_pred_ + (size_t)b * 2 * n : nullptr;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
      const float dr = y_true[t] - yp[t], di = y_true[n + t] - yp[n + t], s2 = sigma[t] * sigma[t];
      s += (dr * dr + di * di) / s2;
      if (gp) { gp[t] = gl * (-2.f * dr / s2) / (float)n; gp[n + t] = gl * (-2.f * di / s2) / (float)n; }
    }
  } else {
    const float* yp = (const float*)y_pred_ + (size_t)b * n;
    float* gp = g_pred_ ? (float*)g_pred_ + (size_t)b * n : nullptr;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
      const float s2 = sigma[t] * sigma[t];
      if (mode == DPI_CHI2_SQ) {
        const float d = y_true[t] - yp[t];
        s += d * d / s2;
        if (gp) gp[t] = gl * (-2.f * d / s2) / (float)n;
      } else {
        const float w = y_true[t] * y_true[t] / s2, d = logf(y_true[t]) - logf(yp[t]);
        s += w * d * d;
        if (gp) gp[t] = gl * w * 2.f * d * (-1.f / yp[t]) / (float)n;
      }
    }
  }
  acc = block_sum((double)s, redd);
  if (threadIdx.x == 0) ((float*)loss_)[b] = (float)(acc / n);
}

size_t eht_ws_floats(const dpi_eht_s* h, int batch, DftPlan* f, DftPlan* bw) {
  DftPlan pf = plan_dft(batch, 2 * h->n_vis, h->P, h->n_sm);
  DftPlan pb = plan_dft(batch, h->P, 2 * h->n_vis, h->n_sm);
  if (f) *f = pf;
  if (bw) *bw = pb;
  size_t part_f = pf.splits > 1 ? (size_t)pf.tiles_m * pf.tiles_n * pf.splits * kBM * 64 : 0;
  size_t part_b = pb.splits > 1 ? (size_t)pb.tiles_m * pb.tiles_n * pb.splits * kBM * 64 : 0;
  return std::max(part_f, part_b) + (size_t)batch * 2 * h->n_vis + 64;
}
// scratch of the tcgen05 DFT (packed A + split partials), placed behind the region above
size_t eht_tc_floats(const dpi_eht_s* h, int batch) {
  if (!h->use_tc) return 0;
  const tc::Plan f = tc::make_plan(batch, 2 * h->n_vis, h->P, h->n_sm), b = tc::make_plan(batch, h->P, 2 * h->n_vis, h->n_sm);
  return std::max(tc::a_pack_floats(f) + tc::part_floats(f), tc::a_pack_floats(b) + tc::part_floats(b)) + 256;
}
float* eht_tc_base(const dpi_eht_s* h, int batch, void* ws) {
  float* base = (float*)((char*)ws + kCtrl * sizeof(unsigned)) + eht_ws_floats(h, batch, nullptr, nullptr);
  return (float*)(((uintptr_t)base + 1023) & ~(uintptr_t)1023);
}

int run_dft_fwd(dpi_eht_s* h, int batch, const float* img, float* vis, void* ws, cudaStream_t st) {
  if (h->use_tc) {
    const tc::Plan p = tc::make_plan(batch, 2 * h->n_vis, h->P, h->n_sm);
    float* a_pack = eht_tc_base(h, batch, ws);
    return tc::gemm(p, img, h->P, h->d_Bf, a_pack, a_pack + tc::a_pack_floats(p), vis, 0, st);
  }
  DftPlan pf;
  eht_ws_floats(h, batch, &pf, nullptr);
  unsigned* ctr = (unsigned*)ws;
  float* part = (float*)((char*)ws + kCtrl * sizeof(unsigned)) + (size_t)batch * 2 * h->n_vis;
  DPI_REQUIRE(pf.splits == 1 || pf.tiles_m * pf.tiles_n <= kCtrl, "eht: too many tiles");
  DPI_CUDA(cudaMemsetAsync(ctr, 0, kCtrl * sizeof(unsigned), st));
  dft_gemm_kernel<false><<<pf.tiles_m * pf.tiles_n * pf.splits, kThreads, 0, st>>>(
      batch, 2 * h->n_vis, h->P, img, h->d_F, vis, 0, pf, part, ctr);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int run_dft_bwd(dpi_eht_s* h, int batch, float* g_img, void* ws, cudaStream_t st) {
  if (h->use_tc) {
    const tc::Plan p = tc::make_plan(batch, h->P, 2 * h->n_vis, h->n_sm);
    const float* gvis_tc = (const float*)((char*)ws + kCtrl * sizeof(unsigned));
    float* a_pack = eht_tc_base(h, batch, ws);
    return tc::gemm(p, gvis_tc, 2 * h->n_vis, h->d_Bb, a_pack, a_pack + tc::a_pack_floats(p), g_img, 1, st);
  }
  DftPlan pb;
  eht_ws_floats(h, batch, nullptr, &pb);
  unsigned* ctr = (unsigned*)ws;
  float* gvis = (float*)((char*)ws + kCtrl * sizeof(unsigned));
  float* part = gvis + (size_t)batch * 2 * h->n_vis;
  DPI_REQUIRE(pb.splits == 1 || pb.tiles_m * pb.tiles_n <= kCtrl, "eht: too many tiles");
  DPI_CUDA(cudaMemsetAsync(ctr, 0, kCtrl * sizeof(unsigned), st));
  dft_gemm_kernel<true><<<pb.tiles_m * pb.tiles_n * pb.splits, kThreads, 0, st>>>(
      batch, h->P, 2 * h->n_vis, gvis, h->d_F, g_img, 1, pb, part, ctr);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

template <typename T>
int upload(T** dst, const std::vector<T>& v) {
  const size_t n = std::max<size_t>(1, v.size());
  DPI_CUDA(cudaMalloc(dst, n * sizeof(T)));
  if (!v.empty()) DPI_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return DPI_OK;
}

}  // namespace
}  // namespace dpi

using namespace dpi;

extern "C" {

int dpi_eht_create(dpi_eht_t* out, int device, int npix, int n_vis, const float* dft_mat, int n_cp,
                   const int64_t* cp_ind, const double* cp_sign, int n_ca, const int64_t* ca_ind) {
  DPI_REQUIRE(out && dft_mat && npix >= 2 && n_vis >= 1 && n_cp >= 0 && n_ca >= 0, "dpi_eht_create: bad argument");
  DPI_REQUIRE((n_cp == 0 || (cp_ind && cp_sign)) && (n_ca == 0 || ca_ind), "dpi_eht_create: null table");
  DeviceGuard g(device);
  DPI_REQUIRE(g.ok, "dpi_eht_create: cannot select device %d", device);
  std::unique_ptr<dpi_eht_s> h(new dpi_eht_s());
  h->device = device; h->npix = npix; h->P = npix * npix; h->n_vis = n_vis; h->n_cp = n_cp; h->n_ca = n_ca;
  h->n_sm = sm_count(device);
  for (int i = 0; i < 3 * n_cp; ++i) DPI_REQUIRE(cp_ind[i] >= 0 && cp_ind[i] < n_vis, "dpi_eht_create: cphase index out of range");
  for (int i = 0; i < 4 * n_ca; ++i) DPI_REQUIRE(ca_ind[i] >= 0 && ca_ind[i] < n_vis, "dpi_eht_create: camp index out of range");
  const size_t nF = (size_t)h->P * 2 * n_vis;
  DPI_CUDA(cudaMalloc(&h->d_F, nF * sizeof(float)));
  DPI_CUDA(cudaMemcpy(h->d_F, dft_mat, nF * sizeof(float), cudaMemcpyHostToDevice));
  {  // DPI_B200_DFT=simt selects the fp32 SIMT GEMM instead of the tcgen05 one
    const char* e = getenv("DPI_B200_DFT");
    h->use_tc = (e && e[0] == 's') ? 0 : 1;
  }
  if (h->use_tc) {
    const int N2 = 2 * n_vis;
    int rc = tc::init();
    if (rc) return rc;
    DPI_CUDA(cudaMalloc(&h->d_Bf, tc::b_pack_floats(N2, h->P) * sizeof(float)));
    DPI_CUDA(cudaMalloc(&h->d_Bb, tc::b_pack_floats(h->P, N2) * sizeof(float)));
    rc = tc::pack_b(h->d_F, 1, N2, N2, h->P, h->d_Bf, 0);        // forward: B[n][k] = F[k][n]
    if (rc) return rc;
    rc = tc::pack_b(h->d_F, N2, 1, h->P, N2, h->d_Bb, 0);        // adjoint: B[n][k] = F[n][k]
    if (rc) return rc;
    DPI_CUDA(cudaDeviceSynchronize());
  }
  std::vector<int> cpi(3 * (size_t)n_cp), cai(4 * (size_t)n_ca);
  std::vector<double> cps(cp_sign, cp_sign + 3 * (size_t)n_cp);
  for (size_t i = 0; i < cpi.size(); ++i) cpi[i] = (int)cp_ind[i];
  for (size_t i = 0; i < cai.size(); ++i) cai[i] = (int)ca_ind[i];
  // transposed tables (entries in increasing closure index, slot order): deterministic backward
  std::vector<int> cp_ptr(n_vis + 1, 0), ca_ptr(n_vis + 1, 0);
  for (int q = 0; q < 3; ++q)
    for (int t = 0; t < n_cp; ++t)
      if (cps[(size_t)q * n_cp + t] != 0.0) cp_ptr[cpi[(size_t)q * n_cp + t] + 1]++;
  for (int q = 0; q < 4; ++q)
    for (int t = 0; t < n_ca; ++t) ca_ptr[cai[(size_t)q * n_ca + t] + 1]++;
  for (int k = 0; k < n_vis; ++k) { cp_ptr[k + 1] += cp_ptr[k]; ca_ptr[k + 1] += ca_ptr[k]; }
  std::vector<int> cp_adj(cp_ptr[n_vis]), ca_adj(ca_ptr[n_vis]);
  std::vector<double> cp_adj_sign(cp_ptr[n_vis]);
  std::vector<float> ca_adj_sign(ca_ptr[n_vis]);
  {
    std::vector<int> fill(cp_ptr.begin(), cp_ptr.end() - 1);
    for (int t = 0; t < n_cp; ++t)
      for (int q = 0; q < 3; ++q) {
        const double s = cps[(size_t)q * n_cp + t];
        if (s == 0.0) continue;
        const int k = cpi[(size_t)q * n_cp + t];
        cp_adj[fill[k]] = t; cp_adj_sign[fill[k]] = s; fill[k]++;
      }
    std::vector<int> fill2(ca_ptr.begin(), ca_ptr.end() - 1);
    for (int t = 0; t < n_ca; ++t)
      for (int q = 0; q < 4; ++q) {
        const int k = cai[(size_t)q * n_ca + t];
        ca_adj[fill2[k]] = t; ca_adj_sign[fill2[k]] = q < 2 ? 1.f : -1.f; fill2[k]++;
      }
  }
  int rc;
  if ((rc = upload(&h->d_cp_ind, cpi)) || (rc = upload(&h->d_cp_sign, cps)) || (rc = upload(&h->d_ca_ind, cai)) ||
      (rc = upload(&h->d_cp_ptr, cp_ptr)) || (rc = upload(&h->d_cp_adj, cp_adj)) ||
      (rc = upload(&h->d_cp_adj_sign, cp_adj_sign)) || (rc = upload(&h->d_ca_ptr, ca_ptr)) ||
      (rc = upload(&h->d_ca_adj, ca_adj)) || (rc = upload(&h->d_ca_adj_sign, ca_adj_sign)))
    return rc;
  handle_register(h.get(), 2);
  *out = h.release();
  return DPI_OK;
}

int dpi_eht_destroy(dpi_eht_t h) {
  if (!h) return DPI_OK;
  DPI_REQUIRE(handle_unregister(h, 2), "dpi_eht_destroy: unknown or already destroyed handle %p", (void*)h);
  DeviceGuard g(h->device);
  cudaFree(h->d_F);
  cudaFree(h->d_Bf);
  cudaFree(h->d_Bb); cudaFree(h->d_cp_ind); cudaFree(h->d_cp_sign); cudaFree(h->d_ca_ind);
  cudaFree(h->d_cp_ptr); cudaFree(h->d_cp_adj); cudaFree(h->d_cp_adj_sign);
  cudaFree(h->d_ca_ptr); cudaFree(h->d_ca_adj); cudaFree(h->d_ca_adj_sign);
  delete h;
  return DPI_OK;
}

size_t dpi_eht_workspace_bytes(dpi_eht_t h, int batch) {
  if (!handle_live(h, 2) || batch < 1) return 0;
  return kCtrl * sizeof(unsigned) + (eht_ws_floats(h, batch, nullptr, nullptr) + eht_tc_floats(h, batch)) * sizeof(float) + 1024;
}

int dpi_eht_forward(dpi_eht_t h, int batch, const float* img, float* vis, float* visamp, double* cphase,
                    float* logcamp, void* workspace, size_t workspace_bytes, dpi_stream_t stream) {
  DPI_REQUIRE(handle_live(h, 2), "dpi_eht_forward: null, unknown or destroyed operator handle");
  DPI_REQUIRE(img && vis && workspace && batch >= 1, "dpi_eht_forward: bad argument");
  DPI_REQUIRE(workspace_bytes >= dpi_eht_workspace_bytes(h, batch), "dpi_eht_forward: workspace too small");
  DeviceGuard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = run_dft_fwd(h, batch, img, vis, workspace, st);
  if (rc) return rc;
  if (visamp || cphase || logcamp) {
    closure_fwd_kernel<<<batch, 256, 0, st>>>(h->n_vis, h->n_cp, h->n_ca, vis, h->d_cp_ind, h->d_cp_sign, h->d_ca_ind,
                                             visamp, cphase, logcamp);
    DPI_LAUNCH_CHECK();
  }
  return DPI_OK;
}

int dpi_eht_backward(dpi_eht_t h, int batch, const float* vis, const float* g_vis, const float* g_visamp,
                     const double* g_cphase, const float* g_logcamp, float* g_img, void* workspace,
                     size_t workspace_bytes, dpi_stream_t stream) {
  DPI_REQUIRE(handle_live(h, 2), "dpi_eht_backward: null, unknown or destroyed operator handle");
  DPI_REQUIRE(vis && g_img && workspace && batch >= 1, "dpi_eht_backward: bad argument");
  DPI_REQUIRE(workspace_bytes >= dpi_eht_workspace_bytes(h, batch), "dpi_eht_backward: workspace too small");
  DeviceGuard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  float* gvis = (float*)((char*)workspace + kCtrl * sizeof(unsigned));
  const size_t smem = (size_t)h->n_cp * sizeof(double) + (size_t)h->n_ca * sizeof(float);
  DPI_REQUIRE(smem <= 48 * 1024, "dpi_eht_backward: closure tables too large for shared memory");
  ClosureLoss cl{};
  closure_bwd_kernel<false><<<batch, 256, smem, st>>>(h->n_vis, h->n_cp, h->n_ca, vis, h->d_cp_ind, h->d_cp_sign,
                                                     h->d_ca_ind, h->d_cp_ptr, h->d_cp_adj, h->d_cp_adj_sign,
                                                     h->d_ca_ptr, h->d_ca_adj, h->d_ca_adj_sign, g_vis, g_visamp,
                                                     g_cphase, g_logcamp, cl, gvis);
  DPI_LAUNCH_CHECK();
  return run_dft_bwd(h, batch, g_img, workspace, st);
}

/* Fused data term of DPI_interferometry.py:221-229: given vis (from dpi_eht_forward with NULL
 * closure outputs), computes loss_cp[b] (fp64) = Loss_angle_diff, loss_ca[b] = Loss_logca_diff2 and
 * g_img = d(w_cp * sum_b loss_cp + w_ca * sum_b loss_ca)/d img in one closure pass + one GEMM. */
int dpi_eht_data_grad(dpi_eht_t h, int batch, const float* vis, const float* cphase_true, const float* sigma_cp,
                      const float* logcamp_true, const float* sigma_ca, float w_cp, float w_ca, double* loss_cp,
                      float* loss_ca, float* g_img, void* workspace, size_t workspace_bytes, dpi_stream_t stream) {
  DPI_REQUIRE(handle_live(h, 2), "dpi_eht_data_grad: null, unknown or destroyed operator handle");
  DPI_REQUIRE(h && vis && cphase_true && sigma_cp && logcamp_true && sigma_ca && loss_cp && loss_ca && g_img &&
                  workspace && batch >= 1, "dpi_eht_data_grad: bad argument");
  DPI_REQUIRE(workspace_bytes >= dpi_eht_workspace_bytes(h, batch), "dpi_eht_data_grad: workspace too small");
  DeviceGuard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  float* gvis = (float*)((char*)workspace + kCtrl * sizeof(unsigned));
  const size_t smem = (size_t)h->n_cp * sizeof(double) + (size_t)h->n_ca * sizeof(float);
  DPI_REQUIRE(smem <= 48 * 1024, "dpi_eht_data_grad: closure tables too large for shared memory");
  ClosureLoss cl{cphase_true, sigma_cp, logcamp_true, sigma_ca, w_cp, w_ca, loss_cp, loss_ca};
  closure_bwd_kernel<true><<<batch, 256, smem, st>>>(h->n_vis, h->n_cp, h->n_ca, vis, h->d_cp_ind, h->d_cp_sign,
                                                    h->d_ca_ind, h->d_cp_ptr, h->d_cp_adj, h->d_cp_adj_sign,
                                                    h->d_ca_ptr, h->d_ca_adj, h->d_ca_adj_sign, nullptr, nullptr,
                                                    nullptr, nullptr, cl, gvis);
  DPI_LAUNCH_CHECK();
  return run_dft_bwd(h, batch, g_img, workspace, st);
}

int dpi_chi2(int mode, int batch, int n, const float* y_true, const void* y_pred, const float* sigma, void* loss,
             const void* g_loss, void* g_pred, dpi_stream_t stream) {
  DPI_REQUIRE(mode >= 0 && mode <= 3 && batch >= 1 && n >= 1 && y_true && y_pred && sigma && loss, "dpi_chi2: bad argument");
  DPI_REQUIRE(!g_pred || g_loss, "dpi_chi2: g_pred requested without g_loss");
  chi2_kernel<<<batch, 256, 0, (cudaStream_t)stream>>>(mode, n, y_true, y_pred, sigma, loss, g_loss, g_pred);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

}  // extern "C"
