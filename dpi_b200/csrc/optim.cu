// optim.cu - global-norm gradient clipping + Adam on flat fp32 buffers.
//   torch.nn.utils.clip_grad_norm_ : coef = min(1, max_norm / (||g|| + 1e-6))
//   torch.optim.Adam (defaults, no weight decay / amsgrad):
//     m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
// as used at DPI_interferometry.py:172,233-235 / DPI_MRI.py:146,205-207.
// HBM-bound: 4 reads + 3 writes of P floats; float4 accesses, grid = multiple of the SM count.
#include <algorithm>

#include "common.cuh"

namespace dpi {
namespace {

constexpr int kT = 256;
constexpr int kMaxBlocks = 1184;  // 8 x 148

__global__ void __launch_bounds__(kT) sqnorm_kernel(long long n, const float* __restrict__ g, double* __restrict__ part,
                                                    unsigned* __restrict__ counter, double* __restrict__ out) {
  __shared__ double red[32];
  __shared__ int s_last;
  double s = 0.0;
  const long long n4 = n / 4;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < n4; i += (long long)gridDim.x * kT) {
    const float4 v = __ldcg(g4 + i);
    float q = v.x * v.x;
    q = fmaf(v.y, v.y, q); q = fmaf(v.z, v.z, q); q = fmaf(v.w, v.w, q);
    s += (double)q;
  }
  if (blockIdx.x == 0)
    for (long long i = n4 * 4 + threadIdx.x; i < n; i += kT) s += (double)g[i] * (double)g[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    __stcg(part + blockIdx.x, s);
    __threadfence();
    const unsigned old = atomicAdd(counter, 1u);
    s_last = (old == gridDim.x - 1);
    if (s_last) *counter = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double t = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += kT) t += __ldcg(part + i);
  t = block_sum(t, red);
  if (threadIdx.x == 0) out[0] = t;
}

__global__ void __launch_bounds__(kT) clip_adam_kernel(long long n, float* __restrict__ p, const float* __restrict__ g,
                                                       float* __restrict__ m, float* __restrict__ v,
                                                       const double* __restrict__ sq, int n_sq, float clip, float lr,
                                                       float b1, float b2, float eps,
                                                       const long long* __restrict__ step_dev, float grad_scale,
                                                       float* __restrict__ norm_out) {
  double tot = 0.0;
  for (int i = 0; i < n_sq; ++i) tot += sq[i];
  // gradients are averaged over ranks before the norm: ||grad_scale * g||
  const float norm = (float)(sqrt(tot) * (double)grad_scale);
  float coef = clip / (norm + 1e-6f);
  coef = coef < 1.0f ? coef : 1.0f;
  if (clip <= 0.f) coef = 1.0f;  // clipping disabled
  const float gs = coef * grad_scale;
  const long long step = *step_dev;
  const double bc1 = 1.0 - pow((double)b1, (double)step), bc2 = 1.0 - pow((double)b2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  if (blockIdx.x == 0 && threadIdx.x == 0 && norm_out) { norm_out[0] = norm; norm_out[1] = coef; }
  const long long n4 = n / 4;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg *= gs;
    mm = fmaf(1.f - b1, gg - mm, mm);         // exp_avg.lerp_(grad, 1 - beta1)
    vv = vv * b2 + (1.f - b2) * gg * gg;      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pp -= step_size * (mm / denom);
  };
  for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < n4; i += (long long)gridDim.x * kT) {
    float4 P = p4[i], M = m4[i], V = v4[i];
    const float4 G = __ldcs(g4 + i);
    upd(P.x, G.x, M.x, V.x); upd(P.y, G.y, M.y, V.y); upd(P.z, G.z, M.z, V.z); upd(P.w, G.w, M.w, V.w);
    p4[i] = P; m4[i] = M; v4[i] = V;
  }
  if (blockIdx.x == 0)
    for (long long i = n4 * 4 + threadIdx.x; i < n; i += kT) upd(p[i], g[i], m[i], v[i]);
}

__global__ void step_increment_kernel(long long* step) { *step += 1; }

}  // namespace
}  // namespace dpi

using namespace dpi;

extern "C" {

size_t dpi_clip_adam_workspace_bytes(int64_t n) { (void)n; return 64 + kMaxBlocks * sizeof(double); }

int dpi_grad_sqnorm(int64_t n, const float* g, double* sq_out, void* workspace, size_t workspace_bytes,
                    dpi_stream_t stream) {
  DPI_REQUIRE(n >= 1 && g && sq_out && workspace, "dpi_grad_sqnorm: bad argument");
  DPI_REQUIRE(workspace_bytes >= dpi_clip_adam_workspace_bytes(n), "dpi_grad_sqnorm: workspace too small");
  DPI_REQUIRE(((uintptr_t)g & 15) == 0, "dpi_grad_sqnorm: gradient buffer must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned* counter = (unsigned*)workspace;
  double* part = (double*)((char*)workspace + 64);
  int sms = current_sm_count();
  long long want = (n / 4 + kT * 4 - 1) / (kT * 4);
  int blocks = (int)std::max(1LL, std::min<long long>(want, std::min(kMaxBlocks, std::max(1, sms) * 8)));
  DPI_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned), st));
  sqnorm_kernel<<<blocks, kT, 0, st>>>(n, g, part, counter, sq_out);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int dpi_clip_adam(int64_t n, float* p, const float* g, float* m, float* v, const double* sq_norms, int n_sq, float clip,
                  float lr, float beta1, float beta2, float eps, const int64_t* step_dev, float grad_scale,
                  float* norm_out, dpi_stream_t stream) {
  DPI_REQUIRE(n >= 1 && p && g && m && v && sq_norms && n_sq >= 1 && step_dev, "dpi_clip_adam: bad argument");
  DPI_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0 || n < 4,
              "dpi_clip_adam: buffers must be 16-byte aligned");
  int sms = current_sm_count();
  long long want = (n / 4 + kT * 2 - 1) / (kT * 2);
  int blocks = (int)std::max(1LL, std::min<long long>(want, std::min(kMaxBlocks, std::max(1, sms) * 8)));
  clip_adam_kernel<<<blocks, kT, 0, (cudaStream_t)stream>>>(n, p, g, m, v, sq_norms, n_sq, clip, lr, beta1, beta2, eps,
                                                           (const long long*)step_dev, grad_scale, norm_out);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

int dpi_step_increment(int64_t* step_dev, dpi_stream_t stream) {
  DPI_REQUIRE(step_dev, "dpi_step_increment: null argument");
  step_increment_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((long long*)step_dev);
  DPI_LAUNCH_CHECK();
  return DPI_OK;
}

}  // extern "C"
