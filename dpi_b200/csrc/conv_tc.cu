// conv_tc.cu - the convolutions of the Glow coupling nets (generative_model/glow_model.py:260-268: Conv2d 3x3 pad 1,
// Conv2d 1x1, ZeroConv2d :238-251 = pad with ONES + Conv2d 3x3 + exp(3 scale)) as tcgen05 GEMMs over feature-major
// image stacks [channel][b * h * w]: positions are the UMMA M, the 3 x 3 window gather (im2col) happens in the
// operand pack, the weight [c_out][c_in * k * k] is already a K-major B operand, bias and the ZeroConv2d scale are
// the epilogue. Forward building block of the Glow path; 3xTF32 like the other flow GEMMs.
#include "flow_tc.cuh"

using namespace dpi;

namespace {
inline size_t align256(size_t floats) { return (floats + 255) / 256 * 256; }
}  // namespace

extern "C" {

size_t dpi_conv2d_fm_workspace_bytes(int batch, int h, int w, int c_in, int c_out, int ksize) {
  if (batch < 1 || h < 1 || w < 1 || c_in < 1 || c_out < 1 || (ksize != 1 && ksize != 3)) return 0;
  ftc::configure();
  const ftc::Gemm g = ftc::make_gemm(batch * h * w, c_out, c_in * ksize * ksize, current_sm_count());
  return (align256(ftc::a_pack_floats(g)) + align256(ftc::b_pack_floats(g)) + align256(ftc::part_floats(g))) * sizeof(float) + 1024;
}

static int conv_run(int batch, int h, int w, int c_in, int c_out, int ksize, float pad_value, const float* x, int64_t ldx,
                    const float* weight, ftc::Epi e, void* workspace, size_t workspace_bytes, dpi_stream_t stream) {
  DPI_REQUIRE(x && weight && e.out && workspace, "dpi_conv2d_fm: null argument");
  DPI_REQUIRE(batch >= 1 && h >= 1 && w >= 1 && c_in >= 1 && c_out >= 1, "dpi_conv2d_fm: bad shape");
  DPI_REQUIRE(ksize == 1 || ksize == 3, "dpi_conv2d_fm: kernel size must be 1 or 3 (got %d)", ksize);
  const long long M = (long long)batch * h * w;
  DPI_REQUIRE(M < (1LL << 30) && ldx >= M && e.ldo >= M, "dpi_conv2d_fm: bad row stride / too many positions");
  const size_t need = dpi_conv2d_fm_workspace_bytes(batch, h, w, c_in, c_out, ksize);
  if (workspace_bytes < need) {
    set_error("dpi_conv2d_fm: workspace too small (%zu < %zu)", workspace_bytes, need);
    return DPI_ERR_WORKSPACE;
  }
  int rc = ftc::init();
  if (rc) return rc;
  const int K = c_in * ksize * ksize;
  const ftc::Gemm g = ftc::make_gemm((int)M, c_out, K, current_sm_count());
  float* base = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
  float* apack = base;
  float* bpack = apack + align256(ftc::a_pack_floats(g));
  float* part = bpack + align256(ftc::b_pack_floats(g));
  ftc::Pack pa{x, (long long)ldx, 0, nullptr, (int)M, K, 0, 0, 0.f};
  if (ksize == 3) { pa.im_h = h; pa.im_w = w; pa.pad = pad_value; }
  const ftc::Pack pb{weight, (long long)K, 1, nullptr, c_out, K, 0, 0, 0.f};
  rc = ftc::pack2(g, pa, apack, pb, bpack, (cudaStream_t)stream);
  if (rc) return rc;
  return ftc::run(g, apack, bpack, part, e, (cudaStream_t)stream);
}

int dpi_conv2d_fm_forward(int batch, int h, int w, int c_in, int c_out, int ksize, float pad_value, const float* x,
                          int64_t ldx, const float* weight, const float* bias, const float* scale, float* y, int64_t ldy,
                          void* workspace, size_t workspace_bytes, dpi_stream_t stream) {
  ftc::Epi e{};
  e.mode = ftc::EPI_T_BIAS; e.out = y; e.ldo = ldy; e.bias = bias; e.scale = scale;
  return conv_run(batch, h, w, c_in, c_out, ksize, pad_value, x, ldx, weight, e, workspace, workspace_bytes, stream);
}

int dpi_conv2d_fm_lrelu(int batch, int h, int w, int c_in, int c_out, int ksize, float pad_value, const float* x, int64_t ldx,
                        const float* weight, const float* bias, float slope, float* y, int64_t ldy, void* workspace,
                        size_t workspace_bytes, dpi_stream_t stream) {
  DPI_REQUIRE(bias && slope > 0.f, "dpi_conv2d_fm_lrelu: bias and a positive slope are required");
  ftc::Epi e{};
  e.mode = ftc::EPI_T_HID; e.out = y; e.ldo = ldy; e.bias = bias; e.bn_mode = 0; e.slope = slope;
  return conv_run(batch, h, w, c_in, c_out, ksize, pad_value, x, ldx, weight, e, workspace, workspace_bytes, stream);
}

int dpi_conv2d_fm_affine(int batch, int h, int w, int c_in, int c_out, const float* x, int64_t ldx, const float* weight,
                         const float* mul, const float* sub, float* y, int64_t ldy, void* workspace, size_t workspace_bytes,
                         dpi_stream_t stream) {
  DPI_REQUIRE(mul && sub, "dpi_conv2d_fm_affine: null argument");
  ftc::Epi e{};
  e.mode = ftc::EPI_T_AFFINE; e.out = y; e.ldo = ldy; e.gamma = mul; e.beta = sub;
  return conv_run(batch, h, w, c_in, c_out, 1, 0.f, x, ldx, weight, e, workspace, workspace_bytes, stream);
}

size_t dpi_conv2d_fm_wgrad_workspace_bytes(int batch, int h, int w, int c_in, int c_out, int ksize) {
  if (batch < 1 || h < 1 || w < 1 || c_in < 1 || c_out < 1 || (ksize != 1 && ksize != 3)) return 0;
  ftc::configure();
  const ftc::Gemm g = ftc::make_gemm(c_out, c_in * ksize * ksize, batch * h * w, current_sm_count());
  return (align256(ftc::a_pack_floats(g)) + align256(ftc::b_pack_floats(g)) + align256(ftc::part_floats(g))) * sizeof(float) + 1024;
}

// gW[co][ci][dy][dx] = sum_pos g[co][pos] * xpad[ci][pos shifted by (dy - 1, dx - 1)]: M = c_out, N = c_in k k, K = positions
int dpi_conv2d_fm_wgrad(int batch, int h, int w, int c_in, int c_out, int ksize, float pad_value, const float* x, int64_t ldx,
                        const float* g_out, int64_t ldg, float* g_weight, void* workspace, size_t workspace_bytes,
                        dpi_stream_t stream) {
  DPI_REQUIRE(x && g_out && g_weight && workspace, "dpi_conv2d_fm_wgrad: null argument");
  DPI_REQUIRE(batch >= 1 && h >= 1 && w >= 1 && c_in >= 1 && c_out >= 1, "dpi_conv2d_fm_wgrad: bad shape");
  DPI_REQUIRE(ksize == 1 || ksize == 3, "dpi_conv2d_fm_wgrad: kernel size must be 1 or 3 (got %d)", ksize);
  const long long M = (long long)batch * h * w;
  DPI_REQUIRE(M < (1LL << 30) && ldx >= M && ldg >= M, "dpi_conv2d_fm_wgrad: bad row stride / too many positions");
  const size_t need = dpi_conv2d_fm_wgrad_workspace_bytes(batch, h, w, c_in, c_out, ksize);
  if (workspace_bytes < need) {
    set_error("dpi_conv2d_fm_wgrad: workspace too small (%zu < %zu)", workspace_bytes, need);
    return DPI_ERR_WORKSPACE;
  }
  int rc = ftc::init();
  if (rc) return rc;
  const int N = c_in * ksize * ksize;
  const ftc::Gemm g = ftc::make_gemm(c_out, N, (int)M, current_sm_count());
  float* base = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
  float* apack = base;
  float* bpack = apack + align256(ftc::a_pack_floats(g));
  float* part = bpack + align256(ftc::b_pack_floats(g));
  const ftc::Pack pa{g_out, (long long)ldg, 1, nullptr, c_out, (int)M, 0, 0, 0.f};
  ftc::Pack pb{x, (long long)ldx, 1, nullptr, N, (int)M, 0, 0, 0.f};
  if (ksize == 3) { pb.im_h = h; pb.im_w = w; pb.pad = pad_value; }
  rc = ftc::pack2(g, pa, apack, pb, bpack, (cudaStream_t)stream);
  if (rc) return rc;
  ftc::Epi e{};
  e.mode = ftc::EPI_ROWMAJOR; e.out = g_weight; e.ldo = N;
  return ftc::run(g, apack, bpack, part, e, (cudaStream_t)stream);
}

}  // extern "C"
