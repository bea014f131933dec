// flow_tc.cuh - the coupling-net GEMMs of the RealNVP flow (generative_model/realnvpfc_model.py:171-187,
// AffineCoupling.net: Linear(Da->H), Linear(H->H), ZeroFC(H->Dout)) and their data / weight gradients on the
// 5th-generation tensor cores, for batches that span at least one 128-row UMMA tile.
//
//   C[M, N] = A[M, K] * B[N, K]^T      fp32 in / out, 3xTF32 (hi/lo split, three tcgen05.mma per k-step, fp32
//                                      accumulation in TMEM): fp32-level accuracy, ~2^-21 per product
//
// Both operands reach shared memory as pre-split, pre-tiled K-major tiles (canonical no-swizzle UMMA layout:
// core matrix 8 rows x 16 B, LBO 128 B, SBO = 32 KT B), one cp.async.bulk per tile and stage, mbarrier expect_tx
// ring; one elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (M 128, N = TN in {64, 128, 256}, K 8); the
// accumulator(s) are read back with tcgen05.ld 32x32b. The flow keeps its activations FEATURE-major (X^T[feature][sample], see flow.cuh),
// so the pack pre-pass is also where the layout turns K-major, where the fixed permutations of the flow are
// applied (a row map on the feature index) and where fp32 is split into TF32 hi / lo.
//
//   forward  : samples are the UMMA M (TMEM lane = sample -> feature-major stores are coalesced along samples)
//   dgrad    : same, with B[n][k] = W[k][n] (transposing pack of the weight)
//   wgrad    : gW[i][j] = sum_m G^T[i][m] X^T[j][m]: both operands are already K-major (K = batch)
//
// Split-K partial tiles are summed in split order by the epilogue kernel (deterministic), which also applies
// the per-layer epilogue: bias + LeakyReLU + BatchNorm (batch or running statistics), BatchNorm + LeakyReLU
// backward with the gamma / beta / bias gradients, the coupling tail (ZeroFC scale, tanh-exp affine, ActNorm,
// logdet partials), the row-mapped scatter with the ActNorm pass-through term, or the row-major weight-gradient
// store. Host-side planning (make_gemm) needs no device and is exposed as dpi_flow_tc_plan for the CPU tests.
#pragma once
#include "common.cuh"

namespace dpi {
namespace ftc {

constexpr int kM = 128;   // UMMA M
constexpr int kMaxStages = 6;

// CTA tile (MH x 128) x TN, k-tile KT per pipeline stage (KT / 8 UMMA k-steps). m_tiles counts 128-row tiles
// (padded to a multiple of MH), k_tiles counts KT-deep tiles.
struct Gemm {
  int M, N, K, TN, KT, MH, m_tiles, n_tiles, k_tiles, splits, kt_per_split, stages;
};
Gemm make_gemm(int M, int N, int K, int n_sm);
inline size_t a_pack_floats(const Gemm& g) { return (size_t)g.m_tiles * g.k_tiles * 2 * kM * g.KT; }
inline size_t b_pack_floats(const Gemm& g) { return (size_t)g.n_tiles * g.k_tiles * 2 * g.TN * g.KT; }
inline size_t part_floats(const Gemm& g) { return (size_t)g.splits * g.m_tiles * g.n_tiles * kM * g.TN; }

// element(r, k) of the operand = k_contig ? src[(map ? map[r] : r) * ld + k] : src[(map ? map[k] : k) * ld + r];
// zero outside rows x K. dst: [row tiles][k tiles][hi, lo][R x 32 canonical].
struct Pack {
  const float* src;
  long long ld;
  int k_contig;
  const int* map;
  int rows, K;
  // im2col (map unused): src is a feature-major image stack [channel][b * h * w] with row stride ld. k_contig = 0:
  // r = position (b, y, x), k = (channel, dy, dx) of a 3 x 3 window; k_contig = 1 (weight-gradient operand): r =
  // (channel, dy, dx), k = position. element = src[channel][b, y + dy - 1, x + dx - 1], or `pad` outside the image
  // (glow_model.py ZeroConv2d pads with ONES, :247)
  int im_h, im_w;      // 0: plain operand
  float pad;
};
// packs A ((MH x 128)-row tiles) and B (TN-row tiles) of one GEMM in one launch
int pack2(const Gemm& g, const Pack& pa, float* dst_a, const Pack& pb, float* dst_b, cudaStream_t st);

enum EpiMode {
  EPI_T_RAW = 0,   // out[n * ldo + m] = c
  EPI_T_HID,       // out[n * ldo + m] = v = lrelu(c + bias[n]); out2 = v (bn_mode 0), BatchNorm with batch statistics
                   // (bn_mode 1: statistics saved, running statistics updated) or with running statistics (bn_mode 2)
  EPI_T_SCATTER,   // out[rowmap[n] * ldo + m] = c + add[n * ldo + m] * exp(*lsi)
  EPI_T_BNBWD,     // out[n * ldo + m] = LeakyReLU'(l) * BatchNorm-backward(c) (bn_mode 1) or LeakyReLU'(l) * c (bn_mode 0),
                   // plus the gradients of gamma / beta / the Linear bias of that feature row
  EPI_ROWMAJOR,    // out[m * ldo + n] = c
  EPI_T_BIAS,      // out[n * ldo + m] = (c + bias[n]) * exp(3 * scale[n])   (bias / scale may be null): conv + ZeroConv2d scale
  EPI_T_AFFINE,    // out[n * ldo + m] = c * gamma[n] - beta[n]   (Glow: 1x1 inverse conv followed by ActNorm.reverse)
  EPI_T_TAIL       // ZeroFC bias / exp(3 scale), tanh-exp affine coupling, ActNorm, logdet partials (AffineCoupling
                   // .reverse / .forward, realnvpfc_model.py:240-249 / :222-231) on the raw ZeroFC GEMM result
};
constexpr int kEpiMaxM = 4096;   // the fused BatchNorm epilogues keep one feature row (all samples) in registers
struct Epi {
  int mode;
  float* out;
  long long ldo;
  const float* bias;
  float* out2;
  int bn_mode;
  const float *gamma, *beta;
  float *rmean, *rvar;          // running statistics (bn_mode 1: updated, 2: read)
  float *mean, *rstd;           // batch statistics of the row (HID bn_mode 1: written, BNBWD: read)
  const int* rowmap;
  const float* add;
  const float* lsi;
  const float* lsaved;          // BNBWD: the saved LeakyReLU output L^T
  float *g_gamma, *g_beta, *g_bias;
  // TAIL: out = O^T (rows j: log_s pre-tanh, rows Db + j: t), y = stage output, yprev = stage input (rows through gmap)
  const float* yprev;
  float* y;
  const int* gmap;
  const float *scale, *b3;      // net.6.scale, net.6.fc.bias
  const float *post_lsi, *post_loc;
  int post_mode;                // 0: y = w * exp(lsi) - loc (reverse); 1: y = (w + loc) / exp(lsi) (forward, next stage's
                                // ActNorm); 2: identity
  int dir, Da, Db;
  float* ldp;                   // logdet partials of this stage: [Db / 16 chunks][ldo]
  // HID / BNBWD: also emit the result as the packed A operand (TF32 hi / lo, canonical K-major tiles of tp_R rows x
  // tp_KT k, k = this epilogue's feature index) of the GEMM that consumes it next - no pack pass in between
  float* tp;
  int tp_R, tp_KT, tp_ktiles;
  float slope, eps;             // HID / BNBWD: LeakyReLU slope and BatchNorm eps; 0 = the RealNVP constants (0.01, 1e-2)
};
constexpr int kTailCols = 16;   // coupling columns per logdet partial
int run(const Gemm& g, const float* a_pack, const float* b_pack, float* part, const Epi& e, cudaStream_t st);

// The weights of EVERY stage packed as B operands in one launch (once per pass instead of once per GEMM).
// Three layers per stage; layer l of stage s reads P + src_off[s * 3 + l] (rows x K, element (r, k) as in Pack with
// the given ld / k_contig) and writes dst + s * stage_floats + dst_off[l], tiled for GEMM cfg[l].
struct WeightPack {
  int S;
  const long long* src_off;     // device: [S * 3] float offsets into the flat parameter buffer
  long long stage_floats;
  long long dst_off[3];
  int rows[3], K[3], ld[3], k_contig[3];
  int TN[3], KT[3];
  int tiles[3];                 // tiles per layer = ceil(rows / TN) * ceil(K / KT)
};
int pack_weights(const WeightPack& w, const float* P, float* dst, cudaStream_t st);
int init();
void configure();   // reads DPI_B200_TC_TN / _KT / _MH (call before make_gemm)

}  // namespace ftc
}  // namespace dpi
