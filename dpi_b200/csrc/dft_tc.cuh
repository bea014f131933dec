// dft_tc.cuh - the dense DFT of the VLBI operator (interferometry_helpers.py torch_complex_matmul :36-39,
// used at :259-260, and its adjoint) as a tcgen05 GEMM.
//
//   C[M, N] = A[M, K] * B[N, K]^T          (fp32 in, fp32 out, fp32-level accuracy through the 3xTF32 split)
//     forward : A = img   [batch, npix^2],   B[n][k] = F[k][n]   (N = 2 n_vis, K = npix^2)
//     adjoint : A = g_vis [batch, 2 n_vis],  B[n][k] = F[n][k]   (N = npix^2,  K = 2 n_vis)
//
// B is a constant of the operator: at create time it is split into TF32 hi / lo parts and stored tile by tile in
// the canonical shared-memory layout of a K-major UMMA operand (no swizzle: core matrix = 8 rows x 16 bytes,
// LBO = 128 B between the 4-k chunks, SBO = 1024 B between 8-row groups), so one cp.async.bulk per 64 x 32 tile
// brings it into shared memory ready for tcgen05.mma. A is split and tiled the same way by a small pre-pass.
// Kernel: 192 threads. Warp 4 / lane 0 = producer (cp.async.bulk + mbarrier expect_tx, 3-stage ring), warp 5 /
// lane 0 = MMA issuer (tcgen05.mma.cta_group::1.kind::tf32, M = 128, N = 64, K = 8; a_lo*b_hi + a_hi*b_lo +
// a_hi*b_hi; tcgen05.commit releases the stage). TMEM holds two 64-column accumulators: each takes the sum of 8
// k-tiles (256 k), then warps 0-3 add it to fp32 registers (tcgen05.ld 32x32b) while the other one fills - the tensor
// core's accumulate truncates, so long sums are kept out of it (dft_tc.cu). K is split over CTAs when M x N alone does
// not fill the GPU; a second tiny kernel sums the split partials in order (deterministic) into the output layout.
#pragma once
#include "common.cuh"

namespace dpi {
namespace tc {

constexpr int kTM = 128;          // UMMA M (batch rows per tile)
constexpr int kTN = 64;           // UMMA N (output columns per tile)
constexpr int kTK = 32;           // k per pipeline stage = 4 UMMA steps of 8
constexpr int kStagesTc = 3;
constexpr int kATileFloats = kTM * kTK;   // 4096
constexpr int kBTileFloats = kTN * kTK;   // 2048
constexpr int kStageBytes = (2 * kATileFloats + 2 * kBTileFloats) * 4;   // A_hi, A_lo, B_hi, B_lo = 49152
constexpr int kSmemBytes = kStagesTc * kStageBytes + 1024;               // + alignment slack

struct Plan {
  int M, N, K, m_tiles, n_tiles, k_tiles, splits, kt_per_split;
};

Plan make_plan(int M, int N, int K, int n_sm);
size_t a_pack_floats(const Plan& p);       // packed A: [m_tiles][k_tiles][hi, lo][128 x 32]
size_t b_pack_floats(int N, int K);        // packed B: [n_tiles][k_tiles][hi, lo][64 x 32]
size_t part_floats(const Plan& p);         // split partials: [splits][m_tiles][n_tiles][128 x 64]

// Create time: dst (b_pack_floats(N, K) floats) <- B[n][k] = src[n * s_n + k * s_k], split + tiled.
int pack_b(const float* src, long long s_n, long long s_k, int N, int K, float* dst, cudaStream_t st);
// Run time. mode 0: out[m][c][k'] planar with n = 2 k' + c (forward DFT: vis [B, 2, n_vis]); mode 1: out[m][n].
int gemm(const Plan& p, const float* A, int lda, const float* Bpack, float* a_pack, float* part, float* out, int mode,
         cudaStream_t st);
int init();                                // one-time kernel attributes

}  // namespace tc
}  // namespace dpi
