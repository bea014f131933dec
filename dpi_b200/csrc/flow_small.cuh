// flow_small.cuh - plan tables of the small-batch (batch <= 64) RealNVP kernel (flow_small.cu).
//
// At batch <= 64 every GEMM of a coupling stage (generative_model/realnvpfc_model.py:171-187) is a few
// MFLOP: the step is a chain of ~7 dependent hops per stage and what matters is the latency of one hop,
// not FLOP/s. flow_small.cu therefore runs the whole pass as ONE persistent kernel of thread-block
// clusters in which each hop is
//     grid barrier wait -> one-shot cp.async of the (L2-resident) activation rows -> fp32 FFMA micro-GEMM
//     -> split-K reduction over the 8 CTAs of a cluster through distributed shared memory -> fused epilogue
//     -> grid barrier arrive -> [shadow: prefetch next hop's weights into smem, weight gradients] .
#pragma once

namespace dpi {

// One "activation GEMM" out[n][m] = sum_k W(n, k) * X[k][m]  (m = sample, <= 64).
struct ActPlan {
  int N, K;        // output features, reduction length
  int NT;          // features per tile (4..64, power of two); F3: rows per tile = 2 * pairs
  int n_tiles;     // ceil(N / NT) (F3: ceil(Db / (NT/2)))
  int ksplit;      // 1: the 8 CTAs of a cluster split K and reduce through DSMEM; 0: one CTA, full K
  int KC;          // K extent of one rank (multiple of 4)
  int KS;          // rows per shared-memory chunk (multiple of 4, <= KC)
  int nq, kp;      // feature quads per tile, k-parts per CTA (nq * kp == 16)
  int nt_log, nq_log;  // log2(NT), log2(nq)
};

struct SmallPlan {
  int ok;                      // 0: shape not supported -> tiled phase program
  int NC, G, cl;               // CTAs, clusters, cluster size
  ActPlan f1, f2, f3, dg2, dg1, dgin;
  int wgv;                     // weight-gradient tile variant: 0 = 16 x 64 (4-way sample split), 1 = 32 x 128
  int ldp_tiles, red_tiles;    // logdet partial slots per stage, ActNorm partial slots per stage
  // dynamic shared memory layout (byte offsets)
  int oX, oW, oX2, oW2, oE, oP, oR, oPm, oIdx, oIdx2, oPD, smem_bytes;
};

}  // namespace dpi
