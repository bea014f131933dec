// flow.cuh - tables shared by the flow planner (host) and the phase-program kernel (device).
//
// The RealNVP flow (generative_model/realnvpfc_model.py:562-603) is executed as a *phase
// program*: a host-planned list of phases, each a set of independent tiles, with a grid-wide
// dependency between consecutive phases. The same device code runs either as one persistent
// cooperative kernel (grid barrier between phases) or as one launch per phase.
#pragma once
#include "common.cuh"
#include "tile_cp.cuh"

namespace dpi {

constexpr int kThreads = kNT;
constexpr int kBM = kTM;
constexpr int kBK = kTK;
constexpr int kCtrlWords = 4096;  // barrier + error flag + split-K tile counters (uint32)

// indices into StageTab::p - order == registration order of one AffineCoupling
enum { P_W1 = 0, P_B1, P_G1, P_BE1, P_W2, P_B2, P_G2, P_BE2, P_SCALE, P_W3, P_B3, P_COUNT };

// One coupling stage = (AffineCoupling, ActNorm) pair. Index 0 is the stage RealNVP.reverse
// runs first: blocks[n_flow-1].coupling2/actnorm2 (:599-600, :459-461).
struct StageTab {
  long long p[P_COUNT];  // float offsets into the flat parameter / gradient buffers
  long long loc, lsi;    // ActNorm.loc / .log_scale_inv
  long long r[4];        // running_mean1, running_var1, running_mean2, running_var2 (flat bn buffer)
  long long pad_;        // keeps sizeof a multiple of 16 (staged into shared memory with int4 copies)
};

enum Op : int {
  OP_FWD_HID = 1,    // Linear + LeakyReLU (+ BatchNorm batch statistics)  net.0-2 / net.3-5
  OP_FWD_OUT,        // ZeroFC + tanh/exp affine coupling + ActNorm + logdet partials
  OP_BN_STATS,       // batch statistics + normalisation when the batch spans several row tiles
  OP_LOGDET_FINAL,   // deterministic reduction of logdet partials
  OP_TRANSPOSE,      // sample-major <-> feature-major at the flow's boundary (+ trailing permutation)
  OP_B0,             // backward of ActNorm + coupling elementwise part
  OP_DGRAD_HID,      // dgrad through W3 / W2 (+ fused BatchNorm/LeakyReLU backward)
  OP_WGRAD,          // weight gradients
  OP_DGRAD_IN,       // dgrad through W1, written through the fused permutation
  OP_BNBWD,          // BatchNorm/LeakyReLU backward for large batches
  OP_BWD_FINAL,      // ActNorm gradients from per-tile partials
  OP_TAIL            // tensor-core path: ZeroFC bias/scale + coupling + ActNorm + logdet partials on the raw GEMM output
};

struct SubOp {
  int op, which;              // which: layer selector (1,2 hidden; 3 output)
  int ntiles;                 // tiles_m * tiles_n * splits (GEMM ops) or strips
  int tiles_m, tiles_n, splits, bn, kchunk;
  int ctr_off;                // first split-K counter (ctrl words)
  long long part_off;         // split-K partial area (floats, relative to oPART)
};
struct Phase {
  int stage;                  // execution position e in [0, S)
  int nsub;
  int pad_[2];
  SubOp sub[2];
};

static_assert(sizeof(Phase) % 16 == 0 && sizeof(StageTab) % 16 == 0, "tables are staged with 16-byte copies");

struct KArgs {
  const StageTab* st;
  const int* gmaps;           // [2][S][D] fused gather maps, by direction and execution position
  const Phase* phases;
  int D, Da, Db, Dout, H, S, has_bn, B, tiles_m, dir, train;
  int ldp_tiles, red_tiles;
  // workspace layout (float offsets from W)
  long long oZT, oYT, oL1T, oL2T, oN1T, oN2T, oOT, oST, oLDP, oGYT, oGPRET, oGNT, oGP1T, oGP2T, oPART, oRED;
  // bases
  const float* P; float* G; float* R; float* W;
  const float* in; float* out; float* logdet;
  const float* g_out; const float* g_ld; float* g_in;
  unsigned* ctrl;
  unsigned long long* tbuf;   // debug: per-phase globaltimer stamps of CTA 0 (start, work done), or nullptr
  int phase_begin, phase_end, persistent;
  int use_tc;                 // 1: tcgen05/TMEM tile core, 0: mma.sync tile core
  int fin_lo, fin_hi;         // OP_BWD_FINAL: stages whose ActNorm gradients are finalised ([0, S) unless the backward
                              // is run in stage ranges)
  int bfin_phase;             // flow_small.cu backward: the phase index that finalises (the last one of the launch)
  const int* pmaps;           // flow_cl.cu: [S][D] consumer slot (cta * 64 + row slot) of every output row in the next stage
  int ldb;                    // row stride (floats) of the feature-major workspace matrices: B, or round_up4(B)
                              // for the small-batch kernel (every row 16-byte aligned)
};

}  // namespace dpi
