// flow_sr.cuh - sample-resident RealNVP.reverse pass (and its backward) for SMALL batches of a mid-sized flow: plan +
// entry points shared with flow.cu. Shape class: the reference's default image flow at its default batch
// (DPI_interferometry.py:127 RealNVP(npix^2 = 1024, n_flow = 16), :60 n_batch = 64; hidden width 128). See flow_sr.cu.
#pragma once
#include "flow.cuh"
#include "flow_nar.cuh"

namespace dpi {
namespace sr {

constexpr int kT = 512;            // threads per CTA
constexpr int kCL = 16;            // CTAs = one cluster (non-portable size): the whole batch lives in it
constexpr int kSP = 4;             // samples per CTA
constexpr int kMaxD = 1024, kMaxH = 128;
constexpr long long kMaxStageFloats = 512 * 1024;   // 2 MB of weights per coupling stage: above this the per-SM weight
                                                    // stream (every CTA reads every weight) loses to the feature-split kernel

struct Plan {
  int ok = 0;
  int smem_fwd = 0, smem_bwd = 0;
  // shared memory (float offsets)
  int oSt = 0, oX0 = 0, oX1 = 0, oXA = 0, oU1 = 0, oU2 = 0, oRed = 0, oPart = 0, oIdx = 0, oMisc = 0, oEsc = 0, oPB = 0;
  int oGP = 0, oAct0 = 0, oAct1 = 0, actRows = 0;         // backward only
  // workspace: k-major copies of the weights for the forward pass, per stage [W1^T Da x H | W2^T H x H | W3^T H x Dout]
  long long oWT = 0, wt_stage = 0;
  nar::Plan pg;                    // rows the chain stores for nar's batched parameter-gradient kernel (oGPRE, oGP2, oGP1, oActPart, G)
};

int query(int device);                                    // 1: one cluster of 16 CTAs with the kernels' shared memory can run
void make_plan(Plan& p, int avail, int D, int H, int B, int S);
long long workspace_floats(const Plan& p, int D, int H, int B, int S);
void place_workspace(Plan& p, long long base, int D, int H, int B, int S);
int forward(const KArgs& a, const Plan& p, cudaStream_t st);
int backward(const KArgs& a, const Plan& p, cudaStream_t st);

}  // namespace sr
}  // namespace dpi
