// flow_nar.cuh - sample-resident RealNVP pass for NARROW flows (ndim <= 32, hidden <= 160): plan + entry points shared
// with flow.cu. Shape class: the alpha-DPI flow over geometric-model parameters (DPIx_interferometry.py:250,
// RealNVP(18, 16, seqfrac 1/16): D = 18, H = 144, batch 2048), the 2-D toy posterior (DPI_2Dtoydist.py:52: D = 2,
// H = 128, batch 512) and posterior sampling (DPIx_interferometry.py:433, batch 8192). See flow_nar.cu.
#pragma once
#include "flow.cuh"

namespace dpi {
namespace nar {

constexpr int kThreadsN = 256;
constexpr int kClusterN = 8;        // CTAs per cluster (first level of the batch-statistics reduction)
constexpr int kMaxD = 32, kMaxH = 160, kMaxSPB = 64, kMaxClusters = 20;

struct Plan {
  int ok = 0;
  int G = 0, ncl = 0, SPB = 0;      // CTAs (multiple of 8), clusters, samples per CTA (multiple of 4)
  int smem_fwd = 0, smem_bwd = 0;   // dynamic shared memory (bytes)
  int Da4 = 0, Hp = 0;              // Da rounded up to 4 (rows of W1^T), H padded to 32 (stride of the per-feature vectors)
  // stage block (float offsets inside one block): first weight tile, the vectors, ZeroFC block, statistics, indices
  int bW1 = 0, bV1 = 0, bV2 = 0, bSC = 0, bST = 0, bIdx = 0, bAct = 0, blockFloats = 0;
  // shared-memory layout (float offsets)
  int oSt = 0, oBlk0 = 0, oBlk1 = 0, oW2 = 0, oX0 = 0, oX1 = 0, oU1 = 0, oU2 = 0, oLd = 0, oPart = 0, oTot = 0;
  int oGP = 0, oAct0 = 0, oAct1 = 0, actRows = 0;      // backward only
  // workspace (float offsets from W): per-stage gradient rows kept for the parameter-gradient kernel, reduction slots,
  // ActNorm partial sums, the transposed weights of the forward pass
  long long oGPRE = 0, oGP2 = 0, oGP1 = 0, oSlots = 0, oActPart = 0, oWT = 0;
};

int query(int device);                                   // max co-resident clusters of 8 (0: unavailable)
void make_plan(Plan& p, int max_clusters, int D, int H, int B, int S, int has_bn);
void make_plan_eval(Plan& p, int avail, int D, int H, int B, int S);   // forward-only, eval mode, any batch
long long workspace_floats_eval(const Plan& p, int D, int H, int S);
void place_workspace_eval(Plan& p, long long base);
long long workspace_floats(const Plan& p, int D, int H, int B, int S);
void place_workspace(Plan& p, long long base, int D, int H, int B, int S);
int forward(const KArgs& a, const Plan& p, cudaStream_t st);
int backward(const KArgs& a, const Plan& p, cudaStream_t st);
int has_backward();
int param_grads(const KArgs& a, const Plan& p, cudaStream_t st);   // batched parameter gradients from stored rows

}  // namespace nar
}  // namespace dpi
