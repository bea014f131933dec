// flow_host.cuh - host-side state of a flow handle, shared by the planner/launcher of the tiled phase
// program (flow.cu) and of the small-batch cluster kernel (flow_small.cu).
#pragma once
#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include "flow.cuh"
#include "flow_small.cuh"
#include "flow_tc.cuh"
#include "flow_cl.cuh"
#include "flow_nar.cuh"
#include "flow_sr.cuh"

namespace dpi {
void* flow_kernel_ptr();
void* flow_strip_kernel_ptr();

struct Program {
  Phase* d_phases = nullptr;
  int n_phases = 0;
  int max_tiles = 0;
  std::vector<int> phase_tiles;
  std::vector<int> ops;
};

struct BatchPlan {
  int B = 0, ldb = 0, tiles_m = 0, ldp_tiles = 0, red_tiles = 0;
  long long oZT, oYT, oL1T, oL2T, oN1T, oN2T, oOT, oST, oLDP, oGYT, oGPRET, oGNT, oGP1T, oGP2T, oPART, oRED, total_floats;
  // [dir][train] forward programs, and the backward program
  Program fwd[2][2];
  Program bwd[2];   // [1]: also produces the gradient wrt the flow input
  // small-batch cluster kernel (batch <= 64): decomposition + shared-memory layout, or small.ok == 0
  SmallPlan small;       // backward pass (and the shared facts: ok, workspace rows)
  SmallPlan small_fwd;   // forward pass (may use a different cluster size)
  // tensor-core path (flow_tc.cu): GEMM plans, the strip-op programs that run between the GEMMs, pack / partial areas
  int tc_ok = 0, tc_ldp_tiles = 0, tc_red_tiles = 0;
  ftc::Gemm g_f1, g_f2, g_f3, g_dg2, g_wg3, g_dg1, g_wg2, g_dgin, g_wg1;
  Program tc_fwd[2][2];
  Program tc_bwd[2];
  long long oTCA = 0, oTCB = 0, oTCP = 0;
  long long oTCA2 = 0, oTCB2 = 0, oTCP2 = 0;   // second set for the weight-gradient GEMMs on the side stream
  // weights of all stages packed once per pass (forward / transposed orientation) and the packed A operands that the
  // epilogues emit for the next GEMM of the chain
  int tc_emit = 0;
  ftc::WeightPack wp_fwd{}, wp_bwd{};
  long long oWPF = 0, oWPB = 0, oEA2 = 0, oEA3 = 0, oED1 = 0, oEDIN = 0;
  // cluster-resident tcgen05 pass (flow_cl.cu): reference-default shape, batch <= 64
  fcl::Plan cl;
  // sample-resident pass for narrow flows (flow_nar.cu): ndim <= 32, hidden <= 160, batch > 64
  nar::Plan nar;
  nar::Plan nar_eval;    // forward-only plan for eval-mode passes of batches the training plan cannot take (sampling)
  // sample-resident pass for batch <= 64 of a mid-sized flow (flow_sr.cu): weights streamed from L2, one cluster of 16
  sr::Plan sr;
};

}  // namespace dpi

struct dpi_flow_s {
  int device = 0, D = 0, Da = 0, Db = 0, Dout = 0, H = 0, n_flow = 0, S = 0, has_bn = 1;
  int n_sm = 0, persistent = 0, ctas_per_sm = 1, coop_max_per_sm = 1, use_tc = 1;
  int small_mode = 1;                         // 0: never use the small-batch kernel
  int nar_clusters = 0;                       // co-resident clusters of 8 of the narrow-flow kernels (0: unavailable)
  int sr_avail = 0;                           // one 16-CTA cluster of the flow_sr kernels can be scheduled
  int cl_avail = 0;                           // one 16-CTA cluster of the flow_cl kernels can be scheduled
  cudaStream_t side = nullptr;                // weight-gradient GEMMs of the tensor-core path run here, beside the dgrad chain
  std::vector<cudaEvent_t> ev;                // [4 S + 2] fork / join events (main <-> side)
  int tc_side = 1;
  int tc_emit = 0;                            // epilogues emit the next GEMM's packed operand, weights packed once per pass
  long long* d_woff = nullptr;                // [S * 3] parameter offsets of W1, W2, W3 by stage-table index
  int tc_min_batch = 96;                      // batches >= this run the coupling GEMMs on tcgen05 (0: never)
  int small_clusters = 0, small_cl = 1;       // CTAs without clusters; small_cl > 1: cluster launches are available
  int small_G[3] = {0, 0, 0};                 // co-resident clusters of 8, 4, 2 CTAs
  int small_ksplit_min = 512;                 // GEMMs with K >= this are split over the ranks of a cluster
  long long n_params = 0, n_bn = 0;
  std::vector<dpi::StageTab> stages;          // reverse execution order
  std::vector<long long> act_off;             // [n_flow][4]
  std::vector<long long> cpl_off;             // [n_flow][2][P_COUNT]
  std::vector<long long> bn_off;              // [n_flow][2][4]
  std::vector<int32_t> h_gmaps;               // [2][S][D] + trailing [D]
  dpi::StageTab* d_stages = nullptr;
  int* d_gmaps = nullptr;
  int* d_pmaps = nullptr;                     // flow_cl push maps [S][D] (nullptr unless the shape qualifies)
  std::mutex mu;
  std::map<int, std::unique_ptr<dpi::BatchPlan>> plans;
  unsigned long long* d_tbuf = nullptr;       // debug timing stamps (2 per phase)
  int tbuf_phases = 0;
  std::vector<int> last_ops;                  // op codes of the last launched program (debug)
};

namespace dpi {
// flow_small.cu
int small_query(dpi_flow_s* h);                                        // fills small_clusters / small_cl
void small_plan(const dpi_flow_s* h, BatchPlan& bp);                   // fills bp.small / small_fwd (ok = 0 if unsupported)
int small_launch(dpi_flow_s* h, const BatchPlan& bp, KArgs& a, int bwd, cudaStream_t stream, int lo = 0, int hi = -1);
}  // namespace dpi
