// flow_cl.cuh - the cluster-resident tcgen05 RealNVP pass (flow_cl.cu): plan + entry points shared with flow.cu.
//
// Shape class: the reference's primary configuration, RealNVP(npix^2 = 1024, n_flow, seqfrac 4) -> H = 128, with a
// batch of at most 64 samples (DPI_interferometry.py:127,190). One thread-block cluster of 16 CTAs runs the whole
// dependent chain of a pass; see flow_cl.cu for the design.
#pragma once
#include "flow.cuh"

namespace dpi {
namespace fcl {

constexpr int kCL = 16;                 // CTAs of the cluster
constexpr int kThreadsCl = 544;
constexpr int kHc = 128;                // hidden width == UMMA M
constexpr int kNS = 64;                 // sample columns of every accumulator (batch zero-padded to 64)
constexpr int kKS1 = 32;                // a-half features per CTA (K slice of net.0)
constexpr int kHS = kHc / kCL;          // hidden features per CTA after a reduce-scatter (8)
constexpr int kJB = 32;                 // coupling columns per CTA; 2 * kJB rows of net.6.fc.weight
constexpr int kDc = 2 * kCL * kKS1;     // 1024

// packed weight tiles of one (stage, CTA): [W1 slice 128 x 32 | W2 slice 128 x 8 | W3 slice 64 x 128], each hi then lo,
// canonical K-major no-swizzle UMMA layout (core matrix 8 rows x 16 B, LBO 128 B, SBO = 32 * K-extent B)
constexpr int kW1Floats = 2 * kHc * kKS1;            // 8192
constexpr int kW2Floats = 2 * kHc * kHS;             // 2048
constexpr int kW3Floats = 2 * (2 * kJB) * kHc;       // 16384
constexpr int kPackFloats = kW1Floats + kW2Floats + kW3Floats;   // 26624 per (stage, CTA), forward and backward alike
constexpr int kPartFloats = 2 * kCL * kHc * kNS;     // two reduce-scatter partial areas [16 sources][128][64]
constexpr int kAgFloats = kCL * 2 * kNS * kHS;       // all-gather staging: per CTA one [64 x 8] operand tile, hi | lo
constexpr int kScratchFloats = kPartFloats + kAgFloats + kCL * kNS + 4096;   // + logdet rows + ActNorm partial sums

struct Plan {
  int ok = 0;
  long long oPackF = 0, oPackB = 0, oScratch = 0;    // float offsets into the workspace
};

// 1 when the device can co-schedule one 16-CTA cluster of each kernel (checked once per handle)
int query(int device);
bool shape_ok(int D, int H, int B);
int forward(const KArgs& a, const Plan& p, int S, cudaStream_t st);
int backward(const KArgs& a, const Plan& p, int S, cudaStream_t st);
int has_backward();

}  // namespace fcl
}  // namespace dpi
