// flow.cu - host side of the RealNVP phase program: parameter layout, fused permutation maps,
// phase planning, launches. Mirrors generative_model/realnvpfc_model.py RealNVP :562-603.
#include <algorithm>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include "flow_host.cuh"

namespace dpi {
namespace {

inline long long pad32(long long n) { return (n + 31) / 32 * 32; }

// Flat parameter layout: nn.Module.parameters() order of the reference RealNVP, every tensor
// start padded to 32 floats (128 B) so rows can be fetched with vector / bulk copies.
void compute_layout(int ndim, int n_flow, int H, int has_bn, std::vector<long long>& act_off,
                    std::vector<long long>& cpl_off, std::vector<long long>& bn_off, long long& n_params,
                    long long& n_bn) {
  const int Da = ndim - ndim / 2, Dout = 2 * (ndim / 2);
  long long o = 0, ob = 0;
  auto take = [&](long long n) { long long r = o; o += pad32(n); return r; };
  act_off.assign((size_t)n_flow * 4, 0);
  cpl_off.assign((size_t)n_flow * 2 * P_COUNT, -1);
  bn_off.assign((size_t)n_flow * 2 * 4, -1);
  for (int i = 0; i < n_flow; ++i) {
    for (int k = 0; k < 4; ++k) act_off[(size_t)i * 4 + k] = take(1);
    for (int c = 0; c < 2; ++c) {
      long long* p = &cpl_off[((size_t)i * 2 + c) * P_COUNT];
      p[P_W1] = take((long long)H * Da); p[P_B1] = take(H);
      if (has_bn) { p[P_G1] = take(H); p[P_BE1] = take(H); }
      p[P_W2] = take((long long)H * H); p[P_B2] = take(H);
      if (has_bn) { p[P_G2] = take(H); p[P_BE2] = take(H); }
      p[P_SCALE] = take(Dout); p[P_W3] = take((long long)Dout * H); p[P_B3] = take(Dout);
      if (has_bn)
        for (int k = 0; k < 4; ++k) { bn_off[((size_t)i * 2 + c) * 4 + k] = ob; ob += pad32(H); }
    }
  }
  n_params = o;
  n_bn = std::max(32LL, ob);
}

struct GemmPlan { int bn, tiles_n, splits, kchunk; };

// Pick the column tile and split-K factor so a phase offers about `target` tiles.
GemmPlan plan_gemm(int tiles_m, int N, int K, int target, bool allow_split, int min_bn = 16) {
  GemmPlan g{};
  const int cands[3] = {64, 32, 16};
  g.bn = min_bn;
  for (int c = 0; c < 3; ++c) {
    const int bn = cands[c];
    if (bn < min_bn) break;
    g.bn = bn;
    if ((long long)tiles_m * ((N + bn - 1) / bn) >= target) break;
  }
  g.tiles_n = (N + g.bn - 1) / g.bn;
  g.splits = 1;
  g.kchunk = (K + kBK - 1) / kBK * kBK;
  if (allow_split) {
    const int mn = tiles_m * g.tiles_n;
    int want = std::max(1, target / std::max(1, mn));
    static const int min_kt = []() { const char* e = getenv("DPI_B200_MIN_KTILES"); return e ? std::max(1, atoi(e)) : 4; }();
    const int max_splits = std::max(1, K / (min_kt * kBK));  // at least min_kt k-tiles per split
    want = std::min(std::min(want, max_splits), 32);
    if (want > 1) {
      int chunk = (K + want - 1) / want;
      chunk = (chunk + kBK - 1) / kBK * kBK;
      g.kchunk = chunk;
      g.splits = (K + chunk - 1) / chunk;
    }
  }
  return g;
}

SubOp gemm_subop(int op, int which, int tiles_m, const GemmPlan& g) {
  SubOp s{};
  s.op = op; s.which = which; s.tiles_m = tiles_m; s.tiles_n = g.tiles_n; s.splits = g.splits;
  s.bn = g.bn; s.kchunk = g.kchunk; s.ntiles = tiles_m * g.tiles_n * g.splits;
  return s;
}
SubOp strip_subop(int op, int which, int ntiles) {
  SubOp s{};
  s.op = op; s.which = which; s.ntiles = ntiles; s.tiles_m = 1; s.tiles_n = ntiles; s.splits = 1; s.bn = 0;
  return s;
}

struct Builder {
  std::vector<Phase> phases;
  long long max_part = 0;
  void add(int stage, SubOp a) {
    Phase p{};
    p.stage = stage; p.nsub = 1; p.sub[0] = a;
    finish(p);
  }
  void add(int stage, SubOp a, SubOp b) {
    Phase p{};
    p.stage = stage; p.nsub = 2; p.sub[0] = a; p.sub[1] = b;
    finish(p);
  }
  void finish(Phase& p) {
    long long part = 0;
    int ctr = 16;
    for (int i = 0; i < p.nsub; ++i) {
      SubOp& s = p.sub[i];
      s.ctr_off = ctr;
      s.part_off = part;
      if (s.splits > 1) {
        ctr += s.tiles_m * s.tiles_n;
        part += (long long)s.tiles_m * s.tiles_n * s.splits * kBM * s.bn;
      }
    }
    max_part = std::max(max_part, part);
    phases.push_back(p);
  }
};

int upload(Builder& b, Program& prog) {
  prog.n_phases = (int)b.phases.size();
  prog.max_tiles = 1;
  prog.phase_tiles.clear();
  for (auto& p : b.phases) {
    int t = p.sub[0].ntiles + (p.nsub > 1 ? p.sub[1].ntiles : 0);
    prog.phase_tiles.push_back(t);
    prog.ops.push_back(p.sub[0].op * 100 + p.sub[0].which * 10 + (p.nsub > 1 ? 1 : 0));
    prog.max_tiles = std::max(prog.max_tiles, t);
    for (int i = 0; i < p.nsub; ++i)
      if (p.sub[i].splits > 1 && p.sub[i].ctr_off + p.sub[i].tiles_m * p.sub[i].tiles_n > kCtrlWords) {
        set_error("flow planner: too many split-K tiles");
        return DPI_ERR_UNSUPPORTED;
      }
  }
  DPI_CUDA(cudaMalloc(&prog.d_phases, sizeof(Phase) * b.phases.size()));
  DPI_CUDA(cudaMemcpy(prog.d_phases, b.phases.data(), sizeof(Phase) * b.phases.size(), cudaMemcpyHostToDevice));
  return DPI_OK;
}

int build_plan(dpi_flow_s* h, int B, BatchPlan& bp) {
  const int D = h->D, Da = h->Da, Db = h->Db, Dout = h->Dout, H = h->H, S = h->S;
  const int target = h->n_sm * std::max(1, h->ctas_per_sm);
  bp.B = B;
  bp.tiles_m = (B + kBM - 1) / kBM;
  const int tm = bp.tiles_m;
  const bool sep_stats = tm > 1;
  const int tr_tiles = ((B + 31) / 32) * ((D + 31) / 32);

  // output layer: BJ = 32 (BN 64) when that still fills the grid, else BJ = 16
  int bj = 32;
  if ((long long)tm * ((Db + 31) / 32) < target) bj = 16;
  const int tiles_j = (Db + bj - 1) / bj;
  bp.ldp_tiles = tiles_j;
  bp.red_tiles = (Db + 7) / 8;

  long long max_part = 0;
  for (int dir = 0; dir < 2; ++dir)
    for (int train = 0; train < 2; ++train) {
      Builder b;
      b.add(0, strip_subop(OP_TRANSPOSE, 0, tr_tiles));
      for (int e = 0; e < S; ++e) {
        b.add(e, gemm_subop(OP_FWD_HID, 1, tm, plan_gemm(tm, H, Da, target, true)));
        if (h->has_bn && train && sep_stats) b.add(e, strip_subop(OP_BN_STATS, 1, (H + 7) / 8));
        b.add(e, gemm_subop(OP_FWD_HID, 2, tm, plan_gemm(tm, H, H, target, true)));
        if (h->has_bn && train && sep_stats) b.add(e, strip_subop(OP_BN_STATS, 2, (H + 7) / 8));
        SubOp o{};
        o.op = OP_FWD_OUT; o.which = 3; o.tiles_m = tm; o.tiles_n = tiles_j; o.splits = 1; o.bn = 2 * bj;
        o.kchunk = H; o.ntiles = tm * tiles_j;
        b.add(e, o);
      }
      b.add(S - 1, strip_subop(OP_LOGDET_FINAL, 0, (B + 31) / 32), strip_subop(OP_TRANSPOSE, dir == 0 ? 2 : 3, tr_tiles));
      max_part = std::max(max_part, b.max_part);
      int rc = upload(b, bp.fwd[dir][train]);
      if (rc) return rc;
    }
  for (int want_gin = 0; want_gin < 2; ++want_gin) {
    Builder b;
    b.add(S - 1, strip_subop(OP_TRANSPOSE, 1, tr_tiles));
    for (int e = S - 1; e >= 0; --e) {
      b.add(e, strip_subop(OP_B0, 0, bp.red_tiles));
      b.add(e, gemm_subop(OP_DGRAD_HID, 2, tm, plan_gemm(tm, H, Dout, target, true)),
            gemm_subop(OP_WGRAD, 3, (H + kBM - 1) / kBM, plan_gemm((H + kBM - 1) / kBM, Dout, B, target, false)));
      if (sep_stats) b.add(e, strip_subop(OP_BNBWD, 2, (H + 7) / 8));
      b.add(e, gemm_subop(OP_DGRAD_HID, 1, tm, plan_gemm(tm, H, H, target, true)),
            gemm_subop(OP_WGRAD, 2, (H + kBM - 1) / kBM, plan_gemm((H + kBM - 1) / kBM, H, B, target, false)));
      if (sep_stats) b.add(e, strip_subop(OP_BNBWD, 1, (H + 7) / 8));
      b.add(e, gemm_subop(OP_DGRAD_IN, 1, tm, plan_gemm(tm, Da, H, target, false)),
            gemm_subop(OP_WGRAD, 1, (Da + kBM - 1) / kBM, plan_gemm((Da + kBM - 1) / kBM, H, B, target, false)));
    }
    if (want_gin) b.add(0, strip_subop(OP_BWD_FINAL, 0, 1), strip_subop(OP_TRANSPOSE, 4, tr_tiles));
    else b.add(0, strip_subop(OP_BWD_FINAL, 0, 1));
    max_part = std::max(max_part, b.max_part);
    int rc = upload(b, bp.bwd[want_gin]);
    if (rc) return rc;
  }

  small_plan(h, bp);
  int ldp_slots = std::max(bp.ldp_tiles, bp.small.ok ? bp.small.ldp_tiles : 0);

  // ---- tensor-core path: strip-op programs between the tcgen05 GEMMs (flow_tc.cu)
  bp.tc_ok = (!bp.small.ok && h->tc_min_batch > 0 && B >= h->tc_min_batch) ? 1 : 0;
  if (bp.tc_ok) {
    const int cols = ftc::kTailCols;                       // coupling columns per logdet partial
    const int chunks = (Db + cols - 1) / cols, mblocks = (B + kThreads - 1) / kThreads;
    bp.tc_ldp_tiles = chunks;
    ldp_slots = std::max(ldp_slots, chunks);
    for (int dir = 0; dir < 2; ++dir)
      for (int train = 0; train < 2; ++train) {
        Builder b;
        b.add(0, strip_subop(OP_TRANSPOSE, 0, tr_tiles));
        for (int e = 0; e < S; ++e) {
          if (h->has_bn && train) {
            b.add(e, strip_subop(OP_BN_STATS, 1, (H + 7) / 8));
            b.add(e, strip_subop(OP_BN_STATS, 2, (H + 7) / 8));
          }
          SubOp tl = strip_subop(OP_TAIL, 0, chunks * mblocks);
          tl.tiles_n = chunks; tl.tiles_m = mblocks; tl.kchunk = cols;
          b.add(e, tl);
        }
        b.add(S - 1, strip_subop(OP_LOGDET_FINAL, 0, (B + 31) / 32), strip_subop(OP_TRANSPOSE, dir == 0 ? 2 : 3, tr_tiles));
        int rc = upload(b, bp.tc_fwd[dir][train]);
        if (rc) return rc;
      }
    for (int want_gin = 0; want_gin < 2; ++want_gin) {
      Builder b;
      b.add(S - 1, strip_subop(OP_TRANSPOSE, 1, tr_tiles));
      // narrow flows (few coupling columns): one CTA per column instead of one warp per column
      const bool wide_b0 = Db <= 128;
      bp.tc_red_tiles = wide_b0 ? Db : bp.red_tiles;
      for (int e = S - 1; e >= 0; --e) {
        b.add(e, strip_subop(OP_B0, wide_b0 ? 1 : 0, bp.tc_red_tiles));
        b.add(e, strip_subop(OP_BNBWD, 2, (H + 7) / 8));
        b.add(e, strip_subop(OP_BNBWD, 1, (H + 7) / 8));
      }
      if (want_gin) b.add(0, strip_subop(OP_BWD_FINAL, 0, 1), strip_subop(OP_TRANSPOSE, 4, tr_tiles));
      else b.add(0, strip_subop(OP_BWD_FINAL, 0, 1));
      int rc = upload(b, bp.tc_bwd[want_gin]);
      if (rc) return rc;
    }
    if (h->tc_side && !h->side) {
      DPI_CUDA(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
      h->ev.resize((size_t)4 * S + 2);
      for (auto& ev : h->ev) DPI_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    }
    const int sm = h->n_sm;
    ftc::configure();
    bp.g_f1 = ftc::make_gemm(B, H, Da, sm);
    bp.g_f2 = ftc::make_gemm(B, H, H, sm);
    bp.g_f3 = ftc::make_gemm(B, Dout, H, sm);
    bp.g_dg2 = ftc::make_gemm(B, H, Dout, sm);
    bp.g_wg3 = ftc::make_gemm(Dout, H, B, sm);
    bp.g_dg1 = ftc::make_gemm(B, H, H, sm);
    bp.g_wg2 = ftc::make_gemm(H, H, B, sm);
    bp.g_dgin = ftc::make_gemm(B, Da, H, sm);
    bp.g_wg1 = ftc::make_gemm(H, Da, B, sm);
    bp.tc_emit = h->tc_emit;
    auto fill = [&](ftc::WeightPack& w, const ftc::Gemm* g[3], const int rows[3], const int K[3], const int ld[3], int kc) {
      w.S = S; w.src_off = h->d_woff;
      long long o = 0;
      for (int l = 0; l < 3; ++l) {
        w.rows[l] = rows[l]; w.K[l] = K[l]; w.ld[l] = ld[l]; w.k_contig[l] = kc;
        w.TN[l] = g[l]->TN; w.KT[l] = g[l]->KT;
        w.tiles[l] = g[l]->n_tiles * g[l]->k_tiles;
        w.dst_off[l] = o;
        o += (long long)ftc::b_pack_floats(*g[l]);
      }
      w.stage_floats = o;
    };
    {   // forward orientation: B[n][k] = W[n][k]
      const ftc::Gemm* g[3] = {&bp.g_f1, &bp.g_f2, &bp.g_f3};
      const int rows[3] = {H, H, Dout}, K[3] = {Da, H, H}, ld[3] = {Da, H, H};
      fill(bp.wp_fwd, g, rows, K, ld, 1);
    }
    {   // backward (dgrad) orientation: B[n][k] = W[k][n]; layer order W1 (dgrad 0), W2 (dgrad 3), W3 (dgrad 6)
      const ftc::Gemm* g[3] = {&bp.g_dgin, &bp.g_dg1, &bp.g_dg2};
      const int rows[3] = {Da, H, H}, K[3] = {H, H, Dout}, ld[3] = {Da, H, H};
      fill(bp.wp_bwd, g, rows, K, ld, 0);
    }
  }

  const long long LB = bp.small.ok ? (B + 3) / 4 * 4 : B;   // workspace row stride
  bp.ldb = (int)LB;
  long long o = 0;
  auto take = [&](long long n) { long long r = o; o += pad32(n); return r; };
  bp.oZT = take((long long)D * LB);
  bp.oYT = take((long long)S * D * LB);
  bp.oL1T = take((long long)S * H * LB);
  bp.oL2T = take((long long)S * H * LB);
  bp.oN1T = take((long long)S * H * LB);
  bp.oN2T = take((long long)S * H * LB);
  bp.oOT = take((long long)S * Dout * LB);
  bp.oST = take((long long)S * 4 * H);
  bp.oLDP = take((long long)S * ldp_slots * LB);
  bp.oGYT = take(2LL * D * LB);
  bp.oGPRET = take((long long)Dout * LB);
  bp.oGNT = take((long long)H * LB);
  bp.oGP1T = take((long long)H * LB);
  bp.oGP2T = take((long long)H * LB);
  bp.oPART = take(std::max(32LL, max_part));
  bp.oRED = take((long long)S * std::max(bp.red_tiles, bp.tc_red_tiles) * 2);
  if (!bp.small.ok && h->nar_clusters > 0) {
    nar::make_plan(bp.nar, h->nar_clusters, D, H, B, S, h->has_bn);
    if (bp.nar.ok) {
      const long long base = take(nar::workspace_floats(bp.nar, D, H, B, S));
      nar::place_workspace(bp.nar, base, D, H, B, S);
    } else {
      nar::make_plan_eval(bp.nar_eval, h->nar_clusters, D, H, B, S);
      if (bp.nar_eval.ok) nar::place_workspace_eval(bp.nar_eval, take(nar::workspace_floats_eval(bp.nar_eval, D, H, S)));
    }
  }
  if (bp.small.ok && h->sr_avail) {
    sr::make_plan(bp.sr, h->sr_avail, D, H, B, S);
    if (bp.sr.ok) {
      const long long base = take(sr::workspace_floats(bp.sr, D, H, B, S));
      sr::place_workspace(bp.sr, base, D, H, B, S);
    }
  }
  bp.cl.ok = (h->cl_avail && bp.small.ok && fcl::shape_ok(D, H, B)) ? 1 : 0;
  if (bp.cl.ok) {
    o = (o + 255) / 256 * 256;                  // packed tiles are fetched with cp.async.bulk
    bp.cl.oPackF = take((long long)S * fcl::kCL * fcl::kPackFloats);
    bp.cl.oPackB = take((long long)S * fcl::kCL * fcl::kPackFloats);
    bp.cl.oScratch = take(fcl::kScratchFloats);
  }
  if (bp.tc_ok) {
    const ftc::Gemm* gs[9] = {&bp.g_f1, &bp.g_f2, &bp.g_f3, &bp.g_dg2, &bp.g_wg3, &bp.g_dg1, &bp.g_wg2, &bp.g_dgin, &bp.g_wg1};
    size_t ma = 0, mb = 0, mp = 0;
    for (const ftc::Gemm* g : gs) {
      ma = std::max(ma, ftc::a_pack_floats(*g));
      mb = std::max(mb, ftc::b_pack_floats(*g));
      mp = std::max(mp, ftc::part_floats(*g));
    }
    o = (o + 255) / 256 * 256;                // packed tiles are fetched with cp.async.bulk: keep them 1 KB aligned
    bp.oTCA = take((long long)ma);
    bp.oTCB = take((long long)mb);
    bp.oTCP = take((long long)mp);
    ma = mb = mp = 0;
    const ftc::Gemm* ws[3] = {&bp.g_wg3, &bp.g_wg2, &bp.g_wg1};
    for (const ftc::Gemm* g : ws) {
      ma = std::max(ma, ftc::a_pack_floats(*g));
      mb = std::max(mb, ftc::b_pack_floats(*g));
      mp = std::max(mp, ftc::part_floats(*g));
    }
    o = (o + 255) / 256 * 256;
    bp.oTCA2 = take((long long)ma);
    bp.oTCB2 = take((long long)mb);
    bp.oTCP2 = take((long long)mp);
    if (bp.tc_emit) {
      o = (o + 255) / 256 * 256;
      bp.oWPF = take((long long)S * bp.wp_fwd.stage_floats);
      bp.oWPB = take((long long)S * bp.wp_bwd.stage_floats);
      bp.oEA2 = take((long long)ftc::a_pack_floats(bp.g_f2));
      bp.oEA3 = take((long long)ftc::a_pack_floats(bp.g_f3));
      bp.oED1 = take((long long)ftc::a_pack_floats(bp.g_dg1));
      bp.oEDIN = take((long long)ftc::a_pack_floats(bp.g_dgin));
    }
  }
  bp.total_floats = o;
  return DPI_OK;
}

int get_plan(dpi_flow_s* h, int B, BatchPlan** out) {
  std::lock_guard<std::mutex> lk(h->mu);
  auto it = h->plans.find(B);
  if (it == h->plans.end()) {
    std::unique_ptr<BatchPlan> bp(new BatchPlan());
    int rc = build_plan(h, B, *bp);
    if (rc) return rc;
    it = h->plans.emplace(B, std::move(bp)).first;
  }
  *out = it->second.get();
  return DPI_OK;
}

void fill_args(dpi_flow_s* h, const BatchPlan& bp, KArgs& a, void* workspace) {
  a.st = h->d_stages; a.gmaps = h->d_gmaps; a.pmaps = h->d_pmaps;
  a.D = h->D; a.Da = h->Da; a.Db = h->Db; a.Dout = h->Dout; a.H = h->H; a.S = h->S; a.has_bn = h->has_bn;
  a.B = bp.B; a.ldb = bp.ldb; a.tiles_m = bp.tiles_m; a.ldp_tiles = bp.ldp_tiles; a.red_tiles = bp.red_tiles;
  a.oZT = bp.oZT; a.oYT = bp.oYT; a.oL1T = bp.oL1T; a.oL2T = bp.oL2T; a.oN1T = bp.oN1T; a.oN2T = bp.oN2T;
  a.oOT = bp.oOT; a.oST = bp.oST; a.oLDP = bp.oLDP; a.oGYT = bp.oGYT; a.oGPRET = bp.oGPRET; a.oGNT = bp.oGNT;
  a.oGP1T = bp.oGP1T; a.oGP2T = bp.oGP2T; a.oPART = bp.oPART; a.oRED = bp.oRED;
  a.fin_lo = 0; a.fin_hi = h->S;
  a.ctrl = (unsigned*)workspace;
  a.W = (float*)((char*)workspace + kCtrlWords * sizeof(unsigned));
}

int launch_program(dpi_flow_s* h, const Program& prog, KArgs& a, cudaStream_t stream) {
  a.phases = prog.d_phases;
  a.use_tc = h->use_tc;
  a.tbuf = nullptr;
  if (h->d_tbuf && prog.n_phases <= 4096) {
    a.tbuf = h->d_tbuf;
    h->tbuf_phases = prog.n_phases;
    h->last_ops = prog.ops;
  }
  DPI_CUDA(cudaMemsetAsync(a.ctrl, 0, kCtrlWords * sizeof(unsigned), stream));
  if (h->persistent) {
    int grid = std::min(prog.max_tiles, h->n_sm * std::min(h->ctas_per_sm, h->coop_max_per_sm));
    grid = std::max(grid, 1);
    a.phase_begin = 0; a.phase_end = prog.n_phases; a.persistent = 1;
    void* params[] = {&a};
    DPI_CUDA(cudaLaunchCooperativeKernel(flow_kernel_ptr(), dim3(grid), dim3(kThreads), params, kTcSmemBytes, stream));
    count_launch();
  } else {
    a.persistent = 0;
    void* params[] = {&a};
    for (int p = 0; p < prog.n_phases; ++p) {
      a.phase_begin = p; a.phase_end = p + 1;
      DPI_CUDA(cudaLaunchKernel(flow_kernel_ptr(), dim3(prog.phase_tiles[p]), dim3(kThreads), params, kTcSmemBytes, stream));
      count_launch();
    }
  }
  return DPI_OK;
}

int launch_phase(const Program& prog, KArgs& a, int p, cudaStream_t stream) {
  a.phases = prog.d_phases;
  a.persistent = 0; a.phase_begin = p; a.phase_end = p + 1;
  void* params[] = {&a};
  DPI_CUDA(cudaLaunchKernel(flow_strip_kernel_ptr(), dim3(prog.phase_tiles[p]), dim3(kThreads), params, 0, stream));
  count_launch();
  return DPI_OK;
}

#define DPI_TRY(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

struct TcBufs {
  float *apack, *bpack, *part;
};

int tc_gemm(const ftc::Gemm& g, const ftc::Pack& pa, const ftc::Pack& pb, const ftc::Epi& ep, const TcBufs& tb,
            cudaStream_t st) {
  DPI_TRY(ftc::pack2(g, pa, tb.apack, pb, tb.bpack, st));
  return ftc::run(g, tb.apack, tb.bpack, tb.part, ep, st);
}

// RealNVP.reverse / .forward (:594-603 / :583-592) with the three Linear layers of every coupling net on tcgen05.
int tc_forward(dpi_flow_s* h, const BatchPlan& bp, KArgs& a, cudaStream_t st) {
  const Program& prog = bp.tc_fwd[a.dir][a.train];
  const int D = h->D, Da = h->Da, Dout = h->Dout, H = h->H, S = h->S, B = bp.B;
  const long long LB = bp.ldb;
  a.ldp_tiles = bp.tc_ldp_tiles;
  a.use_tc = 0; a.tbuf = nullptr;
  DPI_TRY(ftc::init());
  DPI_CUDA(cudaMemsetAsync(a.ctrl, 0, kCtrlWords * sizeof(unsigned), st));
  float* W = a.W;
  const TcBufs tb{W + bp.oTCA, W + bp.oTCB, W + bp.oTCP};
  const bool bn_phase = h->has_bn && a.train;
  // fused operand path: the weights of all stages are packed once (one launch), and each hidden-layer epilogue emits
  // the packed A operand of the next GEMM, so only the first GEMM of a stage still has a pack launch
  const bool emit = bp.tc_emit && (!bn_phase || B <= ftc::kEpiMaxM);
  const ftc::Pack none{nullptr, 0, 1, nullptr, 0, 1};
  float* WPF = W + bp.oWPF;
  float* EA2 = W + bp.oEA2;
  float* EA3 = W + bp.oEA3;
  if (emit) {
    DPI_TRY(ftc::pack_weights(bp.wp_fwd, a.P, WPF, st));
    // the emitted tiles are only written where (sample, feature) exist: their padding must read as zero
    DPI_CUDA(cudaMemsetAsync(EA2, 0, ftc::a_pack_floats(bp.g_f2) * sizeof(float), st));
    DPI_CUDA(cudaMemsetAsync(EA3, 0, ftc::a_pack_floats(bp.g_f3) * sizeof(float), st));
  }
  int pi = 0;
  DPI_TRY(launch_phase(prog, a, pi++, st));
  for (int e = 0; e < S; ++e) {
    const int sidx = a.dir ? S - 1 - e : e;
    const float* wpf = WPF + (long long)sidx * bp.wp_fwd.stage_floats;
    const StageTab& sg = h->stages[a.dir ? S - 1 - e : e];
    const int* gmap = h->d_gmaps + ((size_t)a.dir * S + e) * D;
    const float* yprev = e == 0 ? W + bp.oZT : W + bp.oYT + (long long)(e - 1) * D * LB;
    float* L1 = W + bp.oL1T + (long long)e * H * LB;
    float* L2 = W + bp.oL2T + (long long)e * H * LB;
    float* N1 = W + bp.oN1T + (long long)e * H * LB;
    float* N2 = W + bp.oN2T + (long long)e * H * LB;
    float* O = W + bp.oOT + (long long)e * Dout * LB;
    for (int which = 1; which <= 2; ++which) {
      ftc::Epi ep{};
      ep.mode = ftc::EPI_T_HID; ep.out = which == 1 ? L1 : L2; ep.ldo = LB; ep.bias = a.P + sg.p[which == 1 ? P_B1 : P_B2];
      const bool fused_bn = bn_phase && B <= ftc::kEpiMaxM;   // batch statistics inside the GEMM epilogue
      if (!bn_phase || fused_bn) {
        ep.out2 = which == 1 ? N1 : N2;
        ep.bn_mode = !h->has_bn ? 0 : (a.train ? 1 : 2);
        if (h->has_bn) {
          ep.gamma = a.P + sg.p[which == 1 ? P_G1 : P_G2]; ep.beta = a.P + sg.p[which == 1 ? P_BE1 : P_BE2];
          ep.rmean = a.R + sg.r[(which - 1) * 2]; ep.rvar = a.R + sg.r[(which - 1) * 2 + 1];
          ep.mean = W + bp.oST + ((long long)e * 4 + (which - 1) * 2) * H;
          ep.rstd = ep.mean + H;
        }
      }
      if (emit) {
        const ftc::Gemm& nx = which == 1 ? bp.g_f2 : bp.g_f3;
        ep.tp = which == 1 ? EA2 : EA3; ep.tp_R = nx.MH * ftc::kM; ep.tp_KT = nx.KT; ep.tp_ktiles = nx.k_tiles;
        if (which == 1) {
          DPI_TRY(ftc::pack2(bp.g_f1, ftc::Pack{yprev, LB, 0, gmap, B, Da}, tb.apack, none, tb.bpack, st));
          DPI_TRY(ftc::run(bp.g_f1, tb.apack, wpf + bp.wp_fwd.dst_off[0], tb.part, ep, st));
        } else {
          DPI_TRY(ftc::run(bp.g_f2, EA2, wpf + bp.wp_fwd.dst_off[1], tb.part, ep, st));
        }
      } else if (which == 1)
        DPI_TRY(tc_gemm(bp.g_f1, ftc::Pack{yprev, LB, 0, gmap, B, Da}, ftc::Pack{a.P + sg.p[P_W1], Da, 1, nullptr, H, Da}, ep, tb, st));
      else
        DPI_TRY(tc_gemm(bp.g_f2, ftc::Pack{N1, LB, 0, nullptr, B, H}, ftc::Pack{a.P + sg.p[P_W2], H, 1, nullptr, H, H}, ep, tb, st));
      if (bn_phase) {
        if (!fused_bn) DPI_TRY(launch_phase(prog, a, pi, st));
        ++pi;
      }
    }
    // net.6 (ZeroFC) with the coupling tail as its epilogue
    ftc::Epi ep{};
    ep.mode = ftc::EPI_T_TAIL; ep.out = O; ep.ldo = LB;
    ep.yprev = yprev; ep.y = W + bp.oYT + (long long)e * D * LB; ep.gmap = gmap;
    ep.scale = a.P + sg.p[P_SCALE]; ep.b3 = a.P + sg.p[P_B3];
    ep.dir = a.dir; ep.Da = Da; ep.Db = h->Db;
    if (a.dir == 0) { ep.post_mode = 0; ep.post_lsi = a.P + sg.lsi; ep.post_loc = a.P + sg.loc; }
    else if (e + 1 < S) { const StageTab& nx = h->stages[S - 2 - e]; ep.post_mode = 1; ep.post_lsi = a.P + nx.lsi; ep.post_loc = a.P + nx.loc; }
    else ep.post_mode = 2;
    ep.ldp = W + bp.oLDP + (long long)e * bp.tc_ldp_tiles * LB;
    if (emit) DPI_TRY(ftc::run(bp.g_f3, EA3, wpf + bp.wp_fwd.dst_off[2], tb.part, ep, st));
    else DPI_TRY(tc_gemm(bp.g_f3, ftc::Pack{N2, LB, 0, nullptr, B, H}, ftc::Pack{a.P + sg.p[P_W3], H, 1, nullptr, Dout, H}, ep, tb, st));
    ++pi;   // the strip-op tail of the program is not needed
  }
  return launch_phase(prog, a, pi++, st);
}

// Backward of RealNVP.reverse: data gradients and weight gradients of the coupling nets on tcgen05.
// Stages [lo, hi) in backward order (hi - 1 first); the full pass is [0, S). A partial range that does not end at
// stage 0 finalises only its own stages' ActNorm gradients and leaves the input gradient to the last range.
int tc_backward(dpi_flow_s* h, const BatchPlan& bp, KArgs& a, cudaStream_t st, int lo, int hi) {
  const Program& prog = bp.tc_bwd[(a.g_in && lo == 0) ? 1 : 0];
  const int D = h->D, Da = h->Da, Dout = h->Dout, H = h->H, S = h->S, B = bp.B;
  const long long LB = bp.ldb;
  a.ldp_tiles = bp.tc_ldp_tiles;
  a.red_tiles = bp.tc_red_tiles;
  a.use_tc = 0; a.tbuf = nullptr;
  DPI_TRY(ftc::init());
  DPI_CUDA(cudaMemsetAsync(a.ctrl, 0, kCtrlWords * sizeof(unsigned), st));
  float* W = a.W;
  const TcBufs tb{W + bp.oTCA, W + bp.oTCB, W + bp.oTCP};
  float* GPRE = W + bp.oGPRET;
  float* GN = W + bp.oGNT;
  float* GP1 = W + bp.oGP1T;
  float* GP2 = W + bp.oGP2T;
  // The weight-gradient GEMMs only consume what the dgrad chain has already produced: they run on a side stream
  // (own pack / partial buffers) beside the chain. Events: A = g_pre ready, B = g_p2 ready, C = g_p1 ready (main ->
  // side), D = side has packed the last operand of the stage (side -> main, before the chain overwrites them).
  const bool side = h->tc_side && h->side;
  cudaStream_t sw = side ? h->side : st;
  const TcBufs tw = side ? TcBufs{W + bp.oTCA2, W + bp.oTCB2, W + bp.oTCP2} : tb;
  auto fork = [&](cudaEvent_t ev) -> int {
    if (!side) return DPI_OK;
    DPI_CUDA(cudaEventRecord(ev, st));
    DPI_CUDA(cudaStreamWaitEvent(sw, ev, 0));
    return DPI_OK;
  };
  auto join = [&](cudaEvent_t ev) -> int {
    if (!side) return DPI_OK;
    DPI_CUDA(cudaEventRecord(ev, sw));
    DPI_CUDA(cudaStreamWaitEvent(st, ev, 0));
    return DPI_OK;
  };
  const bool emit = bp.tc_emit && B <= ftc::kEpiMaxM;
  const ftc::Pack none{nullptr, 0, 1, nullptr, 0, 1};
  float* WPB = W + bp.oWPB;
  float* ED1 = W + bp.oED1;
  float* EDIN = W + bp.oEDIN;
  a.fin_lo = lo; a.fin_hi = hi;
  if (emit && hi == S) {
    DPI_TRY(ftc::pack_weights(bp.wp_bwd, a.P, WPB, st));
    DPI_CUDA(cudaMemsetAsync(ED1, 0, ftc::a_pack_floats(bp.g_dg1) * sizeof(float), st));
    DPI_CUDA(cudaMemsetAsync(EDIN, 0, ftc::a_pack_floats(bp.g_dgin) * sizeof(float), st));
  }
  // program layout: [transpose] then per stage (S - 1 first) [B0, BatchNorm backward 2, BatchNorm backward 1], [final]
  int pi = 1 + 3 * (S - hi);
  if (hi == S) DPI_TRY(launch_phase(prog, a, 0, st));
  for (int e = hi - 1; e >= lo; --e) {
    const StageTab& sg = h->stages[e];
    const float* wpb = WPB + (long long)e * bp.wp_bwd.stage_floats;
    const int* gmap = h->d_gmaps + (size_t)e * D;
    const float* yprev = e == 0 ? W + bp.oZT : W + bp.oYT + (long long)(e - 1) * D * LB;
    const float* N1 = W + bp.oN1T + (long long)e * H * LB;
    const float* N2 = W + bp.oN2T + (long long)e * H * LB;
    const float* gy = W + bp.oGYT + (long long)(e & 1) * D * LB;
    float* gyprev = W + bp.oGYT + (long long)((e - 1) & 1) * D * LB;
    cudaEvent_t* ev = side ? &h->ev[(size_t)4 * e] : nullptr;
    // B0 overwrites g_pre (and the chain then g_p2, g_p1): wait until the side stream has packed the previous stage's
    if (side && e < hi - 1) DPI_CUDA(cudaStreamWaitEvent(st, h->ev[(size_t)4 * (e + 1) + 3], 0));
    DPI_TRY(launch_phase(prog, a, pi++, st));   // B0
    const bool fused_bn = B <= ftc::kEpiMaxM;   // BatchNorm + LeakyReLU backward inside the dgrad epilogue
    auto hid_epi = [&](int which) {
      ftc::Epi ep{};
      ep.ldo = LB;
      if (!fused_bn) { ep.mode = ftc::EPI_T_RAW; ep.out = GN; return ep; }
      ep.mode = ftc::EPI_T_BNBWD; ep.out = which == 2 ? GP2 : GP1; ep.bn_mode = h->has_bn;
      ep.lsaved = W + (which == 1 ? bp.oL1T : bp.oL2T) + (long long)e * H * LB;
      if (h->has_bn) {
        ep.mean = W + bp.oST + ((long long)e * 4 + (which - 1) * 2) * H;
        ep.rstd = ep.mean + H;
        ep.gamma = a.P + sg.p[which == 1 ? P_G1 : P_G2];
        ep.g_gamma = a.G + sg.p[which == 1 ? P_G1 : P_G2];
        ep.g_beta = a.G + sg.p[which == 1 ? P_BE1 : P_BE2];
      }
      ep.g_bias = a.G + sg.p[which == 1 ? P_B1 : P_B2];
      if (emit) {   // g_p2 feeds dgrad through net.3, g_p1 feeds dgrad through net.0
        const ftc::Gemm& nx = which == 2 ? bp.g_dg1 : bp.g_dgin;
        ep.tp = which == 2 ? ED1 : EDIN; ep.tp_R = nx.MH * ftc::kM; ep.tp_KT = nx.KT; ep.tp_ktiles = nx.k_tiles;
      }
      return ep;
    };
    ftc::Epi wg{};
    wg.mode = ftc::EPI_ROWMAJOR;
    // through net.6.fc: g_n2[m][n] = sum_k g_pre[k][m] W3[k][n];  gW3[i][j] = sum_m g_pre[i][m] n2[j][m]
    if (side) DPI_TRY(fork(ev[0]));
    wg.out = a.G + sg.p[P_W3]; wg.ldo = H;
    DPI_TRY(tc_gemm(bp.g_wg3, ftc::Pack{GPRE, LB, 1, nullptr, Dout, B}, ftc::Pack{N2, LB, 1, nullptr, H, B}, wg, tw, sw));
    if (emit) {
      DPI_TRY(ftc::pack2(bp.g_dg2, ftc::Pack{GPRE, LB, 0, nullptr, B, Dout}, tb.apack, none, tb.bpack, st));
      DPI_TRY(ftc::run(bp.g_dg2, tb.apack, wpb + bp.wp_bwd.dst_off[2], tb.part, hid_epi(2), st));
    } else
      DPI_TRY(tc_gemm(bp.g_dg2, ftc::Pack{GPRE, LB, 0, nullptr, B, Dout}, ftc::Pack{a.P + sg.p[P_W3], H, 0, nullptr, H, Dout}, hid_epi(2), tb, st));
    if (!fused_bn) DPI_TRY(launch_phase(prog, a, pi, st));   // BatchNorm + LeakyReLU backward of hidden layer 2 -> GP2
    ++pi;
    // through net.3
    if (side) DPI_TRY(fork(ev[1]));
    wg.out = a.G + sg.p[P_W2]; wg.ldo = H;
    DPI_TRY(tc_gemm(bp.g_wg2, ftc::Pack{GP2, LB, 1, nullptr, H, B}, ftc::Pack{N1, LB, 1, nullptr, H, B}, wg, tw, sw));
    if (emit) DPI_TRY(ftc::run(bp.g_dg1, ED1, wpb + bp.wp_bwd.dst_off[1], tb.part, hid_epi(1), st));
    else DPI_TRY(tc_gemm(bp.g_dg1, ftc::Pack{GP2, LB, 0, nullptr, B, H}, ftc::Pack{a.P + sg.p[P_W2], H, 0, nullptr, H, H}, hid_epi(1), tb, st));
    if (!fused_bn) DPI_TRY(launch_phase(prog, a, pi, st));   // hidden layer 1 -> GP1
    ++pi;
    // through net.0, written through the fused permutation, plus the ActNorm pass-through of the a-half
    if (side) DPI_TRY(fork(ev[2]));
    wg.out = a.G + sg.p[P_W1]; wg.ldo = Da;
    {
      const ftc::Pack pa{GP1, LB, 1, nullptr, H, B}, pb{yprev, LB, 1, gmap, Da, B};
      DPI_TRY(ftc::pack2(bp.g_wg1, pa, tw.apack, pb, tw.bpack, sw));
      if (side) DPI_CUDA(cudaEventRecord(ev[3], sw));   // every operand of this stage's weight gradients is packed
      DPI_TRY(ftc::run(bp.g_wg1, tw.apack, tw.bpack, tw.part, wg, sw));
    }
    ftc::Epi sc{};
    sc.mode = ftc::EPI_T_SCATTER; sc.out = gyprev; sc.ldo = LB; sc.rowmap = gmap; sc.add = gy; sc.lsi = a.P + sg.lsi;
    if (emit) DPI_TRY(ftc::run(bp.g_dgin, EDIN, wpb + bp.wp_bwd.dst_off[0], tb.part, sc, st));
    else DPI_TRY(tc_gemm(bp.g_dgin, ftc::Pack{GP1, LB, 0, nullptr, B, H}, ftc::Pack{a.P + sg.p[P_W1], Da, 0, nullptr, Da, H}, sc, tb, st));
  }
  if (side && hi > lo) DPI_TRY(join(h->ev[(size_t)4 * S]));
  return launch_phase(prog, a, 1 + 3 * S, st);
}

}  // namespace
}  // namespace dpi

using namespace dpi;

extern "C" {

int dpi_flow_create(dpi_flow_t* out, int device, int ndim, int n_flow, int hidden, int batch_norm,
                    const int32_t* orders) {
  DPI_REQUIRE(out && orders, "dpi_flow_create: null argument");
  DPI_REQUIRE(ndim >= 2 && n_flow >= 1 && hidden >= 1, "dpi_flow_create: bad dims ndim=%d n_flow=%d hidden=%d", ndim, n_flow, hidden);
  DeviceGuard g(device);
  DPI_REQUIRE(g.ok, "dpi_flow_create: cannot select device %d", device);
  std::unique_ptr<dpi_flow_s> h(new dpi_flow_s());
  h->device = device; h->D = ndim; h->Db = ndim / 2; h->Da = ndim - ndim / 2; h->Dout = 2 * (ndim / 2);
  h->H = hidden; h->n_flow = n_flow; h->S = 2 * n_flow; h->has_bn = batch_norm ? 1 : 0;
  h->n_sm = sm_count(device);
  DPI_REQUIRE(h->n_sm > 0, "dpi_flow_create: no CUDA device %d", device);
  int nb = 0;
  DPI_CUDA(cudaFuncSetAttribute((const void*)flow_kernel_ptr(), cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
  DPI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*)flow_kernel_ptr(), kThreads, kTcSmemBytes));
  h->coop_max_per_sm = std::max(1, nb);
  const char* mode = getenv("DPI_B200_FLOW_MODE");
  h->persistent = (mode && mode[0] == 's') ? 0 : 1;
  // tile core of the large-batch path: mma.sync 3xTF32 (validated). "tc" selects the experimental tcgen05
  // core, whose MN-major operand descriptors still mismatch on hardware (tools/bench_tile.cu).
  const char* tile = getenv("DPI_B200_TILE");
  h->use_tc = (tile && tile[0] == 't') ? 1 : 0;
  small_query(h.get());
  {
    const char* ne = getenv("DPI_B200_NAR");                // 0: never use the sample-resident narrow-flow kernels
    h->nar_clusters = (ne && ne[0] == '0') ? 0 : nar::query(device);
  }
  {
    // The sample-resident small-batch pass (flow_sr.cu) is opt-in (DPI_B200_SR=1): parity-green, but measured at 664 us per
    // RealNVP(1024, 16) reverse pass / 835 us backward against 508 / 675 us for the feature-split kernel
    // (profiles/sr_r02.md: LDG.128 and FFMA issue add up instead of overlapping, 63 B/clk per SM in the micro-GEMM)
    const char* se = getenv("DPI_B200_SR");
    h->sr_avail = (se && se[0] == '1') ? sr::query(device) : 0;
  }
  {
    // The cluster-resident tcgen05 pass (flow_cl.cu) is opt-in (DPI_B200_CL=1): parity-green, but measured at 585 us
    // per RealNVP(1024, 16) reverse pass against 501 us for the fp32 small-batch kernel (profiles/cl_fwd_r02.md)
    const char* cle = getenv("DPI_B200_CL");
    h->cl_avail = (cle && cle[0] == '1') ? fcl::query(device) : 0;
  }
  const char* tcm = getenv("DPI_B200_FLOW_TC");          // minimum batch of the tcgen05 GEMM path; 0 disables it
  if (tcm) h->tc_min_batch = std::max(0, atoi(tcm));
  const char* tce = getenv("DPI_B200_TC_EMIT");            // 1: fused operand emission + once-per-pass weight packs
  if (tce) h->tc_emit = atoi(tce) ? 1 : 0;
  const char* tcs = getenv("DPI_B200_TC_SIDE");            // 0: weight-gradient GEMMs stay on the caller's stream
  if (tcs) h->tc_side = atoi(tcs) ? 1 : 0;
  const char* cps = getenv("DPI_B200_CTAS_PER_SM");
  h->ctas_per_sm = cps ? std::max(1, atoi(cps)) : 1;

  const int D = h->D, Da = h->Da, Dout = h->Dout, H = h->H;
  for (int i = 0; i < n_flow; ++i)
    for (int j = 0; j < ndim; ++j) DPI_REQUIRE(orders[(size_t)i * ndim + j] >= 0 && orders[(size_t)i * ndim + j] < ndim, "dpi_flow_create: orders[%d][%d] out of range", i, j);

  compute_layout(D, n_flow, H, h->has_bn, h->act_off, h->cpl_off, h->bn_off, h->n_params, h->n_bn);

  // ---- stage table in reverse execution order
  h->stages.resize(h->S);
  for (int s = 0; s < h->S; ++s) {
    const int i = n_flow - 1 - s / 2, c = (s % 2 == 0) ? 1 : 0;  // coupling2 first (:459-464)
    StageTab& st = h->stages[s];
    for (int k = 0; k < P_COUNT; ++k) st.p[k] = h->cpl_off[((size_t)i * 2 + c) * P_COUNT + k];
    st.loc = h->act_off[(size_t)i * 4 + (c ? 2 : 0)];
    st.lsi = h->act_off[(size_t)i * 4 + (c ? 3 : 1)];
    for (int k = 0; k < 4; ++k) st.r[k] = h->bn_off[((size_t)i * 2 + c) * 4 + k];
  }

  // ---- fused gather maps
  h->h_gmaps.assign((size_t)(2 * h->S + 1) * D, 0);
  std::vector<int32_t> inv(D);
  for (int e = 0; e < h->S; ++e) {
    int32_t* g0 = &h->h_gmaps[((size_t)0 * h->S + e) * D];
    const int i0 = n_flow - 1 - e / 2;
    if (e % 2 == 0) {
      const int32_t* ord = orders + (size_t)i0 * D;
      for (int j = 0; j < D; ++j) inv[ord[j]] = j;            // Order_inverse (:553-559)
      for (int j = 0; j < D; ++j) g0[j] = inv[D - 1 - j];      // x[:, inv][:, flip]
    } else {
      for (int j = 0; j < D; ++j) g0[j] = D - 1 - j;
    }
    int32_t* g1 = &h->h_gmaps[((size_t)1 * h->S + e) * D];
    const int i1 = e / 2;
    if (e % 2 == 0) {
      if (i1 == 0) for (int j = 0; j < D; ++j) g1[j] = j;
      else {
        const int32_t* ord = orders + (size_t)(i1 - 1) * D;
        for (int j = 0; j < D; ++j) g1[j] = D - 1 - ord[j];   // y[:, flip][:, order]
      }
    } else {
      for (int j = 0; j < D; ++j) g1[j] = D - 1 - j;
    }
  }
  {
    int32_t* gf = &h->h_gmaps[(size_t)2 * h->S * D];
    const int32_t* ord = orders + (size_t)(n_flow - 1) * D;
    for (int j = 0; j < D; ++j) gf[j] = D - 1 - ord[j];
  }
  {
    std::vector<long long> woff((size_t)h->S * 3);
    for (int st = 0; st < h->S; ++st) {
      woff[(size_t)st * 3 + 0] = h->stages[st].p[P_W1];
      woff[(size_t)st * 3 + 1] = h->stages[st].p[P_W2];
      woff[(size_t)st * 3 + 2] = h->stages[st].p[P_W3];
    }
    DPI_CUDA(cudaMalloc(&h->d_woff, sizeof(long long) * woff.size()));
    DPI_CUDA(cudaMemcpy(h->d_woff, woff.data(), sizeof(long long) * woff.size(), cudaMemcpyHostToDevice));
  }
  if (h->cl_avail && D == fcl::kDc && H == fcl::kHc) {
    // flow_cl.cu pushes every output row of stage e straight into the shared memory of the CTA that consumes it in
    // stage e + 1: CTA = (j' mod Da) / 32, slot = (a-half ? 0 : 32) + j' mod 32 for the consumer's feature j'
    std::vector<int32_t> pm((size_t)h->S * D, 0);
    for (int e = 0; e + 1 < h->S; ++e) {
      const int32_t* g = &h->h_gmaps[(size_t)(e + 1) * D];      // reverse direction, stage e + 1
      for (int jp = 0; jp < D; ++jp) {
        const int half = jp >= Da ? 1 : 0, jl = jp - half * Da;
        pm[(size_t)e * D + g[jp]] = (jl / fcl::kKS1) * 64 + half * 32 + jl % fcl::kKS1;
      }
    }
    DPI_CUDA(cudaMalloc(&h->d_pmaps, sizeof(int32_t) * pm.size()));
    DPI_CUDA(cudaMemcpy(h->d_pmaps, pm.data(), sizeof(int32_t) * pm.size(), cudaMemcpyHostToDevice));
  }
  DPI_CUDA(cudaMalloc(&h->d_stages, sizeof(StageTab) * h->S));
  DPI_CUDA(cudaMemcpy(h->d_stages, h->stages.data(), sizeof(StageTab) * h->S, cudaMemcpyHostToDevice));
  DPI_CUDA(cudaMalloc(&h->d_gmaps, sizeof(int32_t) * h->h_gmaps.size()));
  DPI_CUDA(cudaMemcpy(h->d_gmaps, h->h_gmaps.data(), sizeof(int32_t) * h->h_gmaps.size(), cudaMemcpyHostToDevice));
  handle_register(h.get(), 1);
  *out = h.release();
  return DPI_OK;
}

int dpi_flow_layout(int ndim, int n_flow, int hidden, int batch_norm, int64_t* param_offsets, int64_t* bn_offsets,
                    int64_t* n_params, int64_t* n_bn) {
  DPI_REQUIRE(ndim >= 2 && n_flow >= 1 && hidden >= 1, "dpi_flow_layout: bad dims");
  std::vector<long long> act, cpl, bn;
  long long np = 0, nb = 0;
  compute_layout(ndim, n_flow, hidden, batch_norm ? 1 : 0, act, cpl, bn, np, nb);
  if (param_offsets)
    for (int i = 0; i < n_flow; ++i) {
      for (int k = 0; k < 4; ++k) param_offsets[(size_t)i * (4 + 2 * P_COUNT) + k] = act[(size_t)i * 4 + k];
      for (int k = 0; k < 2 * P_COUNT; ++k) param_offsets[(size_t)i * (4 + 2 * P_COUNT) + 4 + k] = cpl[(size_t)i * 2 * P_COUNT + k];
    }
  if (bn_offsets)
    for (size_t i = 0; i < bn.size(); ++i) bn_offsets[i] = bn[i];
  if (n_params) *n_params = np;
  if (n_bn) *n_bn = nb;
  return DPI_OK;
}

int dpi_flow_destroy(dpi_flow_t h) {
  if (!h) return DPI_OK;
  DPI_REQUIRE(handle_unregister(h, 1), "dpi_flow_destroy: unknown or already destroyed handle %p", (void*)h);
  DeviceGuard g(h->device);
  for (auto& kv : h->plans) {
    for (int d = 0; d < 2; ++d)
      for (int t = 0; t < 2; ++t) cudaFree(kv.second->fwd[d][t].d_phases);
    cudaFree(kv.second->bwd[0].d_phases);
    cudaFree(kv.second->bwd[1].d_phases);
    for (int d = 0; d < 2; ++d)
      for (int t = 0; t < 2; ++t) cudaFree(kv.second->tc_fwd[d][t].d_phases);
    cudaFree(kv.second->tc_bwd[0].d_phases);
    cudaFree(kv.second->tc_bwd[1].d_phases);
  }
  for (auto& ev : h->ev) cudaEventDestroy(ev);
  if (h->side) cudaStreamDestroy(h->side);
  cudaFree(h->d_woff);
  cudaFree(h->d_stages);
  cudaFree(h->d_gmaps);
  cudaFree(h->d_pmaps);
  cudaFree(h->d_tbuf);
  delete h;
  return DPI_OK;
}

int64_t dpi_flow_param_count(dpi_flow_t h) { return handle_live(h, 1) ? h->n_params : -1; }
int64_t dpi_flow_bn_count(dpi_flow_t h) { return handle_live(h, 1) ? h->n_bn : -1; }

int64_t dpi_flow_param_offset(dpi_flow_t h, int flow, int tensor_id) {
  if (!handle_live(h, 1) || flow < 0 || flow >= h->n_flow || tensor_id < 0 || tensor_id >= 4 + 2 * P_COUNT) return -1;
  if (tensor_id < 4) return h->act_off[(size_t)flow * 4 + tensor_id];
  const int c = (tensor_id - 4) / P_COUNT, k = (tensor_id - 4) % P_COUNT;
  return h->cpl_off[((size_t)flow * 2 + c) * P_COUNT + k];
}

int64_t dpi_flow_bn_offset(dpi_flow_t h, int flow, int coupling, int stat_id) {
  if (!handle_live(h, 1) || flow < 0 || flow >= h->n_flow || coupling < 0 || coupling > 1 || stat_id < 0 || stat_id > 3) return -1;
  return h->bn_off[((size_t)flow * 2 + coupling) * 4 + stat_id];
}

int dpi_flow_gather_map(dpi_flow_t h, int direction, int stage, int32_t* out) {
  DPI_REQUIRE(handle_live(h, 1) && out, "dpi_flow_gather_map: null argument or dead handle");
  DPI_REQUIRE(direction >= 0 && direction <= 2 && stage >= 0 && (direction == 2 ? stage == 0 : stage < h->S),
              "dpi_flow_gather_map: bad direction/stage");
  const int32_t* src = &h->h_gmaps[((size_t)direction * h->S + stage) * h->D];
  std::copy(src, src + h->D, out);
  return DPI_OK;
}

size_t dpi_flow_workspace_bytes(dpi_flow_t h, int batch) {
  if (!handle_live(h, 1) || batch < 1) return 0;
  DeviceGuard g(h->device);
  BatchPlan* bp = nullptr;
  if (get_plan(h, batch, &bp)) return 0;
  return kCtrlWords * sizeof(unsigned) + (size_t)bp->total_floats * sizeof(float);
}

int64_t dpi_flow_ws_offset(dpi_flow_t h, int batch, int what, int stage) {
  if (!handle_live(h, 1) || batch < 1 || stage < 0 || stage >= h->S) return -1;
  DeviceGuard g(h->device);
  BatchPlan* bp = nullptr;
  if (get_plan(h, batch, &bp)) return -1;
  long long o;
  const long long B = bp->ldb;   // row stride of the saved matrices
  if (what == 5) return bp->ldb;
  switch (what) {
    case 0: o = bp->oYT + stage * B * h->D; break;
    case 1: o = bp->oL1T + stage * B * h->H; break;
    case 2: o = bp->oL2T + stage * B * h->H; break;
    case 3: o = bp->oOT + stage * B * h->Dout; break;
    case 4: o = bp->oST + (long long)stage * 4 * h->H; break;
    default: return -1;
  }
  return (int64_t)(kCtrlWords * sizeof(unsigned)) + o * (int64_t)sizeof(float);
}

int dpi_flow_tc_plan(int M, int N, int K, int n_sm, int32_t* out) {
  DPI_REQUIRE(out && M >= 1 && N >= 1 && K >= 1 && n_sm >= 1, "dpi_flow_tc_plan: bad argument");
  ftc::configure();
  const ftc::Gemm g = ftc::make_gemm(M, N, K, n_sm);
  out[0] = g.TN; out[1] = g.KT; out[2] = g.MH; out[3] = g.m_tiles; out[4] = g.n_tiles; out[5] = g.k_tiles;
  out[6] = g.splits; out[7] = g.kt_per_split; out[8] = g.stages;
  out[9] = g.stages * (2 * g.MH * ftc::kM * g.KT + 2 * g.TN * g.KT) * 4 + 1024;
  return DPI_OK;
}

int dpi_flow_debug_timing(dpi_flow_t h, int enable) {
  DPI_REQUIRE(handle_live(h, 1), "dpi_flow_debug_timing: null or dead handle");
  DeviceGuard g(h->device);
  std::lock_guard<std::mutex> lk(h->mu);
  if (enable && !h->d_tbuf) {
    DPI_CUDA(cudaMalloc(&h->d_tbuf, sizeof(unsigned long long) * 2 * 4096));
    DPI_CUDA(cudaMemset(h->d_tbuf, 0, sizeof(unsigned long long) * 2 * 4096));
  } else if (!enable && h->d_tbuf) {
    cudaDeviceSynchronize();
    cudaFree(h->d_tbuf);
    h->d_tbuf = nullptr;
  }
  return DPI_OK;
}

int dpi_flow_debug_read(dpi_flow_t h, uint64_t* stamps, int32_t* ops, int max_phases) {
  DPI_REQUIRE(handle_live(h, 1) && stamps && ops, "dpi_flow_debug_read: null argument or dead handle");
  DPI_REQUIRE(h->d_tbuf, "dpi_flow_debug_read: timing not enabled");
  DeviceGuard g(h->device);
  DPI_CUDA(cudaDeviceSynchronize());
  const int n = std::min(h->tbuf_phases, max_phases);
  // a caller that offers the whole buffer also gets the fine-grained stamps stored behind the phase stamps
  DPI_CUDA(cudaMemcpy(stamps, h->d_tbuf, sizeof(unsigned long long) * 2 * (max_phases >= 4096 ? 4096 : n),
                      cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; ++i) ops[i] = h->last_ops[i];
  return n;
}

int dpi_flow_set_mode(dpi_flow_t h, int persistent, int ctas_per_sm) {
  DPI_REQUIRE(handle_live(h, 1), "dpi_flow_set_mode: null or dead handle");
  std::lock_guard<std::mutex> lk(h->mu);
  h->persistent = persistent ? 1 : 0;
  if (ctas_per_sm > 0 && ctas_per_sm != h->ctas_per_sm) {
    h->ctas_per_sm = ctas_per_sm;
    // tile plans depend on the grid size: drop cached plans (their device tables stay alive for
    // any launch already enqueued; freed at destroy time would leak, so free them now after a sync)
    cudaDeviceSynchronize();
    for (auto& kv : h->plans) {
      for (int d = 0; d < 2; ++d)
        for (int t = 0; t < 2; ++t) cudaFree(kv.second->fwd[d][t].d_phases);
      cudaFree(kv.second->bwd[0].d_phases);
      cudaFree(kv.second->bwd[1].d_phases);
    }
    h->plans.clear();
  }
  return DPI_OK;
}

int dpi_flow_run(dpi_flow_t h, int direction, int train, int batch, const float* params, float* bn_running,
                 const float* in, float* out, float* logdet, void* workspace, size_t workspace_bytes,
                 dpi_stream_t stream) {
  DPI_REQUIRE(handle_live(h, 1), "dpi_flow_run: null, unknown or destroyed flow handle");
  DPI_REQUIRE(params && in && out && logdet && workspace, "dpi_flow_run: null argument");
  DPI_REQUIRE(direction == 0 || direction == 1, "dpi_flow_run: direction must be 0 (reverse) or 1 (forward)");
  DPI_REQUIRE(batch >= 1, "dpi_flow_run: batch must be >= 1");
  DPI_REQUIRE(!(h->has_bn && !bn_running), "dpi_flow_run: bn_running required for a BatchNorm flow");
  DPI_REQUIRE(!(h->has_bn && train && batch < 2),
              "dpi_flow_run: Expected more than 1 value per channel when training (batch=%d)", batch);
  DeviceGuard g(h->device);
  BatchPlan* bp = nullptr;
  int rc = get_plan(h, batch, &bp);
  if (rc) return rc;
  const size_t need = kCtrlWords * sizeof(unsigned) + (size_t)bp->total_floats * sizeof(float);
  if (workspace_bytes < need) {
    set_error("dpi_flow_run: workspace too small (%zu < %zu)", workspace_bytes, need);
    return DPI_ERR_WORKSPACE;
  }
  KArgs a{};
  fill_args(h, *bp, a, workspace);
  a.dir = direction; a.train = train ? 1 : 0;
  a.P = params; a.G = nullptr; a.R = bn_running; a.in = in; a.out = out; a.logdet = logdet;
  a.g_out = nullptr; a.g_ld = nullptr; a.g_in = nullptr;
  if (bp->cl.ok && direction == 0) {
    a.tbuf = h->d_tbuf;
    return fcl::forward(a, bp->cl, h->S, (cudaStream_t)stream);
  }
  if (bp->sr.ok && direction == 0) {
    a.tbuf = h->d_tbuf;
    return sr::forward(a, bp->sr, (cudaStream_t)stream);
  }
  if (bp->small.ok) return small_launch(h, *bp, a, 0, (cudaStream_t)stream);
  if (bp->nar.ok && direction == 0) {
    a.tbuf = h->d_tbuf;
    return nar::forward(a, bp->nar, (cudaStream_t)stream);
  }
  if (bp->nar_eval.ok && direction == 0 && !a.train) return nar::forward(a, bp->nar_eval, (cudaStream_t)stream);
  if (bp->tc_ok) return tc_forward(h, *bp, a, (cudaStream_t)stream);
  return launch_program(h, bp->fwd[direction][a.train], a, (cudaStream_t)stream);
}

int dpi_flow_backward(dpi_flow_t h, int batch, const float* params, float* grads, const float* in,
                      const float* g_out, const float* g_logdet, float* g_in, void* workspace,
                      size_t workspace_bytes, dpi_stream_t stream) {
  DPI_REQUIRE(handle_live(h, 1), "dpi_flow_backward: null, unknown or destroyed flow handle");
  DPI_REQUIRE(params && grads && in && g_out && g_logdet && workspace, "dpi_flow_backward: null argument");
  DeviceGuard g(h->device);
  BatchPlan* bp = nullptr;
  int rc = get_plan(h, batch, &bp);
  if (rc) return rc;
  const size_t need = kCtrlWords * sizeof(unsigned) + (size_t)bp->total_floats * sizeof(float);
  if (workspace_bytes < need) {
    set_error("dpi_flow_backward: workspace too small (%zu < %zu)", workspace_bytes, need);
    return DPI_ERR_WORKSPACE;
  }
  KArgs a{};
  fill_args(h, *bp, a, workspace);
  a.dir = 0; a.train = 1;
  a.P = params; a.G = grads; a.R = nullptr; a.in = in; a.out = nullptr; a.logdet = nullptr;
  a.g_out = g_out; a.g_ld = g_logdet; a.g_in = g_in;
  if (bp->cl.ok && fcl::has_backward()) {
    a.tbuf = h->d_tbuf;
    return fcl::backward(a, bp->cl, h->S, (cudaStream_t)stream);
  }
  if (bp->sr.ok) {
    a.tbuf = h->d_tbuf;
    return sr::backward(a, bp->sr, (cudaStream_t)stream);
  }
  if (bp->small.ok) return small_launch(h, *bp, a, 1, (cudaStream_t)stream);
  if (bp->nar.ok && nar::has_backward()) return nar::backward(a, bp->nar, (cudaStream_t)stream);
  if (bp->tc_ok) return tc_backward(h, *bp, a, (cudaStream_t)stream, 0, h->S);
  return launch_program(h, bp->bwd[g_in ? 1 : 0], a, (cudaStream_t)stream);
}

int dpi_flow_backward_range(dpi_flow_t h, int batch, const float* params, float* grads, const float* in,
                            const float* g_out, const float* g_logdet, float* g_in, void* workspace,
                            size_t workspace_bytes, int stage_lo, int stage_hi, dpi_stream_t stream) {
  DPI_REQUIRE(handle_live(h, 1), "dpi_flow_backward_range: null, unknown or destroyed flow handle");
  DPI_REQUIRE(params && grads && in && g_out && g_logdet && workspace, "dpi_flow_backward_range: null argument");
  DPI_REQUIRE(0 <= stage_lo && stage_lo < stage_hi && stage_hi <= h->S, "dpi_flow_backward_range: bad stage range [%d, %d) of %d",
              stage_lo, stage_hi, h->S);
  if (stage_lo == 0 && stage_hi == h->S)
    return dpi_flow_backward(h, batch, params, grads, in, g_out, g_logdet, g_in, workspace, workspace_bytes, stream);
  DeviceGuard g(h->device);
  BatchPlan* bp = nullptr;
  int rc = get_plan(h, batch, &bp);
  if (rc) return rc;
  const bool small_path = bp->small.ok && !bp->sr.ok && !(bp->cl.ok && fcl::has_backward());
  if (!bp->tc_ok && !small_path) {
    set_error("dpi_flow_backward_range: partial stage ranges need the tcgen05 path (batch >= %d) or the small-batch kernel; batch %d "
              "runs on a path with one launch per pass", h->tc_min_batch, batch);
    return DPI_ERR_UNSUPPORTED;
  }
  const size_t need = kCtrlWords * sizeof(unsigned) + (size_t)bp->total_floats * sizeof(float);
  if (workspace_bytes < need) {
    set_error("dpi_flow_backward_range: workspace too small (%zu < %zu)", workspace_bytes, need);
    return DPI_ERR_WORKSPACE;
  }
  KArgs a{};
  fill_args(h, *bp, a, workspace);
  a.dir = 0; a.train = 1;
  a.P = params; a.G = grads; a.R = nullptr; a.in = in; a.out = nullptr; a.logdet = nullptr;
  a.g_out = g_out; a.g_ld = g_logdet; a.g_in = g_in;
  if (small_path) return small_launch(h, *bp, a, 1, (cudaStream_t)stream, stage_lo, stage_hi);
  return tc_backward(h, *bp, a, (cudaStream_t)stream, stage_lo, stage_hi);
}

int dpi_flow_range_ok(dpi_flow_t h, int batch) {
  if (!handle_live(h, 1) || batch < 1) return 0;
  DeviceGuard g(h->device);
  BatchPlan* bp = nullptr;
  if (get_plan(h, batch, &bp)) return 0;
  if (bp->tc_ok) return 1;
  return (bp->small.ok && !bp->sr.ok && !(bp->cl.ok && fcl::has_backward())) ? 2 : 0;
}

int dpi_flow_cl_active(dpi_flow_t h, int batch) {
  if (!handle_live(h, 1) || batch < 1) return 0;
  DeviceGuard g(h->device);
  BatchPlan* bp = nullptr;
  if (get_plan(h, batch, &bp)) return 0;
  return bp->cl.ok ? (1 | (fcl::has_backward() ? 2 : 0)) : 0;
}

int dpi_flow_sr_active(dpi_flow_t h, int batch) {
  if (!handle_live(h, 1) || batch < 1) return 0;
  DeviceGuard g(h->device);
  BatchPlan* bp = nullptr;
  if (get_plan(h, batch, &bp)) return 0;
  return (bp->sr.ok && !bp->cl.ok) ? 1 : 0;
}

int dpi_flow_nar_active(dpi_flow_t h, int batch) {
  if (!handle_live(h, 1) || batch < 1) return 0;
  DeviceGuard g(h->device);
  BatchPlan* bp = nullptr;
  if (get_plan(h, batch, &bp)) return 0;
  return bp->nar.ok ? (1 | (nar::has_backward() ? 2 : 0)) : (bp->nar_eval.ok ? 4 : 0);
}

int dpi_flow_tc_active(dpi_flow_t h, int batch) {
  if (!handle_live(h, 1) || batch < 1) return 0;
  DeviceGuard g(h->device);
  BatchPlan* bp = nullptr;
  if (get_plan(h, batch, &bp)) return 0;
  return bp->tc_ok;
}

}  // extern "C"
