// Kernels of the M-mode engine.  One warp per (chain, locus) pair for the genealogy work, one warp per
// chain for the prior sweep; a single thread replays the MC3 swap attempts (latency-only work).
//
//   k_eval_pairs     static evaluation of the current state (treeweight + P(D|G))
//   k_eval_chains    per chain: sum the weights over loci, integrate the prior from scratch
//   k_propose        updategenealogy steps 2-11: proposal + weights + likelihood,
//                    written to the pair's OTHER buffer (rejection is then free)
//   k_accept         per chain, loci in order: all-locus sums, integrated prior, MH accept (steps 12-14);
//                    the loci of one chain are coupled through the prior
//   k_swap           replicated replay of the step's swap attempts on (beta, S) of all chains
#pragma once
#include "ima_genealogy.h"
#include "ima_devapi.h"

namespace ima {

// the model tables live in __constant__ memory; this header is included by exactly one translation
// unit (ima_engine.cu), which therefore owns the definition
IMA_CONSTANT DevModel c_model;
#define IMA_MODEL c_model

constexpr int kWarpsPerBlock = 4;

IMA_DEV void rng_for(Philox &rng, const EngineView &E, uint32_t stream_id, uint32_t purpose) {
  const unsigned long long step = *E.nsteps;
  rng.init(E.seed, stream_id, (uint32_t)step, purpose | ((uint32_t)(step >> 32) << 8));
}

// P(D|G) of the staged genealogy for the locus's mutation model; every lane returns the value
IMA_DEV double pair_likelihood(const EngineView &E, const DevLocus &L, const PairBuf &B, int p, PairSm &S, double *pdg_a_out) {
  const double *u = E.uvals + (size_t)p * kMaxLinked;
  if (L.model == kInfiniteSites) {
    const double v = likelihood_is(E, L, S, u[0]);
    pdg_a_out[0] = v;
    return v;
  }
  if (has_stepwise(L.model)) {
    double tot = 0.0;
    if (L.model == kJointISSW) {                          // part 0 is the infinite-sites part (calc_prob_data / update_gtree.cpp:871-883)
      tot = likelihood_is(E, L, S, u[0]);
      pdg_a_out[0] = tot;
      if (tot == kRejectIS) return kRejectIS;
    }
    for (int ai = sw_first(L.model); ai < L.nlinked; ai++) {
      const short *A = B.A + ((size_t)p * kMaxLinked + ai) * E.d.NL;
      double *dl = B.dlikeA + ((size_t)p * kMaxLinked + ai) * E.d.NL;
      const double v = likelihood_sw(L, S, A, dl, u[ai]);
      pdg_a_out[ai] = v;
      tot += v;
    }
    return tot;
  }
  if (L.model == kHKY) {
    const double v = likelihood_hky(E, L, S, p, u[0], E.kappa[p], E.pi + (size_t)p * 4);
    pdg_a_out[0] = v;
    return v;
  }
  pdg_a_out[0] = 0.0;
  return 0.0;
}

IMA_KERNEL void k_eval_pairs(EngineView E) {
  IMA_SMEM_DECL
  const int p = ima_block() * kWarpsPerBlock + ima_warp_in_block();
  if (p >= E.d.P) return;
  const DevModel &M = IMA_MODEL;
  const int c = p / E.d.nloci, li = p - c * E.d.nloci;
  const DevLocus &L = E.loci[li];
  PairSm S = carve_pair_smem(IMA_SMEM + (size_t)ima_warp_in_block() * pair_smem_bytes(E.d), E.d);
  const PairBuf &B = E.buf[E.cur[p]];
  const double *tv = E.tvals + (size_t)c * kMaxPeriods;
  stage_pair(E, B, p, L.nl, S);
  const bool ok = eval_weights(M, E.d, L, tv, S);
  double pdga[kMaxLinked];
  const double pdg = ok ? pair_likelihood(E, L, B, p, S, pdga) : 0.0;
  const int lane = Warp::lane();
  for (int i = lane; i < E.d.NI; i += IMA_WARP) B.gwi[(size_t)p * E.d.NI + i] = S.gwi[i];
  for (int i = lane; i < E.d.ND; i += IMA_WARP) B.gwd[(size_t)p * E.d.ND + i] = S.gwd[i];
  if (lane == 0) {
    B.sd[(size_t)p * 4 + 1] = S.ctl_d[kCdLength];
    B.sd[(size_t)p * 4 + 2] = S.ctl_d[kCdTlength];
    B.sd[(size_t)p * 4 + 3] = pdg;
    if (B.pdg_a) for (int ai = 0; ai < L.nlinked; ai++) B.pdg_a[(size_t)p * kMaxLinked + ai] = pdga[ai];
    E.prop_flags[p] = ok ? (uint32_t)S.ctl_i[kCiFlags] : (uint32_t)kFlagOverflow;
  }
}

// gather (c, f, hc) of one parameter from a weight record (update_gtree_common.cpp:1985-1994, 2018-2030)
IMA_DEV void gather_q(const DevModel &M, int t, const int *wi, const double *wd, int &c, double &f, double &hc) {
  c = 0; f = 0.0; hc = 0.0;
  for (int j = 0; j < M.q_n[t]; j++) { const int x = M.q_idx[t][j]; c += wi[x]; f += wd[x]; hc += wd[M.ncc + x]; }
}
IMA_DEV void gather_m(const DevModel &M, int t, const int *wi, const double *wd, int &c, double &f) {
  c = 0; f = 0.0;
  for (int j = 0; j < M.m_n[t]; j++) { const int x = M.m_idx[t][j]; c += wi[M.ncc + x]; f += wd[2 * M.ncc + x]; }
}
IMA_DEV bool migration_allowed(const DevModel &M, const int *wi) {
  if (M.nomigration == 0)
    for (int i = 0; i < M.nomig_n; i++) if (wi[M.ncc + M.nomig_idx[i]] != 0) return false;
  return true;
}
IMA_DEV double mig_term(const DevModel &M, const MathCtx &mc, int t, int c, double f) {
  return M.expoprior ? integrate_migration_term_expo(mc, c, f, M.m_mean[t]) : integrate_migration_term(mc, c, f, M.m_max[t], M.m_min[t]);
}

// shared-memory scratch of the per-chain kernels
struct ChainSm { int *ai, *ci, *ic; double *ad, *cd, *q, *cq, *dc; };
IMA_HD size_t chain_smem_bytes(const EngineDims &d) {
  return 2 * align8(sizeof(int) * d.NI) + 2 * align8(sizeof(double) * d.ND) + 2 * align8(sizeof(double) * 2 * kMaxParams) + 64 + 32;
}
IMA_DEV ChainSm carve_chain_smem(unsigned char *base, const EngineDims &d) {
  ChainSm s; unsigned char *p = base;
  s.ad = (double *)p; p += align8(sizeof(double) * d.ND);
  s.cd = (double *)p; p += align8(sizeof(double) * d.ND);
  s.q = (double *)p; p += align8(sizeof(double) * 2 * kMaxParams);
  s.cq = (double *)p; p += align8(sizeof(double) * 2 * kMaxParams);
  s.dc = (double *)p; p += 64;
  s.ai = (int *)p; p += align8(sizeof(int) * d.NI);
  s.ci = (int *)p; p += align8(sizeof(int) * d.NI);
  s.ic = (int *)p;
  return s;
}

// S of swapweight (swapchains.cpp:12-34): sum over loci of pdg, plus probg unless thermodynamic mode
IMA_DEV double chain_swapsum(const EngineView &E, const DevModel &M, int c, double probg) {
  double s = 0.0;
  for (int li = Warp::lane(); li < E.d.nloci; li += IMA_WARP) {
    const int p = c * E.d.nloci + li;
    s += E.buf[E.cur[p]].sd[(size_t)p * 4 + 3];
  }
  s = Warp::sum(s);
  return M.thermo ? s : s + probg;
}

IMA_KERNEL void k_eval_chains(EngineView E) {
  IMA_SMEM_DECL
  const int c = ima_block() * kWarpsPerBlock + ima_warp_in_block();
  if (c >= E.d.nchains) return;
  const DevModel &M = IMA_MODEL;
  const int lane = Warp::lane(), NI = E.d.NI, ND = E.d.ND;
  ChainSm S = carve_chain_smem(IMA_SMEM + (size_t)ima_warp_in_block() * chain_smem_bytes(E.d), E.d);
  // sum_treeinfo over loci in locus order (init_p, mcmcfile.cpp:188)
  for (int i = lane; i < NI; i += IMA_WARP) {
    int a = 0;
    for (int li = 0; li < E.d.nloci; li++) { const int p = c * E.d.nloci + li; a += E.buf[E.cur[p]].gwi[(size_t)p * NI + i]; }
    S.ai[i] = a; E.all_i[(size_t)c * NI + i] = a;
  }
  for (int i = lane; i < ND; i += IMA_WARP) {
    double a = 0.0;
    for (int li = 0; li < E.d.nloci; li++) { const int p = c * E.d.nloci + li; a += E.buf[E.cur[p]].gwd[(size_t)p * ND + i]; }
    S.ad[i] = a; E.all_d[(size_t)c * ND + i] = a;
  }
  Warp::sync();
  // initialize_integrate_tree_prob (update_gtree_common.cpp:2056-2134), one lane per parameter
  double probg = 0.0;
  const int nterms = M.nq + (M.nomigration ? 0 : M.nm);
  for (int t = 0; t < nterms; t++) {                    // whole warp per term (see k_accept)
    double v;
    if (t < M.nq) {
      int cc; double f, hc;
      gather_q(M, t, S.ai, S.ad, cc, f, hc);
      v = integrate_coalescent_term_coop(E.mc, cc, f, hc, M.q_max[t], M.q_min[t]);
      if (lane == 0) E.qint[(size_t)c * kMaxParams + t] = v;
    } else {
      const int tm = t - M.nq;
      int cm; double f;
      gather_m(M, tm, S.ai, S.ad, cm, f);
      v = M.expoprior ? integrate_migration_term_expo(E.mc, cm, f, M.m_mean[tm]) : integrate_migration_term_coop(E.mc, cm, f, M.m_max[tm], M.m_min[tm]);
      if (lane == 0) E.mint[(size_t)c * kMaxParams + tm] = v;
    }
    probg += v;
  }
  if (!migration_allowed(M, S.ai)) probg = -kMyDblMax;
  double pd = 0.0;
  for (int li = lane; li < E.d.nloci; li += IMA_WARP) { const int p = c * E.d.nloci + li; pd += E.buf[E.cur[p]].sd[(size_t)p * 4 + 3]; }
  pd = Warp::sum(pd);
  const double ssum = chain_swapsum(E, M, c, probg);
  if (lane == 0) { E.probg[c] = probg; E.pdgsum[c] = pd; E.swapsum[c] = ssum; }
}

#ifndef IMA_PROPOSE_MINBLOCKS
#define IMA_PROPOSE_MINBLOCKS 5      // <= 102 registers/thread, 20 resident warps per SM: best of 5/6/8 measured on B200
#endif
#if IMA_CUDA
#define IMA_PROPOSE_BOUNDS __launch_bounds__(kWarpsPerBlock * 32, IMA_PROPOSE_MINBLOCKS)
#else
#define IMA_PROPOSE_BOUNDS
#endif
// loci [l0, l1) of every chain (one warp per pair); see launch_update for why a step is cut into pieces
IMA_KERNEL void IMA_PROPOSE_BOUNDS k_propose(EngineView E, int l0, int l1) {
  IMA_SMEM_DECL
  const int idx = ima_block() * kWarpsPerBlock + ima_warp_in_block();
  const int nsub = l1 - l0;
  if (idx >= E.d.nchains * nsub) return;
  const DevModel &M = IMA_MODEL;
  const int c = idx / nsub, li = l0 + (idx - c * nsub);
  const int p = c * E.d.nloci + li;
  const DevLocus &L = E.loci[li];
  PairSm S = carve_pair_smem(IMA_SMEM + (size_t)ima_warp_in_block() * pair_smem_bytes(E.d), E.d);
  const int cb = E.cur[p];
  const PairBuf &B = E.buf[cb];
  const PairBuf &Bn = E.buf[cb ^ 1];
  const double *tv = E.tvals + (size_t)c * kMaxPeriods;
  const int lane = Warp::lane();
  stage_pair(E, B, p, L.nl, S);
  if (lane == 0) {
    Philox rng;
    rng_for(rng, E, (uint32_t)((E.d.chain0 + c) * E.d.nloci + li), kRngPropose);
    propose_move(M, E.d, tv, L.ng, L.nl, rng, S);
  }
  Warp::sync();
  uint32_t flags = (uint32_t)S.ctl_i[kCiFlags];
  bool ok = !(flags & kFlagOverflow);
  if (ok) ok = eval_weights(M, E.d, L, tv, S);
  const int total_mig = ok ? S.ctl_i[kCiMignum] : 0;
  if (ok && total_mig > E.d.CAP) ok = false;
  double pdga[kMaxLinked];
  double pdg = 0.0;
  if (ok && has_stepwise(L.model)) {
    // stepwise loci: incremental update of the allele states and branch terms (update_gtree.cpp:857-868)
    const size_t ao = (size_t)p * kMaxLinked * E.d.NL;
    for (int i = lane; i < L.nlinked * E.d.NL; i += IMA_WARP) { Bn.A[ao + i] = B.A[ao + i]; Bn.dlikeA[ao + i] = B.dlikeA[ao + i]; }
#if IMA_CUDA
    __threadfence_block();
#endif
    Warp::sync();
    double pis = 0.0;
    if (L.model == kJointISSW) {                          // the infinite-sites part is evaluated in full, as for an I locus
      pis = likelihood_is(E, L, S, E.uvals[(size_t)p * kMaxLinked]);
      if (lane == 0) { Bn.pdg_a[(size_t)p * kMaxLinked] = pis; if (pis == kRejectIS) S.ctl_i[kCiFlags] |= kFlagRejectIS; }
      Warp::sync();
    }
    if (lane == 0) {
      Philox rng;
      rng_for(rng, E, (uint32_t)((E.d.chain0 + c) * E.d.nloci + li), kRngAlleles);
      double atermsum = 0.0, tot = pis;
      for (int ai = sw_first(L.model); ai < L.nlinked; ai++) {
        double aterm = 0.0;
        const double dl = sw_update_alleles(L, S, ai, rng, B.A + ao + (size_t)ai * E.d.NL, B.dlikeA + ao + (size_t)ai * E.d.NL,
                                            Bn.A + ao + (size_t)ai * E.d.NL, Bn.dlikeA + ao + (size_t)ai * E.d.NL, S.ctl_i[kCiEdge],
                                            S.ctl_i[kCiFreed], S.ctl_i[kCiOldsis], S.ctl_i[kCiNewsis], S.ctl_i[kCiOldDownDown],
                                            E.uvals[(size_t)p * kMaxLinked + ai], &aterm);
        const double v = B.pdg_a[(size_t)p * kMaxLinked + ai] + dl;
        Bn.pdg_a[(size_t)p * kMaxLinked + ai] = v;
        tot += v;
        atermsum += aterm;
      }
      S.ctl_d[kCdAterm] = atermsum;
      S.ctl_d[kCdPdg] = tot;
    }
    Warp::sync();
    pdg = S.ctl_d[kCdPdg];
    flags = (uint32_t)S.ctl_i[kCiFlags];
    if (!(pdg > -DBL_MAX)) flags |= kFlagRejectIS;       // a branch term of -inf (bessi == 0): the move cannot be accepted
    if (!(flags & (kFlagRejectIS | kFlagBadTree))) store_pair(E, Bn, p, L.nl, S, total_mig);
  } else if (ok) {
    pdg = pair_likelihood(E, L, Bn, p, S, pdga);
    flags = (uint32_t)S.ctl_i[kCiFlags];
    if (pdg == kRejectIS) flags |= kFlagRejectIS;
    if (lane == 0) S.ctl_d[kCdPdg] = pdg;
    Warp::sync();
    if (!(flags & (kFlagRejectIS | kFlagBadTree))) store_pair(E, Bn, p, L.nl, S, total_mig);
  } else {
    flags |= kFlagOverflow;
  }
  if (lane == 0) {
    E.prop_flags[p] = flags;
    E.prop_extra[p] = S.ctl_d[kCdMigw] + S.ctl_d[kCdSlidew] + S.ctl_d[kCdAterm];
    E.prop_dbg[(size_t)p * 4 + 0] = S.ctl_d[kCdMigw]; E.prop_dbg[(size_t)p * 4 + 1] = S.ctl_d[kCdSlidew];
    E.prop_dbg[(size_t)p * 4 + 2] = S.ctl_d[kCdSlideDist]; E.prop_dbg[(size_t)p * 4 + 3] = (double)S.ctl_i[kCiEdge];
  }
}

// what the accept sweep needs from one pair: per-lane slice of the weight records plus the scalars
struct AcceptRecord { int dI, cb; uint32_t flags; double oD, nD, oldpdg, newpdg, extra; };
IMA_DEV void fetch_accept_record(const EngineView &E, int p, int lane, int NI, int ND, AcceptRecord &r) {
  r.flags = E.prop_flags[p];
  r.cb = E.cur[p];
  const PairBuf &O = E.buf[r.cb], &N = E.buf[r.cb ^ 1];
  r.dI = (lane < NI) ? N.gwi[(size_t)p * NI + lane] - O.gwi[(size_t)p * NI + lane] : 0;
  r.oD = (lane < ND) ? O.gwd[(size_t)p * ND + lane] : 0.0;
  r.nD = (lane < ND) ? N.gwd[(size_t)p * ND + lane] : 0.0;
  r.oldpdg = O.sd[(size_t)p * 4 + 3];
  r.newpdg = N.sd[(size_t)p * 4 + 3];
  r.extra = E.prop_extra[p];
}

// ---- accept sweep ---------------------------------------------------------------------------------------
// One block per chain.  The loci of a chain must be decided in order: they are coupled through the integrated
// prior, so locus li sees the sums left by every accepted update before it.  Two things
// shorten that dependent chain without changing its result:
//   * inside a locus the nq + nm prior terms are independent: each goes to its own warp, which runs the term's
//     series / continued fraction 32 terms per round (ima_math.h *_coop);
//   * the next B-1 loci are evaluated SPECULATIVELY in the same round, against the same sums.  Decisions are then
//     taken in locus order; the first acceptance changes the sums, so every speculative locus after it is
//     thrown away and re-evaluated in the next round.  About two thirds of the updates are rejected, so a round
//     decides 1 + q + q^2 ... loci on average (q = rejection rate).  The random number of a locus depends only on
//     (chain, locus, step), so the outcome is identical to the one-locus-at-a-time sweep.
// A loader warp streams the pairs' weight records (and draws the uniforms) a few loci ahead into a ring in
// shared memory, so global-memory latency is off the critical path.
// Loci [l0, l1): the sums travel through global memory between launches.
#if IMA_CUDA
constexpr int kSpecMax = 3;          // speculative depth B
constexpr int kTermWarps = 5;        // warps per speculative locus (terms are strided over them)
IMA_DEV void block_sync() { __syncthreads(); }
#define IMA_FOR_WARPS(wv, nw) for (int wv = ima_warp_in_block(), once_ = 1; once_; once_ = 0)
#else
constexpr int kSpecMax = 3;
constexpr int kTermWarps = 5;
IMA_DEV void block_sync() {}
#define IMA_FOR_WARPS(wv, nw) for (int wv = 0; wv < (nw); wv++)      // host emulation: one thread plays every warp in turn
#endif
constexpr int kRing = 8;             // loci buffered ahead

struct AcceptSm {
  int *ai; double *ad, *q;                                   // all-locus sums and current prior terms
  double *cq;                                                // [kSpecMax][2*kMaxParams] candidate terms per speculative locus
  int *r_dI; double *r_oD, *r_nD, *r_sc; int *r_ic;          // ring: [kRing] records
  int *ctl;                                                  // [4]: accepted group (or -1), advance
  double *dctl;                                              // [2]: new probg
};
IMA_HD size_t accept_smem_bytes(const EngineDims &d) {
  return align8(sizeof(int) * d.NI) + align8(sizeof(double) * d.ND) + align8(sizeof(double) * 2 * kMaxParams) +
         align8(sizeof(double) * kSpecMax * 2 * kMaxParams) + kRing * (align8(sizeof(int) * d.NI) + 2 * align8(sizeof(double) * d.ND) + 5 * 8 + 16) + 64;
}
IMA_DEV AcceptSm carve_accept_smem(unsigned char *base, const EngineDims &d) {
  AcceptSm s; unsigned char *p = base;
  s.ad = (double *)p; p += align8(sizeof(double) * d.ND);
  s.q = (double *)p; p += align8(sizeof(double) * 2 * kMaxParams);
  s.cq = (double *)p; p += align8(sizeof(double) * kSpecMax * 2 * kMaxParams);
  s.r_oD = (double *)p; p += kRing * align8(sizeof(double) * d.ND);
  s.r_nD = (double *)p; p += kRing * align8(sizeof(double) * d.ND);
  s.r_sc = (double *)p; p += kRing * 5 * 8;                  // oldpdg, newpdg, extra, uniform, (pad)
  s.dctl = (double *)p; p += 16;
  s.ai = (int *)p; p += align8(sizeof(int) * d.NI);
  s.r_dI = (int *)p; p += kRing * align8(sizeof(int) * d.NI);
  s.r_ic = (int *)p; p += kRing * 16;                        // flags, cb
  s.ctl = (int *)p;
  return s;
}

// depth 3: one 16-warp block per SM (<= 128 registers); depth 1 and 2: two blocks per SM (<= 93 registers), which is
// what a GPU holding more chains than it has SMs needs
#if IMA_CUDA
#define IMA_ACCEPT_BOUNDS(B) __launch_bounds__(((B) * kTermWarps + 1) * 32, (B) == 3 ? 1 : 2)
#else
#define IMA_ACCEPT_BOUNDS(B)
#endif
template <int B>
IMA_KERNEL void IMA_ACCEPT_BOUNDS(B) k_accept(EngineView E, int l0, int l1) {
  IMA_SMEM_DECL
  const int c = ima_block();
  if (c >= E.d.nchains) return;
  const DevModel &M = IMA_MODEL;
  const int NW = B * kTermWarps + 1, LOADER = B * kTermWarps;       // warps in the block; the last one loads
  const int lane = Warp::lane(), NI = E.d.NI, ND = E.d.ND, ncc = M.ncc;
  const int sI = (int)(align8(sizeof(int) * NI) / sizeof(int)), sD = (int)(align8(sizeof(double) * ND) / sizeof(double));
  AcceptSm S = carve_accept_smem(IMA_SMEM, E.d);
  const int nterms = M.nq + (M.nomigration ? 0 : M.nm);
  // ring slot of locus l: l % kRing.  The loader fills loci [from, upto): every load of every locus of the batch is
  // issued before any is consumed (independent addresses), and the uniforms are drawn one locus per lane.
  auto load_records = [&](int from, int upto) {
    for (int l = from + lane; l < upto; l += IMA_WARP) {
      const int p = c * E.d.nloci + l, slot = l % kRing;
      const int cb = E.cur[p];
      const PairBuf &O = E.buf[cb], &N = E.buf[cb ^ 1];
      S.r_ic[slot * 4 + 0] = (int)E.prop_flags[p]; S.r_ic[slot * 4 + 1] = cb;
      S.r_sc[slot * 5 + 0] = O.sd[(size_t)p * 4 + 3]; S.r_sc[slot * 5 + 1] = N.sd[(size_t)p * 4 + 3]; S.r_sc[slot * 5 + 2] = E.prop_extra[p];
      Philox rng;
      rng_for(rng, E, (uint32_t)((E.d.chain0 + c) * E.d.nloci + l), kRngAccept);
      S.r_sc[slot * 5 + 3] = rng.uniform();
    }
    const int nl = upto - from, per = NI + 2 * ND;
    for (int k = lane; k < nl * per; k += IMA_WARP) {
      const int l = from + k / per, i = k - (l - from) * per;
      const int p = c * E.d.nloci + l, slot = l % kRing;
      const int cb = E.cur[p];
      const PairBuf &O = E.buf[cb], &N = E.buf[cb ^ 1];
      if (i < NI) S.r_dI[slot * sI + i] = N.gwi[(size_t)p * NI + i] - O.gwi[(size_t)p * NI + i];
      else if (i < NI + ND) S.r_oD[slot * sD + (i - NI)] = O.gwd[(size_t)p * ND + (i - NI)];
      else S.r_nD[slot * sD + (i - NI - ND)] = N.gwd[(size_t)p * ND + (i - NI - ND)];
    }
  };
  IMA_FOR_WARPS(w, NW) {
    const int tid = w * IMA_WARP + lane, nth = NW * IMA_WARP;
    for (int i = tid; i < NI; i += nth) S.ai[i] = E.all_i[(size_t)c * NI + i];
    for (int i = tid; i < ND; i += nth) S.ad[i] = E.all_d[(size_t)c * ND + i];
    for (int i = tid; i < M.nq; i += nth) S.q[i] = E.qint[(size_t)c * kMaxParams + i];
    for (int i = tid; i < M.nm; i += nth) S.q[kMaxParams + i] = E.mint[(size_t)c * kMaxParams + i];
    if (w == LOADER) load_records(l0, (l0 + kRing < l1) ? l0 + kRing : l1);
  }
  const double beta = E.beta[c];
  double probg = E.probg[c], pdgsum = E.pdgsum[c];
  unsigned long long dropped = 0;
  int filled = (l0 + kRing < l1) ? l0 + kRing : l1;
  block_sync();
  constexpr uint32_t kNoGo = kFlagRejectIS | kFlagOverflow | kFlagBadTree;
  for (int li = l0; li < l1;) {
    // A proposal that arrives flagged (the data rule it out, or it was dropped) is rejected whatever the prior says: such
    // loci are passed over without spending a speculative slot on them.  cand[g] = offset from li of the g-th locus that
    // needs a decision among the loci already in the ring, span = loci consumed when none of them is accepted.
    int cand[B], nb = 0;
    const int visible = filled - li;
    for (int k = 0; k < visible && nb < B; k++)
      if (!((uint32_t)S.r_ic[((li + k) % kRing) * 4] & kNoGo)) cand[nb++] = k;
    for (int g = nb; g < B; g++) cand[g] = 0;
    const int span = nb == B ? cand[B - 1] + 1 : visible;
    // ---- phase 1: term t of speculative candidate g on warp g*kTermWarps + (t % kTermWarps) ------------------
    const int upto = (li + kRing < l1) ? li + kRing : l1;   // loci < li are decided: their ring slots are free again
    IMA_FOR_WARPS(w, NW) {
      const int g = w / kTermWarps, t0 = w - g * kTermWarps;
      if (w == LOADER) load_records(filled, upto);
      if (w != LOADER && g < nb) {
        const int slot = (li + cand[g]) % kRing;
        {
          const int *dI = S.r_dI + slot * sI;
          const double *oD = S.r_oD + slot * sD, *nD = S.r_nD + slot * sD;
          // candidate sums = sum_subtract_treeinfo (ginfo.cpp:248-285) on the entries this term reads: subtract old, add
          // new, clamp fc and fm at 0; then integrate_tree_prob's reuse rule (:1997-2000, :2031-2034) or the integral
          for (int t = t0; t < nterms; t += kTermWarps) {
            double v;
            if (t < M.nq) {
              int cn = 0, co = 0; double fn = 0.0, fo = 0.0, hn = 0.0;
              for (int j = 0; j < M.q_n[t]; j++) {
                const int x = M.q_idx[t][j];
                co += S.ai[x]; cn += S.ai[x] + dI[x];
                fo += S.ad[x];
                double f = S.ad[x]; f -= oD[x]; f += nD[x]; if (0.0 > f) f = 0.0;
                fn += f;
                double hh = S.ad[ncc + x]; hh -= oD[ncc + x]; hh += nD[ncc + x];
                hn += hh;
              }
              v = (cn == co && fn == fo) ? S.q[t] : integrate_coalescent_term_coop(E.mc, cn, fn, hn, M.q_max[t], M.q_min[t]);
              if (lane == 0) S.cq[g * 2 * kMaxParams + t] = v;
            } else {
              const int tm = t - M.nq;
              int cn = 0, co = 0; double fn = 0.0, fo = 0.0;
              for (int j = 0; j < M.m_n[tm]; j++) {
                const int x = M.m_idx[tm][j];
                co += S.ai[ncc + x]; cn += S.ai[ncc + x] + dI[ncc + x];
                fo += S.ad[2 * ncc + x];
                double f = S.ad[2 * ncc + x]; f -= oD[2 * ncc + x]; f += nD[2 * ncc + x]; if (0.0 > f) f = 0.0;
                fn += f;
              }
              v = (cn == co && fn == fo) ? S.q[kMaxParams + tm]
                  : (M.expoprior ? integrate_migration_term_expo(E.mc, cn, fn, M.m_mean[tm])
                                 : integrate_migration_term_coop(E.mc, cn, fn, M.m_max[tm], M.m_min[tm]));
              if (lane == 0) S.cq[g * 2 * kMaxParams + kMaxParams + tm] = v;
            }
          }
        }
      }
    }
    block_sync();
    // ---- phase 2: decisions in locus order (update_gtree.cpp:917-927) ----------------------------------------
    IMA_FOR_WARPS(w, NW) {
      if (w == 0) {
        // lane g evaluates the MH ratio of speculative locus li+g (all against the same sums); the first accepting
        // lane in locus order wins
        bool acc = false;
        double newprobg = 0.0;
        if (lane < nb) {
          const int g = lane, slot = (li + cand[lane]) % kRing;
          {
            for (int t = 0; t < M.nq; t++) newprobg += S.cq[g * 2 * kMaxParams + t];
            if (!M.nomigration) for (int t = 0; t < M.nm; t++) newprobg += S.cq[g * 2 * kMaxParams + kMaxParams + t];
            if (M.nomigration == 0)
              for (int i = 0; i < M.nomig_n; i++)
                if (S.ai[ncc + M.nomig_idx[i]] + S.r_dI[slot * sI + ncc + M.nomig_idx[i]] != 0) newprobg = -kMyDblMax;
            const double tpw = newprobg - probg, dpdg = S.r_sc[slot * 5 + 1] - S.r_sc[slot * 5 + 0], extra = S.r_sc[slot * 5 + 2];
            double mh;
            if (M.thermo) mh = exp(beta * M.gbeta * dpdg + tpw + extra);
            else mh = exp(beta * (tpw + M.gbeta * dpdg) + extra);
            acc = S.r_sc[slot * 5 + 3] < fmin(1.0, mh);
          }
        }
#if IMA_CUDA
        const int accepted = Warp::first(acc);
        const double np = Warp::bcast(newprobg, accepted < 0 ? 0 : accepted);
        if (lane == 0) {
          S.ctl[0] = accepted; S.ctl[1] = accepted < 0 ? span : cand[accepted < 0 ? 0 : accepted] + 1; S.ctl[2] = accepted < 0 ? 0 : cand[accepted];
          if (accepted >= 0) S.dctl[0] = np;
        }
#else
        // one lane: walk the speculative loci in order
        int accepted = -1;
        for (int g = 0; g < nb && accepted < 0; g++) {
          const int slot = (li + cand[g]) % kRing;
          double npg = 0.0;
          for (int t = 0; t < M.nq; t++) npg += S.cq[g * 2 * kMaxParams + t];
          if (!M.nomigration) for (int t = 0; t < M.nm; t++) npg += S.cq[g * 2 * kMaxParams + kMaxParams + t];
          if (M.nomigration == 0)
            for (int i = 0; i < M.nomig_n; i++)
              if (S.ai[ncc + M.nomig_idx[i]] + S.r_dI[slot * sI + ncc + M.nomig_idx[i]] != 0) npg = -kMyDblMax;
          const double tpw = npg - probg, dpdg = S.r_sc[slot * 5 + 1] - S.r_sc[slot * 5 + 0], extra = S.r_sc[slot * 5 + 2];
          double mh;
          if (M.thermo) mh = exp(beta * M.gbeta * dpdg + tpw + extra);
          else mh = exp(beta * (tpw + M.gbeta * dpdg) + extra);
          if (S.r_sc[slot * 5 + 3] < fmin(1.0, mh)) { accepted = g; S.dctl[0] = npg; }
        }
        S.ctl[0] = accepted; S.ctl[1] = accepted < 0 ? span : cand[accepted] + 1; S.ctl[2] = accepted < 0 ? 0 : cand[accepted];
        (void)acc; (void)newprobg;
#endif
      }
    }
    block_sync();
    // ---- phase 3: commit the accepted locus (if any) -----------------------------------------------------
    const int accepted = S.ctl[0], adv = S.ctl[1], aoff = S.ctl[2];      // accepted candidate (or -1), loci consumed, its offset
    for (int g = 0; g < adv; g++) if ((uint32_t)S.r_ic[((li + g) % kRing) * 4] & kFlagOverflow) dropped++;
    IMA_FOR_WARPS(w, NW) {
      const int tid = w * IMA_WARP + lane, nth = (NW - 1) * IMA_WARP;
      if (w != LOADER && accepted >= 0) {
        const int slot = (li + aoff) % kRing;
        const int *dI = S.r_dI + slot * sI;
        const double *oD = S.r_oD + slot * sD, *nD = S.r_nD + slot * sD;
        for (int i = tid; i < NI; i += nth) S.ai[i] += dI[i];
        for (int i = tid; i < ND; i += nth) {
          double x = S.ad[i];
          x -= oD[i];
          x += nD[i];
          if ((i < ncc || i >= 2 * ncc) && 0.0 > x) x = 0.0;
          S.ad[i] = x;
        }
        for (int i = tid; i < M.nq; i += nth) S.q[i] = S.cq[accepted * 2 * kMaxParams + i];
        for (int i = tid; i < M.nm; i += nth) S.q[kMaxParams + i] = S.cq[accepted * 2 * kMaxParams + kMaxParams + i];
        if (tid == 0) {
          const int p = c * E.d.nloci + li + aoff;
          const uint32_t flags = (uint32_t)S.r_ic[slot * 4];
          E.cur[p] = (unsigned char)(S.r_ic[slot * 4 + 1] ^ 1);
          E.acc[(size_t)p * 3 + 0]++;
          if (flags & kFlagTopol) E.acc[(size_t)p * 3 + 1]++;
          if (flags & kFlagTmrca) E.acc[(size_t)p * 3 + 2]++;
        }
      }
    }
    filled = upto;
    if (accepted >= 0) {
      const int slot = (li + aoff) % kRing;
      probg = S.dctl[0];
      pdgsum -= S.r_sc[slot * 5 + 0];
      pdgsum += S.r_sc[slot * 5 + 1];
    }
    li += adv;
    block_sync();
  }
  IMA_FOR_WARPS(w, NW) {
    const int tid = w * IMA_WARP + lane, nth = NW * IMA_WARP;
    for (int i = tid; i < NI; i += nth) E.all_i[(size_t)c * NI + i] = S.ai[i];
    for (int i = tid; i < ND; i += nth) E.all_d[(size_t)c * ND + i] = S.ad[i];
    for (int i = tid; i < M.nq; i += nth) E.qint[(size_t)c * kMaxParams + i] = S.q[i];
    for (int i = tid; i < M.nm; i += nth) E.mint[(size_t)c * kMaxParams + i] = S.q[kMaxParams + i];
  }
#if IMA_CUDA
  __threadfence_block();
#endif
  block_sync();
  IMA_FOR_WARPS(w, NW) {
    if (w == 0) {
      const double ssum = (l1 == E.d.nloci) ? chain_swapsum(E, M, c, probg) : 0.0;
      if (lane == 0) {
        E.probg[c] = probg; E.pdgsum[c] = pdgsum;
        if (l1 == E.d.nloci) E.swapsum[c] = ssum;
        if (dropped) {
#if IMA_CUDA
          atomicAdd(E.overflow, dropped);
#else
          *E.overflow += dropped;
#endif
        }
      }
    }
  }
}

// MC3 swaps in temperature-rank form (swapchains_bwprocesses, swapchains.cpp:192-523; accept rule
// swapchains.cpp:57-62, 603-605): only betas move.  Every rank replays the same attempts on the same
// all-gathered S with the same counter-based stream, so no further message is needed.
struct SwapView {
  const double *S_global;   // [nchains_global]
  int *rank_of_chain;       // [nchains_global]
  int *chain_of_rank;       // [nchains_global]
  const double *beta_table; // [nchains_global] beta by temperature rank (rank 0 = cold)
  unsigned long long *swap_counts;   // [2] attempts, accepts
  int swaptries, advance_step;
  int smem_chains;          // chains the launch's shared memory can stage (0: work on global memory)
};

IMA_KERNEL void k_swap(EngineView E, SwapView V) {
  IMA_SMEM_DECL
  if (ima_block() != 0 || ima_warp_in_block() != 0) return;
  const int N = E.d.nchains_global;
  const int lane = Warp::lane();
  if (N > 1 && V.swaptries > 0) {
    // the attempts are a dependent chain of small decisions over (beta, S, rank): the warp stages those in shared memory
    // (when they fit) so that one lane walks the attempts without waiting on global memory, and writes the ranks back
    const bool staged = V.smem_chains >= N;
    double *sS = (double *)IMA_SMEM, *sB = sS + (staged ? N : 0);
    int *sC = (int *)(sB + (staged ? N : 0)), *sR = sC + (staged ? N : 0);
    if (staged) {
      for (int i = lane; i < N; i += IMA_WARP) { sS[i] = V.S_global[i]; sB[i] = V.beta_table[i]; sC[i] = V.chain_of_rank[i]; sR[i] = V.rank_of_chain[i]; }
#if IMA_CUDA
      __threadfence_block();
#endif
      Warp::sync();
    }
    const double *Sg = staged ? sS : V.S_global, *Bt = staged ? sB : V.beta_table;
    int *cor = staged ? sC : V.chain_of_rank, *roc = staged ? sR : V.rank_of_chain;
    if (lane == 0) {
      Philox rng;
      rng_for(rng, E, 0xffffffffu, kRngSwap);
      unsigned long long nacc = 0;
      for (int x = 0; x < V.swaptries; x++) {
        const int sa = rng.randint(N);
        int sbmin = 0, sbrange = N;
        if (N >= 2 * kSwapDist + 3) {
          sbmin = sa - kSwapDist > 0 ? sa - kSwapDist : 0;
          sbrange = (N < sa + kSwapDist ? N : sa + kSwapDist) - sbmin;
        }
        int sb;
        do { sb = sbmin + rng.randint(sbrange); } while (sb == sa);
        const int ca = cor[sa], cb = cor[sb];
        const double w = exp((Bt[sa] - Bt[sb]) * (Sg[cb] - Sg[ca]));
        if (w >= 1.0 || w > rng.uniform()) {
          cor[sa] = cb; cor[sb] = ca;
          roc[ca] = sb; roc[cb] = sa;
          nacc++;
        }
      }
      V.swap_counts[0] += (unsigned long long)V.swaptries;
      V.swap_counts[1] += nacc;
    }
#if IMA_CUDA
    __threadfence_block();
#endif
    Warp::sync();
    if (staged) for (int i = lane; i < N; i += IMA_WARP) { V.chain_of_rank[i] = sC[i]; V.rank_of_chain[i] = sR[i]; }
    for (int c = lane; c < E.d.nchains; c += IMA_WARP) E.beta[c] = Bt[roc[E.d.chain0 + c]];
  }
  if (V.advance_step && lane == 0) *E.nsteps += 1;
}

// parity hook: one warp per (a, x) pair
IMA_KERNEL void k_debug_gamma(MathCtx mc, const int *a, const double *x, int n, double *out) {
  const int i = ima_block() * kWarpsPerBlock + ima_warp_in_block();
  if (i >= n) return;
  const double uc = uppergamma_coop(mc, a[i], x[i]);
  const double lc = a[i] > 0 ? lowergamma_coop(mc, a[i], x[i]) : 0.0;
  if (Warp::lane() == 0) {
    out[4 * i + 0] = uppergamma(mc, a[i], x[i]);
    out[4 * i + 1] = a[i] > 0 ? lowergamma(mc, a[i], x[i]) : 0.0;
    out[4 * i + 2] = uc;
    out[4 * i + 3] = lc;
  }
}

// summarginlikecalc (marglike.cpp:51-87): thermosum[i] += allpcalc.pdg of the chain heated at beta_i.  Whole chains
// move between temperatures in the serial reference; here only betas move, so the slot is the temperature rank.
IMA_KERNEL void k_thermo_accumulate(EngineView E, const int *rank_of_chain, double *thermosum) {
  const int i = ima_block() * kWarpsPerBlock * IMA_WARP + ima_warp_in_block() * IMA_WARP + Warp::lane();
  if (i < E.d.nchains) thermosum[rank_of_chain[E.d.chain0 + i]] += E.pdgsum[i];
}

// What the host reads after a step, packed for one copy: out[0 .. 4 nchains) = beta, probg, pdg, S of every local chain;
// then rowlen floats (as doubles) = the cold chain's .ti row, savegsampinf ginfo.cpp:318-377 with its float sums; then
// 1 if the cold chain lives here; then the device error word.
IMA_KERNEL void k_pack_report(EngineView E, const int *chain_of_rank, int rowlen, double *out) {
  if (ima_block() != 0 || ima_warp_in_block() != 0) return;
  const DevModel &M = IMA_MODEL;
  const int lane = Warp::lane(), C = E.d.nchains;
  for (int c = lane; c < C; c += IMA_WARP) {
    out[4 * c] = E.beta[c]; out[4 * c + 1] = E.probg[c]; out[4 * c + 2] = E.pdgsum[c]; out[4 * c + 3] = E.swapsum[c];
  }
  if (lane != 0) return;
  double *row = out + 4 * (size_t)C;
  const int c = chain_of_rank[0] - E.d.chain0;
  const bool here = c >= 0 && c < C;
  row[rowlen] = here ? 1.0 : 0.0;
  row[rowlen + 1] = (double)*E.mc.err;
  if (!here) return;
  const int *wi = E.all_i + (size_t)c * E.d.NI;
  const double *wd = E.all_d + (size_t)c * E.d.ND;
  const int nq = M.nq, nm = M.nomigration ? 0 : M.nm;
  const int fcp = nq, hccp = 2 * nq, mcp = 3 * nq, fmp = mcp + nm, qip = fmp + nm, mip = qip + nq, pdgp = mip + nm;
  for (int i = 0; i < nq; i++) {
    int cc = 0; float f = 0.f, hc = 0.f;
    for (int j = 0; j < M.q_n[i]; j++) { const int x = M.q_idx[i][j]; cc += wi[x]; f += (float)wd[x]; hc += (float)wd[M.ncc + x]; }
    row[i] = (float)cc; row[fcp + i] = f; row[hccp + i] = hc; row[qip + i] = (float)E.qint[(size_t)c * kMaxParams + i];
  }
  for (int i = 0; i < nm; i++) {
    int cm = 0; float f = 0.f;
    for (int j = 0; j < M.m_n[i]; j++) { const int x = M.m_idx[i][j]; cm += wi[M.ncc + x]; f += (float)wd[2 * M.ncc + x]; }
    row[mcp + i] = (float)cm; row[fmp + i] = f; row[mip + i] = (float)E.mint[(size_t)c * kMaxParams + i];
  }
  row[pdgp] = (float)E.pdgsum[c]; row[pdgp + 1] = (float)E.probg[c];
  for (int i = 0; i < M.nsplit; i++) row[pdgp + 2 + i] = (float)E.tvals[(size_t)c * kMaxPeriods + i];
}

IMA_KERNEL void k_copy_swapsum(EngineView E, double *dst) {
  const int i = ima_block() * kWarpsPerBlock * IMA_WARP + ima_warp_in_block() * IMA_WARP + Warp::lane();
  if (i < E.d.nchains) dst[i] = E.swapsum[i];
}

}  // namespace ima
