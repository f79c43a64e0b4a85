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
  if (L.model == kStepwise) {
    double tot = 0.0;
    for (int ai = 0; ai < L.nlinked; ai++) {
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
#define IMA_PROPOSE_MINBLOCKS 6      // 80 registers/thread: 24 resident warps per SM (the kernel is latency-bound)
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
  if (ok && L.model == kStepwise) {
    // stepwise loci: incremental update of the allele states and branch terms (update_gtree.cpp:857-868)
    const size_t ao = (size_t)p * kMaxLinked * E.d.NL;
    for (int i = lane; i < L.nlinked * E.d.NL; i += IMA_WARP) { Bn.A[ao + i] = B.A[ao + i]; Bn.dlikeA[ao + i] = B.dlikeA[ao + i]; }
#if IMA_CUDA
    __threadfence_block();
#endif
    Warp::sync();
    if (lane == 0) {
      Philox rng;
      rng_for(rng, E, (uint32_t)((E.d.chain0 + c) * E.d.nloci + li), kRngAlleles);
      double atermsum = 0.0, tot = 0.0;
      for (int ai = 0; ai < L.nlinked; ai++) {
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

// One block per chain.  The loci of a chain are swept in order (they are coupled through the integrated prior,
// SURVEY.md fact 1); inside a locus the nq + nm prior terms are independent, so each is given to its own warp
// (which evaluates the term's series / continued fraction 32 terms per round, ima_math.h *_coop):
//   warp 0      fetch (prefetched one locus ahead) the pair's weight records, build all +- weights with the clamps
//   warp t      term t: gather (c, f, hc); reuse rule (:1997-2000, :2031-2034) or integrate_*_term
//   thread 0    sum the terms in parameter order, MH decision (update_gtree.cpp:917-927)
//   all         commit on accept
// Loci [l0, l1): the sums travel through global memory between launches, so a step may be cut into pieces.
#if IMA_CUDA
constexpr int kAcceptWarps = 8;
IMA_DEV void block_sync() { __syncthreads(); }
#else
constexpr int kAcceptWarps = 1;      // host emulation: one "warp" takes the terms one after the other
IMA_DEV void block_sync() {}
#endif
enum { kAcFlags = 0, kAcCb, kAcAccept };
enum { kAdOldPdg = 0, kAdNewPdg, kAdExtra, kAdNewProbg, kAdUniform };

#if IMA_CUDA
#define IMA_ACCEPT_BOUNDS __launch_bounds__(kAcceptWarps * 32, 2)
#else
#define IMA_ACCEPT_BOUNDS
#endif
IMA_KERNEL void IMA_ACCEPT_BOUNDS k_accept(EngineView E, int l0, int l1) {
  IMA_SMEM_DECL
  const int c = ima_block();
  if (c >= E.d.nchains) return;
  const DevModel &M = IMA_MODEL;
  const int lane = Warp::lane(), w = ima_warp_in_block(), tid = w * IMA_WARP + lane, nth = kAcceptWarps * IMA_WARP;
  const int NI = E.d.NI, ND = E.d.ND, ncc = M.ncc;
  ChainSm S = carve_chain_smem(IMA_SMEM, E.d);
  for (int i = tid; i < NI; i += nth) S.ai[i] = E.all_i[(size_t)c * NI + i];
  for (int i = tid; i < ND; i += nth) S.ad[i] = E.all_d[(size_t)c * ND + i];
  for (int i = tid; i < M.nq; i += nth) S.q[i] = E.qint[(size_t)c * kMaxParams + i];
  for (int i = tid; i < M.nm; i += nth) S.q[kMaxParams + i] = E.mint[(size_t)c * kMaxParams + i];
  const double beta = E.beta[c];
  double probg = E.probg[c], pdgsum = E.pdgsum[c];
  const int nterms = M.nq + (M.nomigration ? 0 : M.nm);
  unsigned long long dropped = 0;
  AcceptRecord nxt;
  if (w == 0) fetch_accept_record(E, c * E.d.nloci + l0, lane, NI, ND, nxt);
  block_sync();
  for (int li = l0; li < l1; li++) {
    const int p = c * E.d.nloci + li;
    if (w == 0) {
      // records of locus li+1 are fetched while locus li is being decided (the sweep is a dependent chain, so
      // global-memory latency would otherwise sit on the critical path of every locus)
      const AcceptRecord rec = nxt;
      if (li + 1 < l1) fetch_accept_record(E, p + 1, lane, NI, ND, nxt);
      const int cb = rec.cb, nb = cb ^ 1;
      // sum_subtract_treeinfo (ginfo.cpp:248-285): subtract old, add new, clamp fc and fm at 0
      if (lane < NI) S.ci[lane] = S.ai[lane] + rec.dI;
      if (lane < ND) {
        double x = S.ad[lane];
        x -= rec.oD;
        x += rec.nD;
        if ((lane < ncc || lane >= 2 * ncc) && 0.0 > x) x = 0.0;
        S.cd[lane] = x;
      }
      if (NI > IMA_WARP || ND > IMA_WARP) {             // records wider than a warp (>= 4 populations): direct loads
        const int *oi = E.buf[cb].gwi + (size_t)p * NI, *ni = E.buf[nb].gwi + (size_t)p * NI;
        const double *od = E.buf[cb].gwd + (size_t)p * ND, *nd = E.buf[nb].gwd + (size_t)p * ND;
        for (int i = lane + (IMA_WARP == 1 ? 1 : IMA_WARP); i < NI; i += IMA_WARP) S.ci[i] = S.ai[i] + (ni[i] - oi[i]);
        for (int i = lane + (IMA_WARP == 1 ? 1 : IMA_WARP); i < ND; i += IMA_WARP) {
          double x = S.ad[i];
          x -= od[i];
          x += nd[i];
          if ((i < ncc || i >= 2 * ncc) && 0.0 > x) x = 0.0;
          S.cd[i] = x;
        }
      }
      if (lane == 0) {
        S.ic[kAcFlags] = (int)rec.flags; S.ic[kAcCb] = cb;
        S.dc[kAdOldPdg] = rec.oldpdg; S.dc[kAdNewPdg] = rec.newpdg; S.dc[kAdExtra] = rec.extra;
      }
    }
    block_sync();
    const uint32_t flags = (uint32_t)S.ic[kAcFlags];
    if (flags & (kFlagRejectIS | kFlagOverflow | kFlagBadTree)) {
      if (flags & kFlagOverflow) dropped++;
      block_sync();                                      // warp 0 may not overwrite the control words before all have read them
      continue;
    }
    // the MH uniform of this locus is drawn by the last warp while the terms are being evaluated
    if (w == kAcceptWarps - 1 && lane == 0) {
      Philox rng;
      rng_for(rng, E, (uint32_t)((E.d.chain0 + c) * E.d.nloci + li), kRngAccept);
      S.dc[kAdUniform] = rng.uniform();
    }
    // integrate_tree_prob (update_gtree_common.cpp:1944-2053): term t on warp t
    for (int t = w; t < nterms; t += kAcceptWarps) {
      double v;
      if (t < M.nq) {
        int cn, co; double fn, fo, hn, ho;
        gather_q(M, t, S.ci, S.cd, cn, fn, hn);
        gather_q(M, t, S.ai, S.ad, co, fo, ho);
        v = (cn == co && fn == fo) ? S.q[t] : integrate_coalescent_term_coop(E.mc, cn, fn, hn, M.q_max[t], M.q_min[t]);
        if (lane == 0) S.cq[t] = v;
      } else {
        const int tm = t - M.nq;
        int cn, co; double fn, fo;
        gather_m(M, tm, S.ci, S.cd, cn, fn);
        gather_m(M, tm, S.ai, S.ad, co, fo);
        v = (cn == co && fn == fo) ? S.q[kMaxParams + tm]
            : (M.expoprior ? integrate_migration_term_expo(E.mc, cn, fn, M.m_mean[tm])
                           : integrate_migration_term_coop(E.mc, cn, fn, M.m_max[tm], M.m_min[tm]));
        if (lane == 0) S.cq[kMaxParams + tm] = v;
      }
    }
    block_sync();
    if (tid == 0) {
      double newprobg = 0.0;
      for (int t = 0; t < M.nq; t++) newprobg += S.cq[t];
      if (!M.nomigration) for (int t = 0; t < M.nm; t++) newprobg += S.cq[kMaxParams + t];
      if (!migration_allowed(M, S.ci)) newprobg = -kMyDblMax;
      const double tpw = newprobg - probg, dpdg = S.dc[kAdNewPdg] - S.dc[kAdOldPdg], extra = S.dc[kAdExtra];
      double mh;                                        // update_gtree.cpp:917-927
      if (M.thermo) mh = exp(beta * M.gbeta * dpdg + tpw + extra);
      else mh = exp(beta * (tpw + M.gbeta * dpdg) + extra);
      S.ic[kAcAccept] = (S.dc[kAdUniform] < fmin(1.0, mh)) ? 1 : 0;
      S.dc[kAdNewProbg] = newprobg;
    }
    block_sync();
    if (S.ic[kAcAccept]) {
      for (int i = tid; i < NI; i += nth) S.ai[i] = S.ci[i];
      for (int i = tid; i < ND; i += nth) S.ad[i] = S.cd[i];
      for (int i = tid; i < M.nq; i += nth) S.q[i] = S.cq[i];
      for (int i = tid; i < M.nm; i += nth) S.q[kMaxParams + i] = S.cq[kMaxParams + i];
      probg = S.dc[kAdNewProbg];
      pdgsum -= S.dc[kAdOldPdg];
      pdgsum += S.dc[kAdNewPdg];
      if (tid == 0) {
        E.cur[p] = (unsigned char)(S.ic[kAcCb] ^ 1);
        E.acc[(size_t)p * 3 + 0]++;
        if (flags & kFlagTopol) E.acc[(size_t)p * 3 + 1]++;
        if (flags & kFlagTmrca) E.acc[(size_t)p * 3 + 2]++;
      }
    }
    block_sync();
  }
  for (int i = tid; i < NI; i += nth) E.all_i[(size_t)c * NI + i] = S.ai[i];
  for (int i = tid; i < ND; i += nth) E.all_d[(size_t)c * ND + i] = S.ad[i];
  for (int i = tid; i < M.nq; i += nth) E.qint[(size_t)c * kMaxParams + i] = S.q[i];
  for (int i = tid; i < M.nm; i += nth) E.mint[(size_t)c * kMaxParams + i] = S.q[kMaxParams + i];
#if IMA_CUDA
  __threadfence_block();
#endif
  block_sync();
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
};

IMA_KERNEL void k_swap(EngineView E, SwapView V) {
  if (ima_block() != 0 || ima_warp_in_block() != 0) return;
  const int N = E.d.nchains_global;
  const int lane = Warp::lane();
  if (N > 1 && V.swaptries > 0) {
    if (lane == 0) {
    Philox rng;
    rng_for(rng, E, 0xffffffffu, kRngSwap);
    for (int x = 0; x < V.swaptries; x++) {
      const int sa = rng.randint(N);
      int sbmin = 0, sbrange = N;
      if (N >= 2 * kSwapDist + 3) {
        sbmin = sa - kSwapDist > 0 ? sa - kSwapDist : 0;
        sbrange = (N < sa + kSwapDist ? N : sa + kSwapDist) - sbmin;
      }
      int sb;
      do { sb = sbmin + rng.randint(sbrange); } while (sb == sa);
      const int ca = V.chain_of_rank[sa], cb = V.chain_of_rank[sb];
      const double w = exp((V.beta_table[sa] - V.beta_table[sb]) * (V.S_global[cb] - V.S_global[ca]));
      V.swap_counts[0]++;
      if (w >= 1.0 || w > rng.uniform()) {
        V.chain_of_rank[sa] = cb; V.chain_of_rank[sb] = ca;
        V.rank_of_chain[ca] = sb; V.rank_of_chain[cb] = sa;
        V.swap_counts[1]++;
      }
    }
    }
#if IMA_CUDA
    __threadfence_block();
#endif
    Warp::sync();
    for (int c = lane; c < E.d.nchains; c += IMA_WARP) E.beta[c] = V.beta_table[V.rank_of_chain[E.d.chain0 + c]];
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

IMA_KERNEL void k_copy_swapsum(EngineView E, double *dst) {
  const int i = ima_block() * kWarpsPerBlock * IMA_WARP + ima_warp_in_block() * IMA_WARP + Warp::lane();
  if (i < E.d.nchains) dst[i] = E.swapsum[i];
}

}  // namespace ima
