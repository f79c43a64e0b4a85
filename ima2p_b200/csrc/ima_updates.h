// The per-step updates that are not genealogy updates (qupdate, ima_main_mpi.cpp:1867-1945):
//
//   k_rescale_t   changet_RY1 (update_t_RY.cpp:222-517), proposal half: one warp per (chain, locus) pair rescales
//                 the node and migration times around the chain's proposed split time, rebuilds the weights and
//                 the likelihood, and leaves the result in the pair's OTHER buffer (as k_propose does)
//   k_accept_t    changet_RY1, decision half: one warp per chain sums the new weights over its loci, integrates
//                 the prior, applies the Hastings ratio of the rescaling and either flips every locus of the
//                 chain to the other buffer or leaves everything as it was
//   k_changeu     changeu / changekappa (update_mc_params.cpp:23-431): one warp per chain walks the mutation-rate
//                 scalars in order; an infinite-sites likelihood under a new scalar is a closed form of the old one
//
// The loci of a chain all depend on its split times, so a split-time update is all-or-nothing per chain; chains are
// independent.  Random streams are keyed by (global chain, step, purpose) like everything else.
#pragma once
#include "ima_kernels.h"

namespace ima {

struct UpdateView {
  double t_max[kMaxPeriods], t_min[kMaxPeriods];   // split-time priors (T[].pr, initialize.cpp)
  double u_win, u_maxratio;                        // changeu: window and reflection bound on the log ratio
  double kappa_win, kappa_max;
  int nurates;
  const int *ul_l, *ul_a;                          // [nurates] scalar j -> (locus, linked part)  (readata.cpp:832-834)
  int *t_counts;                                   // [P][4] edges above / below, migrations above / below the old split time
  double *t_out;                                   // [nchains][4] period, proposed time, MH term, accepted
  double *u_out;                                   // [nchains][4] debug: new pdg of j, new pdg of k, MH term, accepted
  unsigned long long *stats;                       // [4] t tries, t accepts, u tries, u accepts
  const double *t_forced;                          // [nchains] tests: proposed time to use instead of the draw (or null)
  int t_forced_period, t_force_accept;
  // tests: one forced changeu proposal per launch
  int u_forced;                                    // 0: production sweep; 1: evaluate (u_j, u_k, d, kappas) below on u_chain only
  int u_chain, u_j, u_k, u_every;
  double u_d, u_kappa[2];
};

IMA_DEV double ry_beforesplit(int tnode, double oldt, double newt, double tau_u, double ptime) {     // update_t_RY.cpp:67-80
  if (tnode == 0) return ptime * newt / oldt;
  return tau_u + (ptime - tau_u) * (newt - tau_u) / (oldt - tau_u);
}
IMA_DEV double ry_aftersplit(int tnode, int lastperiod, double oldt, double newt, double tau_d, double ptime) {   // :52-65
  if (tnode == lastperiod - 1) return ptime + newt - oldt;
  return tau_d - (tau_d - newt) * (tau_d - ptime) / (tau_d - oldt);
}

struct TProposal { int period; double oldt, newt, t_u, t_d; };

// period pick, getnewt (update_gtree_common.cpp:2501-2519): every warp of a chain derives the same proposal
IMA_DEV TProposal t_proposal(const EngineView &E, const UpdateView &U, const DevModel &M, int c) {
  const double *tv = E.tvals + (size_t)c * kMaxPeriods;
  Philox rng;
  rng_for(rng, E, (uint32_t)(E.d.chain0 + c), kRngSplitTime);
  TProposal t;
  t.period = rng.randint(M.nsplit);
  const double u = rng.uniform();
  if (U.t_forced) t.period = U.t_forced_period;
  t.oldt = tv[t.period];
  t.t_u = t.period == 0 ? 0.0 : tv[t.period - 1];
  t.t_d = t.period == M.nsplit - 1 ? kTimeMax : tv[t.period + 1];
  const double t_d_prior = U.t_max[t.period] < t.t_d ? U.t_max[t.period] : t.t_d;
  const double t_u_prior = U.t_min[t.period] > t.t_u ? U.t_min[t.period] : t.t_u;
  const double twin = (t_d_prior - t_u_prior) / (log((double)E.d.nloci + 1) * (M.npops - t.period));
  double newt = (t.oldt - twin / 2) + u * twin;
  if (newt >= t_d_prior) newt = 2.0 * t_d_prior - newt;
  else if (newt <= t_u_prior) newt = 2.0 * t_u_prior - newt;
  t.newt = U.t_forced ? U.t_forced[c] : newt;
  return t;
}

IMA_KERNEL void IMA_PROPOSE_BOUNDS k_rescale_t(EngineView E, UpdateView U) {
  IMA_SMEM_DECL
  const int p = ima_block() * kWarpsPerBlock + ima_warp_in_block();
  if (p >= E.d.P) return;
  const DevModel &M = IMA_MODEL;
  const int c = p / E.d.nloci, li = p - c * E.d.nloci;
  const DevLocus &L = E.loci[li];
  PairSm S = carve_pair_smem(IMA_SMEM + (size_t)ima_warp_in_block() * pair_smem_bytes(E.d), E.d);
  const int cb = E.cur[p];
  const PairBuf &B = E.buf[cb];
  const PairBuf &Bn = E.buf[cb ^ 1];
  const int lane = Warp::lane();
  const TProposal t = t_proposal(E, U, M, c);
  double tvn[kMaxPeriods];
  for (int k = 0; k < kMaxPeriods; k++) tvn[k] = E.tvals[(size_t)c * kMaxPeriods + k];
  tvn[t.period] = t.newt;
  stage_pair(E, B, p, L.nl, S);
  // :268-328: every time between the neighbouring split times moves with the split time
  int n_eu = 0, n_ed = 0, n_mu = 0, n_md = 0;
  for (int i = lane; i < L.nl; i += IMA_WARP) {
    if (S.down[i] == -1) continue;
    const double x = S.time[i];
    if (x <= t.oldt && x > t.t_u) { S.time[i] = ry_beforesplit(t.period, t.oldt, t.newt, t.t_u, x); n_eu++; }
    else if (x > t.oldt && x < t.t_d) { S.time[i] = ry_aftersplit(t.period, M.nsplit, t.oldt, t.newt, t.t_d, x); n_ed++; }
  }
  const int mignum = S.ctl_i[kCiMignum];
  for (int i = lane; i < mignum; i += IMA_WARP) {
    const double x = S.pt[i];
    if (x <= t.oldt && x > t.t_u) { S.pt[i] = ry_beforesplit(t.period, t.oldt, t.newt, t.t_u, x); n_mu++; }
    else if (x > t.oldt && x < t.t_d) { S.pt[i] = ry_aftersplit(t.period, M.nsplit, t.oldt, t.newt, t.t_d, x); n_md++; }
  }
  if (lane == 0) {
    const double x = S.ctl_d[kCdRoottime];
    if (x <= t.oldt && x > t.t_u) S.ctl_d[kCdRoottime] = ry_beforesplit(t.period, t.oldt, t.newt, t.t_u, x);
    else if (x > t.oldt && x < t.t_d) S.ctl_d[kCdRoottime] = ry_aftersplit(t.period, M.nsplit, t.oldt, t.newt, t.t_d, x);
  }
  n_eu = Warp::sum(n_eu); n_ed = Warp::sum(n_ed); n_mu = Warp::sum(n_mu); n_md = Warp::sum(n_md);
#if IMA_CUDA
  __threadfence_block();
#endif
  Warp::sync();
  bool ok = eval_weights(M, E.d, L, tvn, S);
  const int total_mig = ok ? S.ctl_i[kCiMignum] : 0;
  if (ok && total_mig > E.d.CAP) ok = false;
  uint32_t flags = ok ? (uint32_t)S.ctl_i[kCiFlags] : (uint32_t)kFlagOverflow;
  if (ok) {
    if (has_stepwise(L.model)) {           // the allele states do not move; the branch terms are recomputed below
      const size_t ao = (size_t)p * kMaxLinked * E.d.NL;
      for (int i = lane; i < L.nlinked * E.d.NL; i += IMA_WARP) Bn.A[ao + i] = B.A[ao + i];
#if IMA_CUDA
      __threadfence_block();
#endif
      Warp::sync();
    }
    double pdga[kMaxLinked];
    const double pdg = pair_likelihood(E, L, Bn, p, S, pdga);
    if (pdg == kRejectIS) flags |= kFlagRejectIS;
    if (lane == 0) {
      S.ctl_d[kCdPdg] = pdg;
      if (Bn.pdg_a) for (int ai = 0; ai < L.nlinked; ai++) Bn.pdg_a[(size_t)p * kMaxLinked + ai] = pdga[ai];
    }
    Warp::sync();
    store_pair(E, Bn, p, L.nl, S, total_mig);
  }
  if (lane == 0) {
    E.prop_flags[p] = flags;
    int *o = U.t_counts + (size_t)p * 4;
    o[0] = n_eu; o[1] = n_ed; o[2] = n_mu; o[3] = n_md;
  }
}

IMA_KERNEL void k_accept_t(EngineView E, UpdateView U) {
  IMA_SMEM_DECL
  const int c = ima_block() * kWarpsPerBlock + ima_warp_in_block();
  if (c >= E.d.nchains) return;
  const DevModel &M = IMA_MODEL;
  const int lane = Warp::lane(), NI = E.d.NI, ND = E.d.ND, nloci = E.d.nloci;
  ChainSm S = carve_chain_smem(IMA_SMEM + (size_t)ima_warp_in_block() * chain_smem_bytes(E.d), E.d);
  const TProposal t = t_proposal(E, U, M, c);
  // setzero + sum_treeinfo over loci in locus order (:256, 331), from the proposed (other) buffers
  for (int i = lane; i < NI; i += IMA_WARP) {
    int a = 0;
    for (int li = 0; li < nloci; li++) { const int p = c * nloci + li; a += E.buf[E.cur[p] ^ 1].gwi[(size_t)p * NI + i]; }
    S.ai[i] = a;
  }
  for (int i = lane; i < ND; i += IMA_WARP) {
    double a = 0.0;
    for (int li = 0; li < nloci; li++) { const int p = c * nloci + li; a += E.buf[E.cur[p] ^ 1].gwd[(size_t)p * ND + i]; }
    S.ad[i] = a;
  }
  Warp::sync();
  // integrate_tree_prob (:410): a term whose (c, f) did not change evaluates to the value it had, so the reuse rule
  // of update_gtree_common.cpp:1997-2000 and a fresh evaluation agree
  double probg = 0.0;
  const int nterms = M.nq + (M.nomigration ? 0 : M.nm);
  for (int k = 0; k < nterms; k++) {
    double v;
    if (k < M.nq) {
      int cc; double f, hc;
      gather_q(M, k, S.ai, S.ad, cc, f, hc);
      v = integrate_coalescent_term_coop(E.mc, cc, f, hc, M.q_max[k], M.q_min[k]);
    } else {
      int cm; double f;
      gather_m(M, k - M.nq, S.ai, S.ad, cm, f);
      v = M.expoprior ? integrate_migration_term_expo(E.mc, cm, f, M.m_mean[k - M.nq]) : integrate_migration_term_coop(E.mc, cm, f, M.m_max[k - M.nq], M.m_min[k - M.nq]);
    }
    if (lane == 0) S.q[k] = v;
    probg += v;
  }
  if (!migration_allowed(M, S.ai)) probg = -kMyDblMax;
  double pdgnew = 0.0;
  int n_eu = 0, n_ed = 0, n_mu = 0, n_md = 0;
  uint32_t bad = 0;
  for (int li = lane; li < nloci; li += IMA_WARP) {
    const int p = c * nloci + li;
    pdgnew += E.buf[E.cur[p] ^ 1].sd[(size_t)p * 4 + 3];
    const int *o = U.t_counts + (size_t)p * 4;
    n_eu += o[0]; n_ed += o[1]; n_mu += o[2]; n_md += o[3];
    bad |= E.prop_flags[p] & (kFlagOverflow | kFlagRejectIS | kFlagBadTree);
  }
  pdgnew = Warp::sum(pdgnew);
  n_eu = Warp::sum(n_eu); n_ed = Warp::sum(n_ed); n_mu = Warp::sum(n_mu); n_md = Warp::sum(n_md);
  const bool anybad = Warp::any(bad != 0);
  // every coalescent node was met on both of its daughter edges (:405-408)
  const int ecu = n_eu / 2, ecd = n_ed / 2;
  const double t_u_hterm = (t.newt - t.t_u) / (t.oldt - t.t_u);
  const double t_d_hterm = t.period == M.nsplit - 1 ? 1.0 : (t.t_d - t.newt) / (t.t_d - t.oldt);
  const double beta = E.beta[c];
  double tpw = M.gbeta * (pdgnew - E.pdgsum[c]), mh;
  const double hast = (ecd + n_md) * log(t_d_hterm) + (ecu + n_mu) * log(t_u_hterm);
  if (M.thermo) mh = beta * tpw + (probg - E.probg[c]) + hast;                    // :413-416
  else { tpw += probg - E.probg[c]; mh = beta * tpw + hast; }                     // :419-421
  Philox rng;
  rng_for(rng, E, (uint32_t)(E.d.nchains_global + E.d.chain0 + c), kRngSplitTime);
  const double lu = log(rng.uniform());
  bool accept = !anybad && lu < (mh < 1.0 ? mh : 1.0);                            // :424-425
  if (U.t_forced && U.t_force_accept >= 0) accept = !anybad && U.t_force_accept != 0;
  if (accept) {
    for (int i = lane; i < NI; i += IMA_WARP) E.all_i[(size_t)c * NI + i] = S.ai[i];
    for (int i = lane; i < ND; i += IMA_WARP) E.all_d[(size_t)c * ND + i] = S.ad[i];
    Warp::sync();
    if (lane == 0) {
      for (int k = 0; k < M.nq; k++) E.qint[(size_t)c * kMaxParams + k] = S.q[k];
      for (int k = M.nq; k < nterms; k++) E.mint[(size_t)c * kMaxParams + k - M.nq] = S.q[k];
      E.probg[c] = probg;
      E.pdgsum[c] = pdgnew;
      E.swapsum[c] = M.thermo ? pdgnew : pdgnew + probg;
      E.tvals[(size_t)c * kMaxPeriods + t.period] = t.newt;
    }
    for (int li = lane; li < nloci; li += IMA_WARP) { const int p = c * nloci + li; E.cur[p] ^= 1; }
  }
  if (lane == 0) {
    double *o = U.t_out + (size_t)c * 4;
    o[0] = t.period; o[1] = t.newt; o[2] = mh; o[3] = accept ? 1.0 : 0.0;
#if IMA_CUDA
    atomicAdd(U.stats + 0, 1ull);
    if (accept) atomicAdd(U.stats + 1, 1ull);
#else
    U.stats[0] += 1; if (accept) U.stats[1] += 1;
#endif
  }
}

// current P(D|G) of one linked part: only stepwise loci keep per-part values (pdg_a); elsewhere the part is the locus
IMA_DEV double part_pdg(const DevLocus &L, const PairBuf &B, int p, int ai) {
  return has_stepwise(L.model) ? B.pdg_a[(size_t)p * kMaxLinked + ai] : B.sd[(size_t)p * 4 + 3];
}
IMA_DEV void set_part_pdg(const DevLocus &L, const PairBuf &B, int p, int ai, double v) {
  if (has_stepwise(L.model)) { B.sd[(size_t)p * 4 + 3] += v - B.pdg_a[(size_t)p * kMaxLinked + ai]; B.pdg_a[(size_t)p * kMaxLinked + ai] = v; }
  else B.sd[(size_t)p * 4 + 3] = v;
}

// P(D|G) of linked part `ai` of pair p under a new scalar (and kappa); the genealogy does not move.
//   infinite sites: -length u + sum_s log(ptime_s u) - sumlogk  ->  old - length (u' - u) + S log(u'/u)
//   stepwise / HKY: recomputed from the staged genealogy (S is scratch for one warp)
IMA_DEV double scalar_likelihood(const EngineView &E, const DevModel &M, int c, int li, int ai, double unew, double logratio,
                                 double kappa_new, PairSm &S, const PairBuf &B, const PairBuf &Bscratch) {
  const DevLocus &L = E.loci[li];
  const int p = c * E.d.nloci + li;
  if (has_infinite_sites(L.model) && ai == 0) {
    const double uold = E.uvals[(size_t)p * kMaxLinked];
    return part_pdg(L, B, p, 0) - B.sd[(size_t)p * 4 + 1] * (unew - uold) + L.nsites * logratio;
  }
  stage_pair(E, B, p, L.nl, S);
  if (has_stepwise(L.model)) {
    const size_t ao = ((size_t)p * kMaxLinked + ai) * E.d.NL;
    return likelihood_sw(L, S, B.A + ao, Bscratch.dlikeA + ao, unew);          // new branch terms go to the other buffer
  }
  if (!eval_weights(M, E.d, L, E.tvals + (size_t)c * kMaxPeriods, S)) return kRejectIS;
  return likelihood_hky(E, L, S, p, unew, kappa_new, E.pi + (size_t)p * 4);
}

IMA_DEV double reflect_kappa(double u, double kappa, double win, double kmax) {       // update_mc_params.cpp:258-272
  double nk;
  if (u > 0.5) { nk = kappa + (2.0 * u - 1.0) * win; if (nk > kmax) nk = 2.0 * kmax - nk; }
  else { nk = kappa - win * u * 2.0; if (nk < 0) nk = -nk; }
  return nk;
}

IMA_KERNEL void k_changeu(EngineView E, UpdateView U) {
  IMA_SMEM_DECL
  const int c = ima_block() * kWarpsPerBlock + ima_warp_in_block();
  if (c >= E.d.nchains) return;
  if (U.u_forced ? c != U.u_chain : ((*E.nsteps + 1) % (unsigned long long)U.u_every) != 0) return;   // every UUPDATEINC+1 steps
  const DevModel &M = IMA_MODEL;
  const int lane = Warp::lane(), nloci = E.d.nloci, nur = U.nurates;
  PairSm S = carve_pair_smem(IMA_SMEM + (size_t)ima_warp_in_block() * pair_smem_bytes(E.d), E.d);
  Philox rng;
  rng_for(rng, E, (uint32_t)(E.d.chain0 + c), kRngScalars);
  const double beta = E.beta[c];
  if (nur == 1) {
    // changekappa (:381-431): a single HKY locus has nothing to trade its scalar against
    if (E.loci[0].model != kHKY) return;
    const int p = c * nloci;
    const PairBuf &B = E.buf[E.cur[p]];
    const double nk = U.u_forced ? U.u_kappa[0] : reflect_kappa(rng.uniform(), E.kappa[p], U.kappa_win, U.kappa_max);
    const double newpdg = scalar_likelihood(E, M, c, 0, 0, E.uvals[(size_t)p * kMaxLinked], 0.0, nk, S, B, E.buf[E.cur[p] ^ 1]);
    const double mh = exp(beta * (newpdg - B.sd[(size_t)p * 4 + 3]));
    const double u = rng.uniform();
    const bool accept = !U.u_forced && newpdg != kRejectIS && (mh >= 1.0 || mh > u);
    if (lane == 0) {
      if (accept) {
        const double dl = newpdg - B.sd[(size_t)p * 4 + 3];
        E.pdgsum[c] += dl; E.swapsum[c] += dl;
        B.sd[(size_t)p * 4 + 3] = newpdg;
        E.kappa[p] = nk;
      }
      double *o = U.u_out + (size_t)c * 4;
      o[0] = newpdg; o[1] = 0.0; o[2] = mh; o[3] = accept ? 1.0 : 0.0;
    }
    return;
  }
  const int jn = U.u_forced ? 1 : nur - (nur == 2);
  for (int jj = 0; jj < jn; jj++) {
    int j = jj, k;
    double d, logd;
    if (U.u_forced) { j = U.u_j; k = U.u_k; d = U.u_d; logd = log(d); }
    else {
      if (nur > 2) { do { k = (int)(rng.uniform() * nur); } while (k == j || k < 0 || k >= nur); }      // :78-90
      else k = 1;
    }
    const int lj = U.ul_l[j], aj = U.ul_a[j], lk = U.ul_l[k], ak = U.ul_a[k];
    const int pj = c * nloci + lj, pk = c * nloci + lk;
    const double olduj = E.uvals[(size_t)pj * kMaxLinked + aj], olduk = E.uvals[(size_t)pk * kMaxLinked + ak];
    if (!U.u_forced) {
      // :201-212: uniform step on the log ratio, reflected at +-maxratio; the two scalars move in opposite directions
      const double r = log(olduj / olduk), u = rng.uniform();
      double newr = u > 0.5 ? r + (2.0 * u - 1.0) * U.u_win : r - U.u_win * u * 2.0;
      if (newr > U.u_maxratio) newr = 2.0 * U.u_maxratio - newr;
      else if (newr < -U.u_maxratio) newr = 2.0 * (-U.u_maxratio) - newr;
      logd = (newr - r) / 2;
      d = exp(logd);
    }
    const double newuj = olduj * d, newuk = olduk / d;
    double newpdg[2], newkappa[2] = {0.0, 0.0}, likenewsum = 0.0;
    bool bad = false;
    for (int i = 0; i < 2; i++) {
      const int li = i ? lk : lj, ai = i ? ak : aj, p = i ? pk : pj;
      const PairBuf &B = E.buf[E.cur[p]];
      if (E.loci[li].model == kHKY)
        newkappa[i] = U.u_forced ? U.u_kappa[i] : reflect_kappa(rng.uniform(), E.kappa[p], U.kappa_win, U.kappa_max);
      newpdg[i] = scalar_likelihood(E, M, c, li, ai, i ? newuk : newuj, i ? -logd : logd, newkappa[i], S, B, E.buf[E.cur[p] ^ 1]);
      if (newpdg[i] == kRejectIS || !(newpdg[i] > -DBL_MAX)) bad = true;
      likenewsum += newpdg[i] - part_pdg(E.loci[li], B, p, ai);
      Warp::sync();
    }
    const double mh = exp(beta * M.gbeta * likenewsum);                               // :291
    const double u = rng.uniform();
    const bool accept = !U.u_forced && !bad && u < (mh < 1.0 ? mh : 1.0);             // :294
    if (accept) {
      for (int i = 0; i < 2; i++) {
        const int li = i ? lk : lj, ai = i ? ak : aj, p = i ? pk : pj;
        const DevLocus &L = E.loci[li];
        const PairBuf &B = E.buf[E.cur[p]], &Bo = E.buf[E.cur[p] ^ 1];
        if (has_stepwise(L.model) && ai >= sw_first(L.model)) {
          const size_t ao = ((size_t)p * kMaxLinked + ai) * E.d.NL;
          for (int e = lane; e < L.nl; e += IMA_WARP) B.dlikeA[ao + e] = Bo.dlikeA[ao + e];
        }
        if (lane == 0) {
          E.uvals[(size_t)p * kMaxLinked + ai] = i ? newuk : newuj;
          set_part_pdg(L, B, p, ai, newpdg[i]);
          if (L.model == kHKY) E.kappa[p] = newkappa[i];
        }
      }
      if (lane == 0) { E.pdgsum[c] += likenewsum; E.swapsum[c] += likenewsum; }
    }
#if IMA_CUDA
    __threadfence_block();
#endif
    Warp::sync();
    if (lane == 0) {
      double *o = U.u_out + (size_t)c * 4;
      o[0] = newpdg[0]; o[1] = newpdg[1]; o[2] = mh; o[3] = accept ? 1.0 : 0.0;
#if IMA_CUDA
      atomicAdd(U.stats + 2, 1ull);
      if (accept) atomicAdd(U.stats + 3, 1ull);
#else
      U.stats[2] += 1; if (accept) U.stats[3] += 1;
#endif
    }
  }
}

}  // namespace ima
