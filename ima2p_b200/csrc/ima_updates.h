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
#include "ima_fastpath.h"

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
  unsigned long long *stats;                       // [4] t tries, t accepts, u tries, u accepts (all chains); then of the chain
                                                   // at beta == 1 only: [kMaxPeriods][2 methods: RY, NW][tries, accepts] from
                                                   // kColdTStat, [nurates][tries, accepts] from kColdUStat
  const double *t_forced;                          // [nchains] tests: proposed time to use instead of the draw (or null)
  int t_forced_period, t_force_accept, t_forced_method;
  int t_methods;                                   // bit 0: Rannala-Yang, bit 1: Nielsen-Wakeley
  // tests: one forced changeu proposal per launch
  int u_forced;                                    // 0: production sweep; 1: evaluate (u_j, u_k, d, kappas) below on u_chain only
  int u_chain, u_j, u_k, u_every;
  int u_levels_in_order;                           // tests: k_changeu_levels gives every proposal a level of its own (the walk in order)
  double u_d, u_kappa[2];
};

constexpr int kColdTStat = 4, kColdUStat = 4 + 4 * kMaxPeriods;
IMA_HD size_t update_stats_len(int nurates) { return (size_t)kColdUStat + 2 * (size_t)nurates; }

IMA_DEV double ry_beforesplit(int tnode, double oldt, double newt, double tau_u, double ptime) {     // update_t_RY.cpp:67-80
  if (tnode == 0) return ptime * newt / oldt;
  return tau_u + (ptime - tau_u) * (newt - tau_u) / (oldt - tau_u);
}
IMA_DEV double ry_aftersplit(int tnode, int lastperiod, double oldt, double newt, double tau_d, double ptime) {   // :52-65
  if (tnode == lastperiod - 1) return ptime + newt - oldt;
  return tau_d - (tau_d - newt) * (tau_d - ptime) / (tau_d - oldt);
}

struct TProposal { int period, method; double oldt, newt, t_u, t_d; };     // method 0 = Rannala-Yang, 1 = Nielsen-Wakeley

// period pick, getnewt (update_gtree_common.cpp:2501-2519): every warp of a chain derives the same proposal
IMA_DEV TProposal t_proposal(const EngineView &E, const UpdateView &U, const DevModel &M, int c) {
  const double *tv = E.tvals + (size_t)c * kMaxPeriods;
  Philox rng;
  rng_for(rng, E, (uint32_t)(E.d.chain0 + c), kRngSplitTime);
  TProposal t;
  t.period = rng.randint(M.nsplit);
  const double u = rng.uniform();
  // ima_main_mpi.cpp:1871-1872: one of the two update types at random (NW adds migration events, which a
  // no-migration model must not have: Rannala-Yang only there)
  t.method = (U.t_methods == 3 && !M.nomigration) ? rng.randint(2) : (U.t_methods == 2 && !M.nomigration ? 1 : 0);
  if (U.t_forced) { t.period = U.t_forced_period; t.method = U.t_forced_method; }
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

// ---- Nielsen-Wakeley split-time update: update_t_NW.cpp -----------------------------------------------------
// The split time moves from oldt to newt and nothing else does; every stretch of edge inside the interval
// (tu, td) = (min, max)(oldt, newt) changes the set of populations it may be in, so its migration events there are
// erased and re-simulated.  One lane walks the edges (as the reference does, the work is a chain of small dependent
// decisions); the weights are rebuilt by the whole warp afterwards.

// nowedgepop update_gtree_common.cpp:1638-1655, population tree times from the chain's CURRENT split times
IMA_DEV int nw_nowpop(const DevModel &M, const double *tv, const PairSm &S, int e, double ptime) {
  int pop = S.pop[e];
  const int s0 = S.ms[e], n = S.mcn[e];
  for (int j = 0; j < n && S.pt[s0 + j] < ptime; j++) pop = S.pp[s0 + j];
  while (pop != -1 && ptime > (M.pt_e[pop] == -1 ? kTimeMax : tv[M.pt_e[pop] - 1])) pop = M.pt_down[pop];
  return pop;
}

IMA_DEV double nw_logpf(int code) {            // the seven values logpfpop / logpfpop_r take (:372-583, MIGSIMFRAC 0.999)
  switch (code) {                                 // written out: a logarithm in the code is ~100 instructions per call site
    case 1: return -0.0010005003335835344;          // log(0.999)
    case 2: return -6.907755278982136;              // log(1.0 - 0.999)
    case 3: return -kLog2;
    case 4: return -0.0005002501667917672;          // log(0.999) / 2
    case 5: return -3.453877639491068;              // log(1.0 - 0.999) / 2
    case 6: return -0.34657359027997265470861606073;      // LOG2HALF, imamp.hpp:189
    default: return 0.0;
  }
}

// getmprob_NW :291-330
IMA_DEV double nw_getmprob(const DevModel &M, int period, double mrate, double mtime, int mcount, int uppop, int dpop, int cm2pop, int numpops) {
  if (period == M.nsplit) return 0.0;
  if (period == M.nsplit - 1) return mcount * log(mrate / mtime) - ((mcount & 1) ? mylogsinh(mrate) : mylogcosh(mrate));
  const double logs = uppop == dpop ? log(1 - mrate * exp(-mrate)) : log(1 - exp(-mrate));
  if (mcount == 0) return -mrate - logs;
  if (mcount == 1) return log(mrate / mtime) - mrate - logs;
  const double lognp = log((double)numpops - 1);
  const double logb = cm2pop == dpop ? -lognp : -log((double)numpops - 2);
  return mcount * log(mrate / mtime) + (2 - mcount) * lognp + logb - mrate - logs;
}

// update_mig_tNW :336-786 for the staged genealogy, by the whole warp; every lane returns the log Hastings ratio of the
// migration events (mproposenum - mproposedenom).  The reference walks the edges one after another; nothing in that
// walk couples two edges except (i) an edge whose upper node lies inside the interval takes its upper populations from
// its daughters' record and (ii) the random draws, so here lanes take edges: a first pass decides the lower end of every
// stretch (one draw stream per edge), a second pass re-simulates every stretch against those decisions.
// One stream per (pair, edge, pass): the first pass (lower-end populations) and the second (re-simulated paths) use
// different purpose tags, so no draw of one edge is ever the draw of another edge or pair.
IMA_DEV void nw_edge_rng(Philox &rng, const EngineView &E, int pair_global, int nl_max, int edge, uint32_t purpose) {
  rng_for(rng, E, (uint32_t)(pair_global * nl_max + edge), purpose);
}

IMA_DEV double nw_update_pair(const EngineView &E, const DevModel &M, const double *tv, int period, double oldt, double newt,
                              int ng, int nl, int pair_global, PairSm &S) {
  const int root = S.ctl_i[kCiRoot], lane = Warp::lane();
  const bool up = newt > oldt;                        // the split moves back in time
  const double tu = up ? oldt : newt, td = up ? newt : oldt;
  const int period_a = up ? period : period + 1, period_b = up ? period + 1 : period, p1 = period + 1;
  const int addp = M.addpop[p1], d0 = M.droppops[p1][0], d1 = M.droppops[p1][1];
  int *rec = S.moff;                                  // per edge: db | da << 5 | codef << 10 | coder << 13 | set << 18
  for (int i = lane; i < nl; i += IMA_WARP) rec[i] = 0;
  if (lane == 0) S.ctl_i[kCiNev] = S.pool_free;       // rewritten lists are appended to the scratch part of the pool
#if IMA_CUDA
  __threadfence_block();
#endif
  Warp::sync();
  // The edges with a stretch inside the interval, compacted (ballot + prefix count) so that the lanes of one trip all have
  // work: a trip of the two passes below is a few thousand dependent instructions whatever the number of lanes in it.
  int *list = (int *)S.mask;                          // NL words, not in use before the likelihood
  int nset = 0;
  for (int base = 0; base < nl; base += IMA_WARP) {
    const int i = base + lane;
    bool in = false;
    if (i < nl) { const double uptime = edge_top_time(S, ng, i); in = S.time[i] > tu && uptime <= td; }
    const unsigned m = Warp::ballot(in);
    if (in) list[nset + Warp::popc(m & ((1u << lane) - 1u))] = i;
    nset += Warp::popc(m);
  }
#if IMA_CUDA
  __threadfence_block();
#endif
  Warp::sync();
  // where their lower ends are before (db) and after (da) the update
  for (int k = lane; k < nset; k += IMA_WARP) {
    const int i = list[k];
    const double uptime = edge_top_time(S, ng, i);
    int db, da, cf = 0, cr = 0, sis = -1;
    Philox rng;
    if (S.time[i] > td) {                             // the edge leaves the interval at its lower end: on its own
      db = nw_nowpop(M, tv, S, i, td);
      if (up) {
        if (db == addp) {
          const int c0 = nw_nowpop(M, tv, S, i, tu);
          nw_edge_rng(rng, E, pair_global, E.d.NL, i, kRngSplitMig);
          if (uptime < tu && (c0 == d0 || c0 == d1)) {
            if (rng.uniform() < 0.999) { da = c0; cf = 1; } else { da = c0 == d0 ? d1 : d0; cf = 2; }
          } else { cf = 3; da = rng.bit() ? d1 : d0; }
        } else da = db;
      } else {
        if (db == d0 || db == d1) {
          da = addp;
          const int c0 = nw_nowpop(M, tv, S, i, tu);
          if (uptime < tu && (c0 == d0 || c0 == d1)) cr = c0 == db ? 1 : 2; else cr = 3;
        } else da = db;
      }
    } else {                                          // the edge ends in a coalescence inside the interval: with its sister
      const int dn = S.down[i];
      sis = S.up0[dn] == i ? S.up1[dn] : S.up0[dn];
      if (sis < i) continue;                          // the record of a pair of sisters is made from the lower-numbered one
      const double uptime1 = edge_top_time(S, ng, sis);
      db = nw_nowpop(M, tv, S, i, S.time[i]);
      if (up) {
        if (db == addp) {
          const int c0 = nw_nowpop(M, tv, S, i, tu), c1 = nw_nowpop(M, tv, S, sis, tu);
          nw_edge_rng(rng, E, pair_global, E.d.NL, i, kRngSplitMig);
          if (uptime < tu && uptime1 < tu && c0 == c1 && (c0 == d0 || c0 == d1)) {
            if (rng.uniform() < 0.999) { da = c0; cf = 4; } else { da = c0 == d0 ? d1 : d0; cf = 5; }
          } else { cf = 6; da = rng.bit() ? d1 : d0; }
        } else da = db;
      } else {
        if (db == d0 || db == d1) {
          const int c0 = nw_nowpop(M, tv, S, i, tu), c1 = nw_nowpop(M, tv, S, sis, tu);
          da = addp;
          if (uptime < tu && uptime1 < tu && c0 == c1 && (c0 == d0 || c0 == d1)) cr = c0 == db ? 4 : 5; else cr = 6;
        } else da = db;
      }
    }
    const int v = db | (da << 5) | (cf << 10) | (cr << 13) | (1 << 18);
    rec[i] = v;
    if (sis >= 0) rec[sis] = v;
  }
#if IMA_CUDA
  __threadfence_block();
#endif
  Warp::sync();
  // per stretch: population at the upper end before / after, migration counts and rates, the new path, Hastings terms
  double num = 0.0, denom = 0.0;
  bool overflow = false;
  for (int k = lane; k < nset; k += IMA_WARP) {
    const int ei = list[k];
    if (!(rec[ei] >> 18)) continue;
    const int db = rec[ei] & 31, da = (rec[ei] >> 5) & 31;
    const double logpf = nw_logpf((rec[ei] >> 10) & 7), logpf_r = nw_logpf((rec[ei] >> 13) & 7);
    const double uptime = edge_top_time(S, ng, ei);
    int upb, upa;
    if (uptime < tu) {
      if (!up) { upb = nw_nowpop(M, tv, S, ei, tu); upa = (upb == d0 || upb == d1) ? addp : upb; }
      else { upb = nw_nowpop(M, tv, S, ei, tu * (1 + DBL_EPSILON)); upa = upb == addp ? nw_nowpop(M, tv, S, ei, tu) : upb; }
    } else {                                          // the upper end is a node inside the interval: its daughters' record
      const int r = rec[S.up0[ei]];
      upb = r & 31; upa = (r >> 5) & 31;
      S.pop[ei] = (short)upa;
    }
    if (ei == root) continue;
    const double bottom = td < S.time[ei] ? td : S.time[ei], top = tu > uptime ? tu : uptime;
    const double mtime = bottom - top;
    const int s0 = S.ms[ei], n = S.mcn[ei];
    int kk = 0, mi = 0, mstart = -1;
    while (kk < n && S.pt[s0 + kk] < td) { if (S.pt[s0 + kk] > tu) { if (mi == 0) mstart = kk; mi++; } kk++; }
    int cm2_b = -1, cm2_a = -1;
    if (kk >= 2 && mstart >= 0 && kk - mstart >= 2) cm2_b = kk == 2 ? upb : (int)S.pp[s0 + kk - 3];
    const int mcount = mi, npopsa = M.npops - period_a, npopsb = M.npops - period_b;
    const double mrate = period_a < M.nsplit ? calcmrate(mcount, mtime) * mtime : 0.0;
    Philox rng;
    nw_edge_rng(rng, E, pair_global, E.d.NL, ei, kRngSplitMigSim);       // its own purpose tag: the simulation draws
    int mnew;
    if (npopsa == 1) mnew = 0;
    else if (npopsa == 2) mnew = poisson_cond(rng, mrate, upa == da ? 0 : 1);
    else mnew = poisson_cond(rng, mrate, upa == da ? 3 : 2);
    const double mrate_r = period_b < M.nsplit ? calcmrate(mnew, mtime) * mtime : 0.0;
    // addmigration_NW :208-289: keep the events above tu and below td, replace those in between
    int above = 0;
    if (uptime < tu) while (above < n && S.pt[s0 + above] < tu) above++;
    int below = above;
    while (below < n && S.pt[s0 + below] < td) below++;
    const int numskip = below - above, numheld = n - below;
    if (!up && period_a == M.nsplit) {
      S.mcn[ei] = (unsigned short)above;
    } else if (mnew > 0 || numskip > 0) {
      const int total = above + mnew + numheld;
#if IMA_CUDA
      const int at = atomicAdd(&S.ctl_i[kCiNev], total);
#else
      const int at = S.ctl_i[kCiNev];
      S.ctl_i[kCiNev] += total;
#endif
      if (at + total > S.pool_end) { overflow = true; continue; }
      for (int j = 0; j < above; j++) { S.pt[at + j] = S.pt[s0 + j]; S.pp[at + j] = S.pp[s0 + j]; }
      if (mnew > 0) {
        Emi em; em.seg = at + above; em.nmig = 0;
        simmpath(M, rng, S, em, S.pool_end, period_a, mnew, mtime, top, upa, da);
        if (mnew >= 2) cm2_a = mnew == 2 ? upa : (int)S.pp[at + above + mnew - 3];
      }
      for (int j = 0; j < numheld; j++) { S.pt[at + above + mnew + j] = S.pt[s0 + below + j]; S.pp[at + above + mnew + j] = S.pp[s0 + below + j]; }
      S.ms[ei] = (unsigned short)at; S.mcn[ei] = (unsigned short)total;
    }
    // forward then reverse, one copy of getmprob_NW's code (a rolled loop of two)
#if IMA_CUDA
#pragma unroll 1
#endif
    for (int h = 0; h < 2; h++) {
      const double mr = h ? mrate_r : mrate;
      if (mr > 0) {
        const double v = (h ? logpf_r : logpf) + nw_getmprob(M, h ? period_b : period_a, mr, mtime, h ? mcount : mnew, h ? upb : upa, h ? db : da, h ? cm2_b : cm2_a, h ? npopsb : npopsa);
        if (h) num += v; else denom += v;
      }
    }
  }
  if (Warp::any(overflow) && lane == 0) S.ctl_i[kCiFlags] |= (int)kFlagOverflow;
  return Warp::sum(num - denom);
}

// evcap / migcap: what the caller's tables hold (the general kernel: EVP events, CAP migration events; the fast one: FEV, FC).
// Returns false when the pair did not fit: the caller decides whether that is a dropped proposal or a pair for the general path.
// One body for both update types, so that the staging, the weights and the store exist once in a kernel's code (the kernels
// that call this run their code once per warp: instruction fetch is what they wait for most).
IMA_DEV bool split_t_pair(const EngineView &E, const UpdateView &U, const DevModel &M, const TProposal &t, int p, int c, int li, PairSm &S, int evcap, int migcap) {
  const DevLocus &L = E.loci[li];
  const int cb = E.cur[p];
  const PairBuf &B = E.buf[cb];
  const PairBuf &Bn = E.buf[cb ^ 1];
  const int lane = Warp::lane();
  const bool nw = t.method == 1;
  double tvn[kMaxPeriods];
  for (int k = 0; k < kMaxPeriods; k++) tvn[k] = E.tvals[(size_t)c * kMaxPeriods + k];
  IMA_PROF_DECL(8)
  stage_pair(E, B, p, L.nl, S);
  IMA_PROF_MARK()
  int n_eu = 0, n_ed = 0, n_mu = 0, n_md = 0;
  if (nw) {
    const double roottime = S.ctl_d[kCdRoottime];
    // update_t_NW.cpp:967-969: a genealogy whose root is younger than both split times is not touched
    const bool touched = (t.newt > t.oldt && roottime > t.oldt) || (t.newt < t.oldt && roottime > t.newt);
    double mw = 0.0;
    if (touched) mw = nw_update_pair(E, M, tvn, t.period, t.oldt, t.newt, L.ng, L.nl, (E.d.chain0 + c) * E.d.nloci + li, S);   // tvn: still the old times
    if (lane == 0) S.ctl_d[kCdMigw] = mw;
  } else {
    // update_t_RY.cpp:268-328: every time between the neighbouring split times moves with the split time
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
    // the four counts in one reduction (each is at most 2 NL + CAP: 16 bits apiece)
    unsigned long long packed = (unsigned long long)n_eu | ((unsigned long long)n_ed << 16) | ((unsigned long long)n_mu << 32) | ((unsigned long long)n_md << 48);
    packed = Warp::sum(packed);
    n_eu = (int)(packed & 0xffffull); n_ed = (int)((packed >> 16) & 0xffffull); n_mu = (int)((packed >> 32) & 0xffffull); n_md = (int)(packed >> 48);
  }
  tvn[t.period] = t.newt;
#if IMA_CUDA
  __threadfence_block();
#endif
  Warp::sync();
  IMA_PROF_MARK()
  bool ok = !(S.ctl_i[kCiFlags] & kFlagOverflow);
  if (ok) ok = eval_weights(M, E.d, L, tvn, S, evcap);
  IMA_PROF_MARK()
  const int total_mig = ok ? S.ctl_i[kCiMignum] : 0;
  if (ok && total_mig > migcap) ok = false;
  uint32_t flags = ok ? (nw ? 0u : (uint32_t)S.ctl_i[kCiFlags]) : (uint32_t)kFlagOverflow;
  if (ok) {
    if (nw) {
      // branch lengths do not change: P(D|G) and everything it is built from are carried over (update_t_NW.cpp:908)
      if (has_stepwise(L.model)) {
        const size_t ao = (size_t)p * kMaxLinked * E.d.NL;
        for (int i = lane; i < L.nlinked * E.d.NL; i += IMA_WARP) { Bn.A[ao + i] = B.A[ao + i]; Bn.dlikeA[ao + i] = B.dlikeA[ao + i]; }
        for (int ai = lane; ai < L.nlinked; ai += IMA_WARP) Bn.pdg_a[(size_t)p * kMaxLinked + ai] = B.pdg_a[(size_t)p * kMaxLinked + ai];
      }
      if (L.model == kHKY)                                  // and so are the stored partials: the same slots stay current
        for (int w = lane; w < E.d.hky_mask_words; w += IMA_WARP) Bn.hky_mask[(size_t)p * E.d.hky_mask_words + w] = B.hky_mask[(size_t)p * E.d.hky_mask_words + w];
      if (lane == 0) S.ctl_d[kCdPdg] = B.sd[(size_t)p * 4 + 3];
    } else {
      if (has_stepwise(L.model)) {           // the allele states do not move; the branch terms are recomputed below
        const size_t ao = (size_t)p * kMaxLinked * E.d.NL;
        for (int i = lane; i < L.nlinked * E.d.NL; i += IMA_WARP) Bn.A[ao + i] = B.A[ao + i];
#if IMA_CUDA
        __threadfence_block();
#endif
        Warp::sync();
      }
      double pdga[kMaxLinked];
      HkyCall hk; hk.mode = kHkyFull; hk.freed = hk.olddd = -1;       // every time moved: every node is recomputed
      hk.mask_cur = B.hky_mask ? B.hky_mask + (size_t)p * E.d.hky_mask_words : nullptr;
      hk.mask_new = Bn.hky_mask ? Bn.hky_mask + (size_t)p * E.d.hky_mask_words : nullptr;
      const double pdg = pair_likelihood(E, L, Bn, p, S, pdga, hk);
      if (pdg == kRejectIS) flags |= kFlagRejectIS;
      if (lane == 0) {
        S.ctl_d[kCdPdg] = pdg;
        if (Bn.pdg_a) for (int ai = 0; ai < L.nlinked; ai++) Bn.pdg_a[(size_t)p * kMaxLinked + ai] = pdga[ai];
      }
    }
    Warp::sync();
    IMA_PROF_MARK()
    store_pair(E, Bn, p, L.nl, S, total_mig);
    IMA_PROF_MARK()
#if defined(IMA_PROF) && IMA_CUDA
    if (lane == 0 && p % 641 == 0 && (current_step(E) % 64) == 0)
      printf("PROFS %s p %d stage %lld update %lld weights %lld like %lld store %lld\n", nw ? "nw" : "ry", p, prof_t_[1] - prof_t_[0], prof_t_[2] - prof_t_[1],
             prof_t_[3] - prof_t_[2], prof_t_[4] - prof_t_[3], prof_t_[5] - prof_t_[4]);
#endif
  }
  if (lane == 0) {
    E.prop_flags[p] = flags;
    if (nw) {
      E.prop_extra[p] = S.ctl_d[kCdMigw];
      if (E.prop_dbg) E.prop_dbg[(size_t)p * 4 + 0] = S.ctl_d[kCdMigw];
    }
    int *o = U.t_counts + (size_t)p * 4;
    o[0] = n_eu; o[1] = n_ed; o[2] = n_mu; o[3] = n_md;
  }
  return ok;
}

// One launch for both split-time updates: every chain drew its update type (t_proposal), every warp follows its chain.
IMA_KERNEL void IMA_PROPOSE_BOUNDS k_split_t(EngineView E, UpdateView U) {
  IMA_SMEM_DECL
  const int idx = ima_block() * kWarpsPerBlock + ima_warp_in_block();
  if (idx >= E.c_n * E.d.nloci) return;
  const DevModel &M = IMA_MODEL;
  const int c = E.c_lo + idx / E.d.nloci, li = idx % E.d.nloci, p = c * E.d.nloci + li;
  PairSm S = carve_pair_smem(IMA_SMEM + (size_t)ima_warp_in_block() * pair_smem_bytes(E.d), E.d);
  const TProposal t = t_proposal(E, U, M, c);
  split_t_pair(E, U, M, t, p, c, li, S, E.d.EVP, E.d.CAP);
}

// ---- the same proposals with the small tables of the fast path (ima_fastpath.h): FC migration events per genealogy plus FS
// entries of scratch for the lists Nielsen-Wakeley rewrites.  A pair that does not fit goes to the redo list and k_split_t_redo
// makes its proposal with the general tables, from the same random streams.
IMA_HD size_t split_smem_bytes(const EngineDims &d) { return weigh_smem_bytes(d, d.FC + d.FS); }
IMA_DEV PairSm carve_split_smem(unsigned char *base, const EngineDims &d) { return carve_weigh_smem(base, d, d.FC + d.FS); }
#ifndef IMA_SPLIT_MINBLOCKS
#define IMA_SPLIT_MINBLOCKS 6
#endif
#if IMA_CUDA
#define IMA_SPLIT_BOUNDS __launch_bounds__(kWarpsPerBlock * 32, IMA_SPLIT_MINBLOCKS)
#else
#define IMA_SPLIT_BOUNDS
#endif
IMA_KERNEL void IMA_SPLIT_BOUNDS k_split_t_fast(EngineView E, UpdateView U) {
  IMA_SMEM_DECL
  const int idx = ima_block() * kWarpsPerBlock + ima_warp_in_block();
  if (idx == 0 && Warp::lane() == 0) *redo_counter(E, 1, (int)((current_step(E) + 1) & 1ull)) = 0;
  if (idx >= E.c_n * E.d.nloci) return;
  const DevModel &M = IMA_MODEL;
  const int c = E.c_lo + idx / E.d.nloci, li = idx % E.d.nloci, p = c * E.d.nloci + li;
  PairSm S = carve_split_smem(IMA_SMEM + (size_t)ima_warp_in_block() * split_smem_bytes(E.d), E.d);
  const TProposal t = t_proposal(E, U, M, c);
  bool ok = E.buf[E.cur[p]].si[(size_t)p * 2 + 1] <= E.d.FC;
  if (ok) ok = split_t_pair(E, U, M, t, p, c, li, S, E.d.FEV, E.d.FC);
  if (!ok && Warp::lane() == 0) { E.prop_flags[p] = kFlagRedo; redo_push(E, 1, p); }
}
IMA_KERNEL void IMA_PROPOSE_BOUNDS k_split_t_redo(EngineView E, UpdateView U) {
  IMA_SMEM_DECL
  const DevModel &M = IMA_MODEL;
  const int n = *redo_counter(E, 1, (int)(current_step(E) & 1ull));
  const int *list = redo_list(E, 1);
  PairSm S = carve_pair_smem(IMA_SMEM + (size_t)ima_warp_in_block() * pair_smem_bytes(E.d), E.d);
  for (int k = ima_block() * kWarpsPerBlock + ima_warp_in_block(); k < n; k += E.redo_grid * kWarpsPerBlock) {
    const int p = list[k], c = p / E.d.nloci, li = p - c * E.d.nloci;
    const TProposal t = t_proposal(E, U, M, c);
    split_t_pair(E, U, M, t, p, c, li, S, E.d.EVP, E.d.CAP);
    Warp::sync();
  }
}

constexpr int kTWarps = 8;            // warps of a k_accept_t block (one block per chain)
constexpr int kAcceptTStageBudget = 96 * 1024;
// staging of the chain's proposed weight records and per-locus scalars: [nloci] other-buffer index, [ND][nloci] doubles,
// [NI][nloci] ints, [nloci] new pdg, [nloci] migration term, [nloci][4] counts, [nloci] flags
IMA_HD size_t accept_t_stage_bytes(const EngineDims &d) {
  return align8((size_t)d.nloci * 4) + (size_t)d.nloci * (8 * (size_t)d.ND + 4 * (size_t)d.NI + 8 + 8 + 16 + 4) + 64;
}
IMA_HD bool accept_t_staged(const EngineDims &d) { return accept_t_stage_bytes(d) <= (size_t)kAcceptTStageBudget; }
IMA_HD size_t accept_t_smem_bytes(const EngineDims &d) { return chain_smem_bytes(d) + (accept_t_staged(d) ? accept_t_stage_bytes(d) : 0); }

IMA_KERNEL void k_accept_t(EngineView E, UpdateView U) {
  IMA_SMEM_DECL
  if (ima_block() >= E.c_n) return;
  const int c = E.c_lo + ima_block();
  const DevModel &M = IMA_MODEL;
  const int lane = Warp::lane(), NI = E.d.NI, ND = E.d.ND, nloci = E.d.nloci;
  ChainSm S = carve_chain_smem(IMA_SMEM, E.d);
#if defined(IMA_PROF) && IMA_CUDA
  auto pclk_ = []() { long long x; asm volatile("mov.u64 %0, %%clock64;" : "=l"(x) :: "memory"); return x; };
  long long pq_[8]; int pn_ = 0; pq_[pn_++] = pclk_();
#define IMA_PT() pq_[pn_++] = pclk_();
#else
#define IMA_PT()
#endif
  const int nterms = M.nq + (M.nomigration ? 0 : M.nm);
  // When the chain's records fit, everything this kernel reads from global memory is brought into shared memory: which buffer
  // every locus proposes into first, then every weight and scalar with kLoadTrips independent loads in flight per thread.  The
  // chain's scalars, the proposal and the uniform -- nothing of which depends on the records -- are computed while they travel.
  const bool staged = accept_t_staged(E.d);
  unsigned char *sp = IMA_SMEM + chain_smem_bytes(E.d);
  int *s_ob = (int *)sp; sp += align8((size_t)nloci * 4);
  double *s_wd = (double *)sp; sp += (size_t)nloci * 8 * ND;
  double *s_pdg = (double *)sp; sp += (size_t)nloci * 8;
  double *s_mw = (double *)sp; sp += (size_t)nloci * 8;
  int *s_wi = (int *)sp; sp += (size_t)nloci * 4 * NI;
  int *s_cnt = (int *)sp; sp += (size_t)nloci * 16;
  int *s_fl = (int *)sp;
  if (staged) {
    IMA_FOR_WARPS(w, kTWarps) {
      const int tid = w * IMA_WARP + lane, nth = kTWarps * IMA_WARP;
      for (int li = tid; li < nloci; li += nth) s_ob[li] = E.cur[c * nloci + li] ^ 1;
    }
  }
  const double beta = E.beta[c], pdgsum_old = E.pdgsum[c], probg_old = E.probg[c], swapsum_old = E.swapsum[c];
  const TProposal t = t_proposal(E, U, M, c);
  Philox rng;
  rng_for(rng, E, (uint32_t)(E.d.nchains_global + E.d.chain0 + c), kRngSplitTime);
  const double uacc = rng.uniform();
  const double t_u_hterm = (t.newt - t.t_u) / (t.oldt - t.t_u);
  const double t_d_hterm = t.period == M.nsplit - 1 ? 1.0 : (t.t_d - t.newt) / (t.t_d - t.oldt);
  double log_u_h, log_d_h, log_uacc;
  log3_coop(t_u_hterm, t_d_hterm, uacc, log_u_h, log_d_h, log_uacc);
  IMA_PT()
  if (staged) {
    block_sync();
    constexpr int kLoadTrips = 4;
    IMA_FOR_WARPS(w, kTWarps) {
      const int tid = w * IMA_WARP + lane, nth = kTWarps * IMA_WARP, per = NI + ND, n = nloci * per;
      for (int k0 = tid; k0 < n; k0 += kLoadTrips * nth) {
        long long raw[kLoadTrips];
#if IMA_CUDA
#pragma unroll
#endif
        for (int u = 0; u < kLoadTrips; u++) {
          const int k = k0 + u * nth;
          if (k < n) {
            const int li = k / per, i = k - li * per, p = c * nloci + li;
            const PairBuf &N = E.buf[s_ob[li]];
            raw[u] = i < NI ? (long long)N.gwi[(size_t)p * NI + i] : dbl_bits(N.gwd[(size_t)p * ND + i - NI]);
          }
        }
#if IMA_CUDA
#pragma unroll
#endif
        for (int u = 0; u < kLoadTrips; u++) {
          const int k = k0 + u * nth;
          if (k < n) {
            const int li = k / per, i = k - li * per;
            if (i < NI) s_wi[i * nloci + li] = (int)raw[u];
            else s_wd[(i - NI) * nloci + li] = bits_dbl(raw[u]);
          }
        }
      }
      for (int li = tid; li < nloci; li += nth) {
        const int p = c * nloci + li;
        const PairBuf &N = E.buf[s_ob[li]];
        const double pdg = N.sd[(size_t)p * 4 + 3], mw = E.prop_extra[p];
        const int *o = U.t_counts + (size_t)p * 4;
        const int o0 = o[0], o1 = o[1], o2 = o[2], o3 = o[3];
        const int fl = (int)(E.prop_flags[p] & (kFlagOverflow | kFlagRejectIS | kFlagBadTree));
        s_pdg[li] = pdg; s_mw[li] = mw;
        s_cnt[li * 4] = o0; s_cnt[li * 4 + 1] = o1; s_cnt[li * 4 + 2] = o2; s_cnt[li * 4 + 3] = o3;
        s_fl[li] = fl;
      }
    }
    block_sync();
  }
  IMA_PT()
  // setzero + sum_treeinfo over the loci (:256, 331), from the proposed (other) buffers: warps take weights, lanes take
  // loci, one warp reduction per weight (the loads of one weight do not wait for each other).  The per-locus scalars the
  // decision needs (new P(D|G), migration term, counts, flags) are three more "weights": S.dc[0..2], S.ic[0..4]
  IMA_FOR_WARPS(w, kTWarps) {
    for (int i = w; i < NI + ND + 3; i += kTWarps) {
      if (i < NI) {
        int a = 0;
        if (staged) for (int li = lane; li < nloci; li += IMA_WARP) a += s_wi[i * nloci + li];
        else for (int li = lane; li < nloci; li += IMA_WARP) { const int p = c * nloci + li; a += E.buf[E.cur[p] ^ 1].gwi[(size_t)p * NI + i]; }
        a = Warp::sum(a);
        if (lane == 0) S.ai[i] = a;
      } else if (i < NI + ND) {
        double a = 0.0;
        if (staged) for (int li = lane; li < nloci; li += IMA_WARP) a += s_wd[(i - NI) * nloci + li];
        else for (int li = lane; li < nloci; li += IMA_WARP) { const int p = c * nloci + li; a += E.buf[E.cur[p] ^ 1].gwd[(size_t)p * ND + i - NI]; }
        a = Warp::sum(a);
        if (lane == 0) S.ad[i - NI] = a;
      } else if (i == NI + ND) {                                  // new P(D|G) and the migration term
        double pdgnew = 0.0, migw = 0.0;
        for (int li = lane; li < nloci; li += IMA_WARP) {
          if (staged) { pdgnew += s_pdg[li]; if (t.method == 1) migw += s_mw[li]; }
          else {
            const int p = c * nloci + li;
            pdgnew += E.buf[E.cur[p] ^ 1].sd[(size_t)p * 4 + 3];
            if (t.method == 1) migw += E.prop_extra[p];
          }
        }
        Warp::sum2(pdgnew, migw);
        if (lane == 0) { S.dc[0] = pdgnew; S.dc[1] = migw; }
      } else if (i == NI + ND + 1) {                              // the four counts (each at most NL or CAP per locus)
        long long a01 = 0, a23 = 0;
        for (int li = lane; li < nloci; li += IMA_WARP) {
          const int *o = staged ? s_cnt + li * 4 : U.t_counts + (size_t)(c * nloci + li) * 4;
          a01 += (long long)o[0] | ((long long)o[1] << 32);
          a23 += (long long)o[2] | ((long long)o[3] << 32);
        }
        a01 = (long long)Warp::sum((unsigned long long)a01); a23 = (long long)Warp::sum((unsigned long long)a23);
        if (lane == 0) { S.ic[0] = (int)(a01 & 0xffffffffll); S.ic[1] = (int)(a01 >> 32); S.ic[2] = (int)(a23 & 0xffffffffll); S.ic[3] = (int)(a23 >> 32); }
      } else {
        uint32_t bad = 0;
        for (int li = lane; li < nloci; li += IMA_WARP)
          bad |= staged ? (uint32_t)s_fl[li] : (E.prop_flags[c * nloci + li] & (kFlagOverflow | kFlagRejectIS | kFlagBadTree));
        const bool anyb = Warp::any(bad != 0);
        if (lane == 0) S.ic[4] = anyb ? 1 : 0;
      }
    }
  }
  block_sync();
  IMA_PT()
  // integrate_tree_prob (:410): a term whose (c, f) did not change evaluates to the value it had, so the reuse rule
  // of update_gtree_common.cpp:1997-2000 and a fresh evaluation agree.  One warp per term.
  IMA_FOR_WARPS(w, kTWarps) {
    for (int k = w; k < nterms; k += kTWarps) {
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
    }
  }
  block_sync();
  IMA_PT()
  if (ima_warp_in_block() != 0) return;              // the decision and the commit are one warp's work
  double probg = 0.0;
  for (int k = 0; k < nterms; k++) probg += S.q[k];
  if (!migration_allowed(M, S.ai)) probg = -kMyDblMax;
  double pdgnew = S.dc[0];
  const double migw = S.dc[1];
  const int n_eu = S.ic[0], n_ed = S.ic[1], n_mu = S.ic[2], n_md = S.ic[3];
  const bool anybad = S.ic[4] != 0;
  // every coalescent node was met on both of its daughter edges (:405-408)
  const int ecu = n_eu / 2, ecd = n_ed / 2;
  double tpw = M.gbeta * (pdgnew - pdgsum_old), mh;
  const double hast = (ecd + n_md) * log_d_h + (ecu + n_mu) * log_u_h;
  if (M.thermo) mh = beta * tpw + (probg - probg_old) + hast;                    // :413-416
  else { tpw += probg - probg_old; mh = beta * tpw + hast; }                     // :419-421
  bool accept;
  if (t.method == 1) {
    // changet_NW update_t_NW.cpp:993-1005: no likelihood term (branch lengths are unchanged), the migration events'
    // Hastings ratio instead; the decision is taken on the natural scale
    const double dprobg = probg - probg_old;
    mh = exp((M.thermo ? dprobg : beta * dprobg) + migw);
    pdgnew = pdgsum_old;
    accept = !anybad && uacc < (mh < 1.0 ? mh : 1.0);
  } else {
    accept = !anybad && log_uacc < (mh < 1.0 ? mh : 1.0);                        // update_t_RY.cpp:424-425
  }
  if (U.t_forced && U.t_force_accept >= 0) accept = !anybad && U.t_force_accept != 0;
  if (E.xch.publisher == 2) publish_chain(E, c, accept ? (M.thermo ? pdgnew : pdgnew + probg) : swapsum_old);
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
    // the flip: the other buffer's index is at hand for a staged chain (no read-modify-write of global memory)
    if (staged) for (int li = lane; li < nloci; li += IMA_WARP) E.cur[c * nloci + li] = (unsigned char)s_ob[li];
    else for (int li = lane; li < nloci; li += IMA_WARP) { const int p = c * nloci + li; E.cur[p] ^= 1; }
  }
  IMA_PT()
#if defined(IMA_PROF) && IMA_CUDA
  if (lane == 0 && (c == 3 || c == 77) && (current_step(E) % 64) == 0)
    printf("PROFT chain %d method %d early %lld load %lld sums %lld terms %lld decide+commit %lld\n", c, t.method, pq_[1] - pq_[0], pq_[2] - pq_[1],
           pq_[3] - pq_[2], pq_[4] - pq_[3], pq_[5] - pq_[4]);
#endif
  if (lane == 0) {
    double *o = U.t_out + (size_t)c * 4;
    o[0] = t.period; o[1] = t.newt; o[2] = mh; o[3] = accept ? 1.0 : 0.0;
#if IMA_CUDA
    atomicAdd(U.stats + 0, 1ull);
    if (accept) atomicAdd(U.stats + 1, 1ull);
#else
    U.stats[0] += 1; if (accept) U.stats[1] += 1;
#endif
    if (beta == 1.0 && !U.t_forced) {
      unsigned long long *cs = U.stats + kColdTStat + ((size_t)t.period * 2 + (t.method == 1 ? 1 : 0)) * 2;
      stat_add(cs, 1ull);
      if (accept) stat_add(cs + 1, 1ull);
    }
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
  // the trial's partials go to the other slots, its slot mask to the other buffer's mask: k_changeu adopts it on acceptance
  HkyCall hk; hk.mode = kHkyFull; hk.freed = hk.olddd = -1;
  hk.mask_cur = B.hky_mask + (size_t)p * E.d.hky_mask_words; hk.mask_new = Bscratch.hky_mask + (size_t)p * E.d.hky_mask_words;
  return likelihood_hky(E, L, S, p, unew, kappa_new, E.pi + (size_t)p * 4, hk);
}

IMA_DEV double reflect_kappa(double u, double kappa, double win, double kmax) {       // update_mc_params.cpp:258-272
  double nk;
  if (u > 0.5) { nk = kappa + (2.0 * u - 1.0) * win; if (nk > kmax) nk = 2.0 * kmax - nk; }
  else { nk = kappa - win * u * 2.0; if (nk < 0) nk = -nk; }
  return nk;
}

// shared memory of the all-infinite-sites fast path of k_changeu, per warp: u, pdg, length, step draw, log accept draw
// (doubles) and the partner (int) of every locus
IMA_HD size_t changeu_smem_doubles(int nloci) { return (size_t)5 * nloci + (nloci + 1) / 2 + 1; }

IMA_DEV void changeu_chain(const EngineView &E, const UpdateView &U, int c, PairSm &S);
IMA_KERNEL void k_changeu(EngineView E, UpdateView U) {
  IMA_SMEM_DECL
  if (ima_block() * kWarpsPerBlock + ima_warp_in_block() >= E.c_n) return;
  const int c = E.c_lo + ima_block() * kWarpsPerBlock + ima_warp_in_block();
  if (U.u_forced ? c == U.u_chain : ((current_step(E) + 1) % (unsigned long long)U.u_every) == 0) {         // every UUPDATEINC+1 steps
    PairSm S = carve_pair_smem(IMA_SMEM + (size_t)ima_warp_in_block() * pair_smem_bytes(E.d), E.d);
    changeu_chain(E, U, c, S);
  }
  if (E.xch.publisher == 3) {                          // the chain's step ends here, whether or not it was the scalars' turn
#if IMA_CUDA
    __threadfence_block();
    __syncwarp();
#endif
    publish_chain(E, c, *(volatile double *)(E.swapsum + c));
  }
}
IMA_DEV void changeu_chain(const EngineView &E, const UpdateView &U, int c, PairSm &S) {
  IMA_SMEM_DECL
  const DevModel &M = IMA_MODEL;
  const int lane = Warp::lane(), nloci = E.d.nloci, nur = U.nurates;
  Philox rng;
  rng_for(rng, E, (uint32_t)(E.d.chain0 + c), kRngScalars);
  const double beta = E.beta[c];
  if (!U.u_forced && nur > 2 && nur == nloci && !E.d.any_sw && !E.d.any_hky) {
    // All loci infinite sites, one scalar each (every shipped input): a proposal touches two numbers per locus and costs
    // a logarithm and two exponentials, so the walk is bound by the latency of fetching them.  The chain's scalars,
    // likelihoods and tree lengths are staged in shared memory by the warp, every proposal's draws (partner, step,
    // accept) are made by the lane of that proposal from its own stream, one lane then walks the proposals in order
    // on shared memory, and the warp writes the result back.
    double *su = (double *)IMA_SMEM + (size_t)ima_warp_in_block() * changeu_smem_doubles(nloci);
    double *spdg = su + nloci, *slen = spdg + nloci, *sdraw = slen + nloci, *slacc = sdraw + nloci;
    int *sk = (int *)(slacc + nloci);
    for (int li = lane; li < nloci; li += IMA_WARP) {
      const int p = c * nloci + li;
      const PairBuf &B = E.buf[E.cur[p]];
      su[li] = E.uvals[(size_t)p * kMaxLinked];
      spdg[li] = B.sd[(size_t)p * 4 + 3];
      slen[li] = B.sd[(size_t)p * 4 + 1];
      Philox r2;
      const unsigned long long step = current_step(E);
      r2.init(E.seed, (uint32_t)((E.d.chain0 + c) * nloci + li), (uint32_t)step, kRngScalars | ((uint32_t)(step >> 32) << 8));
      int k;
      do { k = (int)(r2.uniform() * nur); } while (k == li || k < 0 || k >= nur);          // :78-90
      sk[li] = k;
      sdraw[li] = r2.uniform();
      slacc[li] = log(r2.uniform());
    }
#if IMA_CUDA
    __threadfence_block();
#endif
    Warp::sync();
    if (lane == 0) {
      double total = 0.0;
      unsigned long long nacc = 0;
      for (int j = 0; j < nur; j++) {
        const int k = sk[j];
        const double olduj = su[j], olduk = su[k], r = log(olduj / olduk), u = sdraw[j];
        double newr = u > 0.5 ? r + (2.0 * u - 1.0) * U.u_win : r - U.u_win * u * 2.0;      // :201-212
        if (newr > U.u_maxratio) newr = 2.0 * U.u_maxratio - newr;
        else if (newr < -U.u_maxratio) newr = 2.0 * (-U.u_maxratio) - newr;
        const double logd = (newr - r) / 2, d = exp(logd);
        const double newuj = olduj * d, newuk = olduk / d;
        const double npj = spdg[j] - slen[j] * (newuj - olduj) + E.loci[j].nsites * logd;
        const double npk = spdg[k] - slen[k] * (newuk - olduk) - E.loci[k].nsites * logd;
        const double likenewsum = (npj - spdg[j]) + (npk - spdg[k]);
        const double x = beta * M.gbeta * likenewsum;                                      // log of :291
        if (slacc[j] < (x < 0.0 ? x : 0.0)) {                                              // U < min(1, e^x), :294
          su[j] = newuj; su[k] = newuk; spdg[j] = npj; spdg[k] = npk;
          total += likenewsum;
          nacc++;
          slacc[j] = 1.0;                                    // its draw has been used: the slot now says "accepted" (a log draw is <= 0)
        }
      }
      E.pdgsum[c] += total; E.swapsum[c] += total;
#if IMA_CUDA
      atomicAdd(U.stats + 2, (unsigned long long)nur);
      atomicAdd(U.stats + 3, nacc);
#else
      U.stats[2] += nur; U.stats[3] += nacc;
#endif
    }
#if IMA_CUDA
    __threadfence_block();
#endif
    Warp::sync();
    for (int li = lane; li < nloci; li += IMA_WARP) {
      const int p = c * nloci + li;
      E.uvals[(size_t)p * kMaxLinked] = su[li];
      E.buf[E.cur[p]].sd[(size_t)p * 4 + 3] = spdg[li];
      if (beta == 1.0) {                               // a proposal counts for the scalar and for its partner (ima_main_mpi.cpp:1926-1935)
        const unsigned long long a = slacc[li] == 1.0 ? 1ull : 0ull;
        stat_add(U.stats + kColdUStat + 2 * (size_t)li, 1ull);
        stat_add(U.stats + kColdUStat + 2 * (size_t)sk[li], 1ull);
        if (a) { stat_add(U.stats + kColdUStat + 2 * (size_t)li + 1, 1ull); stat_add(U.stats + kColdUStat + 2 * (size_t)sk[li] + 1, 1ull); }
      }
    }
    return;
  }
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
        for (int w = 0; w < E.d.hky_mask_words; w++)        // the trial's partials become the current ones
          B.hky_mask[(size_t)p * E.d.hky_mask_words + w] = E.buf[E.cur[p] ^ 1].hky_mask[(size_t)p * E.d.hky_mask_words + w];
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
          if (L.model == kHKY) {
            E.kappa[p] = newkappa[i];
            for (int w = 0; w < E.d.hky_mask_words; w++) B.hky_mask[(size_t)p * E.d.hky_mask_words + w] = Bo.hky_mask[(size_t)p * E.d.hky_mask_words + w];
          }
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
      if (beta == 1.0 && !U.u_forced) {                  // counted for the scalar and for its partner (ima_main_mpi.cpp:1926-1935)
        stat_add(U.stats + kColdUStat + 2 * (size_t)j, 1ull);
        stat_add(U.stats + kColdUStat + 2 * (size_t)k, 1ull);
        if (accept) { stat_add(U.stats + kColdUStat + 2 * (size_t)j + 1, 1ull); stat_add(U.stats + kColdUStat + 2 * (size_t)k + 1, 1ull); }
      }
    }
  }
}


// ---- the mutation-scalar walk for data whose likelihood must be recomputed (HKY, stepwise, joint loci) ------------------------
// changeu (update_mc_params.cpp:23-431) walks the scalars in order; a proposal trades scalar j against a partner k and needs
// the likelihood of their two loci under the new scalars -- a whole pruning for an HKY locus.  One warp per chain took a
// millisecond per launch on 50 HKY loci.  Two proposals can only influence each other through a locus they share, so the
// proposals are LEVELLED (level = 1 + the latest earlier proposal touching either of its loci, as k_swap levels its attempts)
// and one level is evaluated by the warps of a block in parallel -- every proposal with its own draws (one stream per chain
// and scalar), the accepted likelihood changes added to the chain's sum in proposal order at the end: the chain is the one
// the walk in order would visit with those draws.
constexpr int kUWarps = 8;
IMA_HD size_t changeu_levels_smem_bytes(const EngineDims &d, int nur) {
  return pair_smem_bytes(d) * kUWarps + (size_t)nur * (5 * sizeof(double) + 2 * sizeof(int)) + align8(sizeof(int) * (size_t)d.nloci) + 64;
}
IMA_KERNEL void k_changeu_levels(EngineView E, UpdateView U) {
  IMA_SMEM_DECL
  if (ima_block() >= E.c_n) return;
  const int c = E.c_lo + ima_block();
  const DevModel &M = IMA_MODEL;
  const int lane = Warp::lane(), nloci = E.d.nloci, nur = U.nurates;
  if (((current_step(E) + 1) % (unsigned long long)U.u_every) == 0) {         // every UUPDATEINC+1 steps
    unsigned char *sp = IMA_SMEM + pair_smem_bytes(E.d) * kUWarps;
    double *s_u = (double *)sp; sp += sizeof(double) * nur;                   // the step draw
    double *s_k0 = (double *)sp; sp += sizeof(double) * nur;                  // kappa draws of the two loci
    double *s_k1 = (double *)sp; sp += sizeof(double) * nur;
    double *s_acc = (double *)sp; sp += sizeof(double) * nur;                 // the accept draw
    double *s_delta = (double *)sp; sp += sizeof(double) * nur;               // accepted change of the likelihood sum
    int *s_k = (int *)sp; sp += sizeof(int) * nur;                            // partner
    int *s_lv = (int *)sp; sp += sizeof(int) * nur;                           // level
    int *s_last = (int *)sp; sp += align8(sizeof(int) * nloci);               // per locus: level of the latest proposal touching it
    int *s_ctl = (int *)sp;                                                   // [0] number of levels
    const double beta = E.beta[c];
    IMA_FOR_WARPS(w, kUWarps) {
      const int tid = w * IMA_WARP + lane, nth = kUWarps * IMA_WARP;
      for (int j = tid; j < nur; j += nth) {
        Philox r2;
        const unsigned long long step = current_step(E);
        r2.init(E.seed, (uint32_t)((E.d.chain0 + c) * nur + j), (uint32_t)step, kRngScalars | ((uint32_t)(step >> 32) << 8));
        int k;
        do { k = (int)(r2.uniform() * nur); } while (k == j || k < 0 || k >= nur);          // :78-90
        s_k[j] = k;
        s_u[j] = r2.uniform(); s_k0[j] = r2.uniform(); s_k1[j] = r2.uniform(); s_acc[j] = r2.uniform();
        s_delta[j] = 0.0;
      }
      for (int l = tid; l < nloci; l += nth) s_last[l] = 0;
    }
    block_sync();
    IMA_FOR_WARPS(w, kUWarps) {
      if (w == 0 && lane == 0) {
        int top = 0;
        for (int j = 0; j < nur; j++) {
          const int lj = U.ul_l[j], lk = U.ul_l[s_k[j]];
          const int lv = U.u_levels_in_order ? j + 1 : 1 + (s_last[lj] > s_last[lk] ? s_last[lj] : s_last[lk]);
          s_lv[j] = lv; s_last[lj] = lv; s_last[lk] = lv;
          if (lv > top) top = lv;
        }
        s_ctl[0] = top;
      }
    }
    block_sync();
    const int nlevels = s_ctl[0];
    for (int lv = 1; lv <= nlevels; lv++) {
      IMA_FOR_WARPS(w, kUWarps) {
        PairSm S = carve_pair_smem(IMA_SMEM + (size_t)w * pair_smem_bytes(E.d), E.d);
        for (int j = 0, n = 0; j < nur; j++) {
          if (s_lv[j] != lv) continue;
          if ((n++ % kUWarps) != w) continue;
          const int k = s_k[j];
          const int lj = U.ul_l[j], aj = U.ul_a[j], lk = U.ul_l[k], ak = U.ul_a[k];
          const int pj = c * nloci + lj, pk = c * nloci + lk;
          const double olduj = E.uvals[(size_t)pj * kMaxLinked + aj], olduk = E.uvals[(size_t)pk * kMaxLinked + ak];
          // :201-212: uniform step on the log ratio, reflected at +-maxratio; the two scalars move in opposite directions
          const double r = log(olduj / olduk), u = s_u[j];
          double newr = u > 0.5 ? r + (2.0 * u - 1.0) * U.u_win : r - U.u_win * u * 2.0;
          if (newr > U.u_maxratio) newr = 2.0 * U.u_maxratio - newr;
          else if (newr < -U.u_maxratio) newr = 2.0 * (-U.u_maxratio) - newr;
          const double logd = (newr - r) / 2, d = exp(logd);
          const double newuj = olduj * d, newuk = olduk / d;
          double newpdg[2], newkappa[2] = {0.0, 0.0}, likenewsum = 0.0;
          bool bad = false;
          for (int i = 0; i < 2; i++) {
            const int li = i ? lk : lj, ai = i ? ak : aj, p = i ? pk : pj;
            const PairBuf &B = E.buf[E.cur[p]];
            if (E.loci[li].model == kHKY) newkappa[i] = reflect_kappa(i ? s_k1[j] : s_k0[j], E.kappa[p], U.kappa_win, U.kappa_max);
            newpdg[i] = scalar_likelihood(E, M, c, li, ai, i ? newuk : newuj, i ? -logd : logd, newkappa[i], S, B, E.buf[E.cur[p] ^ 1]);
            if (newpdg[i] == kRejectIS || !(newpdg[i] > -DBL_MAX)) bad = true;
            likenewsum += newpdg[i] - part_pdg(E.loci[li], B, p, ai);
            Warp::sync();
          }
          const double mh = exp(beta * M.gbeta * likenewsum);                               // :291
          const bool accept = !bad && s_acc[j] < (mh < 1.0 ? mh : 1.0);                     // :294
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
                if (L.model == kHKY) {
                  E.kappa[p] = newkappa[i];
                  for (int x = 0; x < E.d.hky_mask_words; x++) B.hky_mask[(size_t)p * E.d.hky_mask_words + x] = Bo.hky_mask[(size_t)p * E.d.hky_mask_words + x];
                }
              }
            }
            if (lane == 0) s_delta[j] = likenewsum;
          }
#if IMA_CUDA
          __threadfence_block();
#endif
          Warp::sync();
          if (lane == 0) {
            if (j == nur - 1) {
              double *o = U.u_out + (size_t)c * 4;
              o[0] = newpdg[0]; o[1] = newpdg[1]; o[2] = mh; o[3] = accept ? 1.0 : 0.0;
            }
#if IMA_CUDA
            atomicAdd(U.stats + 2, 1ull);
            if (accept) atomicAdd(U.stats + 3, 1ull);
#else
            U.stats[2] += 1; if (accept) U.stats[3] += 1;
#endif
            if (beta == 1.0) {                                 // counted for the scalar and for its partner (ima_main_mpi.cpp:1926-1935)
              stat_add(U.stats + kColdUStat + 2 * (size_t)j, 1ull);
              stat_add(U.stats + kColdUStat + 2 * (size_t)k, 1ull);
              if (accept) { stat_add(U.stats + kColdUStat + 2 * (size_t)j + 1, 1ull); stat_add(U.stats + kColdUStat + 2 * (size_t)k + 1, 1ull); }
            }
          }
        }
      }
#if IMA_CUDA
      __threadfence();                                       // a level reads what the level before it wrote to global memory
#endif
      block_sync();
    }
    IMA_FOR_WARPS(w, kUWarps) {
      if (w == 0 && lane == 0) {
        double pd = E.pdgsum[c], ss = E.swapsum[c];
        for (int j = 0; j < nur; j++) if (s_delta[j] != 0.0) { pd += s_delta[j]; ss += s_delta[j]; }
        E.pdgsum[c] = pd; E.swapsum[c] = ss;
      }
    }
    block_sync();
  }
  if (E.xch.publisher == 3) {                          // the chain's step ends here, whether or not it was the scalars' turn
    IMA_FOR_WARPS(w, kUWarps) {
      if (w == 0) {
#if IMA_CUDA
        __threadfence_block();
        __syncwarp();
#endif
        publish_chain(E, c, *(volatile double *)(E.swapsum + c));
      }
    }
  }
}

}  // namespace ima
