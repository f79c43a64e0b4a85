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
  const unsigned long long step = current_step(E);
  rng.init(E.seed, stream_id, (uint32_t)step, purpose | ((uint32_t)(step >> 32) << 8));
}

// P(D|G) of the staged genealogy for the locus's mutation model; every lane returns the value
// hk: how an HKY locus treats its stored partials (likelihood_hky); ignored by the other models
IMA_DEV double pair_likelihood(const EngineView &E, const DevLocus &L, const PairBuf &B, int p, PairSm &S, double *pdg_a_out, const HkyCall &hk) {
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
    const double v = likelihood_hky(E, L, S, p, u[0], E.kappa[p], E.pi + (size_t)p * 4, hk);
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
  HkyCall hk; hk.mode = kHkyInit; hk.freed = hk.olddd = -1; hk.mask_cur = nullptr;
  hk.mask_new = B.hky_mask ? B.hky_mask + (size_t)p * E.d.hky_mask_words : nullptr;
  const double pdg = ok ? pair_likelihood(E, L, B, p, S, pdga, hk) : 0.0;
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

// ---- the exchange of swap sums between GPUs (struct Exchange) ------------------------------------------------------------
#if !IMA_CUDA
// host emulation (tests): ranks in separate processes share their tables through POSIX shared memory; a waiting rank polls for
// IMA2P_EMU_WAIT_MS milliseconds (default 0: ranks stepped in lockstep by one process never have to wait)
inline bool emu_wait_for(volatile unsigned long long *word, unsigned long long target) {
  const char *w = getenv("IMA2P_EMU_WAIT_MS");
  const long budget_ms = w ? atol(w) : 0;
  for (long waited_us = 0; *word < target; waited_us += 50) {
    if (waited_us / 1000 >= budget_ms) return false;
    struct timespec ts = {0, 50000};
    nanosleep(&ts, nullptr);
  }
  __sync_synchronize();
  return true;
}
#endif
// called by one whole warp when chain c's S for this step is final
IMA_DEV void publish_chain(const EngineView &E, int c, double S) {
  const Exchange &X = E.xch;
  const unsigned long long ep = current_step(E) - X.step0;
  const size_t at = (size_t)(ep & 1ull) * E.d.nchains_global + (size_t)(E.d.chain0 + c);
  for (int r = Warp::lane(); r < X.world; r += IMA_WARP) {
#if IMA_CUDA
    *(volatile double *)(X.peer_S[r] + at) = S;                   // a store into rank r's memory (its own when r == rank)
    __threadfence_system();                                       // the value is visible before the count that announces it
    atomicAdd_system(X.peer_arrived[r] + (ep & 1ull), 1ull);
#else
    X.peer_S[r][at] = S;
    __sync_fetch_and_add(X.peer_arrived[r] + (ep & 1ull), 1ull);          // ranks may be separate processes on shared memory
#endif
  }
}
// the swap kernel's side: wait until every chain of the job has arrived for this step; returns this rank's table of the step
IMA_DEV const double *await_swap_sums(const EngineView &E, int step_bias) {
  const Exchange &X = E.xch;
  const unsigned long long ep = current_step(E) - (unsigned long long)step_bias - X.step0;
  const unsigned long long target = (ep / 2ull + 1ull) * (unsigned long long)E.d.nchains_global;
#if IMA_CUDA
  if (Warp::lane() == 0) {
    volatile unsigned long long *cnt = X.peer_arrived[X.rank] + (ep & 1ull);
    const long long t0 = clock64();
    while (*cnt < target) {
      if (clock64() - t0 > 30000000000ll) { raise(E.mc, kErrExchange); break; }     // a peer never arrived (about fifteen seconds)
    }
    __threadfence_system();
  }
  __syncwarp();
#else
  if (!emu_wait_for(X.peer_arrived[X.rank] + (ep & 1ull), target)) raise(E.mc, kErrExchange);
#endif
  return X.peer_S[X.rank] + (size_t)(ep & 1ull) * E.d.nchains_global;
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

// IMA_PROF (tuning builds only): a few warps / blocks print the clock cycles they spent per stage
#if defined(IMA_PROF) && IMA_CUDA
#define IMA_PROF_DECL(n) long long prof_t_[n]; int prof_i_ = 0; prof_t_[0] = clock64();
#define IMA_PROF_MARK() prof_t_[++prof_i_] = clock64();
#else
#define IMA_PROF_DECL(n)
#define IMA_PROF_MARK()
#endif
#ifndef IMA_PROPOSE_MINBLOCKS
#define IMA_PROPOSE_MINBLOCKS 5      // <= 102 registers/thread, 20 resident warps per SM: best of 5/6/8 measured on B200
#endif
#if IMA_CUDA
#define IMA_PROPOSE_BOUNDS __launch_bounds__(kWarpsPerBlock * 32, IMA_PROPOSE_MINBLOCKS)
#else
#define IMA_PROPOSE_BOUNDS
#endif
// updategenealogy's proposal half for one pair by one warp, any model, any migration load up to the pool capacity: the
// general path.  (The shipped workloads go through the two kernels of ima_fastpath.h and come here only when a pair does not
// fit their smaller tables.)
IMA_DEV void propose_pair_general(const EngineView &E, const DevModel &M, int c, int li, PairSm &S) {
  const int p = c * E.d.nloci + li;
  const int idx = p;                                     // named in the tuning build's cycle stamps
  (void)idx;
  const DevLocus &L = E.loci[li];
  const int cb = E.cur[p];
  const PairBuf &B = E.buf[cb];
  const PairBuf &Bn = E.buf[cb ^ 1];
  const double *tv = E.tvals + (size_t)c * kMaxPeriods;
  const int lane = Warp::lane();
  IMA_PROF_DECL(8)
  stage_pair(E, B, p, L.nl, S);
  IMA_PROF_MARK()
  if (lane == 0) {
    Philox rng;
    rng_for(rng, E, (uint32_t)((E.d.chain0 + c) * E.d.nloci + li), kRngPropose);
    propose_move(M, tv, L.ng, L.nl, rng, S);
  }
  Warp::sync();
  IMA_PROF_MARK()
  uint32_t flags = (uint32_t)S.ctl_i[kCiFlags];
  bool ok = !(flags & kFlagOverflow);
  if (ok) ok = eval_weights(M, E.d, L, tv, S);
  IMA_PROF_MARK()
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
    HkyCall hk; hk.mode = kHkyPartial; hk.freed = S.ctl_i[kCiFreed]; hk.olddd = S.ctl_i[kCiOldDownDown];
    hk.mask_cur = B.hky_mask ? B.hky_mask + (size_t)p * E.d.hky_mask_words : nullptr;
    hk.mask_new = Bn.hky_mask ? Bn.hky_mask + (size_t)p * E.d.hky_mask_words : nullptr;
    pdg = pair_likelihood(E, L, Bn, p, S, pdga, hk);
    IMA_PROF_MARK()
    flags = (uint32_t)S.ctl_i[kCiFlags];
    if (pdg == kRejectIS) flags |= kFlagRejectIS;
    if (lane == 0) S.ctl_d[kCdPdg] = pdg;
    Warp::sync();
    if (!(flags & (kFlagRejectIS | kFlagBadTree))) store_pair(E, Bn, p, L.nl, S, total_mig);
    IMA_PROF_MARK()
#if defined(IMA_PROF) && IMA_CUDA
    if (lane == 0 && idx % 641 == 0 && prof_i_ == 5)
      printf("PROFP %d stage %lld move %lld weights %lld like %lld store %lld flags %u\n", idx, prof_t_[1] - prof_t_[0], prof_t_[2] - prof_t_[1],
             prof_t_[3] - prof_t_[2], prof_t_[4] - prof_t_[3], prof_t_[5] - prof_t_[4], flags);
#endif
  } else {
    flags |= kFlagOverflow;
  }
  if (lane == 0) {
    E.prop_flags[p] = flags;
    E.prop_extra[p] = S.ctl_d[kCdMigw] + S.ctl_d[kCdSlidew] + S.ctl_d[kCdAterm];
    if (E.prop_dbg) {                                     // parity tests only (ima2p_engine_set_debug_records)
      E.prop_dbg[(size_t)p * 4 + 0] = S.ctl_d[kCdMigw]; E.prop_dbg[(size_t)p * 4 + 1] = S.ctl_d[kCdSlidew];
      E.prop_dbg[(size_t)p * 4 + 2] = S.ctl_d[kCdSlideDist]; E.prop_dbg[(size_t)p * 4 + 3] = (double)S.ctl_i[kCiEdge];
    }
  }
}

// every locus of chains [c_lo, c_lo + c_n) (one warp per pair)
IMA_KERNEL void IMA_PROPOSE_BOUNDS k_propose(EngineView E) {
  IMA_SMEM_DECL
  const int idx = ima_block() * kWarpsPerBlock + ima_warp_in_block();
  if (idx >= E.c_n * E.d.nloci) return;
  PairSm S = carve_pair_smem(IMA_SMEM + (size_t)ima_warp_in_block() * pair_smem_bytes(E.d), E.d);
  propose_pair_general(E, IMA_MODEL, E.c_lo + idx / E.d.nloci, idx % E.d.nloci, S);
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
// prior, so locus li sees the sums left by every accepted update before it.  What shortens that
// dependent chain without changing its result:
//   * inside a locus the nq + nm prior terms are independent: each goes to its own warp, which runs the term's
//     series / continued fraction 32 terms per round (ima_math.h *_coop);
//   * the next B-1 loci are evaluated SPECULATIVELY in the same round, against the same sums.  Decisions are then
//     taken in locus order; the first acceptance changes the sums, so every speculative locus after it is
//     thrown away and re-evaluated in the next round.  About two thirds of the updates are rejected, so a round
//     decides 1 + q + q^2 ... loci on average (q = rejection rate).  The random number of a locus depends only on
//     (chain, locus, step), so the outcome is identical to the one-locus-at-a-time sweep;
//   * nothing on the per-round path touches global memory: the weight records of a whole run of loci (all of them
//     when they fit, else chunks of accept_chunk()) are brought into shared memory by every thread of the block at
//     once, with the uniforms drawn one locus per thread, before the rounds start;
//   * ONE barrier per round.  Every warp keeps its own copy of the chain's sums and prior terms; after the barrier that
//     publishes the round's candidate terms every warp takes the round's decisions itself (the same arithmetic on the same
//     numbers: the same outcome in every warp) and commits the accepted locus to its own copy.  No warp waits for another
//     one's commit, and the candidate terms are double-buffered by round so the next round's terms can be written while a
//     slower warp still reads this round's.  Warp 0 alone performs the global side effects (buffer flip, counters).
#if IMA_CUDA
constexpr int kSpecMax = 4;          // speculative depth B
constexpr int kTermWarps = 5;        // warps per speculative locus (terms are strided over them)
IMA_DEV void block_sync() { __syncthreads(); }
#define IMA_FOR_WARPS(wv, nw) for (int wv = ima_warp_in_block(), once_ = 1; once_; once_ = 0)
#else
constexpr int kSpecMax = 4;
constexpr int kTermWarps = 5;
IMA_DEV void block_sync() {}
#define IMA_FOR_WARPS(wv, nw) for (int wv = 0; wv < (nw); wv++)      // host emulation: one thread plays every warp in turn
#endif
constexpr int kAcceptSmemBudget = 100 * 1024;   // two blocks per SM stay possible
constexpr int kAcceptWarpsMax = kSpecMax * kTermWarps;

struct AcceptSm {
  unsigned char *priv;                                       // [warps] private copies: ad[ND], q[NT], ai[NI]
  double *cq;                                                // [2][kSpecMax][NT] candidate terms per speculative locus, by round parity
  int *cflag;                                                // [2][kSpecMax] candidate would put migration where the model forbids it
  int *r_dI; double *r_oD, *r_nD, *r_sc; int *r_ic;          // [chunk] records; r_ic: flags, current buffer, next locus to decide
};
IMA_HD size_t accept_priv_bytes(const EngineDims &d) { return align8(sizeof(double) * d.ND) + align8(sizeof(double) * d.NT) + align8(sizeof(int) * d.NI); }
IMA_HD size_t accept_fixed_bytes(const EngineDims &d) {
  return kAcceptWarpsMax * accept_priv_bytes(d) + align8(sizeof(double) * 2 * kSpecMax * d.NT) + align8(sizeof(int) * 2 * kSpecMax);
}
IMA_HD size_t accept_record_bytes(const EngineDims &d) {
  return align8(sizeof(int) * d.NI) + 2 * align8(sizeof(double) * d.ND) + 4 * 8 + 16;
}
// loci whose records are resident at a time
IMA_HD int accept_chunk(const EngineDims &d) {
  long long k = ((long long)kAcceptSmemBudget - (long long)accept_fixed_bytes(d)) / (long long)accept_record_bytes(d);
  if (k > d.nloci) k = d.nloci;
  if (k < 1) k = 1;
  return (int)k;
}
IMA_HD size_t accept_smem_bytes(const EngineDims &d) { return accept_fixed_bytes(d) + (size_t)accept_chunk(d) * accept_record_bytes(d); }
IMA_DEV AcceptSm carve_accept_smem(unsigned char *base, const EngineDims &d, int K) {
  AcceptSm s; unsigned char *p = base;
  s.priv = p; p += kAcceptWarpsMax * accept_priv_bytes(d);
  s.cq = (double *)p; p += align8(sizeof(double) * 2 * kSpecMax * d.NT);
  s.r_oD = (double *)p; p += (size_t)K * align8(sizeof(double) * d.ND);
  s.r_nD = (double *)p; p += (size_t)K * align8(sizeof(double) * d.ND);
  s.r_sc = (double *)p; p += (size_t)K * 4 * 8;              // oldpdg, newpdg, extra, uniform
  s.r_dI = (int *)p; p += (size_t)K * align8(sizeof(int) * d.NI);
  s.r_ic = (int *)p; p += (size_t)K * 16;                    // flags, cb, next undecided-and-decidable slot, unused
  s.cflag = (int *)p;
  return s;
}

// depth 3 and 4: one block per SM; depth 1 and 2: two blocks per SM, which is what a GPU holding more chains than it
// has SMs needs
#if IMA_CUDA
#define IMA_ACCEPT_BOUNDS(B) __launch_bounds__((B) * kTermWarps * 32, (B) >= 3 ? 1 : 2)
#else
#define IMA_ACCEPT_BOUNDS(B)
#endif
template <int B>
IMA_KERNEL void IMA_ACCEPT_BOUNDS(B) k_accept(EngineView E) {
  IMA_SMEM_DECL
  if (ima_block() >= E.c_n) return;
  const int c = E.c_lo + ima_block();
  const int l0 = 0, l1 = E.d.nloci;
  const DevModel &M = IMA_MODEL;
  const int NW = B * kTermWarps;                                     // warps in the block
  const int lane = Warp::lane(), NI = E.d.NI, ND = E.d.ND, NT = E.d.NT, ncc = M.ncc;
  const int sI = (int)(align8(sizeof(int) * NI) / sizeof(int)), sD = (int)(align8(sizeof(double) * ND) / sizeof(double));
  const int K = accept_chunk(E.d);
  AcceptSm S = carve_accept_smem(IMA_SMEM, E.d, K);
  const int nq = M.nq, nmt = M.nomigration ? 0 : M.nm, nterms = nq + nmt;
  // every warp's own copy of the chain's sums and prior terms (term t of the size parameters at [t], of the migration
  // parameters at [nq + t])
  const size_t pb = accept_priv_bytes(E.d);
  IMA_FOR_WARPS(w, NW) {
    double *ad = (double *)(S.priv + (size_t)w * pb), *q = ad + sD;
    int *ai = (int *)(q + (align8(sizeof(double) * NT) / sizeof(double)));
    for (int i = lane; i < NI; i += IMA_WARP) ai[i] = E.all_i[(size_t)c * NI + i];
    for (int i = lane; i < ND; i += IMA_WARP) ad[i] = E.all_d[(size_t)c * ND + i];
    for (int i = lane; i < nq; i += IMA_WARP) q[i] = E.qint[(size_t)c * kMaxParams + i];
    for (int i = lane; i < nmt; i += IMA_WARP) q[nq + i] = E.mint[(size_t)c * kMaxParams + i];
  }
  const double beta = E.beta[c];
  double probg = E.probg[c], pdgsum = E.pdgsum[c];                    // kept by every warp
  unsigned long long ndropped = 0;                                   // proposals that arrived flagged as dropped (per thread)
  constexpr uint32_t kNoGo = kFlagRejectIS | kFlagOverflow | kFlagBadTree;
  int round = 0;
#if defined(IMA_PROF) && IMA_CUDA
  auto pclk_ = []() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); return t; };
  long long pf_[6] = {0, 0, 0, 0, 0, 0}, pt_ = pclk_(), pt0_ = pt_; int prounds_ = 0;
#define IMA_PF(k) { const long long n_ = pclk_(); pf_[k] += n_ - pt_; pt_ = n_; }
#else
#define IMA_PF(k)
#endif
  for (int ch0 = l0; ch0 < l1; ch0 += K) {
    const int ch1 = (ch0 + K < l1) ? ch0 + K : l1;
    block_sync();                                                    // the previous chunk's records are no longer read
    // every record of loci [ch0, ch1): slot = l - ch0.  All loads are independent, one pass of the whole block.
    IMA_FOR_WARPS(w, NW) {
      const int tid = w * IMA_WARP + lane, nth = NW * IMA_WARP;
      for (int l = ch0 + tid; l < ch1; l += nth) {
        const int p = c * E.d.nloci + l, slot = l - ch0;
        const int cb = E.cur[p];
        const PairBuf &O = E.buf[cb], &N = E.buf[cb ^ 1];
        const uint32_t fl = E.prop_flags[p];
        S.r_ic[slot * 4 + 0] = (int)fl; S.r_ic[slot * 4 + 1] = cb; S.r_ic[slot * 4 + 3] = 0;
        if (fl & kFlagOverflow) ndropped++;
        S.r_sc[slot * 4 + 0] = O.sd[(size_t)p * 4 + 3]; S.r_sc[slot * 4 + 1] = N.sd[(size_t)p * 4 + 3]; S.r_sc[slot * 4 + 2] = E.prop_extra[p];
        Philox rng;
        rng_for(rng, E, (uint32_t)((E.d.chain0 + c) * E.d.nloci + l), kRngAccept);
        // U < min(1, e^x) is taken as log U < min(0, x): the logarithm is off the chain of decisions, the exponential was on it
        S.r_sc[slot * 4 + 3] = log(rng.uniform());
      }
      // kLoadTrips trips per thread in flight: first which buffer is current for each, then the values
      const int nl = ch1 - ch0, per = NI + 2 * ND, nrec = nl * per;
      constexpr int kLoadTrips = 3;
      for (int k0 = tid; k0 < nrec; k0 += kLoadTrips * nth) {
        int cbv[kLoadTrips];
        long long raw[kLoadTrips];
#if IMA_CUDA
#pragma unroll
#endif
        for (int u = 0; u < kLoadTrips; u++) {
          const int k = k0 + u * nth;
          cbv[u] = k < nrec ? (int)E.cur[c * E.d.nloci + ch0 + k / per] : 0;
        }
#if IMA_CUDA
#pragma unroll
#endif
        for (int u = 0; u < kLoadTrips; u++) {
          const int k = k0 + u * nth;
          if (k < nrec) {
            const int slot = k / per, i = k - slot * per;
            const int p = c * E.d.nloci + ch0 + slot;
            const PairBuf &O = E.buf[cbv[u]], &N = E.buf[cbv[u] ^ 1];
            if (i < NI) raw[u] = (long long)(N.gwi[(size_t)p * NI + i] - O.gwi[(size_t)p * NI + i]);
            else if (i < NI + ND) raw[u] = dbl_bits(O.gwd[(size_t)p * ND + (i - NI)]);
            else raw[u] = dbl_bits(N.gwd[(size_t)p * ND + (i - NI - ND)]);
          }
        }
#if IMA_CUDA
#pragma unroll
#endif
        for (int u = 0; u < kLoadTrips; u++) {
          const int k = k0 + u * nth;
          if (k < nrec) {
            const int slot = k / per, i = k - slot * per;
            if (i < NI) S.r_dI[slot * sI + i] = (int)raw[u];
            else if (i < NI + ND) S.r_oD[slot * sD + (i - NI)] = bits_dbl(raw[u]);
            else S.r_nD[slot * sD + (i - NI - ND)] = bits_dbl(raw[u]);
          }
        }
      }
    }
    block_sync();
    // A proposal that arrives flagged (the data rule it out, or it was dropped) is rejected whatever the prior says: such
    // loci are passed over without spending a speculative slot on them.  r_ic[slot][2] = the first slot >= slot that needs a
    // decision (the chunk's length when there is none).
    IMA_FOR_WARPS(w, NW) {
      const int tid = w * IMA_WARP + lane, nth = NW * IMA_WARP, nl = ch1 - ch0;
      for (int sl = tid; sl < nl; sl += nth) {
        int k = sl;
        while (k < nl && ((uint32_t)S.r_ic[k * 4] & kNoGo)) k++;
        S.r_ic[sl * 4 + 2] = k;
      }
    }
    block_sync();
    if (ndropped) {                                                  // every locus is consumed exactly once: counted where it is loaded
#if IMA_CUDA
      atomicAdd(E.overflow, ndropped);
#else
      *E.overflow += ndropped;
#endif
      ndropped = 0;
    }
    IMA_PF(3)
    for (int li = ch0; li < ch1;) {
      // cand[g] = offset from li of the g-th locus that needs a decision, span = loci consumed when none of them is accepted
      int cand[B], nb = 0;
      const int visible = ch1 - li;
      {
        int sl = li - ch0;
        for (int g = 0; g < B; g++) {
          const int k = sl < ch1 - ch0 ? S.r_ic[sl * 4 + 2] : ch1 - ch0;
          if (k < ch1 - ch0) { cand[nb++] = k - (li - ch0); sl = k + 1; } else sl = ch1 - ch0;
        }
        for (int g = nb; g < B; g++) cand[g] = 0;
      }
      const int span = nb == B ? cand[B - 1] + 1 : visible;   // nb == B: up to and including the last candidate; else everything left in the chunk
      double *cq = S.cq + (size_t)(round & 1) * kSpecMax * NT;
      int *cflag = S.cflag + (round & 1) * kSpecMax;
      // ---- phase 1: term t of speculative candidate g on warp g*kTermWarps + (t % kTermWarps) ------------------
      IMA_FOR_WARPS(w, NW) {
        const int g = w / kTermWarps, t0 = w - g * kTermWarps;
        const double *ad = (const double *)(S.priv + (size_t)w * pb), *q = ad + sD;
        const int *ai = (const int *)(q + (align8(sizeof(double) * NT) / sizeof(double)));
        if (g < nb) {
          const int slot = li - ch0 + cand[g];
          const int *dI = S.r_dI + slot * sI;
          const double *oD = S.r_oD + slot * sD, *nD = S.r_nD + slot * sD;
          if (t0 == 0 && lane == 0) {
            int bad = 0;
            if (M.nomigration == 0)
              for (int i = 0; i < M.nomig_n; i++)
                if (ai[ncc + M.nomig_idx[i]] + dI[ncc + M.nomig_idx[i]] != 0) bad = 1;
            cflag[g] = bad;
          }
          // candidate sums = sum_subtract_treeinfo (ginfo.cpp:248-285) on the entries this term reads: subtract old, add
          // new, clamp fc and fm at 0; then integrate_tree_prob's reuse rule (:1997-2000, :2031-2034) or the integral
          for (int t = t0; t < nterms; t += kTermWarps) {
            double v;
            if (t < nq) {
              int cn = 0, co = 0; double fn = 0.0, fo = 0.0, hn = 0.0;
              for (int j = 0; j < M.q_n[t]; j++) {
                const int x = M.q_idx[t][j];
                co += ai[x]; cn += ai[x] + dI[x];
                fo += ad[x];
                double f = ad[x]; f -= oD[x]; f += nD[x]; if (0.0 > f) f = 0.0;
                fn += f;
                double hh = ad[ncc + x]; hh -= oD[ncc + x]; hh += nD[ncc + x];
                hn += hh;
              }
              v = (cn == co && fn == fo) ? q[t] : integrate_coalescent_term_coop(E.mc, cn, fn, hn, M.q_max[t], M.q_min[t]);
            } else {
              const int tm = t - nq;
              int cn = 0, co = 0; double fn = 0.0, fo = 0.0;
              for (int j = 0; j < M.m_n[tm]; j++) {
                const int x = M.m_idx[tm][j];
                co += ai[ncc + x]; cn += ai[ncc + x] + dI[ncc + x];
                fo += ad[2 * ncc + x];
                double f = ad[2 * ncc + x]; f -= oD[2 * ncc + x]; f += nD[2 * ncc + x]; if (0.0 > f) f = 0.0;
                fn += f;
              }
              v = (cn == co && fn == fo) ? q[t]
                  : (M.expoprior ? integrate_migration_term_expo(E.mc, cn, fn, M.m_mean[tm])
                                 : integrate_migration_term_coop(E.mc, cn, fn, M.m_max[tm], M.m_min[tm]));
            }
            if (lane == 0) cq[g * NT + t] = v;
          }
        }
      }
      IMA_PF(0)
      block_sync();
      IMA_PF(1)
      // ---- phase 2: every warp decides in locus order (update_gtree.cpp:917-927) and commits the accepted locus to its copy
      int accepted = -1;
      double np = 0.0;
      IMA_FOR_WARPS(w, NW) {
#if IMA_CUDA
        {
          // lane g evaluates the MH ratio of speculative candidate g (all against the same sums); the first accepting
          // lane in locus order wins
          bool acc = false;
          double newprobg = 0.0;
          if (lane < nb) {
            const int g = lane, slot = li - ch0 + cand[lane];
            for (int t = 0; t < nterms; t++) newprobg += cq[g * NT + t];
            if (cflag[g]) newprobg = -kMyDblMax;
            const double tpw = newprobg - probg, dpdg = S.r_sc[slot * 4 + 1] - S.r_sc[slot * 4 + 0], extra = S.r_sc[slot * 4 + 2];
            double mh;
            if (M.thermo) mh = beta * M.gbeta * dpdg + tpw + extra;
            else mh = beta * (tpw + M.gbeta * dpdg) + extra;
            acc = S.r_sc[slot * 4 + 3] < fmin(0.0, mh);
          }
          accepted = Warp::first(acc);
          np = Warp::bcast(newprobg, accepted < 0 ? 0 : accepted);
        }
#else
        accepted = -1;
        for (int g = 0; g < nb && accepted < 0; g++) {       // one lane: walk the speculative loci in order
          const int slot = li - ch0 + cand[g];
          double npg = 0.0;
          for (int t = 0; t < nterms; t++) npg += cq[g * NT + t];
          if (cflag[g]) npg = -kMyDblMax;
          const double tpw = npg - probg, dpdg = S.r_sc[slot * 4 + 1] - S.r_sc[slot * 4 + 0], extra = S.r_sc[slot * 4 + 2];
          double mh;
          if (M.thermo) mh = beta * M.gbeta * dpdg + tpw + extra;
          else mh = beta * (tpw + M.gbeta * dpdg) + extra;
          if (S.r_sc[slot * 4 + 3] < fmin(0.0, mh)) { accepted = g; np = npg; }
        }
#endif
        if (accepted >= 0) {
          const int slot = li - ch0 + cand[accepted];
          const int *dI = S.r_dI + slot * sI;
          const double *oD = S.r_oD + slot * sD, *nD = S.r_nD + slot * sD;
          double *ad = (double *)(S.priv + (size_t)w * pb), *q = ad + sD;
          int *ai = (int *)(q + (align8(sizeof(double) * NT) / sizeof(double)));
          for (int i = lane; i < NI; i += IMA_WARP) ai[i] += dI[i];
          for (int i = lane; i < ND; i += IMA_WARP) {
            double x = ad[i];
            x -= oD[i];
            x += nD[i];
            if ((i < ncc || i >= 2 * ncc) && 0.0 > x) x = 0.0;
            ad[i] = x;
          }
          for (int i = lane; i < nterms; i += IMA_WARP) q[i] = cq[accepted * NT + i];
#if IMA_CUDA
          __syncwarp();
#endif
          if (w == 0 && lane == 0) S.r_ic[slot * 4 + 3] = 1;          // accepted: the global side effects follow the chunk's rounds
        }
      }
      // what every warp keeps in registers (after the loop: the host emulation plays all warps with one set of variables)
      const int aoff = accepted < 0 ? 0 : cand[accepted];
      const int adv = accepted < 0 ? span : aoff + 1;
      if (accepted >= 0) {
        const int slot = li - ch0 + aoff;
        probg = np;
        pdgsum -= S.r_sc[slot * 4 + 0];
        pdgsum += S.r_sc[slot * 4 + 1];
      }
      IMA_PF(2)
      li += adv;
      round++;
#if defined(IMA_PROF) && IMA_CUDA
      prounds_++;
#endif
    }
    // the chunk's accepted loci: flip their buffers, count them (the whole block, off the chain of decisions)
    block_sync();
    IMA_FOR_WARPS(w, NW) {
      const int tid = w * IMA_WARP + lane, nth = NW * IMA_WARP;
      for (int sl = tid; sl < ch1 - ch0; sl += nth) {
        if (!S.r_ic[sl * 4 + 3]) continue;
        const int p = c * E.d.nloci + ch0 + sl;
        const uint32_t flags = (uint32_t)S.r_ic[sl * 4];
        E.cur[p] = (unsigned char)(S.r_ic[sl * 4 + 1] ^ 1);
#if IMA_CUDA
        atomicAdd(&E.acc[(size_t)p * 3 + 0], 1u);
        if (flags & kFlagTopol) atomicAdd(&E.acc[(size_t)p * 3 + 1], 1u);
        if (flags & kFlagTmrca) atomicAdd(&E.acc[(size_t)p * 3 + 2], 1u);
        if (beta == 1.0) {
          unsigned int *ca = E.cold_acc + (size_t)(ch0 + sl) * 3;
          atomicAdd(ca, 1u);
          if (flags & kFlagTopol) atomicAdd(ca + 1, 1u);
          if (flags & kFlagTmrca) atomicAdd(ca + 2, 1u);
        }
#else
        E.acc[(size_t)p * 3 + 0]++;
        if (flags & kFlagTopol) E.acc[(size_t)p * 3 + 1]++;
        if (flags & kFlagTmrca) E.acc[(size_t)p * 3 + 2]++;
        if (beta == 1.0) {
          unsigned int *ca = E.cold_acc + (size_t)(ch0 + sl) * 3;
          ca[0]++;
          if (flags & kFlagTopol) ca[1]++;
          if (flags & kFlagTmrca) ca[2]++;
        }
#endif
      }
    }
  }
#if defined(IMA_PROF) && IMA_CUDA
  if (c == 3 && lane == 0)
    printf("PROFA chain %d warp %d rounds %d terms %lld barrier %lld decide+commit %lld load %lld total %lld\n", c, ima_warp_in_block(), prounds_, pf_[0], pf_[1],
           pf_[2], pf_[3], pclk_() - pt0_);
#endif
  block_sync();                                                      // the flips above are read by chain_swapsum below
  // warp 0 writes the chain's state back from its copy
  IMA_FOR_WARPS(w, NW) {
    if (w == 0) {
      const double *ad = (const double *)(S.priv), *q = ad + sD;
      const int *ai = (const int *)(q + (align8(sizeof(double) * NT) / sizeof(double)));
      for (int i = lane; i < NI; i += IMA_WARP) E.all_i[(size_t)c * NI + i] = ai[i];
      for (int i = lane; i < ND; i += IMA_WARP) E.all_d[(size_t)c * ND + i] = ad[i];
      for (int i = lane; i < nq; i += IMA_WARP) E.qint[(size_t)c * kMaxParams + i] = q[i];
      for (int i = lane; i < nmt; i += IMA_WARP) E.mint[(size_t)c * kMaxParams + i] = q[nq + i];
#if IMA_CUDA
      __threadfence_block();
      __syncwarp();
#endif
      const double ssum = (l1 == E.d.nloci) ? chain_swapsum(E, M, c, probg) : 0.0;
      if (lane == 0) {
        E.probg[c] = probg; E.pdgsum[c] = pdgsum;
        if (l1 == E.d.nloci) E.swapsum[c] = ssum;
      }
      if (E.xch.publisher == 1) publish_chain(E, c, ssum);
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
  unsigned long long *adj_counts;    // [nchains_global][2] attempts, accepts between temperature ranks r and r + 1 (tempbasedswapcount, swapchains.cpp:760-778)
  int swaptries, advance_step;   // advance_step: what the launch adds to the device step counter afterwards (0, 1, or the steps of a graph)
  int step_bias;            // 1 when the step counter was already advanced (split-phase multi-GPU step): the draws stay keyed by the step they belong to
  int smem_chains;          // chains the launch's shared memory can stage (0: work on global memory)
  int sequential;           // tests: one lane walks the attempts in order (what the level-wise walk must reproduce)
  int use_exchange;         // S of all chains comes from the exchange tables (struct Exchange), after waiting for them
};

constexpr int kSwapBatch = 256;           // attempts drawn per pass
IMA_DEV void stat_add(unsigned long long *p, unsigned long long v) {
#if IMA_CUDA
  atomicAdd(p, v);
#else
  *p += v;
#endif
}
IMA_HD size_t swap_smem_bytes(int staged_chains) { return (size_t)staged_chains * 28 + 16 + (size_t)kSwapBatch * 20; }
IMA_KERNEL void k_swap(EngineView E, SwapView V) {
  IMA_SMEM_DECL
  if (ima_block() != 0 || ima_warp_in_block() != 0) return;
  const int N = E.d.nchains_global;
  const int lane = Warp::lane();
  if (V.use_exchange) V.S_global = await_swap_sums(E, V.step_bias);
  if (N > 1 && V.swaptries > 0) {
    // the attempts are a dependent chain of small decisions over (beta, S, rank): the warp stages those in shared memory
    // (when they fit) so that one lane walks the attempts without waiting on global memory, and writes the ranks back
    const bool staged = V.smem_chains >= N;
    double *sS = (double *)IMA_SMEM, *sB = sS + (staged ? N : 0);
    int *sC = (int *)(sB + (staged ? N : 0)), *sR = sC + (staged ? N : 0);
    if (staged) {
      for (int i = lane; i < N; i += IMA_WARP) { sS[i] = *(const volatile double *)(V.S_global + i); sB[i] = V.beta_table[i]; sC[i] = V.chain_of_rank[i]; sR[i] = V.rank_of_chain[i]; }
#if IMA_CUDA
      __threadfence_block();
#endif
      Warp::sync();
    }
    const double *Sg = staged ? sS : V.S_global, *Bt = staged ? sB : V.beta_table;
    int *cor = staged ? sC : V.chain_of_rank, *roc = staged ? sR : V.rank_of_chain;
    // Every attempt has its own counter block of the step's swap stream, so the lanes draw the attempts of a batch in
    // parallel -- the two temperature ranks (the second uniform over the other ranks of the window, which is what the
    // reference's redraw-until-different loop samples, swapchains.cpp:224-235) and the uniform of the decision -- and one
    // decisions that share no temperature rank commute.  One lane gives every attempt its level (one more than the last earlier
    // attempt that touched either of its ranks: four shared-memory operations per attempt, no arithmetic), then the attempts
    // of a level are decided by the lanes together -- an exponential each -- and the levels follow one another.
    double *dU = (double *)(IMA_SMEM + (staged ? (size_t)N * 24 : 0) + 16);
    int *dA = (int *)(dU + kSwapBatch), *dB = dA + kSwapBatch, *dL = dB + kSwapBatch;
    int *last = dL + kSwapBatch;                              // [N] level of the last attempt that touched a rank (staged tables only)
    const unsigned long long step = current_step(E) - (unsigned long long)V.step_bias;
    unsigned long long nacc = 0;
    for (int x0 = 0; x0 < V.swaptries; x0 += kSwapBatch) {
      const int nb = V.swaptries - x0 < kSwapBatch ? V.swaptries - x0 : kSwapBatch;
      for (int i = lane; i < nb; i += IMA_WARP) {
        Philox rng;
        rng.init(E.seed, 0xffffffffu, (uint32_t)step, kRngSwap | ((uint32_t)(step >> 32) << 8));
        rng.ctr[0] = (uint32_t)(x0 + i);
        const int sa = rng.randint(N);
        int sbmin = 0, sbrange = N;
        if (N >= 2 * kSwapDist + 3) {
          sbmin = sa - kSwapDist > 0 ? sa - kSwapDist : 0;
          sbrange = (N < sa + kSwapDist ? N : sa + kSwapDist) - sbmin;
        }
        int sb = sbmin + rng.randint(sbrange - 1);
        if (sb >= sa) sb++;
        dA[i] = sa; dB[i] = sb; dU[i] = rng.uniform();
      }
#if IMA_CUDA
      __threadfence_block();
#endif
      Warp::sync();
      if (staged && IMA_WARP > 1 && !V.sequential) {
        int nlevels = 0;
        for (int i = lane; i < N; i += IMA_WARP) last[i] = 0;
#if IMA_CUDA
        __threadfence_block();
#endif
        Warp::sync();
        if (lane == 0) {
          for (int i = 0; i < nb; i++) {
            const int sa = dA[i], sb = dB[i];
            const int la = last[sa], lb = last[sb], lv = (la > lb ? la : lb) + 1;
            dL[i] = lv; last[sa] = lv; last[sb] = lv;
            if (lv > nlevels) nlevels = lv;
          }
        }
#if IMA_CUDA
        __threadfence_block();
#endif
        Warp::sync();
        nlevels = Warp::bcast(nlevels, 0);
        for (int lv = 1; lv <= nlevels; lv++) {
          for (int i = lane; i < nb; i += IMA_WARP) {
            if (dL[i] != lv) continue;
            const int sa = dA[i], sb = dB[i];
            const int ca = cor[sa], cb = cor[sb];
            const double w = exp((Bt[sa] - Bt[sb]) * (Sg[cb] - Sg[ca]));
            const bool swapped = w >= 1.0 || w > dU[i];
            if (swapped) {
              cor[sa] = cb; cor[sb] = ca;
              roc[ca] = sb; roc[cb] = sa;
              nacc++;
            }
            if (sa - sb == 1 || sb - sa == 1) {
              unsigned long long *ac = V.adj_counts + (size_t)(sa < sb ? sa : sb) * 2;
              stat_add(ac, 1ull);
              if (swapped) stat_add(ac + 1, 1ull);
            }
          }
#if IMA_CUDA
          __threadfence_block();
#endif
          Warp::sync();
        }
      } else if (lane == 0) {
        for (int i = 0; i < nb; i++) {
          const int sa = dA[i], sb = dB[i];
          const int ca = cor[sa], cb = cor[sb];
          const double w = exp((Bt[sa] - Bt[sb]) * (Sg[cb] - Sg[ca]));
          const bool swapped = w >= 1.0 || w > dU[i];
          if (swapped) {
            cor[sa] = cb; cor[sb] = ca;
            roc[ca] = sb; roc[cb] = sa;
            nacc++;
          }
          if (sa - sb == 1 || sb - sa == 1) {                           // fire and forget, off the chain of decisions
            unsigned long long *ac = V.adj_counts + (size_t)(sa < sb ? sa : sb) * 2;
            stat_add(ac, 1ull);
            if (swapped) stat_add(ac + 1, 1ull);
          }
        }
      }
      Warp::sync();
    }
    const int nacc_all = Warp::sum((int)nacc);                // the lanes of the level-wise walk counted their own
    if (lane == 0) {
      V.swap_counts[0] += (unsigned long long)V.swaptries;
      V.swap_counts[1] += (unsigned long long)nacc_all;
    }
#if IMA_CUDA
    __threadfence_block();
#endif
    Warp::sync();
    if (staged) for (int i = lane; i < N; i += IMA_WARP) { V.chain_of_rank[i] = sC[i]; V.rank_of_chain[i] = sR[i]; }
    for (int c = lane; c < E.d.nchains; c += IMA_WARP) E.beta[c] = Bt[roc[E.d.chain0 + c]];
  }
  if (V.advance_step && lane == 0) *E.nsteps += (unsigned long long)V.advance_step;
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

// ima2p_engine_put_state_packed: the narrow wire form of a pair -- int8 (up0, up1, down, pop) and a uint8 migration count per
// edge, the pools in edge order -- widened to the resident layout (short4 links, (start, count) segments).  One warp per pair.
IMA_KERNEL void k_unpack_state(EngineView E, const signed char *topo8, const unsigned char *mcount) {
  const int p = ima_block() * kWarpsPerBlock + ima_warp_in_block();
  if (p >= E.d.P) return;
  const int lane = Warp::lane(), NL = E.d.NL;
  const PairBuf &B = E.buf[0];
  int carry = 0;
  for (int base = 0; base < NL; base += IMA_WARP) {
    const int e = base + lane;
    const int n = e < NL ? (int)mcount[(size_t)p * NL + e] : 0;
    const int incl = Warp::scan(n);
    if (e < NL) {
      const signed char *t = topo8 + ((size_t)p * NL + e) * 4;
      short4_t o; o.x = t[0]; o.y = t[1]; o.z = t[2]; o.w = t[3];
      B.topo[(size_t)p * NL + e] = o;
      ushort2_t m; m.x = (unsigned short)(carry + incl - n); m.y = (unsigned short)n;
      B.mseg[(size_t)p * NL + e] = m;
    }
    carry += Warp::bcast(incl, IMA_WARP - 1);
  }
}

// ima2p_engine_put_state_block: the whole state of the GPU's chains as ONE host block (one PCIe transfer), sections in the
// order of StateBlock, migration events stored ragged (the events of pair 0, then pair 1, ...).  k_block_offsets turns the
// per-pair event counts into offsets, k_unpack_block widens a pair into the resident layout (one warp per pair).
struct StateBlock { size_t time, sd, uvals, tvals, mig_t, si, mig_p, topo8, mcount, total; };
IMA_HD StateBlock state_block_layout(const EngineDims &d, int nsplit, long long events) {
  StateBlock b; size_t o = 0;
  const size_t P = (size_t)d.P, NL = (size_t)d.NL;
  b.time = o; o += P * NL * 8;
  b.sd = o; o += P * 4 * 8;
  b.uvals = o; o += P * kMaxLinked * 8;
  b.tvals = o; o += (size_t)d.nchains * (nsplit > 0 ? nsplit : 1) * 8;
  b.mig_t = o; o += (size_t)events * 8;
  b.si = o; o += P * 2 * 4;
  b.mig_p = o; o += align8((size_t)events * 2);
  b.topo8 = o; o += align8(P * NL * 4);
  b.mcount = o; o += align8(P * NL);
  b.total = o;
  return b;
}
IMA_KERNEL void k_block_offsets(EngineView E, const unsigned char *block, StateBlock L, int *moff) {
  // exclusive prefix sum of the pairs' event counts (scal_i[p][1]); one block, pairs in chunks of its size
  IMA_SMEM_DECL
  const int *si = (const int *)(block + L.si);
#if IMA_CUDA
  int *sm = (int *)IMA_SMEM;                               // [warps]
  const int lane = Warp::lane(), warp = ima_warp_in_block();
  const int nth = kWarpsPerBlock * IMA_WARP, tid = warp * IMA_WARP + lane;
  int carry = 0;
  for (int base = 0; base < E.d.P; base += nth) {
    const int p = base + tid;
    const int n = p < E.d.P ? si[2 * p + 1] : 0;
    const int incl = Warp::scan(n);
    if (lane == IMA_WARP - 1) sm[warp] = incl;
    __syncthreads();
    int woff = 0, tot = 0;
    for (int w = 0; w < kWarpsPerBlock; w++) { if (w < warp) woff += sm[w]; tot += sm[w]; }
    if (p < E.d.P) moff[p] = carry + woff + incl - n;
    carry += tot;
    __syncthreads();
  }
#else
  if (ima_warp_in_block() != 0) return;                    // host emulation: one lane walks the pairs
  int run = 0;
  for (int p = 0; p < E.d.P; p++) { moff[p] = run; run += si[2 * p + 1]; }
#endif
}
IMA_KERNEL void k_unpack_block(EngineView E, const unsigned char *block, StateBlock L, const int *moff, int nsplit) {
  const int p = ima_block() * kWarpsPerBlock + ima_warp_in_block();
  if (p >= E.d.P) return;
  const int lane = Warp::lane(), NL = E.d.NL, CAP = E.d.CAP;
  const PairBuf &B = E.buf[0];
  const signed char *topo8 = (const signed char *)(block + L.topo8);
  const unsigned char *mcount = block + L.mcount;
  const double *time = (const double *)(block + L.time);
  int carry = 0;
  for (int base = 0; base < NL; base += IMA_WARP) {
    const int e = base + lane;
    const int n = e < NL ? (int)mcount[(size_t)p * NL + e] : 0;
    const int incl = Warp::scan(n);
    if (e < NL) {
      const signed char *t = topo8 + ((size_t)p * NL + e) * 4;
      short4_t o; o.x = t[0]; o.y = t[1]; o.z = t[2]; o.w = t[3];
      B.topo[(size_t)p * NL + e] = o;
      ushort2_t m; m.x = (unsigned short)(carry + incl - n); m.y = (unsigned short)n;
      B.mseg[(size_t)p * NL + e] = m;
      B.time[(size_t)p * NL + e] = time[(size_t)p * NL + e];
    }
    carry += Warp::bcast(incl, IMA_WARP - 1);
  }
  const int *si = (const int *)(block + L.si);
  const double *sd = (const double *)(block + L.sd), *uv = (const double *)(block + L.uvals);
  const int nmig = si[2 * p + 1] < CAP ? si[2 * p + 1] : CAP, m0 = moff[p];
  const double *mt = (const double *)(block + L.mig_t);
  const short *mp = (const short *)(block + L.mig_p);
  for (int i = lane; i < nmig; i += IMA_WARP) { B.mig_t[(size_t)p * CAP + i] = mt[m0 + i]; B.mig_p[(size_t)p * CAP + i] = mp[m0 + i]; }
  for (int i = lane; i < 2; i += IMA_WARP) B.si[(size_t)p * 2 + i] = si[2 * p + i];
  for (int i = lane; i < 4; i += IMA_WARP) B.sd[(size_t)p * 4 + i] = sd[(size_t)p * 4 + i];
  for (int i = lane; i < kMaxLinked; i += IMA_WARP) E.uvals[(size_t)p * kMaxLinked + i] = uv[(size_t)p * kMaxLinked + i];
  if (lane == 0) E.cur[p] = 0;
  const int c = p / E.d.nloci;
  if (p == c * E.d.nloci) {                                   // the first pair of a chain also sets the chain's split times
    const double *tv = (const double *)(block + L.tvals);
    for (int k = lane; k < kMaxPeriods; k += IMA_WARP) E.tvals[(size_t)c * kMaxPeriods + k] = k < nsplit ? tv[(size_t)c * nsplit + k] : kTimeMax;
  }
}

// The cold chain's record of a sharded job, wherever the chain lives: msg = [rowlen floats of the .ti row (as doubles) | probg |
// pdg | pdg of every locus].  The rank that holds the chain at beta = 1 writes it into rank 0's table and then the request's
// sequence number; rank 0 waits for that number and copies the message out for its host.  Every rank launches this kernel.
IMA_KERNEL void k_cold_message(EngineView E, const int *chain_of_rank, int rowlen, unsigned long long seq, double *out_rank0) {
  if (ima_block() != 0 || ima_warp_in_block() != 0) return;
  const DevModel &M = IMA_MODEL;
  const Exchange &X = E.xch;
  const int lane = Warp::lane(), nloci = E.d.nloci;
  const int c = chain_of_rank[0] - E.d.chain0;
  if (c >= 0 && c < E.d.nchains) {
    volatile double *msg = X.cold_msg0;
    if (lane == 0) {
      const int *wi = E.all_i + (size_t)c * E.d.NI;
      const double *wd = E.all_d + (size_t)c * E.d.ND;
      const int nq = M.nq, nm = M.nomigration ? 0 : M.nm;
      const int fcp = nq, hccp = 2 * nq, mcp = 3 * nq, fmp = mcp + nm, qip = fmp + nm, mip = qip + nq, pdgp = mip + nm;
      for (int i = 0; i < nq; i++) {                       // savegsampinf ginfo.cpp:318-377 (sums accumulated in float, as there)
        int cc = 0; float f = 0.f, hc = 0.f;
        for (int j = 0; j < M.q_n[i]; j++) { const int x = M.q_idx[i][j]; cc += wi[x]; f += (float)wd[x]; hc += (float)wd[M.ncc + x]; }
        msg[i] = (float)cc; msg[fcp + i] = f; msg[hccp + i] = hc; msg[qip + i] = (float)E.qint[(size_t)c * kMaxParams + i];
      }
      for (int i = 0; i < nm; i++) {
        int cm = 0; float f = 0.f;
        for (int j = 0; j < M.m_n[i]; j++) { const int x = M.m_idx[i][j]; cm += wi[M.ncc + x]; f += (float)wd[2 * M.ncc + x]; }
        msg[mcp + i] = (float)cm; msg[fmp + i] = f; msg[mip + i] = (float)E.mint[(size_t)c * kMaxParams + i];
      }
      msg[pdgp] = (float)E.pdgsum[c]; msg[pdgp + 1] = (float)E.probg[c];
      for (int i = 0; i < M.nsplit; i++) msg[pdgp + 2 + i] = (float)E.tvals[(size_t)c * kMaxPeriods + i];
      msg[rowlen] = E.probg[c]; msg[rowlen + 1] = E.pdgsum[c];
    }
    for (int li = lane; li < nloci; li += IMA_WARP) {
      const int p = c * nloci + li;
      msg[rowlen + 2 + li] = E.buf[E.cur[p]].sd[(size_t)p * 4 + 3];
    }
#if IMA_CUDA
    __threadfence_system();
    __syncwarp();
    if (lane == 0) { *(volatile unsigned long long *)X.cold_seq0 = seq; __threadfence_system(); }
#else
    __sync_synchronize();
    *X.cold_seq0 = seq;
#endif
  }
  if (X.rank == 0) {
#if IMA_CUDA
    if (lane == 0) {
      volatile unsigned long long *sq = X.cold_seq0;
      const long long t0 = clock64();
      while (*sq < seq) { if (clock64() - t0 > 30000000000ll) { raise(E.mc, kErrExchange); break; } }
      __threadfence_system();
    }
    __syncwarp();
#else
    if (!emu_wait_for(X.cold_seq0, seq)) raise(E.mc, kErrExchange);
#endif
    for (int i = lane; i < X.cold_len; i += IMA_WARP) out_rank0[i] = ((volatile double *)X.cold_msg0)[i];
  }
}

IMA_KERNEL void k_copy_swapsum(EngineView E, double *dst, int advance_step) {
  const int i = ima_block() * kWarpsPerBlock * IMA_WARP + ima_warp_in_block() * IMA_WARP + Warp::lane();
  if (i < E.d.nchains) dst[i] = E.swapsum[i];
  if (advance_step && i == 0) *E.nsteps += 1;
}

}  // namespace ima
