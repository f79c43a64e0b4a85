// updategenealogy's proposal half as two kernels (update_gtree.cpp:723-966, update_gtree_common.cpp:850-1575).
//
// The move itself -- pick an edge, detach, slide, re-attach, simulate the migration path, the two getmprob values -- is a
// chain of ~3,500 dependent scalar instructions; the weights and the likelihood of the result are lane-parallel work.  A warp
// that does both spends half its life with one lane active.  So:
//
//   k_move<PPW>   LANE per pair (PPW pairs per warp): the warp stages its pairs' genealogies in shared memory interleaved
//                 element by element (bank = lane, see Arr), lanes 0..PPW-1 each make their own pair's move with the pair's
//                 own random stream, the warp writes the proposed genealogies, compacted, to the pairs' OTHER buffers.
//   k_weigh       WARP per pair on the proposed genealogy: treeweight, P(D|G), flags; what the accept sweep reads.
//   k_propose_redo  the pairs the two kernels above could not take (more migration events than their small tables hold):
//                 the general one-warp-per-pair path of ima_kernels.h, from a list.  The random streams are keyed by
//                 (chain, locus, step, purpose), so a pair that is done again makes exactly the same move.
//
// The chain a run visits is the same whichever path a pair takes (tests: fast_path_equals_general_path).
#pragma once
#include "ima_kernels.h"

namespace ima {

constexpr int kMoveWarps = 4;            // warps of a k_move block
constexpr uint32_t kFlagRedo = 32u;      // the pair goes to the general path (never seen by the accept sweep)

// ---- shared memory of k_move: PPW interleaved pairs per warp --------------------------------------------------------
IMA_HD size_t move_smem_bytes_per_pair(const EngineDims &d) {
  return align8(sizeof(double) * d.NL) + align8(sizeof(double) * d.FP) + 64 + 4 * align8(sizeof(short) * d.NL) +
         2 * align8(sizeof(unsigned short) * d.NL) + align8(sizeof(short) * d.FP) + 48;
}
template <int PPW> IMA_DEV PairSmT<PPW> carve_move_smem(unsigned char *base, const EngineDims &d) {
  PairSmT<PPW> s;
  unsigned char *p = base;
  auto take = [&](size_t bytes_per_pair) { unsigned char *q = p; p += align8(bytes_per_pair) * PPW; return q; };
  s.time.p = (double *)take(sizeof(double) * d.NL);
  s.pt.p = (double *)take(sizeof(double) * d.FP);
  s.ctl_d.p = (double *)take(64);
  s.up0.p = (short *)take(sizeof(short) * d.NL);
  s.up1.p = (short *)take(sizeof(short) * d.NL);
  s.down.p = (short *)take(sizeof(short) * d.NL);
  s.pop.p = (short *)take(sizeof(short) * d.NL);
  s.ms.p = (unsigned short *)take(sizeof(unsigned short) * d.NL);
  s.mcn.p = (unsigned short *)take(sizeof(unsigned short) * d.NL);
  s.pp.p = (short *)take(sizeof(short) * d.FP);
  s.ctl_i.p = (int *)take(48);
  s.evt = nullptr; s.evi = nullptr; s.evk = nullptr; s.pre = nullptr; s.mask = nullptr; s.moff = nullptr; s.gwi = nullptr; s.gwd = nullptr;
  s.pool_free = 0; s.pool_end = d.FP;
#if defined(IMA_PROF)
  s.prof = nullptr; s.nprof = 0;
#endif
  return s;
}
// the view of the t-th pair of the warp
template <int PPW> IMA_DEV PairSmT<PPW> move_slot(const PairSmT<PPW> &s0, int t) {
  PairSmT<PPW> s = s0;
  s.time.p += t; s.pt.p += t; s.ctl_d.p += t; s.up0.p += t; s.up1.p += t; s.down.p += t; s.pop.p += t;
  s.ms.p += t; s.mcn.p += t; s.pp.p += t; s.ctl_i.p += t;
  return s;
}

// per chain group, kind of update (0 genealogy, 1 split time) and step parity: how many pairs are on the redo list
IMA_DEV int *redo_counter(const EngineView &E, int kind, int parity) { return E.redo_count + ((E.grp * 2 + kind) * 2 + parity); }
// the list of a kind: P entries; a group's entries start at its first pair
IMA_DEV int *redo_list(const EngineView &E, int kind) { return E.redo_list + (size_t)kind * E.d.P + (size_t)E.c_lo * E.d.nloci; }
IMA_DEV void redo_push(const EngineView &E, int kind, int p) {
  const int parity = (int)(current_step(E) & 1ull);
#if IMA_CUDA
  const int at = atomicAdd(redo_counter(E, kind, parity), 1);
#else
  const int at = (*redo_counter(E, kind, parity))++;
#endif
  redo_list(E, kind)[at] = p;
}

#if IMA_CUDA
#define IMA_MOVE_BOUNDS __launch_bounds__(kMoveWarps * 32)
#else
#define IMA_MOVE_BOUNDS
#endif
template <int PPW>
IMA_KERNEL void IMA_MOVE_BOUNDS k_move(EngineView E) {
  IMA_SMEM_DECL
  const int npairs = E.c_n * E.d.nloci;
  const int idx0 = (ima_block() * kMoveWarps + ima_warp_in_block()) * PPW;
  if (idx0 >= npairs) return;
  const DevModel &M = IMA_MODEL;
  const int lane = Warp::lane(), NL = E.d.NL, CAP = E.d.CAP, FP = E.d.FP;
  const PairSmT<PPW> S0 = carve_move_smem<PPW>(IMA_SMEM + (size_t)ima_warp_in_block() * move_smem_bytes_per_pair(E.d) * PPW, E.d);
  const int p0 = E.c_lo * E.d.nloci + idx0;
  const int nmine = npairs - idx0 < PPW ? npairs - idx0 : PPW;
#if defined(IMA_PROF) && IMA_CUDA
  long long pm_[4]; pm_[0] = clock64();
#endif
  // ---- stage: the warp copies its pairs (coalesced reads) into the interleaved layout.  Lane t first fetches what says where
  // pair t lives (current buffer, edges, scalars); the copies of all pairs are then issued without waiting for each other.
  int my_cb = 0, my_nl = 0, my_mig = 0, my_root = 0;
  double my_roottime = 0.0;
  if (lane < nmine) {
    const int p = p0 + lane;
    my_cb = E.cur[p];
    my_nl = E.loci[p % E.d.nloci].nl;
    const int *si = my_cb ? E.buf[1].si : E.buf[0].si;
    my_mig = si[(size_t)p * 2 + 1];
    my_root = si[(size_t)p * 2];
    my_roottime = (my_cb ? E.buf[1].sd : E.buf[0].sd)[(size_t)p * 4];
  }
  const int total = nmine * NL;
  // kStageTrips trips of loads are in flight before the first value is stored (the stores to shared memory may alias the
  // next loads as far as the compiler knows: trip by trip every trip waited a whole memory latency)
  constexpr int kStageTrips = 8;
  for (int base0 = 0; base0 < total; base0 += kStageTrips * IMA_WARP) {     // every lane makes every trip (the shuffles need all of them)
    short4_t q[kStageTrips]; double tm[kStageTrips]; ushort2_t m[kStageTrips];
#if IMA_CUDA
#pragma unroll
#endif
    for (int u = 0; u < kStageTrips; u++) {
      const int idx = base0 + u * IMA_WARP + lane;
      const bool valid = idx < total;
      const int t = valid ? idx / NL : 0;
      const int cb = Warp::bcast(my_cb, t);
      if (valid) {
        const size_t g = (size_t)p0 * NL + idx;                        // pair p0 + t, edge idx - t NL
        q[u] = (cb ? E.buf[1].topo : E.buf[0].topo)[g];
        tm[u] = (cb ? E.buf[1].time : E.buf[0].time)[g];
        m[u] = (cb ? E.buf[1].mseg : E.buf[0].mseg)[g];
      }
    }
#if IMA_CUDA
#pragma unroll
#endif
    for (int u = 0; u < kStageTrips; u++) {
      const int idx = base0 + u * IMA_WARP + lane;
      if (idx < total) {
        const int t = idx / NL, i = idx - t * NL;
        const PairSmT<PPW> S = move_slot(S0, t);
        S.up0[i] = q[u].x; S.up1[i] = q[u].y; S.down[i] = q[u].z; S.pop[i] = q[u].w;
        S.time[i] = tm[u];
        S.ms[i] = m[u].x; S.mcn[i] = m[u].y;
      }
    }
  }
  // migration events: the first IMA_WARP of every pair in flight together, the rest (rare) pair by pair
  {
    constexpr int kMigPairs = PPW < 8 ? PPW : 8;
    for (int t0 = 0; t0 < nmine; t0 += kMigPairs) {
      double mt0[kMigPairs]; short mp0[kMigPairs];
#if IMA_CUDA
#pragma unroll
#endif
      for (int u = 0; u < kMigPairs; u++) {
        const int t = t0 + u < nmine ? t0 + u : nmine - 1;
        const int cb = Warp::bcast(my_cb, t), mignum = Warp::bcast(my_mig, t);
        if (t0 + u < nmine && mignum <= FP && lane < mignum) {
          mt0[u] = ((cb ? E.buf[1].mig_t : E.buf[0].mig_t) + (size_t)(p0 + t) * CAP)[lane];
          mp0[u] = ((cb ? E.buf[1].mig_p : E.buf[0].mig_p) + (size_t)(p0 + t) * CAP)[lane];
        }
      }
#if IMA_CUDA
#pragma unroll
#endif
      for (int u = 0; u < kMigPairs; u++) {
        const int t = t0 + u < nmine ? t0 + u : nmine - 1;
        const int cb = Warp::bcast(my_cb, t), mignum = Warp::bcast(my_mig, t);
        if (t0 + u < nmine && mignum <= FP) {
          const PairSmT<PPW> S = move_slot(S0, t);
          if (lane < mignum) { S.pt[lane] = mt0[u]; S.pp[lane] = mp0[u]; }
          const double *mt = (cb ? E.buf[1].mig_t : E.buf[0].mig_t) + (size_t)(p0 + t) * CAP;
          const short *mp = (cb ? E.buf[1].mig_p : E.buf[0].mig_p) + (size_t)(p0 + t) * CAP;
          for (int i = lane + IMA_WARP; i < mignum; i += IMA_WARP) { S.pt[i] = mt[i]; S.pp[i] = mp[i]; }
        }
      }
    }
  }
  if (lane < nmine) {
    const PairSmT<PPW> S = move_slot(S0, lane);
    S.ctl_i[kCiRoot] = my_root;
    S.ctl_i[kCiMignum] = my_mig;
    S.ctl_i[kCiFlags] = my_mig <= FP ? 0 : (int)kFlagOverflow;
    S.ctl_d[kCdRoottime] = my_roottime;
    S.ctl_d[kCdMigw] = 0.0; S.ctl_d[kCdSlidew] = 0.0; S.ctl_d[kCdAterm] = 0.0; S.ctl_d[kCdSlideDist] = 0.0;
    S.ctl_i[kCiEdge] = -1;
  }
#if IMA_CUDA
  __threadfence_block();
#endif
  Warp::sync();
#if defined(IMA_PROF) && IMA_CUDA
  pm_[1] = clock64();
#endif
  // ---- move: one lane per pair ---------------------------------------------------------------------------------------
  if (lane < nmine) {
    PairSmT<PPW> S = move_slot(S0, lane);
    if (!(S.ctl_i[kCiFlags] & (int)kFlagOverflow)) {
      const int p = p0 + lane;
      const int c = p / E.d.nloci, li = p - c * E.d.nloci;
      const DevLocus &L = E.loci[li];
      S.pool_free = S.ctl_i[kCiMignum];
      Philox rng;
      rng_for(rng, E, (uint32_t)((E.d.chain0 + c) * E.d.nloci + li), kRngPropose);
      propose_move(M, E.tvals + (size_t)c * kMaxPeriods, L.ng, L.nl, rng, S);
    }
  }
#if IMA_CUDA
  __threadfence_block();
#endif
  Warp::sync();
#if defined(IMA_PROF) && IMA_CUDA
  pm_[2] = clock64();
#endif
  // ---- store: the warp writes pair after pair, migration lists compacted in edge order, to the pair's other buffer ---
  for (int t = 0; t < nmine; t++) {
    const int p = p0 + t;
    const int nl = Warp::bcast(my_nl, t);
    const PairBuf &Bn = E.buf[Warp::bcast(my_cb, t) ^ 1];
    const PairSmT<PPW> S = move_slot(S0, t);
    uint32_t flags = (uint32_t)S.ctl_i[kCiFlags];
    int total = 0;
    if (!(flags & kFlagOverflow)) {
      short4_t *topo = Bn.topo + (size_t)p * NL;
      double *time = Bn.time + (size_t)p * NL;
      ushort2_t *mseg = Bn.mseg + (size_t)p * NL;
      double *mt = Bn.mig_t + (size_t)p * CAP;
      short *mp = Bn.mig_p + (size_t)p * CAP;
      for (int base = 0; base < nl; base += IMA_WARP) total += (base + lane < nl) ? (int)S.mcn[base + lane] : 0;
      total = Warp::sum(total);
      // what k_weigh can stage (FC events) and the pool row can hold (CAP): otherwise the general path decides
      if (total > E.d.FC || total > CAP) flags |= kFlagOverflow;
      else {
        int carry = 0;
        for (int base = 0; base < nl; base += IMA_WARP) {
          const int i = base + lane;
          const int n = i < nl ? (int)S.mcn[i] : 0;
          const int incl = Warp::scan(n);
          if (i < nl) {
            const int dst = carry + incl - n, src = S.ms[i];
            short4_t q; q.x = S.up0[i]; q.y = S.up1[i]; q.z = S.down[i]; q.w = S.pop[i];
            topo[i] = q;
            time[i] = S.time[i];
            ushort2_t m; m.x = (unsigned short)dst; m.y = (unsigned short)n;
            mseg[i] = m;
            for (int j = 0; j < n; j++) { mt[dst + j] = S.pt[src + j]; mp[dst + j] = S.pp[src + j]; }
          }
          carry += Warp::bcast(incl, IMA_WARP - 1);
        }
      }
    }
    if (lane == 0) {
      if (flags & kFlagOverflow) {
        E.prop_flags[p] = kFlagRedo;
        redo_push(E, 0, p);
      } else {
        Bn.si[(size_t)p * 2] = S.ctl_i[kCiRoot];
        Bn.si[(size_t)p * 2 + 1] = total;
        Bn.sd[(size_t)p * 4] = S.ctl_d[kCdRoottime];
        E.prop_flags[p] = flags;
        E.prop_extra[p] = S.ctl_d[kCdMigw] + S.ctl_d[kCdSlidew];
        if (E.prop_ids) { E.prop_ids[(size_t)p * 2] = (short)S.ctl_i[kCiFreed]; E.prop_ids[(size_t)p * 2 + 1] = (short)S.ctl_i[kCiOldDownDown]; }
        if (E.prop_dbg) {
          E.prop_dbg[(size_t)p * 4 + 0] = S.ctl_d[kCdMigw]; E.prop_dbg[(size_t)p * 4 + 1] = S.ctl_d[kCdSlidew];
          E.prop_dbg[(size_t)p * 4 + 2] = S.ctl_d[kCdSlideDist]; E.prop_dbg[(size_t)p * 4 + 3] = (double)S.ctl_i[kCiEdge];
        }
      }
    }
  }
#if defined(IMA_PROF) && IMA_CUDA
  pm_[3] = clock64();
  if (lane == 0 && (idx0 / PPW) % 161 == 0) printf("PROFM ppw %d idx0 %d stage %lld move %lld store %lld\n", PPW, idx0, pm_[1] - pm_[0], pm_[2] - pm_[1], pm_[3] - pm_[2]);
#endif
}

// ---- shared memory of k_weigh: one pair per warp, tables sized for FC migration events -------------------------------
// the prefix table doubles as the 2 x 16 transition probabilities of the HKY likelihood: never under 32 doubles
IMA_HD size_t weigh_pre_bytes(const EngineDims &d) { const size_t b = sizeof(unsigned long long) * d.FEV * d.W64; return b < 256 ? 256 : b; }
IMA_HD size_t weigh_smem_bytes(const EngineDims &d, int pool = -1) {
  size_t b = 0;
  if (pool < 0) pool = d.FC;
  b += align8(sizeof(double) * d.NL) + align8(sizeof(double) * pool) + align8(sizeof(double) * d.FEV);
  b += align8(weigh_pre_bytes(d)) + align8(sizeof(double) * d.ND) + 64;
  b += 4 * align8(sizeof(short) * d.NL) + 2 * align8(sizeof(unsigned short) * d.NL) + align8(sizeof(short) * pool);
  b += 2 * align8(sizeof(int) * d.FEV) + align8(sizeof(uint32_t) * d.NL * d.W) + align8(sizeof(int) * (d.NL + 1)) + align8(sizeof(int) * d.NI) + 48;
  return b;
}
IMA_DEV PairSm carve_weigh_smem(unsigned char *base, const EngineDims &d, int pool = -1) {
  PairSm s;
  unsigned char *p = base;
  if (pool < 0) pool = d.FC;
  auto take = [&](size_t bytes) { unsigned char *q = p; p += align8(bytes); return q; };
  s.time.p = (double *)take(sizeof(double) * d.NL);
  s.pt.p = (double *)take(sizeof(double) * pool);
  s.evt = (double *)take(sizeof(double) * d.FEV);
  s.pre = (unsigned long long *)take(weigh_pre_bytes(d));
  s.gwd = (double *)take(sizeof(double) * d.ND);
  s.ctl_d.p = (double *)take(64);
  s.up0.p = (short *)take(sizeof(short) * d.NL);
  s.up1.p = (short *)take(sizeof(short) * d.NL);
  s.down.p = (short *)take(sizeof(short) * d.NL);
  s.pop.p = (short *)take(sizeof(short) * d.NL);
  s.ms.p = (unsigned short *)take(sizeof(unsigned short) * d.NL);
  s.mcn.p = (unsigned short *)take(sizeof(unsigned short) * d.NL);
  s.pp.p = (short *)take(sizeof(short) * pool);
  s.evi = (int *)take(sizeof(int) * d.FEV);
  s.evk = (int *)take(sizeof(int) * d.FEV);
  s.mask = (uint32_t *)take(sizeof(uint32_t) * d.NL * d.W);
  s.moff = (int *)take(sizeof(int) * (d.NL + 1));
  s.gwi = (int *)take(sizeof(int) * d.NI);
  s.ctl_i.p = (int *)take(48);
  s.pool_free = d.FC; s.pool_end = pool;
#if defined(IMA_PROF)
  s.prof = nullptr; s.nprof = 0;
#endif
  return s;
}

#ifndef IMA_WEIGH_MINBLOCKS
#define IMA_WEIGH_MINBLOCKS 8
#endif
#if IMA_CUDA
#define IMA_WEIGH_BOUNDS __launch_bounds__(kWarpsPerBlock * 32, IMA_WEIGH_MINBLOCKS)
#else
#define IMA_WEIGH_BOUNDS
#endif
// kind 0: after k_move (the proposed genealogy is in the pair's other buffer; its likelihood is computed here)
IMA_KERNEL void IMA_WEIGH_BOUNDS k_weigh(EngineView E) {
  IMA_SMEM_DECL
  const int idx = ima_block() * kWarpsPerBlock + ima_warp_in_block();
  const int lane = Warp::lane();
  // the list of the NEXT step's parity starts empty (the list of this step is still to be read by k_propose_redo)
  if (idx == 0 && lane == 0) *redo_counter(E, 0, (int)((current_step(E) + 1) & 1ull)) = 0;
  if (idx >= E.c_n * E.d.nloci) return;
  const DevModel &M = IMA_MODEL;
  const int c = E.c_lo + idx / E.d.nloci, li = idx % E.d.nloci, p = c * E.d.nloci + li;
  uint32_t flags = E.prop_flags[p];
  if (flags & kFlagRedo) return;
  const DevLocus &L = E.loci[li];
  PairSm S = carve_weigh_smem(IMA_SMEM + (size_t)ima_warp_in_block() * weigh_smem_bytes(E.d), E.d);
  const PairBuf &Bn = E.buf[E.cur[p] ^ 1];
  const double *tv = E.tvals + (size_t)c * kMaxPeriods;
#if defined(IMA_PROF) && IMA_CUDA
  long long pf_[16]; S.prof = pf_; S.nprof = 0;
  IMA_SPROF(S)
#endif
  stage_pair(E, Bn, p, L.nl, S);
  bool ok = eval_weights(M, E.d, L, tv, S, E.d.FEV);
  double pdg = 0.0;
  if (ok) {
    double pdga[kMaxLinked];
    const PairBuf &Bc = E.buf[E.cur[p]];
    HkyCall hk; hk.mode = kHkyPartial; hk.freed = hk.olddd = -1;
    if (L.model == kHKY) { hk.freed = E.prop_ids[(size_t)p * 2]; hk.olddd = E.prop_ids[(size_t)p * 2 + 1]; }
    hk.mask_cur = Bc.hky_mask ? Bc.hky_mask + (size_t)p * E.d.hky_mask_words : nullptr;
    hk.mask_new = Bn.hky_mask ? Bn.hky_mask + (size_t)p * E.d.hky_mask_words : nullptr;
    pdg = pair_likelihood(E, L, Bn, p, S, pdga, hk);
    flags |= (uint32_t)S.ctl_i[kCiFlags];
    if (pdg == kRejectIS) flags |= kFlagRejectIS;
    if (!(flags & (kFlagRejectIS | kFlagBadTree))) {
      for (int i = lane; i < E.d.NI; i += IMA_WARP) Bn.gwi[(size_t)p * E.d.NI + i] = S.gwi[i];
      for (int i = lane; i < E.d.ND; i += IMA_WARP) Bn.gwd[(size_t)p * E.d.ND + i] = S.gwd[i];
      if (lane == 0) {
        Bn.sd[(size_t)p * 4 + 1] = S.ctl_d[kCdLength];
        Bn.sd[(size_t)p * 4 + 2] = S.ctl_d[kCdTlength];
        Bn.sd[(size_t)p * 4 + 3] = pdg;
      }
    }
  } else {
    flags = kFlagRedo;                                   // cannot happen while FEV covers FC events; the general path decides
    if (lane == 0) redo_push(E, 0, p);
  }
  if (lane == 0) E.prop_flags[p] = flags;
#if defined(IMA_PROF) && IMA_CUDA
  IMA_SPROF(S)
  if (lane == 0 && idx % 641 == 0) {
    printf("PROFW %d nev %d:", idx, S.ctl_i[kCiNev]);
    for (int i = 1; i < S.nprof; i++) printf(" %lld", pf_[i] - pf_[i - 1]);
    printf("  (stage+scan build sort zero prefix sweep fc | keys search | store)\n");
  }
#endif
}

// the pairs on the redo list of this step, by the general path
IMA_KERNEL void IMA_PROPOSE_BOUNDS k_propose_redo(EngineView E) {
  IMA_SMEM_DECL
  const int n = *redo_counter(E, 0, (int)(current_step(E) & 1ull));
  const int *list = redo_list(E, 0);
  PairSm S = carve_pair_smem(IMA_SMEM + (size_t)ima_warp_in_block() * pair_smem_bytes(E.d), E.d);
  for (int k = ima_block() * kWarpsPerBlock + ima_warp_in_block(); k < n; k += E.redo_grid * kWarpsPerBlock) {
    const int p = list[k];
    propose_pair_general(E, IMA_MODEL, p / E.d.nloci, p % E.d.nloci, S);
    Warp::sync();
  }
}

}  // namespace ima
