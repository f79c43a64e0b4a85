// Per-(chain, locus) device work: one warp, the genealogy staged in shared memory.
//
//   stage_pair       coalesced load of the pair's edge arrays and migration pool into shared memory
//   propose_move     the updategenealogy proposal (update_gtree.cpp:723-827): pick an edge, detach,
//                    slide with reflection, re-attach, simulate the migration path, Hastings terms
//   eval_weights     treeweight (update_gtree_common.cpp:1679-1931): event build, bitonic sort by
//                    time, sweep; also produces the subtree tip masks used by the IS likelihood
//   likelihood_is    infinite sites (calc_prob_data.cpp:731-836) as a bit-mask equality search
//   store_pair       coalesced write of the proposed state into the pair's other buffer
//
// Lane-parallel loops are written `for (i = lane; i < n; i += IMA_WARP)`; strictly sequential parts
// (tree surgery, event sweep) run on lane 0 between Warp::sync() calls.
#pragma once
#include "ima_model.h"

namespace ima {

// Array view of one pair's data in shared memory.  ST == 1: the pair has the arrays to itself (one warp works on the pair).
// ST > 1: ST pairs are interleaved element by element (element i of this pair lives at base[i * ST]) and one LANE works on
// each pair -- whatever elements the lanes of a warp touch, they fall into different shared-memory banks.
template <class T, int ST> struct Arr {
  T *p;
  IMA_DEV T &operator[](int i) const { return p[i * ST]; }
};

// shared-memory view of one pair (pointers into the warp's slice of dynamic shared memory)
template <int ST> struct PairSmT {
  Arr<double, ST> time;            // [NL]
  Arr<short, ST> up0, up1, down, pop;   // [NL]
  Arr<unsigned short, ST> ms, mcn;      // [NL] migration segment start / count (into pt/pp)
  Arr<double, ST> pt;              // pool.  warp-per-pair kernels: [4*CAP], [0,CAP) current lists, [CAP,2CAP) join scratch, [2CAP,3CAP)
                                   // new edge, [3CAP,4CAP) new sister; the lane-per-pair move kernel carves its smaller pool itself
  Arr<short, ST> pp;
  double *evt;             // [EVP] event times (sort keys)
  int *evi;                // [EVP] packed event info
  int *evk;                // [EVP] period of the event | lineages before it << 8
  unsigned long long *pre; // [EVP*W64] exclusive prefix of the per-population lineage deltas (16-bit fields)
  uint32_t *mask;          // [NL*W] subtree tip masks
  int *moff;               // [NL+1] exclusive prefix of mcn
  int *gwi;                // [NI]
  double *gwd;             // [ND]
  Arr<double, ST> ctl_d;   // [8]  roottime, length, tlength, migweight, slideweight, pdg, Aterm, slide distance
  Arr<int, ST> ctl_i;      // [12] root, mignum, flags, nev, edge, freed, oldsis, newsis, parent of freed before the move
  int pool_free, pool_end;         // pool entries [pool_free, pool_end) are scratch: propose_move builds its lists there
#if defined(IMA_PROF)
  long long *prof; int nprof;      // tuning builds: clock marks inside eval_weights / likelihood_is
#endif
};
typedef PairSmT<1> PairSm;

#if defined(IMA_PROF) && IMA_CUDA
#define IMA_SPROF(S) { if ((S).prof) (S).prof[(S).nprof++] = clock64(); }
#else
#define IMA_SPROF(S)
#endif

enum { kCdRoottime = 0, kCdLength, kCdTlength, kCdMigw, kCdSlidew, kCdPdg, kCdAterm, kCdSlideDist };
enum { kCiRoot = 0, kCiMignum, kCiFlags, kCiNev, kCiEdge, kCiFreed, kCiOldsis, kCiNewsis, kCiOldDownDown };

IMA_HD size_t align8(size_t x) { return (x + 7) & ~(size_t)7; }

// bytes of shared memory one warp needs
IMA_HD size_t pair_smem_bytes(const EngineDims &d) {
  size_t b = 0;
  b += align8(sizeof(double) * d.NL);
  b += align8(sizeof(short) * d.NL) * 4;
  b += align8(sizeof(unsigned short) * d.NL) * 2;
  b += align8(sizeof(double) * 4 * d.CAP);
  b += align8(sizeof(short) * 4 * d.CAP);
  b += align8(sizeof(double) * d.EVP);
  b += align8(sizeof(int) * d.EVP) * 2;
  b += align8(sizeof(unsigned long long) * d.EVP * d.W64);
  b += align8(sizeof(uint32_t) * d.NL * d.W);
  b += align8(sizeof(int) * (d.NL + 1));
  b += align8(sizeof(int) * d.NI);
  b += align8(sizeof(double) * d.ND);
  b += align8(sizeof(double) * 8);
  b += align8(sizeof(int) * 12);
  return b;
}

IMA_DEV PairSm carve_pair_smem(unsigned char *base, const EngineDims &d) {
  PairSm s;
  unsigned char *p = base;
  auto take = [&](size_t bytes) { unsigned char *q = p; p += align8(bytes); return q; };
  s.time.p = (double *)take(sizeof(double) * d.NL);
  s.pt.p = (double *)take(sizeof(double) * 4 * d.CAP);
  s.evt = (double *)take(sizeof(double) * d.EVP);
  s.pre = (unsigned long long *)take(sizeof(unsigned long long) * d.EVP * d.W64);
  s.gwd = (double *)take(sizeof(double) * d.ND);
  s.ctl_d.p = (double *)take(sizeof(double) * 8);
  s.up0.p = (short *)take(sizeof(short) * d.NL);
  s.up1.p = (short *)take(sizeof(short) * d.NL);
  s.down.p = (short *)take(sizeof(short) * d.NL);
  s.pop.p = (short *)take(sizeof(short) * d.NL);
  s.ms.p = (unsigned short *)take(sizeof(unsigned short) * d.NL);
  s.mcn.p = (unsigned short *)take(sizeof(unsigned short) * d.NL);
  s.pp.p = (short *)take(sizeof(short) * 4 * d.CAP);
  s.evi = (int *)take(sizeof(int) * d.EVP);
  s.evk = (int *)take(sizeof(int) * d.EVP);
  s.mask = (uint32_t *)take(sizeof(uint32_t) * d.NL * d.W);
  s.moff = (int *)take(sizeof(int) * (d.NL + 1));
  s.gwi = (int *)take(sizeof(int) * d.NI);
  s.ctl_i.p = (int *)take(sizeof(int) * 12);
  s.pool_free = d.CAP; s.pool_end = 4 * d.CAP;
#if defined(IMA_PROF)
  s.prof = nullptr; s.nprof = 0;
#endif
  return s;
}

// ------------------------------------------------------------------------------------------------
// staging
// ------------------------------------------------------------------------------------------------
IMA_DEV void stage_pair(const EngineView &E, const PairBuf &B, int p, int nl, PairSm &S) {
  const int lane = Warp::lane();
  const short4_t *topo = B.topo + (size_t)p * E.d.NL;
  const double *time = B.time + (size_t)p * E.d.NL;
  const ushort2_t *mseg = B.mseg + (size_t)p * E.d.NL;
  // every global load is issued before anything waits for one: the scalars, the first IMA_WARP migration events (read
  // whether or not the genealogy has that many: the pool row exists) and the edges; only a genealogy with more events than
  // lanes makes a second, dependent trip
  const double *mt = B.mig_t + (size_t)p * E.d.CAP;
  const short *mp = B.mig_p + (size_t)p * E.d.CAP;
  const int mignum = B.si[(size_t)p * 2 + 1], root = B.si[(size_t)p * 2];
  const double roottime = B.sd[(size_t)p * 4];
  const double mt0 = lane < E.d.CAP ? mt[lane] : 0.0;
  const short mp0 = lane < E.d.CAP ? mp[lane] : (short)0;
  // the first two trips over the edges (64 edges: every locus of the shipped inputs) are in flight together
  constexpr int kTrips = 2;
  short4_t q[kTrips]; double tm[kTrips]; ushort2_t ms[kTrips];
#if IMA_CUDA
#pragma unroll
#endif
  for (int u = 0; u < kTrips; u++) {
    const int i = lane + u * IMA_WARP;
    if (i < nl) { q[u] = topo[i]; tm[u] = time[i]; ms[u] = mseg[i]; }
  }
#if IMA_CUDA
#pragma unroll
#endif
  for (int u = 0; u < kTrips; u++) {
    const int i = lane + u * IMA_WARP;
    if (i < nl) {
      S.up0[i] = q[u].x; S.up1[i] = q[u].y; S.down[i] = q[u].z; S.pop[i] = q[u].w;
      S.time[i] = tm[u];
      S.ms[i] = ms[u].x; S.mcn[i] = ms[u].y;
    }
  }
  for (int i = lane + kTrips * IMA_WARP; i < nl; i += IMA_WARP) {
    short4_t t = topo[i];
    S.up0[i] = t.x; S.up1[i] = t.y; S.down[i] = t.z; S.pop[i] = t.w;
    S.time[i] = time[i];
    ushort2_t m = mseg[i];
    S.ms[i] = m.x; S.mcn[i] = m.y;
  }
  if (lane < mignum) { S.pt[lane] = mt0; S.pp[lane] = mp0; }
  for (int i = lane + IMA_WARP; i < mignum; i += IMA_WARP) { S.pt[i] = mt[i]; S.pp[i] = mp[i]; }
  if (lane == 0) {
    S.ctl_i[kCiRoot] = root;
    S.ctl_i[kCiMignum] = mignum;
    S.ctl_i[kCiFlags] = 0;
    S.ctl_d[kCdRoottime] = roottime;
    S.ctl_d[kCdMigw] = 0.0; S.ctl_d[kCdSlidew] = 0.0; S.ctl_d[kCdAterm] = 0.0;
  }
  Warp::sync();
}

// exclusive prefix sum of the per-edge migration counts -> S.moff[0..nl]; returns total
IMA_DEV int scan_mig_counts(int nl, PairSm &S) {
  const int lane = Warp::lane();
  int carry = 0;
  for (int base = 0; base < nl; base += IMA_WARP) {
    int i = base + lane;
    int v = (i < nl) ? (int)S.mcn[i] : 0;
    int inc = Warp::scan(v);
    if (i < nl) S.moff[i] = carry + inc - v;
    carry += Warp::bcast(inc, IMA_WARP - 1);
  }
  if (lane == 0) S.moff[nl] = carry;
  Warp::sync();
  return carry;
}

IMA_DEV void store_pair(const EngineView &E, const PairBuf &B, int p, int nl, const PairSm &S, int total_mig) {
  const int lane = Warp::lane();
  short4_t *topo = B.topo + (size_t)p * E.d.NL;
  double *time = B.time + (size_t)p * E.d.NL;
  ushort2_t *mseg = B.mseg + (size_t)p * E.d.NL;
  double *mt = B.mig_t + (size_t)p * E.d.CAP;
  short *mp = B.mig_p + (size_t)p * E.d.CAP;
  for (int i = lane; i < nl; i += IMA_WARP) {
    short4_t t; t.x = S.up0[i]; t.y = S.up1[i]; t.z = S.down[i]; t.w = S.pop[i];
    topo[i] = t;
    time[i] = S.time[i];
    ushort2_t m; m.x = (unsigned short)S.moff[i]; m.y = S.mcn[i];
    mseg[i] = m;
    const int src = S.ms[i], dst = S.moff[i], n = S.mcn[i];
    for (int j = 0; j < n; j++) { mt[dst + j] = S.pt[src + j]; mp[dst + j] = S.pp[src + j]; }   // compaction
  }
  for (int i = lane; i < E.d.NI; i += IMA_WARP) B.gwi[(size_t)p * E.d.NI + i] = S.gwi[i];
  for (int i = lane; i < E.d.ND; i += IMA_WARP) B.gwd[(size_t)p * E.d.ND + i] = S.gwd[i];
  if (lane == 0) {
    B.si[(size_t)p * 2] = S.ctl_i[kCiRoot];
    B.si[(size_t)p * 2 + 1] = total_mig;
    B.sd[(size_t)p * 4 + 0] = S.ctl_d[kCdRoottime];
    B.sd[(size_t)p * 4 + 1] = S.ctl_d[kCdLength];
    B.sd[(size_t)p * 4 + 2] = S.ctl_d[kCdTlength];
    B.sd[(size_t)p * 4 + 3] = S.ctl_d[kCdPdg];
  }
}

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
// findperiod update_gtree_common.cpp:1625-1633 (tv carries the TIMEMAX sentinel at [nsplit])
IMA_DEV int findperiod(const DevModel &M, const double *tv, double t) {
  int k = 0;
  while (k < M.nsplit && tv[k] <= t) k++;
  return k;
}
// follow a lineage's population down the population tree until it exists in `period`
IMA_DEV int pop_in_period(const DevModel &M, int pop, int period) {
  while (M.pt_e[pop] <= period && M.pt_e[pop] != -1) pop = M.pt_down[pop];
  return pop;
}
template <class PS> IMA_DEV double edge_top_time(const PS &S, int ng, int e) { return e < ng ? 0.0 : S.time[S.up0[e]]; }

// ------------------------------------------------------------------------------------------------
// migration-path proposal: struct edgemiginfo (imamp.hpp:627-647) with the list held in the pool
// ------------------------------------------------------------------------------------------------
struct Emi {
  int edgeid, pop, temppop, fpop, b, e, mpall;
  double upt, dnt, mtall;
  double mtimeavail[kMaxPeriods];
  int mp[kMaxPeriods];
  int seg, nmig;      // migration list: pool[seg .. seg+nmig)
};

// IMA_reset_edgemiginfo update_gtree_common.cpp:493-514
IMA_DEV void emi_reset(Emi &em) {
  em.edgeid = -1; em.pop = em.temppop = em.fpop = -1; em.b = em.e = -1; em.mpall = 0;
  em.upt = em.dnt = -1.0; em.mtall = 0.0;
  for (int i = 0; i < kMaxPeriods; i++) { em.mtimeavail[i] = 0.0; em.mp[i] = 0; }
  em.seg = 0; em.nmig = 0;
}

// fillmiginfoperiods update_gtree_common.cpp:1100-1162
IMA_DEV void emi_periods(const DevModel &M, const double *tv, Emi &em) {
  const int last = M.nsplit;
  em.b = 0;
  while (em.upt > tv[em.b]) em.b++;
  em.e = em.b;
  while (em.dnt > tv[em.e]) em.e++;
  if (em.e == em.b) {
    em.mtimeavail[em.b] = (em.b == last) ? 0.0 : em.dnt - em.upt;
  } else {
    em.mtimeavail[em.b] = tv[em.b] - em.upt;
    em.mtimeavail[em.e] = (em.e == last) ? 0.0 : em.dnt - tv[em.e - 1];
    for (int i = em.b + 1; i < em.e; i++) em.mtimeavail[i] = tv[i] - tv[i - 1];
  }
  em.mtall = (em.b < last) ? ((tv[last - 1] < em.dnt ? tv[last - 1] : em.dnt) - em.upt) : 0.0;
}

// fillmiginfo for one edge, update_gtree_common.cpp:1169-1246
template <class PS> IMA_DEV void emi_fill_old(const DevModel &M, const double *tv, const PS &S, int ng, int edge, Emi &em) {
  emi_reset(em);
  em.edgeid = edge;
  em.upt = edge_top_time(S, ng, edge);
  em.pop = em.temppop = S.pop[edge];
  em.dnt = S.time[edge];
  em.fpop = S.pop[S.down[edge]];
  emi_periods(M, tv, em);
  em.seg = S.ms[edge];
  em.nmig = S.mcn[edge];
  int j = em.b;
  for (int i = 0; i < em.nmig; i++) {
    while (S.pt[em.seg + i] > tv[j]) j++;
    em.mp[j]++;
    em.mpall++;
  }
}

// poisson utilities.cpp:516-603 (conditioned draws; normal approximation above 100; ADDMIGMAX cap)
IMA_DEV int poisson_cond(Philox &rng, double param, int condition) {
  int i;
  if (param < 0.25 && condition == 1) {
    double u = rng.uniform();
    double sh = sinh(param);
    if (u < param / sh) return 1;
    return (u < param * (6 + param * param) / (6 * sh)) ? 3 : 5;
  }
  if (param < 0.25 && condition == 2) {
    double raised = exp(-param);
    double rcheck = raised = param * raised / (1 - raised);
    double u = rng.uniform();
    i = 1;
    while (u > rcheck && i < kAddMigMax) {
      raised *= param / (i + 1);
      rcheck += raised;
      i++;
    }
    return i;
  }
  bool stop;
  do {
    if (param >= 100.0) {
      double v = rng.normal(param, param) + 0.5;       // POSROUND of normdev(param, param)
      long r = (long)v;
      i = r > 0 ? (r > 1000000 ? 1000000 : (int)r) : 0;
    } else {
      double raised = exp(-param);
      double u = rng.uniform();
      for (i = 0; u > raised; i++) u *= rng.uniform();
    }
    switch (condition) {
      case 0: stop = !(i & 1); break;
      case 1: stop = (i & 1); break;
      case 2: stop = (i != 0); break;
      case 3: stop = (i != 1); break;
      default: stop = true; break;
    }
  } while (!stop);
  if (i > kAddMigMax) i = (condition == 1) ? kAddMigMax - 1 : kAddMigMax;
  return i;
}

// picktopop / picktopop2 update_gtree_common.cpp:1320-1348
IMA_DEV int picktopop(const DevModel &M, Philox &rng, int nowpop, int period, int notother) {
  const int numpops = M.npops - period;
  int t;
  do { t = M.plist[period][rng.randint(numpops)]; } while (t == nowpop || t == notother);
  return t;
}

// simmpath update_gtree_common.cpp:329-398: numm migration times uniform on the period's stretch of
// the edge, sorted, re-drawn on an exact tie; destinations random, the last two constrained when the
// edge must end in `constrainpop`.  Appends to the list of `em` in the pool; false on pool overflow.
template <class PS> IMA_DEV bool simmpath(const DevModel &M, Philox &rng, PS &S, Emi &em, int cap_end, int period, int numm,
                      double timein, double upt, int pop, int constrainpop) {
  const int start = em.seg + em.nmig;
  if (start + numm > cap_end) return false;
  bool dup;
  do {
    for (int i = 0; i < numm; i++) {                   // insertion sort while drawing (hpsortmig :625)
      double t = upt + rng.uniform() * timein;
      int j = start + i;
      while (j > start && S.pt[j - 1] > t) { S.pt[j] = S.pt[j - 1]; j--; }
      S.pt[j] = t;
    }
    dup = false;
    for (int i = start; i + 1 < start + numm; i++) if (S.pt[i] == S.pt[i + 1]) { dup = true; break; }
  } while (dup);
  int lastpop = pop;
  const int lastm = start + numm - 1;
  for (int i = start; i <= lastm; i++) {
    int to;
    if (constrainpop >= 0 && i >= lastm - 1) to = (i == lastm - 1) ? picktopop(M, rng, lastpop, period, constrainpop) : constrainpop;
    else to = picktopop(M, rng, lastpop, period, -1);
    S.pp[i] = (short)to;
    lastpop = to;
  }
  em.nmig += numm;
  return true;
}

// one period of one edge, shared by mwork_single_edge (:1353-1427) and mwork_two_edges (:1430-1575)
// mode: 0 = free period (any count), 1 = last period of the edge (count conditioned on ending in fpop)
template <class PS> IMA_DEV bool mwork_period(const DevModel &M, Philox &rng, PS &S, Emi &em, const Emi &oldem, int cap_end, int periodi,
                          int mode, double timestart) {
  const double r = calcmrate(oldem.mp[periodi], oldem.mtimeavail[periodi]) * em.mtimeavail[periodi];
  int cond = -1, constrain = -1;
  if (mode == 1) {
    if (M.npops - periodi == 2) cond = (em.temppop == em.fpop) ? 0 : 1;
    else { cond = (em.temppop == em.fpop) ? 3 : 2; constrain = em.fpop; }
  }
  const int k = poisson_cond(rng, r, cond);
  em.mp[periodi] = k;
  if (k > 0) {
    if (!simmpath(M, rng, S, em, cap_end, periodi, k, em.mtimeavail[periodi], timestart, em.temppop, constrain)) return false;
    em.mpall += k;
    em.temppop = S.pp[em.seg + em.nmig - 1];
  }
  return true;
}

// mwork_single_edge update_gtree_common.cpp:1353-1427
template <class PS> IMA_DEV bool mwork_single_edge(const DevModel &M, const double *tv, Philox &rng, PS &S, Emi &em, const Emi &oldem,
                               int cap_end, int lastmigperiod) {
  if (lastmigperiod < em.b) return true;
  double timestart = em.upt;
  int periodi;
  for (periodi = em.b; periodi <= lastmigperiod; periodi++) {
    em.temppop = pop_in_period(M, em.temppop, periodi);
    if (!mwork_period(M, rng, S, em, oldem, cap_end, periodi, periodi < em.e ? 0 : 1, timestart)) return false;
    timestart = tv[periodi];
  }
  if (em.mtimeavail[periodi] > 0 && M.pt_e[em.temppop] == periodi) em.temppop = M.pt_down[em.temppop];
  return true;
}

// mwork_two_edges update_gtree_common.cpp:1430-1575
template <class PS> IMA_DEV bool mwork_two_edges(const DevModel &M, const double *tv, Philox &rng, PS &S, Emi &ee, Emi &se, const Emi &oe,
                             const Emi &os, int cap_e, int cap_s, int lastmigperiod) {
  double tstart[2] = { ee.upt, se.upt };
  const int b = ee.b < se.b ? ee.b : se.b;
  const int lastperiodi = (ee.e == M.nsplit) ? lastmigperiod : lastmigperiod - 1;
  int periodi;
  for (periodi = b; periodi <= lastperiodi; periodi++)
    for (int ii = 0; ii < 2; ii++) {
      Emi &mm = ii == 0 ? ee : se;
      const Emi &oldmm = ii == 0 ? oe : os;
      if (mm.b <= periodi) {
        mm.temppop = pop_in_period(M, mm.temppop, periodi);
        if (!mwork_period(M, rng, S, mm, oldmm, ii == 0 ? cap_e : cap_s, periodi, 0, tstart[ii])) return false;
        tstart[ii] = tv[periodi];
      }
    }
  if (periodi == M.nsplit) {
    ee.fpop = se.fpop = M.rootpop;
    return true;
  }
  // both edges end in this period: choose the population in which they join (:1486-1524)
  if (M.pt_e[ee.temppop] == periodi) ee.temppop = M.pt_down[ee.temppop];
  if (M.pt_e[se.temppop] == periodi) se.temppop = M.pt_down[se.temppop];
  int f;
  if (ee.temppop == se.temppop) {
    f = (rng.uniform() < kMigCloseFrac) ? ee.temppop : picktopop(M, rng, ee.temppop, periodi, -1);
  } else if (M.npops - periodi == 2) {
    f = (rng.uniform() < 0.5) ? ee.temppop : se.temppop;
  } else if (rng.uniform() < kMigCloseFrac) {
    f = (rng.uniform() < 0.5) ? ee.temppop : se.temppop;
  } else {
    f = picktopop(M, rng, ee.temppop, periodi, se.temppop);
  }
  ee.fpop = se.fpop = f;
  for (int ii = 0; ii < 2; ii++) {
    Emi &mm = ii == 0 ? ee : se;
    const Emi &oldmm = ii == 0 ? oe : os;
    if (!mwork_period(M, rng, S, mm, oldmm, ii == 0 ? cap_e : cap_s, periodi, 1, tstart[ii])) return false;
    if (mm.mtimeavail[periodi] > 0 && M.pt_e[mm.temppop] == periodi) mm.temppop = M.pt_down[mm.temppop];
  }
  return true;
}

// last-period term of getmprob (update_gtree_common.cpp:879-939 and :1007-1065)
template <class PS> IMA_DEV double getmprob_last(const DevModel &M, const PS &S, const Emi &mm, const Emi &oldmm, int cm) {
  const int e = mm.e;
  const double r = calcmrate(oldmm.mp[e], oldmm.mtimeavail[e]) * mm.mtimeavail[e];
  const int k = mm.mp[e];
  if (e == M.nsplit - 1)
    return k * log(r / mm.mtimeavail[e]) - ((k & 1) ? mylogsinh(r) : mylogcosh(r));
  int pop = (cm == 0) ? mm.pop : (int)S.pp[mm.seg + cm - 1];
  while (M.pt_e[pop] <= e) pop = M.pt_down[pop];
  const int topop = mm.fpop, popc = M.npops - e - 1;
  const double d = (pop == topop) ? log(1 - r * exp(-r)) : log(1 - exp(-r));
  double n;
  if (k == 0) n = -r;
  else if (k == 1) n = log(r / mm.mtimeavail[e]) - r;
  else {
    const int lastm_2_pop = (k == 2) ? pop : (int)S.pp[mm.seg + mm.mpall - 3];
    const double pathc = (lastm_2_pop == topop) ? -log((double)popc) : -log((double)popc - 1);
    n = k * log(r / mm.mtimeavail[e]) - r + (2 - k) * log((double)popc) + pathc;
  }
  return n - d;
}

// getmprob update_gtree_common.cpp:850-1071: log probability of having simulated the lists of
// (edgem, sisem) given the migration counts of (oldedgem, oldsisem)
template <class PS> IMA_DEV double getmprob(const DevModel &M, const double *tv, const PS &S, const Emi &edgem, const Emi &sisem,
                        const Emi &oldedgem, const Emi &oldsisem) {
  double tempp = 0.0;
  const int last = M.nsplit, npops = M.npops;
  const int lastmigrationperiod = edgem.e < last - 1 ? edgem.e : last - 1;
  if (sisem.mtall <= 0) {
    int cm = 0;
    for (int p = edgem.b; p <= edgem.e; p++)
      if (p < lastmigrationperiod || (p == lastmigrationperiod && edgem.e == last)) {
        const double r = calcmrate(oldedgem.mp[p], oldedgem.mtimeavail[p]) * edgem.mtimeavail[p];
        tempp += edgem.mp[p] * log(r / (edgem.mtimeavail[p] * (npops - (p + 1)))) - r;
        cm += edgem.mp[p];
      }
    if (edgem.e < last) tempp += getmprob_last(M, S, edgem, oldedgem, cm);
    return tempp;
  }
  if (edgem.mtall > 0 && sisem.mtall > 0 && edgem.e < last) {
    int pop[2];
    for (int ii = 0; ii < 2; ii++) {                   // population of each edge where they join (:945-964)
      const Emi &mm = ii == 0 ? edgem : sisem;
      pop[ii] = mm.pop;
      if (mm.mpall > 0 && mm.e > 0) {
        const double t = tv[mm.e - 1];
        int k = -1;
        while (k + 1 < mm.nmig && S.pt[mm.seg + k + 1] < t) k++;
        if (k >= 0) pop[ii] = S.pp[mm.seg + k];
      }
      if (mm.e > 0) while (M.pt_e[pop[ii]] <= mm.e) pop[ii] = M.pt_down[pop[ii]];
    }
    if (pop[0] == pop[1])
      tempp = (pop[0] == edgem.fpop) ? log(kMigCloseFrac) : log((1.0 - kMigCloseFrac) / (double)(npops - edgem.e - 1));
    else if (npops - edgem.e == 2)
      tempp = log(0.5);
    else if (edgem.fpop == pop[0] || edgem.fpop == pop[1])
      tempp = log(0.5 * kMigCloseFrac);
    else
      tempp = log((1.0 - kMigCloseFrac) / (double)(npops - edgem.e - 2));
  }
  for (int ii = 0; ii < 2; ii++) {
    const Emi &mm = ii == 0 ? edgem : sisem;
    const Emi &oldmm = ii == 0 ? oldedgem : oldsisem;
    if (mm.mtall > 0) {
      int cm = 0;
      for (int p = mm.b; p <= mm.e; p++)
        if (p < lastmigrationperiod || (p == lastmigrationperiod && mm.e == last)) {
          const double r = calcmrate(oldmm.mp[p], oldmm.mtimeavail[p]) * mm.mtimeavail[p];
          tempp += mm.mp[p] * log(r / (mm.mtimeavail[p] * (npops - (p + 1)))) - r;
          cm += mm.mp[p];
        }
      if (mm.e < last) tempp += getmprob_last(M, S, mm, oldmm, cm);
    }
  }
  return tempp;
}

// normprob utilities.cpp:463-468
IMA_DEV double log_normprob(double stdev, double val) {
  const double z = val / stdev;
  return log(0.3989422803 * exp(-(z * z) / 2) / stdev);
}

// ------------------------------------------------------------------------------------------------
// the proposal (lane 0): update_gtree.cpp:755-825
// ------------------------------------------------------------------------------------------------
// findjointime update_gtree.cpp:34-76: the time from which two lineages, in populations slidepop and sispop at the
// tops of their edges, are in the same population (no-migration slider)
IMA_DEV double findjointime(const DevModel &M, const double *tv, int slidepop, int sispop, double edgeuptime, double sisuptime) {
  int edgeperiod = findperiod(M, tv, edgeuptime), sisperiod = findperiod(M, tv, sisuptime);
  while (edgeperiod < sisperiod) {
    edgeperiod++;
    if (slidepop == M.droppops[edgeperiod][0] || slidepop == M.droppops[edgeperiod][1]) slidepop = M.pt_down[slidepop];
  }
  while (sisperiod < edgeperiod) {
    sisperiod++;
    if (sispop == M.droppops[sisperiod][0] || sispop == M.droppops[sisperiod][1]) sispop = M.pt_down[sispop];
  }
  while (slidepop != sispop) {
    edgeperiod++;
    if (slidepop == M.droppops[edgeperiod][0] || slidepop == M.droppops[edgeperiod][1]) slidepop = M.pt_down[slidepop];
    if (sispop == M.droppops[edgeperiod][0] || sispop == M.droppops[edgeperiod][1]) sispop = M.pt_down[sispop];
  }
  return edgeperiod == 0 ? 0.0 : tv[edgeperiod - 1];
}

template <class PS> IMA_DEV void propose_move(const DevModel &M, const double *tv, int ng, int nl, Philox &rng, PS &S) {
  int root = S.ctl_i[kCiRoot];
  double roottime = S.ctl_d[kCdRoottime];
  uint32_t flags = 0;
  int edge;
  do { edge = rng.randint(nl); } while (S.down[edge] == -1);                       // :756-759
  const int freed = S.down[edge];
  const int oldsis = (S.up0[freed] == edge) ? S.up1[freed] : S.up0[freed];
  Emi oe, os, ne, ns;
  emi_fill_old(M, tv, S, ng, edge, oe);                                             // :765-772
  if (freed == root) emi_fill_old(M, tv, S, ng, oldsis, os); else emi_reset(os);
  S.mcn[edge] = 0;                                                                  // :777
  const double slidestdv = fmin(kSlideStdvMax, roottime / 3);                      // :782-783
  const double holdslidedist = rng.normal(0.0, slidestdv);
  double slidedist = holdslidedist;

  // joinsisdown :374-454 -- sister swallows the freed edge (lists concatenated in the join scratch)
  // the scratch part of the pool is handed out as needed: the joined list first, the rest in halves to the two new lists
  int rootmove, tmrca = 0, pool_edge = S.pool_free;
  {
    const int n1 = S.mcn[oldsis], n2 = S.mcn[freed];
    if (n1 > 0 && n2 > 0) {
      const int J = S.pool_free;
      if (J + n1 + n2 > S.pool_end) flags |= kFlagOverflow;
      else {
        pool_edge = J + n1 + n2;
        for (int i = 0; i < n1; i++) { S.pt[J + i] = S.pt[S.ms[oldsis] + i]; S.pp[J + i] = S.pp[S.ms[oldsis] + i]; }
        for (int i = 0; i < n2; i++) { S.pt[J + n1 + i] = S.pt[S.ms[freed] + i]; S.pp[J + n1 + i] = S.pp[S.ms[freed] + i]; }
        S.ms[oldsis] = (unsigned short)J;
        S.mcn[oldsis] = (unsigned short)(n1 + n2);
      }
    } else if (n2 > 0) { S.ms[oldsis] = S.ms[freed]; S.mcn[oldsis] = (unsigned short)n2; }
    S.time[oldsis] = S.time[freed];
    const int dd = S.down[freed];
    S.ctl_i[kCiOldDownDown] = dd;
    S.down[oldsis] = (short)dd;
    if (dd != -1) {
      rootmove = 0;
      if (S.up0[dd] == freed) S.up0[dd] = (short)oldsis; else S.up1[dd] = (short)oldsis;
    } else {
      rootmove = 1; tmrca++;
      root = oldsis;
      S.time[oldsis] = kTimeMax;
      S.mcn[oldsis] = 0;
      roottime = edge_top_time(S, ng, oldsis);
    }
  }
  // slider :255-372 as a loop (the reference recurses)
  int newsis = oldsis;
  double tp = S.time[edge];
  for (int iter = 0;; iter++) {
    if (iter > 100000) { flags |= kFlagOverflow; break; }
    if (slidedist < 0 && M.nomigration) {
      // slider_nomigration :78-253: without migration the sliding edge may only meet a sister that is in its own
      // population, so the upper limit is also bounded by the time the two populations join
      slidedist = -slidedist;
      const double edgeuptime = edge_top_time(S, ng, edge), sisuptime = edge_top_time(S, ng, newsis);
      const int slidepop = S.pop[edge], sispop = S.pop[newsis];
      const double popjointime = slidepop != sispop ? findjointime(M, tv, slidepop, sispop, edgeuptime, sisuptime) : 0.0;
      if (popjointime > edgeuptime && popjointime > sisuptime) {
        if (slidedist < tp - popjointime) { tp -= slidedist; break; }
        slidedist -= tp - popjointime; tp = popjointime;                              // reflect, continue downwards
      } else if (sisuptime == 0 || edgeuptime >= sisuptime) {
        if (slidedist < tp - edgeuptime) { tp -= slidedist; break; }
        slidedist -= tp - edgeuptime; tp = edgeuptime;
      } else {
        if (slidedist < tp - sisuptime) { tp -= slidedist; break; }
        slidedist -= tp - sisuptime; tp = sisuptime;
        newsis = rng.bit() ? S.up0[newsis] : S.up1[newsis];
        slidedist = -slidedist;                                                      // keep going up
      }
    } else if (slidedist < 0) {
      slidedist = -slidedist;
      const double uplimit = edge_top_time(S, ng, edge);
      const int su = S.up0[newsis];
      if (su == -1 || uplimit >= S.time[su]) {
        if (slidedist < tp - uplimit) { tp -= slidedist; break; }
        slidedist -= tp - uplimit; tp = uplimit;                                     // reflect, continue downwards
      } else {
        const double sistop = S.time[su];
        if (slidedist < tp - sistop) { tp -= slidedist; break; }
        slidedist -= tp - sistop; tp = sistop;
        newsis = rng.bit() ? S.up0[newsis] : S.up1[newsis];
        slidedist = -slidedist;                                                      // keep going up
      }
    } else {
      if (S.down[newsis] == -1 || tp + slidedist < S.time[newsis]) {
        tp += slidedist;
        if (tp >= kTimeMax) tp = kTimeMax;
        break;
      }
      slidedist -= S.time[newsis] - tp; tp = S.time[newsis];
      if (rng.bit()) newsis = S.down[newsis];
      else {
        const int dn = S.down[newsis];
        newsis = (S.up0[dn] == newsis) ? S.up1[dn] : S.up0[dn];
        slidedist = -slidedist;
      }
    }
  }
  S.time[edge] = tp;
  const int topol = (oldsis != newsis);
  // splitsisdown :456-528
  {
    const double curt = tp;
    S.time[freed] = S.time[newsis];
    S.time[newsis] = curt;
    const int dd = S.down[newsis];
    if (dd != -1) {
      if (S.up0[dd] == newsis) S.up0[dd] = (short)freed; else S.up1[dd] = (short)freed;
    } else {
      root = freed; roottime = curt; rootmove = 1;
    }
    S.down[freed] = (short)dd;
    int i = 0;
    const int n = S.mcn[newsis], s0 = S.ms[newsis];
    while (i < n && S.pt[s0 + i] < curt) i++;
    int nowpop = (i > 0) ? (int)S.pp[s0 + i - 1] : (int)S.pop[newsis];
    nowpop = pop_in_period(M, nowpop, findperiod(M, tv, curt));
    S.pop[freed] = (short)nowpop;
    if (dd != -1) { S.ms[freed] = (unsigned short)(s0 + i); S.mcn[freed] = (unsigned short)(n - i); }
    else S.mcn[freed] = 0;
    S.mcn[newsis] = (unsigned short)i;
    S.down[newsis] = S.down[edge] = (short)freed;
    S.up0[freed] = (short)newsis; S.up1[freed] = (short)edge;
  }
  double slideweight = 0.0;                                                         // :803-812
  if (rootmove) slideweight = -log_normprob(slidestdv, holdslidedist) + log_normprob(fmin(kSlideStdvMax, roottime / 3), holdslidedist);

  // addmigration :588-665
  double migweight = 0.0;
  if (!M.nomigration && !(flags & kFlagOverflow)) {                                 // :815-822 (no migration: nothing to simulate)
    emi_reset(ne); emi_reset(ns);
    ne.edgeid = edge;
    ne.upt = edge_top_time(S, ng, edge);
    ne.fpop = S.pop[freed];
    ne.pop = ne.temppop = S.pop[edge];
    ne.dnt = S.time[edge];
    const int pool_sis = pool_edge + (S.pool_end - pool_edge) / 2;
    ne.seg = pool_edge; ne.nmig = 0;
    emi_periods(M, tv, ne);
    bool ok = true;
    const int lastmigperiod = ne.e < M.nsplit - 1 ? ne.e : M.nsplit - 1;
    if (freed == root) {
      ne.fpop = -1;
      ns.edgeid = newsis;
      ns.upt = edge_top_time(S, ng, newsis);
      ns.fpop = -1;
      ns.pop = ns.temppop = S.pop[newsis];
      ns.dnt = S.time[newsis];
      ns.seg = pool_sis; ns.nmig = 0;
      emi_periods(M, tv, ns);
      // getm :536-568
      if (ne.mtall <= 0) ok = mwork_single_edge(M, tv, rng, S, ns, os, S.pool_end, lastmigperiod);
      else ok = mwork_two_edges(M, tv, rng, S, ne, ns, oe, os, pool_sis, S.pool_end, lastmigperiod);
    } else {
      ns.seg = pool_sis;
      ok = mwork_single_edge(M, tv, rng, S, ne, oe, pool_sis, lastmigperiod);
    }
    if (!ok) flags |= kFlagOverflow;
    else {
      const double fwd = getmprob(M, tv, S, ne, ns, oe, os);
      const double rev = getmprob(M, tv, S, oe, os, ne, ns);
      migweight = rev - fwd;                                                         // :654-664
      // copynewmig_to_gtree update_gtree_common.cpp:1250-1289
      S.ms[edge] = (unsigned short)ne.seg; S.mcn[edge] = (unsigned short)ne.nmig;
      if (ns.edgeid >= 0) {
        S.ms[newsis] = (unsigned short)ns.seg; S.mcn[newsis] = (unsigned short)ns.nmig;
        int pop = ns.nmig > 0 ? (int)S.pp[ns.seg + ns.nmig - 1] : (int)S.pop[newsis];
        pop = pop_in_period(M, pop, findperiod(M, tv, S.time[newsis]));
        S.pop[S.down[newsis]] = (short)pop;
      }
    }
  }
  if (topol) flags |= kFlagTopol;
  if (tmrca) flags |= kFlagTmrca;
  S.ctl_i[kCiRoot] = root;
  S.ctl_i[kCiFlags] = (int)flags;
  S.ctl_i[kCiEdge] = edge; S.ctl_i[kCiFreed] = freed; S.ctl_i[kCiOldsis] = oldsis; S.ctl_i[kCiNewsis] = newsis;
  S.ctl_d[kCdRoottime] = roottime;
  S.ctl_d[kCdMigw] = migweight;
  S.ctl_d[kCdSlidew] = slideweight;
  S.ctl_d[kCdSlideDist] = holdslidedist;
}

// ------------------------------------------------------------------------------------------------
// treeweight: update_gtree_common.cpp:1679-1931
// ------------------------------------------------------------------------------------------------
// packed event: bits 0-1 kind (0 coalescence, 1 migration, 2 population split), 2-6 pop, 7-11 topop, 12.. node
constexpr int kRankSortMax = 128;
IMA_DEV long long dbl_bits(double x) {
#if IMA_CUDA
  return __double_as_longlong(x);
#else
  long long v; memcpy(&v, &x, sizeof v); return v;
#endif
}
IMA_DEV double bits_dbl(long long v) {
#if IMA_CUDA
  return __longlong_as_double(v);
#else
  double x; memcpy(&x, &v, sizeof x); return x;
#endif
}
IMA_DEV int pack_event(int kind, int pop, int topop, int node) { return kind | (pop << 2) | (topop << 7) | (node << 12); }

#if IMA_CUDA
// Sorts the nev <= 32 NQ events (time bits, info) of the scratch table by (time, info) into (evt, evi): see eval_weights.
// By counting: element q * 32 + lane sits in register q of the lane; every lane reads every event of the table (a broadcast
// read of shared memory: the reads do not depend on each other, so they pipeline) and counts the events that sort before its
// own.  A bitonic network over the same registers takes 15 to 28 DEPENDENT stages of shuffles and measured six times this;
// unrolled it was also half of the code of every kernel that weighs a genealogy.
template <int NQ> IMA_DEV void warp_sort_events(const double *bt, const int *bi, int nev, double *evt, int *evi) {
  const int lane = Warp::lane();
  const long long *kb = (const long long *)bt;             // times are not negative: their bit patterns order like the times
  long long key[NQ];
  int val[NQ], rank[NQ];
#pragma unroll
  for (int q = 0; q < NQ; q++) {
    const int g = q * 32 + lane;
    key[q] = g < nev ? kb[g] : 0x7fffffffffffffffll;
    val[q] = g < nev ? bi[g] : 0x7fffffff;
    rank[q] = 0;
  }
#pragma unroll 4
  for (int s = 0; s < nev; s++) {
    const long long ok = kb[s];
    const int ov = bi[s];
#pragma unroll
    for (int q = 0; q < NQ; q++) {
      const int g = q * 32 + lane;
      // no short-circuit: the lanes disagree on every one of these, a branch each would serialise them
      rank[q] += (int)((ok < key[q]) | ((ok == key[q]) & ((ov < val[q]) | ((ov == val[q]) & (s < g)))));
    }
  }
#pragma unroll
  for (int q = 0; q < NQ; q++) {
    const int g = q * 32 + lane;
    if (g < nev) { evt[rank[q]] = __longlong_as_double(key[q]); evi[rank[q]] = val[q]; }
  }
}
#endif

// returns false when the event table does not fit (flagged as overflow by the caller)
IMA_DEV bool eval_weights(const DevModel &M, const EngineDims &d, const DevLocus &L, const double *tv, PairSm &S, int evcap = -1) {
  const int lane = Warp::lane();
  const int ng = L.ng, nl = L.nl;
  const int mignum = scan_mig_counts(nl, S);
  const double roottime = S.ctl_d[kCdRoottime];
  const int nsplitev = findperiod(M, tv, roottime);
  const int nev = (ng - 1) + mignum + nsplitev;
  if (nev > (evcap < 0 ? d.EVP : evcap)) return false;
  // Events sorted by (time, info) -- the reference: indexx quicksort, utilities.cpp:709-793.  Up to kRankSortMax events (every
  // genealogy of the shipped inputs) by counting: the events are built into scratch (the prefix table and the period table are
  // not in use yet), every lane counts, for its own events, the events that sort before them -- independent broadcast reads,
  // no barrier -- and writes them to their places.  Larger tables by a bitonic network in place.
  IMA_SPROF(S)
  const bool small = nev <= kRankSortMax;
  double *bt = small ? (double *)S.pre : S.evt;
  int *bi = small ? S.evk : S.evi;
  // event build (:1741-1799): lane per edge
  for (int i = lane; i < nl; i += IMA_WARP) {
    int nowpop = S.pop[i];
    if (i >= ng) {
      bt[i - ng] = S.time[S.up0[i]];
      bi[i - ng] = pack_event(0, nowpop, 0, i);
    }
    const int n = S.mcn[i], s0 = S.ms[i], o = (ng - 1) + S.moff[i];
    for (int j = 0; j < n; j++) {
      const double t = S.pt[s0 + j];
      nowpop = pop_in_period(M, nowpop, findperiod(M, tv, t));
      const int topop = S.pp[s0 + j];
      bt[o + j] = t;
      bi[o + j] = pack_event(1, nowpop, topop, 0);
      nowpop = topop;
    }
  }
  for (int i = lane; i < nsplitev; i += IMA_WARP) {
    bt[(ng - 1) + mignum + i] = tv[i];
    bi[(ng - 1) + mignum + i] = pack_event(2, 0, 0, i);
  }
  IMA_SPROF(S)
  if (small) {
#if IMA_CUDA
    __threadfence_block();
#endif
    Warp::sync();
    // times are not negative, so their bit patterns order like the times: integer compares (the FP64 pipe is narrow)
#if IMA_CUDA
    if (nev <= 32) warp_sort_events<1>(bt, bi, nev, S.evt, S.evi);
    else if (nev <= 64) warp_sort_events<2>(bt, bi, nev, S.evt, S.evi);
    else warp_sort_events<4>(bt, bi, nev, S.evt, S.evi);
#else
    for (int j0 = lane; j0 < nev; j0 += 2 * IMA_WARP) {            // two events of the lane share every read of the table
      const int j1 = j0 + IMA_WARP;
      const bool two = j1 < nev;
      const long long k0 = dbl_bits(bt[j0]), k1 = two ? dbl_bits(bt[j1]) : 0;
      const int i0 = bi[j0], i1 = two ? bi[j1] : 0;
      int r0 = 0, r1 = 0;
      for (int k = 0; k < nev; k++) {
        const long long kk = dbl_bits(bt[k]);
        const int ik = bi[k];
        r0 += (kk < k0 || (kk == k0 && (ik < i0 || (ik == i0 && k < j0)))) ? 1 : 0;
        r1 += (kk < k1 || (kk == k1 && (ik < i1 || (ik == i1 && k < j1)))) ? 1 : 0;
      }
      S.evt[r0] = bt[j0]; S.evi[r0] = i0;
      if (two) { S.evt[r1] = bt[j1]; S.evi[r1] = i1; }
    }
#endif
#if IMA_CUDA
    __threadfence_block();
#endif
    Warp::sync();
  } else {
    int np2 = 1;
    while (np2 < nev) np2 <<= 1;
    for (int i = nev + lane; i < np2; i += IMA_WARP) { S.evt[i] = DBL_MAX; S.evi[i] = 0x7fffffff; }
    Warp::sync();
    for (int k = 2; k <= np2; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = lane; i < np2; i += IMA_WARP) {
          const int x = i ^ j;
          if (x > i) {
            const double a = S.evt[i], b = S.evt[x];
            const int ia = S.evi[i], ib = S.evi[x];
            const bool gt = (a > b) || (a == b && ia > ib);
            if (gt == ((i & k) == 0)) { S.evt[i] = b; S.evt[x] = a; S.evi[i] = ib; S.evi[x] = ia; }
          }
        }
        Warp::sync();
      }
  }
  IMA_SPROF(S)
  for (int i = lane; i < d.NI; i += IMA_WARP) S.gwi[i] = 0;
  for (int i = lane; i < d.ND; i += IMA_WARP) S.gwd[i] = 0.0;
  Warp::sync();
  // ---- sweep (:1805-1921), lane-parallel --------------------------------------------------------
  // The reference walks the sorted events keeping lineage counts n[pop].  Here every event j gets its
  // own view of the state just before it from exclusive prefix scans: the period k_j (splits before j),
  // the number of lineages (coalescences before j) and, per tree population q, the running sum D_q of
  // the events' +-1 effects.  A population that exists in period k holds
  //     n = sum over q in desc(pop) of (samples_q + D_q)
  // lineages (populations merge at splits and receive no events afterwards), so no sequential state is
  // needed.  Integer counts are bit-exact; the double sums are added lane-wise then by a fixed-order
  // warp reduction (the reference adds them in time order: differences are rounding only).
  const int W64 = (M.ntreepops + 3) >> 2;              // 16-bit fields, four populations per 64-bit word
  {
    int carry_k = 0, carry_c = 0;
    unsigned long long carry_w[(kMaxTreePops + 3) / 4];
    for (int w = 0; w < W64; w++) {                      // the scan starts from the samples (sampled populations only)
      unsigned long long c0 = 0ull;
      for (int f = 0; f < 4; f++) { const int q = w * 4 + f; if (q < M.npops) c0 |= (unsigned long long)L.samppop[q] << (16 * f); }
      carry_w[w] = c0;
    }
    for (int base = 0; base < nev; base += IMA_WARP) {
      const int j = base + lane;
      const bool valid = j < nev;
      const int info = valid ? S.evi[j] : 0;
      const int kind = valid ? (info & 3) : 3, ip = (info >> 2) & 31, jp = (info >> 7) & 31;
      const int is_split = (kind == 2), is_coal = (kind == 0);
      const int inc_k = Warp::scan(is_split), inc_c = Warp::scan(is_coal);
      const int kj = carry_k + inc_k - is_split, cj = carry_c + inc_c - is_coal;
      carry_k += Warp::bcast(inc_k, IMA_WARP - 1);
      carry_c += Warp::bcast(inc_c, IMA_WARP - 1);
      if (valid) S.evk[j] = kj | ((ng - cj) << 8);
      for (int w = 0; w < W64; w++) {
        unsigned long long own = 0ull;
        if (valid) {
          for (int f = 0; f < 4; f++) {
            const int q = w * 4 + f;
            if (q < M.ntreepops) {
              int dq = 1;                                // biased by +1 so that every field stays non-negative
              if (kind == 0 && ip == q) dq -= 1;
              if (kind == 1) { if (ip == q) dq -= 1; if (jp == q) dq += 1; }
              own |= (unsigned long long)dq << (16 * f);
            }
          }
        }
        unsigned long long inc = own;
#if IMA_CUDA
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        const unsigned long long tot = __shfl_sync(0xffffffffu, inc, 31);
#else
        const unsigned long long tot = inc;
#endif
        if (valid) S.pre[(size_t)j * W64 + w] = carry_w[w] + inc - own;
        carry_w[w] += tot;
      }
    }
  }
  Warp::sync();
  IMA_SPROF(S)
  // number of lineages in tree population `pop` just before event j (the samples are in the scan's starting value)
  const int *tab = d.tab;
  const int ntp = M.ntreepops, npops = M.npops, nsplit = M.nsplit, ncc = M.ncc;
  auto lineages = [&](int pop, int j) {
    int n = 0;
    const unsigned dm = (unsigned)tab[kTabDesc + pop];
    for (int q = 0; q < ntp; q++)
      if (dm & (1u << q)) n += (int)((S.pre[(size_t)j * W64 + (q >> 2)] >> (16 * (q & 3))) & 0xffffull) - j;   // remove the +1 bias
    return n;
  };
  const double h2term = L.h2term;
  const double lastsplitt = nsplit > 0 ? tv[nsplit - 1] : kTimeMax;
  double length = 0.0, tlength = 0.0;
  int bad = 0;
  for (int j = lane; j < nev; j += IMA_WARP) {
    const double t = S.evt[j], lasttime = j > 0 ? S.evt[j - 1] : 0.0;
    const double dt = t - lasttime;
    const int ek = S.evk[j], kj = ek & 0xff, nsum = ek >> 8;
    const double timeadd = nsum * dt;
    length += timeadd;
    if (t < lastsplitt) tlength += timeadd;
    else if (lasttime < lastsplitt) tlength += nsum * (lastsplitt - lasttime);
    const int info = S.evi[j], kind = info & 3, ip = (info >> 2) & 31, jp = (info >> 7) & 31;
    const int np = npops - kj;
    const int *pl = tab + kTabPlist + kj * kMaxPops;
    if (kind == 0) {
      int ii = 0;
      while (ii < np && pl[ii] != ip) ii++;
      if (ii >= np || lineages(ip, j) < 2) bad = 1;
      else {
#if IMA_CUDA
        atomicAdd(&S.gwi[tab[kTabCcOff + kj] + ii], 1);
#else
        S.gwi[tab[kTabCcOff + kj] + ii]++;
#endif
      }
    } else if (kind == 1) {
      int ii = 0, jj = 0;
      while (ii < np && pl[ii] != ip) ii++;
      while (jj < np && pl[jj] != jp) jj++;
      if (ii >= np || jj >= np || kj >= nsplit || lineages(ip, j) < 1) bad = 1;
      else {
#if IMA_CUDA
        atomicAdd(&S.gwi[ncc + tab[kTabMcOff + kj] + ii * np + jj], 1);
#else
        S.gwi[ncc + tab[kTabMcOff + kj] + ii * np + jj]++;
#endif
      }
    }
  }
  bad = Warp::any(bad != 0) ? 1 : 0;
  // fc[k][ii] += n(n-1) dt / (2h) and fm[k][ii][*] += n dt.  Two targets (k, ii), (k, ii + 1) per pass over the events, and
  // the reductions of a pass issued together: independent shuffle chains overlap, one after another they were a quarter of
  // the sweep.  Every sum adds the same numbers in the same order as the one-target-at-a-time loop.
  bool first_pass = true;
  for (int k = 0; k <= nsplit; k++)
    for (int ii = 0; ii < npops - k; ii += 2) {
      const bool two = ii + 1 < npops - k;
      const int ip0 = tab[kTabPlist + k * kMaxPops + ii], ip1 = two ? tab[kTabPlist + k * kMaxPops + ii + 1] : ip0;
      double fc0 = 0.0, fm0 = 0.0, fc1 = 0.0, fm1 = 0.0;
      for (int j = lane; j < nev; j += IMA_WARP)
        if ((S.evk[j] & 0xff) == k) {
          const double dt = S.evt[j] - (j > 0 ? S.evt[j - 1] : 0.0);
          const int n0 = lineages(ip0, j);
          fc0 += ((double)n0 * ((double)n0 - 1)) * dt * h2term;
          fm0 += n0 * dt;
          if (two) {
            const int n1 = lineages(ip1, j);
            fc1 += ((double)n1 * ((double)n1 - 1)) * dt * h2term;
            fm1 += n1 * dt;
          }
        }
      if (first_pass) {                                   // the two lengths ride along with the first pass
        Warp::sum4(fc0, fm0, length, tlength);
        if (two) Warp::sum2(fc1, fm1);
        first_pass = false;
      } else if (two) Warp::sum4(fc0, fm0, fc1, fm1);
      else Warp::sum2(fc0, fm0);
      if (lane == 0) {
        S.gwd[wd_fc(M, k, ii)] = fc0;
        if (two) S.gwd[wd_fc(M, k, ii + 1)] = fc1;
        if (!M.nomigration && k < nsplit)
          for (int jj = 0; jj < npops - k; jj++) {
            if (jj != ii) S.gwd[wd_fm(M, k, ii, jj)] = fm0;
            if (two && jj != ii + 1) S.gwd[wd_fm(M, k, ii + 1, jj)] = fm1;
          }
      }
    }
  Warp::sync();
  IMA_SPROF(S)
  const double hlog = L.hlog;
  if (hlog != 0.0)
    for (int i = lane; i < M.ncc; i += IMA_WARP) S.gwd[M.ncc + i] += hlog * S.gwi[i];
  if (lane == 0) {
    if (bad) S.ctl_i[kCiFlags] |= (int)kFlagBadTree;
    S.ctl_d[kCdLength] = length;
    S.ctl_d[kCdTlength] = tlength;
    S.ctl_i[kCiMignum] = mignum;
    S.ctl_i[kCiNev] = nev;
  }
  Warp::sync();
  return true;
}

// Subtree tip sets as canonical keys: every tip walks to the root OR-ing its bit into the edges it passes
// (order-independent, so deterministic); a set that contains tip 0 is replaced by its complement, which makes
// "equals the carrier set or its complement" a single comparison against the (canonical) site key.
IMA_DEV void build_tip_keys(const DevLocus &L, PairSm &S) {
  const int lane = Warp::lane();
  const int ng = L.ng, nl = L.nl, W = L.nwords;
  for (int i = lane; i < nl * W; i += IMA_WARP) {
    const int e = i / W, w = i - e * W;
    S.mask[i] = (e < ng && (e >> 5) == w) ? (1u << (e & 31)) : 0u;
  }
  Warp::sync();
  for (int tip = lane; tip < ng; tip += IMA_WARP) {
    const uint32_t bit = 1u << (tip & 31);
    const int w = tip >> 5;
    for (int e = S.down[tip], guard = 0; e != -1 && guard < nl; e = S.down[e], guard++) {
#if IMA_CUDA
      atomicOr(&S.mask[e * W + w], bit);
#else
      S.mask[e * W + w] |= bit;
#endif
    }
  }
  Warp::sync();
  for (int e = lane; e < nl; e += IMA_WARP)
    if (S.mask[e * W] & 1u)
      for (int w = 0; w < W; w++) {
        const uint32_t full = (w == W - 1 && (ng & 31)) ? ((1u << (ng & 31)) - 1u) : 0xffffffffu;
        S.mask[e * W + w] = ~S.mask[e * W + w] & full;
      }
  Warp::sync();
}

// ------------------------------------------------------------------------------------------------
// infinite sites: calc_prob_data.cpp:731-836.  A site is compatible iff exactly one branch carries
// its mutation, i.e. iff some non-root edge's subtree tip set equals the site's carrier set or its
// complement (same accept/reject and same branch as labelgtree's Fitch pass).
// Every lane returns the likelihood (or kRejectIS).
// ------------------------------------------------------------------------------------------------
IMA_DEV double likelihood_is(const EngineView &E, const DevLocus &L, PairSm &S, double mutrate) {
  const int lane = Warp::lane();
  const int ng = L.ng, nl = L.nl, W = L.nwords, root = S.ctl_i[kCiRoot];
  IMA_SPROF(S)
  build_tip_keys(L, S);
  IMA_SPROF(S)
  const uint32_t *sm = E.sitemask + L.sitemask_off;     // canonical site keys (set_locus)
  double acc = 0.0;
  bool reject = false;
  for (int s = lane; s < L.nsites; s += IMA_WARP) {
    int found = -1;
    if (W == 1) {
      const uint32_t key = sm[s];
      for (int b = 0; b < nl; b++) if (S.mask[b] == key) found = b;      // the root's key is 0: never a site key
    } else {
      for (int b = 0; b < nl && found < 0; b++) {
        bool eq = true;
        for (int w = 0; w < W; w++) eq = eq && (S.mask[b * W + w] == sm[(size_t)s * W + w]);
        if (eq) found = b;
      }
    }
    if (found < 0) { reject = true; continue; }
    double ptime = S.time[found] - edge_top_time(S, ng, found);
    const int dn = S.down[found];
    if (dn == root) {                                  // mutation on either root branch: both lengths (:791-801)
      const int a = S.up0[dn], b = S.up1[dn];
      ptime = (S.time[a] - edge_top_time(S, ng, a)) + (S.time[b] - edge_top_time(S, ng, b));
    }
    acc += log(ptime * mutrate);
  }
  reject = Warp::any(reject);
  acc = Warp::sum(acc);
  if (reject) return kRejectIS;
  return -S.ctl_d[kCdLength] * mutrate + acc - L.sumlogk;
}

// ------------------------------------------------------------------------------------------------
// HKY: calc_prob_data.cpp:26-41 (pijt), 118-471 (makefrac), 473-493 (getstandfactor), 583-607.
// Like the reference, every internal node keeps its partial likelihoods (frac) and scale factors, and a proposal recomputes
// only the nodes above the edges it touched (makefrac's e1..e4 rule, :137-164): the new parent of the moved edge, the old
// parent of its old sister, and everything between them and the root.  The reference keeps frac / newfrac per node and
// copies newfrac -> frac on acceptance (copyfraclike, update_gtree_common.cpp:2139-2154); here every node has TWO slots in a
// per-pair slab and every genealogy buffer carries a bit mask saying which slot is the node's current one.  A proposal writes
// the recomputed nodes into their other slots and hands the proposed buffer the mask with those bits flipped: acceptance is the
// buffer flip it always was, rejection leaves the current slots untouched, nothing is ever copied.
// Slab of a pair: [node][slot][5][patterns] doubles (4 partials and the scale of every compressed site pattern; lanes take
// patterns, so every load and store is a coalesced row).  Nodes are visited in coalescence order (children first -- the order
// eval_weights' sorted event list already gives); the 32 lanes first compute the 2 x 16 transition probabilities of the
// node's two child branches (one entry each), then take one compressed site pattern each.
// ------------------------------------------------------------------------------------------------
IMA_DEV double hky_pijt(const double *pi, double mutrate, double t, double kappa, int from, int to) {
  const double PIj = (to == 0 || to == 2) ? pi[0] + pi[2] : pi[1] + pi[3];
  const double A = 1.0 + PIj * (kappa - 1.0);
  if (from == to) return pi[to] + pi[to] * exp(-mutrate * t) * (1.0 / PIj - 1.0) + exp(-mutrate * t * A) * ((PIj - pi[to]) / PIj);
  if (from + to == 2 || from + to == 4) return pi[to] + pi[to] * (1.0 / PIj - 1.0) * exp(-mutrate * t) - (pi[to] / PIj) * exp(-mutrate * t * A);
  return pi[to] * (1.0 - exp(-mutrate * t));
}

// how a call treats the slots: kHkyInit -- every node into slot 0, mask 0 (a state that was just loaded); kHkyFull -- every
// node into its other slot (a rescaled genealogy, new mutation scalar or kappa); kHkyPartial -- the nodes above `freed` and
// `olddd` (the junction node after the move; the parent of the freed node before it, -1 when that was the root)
enum { kHkyInit = 0, kHkyFull = 1, kHkyPartial = 2 };
struct HkyCall { int mode, freed, olddd; const uint32_t *mask_cur; uint32_t *mask_new; };

IMA_DEV double likelihood_hky(const EngineView &E, const DevLocus &L, PairSm &S, int p, double u, double kappa, const double *pi, const HkyCall &hk) {
  const int lane = Warp::lane();
  const int ng = L.ng, nl = L.nl, ns = L.nsites, nev = S.ctl_i[kCiNev], root = S.ctl_i[kCiRoot];
  const size_t hs = (size_t)E.d.hky_sites;
  double *slab = E.hky_frac + (size_t)p * E.d.hky_stride;          // [internal node][slot][5][hky_sites]
  const unsigned char *seq = E.seq + L.seq_off;
  const int *mult = E.mult + L.mult_off;
  double sf = 0.0;                                                 // getstandfactor :473-493
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
      if (i != j) sf += (i + j == 2 || i + j == 4) ? pi[i] * pi[j] * kappa : pi[i] * pi[j];
  const double mu = u / (L.totsites * sf);                         // :593
  // which nodes are recomputed (the tip-mask table of the infinite-sites likelihood is free on an HKY locus): dirty[e] for
  // every internal edge e; then newslot[e] = the slot that holds e's partials in the proposed genealogy
  uint32_t *dirty = S.mask;
  for (int e = lane; e < nl; e += IMA_WARP) dirty[e] = (hk.mode != kHkyPartial && e >= ng) ? 1u : 0u;
#if IMA_CUDA
  __threadfence_block();
#endif
  Warp::sync();
  if (hk.mode == kHkyPartial && lane == 0) {
    for (int e = hk.freed, guard = 0; e >= ng && guard < nl && !dirty[e]; e = S.down[e], guard++) dirty[e] = 1u;
    for (int e = hk.olddd, guard = 0; e >= ng && guard < nl && !dirty[e]; e = S.down[e], guard++) dirty[e] = 1u;
  }
#if IMA_CUDA
  __threadfence_block();
#endif
  Warp::sync();
  auto cur_slot = [&](int e) { return hk.mode == kHkyInit ? 1 : (int)((hk.mask_cur[(e - ng) >> 5] >> ((e - ng) & 31)) & 1u); };
  auto new_slot = [&](int e) { return cur_slot(e) ^ (int)dirty[e]; };
  double *P = (double *)S.pre;                                     // 32 doubles; the prefix table is free after the sweep
  for (int j = 0; j < nev; j++) {
    const int info = S.evi[j];
    if ((info & 3) != 0) continue;
    const int node = info >> 12;
    if (!dirty[node]) continue;
    const int a = S.up0[node], b = S.up1[node];
    const double ta = S.time[a] - edge_top_time(S, ng, a), tb = S.time[b] - edge_top_time(S, ng, b);
    for (int e = lane; e < 32; e += IMA_WARP) P[e] = hky_pijt(pi, mu, (e >> 4) ? tb : ta, kappa, (e >> 2) & 3, e & 3);
    Warp::sync();
    double *out = slab + ((size_t)(node - ng) * 2 + new_slot(node)) * 5 * hs;
    const double *fa = a < ng ? nullptr : slab + ((size_t)(a - ng) * 2 + new_slot(a)) * 5 * hs;
    const double *fb = b < ng ? nullptr : slab + ((size_t)(b - ng) * 2 + new_slot(b)) * 5 * hs;
    for (int s = lane; s < ns; s += IMA_WARP) {
      double v[4], mx = 0.0, ca[4], cb[4];
      if (fa) for (int k = 0; k < 4; k++) ca[k] = fa[k * hs + s];
      if (fb) for (int k = 0; k < 4; k++) cb[k] = fb[k * hs + s];
      for (int from = 0; from < 4; from++) {
        double sa, sb;
        if (!fa) sa = P[from * 4 + seq[(size_t)a * ns + s]];
        else { sa = 0.0; for (int k = 0; k < 4; k++) sa += P[from * 4 + k] * ca[k]; }
        if (!fb) sb = P[16 + from * 4 + seq[(size_t)b * ns + s]];
        else { sb = 0.0; for (int k = 0; k < 4; k++) sb += P[16 + from * 4 + k] * cb[k]; }
        v[from] = sa * sb;
        if (v[from] > mx) mx = v[from];
      }
      for (int k = 0; k < 4; k++) out[k * hs + s] = v[k] / mx;
      out[4 * hs + s] = (fa ? fa[4 * hs + s] : 0.0) + (fb ? fb[4 * hs + s] : 0.0) + log(mx);
    }
#if IMA_CUDA
    __threadfence_block();
#endif
    Warp::sync();
  }
  const double *fr = slab + ((size_t)(root - ng) * 2 + new_slot(root)) * 5 * hs;
  double acc = 0.0;
  for (int s = lane; s < ns; s += IMA_WARP) {
    double fracp = 0.0;
    for (int k = 0; k < 4; k++) fracp += pi[k] * fr[k * hs + s];
    acc += mult[s] * (log(fracp) + fr[4 * hs + s]);
  }
  // the proposed genealogy's slot mask
  const int MW = E.d.hky_mask_words;
  for (int w = lane; w < MW; w += IMA_WARP) {
    uint32_t m = 0u;
    for (int k = 0; k < 32; k++) { const int e = ng + w * 32 + k; if (e < nl && new_slot(e)) m |= 1u << k; }
    hk.mask_new[w] = m;
  }
  return Warp::sum(acc);
}

// finishSWupdateA update_gtree_common.cpp:2218-2362 (lane 0): after the edge has been re-attached, draw the
// allele state of the new junction node around its neighbours (geometric step), refresh the branch terms of the
// (at most four) branches whose ends changed, and return the change of log-likelihood; *aterm is the Hastings
// term of the allele draw.  Aold/dlold: the pair's current arrays; Anew/dlnew: the proposal buffer (already a copy).
IMA_DEV double sw_update_alleles(const DevLocus &L, const PairSm &S, int ai, Philox &rng, const short *Aold, const double *dlold,
                                 short *Anew, double *dlnew, int edge, int downedge, int sisedge, int newsisedge,
                                 int old_downdown, double u, double *aterm) {
  const int ng = L.ng;
  const int oldA = Anew[downedge];                     // the junction keeps its number; its old state is the starting point
  const double holdsis = (newsisedge != sisedge) ? dlold[newsisedge] : 0.0;
  // old terms of edge, sister and (when it was not the root) the freed edge: copyedge[0..2].dlikeA (storeAinfo :822-847)
  const double oldlikeadj = dlold[edge] + dlold[sisedge] + (old_downdown != -1 ? dlold[downedge] : 0.0) + holdsis;
  const int e[3] = { edge, newsisedge, downedge };
  double t[3];
  int wsumdiff = 0, j = 0;
  for (int i = 0; i < 3; i++)
    if (S.down[e[i]] != -1) {
      t[i] = S.time[e[i]] - edge_top_time(S, ng, e[i]);
      const int d = (i < 2 ? Anew[e[i]] : Anew[S.down[e[i]]]) - oldA;
      wsumdiff += d < 0 ? -d : d;
      j++;
    }
  double geonew = j / ((double)(wsumdiff + j));
  if (geonew > 0.95) geonew = 0.95;
  int dA = (int)ceil(log(rng.uniform()) / log(1.0 - geonew)) - 1;           // geometric(p) - 1, utilities.cpp:608-617
  if (rng.bit()) dA = -dA;
  int newA;
  if (dA >= 0) newA = (oldA + dA < L.maxA[ai]) ? oldA + dA : L.maxA[ai];
  else newA = (oldA + dA > L.minA[ai]) ? oldA + dA : L.minA[ai];
  Anew[downedge] = (short)newA;
  dA = newA - oldA;
  if (S.down[sisedge] != -1 && sisedge != newsisedge) {                      // old sister now runs on through the old junction
    const double ts = S.time[sisedge] - edge_top_time(S, ng, sisedge);
    dlnew[sisedge] = -(ts * u) + log(bessi(Anew[sisedge] - Anew[S.down[sisedge]], ts * u));
  } else {
    dlnew[sisedge] = 0.0;
  }
  double likeadj = dlnew[sisedge];
  for (int i = 0; i < 3; i++)
    if (S.down[e[i]] != -1) {
      const int d = (i < 2 ? Anew[e[i]] : Anew[S.down[e[i]]]) - newA;
      dlnew[e[i]] = -(t[i] * u) + log(bessi(d, t[i] * u));
      likeadj += dlnew[e[i]];
    } else {
      dlnew[e[i]] = 0.0;                               // the root edge carries no term (update_gtree.cpp:448-450, 486-488)
    }
  // reverse move: the old junction (children edge and old sister, parent old_downdown) seen from newA
  wsumdiff = 0; j = 0;
  { int d = Aold[edge] - newA; wsumdiff += d < 0 ? -d : d; j++; d = Aold[sisedge] - newA; wsumdiff += d < 0 ? -d : d; j++; }
  if (old_downdown != -1) { const int d = Aold[old_downdown] - newA; wsumdiff += d < 0 ? -d : d; j++; }
  double geoold = j / ((double)(wsumdiff + j));
  if (geoold > 0.95) geoold = 0.95;
  const int adA = dA < 0 ? -dA : dA;
  *aterm = (adA * log(1 - geoold) + log(geoold)) - (adA * log(1 - geonew) + log(geonew));
  return likeadj - oldlikeadj;
}

// stepwise: calc_prob_data.cpp:841-909 (full evaluation of one linked portion); A/dlikeA in global memory
IMA_DEV double likelihood_sw(const DevLocus &L, const PairSm &S, const short *A, double *dlikeA, double u) {
  const int lane = Warp::lane();
  double like = 0.0;
  bool zero = false;
  for (int i = lane; i < L.nl; i += IMA_WARP) {
    double dl = 0.0;
    if (S.down[i] != -1) {
      const int dd = A[i] - A[S.down[i]];
      const double t = S.time[i] - edge_top_time(S, L.ng, i);
      const double bv = bessi(dd, t * u);
      if (!(bv > 0.0)) { zero = true; dl = -1e+100; }
      else dl = -(t * u) + log(bv);
      like += dl;
    }
    dlikeA[i] = dl;
  }
  zero = Warp::any(zero);
  like = Warp::sum(like);
  return zero ? -DBL_MAX : like;
}

}  // namespace ima
