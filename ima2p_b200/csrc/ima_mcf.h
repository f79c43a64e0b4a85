// The reference's MCMC state file (.mcf): writemcf / readmcf, mcmcfile.cpp:203-442.  A stream of records
// "name type count values..." (awrite/aread :31-130; type 0 int "%d ", type 3 double "%.10lg "), chain after chain:
// tvalue x nsplit; per locus uvalue x nlinked, kappavalue (HKY), pi[4]; per edge up[2], down, mut, pop, (A[], dlikeA[]
// for stepwise / joint loci), time, the migration list as mig[].mt (+ mig[].mp while mt > 0, closed by mt = -1), and for
// the internal nodes of an HKY locus two 4-vectors per site pattern (conditional likelihoods; both sides recompute them
// after loading, the reference in init_p).  Included by ima_engine.cu, which owns the Engine.
#pragma once

namespace ima {

struct McfReader {
  FILE *f;
  std::string err;
  bool expect(const char *name, int type, int *count) {
    char got[128];
    int t = -1, n = -1;
    if (fscanf(f, "%127s %d %d ", got, &t, &n) != 3) { err = std::string("mcf file ends before ") + name; return false; }
    if (strcmp(got, name) != 0) { err = std::string("variable names do not match: ") + name + "  <> " + got; return false; }
    if (t != type) { err = std::string("variable types do not match for ") + name; return false; }
    *count = n;
    return true;
  }
  bool ints(const char *name, int want, int *out) {
    int n;
    if (!expect(name, 0, &n)) return false;
    if (n != want) { err = std::string("unexpected count for ") + name; return false; }
    for (int i = 0; i < n; i++) if (fscanf(f, "%d ", out + i) != 1) { err = std::string("bad value in ") + name; return false; }
    return true;
  }
  bool doubles(const char *name, int want, double *out) {
    int n;
    if (!expect(name, 3, &n)) return false;
    if (n != want) { err = std::string("unexpected count for ") + name; return false; }
    for (int i = 0; i < n; i++) if (fscanf(f, "%lg ", out + i) != 1) { err = std::string("bad value in ") + name; return false; }
    return true;
  }
};

static void mcf_ints(FILE *f, const char *name, int n, const int *v) {
  fprintf(f, "%s %d %d ", name, 0, n);
  for (int i = 0; i < n; i++) fprintf(f, "%d ", v[i]);
  fprintf(f, "\n");
}
static void mcf_doubles(FILE *f, const char *name, int n, const double *v) {
  fprintf(f, "%s %d %d ", name, 3, n);
  for (int i = 0; i < n; i++) fprintf(f, "%.10lg ", v[i]);
  fprintf(f, "\n");
}

}  // namespace ima

extern "C" {

// writemcf (mcmcfile.cpp:203-296) for the chains this engine holds, in chain order
int ima2p_engine_write_mcf(ima2p_engine *h, const char *path) {
  if (!h || !h->eng.finalized || !path) return fail(IMA2P_E_ARG, "write_mcf: bad argument");
  Engine &e = h->eng;
  FILE *f = fopen(path, "w");
  if (!f) return fail(IMA2P_E_ARG, "Error creating mcffile");
  const int NL = e.d.NL, CAP = e.d.CAP;
  std::vector<int> up0(NL), up1(NL), down(NL), pop(NL), moff(NL + 1), mp(CAP + 1), A((size_t)kMaxLinked * NL);
  std::vector<double> time(NL), mt(CAP + 1), dl((size_t)kMaxLinked * NL), pa(kMaxLinked);
  int rc = IMA2P_OK;
  for (int ci = 0; ci < e.d.nchains && rc == IMA2P_OK; ci++) {
    double tv[kMaxPeriods];
    if ((rc = ima2p_engine_get_split_times(h, ci, tv)) != IMA2P_OK) break;
    for (int k = 0; k < e.model.nsplit; k++) mcf_doubles(f, "tvalue", 1, &tv[k]);
    for (int li = 0; li < e.d.nloci && rc == IMA2P_OK; li++) {
      const DevLocus &L = e.loci[li].d;
      const size_t p = (size_t)ci * e.d.nloci + li;
      double u[kMaxLinked], kappa = 0.0;
      int root = 0;
      double roottime = 0.0;
      if ((rc = ima2p_engine_get_scalars(h, ci, li, u, &kappa)) != IMA2P_OK) break;
      if ((rc = ima2p_engine_get_genealogy(h, ci, li, 0, up0.data(), up1.data(), down.data(), pop.data(), time.data(), moff.data(), mt.data(),
                                           mp.data(), CAP, &root, &roottime)) != IMA2P_OK) break;
      const bool sw = has_stepwise(L.model);
      if (sw && (rc = ima2p_engine_get_alleles(h, ci, li, 0, A.data(), dl.data(), pa.data())) != IMA2P_OK) break;
      for (int a = 0; a < L.nlinked; a++) mcf_doubles(f, "uvalue", 1, &u[a]);
      if (L.model == kHKY) mcf_doubles(f, "kappavalue", 1, &kappa);
      mcf_doubles(f, "pi[4]", 4, &e.h_pi[p * 4]);
      for (int i = 0; i < L.nl; i++) {
        const int up[2] = {up0[i], up1[i]}, zero = 0;
        mcf_ints(f, "up[2]", 2, up);
        mcf_ints(f, "down", 1, &down[i]);
        mcf_ints(f, "mut", 1, &zero);                       // scratch of the labelling pass in the reference
        mcf_ints(f, "pop", 1, &pop[i]);
        if (sw) {
          int Ai[kMaxLinked]; double di[kMaxLinked];
          for (int a = 0; a < L.nlinked; a++) { Ai[a] = A[(size_t)a * L.nl + i]; di[a] = dl[(size_t)a * L.nl + i]; }
          mcf_ints(f, "A[]", L.nlinked, Ai);
          mcf_doubles(f, "dlikeA[]", L.nlinked, di);
        }
        mcf_doubles(f, "time", 1, &time[i]);
        for (int j = moff[i]; j < moff[i + 1]; j++) { mcf_doubles(f, "mig[].mt", 1, &mt[j]); mcf_ints(f, "mig[].mp", 1, &mp[j]); }
        const double end = -1.0;
        mcf_doubles(f, "mig[].mt", 1, &end);
        if (L.model == kHKY && i >= L.ng) {
          const double z[4] = {0, 0, 0, 0};
          for (int j = 0; j < L.nsites; j++) {
            mcf_doubles(f, "C[ci]->G[li].gtree[i].hkyi.frac[j]", 4, z);
            mcf_doubles(f, "C[ci]->G[li].gtree[i].hkyi.newfrac[j]", 4, z);
          }
        }
      }
    }
  }
  fclose(f);
  return rc;
}

// readmcf (mcmcfile.cpp:310-442): loads the file's chains into this engine's chains in order (when the file holds fewer
// it is read again from the top, as the reference does), uploads and evaluates (init_p)
int ima2p_engine_read_mcf(ima2p_engine *h, const char *path) {
  if (!h || !h->eng.finalized || !path) return fail(IMA2P_E_ARG, "read_mcf: bad argument");
  Engine &e = h->eng;
  McfReader R{fopen(path, "r"), ""};
  if (!R.f) return fail(IMA2P_E_ARG, "Error opening mcffile");
  const int NL = e.d.NL;
  std::vector<int> up0(NL), up1(NL), down(NL), pop(NL), moff(NL + 1), mp, A((size_t)kMaxLinked * NL);
  std::vector<double> time(NL), mt;
  int rc = IMA2P_OK, lastci = -1;
  for (int ci = 0; ci < e.d.nchains && rc == IMA2P_OK; ci++) {
    double tv[kMaxPeriods];
    for (int k = 0; k < e.model.nsplit; k++) if (!R.doubles("tvalue", 1, &tv[k])) { rc = IMA2P_E_ARG; break; }
    if (rc) break;
    if ((rc = ima2p_engine_set_chain(h, ci, tv)) != IMA2P_OK) break;
    for (int li = 0; li < e.d.nloci && rc == IMA2P_OK; li++) {
      const DevLocus &L = e.loci[li].d;
      double u[kMaxLinked] = {1, 1, 1, 1}, kappa = 0.0, pi[4], dummy[kMaxLinked], frac[4];
      int root = -1, mut, Ai[kMaxLinked];
      bool ok = true;
      for (int a = 0; a < L.nlinked && ok; a++) ok = R.doubles("uvalue", 1, &u[a]);
      if (ok && L.model == kHKY) ok = R.doubles("kappavalue", 1, &kappa);
      ok = ok && R.doubles("pi[4]", 4, pi);
      mt.clear(); mp.clear();
      for (int i = 0; i < L.nl && ok; i++) {
        int up[2];
        ok = R.ints("up[2]", 2, up) && R.ints("down", 1, &down[i]) && R.ints("mut", 1, &mut) && R.ints("pop", 1, &pop[i]);
        if (!ok) break;
        up0[i] = up[0]; up1[i] = up[1];
        if (down[i] == -1) root = i;
        if (has_stepwise(L.model)) {
          ok = R.ints("A[]", L.nlinked, Ai) && R.doubles("dlikeA[]", L.nlinked, dummy);
          for (int a = 0; a < L.nlinked && ok; a++) A[(size_t)a * L.nl + i] = Ai[a];
        }
        ok = ok && R.doubles("time", 1, &time[i]);
        moff[i] = (int)mt.size();
        for (; ok;) {
          double t;
          int pp;
          ok = R.doubles("mig[].mt", 1, &t);
          if (!ok || !(t > 0)) break;
          ok = R.ints("mig[].mp", 1, &pp);
          mt.push_back(t); mp.push_back(pp);
        }
        if (ok && L.model == kHKY && i >= L.ng)
          for (int j = 0; j < L.nsites && ok; j++)
            ok = R.doubles("C[ci]->G[li].gtree[i].hkyi.frac[j]", 4, frac) && R.doubles("C[ci]->G[li].gtree[i].hkyi.newfrac[j]", 4, frac);
      }
      if (!ok || root < 0) { rc = fail(IMA2P_E_ARG, R.err.empty() ? "mcf file: genealogy without a root" : R.err.c_str()); break; }
      moff[L.nl] = (int)mt.size();
      mt.push_back(0.0); mp.push_back(0);
      if (up0[root] < 0 || up0[root] >= L.nl) { rc = fail(IMA2P_E_ARG, "mcf file: the root of a genealogy is a tip"); break; }
      const double roottime = time[up0[root]];                                    // :413
      rc = ima2p_engine_set_genealogy(h, ci, li, up0.data(), up1.data(), down.data(), pop.data(), time.data(), moff.data(), mt.data(),
                                      mp.data(), root, roottime, u, kappa, pi, has_stepwise(L.model) ? A.data() : nullptr);
    }
    if (rc) break;
    // :415-433: at the end of the file with chains still to fill, start over from its first chain
    const int c = fgetc(R.f);
    if (c == EOF) {
      if (ci < e.d.nchains - 1) {
        if (ci == lastci) { rc = fail(IMA2P_E_ARG, "mcf file holds no complete chain"); break; }
        fclose(R.f);
        R.f = fopen(path, "r");
        if (!R.f) return fail(IMA2P_E_ARG, "Error reopening mcffile");
        lastci = ci;
      }
    } else ungetc(c, R.f);
  }
  if (R.f) fclose(R.f);
  if (rc) return rc;
  if ((rc = ima2p_engine_upload(h)) != IMA2P_OK) return rc;
  return ima2p_engine_eval(h);
}

}  // extern "C"
