// Model tables from the population tree string and the priors, host C++: what setup_poptree (build_poptree.cpp:628-709;
// poptreeread :508-539, fillplist :391-448) and setup_iparams (initialize.cpp:201-727) build for the default model --
// one size parameter per population of the tree, a pair of migration parameters for every pair of populations that
// coexist, each spanning all the periods its two populations share.  The -j model options that merge or drop
// parameters are not covered (IMA2P_E_UNSUPPORTED where they would be asked for); exponential migration priors (-j7)
// and the no-migration model (-m 0) are.
#include "../../include/ima2p_b200.h"
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

extern "C" void ima2p_internal_set_error(const char *msg);

struct ima2p_modelspec {
  int npops = 0, nsplit = 0, ntreepops = 0, rootpop = 0, nomigration = 0, expoprior = 0, thermo = 0;
  double gbeta = 1.0;
  std::vector<int> plist, addpop, droppops, pt_b, pt_e, pt_down;
  std::vector<int> q_off, q_p, q_r, m_off, m_p, m_r, m_c;
  std::vector<double> q_max, q_min, m_max, m_min, m_mean;
};

namespace {

int mfail(int code, const std::string &m) { ima2p_internal_set_error(m.c_str()); return code; }

// "((0,1):3,2):4" -> children of every internal node; returns the node's label, or -1 on a syntax error
int parse_node(const char *&s, int npops, std::vector<int> &c0, std::vector<int> &c1) {
  if (*s == '(') {
    s++;
    const int a = parse_node(s, npops, c0, c1);
    if (a < 0 || *s != ',') return -1;
    s++;
    const int b = parse_node(s, npops, c0, c1);
    if (b < 0 || *s != ')') return -1;
    s++;
    if (*s != ':') return -1;
    s++;
    char *end = nullptr;
    const long n = strtol(s, &end, 10);
    if (end == s || n < npops || n > 2 * npops - 2 || c0[n] != -1) return -1;
    s = end;
    c0[n] = a; c1[n] = b;
    return (int)n;
  }
  char *end = nullptr;
  const long n = strtol(s, &end, 10);
  if (end == s || n < 0 || n >= npops) return -1;
  s = end;
  return (int)n;
}

}  // namespace

extern "C" {

int ima2p_modelspec_create(ima2p_modelspec **out, int npops, const char *tree, double qmax, double mmax, int expo_prior,
                           double m_mean, int thermo, double gbeta) {
  if (!out || npops < 1 || npops > 10 || !(qmax > 0) || mmax < 0) return mfail(IMA2P_E_ARG, "modelspec_create: bad argument");
  ima2p_modelspec *S = new ima2p_modelspec();
  S->npops = npops; S->nsplit = npops - 1; S->ntreepops = 2 * npops - 1; S->rootpop = 2 * npops - 2;
  S->nomigration = (mmax == 0 && !expo_prior) || npops == 1;        // -m 0, ima_main_mpi.cpp:820-823
  S->expoprior = expo_prior; S->thermo = thermo; S->gbeta = gbeta;
  const int nt = S->ntreepops;
  std::vector<int> c0(nt, -1), c1(nt, -1);
  if (npops > 1) {
    if (!tree) { delete S; return mfail(IMA2P_E_ARG, "modelspec_create: a population tree string is needed"); }
    std::string t;
    for (const char *p = tree; *p; p++) if (*p != ' ') t.push_back(*p);
    const char *s = t.c_str();
    const int root = parse_node(s, npops, c0, c1);
    if (root != S->rootpop || *s != '\0') { delete S; return mfail(IMA2P_E_ARG, std::string("population tree string not understood: ") + tree); }
    for (int n = npops; n < nt; n++) if (c0[n] < 0) { delete S; return mfail(IMA2P_E_ARG, "population tree string: an ancestral population is missing"); }
  }
  // poptreeread: ancestral population npops + k - 1 begins in period k, where its two daughters end
  S->pt_b.assign(nt, 0); S->pt_e.assign(nt, -1); S->pt_down.assign(nt, -1);
  for (int n = npops; n < nt; n++) {
    const int b = n - npops + 1;
    S->pt_b[n] = b;
    for (int ch : {c0[n], c1[n]}) {
      if (S->pt_b[ch] >= b) { delete S; return mfail(IMA2P_E_ARG, "population tree string: an ancestor must carry a larger number than its descendants"); }
      S->pt_e[ch] = b; S->pt_down[ch] = n;
    }
  }
  // fillplist: the populations of every period in increasing number; what each period adds and drops
  S->plist.assign((size_t)npops * npops, -1);
  S->addpop.assign(S->nsplit + 1, -1); S->droppops.assign((size_t)(S->nsplit + 1) * 2, -1);
  std::vector<std::vector<int>> per(npops);
  for (int k = 0; k < npops; k++) {
    for (int p = 0; p < nt; p++) if (S->pt_b[p] <= k && (S->pt_e[p] == -1 || S->pt_e[p] > k)) per[k].push_back(p);
    if ((int)per[k].size() != npops - k) { delete S; return mfail(IMA2P_E_ARG, "population tree string: wrong number of populations in a period"); }
    for (int i = 0; i < npops - k; i++) S->plist[(size_t)k * npops + i] = per[k][i];
    if (k > 0) {
      int q = 0;
      for (int p = 0; p < nt; p++) {
        if (S->pt_e[p] == k) S->droppops[(size_t)k * 2 + q++] = p;
        if (S->pt_b[p] == k && p >= npops) S->addpop[k] = p;
      }
    }
  }
  auto index_in = [&](int k, int p) { for (int i = 0; i < npops - k; i++) if (per[k][i] == p) return i; return -1; };
  // population size parameters: one per population of the tree, at every (period, row) it occupies (initialize.cpp:266-322)
  S->q_off.push_back(0);
  for (int p = 0; p < nt; p++) {
    for (int k = 0; k < npops; k++) { const int r = index_in(k, p); if (r >= 0) { S->q_p.push_back(k); S->q_r.push_back(r); } }
    S->q_off.push_back((int)S->q_p.size());
    S->q_max.push_back(qmax); S->q_min.push_back(0.0);
  }
  // migration parameters (:360-560): in period k, every pair of which at least one population is new in that period (all
  // pairs in period 0) gets one parameter per direction, lasting while both populations last
  S->m_off.push_back(0);
  if (!S->nomigration)
    for (int k = 0; k < S->nsplit; k++)
      for (int i = 0; i < npops - k - 1; i++)
        for (int j = i + 1; j < npops - k; j++) {
          const int pi = per[k][i], pj = per[k][j];
          if (k > 0 && index_in(k - 1, pi) >= 0 && index_in(k - 1, pj) >= 0) continue;
          for (int dir = 0; dir < 2; dir++) {
            const int from = dir ? pj : pi, to = dir ? pi : pj;
            for (int kk = k; kk < S->nsplit && index_in(kk, from) >= 0 && index_in(kk, to) >= 0; kk++) {
              S->m_p.push_back(kk); S->m_r.push_back(index_in(kk, from)); S->m_c.push_back(index_in(kk, to));
            }
            S->m_off.push_back((int)S->m_p.size());
            S->m_max.push_back(mmax); S->m_min.push_back(0.0); S->m_mean.push_back(expo_prior ? m_mean : 0.0);
          }
        }
  *out = S;
  return IMA2P_OK;
}

void ima2p_modelspec_free(ima2p_modelspec *s) { delete s; }

// dims[6] = npops, nsplit, numtreepops, numpopsizeparams, nummigrateparams, total weight positions of the migration parameters
int ima2p_modelspec_dims(const ima2p_modelspec *s, int *dims) {
  if (!s || !dims) return mfail(IMA2P_E_ARG, "modelspec_dims: bad argument");
  dims[0] = s->npops; dims[1] = s->nsplit; dims[2] = s->ntreepops; dims[3] = (int)s->q_max.size(); dims[4] = (int)s->m_max.size();
  dims[5] = (int)s->m_p.size();
  return IMA2P_OK;
}

// the tables in the layout ima2p_engine_set_model takes (any pointer may be NULL); pt_b/pt_e/pt_down[numtreepops]
int ima2p_modelspec_tables(const ima2p_modelspec *s, int *plist, int *addpop, int *droppops, int *pt_b, int *pt_e, int *pt_down,
                           int *q_off, int *q_p, int *q_r, int *m_off, int *m_p, int *m_r, int *m_c) {
  if (!s) return mfail(IMA2P_E_ARG, "modelspec_tables: bad argument");
  auto cp = [](int *dst, const std::vector<int> &v) { if (dst) memcpy(dst, v.data(), v.size() * sizeof(int)); };
  cp(plist, s->plist); cp(addpop, s->addpop); cp(droppops, s->droppops); cp(pt_b, s->pt_b); cp(pt_e, s->pt_e); cp(pt_down, s->pt_down);
  cp(q_off, s->q_off); cp(q_p, s->q_p); cp(q_r, s->q_r); cp(m_off, s->m_off); cp(m_p, s->m_p); cp(m_r, s->m_r); cp(m_c, s->m_c);
  return IMA2P_OK;
}

int ima2p_engine_set_model_spec(ima2p_engine *e, const ima2p_modelspec *s) {
  if (!e || !s) return mfail(IMA2P_E_ARG, "set_model_spec: bad argument");
  return ima2p_engine_set_model(e, s->npops, s->nsplit, s->plist.data(), s->addpop.data(), s->droppops.data(), s->pt_e.data(), s->pt_down.data(),
                                s->rootpop, (int)s->q_max.size(), s->q_off.data(), s->q_p.data(), s->q_r.data(), s->q_max.data(), s->q_min.data(),
                                (int)s->m_max.size(), s->m_off.data(), s->m_p.data(), s->m_r.data(), s->m_c.data(), s->m_max.data(), s->m_min.data(),
                                s->m_mean.data(), 0, nullptr, nullptr, nullptr, s->nomigration, s->expoprior, s->thermo, s->gbeta);
}

}  // extern "C"
