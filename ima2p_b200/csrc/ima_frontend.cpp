// Command-line front end over the C ABI: an M-mode run from a .u file to a .ti file, with the reference's flags for the
// things this repository implements (ima_main_mpi.cpp:252-1561 option scan, :4317-4560 main loop):
//
//   IMa2p_b200 -i data.u -o out -q QMAX -m MMAX -t TMAX -b BURNSTEPS -l GENEALOGIES [-d STEPS_BETWEEN_SAVES (100)]
//              [-hn CHAINS] [-hfg | -hfl | -hfs] [-ha A] [-hb B] [-s SEED] [-j7] [-r3 (write out.mcf at the end)] [-f file.mcf]
//
// What it does in the reference's order: readdata -> setup_poptree / setup_iparams -> a starting genealogy for every
// locus (any valid one: infinite-sites loci get a perfect phylogeny of their 0/1 columns, everything coalescing in the
// root population above the last split time; see start_genealogy) -> setheat -> burn-in -> every -d steps the cold chain's
// row (savegsampinf) is appended to out.ti -> a short report.  Values are arguments attached to the flag ("-q10") or
// the next word ("-q 10"), as the reference accepts.  Options of the reference outside this path are refused, not ignored.
//
//   IMa2p_b200 -r0 -v BASE -i data.u -o out -q QMAX -m MMAX -t TMAX [-j7] [-p5] [-p6] [-c2]
//
// is L mode (LOAD-GENEALOGY, ima_main_mpi.cpp:3216-3440, 4037-4100): the rows of BASE.ti are loaded onto the device and the
// report sections that are sums over every sampled genealogy are written in the reference's own layout -- the means /
// variances / correlations table (output.cpp:687-838), with -p6 the greater-than tables (gtint.cpp:336-447), the marginal
// peak table (surface_call_functions.cpp:175-732 over surface_search_functions.cpp:41-187) and the histogram group of the
// population-size and migration parameters (histograms.cpp:81-99, 146-431, 541-567).
#include "../../include/ima2p_b200.h"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <ctime>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>
#include <time.h>
#include <unistd.h>

namespace {

time_t g_starttime = 0;                    // set when main starts: the report's "Time Elapsed" lines
bool g_throw_instead_of_exit = false;      // set around the closing report of an M-mode run: its failure must not lose the run
[[noreturn]] void die(const std::string &m, int code = 1) {
  if (g_throw_instead_of_exit) throw std::runtime_error(m);
  fprintf(stderr, "IMa2: %s\n", m.c_str());
  exit(code);
}
void ck(int rc, const char *what) {
  if (rc != IMA2P_OK) die(std::string(what) + ": " + ima2p_last_error(), rc < 0 ? -rc : rc);
}

struct Locus { int info[8]; double hval; std::vector<int> samppop, seq, mult, A, minA, maxA; double pi[4]; };

// ---- several ranks, one GPU each (the reference: mpirun -np P IMa2p -hn N, ima_main_mpi.cpp:4317-4560) -----------------------
// Started once per GPU by any launcher that sets RANK / WORLD_SIZE / LOCAL_RANK (torchrun, mpirun wrappers, a shell loop).  -hn
// is the number of chains PER RANK, as in the reference.  The ranks find each other through small files in a directory all of
// them see (IMA2P_RENDEZVOUS_DIR, default /dev/shm): each publishes the 64-byte handle of its exchange table, opens the others'
// and from then on the kernels exchange the swap sums and the cold chain's record themselves (include/ima2p_b200.h).
struct Ranks {
  int world = 1, rank = 0, local = 0;
  std::string key;
  void init() {
    if (const char *w = getenv("WORLD_SIZE")) world = atoi(w) > 1 ? atoi(w) : 1;
    if (world == 1) return;
    rank = getenv("RANK") ? atoi(getenv("RANK")) : 0;
    local = getenv("LOCAL_RANK") ? atoi(getenv("LOCAL_RANK")) : rank;
    if (rank < 0 || rank >= world) die("RANK must be in [0, WORLD_SIZE)", 5);
    const char *dir = getenv("IMA2P_RENDEZVOUS_DIR");
    const char *job = getenv("TORCHELASTIC_RUN_ID");
    const char *port = getenv("MASTER_PORT");
    // the launcher's process id tells one launch from the next when the launcher gives the job no name
    key = std::string(dir ? dir : "/dev/shm") + "/ima2p_" + (job && strcmp(job, "none") ? job : "job") + "_" + std::to_string((long)getppid()) + "_" + (port ? port : "0");
  }
  std::string file(const std::string &what, int r) const { return key + "." + what + "." + std::to_string(r); }
  void put(const std::string &what, const void *data, size_t n) const {
    const std::string tmp = file(what, rank) + ".tmp";
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f || fwrite(data, 1, n, f) != n) die("rendezvous: cannot write " + tmp, 2);
    fclose(f);
    if (rename(tmp.c_str(), file(what, rank).c_str()) != 0) die("rendezvous: cannot publish " + file(what, rank), 2);
  }
  std::vector<unsigned char> get(const std::string &what, int r, size_t n) const {
    std::vector<unsigned char> buf(n);
    for (int tries = 0; tries < 240000; tries++) {              // up to about two minutes
      FILE *f = fopen(file(what, r).c_str(), "rb");
      if (f) { const size_t k = fread(buf.data(), 1, n, f); fclose(f); if (k == n) return buf; }
      struct timespec ts = {0, 500000};
      nanosleep(&ts, nullptr);
    }
    die("rendezvous: rank " + std::to_string(r) + " did not show up (" + file(what, r) + ")", 2);
  }
  void barrier(const std::string &name) const {
    if (world == 1) return;
    const char one = 1;
    put("bar_" + name, &one, 1);
    for (int r = 0; r < world; r++) get("bar_" + name, r, 1);
  }
  // the end of a run: every other rank says it is done and leaves; rank 0 waits for all of them and removes every file
  void finish() const {
    if (world == 1) return;
    const char one = 1;
    if (rank != 0) { put("bar_done", &one, 1); return; }
    for (int r = 1; r < world; r++) get("bar_done", r, 1);
    for (int r = 0; r < world; r++)
      for (const char *w : {"xch", "cnt", "bar_attached", "bar_done"}) remove(file(w, r).c_str());
  }
};

// One valid genealogy for a locus, every coalescence above `tbase` (so every lineage has reached the root population
// by the population tree alone and no migration event is needed).  Infinite-sites columns (0 = the first gene's base)
// define nested-or-disjoint gene sets because the first gene is in none of them; those sets are made clades.
struct Tree { std::vector<int> up0, up1, down, pop; std::vector<double> time; int root; double roottime; };

Tree start_genealogy(const Locus &L, int npops, int rootpop, double tbase, unsigned &rng) {
  const int n = L.info[1], ns = (L.info[0] == IMA2P_MODEL_IS || L.info[0] == IMA2P_MODEL_JOINT) ? L.info[2] : 0, nl = 2 * n - 1;
  Tree T;
  T.up0.assign(nl, -1); T.up1.assign(nl, -1); T.down.assign(nl, -1); T.pop.assign(nl, rootpop); T.time.assign(nl, 0.0);
  for (int i = 0, p = 0, e = L.samppop[0]; i < n; i++) { while (i >= e) e += L.samppop[++p]; T.pop[i] = p; }
  // distinct carrier sets, largest first
  std::vector<std::vector<char>> sets;
  for (int s = 0; s < ns; s++) {
    std::vector<char> c(n);
    int k = 0;
    for (int j = 0; j < n; j++) { c[j] = (char)L.seq[(size_t)j * ns + s]; k += c[j]; }
    if (k < 2) continue;                               // a singleton is a tip branch already
    bool dup = false;
    for (auto &o : sets) dup |= o == c;
    if (!dup) sets.push_back(c);
  }
  auto size_of = [&](const std::vector<char> &c) { int k = 0; for (char x : c) k += x; return k; };
  for (size_t a = 0; a < sets.size(); a++) for (size_t b = a + 1; b < sets.size(); b++) if (size_of(sets[b]) > size_of(sets[a])) std::swap(sets[a], sets[b]);
  // member lists per set; the implicit outermost set holds every gene
  std::vector<int> node_of(n);                         // current tree node standing for gene j's lineage
  std::vector<int> ntips(nl, 1);
  for (int j = 0; j < n; j++) node_of[j] = j;
  int next = n;
  auto join = [&](int a, int b) {
    if (next >= nl) die("starting genealogy: data not compatible with the infinite sites model", 36);
    const int k = next++;
    T.up0[k] = a; T.up1[k] = b; T.down[a] = T.down[b] = k;
    ntips[k] = ntips[a] + ntips[b];
    return k;
  };
  // smallest sets first: each set's current lineages are merged into one clade (sets inside it are single lineages by then)
  for (int si = (int)sets.size() - 1; si >= -1; si--) {
    std::vector<int> members;
    for (int j = 0; j < n; j++) if (si < 0 || sets[si][j]) { bool seen = false; for (int m : members) seen |= m == node_of[j]; if (!seen) members.push_back(node_of[j]); }
    while (members.size() > 1) {
      rng = rng * 1664525u + 1013904223u;
      const size_t a = (rng >> 8) % members.size();
      size_t b = (rng >> 20) % (members.size() - 1);
      if (b >= a) b++;
      const int k = join(members[a], members[b]);
      members[a] = k; members.erase(members.begin() + b);
    }
    for (int j = 0; j < n; j++) if (si < 0 || sets[si][j]) node_of[j] = members[0];
    if (si < 0) T.root = members[0];
  }
  if (next != nl) die("starting genealogy: data not compatible with the infinite sites model", 36);
  // node heights grow with the number of tips below, all above tbase
  const double step = tbase > 0 ? 0.05 * tbase : 0.05;
  std::vector<double> height(nl, 0.0);
  for (int k = n; k < nl; k++) height[k] = (tbase > 0 ? 1.1 * tbase : 0.1) + step * (ntips[k] - 1);
  for (int e = 0; e < nl; e++) T.time[e] = T.down[e] == -1 ? 1000000.0 : height[T.down[e]];
  T.roottime = height[T.root];
  return T;
}

// The constant of the infinite-sites likelihood as the reference fixes it (calc_sumlogk, calc_prob_data.cpp:609-719, called
// once per locus on the genealogy the run starts with): every site is counted at the node whose two daughters carry its two
// states, and the constant is the sum over nodes of log(count!) -- mutations on two sister branches fall on the same node.
double sumlogk_of(const Locus &L, const Tree &T) {
  const int n = L.info[1], ns = L.info[2], nl = 2 * n - 1;
  std::vector<std::vector<char>> below(nl, std::vector<char>(n, 0));          // tips under every edge
  for (int j = 0; j < n; j++) for (int e = j; e != -1; e = T.down[e]) below[e][j] = 1;
  std::vector<int> mutcount(nl, 0);
  for (int s = 0; s < ns; s++) {
    int node = -1;
    for (int e = 0; e < nl && node < 0; e++) {
      if (T.down[e] == -1) continue;
      bool same = true, comp = true;
      for (int j = 0; j < n && (same || comp); j++) { const bool c = L.seq[(size_t)j * ns + s] != 0; same &= c == (below[e][j] != 0); comp &= c != (below[e][j] != 0); }
      if (same || comp) node = T.down[e];
    }
    if (node < 0) die("starting genealogy: data not compatible with the infinite sites model", 36);
    mutcount[node]++;
  }
  double sum = 0.0;
  for (int k = n; k < nl; k++) sum += lgamma(mutcount[k] + 1.0);
  return sum;
}


// ---- L mode report ------------------------------------------------------------------------------------------------
constexpr int kGrid = 1000;                    // GRIDSIZE imamp.hpp

// histformatdouble histograms.cpp:45-78
std::string histfmt(double v) {
  char b[64];
  if (v < -1e9 || v > 1e9) { snprintf(b, sizeof b, "%-9.0lg", v); return b; }
  const double a = fabs(v);
  if (a < 1e-4) { snprintf(b, sizeof b, "%-9.8lf", v); if (!strcmp(b, "0.00000000")) return "0.0"; }
  else if (a < 1e-3) snprintf(b, sizeof b, "%-9.7lf", v);
  else if (a < 1e-2) snprintf(b, sizeof b, "%-9.6lf", v);
  else if (a < 1e-1) snprintf(b, sizeof b, "%-9.5lf", v);
  else if (a < 1e0) snprintf(b, sizeof b, "%-9.4lf", v);
  else if (a < 1e1) snprintf(b, sizeof b, "%-9.3lf", v);
  else if (a < 1e2) snprintf(b, sizeof b, "%-9.2lf", v);
  else if (a < 1e3) snprintf(b, sizeof b, "%-9.1lf", v);
  else snprintf(b, sizeof b, "%-9.0lf", v);
  return b;
}

// print_means_variances_correlations output.cpp:746-838 (the printing half; the sums come from the device)
void print_moments(FILE *f, const std::vector<std::string> &name, int nq, const std::vector<double> &m_max, const std::vector<double> &mean,
                   const std::vector<double> &var, const std::vector<double> &corr, int npops) {
  const int np = (int)name.size();
  fprintf(f, "\nMEANS, VARIANCES and CORRELATIONS OF PARAMETERS ('$' r > 0.4  '*' r > 0.75)\n");
  fprintf(f, "=============================================================================\n");
  fprintf(f, "Param:");
  for (int p = 0; p < np; p++) if (p < nq || m_max[p - nq] > 0.000001) fprintf(f, "\t%s", name[p].c_str());
  fprintf(f, "\nMean:");
  for (int p = 0; p < np; p++) fprintf(f, "\t%-.3lf", mean[p]);
  fprintf(f, "\nStdv:");
  for (int p = 0; p < np; p++) fprintf(f, "\t%-.3lf", sqrt(var[p]));
  if (npops > 1) {
    fprintf(f, "\n\nCorrelations\n");
    for (int p = 0; p < np; p++) if (p < nq || m_max[p - nq] > 0.000001) fprintf(f, "\t%s", name[p].c_str());
    fprintf(f, "\n");
    for (int p = 0; p < np; p++) {
      fprintf(f, "%s", name[p].c_str());
      for (int q = 0; q < np; q++) {
        if (q == p) { fprintf(f, "\t  - "); continue; }
        const double c = q > p ? corr[(size_t)p * np + q] : corr[(size_t)q * np + p];
        if (fabs(c) < 0.4) fprintf(f, "\t%-.3lf", c);
        else fprintf(f, "\t%-.3lf%s", c, fabs(c) < 0.75 ? "$" : "*");
      }
      fprintf(f, "\n");
    }
  }
  fprintf(f, "\n");
}

// print_greater_than_tests gtint.cpp:336-447
void print_greater_than(FILE *f, ima2p_lmode *LM, const std::vector<std::string> &name, int nq, int nm, int expo) {
  bool warn = false;
  auto table = [&](int kind, int n, int off) {
    std::vector<double> g((size_t)n * n, -1.0);
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) ck(ima2p_lmode_greater_than(LM, kind, i, j, &g[(size_t)i * n + j]), "greater-than probabilities");
    for (int i = 0; i < n; i++) fprintf(f, "\t%s", name[off + i].c_str());
    fprintf(f, "\n");
    for (int i = 0; i < n; i++) {
      fprintf(f, "%s", name[off + i].c_str());
      for (int j = 0; j < n; j++) {
        if (i == j) fprintf(f, "\t  - ");
        else if (g[(size_t)i * n + j] < 0.0) fprintf(f, "\tna");
        else if (fabs(1.0 - (g[(size_t)i * n + j] + g[(size_t)j * n + i])) >= 0.02) { fprintf(f, "\t%.3lf?", g[(size_t)i * n + j]); warn = true; }
        else fprintf(f, "\t%.3lf", g[(size_t)i * n + j]);
      }
      fprintf(f, "\n");
    }
  };
  fprintf(f, "\nPARAMETER COMPARISONS, PROBABILITY THAT ROW PARAMETER IS GREATER THAN COLUMN PARAMETER\n");
  fprintf(f, "========================================================================================\n");
  fprintf(f, "Population Sizes\n");
  table(0, nq, 0);
  fprintf(f, "\nMigration Rates\n");
  if (expo) fprintf(f, "  NOT IMPLEMENTED FOR MIGRATION RATES WITH EXPONENTIAL PRIORS \n");
  else table(1, nm, nq);
  if (warn) fprintf(f, "  \"?\" indicates that reciprocal values do not sum to approximately 1, possibly due to a small sample of genealogies\n");
  fprintf(f, "\n\n");
}

// writehistogram histograms.cpp:146-431 with dosmooth = 0 and unit scale adjustments, as prepare_parameter_histograms sets them
void write_histograms(FILE *f, const std::vector<std::string> &name, const std::vector<std::vector<double>> &x, const std::vector<std::vector<double>> &yraw,
                      const std::vector<double> &yscale = std::vector<double>(), bool dosmooth = false,
                      const std::vector<double> &before = std::vector<double>(), const std::vector<double> &after = std::vector<double>()) {
  const int nh = (int)name.size();
  // y as printed = yscaleadjust * counts; the smoothing passes of the reference work on the unscaled curve (:253-262, :306-317)
  std::vector<std::vector<double>> y(yraw);
  for (int j = 0; j < nh && !yscale.empty(); j++) for (int i = 0; i < kGrid; i++) y[j][i] = yscale[j] * yraw[j][i];
  std::vector<double> smthprob(nh, 0.0);
  std::vector<double> xysum(nh, 0.0), ysum(nh, 0.0), hpdlo(nh, 0.0), hpdhi(nh, 0.0);
  std::vector<char> c1(nh, ' '), c2(nh, ' ');
  fprintf(f, " Summaries\n\tValue  ");
  for (int j = 0; j < nh; j++) fprintf(f, "\t %s", name[j].c_str());
  fprintf(f, "\n\tMinbin ");
  for (int j = 0; j < nh; j++) { int i = 0; while (i < kGrid - 1 && y[j][i] <= 0) i++; fprintf(f, "\t%s", histfmt(x[j][i]).c_str()); }
  fprintf(f, "\n\tMaxbin");
  for (int j = 0; j < nh; j++) { int i = kGrid - 1; while (i > 0 && y[j][i] <= 0) i--; fprintf(f, "\t%s", histfmt(x[j][i]).c_str()); }
  fprintf(f, "\n\tHiPt  ");
  std::vector<int> imax(nh, 0);
  for (int j = 0; j < nh; j++) {
    double maxval = -1;
    for (int i = 0; i < kGrid; i++) {
      xysum[j] += x[j][i] * y[j][i];
      ysum[j] += y[j][i];
      if (maxval < y[j][i]) { maxval = y[j][i]; imax[j] = i; }
    }
    fprintf(f, "\t%s", histfmt(x[j][imax[j]]).c_str());
  }
  if (dosmooth) {                                                     // HiSmth :240-270: peak of the curve smoothed over 10 cells
    fprintf(f, "\n\tHiSmth");
    for (int j = 0; j < nh; j++) {
      double maxval = -1; int im = 0;
      for (int i = 0; i < kGrid; i++) {
        int cell = 10 < 2 * i ? 10 : 2 * i;
        cell = cell < 2 * (kGrid - 1 - i) ? cell : 2 * (kGrid - 1 - i);
        double den = 0, sm = 0;
        for (int k = (i - cell / 2 > 0 ? i - cell / 2 : 0); k <= (kGrid - 1 < i + cell / 2 ? kGrid - 1 : i + cell / 2); k++) {
          const double term = 1.0 / (0.5 + abs(k - i));
          sm += yraw[j][k] * term; den += term;
        }
        sm /= den;
        if (maxval < sm) { maxval = sm; smthprob[j] = maxval; im = i; }
      }
      fprintf(f, "\t%s", histfmt(x[j][im]).c_str());
    }
  }
  fprintf(f, "\n\tMean  ");
  for (int j = 0; j < nh; j++) fprintf(f, "\t%s", histfmt(xysum[j] / ysum[j]).c_str());
  fprintf(f, "\n\t95%%Lo  ");
  for (int j = 0; j < nh; j++) {
    int i = 0; double sum = 0;
    while (i < kGrid && (sum + y[j][i]) / ysum[j] <= 0.025) { sum += y[j][i]; i++; }
    fprintf(f, "\t%s", histfmt(x[j][i < kGrid ? i : kGrid - 1]).c_str());
  }
  fprintf(f, "\n\t95%%Hi  ");
  for (int j = 0; j < nh; j++) {
    int i = kGrid - 1; double sum = 0;
    while ((sum + y[j][i]) / ysum[j] <= 0.025 && i > 0) { sum += y[j][i]; i--; }
    fprintf(f, "\t%s", histfmt(x[j][i]).c_str());
  }
  // highest posterior density interval: smooth over 30 cells, sort by height, accumulate 95% from the top (:296-372)
  for (int j = 0; j < nh; j++) {
    std::vector<std::pair<double, double>> h(kGrid);       // (p, v)
    double tempsum = 0;
    for (int i = 0; i < kGrid; i++) {
      int cell = 30 < 2 * i ? 30 : 2 * i;
      cell = cell < 2 * (kGrid - 1 - i) ? cell : 2 * (kGrid - 1 - i);
      double den = 0, sm = 0;
      for (int k = (i - cell / 2 > 0 ? i - cell / 2 : 0); k <= (kGrid - 1 < i + cell / 2 ? kGrid - 1 : i + cell / 2); k++) {
        const double term = 1.0 / (0.5 + abs(k - i));
        sm += yraw[j][k] * term; den += term;
      }
      sm /= den;
      tempsum += sm;
      h[i] = std::make_pair(sm, x[j][i]);
    }
    const double vminhold = h[0].second;
    std::sort(h.begin(), h.end());                         // by height, ties by value: the order shellhist leaves (utilities.cpp:674-702)
    double hpdmax = -1, hpdmin = 1e10, sum = h[kGrid - 1].first;
    int i = kGrid - 1;
    while (i >= 0 && sum <= 0.95 * tempsum) {
      if (h[i].second > hpdmax) hpdmax = h[i].second;
      if (h[i].second < hpdmin) hpdmin = h[i].second;
      i--;
      if (i >= 0) sum += h[i].first;
    }
    hpdlo[j] = hpdmin <= vminhold ? 0.0 : hpdmin;
    hpdhi[j] = hpdmax;
    while (i > 0 && (h[i].second < hpdmin || h[i].second > hpdmax)) i--;
    if (i > 0) c2[j] = '?';
    if (smthprob[j] * 0.05 < yraw[j][0] && smthprob[j] * 0.05 < yraw[j][kGrid - 1]) c1[j] = '#';      // smthprobvals is 0 without smoothing (:367)
  }
  fprintf(f, "\n\tHPD95Lo");
  for (int j = 0; j < nh; j++) fprintf(f, "\t%s%c%c", histfmt(hpdlo[j]).c_str(), c1[j], c2[j]);
  fprintf(f, "\n\tHPD95Hi");
  for (int j = 0; j < nh; j++) fprintf(f, "\t%s%c%c", histfmt(hpdhi[j]).c_str(), c1[j], c2[j]);
  fprintf(f, "\n\n\tParameter");
  for (int j = 0; j < nh; j++) fprintf(f, "\t%s\tP", name[j].c_str());
  fprintf(f, "\n\tHiPt");
  for (int j = 0; j < nh; j++) fprintf(f, "\t%s\t%s", histfmt(x[j][imax[j]]).c_str(), histfmt(y[j][imax[j]]).c_str());
  fprintf(f, "\n\n");
  for (int i = 0; i < kGrid; i++) {
    fprintf(f, "\t%4d", i);
    for (int j = 0; j < nh; j++) fprintf(f, "\t%s\t%s", histfmt(x[j][i]).c_str(), histfmt(y[j][i]).c_str());
    fprintf(f, "\n");
  }
  fprintf(f, " SumP\t");
  for (int j = 0; j < nh; j++) fprintf(f, "\t\t%s", histfmt(ysum[j]).c_str());
  fprintf(f, "\n Before\t");
  for (int j = 0; j < nh; j++) fprintf(f, "\t\t%s", histfmt(before.empty() ? 0.0 : before[j] * (yscale.empty() ? 1.0 : yscale[j])).c_str());
  fprintf(f, "\n After\t");
  for (int j = 0; j < nh; j++) fprintf(f, "\t\t%s", histfmt(after.empty() ? 0.0 : after[j] * (yscale.empty() ? 1.0 : yscale[j])).c_str());
  fprintf(f, "\n");
}


// ---- marginal peak search, in lock step.  The reference finds the peak of one marginal curve at a time and calls marginp once
// per iterate (surface_search_functions.cpp:41-187 mnbrakmod / goldenmod under surface_call_functions.cpp:175-274 marginalopt;
// :277-297 margin95 over :82-104 marginbis; popmig.cpp:366-429 for the 2NM terms).  Every such search only ever asks "the curve's
// value at my next abscissa", and the searches of different parameters, row sets and brackets do not depend on one another.  So each
// search is a small resumable object (`want()` = the abscissa it needs, `take(f)` = here is the value, advance to the next one), and
// SearchPool::run advances all of them together: one round = one device pass (ima2p_lmode_marginal_many) over the current abscissa
// of every search still running.  A search sees the same sequence of values the reference's serial loop would have produced, so
// iterates, peaks and bounds are the reference's to the last bit; the number of device passes is the longest search's length instead
// of the sum of all lengths.
struct Curve {                       // which function a search evaluates
  int kind;                          // 0 marginp over [first, last); 1 log margincalc - yadjust (all rows); 2 marginpopmig (2NM term)
  int param, first, last, thetai;
  double yadjust;
};

// Downhill bracketing with parabolic extrapolation (the rule of mnbrakmod :41-131): from two abscissae, walk downhill until the
// middle point lies below both ends.  at = which evaluation is outstanding.
struct Bracketing {
  enum At { kFirst, kSecond, kThird, kInside, kBeyond, kShift, kDone } at = kDone;
  double a = 0, b = 0, c = 0, fa = 0, fb = 0, fc = 0, u = 0;
  void start(double a0, double b0) { a = a0; b = b0; at = kFirst; }
  bool done() const { return at == kDone; }
  double want() const { return at == kFirst ? a : at == kSecond ? b : at == kThird ? c : u; }
  void take(double f) {
    const double gold = 1.618034;
    switch (at) {
    case kFirst: fa = f; at = kSecond; return;
    case kSecond:
      fb = f;
      if (fb > fa) { std::swap(a, b); std::swap(fa, fb); }
      c = fabs(b + gold * (b - a));
      at = kThird;
      return;
    case kThird: fc = f; break;
    case kInside:                                        // the parabola's minimum lay between b and c
      if (f < fc) { a = b; fa = fb; b = u; fb = f; break; }
      if (f > fb) { c = u; fc = f; break; }
      u = c + gold * (c - b);                            // no use: default magnification
      at = kShift;
      return;
    case kBeyond:                                        // it lay between c and the allowed limit
      if (f < fc) { b = c; c = u; u = c + gold * (c - b); fb = fc; fc = f; at = kShift; return; }
      shift(f);
      break;
    case kShift: shift(f); break;
    case kDone: return;
    }
    next_trial();
  }
 private:
  void shift(double fu) { a = b; b = c; c = u; fa = fb; fb = fc; fc = fu; }
  void next_trial() {
    const double gold = 1.618034, glimit = 100.0, tiny = (double)(float)1.0e-20;
    if (!(fb >= fc && fb > -DBL_MAX && !(fb == 0 && fc == 0))) { at = kDone; return; }
    const double rr = (b - a) * (fb - fc), q = (b - c) * (fb - fa);
    const double dq = fabs(q - rr) > tiny ? fabs(q - rr) : tiny;
    u = b - ((b - c) * q - (b - a) * rr) / (2.0 * (q - rr > 0.0 ? fabs(dq) : -fabs(dq)));
    const double ulim = b + glimit * (c - b);
    if ((b - u) * (u - c) > 0.0) at = kInside;
    else if ((c - u) * (u - ulim) > 0.0) at = kBeyond;
    else { if ((u - ulim) * (ulim - c) >= 0.0) u = ulim; else u = c + gold * (c - b); at = kShift; }
  }
};

// Golden-section descent inside a bracket (the rule of goldenmod :133-187)
struct GoldenSection {
  enum At { kInner1, kInner2, kLeft, kRight, kDone } at = kDone;
  double x0 = 0, x1 = 0, x2 = 0, x3 = 0, f1 = 0, f2 = 0, tol = 1e-7;
  void start(double ax, double bx, double cx, double tolerance) {
    const double cc = 1.0 - 0.61803399;
    tol = tolerance; x0 = ax; x3 = cx;
    if (fabs(cx - bx) > fabs(bx - ax)) { x1 = bx; x2 = bx + cc * (cx - bx); }
    else { x2 = bx; x1 = bx - cc * (bx - ax); }
    at = kInner1;
  }
  bool done() const { return at == kDone; }
  double want() const { return (at == kInner1 || at == kLeft) ? x1 : x2; }
  void take(double f) {
    const double r = 0.61803399, cc = 1.0 - r;
    if (at == kInner1) { f1 = f; at = kInner2; return; }
    if (at == kInner2 || at == kRight) f2 = f; else f1 = f;
    if (!(fabs(x3 - x0) > tol * (fabs(x1) + fabs(x2)) && x3 > -DBL_MAX)) { at = kDone; return; }
    if (f2 < f1) { x0 = x1; x1 = x2; x2 = r * x1 + cc * x3; f1 = f2; at = kRight; }
    else { x3 = x2; x2 = x1; x1 = r * x2 + cc * x0; f2 = f1; at = kLeft; }
  }
  double xmin() const { return f1 < f2 ? x1 : x2; }
  double fmin() const { return f1 < f2 ? f1 : f2; }
};

// Root by halving (the rule of marginbis :82-104: BISTOL 1e-4, at most 40 halvings)
struct Bisection {
  enum At { kLow, kHigh, kMid, kDone } at = kDone;
  double x1 = 0, x2 = 0, fl = 0, dx = 0, rtb = 0, xmid = 0, root = DBL_MAX;
  int j = 0;
  void start(double lo, double hi) { x1 = lo; x2 = hi; at = kLow; }
  bool done() const { return at == kDone; }
  double want() const { return at == kLow ? x1 : at == kHigh ? x2 : xmid; }
  void take(double f) {
    if (at == kLow) { fl = f; at = kHigh; return; }
    if (at == kHigh) {
      if (fl * f >= 0.0) { root = DBL_MIN; at = kDone; return; }
      if (fl < 0.0) { dx = x2 - x1; rtb = x1; } else { dx = x1 - x2; rtb = x2; }
      j = 1; xmid = rtb + (dx *= 0.5); at = kMid;
      return;
    }
    if (f <= 0.0) rtb = xmid;
    if (fabs(dx) < 1e-4 || f == 0.0) { root = rtb; at = kDone; return; }
    if (++j > 40) { root = DBL_MAX; at = kDone; return; }
    xmid = rtb + (dx *= 0.5);
  }
};

// One peak of one curve: the two bracketings of marginalopt (from the prior's upper end and from its lower end, :192-201; they
// run side by side), its test that both found the same valley (:203-262, as written, with the unconditional clamp of :233-237),
// then the golden section.  With `bracket` given (2NM terms: the bin of a 100-point scan), only the golden section.
struct PeakSearch {
  Curve curve;
  double prior = 0, mlval = 0, peakloc = 0;
  Bracketing down, up;
  GoldenSection gs;
  int stage = 0;                      // 0 bracketing, 1 golden section, 2 finished
  void start(const Curve &c, double prior_max) {
    curve = c; prior = prior_max;
    down.start(prior, prior / 2);
    up.start(0.0000001, prior / 2);
    stage = 0;
  }
  void start_in(const Curve &c, double ax, double bx, double cx) { curve = c; gs.start(ax, bx, cx, 1e-7); stage = 1; }
  bool done() const { return stage == 2; }
  // the abscissae wanted this round (the two bracketings are independent: up to two)
  int wants(double *x) const {
    int n = 0;
    if (stage == 0) { if (!down.done()) x[n++] = down.want(); if (!up.done()) x[n++] = up.want(); }
    else if (stage == 1) x[n++] = gs.want();
    return n;
  }
  void take(const double *f) {
    int n = 0;
    if (stage == 0) {
      if (!down.done()) down.take(f[n++]);
      if (!up.done()) up.take(f[n++]);
      if (down.done() && up.done()) same_valley();
    } else if (stage == 1) {
      gs.take(f[n++]);
      if (gs.done()) { mlval = -gs.fmin(); peakloc = gs.xmin(); stage = 2; }
    }
  }
 private:
  static double lowest(double a, double b, double c) { double v = 0; if (a < b && a < c) v = a; if (b < a && b < c) v = b; if (c < a && c < b) v = c; return v; }
  void same_valley() {
    double axt = down.a, bxt = down.b, cxt = down.c, ax = up.a, bx = up.b, cx = up.c;
    const double min0 = lowest(axt, bxt, cxt), min1 = lowest(ax, bx, cx);
    double max0 = 0, max1 = 0;
    if (axt > bxt && axt > cxt) max0 = axt = std::min(axt, prior);
    if (bxt > axt && bxt > cxt) max0 = bxt = std::min(bxt, prior);
    if (cxt > axt && cxt > bxt) max0 = cxt = std::min(cxt, prior);
    max1 = ax = std::min(ax, prior);                       // :233-237: clamps ax and takes it whatever the order
    if (bx > ax && bx > cx) max1 = bx = std::min(bx, prior);
    if (cx > ax && cx > bx) max1 = cx = std::min(cx, prior);
    if (max0 <= min1 || max1 <= min0) { peakloc = -1; stage = 2; return; }
    gs.start(ax, bx, cx, 1e-7);
    stage = 1;
  }
};

// Advances any number of searches together: every round, one device pass for the marginal curves of all of them
// (ima2p_lmode_marginal_many) and one marginpopmig call per 2NM curve.
struct SearchPool {
  ima2p_lmode *LM;
  int passes = 0;
  long long points = 0;
  struct Slot { Curve c; std::function<int(double *)> wants; std::function<void(const double *)> take; };
  std::vector<Slot> slots;
  template <class S> void add(S *s, const Curve &c) {
    slots.push_back(Slot{c, [s](double *x) { if (s->done()) return 0; x[0] = s->want(); return 1; }, [s](const double *f) { s->take(f[0]); }});
  }
  void add(PeakSearch *s) { slots.push_back(Slot{s->curve, [s](double *x) { return s->wants(x); }, [s](const double *f) { s->take(f); }}); }
  void run() {
    std::vector<int> kind, param, first, last, owner, nwant(slots.size());
    std::vector<double> x, yadj, f;
    for (;;) {
      kind.clear(); param.clear(); first.clear(); last.clear(); x.clear(); yadj.clear(); owner.clear();
      for (size_t i = 0; i < slots.size(); i++) {
        double w[2];
        nwant[i] = slots[i].wants(w);
        for (int k = 0; k < nwant[i]; k++) {
          const Curve &c = slots[i].c;
          kind.push_back(c.kind); param.push_back(c.param); first.push_back(c.first); last.push_back(c.last);
          x.push_back(w[k]); yadj.push_back(c.yadjust); owner.push_back((int)i);
        }
      }
      const int n = (int)x.size();
      if (!n) break;
      f.assign(n, 0.0);
      // the marginal curves of this round in one pass; 2NM curves (their own kernels) one call each
      std::vector<int> mk, mp, mf, ml, at;
      std::vector<double> mx, my, mo;
      for (int q = 0; q < n; q++) {
        if (kind[q] == 2) {
          const Curve &c = slots[owner[q]].c;
          ck(ima2p_lmode_marginpopmig(LM, c.thetai, c.param, c.first, c.last, &x[q], 1, &f[q]), "2NM density");
        } else { mk.push_back(kind[q]); mp.push_back(param[q]); mf.push_back(first[q]); ml.push_back(last[q]); mx.push_back(x[q]); my.push_back(yadj[q]); at.push_back(q); }
      }
      if (!mx.empty()) {
        mo.assign(mx.size(), 0.0);
        ck(ima2p_lmode_marginal_many(LM, (int)mx.size(), mk.data(), mp.data(), mf.data(), ml.data(), mx.data(), my.data(), mo.data()), "marginal density");
        for (size_t k = 0; k < at.size(); k++) f[at[k]] = mo[k];
        passes++; points += (long long)mx.size();
      }
      int q = 0;
      for (size_t i = 0; i < slots.size(); i++) if (nwant[i]) { slots[i].take(&f[q]); q += nwant[i]; }
    }
    slots.clear();
  }
};

struct PopMigTerm { int thetai, mi; std::string peakname, histname; };
// the bin of a 100-point scan that holds the lowest value of a 2NM curve (popmig.cpp:384-417): the bracket its golden section starts in
void popmig_bracket(ima2p_lmode *LM, const PopMigTerm &t, int first, int last, double ub, double *ax, double *bx, double *cx) {
  const int bins = 100;
  const double kMinParam = 0.0000001, w = ub / (double)bins;
  std::vector<double> x(bins), fa(bins);
  for (int j = 0; j < bins; j++) x[j] = kMinParam + j * w;
  ck(ima2p_lmode_marginpopmig(LM, t.thetai, t.mi, first, last, x.data(), bins, fa.data()), "2NM density");
  double maxf = 1e100; int maxj = -1;
  for (int j = 0; j < bins; j++) if (fa[j] < maxf) { maxf = fa[j]; maxj = j; }
  if (maxj == 0) { *ax = kMinParam; *cx = kMinParam + w; *bx = (*ax + *cx) / 2.0; }
  else if (maxj == bins - 1) { *ax = kMinParam + ((double)bins - 1) * w; *bx = kMinParam + ((double)bins - 2) * w; *cx = ub; }
  else { *bx = kMinParam + ((double)maxj - 1) * w; *cx = kMinParam + ((double)maxj + 1) * w; *ax = kMinParam + ((double)maxj) * w; }
}

// findmarginpeaks :316-732
void print_marginal_peaks(FILE *f, ima2p_lmode *LM, long long G, const std::vector<std::string> &name, int nq, int nm, int nsplit,
                          const std::vector<double> &prior_max, const std::vector<int> &pb, const std::vector<int> &pe,
                          const std::vector<PopMigTerm> &terms, const std::vector<double> &term_upper) {
  const int p = nq + nm, NT = 2;                      // NUMTREEINT
  fprintf(f, "\nMarginal Peak Locations and Probabilities\n=========================================\n");
  fprintf(f, "  peak locations are estimated using a peak finding algorithm, which may\n");
  fprintf(f, "  fail if the curve has multiple peaks. All peaks can also be found, and\n");
  fprintf(f, "  related analyses conducted, by plotting the histograms\n\n");
  if (G <= 10) { fprintf(f, " TOO FEW TREES SAVED - MARGINAL VALUES NOT FOUND \n\n"); return; }
  std::vector<std::vector<double>> mlval(NT + 1, std::vector<double>(p, 0.0)), peakloc(NT + 1, std::vector<double>(p, 0.0));
  const int nt = (int)terms.size();                   // 2NM terms (-p5), in the order of the migration parameters' populations
  std::vector<std::vector<double>> pmml(NT + 1, std::vector<double>(nt, 0.0)), pmpk(NT + 1, std::vector<double>(nt, 0.0));
  // row sets: the two halves of the run and all rows (:352-372)
  int set_first[NT + 1], set_last[NT + 1];
  {
    int firsttree = 0, lasttree = (int)G / NT;
    for (int j = 0; j < NT; j++) {
      set_first[j] = firsttree; set_last[j] = lasttree;
      firsttree = lasttree + 1;
      lasttree += (int)G / NT;
      if (lasttree > G) lasttree = (int)G;
    }
    set_first[NT] = 0; set_last[NT] = (int)G;
  }
  // every peak search of the table -- (NT + 1) row sets x (p parameters + nt 2NM terms) -- advances in one pool
  SearchPool pool{LM};
  std::vector<PeakSearch> ps((size_t)(NT + 1) * p), pms((size_t)(NT + 1) * nt);
  for (int j = 0; j <= NT; j++) {
    for (int i = 0; i < p; i++) {
      ps[(size_t)j * p + i].start(Curve{0, i, set_first[j], set_last[j], -1, 0.0}, prior_max[i]);
      pool.add(&ps[(size_t)j * p + i]);
    }
    for (int i = 0; i < nt; i++) {
      double ax, bx, cx;
      popmig_bracket(LM, terms[i], set_first[j], set_last[j], term_upper[i], &ax, &bx, &cx);
      pms[(size_t)j * nt + i].start_in(Curve{2, terms[i].mi, set_first[j], set_last[j], terms[i].thetai, 0.0}, ax, bx, cx);
      pool.add(&pms[(size_t)j * nt + i]);
    }
  }
  pool.run();
  for (int j = 0; j <= NT; j++) {
    for (int i = 0; i < p; i++) { mlval[j][i] = ps[(size_t)j * p + i].mlval; peakloc[j][i] = ps[(size_t)j * p + i].peakloc; }
    for (int i = 0; i < nt; i++) { pmml[j][i] = pms[(size_t)j * nt + i].mlval; pmpk[j][i] = pms[(size_t)j * nt + i].peakloc; }
  }
  // likelihood-ratio tests of the migration rates and 2NM terms: the curve at its peak and at zero (:380-399), one pass for all
  std::vector<double> migtest(nm, 0.0), pmtest(nt, 0.0);
  {
    std::vector<int> kind(2 * nm, 0), par(2 * nm), fi(2 * nm, 0), la(2 * nm, (int)G);
    std::vector<double> x(2 * nm), v(2 * nm, 0.0);
    for (int i = 0; i < nm; i++) { par[2 * i] = par[2 * i + 1] = i + nq; x[2 * i] = peakloc[NT][i + nq]; x[2 * i + 1] = 0.0000001; }
    if (nm) ck(ima2p_lmode_marginal_many(LM, 2 * nm, kind.data(), par.data(), fi.data(), la.data(), x.data(), nullptr, v.data()), "marginal density");
    for (int i = 0; i < nm; i++) migtest[i] = 2 * log((-v[2 * i]) / (-v[2 * i + 1]));
    for (int i = 0; i < nt; i++) {
      const double xs[2] = {pmpk[NT][i], 0.0000001};
      double f2[2];
      ck(ima2p_lmode_marginpopmig(LM, terms[i].thetai, terms[i].mi, 0, (int)G, xs, 2, f2), "2NM density");
      pmtest[i] = 2 * log((-f2[0]) / (-f2[1]));
    }
  }
  // the 95% bounds (:277-297): for every parameter with a peak, two roots of log margincalc - (log peak height - 1.92), by
  // halving, all 2 p of them in lock step
  std::vector<Bisection> lo95(p), hi95(p);
  for (int i = 0; i < p; i++) {
    if (peakloc[NT][i] < 0) continue;
    const Curve c{1, i, 0, (int)G, -1, log(mlval[NT][i]) - 1.92};
    lo95[i].start(0.0000001, peakloc[NT][i]);
    hi95[i].start(peakloc[NT][i], prior_max[i]);
    pool.add(&lo95[i], c);
    pool.add(&hi95[i], c);
  }
  pool.run();
  bool errnote = false, signote = false;
  static const char *sig[4] = {"ns", "*", "**", "***"};
  for (int k = 0; k <= nsplit; k++) {
    fprintf(f, "\nPeriod %d\n--------\n", k);
    const int iihi = k < nsplit ? 2 : 1;
    for (int ii = 1; ii <= iihi; ii++) {
      const int ilo = ii == 1 ? 0 : nq, ihi = ii == 1 ? nq : p;
      int nprint = 0;
      if (ii == 1) {
        fprintf(f, "Population Size Parameters\n Param:");
        for (int i = 0; i < nq; i++) if (pb[i] == k) fprintf(f, "\t %s\tP", name[i].c_str());
        fprintf(f, "\n");
      } else {
        for (int i = 0; i < nm; i++) nprint += pb[nq + i] == k;
        if (nprint) {
          fprintf(f, "Migration Rate Parameters\n Param:");
          for (int i = 0; i < nm; i++) if (prior_max[nq + i] > 0.000001 && pb[nq + i] == k) fprintf(f, "\t %s\tP", name[nq + i].c_str());
          fprintf(f, "\n");
        }
      }
      if (ii == 2 && nprint == 0) continue;
      for (int j = 0; j <= NT; j++) {
        if (j < NT) fprintf(f, " Set%d", j); else fprintf(f, " All");
        for (int i = ilo; i < ihi; i++)
          if (pb[i] == k) {
            if (peakloc[j][i] >= 0) fprintf(f, "\t%7.3lf\t%7.3lf", peakloc[j][i], mlval[j][i]);
            else { fprintf(f, "\terror*\t"); errnote = true; }
          }
        fprintf(f, "\n");
      }
      fprintf(f, " LR95%%Lo");
      for (int i = ilo; i < ihi; i++)
        if (pb[i] == k) {
          if (peakloc[NT][i] >= 0) {
            const double t = lo95[i].root;
            if (t <= 0) fprintf(f, "\t<min\t");
            else if (t >= DBL_MAX || t <= DBL_MIN) fprintf(f, "\tna\t");
            else fprintf(f, "\t%7.3lf\t", t);
          } else fprintf(f, "\t\t");
        }
      fprintf(f, "\n LR95%%Hi");
      for (int i = ilo; i < ihi; i++)
        if (pb[i] == k) {
          if (peakloc[NT][i] >= 0) {
            const double t = hi95[i].root;
            if (t >= prior_max[i] || t <= DBL_MIN) fprintf(f, "\t>max\t");
            else fprintf(f, "\t%7.3lf\t", t);
          } else fprintf(f, "\t\t");
        }
      fprintf(f, "\n");
      if (ii == 2) {
        fprintf(f, " LLRtest ");
        for (int i = 0; i < nm; i++)
          if (pb[nq + i] == k && prior_max[nq + i] > 0.000001) {
            if (fabs(migtest[i]) > 1e6) fprintf(f, "bad value\t");
            else {
              const int lev = migtest[i] > 9.54954 ? 3 : migtest[i] > 5.41189 ? 2 : migtest[i] > 2.70554 ? 1 : 0;
              if (lev == 3) signote = true;
              fprintf(f, "%7.3lf%s\t", migtest[i], sig[lev]);
            }
          }
        fprintf(f, "\n");
      }
      fprintf(f, " LastPeriod");
      for (int i = ilo; i < ihi; i++)
        if (pb[i] == k && (ii == 1 || prior_max[i] > 0.000001)) fprintf(f, "\t%d\t", pe[i]);
      fprintf(f, "\n");
    }
    if (nt && k < nsplit) {                           // case 3 of the reference's loop: the 2NM terms of this period
      int nprint = 0;
      for (int i = 0; i < nm; i++) nprint += pb[nq + i] == k;
      if (nprint) {
        fprintf(f, "Population Migration (2NM) Terms\n Term:");
        for (int i = 0; i < nt; i++) if (prior_max[nq + terms[i].mi] > 0.000001 && pb[nq + terms[i].mi] == k) fprintf(f, "\t %s\tP", terms[i].peakname.c_str());
        fprintf(f, "\n");
        for (int j = 0; j <= NT; j++) {
          if (j < NT) fprintf(f, " Set%d", j); else fprintf(f, " All");
          for (int i = 0; i < nt; i++)
            if (pb[nq + terms[i].mi] == k) {
              if (pmpk[j][i] >= 0) fprintf(f, "\t%7.3lf\t%7.3lf", pmpk[j][i], pmml[j][i]);
              else { fprintf(f, "\terror*\t"); errnote = true; }
            }
          fprintf(f, "\n");
        }
        fprintf(f, " LLRtest ");
        for (int i = 0; i < nt; i++)
          if (pb[nq + terms[i].mi] == k && prior_max[nq + terms[i].mi] > 0.000001) {
            if (fabs(pmtest[i]) > 1e6) fprintf(f, "bad value\t");
            else {
              const int lev = pmtest[i] > 9.54954 ? 3 : pmtest[i] > 5.41189 ? 2 : pmtest[i] > 2.70554 ? 1 : 0;
              if (lev == 3) signote = true;
              fprintf(f, "%7.3lf%s\t", pmtest[i], sig[lev]);
            }
          }
        fprintf(f, "\n LastPeriod");
        for (int i = 0; i < nt; i++)
          if (pb[nq + terms[i].mi] == k && prior_max[nq + terms[i].mi] > 0.000001) fprintf(f, "\t%d\t", pe[nq + terms[i].mi]);
        fprintf(f, "\n");
      }
    }
  }
  if (errnote) fprintf(f, "*  peak not found possibly due to multiple peaks (check plot of marginal density) \n");
  if (signote) fprintf(f, " migration rate likelihood ratio test - see Nielsen and Wakeley (2001)\n migration significance levels :  * p < 0.05;   **  p < 0.01,   *** p < 0.001\n");
  fprintf(f, "\n");
}


// ---- joint posterior peak: jointfind.cpp:599-812 (differential evolution: startpop, nextgen, difeloop, copybest, modelloop)
// and :815-879, 1087-1184 (the table).  The reference evaluates jointp for one individual after the other; a generation's
// trial vectors do not depend on each other (they are all built from the previous population), so the whole generation
// goes to the device in one ima2p_lmode_jointp call.  The random numbers are this
// program's own, the peak it converges to is the reference's (spread tolerance 1e-7, found twice before stopping).
std::string logpfmt(double v) {                     // logpstrformat jointfind.cpp:295-318
  char b[64];
  const double a = fabs(v);
  if (a < 1e-2) snprintf(b, sizeof b, "%.6lf", v);
  else if (a < 1e-1) snprintf(b, sizeof b, "%.5lf", v);
  else if (a < 1e-0) snprintf(b, sizeof b, "%.4lf", v);
  else if (a < 1e1) snprintf(b, sizeof b, "%.3lf", v);
  else if (a < 1e2) snprintf(b, sizeof b, "%.2lf", v);
  else if (a < 1e3) snprintf(b, sizeof b, "%.1lf", v);
  else snprintf(b, sizeof b, "%.0lf", v);
  return b;
}

// A nested model (-w file, jointfind.cpp:40-135): some parameters of the full model are tied to another one ("equal") or fixed
// ("constant"); the search runs over the parameters that are left.  map[i] says where full parameter i comes from: a position in
// the vector of free parameters (>= 0) or minus a constant (< 0) -- the reference's xmap.
struct NestedModel {
  std::string name;
  int type = -1;                     // 0: sizes and migration rates (two populations); 1: sizes only; 2: migration rates only
  std::vector<double> map;
  std::vector<int> listed;           // parameters named on an equal / constant line (printed in brackets)
  int nfree = 0;                     // free parameters within the model type's family (#terms)
  bool boundary = false;             // a constant sits on a bound of the prior: the 2LLR distribution is a mixture
  void expand(const double *freevals, double *full) const {            // mapvals :563-574
    for (size_t i = 0; i < map.size(); i++) full[i] = map[i] < 0 ? -map[i] : freevals[(int)map[i]];
  }
  void reduce(const double *full, double *freevals) const {            // reversemapvals :547-560
    int to = 0;
    for (size_t i = 0; i < map.size(); i++) if (to == (int)map[i]) freevals[to++] = full[i];
  }
};

// setup_mapping :380-543.  File: a line with the number of models, then per model a line "model <name>" followed by lines
// "equal p|m i j k ..." (j, k, ... always take the value of i) and "constant p|m value i j ..." (fixed at value, clamped into the
// prior); p indices are population numbers, m indices count the migration parameters; anything after the numbers is comment.
std::vector<NestedModel> read_nested_models(const std::string &fname, int npops, int nq, int nm, double thetaprior, double mprior) {
  FILE *f = fopen(fname.c_str(), "r");
  if (!f) die("Error opening nested model file: " + fname, 1);        // IMERR_READFILEOPENFAIL
  const int np = nq + nm;
  std::vector<NestedModel> models;
  struct Tie { bool constant; double head; std::vector<int> members; };
  std::vector<Tie> ties;
  std::vector<char> removed;
  auto finish = [&]() {
    if (models.empty()) return;
    NestedModel &M = models.back();
    const int k = (int)models.size() - 1;
    if (ties.empty()) die("nested model " + std::to_string(k) + " does not include any 'constant' or 'equal' specifications", 16);
    if (npops == 2 && M.type != 0) die("nested model " + std::to_string(k) + " has m or p type but should be both when there are only two sampled populations", 16);
    if (npops > 2 && M.type <= 0) die("nested model " + std::to_string(k) + " has type 0 but should be m or p type", 16);
    int rank = 0;
    for (int i = 0; i < np; i++) if (!removed[i]) M.map[i] = rank++;           // the free parameters keep their order
    for (const Tie &t : ties)
      for (int i : t.members) M.map[i] = t.constant ? -t.head : (double)(int)M.map[(int)t.head];
    const int family = M.type == 1 ? nq : M.type == 2 ? nm : np;
    int gone = 0;
    for (int i = 0; i < np; i++) gone += removed[i];
    M.nfree = family - gone;
  };
  char line[1024];
  bool counted = false;
  while (fgets(line, sizeof line, f)) {
    if (!counted) { counted = isdigit((unsigned char)line[0]) != 0; continue; }          // the count itself is implied by the model lines
    if (!isalpha((unsigned char)line[0])) continue;
    std::string word;
    char *c = line;
    while (*c && !isspace((unsigned char)*c)) word += (char)tolower((unsigned char)*c++);
    while (*c && isspace((unsigned char)*c)) c++;
    if (word == "model") {
      finish();
      NestedModel M;
      M.name = c;
      while (!M.name.empty() && (M.name.back() == '\n' || M.name.back() == '\r')) M.name.pop_back();
      M.type = npops == 2 ? 0 : -1;
      M.map.assign(np, -1.0); M.listed.assign(np, 0);
      models.push_back(M);
      ties.clear(); removed.assign(np, 0);
    } else if ((word == "equal" || word == "constant") && !models.empty()) {
      NestedModel &M = models.back();
      const char family = *c;
      if (family != 'm' && family != 'p') continue;
      if (M.type == -1) M.type = family == 'm' ? 2 : 1;
      else if ((family == 'm' && M.type == 1) || (family == 'p' && M.type == 2)) die("nested model " + std::to_string(models.size() - 1) + " specified as both p and m types", 16);
      const int base = family == 'm' ? nq : 0;
      c++;
      char *end = c;
      Tie t;
      t.constant = word == "constant";
      t.head = strtod(c, &end);
      if (end == c) continue;
      c = end;
      if (!t.constant) t.head += base;
      else if (t.head <= 0.0000001) { t.head = 0.0000001; M.boundary = true; }           // MINPARAMVAL
      else if (family == 'm' && t.head >= mprior) { t.head = mprior; M.boundary = true; }
      else if (family == 'p' && t.head >= thetaprior) { t.head = thetaprior; M.boundary = true; }
      for (;;) {
        const long v = strtol(c, &end, 10);
        if (end == c) break;
        c = end;
        const int i = (int)v + base;
        if (i < 0 || i >= np) die("nested model file: parameter number out of range", 16);
        t.members.push_back(i);
        M.listed[i] = 1;
        removed[i] = 1;
      }
      ties.push_back(t);
    }
  }
  finish();
  fclose(f);
  return models;
}

// One search of the differential evolution (modelloop jointfind.cpp:744-812 with difeloop :702-716, nextgen :599-684, startpop
// :686-699) over nvar free parameters that sit at positions [lo, lo + nvar) of the vectors; under a nested model a vector is
// expanded to the full parameter list before it goes to the device.  best[0..np) = the peak (as searched), best[np] = -log joint
// density there.  A whole generation is one device call.
void joint_model_search(ima2p_lmode *LM, int np, int lo, int nvar, const std::vector<double> &upper_full, const NestedModel *nested,
                        unsigned long long &st, std::vector<double> &best) {
  const int depop = nvar * 100, hi = lo + nvar;      // DEFAULTPOPSIZEMULTIPLIER (:765)
  const double recrate = 0.9, fweight = 0.8;
  auto uni = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return ((double)(st >> 11) + 0.5) / 9007199254740992.0; };
  std::vector<double> lowerb(np, 0.0000001), upperb(upper_full);
  if (nested) {                                      // bounds of the free parameters (:691-692)
    std::vector<double> lf(np, 0.0000001);
    nested->reduce(lf.data(), lowerb.data());
    nested->reduce(upper_full.data(), upperb.data());
  }
  // entries outside [lo, hi) are not part of the search: the device does not read what they map to, they stay at 1
  std::vector<double> pop((size_t)depop * np, 1.0), trial((size_t)depop * np, 1.0), full, fpop(depop), ftrial(depop);
  if (nested) full.assign((size_t)depop * np, 1.0);
  best.assign(np + 1, 0.0);
  auto evaluate = [&](std::vector<double> &x, std::vector<double> &fx) {
    if (nested) for (int i = 0; i < depop; i++) nested->expand(&x[(size_t)i * np], &full[(size_t)i * np]);
    ck(ima2p_lmode_jointp(LM, nested ? full.data() : x.data(), depop, 0, fx.data(), nullptr), "joint density");
  };
  auto startpop = [&]() {
    for (int i = 0; i < depop; i++) for (int j = lo; j < hi; j++) pop[(size_t)i * np + j] = lowerb[j] + uni() * (upperb[j] - lowerb[j]);
    evaluate(pop, fpop);
  };
  double global_pd = 1e200;
  int newstart = 0, countloop = 0;
  startpop();
  for (int i = 0; i < depop; i++) if (fpop[i] < global_pd) global_pd = fpop[i];
  do {
    if (newstart > 0) startpop();
    double lowpd, hipd;
    do {                                             // difeloop: generations until the population has collapsed on a peak
      for (int i = 0; i < depop; i++) {
        const int A = (int)(uni() * depop) % depop, B = (int)(uni() * depop) % depop, Cv = (int)(uni() * depop) % depop;
        for (int j = lo; j < hi; j++) {
          if (uni() < recrate) {
            const double c = pop[(size_t)Cv * np + j];
            double t = c + fweight * (pop[(size_t)A * np + j] - pop[(size_t)B * np + j]);
            if (t < lowerb[j]) t = c - uni() * (c - lowerb[j]);           // move only part of the way towards the bound
            if (t > upperb[j]) t = c + uni() * (upperb[j] - c);
            trial[(size_t)i * np + j] = t;
          } else trial[(size_t)i * np + j] = pop[(size_t)i * np + j];
        }
      }
      evaluate(trial, ftrial);
      lowpd = 1e200; hipd = -1e200;
      for (int i = 0; i < depop; i++) {
        if (ftrial[i] < fpop[i]) { fpop[i] = ftrial[i]; for (int j = lo; j < hi; j++) pop[(size_t)i * np + j] = trial[(size_t)i * np + j]; }
        if (fpop[i] < lowpd) lowpd = fpop[i];
        if (fpop[i] > hipd) hipd = fpop[i];
      }
    } while (hipd - lowpd > 1.0e-7);                 // SPREADTOL
    const double local_pd = lowpd;
    if (fabs(local_pd - global_pd) < 1.0e-6) countloop++;               // PLOOPTOL
    else if (local_pd < global_pd) countloop = 0;
    if (local_pd < global_pd) {
      int k = 0;
      for (int i = 1; i < depop; i++) if (fpop[i] < fpop[k]) k = i;
      for (int j = 0; j < np; j++) best[j] = pop[(size_t)k * np + j];
      best[np] = fpop[k];
      global_pd = local_pd;
    }
    newstart++;
  } while (countloop < 2 && newstart < 10);          // LOOPMATCHCRITERIA, MAXRESTART
}

// findjointpeaks jointfind.cpp:1087-1170: the FULL model of a two-population analysis, or the two searches of a three-population
// analysis -- all population sizes (nowmodeltype 1), then all migration rates (2); jointp is a function of that family only
// (ima2p_lmode_set_joint_model).  After each full model, the nested models of its type (-w file) with their likelihood-ratio
// statistics against it.  More populations: the reference refuses too (:1076-1081).
void print_joint_peak(FILE *f, ima2p_lmode *LM, long long G, const std::vector<std::string> &name, int nq, int nm, const std::vector<double> &prior_max,
                      int npops, unsigned long long seed, const std::string &nestfname, double thetaprior, double mprior) {
  fprintf(f, "Joint Peak Locations and Posterior Probabilities\n================================================\n");
  fprintf(f, "  estimates based on %lld sampled genealogies\n", G);
  if (npops != 2 && npops != 3) { fprintf(f, "  the joint search is defined for two- and three-population models\n\n"); return; }
  std::vector<NestedModel> nested;
  if (!nestfname.empty()) {
    nested = read_nested_models(nestfname, npops, nq, nm, thetaprior, mprior);
    fprintf(f, "  nested model filename:%s\n", nestfname.c_str());
  }
  static const char *modelstart[3] = {"FULL", "ALL POPULATION SIZE PARAMETERS", "ALL MIGRATION PARAMETERS"};     // modelstartstr :163
  const int type0 = npops == 2 ? 0 : 1, type1 = npops == 2 ? 0 : 2;
  fprintf(f, "\nModel#  Model Description\n");
  for (int t = type0, j = 0; t <= type1; t++) {
    fprintf(f, "%2d     %s\n", ++j, modelstart[t]);
    for (const NestedModel &M : nested) if (M.type == t) fprintf(f, "%2d   %s\n", ++j, M.name.c_str());
  }
  fprintf(f, "\nModel#\tlog(P)\t#terms\tdf\t2LLR\tESS");
  const int np = nq + nm;
  for (int i = 0; i < np; i++) if (i < nq || prior_max[i] > 0.000001) fprintf(f, "\t%s", name[i].c_str());
  fprintf(f, "\n");
  unsigned long long st = seed * 6364136223846793005ull + 1442695040888963407ull;
  bool any_boundary = false;
  auto print_vals = [&](const double *v, int lo, int hi, const std::vector<int> *listed) {      // printjointpeakvals :814-840
    for (int i = 0; i < np; i++) {
      if (i < lo || i >= hi) { fprintf(f, "\t-"); continue; }
      const bool br = listed && (*listed)[i];
      fprintf(f, v[i] < 0.001 ? (br ? "\t[%.5lf]" : "\t%.5lf") : (br ? "\t[%.4lf]" : "\t%.4lf"), v[i]);
    }
    fprintf(f, "\n");
  };
  for (int t = type0, j = 0; t <= type1; t++) {
    const int lo = t == 2 ? nq : 0, hi = t == 1 ? nq : np;
    ck(ima2p_lmode_set_joint_model(LM, t), "joint model");
    std::vector<double> best, full(np);
    joint_model_search(LM, np, lo, hi - lo, prior_max, nullptr, st, best);
    double q = 0, e = 0;
    ck(ima2p_lmode_jointp(LM, best.data(), 1, 1, &q, &e), "joint density");
    const double holdml = best[np];
    fprintf(f, "%d\t%s\t%d\t-\t-\t%s", ++j, logpfmt(-best[np]).c_str(), hi - lo, logpfmt(e).c_str());
    print_vals(best.data(), lo, hi, nullptr);
    for (const NestedModel &M : nested) if (M.type == t) {
      joint_model_search(LM, np, lo, M.nfree, prior_max, &M, st, best);
      M.expand(best.data(), full.data());
      ck(ima2p_lmode_jointp(LM, full.data(), 1, 1, &q, &e), "joint density");
      fprintf(f, "%d\t%s\t%d\t%d%s\t%s\t%s", ++j, logpfmt(-best[np]).c_str(), M.nfree, (hi - lo) - M.nfree, M.boundary ? "*" : "",
              logpfmt(2 * (best[np] - holdml)).c_str(), logpfmt(e).c_str());
      print_vals(full.data(), lo, hi, &M.listed);
      any_boundary |= M.boundary;
    }
  }
  ck(ima2p_lmode_set_joint_model(LM, 0), "joint model");
  if (any_boundary) fprintf(f, "    * test distribution of 2LLR is a mixture\n");
  fprintf(f, "\n");
}

// The report sections that are sums over the sampled genealogies (printoutput, ima_main_mpi.cpp:4080-4110), written to f from
// `nrows` rows: used by L mode on the rows of a .ti file and by M mode on the rows the run has just saved.
// ---- the opening sections of the report, in the reference's layout ------------------------------------------------
// "INPUT AND STARTING INFORMATION": print_outputfile_info_string (ima_main_mpi.cpp:1565-1671) followed by what readdata
// (readata.cpp:916-1081, locus lines :640-832), reportparamcounts (initialize.cpp:2059-2069) and add_priorinfo_to_output
// (initialize.cpp:1557-1575) append to the same string while the run is set up.
struct RunInfo {
  std::string command_line, heating, calc, model, output;      // scan_commandline's echo strings (ima_main_mpi.cpp:449-487)
  std::string infile, outfile, ti, mcf_in, mcf_out;
  unsigned long long seed; long burn, nsave, every; int nchains, heatmode; double ha, hb;
  bool lmode, expo;
};

std::string sprintf_s(const char *fmt, ...) {
  char b[1024];
  va_list ap; va_start(ap, fmt); vsnprintf(b, sizeof b, fmt, ap); va_end(ap);
  return b;
}

std::string start_info(const RunInfo &R, ima2p_dataset *D, int npops, int nloci, const char *tree, int nq, int nm, int nsplit, double qmax, double mmax,
                       double tmax) {
  std::string s = "IMa2p_b200 - B200 engine for the IMa2p hot path; report sections in the layout of IMa2p version 1.0\n\n";
  s += "\nINPUT AND STARTING INFORMATION \n================================\n";
  s += "\nCommand line string : " + R.command_line + " \n";
  s += "  Input filename : " + R.infile + " \n  Output filename: " + R.outfile + " \n";
  s += sprintf_s("  Random number seed : %llu \n", R.seed);
  s += "  Heating terms on command line : " + R.heating + " \n  Calculation options on command line : " + R.calc + " \n";
  s += "  Model options on command line : " + R.model + " \n  Output options on command line : " + R.output + " \n";
  if (!R.lmode) {
    s += "- Run Duration - \n";
    s += sprintf_s("     Burn period, # steps: %li \n", R.burn);
    s += sprintf_s("     Record period, #saves: %d  #steps each: %li   total #steps: %li \n", (int)R.nsave, R.every, R.nsave * R.every);
    s += "- Metropolis Coupling -\n";
    if (R.nchains > 1) {
      s += sprintf_s("     Metropolis Coupling implemented using %d chains \n", R.nchains);
      if (R.heatmode == 0) s += sprintf_s("     Linear Increment Model   term: %.3f\n", R.ha);
      else if (R.heatmode == 1) s += sprintf_s("     Geometric Increment Model   term1: %.3f  term2: %.3f\n", R.ha, R.hb);
    } else s += "     None \n";
  }
  if (!R.mcf_in.empty()) s += "Initial Markov chain state space loaded from file: " + R.mcf_in + "\n";
  if (R.expo) s += "Exponential priors used for migration rate parameters\n";
  if (!R.mcf_out.empty()) s += "State of Markov chain saved to file : " + R.mcf_out + "\n";
  if (!R.lmode) s += "All genealogy information used for surface estimation printed to file: " + R.ti + "\n";
  char buf[2048];
  for (int k = 0; ima2p_dataset_text(D, 0, k, buf, sizeof buf) == IMA2P_OK; k++) s += std::string(k == 0 ? "\n" : "") + "Text from input file: " + buf + "\n";
  s += sprintf_s("\nNumber of sampled populations given in input file: %d \n- Population Names - \n", npops);
  for (int i = 0; i < npops; i++) { ck(ima2p_dataset_text(D, 1, i, buf, sizeof buf), "population name"); s += sprintf_s("Population %d : %s \n", i, buf); }
  s += std::string("\nPopulation Tree : ") + tree + "\n";
  s += sprintf_s("\nLocus Information\n-----------------\n\nNumber of loci: %d \nLocus#\tLocusname", nloci);
  for (int i = 0; i < npops; i++) s += sprintf_s("\tPop%d#", i);
  s += "\tModel\tInheritanceScalar\tMutationRatesPerYear\n";
  int nurates = 0, nkappas = 0;
  for (int li = 0; li < nloci; li++) {
    int info[8]; double hval; std::vector<int> samppop(npops); char name[64];
    ck(ima2p_dataset_locus(D, li, info, &hval, samppop.data(), name, sizeof name), "locus");
    s += sprintf_s("%d\t%s", li, name);
    for (int i = 0; i < npops; i++) s += sprintf_s("\t%3d", samppop[i]);
    const bool many = (info[7] & 1) != 0;
    s += info[0] == IMA2P_MODEL_HKY ? "\tHKY" : info[0] == IMA2P_MODEL_SW ? (many ? "\tSW_M" : "\tSW") : info[0] == IMA2P_MODEL_JOINT ? (many ? "\tIS+SW_M" : "\tIS+SW") : "\tIS";
    s += sprintf_s((info[7] & 2) ? "\t%5.3lf" : "\t%lf", hval);
    if (info[6] > 0) {
      std::vector<double> ur(info[6]);
      const int n = info[1], ns = info[2] > 0 ? info[2] : 1, nl = info[5];
      std::vector<int> seq((size_t)n * ns), mult(ns), A((size_t)nl * n), mn(nl), mx(nl); double pi[4];
      ck(ima2p_dataset_locus_data(D, li, seq.data(), mult.data(), A.data(), mn.data(), mx.data(), pi, ur.data()), "locus data");
      for (double u : ur) s += sprintf_s("\t%lg", u);
    }
    s += "\n";
    nurates += info[5];
    nkappas += info[0] == IMA2P_MODEL_HKY;
  }
  s += sprintf_s("\nParameter Counts\n----------------\n   Population sizes : %d\n   Migration rates  : %d\n   Parameters in the MCMC simulation\n", nq, nm);
  s += sprintf_s("      Splitting times : %d\n      Mutation scalars: %d\n      HKY Kappa (ti/tv) ratios: %d\n", nsplit, nurates, nkappas);
  s += sprintf_s("\nParameter Priors\n-----------------\n  Population size parameters maximum value : %.4lf \n", qmax);
  s += sprintf_s(R.expo ? "  Migration rate parameters exponential distribution mean : %.4lf \n" : "  Migration rate parameters maximum value: %.4lf \n", mmax);
  s += sprintf_s("  Splitting time : %.4lf\n\n", tmax);
  return s;
}

// "%.2e" with the exponent's sign and leading zeros dropped (shorten_e_num, utilities.cpp:1909-1921)
std::string short_e(double v) {
  char b[32];
  snprintf(b, sizeof b, "%.2e", (float)v);
  std::string s = b;
  const size_t e = s.find('e');
  while (e + 1 < s.size() && (s[e + 1] == '0' || s[e + 1] == '+')) s.erase(e + 1, 1);
  return s;
}

// printacceptancerates (output.cpp:528-571): one row per record, (#Tries, #Accp, %) per update type
struct RateRow { std::string name; std::vector<unsigned long long> tries, accp; };
void print_rates(FILE *f, const char *title, const std::vector<std::string> &types, const std::vector<RateRow> &rows) {
  fprintf(f, "\n%s\n", title);
  for (size_t i = 0; i < strlen(title); i++) fprintf(f, "-");
  fprintf(f, "\nUpdate Type:");
  for (const auto &t : types) fprintf(f, "\t%s\t", t.c_str());
  fprintf(f, "\n            ");
  for (size_t i = 0; i < types.size(); i++) fprintf(f, "\t#Tries\t#Accp\t%%");
  fprintf(f, "\n");
  for (const auto &r : rows) {
    fprintf(f, " %s", r.name.c_str());
    for (size_t i = r.name.size(); i < 13; i++) fprintf(f, " ");
    for (size_t i = 0; i < types.size(); i++) {
      fprintf(f, "\t%s\t%s", short_e((double)r.tries[i]).c_str(), short_e((double)r.accp[i]).c_str());
      if (r.tries[i] > 0) fprintf(f, "\t%.2f", (float)100 * r.accp[i] / r.tries[i]);
      else fprintf(f, "\tna");
    }
    fprintf(f, "\n");
  }
}

// asciicurve (output.cpp:428-524): a 75 x 50 character plot of a histogram's 1,000 points, y scaled to its maximum
void ascii_curve(FILE *f, const std::vector<double> &x, const std::vector<double> &y, const std::string &label, bool logscale, double recordstep) {
  constexpr int kLeft = 10, kPlotX = 75, kMaxX = kLeft + kPlotX, kPlotY = 50, kMaxY = kPlotY + 3;
  std::vector<std::string> g(kMaxY);
  double ymax = -1e10;
  const double ymin = 0.0;                               // "don't shift plot on y axis"
  for (int i = 0; i < kGrid; i++) if (ymax < y[i]) ymax = y[i];
  ymax /= recordstep;
  int xmax = kGrid - 1;
  if (!logscale) {
    xmax = -1;
    for (int i = kGrid - 1; i >= 0 && xmax == -1; i--) if (fabs(y[i]) > 1e-6) xmax = i;     // ASCIICURVEMINVAL
    if (xmax < 0) xmax = kGrid - 1;
  }
  auto pad = [&](std::string &s, size_t n, char c = ' ') { while (s.size() < n) s.push_back(c); };
  char tc[32];
  g[0] = label + " curve"; pad(g[0], kMaxX);
  snprintf(tc, sizeof tc, "%8.4g", ymax); g[1] = tc; pad(g[1], kMaxX);
  snprintf(tc, sizeof tc, "%8.4f", ymin); g[kPlotY] = tc; pad(g[kPlotY], kMaxX);
  g[kPlotY + 1] = std::string(kLeft, ' '); pad(g[kPlotY + 1], kMaxX, '-');
  g[kPlotY + 2] = std::string(kLeft, ' ');
  snprintf(tc, sizeof tc, "%8.4f", x[0]); g[kPlotY + 2] += tc;
  snprintf(tc, sizeof tc, "%8.4f", x[xmax]);
  if (logscale) g[kPlotY + 2] += "           Log Scale";
  pad(g[kPlotY + 2], kMaxX - strlen(tc) - 1);
  g[kPlotY + 2] += tc; pad(g[kPlotY + 2], kMaxX);
  for (int i = 2; i < kPlotY; i++) pad(g[i], kMaxX);
  for (int i = 1; i < kPlotY + 1; i++) g[i][kLeft] = '|';
  for (int i = 0; i <= xmax; i++) {
    const int yspot = (int)1 + kPlotY - (int)(kPlotY * (y[i] / recordstep - ymin) / (ymax - ymin));
    const int xspot = logscale ? (int)kLeft + 1 + (int)((kPlotX - 2) * (log(x[i]) - log(x[0])) / (2 * log(x[xmax])))
                               : (int)kLeft + 1 + (int)((kPlotX - 2) * (x[i] - x[0]) / (x[xmax] - x[0]));
    if (xspot < kMaxX && xspot >= kLeft + 1 && yspot < kMaxY && yspot >= 1 && xspot < (int)g[yspot].size()) g[yspot][xspot] = '*';
  }
  for (int i = 0; i < kMaxY; i++) fprintf(f, "%s \n", g[i].c_str());
  fprintf(f, "\n");
}

void report_sections(FILE *f, std::map<std::string, std::string> &opt, ima2p_modelspec *S, int npops, double qmax, double mmax, int expo,
                     const float *rowdata, long long nrows, bool loaded_from_ti) {
  int md[6];
  ima2p_modelspec_dims(S, md);
  const int nsplit = md[1], nq = md[3], nm = md[4], nwp = md[5], np = nq + nm;
  // parameter names as setup_iparams writes them: q<pop>, m<from>><to> (initialize.cpp:230-232, 466-470)
  std::vector<int> plist((size_t)npops * npops), addpop(npops + 1), droppops((size_t)(npops + 1) * 2), ptb(2 * npops), pte(2 * npops), ptd(2 * npops),
      qoff(nq + 1), qp(2 * npops * npops), qr(2 * npops * npops), moff(nm + 1), mp(nwp > 0 ? nwp : 1), mr(nwp > 0 ? nwp : 1), mc(nwp > 0 ? nwp : 1);
  ck(ima2p_modelspec_tables(S, plist.data(), addpop.data(), droppops.data(), ptb.data(), pte.data(), ptd.data(), qoff.data(), qp.data(), qr.data(), moff.data(),
                            mp.data(), mr.data(), mc.data()), "model tables");
  std::vector<std::string> name;
  for (int i = 0; i < nq; i++) name.push_back("q" + std::to_string(i));
  for (int i = 0; i < nm; i++) {
    const int k = moff[i];
    name.push_back("m" + std::to_string(plist[(size_t)mp[k] * npops + mr[k]]) + ">" + std::to_string(plist[(size_t)mp[k] * npops + mc[k]]));
  }
  std::vector<double> qmx(nq, qmax), qmn(nq, 0.0), mmx(nm, expo ? 20.0 * mmax : mmax), mmn(nm, 0.0), mmean(nm, expo ? mmax : 0.0);
  const int rowlen = 3 * nq + 2 * nm + nq + nm + 2 + nsplit;            // calc_gsampinf_length ginfo.cpp:306-316
  ima2p_lmode *LM = nullptr;
  ck(ima2p_lmode_create(&LM, 0, nq, nm, nsplit, qmx.data(), qmn.data(), mmx.data(), mmn.data(), mmean.data(), expo), "L mode");
  ck(ima2p_lmode_load(LM, rowdata, (int)nrows, rowlen, nrows), "uploading the genealogies");
  if (opt.count("p") && opt["p"].find('6') != std::string::npos) print_greater_than(f, LM, name, nq, nm, expo);
  if (!expo) {                                                          // ima_main_mpi.cpp:4086: not done for the exponential prior
    std::vector<double> mean(np), var(np), corr((size_t)np * np);
    ck(ima2p_lmode_moments(LM, mean.data(), var.data(), corr.data(), nullptr), "moments");
    print_moments(f, name, nq, mmx, mean, var, corr, npops);
  }
  std::vector<PopMigTerm> terms;
  std::vector<double> term_upper;
  {
    // period in which a parameter first appears / ends: a size parameter lives with its population (poptree b, e), a
    // migration parameter over the periods of its weight positions (initialize.cpp:237-243, 470-520)
    std::vector<double> pmax;
    std::vector<int> pb, pe;
    for (int i = 0; i < nq; i++) { pmax.push_back(qmx[i]); pb.push_back(ptb[i]); pe.push_back(pte[i]); }
    for (int i = 0; i < nm; i++) { pmax.push_back(mmx[i]); pb.push_back(mp[moff[i]]); pe.push_back(mp[moff[i + 1] - 1]); }
    // 2NM terms (-p5): every population of the tree but the root with each migration parameter leaving it
    // (surface_call_functions.cpp:409-437, histograms.cpp:751-766)
    if (opt.count("p") && opt["p"].find('5') != std::string::npos && nm > 0)
      for (int thetai = 0; thetai < 2 * npops - 2; thetai++)
        for (int mi = 0; mi < nm; mi++)
          if (atoi(name[nq + mi].c_str() + 1) == thetai) {
            PopMigTerm t; t.thetai = thetai; t.mi = mi;
            t.peakname = "2N" + std::to_string(thetai) + "M" + name[nq + mi].substr(1);
            t.histname = "2N" + std::to_string(thetai) + name[nq + mi];
            terms.push_back(t);
            term_upper.push_back(expo ? 20.0 * mmean[mi] : qmx[thetai] * mmx[mi] / 2.0);
          }
    print_marginal_peaks(f, LM, nrows, name, nq, nm, nsplit, pmax, pb, pe, terms, term_upper);
    // -c2 FINDJOINTPOSTERIOR (L mode only, ima_main_mpi.cpp:1466); -w <nested model file> implies it (:934-937)
    if (loaded_from_ti && ((opt.count("c") && opt["c"].find('2') != std::string::npos) || opt.count("w")))
      print_joint_peak(f, LM, nrows, name, nq, nm, pmax, npops, opt.count("s") ? strtoull(opt["s"].c_str(), nullptr, 10) : 1ull,
                       opt.count("w") ? opt["w"] : std::string(), atof(opt["q"].c_str()), opt.count("m") ? atof(opt["m"].c_str()) : 0.0);
  }
  // fillvec histograms.cpp:81-99: margincalc at the GRIDSIZE mid-bin points of every parameter (initialize.cpp:189-193, 237-242)
  std::vector<std::vector<double>> xs, ys;
  std::vector<std::string> hname;
  for (int p = 0; p < np; p++) {
    const double mx = p < nq ? qmx[p] : mmx[p - nq], mn = 0.0;
    if (p >= nq && !(mx > 0.000001)) continue;
    std::vector<double> x(kGrid), y(kGrid);
    for (int j = 0; j < kGrid; j++) x[j] = mn + ((j + 0.5) * (mx - mn)) / kGrid;
    ck(ima2p_lmode_margincalc(LM, p, x.data(), kGrid, 0.0, 0, y.data()), "marginal densities");
    xs.push_back(x); ys.push_back(y); hname.push_back(name[p]);
  }
  // printhistograms histograms.cpp:895-909
  fprintf(f, "\nHISTOGRAMS\n==========\n  Each histogram is given as %d pairs of values (i.e. two columns side by side).\n", kGrid);
  fprintf(f, "  In each case the left column is the value of the parameter or term (i.e. x value)\n  and the right column is the estimated posterior probability (i.e. y value).\n");
  fprintf(f, "  HPD (Highest Posterior Density) intervals are estimated, and may be incorrect: \n  Possible HPD footnotes: \n");
  fprintf(f, "       '?' HPD interval may be incorrect due to multiple peaks\n");
  fprintf(f, "       '#' HPD may not be useful - posterior density does not reach low levels near either the upper or the lower limit of the prior\n");
  fprintf(f, "\nNUMBER OF GROUPS OF HISTOGRAM TABLES : %d\n\n", 2 + (int)(!terms.empty()));
  std::vector<std::vector<double>> acx, acy;          // the split-time histograms, kept for the ASCII curves
  std::vector<std::string> acn;
  if (nsplit > 0) {
    // histogram group 1 in L mode: the split times of the loaded rows binned as recordval does (ima_main_mpi.cpp:2746-2781,
    // 3420-3431), scaled to a density (histograms.cpp:496-509); mutation-scalar histograms need an M-mode run
    const double tmax = atof(opt["t"].c_str());
    std::vector<std::vector<double>> tx(nsplit, std::vector<double>(kGrid)), ty(nsplit, std::vector<double>(kGrid, 0.0));
    std::vector<double> tscale(nsplit, (kGrid / tmax) / (double)nrows), tb(nsplit, 0.0), ta(nsplit, 0.0);
    std::vector<std::string> tn;
    for (int k = 0; k < nsplit; k++) {
      tn.push_back("t" + std::to_string(k));
      for (int j = 0; j < kGrid; j++) tx[k][j] = 0.0 + ((j + 0.5) * (tmax * 1.0 - 0.0)) / kGrid;
      for (long long r = 0; r < nrows; r++) {
        const int b = (int)(kGrid * rowdata[(size_t)r * rowlen + (rowlen - nsplit) + k] / tmax);
        if (b < 0) tb[k] += 1; else if (b >= kGrid) ta[k] += 1; else ty[k][b] += 1;
      }
    }
    fprintf(f, "HISTOGRAM GROUP 1: MARGINAL DISTRIBUTION VALUES AND HISTOGRAMS OF PARAMETERS IN MCMC\n");
    fprintf(f, "----------------------------------------------------------------------------------\n");
    fprintf(f, "    curve height is an estimate of marginal posterior probability\n");
    if (loaded_from_ti) fprintf(f, "  IMa LOAD TREES MODE  - splittime values loaded from *.ti file, mutation rate scalar histograms are not available \n");
    else fprintf(f, "  split times of the saved genealogies\n");
    write_histograms(f, tn, tx, ty, tscale, true, tb, ta);
    acn = tn; acx = tx; acy = ty;
  }
  fprintf(f, "\n\nHISTOGRAM GROUP 2: MARGINAL DISTRIBUTION VALUES AND HISTOGRAMS OF POPULATION SIZE AND MIGRATION PARAMETERS\n");
  fprintf(f, "--------------------------------------------------------------------------------------------------------\n");
  fprintf(f, "       curve height is an estimate of marginal posterior probability\n");
  write_histograms(f, hname, xs, ys);
  if (!terms.empty()) {
    // print_populationmigrationrate_histograms histograms.cpp:648-815: the grid ends at the last bin whose density exceeds 1e-9
    std::vector<std::vector<double>> px, py;
    std::vector<std::string> pn;
    for (size_t i = 0; i < terms.size(); i++) {
      std::vector<double> x(kGrid), y(kGrid);
      for (int j = 0; j < kGrid; j++) x[j] = (j + 0.5) * term_upper[i] / kGrid;
      ck(ima2p_lmode_popmig(LM, terms[i].thetai, terms[i].mi, x.data(), kGrid, 0, y.data()), "2NM density");
      int j = kGrid - 1;
      while (j >= 0 && !(y[j] > 1e-9)) j--;
      const double maxx = j < 0 ? term_upper[i] : x[j];
      for (int q = 0; q < kGrid; q++) x[q] = (q + 0.5) * maxx / kGrid;
      ck(ima2p_lmode_popmig(LM, terms[i].thetai, terms[i].mi, x.data(), kGrid, 0, y.data()), "2NM density");
      px.push_back(x); py.push_back(y); pn.push_back(terms[i].histname);
    }
    fprintf(f, "\n\nHISTOGRAM GROUP 3: POPULATION MIGRATION (2NM) POSTERIOR PROBABILITY HISTOGRAMS\n");
    fprintf(f, "-----------------------------------------------------------------------------\n");
    fprintf(f, "     curve height is an estimate of the posterior probability\n");
    fprintf(f, "      each term is the product of a population parameter (e.g. q0) and a migration rate (e.g.m0>1) \n");
    fprintf(f, "      migration rates are in the coalescent (backwards in times), so that a population migration rate of \n");
    fprintf(f, "        q1m0>1  is the population rate (forward in time) at which population 1 receives migrants from population 0\n");
    write_histograms(f, pn, px, py);
  }
  // callasciicurves ima_main_mpi.cpp:3905-4000: population sizes, migration rates with a prior above MPRIORMIN, split times
  // (their counts over the recorded steps)
  fprintf(f, "\n\nASCII Curves - Approximate Posterior Densities \n===================================================\n");
  for (size_t i = 0; i < hname.size(); i++) ascii_curve(f, xs[i], ys[i], hname[i], false, 1.0);
  for (size_t i = 0; i < acn.size(); i++) ascii_curve(f, acx[i], acy[i], acn[i], false, (double)nrows);
  {
    const int seconds = (int)difftime(time(nullptr), g_starttime);          // ima_main_mpi.cpp:4149-4155
    fprintf(f, "Time Elapsed : %d hours, %d minutes, %d seconds \n\n", seconds / 3600, seconds / 60 - 60 * (seconds / 3600), seconds - 60 * (seconds / 60));
  }
  fprintf(f, "\nEND OF OUTPUT\n");
  ima2p_lmode_destroy(LM);
}

int run_lmode(std::map<std::string, std::string> &opt, ima2p_modelspec *S, int npops, double qmax, double mmax, int expo, const std::string &info) {
  int md[6];
  ima2p_modelspec_dims(S, md);
  const int rowlen = 3 * md[3] + 2 * md[4] + md[3] + md[4] + 2 + md[1];
  const std::string ti = opt["v"] + ".ti";
  long long nrows = 0;
  ck(ima2p_ti_load(ti.c_str(), rowlen, nullptr, 0, &nrows), "loading the genealogy file");
  if (nrows < 1) die("no genealogies in " + ti, 13);
  if (nrows > 1100000) nrows = 1100000;
  std::vector<float> rows((size_t)nrows * rowlen);
  ck(ima2p_ti_load(ti.c_str(), rowlen, rows.data(), nrows, &nrows), "loading the genealogy file");
  FILE *f = fopen(opt["o"].c_str(), "w");
  if (!f) die("cannot create the output file", 2);
  fprintf(f, "%s\n\nLOAD TREES (L) MODE INFORMATION\n============================================================================\n", info.c_str());
  fprintf(f, "  Base filename for loading files with sampled genealogies: %s*.ti\n  loaded %lld genealogies from genealogy file  %s\n", opt["v"].c_str(), nrows, ti.c_str());
  report_sections(f, opt, S, npops, qmax, mmax, expo, rows.data(), nrows, true);
  fclose(f);
  printf("IMa2p_b200: L mode, %lld genealogies from %s, report in %s\n", nrows, ti.c_str(), opt["o"].c_str());
  return 0;
}

}  // namespace

int main(int argc, char **argv) {
  g_starttime = time(nullptr);
  std::map<std::string, std::string> opt;
  // -cap N (this build only): migration events per genealogy the device pools start with (they grow when they fill)
  static const char *known[] = {"hn", "hf", "ha", "hb", "i", "o", "q", "m", "t", "b", "l", "d", "s", "j", "r", "f", "p", "z", "v", "cap", "c", "w", nullptr};
  RunInfo R{};
  for (int a = 1; a < argc; a++) {
    if (argv[a][0] != '-') die(std::string("command line: unexpected word ") + argv[a], 5);
    std::string w = argv[a] + 1;
    if (w[0] == 'W') w[0] = 'w';                    // the reference's options are case-blind (ima_main_mpi.cpp:640)
    const char c1 = (char)toupper((unsigned char)w[0]);
    if (c1 == 'H') R.heating += std::string(" ") + argv[a];
    if (c1 == 'J') R.model += std::string(" ") + argv[a];
    if (c1 == 'C') R.calc += std::string(" ") + argv[a];
    if (c1 == 'P') R.output += std::string(" ") + argv[a];
    std::string key;
    for (int k = 0; known[k]; k++) if (w.compare(0, strlen(known[k]), known[k]) == 0) { key = known[k]; break; }
    if (key.empty()) die("command line: option -" + w + " is not part of this build (M-mode hot path only)", 5);
    std::string val = w.substr(key.size());
    const bool flag_only = key == "hf" || key == "j" || key == "r" || key == "p" || key == "c";
    if (val.empty() && !flag_only) { if (a + 1 >= argc) die("command line: -" + key + " needs a value", 5); val = argv[++a]; }
    if (key == "j" && val != "7") die("model option -j" + val + " is not part of this build", 5);
    if (key == "c" && val != "2") die("calculation option -c" + val + " is not part of this build", 5);
    opt[key] = val;
  }
  const bool lmode = opt.count("r") && opt["r"] == "0";
  if (lmode && !opt.count("v")) die("-r0 invoked without -v information, i.e. no base name for files containing genealogys was given on the command line", 8);
  if (lmode) { opt.emplace("b", "0"); opt.emplace("l", "1"); }
  for (const char *need : {"i", "o", "q", "t", "b", "l"}) if (!opt.count(need)) die(std::string("command line: -") + need + " is required", 5);
  const double qmax = atof(opt["q"].c_str()), mmax = opt.count("m") ? atof(opt["m"].c_str()) : 0.0, tmax = atof(opt["t"].c_str());
  Ranks ranks;
  ranks.init();
  const int nlocal = opt.count("hn") ? atoi(opt["hn"].c_str()) : 1, nchains = nlocal * ranks.world, expo = opt.count("j") ? 1 : 0;      // -hn: chains per rank
  const long burn = atol(opt["b"].c_str()), nsave = atol(opt["l"].c_str()), every = opt.count("d") ? atol(opt["d"].c_str()) : 100;
  // no -s: seeded from the clock as the reference does (ima_main_mpi.cpp:1413), and said so in the report; a run continued from
  // a state file (-f) mixes the time into the seed it was given so that it does not replay the first run's random streams
  unsigned long long seed = opt.count("s") ? strtoull(opt["s"].c_str(), nullptr, 10) : (unsigned long long)time(nullptr);
  if (opt.count("f") && opt.count("s")) seed = seed * 6364136223846793005ull + (unsigned long long)time(nullptr);
  if (nchains < 1 || burn < 0 || nsave < 1 || every < 1 || !(tmax > 0)) die("command line: bad value", 5);
  for (int a = 1; a < argc; a++) R.command_line += std::string(" ") + argv[a];
  R.infile = opt["i"]; R.outfile = opt["o"]; R.ti = opt["o"] + ".ti"; R.seed = seed; R.burn = burn; R.nsave = nsave; R.every = every; R.nchains = nchains;
  R.heatmode = opt.count("hf") ? (opt["hf"] == "g" ? 1 : opt["hf"] == "s" ? 2 : 0) : 0;
  R.ha = opt.count("ha") ? atof(opt["ha"].c_str()) : 0.05; R.hb = opt.count("hb") ? atof(opt["hb"].c_str()) : 0.0;
  R.lmode = lmode; R.expo = expo != 0;
  if (opt.count("f")) R.mcf_in = opt["f"];
  if (opt.count("r") && !lmode) R.mcf_out = opt["o"] + ".mcf";
  const time_t starttime = g_starttime;

  ima2p_dataset *D = nullptr;
  ck(ima2p_dataset_read(opt["i"].c_str(), &D), "reading data");
  int npops = 0, nloci = 0;
  char tree[256];
  ck(ima2p_dataset_dims(D, &npops, &nloci, tree, sizeof tree), "data");
  ima2p_modelspec *S = nullptr;
  ck(ima2p_modelspec_create(&S, npops, tree, qmax, expo ? 20.0 * mmax : mmax, expo, expo ? mmax : 0.0, 0, 1.0), "model");   // -j7: -m is the mean, plotted to 20 means
  if (lmode) {
    int lmd[6];
    ima2p_modelspec_dims(S, lmd);
    const int rc = run_lmode(opt, S, npops, qmax, mmax, expo, start_info(R, D, npops, nloci, tree, lmd[3], lmd[4], lmd[1], qmax, mmax, tmax));
    ima2p_modelspec_free(S);
    ima2p_dataset_free(D);
    return rc;
  }
  int md[6];
  ima2p_modelspec_dims(S, md);
  const int nsplit = md[1], rootpop = 2 * npops - 2;

  std::vector<Locus> loci(nloci);
  for (int li = 0; li < nloci; li++) {
    Locus &L = loci[li];
    L.samppop.resize(npops);
    char name[64];
    ck(ima2p_dataset_locus(D, li, L.info, &L.hval, L.samppop.data(), name, sizeof name), "locus");
    const int n = L.info[1], ns = L.info[2], nlinked = L.info[5];
    L.seq.assign((size_t)n * (ns > 0 ? ns : 1), 0); L.mult.assign(ns > 0 ? ns : 1, 1); L.A.assign((size_t)nlinked * n, 0);
    L.minA.assign(nlinked, 0); L.maxA.assign(nlinked, 0);
    ck(ima2p_dataset_locus_data(D, li, L.seq.data(), L.mult.data(), L.A.data(), L.minA.data(), L.maxA.data(), L.pi, nullptr), "locus data");
    if (nlinked > IMA2P_MAX_LINKED) die("more linked stepwise parts than this build keeps on the device", 5);
  }

  // set_tvalues (initialize.cpp:1945-1960): split times evenly spaced inside the prior; the genealogies a run starts with
  // (with -f they only fix the infinite-sites constant, as the reference's do)
  std::vector<double> tv(nsplit > 0 ? nsplit : 1);
  for (int k = 0; k < nsplit; k++) tv[k] = (k + 1.0) / (nsplit + 1.0) * tmax;
  std::vector<Tree> start(nloci);
  {
    unsigned rng = (unsigned)seed * 2654435761u + 12345u;
    for (int li = 0; li < nloci; li++) start[li] = start_genealogy(loci[li], npops, rootpop, nsplit > 0 ? tv[nsplit - 1] : 0.0, rng);
  }
  // Migration capacity: the reference grows an edge's list whenever it fills (checkmig, utilities.cpp:1365-1383).  Here the
  // pools start with `capacity` events per genealogy (-cap, default 96) and double when a proposal did not fit; a proposal
  // that did not fit was rejected unseen, which is reported on stderr with the count, never silently.
  int capacity = opt.count("cap") ? atoi(opt["cap"].c_str()) : 96;
  if (capacity < 8 || capacity > 8000) die("command line: -cap must be between 8 and 8000", 5);
  ima2p_engine *E = nullptr;
  ck(ima2p_engine_create(&E, ranks.local, nlocal, nchains, ranks.rank * nlocal, nloci, capacity, seed), "engine");
  ck(ima2p_engine_set_model_spec(E, S), "model");
  for (int li = 0; li < nloci; li++) {
    Locus &L = loci[li];
    const int model = L.info[0], n = L.info[1], ns = L.info[2], nlinked = L.info[5];
    const double sumlogk = (model == IMA2P_MODEL_IS || model == IMA2P_MODEL_JOINT) ? sumlogk_of(L, start[li]) : 0.0;
    std::vector<int> lo(nlinked, 0), hi(nlinked, 0);          // allele range rule of build_gtree.cpp:519-520, 721-722
    for (int a = (model == IMA2P_MODEL_JOINT); a < nlinked && (model == IMA2P_MODEL_SW || model == IMA2P_MODEL_JOINT); a++) {
      const int mn = L.minA[a], mx = L.maxA[a];
      lo[a] = (mx + mn) / 2 - (mx - mn) > 1 ? (mx + mn) / 2 - (mx - mn) : 1;
      hi[a] = (mx + mn) / 2 + (mx - mn) < 1000 ? (mx + mn) / 2 + (mx - mn) : 1000;
    }
    ck(ima2p_engine_set_locus(E, li, model, n, ns, L.info[3], L.hval, L.samppop.data(), ns > 0 ? L.seq.data() : nullptr,
                              model == IMA2P_MODEL_HKY ? L.mult.data() : nullptr, nlinked, lo.data(), hi.data(), sumlogk), "locus");
  }
  ck(ima2p_engine_finalize(E), "finalize");
  if (nchains > 1) {
    const int mode = opt.count("hf") ? (opt["hf"] == "g" ? 1 : opt["hf"] == "s" ? 2 : 0) : 0;
    ck(ima2p_engine_set_heating(E, mode, opt.count("ha") ? atof(opt["ha"].c_str()) : 0.05, opt.count("hb") ? atof(opt["hb"].c_str()) : 0.0), "heating");
  }
  std::vector<double> tmaxv(nsplit > 0 ? nsplit : 1, tmax), tminv(nsplit > 0 ? nsplit : 1, 0.0);
  ck(ima2p_engine_set_update_priors(E, tmaxv.data(), tminv.data(), 0.0, 0.0, 0.0, 0.0), "priors");

  if (opt.count("f")) {
    ck(ima2p_engine_read_mcf(E, (ranks.world > 1 ? opt["f"] + "." + std::to_string(ranks.rank) : opt["f"]).c_str()), "loading the state file");
  } else {
    for (int li = 0; li < nloci; li++) {
      const Locus &L = loci[li];
      const int n = L.info[1], nl = 2 * n - 1, nlinked = L.info[5];
      const Tree &T = start[li];
      std::vector<int> moff(nl + 1, 0), mp(1, 0), A((size_t)nlinked * nl, 0);
      std::vector<double> mt(1, 0.0), u(IMA2P_MAX_LINKED, 1.0);
      for (int a = 0; a < nlinked; a++) {                     // internal allele states: those of a descendant tip
        for (int i = 0; i < n; i++) A[(size_t)a * nl + i] = L.A[(size_t)a * n + i];
        for (int k = n; k < nl; k++) A[(size_t)a * nl + k] = A[(size_t)a * nl + T.up0[k]];
      }
      for (int c = 0; c < nlocal; c++) {
        if (li == 0) ck(ima2p_engine_set_chain(E, c, tv.data()), "chain");
        ck(ima2p_engine_set_genealogy(E, c, li, T.up0.data(), T.up1.data(), T.down.data(), T.pop.data(), T.time.data(), moff.data(), mt.data(),
                                      mp.data(), T.root, T.roottime, u.data(), 2.0, L.pi, nlinked > 0 && (L.info[0] == IMA2P_MODEL_SW || L.info[0] == IMA2P_MODEL_JOINT) ? A.data() : nullptr),
           "starting genealogy");
      }
    }
    ck(ima2p_engine_upload(E), "upload");
    ck(ima2p_engine_eval(E), "evaluating the starting state");
  }
  ck(ima2p_engine_set_update_schedule(E, nsplit > 0 ? 3 : 0, 5), "schedule");

  const int swaptries = nchains > 1 ? (nchains / 10 > 1 ? nchains / 10 : 1) : 0;                 // ima_main_mpi.cpp:1378
  if (ranks.world > 1) {
    // exchange tables: publish mine, open the others', attach; nobody steps before everybody is attached
    void *table = nullptr; uint64_t tbytes = 0;
    ck(ima2p_engine_exchange_create(E, &table, &tbytes), "exchange table");
    unsigned char hd[64];
    ck(ima2p_ipc_export(table, hd), "exchange handle");
    ranks.put("xch", hd, 64);
    std::vector<void *> tables(ranks.world, nullptr);
    for (int r = 0; r < ranks.world; r++) {
      if (r == ranks.rank) continue;
      const std::vector<unsigned char> h = ranks.get("xch", r, 64);
      ck(ima2p_ipc_import(ranks.local, h.data(), &tables[r]), "opening a peer's exchange table");
    }
    ck(ima2p_engine_exchange_attach(E, tables.data()), "attaching the exchange");
    ranks.barrier("attached");
  }
  auto run_steps = [&](int n, const char *what) {
    if (ranks.world > 1) ck(ima2p_engine_run_sharded(E, n, swaptries, nullptr), what);
    else ck(ima2p_engine_run(E, n, swaptries, nullptr), what);
  };
  int dims[5];
  ima2p_engine_dims(E, dims);
  const int rowlen = dims[4];
  const std::string ti = opt["o"] + ".ti";
  std::string header = "Command line string : ";
  for (int a = 0; a < argc; a++) header += std::string(argv[a]) + " ";
  if (ranks.rank == 0) ck(ima2p_ti_create(ti.c_str(), header.c_str()), "creating the .ti file");
  if (ranks.rank == 0) printf("IMa2p_b200: %d populations %s, %d loci, %d chains, burn %ld steps, %ld genealogies every %ld steps\n", npops, tree, nloci, nchains, burn, nsave, every);
  // after every stretch of steps: did a proposal fail to fit the migration pools?  Then the pools double (or the run ends with
  // the reference's own error when a genealogy of that size cannot be held at all: IMERR_MIGARRAYTOOBIG, utilities.hpp:43)
  unsigned long long dropped_seen = 0;
  auto check_capacity = [&](const char *phase) {
    uint64_t c8[8];
    ck(ima2p_engine_counters(E, c8), "counters");
    if (c8[7] == dropped_seen) return;
    const unsigned long long lost = c8[7] - dropped_seen;
    dropped_seen = c8[7];
    const int bigger = capacity * 2;
    if (ima2p_engine_grow_capacity(E, bigger) != IMA2P_OK) {
      fprintf(stderr, "IMa2: %llu genealogy proposals needed more than %d migration events (%s)\n", lost, capacity, phase);
      die(" too many migrations in array " + std::to_string(capacity), 22);
    }
    fprintf(stderr, "IMa2: %llu genealogy proposals needed more than %d migration events and were rejected unseen (%s); the pools now hold %d -- "
            "start with -cap %d to avoid this\n", lost, capacity, phase, bigger, bigger);
    capacity = bigger;
  };
  for (long done = 0; done < burn;) { const int n = burn - done > 1000 ? 1000 : (int)(burn - done); run_steps(n, "burn-in"); done += n; check_capacity("burn-in"); }
  // the reference's update-rate and swap tables start counting after the burn-in (reset_after_burn)
  uint64_t cnt0[8], ucnt0[4];
  const int nur = [&] { int k = 0; for (auto &L : loci) k += L.info[5]; return k; }();
  std::vector<uint64_t> cg0((size_t)nloci * 3), ct0((size_t)(nsplit > 0 ? nsplit : 1) * 4), cu0((size_t)nur * 2), ca0((size_t)nchains * 2, 0), cg(cg0), ct(ct0), cu(cu0), ca(ca0);
  ck(ima2p_engine_counters(E, cnt0), "counters");
  ck(ima2p_engine_update_counters(E, ucnt0), "counters");
  ck(ima2p_engine_cold_counters(E, cg0.data(), ct0.data(), cu0.data(), ca0.data()), "counters");
  double hilike = -1e20, hiprob = -1e20;                                   // checkhighs (output.cpp:207-240), at the recorded steps
  std::vector<double> hilocus(nloci, -1e20), locus_pdg(nloci);
  std::vector<double> chain4((size_t)nchains * 4);
  std::vector<float> rows, row(rowlen), allrows;
  std::vector<double> tsum(nsplit > 0 ? nsplit : 1, 0.0);
  long saved = 0;
  while (saved < nsave) {
    run_steps((int)every, "run");
    check_capacity("sampling");
    if (ranks.world > 1) {
      // the cold chain may live on any rank: its record comes to rank 0 through the exchange (ima2p_engine_cold_message)
      std::vector<double> msg((size_t)rowlen + 2 + nloci);
      ck(ima2p_engine_cold_message(E, msg.data(), nullptr), "reading the cold chain");
      saved++;
      if (ranks.rank != 0) continue;
      saved--;
      for (int i = 0; i < rowlen; i++) row[i] = (float)msg[i];
      if (msg[rowlen] > hiprob) hiprob = msg[rowlen];
      if (msg[rowlen + 1] > hilike) hilike = msg[rowlen + 1];
      for (int li = 0; li < nloci; li++) if (msg[rowlen + 2 + li] > hilocus[li]) hilocus[li] = msg[rowlen + 2 + li];
    } else {
      int present = 0;
      ck(ima2p_engine_step_report(E, chain4.data(), row.data(), &present, nullptr), "reading the cold chain");
      if (!present) die("the cold chain is not on this device");
      for (int c = 0; c < nchains; c++) {
        if (chain4[(size_t)c * 4] != 1.0) continue;
        if (chain4[(size_t)c * 4 + 1] > hiprob) hiprob = chain4[(size_t)c * 4 + 1];
        if (chain4[(size_t)c * 4 + 2] > hilike) hilike = chain4[(size_t)c * 4 + 2];
        ck(ima2p_engine_fetch_chain_pdg(E, c, locus_pdg.data()), "reading the likelihoods");
        for (int li = 0; li < nloci; li++) if (locus_pdg[li] > hilocus[li]) hilocus[li] = locus_pdg[li];
      }
    }
    rows.insert(rows.end(), row.begin(), row.end());
    allrows.insert(allrows.end(), row.begin(), row.end());
    for (int k = 0; k < nsplit; k++) tsum[k] += row[rowlen - nsplit + k];
    saved++;
    if (rows.size() >= (size_t)rowlen * 256 || saved == nsave) { ck(ima2p_ti_append(ti.c_str(), rows.data(), (long long)(rows.size() / rowlen), rowlen), "writing the .ti file"); rows.clear(); }
  }
  uint64_t cnt[8], ucnt[4];
  ck(ima2p_engine_counters(E, cnt), "counters");
  ck(ima2p_engine_update_counters(E, ucnt), "counters");
  const std::string outname = opt["o"];
  ck(ima2p_engine_cold_counters(E, cg.data(), ct.data(), cu.data(), ca.data()), "counters");
  if (opt.count("r") && ranks.world > 1) ck(ima2p_engine_write_mcf(E, (outname + ".mcf." + std::to_string(ranks.rank)).c_str()), "writing the state file");
  if (ranks.world > 1) {
    // Update counts are kept where the updates were made (the cold chain moves between ranks): every rank publishes what it
    // counted since the burn-in, rank 0 adds them up.  Steps and swaps are replayed identically on every rank: not added.
    std::vector<uint64_t> mine;
    for (int i = 0; i < 8; i++) mine.push_back(cnt[i]);
    for (int i = 0; i < 4; i++) mine.push_back(ucnt[i]);
    for (size_t i = 0; i < cg.size(); i++) mine.push_back(cg[i] - cg0[i]);
    for (size_t i = 0; i < ct.size(); i++) mine.push_back(ct[i] - ct0[i]);
    for (size_t i = 0; i < cu.size(); i++) mine.push_back(cu[i] - cu0[i]);
    ranks.put("cnt", mine.data(), mine.size() * 8);
    if (ranks.rank != 0) {
      ranks.finish();
      ima2p_engine_destroy(E); ima2p_modelspec_free(S); ima2p_dataset_free(D);
      return 0;
    }
    for (size_t i = 0; i < cg.size(); i++) { cg[i] -= cg0[i]; cg0[i] = 0; }
    for (size_t i = 0; i < ct.size(); i++) { ct[i] -= ct0[i]; ct0[i] = 0; }
    for (size_t i = 0; i < cu.size(); i++) { cu[i] -= cu0[i]; cu0[i] = 0; }
    for (int r = 1; r < ranks.world; r++) {
      const std::vector<unsigned char> raw = ranks.get("cnt", r, mine.size() * 8);
      const uint64_t *o = (const uint64_t *)raw.data();
      for (int i : {1, 2, 3, 4, 7}) cnt[i] += o[i];
      for (int i = 0; i < 4; i++) ucnt[i] += o[8 + i];
      size_t at = 12;
      for (size_t i = 0; i < cg.size(); i++) cg[i] += o[at++];
      for (size_t i = 0; i < ct.size(); i++) ct[i] += o[at++];
      for (size_t i = 0; i < cu.size(); i++) cu[i] += o[at++];
    }
  }
  FILE *f = fopen(outname.c_str(), "w");
  if (!f) die("cannot create the output file", 2);
  const unsigned long long poststeps = cnt[0] - cnt0[0];
  fprintf(f, "%s", start_info(R, D, npops, nloci, tree, md[3], md[4], nsplit, qmax, mmax, tmax).c_str());
  // printrunbasics (output.cpp:166-206)
  fprintf(f, "\n\nMCMC INFORMATION\n===========================\n\n");
  fprintf(f, "Number of steps in burnin: %10d\nNumber of steps in chain following burnin: %10d \n", (int)burn, (int)poststeps);
  fprintf(f, "Number of steps between recording : %d  Number of record steps: %d \n", (int)every, (int)saved);
  fprintf(f, "Number of steps between saving genealogy information: %d  Number of genealogies saved per locus: %d \n", (int)every, (int)saved);
  const int seconds = (int)difftime(time(nullptr), starttime);
  fprintf(f, "\nTime Elapsed : %d hours, %d minutes, %d seconds \n\n", seconds / 3600, seconds / 60 - 60 * (seconds / 3600), seconds - 60 * (seconds / 60));
  fprintf(f, "Highest Sampled Joint P(G) (log) : %10.3f \nHighest Joint P(D|G) (log) : %10.3f \n\n\n", hiprob, hilike);
  fprintf(f, "Highest P(D|G) (log) for each Locus \n\tLocus\tP(D|G)\n");
  for (int li = 0; li < nloci; li++) fprintf(f, "\t%d\t%.3f\n", li, hilocus[li]);
  fprintf(f, "\n");
  // callprintacceptancerates (ima_main_mpi.cpp:3473-3899): the cold chain's tries and accepts since the burn-in
  if (nsplit > 0) {
    std::vector<RateRow> tr;
    for (int k = 0; k < nsplit; k++)
      tr.push_back({"t" + std::to_string(k), {ct[k * 4 + 2] - ct0[k * 4 + 2], ct[k * 4] - ct0[k * 4]}, {ct[k * 4 + 3] - ct0[k * 4 + 3], ct[k * 4 + 1] - ct0[k * 4 + 1]}});
    print_rates(f, "Update Rates -- Population Splitting Times", {"NielsenWakeley", "RannalaYang"}, tr);
  }
  {
    std::vector<RateRow> gr;
    for (int li = 0; li < nloci; li++)
      gr.push_back({"gtree_" + std::to_string(li), {poststeps, poststeps, poststeps}, {cg[li * 3] - cg0[li * 3], cg[li * 3 + 1] - cg0[li * 3 + 1], cg[li * 3 + 2] - cg0[li * 3 + 2]}});
    print_rates(f, "Update Rates -- Genealogies", {"branch     ", "topology   ", "tmrca      "}, gr);
  }
  if (nur > 1) {
    std::vector<RateRow> ur;
    int j = 0;
    for (int li = 0; li < nloci; li++)
      for (int a = 0; a < loci[li].info[5]; a++, j++) {
        const bool sw = loci[li].info[0] == IMA2P_MODEL_SW || (loci[li].info[0] == IMA2P_MODEL_JOINT && a > 0);
        ur.push_back({sw ? std::to_string(li) + "SW" + std::to_string(a) : std::to_string(li) + "u ", {cu[j * 2] - cu0[j * 2]}, {cu[j * 2 + 1] - cu0[j * 2 + 1]}});     // initialize.cpp:1453-1466
      }
    print_rates(f, "Update Rates -- Mutation Rate Scalars", {"scalar update"}, ur);
    // kappa of an HKY locus is proposed and decided together with the locus' scalar (update_mc_params.cpp:258-277, 310-317):
    // its record counts the same proposals
    std::vector<RateRow> kr;
    j = 0;
    for (int li = 0; li < nloci; j += loci[li].info[5], li++)
      if (loci[li].info[0] == IMA2P_MODEL_HKY) kr.push_back({std::to_string(li) + "_Ka", {cu[j * 2] - cu0[j * 2]}, {cu[j * 2 + 1] - cu0[j * 2 + 1]}});     // initialize.cpp:1502
    if (!kr.empty()) print_rates(f, "Update Rates -- HKY Model Kappa parameter", {"kappa update"}, kr);
  }
  if (nchains > 1) {
    // printchaininfo (swapchains.cpp:664-690, 815-823): swaps between adjacent temperatures (only betas move here, so the
    // per-chain table of the serial build has no counterpart)
    fprintf(f, "\nCHAIN SWAPPING:");
    if (R.heatmode == 0) fprintf(f, " Linear Increment  term: %.4f\n", R.ha);
    else if (R.heatmode == 1) fprintf(f, " Geometric Increment  term1: %.4f term2: %.4f\n", R.ha, R.hb);
    else fprintf(f, "\n");
    fprintf(f, "-----------------------------------------------------------------------------\n");
    std::vector<double> betas(nchains);
    ck(ima2p_engine_get_betas(E, betas.data()), "betas");
    std::sort(betas.begin(), betas.end(), [](double a, double b) { return a > b; });
    fprintf(f, "Temp1    Temp2    #Swaps    #Attempts    Rate\n");
    for (int r = 0; r + 1 < nchains; r++) {
      const unsigned long long att = ca[r * 2] - ca0[r * 2], sw = ca[r * 2 + 1] - ca0[r * 2 + 1];
      if (att > 0) fprintf(f, " %7.4f    %7.4f    %5llu    %5llu    %7.4lf\n", betas[r], betas[r + 1], sw, att, (float)sw / (float)att);
      else fprintf(f, " %7.4f    %7.4f    %5llu    %5llu    na\n", betas[r], betas[r + 1], sw, att);
    }
  }
  fprintf(f, "\nENGINE INFORMATION (all chains, whole run)\n------------------------------------------\nsteps %llu  genealogy updates %llu  accepted %.4f  topology changes %.4f\n", (unsigned long long)cnt[0],
          (unsigned long long)cnt[1], (double)cnt[2] / (double)(cnt[1] ? cnt[1] : 1), (double)cnt[3] / (double)(cnt[1] ? cnt[1] : 1));
  fprintf(f, "chain swaps %llu of %llu attempts\nsplit-time updates accepted %llu of %llu   mutation-scalar updates accepted %llu of %llu\n", (unsigned long long)cnt[6],
          (unsigned long long)cnt[5], (unsigned long long)ucnt[1], (unsigned long long)ucnt[0], (unsigned long long)ucnt[3], (unsigned long long)ucnt[2]);
  fprintf(f, "proposals dropped for migration capacity %llu\ngenealogies saved %ld in %s\n", (unsigned long long)cnt[7], saved, ti.c_str());
  for (int k = 0; k < nsplit; k++) fprintf(f, "mean of t%d over the saved genealogies %.6f\n", k, tsum[k] / saved);
  // the same sections an L-mode run on out.ti would write, from the rows just saved (printoutput at the end of an M-mode run)
  g_throw_instead_of_exit = true;
  try {
    report_sections(f, opt, S, npops, qmax, mmax, expo, allrows.data(), saved, false);
  } catch (const std::exception &ex) {
    fprintf(f, "\nthe report sections over the saved genealogies could not be completed: %s\n(the genealogies are in %s; run L mode, -r0 -v, on them)\n", ex.what(), ti.c_str());
  }
  g_throw_instead_of_exit = false;
  fclose(f);
  if (opt.count("r") && ranks.world == 1) ck(ima2p_engine_write_mcf(E, (outname + ".mcf").c_str()), "writing the state file");
  ranks.finish();
  printf("IMa2p_b200: done, %ld genealogies in %s\n", saved, ti.c_str());
  ima2p_engine_destroy(E);
  ima2p_modelspec_free(S);
  ima2p_dataset_free(D);
  return 0;
}
