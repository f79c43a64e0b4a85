/* TEST INFRASTRUCTURE -- reference-side oracle harness (never shipped, never linked into the product).
 *
 * Builds the *unmodified* reference (arunsethuraman/ima2p, /root/reference/src) into one binary with
 * its own main() renamed, so that the parity fixtures under tests/golden/ and the CPU baseline of
 * bench.py come from the reference's own code:  `#include "ima_main_mpi.cpp"` pulls the reference
 * driver in where it lies (nothing is copied).  The route follows SURVEY.md Appendix B.
 *
 *   ref_harness MODE OUT [key=value ...] -- <IMa2p command line>
 *
 *   state    burn=N                      full model tables + every chain x locus genealogy + values
 *                                        recomputed from scratch the way init_p() does
 *                                        (mcmcfile.cpp:130-193)
 *   updates  burn=N n=M                  accepted updategenealogy() proposals: tree before/after and
 *                                        the forward/reverse getmprob() values (update_gtree.cpp:654-663)
 *   kat                                  tables of the numerics (utilities.cpp, update_gtree_common.cpp:108-304)
 *   lmode    burn=N rows=G every=K       sampled .ti rows (savegsampinf) and marginp/margincalc/jointp values
 *   bench    burn=N iters=K full=0|1     times the updategenealogy() loop (or whole qupdate steps)
 *   lbench   rows=G evals=E              times margincalc / jointp over G synthetic-from-run rows
 *   stock    [seed=K]                    the reference's own main() on the IMa2p command line, unchanged (the serial build
 *                                        always seeds its generator with seed * currentid = 0; seed=K makes replicates possible)
 */
#define main ima2p_reference_main
#define setseeds harness_seed_hook      /* start()'s call (ima_main_mpi.cpp:1721) goes through the hook below */
#include "ima_main_mpi.cpp"
#undef setseeds
#undef main
void setseeds (int seed);
static long g_stock_seed = -1;
void harness_seed_hook (int seed) { setseeds (g_stock_seed >= 0 ? (int) g_stock_seed : seed); }
#include "update_gtree_common.hpp"      /* getmprob :850, calcmrate :462 */

#include <chrono>
#include <vector>
#include <string>
#include <map>

/* exported by shims.cpp */
double harness_integrate_coalescent_term (int cc, double fc, double hcc, double max, double min);
double harness_integrate_migration_term (int cm, double fm, double max, double min);
double harness_integrate_migration_term_expo_prior (int cm, double fm, double exmean);
double harness_swapweight (int ci, int cj);
double harness_swapweight_bwprocesses (double sumi, double sumj, double betai, double betaj);
double harness_marginp (int param, int firsttree, int lasttree, double x);
void harness_jointp_setup (void);
void harness_jointp_set_type (int type);
double harness_calcx (int ei, int pnum, int mode);
double harness_greater_than (int kind, int i, int j);
void print_greater_than_tests (FILE * outfile);
double calc_popmig (int thetai, int mi, double x, int prob_or_like);
double calc_pop_expomig (int thetai, int mi, double x, int prob_or_like);
double marginpopmig (int mi, int firsttree, int lasttree, double x, int thetai);
double marginpop_expomig (int mi, int firsttree, int lasttree, double x, int thetai);
void print_means_variances_correlations (FILE * outfile);
double harness_get_sumlogk (int li);
double jointp (double *x, int calc_ess, double *effective_n);
extern struct edgemiginfo oldedgemig, oldsismig, newedgemig, newsismig;
void harness_init_p (void);                             /* mcmcfile.cpp:130, file-static */
extern int rootmove;

double harness_ry_last_mh (void);
double harness_mc_last_mh (void);
extern double thermosum[];                              /* marglike.cpp:23 */

static FILE *jo;
static std::map<std::string, std::string> kv;

/* hooks named by the uniform() macro of the update_t_RY / update_mc_params shims */
static int g_force_accept = 0;
static std::vector<double> g_ulog;
double harness_uniform_ry ()
{
  double u = uniform ();
  return g_force_accept ? 1e-300 : u;       /* log(1e-300) is below any finite Metropolis-Hastings term */
}
static long g_nw_count = 0, g_nw_force = -1;
double harness_nw_last_mh (void);
double harness_uniform_nw ()
{
  double u = uniform ();
  g_nw_count++;
  return g_nw_count == g_nw_force ? 0.0 : u;
}
double harness_uniform_mc ()
{
  double u = uniform ();
  g_ulog.push_back (u);
  return u;
}

static long
kvl (const char *k, long dflt)
{
  auto it = kv.find (k);
  return it == kv.end ()? dflt : atol (it->second.c_str ());
}

static double
nowsec ()
{
  return std::chrono::duration<double> (std::chrono::steady_clock::now ().time_since_epoch ()).count ();
}

static void
jd (double v)
{
  if (v != v)
    fprintf (jo, "\"nan\"");
  else if (v > DBL_MAX)
    fprintf (jo, "\"inf\"");
  else if (v < -DBL_MAX)
    fprintf (jo, "\"-inf\"");
  else
    fprintf (jo, "%.17g", v);
}

static void
jdarr (const char *name, const double *a, int n, const char *tail)
{
  fprintf (jo, "\"%s\":[", name);
  for (int i = 0; i < n; i++)
  {
    if (i)
      fputc (',', jo);
    jd (a[i]);
  }
  fprintf (jo, "]%s", tail);
}

static void
jiarr (const char *name, const int *a, int n, const char *tail)
{
  fprintf (jo, "\"%s\":[", name);
  for (int i = 0; i < n; i++)
    fprintf (jo, "%s%d", i ? "," : "", a[i]);
  fprintf (jo, "]%s", tail);
}

/* gweight flattened: cc/fc/hcc[k][i] for k=0..numsplittimes, i<npops-k ; mc/fm[k][i][j] k<lastperiodnumber */
static void
dump_gweight (const char *name, struct genealogy_weights *gw, const char *tail)
{
  int k, i, j, first;
  fprintf (jo, "\"%s\":{\"cc\":[", name);
  for (first = 1, k = 0; k <= numsplittimes; k++)
    for (i = 0; i < npops - k; i++, first = 0)
      fprintf (jo, "%s%d", first ? "" : ",", gw->cc[k][i]);
  fprintf (jo, "],\"fc\":[");
  for (first = 1, k = 0; k <= numsplittimes; k++)
    for (i = 0; i < npops - k; i++, first = 0)
    {
      if (!first)
        fputc (',', jo);
      jd (gw->fc[k][i]);
    }
  fprintf (jo, "],\"hcc\":[");
  for (first = 1, k = 0; k <= numsplittimes; k++)
    for (i = 0; i < npops - k; i++, first = 0)
    {
      if (!first)
        fputc (',', jo);
      jd (gw->hcc[k][i]);
    }
  fprintf (jo, "],\"mc\":[");
  if (!modeloptions[NOMIGRATION])
    for (first = 1, k = 0; k < lastperiodnumber; k++)
      for (i = 0; i < npops - k; i++)
        for (j = 0; j < npops - k; j++, first = 0)
          fprintf (jo, "%s%d", first ? "" : ",", gw->mc[k][i][j]);
  fprintf (jo, "],\"fm\":[");
  if (!modeloptions[NOMIGRATION])
    for (first = 1, k = 0; k < lastperiodnumber; k++)
      for (i = 0; i < npops - k; i++)
        for (j = 0; j < npops - k; j++, first = 0)
        {
          if (!first)
            fputc (',', jo);
          jd (gw->fm[k][i][j]);
        }
  fprintf (jo, "]}%s", tail);
}

static void
dump_tree (int ci, int li, const char *tail)
{
  struct genealogy *G = &C[ci]->G[li];
  struct edge *gt = G->gtree;
  int nl = L[li].numlines, i, j, ai;
  fprintf (jo, "{\"root\":%d,\"roottime\":", G->root);
  jd (G->roottime);
  fprintf (jo, ",\"up0\":[");
  for (i = 0; i < nl; i++)
    fprintf (jo, "%s%d", i ? "," : "", gt[i].up[0]);
  fprintf (jo, "],\"up1\":[");
  for (i = 0; i < nl; i++)
    fprintf (jo, "%s%d", i ? "," : "", gt[i].up[1]);
  fprintf (jo, "],\"down\":[");
  for (i = 0; i < nl; i++)
    fprintf (jo, "%s%d", i ? "," : "", gt[i].down);
  fprintf (jo, "],\"pop\":[");
  for (i = 0; i < nl; i++)
    fprintf (jo, "%s%d", i ? "," : "", gt[i].pop);
  fprintf (jo, "],\"time\":[");
  for (i = 0; i < nl; i++)
  {
    if (i)
      fputc (',', jo);
    jd (gt[i].time);
  }
  fprintf (jo, "],\"mig\":[");
  for (i = 0; i < nl; i++)
  {
    fprintf (jo, "%s[", i ? "," : "");
    for (j = 0; gt[i].mig[j].mt > -0.5; j++)
    {
      if (j)
        fputc (',', jo);
      jd (gt[i].mig[j].mt);
      fprintf (jo, ",%d", gt[i].mig[j].mp);
    }
    fputc (']', jo);
  }
  fprintf (jo, "]");
  if (L[li].model == STEPWISE || L[li].model == JOINT_IS_SW)
  {
    fprintf (jo, ",\"A\":[");
    for (ai = 0; ai < L[li].nlinked; ai++)
    {
      fprintf (jo, "%s[", ai ? "," : "");
      for (i = 0; i < nl; i++)
        fprintf (jo, "%s%d", i ? "," : "", (ai == 0 && L[li].model == JOINT_IS_SW) ? 0 : gt[i].A[ai]);
      fputc (']', jo);
    }
    fprintf (jo, "],\"dlikeA\":[");
    for (ai = 0; ai < L[li].nlinked; ai++)
    {
      fprintf (jo, "%s[", ai ? "," : "");
      for (i = 0; i < nl; i++)
      {
        if (i)
          fputc (',', jo);
        jd ((ai == 0 && L[li].model == JOINT_IS_SW) ? 0.0 : gt[i].dlikeA[ai]);
      }
      fputc (']', jo);
    }
    fprintf (jo, "]");
  }
  fprintf (jo, "}%s", tail);
}

static void
dump_model (void)
{
  int i, k, li, j;
  fprintf (jo, "\"model\":{\"npops\":%d,\"numtreepops\":%d,\"numsplittimes\":%d,\"numpopsizeparams\":%d,"
           "\"nummigrateparams\":%d,\"nomigration\":%d,\"expomigrationprior\":%d,\"gsampinflength\":%d,\"gbeta\":",
           npops, numtreepops, numsplittimes, numpopsizeparams, nummigrateparams,
           modeloptions[NOMIGRATION], modeloptions[EXPOMIGRATIONPRIOR], calc_gsampinf_length ());
  jd (gbeta);
  fprintf (jo, ",\"calcmarginallikelihood\":%d,\"rootpop\":%d,", calcoptions[CALCMARGINALLIKELIHOOD], C[0]->rootpop);
  fprintf (jo, "\"plist\":[");
  for (k = 0; k < npops; k++)
  {
    fprintf (jo, "%s[", k ? "," : "");
    for (i = 0; i < npops - k; i++)
      fprintf (jo, "%s%d", i ? "," : "", C[0]->plist[k][i]);
    fputc (']', jo);
  }
  fprintf (jo, "],\"addpop\":[");
  for (k = 0; k <= numsplittimes; k++)
    fprintf (jo, "%s%d", k ? "," : "", k == 0 ? -1 : C[0]->addpop[k]);
  fprintf (jo, "],\"droppops\":[");
  for (k = 0; k <= numsplittimes; k++)
    fprintf (jo, "%s[%d,%d]", k ? "," : "", k == 0 ? -1 : C[0]->droppops[k][0], k == 0 ? -1 : C[0]->droppops[k][1]);
  fprintf (jo, "],\"poptree\":[");
  for (i = 0; i < numtreepops; i++)
    fprintf (jo, "%s{\"b\":%d,\"e\":%d,\"down\":%d}", i ? "," : "", C[0]->poptree[i].b, C[0]->poptree[i].e,
             C[0]->poptree[i].down);
  fprintf (jo, "],\"itheta\":[");
  for (i = 0; i < numpopsizeparams; i++)
  {
    fprintf (jo, "%s{\"max\":", i ? "," : "");
    jd (itheta[i].pr.max);
    fprintf (jo, ",\"min\":");
    jd (itheta[i].pr.min);
    fputc (',', jo);
    jiarr ("p", itheta[i].wp.p, itheta[i].wp.n, ",");
    jiarr ("r", itheta[i].wp.r, itheta[i].wp.n, "}");
  }
  fprintf (jo, "],\"imig\":[");
  for (i = 0; i < nummigrateparams; i++)
  {
    fprintf (jo, "%s{\"max\":", i ? "," : "");
    jd (imig[i].pr.max);
    fprintf (jo, ",\"min\":");
    jd (imig[i].pr.min);
    fprintf (jo, ",\"mean\":");
    jd (modeloptions[EXPOMIGRATIONPRIOR] ? imig[i].pr.mean : 0.0);
    fputc (',', jo);
    jiarr ("p", imig[i].wp.p, imig[i].wp.n, ",");
    jiarr ("r", imig[i].wp.r, imig[i].wp.n, ",");
    jiarr ("c", imig[i].wp.c, imig[i].wp.n, "}");
  }
  fprintf (jo, "],\"nomigrationchecklist\":{");
  jiarr ("p", nomigrationchecklist.p, nomigrationchecklist.n, ",");
  jiarr ("r", nomigrationchecklist.r, nomigrationchecklist.n, ",");
  jiarr ("c", nomigrationchecklist.c, nomigrationchecklist.n, "}},\n");
  fprintf (jo, "\"loci\":[");
  for (li = 0; li < nloci; li++)
  {
    fprintf (jo, "%s{\"model\":%d,\"numgenes\":%d,\"numlines\":%d,\"numsites\":%d,\"totsites\":%d,\"numbases\":%d,\"nlinked\":%d,\"hval\":",
             li ? ",\n" : "", L[li].model, L[li].numgenes, L[li].numlines, L[li].numsites, L[li].totsites,
             L[li].numbases, L[li].nlinked);
    jd (L[li].hval);
    fprintf (jo, ",\"sumlogk\":");
    jd (harness_get_sumlogk (li));
    fputc (',', jo);
    jiarr ("samppop", L[li].samppop, npops, ",");
    jiarr ("umodel", L[li].umodel, L[li].nlinked, ",");
    jiarr ("minA", L[li].minA, L[li].nlinked, ",");
    jiarr ("maxA", L[li].maxA, L[li].nlinked, ",");
    if (L[li].model == HKY)
      jiarr ("mult", L[li].mult, L[li].numsites, ",");
    fprintf (jo, "\"seq\":[");
    if (L[li].model == HKY || L[li].model == INFINITESITES || L[li].model == JOINT_IS_SW)
      for (j = 0; j < L[li].numgenes; j++)
      {
        fprintf (jo, "%s[", j ? "," : "");
        for (i = 0; i < L[li].numsites; i++)
          fprintf (jo, "%s%d", i ? "," : "", L[li].seq[j][i]);
        fputc (']', jo);
      }
    fprintf (jo, "]}");
  }
  fprintf (jo, "],\n");
}

/* from-scratch evaluation of every chain, as init_p() (mcmcfile.cpp:130-193) does after a reload */
static void
recompute_all (void)
{
  harness_init_p ();
}

static void dump_chain (int ci);
static void
dump_state (void)
{
  int ci, li;
  fprintf (jo, "{");
  dump_model ();
  fprintf (jo, "\"chains\":[");
  for (ci = 0; ci < numchains; ci++)
  {
    fprintf (jo, "%s", ci ? ",\n" : "");
    dump_chain (ci);
  }
  fprintf (jo, "]}\n");
}

static void
dump_chain (int ci)
{
  int li;
  {
    fprintf (jo, "{\"beta\":");
    jd (beta[ci]);
    fputc (',', jo);
    jdarr ("tvals", C[ci]->tvals, numsplittimes, ",");
    dump_gweight ("allgweight", &C[ci]->allgweight, ",");
    jdarr ("qintegrate", C[ci]->allpcalc.qintegrate, numpopsizeparams, ",");
    jdarr ("mintegrate", C[ci]->allpcalc.mintegrate, nummigrateparams, ",");
    fprintf (jo, "\"probg\":");
    jd (C[ci]->allpcalc.probg);
    fprintf (jo, ",\"pdg\":");
    jd (C[ci]->allpcalc.pdg);
    fprintf (jo, ",\"G\":[");
    for (li = 0; li < nloci; li++)
    {
      struct genealogy *G = &C[ci]->G[li];
      fprintf (jo, "%s{", li ? ",\n" : "");
      jdarr ("uvals", G->uvals, L[li].nlinked, ",");
      fprintf (jo, "\"kappa\":");
      jd (L[li].model == HKY ? G->kappaval : 0.0);
      fputc (',', jo);
      jdarr ("pi", G->pi, 4, ",");
      fprintf (jo, "\"pdg\":");
      jd (G->pdg);
      fputc (',', jo);
      jdarr ("pdg_a", G->pdg_a, L[li].nlinked, ",");
      fprintf (jo, "\"length\":");
      jd (G->length);
      fprintf (jo, ",\"tlength\":");
      jd (G->tlength);
      fprintf (jo, ",\"mignum\":%d,", G->mignum);
      dump_gweight ("gweight", &G->gweight, ",");
      fprintf (jo, "\"tree\":");
      dump_tree (ci, li, "}");
    }
    fprintf (jo, "]}");
  }
}

static std::string g_start_trees;
static void
capture_start_trees (void)      /* chain 0's genealogies as JSON, kept for the trace fixture */
{
  char *buf = NULL;
  size_t blen = 0;
  FILE *keep = jo;
  jo = open_memstream (&buf, &blen);
  fputc ('[', jo);
  for (int li = 0; li < nloci; li++)
    dump_tree (0, li, li + 1 < nloci ? "," : "");
  fputc (']', jo);
  fclose (jo);
  jo = keep;
  g_start_trees = buf;
  free (buf);
}

static void
dump_tree_all (void)
{
  fputs (g_start_trees.c_str (), jo);
}

static void
do_burn (long burn)
{
  for (long s = 0; s < burn; s++)
  {
    qupdate (0, 0, 1);
    step++;
  }
}

static void
dump_emi (const char *name, struct edgemiginfo *em, const char *tail)
{
  int j;
  fprintf (jo, "\"%s\":{\"edgeid\":%d,\"pop\":%d,\"fpop\":%d,\"b\":%d,\"e\":%d,\"mpall\":%d,\"upt\":", name,
           em->edgeid, em->pop, em->fpop, em->b, em->e, em->mpall);
  jd (em->upt);
  fprintf (jo, ",\"dnt\":");
  jd (em->dnt);
  fprintf (jo, ",\"mtall\":");
  jd (em->mtall);
  fputc (',', jo);
  jdarr ("mtimeavail", em->mtimeavail, npops, ",");
  jiarr ("mp", em->mp, npops, ",");
  fprintf (jo, "\"mig\":[");
  for (j = 0; em->mig[j].mt > -0.5; j++)
  {
    if (j)
      fputc (',', jo);
    jd (em->mig[j].mt);
    fprintf (jo, ",%d", em->mig[j].mp);
  }
  fprintf (jo, "]}%s", tail);
}

static void
mode_updates (long burn, long n)
{
  int ci, li, topol, tmrca, acc, first = 1;
  long done = 0, tries = 0;
  do_burn (burn);
  recompute_all ();
  fprintf (jo, "{");
  dump_model ();
  fprintf (jo, "\"tvals\":[");
  for (ci = 0; ci < numchains; ci++)
  {
    fprintf (jo, "%s[", ci ? "," : "");
    for (int k = 0; k < numsplittimes; k++)
    {
      if (k)
        fputc (',', jo);
      jd (C[ci]->tvals[k]);
    }
    fputc (']', jo);
  }
  fprintf (jo, "],\n\"updates\":[");
  while (done < n && tries < 200 * n)
  {
    for (ci = 0; ci < numchains && done < n; ci++)
      for (li = 0; li < nloci && done < n; li++)
      {
        /* the "before" tree must be written before the call; keep it in a memory stream */
        char *buf = NULL;
        size_t blen = 0;
        FILE *keep = jo;
        jo = open_memstream (&buf, &blen);
        dump_tree (ci, li, "");
        fclose (jo);
        jo = keep;
        double oldpdg = C[ci]->G[li].pdg, oldprobg = C[ci]->allpcalc.probg;
        acc = updategenealogy (ci, li, &topol, &tmrca);
        tries++;
        if (acc)
        {
          double fwd = getmprob (ci, &newedgemig, &newsismig, &oldedgemig, &oldsismig);
          double rev = getmprob (ci, &oldedgemig, &oldsismig, &newedgemig, &newsismig);
          fprintf (jo, "%s{\"ci\":%d,\"li\":%d,\"topolchange\":%d,\"tmrcachange\":%d,\"rootmove\":%d,\"fwd\":", first ? "" : ",\n", ci,
                   li, topol, tmrca, rootmove);
          first = 0;
          jd (fwd);
          fprintf (jo, ",\"rev\":");
          jd (rev);
          fprintf (jo, ",\"oldpdg\":");
          jd (oldpdg);
          fprintf (jo, ",\"newpdg\":");
          jd (C[ci]->G[li].pdg);
          fprintf (jo, ",\"oldprobg\":");
          jd (oldprobg);
          fprintf (jo, ",\"newprobg\":");
          jd (C[ci]->allpcalc.probg);
          fputc (',', jo);
          dump_emi ("oldedgemig", &oldedgemig, ",");
          dump_emi ("oldsismig", &oldsismig, ",");
          dump_emi ("newedgemig", &newedgemig, ",");
          dump_emi ("newsismig", &newsismig, ",");
          dump_gweight ("newgweight", &C[ci]->G[li].gweight, ",");
          fprintf (jo, "\"before\":%s,\"after\":", buf);
          dump_tree (ci, li, "}");
          done++;
        }
        free (buf);
      }
  }
  fprintf (jo, "],\"tries\":%ld}\n", tries);
}

static void
mode_kat (void)
{
  int a, i, first;
  double x;
  /* incomplete gamma on a grid that straddles the x < a+1 switch (utilities.cpp:1053-1122) */
  static const double xs[] = { 1e-6, 1e-3, 0.01, 0.1, 0.5, 0.9, 1.0, 1.5, 2.0, 3.2, 5.0, 7.5, 10.0, 17.0, 33.0, 64.0,
    100.0, 250.0, 500.0, 999.0, 1001.0, 1500.0, 2500.0, 4000.0, 1e4
  };
  static const int as[] = { 0, 1, 2, 3, 4, 5, 7, 10, 16, 29, 50, 99, 100, 250, 500, 999, 1000, 1450, 2000, 3000 };
  const int nx = sizeof (xs) / sizeof (xs[0]), na = sizeof (as) / sizeof (as[0]);
  fprintf (jo, "{\"uppergamma\":[");
  for (first = 1, a = 0; a < na; a++)
    for (i = 0; i < nx + 4; i++, first = 0)
    {
      /* extra points hugging x = a+1 */
      x = i < nx ? xs[i] : (as[a] + 1.0) * (i == nx ? 0.999 : i == nx + 1 ? 1.0 : i == nx + 2 ? 1.001 : 0.5);
      if (as[a] == 0 && x <= 0)
        continue;
      fprintf (jo, "%s[%d,", first ? "" : ",", as[a]);
      jd (x);
      fputc (',', jo);
      jd (uppergamma (as[a], x));
      fputc (']', jo);
    }
  fprintf (jo, "],\n\"lowergamma\":[");
  for (first = 1, a = 1; a < na; a++)
    for (i = 0; i < nx + 4; i++, first = 0)
    {
      x = i < nx ? xs[i] : (as[a] + 1.0) * (i == nx ? 0.999 : i == nx + 1 ? 1.0 : i == nx + 2 ? 1.001 : 0.5);
      fprintf (jo, "%s[%d,", first ? "" : ",", as[a]);
      jd (x);
      fputc (',', jo);
      jd (lowergamma (as[a], x));
      fputc (']', jo);
    }
  fprintf (jo, "],\n\"bessi\":[");
  {
    static const double bx[] = { 0.0, 1e-8, 1e-3, 0.1, 0.5, 1.0, 2.5, 3.74, 3.75, 3.76, 5.0, 10.0, 25.0, 60.0, 150.0, 400.0, 699.0,
      700.0, 701.0
    };
    for (first = 1, a = 0; a <= 40; a += (a < 6 ? 1 : 5))
      for (i = 0; i < (int) (sizeof (bx) / sizeof (bx[0])); i++, first = 0)
      {
        fprintf (jo, "%s[%d,", first ? "" : ",", a);
        jd (bx[i]);
        fputc (',', jo);
        jd (bessi (a, bx[i]));
        fputc (']', jo);
      }
  }
  fprintf (jo, "],\n\"eexp\":[");
  {
    static const double ex[] = { -1e4, -745.2, -700.0, -312.7, -100.0, -23.5, -10.0, -2.302585092994046, -1.0, -0.5, -1e-9, 0.0, 1e-9,
      0.3, 0.6931471805599453, 1.0, 2.302585092994046, 7.7, 42.0, 100.0, 333.3, 700.0, 709.0, 1e4
    };
    for (i = 0; i < (int) (sizeof (ex) / sizeof (ex[0])); i++)
    {
      double m;
      int z;
      eexp (ex[i], &m, &z);
      fprintf (jo, "%s[", i ? "," : "");
      jd (ex[i]);
      fputc (',', jo);
      jd (m);
      fprintf (jo, ",%d]", z);
    }
  }
  fprintf (jo, "],\n\"logfact\":[");
  for (i = 0; i < 64; i++)
  {
    if (i)
      fputc (',', jo);
    jd (logfact[i * i]);
  }
  /* integrate_coalescent_term / integrate_migration_term: all branches (update_gtree_common.cpp:108-296) */
  fprintf (jo, "],\n\"integrate_coalescent_term\":[");
  {
    static const int ccs[] = { 0, 1, 2, 3, 10, 29, 95, 400, 1450, 2900 };
    static const double fcs[] = { 0.0, 1e-9, 1e-3, 0.37, 2.0, 9.9, 30.8, 343.7, 1500.0, 9000.0, 60000.0 };
    static const double maxs[] = { 1.0, 10.0, 50.0 };
    for (first = 1, a = 0; a < 10; a++)
      for (i = 0; i < 11; i++)
        for (int m = 0; m < 3; m++)
        {
          if (ccs[a] > 0 && fcs[i] <= 0)
            continue;
          double hcc = (m == 1) ? 0.0 : ccs[a] * log (0.75);
          fprintf (jo, "%s[%d,", first ? "" : ",", ccs[a]);
          first = 0;
          jd (fcs[i]);
          fputc (',', jo);
          jd (hcc);
          fputc (',', jo);
          jd (maxs[m]);
          fputc (',', jo);
          jd (harness_integrate_coalescent_term (ccs[a], fcs[i], hcc, maxs[m], 0.0));
          fputc (']', jo);
        }
  }
  fprintf (jo, "],\n\"integrate_migration_term\":[");
  {
    static const int cms[] = { 0, 1, 2, 3, 7, 20, 64, 300, 1000 };
    static const double fms[] = { 0.0, 5e-7, 2e-6, 1e-3, 0.4, 6.4, 7.07, 55.0, 700.0, 5000.0, 40000.0 };
    static const double maxs[] = { 1e-6, 0.1, 1.0, 5.0 };
    for (first = 1, a = 0; a < 9; a++)
      for (i = 0; i < 11; i++)
        for (int m = 0; m < 4; m++)
        {
          if (cms[a] > 0 && fms[i] <= 0)
            continue;
          fprintf (jo, "%s[%d,", first ? "" : ",", cms[a]);
          first = 0;
          jd (fms[i]);
          fputc (',', jo);
          jd (maxs[m]);
          fputc (',', jo);
          jd (harness_integrate_migration_term (cms[a], fms[i], maxs[m], 0.0));
          fputc (',', jo);
          jd (harness_integrate_migration_term_expo_prior (cms[a], fms[i], maxs[m]));
          fputc (']', jo);
        }
  }
  fprintf (jo, "],\n\"calcmrate\":[");
  {
    static const int mcs[] = { 0, 1, 4 };
    static const double mts[] = { -1.0, 0.0, 0.3, 1.0, 2.5 };
    for (first = 1, a = 0; a < 3; a++)
      for (i = 0; i < 5; i++, first = 0)
      {
        fprintf (jo, "%s[%d,", first ? "" : ",", mcs[a]);
        jd (mts[i]);
        fputc (',', jo);
        jd (calcmrate (mcs[a], mts[i]));
        fputc (']', jo);
      }
  }
  fprintf (jo, "],\n\"swapweight_bw\":[");
  {
    static const double s[] = { -1234.5, -1230.25, -99.0, 10.5 };
    static const double b[] = { 1.0, 0.96, 0.5, 0.02 };
    for (first = 1, a = 0; a < 4; a++)
      for (i = 0; i < 4; i++)
        for (int c = 0; c < 4; c++)
          for (int d = 0; d < 4; d++, first = 0)
          {
            fprintf (jo, "%s[", first ? "" : ",");
            jd (s[a]);
            fputc (',', jo);
            jd (s[i]);
            fputc (',', jo);
            jd (b[c]);
            fputc (',', jo);
            jd (b[d]);
            fputc (',', jo);
            jd (harness_swapweight_bwprocesses (s[a], s[i], b[c], b[d]));
            fputc (']', jo);
          }
  }
  /* swapweight on the loaded chains (swapchains.cpp:12-34) */
  fprintf (jo, "],\n\"swapweight\":[");
  for (first = 1, a = 0; a < numchains; a++)
    for (i = 0; i < numchains; i++)
      if (a != i)
      {
        fprintf (jo, "%s[%d,%d,", first ? "" : ",", a, i);
        first = 0;
        jd (harness_swapweight (a, i));
        fputc (']', jo);
      }
  fprintf (jo, "],\"betas\":[");
  for (a = 0; a < numchains; a++)
  {
    if (a)
      fputc (',', jo);
    jd (beta[a]);
  }
  fprintf (jo, "],\"chainsum\":[");
  for (a = 0; a < numchains; a++)
  {
    double s = 0;
    for (i = 0; i < nloci; i++)
      s += C[a]->G[i].pdg;
    s += C[a]->allpcalc.probg;
    if (a)
      fputc (',', jo);
    jd (s);
  }
  fprintf (jo, "]}\n");
}

/* Collect G rows of the cold chain the way savegenealogyinfo() does (ima_main_mpi.cpp:3035-3211 ->
 * ginfo.cpp:318-377), directly into the reference's global gsampinf. */
static void
collect_rows (long rows, long every)
{
  long g, s;
  gsampinflength = calc_gsampinf_length ();
  gsampinf = static_cast<float **> (malloc (rows * sizeof (float *)));
  for (g = 0; g < rows; g++)
  {
    gsampinf[g] = static_cast<float *> (malloc (gsampinflength * sizeof (float)));
    for (s = 0; s < every; s++)
    {
      qupdate (0, 0, 1);
      step++;
    }
    savegsampinf (gsampinf[g], whichiscoldchain ());
  }
  genealogiessaved = (int) rows;
}

static void
mode_lmode (long burn, long rows, long every)
{
  long g;
  int p, i, first, np;
  do_burn (burn);
  collect_rows (rows, every);
  np = numpopsizeparams + nummigrateparams;
  if (kv.count ("ti"))
  {
    /* the rows as the reference writes them to a .ti file: savegenealogyfile (output.cpp:662-685) appends to a file
     * that the run created with a header ending in VALUESSTART (ima_main_mpi.cpp:2123-2141) */
    FILE *tf = fopen (kv["ti"].c_str (), "w");
    fprintf (tf, "header written by the harness\n\nVALUESSTART\n");
    fclose (tf);
    int last = -1;
    savegenealogyfile (const_cast<char *> (kv["ti"].c_str ()), NULL, &last, gsampinflength);
  }
  fprintf (jo, "{");
  dump_model ();
  fprintf (jo, "\"rows\":[");
  for (g = 0; g < rows; g++)
  {
    fprintf (jo, "%s[", g ? ",\n" : "");
    for (i = 0; i < gsampinflength; i++)
      fprintf (jo, "%s%.9g", i ? "," : "", (double) gsampinf[g][i]);
    fputc (']', jo);
  }
  fprintf (jo, "],\n\"margincalc\":[");
  for (first = 1, p = 0; p < np; p++)
  {
    double mx = p < numpopsizeparams ? itheta[p].pr.max : imig[p - numpopsizeparams].pr.max;
    for (i = 0; i < 40; i++, first = 0)
    {
      /* GRIDSIZE-style mid-bin points (histograms.cpp:81-99) thinned to 40, plus points near 0 and max */
      double x = mx * (i + 0.5) / 40.0;
      if (i == 0)
        x = mx * 0.5 / GRIDSIZE;
      if (i == 39)
        x = mx * (GRIDSIZE - 0.5) / GRIDSIZE;
      fprintf (jo, "%s[%d,", first ? "" : ",", p);
      jd (x);
      fputc (',', jo);
      jd (margincalc (x, 0.0, p, 0));
      fputc (',', jo);
      jd (margincalc (x, 0.25, p, 1));
      fputc (',', jo);
      jd (harness_marginp (p, 0, (int) rows, x));
      fputc (',', jo);
      jd (harness_marginp (p, (int) (rows / 3), (int) (2 * rows / 3), x));
      fputc (']', jo);
    }
  }
  fprintf (jo, "],\n\"jointp\":[");
  harness_jointp_setup ();
  {
    unsigned long long lcg = 88172645463325252ULL;
    for (i = 0; i < 48; i++)
    {
      double xv[32], ess = 0;
      for (p = 0; p < np; p++)
      {
        double mx = p < numpopsizeparams ? itheta[p].pr.max : imig[p - numpopsizeparams].pr.max;
        lcg ^= lcg << 13;
        lcg ^= lcg >> 7;
        lcg ^= lcg << 17;
        double u = (double) (lcg >> 11) / 9007199254740992.0;
        /* concentrate half of the points near the bulk of the posterior so that many terms are kept */
        xv[p] = (i & 1) ? mx * (0.02 + 0.3 * u) : mx * (1e-3 + 0.998 * u);
      }
      double q = jointp (xv, 1, &ess);
      fprintf (jo, "%s{", i ? ",\n" : "");
      jdarr ("x", xv, np, ",");
      fprintf (jo, "\"q\":");
      jd (q);
      fprintf (jo, ",\"ess\":");
      jd (ess);
      fputc ('}', jo);
    }
  }
  fprintf (jo, "]");
  if (npops > 2)
  {
    /* the two full models of a three-population joint search: the same kind of points under nowmodeltype 1 and 2 */
    for (int type = 1; type <= 2; type++)
    {
      unsigned long long lcg = 88172645463325252ULL + (unsigned long long) type;
      harness_jointp_set_type (type);
      fprintf (jo, ",\n\"jointp_type%d\":[", type);
      for (i = 0; i < 32; i++)
      {
        double xv[64], ess = 0;
        for (p = 0; p < np; p++)
        {
          double mx = p < numpopsizeparams ? itheta[p].pr.max : imig[p - numpopsizeparams].pr.max;
          lcg ^= lcg << 13;
          lcg ^= lcg >> 7;
          lcg ^= lcg << 17;
          double u = (double) (lcg >> 11) / 9007199254740992.0;
          xv[p] = (i & 1) ? mx * (0.02 + 0.3 * u) : mx * (1e-3 + 0.998 * u);
        }
        double q = jointp (xv, 1, &ess);
        fprintf (jo, "%s{", i ? ",\n" : "");
        jdarr ("x", xv, np, ",");
        fprintf (jo, "\"q\":");
        jd (q);
        fprintf (jo, ",\"ess\":");
        jd (ess);
        fputc ('}', jo);
      }
      fprintf (jo, "]");
    }
    harness_jointp_set_type (0);
  }
  if (kv.count ("extra"))
  {
    /* section 8 (f3): the other evaluators that stream over the rows.  calcx sums in row order exactly as
     * print_means_variances_correlations accumulates them (output.cpp:704-728), the table it prints, and the 2NM
     * densities calc_popmig / marginpopmig (popmig.cpp:9-97,176-268) or their exponential-prior forms (:101-170,:272-357) */
    int q, expo = modeloptions[EXPOMIGRATIONPRIOR];
    std::vector<double> m0 (np, 0.0), m1 (np, 0.0), cr (np * np, 0.0);
    for (g = 0; g < rows; g++)
      for (p = 0; p < np; p++)
      {
        m0[p] += harness_calcx ((int) g, p, 0);
        m1[p] += harness_calcx ((int) g, p, 1);
      }
    for (g = 0; g < rows; g++)
      for (p = 0; p < np - 1; p++)
        for (q = p + 1; q < np; q++)
          cr[p * np + q] += harness_calcx ((int) g, p, 0) * harness_calcx ((int) g, q, 0);
    fprintf (jo, ",\n\"calcx\":{");
    jdarr ("sum0", &m0[0], np, ",");
    jdarr ("sum1", &m1[0], np, ",");
    jdarr ("cross", &cr[0], np * np, ",");
    fprintf (jo, "\"sample\":[");
    for (g = 0; g < rows && g < 8; g++)
      for (p = 0; p < np; p++)
      {
        fprintf (jo, "%s[%ld,%d,", (g || p) ? "," : "", g, p);
        jd (harness_calcx ((int) g, p, 0));
        fputc (',', jo);
        jd (harness_calcx ((int) g, p, 1));
        fputc (']', jo);
      }
    fprintf (jo, "],\"table\":\"");
    {
      FILE *tf = tmpfile ();
      int ch;
      print_means_variances_correlations (tf);
      rewind (tf);
      while ((ch = fgetc (tf)) != EOF)
      {
        if (ch == '\n') fputs ("\\n", jo);
        else if (ch == '\t') fputs ("\\t", jo);
        else if (ch == '"' || ch == '\\') { fputc ('\\', jo); fputc (ch, jo); }
        else fputc (ch, jo);
      }
      fclose (tf);
    }
    fprintf (jo, "\"},\n\"popmig\":[");
    for (first = 1, p = 0; p < numpopsizeparams; p++)
      for (q = 0; q < nummigrateparams; q++)
      {
        double hi = expo ? EXPOMIGPLOTSCALE * imig[q].pr.mean : itheta[p].pr.max * imig[q].pr.max / 2.0;
        for (i = 0; i < 14; i++, first = 0)
        {
          double x = hi * (i + 0.5) / 14.0;
          if (i == 0)
            x = hi * 0.5 / GRIDSIZE;
          if (i == 13)
            x = hi * (GRIDSIZE - 0.5) / GRIDSIZE;
          fprintf (jo, "%s[%d,%d,", first ? "" : ",", p, q);
          jd (x);
          fputc (',', jo);
          jd (expo ? calc_pop_expomig (p, q, x, 0) : calc_popmig (p, q, x, 0));
          fputc (',', jo);
          jd (expo ? calc_pop_expomig (p, q, x, 1) : calc_popmig (p, q, x, 1));
          fputc (',', jo);
          jd (expo ? marginpop_expomig (q, 0, (int) rows, x, p) : marginpopmig (q, 0, (int) rows, x, p));
          fputc (',', jo);
          jd (expo ? marginpop_expomig (q, (int) (rows / 3), (int) (2 * rows / 3), x, p) : marginpopmig (q, (int) (rows / 3), (int) (2 * rows / 3), x, p));
          fputc (']', jo);
        }
      }
    fprintf (jo, "]");
    /* greater-than probabilities (gtint.cpp:128-330), for the pairs print_greater_than_tests computes (:353-373) */
    fprintf (jo, ",\n\"greater_than\":[");
    for (first = 1, p = 0; p < numpopsizeparams; p++)
      for (q = 0; q < numpopsizeparams; q++)
        if (p != q && itheta[p].pr.max == itheta[q].pr.max)
        {
          fprintf (jo, "%s[0,%d,%d,", first ? "" : ",", p, q);
          jd (harness_greater_than (0, p, q));
          fputc (']', jo);
          first = 0;
        }
    if (!expo)
      for (p = 0; p < nummigrateparams; p++)
        for (q = 0; q < nummigrateparams; q++)
          if (p != q && imig[p].pr.max > MINPARAMVAL && imig[q].pr.max > MINPARAMVAL && imig[p].pr.max == imig[q].pr.max)
          {
            fprintf (jo, "%s[1,%d,%d,", first ? "" : ",", p, q);
            jd (harness_greater_than (1, p, q));
            fputc (']', jo);
            first = 0;
          }
    fprintf (jo, "]");
  }
  fprintf (jo, "}\n");
}

/* `chunks` timed chunks of `iters` sweeps each; a sweep = updategenealogy() for every chain x locus
 * (full=0, the loop of qupdate ima_main_mpi.cpp:1821-1841) or one whole qupdate() step (full=1) */
static void
mode_bench (long burn, long iters, long full, long chunks, long gburn)
{
  long it, ch, acc = 0, tries = 0;
  int ci, li, a, b;
  std::vector<double> secs;
  do_burn (burn);
  /* gburn untimed sweeps of updategenealogy() only: the split times stay where start() put them
   * ((i+1)/(nsplit+1) of the prior maximum, initialize.cpp:1959), which is what the GPU arm also uses */
  for (it = 0; it < gburn; it++)
    for (ci = 0; ci < numchains; ci++)
      for (li = 0; li < nloci; li++)
        updategenealogy (ci, li, &a, &b);
  /* qupdate() counts the cold chain's genealogy updates itself (ima_main_mpi.cpp:1827,1835) */
  long cold0[2] = { 0, 0 }, cold1[2] = { 0, 0 };
  for (li = 0; li < nloci; li++)
  {
    cold0[0] += (long) L[li].g_rec->upinf[IM_UPDATE_GENEALOGY_ANY].tries;
    cold0[1] += (long) L[li].g_rec->upinf[IM_UPDATE_GENEALOGY_ANY].accp;
  }
  double t00 = nowsec ();
  for (ch = 0; ch < chunks; ch++)
  {
    double t0 = nowsec ();
    for (it = 0; it < iters; it++)
    {
      if (full)
      {
        qupdate (0, 0, 1);
        step++;
        tries += (long) numchains *nloci;
      }
      else
        for (ci = 0; ci < numchains; ci++)
          for (li = 0; li < nloci; li++)
          {
            acc += updategenealogy (ci, li, &a, &b);
            tries++;
          }
    }
    secs.push_back (nowsec () - t0);
  }
  double t1 = nowsec ();
  for (li = 0; li < nloci; li++)
  {
    cold1[0] += (long) L[li].g_rec->upinf[IM_UPDATE_GENEALOGY_ANY].tries;
    cold1[1] += (long) L[li].g_rec->upinf[IM_UPDATE_GENEALOGY_ANY].accp;
  }
  if (full)
  {

    acc = cold1[1] - cold0[1];
  }
  fprintf (jo, "{\"mode\":\"%s\",\"chains\":%d,\"loci\":%d,\"iters\":%ld,\"chunks\":%ld,\"updates\":%ld,\"updates_per_chunk\":%ld,"
           "\"accepted\":%ld,\"accept_base\":%ld,\"seconds\":%.6f,\"updates_per_sec\":%.3f,\"chunk_seconds\":[", full ? "qupdate" : "updategenealogy",
           numchains, nloci, iters, chunks, tries, (long) numchains * nloci * iters, acc, full ? cold1[0] - cold0[0] : tries, t1 - t00,
           tries / (t1 - t00));
  for (ch = 0; ch < chunks; ch++)
    fprintf (jo, "%s%.6f", ch ? "," : "", secs[ch]);
  fprintf (jo, "]}\n");
}

/* split-time updates: changet_RY1() (update_t_RY.cpp:222-517) with the accept draw forced, so that the proposed
 * state is visible afterwards; the Metropolis-Hastings term is what the reference passed to DMIN at :425 */
static void
mode_tupdates (long burn, long n, long between)
{
  do_burn (burn);
  recompute_all ();
  fprintf (jo, "{");
  dump_model ();
  fprintf (jo, "\"tprior_max\":[");
  for (int k = 0; k < numsplittimes; k++)
  {
    if (k) fputc (',', jo);
    jd (T[k].pr.max);
  }
  fprintf (jo, "],\"tprior_min\":[");
  for (int k = 0; k < numsplittimes; k++)
  {
    if (k) fputc (',', jo);
    jd (T[k].pr.min);
  }
  fprintf (jo, "],\n\"getnewt\":[");
  for (int i = 0; i < 40; i++)
  {
    /* getnewt (update_gtree_common.cpp:2501-2519): the same generator state gives U and then newt(U) */
    int period = i % numsplittimes;
    double tu = 0.05 * (i % 3), td = T[period].pr.max - 0.1 * (i % 4), oldt = tu + (td - tu) * (0.03 + 0.94 * ((i * 7) % 40) / 40.0);
    setseeds (1000 + i);
    double u = uniform ();
    setseeds (1000 + i);
    double nt = getnewt (period, tu, td, oldt, 1);
    fprintf (jo, "%s[%d,", i ? "," : "", period);
    jd (tu); fputc (',', jo); jd (td); fputc (',', jo); jd (oldt); fputc (',', jo); jd (u); fputc (',', jo); jd (nt);
    fputc (']', jo);
  }
  setseeds (77);
  fprintf (jo, "],\n\"records\":[");
  for (long it = 0; it < n; it++)
  {
    int ci = (int) (it % numchains), period = (int) ((it / numchains) % numsplittimes);
    for (long b = 0; b < between; b++)
    {
      qupdate (0, 0, 1);
      step++;
    }
    recompute_all ();
    fprintf (jo, "%s{\"ci\":%d,\"period\":%d,\"before\":", it ? ",\n" : "", ci, period);
    dump_chain (ci);
    g_force_accept = 1;
    int acc = changet_RY1 (ci, period);
    g_force_accept = 0;
    fprintf (jo, ",\"accepted\":%d,\"mh\":", acc);
    jd (harness_ry_last_mh ());
    fprintf (jo, ",\"after\":");
    dump_chain (ci);
    fputc ('}', jo);
  }
  fprintf (jo, "]}\n");
}

/* split-time updates, Nielsen-Wakeley: changet_NW() (update_t_NW.cpp:919-1032).  The call is made from a fixed seed; if it
 * rejects (the reference then restores everything) it is repeated from the same seed with its last uniform() -- the
 * accept draw -- replaced by 0, so that the proposed state stays visible */
static void
mode_nwupdates (long burn, long n, long between)
{
  do_burn (burn);
  recompute_all ();
  fprintf (jo, "{");
  dump_model ();
  fprintf (jo, "\"records\":[");
  int first = 1;
  for (long it = 0; it < n; it++)
  {
    int ci = (int) (it % numchains), period = (int) ((it / numchains) % numsplittimes);
    for (long b = 0; b < between; b++)
    {
      qupdate (0, 0, 1);
      step++;
    }
    recompute_all ();
    char *buf = NULL;
    size_t blen = 0;
    FILE *keep = jo;
    jo = open_memstream (&buf, &blen);
    dump_chain (ci);
    fclose (jo);
    jo = keep;
    setseeds (5000 + (int) it);
    g_nw_count = 0;
    g_nw_force = -1;
    int acc = changet_NW (ci, period), natural = acc;
    long ncalls = g_nw_count;
    if (!acc)
    {
      setseeds (5000 + (int) it);
      g_nw_count = 0;
      g_nw_force = ncalls;
      acc = changet_NW (ci, period);
      g_nw_force = -1;
    }
    if (acc)
    {
      fprintf (jo, "%s{\"ci\":%d,\"period\":%d,\"natural\":%d,\"mh\":", first ? "" : ",\n", ci, period, natural);
      jd (harness_nw_last_mh ());
      fprintf (jo, ",\"before\":%s,\"after\":", buf);
      dump_chain (ci);
      fputc ('}', jo);
      first = 0;
    }
    free (buf);
  }
  fprintf (jo, "]}\n");
}

/* mutation-scalar updates: changeu() (update_mc_params.cpp:23-370), unforced; every uniform() the call drew is
 * logged, the Metropolis-Hastings term is DMIN's second argument at :294 */
static void
mode_uupdates (long burn, long n, long between)
{
  do_burn (burn);
  recompute_all ();
  fprintf (jo, "{");
  dump_model ();
  fprintf (jo, "\"nurates\":%d,\"ul\":[", nurates);
  for (int i = 0; i < nurates; i++)
    fprintf (jo, "%s[%d,%d]", i ? "," : "", ul[i].l, ul[i].u);
  fprintf (jo, "],\"u_prmax\":");
  jd (L[0].u_rec[0].pr.max);
  fprintf (jo, ",\"u_win\":");
  jd (L[0].u_rec[0].win);
  fprintf (jo, ",\"kappa_win\":");
  jd (L[0].model == HKY ? L[0].kappa_rec->win : 0.0);
  fprintf (jo, ",\"kappa_max\":");
  jd (L[0].model == HKY ? L[0].kappa_rec->pr.max : 0.0);
  fprintf (jo, ",\n\"records\":[");
  for (long it = 0; it < n; it++)
  {
    int ci = (int) (it % numchains), j = (int) ((it / numchains) % (nurates - (nurates == 2))), k = -1;
    for (long b = 0; b < between; b++)
    {
      qupdate (0, 0, 1);
      step++;
    }
    recompute_all ();
    fprintf (jo, "%s{\"ci\":%d,\"j\":%d,\"before\":", it ? ",\n" : "", ci, j);
    dump_chain (ci);
    g_ulog.clear ();
    int acc = changeu (ci, j, &k);
    fprintf (jo, ",\"k\":%d,\"accepted\":%d,\"mh\":", k, acc);
    jd (harness_mc_last_mh ());
    fputc (',', jo);
    jdarr ("U", g_ulog.data (), (int) g_ulog.size (), ",");
    fprintf (jo, "\"after\":");
    dump_chain (ci);
    fputc ('}', jo);
  }
  fprintf (jo, "]}\n");
}

/* thermomarginlikecalc (marglike.cpp:121-150) on synthetic per-temperature sums */
static void
mode_thermo (void)
{
  static const int ns[] = { 3, 4, 5, 8, 9, 16, 33, 128 };
  const int keepn = numchains;
  fprintf (jo, "{\"thermo\":[");
  for (size_t t = 0; t < sizeof (ns) / sizeof (ns[0]); t++)
  {
    const int n = ns[t], k = 10 + 7 * (int) t;
    numchains = n;
    for (int i = 0; i < n; i++)
      thermosum[i] = -k * (300.0 + 40.0 * sin (0.7 * i + t) + 2.5 * i);
    fprintf (jo, "%s{\"k\":%d,", t ? ",\n" : "", k);
    jdarr ("sums", thermosum, n, ",");
    fprintf (jo, "\"value\":");
    jd (thermomarginlikecalc (k));
    fputc ('}', jo);
  }
  numchains = keepn;
  fprintf (jo, "]}\n");
}

/* long-run summary statistics of the reference sampler with split times and mutation scalars held at their start
 * values: `sweeps` sweeps of updategenealogy() over every chain x locus after `gburn` untimed ones; per locus the
 * mean (over sweeps and chains) of tree length, root time, migration count and per-population coalescence counts,
 * with batch-means standard errors (nbatch batches).  Used for the statistical parity fixture. */
/* diagnostic schedules built from the reference's own update functions: full=2 the engine's schedule (RY1 every
 * step, changeu every 5th), full=3 NW only, full=4 RY1 only without changeu */
/* tries / accepts of the split-time updates per period: [period][RY tries, RY accepts, NW tries, NW accepts], all chains */
static std::vector<double> g_trate;

static void
ry_only_rest (long it, long full)
{
  int k;
  if (g_trate.empty ())
    g_trate.assign ((size_t) 4 * (numsplittimes > 0 ? numsplittimes : 1), 0.0);
  for (int ci = 0; ci < numchains; ci++)
  {
    int period = randposint (numsplittimes);
    if (full == 3)
    {
      g_trate[4 * period + 2] += 1;
      g_trate[4 * period + 3] += changet_NW (ci, period) ? 1 : 0;
    }
    else
    {
      g_trate[4 * period + 0] += 1;
      g_trate[4 * period + 1] += changet_RY1 (ci, period) ? 1 : 0;
    }
  }
  if ((it + 1) % 5 == 0 && nurates > 1 && full != 4)
    for (int ci = 0; ci < numchains; ci++)
      for (int j = 0; j < (nurates - (nurates == 2)); j++)
        changeu (ci, j, &k);
}

static void
mode_trace (long gburn, long sweeps, long nbatch, long full)
{
  int ci, li, a, b, k;
  long it, bi;
  /* full=1: whole qupdate() steps (genealogies, split times by RY1 or NW, mutation scalars); the split times and
   * the log scalars are then summarised too */
  std::vector<double> t0v, u0v;
  /* the genealogies the run starts from (they belong to the split times in "tvals") */
  fprintf (jo, "{\"start\":");
  dump_tree_all ();
  fputc (',', jo);
  for (k = 0; k < numsplittimes; k++)
    t0v.push_back (C[0]->tvals[k]);
  for (li = 0; li < nloci; li++)
    u0v.push_back (C[0]->G[li].uvals[0]);
  for (it = 0; it < gburn; it++)
  {
    if (full == 1)
    {
      qupdate (0, 0, 1);
      step++;
      continue;
    }
    for (ci = 0; ci < numchains; ci++)
      for (li = 0; li < nloci; li++)
        updategenealogy (ci, li, &a, &b);
    if (full >= 2)
      ry_only_rest (it, full);
  }
  /* split-time update counts start after the burn-in: qupdate() counts the cold chain in T[].upinf, the diagnostic
   * schedules count every chain in g_trate */
  std::vector<double> trate0 ((size_t) 4 * (numsplittimes > 0 ? numsplittimes : 1), 0.0);
  for (k = 0; k < numsplittimes; k++)
    if (full == 1)
    {
      trate0[4 * k + 0] = T[k].upinf[IM_UPDATE_TIME_RY1].tries; trate0[4 * k + 1] = T[k].upinf[IM_UPDATE_TIME_RY1].accp;
      trate0[4 * k + 2] = T[k].upinf[IM_UPDATE_TIME_NW].tries; trate0[4 * k + 3] = T[k].upinf[IM_UPDATE_TIME_NW].accp;
    }
    else if (!g_trate.empty ())
      for (a = 0; a < 4; a++)
        trate0[4 * k + a] = g_trate[4 * k + a];
  const int NS = 6;             /* length, roottime, mignum, cc0, cc1, cc2(+) */
  std::vector<double> bsum ((size_t) nbatch * nloci * NS, 0.0);
  std::vector<double> tsum ((size_t) nbatch * (numsplittimes + 1), 0.0), usum ((size_t) nbatch * nloci, 0.0);
  long per = sweeps / nbatch, acc = 0, tries = 0;
  for (bi = 0; bi < nbatch; bi++)
    for (it = 0; it < per; it++)
    {
      if (full == 1)
      {
        qupdate (0, 0, 1);
        step++;
      }
      if (full >= 2)
      {
        for (ci = 0; ci < numchains; ci++)
          for (li = 0; li < nloci; li++)
            updategenealogy (ci, li, &a, &b);
        ry_only_rest (bi * per + it, full);
      }
      for (ci = 0; ci < numchains; ci++)
      {
        for (k = 0; k < numsplittimes; k++)
          tsum[(size_t) bi * (numsplittimes + 1) + k] += C[ci]->tvals[k];
        for (li = 0; li < nloci; li++)
        {
          if (!full)
          {
            acc += updategenealogy (ci, li, &a, &b);
            tries++;
          }
          usum[(size_t) bi * nloci + li] += log (C[ci]->G[li].uvals[0]);
          struct genealogy *G = &C[ci]->G[li];
          double *o = &bsum[((size_t) bi * nloci + li) * NS];
          o[0] += G->length;
          o[1] += G->roottime;
          o[2] += G->mignum;
          o[3] += G->gweight.cc[0][0];
          o[4] += npops > 1 ? G->gweight.cc[0][1] : 0;
          for (k = 1; k <= numsplittimes; k++)
            o[5] += G->gweight.cc[k][0];
        }
      }
    }
  dump_model ();
  fprintf (jo, "\"sweeps\":%ld,\"chains\":%d,\"nbatch\":%ld,\"full\":%ld,\"accept\":%.6f,\"tvals\":[", per * nbatch, numchains, nbatch, full,
           tries ? (double) acc / tries : -1.0);
  for (k = 0; k < numsplittimes; k++)
  {
    if (k)
      fputc (',', jo);
    jd (t0v[k]);
  }
  fprintf (jo, "],\"uvals\":[");
  for (li = 0; li < nloci; li++)
  {
    if (li)
      fputc (',', jo);
    jd (u0v[li]);
  }
  fprintf (jo, "],\"tprior_max\":[");
  for (k = 0; k < numsplittimes; k++)
  {
    if (k)
      fputc (',', jo);
    jd (T[k].pr.max);
  }
  fprintf (jo, "],\"t_rates\":[");
  for (k = 0; k < numsplittimes; k++)
  {
    double now[4] = { 0, 0, 0, 0 };
    if (full == 1)
    {
      now[0] = T[k].upinf[IM_UPDATE_TIME_RY1].tries; now[1] = T[k].upinf[IM_UPDATE_TIME_RY1].accp;
      now[2] = T[k].upinf[IM_UPDATE_TIME_NW].tries; now[3] = T[k].upinf[IM_UPDATE_TIME_NW].accp;
    }
    else if (!g_trate.empty ())
      for (a = 0; a < 4; a++)
        now[a] = g_trate[4 * k + a];
    fprintf (jo, "%s[%.0f,%.0f,%.0f,%.0f]", k ? "," : "", now[0] - trate0[4 * k], now[1] - trate0[4 * k + 1], now[2] - trate0[4 * k + 2],
             now[3] - trate0[4 * k + 3]);
  }
  fprintf (jo, "],\"t_batch_means\":[");
  for (bi = 0; bi < nbatch; bi++)
  {
    fprintf (jo, "%s[", bi ? "," : "");
    for (k = 0; k < numsplittimes; k++)
    {
      if (k)
        fputc (',', jo);
      jd (tsum[(size_t) bi * (numsplittimes + 1) + k] / ((double) per * numchains));
    }
    fputc (']', jo);
  }
  fprintf (jo, "],\"logu_batch_means\":[");
  for (bi = 0; bi < nbatch; bi++)
  {
    fprintf (jo, "%s[", bi ? "," : "");
    for (li = 0; li < nloci; li++)
    {
      if (li)
        fputc (',', jo);
      jd (usum[(size_t) bi * nloci + li] / ((double) per * numchains));
    }
    fputc (']', jo);
  }
  fprintf (jo, "],\"batch_means\":[");
  for (bi = 0; bi < nbatch; bi++)
  {
    fprintf (jo, "%s[", bi ? "," : "");
    for (li = 0; li < nloci; li++)
    {
      fprintf (jo, "%s[", li ? "," : "");
      for (k = 0; k < NS; k++)
      {
        if (k)
          fputc (',', jo);
        jd (bsum[((size_t) bi * nloci + li) * NS + k] / ((double) per * numchains));
      }
      fputc (']', jo);
    }
    fputc (']', jo);
  }
  fprintf (jo, "]}\n");
}

static void
mode_lbench (long burn, long rows, long evals)
{
  long g, e;
  int i, np;
  /* a few hundred real rows, then bootstrap them up to `rows` */
  long base = rows < 256 ? rows : 256;
  do_burn (burn);
  collect_rows (base, 2);
  gsampinf = static_cast<float **> (realloc (gsampinf, rows * sizeof (float *)));
  for (g = base; g < rows; g++)
    gsampinf[g] = gsampinf[(g * 2654435761UL) % base];
  genealogiessaved = (int) rows;
  np = numpopsizeparams + nummigrateparams;
  double acc = 0, t0 = nowsec ();
  for (e = 0; e < evals; e++)
  {
    i = (int) (e % np);
    double mx = i < numpopsizeparams ? itheta[i].pr.max : imig[i - numpopsizeparams].pr.max;
    acc += margincalc (mx * (0.5 + (e % 997)) / 1000.0, 0.0, i, 0);
  }
  double t1 = nowsec ();
  harness_jointp_setup ();
  long jevals = evals / 8 > 0 ? evals / 8 : 1;
  double t2 = nowsec ();
  for (e = 0; e < jevals; e++)
  {
    double xv[32], ess;
    for (i = 0; i < np; i++)
    {
      double mx = i < numpopsizeparams ? itheta[i].pr.max : imig[i - numpopsizeparams].pr.max;
      xv[i] = mx * (0.05 + 0.0007 * ((e * 31 + i * 17) % 997));
    }
    acc += jointp (xv, 0, &ess);
  }
  double t3 = nowsec ();
  fprintf (jo, "{\"rows\":%ld,\"margincalc_evals\":%ld,\"margincalc_seconds\":%.6f,\"margincalc_geneval_per_sec\":%.3f,"
           "\"jointp_evals\":%ld,\"jointp_seconds\":%.6f,\"jointp_geneval_per_sec\":%.3f,\"checksum\":%.17g}\n", rows, evals, t1 - t0,
           (double) evals * rows / (t1 - t0), jevals, t3 - t2, (double) jevals * rows / (t3 - t2), acc);
}

int
main (int argc, char *argv[])
{
  int i, split = -1;
  if (argc < 4)
  {
    fprintf (stderr, "usage: ref_harness MODE OUT [key=value ...] -- <IMa2p args>\n");
    return 2;
  }
  for (i = 3; i < argc; i++)
  {
    if (!strcmp (argv[i], "--"))
    {
      split = i;
      break;
    }
    char *eq = strchr (argv[i], '=');
    if (eq)
      kv[std::string (argv[i], eq - argv[i])] = std::string (eq + 1);
  }
  if (split < 0)
  {
    fprintf (stderr, "missing -- before the IMa2p command line\n");
    return 2;
  }
  std::string mode = argv[1];
  jo = fopen (argv[2], "w");
  if (!jo)
  {
    perror (argv[2]);
    return 2;
  }
  std::vector<char *>av;
  av.push_back (argv[0]);
  for (i = split + 1; i < argc; i++)
    av.push_back (argv[i]);
  if (mode == "stock")           /* the reference's own main(), end to end (M or L mode); OUT receives its exit status */
  {
    g_stock_seed = kv.count ("seed") ? atol (kv["seed"].c_str ()) : -1;
    int rc = ima2p_reference_main ((int) av.size (), av.data ());
    fprintf (jo, "{\"exit\":%d}\n", rc);
    fclose (jo);
    return rc;
  }
  numprocesses = 1;
  init_IMA ();
  start ((int) av.size (), av.data (), 0);
  if (kvl ("seed", -1) >= 0)
    setseeds ((int) kvl ("seed", 0));     /* -s is ineffective in the serial build */
  init_after_start_IMA ();
  step = 0;
  recordstep = 0;
  long burn = kvl ("burn", 0);
  if (mode == "state")
  {
    do_burn (burn);
    recompute_all ();
    dump_state ();
  }
  else if (mode == "updates")
    mode_updates (burn, kvl ("n", 100));
  else if (mode == "kat")
  {
    do_burn (burn);
    recompute_all ();
    mode_kat ();
  }
  else if (mode == "lmode")
    mode_lmode (burn, kvl ("rows", 500), kvl ("every", 5));
  else if (mode == "bench")
    mode_bench (burn, kvl ("iters", 10), kvl ("full", 0), kvl ("chunks", 1), kvl ("gburn", 0));
  else if (mode == "trace")
  {
    capture_start_trees ();
    mode_trace (kvl ("gburn", 1000), kvl ("sweeps", 20000), kvl ("nbatch", 20), kvl ("full", 0));
  }
  else if (mode == "tupdates")
    mode_tupdates (burn, kvl ("n", 40), kvl ("between", 3));
  else if (mode == "mcf")
  {
    /* the reference's own state file: written, read back (readmcf ends in init_p), then dumped -- both sides of the
     * parity test parse the same text */
    do_burn (burn);
    if (kv.count ("load") == 0)
      writemcf (const_cast<char *> (kv["mcf"].c_str ()));
    readmcf (const_cast<char *> (kv["mcf"].c_str ()));
    dump_state ();
  }
  else if (mode == "nwupdates")
    mode_nwupdates (burn, kvl ("n", 40), kvl ("between", 3));
  else if (mode == "uupdates")
    mode_uupdates (burn, kvl ("n", 40), kvl ("between", 3));
  else if (mode == "thermo")
    mode_thermo ();
  else if (mode == "lbench")
    mode_lbench (burn, kvl ("rows", 100000), kvl ("evals", 50));
  else
  {
    fprintf (stderr, "unknown mode %s\n", mode.c_str ());
    return 2;
  }
  fclose (jo);
  return 0;
}
