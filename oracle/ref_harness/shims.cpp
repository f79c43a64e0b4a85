/* TEST INFRASTRUCTURE -- reference-side harness shim (never shipped, never linked into the product).
 *
 * Each shim translation unit #includes exactly ONE reference source file *where it lies* under
 * /root/reference/src (nothing is copied into this repo) and appends exported wrappers for the
 * file-static functions / variables the parity fixtures need.  Which file is included is selected
 * with -DSHIM_<NAME> by oracle/build_ref.sh; the unmodified reference file of the same name is then
 * left out of the link.
 */
#if defined(SHIM_GTREE_COMMON)
#include "update_gtree_common.cpp"
/* update_gtree_common.cpp:108,205,298 are file-static */
double harness_integrate_coalescent_term (int cc, double fc, double hcc, double max, double min)
{ return integrate_coalescent_term (cc, fc, hcc, max, min); }
double harness_integrate_migration_term (int cm, double fm, double max, double min)
{ return integrate_migration_term (cm, fm, max, min); }
double harness_integrate_migration_term_expo_prior (int cm, double fm, double exmean)
{ return integrate_migration_term_expo_prior (cm, fm, exmean); }

#elif defined(SHIM_SWAPCHAINS)
#include "swapchains.cpp"
/* swapchains.cpp:12,57 are file-static */
double harness_swapweight (int ci, int cj) { return swapweight (ci, cj); }
double harness_swapweight_bwprocesses (double sumi, double sumj, double betai, double betaj)
{ return swapweight_bwprocesses (sumi, sumj, betai, betaj); }
double harness_calcpartialswapweight (int c) { return calcpartialswapweight (c); }

#elif defined(SHIM_SURFACE)
#include "surface_call_functions.cpp"
/* surface_call_functions.cpp:25 is file-static */
double harness_marginp (int param, int firsttree, int lasttree, double x)
{ return marginp (param, firsttree, lasttree, x, 0); }

#elif defined(SHIM_JOINTFIND)
#include "jointfind.cpp"
/* jointfind.cpp:159,183,184 are file-static; findjointpeaks() :1074-1115 does this set-up before
 * the first jointp() call, ima_main_mpi.cpp allocates eexpsum in L mode. */
void harness_jointp_setup (void)
{
  nparams = numpopsizeparams + nummigrateparams;
  setuplist ();
  nparamrange[0] = 0;
  nparamrange[1] = nparams;
  nowmodeltype = 0;
  eexpsum = (struct extendnum *) malloc ((genealogiessaved + 1) * sizeof (struct extendnum));
}
/* the two "full" models of a three-population analysis (findjointpeaks :1118-1133): 1 = all population sizes,
 * 2 = all migration rates; 0 = the two-population full model */
void harness_jointp_set_type (int type)
{
  nowmodeltype = type;
  nparamrange[0] = type == 2 ? numpopsizeparams : 0;
  nparamrange[1] = type == 1 ? numpopsizeparams : numpopsizeparams + nummigrateparams;
}

#elif defined(SHIM_CALC_PROB_DATA)
#include "calc_prob_data.cpp"
/* calc_prob_data.cpp:8 is file-static */
double harness_get_sumlogk (int li) { return sumlogk[li] ? *sumlogk[li] : 0.0; }

#elif defined(SHIM_MCMCFILE)
#include "mcmcfile.cpp"
/* mcmcfile.cpp:130 is file-static */
void harness_init_p (void) { init_p (); }
#elif defined(SHIM_T_RY)
/* every uniform() call of this file (the accept draw, update_t_RY.cpp:424) goes through a hook defined in
 * harness_main.cpp, which can force acceptance; DMIN's second argument at :425 is the Metropolis-Hastings term */
#define uniform harness_uniform_ry
#include "update_t_RY.cpp"
#undef uniform
double harness_ry_last_mh (void) { return dminarg2; }

#elif defined(SHIM_MC_PARAMS)
/* uniform() calls of this file are logged (pick of k, the ratio draw, kappa draws, the accept draw) */
#define uniform harness_uniform_mc
#include "update_mc_params.cpp"
#undef uniform
double harness_mc_last_mh (void) { return dminarg2; }

#elif defined(SHIM_T_NW)
/* uniform() calls of this file: the MIGSIMFRAC draws, the migration times, and last the accept draw (:1007);
 * the hook counts them and can replace the n-th by 0 */
#define uniform harness_uniform_nw
#include "update_t_NW.cpp"
#undef uniform
double harness_nw_last_mh (void) { return dminarg2; }

#elif defined(SHIM_OUTPUT)
#include "output.cpp"
/* output.cpp:14 is file-static */
double harness_calcx (int ei, int pnum, int mode) { return calcx (ei, pnum, mode); }

#elif defined(SHIM_GTINT)
#define USETREESMAX_H 20000       /* USETREESMAX gtint.cpp:24 (undefined again at the end of that file) */
#include "gtint.cpp"
/* gtint.cpp:19-20 (gtmig, gtpops) are file-static, and so is the row-thinning state print_greater_than_tests sets
 * (:341-351): the wrapper sets it the same way before calling them */
double harness_greater_than (int kind, int i, int j)
{
  if (genealogiessaved > USETREESMAX_H)
  {
    treeinc = (int) genealogiessaved / (int) USETREESMAX_H;
    numtreesused = USETREESMAX_H;
    hitreenum = treeinc * USETREESMAX_H;
  }
  else
  {
    treeinc = 1;
    hitreenum = numtreesused = genealogiessaved;
  }
  return kind == 0 ? gtpops (i, j) : gtmig (i, j);
}

#else
#error "select a shim"
#endif
