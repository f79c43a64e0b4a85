/* TEST INFRASTRUCTURE -- CPU oracle for the IMa2p hot path (see ima_oracle.h).
 *
 * Plain-C restatement of the reference algorithms.  Every function names the reference file:line it
 * follows (paths relative to the reference's src/).  Pinned against fixtures generated from the
 * unmodified reference: tests/test_oracle_golden.py.  Not part of the product.
 */
#include "ima_oracle.h"
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define LOG2 0.69314718055994530941723212146    /* imamp.hpp:189 */
#define LOG10 2.3025850929940456840     /* imamp.hpp:188 */
#define LOG_DBL_MAX 7.0978271289338397e+02      /* imamp.hpp:191 */
#define MYDBL_MAX (DBL_MAX/1e10)        /* imamp.hpp:170 */
#define IM_BESSI_MIN (-1e+100)  /* imamp.hpp:172 */
#define MPRIORMIN 0.000001      /* imamp.hpp:137 */
#define ABSMIGMAX 5000          /* imamp.hpp:134 */
#define LOGFACT_N (100 * ABSMIGMAX + 1) /* imamp.hpp:1107 */
#define MIGCLOSEFRAC 0.9        /* update_gtree_common.cpp:23 */
#define INTEGERROUND(x) ((x)>=0?(long)((x)+0.5):(long)((x)-0.5))        /* imamp.hpp:180 */

enum
{ ORA_ERR_LOGDIFF = 1, ORA_ERR_GAMMA = 2, ORA_ERR_IS = 3, ORA_ERR_TREE = 4, ORA_ERR_ARG = 5 };

static __thread int g_err = 0;
int
ora_last_error (void)
{
  return g_err;
}

void
ora_clear_error (void)
{
  g_err = 0;
}

struct ora_model
{
  int npops, nsplit, ntreepops, rootpop;
  int plist[ORA_MAXPOPS][ORA_MAXPOPS];
  int addpop[ORA_MAXPERIODS], droppops[ORA_MAXPERIODS][2];
  int pt_e[ORA_MAXTREEPOPS], pt_down[ORA_MAXTREEPOPS];
  int cc_off[ORA_MAXPERIODS + 1], mc_off[ORA_MAXPERIODS + 1], ncc, nmc;
  int nq, nm;
  int q_n[ORA_MAXPARAMS], q_idx[ORA_MAXPARAMS][ORA_MAXWP];
  int m_n[ORA_MAXPARAMS], m_idx[ORA_MAXPARAMS][ORA_MAXWP];
  double q_max[ORA_MAXPARAMS], q_min[ORA_MAXPARAMS];
  double m_max[ORA_MAXPARAMS], m_min[ORA_MAXPARAMS], m_mean[ORA_MAXPARAMS];
  int nomig_n, nomig_idx[ORA_MAXPARAMS];
  int nomigration, expoprior, thermo;
  double gbeta;
};

static int
ccidx (const ora_model * m, int k, int i)
{
  return m->cc_off[k] + i;
}

static int
mcidx (const ora_model * m, int k, int i, int j)
{
  return m->mc_off[k] + i * (m->npops - k) + j;
}

ora_model *
ora_model_create (int npops, int nsplit, const int *plist, const int *addpop, const int *droppops, const int *pt_e,
                  const int *pt_down, int rootpop, int nq, const int *q_off, const int *q_p, const int *q_r,
                  const double *q_max, const double *q_min, int nm, const int *m_off, const int *m_p, const int *m_r,
                  const int *m_c, const double *m_max, const double *m_min, const double *m_mean, int nomig_n,
                  const int *nomig_p, const int *nomig_r, const int *nomig_c, int nomigration, int expoprior, int thermo,
                  double gbeta)
{
  int k, i, j;
  if (npops < 1 || npops > ORA_MAXPOPS || nq > ORA_MAXPARAMS || nm > ORA_MAXPARAMS || nomig_n > ORA_MAXPARAMS)
    return NULL;
  ora_model *m = (ora_model *) calloc (1, sizeof (ora_model));
  m->npops = npops;
  m->nsplit = nsplit;
  m->ntreepops = 2 * npops - 1;
  m->rootpop = rootpop;
  for (k = 0; k < npops; k++)
    for (i = 0; i < npops; i++)
      m->plist[k][i] = plist[k * npops + i];
  for (k = 0; k <= nsplit; k++)
  {
    m->addpop[k] = addpop[k];
    m->droppops[k][0] = droppops[2 * k];
    m->droppops[k][1] = droppops[2 * k + 1];
  }
  for (i = 0; i < m->ntreepops; i++)
  {
    m->pt_e[i] = pt_e[i];
    m->pt_down[i] = pt_down[i];
  }
  m->cc_off[0] = m->mc_off[0] = 0;
  for (k = 0; k <= nsplit; k++)
  {
    m->cc_off[k + 1] = m->cc_off[k] + (npops - k);
    m->mc_off[k + 1] = m->mc_off[k] + (npops - k) * (npops - k);
  }
  m->ncc = m->cc_off[nsplit + 1];
  m->nmc = m->mc_off[nsplit];
  m->nq = nq;
  m->nm = nm;
  for (i = 0; i < nq; i++)
  {
    m->q_n[i] = q_off[i + 1] - q_off[i];
    for (j = 0; j < m->q_n[i]; j++)
      m->q_idx[i][j] = ccidx (m, q_p[q_off[i] + j], q_r[q_off[i] + j]);
    m->q_max[i] = q_max[i];
    m->q_min[i] = q_min[i];
  }
  for (i = 0; i < nm; i++)
  {
    m->m_n[i] = m_off[i + 1] - m_off[i];
    for (j = 0; j < m->m_n[i]; j++)
      m->m_idx[i][j] = mcidx (m, m_p[m_off[i] + j], m_r[m_off[i] + j], m_c[m_off[i] + j]);
    m->m_max[i] = m_max[i];
    m->m_min[i] = m_min[i];
    m->m_mean[i] = m_mean[i];
  }
  m->nomig_n = nomig_n;
  for (i = 0; i < nomig_n; i++)
    m->nomig_idx[i] = mcidx (m, nomig_p[i], nomig_r[i], nomig_c[i]);
  m->nomigration = nomigration;
  m->expoprior = expoprior;
  m->thermo = thermo;
  m->gbeta = gbeta;
  return m;
}

void
ora_model_destroy (ora_model * m)
{
  free (m);
}

int
ora_ncc (const ora_model * m)
{
  return m->ncc;
}

int
ora_nmc (const ora_model * m)
{
  return m->nmc;
}

/* ------------------------------------------------------------------------------------------- */
/* numerics: utilities.cpp                                                                      */
/* ------------------------------------------------------------------------------------------- */

static double *g_logfact = NULL;

/* setlogfact utilities.cpp:1405-1414 (same running sum, so the table is bit-identical) */
static void
ensure_logfact (void)
{
  int i;
  if (g_logfact)
    return;
  double *t = (double *) malloc (LOGFACT_N * sizeof (double));
  t[0] = 0;
  for (i = 1; i < LOGFACT_N; i++)
    t[i] = t[i - 1] + log ((double) i);
  g_logfact = t;
}

double
ora_logfact (int n)
{
  ensure_logfact ();
  if (n < 0 || n >= LOGFACT_N)
  {
    g_err = ORA_ERR_ARG;
    return NAN;
  }
  return g_logfact[n];
}

#define ITMAX 1000              /* utilities.cpp:808 */
#define EPS 3.0e-7              /* utilities.cpp:809 */
#define FPMIN 1.0e-30           /* utilities.cpp:810 */

/* gcf utilities.cpp:811-841 */
static double
gcf (double a, double x, double *gln)
{
  int i;
  double an, b, c, d, del, h;
  *gln = g_logfact[(int) a - 1];
  b = x + 1.0 - a;
  c = 1.0 / FPMIN;
  d = 1.0 / b;
  h = d;
  for (i = 1; i <= ITMAX; i++)
  {
    an = -i * (i - a);
    b += 2.0;
    d = an * d + b;
    if (fabs (d) < FPMIN)
      d = FPMIN;
    c = b + an / c;
    if (fabs (c) < FPMIN)
      c = FPMIN;
    d = 1.0 / d;
    del = d * c;
    h *= del;
    if (fabs (del - 1.0) < EPS)
      break;
  }
  if (i > ITMAX)
    g_err = ORA_ERR_GAMMA;
  return exp (-x + a * log (x) - (*gln)) * h;
}

/* gcflog utilities.cpp:846-877 */
static double
gcflog (double a, double x, double *gln)
{
  int i;
  double an, b, c, d, del, h;
  *gln = g_logfact[(int) a - 1];
  b = x + 1.0 - a;
  c = 1.0 / FPMIN;
  d = 1.0 / b;
  h = d;
  for (i = 1; i <= ITMAX; i++)
  {
    an = -i * (i - a);
    b += 2.0;
    d = an * d + b;
    if (fabs (d) < FPMIN)
      d = FPMIN;
    c = b + an / c;
    if (fabs (c) < FPMIN)
      c = FPMIN;
    d = 1.0 / d;
    del = d * c;
    h *= del;
    if (fabs (del - 1.0) < EPS)
      break;
  }
  if (i > ITMAX)
    g_err = ORA_ERR_GAMMA;
  return (-x + a * log (x) - (*gln)) + log (h);
}

/* gser utilities.cpp:882-912 */
static double
gser (int a, double x, double *gln)
{
  int n;
  double sum, del, ap;
  *gln = g_logfact[a - 1];
  if (x <= 0.0)
  {
    if (x < 0.0)
      g_err = ORA_ERR_GAMMA;
    return 0.0;
  }
  ap = a;
  del = sum = 1.0 / a;
  for (n = 1; n <= ITMAX; n++)
  {
    ++ap;
    del *= x / ap;
    sum += del;
    if (fabs (del) < fabs (sum) * EPS)
      return sum * exp (-x + a * log (x) - (*gln));
  }
  g_err = ORA_ERR_GAMMA;
  return 0.0;
}

/* gserlog utilities.cpp:917-953 */
static double
gserlog (int a, double x, double *gln)
{
  int n;
  double sum, del, ap;
  *gln = g_logfact[a - 1];
  if (x <= 0.0)
  {
    if (x < 0.0)
      g_err = ORA_ERR_GAMMA;
    return 0.0;
  }
  ap = a;
  del = sum = 1.0 / a;
  for (n = 1; n <= ITMAX; n++)
  {
    ++ap;
    del *= x / ap;
    sum += del;
    if (fabs (del) < fabs (sum) * EPS)
      return log (sum) + (-x + a * log (x) - (*gln));
  }
  g_err = ORA_ERR_GAMMA;
  return 0.0;
}

#define MAXIT 100               /* utilities.cpp:958 */
#define EULER 0.5772156649      /* utilities.cpp:959 */
/* expint utilities.cpp:962-1045 */
static double
expint (int n, double x, int *islog)
{
  int i, ii, nm1;
  double a, b, c, d, del, fact, h, psi, ans = 0;
  *islog = 0;
  nm1 = n - 1;
  if (n < 0 || x < 0.0 || (x == 0.0 && (n == 0 || n == 1)))
  {
    g_err = ORA_ERR_GAMMA;
    return NAN;
  }
  if (n == 0)
    return exp (-x) / x;
  if (x == 0.0)
    return 1.0 / nm1;
  if (x > 1.0)
  {
    b = x + n;
    c = 1.0 / FPMIN;
    d = 1.0 / b;
    h = d;
    for (i = 1; i <= MAXIT; i++)
    {
      a = -i * (nm1 + i);
      b += 2.0;
      d = 1.0 / (a * d + b);
      c = b + a / c;
      del = c * d;
      h *= del;
      if (fabs (del - 1.0) < EPS)
      {
        *islog = 1;
        return log (h) - x;
      }
    }
    g_err = ORA_ERR_GAMMA;
    return NAN;
  }
  ans = (nm1 != 0 ? 1.0 / nm1 : -log (x) - EULER);
  fact = 1.0;
  for (i = 1; i <= MAXIT; i++)
  {
    fact *= -x / i;
    if (i != nm1)
      del = -fact / (i - nm1);
    else
    {
      psi = -EULER;
      for (ii = 1; ii <= nm1; ii++)
        psi += 1.0 / ii;
      del = fact * (-log (x) + psi);
    }
    ans += del;
    if (fabs (del) < fabs (ans) * EPS)
      return ans;
  }
  g_err = ORA_ERR_GAMMA;
  return NAN;
}

/* uppergamma utilities.cpp:1053-1090 */
double
ora_uppergamma (int a, double x)
{
  int logindicator;
  double gamser, gammcf, gln, p, temp;
  ensure_logfact ();
  if (x < 0.0 || a < 0.0)
  {
    g_err = ORA_ERR_GAMMA;
    return NAN;
  }
  if (a == 0)
  {
    temp = expint (1, x, &logindicator);
    p = logindicator ? temp : log (temp);
  }
  else if (x < (a + 1.0))
  {
    gamser = gser (a, x, &gln);
    p = gln + log (1.0 - gamser);
  }
  else
  {
    gammcf = gcflog ((double) a, x, &gln);
    p = gln + gammcf;
  }
  if (p < -1e200)
    p = -1e200;
  return p;
}

/* lowergamma utilities.cpp:1092-1122 */
double
ora_lowergamma (int a, double x)
{
  double gamser, gammcf, gln, p;
  ensure_logfact ();
  if (x < 0.0 || a <= 0.0)
  {
    g_err = ORA_ERR_GAMMA;
    return NAN;
  }
  if (x < (a + 1.0))
  {
    gamser = gserlog (a, x, &gln);
    p = gln + gamser;
  }
  else
  {
    gammcf = gcf ((double) a, x, &gln);
    p = gln + log (1 - gammcf);
  }
  if (p < -1e200)
    p = -1e200;
  return p;
}

/* bessi0 utilities.cpp:54-91 */
static double
bessi0 (double x)
{
  double ax, ans, y;
  if ((ax = fabs (x)) < 3.75)
  {
    y = x / 3.75;
    y *= y;
    ans = 1.0 + y * (3.5156229 + y * (3.0899424 + y * (1.2067492 + y * (0.2659732 + y * (0.360768e-1 + y * 0.45813e-2)))));
  }
  else
  {
    y = 3.75 / ax;
    ans = (exp (ax) / sqrt (ax)) * (0.39894228 + y * (0.1328592e-1 + y * (0.225319e-2 + y * (-0.157565e-2 + y * (0.916281e-2 +
                                    y * (-0.2057706e-1 + y * (0.2635537e-1 + y * (-0.1647633e-1 + y * 0.392377e-2))))))));
  }
  return ans;
}

/* bessi1 utilities.cpp:93-124 */
static double
bessi1 (double x)
{
  double ax, ans, y;
  if ((ax = fabs (x)) < 3.75)
  {
    y = x / 3.75;
    y *= y;
    ans = ax * (0.5 + y * (0.87890594 + y * (0.51498869 + y * (0.15084934 + y * (0.2658733e-1 + y * (0.301532e-2 +
                y * 0.32411e-3))))));
  }
  else
  {
    y = 3.75 / ax;
    ans = 0.2282967e-1 + y * (-0.2895312e-1 + y * (0.1787654e-1 - y * 0.420059e-2));
    ans = 0.39894228 + y * (-0.3988024e-1 + y * (-0.362018e-2 + y * (0.163801e-2 + y * (-0.1031555e-1 + y * ans))));
    ans *= (exp (ax) / sqrt (ax));
  }
  return x < 0.0 ? -ans : ans;
}

/* bessi utilities.cpp:1452-1491 */
double
ora_bessi (int n, double x)
{
  int j;
  double bi, bim, bip, tox, ans;
  n = abs (n);
  if (x > 700)
    return MYDBL_MAX;
  if (n == 0)
    return bessi0 (x);
  if (n == 1)
    return bessi1 (x);
  if (x == 0.0)
    return 0.0;
  tox = 2.0 / fabs (x);
  bip = ans = 0.0;
  bi = 1.0;
  for (j = 2 * (n + (int) sqrt (40.0 * n)); j > 0; j--)
  {
    bim = bip + j * tox * bi;
    bip = bi;
    bi = bim;
    if (fabs (bi) > 1.0e10)
    {
      ans *= 1.0e-10;
      bi *= 1.0e-10;
      bip *= 1.0e-10;
    }
    if (j == n)
      ans = bip;
  }
  ans *= bessi0 (x) / bi;
  return x < 0.0 && (n & 1) ? -ans : ans;
}

/* eexp utilities.cpp:1501-1539 */
void
ora_eexp (double x, double *m, int *z)
{
  static const double us[10] = { 1.0, 0.5, 0.16666666666666666666666666667,
    0.04166666666666666666666666667, 0.00833333333333333333333333333,
    0.001388888888888888888888888889, 0.000198412698412698412698412698,
    0.000024801587301587301587301587301, 2.75573192239858906525573192239859e-6,
    2.75573192239858906525573192239e-7
  };
  double u, zr, temp;
  int n;
  n = (int) floor (x / LOG2);
  zr = 0.30102999566398119521 * (double) n;
  *z = (int) zr;
  zr -= (double) *z;
  u = x - (((double) n) * LOG2);
  temp = 1 + u * (us[0] + u * (us[1] + u * (us[2] + u * (us[3] + u * (us[4] + u * (us[5] + u * (us[6] + u * (us[7] + u *
                  (us[8] + u * (us[9]))))))))));
  *m = temp * pow (10.0, zr);
  if (fabs (*m) > 10)
  {
    *m = *m / 10.0;
    *z = *z + 1;
  }
  if (fabs (*m) < 1)
  {
    *m = *m * 10.0;
    *z = *z - 1;
  }
}

/* utilities.cpp:239-255 */
double
ora_mylogcosh (double x)
{
  return x < 100 ? log (cosh (x)) : x - LOG2;
}

double
ora_mylogsinh (double x)
{
  return x < 100 ? log (sinh (x)) : x - LOG2;
}

/* calcmrate update_gtree_common.cpp:462-482 */
double
ora_calcmrate (int mc, double mt)
{
  if (mt <= 0.0)
    return 1.0;
  if (mc == 0)
    return mt < 1 ? 0.1 : 0.1 / mt;
  return mt < 1 ? (double) mc : ((double) mc) / mt;
}

/* LogDiff imamp.hpp:257-263; returns 0 and records an error when a <= b (the reference exits) */
static int
logdiff (double *v, double a, double b)
{
  if (a <= b)
  {
    g_err = ORA_ERR_LOGDIFF;
    *v = NAN;
    return 0;
  }
  if (a - b < LOG_DBL_MAX)
    *v = b + log (exp (a - b) - 1.0);
  else
    *v = a;
  return 1;
}

/* ------------------------------------------------------------------------------------------- */
/* integrated prior: update_gtree_common.cpp:108-304 (MORESTABLE defined, imamp.hpp:103)        */
/* ------------------------------------------------------------------------------------------- */

double
ora_integrate_coalescent_term (int cc, double fc, double hcc, double max, double min)
{
  double p, a, b, c, d, ug, lg, ugalt, fullg;
  ensure_logfact ();
  if (cc > 0)
  {
    if (min == 0)
    {
      ug = ora_uppergamma (cc - 1, 2 * fc / max);
      if (cc > 1)
      {
        fullg = g_logfact[cc - 2];
        if (fullg - ug < 1e-15 || fullg - ug > LOG_DBL_MAX)
        {
          lg = ora_lowergamma (cc - 1, 2 * fc / max);
          if (fullg > lg)
          {
            logdiff (&ugalt, fullg, lg);
            if (fabs (ugalt - ug) > 1e-10)
              ug = ugalt;
          }
        }
      }
      p = ug + LOG2 - hcc + (1 - cc) * log (fc);
    }
    else
    {
      a = ora_uppergamma (cc - 1, 2 * fc / max);
      b = ora_uppergamma (cc - 1, 2 * fc / min);
      if (!logdiff (&p, a, b))
        return NAN;
      p += (LOG2 - hcc + (1 - cc) * log (fc));
    }
  }
  else if (2 * fc / max > 0)
  {
    if (min == 0)
    {
      a = log (max) - 2.0 * fc / max;
      b = LOG2 + log (fc) + ora_uppergamma (0, 2.0 * fc / max);
      if (!logdiff (&p, a, b))
        return NAN;
    }
    else
    {
      a = ora_uppergamma (0, 2 * fc / max);
      b = ora_uppergamma (0, 2 * fc / min);
      if (!logdiff (&c, a, b))
        return NAN;
      c += LOG2 + log (fc);
      a = log (max) - 2.0 * fc / max;
      b = log (min) - 2.0 * fc / min;
      if (!logdiff (&d, a, b))
        return NAN;
      if (!logdiff (&p, d, c))
        return NAN;
    }
  }
  else
    p = log (max - min);
  return p;
}

double
ora_integrate_migration_term (int cm, double fm, double max, double min)
{
  double p, a, b, c, ug, lg, lgalt, fullg;
  ensure_logfact ();
  if (cm > 0)
  {
    if (min == 0)
    {
      lg = ora_lowergamma (cm + 1, fm * max);
      fullg = g_logfact[cm];
      if (fullg - lg < 1e-15 || fullg - lg > LOG_DBL_MAX)
      {
        ug = ora_uppergamma (cm + 1, fm * max);
        if (fullg > ug)
        {
          logdiff (&lgalt, fullg, ug);
          if (fabs (lgalt - lg) > 1e-12)
            lg = lgalt;
        }
      }
      p = (-1 - cm) * log (fm) + lg;
    }
    else
    {
      a = ora_uppergamma (cm + 1, fm * min);
      b = ora_uppergamma (cm + 1, fm * max);
      if (!logdiff (&c, a, b))
        return NAN;
      p = (-1 - cm) * log (fm) + c;
    }
  }
  else if (fm > MPRIORMIN)
  {
    if (min == 0)
    {
      if (max == MPRIORMIN)
        p = 0;
      else
      {
        a = 0.0;
        b = -fm * max;
        if (!logdiff (&c, a, b))
          return NAN;
        p = c - log (fm);
      }
    }
    else
    {
      a = -fm * min;
      b = -fm * max;
      if (!logdiff (&c, a, b))
        return NAN;
      p = c - log (fm);
    }
  }
  else
    p = log (max - min);
  return p;
}

double
ora_integrate_migration_term_expo_prior (int cm, double fm, double exmean)
{
  ensure_logfact ();
  return -log (exmean) + (-(cm + 1) * log (fm + 1.0 / exmean)) + g_logfact[cm];
}

/* forbidden-migration check, update_gtree_common.cpp:1958-1979 */
static int
checkm (const ora_model * m, const int *mc)
{
  int i;
  if (m->nomigration == 0 && m->nomig_n > 0)
    for (i = 0; i < m->nomig_n; i++)
      if (mc[m->nomig_idx[i]] != 0)
        return 0;
  return 1;
}

/* initialize_integrate_tree_prob update_gtree_common.cpp:2056-2134 */
double
ora_initialize_integrate_tree_prob (const ora_model * m, const int *cc, const double *fc, const double *hcc, const int *mc,
                                    const double *fm, double *qint, double *mint)
{
  double psum, f, hc;
  int i, j, c;
  if (!checkm (m, mc))
    return -MYDBL_MAX;
  psum = 0;
  for (i = 0; i < m->nq; i++)
  {
    c = 0;
    f = hc = 0.0;
    for (j = 0; j < m->q_n[i]; j++)
    {
      c += cc[m->q_idx[i][j]];
      f += fc[m->q_idx[i][j]];
      hc += hcc[m->q_idx[i][j]];
    }
    qint[i] = ora_integrate_coalescent_term (c, f, hc, m->q_max[i], m->q_min[i]);
    psum += qint[i];
  }
  if (!m->nomigration)
    for (i = 0; i < m->nm; i++)
    {
      c = 0;
      f = 0.0;
      for (j = 0; j < m->m_n[i]; j++)
      {
        c += mc[m->m_idx[i][j]];
        f += fm[m->m_idx[i][j]];
      }
      if (m->expoprior)
        mint[i] = ora_integrate_migration_term_expo_prior (c, f, m->m_mean[i]);
      else
        mint[i] = ora_integrate_migration_term (c, f, m->m_max[i], m->m_min[i]);
      psum += mint[i];
    }
  return psum;
}

/* integrate_tree_prob update_gtree_common.cpp:1944-2053 */
double
ora_integrate_tree_prob (const ora_model * m, const int *cc, const double *fc, const double *hcc, const int *mc,
                         const double *fm, const int *hold_cc, const double *hold_fc, const int *hold_mc,
                         const double *hold_fm, const double *hold_qint, const double *hold_mint, double *qint,
                         double *mint)
{
  double psum, f, holdf, hc;
  int i, j, c, holdc;
  if (!checkm (m, mc))
    return -MYDBL_MAX;
  psum = 0;
  for (i = 0; i < m->nq; i++)
  {
    c = holdc = 0;
    f = holdf = hc = 0.0;
    for (j = 0; j < m->q_n[i]; j++)
    {
      c += cc[m->q_idx[i][j]];
      holdc += hold_cc[m->q_idx[i][j]];
      f += fc[m->q_idx[i][j]];
      holdf += hold_fc[m->q_idx[i][j]];
      hc += hcc[m->q_idx[i][j]];
    }
    if (c == holdc && f == holdf)
      qint[i] = hold_qint[i];
    else
      qint[i] = ora_integrate_coalescent_term (c, f, hc, m->q_max[i], m->q_min[i]);
    psum += qint[i];
  }
  if (!m->nomigration)
    for (i = 0; i < m->nm; i++)
    {
      c = holdc = 0;
      f = holdf = 0.0;
      for (j = 0; j < m->m_n[i]; j++)
      {
        c += mc[m->m_idx[i][j]];
        holdc += hold_mc[m->m_idx[i][j]];
        f += fm[m->m_idx[i][j]];
        holdf += hold_fm[m->m_idx[i][j]];
      }
      if (c == holdc && f == holdf)
        mint[i] = hold_mint[i];
      else if (m->expoprior)
        mint[i] = ora_integrate_migration_term_expo_prior (c, f, m->m_mean[i]);
      else
        mint[i] = ora_integrate_migration_term (c, f, m->m_max[i], m->m_min[i]);
      psum += mint[i];
    }
  return psum;
}

/* sum_subtract_treeinfo ginfo.cpp:248-285 (same operation order: subtract, add, clamp) */
void
ora_sum_subtract_treeinfo (const ora_model * m, int *all_cc, double *all_fc, double *all_hcc, int *all_mc, double *all_fm,
                           const int *p_cc, const double *p_fc, const double *p_hcc, const int *p_mc, const double *p_fm,
                           const int *n_cc, const double *n_fc, const double *n_hcc, const int *n_mc, const double *n_fm)
{
  int i;
  for (i = 0; i < m->ncc; i++)
  {
    all_cc[i] += p_cc[i] - n_cc[i];
    all_fc[i] -= n_fc[i];
    all_fc[i] += p_fc[i];
    if (!(all_fc[i] > 0))
      all_fc[i] = 0;
    all_hcc[i] -= n_hcc[i];
    all_hcc[i] += p_hcc[i];
  }
  if (!m->nomigration)
    for (i = 0; i < m->nmc; i++)
    {
      all_mc[i] += p_mc[i] - n_mc[i];
      all_fm[i] -= n_fm[i];
      all_fm[i] += p_fm[i];
      if (!(all_fm[i] > 0))
        all_fm[i] = 0;
    }
}

/* ------------------------------------------------------------------------------------------- */
/* treeweight: update_gtree_common.cpp:1679-1931                                                */
/* ------------------------------------------------------------------------------------------- */

/* findperiod update_gtree_common.cpp:1625-1633 */
static int
findperiod (const ora_model * m, const double *tvals, double t)
{
  int k = 0;
  while (k < m->nsplit && tvals[k] <= t)
    k++;
  return k;
}

typedef struct
{
  double time;
  int pop, topop, cmt, seq;
} ora_event;

static int
event_cmp (const void *a, const void *b)
{
  const ora_event *x = (const ora_event *) a, *y = (const ora_event *) b;
  if (x->time < y->time)
    return -1;
  if (x->time > y->time)
    return 1;
  return x->seq - y->seq;
}

int
ora_treeweight (const ora_model * m, const double *tvals, int numgenes, const int *samppop, double hval, const int *up0,
                const int *up1, const int *down, const int *pop, const double *time, const int *mig_off,
                const double *mig_t, const int *mig_p, int root, double roottime, int *cc, double *fc, double *hcc,
                int *mc, double *fm, double *out_length_tlength)
{
  int ng = numgenes, nl = 2 * numgenes - 1, npops = m->npops;
  int i, ii, j, jj, k, n[ORA_MAXTREEPOPS], nsum, ncount, mignum, ec, ip, jp, nowpop, per;
  double t, timeinterval, fmtemp, lasttime, h2term, hlog, lastsplitt, timeadd, length, tlength;
  ora_event *ev;
  (void) up1;
  (void) down;
  (void) root;
  memset (cc, 0, m->ncc * sizeof (int));
  memset (fc, 0, m->ncc * sizeof (double));
  memset (hcc, 0, m->ncc * sizeof (double));
  memset (mc, 0, m->nmc * sizeof (int));
  memset (fm, 0, m->nmc * sizeof (double));
  mignum = mig_off[nl] - mig_off[0];
  ncount = ng - 1 + mignum + findperiod (m, tvals, roottime);
  ev = (ora_event *) malloc ((ncount + 2) * sizeof (ora_event));
  ec = 0;
  for (i = 0; i < nl; i++)      /* :1741-1786 */
  {
    nowpop = pop[i];
    if (i >= ng)
    {
      t = time[up0[i]];
      ev[ec].time = t;
      ev[ec].cmt = 0;
      ev[ec].pop = nowpop;
      ev[ec].topop = -1;
      ev[ec].seq = ec;
      ec++;
    }
    for (j = mig_off[i]; j < mig_off[i + 1]; j++)
    {
      t = mig_t[j];
      per = findperiod (m, tvals, t);
      while (m->pt_e[nowpop] <= per && m->pt_e[nowpop] != -1)
        nowpop = m->pt_down[nowpop];
      ev[ec].time = t;
      ev[ec].pop = nowpop;
      ev[ec].topop = mig_p[j];
      ev[ec].cmt = 1;
      ev[ec].seq = ec;
      nowpop = mig_p[j];
      ec++;
    }
  }
  k = findperiod (m, tvals, roottime);  /* :1788-1799 */
  for (i = 0; i < k; i++)
  {
    ev[ec].time = tvals[i];
    ev[ec].cmt = -1;
    ev[ec].pop = ev[ec].topop = -1;
    ev[ec].seq = ec;
    ec++;
  }
  if (ec != ncount)
  {
    free (ev);
    g_err = ORA_ERR_TREE;
    return -1;
  }
  qsort (ev, ec, sizeof (ora_event), event_cmp);        /* indexx :1801 */
  for (i = 0; i < npops; i++)
    n[i] = samppop[i];
  for (i = npops; i < 2 * npops - 1; i++)
    n[i] = 0;
  nsum = ng;
  lasttime = 0;
  length = tlength = 0;
  h2term = 1 / (2 * hval);
  lastsplitt = m->nsplit > 0 ? tvals[m->nsplit - 1] : ORA_TIMEMAX;
  k = 0;
  for (j = 0; j < ec; j++)      /* :1826-1921 */
  {
    ora_event *e = &ev[j];
    timeinterval = e->time - lasttime;
    timeadd = nsum * timeinterval;
    length += timeadd;
    if (e->time < lastsplitt)
      tlength += timeadd;
    else if (lasttime < lastsplitt)
      tlength += nsum * (lastsplitt - lasttime);
    lasttime = e->time;
    for (ii = 0; ii < npops - k; ii++)
    {
      ip = npops > 1 ? m->plist[k][ii] : 0;
      fc[ccidx (m, k, ii)] += ((double) n[ip] * ((double) n[ip] - 1)) * timeinterval * h2term;   /* nnminus1 :543 */
      if (!m->nomigration && k < m->nsplit)
      {
        fmtemp = n[ip] * timeinterval;
        for (jj = 0; jj < npops - k; jj++)
          if (jj != ii)
            fm[mcidx (m, k, ii, jj)] += fmtemp;
      }
    }
    switch (e->cmt)
    {
    case 0:
      if (npops > 1)
      {
        ip = e->pop;
        ii = 0;
        while (ii < npops - k && m->plist[k][ii] != ip)
          ii++;
        if (ii >= npops - k || n[ip] < 2)
        {
          free (ev);
          g_err = ORA_ERR_TREE;
          return -1;
        }
      }
      else
        ip = ii = 0;
      cc[ccidx (m, k, ii)]++;
      n[ip]--;
      nsum--;
      break;
    case 1:
      ip = e->pop;
      jp = e->topop;
      ii = jj = 0;
      while (ii < npops - k && m->plist[k][ii] != ip)
        ii++;
      while (jj < npops - k && m->plist[k][jj] != jp)
        jj++;
      if (ii >= npops - k || jj >= npops - k || n[ip] < 1 || k >= m->nsplit)
      {
        free (ev);
        g_err = ORA_ERR_TREE;
        return -1;
      }
      mc[mcidx (m, k, ii, jj)]++;
      n[ip]--;
      n[jp]++;
      break;
    case -1:
      k++;
      n[m->addpop[k]] = n[m->droppops[k][0]] + n[m->droppops[k][1]];
      n[m->droppops[k][0]] = n[m->droppops[k][1]] = 0;
      break;
    }
  }
  hlog = log (hval);
  if (hlog != 0.0)
    for (i = 0; i < m->ncc; i++)
      hcc[i] += hlog * cc[i];
  free (ev);
  if (nsum != 1)
  {
    g_err = ORA_ERR_TREE;
    return -1;
  }
  out_length_tlength[0] = length;
  out_length_tlength[1] = tlength;
  return mignum;
}

/* ------------------------------------------------------------------------------------------- */
/* infinite sites: calc_prob_data.cpp:537-581 (labelgtree), 609-719 (calc_sumlogk), 731-836      */
/* ------------------------------------------------------------------------------------------- */

static void
labelgtree (int *mut, const int *up0, const int *up1, const int *down, int edge)
{
  int dow1, sis, flag = 1;
  while (flag)
  {
    flag = 0;
    dow1 = down[edge];
    if (dow1 == -1)
      break;
    if ((sis = up0[dow1]) == edge)
      sis = up1[dow1];
    if (mut[dow1] != mut[edge])
    {
      if (mut[edge] == 2)
      {
        if (mut[sis] != -1)
          mut[dow1] = mut[sis];
        else
          mut[dow1] = 2;
      }
      else if ((mut[sis] != -1 && mut[sis] != 2) && (mut[sis] != mut[edge]))
        mut[dow1] = 2;
      else
        mut[dow1] = mut[edge];
      flag = 1;
    }
    else if (mut[dow1] == 2)
    {
      if (mut[sis] != -1 && mut[sis] != 2)
      {
        mut[dow1] = mut[sis];
        flag = 1;
      }
    }
    edge = dow1;
  }
}

double
ora_calc_sumlogk (int numgenes, int numsites, const int *seq, const int *up0, const int *up1, const int *down)
{
  int ng = numgenes, nl = 2 * ng - 1, ret, node, site, a, b, i, j;
  int *mut = (int *) malloc (nl * sizeof (int));
  int *mutcount = (int *) calloc (ng, sizeof (int));
  double fact, sum = 0.0;
  for (site = 0; site < numsites; site++)
  {
    for (j = 0; j < ng; j++)
      mut[j] = seq[j * numsites + site];
    for (j = ng; j < nl; j++)
      mut[j] = -1;
    for (j = 0; j < ng; j++)
      labelgtree (mut, up0, up1, down, j);
    ret = -1;
    for (node = ng; node < nl; node++)
    {
      a = up0[node];
      b = up1[node];
      if ((mut[a] == 0 && mut[b] == 1) || (mut[a] == 1 && mut[b] == 0))
      {
        if (node != ret && ret != -1)
        {
          /* :654-658 returns with *psumlogk untouched (calloc'ed 0 by the caller :762) */
          free (mut);
          free (mutcount);
          return 0.0;
        }
        ret = node;
        mutcount[node - ng]++;
      }
    }
    if (ret == -1)
    {
      g_err = ORA_ERR_IS;
      free (mut);
      free (mutcount);
      return NAN;
    }
  }
  for (i = 0; i < nl - ng; i++)
  {
    fact = 1.0;
    for (j = 1; j <= mutcount[i]; j++)
      fact *= (double) j;
    sum += log (fact);
  }
  free (mut);
  free (mutcount);
  return sum;
}

double
ora_likelihoodIS (int numgenes, int numsites, const int *seq, const int *up0, const int *up1, const int *down,
                  const double *time, double length, double mutrate, double sumlogk)
{
  int ng = numgenes, nl = 2 * ng - 1, ret, j, node, site, a, b, upup;
  double ptime = 0, p;
  int *mut = (int *) malloc (nl * sizeof (int));
  p = -length * mutrate;
  for (site = 0; site < numsites; site++)
  {
    for (j = 0; j < ng; j++)
      mut[j] = seq[j * numsites + site];
    for (j = ng; j < nl; j++)
      mut[j] = -1;
    for (j = 0; j < ng; j++)
      labelgtree (mut, up0, up1, down, j);
    ret = -1;
    for (node = ng; node < nl; node++)
    {
      a = up0[node];
      b = up1[node];
      if ((mut[a] == 0 && mut[b] == 1) || (mut[a] == 1 && mut[b] == 0))
      {
        if (node != ret && ret != -1)
        {
          free (mut);
          return ORA_REJECT_IS;
        }
        if (down[node] == -1)
        {
          if ((upup = up0[a]) == -1)
            ptime = time[a];
          else
            ptime = time[a] - time[upup];
          if ((upup = up0[b]) == -1)
            ptime = ptime + time[b];
          else
            ptime = ptime + time[b] - time[upup];
        }
        else if (mut[down[node]] == mut[a])
        {
          if ((upup = up0[b]) == -1)
            ptime = time[b];
          else
            ptime = time[b] - time[upup];
        }
        else
        {
          if ((upup = up0[a]) == -1)
            ptime = time[a];
          else
            ptime = time[a] - time[upup];
        }
        ret = node;
      }
    }
    if (ret == -1)
    {
      g_err = ORA_ERR_IS;
      free (mut);
      return NAN;
    }
    p = p + log (ptime * mutrate);
  }
  p -= sumlogk;
  free (mut);
  return p;
}

/* ------------------------------------------------------------------------------------------- */
/* HKY: calc_prob_data.cpp:26-41 (pijt), 118-471 (makefrac, full recompute e1 == -1), 473-493,  */
/* 583-607                                                                                      */
/* ------------------------------------------------------------------------------------------- */

static double
pijt (const double *pi, double mutrate, double t, double kappa, int from, int to)
{
  double A, PIj;
  if (to == 0 || to == 2)
    PIj = pi[0] + pi[2];
  else
    PIj = pi[1] + pi[3];
  A = 1.0 + PIj * (kappa - 1.0);
  if (from == to)
    return pi[to] + pi[to] * exp (-mutrate * t) * (1.0 / PIj - 1.0) + exp (-mutrate * t * A) * ((PIj - pi[to]) / PIj);
  else if (from + to == 2 || from + to == 4)
    return pi[to] + pi[to] * (1.0 / PIj - 1.0) * exp (-mutrate * t) - (pi[to] / PIj) * exp (-mutrate * t * A);
  else
    return pi[to] * (1.0 - exp (-mutrate * t));
}

double
ora_likelihoodHKY (int numgenes, int numsites, int totsites, const int *seq, const int *mult, const int *up0,
                   const int *up1, const int *down, const double *time, int root, const double *pi, double mutrate,
                   double kappa)
{
  int ng = numgenes, nl = 2 * ng - 1, i, j, k, s, node, done, guard;
  double standfactor = 0, p = 0, fracp, max;
  double *frac = (double *) calloc ((size_t) nl * numsites * 4, sizeof (double));
  double *scale = (double *) calloc ((size_t) nl * numsites, sizeof (double));
  char *ready = (char *) calloc (nl, 1);
  (void) down;
  for (i = 0; i < 4; i++)       /* getstandfactor :473-493 */
    for (j = 0; j < 4; j++)
      if (i != j)
      {
        if (i + j == 2 || i + j == 4)
          standfactor += pi[i] * pi[j] * kappa;
        else
          standfactor += pi[i] * pi[j];
      }
  mutrate = mutrate / (totsites * standfactor); /* :593 */
  for (i = 0; i < ng; i++)
    ready[i] = 1;
  /* post-order evaluation of every internal node (makefrac recursion with e1 == -1) */
  for (done = ng, guard = 0; done < nl && guard <= nl; guard++)
    for (node = ng; node < nl; node++)
    {
      int a = up0[node], b = up1[node];
      if (ready[node] || !ready[a] || !ready[b])
        continue;
      double ta = a < ng ? time[a] : time[a] - time[up0[a]];
      double tb = b < ng ? time[b] : time[b] - time[up0[b]];
      for (s = 0; s < numsites; s++)
      {
        double *nf = &frac[((size_t) node * numsites + s) * 4];
        max = 0.0;
        for (j = 0; j < 4; j++)
        {
          double sa, sb;
          if (a < ng)
            sa = pijt (pi, mutrate, ta, kappa, j, seq[a * numsites + s]);
          else
            for (sa = 0, k = 0; k < 4; k++)
              sa += pijt (pi, mutrate, ta, kappa, j, k) * frac[((size_t) a * numsites + s) * 4 + k];
          if (b < ng)
            sb = pijt (pi, mutrate, tb, kappa, j, seq[b * numsites + s]);
          else
            for (sb = 0, k = 0; k < 4; k++)
              sb += pijt (pi, mutrate, tb, kappa, j, k) * frac[((size_t) b * numsites + s) * 4 + k];
          nf[j] = sa * sb;
          if (nf[j] > max)
            max = nf[j];
        }
        for (j = 0; j < 4; j++)
          nf[j] = nf[j] / max;
        scale[(size_t) node * numsites + s] =
          (a < ng ? 0.0 : scale[(size_t) a * numsites + s]) + (b < ng ? 0.0 : scale[(size_t) b * numsites + s]) + log (max);
      }
      ready[node] = 1;
      done++;
    }
  for (s = 0; s < numsites; s++)        /* :597-605 */
  {
    fracp = 0;
    for (j = 0; j < 4; j++)
      fracp += pi[j] * frac[((size_t) root * numsites + s) * 4 + j];
    p += mult[s] * (log (fracp) + scale[(size_t) root * numsites + s]);
  }
  free (frac);
  free (scale);
  free (ready);
  return p;
}

/* likelihoodSW calc_prob_data.cpp:841-909 */
double
ora_likelihoodSW (int numgenes, const int *up0, const int *down, const double *time, const int *A, double u, double tr,
                  double *dlikeA)
{
  int nl = 2 * numgenes - 1, i, d, iszero = 0;
  double t, like = 0.0, bessiv;
  for (i = 0; i < nl; i++)
  {
    if (down[i] != -1)
    {
      d = A[i] - A[down[i]];
      if (up0[i] == -1)
        t = time[i] * tr;
      else
        t = tr * (time[i] - time[up0[i]]);
      bessiv = ora_bessi (d, t * u);
      if (!(bessiv > 0.0))
      {
        iszero = 1;
        dlikeA[i] = IM_BESSI_MIN;
      }
      else
        dlikeA[i] = -(t * u) + log (bessiv);
      like += dlikeA[i];
    }
    else
      dlikeA[i] = 0.0;
  }
  if (iszero)
    like = -DBL_MAX;
  return like;
}

/* ------------------------------------------------------------------------------------------- */
/* migration-path proposal probabilities                                                        */
/* ------------------------------------------------------------------------------------------- */

typedef struct
{
  int edgeid, pop, fpop, b, e, mpall;
  double upt, dnt, mtall;
  double mtimeavail[ORA_MAXPOPS + 1];
  int mp[ORA_MAXPOPS + 1];
  int nmig;
  const double *mig_t;
  const int *mig_p;
} ora_emi;

/* IMA_reset_edgemiginfo update_gtree_common.cpp:493-514 */
static void
emi_reset (ora_emi * em)
{
  memset (em, 0, sizeof (*em));
  em->edgeid = -1;
  em->b = em->e = -1;
  em->pop = em->fpop = -1;
  em->upt = em->dnt = -1.0;
}

/* fillmiginfoperiods update_gtree_common.cpp:1100-1162; tv has the TIMEMAX sentinel at [nsplit] */
static void
emi_periods (const ora_model * m, const double *tv, ora_emi * em)
{
  int i, last = m->nsplit;
  em->b = 0;
  while (em->upt > tv[em->b])
    em->b++;
  em->e = em->b;
  while (em->dnt > tv[em->e])
    em->e++;
  if (em->e == em->b)
  {
    if (em->b == last)
      em->mtimeavail[em->b] = 0;
    else
      em->mtimeavail[em->b] = em->dnt - em->upt;
  }
  else
  {
    em->mtimeavail[em->b] = tv[em->b] - em->upt;
    if (em->e == last)
      em->mtimeavail[em->e] = 0;
    else
      em->mtimeavail[em->e] = em->dnt - tv[em->e - 1];
    for (i = em->b + 1; i < em->e; i++)
      em->mtimeavail[i] = tv[i] - tv[i - 1];
  }
  if (em->b < last)
    em->mtall = (tv[last - 1] < em->dnt ? tv[last - 1] : em->dnt) - em->upt;
  else
    em->mtall = 0;
}

/* fillmiginfo update_gtree_common.cpp:1169-1246 (one edge) and the new-edge set-up of
 * addmigration update_gtree.cpp:597-636: period table + per-period counts of the edge's list */
static void
emi_fill (const ora_model * m, const double *tv, ora_emi * em, int numgenes, int edge, const int *up0, const int *pop,
          const double *time, const int *mig_off, const double *mig_t, const int *mig_p, int fpop)
{
  int i, j;
  emi_reset (em);
  em->edgeid = edge;
  em->upt = edge < numgenes ? 0 : time[up0[edge]];
  em->pop = pop[edge];
  em->dnt = time[edge];
  em->fpop = fpop;
  emi_periods (m, tv, em);
  em->nmig = mig_off[edge + 1] - mig_off[edge];
  em->mig_t = mig_t + mig_off[edge];
  em->mig_p = mig_p + mig_off[edge];
  j = em->b;
  for (i = 0; i < em->nmig; i++)
  {
    while (em->mig_t[i] > tv[j])
      j++;
    em->mp[j]++;
    em->mpall++;
  }
}

/* getmprob update_gtree_common.cpp:850-1071 */
static double
getmprob (const ora_model * m, const double *tv, const ora_emi * edgem, const ora_emi * sisem, const ora_emi * oldedgem,
          const ora_emi * oldsisem)
{
  double tempp = 0, n = 0, d, t, r, pathc;
  int periodi, cm[2], pop[2], topop, popc, lastm_2_pop, ii, lastmigrationperiod;
  int npops = m->npops, last = m->nsplit;
  const ora_emi *mm, *oldmm;
  lastmigrationperiod = edgem->e < last - 1 ? edgem->e : last - 1;
  if (sisem->mtall <= 0)
  {
    cm[0] = 0;
    for (periodi = edgem->b; periodi <= edgem->e; periodi++)
      if ((periodi < lastmigrationperiod) || (periodi == lastmigrationperiod && edgem->e == last))
      {
        r = ora_calcmrate (oldedgem->mp[periodi], oldedgem->mtimeavail[periodi]) * edgem->mtimeavail[periodi];
        tempp += edgem->mp[periodi] * log (r / (edgem->mtimeavail[periodi] * (npops - (periodi + 1)))) - r;
        cm[0] += edgem->mp[periodi];
      }
    if (edgem->e < last)
    {
      r = ora_calcmrate (oldedgem->mp[edgem->e], oldedgem->mtimeavail[edgem->e]) * edgem->mtimeavail[edgem->e];
      if (edgem->e == last - 1)
      {
        if (edgem->mp[edgem->e] & 1)
          tempp += edgem->mp[edgem->e] * log (r / edgem->mtimeavail[edgem->e]) - ora_mylogsinh (r);
        else
          tempp += edgem->mp[edgem->e] * log (r / edgem->mtimeavail[edgem->e]) - ora_mylogcosh (r);
      }
      else
      {
        if (cm[0] == 0)
          pop[0] = edgem->pop;
        else
          pop[0] = edgem->mig_p[cm[0] - 1];
        while (m->pt_e[pop[0]] <= edgem->e)
          pop[0] = m->pt_down[pop[0]];
        topop = edgem->fpop;
        popc = (npops - edgem->e - 1);
        if (pop[0] == topop)
          d = log (1 - r * exp (-r));
        else
          d = log ((1 - exp (-r)));
        switch (edgem->mp[edgem->e])
        {
        case 0:
          n = -r;
          break;
        case 1:
          n = log (r / edgem->mtimeavail[edgem->e]) - r;
          break;
        default:
          if (edgem->mp[edgem->e] == 2)
            lastm_2_pop = pop[0];
          else
            lastm_2_pop = edgem->mig_p[edgem->mpall - 3];
          if (lastm_2_pop == topop)
            pathc = -log ((double) popc);
          else
            pathc = -log ((double) popc - 1);
          n = edgem->mp[edgem->e] * log (r / edgem->mtimeavail[edgem->e]) - r + (2 - edgem->mp[edgem->e]) * log ((double) popc) + pathc;
        }
        tempp += n - d;
      }
    }
  }
  else
  {
    if (edgem->mtall > 0 && sisem->mtall > 0 && edgem->e < last)
    {
      for (ii = 0; ii < 2; ii++)
      {
        mm = (ii == 0) ? edgem : sisem;
        pop[ii] = mm->pop;
        if (mm->mpall > 0 && mm->e > 0)
        {
          t = tv[mm->e - 1];
          periodi = -1;
          while (periodi + 1 < mm->nmig && mm->mig_t[periodi + 1] >= 0 && mm->mig_t[periodi + 1] < t)
            periodi++;
          if (periodi >= 0)
            pop[ii] = mm->mig_p[periodi];
        }
        if (mm->e > 0)
          while (m->pt_e[pop[ii]] <= mm->e)
            pop[ii] = m->pt_down[pop[ii]];
      }
      if (pop[0] == pop[1])
      {
        if (pop[0] == edgem->fpop)
          tempp = log (MIGCLOSEFRAC);
        else
          tempp = log ((1.0 - MIGCLOSEFRAC) / (double) (npops - edgem->e - 1));
      }
      else
      {
        if (npops - edgem->e == 2)
          tempp = log (0.5);
        else
        {
          if (edgem->fpop == pop[0] || edgem->fpop == pop[1])
            tempp = log (0.5 * MIGCLOSEFRAC);
          else
            tempp = log ((1.0 - MIGCLOSEFRAC) / (double) (npops - edgem->e - 2));
        }
      }
    }
    else
      tempp = 0.0;
    lastmigrationperiod = edgem->e < last - 1 ? edgem->e : last - 1;
    for (ii = 0; ii < 2; ii++)
    {
      mm = (ii == 0) ? edgem : sisem;
      oldmm = (ii == 0) ? oldedgem : oldsisem;
      if (mm->mtall > 0)
      {
        cm[ii] = 0;
        for (periodi = mm->b; periodi <= mm->e; periodi++)
          if ((periodi < lastmigrationperiod) || (periodi == lastmigrationperiod && mm->e == last))
          {
            r = ora_calcmrate (oldmm->mp[periodi], oldmm->mtimeavail[periodi]) * mm->mtimeavail[periodi];
            tempp += mm->mp[periodi] * log (r / (mm->mtimeavail[periodi] * (npops - (periodi + 1)))) - r;
            cm[ii] += mm->mp[periodi];
          }
        if (mm->e < last)
        {
          r = ora_calcmrate (oldmm->mp[mm->e], oldmm->mtimeavail[mm->e]) * mm->mtimeavail[mm->e];
          if (mm->e == last - 1)
          {
            if (mm->mp[mm->e] & 1)
              tempp += mm->mp[mm->e] * log (r / mm->mtimeavail[mm->e]) - ora_mylogsinh (r);
            else
              tempp += mm->mp[mm->e] * log (r / mm->mtimeavail[mm->e]) - ora_mylogcosh (r);
          }
          else
          {
            if (cm[ii] == 0)
              pop[ii] = mm->pop;
            else
              pop[ii] = mm->mig_p[cm[ii] - 1];
            while (m->pt_e[pop[ii]] <= mm->e)
              pop[ii] = m->pt_down[pop[ii]];
            topop = mm->fpop;
            popc = (npops - mm->e - 1);
            if (pop[ii] == topop)
              d = log (1 - r * exp (-r));
            else
              d = log ((1 - exp (-r)));
            switch (mm->mp[mm->e])
            {
            case 0:
              n = -r;
              break;
            case 1:
              n = log (r / mm->mtimeavail[mm->e]) - r;
              break;
            default:
              if (mm->mp[mm->e] == 2)
                lastm_2_pop = pop[ii];
              else
                lastm_2_pop = mm->mig_p[mm->mpall - 3];
              if (lastm_2_pop == topop)
                pathc = -log ((double) popc);
              else
                pathc = -log ((double) popc - 1);
              n = mm->mp[mm->e] * log (r / mm->mtimeavail[mm->e]) - r + (2 - mm->mp[mm->e]) * log ((double) popc) + pathc;
              break;
            }
            tempp += n - d;
          }
        }
      }
    }
  }
  return tempp;
}

/* old edge infos as fillmiginfo (update_gtree.cpp:765-772), new ones as addmigration
 * (update_gtree.cpp:597-651), then the two getmprob calls of update_gtree.cpp:654-663 */
int
ora_migration_proposal_logprobs (const ora_model * m, const double *tvals, int numgenes, const int *b_up0, const int *b_up1,
                                 const int *b_down, const int *b_pop, const double *b_time, const int *b_mig_off,
                                 const double *b_mig_t, const int *b_mig_p, int b_root, const int *a_up0, const int *a_up1,
                                 const int *a_down, const int *a_pop, const double *a_time, const int *a_mig_off,
                                 const double *a_mig_t, const int *a_mig_p, int a_root, int edge, double *out)
{
  double tv[ORA_MAXPERIODS + 1];
  ora_emi oe, os, ne, ns;
  int i, sis;
  for (i = 0; i < m->nsplit; i++)
    tv[i] = tvals[i];
  tv[m->nsplit] = ORA_TIMEMAX;  /* initialize.cpp:1948 */
  emi_fill (m, tv, &oe, numgenes, edge, b_up0, b_pop, b_time, b_mig_off, b_mig_t, b_mig_p, b_pop[b_down[edge]]);
  if (b_down[edge] == b_root)
  {
    sis = b_up0[b_down[edge]] == edge ? b_up1[b_down[edge]] : b_up0[b_down[edge]];
    emi_fill (m, tv, &os, numgenes, sis, b_up0, b_pop, b_time, b_mig_off, b_mig_t, b_mig_p, b_pop[b_down[sis]]);
  }
  else
    emi_reset (&os);
  if (a_down[edge] == a_root)
  {
    /* both edges end in the population of the new root (chosen in mwork_two_edges :1479-1524 and written
     * back by copynewmig_to_gtree :1279-1287) */
    sis = a_up0[a_down[edge]] == edge ? a_up1[a_down[edge]] : a_up0[a_down[edge]];
    emi_fill (m, tv, &ne, numgenes, edge, a_up0, a_pop, a_time, a_mig_off, a_mig_t, a_mig_p, a_pop[a_root]);
    emi_fill (m, tv, &ns, numgenes, sis, a_up0, a_pop, a_time, a_mig_off, a_mig_t, a_mig_p, a_pop[a_root]);
  }
  else
  {
    emi_fill (m, tv, &ne, numgenes, edge, a_up0, a_pop, a_time, a_mig_off, a_mig_t, a_mig_p, a_pop[a_down[edge]]);
    emi_reset (&ns);
  }
  out[0] = getmprob (m, tv, &ne, &ns, &oe, &os);
  out[1] = getmprob (m, tv, &oe, &os, &ne, &ns);
  return 0;
}

/* update_gtree.cpp:782,803-812 with normprob utilities.cpp:463-468 */
double
ora_slideweight (double slidedist, double oldroottime, double newroottime)
{
  const double c = 0.3989422803;
  double s0 = oldroottime / 3 < 20 ? oldroottime / 3 : 20;
  double s1 = newroottime / 3 < 20 ? newroottime / 3 : 20;
  double z0 = slidedist / s0, z1 = slidedist / s1;
  double w = -log (c * exp (-(z0 == 0.0 ? 0.0 : z0 * z0) / 2) / s0);
  w += log (c * exp (-(z1 == 0.0 ? 0.0 : z1 * z1) / 2) / s1);
  return w;
}

/* ------------------------------------------------------------------------------------------- */
/* MC3: swapchains.cpp:12-62, 71-178                                                            */
/* ------------------------------------------------------------------------------------------- */

double
ora_swapweight (double sumi, double sumj, double betai, double betaj)
{
  return exp ((betai - betaj) * (sumj - sumi));
}

/* beta[] of setheat (swapchains.cpp:119-163) for one process holding all chains */
void
ora_setheat (int heatmode, double hval1, double hval2, int nchains, double *betas)
{
  int ci;
  if (nchains == 1)
  {
    betas[0] = 1.0;
    return;
  }
  for (ci = 0; ci < nchains; ci++)
    switch (heatmode)
    {
    case 0:
      if (hval1 < 0.05)
        hval1 = 0.05;
      betas[ci] = 1.0 / (1.0 + hval1 * ci);
      break;
    case 1:
      betas[ci] = 1 - (1 - hval2) * (ci) * pow (hval1, (double) (nchains - 1 - (ci))) / (double) (nchains - 1);
      break;
    default:
      betas[ci] = 1.0 - ci * (1.0 / (nchains - 1));
      break;
    }
}

/* ------------------------------------------------------------------------------------------- */
/* .ti row: ginfo.cpp:288-377; column offsets initialize.cpp:710-719                             */
/* ------------------------------------------------------------------------------------------- */

void
ora_savegsampinf (const ora_model * m, const int *cc, const double *fc, const double *hcc, const int *mc, const double *fm,
                  const double *qint, const double *mint, double pdg, double probg, const double *tvals, float *row)
{
  int i, j, c, nq = m->nq, nm = m->nomigration ? 0 : m->nm;
  int ccp = 0, fcp = nq, hccp = 2 * nq, mcp = 3 * nq, fmp = mcp + nm, qip = fmp + nm, mip = qip + nq, pdgp = mip + nm,
    probgp = pdgp + 1, tp = probgp + 1;
  float f, hc;
  for (i = 0; i < nq; i++)
  {
    c = 0;
    f = hc = 0.0;
    for (j = 0; j < m->q_n[i]; j++)
    {
      c += cc[m->q_idx[i][j]];
      f += (float) fc[m->q_idx[i][j]];
      hc += (float) hcc[m->q_idx[i][j]];
    }
    row[ccp + i] = (float) c;
    row[fcp + i] = f;
    row[hccp + i] = hc;
  }
  for (i = 0; i < nm; i++)
  {
    c = 0;
    f = 0.0;
    for (j = 0; j < m->m_n[i]; j++)
    {
      c += mc[m->m_idx[i][j]];
      f += (float) fm[m->m_idx[i][j]];
    }
    row[mcp + i] = (float) c;
    row[fmp + i] = f;
  }
  for (i = 0; i < nq; i++)
    row[qip + i] = (float) qint[i];
  for (i = 0; i < nm; i++)
    row[mip + i] = (float) mint[i];
  row[pdgp] = (float) pdg;
  row[probgp] = (float) probg;
  for (i = 0; i < m->nsplit; i++)
    row[tp + i] = (float) tvals[i];
}

/* ------------------------------------------------------------------------------------------- */
/* L mode                                                                                       */
/* ------------------------------------------------------------------------------------------- */

/* marginp surface_call_functions.cpp:25-80 */
double
ora_marginp (const ora_model * m, const float *rows, int rowlen, int param, int firsttree, int lasttree, double x)
{
  int ei, p, nq = m->nq, nm = m->nm;
  int ccp = 0, fcp = nq, hccp = 2 * nq, mcp = 3 * nq, fmp = mcp + nm, qip = fmp + nm, mip = qip + nq;
  double hval, temp, max, min, sumtemp = 0, meani = 0;
  if (param < nq)
  {
    max = m->q_max[param];
    min = m->q_min[param];
  }
  else
  {
    max = m->m_max[param - nq];
    min = m->m_min[param - nq];
    if (m->expoprior)
      meani = 1.0 / m->m_mean[param - nq];
  }
  if (x < min || x > max)
    return 1;
  for (ei = firsttree; ei < lasttree; ei++)
  {
    const float *g = rows + (size_t) ei * rowlen;
    if (param < nq)
    {
      p = param;
      hval = g[hccp + p];
      temp = -g[qip + p] + g[ccp + p] * (LOG2 - log (x)) - hval - 2 * g[fcp + p] / x;
      sumtemp += exp (temp);
    }
    else
    {
      p = param - nq;
      if (m->expoprior)
        temp = -g[mip + p] + log (meani) - x * meani + INTEGERROUND (g[mcp + p]) * log (x) - g[fmp + p] * x;
      else
        temp = -g[mip + p] + INTEGERROUND (g[mcp + p]) * log (x) - g[fmp + p] * x;
      sumtemp += exp (temp);
    }
  }
  sumtemp /= (lasttree - firsttree + (firsttree == 0));
  return -sumtemp;
}

/* margincalc surface_call_functions.cpp:119-173 */
double
ora_margincalc (const ora_model * m, const float *rows, int rowlen, int nrows, double x, double yadjust, int pi, int logi)
{
  int ei, p, nq = m->nq, nm = m->nm;
  int ccp = 0, fcp = nq, hccp = 2 * nq, mcp = 3 * nq, fmp = mcp + nm, qip = fmp + nm, mip = qip + nq;
  double hval, sum = 0, temp, meani = 0;
  if (m->expoprior && pi >= nq && pi < nq + nm)
    meani = 1.0 / m->m_mean[pi - nq];
  for (ei = 0; ei < nrows; ei++)
  {
    const float *g = rows + (size_t) ei * rowlen;
    if (pi < nq)
    {
      p = pi;
      hval = g[hccp + p];
      temp = -g[qip + p] + INTEGERROUND (g[ccp + p]) * (LOG2 - log (x)) - hval - 2 * g[fcp + p] / x;
      sum += exp (temp);
    }
    else
    {
      p = pi - nq;
      if (m->expoprior)
        sum += exp (-g[mip + p] + log (meani) - x * meani + INTEGERROUND (g[mcp + p]) * log (x) - g[fmp + p] * x);
      else
        sum += exp (-g[mip + p] + INTEGERROUND (g[mcp + p]) * log (x) - g[fmp + p] * x);
    }
  }
  sum /= nrows;
  if (logi)
    sum = sum <= 0 ? -1e200 : log (sum);
  sum -= yadjust;
  return sum;
}

/* jointp jointfind.cpp:885-1047.  modeltype 0: the full model of a two-population analysis (p starts at -probg, all
 * parameters, :973, :1110-1113); 1 / 2: the two "full" models of a three-population analysis -- all population sizes
 * (p starts at minus the sum of the size integrals, :949-950, :976; parameters [0, nq)) and all migration rates (minus the
 * sum of the migration integrals, :951-952, :978; parameters [nq, nq + nm)), findjointpeaks :1118-1133.  The sorted linked list of
 * :191-270 is replaced by its observable behaviour: a term is *inserted* iff it lies within
 * PRANGELOG of the running maximum at the time it is met (:1005); the summation walks the inserted
 * terms downwards from the maximum while they lie within PRANGELOG of it, but visits at most
 * (number inserted - 1 + ... ) = iin terms, iin being the count of insertions after the first (:1011-1022). */
static int
dbl_desc (const void *a, const void *b)
{
  double x = *(const double *) a, y = *(const double *) b;
  return x < y ? 1 : x > y ? -1 : 0;
}

double
ora_jointp (const ora_model * m, const float *rows, int rowlen, int nrows, const double *x, int calc_ess,
            double *effective_n)
{
  return ora_jointp_model (m, rows, rowlen, nrows, x, 0, calc_ess, effective_n);
}

double
ora_jointp_model (const ora_model * m, const float *rows, int rowlen, int nrows, const double *x, int modeltype,
                  int calc_ess, double *effective_n)
{
  int gi, i, i1, nq = m->nq, nm = m->nm, np = nq + nm;
  int ccp = 0, fcp = nq, hccp = 2 * nq, mcp = 3 * nq, fmp = mcp + nm, qip = fmp + nm, mip = qip + nq, probgp = fmp + nm + nq + nm + 1;
  int lo = modeltype == 2 ? nq : 0, hi = modeltype == 1 ? nq : np;
  double logx[ORA_MAXPARAMS], divx[ORA_MAXPARAMS], log2diffx[ORA_MAXPARAMS];
  double p, last = 0, acumm = 0, acumm_sqr = 0, sum;
  double *kept = (double *) malloc ((size_t) nrows * sizeof (double));
  double *em = (double *) malloc ((size_t) nrows * sizeof (double));
  int *ez = (int *) malloc ((size_t) nrows * sizeof (int));
  int nkept = 0, iin, gin, maxz = -10000000, zadj;
  for (i = 0; i < nq; i++)
    log2diffx[i] = LOG2 - log (x[i]);
  for (i = 0; i < np; i++)
  {
    logx[i] = log (x[i]);
    divx[i] = 1.0 / x[i];
  }
  for (gi = 0; gi < nrows; gi++)
  {
    const float *g = rows + (size_t) gi * rowlen;
    if (modeltype == 0)
      p = -g[probgp];
    else
    {
      double acc = 0;
      if (modeltype == 1)
        for (i = 0; i < nq; i++)
          acc += g[qip + i];
      else
        for (i = 0; i < nm; i++)
          acc += g[mip + i];
      p = -acc;
    }
    for (i = lo; i < hi; i++)
    {
      if (i < nq)
        p += g[ccp + i] * log2diffx[i] - g[hccp + i] - (2.0 * g[fcp + i]) * divx[i];
      else
      {
        i1 = i - nq;
        p += g[mcp + i1] * logx[i] - g[fmp + i1] * x[i];
      }
    }
    if (gi == 0)
    {
      kept[nkept++] = p;
      last = p;
    }
    else if (last - p < 10)
    {
      kept[nkept++] = p;
      if (p > last)
        last = p;
    }
  }
  iin = nkept - 1;
  qsort (kept, nkept, sizeof (double), dbl_desc);
  gi = 0;
  do
  {
    ora_eexp (kept[gi], &em[gi], &ez[gi]);
    if (ez[gi] > maxz)
      maxz = ez[gi];
    gi++;
  }
  while (gi < iin && last - kept[gi] < 10);
  gin = gi;
  maxz -= 10;
  for (gi = 0; gi < gin; gi++)
  {
    zadj = ez[gi] - maxz;
    if (zadj > -308 && zadj < 308)
      em[gi] *= pow (10.0, zadj);
    else if (zadj <= -308)
      em[gi] = 0.0;
    else
      em[gi] = DBL_MAX;
    acumm += em[gi];
    if (calc_ess)
      acumm_sqr += em[gi] * em[gi];
  }
  if (calc_ess && effective_n)
    *effective_n = acumm * acumm / acumm_sqr;
  sum = log (acumm) + maxz * LOG10;
  free (kept);
  free (em);
  free (ez);
  return log ((double) nrows) - sum;
}


/* ------------------------------------------------------------------------------------------- */
/* changet_RY1: update_t_RY.cpp:52-80 (aftersplit, beforesplit), 222-517; getnewt               */
/* update_gtree_common.cpp:2501-2519                                                            */
/* ------------------------------------------------------------------------------------------- */

double
ora_getnewt (double U, int nloci, int npops, int timeperiod, double t_u_prior, double t_d_prior, double oldt)
{
  double twin, newt;
  twin = (t_d_prior - t_u_prior) / (log ((double) nloci + 1) * (npops - timeperiod));
  newt = (oldt - twin / 2) + U * twin;
  if (newt >= t_d_prior)
    newt = 2.0 * t_d_prior - newt;
  else if (newt <= t_u_prior)
    newt = 2.0 * t_u_prior - newt;
  return newt;
}

static double
ry_aftersplit (int tnode, int lastperiodnumber, double oldt, double newt, double tau_d, double ptime)
{
  if (tnode == lastperiodnumber - 1)
    return ptime + newt - oldt;
  return tau_d - (tau_d - newt) * (tau_d - ptime) / (tau_d - oldt);
}

static double
ry_beforesplit (int tnode, double oldt, double newt, double tau_u, double ptime)
{
  if (tnode == 0)
    return ptime * newt / oldt;
  return tau_u + (ptime - tau_u) * (newt - tau_u) / (oldt - tau_u);
}

void
ora_ry1_rescale (int numlines, const int *down, double *time, int nmig, double *mig_t, double *roottime,
                 int timeperiod, int lastperiodnumber, double oldt, double newt, double t_u, double t_d, int *counts)
{
  int i;
  /* :268-320; migration events sit on non-root edges only, so the per-edge loop over them is a flat one here */
  for (i = 0; i < numlines; i++)
    if (down[i] != -1)
    {
      if (time[i] <= oldt && time[i] > t_u)
      {
        time[i] = ry_beforesplit (timeperiod, oldt, newt, t_u, time[i]);
        counts[0]++;
      }
      else if (time[i] > oldt && time[i] < t_d)
      {
        time[i] = ry_aftersplit (timeperiod, lastperiodnumber, oldt, newt, t_d, time[i]);
        counts[1]++;
      }
    }
  for (i = 0; i < nmig; i++)
  {
    if (mig_t[i] <= oldt && mig_t[i] > t_u)
    {
      mig_t[i] = ry_beforesplit (timeperiod, oldt, newt, t_u, mig_t[i]);
      counts[2]++;
    }
    else if (mig_t[i] > oldt && mig_t[i] < t_d)
    {
      mig_t[i] = ry_aftersplit (timeperiod, lastperiodnumber, oldt, newt, t_d, mig_t[i]);
      counts[3]++;
    }
  }
  /* :321-328 */
  if (*roottime <= oldt && *roottime > t_u)
    *roottime = ry_beforesplit (timeperiod, oldt, newt, t_u, *roottime);
  else if (*roottime > oldt && *roottime < t_d)
    *roottime = ry_aftersplit (timeperiod, lastperiodnumber, oldt, newt, t_d, *roottime);
}

double
ora_ry1_hastings (int timeperiod, int lastperiodnumber, double oldt, double newt, double t_u, double t_d, int ecu,
                  int ecd, int emu, int emd)
{
  double t_u_hterm = (newt - t_u) / (oldt - t_u), t_d_hterm;
  if (timeperiod == lastperiodnumber - 1)
    t_d_hterm = 1;
  else
    t_d_hterm = (t_d - newt) / (t_d - oldt);
  return (ecd + emd) * log (t_d_hterm) + (ecu + emu) * log (t_u_hterm);
}

/* ------------------------------------------------------------------------------------------- */
/* changeu / changekappa proposals: update_mc_params.cpp:201-212, 258-272                       */
/* ------------------------------------------------------------------------------------------- */

double
ora_changeu_newr (double U, double r, double windowsize, double maxratio, double *d)
{
  double newr;
  if (U > 0.5)
    newr = r + (2.0 * U - 1.0) * windowsize;
  else
    newr = r - windowsize * U * 2.0;
  if (newr > maxratio)
    newr = 2.0 * maxratio - newr;
  else if (newr < -maxratio)
    newr = 2.0 * (-maxratio) - newr;
  *d = exp ((newr - r) / 2);
  return newr;
}

double
ora_new_kappa (double U, double kappa, double win, double max)
{
  double nk;
  if (U > 0.5)
  {
    nk = kappa + (2.0 * U - 1.0) * win;
    if (nk > max)
      nk = 2.0 * max - nk;
  }
  else
  {
    nk = kappa - win * U * 2.0;
    if (nk < 0)
      nk = -nk;
  }
  return nk;
}

/* ------------------------------------------------------------------------------------------- */
/* thermomarginlikecalc: marglike.cpp:121-150 (Simpson's rule; the beta = 0 chain contributes 0) */
/* ------------------------------------------------------------------------------------------- */

double
ora_thermomarginlike (const double *thermosum, int numchains, int k)
{
  int i;
  double width = 1.0 / (float) (numchains - 1), sum = 0.0;
  for (i = 0; i <= numchains - 1; i += 2)
    if (i != numchains - 1)
      sum += 4.0 * (thermosum[i] / k);
  for (i = 1; i <= numchains - 2; i += 2)
    sum += 2.0 * (thermosum[i] / k);
  return width * sum / 3.0;
}


/* ------------------------------------------------------------------------------------------- */
/* changet_NW, replayed: update_t_NW.cpp:291-330 (getmprob_NW), 336-786 (update_mig_tNW);       */
/* nowedgepop update_gtree_common.cpp:1638-1655                                                 */
/* ------------------------------------------------------------------------------------------- */

static int
nw_nowedgepop (const ora_model * m, const double *tv, int pop, const int *off, const double *mt, const int *mp, int e,
               double ptime)
{
  int j;
  for (j = off[e]; j < off[e + 1] && mt[j] < ptime; j++)
    pop = mp[j];
  while (pop != -1 && ptime > (m->pt_e[pop] == -1 ? ORA_TIMEMAX : tv[m->pt_e[pop] - 1]))
    pop = m->pt_down[pop];
  return pop;
}

static double
nw_getmprob (const ora_model * m, int period, double mrate, double mtime, int mcount, int uppop, int dpop, int cm2pop,
             int numpops)
{
  double logb, logs, lognp;
  if (period == m->nsplit)
    return 0;
  if (period == m->nsplit - 1)
  {
    if (mcount & 1)
      return mcount * log (mrate / mtime) - ora_mylogsinh (mrate);
    return mcount * log (mrate / mtime) - ora_mylogcosh (mrate);
  }
  if (uppop == dpop)
    logs = log (1 - mrate * exp (-mrate));
  else
    logs = log (1 - exp (-mrate));
  switch (mcount)
  {
  case 0:
    return -mrate - logs;
  case 1:
    return log (mrate / mtime) - mrate - logs;
  default:
    lognp = log ((double) numpops - 1);
    if (cm2pop == dpop)
      logb = -lognp;
    else
      logb = -log ((double) numpops - 2);
    return mcount * log (mrate / mtime) + (2 - mcount) * lognp + logb - mrate - logs;
  }
}

#define NW_MIGSIMFRAC 0.999
#define NW_LOG2HALF 0.34657359027997265470861606073

double
ora_nw_migweight (const ora_model * m, const double *tvals, int period, double newt, int numgenes, int root,
                  const int *up0, const int *up1, const int *down, const double *time, const int *b_pop,
                  const int *b_mig_off, const double *b_mig_t, const int *b_mig_p, const int *a_pop,
                  const int *a_mig_off, const double *a_mig_t, const int *a_mig_p)
{
  const int ng = numgenes, nl = 2 * ng - 1, p1 = period + 1;
  const double oldt = tvals[period];
  const int up = newt > oldt;
  const double tu = up ? oldt : newt, td = up ? newt : oldt;
  const int period_a = up ? period : period + 1, period_b = up ? period + 1 : period;
  const int addp = m->addpop[p1], d0 = m->droppops[p1][0], d1 = m->droppops[p1][1];
  double tvn[ORA_MAXPERIODS + 1], num = 0, denom = 0;
  int i, k, *setf = (int *) malloc (nl * sizeof (int)), *rdb = (int *) malloc (nl * sizeof (int)), *rda =
    (int *) malloc (nl * sizeof (int)), *first = (int *) calloc (nl, sizeof (int)), *two = (int *) calloc (nl, sizeof (int));
  double *lpf = (double *) calloc (nl, sizeof (double)), *lpfr = (double *) calloc (nl, sizeof (double));
  for (k = 0; k <= m->nsplit; k++)
    tvn[k] = k < m->nsplit ? tvals[k] : ORA_TIMEMAX;
  tvn[period] = newt;
#define UPT(e) ((e) < ng ? 0.0 : time[up0[e]])
#define POPB(e, t) nw_nowedgepop (m, tvals, b_pop[e], b_mig_off, b_mig_t, b_mig_p, e, t)
#define POPA(e, t) nw_nowedgepop (m, tvn, a_pop[e], a_mig_off, a_mig_t, a_mig_p, e, t)
#define ISDROP(p) ((p) == d0 || (p) == d1)
  for (i = 0; i < nl; i++)
    setf[i] = -1;
  /* :355-583, the simulated population below the interval taken from the genealogy after the move */
  for (i = 0; i < nl; i++)
  {
    double uptime = UPT (i);
    if (!(time[i] > tu && uptime <= td) || setf[i] != -1)
      continue;
    int db, da, c0, c1, sis = -1;
    double f = 0, fr = 0;
    if (time[i] > td)
    {
      db = POPB (i, td);
      da = db;
      if (up)
      {
        if (db == addp)
        {
          c0 = POPB (i, tu);
          da = POPA (i, td);
          if (uptime < tu && ISDROP (c0))
            f = da == c0 ? log (NW_MIGSIMFRAC) : log (1.0 - NW_MIGSIMFRAC);
          else
            f = -0.69314718055994530941723212146;   /* LOG2, imamp.hpp:188 */
        }
      }
      else if (ISDROP (db))
      {
        da = addp;
        c0 = POPB (i, tu);
        if (uptime < tu && ISDROP (c0))
          fr = c0 == db ? log (NW_MIGSIMFRAC) : log (1.0 - NW_MIGSIMFRAC);
        else
          fr = -0.69314718055994530941723212146;   /* LOG2, imamp.hpp:188 */
      }
    }
    else
    {
      sis = up0[down[i]] == i ? up1[down[i]] : up0[down[i]];
      double uptime1 = UPT (sis);
      two[i] = 1;
      db = POPB (i, time[i]);
      da = db;
      if (up)
      {
        if (db == addp)
        {
          c0 = POPB (i, tu);
          c1 = POPB (sis, tu);
          da = POPA (i, time[i]);
          if (uptime < tu && uptime1 < tu && c0 == c1 && ISDROP (c0))
            f = (da == c0 ? log (NW_MIGSIMFRAC) : log (1.0 - NW_MIGSIMFRAC)) / 2.0;
          else
            f = -NW_LOG2HALF;
        }
      }
      else if (ISDROP (db))
      {
        c0 = POPB (i, tu);
        c1 = POPB (sis, tu);
        da = addp;
        if (uptime < tu && uptime1 < tu && c0 == c1 && ISDROP (c0))
          fr = (c0 == db ? log (NW_MIGSIMFRAC) : log (1.0 - NW_MIGSIMFRAC)) / 2.0;
        else
          fr = -NW_LOG2HALF;
      }
    }
    setf[i] = i;
    first[i] = 1;
    rdb[i] = db;
    rda[i] = da;
    lpf[i] = f;
    lpfr[i] = fr;
    if (sis >= 0)
    {
      setf[sis] = i;
      rdb[sis] = db;
      rda[sis] = da;
    }
  }
  /* :586-757 */
  for (i = 0; i < nl; i++)
  {
    if (!first[i])
      continue;
    for (k = 0; k <= two[i]; k++)
    {
      int ei = k ? (up0[down[i]] == i ? up1[down[i]] : up0[down[i]]) : i, upb, upa, j, mi, mstart, kk, cm2_b = -1, cm2_a = -1;
      int mcount, mnew, npopsa = m->npops - period_a, npopsb = m->npops - period_b;
      double uptime = UPT (ei), mtime, mrate, mrate_r, top;
      if (uptime < tu)
      {
        if (!up)
        {
          upb = POPB (ei, tu);
          upa = ISDROP (upb) ? addp : upb;
        }
        else
        {
          upb = POPB (ei, tu * (1 + DBL_EPSILON));
          upa = upb == addp ? POPB (ei, tu) : upb;
        }
      }
      else
      {
        upb = rdb[up0[ei]];
        upa = rda[up0[ei]];
      }
      if (ei == root)
        continue;
      top = tu > uptime ? tu : uptime;
      mtime = (td < time[ei] ? td : time[ei]) - top;
      for (j = b_mig_off[ei], kk = 0, mi = 0, mstart = -1; j < b_mig_off[ei + 1] && b_mig_t[j] < td; j++, kk++)
        if (b_mig_t[j] > tu)
        {
          if (mi == 0)
            mstart = kk;
          mi++;
        }
      if (kk >= 2 && mstart >= 0 && kk - mstart >= 2)
        cm2_b = kk == 2 ? upb : b_mig_p[b_mig_off[ei] + kk - 3];
      mcount = mi;
      mrate = period_a < m->nsplit ? ora_calcmrate (mcount, mtime) * mtime : 0;
      /* the number of events simulated, and the population before the second to last of them, from the result */
      for (j = a_mig_off[ei], kk = 0, mnew = 0; j < a_mig_off[ei + 1] && a_mig_t[j] < td; j++, kk++)
        if (a_mig_t[j] > tu)
          mnew++;
      if (mnew >= 2)
        cm2_a = mnew == 2 ? upa : a_mig_p[a_mig_off[ei] + kk - 3];
      mrate_r = period_b < m->nsplit ? ora_calcmrate (mnew, mtime) * mtime : 0;
      if (mrate > 0)
        denom += lpf[setf[ei]] + nw_getmprob (m, period_a, mrate, mtime, mnew, upa, rda[ei], cm2_a, npopsa);
      if (mrate_r > 0)
        num += lpfr[setf[ei]] + nw_getmprob (m, period_b, mrate_r, mtime, mcount, upb, rdb[ei], cm2_b, npopsb);
    }
  }
  free (setf); free (rdb); free (rda); free (first); free (two); free (lpf); free (lpfr);
  return num - denom;
#undef UPT
#undef POPB
#undef POPA
#undef ISDROP
}

/* ---- section 8 (f3): the other evaluators over the sampled-genealogy rows ---------------------------------- */
#define MINPARAMVAL 0.0000001   /* imamp.hpp:130 */
#define SQR(x) ((x)*(x))

/* calcx output.cpp:14-134 */
double
ora_calcx (const ora_model * m, const float *rows, int rowlen, int ei, int pnum, int mode)
{
  int nq = m->nq, nm = m->nm, cc, mc, p = pnum;
  int ccp = 0, fcp = nq, hccp = 2 * nq, mcp = 3 * nq, fmp = mcp + nm, qip = fmp + nm, mip = qip + nq;
  const float *g = rows + (size_t) ei * rowlen;
  double fc, fm, hval, denom, max, tempval;
  if (p < nq)
  {
    if (m->q_max[p] <= MINPARAMVAL)
      return -1;
    cc = (int) g[ccp + p];
    fc = g[fcp + p];
    hval = g[hccp + p];
    denom = g[qip + p];
    max = m->q_max[p];
    if (mode == 0)
    {
      if (cc == 0 && fc == 0)
        tempval = (SQR (max) / 2) / exp (denom);
      else if (cc > 1)
        tempval = exp (2 * LOG2 - hval + (2 - cc) * log (fc) + ora_uppergamma (cc - 2, 2 * fc / max) - denom);
      else if (cc == 1)
        tempval = exp (LOG2 - hval + log (max * exp (-2 * fc / max) - 2 * fc * exp (ora_uppergamma (0, 2 * fc / max))) - denom);
      else
        tempval = exp (log ((max / 2) * (max - 2 * fc) * exp (-2 * fc / max) + 2 * SQR (fc) * exp (ora_uppergamma (0, 2 * fc / max))) - denom);
    }
    else
    {
      if (cc == 0 && fc == 0)
        tempval = (max * SQR (max) / 3) / exp (denom);
      else if (cc > 2)
        tempval = exp (ora_uppergamma (cc - 3, 2 * fc / max) + 3 * LOG2 - hval + (3 - cc) * log (fc) - denom);
      else if (cc == 2)
        tempval = exp (2 * LOG2 - hval + log (max * exp (-2 * fc / max) - 2 * fc * exp (ora_uppergamma (0, 2 * fc / max))) - denom);
      else if (cc == 1)
        tempval = exp (-hval + log (max * (max - 2 * fc) * exp (-2 * fc / max) + 4 * SQR (fc) * exp (ora_uppergamma (0, 2 * fc / max))) - denom);
      else
        tempval = exp (-log (3.0) + log (max * (2 * SQR (fc) - fc * max + SQR (max)) * exp (-2 * fc / max) - 4 * pow ((double) fc, 3.0) * exp (ora_uppergamma (0, 2 * fc / max))) - denom);
    }
  }
  else
  {
    p -= nq;
    if (m->m_max[p] <= MINPARAMVAL)
      return -1;
    mc = (int) g[mcp + p];
    fm = g[fmp + p];
    denom = g[mip + p];
    max = m->m_max[p];
    if (mode == 0)
    {
      if (mc == 0 && fm == 0)
        tempval = (SQR (max) / 2) / exp (denom);
      else if (mc > 0)
        tempval = exp (ora_lowergamma (mc + 2, fm * max) - (mc + 2) * log (fm) - denom);
      else
        tempval = (1 - (1 + fm * max) * exp (-fm * max)) / SQR (fm) / exp (denom);
    }
    else
    {
      if (mc == 0 && fm == 0)
        tempval = (pow (max, 3.0) / 3) / exp (denom);
      else
        tempval = exp (ora_lowergamma (mc + 3, fm * max) - (mc + 3) * log (fm) - denom);
    }
  }
  return tempval;
}

/* the accumulation of print_means_variances_correlations output.cpp:704-728: sums[2 np + np*np] = sum0, sum1, cross (p < q) */
void
ora_moment_sums (const ora_model * m, const float *rows, int rowlen, int nrows, double *sums)
{
  int np = m->nq + m->nm, i, p, q;
  for (i = 0; i < 2 * np + np * np; i++)
    sums[i] = 0;
  for (i = 0; i < nrows; i++)
    for (p = 0; p < np; p++)
    {
      sums[p] += ora_calcx (m, rows, rowlen, i, p, 0);
      sums[np + p] += ora_calcx (m, rows, rowlen, i, p, 1);
    }
  for (i = 0; i < nrows; i++)
    for (p = 0; p < np - 1; p++)
      for (q = p + 1; q < np; q++)
        sums[2 * np + p * np + q] += ora_calcx (m, rows, rowlen, i, p, 0) * ora_calcx (m, rows, rowlen, i, q, 0);
}

/* row sum of calc_popmig / marginpopmig popmig.cpp:26-90, 209-263 (uniform migration prior) over rows [first, last) */
double
ora_popmig_sum (const ora_model * m, const float *rows, int rowlen, int first, int last, int thetai, int mi, double x)
{
  int nq = m->nq, nm = m->nm, ei, cc, mc;
  int ccp = thetai, fcp = nq + thetai, hcp = 2 * nq + thetai, mcp = 3 * nq + mi, fmp = 3 * nq + nm + mi, qip = 3 * nq + 2 * nm + thetai,
    mip = 4 * nq + 2 * nm + mi;
  double sum = 0, temp1, temp2, fc, fm, hc, qintg, mintg, mmax = m->m_max[mi], qmax = m->q_max[thetai], a, b;
  for (ei = first; ei < last; ei++)
  {
    const float *g = rows + (size_t) ei * rowlen;
    cc = (int) g[ccp];
    fc = (double) g[fcp];
    hc = (double) g[hcp];
    mc = (int) g[mcp];
    fm = (double) g[fmp];
    qintg = (double) g[qip];
    mintg = (double) g[mip];
    if (fc == 0 && cc == 0 && fm > 0)
    {
      temp1 = LOG2 - (mc * log (fm)) - hc - qintg - mintg;
      temp2 = log (exp (ora_uppergamma (mc, 2 * fm * x / qmax)) - exp (ora_uppergamma (mc, mmax * fm)));
    }
    else if (fm == 0 && mc == 0 && fc > 0)
    {
      temp1 = LOG2 - (cc * log (fc)) - hc - qintg - mintg;
      temp2 = log (exp (ora_uppergamma (cc, 2 * fc / qmax)) - exp (ora_uppergamma (cc, fc * mmax / x)));
    }
    else if (fc == 0 && cc == 0 && mc == 0 && fm == 0)
    {
      temp1 = log (2 * log (mmax * qmax / (2 * x))) - hc - qintg - mintg;
      temp2 = 0;
    }
    else
    {
      temp1 = LOG2 + (mc * log (x)) - ((cc + mc) * log (fc + fm * x)) - hc - qintg - mintg;
      a = ora_uppergamma (cc + mc, 2 * (fc + fm * x) / qmax);
      b = ora_uppergamma (cc + mc, mmax * (fm + fc / x));
      if (a == b)
      {
        b = ora_lowergamma (cc + mc, 2 * (fc + fm * x) / qmax);
        a = ora_lowergamma (cc + mc, mmax * (fm + fc / x));
      }
      if (a > b)
        logdiff (&temp2, a, b);
      else
        temp1 = temp2 = 0.0;
    }
    if ((temp1 + temp2 < 700) && (temp1 + temp2 > -700))
      sum += exp (temp1 + temp2);
  }
  return sum;
}

/* greater-than probabilities gtint.cpp:26-330.  The integrands read file-static row values there; here they take them in
 * a struct.  qtrap / trapzd (:83-125) keep the running estimate s between refinements exactly as the static does. */
struct gtrow
{
  int cci, ccj, wi, wj;
  double fci, fcj, hval, denom, qmax, fmi, fmj, mmax;
};

static double
gt_mig_integrand (const struct gtrow *r, double mi)     /* mgt_wj_gt_0 :26-47 */
{
  double temp1, temp2, a, b;
  if (mi < MINPARAMVAL)
    return 0.0;
  a = ora_logfact (r->wj);
  b = ora_uppergamma (r->wj + 1, r->fmj * mi);
  if (a <= b)
    return 0.0;
  if ((a - b) < 1e-15)
    temp1 = ora_lowergamma (r->wj + 1, r->fmj * mi);
  else
    logdiff (&temp1, a, b);
  temp2 = r->wi * log (mi) - r->fmi * mi - (r->wj + 1) * log (r->fmj) + temp1;
  temp2 -= r->denom;
  return exp (temp2);
}

static double
gt_pop_integrand (const struct gtrow *r, double qi)     /* pgt_fcj_gt_0 :49-79 */
{
  double fcj2, fci2, temp1, temp2, temp3, a, b;
  if (qi < MINPARAMVAL)
    return 0.0;
  fcj2 = 2 * r->fcj;
  fci2 = 2 * r->fci;
  if (r->ccj == 0)
  {
    a = log (qi) - fcj2 / qi;
    b = log (fcj2) + ora_uppergamma (0, fcj2 / qi);
    if (a > b)
    {
      logdiff (&temp1, a, b);
      temp2 = -fci2 / qi + r->cci * log (2 / qi);
      temp3 = temp1 + temp2 - r->hval - r->denom;
      return exp (temp3);
    }
    return 0.0;
  }
  temp1 = ora_uppergamma (r->ccj - 1, fcj2 / qi);
  temp2 = LOG2 + r->cci * log (2 / qi) + (1 - r->ccj) * log (r->fcj) - fci2 / qi;
  temp3 = temp2 + temp1 - r->hval - r->denom;
  return exp (temp3);
}

static double
gt_qtrap (double (*func) (const struct gtrow *, double), const struct gtrow *r, double a, double b)
{
  int j, k, it;
  double s = 0, olds = -1.0e100, x, tnm, sum, del;
  for (j = 1; j <= 20; j++)
  {
    if (j == 1)
      s = 0.5 * (b - a) * (func (r, a) + func (r, b));
    else
    {
      for (it = 1, k = 1; k < j - 1; k++)
        it <<= 1;
      tnm = it;
      del = (b - a) / tnm;
      x = a + 0.5 * del;
      for (sum = 0.0, k = 1; k <= it; k++, x += del)
        sum += func (r, x);
      s = 0.5 * (s + (b - a) * sum / tnm);
    }
    if (j > 5)
      if (fabs (s - olds) < 1.0e-4 * fabs (olds) || (s == 0.0 && olds == 0.0))
        return s;
    olds = s;
  }
  return s;
}

/* gtpops (kind 0, :248-318) / gtmig (kind 1, :128-245) over the rows print_greater_than_tests uses (:341-351) */
double
ora_greater_than (const ora_model * m, const float *rows, int rowlen, int nrows, int kind, int pi, int pj)
{
  int nq = m->nq, nm = m->nm, ei, treeinc = 1, hitreenum = nrows, numtreesused = nrows;
  int ccp = 0, fcp = nq, hccp = 2 * nq, mcp = 3 * nq, fmp = mcp + nm, qip = fmp + nm, mip = qip + nq;
  double sum = 0, temp, temp1, temp2, temp3, temp4, a, b, c;
  struct gtrow r;
  if (nrows > 20000)
  {
    treeinc = nrows / 20000;
    numtreesused = 20000;
    hitreenum = treeinc * 20000;
  }
  for (ei = 0; ei < hitreenum; ei += treeinc)
  {
    const float *g = rows + (size_t) ei * rowlen;
    if (kind == 0)
    {
      double qmax = r.qmax = m->q_max[pi], fci, hval, denom;
      int cci;
      cci = r.cci = (int) g[ccp + pi];
      r.ccj = (int) g[ccp + pj];
      fci = r.fci = g[fcp + pi];
      r.fcj = g[fcp + pj];
      hval = r.hval = g[hccp + pi] + g[hccp + pj];      /* float sums, as :266-267 */
      denom = r.denom = g[qip + pi] + g[qip + pj];
      if (r.ccj == 0 && r.fcj == 0)
      {
        if (fci == 0)
          temp = exp (2.0 * log (qmax) - LOG2 - hval - denom);
        else if (cci >= 2)
        {
          temp1 = 2 * LOG2 + (2 - cci) * log (fci);
          temp2 = ora_uppergamma (cci - 2, 2 * fci / qmax);
          temp = exp (temp1 + temp2 - hval - denom);
        }
        else if (cci == 1)
        {
          temp1 = 4 * fci * exp (ora_uppergamma (0, 2 * fci / qmax));
          temp2 = 2 * qmax * exp (-2 * fci / qmax) - temp1;
          temp = exp (log (temp2) - hval - denom);
        }
        else
        {
          temp1 = exp (ora_uppergamma (0, 2 * fci / qmax));
          temp2 = (qmax / 2) * (qmax - 2 * fci) * exp (-2 * fci / qmax) + 2 * fci * fci * temp1;
          temp = exp (log (temp2) - hval - denom);
        }
      }
      else
        temp = gt_qtrap (gt_pop_integrand, &r, MINPARAMVAL, qmax);
    }
    else
    {
      double mmax = r.mmax = m->m_max[pi], fmi, fmj, denom;
      int wi;
      fmi = r.fmi = g[fmp + pi];
      fmj = r.fmj = g[fmp + pj];
      wi = r.wi = (int) g[mcp + pi];
      r.wj = (int) g[mcp + pj];
      denom = r.denom = g[mip + pi] + g[mip + pj];      /* float sum, as :144 */
      if (r.wj == 0)
      {
        if (fmj > 0.0)
        {
          if (wi > 0)
          {
            a = ora_logfact (wi);
            b = ora_uppergamma (wi + 1, (fmi + fmj) * mmax);
            c = ora_uppergamma (wi + 1, fmi * mmax);
            if (a <= b || a <= c)
              temp = 0.0;
            else
            {
              if ((a - b) < 1e-15)
                temp1 = ora_lowergamma (wi + 1, (fmi + fmj) * mmax);
              else
                logdiff (&temp1, a, b);
              temp1 += -(wi + 1) * log (fmi + fmj);
              if ((a - c) < 1e-15)
                temp2 = ora_lowergamma (wi + 1, fmi * mmax);
              else
                logdiff (&temp2, a, c);
              temp2 += -(wi + 1) * log (fmi);
              if (temp2 <= temp1)
                temp = 0.0;
              else
              {
                logdiff (&temp3, temp2, temp1);
                temp4 = temp3 - log (fmj) - denom;
                temp = exp (temp4);
              }
            }
          }
          else if (fmi > 0.0)
          {
            temp1 = fmi * (exp (-(fmi + fmj) * mmax) - exp (-fmi * mmax));
            temp2 = fmj * (1 - exp (-fmi * mmax));
            temp3 = (temp1 + temp2) / (fmi * fmj * (fmi + fmj));
            temp = exp (log (temp3) - denom);
          }
          else
          {
            temp1 = (fmj * mmax - 1.0 + exp (-fmj * mmax)) / (fmj * fmj);
            temp = exp (log (temp1) - denom);
          }
        }
        else if (wi > 0)
        {
          a = ora_logfact (wi + 1);
          b = ora_uppergamma (wi + 2, fmi * mmax);
          if (a <= b)
            temp = 0.0;
          else
          {
            if ((a - b) < 1e-15)
              temp1 = ora_lowergamma (wi + 2, fmi * mmax);
            else
              logdiff (&temp1, a, b);
            temp2 = -(wi + 2) * log (fmi);
            temp = exp (temp2 + temp1 - denom);
          }
        }
        else if (fmi > 0.0)
        {
          temp1 = (1.0 - exp (-fmi * mmax) * (fmi * mmax + 1.0)) / (fmi * fmi);
          temp = exp (log (temp1) - denom);
        }
        else
          temp = exp (log (mmax * mmax / 2.0) - denom);
      }
      else
        temp = gt_qtrap (gt_mig_integrand, &r, MINPARAMVAL, mmax);
    }
    if (temp > 1.0)
      temp = 1.0;
    sum += temp;
  }
  return sum / numtreesused;
}
