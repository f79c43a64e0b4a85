// Device numerics of the IMa2p hot path.
//
// "Within 1e-9 of the reference" means reproducing the reference's own recurrences and stopping rules
// (SURVEY.md fact 3): incomplete-gamma series/continued fractions that stop at EPS 3e-7
// (utilities.cpp:808-1122), Abramowitz-Stegun Bessel polynomials (utilities.cpp:54-124,1452-1491), the
// 10-term eexp polynomial (utilities.cpp:1501-1539) and the LogDiff switches of the integrated prior
// (update_gtree_common.cpp:108-304).  This translation unit is compiled with --fmad=false so that
// no multiply-add is contracted (the reference is an FMA-free x86-64 build).
#pragma once
#include "ima_platform.h"

namespace ima {

struct MathCtx {
  const double *logfact;   // logfact[i] = sum_{j<=i} log j, same running sum as setlogfact (utilities.cpp:1405-1414)
  int logfact_n;
  int *err;                // device error word (first error wins); 0 = none
};

enum DevErr { kErrNone = 0, kErrLogDiff = 14, kErrGamma = 15, kErrRange = 16, kErrExchange = 17 };

IMA_DEV void raise(const MathCtx &mc, int code) {
#if IMA_CUDA
  atomicCAS(mc.err, 0, code);
#else
  if (*mc.err == 0) *mc.err = code;
#endif
}

IMA_DEV double lfact(const MathCtx &mc, int n) {
  if (n < 0 || n >= mc.logfact_n) { raise(mc, kErrRange); return 0.0; }
  return mc.logfact[n];
}

constexpr int kItMax = 1000;       // ITMAX
constexpr double kEps = 3.0e-7;    // EPS
constexpr double kFpMin = 1.0e-30; // FPMIN

// Lentz continued fraction shared by gcf / gcflog (utilities.cpp:811-877); returns h
IMA_DEV double gamma_cf(const MathCtx &mc, double a, double x) {
  double b = x + 1.0 - a, c = 1.0 / kFpMin, d = 1.0 / b, h = d;
  int i;
  for (i = 1; i <= kItMax; i++) {
    double an = -i * (i - a);
    b += 2.0;
    d = an * d + b;
    if (fabs(d) < kFpMin) d = kFpMin;
    c = b + an / c;
    if (fabs(c) < kFpMin) c = kFpMin;
    d = 1.0 / d;
    double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < kEps) break;
  }
  if (i > kItMax) raise(mc, kErrGamma);
  return h;
}

// series shared by gser / gserlog (utilities.cpp:882-953); returns sum (0 when x <= 0)
IMA_DEV double gamma_series(const MathCtx &mc, int a, double x) {
  if (x <= 0.0) { if (x < 0.0) raise(mc, kErrGamma); return 0.0; }
  double ap = a, del = 1.0 / a, sum = del;
  for (int n = 1; n <= kItMax; n++) {
    ap += 1.0;
    del *= x / ap;
    sum += del;
    if (fabs(del) < fabs(sum) * kEps) return sum;
  }
  raise(mc, kErrGamma);
  return 0.0;
}

// expint(1, x) as used by uppergamma(0, x) (utilities.cpp:962-1045, n == 1); result in log form
IMA_DEV double log_expint1(const MathCtx &mc, double x) {
  const double euler = 0.5772156649;
  if (x <= 0.0) { raise(mc, kErrGamma); return 0.0; }
  if (x > 1.0) {
    double b = x + 1.0, c = 1.0 / kFpMin, d = 1.0 / b, h = d;
    for (int i = 1; i <= 100; i++) {
      double a = -(double)i * (double)i;       // -i*(nm1+i), nm1 = 0
      b += 2.0;
      d = 1.0 / (a * d + b);
      c = b + a / c;
      double del = c * d;
      h *= del;
      if (fabs(del - 1.0) < kEps) return log(h) - x;
    }
    raise(mc, kErrGamma);
    return 0.0;
  }
  double ans = -log(x) - euler, fact = 1.0;
  for (int i = 1; i <= 100; i++) {
    fact *= -x / i;
    double del = -fact / i;                    // i != nm1 (= 0) always
    ans += del;
    if (fabs(del) < fabs(ans) * kEps) return log(ans);
  }
  raise(mc, kErrGamma);
  return 0.0;
}

// uppergamma utilities.cpp:1053-1090: log Gamma(a, x), integer a >= 0
IMA_DEV double uppergamma(const MathCtx &mc, int a, double x) {
  double p;
  if (x < 0.0 || a < 0) { raise(mc, kErrGamma); return 0.0; }
  if (a == 0) {
    p = log_expint1(mc, x);
  } else {
    double gln = lfact(mc, a - 1);
    if (x < a + 1.0) {
      double s = gamma_series(mc, a, x);
      double gamser = (x <= 0.0) ? 0.0 : s * exp(-x + a * log(x) - gln);
      p = gln + log(1.0 - gamser);
    } else {
      double h = gamma_cf(mc, (double)a, x);
      p = gln + ((-x + a * log(x) - gln) + log(h));
    }
  }
  if (p < -1e200) p = -1e200;
  return p;
}

// lowergamma utilities.cpp:1092-1122: log gamma(a, x), integer a >= 1
IMA_DEV double lowergamma(const MathCtx &mc, int a, double x) {
  double p;
  if (x < 0.0 || a <= 0) { raise(mc, kErrGamma); return 0.0; }
  double gln = lfact(mc, a - 1);
  if (x < a + 1.0) {
    double s = gamma_series(mc, a, x);
    double gamserlog = (x <= 0.0) ? 0.0 : log(s) + (-x + a * log(x) - gln);
    p = gln + gamserlog;
  } else {
    double h = gamma_cf(mc, (double)a, x);
    double gammcf = exp(-x + a * log(x) - gln) * h;
    p = gln + log(1 - gammcf);
  }
  if (p < -1e200) p = -1e200;
  return p;
}

// LogDiff imamp.hpp:257-263; a <= b is fatal in the reference -> error word + huge negative value
IMA_DEV bool logdiff(const MathCtx &mc, double &v, double a, double b) {
  if (a <= b) { raise(mc, kErrLogDiff); v = -kMyDblMax; return false; }
  v = (a - b < kLogDblMax) ? b + log(exp(a - b) - 1.0) : a;
  return true;
}

// integrate_coalescent_term update_gtree_common.cpp:108-203 (MORESTABLE)
IMA_DEV double integrate_coalescent_term(const MathCtx &mc, int cc, double fc, double hcc, double max, double min) {
  double p, a, b, c, d;
  if (cc > 0) {
    if (min == 0) {
      double ug = uppergamma(mc, cc - 1, 2 * fc / max);
      if (cc > 1) {
        double fullg = lfact(mc, cc - 2);
        if (fullg - ug < 1e-15 || fullg - ug > kLogDblMax) {
          double lg = lowergamma(mc, cc - 1, 2 * fc / max);
          if (fullg > lg) {
            double ugalt;
            logdiff(mc, ugalt, fullg, lg);
            if (fabs(ugalt - ug) > 1e-10) ug = ugalt;
          }
        }
      }
      p = ug + kLog2 - hcc + (1 - cc) * log(fc);
    } else {
      a = uppergamma(mc, cc - 1, 2 * fc / max);
      b = uppergamma(mc, cc - 1, 2 * fc / min);
      if (!logdiff(mc, p, a, b)) return p;
      p += (kLog2 - hcc + (1 - cc) * log(fc));
    }
  } else if (2 * fc / max > 0) {
    if (min == 0) {
      a = log(max) - 2.0 * fc / max;
      b = kLog2 + log(fc) + uppergamma(mc, 0, 2.0 * fc / max);
      logdiff(mc, p, a, b);
    } else {
      a = uppergamma(mc, 0, 2 * fc / max);
      b = uppergamma(mc, 0, 2 * fc / min);
      if (!logdiff(mc, c, a, b)) return c;
      c += kLog2 + log(fc);
      a = log(max) - 2.0 * fc / max;
      b = log(min) - 2.0 * fc / min;
      if (!logdiff(mc, d, a, b)) return d;
      logdiff(mc, p, d, c);
    }
  } else {
    p = log(max - min);
  }
  return p;
}

// integrate_migration_term update_gtree_common.cpp:205-296 (MORESTABLE)
IMA_DEV double integrate_migration_term(const MathCtx &mc, int cm, double fm, double max, double min) {
  double p, a, b, c;
  if (cm > 0) {
    if (min == 0) {
      double lg = lowergamma(mc, cm + 1, fm * max);
      double fullg = lfact(mc, cm);
      if (fullg - lg < 1e-15 || fullg - lg > kLogDblMax) {
        double ug = uppergamma(mc, cm + 1, fm * max);
        if (fullg > ug) {
          double lgalt;
          logdiff(mc, lgalt, fullg, ug);
          if (fabs(lgalt - lg) > 1e-12) lg = lgalt;
        }
      }
      p = (-1 - cm) * log(fm) + lg;
    } else {
      a = uppergamma(mc, cm + 1, fm * min);
      b = uppergamma(mc, cm + 1, fm * max);
      if (!logdiff(mc, c, a, b)) return c;
      p = (-1 - cm) * log(fm) + c;
    }
  } else if (fm > kMPriorMin) {
    if (min == 0) {
      if (max == kMPriorMin) {
        p = 0;
      } else {
        if (!logdiff(mc, c, 0.0, -fm * max)) return c;
        p = c - log(fm);
      }
    } else {
      if (!logdiff(mc, c, -fm * min, -fm * max)) return c;
      p = c - log(fm);
    }
  } else {
    p = log(max - min);
  }
  return p;
}

// integrate_migration_term_expo_prior update_gtree_common.cpp:298-304
IMA_DEV double integrate_migration_term_expo(const MathCtx &mc, int cm, double fm, double exmean) {
  return -log(exmean) + (-(cm + 1) * log(fm + 1.0 / exmean)) + lfact(mc, cm);
}

// ------------------------------------------------------------------------------------------------
// Warp-cooperative forms: all lanes of the warp call them with the SAME arguments and all receive the
// same result.  The series and the continued fraction, which dominate the reference's cost, are evaluated
// 32 terms at a time: term ratios in parallel, running products / sums by warp scans, and the reference's
// stopping rule applied to every term so that the series stops at the same index.  (Scan products associate
// differently from the sequential recurrences: differences are a few ulp, far inside the 1e-9 parity bar.)
// ------------------------------------------------------------------------------------------------
IMA_DEV double gamma_series_coop(const MathCtx &mc, int a, double x) {
  if (x <= 0.0) { if (x < 0.0) raise(mc, kErrGamma); return 0.0; }
  const int lane = Warp::lane();
  double del0 = 1.0 / a, sum0 = del0;
  for (int n0 = 0; n0 < kItMax; n0 += IMA_WARP) {
    const int n = n0 + 1 + lane;
    const double del = del0 * Warp::scan_mul(x / ((double)a + n));
    const double sum = sum0 + Warp::scan_add(del);
    const int first = Warp::first(n <= kItMax && fabs(del) < fabs(sum) * kEps);
    if (first >= 0) return Warp::bcast(sum, first);
    del0 = Warp::bcast(del, IMA_WARP - 1);
    sum0 = Warp::bcast(sum, IMA_WARP - 1);
  }
  raise(mc, kErrGamma);
  return 0.0;
}

// 2x2 matrices for the Lentz recurrences d_i = 1/(an_i d_{i-1} + b_i), c_i = b_i + an_i/c_{i-1}
// (utilities.cpp:822-836): with d = N/D and c = U/V they are linear, so 32 steps compose by a matrix scan.
struct Mat2 { double a, b, c, d; };
IMA_DEV Mat2 mat2_mul(const Mat2 &L, const Mat2 &R) {   // L * R
  Mat2 o;
  o.a = L.a * R.a + L.b * R.c; o.b = L.a * R.b + L.b * R.d;
  o.c = L.c * R.a + L.d * R.c; o.d = L.c * R.b + L.d * R.d;
  return o;
}
IMA_DEV Mat2 mat2_scan(Mat2 m) {   // inclusive: lane k gets m_k * m_{k-1} * ... * m_0
#if IMA_CUDA
  for (int o = 1; o < 32; o <<= 1) {
    Mat2 t;
    t.a = Warp::shfl_up(m.a, o); t.b = Warp::shfl_up(m.b, o); t.c = Warp::shfl_up(m.c, o); t.d = Warp::shfl_up(m.d, o);
    if (Warp::lane() >= o) m = mat2_mul(m, t);
  }
#endif
  return m;
}

IMA_DEV double gamma_cf_coop(const MathCtx &mc, double a, double x) {
  const int lane = Warp::lane();
  const double b0 = x + 1.0 - a;
  double dprev = 1.0 / b0, cprev = 1.0 / kFpMin, h0 = dprev;
  for (int i0 = 1; i0 <= kItMax; i0 += IMA_WARP) {
    const int i = i0 + lane;
    const double an = -i * (i - a), b = b0 + 2.0 * i;
    // c-recurrence matrix C_i = [[b, an], [1, 0]] acting on (U, V); the d-recurrence matrix is J C_i J (J swaps the two
    // coordinates), so the product of the d-matrices is J (product of the C_i) J: one matrix scan serves both
    Mat2 pc;
    pc.a = b; pc.b = an; pc.c = 1.0; pc.d = 0.0;
    pc = mat2_scan(pc);
    const double U = pc.a * cprev + pc.b, V = pc.c * cprev + pc.d;       // applied to (cprev, 1)
    const double N = pc.d * dprev + pc.c, D = pc.b * dprev + pc.a;       // J P J applied to (dprev, 1)
    // the reference clamps |an d + b| and |c| at FPMIN; if that would ever trigger, redo the call sequentially
    const bool guard = (fabs(D) < kFpMin * fabs(N)) || (fabs(U) < kFpMin * fabs(V)) || !(fabs(D) < DBL_MAX) || !(fabs(U) < DBL_MAX);
    if (Warp::any(guard)) return gamma_cf(mc, a, x);
    const double d = N / D, c = U / V;
    const double del = d * c;
    const double h = h0 * Warp::scan_mul(del);
    const int first = Warp::first(i <= kItMax && fabs(del - 1.0) < kEps);
    if (first >= 0) return Warp::bcast(h, first);
    dprev = Warp::bcast(d, IMA_WARP - 1);
    cprev = Warp::bcast(c, IMA_WARP - 1);
    h0 = Warp::bcast(h, IMA_WARP - 1);
  }
  raise(mc, kErrGamma);
  return h0;
}

// The same continued fraction by its convergents A_i / B_i (A_i = b_i A_{i-1} + an_i A_{i-2}, the same for B): Lentz's
// c_i = A_i / A_{i-1}, d_i = B_{i-1} / B_i and h_i = A_i / B_i, so the walk needs no division until it stops, and the stopping
// rule |c d - 1| < EPS is |A_i B_{i-1} - A_{i-1} B_i| < EPS |A_{i-1} B_i|.  Where the integrated prior calls it (x well above a:
// the all-locus sums) the fraction stops after 3 to 6 terms, which one lane's dependent chain reaches sooner than a 32-wide
// matrix scan does; a call that is still running after kCfSeq terms is handed to the scan, one where the reference's FPMIN
// clamps could matter to the reference's own walk.
constexpr int kCfSeq = 40;
IMA_DEV double gamma_cf_fast(const MathCtx &mc, double a, double x) {
  double Am = kFpMin, A = 1.0, Bm = 1.0, B = x + 1.0 - a, b = B;      // c_0 = 1 / FPMIN, d_0 = h_0 = 1 / b_0
  for (int i = 1; i <= kCfSeq; i++) {
    const double an = -i * (i - a);
    b += 2.0;
    const double An = b * A + an * Am, Bn = b * B + an * Bm;
    if (fabs(Bn) < kFpMin * fabs(B) || fabs(An) < kFpMin * fabs(A) || !(fabs(An) < 1e250) || !(fabs(Bn) < 1e250)) return gamma_cf(mc, a, x);
    const double num = An * B, den = A * Bn;
    Am = A; A = An; Bm = B; B = Bn;
    if (fabs(num - den) < kEps * fabs(den)) return A / B;
  }
  return gamma_cf_coop(mc, a, x);
}

IMA_DEV double uppergamma_coop(const MathCtx &mc, int a, double x) {
  double p;
  if (x < 0.0 || a < 0) { raise(mc, kErrGamma); return 0.0; }
  if (a == 0) {
    p = log_expint1(mc, x);
  } else {
    const double gln = lfact(mc, a - 1);
    if (x < a + 1.0) {
      const double s = gamma_series_coop(mc, a, x);
      const double gamser = (x <= 0.0) ? 0.0 : s * exp(-x + a * log(x) - gln);
      p = gln + log(1.0 - gamser);
    } else {
      const double h = gamma_cf_fast(mc, (double)a, x);
      p = gln + ((-x + a * log(x) - gln) + log(h));
    }
  }
  if (p < -1e200) p = -1e200;
  return p;
}

IMA_DEV double lowergamma_coop(const MathCtx &mc, int a, double x) {
  double p;
  if (x < 0.0 || a <= 0) { raise(mc, kErrGamma); return 0.0; }
  const double gln = lfact(mc, a - 1);
  if (x < a + 1.0) {
    const double s = gamma_series_coop(mc, a, x);
    const double gamserlog = (x <= 0.0) ? 0.0 : log(s) + (-x + a * log(x) - gln);
    p = gln + gamserlog;
  } else {
    const double h = gamma_cf_fast(mc, (double)a, x);
    const double gammcf = exp(-x + a * log(x) - gln) * h;
    p = gln + log(1 - gammcf);
  }
  if (p < -1e200) p = -1e200;
  return p;
}

// The reference falls back from one incomplete gamma to the complementary one at the SAME (a, x) when the first is
// indistinguishable from the complete gamma (update_gtree_common.cpp:128-141, :222-237).  Both evaluate the same
// series / continued fraction and the same exponent, so the cooperative forms compute them once: GammaCore holds the
// shared pieces, upper_of / lower_of finish either function with the arithmetic of uppergamma / lowergamma above.
struct GammaCore { double gln, t, v; bool series; };    // v = series sum (x < a+1) or continued fraction h
IMA_DEV GammaCore gamma_core_coop(const MathCtx &mc, int a, double x) {      // a >= 1, x >= 0
  GammaCore g;
  g.gln = lfact(mc, a - 1);
  g.series = x < a + 1.0;
  g.v = g.series ? gamma_series_coop(mc, a, x) : gamma_cf_fast(mc, (double)a, x);
  g.t = -x + a * log(x) - g.gln;
  return g;
}
IMA_DEV double upper_of(const GammaCore &g, double x) {
  double p;
  if (g.series) {
    const double gamser = (x <= 0.0) ? 0.0 : g.v * exp(g.t);
    p = g.gln + log(1.0 - gamser);
  } else {
    p = g.gln + (g.t + log(g.v));
  }
  if (p < -1e200) p = -1e200;
  return p;
}
IMA_DEV double lower_of(const GammaCore &g, double x) {
  double p;
  if (g.series) {
    const double gamserlog = (x <= 0.0) ? 0.0 : log(g.v) + g.t;
    p = g.gln + gamserlog;
  } else {
    const double gammcf = exp(g.t) * g.v;
    p = g.gln + log(1 - gammcf);
  }
  if (p < -1e200) p = -1e200;
  return p;
}

// Several independent logarithms / exponentials in the time of one: lane k evaluates the k-th argument, the results are
// broadcast.  (The value of log(x) does not depend on the lane that computes it, so nothing changes numerically.)
IMA_DEV void log3_coop(double a, double b, double c, double &la, double &lb, double &lc) {
#if IMA_CUDA
  const int lane = Warp::lane();
  const double r = log(lane == 1 ? b : (lane == 2 ? c : a));
  la = Warp::bcast(r, 0); lb = Warp::bcast(r, 1); lc = Warp::bcast(r, 2);
#else
  la = log(a); lb = log(b); lc = log(c);
#endif
}
IMA_DEV void log2_coop(double a, double b, double &la, double &lb) {
#if IMA_CUDA
  const double r = log(Warp::lane() == 1 ? b : a);
  la = Warp::bcast(r, 0); lb = Warp::bcast(r, 1);
#else
  la = log(a); lb = log(b);
#endif
}
IMA_DEV void exp2_coop(double a, double b, double &ea, double &eb) {
#if IMA_CUDA
  const double r = exp(Warp::lane() == 1 ? b : a);
  ea = Warp::bcast(r, 0); eb = Warp::bcast(r, 1);
#else
  ea = exp(a); eb = exp(b);
#endif
}

// The incomplete gamma of the integrated prior with the reference's fallback to the complementary function
// (update_gtree_common.cpp:128-141 for the upper one of a coalescent term, :222-237 for the lower one of a migration term), a >= 1,
// x > 0, and log(extra) on the side.  Same series / continued fraction, same arithmetic per value as upper_of / lower_of /
// logdiff; the dependent chain is three transcendental evaluations long instead of seven: {log x, log v, log extra}, then
// {exp t, exp(full - direct)}, then {log(1 - v exp t), log(exp(full - direct) - 1)}, where `direct` is whichever of the two
// functions needs no exponential (the lower one from the series, the upper one from the continued fraction) -- the second
// member of each pair is what the fallback needs, computed whether or not the fallback is taken.
IMA_DEV double gamma_with_fallback_coop(const MathCtx &mc, int a, double x, bool want_upper, double tol, double extra, double &lextra) {
  const double gln = lfact(mc, a - 1);
  const bool series = x < a + 1.0;
  const double v = series ? gamma_series_coop(mc, a, x) : gamma_cf_fast(mc, (double)a, x);
  double lx, lv;
  log3_coop(x, v, extra, lx, lv, lextra);
  const double t = -x + a * lx - gln;
  double direct = gln + (lv + t);                         // lower (series) / upper (continued fraction)
  if (direct < -1e200) direct = -1e200;
  const double fullg = gln;
  const bool primary_is_indirect = (want_upper == series);
  // the wanted function is the direct one and it is neither indistinguishable from the complete gamma nor out of range (the
  // upper function of a coalescent term with x above a, the usual case of the all-locus sums): no exponential is needed at all
  if (!primary_is_indirect && !(fullg - direct < 1e-15 || fullg - direct > kLogDblMax)) return direct;
  double e1, e2, l1, l2;
  exp2_coop(t, fullg - direct, e1, e2);
  log2_coop(1.0 - v * e1, e2 - 1.0, l1, l2);
  double indirect = gln + l1;                             // upper (series) / lower (continued fraction)
  if (indirect < -1e200) indirect = -1e200;
  double p = primary_is_indirect ? indirect : direct;
  if (fullg - p < 1e-15 || fullg - p > kLogDblMax) {
    const double other = primary_is_indirect ? direct : indirect;
    if (fullg > other) {
      double alt;
      if (primary_is_indirect) alt = (fullg - other < kLogDblMax) ? other + l2 : fullg;     // LogDiff(fullg, other) from the pieces at hand
      else logdiff(mc, alt, fullg, other);
      if (fabs(alt - p) > tol) p = alt;
    }
  }
  return p;
}

// integrate_coalescent_term / integrate_migration_term with the cooperative gamma functions
IMA_DEV double integrate_coalescent_term_coop(const MathCtx &mc, int cc, double fc, double hcc, double max, double min) {
  double p, a, b, c, d;
  if (cc > 0) {
    if (min == 0) {
      const double x = 2 * fc / max;
      double ug;
      if (cc > 1 && x > 0.0) {
        double lfc;
        ug = gamma_with_fallback_coop(mc, cc - 1, x, true, 1e-10, fc, lfc);
        return ug + kLog2 - hcc + (1 - cc) * lfc;
      } else if (cc > 1 && x >= 0.0) {
        const GammaCore g = gamma_core_coop(mc, cc - 1, x);
        ug = upper_of(g, x);
        const double fullg = g.gln;                        // lfact(cc - 2)
        if (fullg - ug < 1e-15 || fullg - ug > kLogDblMax) {
          const double lg = lower_of(g, x);
          if (fullg > lg) {
            double ugalt;
            logdiff(mc, ugalt, fullg, lg);
            if (fabs(ugalt - ug) > 1e-10) ug = ugalt;
          }
        }
      } else {
        ug = uppergamma_coop(mc, cc - 1, x);
      }
      p = ug + kLog2 - hcc + (1 - cc) * log(fc);
    } else {
      a = uppergamma_coop(mc, cc - 1, 2 * fc / max);
      b = uppergamma_coop(mc, cc - 1, 2 * fc / min);
      if (!logdiff(mc, p, a, b)) return p;
      p += (kLog2 - hcc + (1 - cc) * log(fc));
    }
  } else if (2 * fc / max > 0) {
    if (min == 0) {
      a = log(max) - 2.0 * fc / max;
      b = kLog2 + log(fc) + uppergamma_coop(mc, 0, 2.0 * fc / max);
      logdiff(mc, p, a, b);
    } else {
      a = uppergamma_coop(mc, 0, 2 * fc / max);
      b = uppergamma_coop(mc, 0, 2 * fc / min);
      if (!logdiff(mc, c, a, b)) return c;
      c += kLog2 + log(fc);
      a = log(max) - 2.0 * fc / max;
      b = log(min) - 2.0 * fc / min;
      if (!logdiff(mc, d, a, b)) return d;
      logdiff(mc, p, d, c);
    }
  } else {
    p = log(max - min);
  }
  return p;
}

IMA_DEV double integrate_migration_term_coop(const MathCtx &mc, int cm, double fm, double max, double min) {
  double p, a, b, c;
  if (cm > 0) {
    if (min == 0) {
      const double x = fm * max;
      if (x < 0.0) { raise(mc, kErrGamma); return 0.0; }
      if (x > 0.0) {
        double lfm;
        const double lgf = gamma_with_fallback_coop(mc, cm + 1, x, false, 1e-12, fm, lfm);
        return (-1 - cm) * lfm + lgf;
      }
      const GammaCore g = gamma_core_coop(mc, cm + 1, x);
      double lg = lower_of(g, x);
      const double fullg = g.gln;                          // lfact(cm)
      if (fullg - lg < 1e-15 || fullg - lg > kLogDblMax) {
        const double ug = upper_of(g, x);
        if (fullg > ug) {
          double lgalt;
          logdiff(mc, lgalt, fullg, ug);
          if (fabs(lgalt - lg) > 1e-12) lg = lgalt;
        }
      }
      p = (-1 - cm) * log(fm) + lg;
    } else {
      a = uppergamma_coop(mc, cm + 1, fm * min);
      b = uppergamma_coop(mc, cm + 1, fm * max);
      if (!logdiff(mc, c, a, b)) return c;
      p = (-1 - cm) * log(fm) + c;
    }
  } else if (fm > kMPriorMin) {
    if (min == 0) {
      if (max == kMPriorMin) {
        p = 0;
      } else {
        if (!logdiff(mc, c, 0.0, -fm * max)) return c;
        p = c - log(fm);
      }
    } else {
      if (!logdiff(mc, c, -fm * min, -fm * max)) return c;
      p = c - log(fm);
    }
  } else {
    p = log(max - min);
  }
  return p;
}

// utilities.cpp:239-255
IMA_DEV double mylogcosh(double x) { return x < 100 ? log(cosh(x)) : x - kLog2; }
IMA_DEV double mylogsinh(double x) { return x < 100 ? log(sinh(x)) : x - kLog2; }

// calcmrate update_gtree_common.cpp:462-482
IMA_DEV double calcmrate(int mcnt, double mt) {
  if (mt <= 0.0) return 1.0;
  if (mcnt == 0) return mt < 1 ? 0.1 : 0.1 / mt;
  return mt < 1 ? (double)mcnt : ((double)mcnt) / mt;
}

// bessi0 / bessi1 / bessi: utilities.cpp:54-124, 1452-1491
IMA_DEV double bessi0(double x) {
  double ax = fabs(x), y;
  if (ax < 3.75) {
    y = x / 3.75; y *= y;
    return 1.0 + y * (3.5156229 + y * (3.0899424 + y * (1.2067492 + y * (0.2659732 + y * (0.360768e-1 + y * 0.45813e-2)))));
  }
  y = 3.75 / ax;
  return (exp(ax) / sqrt(ax)) * (0.39894228 + y * (0.1328592e-1 + y * (0.225319e-2 + y * (-0.157565e-2 + y * (0.916281e-2 +
         y * (-0.2057706e-1 + y * (0.2635537e-1 + y * (-0.1647633e-1 + y * 0.392377e-2))))))));
}
IMA_DEV double bessi1(double x) {
  double ax = fabs(x), y, ans;
  if (ax < 3.75) {
    y = x / 3.75; y *= y;
    ans = ax * (0.5 + y * (0.87890594 + y * (0.51498869 + y * (0.15084934 + y * (0.2658733e-1 + y * (0.301532e-2 + y * 0.32411e-3))))));
  } else {
    y = 3.75 / ax;
    ans = 0.2282967e-1 + y * (-0.2895312e-1 + y * (0.1787654e-1 - y * 0.420059e-2));
    ans = 0.39894228 + y * (-0.3988024e-1 + y * (-0.362018e-2 + y * (0.163801e-2 + y * (-0.1031555e-1 + y * ans))));
    ans *= (exp(ax) / sqrt(ax));
  }
  return x < 0.0 ? -ans : ans;
}
IMA_DEV double bessi(int n, double x) {
  n = n < 0 ? -n : n;
  if (x > 700) return kMyDblMax;
  if (n == 0) return bessi0(x);
  if (n == 1) return bessi1(x);
  if (x == 0.0) return 0.0;
  double tox = 2.0 / fabs(x), bip = 0.0, ans = 0.0, bi = 1.0;
  for (int j = 2 * (n + (int)sqrt(40.0 * n)); j > 0; j--) {
    double bim = bip + j * tox * bi;
    bip = bi;
    bi = bim;
    if (fabs(bi) > 1.0e10) { ans *= 1.0e-10; bi *= 1.0e-10; bip *= 1.0e-10; }
    if (j == n) ans = bip;
  }
  ans *= bessi0(x) / bi;
  return (x < 0.0 && (n & 1)) ? -ans : ans;
}

// eexp utilities.cpp:1501-1539: exp(x) = m * 10^z with 1 <= |m| < 10
IMA_DEV void eexp(double x, double &m, int &z) {
  int n = (int)floor(x / kLog2);
  double zr = 0.30102999566398119521 * (double)n;
  z = (int)zr;
  zr -= (double)z;
  double u = x - (((double)n) * kLog2);
  double t = 1 + u * (1.0 + u * (0.5 + u * (0.16666666666666666666666666667 + u * (0.04166666666666666666666666667 +
             u * (0.00833333333333333333333333333 + u * (0.001388888888888888888888888889 +
             u * (0.000198412698412698412698412698 + u * (0.000024801587301587301587301587301 +
             u * (2.75573192239858906525573192239859e-6 + u * (2.75573192239858906525573192239e-7))))))))));
  m = t * pow(10.0, zr);
  if (fabs(m) > 10) { m = m / 10.0; z = z + 1; }
  if (fabs(m) < 1) { m = m * 10.0; z = z - 1; }
}

}  // namespace ima
