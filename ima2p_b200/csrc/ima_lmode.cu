// L mode: densities over every sampled genealogy (.ti rows, ginfo.cpp:288-304).
//
// The reference walks float **gsampinf row by row for every evaluation point
// (surface_call_functions.cpp:25-173, jointfind.cpp:885-1047).  Here the rows are transposed once into a
// column-major (SoA) table in HBM so that a warp reads 32 consecutive genealogies of one column in one
// coalesced request, and many evaluation points are processed per pass over the rows:
//
//   k_marginal      one block = a chunk of rows x a tile of kXT evaluation points held in registers;
//                   per-block partial sums, reduced in block order by k_reduce_partials (deterministic)
//   k_joint_terms   p_g for a batch of parameter vectors, written to a [nvec][G] buffer + per-chunk maxima
//   k_joint_scan    jointp's keep-set in its observable form: a term is inserted iff it lies within
//                   PRANGELOG = 10 of the running maximum of the rows before it (jointfind.cpp:1005); the sum
//                   takes the inserted terms within 10 of the final maximum, minus the smallest one when every
//                   inserted term qualified (loop bound gi < iin, :1011-1022); terms go through eexp (:1014)
#include "ima_devapi.h"
#include "ima_math.h"
#include "../../include/ima2p_b200.h"
#include <string>
#include <vector>
#include <new>

namespace ima {

#if !IMA_CUDA
extern thread_local EmuCtx g_emu;
#endif

constexpr int kXT = 8;               // evaluation points per thread
constexpr int kLmWarps = 8;          // warps per block
constexpr int kRowsPerBlock = 4096;  // rows per chunk
constexpr int kJointVecMax = 32;     // parameter vectors per batch

struct LmView {
  const float *cols;     // [rowlen][G] column-major
  long long G, G_total;
  int rowlen, nq, nm, nsplit, expoprior;
  int ccp, fcp, hccp, mcp, fmp, qip, mip, pdgp, probgp;
  double m_meaninv[kMaxParams];
};

IMA_DEV double integerround(double x) { return x >= 0 ? (double)(long)(x + 0.5) : (double)(long)(x - 0.5); }   // imamp.hpp:180

// margincalc / marginp term (surface_call_functions.cpp:50-73, 139-162)
IMA_KERNEL void k_marginal(LmView V, int param, const double *x, int nx, long long first, long long last, int round_counts, double *partials) {
  IMA_SMEM_DECL
  const int lane = Warp::lane(), warp = ima_warp_in_block();
  const int nxt = (nx + kXT - 1) / kXT;
  const int chunk = ima_block() / nxt, xt = ima_block() - chunk * nxt;
  const long long r0 = first + (long long)chunk * kRowsPerBlock;
  long long r1 = r0 + kRowsPerBlock;
  if (r1 > last) r1 = last;
  double xs[kXT], a1[kXT], a2[kXT], acc[kXT];
  const bool theta = param < V.nq;
  const int p = theta ? param : param - V.nq;
  for (int j = 0; j < kXT; j++) {
    const int ix = xt * kXT + j;
    xs[j] = ix < nx ? x[ix] : 1.0;
    a1[j] = theta ? (kLog2 - log(xs[j])) : log(xs[j]);
    a2[j] = (!theta && V.expoprior) ? (log(V.m_meaninv[p]) - xs[j] * V.m_meaninv[p]) : 0.0;
    acc[j] = 0.0;
  }
  const float *c0 = V.cols + (size_t)((theta ? V.ccp : V.mcp) + p) * V.G;
  const float *c1 = V.cols + (size_t)((theta ? V.fcp : V.fmp) + p) * V.G;
  const float *c2 = V.cols + (size_t)((theta ? V.qip : V.mip) + p) * V.G;
  const float *c3 = V.cols + (size_t)(V.hccp + (theta ? p : 0)) * V.G;
  for (long long r = r0 + warp * IMA_WARP + lane; r < r1; r += kLmWarps * IMA_WARP) {
    double cnt = c0[r];
    const double f = c1[r], integ = c2[r];
    if (round_counts) cnt = integerround(cnt);
    if (theta) {
      const double h = c3[r];
#pragma unroll
      for (int j = 0; j < kXT; j++) acc[j] += exp(-integ + cnt * a1[j] - h - 2 * f / xs[j]);
    } else if (V.expoprior) {
#pragma unroll
      for (int j = 0; j < kXT; j++) acc[j] += exp(-integ + (a2[j] + cnt * a1[j]) - f * xs[j]);   // :66-68 / :156-158 up to association
    } else {
#pragma unroll
      for (int j = 0; j < kXT; j++) acc[j] += exp(-integ + cnt * a1[j] - f * xs[j]);
    }
  }
  double *sm = (double *)IMA_SMEM;     // [kLmWarps][kXT]
  for (int j = 0; j < kXT; j++) {
    const double s = Warp::sum(acc[j]);
    if (lane == 0) sm[warp * kXT + j] = s;
  }
#if IMA_CUDA
  __syncthreads();
  if (threadIdx.x < kXT) {
    double s = 0.0;
    for (int w = 0; w < kLmWarps; w++) s += sm[w * kXT + threadIdx.x];
    partials[(size_t)chunk * (nxt * kXT) + xt * kXT + threadIdx.x] = s;
  }
#else
  // host emulation runs the "warps" of a block one after the other: the last one folds
  if (warp == kLmWarps - 1)
    for (int j = 0; j < kXT; j++) {
      double s = 0.0;
      for (int w = 0; w < kLmWarps; w++) s += sm[w * kXT + j];
      partials[(size_t)chunk * (nxt * kXT) + xt * kXT + j] = s;
    }
#endif
}

// out[i] = sum over chunks (in chunk order) of partials[chunk][i]
IMA_KERNEL void k_reduce_partials(const double *partials, int nchunks, int width, int n, double *out) {
  const int i = ima_block() * kLmWarps * IMA_WARP + ima_warp_in_block() * IMA_WARP + Warp::lane();
  if (i >= n) return;
  double s = 0.0;
  for (int c = 0; c < nchunks; c++) s += partials[(size_t)c * width + i];
  out[i] = s;
}

// Many independent one-point marginal evaluations in one launch: the lock-step peak and bound searches of the front end (one
// search per parameter and row set, each wanting its next abscissa) hand all their current points to one pass.  Request q owns
// the blocks [block0[q], block0[q+1]): one block per chunk of its row range, one evaluation point per block.  Lane, warp and chunk
// order of the additions are k_marginal's with nx = 1, so a value is bit for bit the one the single-request call gives.
struct MargReq { long long first, last; double x; int param, round_counts, block0, pad; };

IMA_KERNEL void k_marginal_many(LmView V, const MargReq *req, int nreq, double *partials) {
  IMA_SMEM_DECL
  const int lane = Warp::lane(), warp = ima_warp_in_block();
  int lo = 0, hi = nreq - 1;                   // the request this block belongs to (block0 ascending)
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (req[mid].block0 <= (int)ima_block()) lo = mid; else hi = mid - 1; }
  const MargReq R = req[lo];
  const int chunk = ima_block() - R.block0;
  const long long r0 = R.first + (long long)chunk * kRowsPerBlock;
  long long r1 = r0 + kRowsPerBlock;
  if (r1 > R.last) r1 = R.last;
  const bool theta = R.param < V.nq;
  const int p = theta ? R.param : R.param - V.nq;
  const double xs = R.x;
  const double a1 = theta ? (kLog2 - log(xs)) : log(xs);
  const double a2 = (!theta && V.expoprior) ? (log(V.m_meaninv[p]) - xs * V.m_meaninv[p]) : 0.0;
  const float *c0 = V.cols + (size_t)((theta ? V.ccp : V.mcp) + p) * V.G;
  const float *c1 = V.cols + (size_t)((theta ? V.fcp : V.fmp) + p) * V.G;
  const float *c2 = V.cols + (size_t)((theta ? V.qip : V.mip) + p) * V.G;
  const float *c3 = V.cols + (size_t)(V.hccp + (theta ? p : 0)) * V.G;
  double acc = 0.0;
  for (long long r = r0 + warp * IMA_WARP + lane; r < r1; r += kLmWarps * IMA_WARP) {
    double cnt = c0[r];
    const double f = c1[r], integ = c2[r];
    if (R.round_counts) cnt = integerround(cnt);
    if (theta) acc += exp(-integ + cnt * a1 - (double)c3[r] - 2 * f / xs);
    else if (V.expoprior) acc += exp(-integ + (a2 + cnt * a1) - f * xs);
    else acc += exp(-integ + cnt * a1 - f * xs);
  }
  double *sm = (double *)IMA_SMEM;     // [kLmWarps]
  const double s = Warp::sum(acc);
  if (lane == 0) sm[warp] = s;
#if IMA_CUDA
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kLmWarps; w++) t += sm[w];
    partials[ima_block()] = t;
  }
#else
  if (warp == kLmWarps - 1) {
    double t = 0.0;
    for (int w = 0; w < kLmWarps; w++) t += sm[w];
    partials[ima_block()] = t;
  }
#endif
}

// out[q] = the chunk partials of request q added in chunk order
IMA_KERNEL void k_reduce_many(const MargReq *req, int nreq, int nblocks, const double *partials, double *out) {
  const int q = ima_block() * kLmWarps * IMA_WARP + ima_warp_in_block() * IMA_WARP + Warp::lane();
  if (q >= nreq) return;
  const int b1 = q + 1 < nreq ? req[q + 1].block0 : nblocks;
  double s = 0.0;
  for (int b = req[q].block0; b < b1; b++) s += partials[b];
  out[q] = s;
}

// what jointp precomputes per parameter vector (jointfind.cpp:955-970), two numbers per parameter: log(2/x) and 1/x for a
// population size, log x and x for a migration rate; layout [vector][2][np]

// p_g of jointp (jointfind.cpp:971-996, two populations / full model).  One block = a chunk of rows x a tile of kVT parameter
// vectors; any number of vectors per launch (grid = chunks x tiles).  The tile's accumulators and running maxima live in registers
// (loops over the tile fully unrolled); the per-vector coefficients of parameter i (log(2/x), 1/x for a size, log x, x for a
// migration rate) are staged once per block in shared memory and read as broadcasts; a row's columns are read once per tile,
// parameter by parameter, so nothing is indexed dynamically and nothing spills.  Every p is the same sequence of additions as
// the reference's loop over the parameters (:974-992).
constexpr int kVT = 16;              // parameter vectors per block of k_joint_terms
IMA_HD size_t joint_terms_smem(int np) { return (size_t)(2 * kVT * np + kLmWarps * kVT) * sizeof(double); }

IMA_KERNEL void k_joint_terms(LmView V, const double *coef, int nvec, int nchunks, int modeltype, double *pbuf, double *chunkmax) {
  IMA_SMEM_DECL
  const int lane = Warp::lane(), warp = ima_warp_in_block();
  const int tile = ima_block() / nchunks, chunk = ima_block() - tile * nchunks;
  const int v0 = tile * kVT, nt = nvec - v0 < kVT ? nvec - v0 : kVT;
  const long long r0 = (long long)chunk * kRowsPerBlock;
  long long r1 = r0 + kRowsPerBlock;
  if (r1 > V.G) r1 = V.G;
  const int np = V.nq + V.nm;
  double *ca = (double *)IMA_SMEM, *cb = ca + kVT * np, *sm = cb + kVT * np;     // [np][kVT], [np][kVT], [kLmWarps][kVT]
#if IMA_CUDA
  for (int k = threadIdx.x; k < kVT * np; k += blockDim.x) {
#else
  for (int k = 0; k < kVT * np; k++) {                   // the emulation runs a block's warps one after the other
#endif
    const int i = k / kVT, j = k - i * kVT;
    const double *X = coef + (size_t)(v0 + (j < nt ? j : 0)) * 2 * np;
    ca[k] = X[i];
    cb[k] = X[np + i];
  }
#if IMA_CUDA
  __syncthreads();
#endif
  double vmax[kVT];
#pragma unroll
  for (int j = 0; j < kVT; j++) vmax[j] = -DBL_MAX;
  for (long long r = r0 + warp * IMA_WARP + lane; r < r1; r += kLmWarps * IMA_WARP) {
    // modeltype 0: two populations, all parameters, p starts at -probg (:973); 1 / 2: the size-only and migration-only models
    // of a three-population search, p starts at minus the sum of that family's integrals (:949-952, :976-978)
    double probg = 0.0;
    if (modeltype == 0) probg = V.cols[(size_t)V.probgp * V.G + r];
    else if (modeltype == 1) for (int i = 0; i < V.nq; i++) probg += V.cols[(size_t)(V.qip + i) * V.G + r];
    else for (int i = 0; i < V.nm; i++) probg += V.cols[(size_t)(V.mip + i) * V.G + r];
    double p[kVT];
#pragma unroll
    for (int j = 0; j < kVT; j++) p[j] = -probg;
    const int nq_used = modeltype == 2 ? 0 : V.nq, nm_used = modeltype == 1 ? 0 : V.nm;
    for (int i = 0; i < nq_used; i++) {
      const double cc = V.cols[(size_t)(V.ccp + i) * V.G + r], hc = V.cols[(size_t)(V.hccp + i) * V.G + r];
      const double fc2 = 2.0 * V.cols[(size_t)(V.fcp + i) * V.G + r];
      const double *a = ca + i * kVT, *b = cb + i * kVT;
#pragma unroll
      for (int j = 0; j < kVT; j++) p[j] += cc * a[j] - hc - fc2 * b[j];
    }
    for (int i = 0; i < nm_used; i++) {
      const double mc = V.cols[(size_t)(V.mcp + i) * V.G + r], fm = V.cols[(size_t)(V.fmp + i) * V.G + r];
      const double *a = ca + (V.nq + i) * kVT, *b = cb + (V.nq + i) * kVT;
#pragma unroll
      for (int j = 0; j < kVT; j++) p[j] += mc * a[j] - fm * b[j];
    }
#pragma unroll
    for (int j = 0; j < kVT; j++)
      if (j < nt) {
        pbuf[(size_t)(v0 + j) * V.G + r] = p[j];
        if (p[j] > vmax[j]) vmax[j] = p[j];
      }
  }
#pragma unroll
  for (int j = 0; j < kVT; j++) {
    double m = vmax[j];
#if IMA_CUDA
    for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, m, o); m = t > m ? t : m; }
#endif
    if (lane == 0) sm[warp * kVT + j] = m;
  }
#if IMA_CUDA
  __syncthreads();
  if ((int)threadIdx.x < nt) {
    double m = -DBL_MAX;
    for (int w = 0; w < kLmWarps; w++) m = sm[w * kVT + threadIdx.x] > m ? sm[w * kVT + threadIdx.x] : m;
    chunkmax[(size_t)(v0 + threadIdx.x) * nchunks + chunk] = m;
  }
#else
  if (warp == kLmWarps - 1)
    for (int j = 0; j < nt; j++) {
      double m = -DBL_MAX;
      for (int w = 0; w < kLmWarps; w++) m = sm[w * kVT + j] > m ? sm[w * kVT + j] : m;
      chunkmax[(size_t)(v0 + j) * nchunks + chunk] = m;
    }
#endif
}

// per vector: exclusive prefix maxima of the chunks (seeded with the maximum of the rows held by earlier ranks) and the maximum
// over this rank's rows.  One warp per vector, 32 chunks a trip (a prefix maximum is exact: the serial walk's values)
IMA_KERNEL void k_joint_prefix(const double *chunkmax, int nchunks, int nvec, const double *seed_before, double *chunkprefix, double *localmax) {
  const int v = ima_block() * kLmWarps + ima_warp_in_block(), lane = Warp::lane();
  if (v >= nvec) return;
  double run = seed_before ? seed_before[v] : -DBL_MAX;
  for (int base = 0; base < nchunks; base += IMA_WARP) {
    const int c = base + lane;
    double pre = c < nchunks ? chunkmax[(size_t)v * nchunks + c] : -DBL_MAX;
    for (int o = 1; o < IMA_WARP; o <<= 1) { const double t = Warp::shfl_up(pre, o); if (lane >= o && t > pre) pre = t; }
    double before = Warp::shfl_up(pre, 1);
    if (lane == 0) before = -DBL_MAX;
    if (run > before) before = run;
    if (c < nchunks) chunkprefix[(size_t)v * nchunks + c] = before;
    const double groupmax = Warp::bcast(pre, IMA_WARP - 1);
    if (groupmax > run) run = groupmax;
  }
  if (lane == 0) localmax[v] = run;      // includes the seed
}

// partial record of one chunk: inserted count, kept count, sum, sum of squares, smallest kept p, its scaled term
constexpr int kJP = 6;

IMA_DEV int lowest_bit(unsigned m) {
#if IMA_CUDA
  return __ffs((int)m) - 1;
#else
  int i = 0;
  while (!(m & 1u)) { m >>= 1; i++; }
  return i;
#endif
}

IMA_KERNEL void k_joint_scan(LmView V, const double *pbuf, int nvec, const double *chunkprefix, const double *globalmax,
                             long long global_row0, const double *pow10, double *partials) {
  IMA_SMEM_DECL
  // one warp per (chunk, vector): rows of the chunk are walked in order, 32 at a time, with a running maximum
  const int nchunks = (int)((V.G + kRowsPerBlock - 1) / kRowsPerBlock);
  const int job = ima_block() * kLmWarps + ima_warp_in_block();
  if (job >= nchunks * nvec) return;
  const int v = job / nchunks, chunk = job - v * nchunks;
  const int lane = Warp::lane();
  const long long r0 = (long long)chunk * kRowsPerBlock;
  long long r1 = r0 + kRowsPerBlock;
  if (r1 > V.G) r1 = V.G;
  const double gmax = globalmax[v];
  double mmax; int zmax;
  eexp(gmax, mmax, zmax);
  const int maxz = zmax - 10;                       // OCUTOFF :1024
  double run = chunkprefix[(size_t)v * nchunks + chunk];
  double inserted = 0.0, kept = 0.0, sum = 0.0, sumsq = 0.0, minp = DBL_MAX, minterm = 0.0;
  const double *pb = pbuf + (size_t)v * V.G;
  // Rows that go into the sum are few and scattered, and their term (eexp: a division, a polynomial, a fractional power of ten)
  // is by far the longest code of the loop: taken where it is found, a group with one such row pays for it in full.  So the
  // rows are queued (per warp, in shared memory) and their terms are taken 32 at a time, one per lane.
  double *queue = (double *)IMA_SMEM + (size_t)ima_warp_in_block() * 2 * IMA_WARP;
  int qn = 0;
  auto take_term = [&](double pk) {
    double m; int z;
    eexp(pk, m, z);
    const int zadj = z - maxz;
    const double term = (zadj > -308 && zadj < 308) ? m * pow10[zadj + 308] : (zadj <= -308 ? 0.0 : DBL_MAX);      // the table of :907-908
    kept += 1.0; sum += term; sumsq += term * term;
    if (pk < minp) { minp = pk; minterm = term; }
  };
  auto group = [&](long long r, bool valid, double p) {
    // a row at least 10 below the carried maximum is neither inserted (its own running maximum is no smaller) nor a new maximum:
    // only the other rows -- the candidates -- matter, to themselves and to the rows after them.  No candidate in the group:
    // nothing to do, the common case away from the bulk of the posterior.
    const unsigned cm = Warp::ballot(valid && (run - p < 10 || global_row0 + r == 0));
    if (!cm) return;
    // running maximum of the rows before r (`before`) and of the whole group
    double before = run, groupmax = run;
    if (Warp::popc(cm) <= 4) {                       // few candidates: each tells the lanes after it its value
      for (unsigned m = cm; m; m &= m - 1u) {
        const int c = lowest_bit(m);
        const double pc = Warp::bcast(p, c);
        if (lane > c && pc > before) before = pc;
        if (pc > groupmax) groupmax = pc;
      }
    } else {                                         // many: prefix maximum over the 32 lanes by shuffles
      double pre = p;
      for (int o = 1; o < IMA_WARP; o <<= 1) { const double t = Warp::shfl_up(pre, o); if (lane >= o && t > pre) pre = t; }
      const double prev = Warp::shfl_up(pre, 1);
      if (lane > 0 && prev > before) before = prev;
      const double gm = Warp::bcast(pre, IMA_WARP - 1);
      if (gm > groupmax) groupmax = gm;
    }
    bool keep = false;
    if (valid) {
      const bool first_row = (global_row0 + r == 0);
      if (first_row || before - p < 10) {            // :1005 (row 0 is always the list head :998-1003)
        inserted += 1.0;
        keep = gmax - p < 10;                        // :1022
      }
    }
    if (groupmax > run) run = groupmax;
    const unsigned km = Warp::ballot(keep);
    if (km) {
      if (keep) queue[qn + Warp::popc(km & ((1u << lane) - 1u))] = p;
      qn += Warp::popc(km);
      Warp::sync();
      if (qn >= IMA_WARP) {                          // a full set: one term per lane
        take_term(queue[lane]);
        const int left = qn - IMA_WARP;
        const double carry = lane < left ? queue[IMA_WARP + lane] : 0.0;
        Warp::sync();
        if (lane < left) queue[lane] = carry;
        qn = left;
        Warp::sync();
      }
    }
    };
  // four groups of rows are loaded before the first of them is processed: the scan is one dependent walk per warp, and with a
  // single 256-byte load in flight per warp it ran at the memory latency, not the bandwidth (ncu: long_scoreboard 15 of 19)
  constexpr int kAhead = 4;
  for (long long base = r0; base < r1; base += kAhead * IMA_WARP) {
    double pa[kAhead];
#pragma unroll
    for (int k = 0; k < kAhead; k++) {
      const long long r = base + k * IMA_WARP + lane;
      pa[k] = r < r1 ? pb[r] : -DBL_MAX;
    }
#pragma unroll
    for (int k = 0; k < kAhead; k++) {
      const long long r = base + k * IMA_WARP + lane;
      if (base + k * IMA_WARP < r1) group(r, r < r1, pa[k]);
    }
  }
  if (lane < qn) take_term(queue[lane]);
  inserted = Warp::sum(inserted); kept = Warp::sum(kept); sum = Warp::sum(sum); sumsq = Warp::sum(sumsq);
#if IMA_CUDA
  for (int o = 16; o > 0; o >>= 1) {
    const double op = __shfl_xor_sync(0xffffffffu, minp, o), ot = __shfl_xor_sync(0xffffffffu, minterm, o);
    if (op < minp) { minp = op; minterm = ot; }
  }
#endif
  if (lane == 0) {
    double *o = partials + ((size_t)v * nchunks + chunk) * kJP;
    o[0] = inserted; o[1] = kept; o[2] = sum; o[3] = sumsq; o[4] = minp; o[5] = minterm;
  }
}

// fold the chunk records of each vector: one warp per vector, lanes over the chunks, then a fixed butterfly (deterministic)
IMA_KERNEL void k_joint_fold(const double *partials, int nchunks, int nvec, double *out) {
  const int v = ima_block() * kLmWarps + ima_warp_in_block(), lane = Warp::lane();
  if (v >= nvec) return;
  double ins = 0, kept = 0, sum = 0, sq = 0, minp = DBL_MAX, minterm = 0;
  for (int c = lane; c < nchunks; c += IMA_WARP) {
    const double *o = partials + ((size_t)v * nchunks + c) * kJP;
    ins += o[0]; kept += o[1]; sum += o[2]; sq += o[3];
    if (o[4] < minp) { minp = o[4]; minterm = o[5]; }
  }
  Warp::sum4(ins, kept, sum, sq);
#if IMA_CUDA
  for (int o = 16; o > 0; o >>= 1) {
    const double op = __shfl_xor_sync(0xffffffffu, minp, o), ot = __shfl_xor_sync(0xffffffffu, minterm, o);
    if (op < minp) { minp = op; minterm = ot; }
  }
#endif
  if (lane == 0) {
    double *r = out + (size_t)v * kJP;
    r[0] = ins; r[1] = kept; r[2] = sum; r[3] = sq; r[4] = minp; r[5] = minterm;
  }
}

// device-resident exchange of the sharded jointp (ima2p_lmode_joint_begin / _middle): from the local maxima of all ranks, the
// maximum over the rows held by lower ranks (the seed of this rank's running maximum) and the maximum over all rows
IMA_KERNEL void k_joint_seed(const double *allmax, int world, int rank, int nvec, double *seed, double *gmax) {
  const int v = ima_block() * kLmWarps * IMA_WARP + ima_warp_in_block() * IMA_WARP + Warp::lane();
  if (v >= nvec) return;
  double s = -DBL_MAX, g = -DBL_MAX;
  for (int r = 0; r < world; r++) {
    const double m = allmax[(size_t)r * nvec + v];
    if (r < rank && m > s) s = m;
    if (m > g) g = m;
  }
  seed[v] = s; gmax[v] = g;
}
// records of this rank with the global maximum beside them: [nvec][8] = the six of k_joint_fold, the global maximum, 0
IMA_KERNEL void k_joint_pack(const double *rec6, const double *gmax, int nvec, double *out8) {
  const int v = ima_block() * kLmWarps * IMA_WARP + ima_warp_in_block() * IMA_WARP + Warp::lane();
  if (v >= nvec) return;
  for (int k = 0; k < kJP; k++) out8[(size_t)v * 8 + k] = rec6[(size_t)v * kJP + k];
  out8[(size_t)v * 8 + 6] = gmax[v]; out8[(size_t)v * 8 + 7] = 0.0;
}


// ---- section 8 (f3): the other evaluators that stream over the rows -------------------------------------------------
// calcx moments (output.cpp:14-134, 687-745) and the densities of the product 2NM (popmig.cpp:9-357).  Same shape as
// k_marginal -- one pass over the column-major rows, per-block partials folded in block order -- but every row costs a
// few incomplete gamma functions (scalar forms of ima_math.h, one lane each), so these are bound by FP64 throughput.
struct LmPriors { double q_max[kMaxParams], m_max[kMaxParams], m_mean[kMaxParams]; };
constexpr double kMinParamVal = 0.0000001;       // MINPARAMVAL imamp.hpp:130
constexpr int kMomentsMaxParams = 32;

// calcx output.cpp:14-134: E[x] (mode 0) or E[x^2] (mode 1) of parameter pnum given genealogy row r
IMA_DEV double calcx_row(const LmView &V, const MathCtx &mc, const LmPriors &P, long long r, int pnum, int mode) {
  double tempval;
  if (pnum < V.nq) {
    const int p = pnum;
    const double max = P.q_max[p];
    if (max <= kMinParamVal) return -1;
    const int cc = (int)V.cols[(size_t)(V.ccp + p) * V.G + r];
    const double fc = V.cols[(size_t)(V.fcp + p) * V.G + r];
    const double hval = V.cols[(size_t)(V.hccp + p) * V.G + r];
    const double denom = V.cols[(size_t)(V.qip + p) * V.G + r];
    if (mode == 0) {
      if (cc == 0 && fc == 0) tempval = (max * max / 2) / exp(denom);
      else if (cc > 1) tempval = exp(2 * kLog2 - hval + (2 - cc) * log(fc) + uppergamma(mc, cc - 2, 2 * fc / max) - denom);
      else if (cc == 1) tempval = exp(kLog2 - hval + log(max * exp(-2 * fc / max) - 2 * fc * exp(uppergamma(mc, 0, 2 * fc / max))) - denom);
      else tempval = exp(log((max / 2) * (max - 2 * fc) * exp(-2 * fc / max) + 2 * (fc * fc) * exp(uppergamma(mc, 0, 2 * fc / max))) - denom);
    } else {
      if (cc == 0 && fc == 0) tempval = (max * (max * max) / 3) / exp(denom);
      else if (cc > 2) tempval = exp(uppergamma(mc, cc - 3, 2 * fc / max) + 3 * kLog2 - hval + (3 - cc) * log(fc) - denom);
      else if (cc == 2) tempval = exp(2 * kLog2 - hval + log(max * exp(-2 * fc / max) - 2 * fc * exp(uppergamma(mc, 0, 2 * fc / max))) - denom);
      else if (cc == 1) tempval = exp(-hval + log(max * (max - 2 * fc) * exp(-2 * fc / max) + 4 * (fc * fc) * exp(uppergamma(mc, 0, 2 * fc / max))) - denom);
      else tempval = exp(-log(3.0) + log(max * (2 * (fc * fc) - fc * max + (max * max)) * exp(-2 * fc / max) - 4 * pow(fc, 3.0) * exp(uppergamma(mc, 0, 2 * fc / max))) - denom);
    }
  } else {
    const int p = pnum - V.nq;
    const double max = P.m_max[p];
    if (max <= kMinParamVal) return -1;
    const int mcnt = (int)V.cols[(size_t)(V.mcp + p) * V.G + r];
    const double fm = V.cols[(size_t)(V.fmp + p) * V.G + r];
    const double denom = V.cols[(size_t)(V.mip + p) * V.G + r];
    if (mode == 0) {
      if (mcnt == 0 && fm == 0) tempval = (max * max / 2) / exp(denom);
      else if (mcnt > 0) tempval = exp(lowergamma(mc, mcnt + 2, fm * max) - (mcnt + 2) * log(fm) - denom);
      else tempval = (1 - (1 + fm * max) * exp(-fm * max)) / (fm * fm) / exp(denom);
    } else {
      if (mcnt == 0 && fm == 0) tempval = (pow(max, 3.0) / 3) / exp(denom);
      else tempval = exp(lowergamma(mc, mcnt + 3, fm * max) - (mcnt + 3) * log(fm) - denom);
    }
  }
  return tempval;
}

IMA_KERNEL void k_upper0(MathCtx mc, const double *x, int n, double *out) {
  const int i = ima_block() * IMA_WARP + Warp::lane();
  if (i < n) out[i] = uppergamma(mc, 0, x[i]);
}

// partials[chunk][2 np + np (np-1)/2]: sums over the chunk's rows of calcx(.,p,0), calcx(.,p,1) and of the products
// calcx(.,p,0) calcx(.,q,0), p < q (output.cpp:704-728)
IMA_KERNEL void k_moments(LmView V, MathCtx mc, const LmPriors *pri, double *partials) {
  IMA_SMEM_DECL
  const int lane = Warp::lane(), warp = ima_warp_in_block();
  const int chunk = ima_block();
  const long long r0 = (long long)chunk * kRowsPerBlock;
  long long r1 = r0 + kRowsPerBlock;
  if (r1 > V.G) r1 = V.G;
  const int np = V.nq + V.nm, nacc = 2 * np + np * (np - 1) / 2;
  double *sm = (double *)IMA_SMEM + (size_t)warp * nacc;       // [kLmWarps][nacc]
  for (int i = lane; i < nacc; i += IMA_WARP) sm[i] = 0.0;
  Warp::sync();
  const LmPriors &P = *pri;
  for (long long base = r0 + (long long)warp * IMA_WARP; base < r1; base += kLmWarps * IMA_WARP) {
    const long long r = base + lane;
    const bool valid = r < r1;
    double x0[kMomentsMaxParams];
    for (int p = 0; p < np; p++) {
      const double a = valid ? calcx_row(V, mc, P, r, p, 0) : 0.0, b = valid ? calcx_row(V, mc, P, r, p, 1) : 0.0;
      x0[p] = a;
      const double sa = Warp::sum(a), sb = Warp::sum(b);
      if (lane == 0) { sm[p] += sa; sm[np + p] += sb; }
    }
    int k = 2 * np;
    for (int p = 0; p < np - 1; p++)
      for (int q = p + 1; q < np; q++, k++) {
        const double s = Warp::sum(x0[p] * x0[q]);
        if (lane == 0) sm[k] += s;
      }
  }
#if IMA_CUDA
  __syncthreads();
  const double *all = (const double *)IMA_SMEM;
  for (int i = threadIdx.x; i < nacc; i += blockDim.x) {
    double s = 0.0;
    for (int w = 0; w < kLmWarps; w++) s += all[(size_t)w * nacc + i];
    partials[(size_t)chunk * nacc + i] = s;
  }
#else
  if (warp == kLmWarps - 1) {
    const double *all = (const double *)IMA_SMEM;
    for (int i = 0; i < nacc; i++) {
      double s = 0.0;
      for (int w = 0; w < kLmWarps; w++) s += all[(size_t)w * nacc + i];
      partials[(size_t)chunk * nacc + i] = s;
    }
  }
#endif
}

// one row's term of calc_popmig / marginpopmig (popmig.cpp:29-89, 213-262); false when the reference skips the row
IMA_DEV bool popmig_term(const LmView &V, const MathCtx &mc, long long r, int thetai, int mi, double x, double qmax, double mmax, double &val) {
  const int cc = (int)V.cols[(size_t)(V.ccp + thetai) * V.G + r];
  const double fc = V.cols[(size_t)(V.fcp + thetai) * V.G + r];
  const double hc = V.cols[(size_t)(V.hccp + thetai) * V.G + r];
  const int mcnt = (int)V.cols[(size_t)(V.mcp + mi) * V.G + r];
  const double fm = V.cols[(size_t)(V.fmp + mi) * V.G + r];
  const double qintg = V.cols[(size_t)(V.qip + thetai) * V.G + r];
  const double mintg = V.cols[(size_t)(V.mip + mi) * V.G + r];
  double temp1, temp2;
  if (fc == 0 && cc == 0 && fm > 0) {
    temp1 = kLog2 - (mcnt * log(fm)) - hc - qintg - mintg;
    temp2 = log(exp(uppergamma(mc, mcnt, 2 * fm * x / qmax)) - exp(uppergamma(mc, mcnt, mmax * fm)));
  } else if (fm == 0 && mcnt == 0 && fc > 0) {
    temp1 = kLog2 - (cc * log(fc)) - hc - qintg - mintg;
    temp2 = log(exp(uppergamma(mc, cc, 2 * fc / qmax)) - exp(uppergamma(mc, cc, fc * mmax / x)));
  } else if (fc == 0 && cc == 0 && mcnt == 0 && fm == 0) {
    temp1 = log(2 * log(mmax * qmax / (2 * x))) - hc - qintg - mintg;
    temp2 = 0;
  } else {
    temp1 = kLog2 + (mcnt * log(x)) - ((cc + mcnt) * log(fc + fm * x)) - hc - qintg - mintg;
    double a = uppergamma(mc, cc + mcnt, 2 * (fc + fm * x) / qmax);
    double b = uppergamma(mc, cc + mcnt, mmax * (fm + fc / x));
    if (a == b) {                     // both saturated: the difference of the lower gammas carries the information (:62-66)
      b = lowergamma(mc, cc + mcnt, 2 * (fc + fm * x) / qmax);
      a = lowergamma(mc, cc + mcnt, mmax * (fm + fc / x));
    }
    if (a > b) logdiff(mc, temp2, a, b);               // LogDiff imamp.hpp:257-263
    else temp1 = temp2 = 0.0;
  }
  if ((temp1 + temp2 < 700) && (temp1 + temp2 > -700)) { val = exp(temp1 + temp2); return true; }
  return false;
}

// sums over rows [first, last) of the 2NM density terms at nx points; layout of blocks and partials as k_marginal
IMA_KERNEL void k_popmig(LmView V, MathCtx mc, const LmPriors *pri, int thetai, int mi, const double *x, int nx, long long first, long long last,
                         double *partials) {
  IMA_SMEM_DECL
  const int lane = Warp::lane(), warp = ima_warp_in_block();
  const int nxt = (nx + kXT - 1) / kXT;
  const int chunk = ima_block() / nxt, xt = ima_block() - chunk * nxt;
  const long long r0 = first + (long long)chunk * kRowsPerBlock;
  long long r1 = r0 + kRowsPerBlock;
  if (r1 > last) r1 = last;
  const double qmax = pri->q_max[thetai], mmax = pri->m_max[mi];
  double acc[kXT];
  for (int j = 0; j < kXT; j++) acc[j] = 0.0;
  for (long long r = r0 + warp * IMA_WARP + lane; r < r1; r += kLmWarps * IMA_WARP)
    for (int j = 0; j < kXT; j++) {
      const int ix = xt * kXT + j;
      double v;
      if (ix < nx && popmig_term(V, mc, r, thetai, mi, x[ix], qmax, mmax, v)) acc[j] += v;
    }
  double *sm = (double *)IMA_SMEM;     // [kLmWarps][kXT]
  for (int j = 0; j < kXT; j++) {
    const double s = Warp::sum(acc[j]);
    if (lane == 0) sm[warp * kXT + j] = s;
  }
#if IMA_CUDA
  __syncthreads();
  if (threadIdx.x < kXT) {
    double s = 0.0;
    for (int w = 0; w < kLmWarps; w++) s += sm[w * kXT + threadIdx.x];
    partials[(size_t)chunk * (nxt * kXT) + xt * kXT + threadIdx.x] = s;
  }
#else
  if (warp == kLmWarps - 1)
    for (int j = 0; j < kXT; j++) {
      double s = 0.0;
      for (int w = 0; w < kLmWarps; w++) s += sm[w * kXT + j];
      partials[(size_t)chunk * (nxt * kXT) + xt * kXT + j] = s;
    }
#endif
}

// exponential migration prior (popmig.cpp:101-170, 272-357): the log term of every row (temp2) goes to tbuf[ix][row - first]
// and the largest base-10 exponent eexp gives any of them to zmax[ix][chunk]; k_expomig_sum then adds the mantissas on the
// common exponent maxz - OCUTOFF
IMA_DEV double expomig_term(const LmView &V, const MathCtx &mc, long long r, int thetai, int mi, double x, double qmax, double mmean) {
  const int cc = (int)V.cols[(size_t)(V.ccp + thetai) * V.G + r];
  const double fc = V.cols[(size_t)(V.fcp + thetai) * V.G + r];
  const double hc = V.cols[(size_t)(V.hccp + thetai) * V.G + r];
  const int mcnt = (int)V.cols[(size_t)(V.mcp + mi) * V.G + r];
  const double fm = V.cols[(size_t)(V.fmp + mi) * V.G + r];
  const double qintg = V.cols[(size_t)(V.qip + thetai) * V.G + r];
  const double mintg = V.cols[(size_t)(V.mip + mi) * V.G + r];
  const double temp1 = x + fc * mmean + fm * mmean * x;
  double temp2 = kLog2 - hc - log(mmean) - qintg - mintg;
  double temp3 = 2 * temp1 / (mmean * qmax);
  if (mcnt == 0 && cc == 0) {
    temp3 = uppergamma(mc, 0, temp3);
    temp2 += temp3;
  } else if (cc == 0) {
    temp3 = uppergamma(mc, mcnt, temp3);
    temp2 += temp3 - mcnt * log(fm + 1 / mmean);
  } else if (mcnt == 0) {
    temp3 = uppergamma(mc, cc, temp3);
    temp2 += temp3 + -cc * log(fc + x * (fm + 1 / mmean));
  } else {
    temp3 = uppergamma(mc, mcnt + cc, temp3);
    temp2 += temp3 - cc * log(x) + (cc + mcnt) * log(mmean * x / temp1);
  }
  return temp2;
}

IMA_KERNEL void k_expomig_terms(LmView V, MathCtx mc, const LmPriors *pri, int thetai, int mi, const double *x, int nx, long long first, long long last,
                                double *tbuf, double *zmax) {
  IMA_SMEM_DECL
  const int lane = Warp::lane(), warp = ima_warp_in_block();
  const int chunk = ima_block();
  const int nchunks = (int)((last - first + kRowsPerBlock - 1) / kRowsPerBlock);
  const long long r0 = first + (long long)chunk * kRowsPerBlock;
  long long r1 = r0 + kRowsPerBlock;
  if (r1 > last) r1 = last;
  const double qmax = pri->q_max[thetai], mmean = pri->m_mean[mi];
  double *sm = (double *)IMA_SMEM;     // [kLmWarps][kJointVecMax]
  for (int ix = 0; ix < nx; ix++) {
    double zm = -1e300;
    for (long long r = r0 + warp * IMA_WARP + lane; r < r1; r += kLmWarps * IMA_WARP) {
      const double t2 = expomig_term(V, mc, r, thetai, mi, x[ix], qmax, mmean);
      tbuf[(size_t)ix * (last - first) + (r - first)] = t2;
      double m; int z;
      eexp(t2, m, z);
      if ((double)z > zm) zm = (double)z;
    }
#if IMA_CUDA
    for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, zm, o); zm = t > zm ? t : zm; }
#endif
    if (lane == 0) sm[warp * kJointVecMax + ix] = zm;
  }
#if IMA_CUDA
  __syncthreads();
  if ((int)threadIdx.x < nx) {
    double m = -1e300;
    for (int w = 0; w < kLmWarps; w++) m = sm[w * kJointVecMax + threadIdx.x] > m ? sm[w * kJointVecMax + threadIdx.x] : m;
    zmax[(size_t)threadIdx.x * nchunks + chunk] = m;
  }
#else
  if (warp == kLmWarps - 1)
    for (int ix = 0; ix < nx; ix++) {
      double m = -1e300;
      for (int w = 0; w < kLmWarps; w++) m = sm[w * kJointVecMax + ix] > m ? sm[w * kJointVecMax + ix] : m;
      zmax[(size_t)ix * nchunks + chunk] = m;
    }
#endif
}

// partials[chunk][ix] = sum over the chunk's rows of m 10^(z - (maxz - OCUTOFF)) (popmig.cpp:157-162)
IMA_KERNEL void k_expomig_sum(const double *tbuf, int nx, long long nrows, const double *maxz, double *partials) {
  IMA_SMEM_DECL
  const int lane = Warp::lane(), warp = ima_warp_in_block();
  const int chunk = ima_block();
  const long long r0 = (long long)chunk * kRowsPerBlock;
  long long r1 = r0 + kRowsPerBlock;
  if (r1 > nrows) r1 = nrows;
  double *sm = (double *)IMA_SMEM;     // [kLmWarps][kJointVecMax]
  for (int ix = 0; ix < nx; ix++) {
    const int base = (int)maxz[ix] - 10;
    double acc = 0.0;
    for (long long r = r0 + warp * IMA_WARP + lane; r < r1; r += kLmWarps * IMA_WARP) {
      double m; int z;
      eexp(tbuf[(size_t)ix * nrows + r], m, z);
      acc += m * pow(10.0, (double)(z - base));
    }
    acc = Warp::sum(acc);
    if (lane == 0) sm[warp * kJointVecMax + ix] = acc;
  }
#if IMA_CUDA
  __syncthreads();
  if ((int)threadIdx.x < nx) {
    double s = 0.0;
    for (int w = 0; w < kLmWarps; w++) s += sm[w * kJointVecMax + threadIdx.x];
    partials[(size_t)chunk * kJointVecMax + threadIdx.x] = s;
  }
#else
  if (warp == kLmWarps - 1)
    for (int ix = 0; ix < nx; ix++) {
      double s = 0.0;
      for (int w = 0; w < kLmWarps; w++) s += sm[w * kJointVecMax + ix];
      partials[(size_t)chunk * kJointVecMax + ix] = s;
    }
#endif
}


// ---- greater-than probabilities gtint.cpp:26-330: P(parameter i > parameter j) as the mean over (a thinned set of)
// rows of a closed form or of a trapezoid quadrature (qtrap, EPS 1e-4, JMAX 20) of the reference's integrands.  One
// lane per row; the quadrature is evaluated in the reference's own order so that it stops at the same refinement.
struct GtRow { int cci, ccj, wi, wj; double fci, fcj, hval, denom, qmax, fmi, fmj, mmax; };
constexpr double kGtSwitchLower = 1e-15;          // SWITCH_TO_LOWERGAMMA_CRIT gtint.cpp:23

IMA_DEV double mgt_wj_gt_0(const MathCtx &mc, const GtRow &g, double mi) {      // gtint.cpp:26-47
  if (mi < kMinParamVal) return 0.0;
  const double a = lfact(mc, g.wj);
  const double b = uppergamma(mc, g.wj + 1, g.fmj * mi);
  if (a <= b) return 0.0;
  double temp1;
  if ((a - b) < kGtSwitchLower) temp1 = lowergamma(mc, g.wj + 1, g.fmj * mi);
  else logdiff(mc, temp1, a, b);
  double temp2 = g.wi * log(mi) - g.fmi * mi - (g.wj + 1) * log(g.fmj) + temp1;
  temp2 -= g.denom;
  return exp(temp2);
}

IMA_DEV double pgt_fcj_gt_0(const MathCtx &mc, const GtRow &g, double qi) {      // gtint.cpp:49-79
  if (qi < kMinParamVal) return 0.0;
  const double fcj2 = 2 * g.fcj, fci2 = 2 * g.fci;
  if (g.ccj == 0) {
    const double a = log(qi) - fcj2 / qi;
    const double b = log(fcj2) + uppergamma(mc, 0, fcj2 / qi);
    if (a > b) {
      double temp1;
      logdiff(mc, temp1, a, b);
      const double temp2 = -fci2 / qi + g.cci * log(2 / qi);
      return exp(temp1 + temp2 - g.hval - g.denom);
    }
    return 0.0;
  }
  const double temp1 = uppergamma(mc, g.ccj - 1, fcj2 / qi);
  const double temp2 = kLog2 + g.cci * log(2 / qi) + (1 - g.ccj) * log(g.fcj) - fci2 / qi;
  return exp(temp2 + temp1 - g.hval - g.denom);
}

// qtrap / trapzd gtint.cpp:83-125 (Numerical Recipes): refinement j adds 2^(j-2) midpoints, summed in order
template <int MIG>
IMA_DEV double gt_qtrap(const MathCtx &mc, const GtRow &g, double a, double b) {
  auto f = [&](double x) { return MIG ? mgt_wj_gt_0(mc, g, x) : pgt_fcj_gt_0(mc, g, x); };
  double s = 0.0, olds = -1.0e100;
  for (int j = 1; j <= 20; j++) {
    if (j == 1) s = 0.5 * (b - a) * (f(a) + f(b));
    else {
      int it = 1;
      for (int k = 1; k < j - 1; k++) it <<= 1;
      const double tnm = it, del = (b - a) / tnm;
      double x = a + 0.5 * del, sum = 0.0;
      for (int k = 1; k <= it; k++, x += del) sum += f(x);
      s = 0.5 * (s + (b - a) * sum / tnm);
    }
    if (j > 5 && (fabs(s - olds) < 1.0e-4 * fabs(olds) || (s == 0.0 && olds == 0.0))) return s;
    olds = s;
  }
  return s;
}

IMA_DEV double gtmig_row(const LmView &V, const MathCtx &mc, long long r, int mi, int mj, double mmax) {   // gtint.cpp:137-245
  GtRow g;
  g.mmax = mmax;
  g.fmi = V.cols[(size_t)(V.fmp + mi) * V.G + r];
  g.fmj = V.cols[(size_t)(V.fmp + mj) * V.G + r];
  g.wi = (int)V.cols[(size_t)(V.mcp + mi) * V.G + r];
  g.wj = (int)V.cols[(size_t)(V.mcp + mj) * V.G + r];
  g.denom = V.cols[(size_t)(V.mip + mi) * V.G + r] + V.cols[(size_t)(V.mip + mj) * V.G + r];      // float sum, as gtint.cpp:144
  const double fmi = g.fmi, fmj = g.fmj, denom = g.denom;
  const int wi = g.wi;
  double temp;
  if (g.wj == 0) {
    if (fmj > 0.0) {
      if (wi > 0) {
        const double a = lfact(mc, wi);
        const double b = uppergamma(mc, wi + 1, (fmi + fmj) * mmax);
        const double c = uppergamma(mc, wi + 1, fmi * mmax);
        if (a <= b || a <= c) temp = 0.0;
        else {
          double temp1, temp2;
          if ((a - b) < kGtSwitchLower) temp1 = lowergamma(mc, wi + 1, (fmi + fmj) * mmax);
          else logdiff(mc, temp1, a, b);
          temp1 += -(wi + 1) * log(fmi + fmj);
          if ((a - c) < kGtSwitchLower) temp2 = lowergamma(mc, wi + 1, fmi * mmax);
          else logdiff(mc, temp2, a, c);
          temp2 += -(wi + 1) * log(fmi);
          if (temp2 <= temp1) temp = 0.0;
          else {
            double temp3;
            logdiff(mc, temp3, temp2, temp1);
            temp = exp(temp3 - log(fmj) - denom);
          }
        }
      } else if (fmi > 0.0) {
        const double temp1 = fmi * (exp(-(fmi + fmj) * mmax) - exp(-fmi * mmax));
        const double temp2 = fmj * (1 - exp(-fmi * mmax));
        const double temp3 = (temp1 + temp2) / (fmi * fmj * (fmi + fmj));
        temp = exp(log(temp3) - denom);
      } else {
        const double temp1 = (fmj * mmax - 1.0 + exp(-fmj * mmax)) / (fmj * fmj);
        temp = exp(log(temp1) - denom);
      }
    } else {
      if (wi > 0) {
        const double a = lfact(mc, wi + 1);
        const double b = uppergamma(mc, wi + 2, fmi * mmax);
        if (a <= b) temp = 0.0;
        else {
          double temp1;
          if ((a - b) < kGtSwitchLower) temp1 = lowergamma(mc, wi + 2, fmi * mmax);
          else logdiff(mc, temp1, a, b);
          const double temp2 = -(wi + 2) * log(fmi);
          temp = exp(temp2 + temp1 - denom);
        }
      } else if (fmi > 0.0) {
        const double temp1 = (1.0 - exp(-fmi * mmax) * (fmi * mmax + 1.0)) / (fmi * fmi);
        temp = exp(log(temp1) - denom);
      } else {
        temp = exp(log(mmax * mmax / 2.0) - denom);
      }
    }
  } else {
    temp = gt_qtrap<1>(mc, g, kMinParamVal, mmax);
  }
  return temp < 1.0 ? temp : 1.0;          // DMIN(1.0, temp) :240
}

IMA_DEV double gtpops_row(const LmView &V, const MathCtx &mc, long long r, int pi, int pj, double qmax) {   // gtint.cpp:248-318
  GtRow g;
  g.qmax = qmax;
  g.cci = (int)V.cols[(size_t)(V.ccp + pi) * V.G + r];
  g.ccj = (int)V.cols[(size_t)(V.ccp + pj) * V.G + r];
  g.fci = V.cols[(size_t)(V.fcp + pi) * V.G + r];
  g.fcj = V.cols[(size_t)(V.fcp + pj) * V.G + r];
  // the reference adds the two float row entries before widening (gtint.cpp:266-267): a float sum
  g.hval = V.cols[(size_t)(V.hccp + pi) * V.G + r] + V.cols[(size_t)(V.hccp + pj) * V.G + r];
  g.denom = V.cols[(size_t)(V.qip + pi) * V.G + r] + V.cols[(size_t)(V.qip + pj) * V.G + r];
  const double fci = g.fci, hval = g.hval, denom = g.denom;
  const int cci = g.cci;
  double temp;
  if (g.ccj == 0 && g.fcj == 0) {
    if (fci == 0) temp = exp(2.0 * log(qmax) - kLog2 - hval - denom);
    else if (cci >= 2) {
      const double temp1 = 2 * kLog2 + (2 - cci) * log(fci);
      const double temp2 = uppergamma(mc, cci - 2, 2 * fci / qmax);
      temp = exp(temp1 + temp2 - hval - denom);
    } else if (cci == 1) {
      const double temp1 = 4 * fci * exp(uppergamma(mc, 0, 2 * fci / qmax));
      const double temp2 = 2 * qmax * exp(-2 * fci / qmax) - temp1;
      temp = exp(log(temp2) - hval - denom);
    } else {
      const double temp1 = exp(uppergamma(mc, 0, 2 * fci / qmax));
      const double temp2 = (qmax / 2) * (qmax - 2 * fci) * exp(-2 * fci / qmax) + 2 * fci * fci * temp1;
      temp = exp(log(temp2) - hval - denom);
    }
  } else {
    temp = gt_qtrap<0>(mc, g, kMinParamVal, qmax);
  }
  return temp < 1.0 ? temp : 1.0;
}

// partials[chunk] = sum over the chunk's used rows (row = index * treeinc) of the row terms; kind 0 = gtpops, 1 = gtmig
IMA_KERNEL void k_greater_than(LmView V, MathCtx mc, const LmPriors *pri, int kind, int pi, int pj, int treeinc, int nused, double *partials) {
  IMA_SMEM_DECL
  const int lane = Warp::lane(), warp = ima_warp_in_block();
  const int chunk = ima_block();
  const int rowsper = kLmWarps * IMA_WARP;                         // one row per thread: the quadrature is long
  const int i = chunk * rowsper + warp * IMA_WARP + lane;
  double v = 0.0;
  if (i < nused) {
    const long long r = (long long)i * treeinc;
    v = kind == 0 ? gtpops_row(V, mc, r, pi, pj, pri->q_max[pi]) : gtmig_row(V, mc, r, pi, pj, pri->m_max[pi]);
  }
  double *sm = (double *)IMA_SMEM;     // [kLmWarps]
  v = Warp::sum(v);
  if (lane == 0) sm[warp] = v;
#if IMA_CUDA
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < kLmWarps; w++) s += sm[w];
    partials[chunk] = s;
  }
#else
  if (warp == kLmWarps - 1) {
    double s = 0.0;
    for (int w = 0; w < kLmWarps; w++) s += sm[w];
    partials[chunk] = s;
  }
#endif
}

}  // namespace ima
extern "C" void ima2p_internal_set_error(const char *msg);   // ima_engine.cu: one error string for the whole library
namespace ima {
static int lfail(int code, const std::string &m) { ima2p_internal_set_error(m.c_str()); return code; }

struct Lmode {
  int device = 0;
  LmView v{};
  double q_max[kMaxParams], q_min[kMaxParams], m_max[kMaxParams], m_min[kMaxParams], m_mean[kMaxParams];
  float *d_cols = nullptr;
  double *d_x = nullptr, *d_partials = nullptr, *d_out = nullptr, *d_pbuf = nullptr, *d_chunkmax = nullptr, *d_prefix = nullptr,
         *d_lmax = nullptr, *d_jpart = nullptr, *d_jout = nullptr, *d_seed = nullptr, *d_gmax = nullptr, *d_ltmp = nullptr, *d_wlmax = nullptr, *d_wrec = nullptr,
         *w_pbuf = nullptr, *w_chunkmax = nullptr, *w_prefix = nullptr, *w_jpart = nullptr, *w_jout = nullptr;   // wide batches (joint_begin / _middle)
  double *w_xs = nullptr, *d_xs = nullptr;      // coefficient tables of the vectors of a call
  size_t cap_x = 0, cap_partials = 0;
  double *d_pow10 = nullptr;                   // 10^i, i = -308..308, from the host's pow as the reference's table (jointfind.cpp:907-908)
  int joint_model = 0;                         // which parameters jointp is a function of (ima2p_lmode_set_joint_model)
  MargReq *d_req = nullptr; double *d_mpart = nullptr, *d_mout = nullptr;      // marginal_many: requests, per-block partials, sums
  size_t cap_req = 0, cap_mpart = 0;
  struct LmPriors *d_pri = nullptr;            // section 8 (f3) evaluators: priors, logfact table and error word, made on first use
  double *d_logfact = nullptr, *d_msums = nullptr;
  int *d_err = nullptr;
  MathCtx mc{};
  std::vector<void *> allocs;
#if IMA_CUDA
  cudaStream_t stream = nullptr;
#endif
  template <class T> T *alloc(size_t n) { T *p = (T *)dev_alloc(n * sizeof(T)); if (p) allocs.push_back(p); return p; }
  ~Lmode() {
#if IMA_CUDA
    if (stream) cudaStreamDestroy(stream);
#endif
    for (void *p : allocs) dev_free(p);
  }
};

static stream_t lm_stream(Lmode *l, void *s) {
#if IMA_CUDA
  return s ? (cudaStream_t)s : l->stream;
#else
  (void)l; (void)s; return nullptr;
#endif
}
static bool lm_use(Lmode *l) {
#if IMA_CUDA
  return IMA_CUDA_OK(cudaSetDevice(l->device));
#else
  (void)l; return true;
#endif
}

}  // namespace ima

using namespace ima;
struct ima2p_lmode { Lmode lm; };

extern "C" {

int ima2p_lmode_create(ima2p_lmode **out, int device, int nq, int nm, int nsplit, const double *q_max, const double *q_min,
                       const double *m_max, const double *m_min, const double *m_mean, int expoprior) {
  if (!out || nq < 1 || nq > kMaxParams || nm < 0 || nm > kMaxParams || 3 * nq + 2 * nm > 2 * kMaxParams) return lfail(IMA2P_E_ARG, "lmode_create: bad argument");
#if IMA_CUDA
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return lfail(IMA2P_E_CUDA, "no CUDA device: ima2p_b200 has no CPU path");
  if (device < 0 || device >= ndev) return lfail(IMA2P_E_ARG, "device index out of range");
#endif
  ima2p_lmode *h = new (std::nothrow) ima2p_lmode();
  if (!h) return lfail(IMA2P_E_ARG, "out of host memory");
  Lmode &l = h->lm;
  l.device = device;
  LmView &v = l.v;
  v.nq = nq; v.nm = nm; v.nsplit = nsplit; v.expoprior = expoprior;
  // column offsets: initialize.cpp:710-719
  v.ccp = 0; v.fcp = nq; v.hccp = 2 * nq; v.mcp = 3 * nq; v.fmp = v.mcp + nm; v.qip = v.fmp + nm; v.mip = v.qip + nq;
  v.pdgp = v.mip + nm; v.probgp = v.pdgp + 1;
  v.rowlen = v.probgp + 1 + nsplit;
  for (int i = 0; i < nq; i++) { l.q_max[i] = q_max[i]; l.q_min[i] = q_min[i]; }
  for (int i = 0; i < nm; i++) { l.m_max[i] = m_max[i]; l.m_min[i] = m_min[i]; l.m_mean[i] = m_mean ? m_mean[i] : 0.0; v.m_meaninv[i] = (expoprior && m_mean) ? 1.0 / m_mean[i] : 0.0; }
  if (!lm_use(&l)) { delete h; return lfail(IMA2P_E_CUDA, "cudaSetDevice failed"); }
#if IMA_CUDA
  if (!IMA_CUDA_OK(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking))) { delete h; return lfail(IMA2P_E_CUDA, "stream create failed"); }
#endif
  *out = h;
  return IMA2P_OK;
}

void ima2p_lmode_destroy(ima2p_lmode *h) { if (h) { lm_use(&h->lm); delete h; } }

int ima2p_lmode_load(ima2p_lmode *h, const float *rows, int nrows, int rowlen, long long nrows_total) {
  if (!h || !rows || nrows < 1 || nrows_total < nrows) return lfail(IMA2P_E_ARG, "lmode_load: bad argument");
  Lmode &l = h->lm;
  if (rowlen != l.v.rowlen) return lfail(IMA2P_E_ARG, "lmode_load: row length does not match the model (calc_gsampinf_length, ginfo.cpp:306-316)");
  if (l.d_cols) return lfail(IMA2P_E_ARG, "lmode_load: rows already loaded");
  if (!lm_use(&l)) return lfail(IMA2P_E_CUDA, "cudaSetDevice failed");
  std::vector<float> cols((size_t)nrows * rowlen);
  for (int r = 0; r < nrows; r++) for (int c = 0; c < rowlen; c++) cols[(size_t)c * nrows + r] = rows[(size_t)r * rowlen + c];
  l.d_cols = l.alloc<float>(cols.size());
  const int nchunks = (nrows + kRowsPerBlock - 1) / kRowsPerBlock;
  l.d_pbuf = l.alloc<double>((size_t)kJointVecMax * nrows);
  l.d_chunkmax = l.alloc<double>((size_t)kJointVecMax * nchunks);
  l.d_prefix = l.alloc<double>((size_t)kJointVecMax * nchunks);
  l.d_lmax = l.alloc<double>(kJointVecMax);
  l.d_jpart = l.alloc<double>((size_t)kJointVecMax * nchunks * kJP);
  l.d_jout = l.alloc<double>((size_t)kJointVecMax * kJP);
  l.d_xs = l.alloc<double>((size_t)kJointVecMax * 2 * kMaxParams);
  l.d_pow10 = l.alloc<double>(617);
  if (!l.d_cols || !l.d_pbuf || !l.d_xs || !l.d_pow10) return lfail(IMA2P_E_CUDA, "device allocation failed");
  stream_t s = lm_stream(&l, nullptr);
  double p10[617];
  for (int i = -308; i <= 308; i++) p10[i + 308] = pow(10.0, (double)i);
  if (!h2d(l.d_pow10, p10, sizeof p10, s)) return lfail(IMA2P_E_CUDA, "table upload failed");
  if (!h2d(l.d_cols, cols.data(), cols.size() * sizeof(float), s) || !dev_sync(s)) return lfail(IMA2P_E_CUDA, "row upload failed");
  l.v.cols = l.d_cols; l.v.G = nrows; l.v.G_total = nrows_total;
  return IMA2P_OK;
}

int ima2p_lmode_marginal_sums(ima2p_lmode *h, int param, const double *x, int nx, int first, int last, int round_counts,
                              double *host_sums, double *dev_sums, void *cuda_stream) {
  if (!h || !h->lm.d_cols || !x || nx < 1) return lfail(IMA2P_E_ARG, "marginal_sums: bad argument / rows not loaded");
  Lmode &l = h->lm;
  if (param < 0 || param >= l.v.nq + l.v.nm || first < 0 || last > l.v.G || first >= last) return lfail(IMA2P_E_ARG, "marginal_sums: bad range");
  if (!lm_use(&l)) return lfail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = lm_stream(&l, cuda_stream);
  const int nxt = (nx + kXT - 1) / kXT, width = nxt * kXT;
  const int nchunks = (int)(((long long)last - first + kRowsPerBlock - 1) / kRowsPerBlock);
  if ((size_t)nx > l.cap_x) { l.d_x = l.alloc<double>(nx); l.d_out = l.alloc<double>(width); l.cap_x = nx; }
  if ((size_t)nchunks * width > l.cap_partials) { l.d_partials = l.alloc<double>((size_t)nchunks * width); l.cap_partials = (size_t)nchunks * width; }
  if (!l.d_x || !l.d_out || !l.d_partials) return lfail(IMA2P_E_CUDA, "device allocation failed");
  if (!h2d(l.d_x, x, nx * sizeof(double), s)) return lfail(IMA2P_E_CUDA, "upload failed");
  IMA_LAUNCH(k_marginal, nchunks * nxt, kLmWarps, kLmWarps * kXT * sizeof(double), s, l.v, param, l.d_x, nx, (long long)first, (long long)last, round_counts, l.d_partials);
  double *outp = dev_sums ? dev_sums : l.d_out;
  IMA_LAUNCH(k_reduce_partials, (nx + kLmWarps * IMA_WARP - 1) / (kLmWarps * IMA_WARP), kLmWarps, 0, s, l.d_partials, nchunks, width, nx, outp);
#if IMA_CUDA
  if (!IMA_CUDA_OK(cudaGetLastError())) return lfail(IMA2P_E_CUDA, "kernel launch failed (marginal)");
#endif
  if (host_sums) { if (!d2h(host_sums, outp, nx * sizeof(double), s) || !dev_sync(s)) return lfail(IMA2P_E_CUDA, "download failed"); }
  return IMA2P_OK;
}

// margincalc surface_call_functions.cpp:119-173 (all rows, INTEGERROUND counts, optional log and offset)
int ima2p_lmode_margincalc(ima2p_lmode *h, int param, const double *x, int nx, double yadjust, int logi, double *out) {
  if (!h || !out) return lfail(IMA2P_E_ARG, "margincalc: bad argument");
  Lmode &l = h->lm;
  int rc = ima2p_lmode_marginal_sums(h, param, x, nx, 0, (int)l.v.G, 1, out, nullptr, nullptr);
  if (rc) return rc;
  for (int i = 0; i < nx; i++) {
    double s = out[i] / (double)l.v.G;
    if (logi) s = s <= 0 ? -1e200 : log(s);
    out[i] = s - yadjust;
  }
  return IMA2P_OK;
}

// marginp surface_call_functions.cpp:25-80 (row range, unrounded theta counts, divisor quirk :77, returns -mean)
int ima2p_lmode_marginp(ima2p_lmode *h, int param, int firsttree, int lasttree, const double *x, int nx, double *out) {
  if (!h || !out) return lfail(IMA2P_E_ARG, "marginp: bad argument");
  Lmode &l = h->lm;
  if (param < 0 || param >= l.v.nq + l.v.nm) return lfail(IMA2P_E_ARG, "marginp: bad parameter index");
  const double mx = param < l.v.nq ? l.q_max[param] : l.m_max[param - l.v.nq], mn = param < l.v.nq ? l.q_min[param] : l.m_min[param - l.v.nq];
  int rc = ima2p_lmode_marginal_sums(h, param, x, nx, firsttree, lasttree, param < l.v.nq ? 0 : 1, out, nullptr, nullptr);
  if (rc) return rc;
  for (int i = 0; i < nx; i++) {
    if (x[i] < mn || x[i] > mx) out[i] = 1;          // OFFSCALEVAL :45-46
    else out[i] = -(out[i] / (lasttree - firsttree + (firsttree == 0)));
  }
  return IMA2P_OK;
}

// The current points of many independent one-dimensional searches in one device pass (k_marginal_many): request q evaluates
// parameter param[q] at x[q]; kind[q] = 0: marginp over rows [first[q], last[q]) (:25-80), kind[q] = 1: log margincalc over all rows
// minus yadjust[q] (:119-173, the function margin95 finds the root of).  One upload, two launches, one download for the batch.
int ima2p_lmode_marginal_many(ima2p_lmode *h, int n, const int *kind, const int *param, const int *first, const int *last, const double *x,
                              const double *yadjust, double *out) {
  if (!h || !h->lm.d_cols || n < 0 || (n && (!kind || !param || !x || !out))) return lfail(IMA2P_E_ARG, "marginal_many: bad argument / rows not loaded");
  if (n == 0) return IMA2P_OK;
  Lmode &l = h->lm;
  std::vector<MargReq> req(n);
  int nblocks = 0;
  for (int q = 0; q < n; q++) {
    MargReq &R = req[q];
    if (param[q] < 0 || param[q] >= l.v.nq + l.v.nm) return lfail(IMA2P_E_ARG, "marginal_many: bad parameter index");
    if (kind[q] == 0) {
      if (!first || !last || first[q] < 0 || last[q] > l.v.G || first[q] >= last[q]) return lfail(IMA2P_E_ARG, "marginal_many: bad range");
      R.first = first[q]; R.last = last[q]; R.round_counts = param[q] < l.v.nq ? 0 : 1;
    } else if (kind[q] == 1) {
      if (!yadjust) return lfail(IMA2P_E_ARG, "marginal_many: kind 1 needs yadjust");
      R.first = 0; R.last = l.v.G; R.round_counts = 1;
    } else return lfail(IMA2P_E_ARG, "marginal_many: kind is 0 (marginp) or 1 (log margincalc)");
    R.x = x[q]; R.param = param[q]; R.block0 = nblocks; R.pad = 0;
    nblocks += (int)((R.last - R.first + kRowsPerBlock - 1) / kRowsPerBlock);
  }
  if (!lm_use(&l)) return lfail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = lm_stream(&l, nullptr);
  if ((size_t)n > l.cap_req) { l.d_req = l.alloc<MargReq>(n); l.d_mout = l.alloc<double>(n); l.cap_req = n; }
  if ((size_t)nblocks > l.cap_mpart) { l.d_mpart = l.alloc<double>(nblocks); l.cap_mpart = nblocks; }
  if (!l.d_req || !l.d_mout || !l.d_mpart) return lfail(IMA2P_E_CUDA, "device allocation failed");
  if (!h2d(l.d_req, req.data(), n * sizeof(MargReq), s)) return lfail(IMA2P_E_CUDA, "upload failed");
  IMA_LAUNCH(k_marginal_many, nblocks, kLmWarps, kLmWarps * sizeof(double), s, l.v, l.d_req, n, l.d_mpart);
  IMA_LAUNCH(k_reduce_many, (n + kLmWarps * IMA_WARP - 1) / (kLmWarps * IMA_WARP), kLmWarps, 0, s, l.d_req, n, nblocks, l.d_mpart, l.d_mout);
#if IMA_CUDA
  if (!IMA_CUDA_OK(cudaGetLastError())) return lfail(IMA2P_E_CUDA, "kernel launch failed (marginal_many)");
#endif
  if (!d2h(out, l.d_mout, n * sizeof(double), s) || !dev_sync(s)) return lfail(IMA2P_E_CUDA, "download failed");
  for (int q = 0; q < n; q++) {
    const int pi = param[q];
    if (kind[q] == 0) {
      const double mx = pi < l.v.nq ? l.q_max[pi] : l.m_max[pi - l.v.nq], mn = pi < l.v.nq ? l.q_min[pi] : l.m_min[pi - l.v.nq];
      if (x[q] < mn || x[q] > mx) out[q] = 1;          // OFFSCALEVAL :45-46
      else out[q] = -(out[q] / (last[q] - first[q] + (first[q] == 0)));
    } else {
      double v = out[q] / (double)l.v.G;
      v = v <= 0 ? -1e200 : log(v);
      out[q] = v - yadjust[q];
    }
  }
  return IMA2P_OK;
}

// two-phase joint evaluation, also the building block of the multi-GPU form:
//   phase 1: terms + this rank's maximum per vector (seeded with the maximum of the rows of earlier ranks)
//   phase 2: keep-set records given the global maximum
static std::vector<double> joint_coefficients(const Lmode &l, const double *x, int nvec) {
  const int nq = l.v.nq, np = nq + l.v.nm;
  std::vector<double> c((size_t)nvec * 2 * np);
  for (int v = 0; v < nvec; v++)
    for (int i = 0; i < np; i++) {                     // jointfind.cpp:955-970
      const double xv = x[(size_t)v * np + i];
      c[((size_t)v * 2) * np + i] = i < nq ? kLog2 - log(xv) : log(xv);
      c[((size_t)v * 2 + 1) * np + i] = i < nq ? 1.0 / xv : xv;
    }
  return c;
}
int ima2p_lmode_joint_phase1(ima2p_lmode *h, const double *x, int nvec, const double *seed_before, double *localmax_out) {
  if (!h || !h->lm.d_cols || !x || nvec < 1 || nvec > kJointVecMax) return lfail(IMA2P_E_ARG, "joint_phase1: bad argument");
  Lmode &l = h->lm;
  if (!lm_use(&l)) return lfail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = lm_stream(&l, nullptr);
  const int np = l.v.nq + l.v.nm, nchunks = (int)((l.v.G + kRowsPerBlock - 1) / kRowsPerBlock);
  const std::vector<double> xs = joint_coefficients(l, x, nvec);
  double *d_seed = nullptr;
  if (seed_before) { d_seed = l.d_jout; if (!h2d(d_seed, seed_before, nvec * sizeof(double), s)) return lfail(IMA2P_E_CUDA, "upload failed"); }
  if (!h2d(l.d_xs, xs.data(), xs.size() * sizeof(double), s)) return lfail(IMA2P_E_CUDA, "upload failed");
  IMA_LAUNCH(k_joint_terms, nchunks * ((nvec + kVT - 1) / kVT), kLmWarps, joint_terms_smem(np), s, l.v, (const double *)l.d_xs, nvec, nchunks, l.joint_model, l.d_pbuf, l.d_chunkmax);
  IMA_LAUNCH(k_joint_prefix, (nvec + kLmWarps - 1) / kLmWarps, kLmWarps, 0, s, l.d_chunkmax, nchunks, nvec, d_seed, l.d_prefix, l.d_lmax);
#if IMA_CUDA
  if (!IMA_CUDA_OK(cudaGetLastError())) return lfail(IMA2P_E_CUDA, "kernel launch failed (joint terms)");
#endif
  if (!d2h(localmax_out, l.d_lmax, nvec * sizeof(double), s) || !dev_sync(s)) return lfail(IMA2P_E_CUDA, "download failed");
  return IMA2P_OK;
}

// The same two phases with everything left on the device, for callers that exchange between ranks with device collectives
// (NCCL all-gather of the local maxima, then of the records): nothing crosses PCIe and nothing synchronises between the phases,
// so batch after batch can be queued on one stream.
//   begin : terms of nvec (<= 32) vectors over the local rows; dev_localmax_out[nvec] (device) = maximum over the local rows
//   middle: dev_allmax[world][nvec] (device, the gathered local maxima) -> seed and global maximum on the device, prefixes,
//           scan, fold; dev_records_out[nvec][8] (device) = the six record fields, the global maximum, 0
// The caller gathers the records of all ranks and closes every vector with ima2p_lmode_joint_finish on their sums.
// up to kJointCallMax vectors per call; every kernel is launched once for all of them
constexpr int kJointCallMax = 512;
static int joint_wide_buffers(Lmode &l) {
  if (l.w_pbuf) return IMA2P_OK;
  const size_t G = (size_t)l.v.G, nchunks = (G + kRowsPerBlock - 1) / kRowsPerBlock;
  l.w_pbuf = l.alloc<double>((size_t)kJointCallMax * G);
  l.w_chunkmax = l.alloc<double>((size_t)kJointCallMax * nchunks);
  l.w_prefix = l.alloc<double>((size_t)kJointCallMax * nchunks);
  l.w_jpart = l.alloc<double>((size_t)kJointCallMax * nchunks * kJP);
  l.w_jout = l.alloc<double>((size_t)kJointCallMax * kJP);
  l.w_xs = l.alloc<double>((size_t)kJointCallMax * 2 * kMaxParams);
  l.d_seed = l.alloc<double>(kJointCallMax); l.d_gmax = l.alloc<double>(kJointCallMax); l.d_ltmp = l.alloc<double>(kJointCallMax);
  l.d_wlmax = l.alloc<double>(kJointCallMax); l.d_wrec = l.alloc<double>((size_t)kJointCallMax * 8);
  if (!l.w_pbuf || !l.w_chunkmax || !l.w_prefix || !l.w_jpart || !l.w_jout || !l.w_xs || !l.d_seed || !l.d_gmax || !l.d_ltmp || !l.d_wlmax || !l.d_wrec)
    return lfail(IMA2P_E_CUDA, "device allocation failed (jointp, wide batch)");
  return IMA2P_OK;
}
int ima2p_lmode_joint_begin(ima2p_lmode *h, const double *x, int nvec, double *dev_localmax_out, void *cuda_stream) {
  if (!h || !h->lm.d_cols || !x || nvec < 1 || nvec > kJointCallMax || !dev_localmax_out) return lfail(IMA2P_E_ARG, "joint_begin: bad argument (at most 512 vectors per call)");
  Lmode &l = h->lm;
  if (!lm_use(&l)) return lfail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = lm_stream(&l, cuda_stream);
  int rc = joint_wide_buffers(l);
  if (rc) return rc;
  const int np = l.v.nq + l.v.nm, nchunks = (int)((l.v.G + kRowsPerBlock - 1) / kRowsPerBlock);
  const std::vector<double> xs = joint_coefficients(l, x, nvec);
  if (!h2d(l.w_xs, xs.data(), xs.size() * sizeof(double), s)) return lfail(IMA2P_E_CUDA, "upload failed");
  // every buffer is [vector][...]: one launch of each kernel serves all vectors of the call
  IMA_LAUNCH(k_joint_terms, nchunks * ((nvec + kVT - 1) / kVT), kLmWarps, joint_terms_smem(np), s, l.v, (const double *)l.w_xs, nvec, nchunks, l.joint_model, l.w_pbuf, l.w_chunkmax);
  IMA_LAUNCH(k_joint_prefix, (nvec + kLmWarps - 1) / kLmWarps, kLmWarps, 0, s, (const double *)l.w_chunkmax, nchunks, nvec, (const double *)nullptr,
             l.w_prefix, dev_localmax_out);
#if IMA_CUDA
  if (!IMA_CUDA_OK(cudaGetLastError())) return lfail(IMA2P_E_CUDA, "kernel launch failed (joint begin)");
#endif
  return IMA2P_OK;
}

int ima2p_lmode_joint_middle(ima2p_lmode *h, int nvec, const double *dev_allmax, int world, int rank, long long global_row0,
                             double *dev_records_out, void *cuda_stream) {
  if (!h || !h->lm.d_cols || nvec < 1 || nvec > kJointCallMax || !dev_allmax || !dev_records_out || world < 1 || rank < 0 || rank >= world)
    return lfail(IMA2P_E_ARG, "joint_middle: bad argument");
  Lmode &l = h->lm;
  if (!lm_use(&l)) return lfail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = lm_stream(&l, cuda_stream);
  int rc = joint_wide_buffers(l);
  if (rc) return rc;
  const int nchunks = (int)((l.v.G + kRowsPerBlock - 1) / kRowsPerBlock);
  const int gall = (nvec + kLmWarps * IMA_WARP - 1) / (kLmWarps * IMA_WARP);
  IMA_LAUNCH(k_joint_seed, gall, kLmWarps, 0, s, dev_allmax, world, rank, nvec, l.d_seed, l.d_gmax);
  IMA_LAUNCH(k_joint_prefix, (nvec + kLmWarps - 1) / kLmWarps, kLmWarps, 0, s, (const double *)l.w_chunkmax, nchunks, nvec, (const double *)l.d_seed, l.w_prefix, l.d_ltmp);
  IMA_LAUNCH(k_joint_scan, (nchunks * nvec + kLmWarps - 1) / kLmWarps, kLmWarps, kLmWarps * 2 * IMA_WARP * sizeof(double), s, l.v, (const double *)l.w_pbuf, nvec, (const double *)l.w_prefix,
             (const double *)l.d_gmax, global_row0, (const double *)l.d_pow10, l.w_jpart);
  IMA_LAUNCH(k_joint_fold, (nvec + kLmWarps - 1) / kLmWarps, kLmWarps, 0, s, (const double *)l.w_jpart, nchunks, nvec, l.w_jout);
  IMA_LAUNCH(k_joint_pack, gall, kLmWarps, 0, s, (const double *)l.w_jout, (const double *)l.d_gmax, nvec, dev_records_out);
#if IMA_CUDA
  if (!IMA_CUDA_OK(cudaGetLastError())) return lfail(IMA2P_E_CUDA, "kernel launch failed (joint middle)");
#endif
  return IMA2P_OK;
}

// recompute the running-maximum prefixes of the batch evaluated by the last phase-1 call with a (new) seed: the
// maximum over the rows held by lower ranks, known only after the ranks have exchanged their local maxima
int ima2p_lmode_joint_reseed(ima2p_lmode *h, int nvec, const double *seed_before, double *localmax_out) {
  if (!h || !h->lm.d_cols || nvec < 1 || nvec > kJointVecMax || !localmax_out) return lfail(IMA2P_E_ARG, "joint_reseed: bad argument");
  Lmode &l = h->lm;
  if (!lm_use(&l)) return lfail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = lm_stream(&l, nullptr);
  const int nchunks = (int)((l.v.G + kRowsPerBlock - 1) / kRowsPerBlock);
  double *d_seed = nullptr;
  if (seed_before) { d_seed = l.d_jout; if (!h2d(d_seed, seed_before, nvec * sizeof(double), s)) return lfail(IMA2P_E_CUDA, "upload failed"); }
  IMA_LAUNCH(k_joint_prefix, (nvec + kLmWarps - 1) / kLmWarps, kLmWarps, 0, s, l.d_chunkmax, nchunks, nvec, d_seed, l.d_prefix, l.d_lmax);
  if (!d2h(localmax_out, l.d_lmax, nvec * sizeof(double), s) || !dev_sync(s)) return lfail(IMA2P_E_CUDA, "download failed");
  return IMA2P_OK;
}

int ima2p_lmode_joint_phase2(ima2p_lmode *h, int nvec, const double *globalmax, long long global_row0, double *records_out /* [nvec][6] */) {
  if (!h || !h->lm.d_cols || nvec < 1 || nvec > kJointVecMax || !globalmax || !records_out) return lfail(IMA2P_E_ARG, "joint_phase2: bad argument");
  Lmode &l = h->lm;
  if (!lm_use(&l)) return lfail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = lm_stream(&l, nullptr);
  const int nchunks = (int)((l.v.G + kRowsPerBlock - 1) / kRowsPerBlock);
  if (!h2d(l.d_lmax, globalmax, nvec * sizeof(double), s)) return lfail(IMA2P_E_CUDA, "upload failed");
  IMA_LAUNCH(k_joint_scan, (nchunks * nvec + kLmWarps - 1) / kLmWarps, kLmWarps, kLmWarps * 2 * IMA_WARP * sizeof(double), s, l.v, l.d_pbuf, nvec, l.d_prefix, l.d_lmax, global_row0, (const double *)l.d_pow10, l.d_jpart);
  IMA_LAUNCH(k_joint_fold, (nvec + kLmWarps - 1) / kLmWarps, kLmWarps, 0, s, l.d_jpart, nchunks, nvec, l.d_jout);
#if IMA_CUDA
  if (!IMA_CUDA_OK(cudaGetLastError())) return lfail(IMA2P_E_CUDA, "kernel launch failed (joint scan)");
#endif
  if (!d2h(records_out, l.d_jout, (size_t)nvec * kJP * sizeof(double), s) || !dev_sync(s)) return lfail(IMA2P_E_CUDA, "download failed");
  return IMA2P_OK;
}

// closing arithmetic of jointp (jointfind.cpp:1011-1046) from the folded record of one vector
void ima2p_lmode_joint_finish(const double *rec6, double globalmax, long long nrows_total, int calc_ess, double *q, double *ess) {
  double ins = rec6[0], kept = rec6[1], sum = rec6[2], sq = rec6[3];
  if (kept == ins && ins >= 2) { sum -= rec6[5]; sq -= rec6[5] * rec6[5]; }   // loop bound gi < iin drops the smallest kept term
  double m; int z;
  // same eexp as the device (host copy of utilities.cpp:1501-1539)
  {
    int n = (int)floor(globalmax / kLog2);
    double zr = 0.30102999566398119521 * (double)n;
    z = (int)zr; zr -= (double)z;
    double u = globalmax - (((double)n) * kLog2);
    double t = 1 + u * (1.0 + u * (0.5 + u * (0.16666666666666666666666666667 + u * (0.04166666666666666666666666667 +
               u * (0.00833333333333333333333333333 + u * (0.001388888888888888888888888889 + u * (0.000198412698412698412698412698 +
               u * (0.000024801587301587301587301587301 + u * (2.75573192239858906525573192239859e-6 + u * (2.75573192239858906525573192239e-7))))))))));
    m = t * pow(10.0, zr);
    if (fabs(m) > 10) { m /= 10.0; z += 1; }
    if (fabs(m) < 1) { m *= 10.0; z -= 1; }
  }
  const int maxz = z - 10;
  if (calc_ess && ess) *ess = sum * sum / sq;
  *q = log((double)nrows_total) - (log(sum) + maxz * 2.3025850929940456840);
}

// the records of joint_middle gathered from every rank, [world][nvec][8]: sums over ranks, the smallest kept term from the rank
// that holds it, then joint_finish per vector (host arithmetic on world x nvec small records)
void ima2p_lmode_joint_finish_gathered(const double *rec8, int world, int nvec, long long nrows_total, int calc_ess, double *q, double *ess) {
  for (int v = 0; v < nvec; v++) {
    double tot[6] = {0, 0, 0, 0, DBL_MAX, 0};
    for (int r = 0; r < world; r++) {
      const double *o = rec8 + ((size_t)r * nvec + v) * 8;
      tot[0] += o[0]; tot[1] += o[1]; tot[2] += o[2]; tot[3] += o[3];
      if (o[4] < tot[4]) { tot[4] = o[4]; tot[5] = o[5]; }
    }
    double e = 0.0;
    ima2p_lmode_joint_finish(tot, rec8[(size_t)v * 8 + 6], nrows_total, calc_ess, q + v, &e);
    if (ess) ess[v] = e;
  }
}

// nowmodeltype of findjointpeaks (jointfind.cpp:1104-1133): 0 = all parameters (two populations), 1 = the population sizes only,
// 2 = the migration rates only (the two searches of a three-population analysis); applies to every joint evaluation that follows.
// The entries of x outside the model's parameter range are not read.
int ima2p_lmode_set_joint_model(ima2p_lmode *h, int modeltype) {
  if (!h || modeltype < 0 || modeltype > 2) return lfail(IMA2P_E_ARG, "set_joint_model: model type is 0, 1 or 2");
  h->lm.joint_model = modeltype;
  return IMA2P_OK;
}

int ima2p_lmode_jointp(ima2p_lmode *h, const double *x, int nvec, int calc_ess, double *out_q, double *out_ess) {
  if (!h || !x || nvec < 1 || !out_q) return lfail(IMA2P_E_ARG, "jointp: bad argument");
  Lmode &l = h->lm;
  if (l.v.G != l.v.G_total) return lfail(IMA2P_E_ARG, "jointp: this handle holds a shard; use the two-phase form");
  const int np = l.v.nq + l.v.nm;
  // the device-resident form with a world of one: up to 512 vectors per pass, seven launches and one copy back per pass
  int rc = joint_wide_buffers(l);
  if (rc) return rc;
  stream_t s = lm_stream(&l, nullptr);
  std::vector<double> rec((size_t)kJointCallMax * 8);
  for (int v0 = 0; v0 < nvec; v0 += kJointCallMax) {
    const int nb = nvec - v0 < kJointCallMax ? nvec - v0 : kJointCallMax;
    if ((rc = ima2p_lmode_joint_begin(h, x + (size_t)v0 * np, nb, l.d_wlmax, nullptr))) return rc;
    if ((rc = ima2p_lmode_joint_middle(h, nb, l.d_wlmax, 1, 0, 0, l.d_wrec, nullptr))) return rc;
    if (!d2h(rec.data(), l.d_wrec, (size_t)nb * 8 * sizeof(double), s) || !dev_sync(s)) return lfail(IMA2P_E_CUDA, "download failed");
    for (int v = 0; v < nb; v++) {
      double e = 0.0;
      ima2p_lmode_joint_finish(rec.data() + (size_t)v * 8, rec[(size_t)v * 8 + 6], l.v.G_total, calc_ess, &out_q[v0 + v], &e);
      if (out_ess) out_ess[v0 + v] = e;
    }
  }
  return IMA2P_OK;
}

}  // extern "C"

namespace ima {
// ---- section 8 (f3) entry points ---------------------------------------------------------------------------------
static int lm_prepare_extra(Lmode &l) {
  if (l.d_pri) return IMA2P_OK;
  const int nlf = 100 * 5000 + 1;                      // logfact: same running sum as setlogfact (utilities.cpp:1405-1414)
  std::vector<double> lf(nlf);
  lf[0] = 0;
  for (int i = 1; i < nlf; i++) lf[i] = lf[i - 1] + log((double)i);
  LmPriors pr;
  for (int i = 0; i < kMaxParams; i++) { pr.q_max[i] = l.q_max[i]; pr.m_max[i] = l.m_max[i]; pr.m_mean[i] = l.m_mean[i]; }
  l.d_logfact = l.alloc<double>(nlf);
  l.d_err = l.alloc<int>(1);
  LmPriors *dp = l.alloc<LmPriors>(1);
  if (!l.d_logfact || !l.d_err || !dp) return lfail(IMA2P_E_CUDA, "device allocation failed");
  stream_t s = lm_stream(&l, nullptr);
  int zero = 0;
  if (!h2d(l.d_logfact, lf.data(), nlf * sizeof(double), s) || !h2d(dp, &pr, sizeof pr, s) || !h2d(l.d_err, &zero, sizeof zero, s) || !dev_sync(s))
    return lfail(IMA2P_E_CUDA, "upload failed");
  l.mc.logfact = l.d_logfact; l.mc.logfact_n = nlf; l.mc.err = l.d_err;
  l.d_pri = dp;
  return IMA2P_OK;
}
static int lm_check_err(Lmode &l, stream_t s, const char *what) {
  int code = 0;
  if (!d2h(&code, l.d_err, sizeof code, s) || !dev_sync(s)) return lfail(IMA2P_E_CUDA, "download failed");
  if (code) {
    int zero = 0;
    h2d(l.d_err, &zero, sizeof zero, s); dev_sync(s);
    return lfail(IMA2P_E_DEVICE, std::string(what) + ": device error word raised (14 = LogDiff a<=b, 15 = incomplete gamma, 16 = logfact range)");
  }
  return IMA2P_OK;
}
// uppergamma(0, x) = log E1(x) evaluated by the device routine (one lane)
static int lm_upper0(Lmode &l, double x, double *out) {
  stream_t s = lm_stream(&l, nullptr);
  if ((size_t)1 > l.cap_x) { l.d_x = l.alloc<double>(8); l.d_out = l.alloc<double>(8); l.cap_x = 8; }
  if (!l.d_x || !l.d_out) return lfail(IMA2P_E_CUDA, "device allocation failed");
  if (!h2d(l.d_x, &x, sizeof x, s)) return lfail(IMA2P_E_CUDA, "upload failed");
  IMA_LAUNCH(k_upper0, 1, 1, 0, s, l.mc, l.d_x, 1, l.d_out);
  if (!d2h(out, l.d_out, sizeof *out, s) || !dev_sync(s)) return lfail(IMA2P_E_CUDA, "download failed");
  return IMA2P_OK;
}
static bool lm_grow_partials(Lmode &l, size_t n) {
  if (n > l.cap_partials) { l.d_partials = l.alloc<double>(n); l.cap_partials = n; }
  return l.d_partials != nullptr;
}

}  // namespace ima
using namespace ima;
extern "C" {

// print_means_variances_correlations output.cpp:687-745: sums of calcx over every row, then the reference's finishing
// (means, variances = E[x^2] - mean^2, correlations of the p < q pairs; -1 marks a parameter whose prior maximum is ~0)
int ima2p_lmode_moments(ima2p_lmode *h, double *means, double *variances, double *correlations, double *raw_sums) {
  if (!h || !h->lm.d_cols || !means || !variances) return lfail(IMA2P_E_ARG, "moments: bad argument / rows not loaded");
  Lmode &l = h->lm;
  const int np = l.v.nq + l.v.nm, nacc = 2 * np + np * (np - 1) / 2;
  if (np > kMomentsMaxParams) return lfail(IMA2P_E_ARG, "moments: more than 32 parameters");
  if (!lm_use(&l)) return lfail(IMA2P_E_CUDA, "cudaSetDevice failed");
  int rc = lm_prepare_extra(l);
  if (rc) return rc;
  stream_t s = lm_stream(&l, nullptr);
  const int nchunks = (int)((l.v.G + kRowsPerBlock - 1) / kRowsPerBlock);
  if (!lm_grow_partials(l, (size_t)nchunks * nacc)) return lfail(IMA2P_E_CUDA, "device allocation failed");
  if (!l.d_msums) l.d_msums = l.alloc<double>(2 * kMomentsMaxParams + kMomentsMaxParams * (kMomentsMaxParams - 1) / 2);
  double *d_sums = l.d_msums;
  if (!d_sums) return lfail(IMA2P_E_CUDA, "device allocation failed");
  IMA_LAUNCH(k_moments, nchunks, kLmWarps, (size_t)kLmWarps * nacc * sizeof(double), s, l.v, l.mc, l.d_pri, l.d_partials);
  IMA_LAUNCH(k_reduce_partials, (nacc + kLmWarps * IMA_WARP - 1) / (kLmWarps * IMA_WARP), kLmWarps, 0, s, l.d_partials, nchunks, nacc, nacc, d_sums);
#if IMA_CUDA
  if (!IMA_CUDA_OK(cudaGetLastError())) return lfail(IMA2P_E_CUDA, "kernel launch failed (moments)");
#endif
  std::vector<double> sums(nacc);
  if (!d2h(sums.data(), d_sums, nacc * sizeof(double), s) || !dev_sync(s)) return lfail(IMA2P_E_CUDA, "download failed");
  if ((rc = lm_check_err(l, s, "moments"))) return rc;
  std::vector<double> raw((size_t)2 * np + (size_t)np * np, 0.0);
  for (int p = 0; p < np; p++) { raw[p] = sums[p]; raw[np + p] = sums[np + p]; }
  {
    int k = 2 * np;
    for (int p = 0; p < np - 1; p++) for (int q = p + 1; q < np; q++, k++) raw[2 * np + p * np + q] = sums[k];
  }
  if (raw_sums) for (size_t i = 0; i < raw.size(); i++) raw_sums[i] = raw[i];
  ima2p_lmode_moments_finish(np, raw.data(), l.v.G, means, variances, correlations);
  return IMA2P_OK;
}

// the closing arithmetic of print_means_variances_correlations (output.cpp:709-739) on row sums (this GPU's, or the sums
// over all ranks after an all-reduce): raw = {sum0[np], sum1[np], cross[np][np]}
void ima2p_lmode_moments_finish(int np, const double *raw, long long nrows_total, double *means, double *variances, double *correlations) {
  const double G = (double)nrows_total;
  for (int p = 0; p < np; p++) {
    means[p] = raw[p]; variances[p] = raw[np + p];
    if (means[p] >= 0.0) means[p] /= G;                                          // output.cpp:712-713
    if (variances[p] >= 0.0) { variances[p] /= G; variances[p] -= means[p] * means[p]; }   // :714-718
  }
  if (correlations) {
    for (int i = 0; i < np * np; i++) correlations[i] = 0.0;
    for (int p = 0; p < np - 1; p++)
      for (int q = p + 1; q < np; q++) {
        double c = raw[2 * np + p * np + q];
        if (c >= 0.0) { c /= G; c -= means[p] * means[q]; c /= sqrt(variances[p] * variances[q]); }   // :731-737
        else c = -1.0;
        correlations[p * np + q] = c;
      }
  }
}

// sums over rows [first, last) of the 2NM density terms; uniform prior: one pass; exponential prior: terms, maximum
// exponent, mantissa sum.  out[ix] = the row sum (uniform) or exp(log(acumm) + (maxz - OCUTOFF) LOG10) (exponential),
// not yet divided by the number of rows
static int lm_popmig_sums(Lmode &l, int thetai, int mi, const double *x, int nx, long long first, long long last, double *out) {
  if (thetai < 0 || thetai >= l.v.nq || mi < 0 || mi >= l.v.nm || first < 0 || last > l.v.G || first >= last || nx < 1)
    return lfail(IMA2P_E_ARG, "popmig: bad argument");
  if (!lm_use(&l)) return lfail(IMA2P_E_CUDA, "cudaSetDevice failed");
  int rc = lm_prepare_extra(l);
  if (rc) return rc;
  stream_t s = lm_stream(&l, nullptr);
  const long long nrows = last - first;
  const int nchunks = (int)((nrows + kRowsPerBlock - 1) / kRowsPerBlock);
  if ((size_t)nx > l.cap_x) { l.d_x = l.alloc<double>(nx); l.d_out = l.alloc<double>(((nx + kXT - 1) / kXT) * kXT); l.cap_x = nx; }
  if (!l.d_x || !l.d_out) return lfail(IMA2P_E_CUDA, "device allocation failed");
  if (!h2d(l.d_x, x, nx * sizeof(double), s)) return lfail(IMA2P_E_CUDA, "upload failed");
  if (!l.v.expoprior) {
    const int nxt = (nx + kXT - 1) / kXT, width = nxt * kXT;
    if (!lm_grow_partials(l, (size_t)nchunks * width)) return lfail(IMA2P_E_CUDA, "device allocation failed");
    IMA_LAUNCH(k_popmig, nchunks * nxt, kLmWarps, kLmWarps * kXT * sizeof(double), s, l.v, l.mc, l.d_pri, thetai, mi, l.d_x, nx, first, last, l.d_partials);
    IMA_LAUNCH(k_reduce_partials, (nx + kLmWarps * IMA_WARP - 1) / (kLmWarps * IMA_WARP), kLmWarps, 0, s, l.d_partials, nchunks, width, nx, l.d_out);
#if IMA_CUDA
    if (!IMA_CUDA_OK(cudaGetLastError())) return lfail(IMA2P_E_CUDA, "kernel launch failed (popmig)");
#endif
    if (!d2h(out, l.d_out, nx * sizeof(double), s) || !dev_sync(s)) return lfail(IMA2P_E_CUDA, "download failed");
  } else {
    // batches of kJointVecMax points reuse the joint-density term buffer ([kJointVecMax][rows])
    std::vector<double> zm((size_t)kJointVecMax * nchunks), mz(kJointVecMax), part((size_t)nchunks * kJointVecMax);
    if (!lm_grow_partials(l, (size_t)nchunks * kJointVecMax)) return lfail(IMA2P_E_CUDA, "device allocation failed");
    for (int x0 = 0; x0 < nx; x0 += kJointVecMax) {
      const int nb = nx - x0 < kJointVecMax ? nx - x0 : kJointVecMax;
      IMA_LAUNCH(k_expomig_terms, nchunks, kLmWarps, kLmWarps * kJointVecMax * sizeof(double), s, l.v, l.mc, l.d_pri, thetai, mi, l.d_x + x0, nb, first, last,
                 l.d_pbuf, l.d_chunkmax);
      if (!d2h(zm.data(), l.d_chunkmax, (size_t)nb * nchunks * sizeof(double), s) || !dev_sync(s)) return lfail(IMA2P_E_CUDA, "download failed");
      for (int i = 0; i < nb; i++) { double m = -1e300; for (int c = 0; c < nchunks; c++) m = zm[(size_t)i * nchunks + c] > m ? zm[(size_t)i * nchunks + c] : m; mz[i] = m; }
      if (!h2d(l.d_lmax, mz.data(), nb * sizeof(double), s)) return lfail(IMA2P_E_CUDA, "upload failed");
      IMA_LAUNCH(k_expomig_sum, nchunks, kLmWarps, kLmWarps * kJointVecMax * sizeof(double), s, l.d_pbuf, nb, nrows, l.d_lmax, l.d_partials);
#if IMA_CUDA
      if (!IMA_CUDA_OK(cudaGetLastError())) return lfail(IMA2P_E_CUDA, "kernel launch failed (expomig)");
#endif
      if (!d2h(part.data(), l.d_partials, (size_t)nchunks * kJointVecMax * sizeof(double), s) || !dev_sync(s)) return lfail(IMA2P_E_CUDA, "download failed");
      for (int i = 0; i < nb; i++) {
        double acumm = 0.0;
        for (int c = 0; c < nchunks; c++) acumm += part[(size_t)c * kJointVecMax + i];
        out[x0 + i] = log(acumm) + (mz[i] - 10) * 2.3025850929940456840;      // :162, LOG10
      }
    }
  }
  return lm_check_err(l, s, "popmig");
}

// row sums of the 2NM density terms over this GPU's rows [first, last) (uniform migration prior): additive over ranks, the
// caller all-reduces them and divides by the number of rows as calc_popmig / marginpopmig do
int ima2p_lmode_popmig_sums(ima2p_lmode *h, int thetai, int mi, const double *x, int nx, int first, int last, double *out) {
  if (!h || !h->lm.d_cols || !x || !out) return lfail(IMA2P_E_ARG, "popmig_sums: bad argument / rows not loaded");
  if (h->lm.v.expoprior) return lfail(IMA2P_E_ARG, "popmig_sums: with the exponential prior the terms are scaled by their largest exponent and are not additive");
  return lm_popmig_sums(h->lm, thetai, mi, x, nx, first, last, out);
}

// calc_popmig popmig.cpp:9-97 / calc_pop_expomig :101-170 (the choice follows the model's migration prior)
int ima2p_lmode_popmig(ima2p_lmode *h, int thetai, int mi, const double *x, int nx, int prob_or_like, double *out) {
  if (!h || !h->lm.d_cols || !x || !out) return lfail(IMA2P_E_ARG, "popmig: bad argument / rows not loaded");
  Lmode &l = h->lm;
  int rc = lm_popmig_sums(l, thetai, mi, x, nx, 0, l.v.G, out);
  if (rc) return rc;
  const double qmax = l.q_max[thetai];
  for (int i = 0; i < nx; i++) {
    double sum;
    if (!l.v.expoprior) {
      sum = out[i] / (double)l.v.G;
      if (prob_or_like) sum /= 2 * (log(qmax) + log(l.m_max[mi]) - log(2 * x[i])) / (qmax * l.m_max[mi]);
    } else {
      sum = exp(out[i] - log((double)l.v.G));
      if (prob_or_like) {
        double ug = 0.0;                                  // prior density of 2NM: 2 exp(uppergamma(0, 2x/(mmean qmax))) / (qmax mmean) (:166)
        if ((rc = lm_upper0(l, 2 * x[i] / (l.m_mean[mi] * qmax), &ug))) return rc;
        sum /= 2 * exp(ug) / (qmax * l.m_mean[mi]);
      }
    }
    out[i] = sum;
  }
  return IMA2P_OK;
}

// marginpopmig popmig.cpp:176-268 / marginpop_expomig :272-357: minus the mean over rows [firsttree, lasttree), with the
// reference's divisor and its OFFSCALEVAL = 1 outside the plotted range
int ima2p_lmode_marginpopmig(ima2p_lmode *h, int thetai, int mi, int firsttree, int lasttree, const double *x, int nx, double *out) {
  if (!h || !h->lm.d_cols || !x || !out) return lfail(IMA2P_E_ARG, "marginpopmig: bad argument / rows not loaded");
  Lmode &l = h->lm;
  int rc = lm_popmig_sums(l, thetai, mi, x, nx, firsttree, lasttree, out);
  if (rc) return rc;
  const double hi = l.v.expoprior ? 20 * l.m_mean[mi] : l.q_max[thetai] * l.m_max[mi] / 2.0;     // EXPOMIGPLOTSCALE imamp.hpp:156
  const double div = (double)lasttree - firsttree + (firsttree == 0);
  for (int i = 0; i < nx; i++) {
    if (x[i] < 0 || x[i] > hi) out[i] = 1.0;
    else if (!l.v.expoprior) out[i] = -(out[i] / div);
    else out[i] = -exp(out[i] - log(div));
  }
  return IMA2P_OK;
}

}  // extern "C"

extern "C" {
// gtpops / gtmig gtint.cpp:128-330 with the row thinning of print_greater_than_tests (:341-351, at most 20000 rows).
// kind 0: population sizes, 1: migration rates.  *out = -1 for the pairs the reference does not compute (i == j, different
// prior maxima, a migration maximum of ~0); migration rates under the exponential prior are not implemented there either.
int ima2p_lmode_greater_than(ima2p_lmode *h, int kind, int i, int j, double *out) {
  if (!h || !h->lm.d_cols || !out || (kind != 0 && kind != 1)) return lfail(IMA2P_E_ARG, "greater_than: bad argument / rows not loaded");
  Lmode &l = h->lm;
  const int n = kind == 0 ? l.v.nq : l.v.nm;
  if (i < 0 || j < 0 || i >= n || j >= n) return lfail(IMA2P_E_ARG, "greater_than: bad parameter index");
  if (kind == 1 && l.v.expoprior) return lfail(IMA2P_E_ARG, "greater_than: not defined for migration rates with exponential priors (gtint.cpp:404)");
  const double *mx = kind == 0 ? l.q_max : l.m_max;
  if (i == j || mx[i] != mx[j] || (kind == 1 && !(mx[i] > kMinParamVal && mx[j] > kMinParamVal))) { *out = -1.0; return IMA2P_OK; }
  if (!lm_use(&l)) return lfail(IMA2P_E_CUDA, "cudaSetDevice failed");
  int rc = lm_prepare_extra(l);
  if (rc) return rc;
  stream_t s = lm_stream(&l, nullptr);
  const long long G = l.v.G;
  int treeinc = 1; long long nused = G;
  if (G > 20000) { treeinc = (int)(G / 20000); nused = 20000; }      // USETREESMAX gtint.cpp:24
  const int rowsper = kLmWarps * IMA_WARP, nchunks = (int)((nused + rowsper - 1) / rowsper);
  if (!lm_grow_partials(l, (size_t)nchunks)) return lfail(IMA2P_E_CUDA, "device allocation failed");
  if (l.cap_x < 1) { l.d_x = l.alloc<double>(8); l.d_out = l.alloc<double>(8); l.cap_x = 8; }
  if (!l.d_out) return lfail(IMA2P_E_CUDA, "device allocation failed");
  IMA_LAUNCH(k_greater_than, nchunks, kLmWarps, kLmWarps * sizeof(double), s, l.v, l.mc, l.d_pri, kind, i, j, treeinc, (int)nused, l.d_partials);
  IMA_LAUNCH(k_reduce_partials, 1, kLmWarps, 0, s, l.d_partials, nchunks, 1, 1, l.d_out);
#if IMA_CUDA
  if (!IMA_CUDA_OK(cudaGetLastError())) return lfail(IMA2P_E_CUDA, "kernel launch failed (greater_than)");
#endif
  double sum = 0.0;
  if (!d2h(&sum, l.d_out, sizeof sum, s) || !dev_sync(s)) return lfail(IMA2P_E_CUDA, "download failed");
  if ((rc = lm_check_err(l, s, "greater_than"))) return rc;
  *out = sum / (double)nused;
  return IMA2P_OK;
}
// ---- measured FP64 peaks of the device (the L-mode evaluators and the prior sweep are bound by FP64
// arithmetic and by exp, not by HBM; their fractions are quoted against what this GPU does, measured here) ---------------
}  // extern "C"
namespace ima {
// 8 independent fused multiply-add chains per thread (fma() is exempt from --fmad=false: it is asked for)
IMA_KERNEL void k_peak_fma(double *out, int iters) {
  double a[8];
  const double x = 1.0 + 1e-9 * (ima_block() + 1), y = 1e-12 * (Warp::lane() + 1);
  for (int k = 0; k < 8; k++) a[k] = 1.0 + k;
  for (int i = 0; i < iters; i++)
    for (int k = 0; k < 8; k++) a[k] = fma(a[k], x, y);
  double s = 0.0;
  for (int k = 0; k < 8; k++) s += a[k];
  if (s == 123.456) out[0] = s;                      // never true: keeps the loop alive
}
// 4 independent exp evaluations per thread and iteration
IMA_KERNEL void k_peak_exp(double *out, int iters) {
  double a[4];
  for (int k = 0; k < 4; k++) a[k] = -1e-3 * (k + 1 + Warp::lane());
  for (int i = 0; i < iters; i++)
    for (int k = 0; k < 4; k++) a[k] = exp(a[k]) - 1.0001;
  double s = 0.0;
  for (int k = 0; k < 4; k++) s += a[k];
  if (s == 123.456) out[0] = s;
}
}  // namespace ima
extern "C" {
// out[0] = FP64 fused multiply-adds per second (x 2 = flop/s), out[1] = FP64 exp evaluations per second, whole device
int ima2p_debug_fp64_peaks(int device, double *out2) {
  if (!out2) return lfail(IMA2P_E_ARG, "fp64_peaks: bad argument");
  out2[0] = out2[1] = 0.0;
#if IMA_CUDA
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return lfail(IMA2P_E_CUDA, "no CUDA device: ima2p_b200 has no CPU path");
  if (!IMA_CUDA_OK(cudaSetDevice(device))) return lfail(IMA2P_E_CUDA, "cudaSetDevice failed");
  cudaDeviceProp prop;
  if (!IMA_CUDA_OK(cudaGetDeviceProperties(&prop, device))) return lfail(IMA2P_E_CUDA, "cudaGetDeviceProperties failed");
  double *d = (double *)dev_alloc(8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = prop.multiProcessorCount * 8, warps = 8, it_fma = 20000, it_exp = 2000;
  float ms = 0.f;
  IMA_LAUNCH(k_peak_fma, blocks, warps, 0, 0, d, 100);
  cudaEventRecord(e0, 0);
  IMA_LAUNCH(k_peak_fma, blocks, warps, 0, 0, d, it_fma);
  cudaEventRecord(e1, 0);
  cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1);
  out2[0] = (double)blocks * warps * 32 * 8.0 * it_fma / (ms * 1e-3);
  IMA_LAUNCH(k_peak_exp, blocks, warps, 0, 0, d, 100);
  cudaEventRecord(e0, 0);
  IMA_LAUNCH(k_peak_exp, blocks, warps, 0, 0, d, it_exp);
  cudaEventRecord(e1, 0);
  cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1);
  out2[1] = (double)blocks * warps * 32 * 4.0 * it_exp / (ms * 1e-3);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  dev_free(d);
  return IMA2P_OK;
#else
  (void)device;
  return lfail(IMA2P_E_UNSUPPORTED, "fp64_peaks: needs the device");
#endif
}
}  // extern "C"
