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

struct JointXs { double log2diffx[kMaxParams], logx[kMaxParams], divx[kMaxParams], x[kMaxParams]; };

// p_g of jointp (jointfind.cpp:971-996, two populations / full model): one thread per row, all vectors of the batch
IMA_KERNEL void k_joint_terms(LmView V, const JointXs *xs, int nvec, double *pbuf, double *chunkmax) {
  IMA_SMEM_DECL
  const int lane = Warp::lane(), warp = ima_warp_in_block();
  const int chunk = ima_block();
  const long long r0 = (long long)chunk * kRowsPerBlock;
  long long r1 = r0 + kRowsPerBlock;
  if (r1 > V.G) r1 = V.G;
  double *sm = (double *)IMA_SMEM;     // [kLmWarps][kJointVecMax]
  double vmax[kJointVecMax];
  for (int v = 0; v < nvec; v++) vmax[v] = -DBL_MAX;
  const int np = V.nq + V.nm;
  for (long long r = r0 + warp * IMA_WARP + lane; r < r1; r += kLmWarps * IMA_WARP) {
    float g[2 * kMaxParams];
    // columns are read once per row and reused for every vector of the batch
    const double probg = V.cols[(size_t)V.probgp * V.G + r];
    for (int i = 0; i < V.nq; i++) { g[3 * i] = V.cols[(size_t)(V.ccp + i) * V.G + r]; g[3 * i + 1] = V.cols[(size_t)(V.hccp + i) * V.G + r]; g[3 * i + 2] = V.cols[(size_t)(V.fcp + i) * V.G + r]; }
    for (int i = 0; i < V.nm; i++) { g[3 * V.nq + 2 * i] = V.cols[(size_t)(V.mcp + i) * V.G + r]; g[3 * V.nq + 2 * i + 1] = V.cols[(size_t)(V.fmp + i) * V.G + r]; }
    for (int v = 0; v < nvec; v++) {
      const JointXs &X = xs[v];
      double p = -probg;
      for (int i = 0; i < np; i++) {
        if (i < V.nq) p += g[3 * i] * X.log2diffx[i] - g[3 * i + 1] - (2.0 * g[3 * i + 2]) * X.divx[i];
        else { const int i1 = i - V.nq; p += g[3 * V.nq + 2 * i1] * X.logx[i] - g[3 * V.nq + 2 * i1 + 1] * X.x[i]; }
      }
      pbuf[(size_t)v * V.G + r] = p;
      if (p > vmax[v]) vmax[v] = p;
    }
  }
  for (int v = 0; v < nvec; v++) {
    double m = vmax[v];
#if IMA_CUDA
    for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, m, o); m = t > m ? t : m; }
#endif
    if (lane == 0) sm[warp * kJointVecMax + v] = m;
  }
#if IMA_CUDA
  __syncthreads();
  if ((int)threadIdx.x < nvec) {
    double m = -DBL_MAX;
    for (int w = 0; w < kLmWarps; w++) m = sm[w * kJointVecMax + threadIdx.x] > m ? sm[w * kJointVecMax + threadIdx.x] : m;
    chunkmax[(size_t)threadIdx.x * gridDim.x + chunk] = m;
  }
#else
  if (warp == kLmWarps - 1)
    for (int v = 0; v < nvec; v++) {
      double m = -DBL_MAX;
      for (int w = 0; w < kLmWarps; w++) m = sm[w * kJointVecMax + v] > m ? sm[w * kJointVecMax + v] : m;
      chunkmax[(size_t)v * ((V.G + kRowsPerBlock - 1) / kRowsPerBlock) + chunk] = m;
    }
#endif
}

// per vector: exclusive prefix maxima of the chunks (seeded with the maximum of the rows held by earlier ranks)
// and the maximum over this rank's rows; one thread per vector (a few hundred chunks)
IMA_KERNEL void k_joint_prefix(const double *chunkmax, int nchunks, int nvec, const double *seed_before, double *chunkprefix, double *localmax) {
  const int v = ima_block() * kLmWarps * IMA_WARP + ima_warp_in_block() * IMA_WARP + Warp::lane();
  if (v >= nvec) return;
  double run = seed_before ? seed_before[v] : -DBL_MAX;
  for (int c = 0; c < nchunks; c++) {
    chunkprefix[(size_t)v * nchunks + c] = run;
    const double m = chunkmax[(size_t)v * nchunks + c];
    if (m > run) run = m;
  }
  localmax[v] = run;      // includes the seed
}

// partial record of one chunk: inserted count, kept count, sum, sum of squares, smallest kept p, its scaled term
constexpr int kJP = 6;

IMA_KERNEL void k_joint_scan(LmView V, const double *pbuf, int nvec, const double *chunkprefix, const double *globalmax,
                             long long global_row0, double *partials) {
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
  for (long long base = r0; base < r1; base += IMA_WARP) {
    const long long r = base + lane;
    const bool valid = r < r1;
    const double p = valid ? pb[r] : -DBL_MAX;
    // running maximum of the rows before r: prefix max within the group of 32, then the carried maximum
    double pre = p;
#if IMA_CUDA
    for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= o && t > pre) pre = t; }
    double before = __shfl_up_sync(0xffffffffu, pre, 1);
    if (lane == 0) before = -DBL_MAX;
    const double groupmax = __shfl_sync(0xffffffffu, pre, 31);
#else
    double before = -DBL_MAX;
    const double groupmax = pre;
#endif
    if (run > before) before = run;
    if (valid) {
      const bool first_row = (global_row0 + r == 0);
      if (first_row || before - p < 10) {            // :1005 (row 0 is always the list head :998-1003)
        inserted += 1.0;
        if (gmax - p < 10) {                         // :1022
          double m; int z;
          eexp(p, m, z);
          const int zadj = z - maxz;
          double term = (zadj > -308 && zadj < 308) ? m * pow(10.0, (double)zadj) : (zadj <= -308 ? 0.0 : DBL_MAX);
          kept += 1.0; sum += term; sumsq += term * term;
          if (p < minp) { minp = p; minterm = term; }
        }
      }
    }
    if (groupmax > run) run = groupmax;
  }
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

// fold the chunk records of each vector in chunk order
IMA_KERNEL void k_joint_fold(const double *partials, int nchunks, int nvec, double *out) {
  const int v = ima_block() * kLmWarps * IMA_WARP + ima_warp_in_block() * IMA_WARP + Warp::lane();
  if (v >= nvec) return;
  double ins = 0, kept = 0, sum = 0, sq = 0, minp = DBL_MAX, minterm = 0;
  for (int c = 0; c < nchunks; c++) {
    const double *o = partials + ((size_t)v * nchunks + c) * kJP;
    ins += o[0]; kept += o[1]; sum += o[2]; sq += o[3];
    if (o[4] < minp) { minp = o[4]; minterm = o[5]; }
  }
  double *r = out + (size_t)v * kJP;
  r[0] = ins; r[1] = kept; r[2] = sum; r[3] = sq; r[4] = minp; r[5] = minterm;
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
         *d_lmax = nullptr, *d_jpart = nullptr, *d_jout = nullptr;
  JointXs *d_xs = nullptr;
  size_t cap_x = 0, cap_partials = 0;
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
  l.d_xs = l.alloc<JointXs>(kJointVecMax);
  if (!l.d_cols || !l.d_pbuf || !l.d_xs) return lfail(IMA2P_E_CUDA, "device allocation failed");
  stream_t s = lm_stream(&l, nullptr);
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

// two-phase joint evaluation, also the building block of the multi-GPU form:
//   phase 1: terms + this rank's maximum per vector (seeded with the maximum of the rows of earlier ranks)
//   phase 2: keep-set records given the global maximum
int ima2p_lmode_joint_phase1(ima2p_lmode *h, const double *x, int nvec, const double *seed_before, double *localmax_out) {
  if (!h || !h->lm.d_cols || !x || nvec < 1 || nvec > kJointVecMax) return lfail(IMA2P_E_ARG, "joint_phase1: bad argument");
  Lmode &l = h->lm;
  if (!lm_use(&l)) return lfail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = lm_stream(&l, nullptr);
  const int np = l.v.nq + l.v.nm, nchunks = (int)((l.v.G + kRowsPerBlock - 1) / kRowsPerBlock);
  std::vector<JointXs> xs(nvec);
  for (int v = 0; v < nvec; v++)
    for (int i = 0; i < np; i++) {                     // jointfind.cpp:955-970
      const double xv = x[(size_t)v * np + i];
      xs[v].x[i] = xv; xs[v].logx[i] = log(xv); xs[v].divx[i] = 1.0 / xv;
      if (i < l.v.nq) xs[v].log2diffx[i] = kLog2 - log(xv);
    }
  double *d_seed = nullptr;
  if (seed_before) { d_seed = l.d_jout; if (!h2d(d_seed, seed_before, nvec * sizeof(double), s)) return lfail(IMA2P_E_CUDA, "upload failed"); }
  if (!h2d(l.d_xs, xs.data(), nvec * sizeof(JointXs), s)) return lfail(IMA2P_E_CUDA, "upload failed");
  IMA_LAUNCH(k_joint_terms, nchunks, kLmWarps, kLmWarps * kJointVecMax * sizeof(double), s, l.v, l.d_xs, nvec, l.d_pbuf, l.d_chunkmax);
  IMA_LAUNCH(k_joint_prefix, (nvec + kLmWarps * IMA_WARP - 1) / (kLmWarps * IMA_WARP), kLmWarps, 0, s, l.d_chunkmax, nchunks, nvec, d_seed, l.d_prefix, l.d_lmax);
#if IMA_CUDA
  if (!IMA_CUDA_OK(cudaGetLastError())) return lfail(IMA2P_E_CUDA, "kernel launch failed (joint terms)");
#endif
  if (!d2h(localmax_out, l.d_lmax, nvec * sizeof(double), s) || !dev_sync(s)) return lfail(IMA2P_E_CUDA, "download failed");
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
  IMA_LAUNCH(k_joint_prefix, (nvec + kLmWarps * IMA_WARP - 1) / (kLmWarps * IMA_WARP), kLmWarps, 0, s, l.d_chunkmax, nchunks, nvec, d_seed, l.d_prefix, l.d_lmax);
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
  IMA_LAUNCH(k_joint_scan, (nchunks * nvec + kLmWarps - 1) / kLmWarps, kLmWarps, 0, s, l.v, l.d_pbuf, nvec, l.d_prefix, l.d_lmax, global_row0, l.d_jpart);
  IMA_LAUNCH(k_joint_fold, (nvec + kLmWarps * IMA_WARP - 1) / (kLmWarps * IMA_WARP), kLmWarps, 0, s, l.d_jpart, nchunks, nvec, l.d_jout);
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

int ima2p_lmode_jointp(ima2p_lmode *h, const double *x, int nvec, int calc_ess, double *out_q, double *out_ess) {
  if (!h || !x || nvec < 1 || !out_q) return lfail(IMA2P_E_ARG, "jointp: bad argument");
  Lmode &l = h->lm;
  if (l.v.G != l.v.G_total) return lfail(IMA2P_E_ARG, "jointp: this handle holds a shard; use the two-phase form");
  const int np = l.v.nq + l.v.nm;
  for (int v0 = 0; v0 < nvec; v0 += kJointVecMax) {
    const int nb = nvec - v0 < kJointVecMax ? nvec - v0 : kJointVecMax;
    double lmax[kJointVecMax], rec[kJointVecMax * kJP];
    int rc = ima2p_lmode_joint_phase1(h, x + (size_t)v0 * np, nb, nullptr, lmax);
    if (rc) return rc;
    rc = ima2p_lmode_joint_phase2(h, nb, lmax, 0, rec);
    if (rc) return rc;
    for (int v = 0; v < nb; v++) {
      double e = 0.0;
      ima2p_lmode_joint_finish(rec + (size_t)v * kJP, lmax[v], l.v.G_total, calc_ess, &out_q[v0 + v], &e);
      if (out_ess) out_ess[v0 + v] = e;
    }
  }
  return IMA2P_OK;
}

}  // extern "C"
