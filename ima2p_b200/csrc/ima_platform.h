// Platform layer for the IMa2p B200 engine kernels.
//
// The product is CUDA for sm_100a: one warp works on one (chain, locus) pair, its genealogy staged in
// shared memory.  The same kernel bodies can also be compiled by a plain C++ compiler with
// -DIMA_HOSTEMU (a "warp" of ONE lane, blocks run in a loop).  That build exists ONLY so that the
// CPU-side test-suite can exercise the kernel logic against the oracle where there is no GPU
// (tests/hostemu/); the package never loads it and it is not a fallback: libima2p_b200.so always
// requires a CUDA device.
#pragma once
#include <stdint.h>
#include <math.h>
#include <float.h>
#include <string.h>

#if defined(__CUDACC__) && !defined(IMA_HOSTEMU)
#define IMA_CUDA 1
#define IMA_HD __host__ __device__ __forceinline__
#define IMA_DEV __device__ __forceinline__
#define IMA_DEV_NOINLINE __device__ __noinline__
#define IMA_WARP 32
#else
#define IMA_CUDA 0
#define IMA_HD inline
#define IMA_DEV inline
#define IMA_DEV_NOINLINE inline
#define IMA_WARP 1
#endif

namespace ima {

// ---- limits (reference: imamp.hpp:117-173) ----
constexpr int kMaxPops = 10;             // MAXPOPS
constexpr int kMaxPeriods = kMaxPops + 1;
constexpr int kMaxTreePops = 2 * kMaxPops - 1;
constexpr int kMaxParams = 64;           // population-size or migration parameters per model
constexpr int kMaxWp = 12;               // weight positions summed per parameter
constexpr int kMaxLinked = 4;            // linked stepwise portions kept on device (reference: MAXLINKED 15)
constexpr double kTimeMax = 1000000.0;   // TIMEMAX
constexpr double kRejectIS = -1000000000.0;      // REJECTINFINITESITESCONSTANT
constexpr double kLog2 = 0.69314718055994530941723212146;
constexpr double kLogDblMax = 7.0978271289338397e+02;
constexpr double kMyDblMax = DBL_MAX / 1e10;
constexpr double kMPriorMin = 0.000001;
constexpr double kMigCloseFrac = 0.9;    // update_gtree_common.cpp:23
constexpr double kSlideStdvMax = 20.0;   // update_gtree.cpp:687
constexpr int kAddMigMax = 1000;         // ADDMIGMAX
constexpr int kSwapDist = 7;             // swapchains.cpp:220

enum MutModel { kInfiniteSites = 0, kHKY = 1, kStepwise = 2, kJointISSW = 3 };
// linked parts [sw_first, nlinked) of a locus are stepwise: all of them (S), all but part 0 (J: part 0 is infinite sites), none
IMA_HD bool has_stepwise(int model) { return model == kStepwise || model == kJointISSW; }
IMA_HD bool has_infinite_sites(int model) { return model == kInfiniteSites || model == kJointISSW; }
IMA_HD int sw_first(int model) { return model == kJointISSW ? 1 : 0; }

// per-pair proposal flags
enum : uint32_t {
  kFlagRejectIS = 1u,        // infinite-sites incompatibility => reject (update_gtree.cpp:855)
  kFlagOverflow = 2u,        // migration pool capacity exceeded => proposal dropped (counted)
  kFlagTopol = 4u,
  kFlagTmrca = 8u,
  kFlagBadTree = 16u         // internal consistency check failed (never expected)
};

// ---- warp abstraction ----
struct Warp {
#if IMA_CUDA
  IMA_DEV static int lane() { return threadIdx.x & 31; }
  IMA_DEV static void sync() { __syncwarp(); }
  IMA_DEV static double sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
  IMA_DEV static int sum(int v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
  // several sums at once: the shuffle chains are independent and overlap
  IMA_DEV static void sum2(double &a, double &b) {
    for (int o = 16; o > 0; o >>= 1) {
      const double ta = __shfl_xor_sync(0xffffffffu, a, o), tb = __shfl_xor_sync(0xffffffffu, b, o);
      a += ta; b += tb;
    }
  }
  IMA_DEV static void sum4(double &a, double &b, double &c, double &d) {
    for (int o = 16; o > 0; o >>= 1) {
      const double ta = __shfl_xor_sync(0xffffffffu, a, o), tb = __shfl_xor_sync(0xffffffffu, b, o);
      const double tc = __shfl_xor_sync(0xffffffffu, c, o), td = __shfl_xor_sync(0xffffffffu, d, o);
      a += ta; b += tb; c += tc; d += td;
    }
  }
  IMA_DEV static unsigned long long sum(unsigned long long v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
  IMA_DEV static int max(int v) {
    for (int o = 16; o > 0; o >>= 1) { int t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
    return v;
  }
  IMA_DEV static int bcast(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
  IMA_DEV static double bcast(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
  IMA_DEV static bool any(bool p) { return __any_sync(0xffffffffu, p) != 0; }
  IMA_DEV static unsigned ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
  IMA_DEV static int popc(unsigned m) { return __popc(m); }
  // inclusive prefix sum over the 32 lanes
  IMA_DEV static int scan(int v) {
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, v, o); if (lane() >= o) v += t; }
    return v;
  }
  IMA_DEV static double scan_add(double v) {
    for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(0xffffffffu, v, o); if (lane() >= o) v += t; }
    return v;
  }
  IMA_DEV static double scan_mul(double v) {
    for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(0xffffffffu, v, o); if (lane() >= o) v *= t; }
    return v;
  }
  IMA_DEV static double shfl_up(double v, int o) { return __shfl_up_sync(0xffffffffu, v, o); }
  // index of the first lane whose predicate holds, or -1
  IMA_DEV static int first(bool p) { unsigned m = __ballot_sync(0xffffffffu, p); return m ? __ffs(m) - 1 : -1; }
#else
  static int lane() { return 0; }
  static void sync() {}
  static double sum(double v) { return v; }
  static int sum(int v) { return v; }
  static unsigned long long sum(unsigned long long v) { return v; }
  static void sum2(double &, double &) {}
  static void sum4(double &, double &, double &, double &) {}
  static int max(int v) { return v; }
  static int bcast(int v, int) { return v; }
  static double bcast(double v, int) { return v; }
  static bool any(bool p) { return p; }
  static unsigned ballot(bool p) { return p ? 1u : 0u; }
  static int popc(unsigned m) { int n = 0; while (m) { n += (int)(m & 1u); m >>= 1; } return n; }
  static int scan(int v) { return v; }
  static double scan_add(double v) { return v; }
  static double scan_mul(double v) { return v; }
  static double shfl_up(double v, int) { return v; }
  static int first(bool p) { return p ? 0 : -1; }
#endif
};

// ---- Philox4x32-10 counter-based generator (Salmon et al. 2011).  One stream per
// (global chain, locus, step, purpose): results do not depend on how chains are sharded over GPUs. ----
struct Philox {
  uint32_t key[2];
  uint32_t ctr[4];
  uint32_t out[4];
  int have;

  IMA_HD static void mulhilo(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo) {
    uint64_t p = (uint64_t)a * (uint64_t)b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
  }
  IMA_HD void init(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2) {
    key[0] = (uint32_t)seed;
    key[1] = (uint32_t)(seed >> 32);
    ctr[0] = 0; ctr[1] = c0; ctr[2] = c1; ctr[3] = c2;
    have = 0;
  }
  IMA_HD void refill() {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
#pragma unroll
    for (int r = 0; r < 10; r++) {
      uint32_t hi0, lo0, hi1, lo1;
      mulhilo(0xD2511F53u, c0, hi0, lo0);
      mulhilo(0xCD9E8D57u, c2, hi1, lo1);
      uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    ctr[0]++;
    have = 4;
  }
  IMA_HD uint32_t next32() {
    if (have == 0) refill();
    return out[4 - (have--)];
  }
  // uniform on the open interval (0,1) with 32-bit resolution, exactly the grid of the reference's
  // genrand_real3 (utilities.cpp:418 -> ((double)genrand_int32() + 0.5) / 2^32)
  IMA_HD double uniform() { return ((double)next32() + 0.5) * (1.0 / 4294967296.0); }
  IMA_HD int randint(int n) { int v = (int)floor(uniform() * n); return v < n ? v : n - 1; }  // randposint :433
  IMA_HD int bit() { return (int)(next32() >> 31); }                                           // bitran :443
  // normdev utilities.cpp:472-500 (polar Box-Muller; the cached second deviate is not kept)
  IMA_HD double normal(double mean, double stdev) {
    double v1, v2, rsq;
    do {
      v1 = 2.0 * uniform() - 1.0;
      v2 = 2.0 * uniform() - 1.0;
      rsq = v1 * v1 + v2 * v2;
    } while (rsq >= 1.0 || rsq == 0.0);
    double fac = sqrt(-2.0 * log(rsq) / rsq);
    return v2 * fac * stdev + mean;
  }
};

enum RngPurpose : uint32_t { kRngPropose = 1, kRngAccept = 2, kRngSwap = 3, kRngAlleles = 4, kRngSplitTime = 5, kRngScalars = 6, kRngSplitMig = 7, kRngSplitMigSim = 8 };

}  // namespace ima
