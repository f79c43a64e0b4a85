// Device-resident tables and state layout of the engine (see DESIGN.md "Data layout in HBM").
#pragma once
#include "ima_platform.h"
#include "ima_math.h"

namespace ima {

// Chain-invariant model tables: population tree, per-period population lists, parameter ->
// weight-position lists and priors (what setup_iparams builds, initialize.cpp:201-727).  Lives in
// __constant__ memory on the device.
struct DevModel {
  int npops, nsplit, ntreepops, rootpop;
  int nq, nm, ncc, nmc;                 // NI = ncc + nmc ints, ND = 2*ncc + nmc doubles per weight record
  int nomigration, expoprior, thermo;
  double gbeta;
  signed char plist[kMaxPops][kMaxPops];
  signed char addpop[kMaxPeriods], droppops[kMaxPeriods][2];
  signed char pt_e[kMaxTreePops], pt_down[kMaxTreePops];
  int desc_mask[kMaxTreePops];          // bit q set: tree population q is this population or one of its descendants
  short cc_off[kMaxPeriods + 1], mc_off[kMaxPeriods + 1];
  signed char q_n[kMaxParams], m_n[kMaxParams];
  short q_idx[kMaxParams][kMaxWp], m_idx[kMaxParams][kMaxWp];
  double q_max[kMaxParams], q_min[kMaxParams];
  double m_max[kMaxParams], m_min[kMaxParams], m_mean[kMaxParams];
  int nomig_n;
  short nomig_idx[kMaxParams];
};

// layout of EngineDims::tab
constexpr int kTabDesc = 0, kTabCcOff = 20, kTabMcOff = 32, kTabPlist = 44, kTabInts = 44 + kMaxPops * kMaxPops;

// weight record layout: ints  [cc(ncc) | mc(nmc)],  doubles [fc(ncc) | hcc(ncc) | fm(nmc)]
IMA_HD int wi_cc(const DevModel &M, int k, int i) { return M.cc_off[k] + i; }
IMA_HD int wi_mc(const DevModel &M, int k, int i, int j) { return M.ncc + M.mc_off[k] + i * (M.npops - k) + j; }
IMA_HD int wd_fc(const DevModel &M, int k, int i) { return M.cc_off[k] + i; }
IMA_HD int wd_hcc(const DevModel &M, int k, int i) { return M.ncc + M.cc_off[k] + i; }
IMA_HD int wd_fm(const DevModel &M, int k, int i, int j) { return 2 * M.ncc + M.mc_off[k] + i * (M.npops - k) + j; }

// Per-locus read-only data, shared by all chains (struct locus, imamp.hpp:894-936).
struct DevLocus {
  int model, ng, nl, nsites, nwords, totsites, nlinked;
  int samppop[kMaxPops];
  int minA[kMaxLinked], maxA[kMaxLinked];
  double hval, sumlogk;
  double hlog, h2term;      // log(hval) and 1 / (2 hval), made once by set_locus
  long long sitemask_off;   // uint32 [nsites][nwords]: carrier-tip bit masks of the segregating sites (IS)
  long long seq_off;        // uint8  [ng][nsites]: bases 0..3 of the compressed site patterns (HKY)
  long long mult_off;       // int    [nsites]: pattern multiplicities (HKY)
};

struct alignas(8) short4_t { short x, y, z, w; };       // up0, up1, down, pop (one 64-bit load)
struct alignas(4) ushort2_t { unsigned short x, y; };   // migration segment (start, count) in the pair's pool

// One of the two state buffers.  Pair-major arrays: pair p = local_chain * nloci + locus.
struct PairBuf {
  short4_t *topo;       // [P][NL]
  double *time;         // [P][NL]   time at the bottom of the edge (root: TIMEMAX)
  ushort2_t *mseg;      // [P][NL]
  double *mig_t;        // [P][CAP]
  short *mig_p;         // [P][CAP]
  double *sd;           // [P][4]    roottime, length, tlength, pdg
  int *si;              // [P][2]    root, mignum
  int *gwi;             // [P][NI]
  double *gwd;          // [P][ND]
  short *A;             // [P][kMaxLinked][NL]   stepwise allele states (only when some locus is stepwise)
  double *dlikeA;       // [P][kMaxLinked][NL]
  double *pdg_a;        // [P][kMaxLinked]
  uint32_t *hky_mask;   // [P][hky_mask_words] which of its two slots holds every internal node's HKY partials in THIS buffer's genealogy
};

struct EngineDims {
  int nchains;          // chains held by this GPU
  int nchains_global;   // chains over all ranks
  int chain0;           // global index of local chain 0
  int nloci, P;
  int NT;               // prior terms: size parameters + migration parameters
  int NL, CAP, NI, ND, EVP, W, S, W64;     // maxima over loci: numlines, pool capacity, record sizes, event slots, mask words, sites
  int FP, FC, FEV, FS;  // the small tables of the two-kernel proposal path (ima_fastpath.h): pool entries per pair in k_move,
                        // migration events per genealogy and event slots in k_weigh, scratch entries of the fast split-time kernel;
                        // a pair that needs more takes the general path
  int any_sw, any_hky;
  long long hky_stride; // doubles of HKY partials per pair: (max genes - 1) nodes * 2 slots * 5 * hky_sites
  int hky_sites, hky_mask_words;   // largest number of compressed site patterns of an HKY locus; 32-bit words of a pair's slot mask
  const int *tab;       // model tables the event sweep indexes with lane-dependent subscripts, in global memory (constant
                        // memory serialises such reads): see kTab* below
};

// Multi-GPU exchange of the swap sums (swapchains_bwprocesses, swapchains.cpp:192-523, exchanges them by MPI messages): every
// rank holds S[2 step parities][all chains] and two arrival counters in its own memory, mapped into every peer (NVLink peer
// access, cudaIpc between processes).  The kernel that finishes a chain's step stores the chain's S into EVERY rank's table and
// then bumps that rank's counter; the swap kernel of a rank waits until its counter says all chains have arrived.  No host
// round trip, no collective call: the exchange is the epilogue of the compute kernel and the prologue of the swap kernel.
constexpr int kMaxRanks = 16;
struct Exchange {
  double *peer_S[kMaxRanks];                      // rank r's table, as this GPU addresses it
  unsigned long long *peer_arrived[kMaxRanks];    // rank r's counters [2]
  unsigned long long step0;                       // device step counter when the exchange was attached (same on every rank)
  int world, rank;
  int publisher;                                  // which kernel of the step publishes: 1 k_accept, 2 k_accept_t, 3 k_changeu (0: nobody)
  // the cold chain's record (what rank 0 writes to the .ti file and the report) travels the same way: the rank that holds the
  // chain at beta = 1 stores it into rank 0's table and then the sequence number of the request
  double *cold_msg0;                              // rank 0's message area [cold_len], as this GPU addresses it
  unsigned long long *cold_seq0;                  // rank 0's sequence word
  int cold_len;
};

// Everything a kernel needs, passed by value.
struct EngineView {
  EngineDims d;
  MathCtx mc;
  const DevLocus *loci;
  const uint32_t *sitemask;
  const unsigned char *seq;
  const int *mult;
  PairBuf buf[2];
  unsigned char *cur;       // [P] which buffer holds the current state
  double *uvals;            // [P][kMaxLinked] mutation-rate scalars
  double *kappa;            // [P]
  double *pi;               // [P][4]
  double *hky_frac;         // [P][hky_stride] partial likelihoods and scale factors of every internal node, two slots each (likelihood_hky)
  // per chain
  double *tvals;            // [nchains][kMaxPeriods] split times, TIMEMAX sentinel at [nsplit]
  double *beta;             // [nchains]
  int *all_i;               // [nchains][NI]
  double *all_d;            // [nchains][ND]
  double *qint;             // [nchains][kMaxParams]
  double *mint;             // [nchains][kMaxParams]
  double *probg;            // [nchains]
  double *pdgsum;           // [nchains]  allpcalc.pdg
  double *swapsum;          // [nchains]  sum_li pdg (+ probg): the S of swapweight
  // proposal hand-off (propose kernel -> accept kernel)
  double *prop_extra;       // [P] migweight + slideweight + Aterm
  uint32_t *prop_flags;     // [P]
  short *prop_ids;          // [P][2] HKY loci: junction node after the move, parent of the freed node before it (what k_weigh needs of the move)
  double *prop_dbg;         // [P][4] migweight, slideweight, slide distance drawn, edge moved (parity tests)
  // counters
  unsigned int *acc;        // [P][3] accepted: any, topology, tmrca
  unsigned int *cold_acc;   // [nloci][3] the same, of the chain at beta == 1 only (the reference's update-rate tables count chain 0)
  unsigned long long *nsteps;   // [1] steps done (device-side step counter, feeds the RNG streams)
  unsigned long long *overflow; // [1] proposals dropped because the migration pool was full
  unsigned long long seed;
  // what this launch covers: chains [c_lo, c_lo + c_n) of the GPU's chains (chain groups run on their own streams and
  // overlap each other's accept sweeps), and the step the launch belongs to relative to the device counter (a graph of
  // several steps advances the counter once, at its end)
  int c_lo, c_n, step_off;
  int grp, redo_grid;           // chain group of the launch (its redo counters), blocks of the redo kernels
  int *redo_count;              // [groups][2 kinds][2 step parities]
  int *redo_list;               // [2 kinds][P]
  Exchange xch;
};

IMA_HD unsigned long long current_step(const EngineView &E) { return *E.nsteps + (unsigned long long)E.step_off; }

}  // namespace ima
