/*
 * ima2p_b200 -- C ABI of the B200-native engine for IMa2p's data-parallel hot path.
 *
 * The reference (arunsethuraman/ima2p) has no plugin/FFI interface: its hot path is a set of free
 * functions over process-global state, called with (chain, locus) integer handles
 * (SURVEY.md section 8b).  Each entry point below is the batched, handle-based replacement of one such
 * seam function; the reference file:line it replaces is cited.  Plain pointers and sizes only.
 * All functions return 0 on success or a negative IMA2P_E_* code; ima2p_last_error() gives the text.
 * There is no CPU fallback: every call needs a CUDA device (IMA2P_E_CUDA otherwise).
 *
 * Conventions
 *   - edges 0..n-1 are tips, n..2n-2 internal; edge.time is the time at the BOTTOM of the edge, the
 *     root edge has down = -1 and time = 1e6 (TIMEMAX) (imamp.hpp:583-679).
 *   - migration lists are CSR: mig_off[numlines+1], mig_t[], mig_p[] (imamp.hpp:601-605, (mt, mp) pairs).
 *   - genealogy weights (struct genealogy_weights, imamp.hpp:878-887) are flat:
 *       ints    wi[NI] = cc[k][i] (k = 0..nsplit, i < npops-k) followed by mc[k][i][j] (k < nsplit)
 *       doubles wd[ND] = fc[k][i], then hcc[k][i], then fm[k][i][j]        NI = ncc+nmc, ND = 2*ncc+nmc
 *   - chains are indexed locally (0..nchains_local-1); chain0 is the global index of local chain 0.
 */
#ifndef IMA2P_B200_H
#define IMA2P_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IMA2P_OK 0
#define IMA2P_E_ARG (-1)        /* bad argument / call order */
#define IMA2P_E_CUDA (-2)       /* CUDA runtime failure or no device */
#define IMA2P_E_UNSUPPORTED (-3)        /* model feature not on the device path yet */
#define IMA2P_E_DEVICE (-4)     /* error word raised by a kernel: IMERR_LOGDIFF / IMERR_*GAMMA equivalents
                                   (utilities.hpp:18-72); the reference exit()s there */
#define IMA2P_E_CAPACITY (-5)   /* a genealogy does not fit the configured migration capacity
                                   (reference: IMERR_MIGARRAYTOOBIG / IMERR_TOOMANYMIG) */

#define IMA2P_MODEL_IS 0        /* INFINITESITES */
#define IMA2P_MODEL_HKY 1
#define IMA2P_MODEL_SW 2        /* STEPWISE */
#define IMA2P_MODEL_JOINT 3     /* JOINT_IS_SW: part 0 infinite sites, parts 1.. stepwise */
#define IMA2P_MAX_LINKED 4       /* linked parts per locus kept on the device (reference MAXLINKED 15) */

typedef struct ima2p_engine ima2p_engine;
typedef struct ima2p_lmode ima2p_lmode;

const char *ima2p_version (void);
const char *ima2p_last_error (void);

/* ---- set-up: replaces the global state built by setup() (initialize.cpp:2074) ------------------ */
int ima2p_engine_create (ima2p_engine ** out, int device, int nchains_local, int nchains_global, int chain0,
                         int nloci, int mig_capacity, uint64_t seed);
void ima2p_engine_destroy (ima2p_engine * e);

/* population tree, period lists, parameter -> weight-position lists, priors
 * (setup_poptree build_poptree.cpp:628-709; setup_iparams initialize.cpp:201-727; nomigrationchecklist :760-809) */
int ima2p_engine_set_model (ima2p_engine * e, int npops, int nsplit, const int *plist /* [npops*npops], -1 padded */ ,
                            const int *addpop /* [nsplit+1] */ , const int *droppops /* [(nsplit+1)*2] */ ,
                            const int *pt_e, const int *pt_down /* [2*npops-1] */ , int rootpop,
                            int nq, const int *q_off, const int *q_p, const int *q_r, const double *q_max,
                            const double *q_min, int nm, const int *m_off, const int *m_p, const int *m_r,
                            const int *m_c, const double *m_max, const double *m_min, const double *m_mean,
                            int nomig_n, const int *nomig_p, const int *nomig_r, const int *nomig_c,
                            int nomigration, int expoprior, int thermo, double gbeta);

/* The same tables built from the population tree string and the priors, for the default model: one size parameter per
 * population of the tree (-q qmax), two migration parameters for every pair of populations that coexist (-m mmax; 0 =
 * no migration; expo_prior: -j7 with mean m_mean), setup_poptree build_poptree.cpp:391-448,508-539,628-709 and
 * setup_iparams initialize.cpp:201-727.  dims[6] = npops, nsplit, numtreepops, numpopsizeparams, nummigrateparams,
 * migration weight positions. */
typedef struct ima2p_modelspec ima2p_modelspec;
int ima2p_modelspec_create (ima2p_modelspec ** out, int npops, const char *tree, double qmax, double mmax,
                            int expo_prior, double m_mean, int thermo, double gbeta);
void ima2p_modelspec_free (ima2p_modelspec * s);
int ima2p_modelspec_dims (const ima2p_modelspec * s, int *dims);
int ima2p_modelspec_tables (const ima2p_modelspec * s, int *plist, int *addpop, int *droppops, int *pt_b, int *pt_e,
                            int *pt_down, int *q_off, int *q_p, int *q_r, int *m_off, int *m_p, int *m_r, int *m_c);
int ima2p_engine_set_model_spec (ima2p_engine * e, const ima2p_modelspec * s);

/* struct locus (imamp.hpp:894-936) as produced by readdata (readata.cpp:1037): seq is [numgenes][numsites]
 * (0/1 segregating sites for IS, bases 0..3 of the compressed patterns for HKY), mult[numsites] (HKY). */
int ima2p_engine_set_locus (ima2p_engine * e, int li, int model, int numgenes, int numsites, int totsites,
                            double hval, const int *samppop, const int *seq, const int *mult, int nlinked,
                            const int *minA, const int *maxA, double sumlogk);

/* allocate the HBM-resident state (call once after the model and every locus are set) */
int ima2p_engine_finalize (ima2p_engine * e);

/* heating schedule, setheat (swapchains.cpp:71-178): heatmode 0 linear, 1 geometric, 2 even */
int ima2p_engine_set_heating (ima2p_engine * e, int heatmode, double hval1, double hval2);
int ima2p_engine_set_betas (ima2p_engine * e, const double *betas_global /* [nchains_global] */ );

/* ---- state upload / download: replaces C[ci]->G[li] (imamp.hpp:956-1012) and .mcf reload (mcmcfile.cpp:310-442) -- */
int ima2p_engine_set_chain (ima2p_engine * e, int ci, const double *tvals /* [nsplit] */ );
int ima2p_engine_set_genealogy (ima2p_engine * e, int ci, int li, const int *up0, const int *up1, const int *down,
                                const int *pop, const double *time, const int *mig_off, const double *mig_t,
                                const int *mig_p, int root, double roottime, const double *uvals, double kappa,
                                const double *pi, const int *A /* [nlinked][numlines] or NULL */ );
/* which = 0: the current genealogy; 1: the pair's other buffer, i.e. after a step the proposed genealogy when
 * it was rejected, or the previous genealogy when it was accepted (restoreedges, update_gtree_common.cpp:705-814,
 * becomes a buffer flip) */
int ima2p_engine_get_genealogy (ima2p_engine * e, int ci, int li, int which, int *up0, int *up1, int *down, int *pop,
                                double *time, int *mig_off, double *mig_t, int *mig_p, int mig_room, int *root,
                                double *roottime);
/* stepwise loci: allele states A[nlinked][numlines] at the top of every edge, the per-branch terms dlikeA and the
 * per-portion likelihoods pdg_a (struct edge A/dlikeA imamp.hpp:651-679, genealogy.pdg_a :956-987) */
int ima2p_engine_get_alleles (ima2p_engine * e, int ci, int li, int which, int *A, double *dlikeA, double *pdg_a);
/* push everything staged by set_chain / set_genealogy to the device (one H2D per array) */
int ima2p_engine_upload (ima2p_engine * e);

/* ---- evaluation of a loaded state: what init_p() does after a reload (mcmcfile.cpp:130-193) ----
 * per (chain, locus): treeweight (update_gtree_common.cpp:1679-1931) + likelihoodIS/HKY/SW
 * (calc_prob_data.cpp:583-607, 731-836, 841-909); per chain: sum_treeinfo + initialize_integrate_tree_prob
 * (update_gtree_common.cpp:2056-2134). */
int ima2p_engine_eval (ima2p_engine * e);
/* out_d = {pdg, length, tlength, roottime}, out_i = {mignum, root} */
int ima2p_engine_get_pair (ima2p_engine * e, int ci, int li, int *wi, double *wd, double *out_d, int *out_i);
/* out_d = {probg, pdg, beta} (struct probcalc imamp.hpp:858-865) */
int ima2p_engine_get_chain (ima2p_engine * e, int ci, int *all_wi, double *all_wd, double *qintegrate,
                            double *mintegrate, double *out_d);
int ima2p_engine_dims (ima2p_engine * e, int *out /* {NI, ND, NL, CAP, rowlen} */ );

/* ---- M mode -------------------------------------------------------------------------------------
 * one step = updategenealogy(ci, li) for every chain x locus (qupdate ima_main_mpi.cpp:1821-1841 ->
 * update_gtree.cpp:723-966) followed by swaptries MC3 temperature swaps (swapchains.cpp:526-653;
 * temperature-rank form of swapchains_bwprocesses :192-523).  Single-GPU form: */
int ima2p_engine_run (ima2p_engine * e, int nsteps, int swaptries, void *cuda_stream);
/* How ima2p_engine_run issues its steps.  Chains only meet at the swaps (swapchains.cpp:526-653) and a chain's loci are
 * decided in order (qupdate's loop over li, ima_main_mpi.cpp:1821-1841), so the GPU's chains are cut into `groups`
 * groups on their own streams -- the accept sweep of one group overlaps the proposals of the others -- and `depth`
 * consecutive steps form one CUDA graph in which a group starts its next proposals without waiting for the other
 * groups' swaps.  decisions_first != 0 puts every group's decision kernels on a high-priority stream of their own.
 * The run is bit for bit the same for every setting. */
int ima2p_engine_set_pipeline (ima2p_engine * e, int groups, int depth, int decisions_first);
/* Kernel launches one step of ima2p_engine_run makes with the current settings (per chain group: the proposal kernels, the
 * accept sweep, the split-time proposals and their decision, the mutation scalars; plus the swaps once), for a caller that
 * reports them.  The calls it replaces are the body of qupdate(), ima_main_mpi.cpp:1808-1945. */
int ima2p_engine_launches_per_step (ima2p_engine * e, int swaptries);
/* Which kernels make updategenealogy's proposal (update_gtree.cpp:723-827): fast != 0 (the default where it applies) a
 * lane-per-pair move kernel followed by a warp-per-pair weights / likelihood kernel, with pairs_per_warp lanes of a move warp at
 * work (1, 2, 4, 8, 16, or 0 = chosen from the number of pairs); fast == 0 the general one-warp-per-pair kernel for every pair.
 * Pairs that do not fit the fast kernels' tables take the general path either way; the chain does not depend on the choice. */
int ima2p_engine_set_proposal_path (ima2p_engine * e, int fast, int pairs_per_warp);
/* ---- chains sharded over several GPUs: the reference's `mpirun -np P IMa2p -hn ...` (ima_main_mpi.cpp:4317-4560), whose
 * processes exchange swap sums by MPI messages (swapchains_bwprocesses, swapchains.cpp:192-523).  Here every rank keeps a
 * small table (S of every chain of the job, two step parities, and arrival counters) that its peers map -- NVLink peer
 * access inside one process, cudaIpc handles between processes -- and the kernels do the exchange themselves: the kernel that
 * finishes a chain's step stores its S into every rank's table, the swap kernel waits for all chains to arrive.  Set-up on
 * every rank: exchange_create, pass the table to the peers (ipc_export -> any channel -> ipc_import), exchange_attach with
 * all ranks' tables (tables[own rank] may be NULL); every rank attaches before any rank steps, and all ranks make the same
 * calls.  ima2p_engine_run_sharded = ima2p_engine_run for a shard: the whole step, exchange included, is one CUDA graph. */
int ima2p_engine_exchange_create (ima2p_engine * e, void **table, uint64_t * bytes);
int ima2p_engine_exchange_attach (ima2p_engine * e, void *const *tables);
int ima2p_ipc_export (const void *device_pointer, unsigned char *handle64);
int ima2p_ipc_import (int device, const unsigned char *handle64, void **device_pointer);
int ima2p_engine_run_sharded (ima2p_engine * e, int nsteps, int swaptries, void *cuda_stream);
/* the cold chain's record of a sharded job, called by every rank at the same step boundary: on rank 0 out_msg[rowlen + 2 + nloci] =
 * the .ti row (savegsampinf ginfo.cpp:318-377), probg, P(D|G), P(D|G) per locus -- stored into rank 0's table by whichever rank
 * holds the chain at beta = 1; on the other ranks out_msg is not touched */
int ima2p_engine_cold_message (ima2p_engine * e, double *out_msg, void *cuda_stream);
/* the two halves of a shard's step for callers that keep the ranks in lockstep themselves (every rank's update, then every
 * rank's swap): one process driving several GPUs, and the tests */
int ima2p_engine_sharded_update (ima2p_engine * e, void *cuda_stream);
int ima2p_engine_sharded_swap (ima2p_engine * e, int swaptries, void *cuda_stream);
/* Migration capacity while the engine runs.  The reference grows an edge's migration list whenever it fills (checkmig,
 * utilities.cpp:1365-1383; IMERR_MIGARRAYTOOBIG beyond ABSMIGMAX 5000).  Here a proposal whose genealogy would hold more than
 * mig_capacity events is dropped and counted (ima2p_engine_counters field 7); a caller that sees the count move calls this at a
 * step boundary: the pools are re-made with the new capacity, the resident genealogies copied over, nothing else changes.
 * IMA2P_E_CAPACITY when a genealogy of that size no longer fits the kernels' shared-memory tables. */
int ima2p_engine_grow_capacity (ima2p_engine * e, int new_capacity);
/* parity tests: keep the per-proposal record that ima2p_engine_get_proposal reads (off by default) */
int ima2p_engine_set_debug_records (ima2p_engine * e, int on);
/* speculative depth of the accept sweep (1..3): how many consecutive loci of a chain are evaluated per round against
 * the same all-locus sums; results are identical for every depth (see csrc/ima_kernels.h k_accept) */
int ima2p_engine_set_speculation (ima2p_engine * e, int depth);
/* same steps, launched kernel by kernel with CUDA events on the launching stream around each kernel;
 * kernel_ms[IMA2P_TIMED_SLOTS] = summed device time of {0 the proposal kernels together, 1 accept, 2 swap, 3 split-time proposals,
 * 4 accept_t, 5 changeu, 6 k_move, 7 k_weigh, 8 k_propose_redo, 9-11 unused} (roofline accounting) */
#define IMA2P_TIMED_SLOTS 12
int ima2p_engine_run_timed (ima2p_engine * e, int nsteps, int swaptries, void *cuda_stream, float *kernel_ms);
/* Multi-GPU form (one process per GPU): genealogy updates of the local chains, then the per-chain
 * S = sum_li pdg + probg (swapweight, swapchains.cpp:12-34) is written to dev_S_local[nchains_local]
 * (device memory owned by the caller); the caller all-gathers it over NCCL into
 * dev_S_global[nchains_global] and every rank replays the same swap attempts. */
int ima2p_engine_update_genealogies (ima2p_engine * e, double *dev_S_local, void *cuda_stream);
int ima2p_engine_swap_replay (ima2p_engine * e, const double *dev_S_global, int swaptries, void *cuda_stream);
/* The same step in split phases, so that the exchange of step s hides behind the proposals of step s+1 (updategenealogy's
 * proposal half does not read beta): step_propose launches the proposals of every local pair; step_decide the accept sweep,
 * the split-time / scalar updates, writes S to dev_S_local and advances the step counter; swap_replay_late is swap_replay
 * for a step whose counter has already been advanced (same draws).  The caller orders them with stream events:
 * decide(s) -> [all-gather, swap_replay_late](s) -> decide(s+1), while propose(s+1) only follows decide(s). */
int ima2p_engine_step_propose (ima2p_engine * e, void *cuda_stream);
int ima2p_engine_step_decide (ima2p_engine * e, double *dev_S_local, void *cuda_stream);
int ima2p_engine_swap_replay_late (ima2p_engine * e, const double *dev_S_global, int swaptries, void *cuda_stream);

/* last proposal of a pair: out4[5] = {migweight (update_gtree.cpp:663), slideweight (:803-812), slide distance
 * drawn (:783), edge moved, migweight + slideweight + Atermsum (the non-likelihood part of the MH exponent, :919-924)}; flags bit0 infinite-sites reject, bit1 dropped for capacity, bit2 topology changed,
 * bit3 root moved; buffer = index of the buffer that is current (flips on accept) */
int ima2p_engine_get_proposal (ima2p_engine * e, int ci, int li, double *out4, unsigned int *flags, int *buffer);

/* parity hook for the device numerics (uppergamma / lowergamma utilities.cpp:1053-1122): out[4*i..] =
 * {uppergamma, lowergamma} in their one-lane form and in their warp-cooperative form for (a[i], x[i]) */
int ima2p_debug_gamma (int device, const int *a, const double *x, int n, double *out);

/* The rest of one qupdate step (ima_main_mpi.cpp:1867-1945), local to each chain:
 *   t_updates      : a split-time update of every chain in every step: 1 = changet_RY1 (update_t_RY.cpp:222-517),
 *                    2 = changet_NW (update_t_NW.cpp:919-1032), 3 = one of the two at random per chain, as the
 *                    reference does (ima_main_mpi.cpp:1871-1872); 0 = none;
 *   u_every  > 0   : changeu for every mutation-rate scalar (update_mc_params.cpp:23-370; changekappa :381-431 when a
 *                    single HKY locus is all there is) in every u_every-th step (the reference: 5, UUPDATEINC 4).
 * Both default to off; ima2p_engine_run / update_genealogies then perform them after the genealogy updates.
 * set_update_priors: split-time prior bounds T[].pr (t_max/t_min[nsplit], NULL keeps the current ones), the mutation
 * scalar prior bound log(UMAX) and window (<= 0 keeps / derives initialize.cpp:1448-1450), kappa window and bound. */
int ima2p_engine_set_update_schedule (ima2p_engine * e, int t_updates, int u_every);
int ima2p_engine_set_update_priors (ima2p_engine * e, const double *t_max, const double *t_min, double u_prior_max,
                                    double u_window, double kappa_window, double kappa_max);
/* tries / accepts: out4 = split-time tries, accepts, mutation-scalar tries, accepts */
int ima2p_engine_update_counters (ima2p_engine * e, uint64_t * out4);
/* What the reference's update-rate tables and swap table report (callprintacceptancerates ima_main_mpi.cpp:3473-3900 over
 * the cold chain's update_rate_calc records; printchaininfo swapchains.cpp:760-778), counted since the engine was created
 * for the chain at beta == 1 while it lives on this device; any pointer may be NULL:
 *   genealogy[nloci][3]  accepted updategenealogy calls: any, topology-changing, tmrca-changing (tries = steps);
 *   split[nsplit][4]     per split time: Rannala-Yang tries, accepts, Nielsen-Wakeley tries, accepts;
 *   scalars[nurates][2]  per mutation-rate scalar (readata.cpp:832-834 order): tries, accepts, a proposal counting for
 *                        both scalars it trades between as in qupdate (ima_main_mpi.cpp:1926-1935);
 *   adjacent[nchains_global - 1][2]  swap attempts, swaps between temperature ranks r and r + 1 (tempbasedswapcount).
 * With several GPUs the first three are this rank's share (sum them over ranks); `adjacent` is the same on every rank, the
 * swap replay being replicated. */
int ima2p_engine_cold_counters (ima2p_engine * e, uint64_t * genealogy, uint64_t * split, uint64_t * scalars,
                                uint64_t * adjacent);
/* current split times C[ci]->tvals[nsplit] of one chain */
int ima2p_engine_get_split_times (ima2p_engine * e, int chain, double *tvals);
/* all of them at once: tvals[nchains][nsplit], uvals[P][IMA2P_MAX_LINKED], kappa[P] (any pointer may be NULL) */
int ima2p_engine_fetch_parameters (ima2p_engine * e, double *tvals, double *uvals, double *kappa);
/* mutation-rate scalars (uvals[IMA2P_MAX_LINKED]) and kappa of one (chain, locus) */
int ima2p_engine_get_scalars (ima2p_engine * e, int chain, int locus, double *uvals, double *kappa);
/* parity hooks (tests): one changet_RY1 (method 0) or changet_NW (method 1) with the proposed times given,
 * newt[nchains] (NULL: drawn); force_accept -1
 * draws the decision, 0 rejects, 1 accepts; out[nchains][4] = period, proposed time, log MH term, accepted.
 * debug_changeu evaluates, and never applies, the proposal u_j *= d, u_k /= d on one chain:
 * out[4] = new P(D|G) of j's part, of k's part, MH term (update_mc_params.cpp:291), 0 */
int ima2p_engine_debug_split_time (ima2p_engine * e, int method, int period, const double *newt, int force_accept,
                                   double *out);
int ima2p_engine_debug_changeu (ima2p_engine * e, int chain, int j, int k, double d, double kappa_j, double kappa_k,
                                double *out);

/* Thermodynamic integration (marglike.cpp:51-87 summarginlikecalc, :121-150 thermomarginlikecalc).
 * accumulate: thermosum[slot of the chain's beta] += allpcalc.pdg for every local chain (call once per recorded step);
 * thermo_sums: this rank's share of thermosum[nchains_global] (sum the shares over ranks), optionally reset;
 * thermo_marginlike: the reference's Simpson rule over k recorded steps (host arithmetic). */
int ima2p_engine_thermo_accumulate (ima2p_engine * e, void *cuda_stream);
int ima2p_engine_thermo_sums (ima2p_engine * e, double *thermosum_global, int reset);
int ima2p_thermo_marginlike (const double *thermosum, int numchains, int k, double *out);

/* counters: out = {steps, updates tried, accepted, topology-changing accepted, tmrca-changing accepted,
 *                  swap attempts, swaps accepted, proposals dropped for capacity} */
int ima2p_engine_counters (ima2p_engine * e, uint64_t * out8);
int ima2p_engine_get_betas (ima2p_engine * e, double *betas_global);
/* savegsampinf (ginfo.cpp:318-377): the 4*nq+3*nm+2+nsplit float row of the chain with beta == 1;
 * returns 1 in *present when that chain lives on this GPU */
int ima2p_engine_cold_row (ima2p_engine * e, float *row, int *present);
int ima2p_engine_sync (ima2p_engine * e);

/* bulk state I/O in the engine's own packed layout (host buffers; used for end-to-end timing):
 * sizes from ima2p_engine_state_bytes: {topo, time, mseg, mig_t, mig_p, scal_i, scal_d, uvals} */
int ima2p_engine_state_bytes (ima2p_engine * e, uint64_t * out8);
int ima2p_engine_put_state (ima2p_engine * e, const void *topo, const void *time, const void *mseg, const void *mig_t,
                            const void *mig_p, const void *scal_i, const void *scal_d, const void *uvals,
                            const double *tvals /* [nchains][nsplit] */ , void *cuda_stream);
int ima2p_engine_fetch_state (ima2p_engine * e, void *topo, void *time, void *mseg, void *mig_t, void *mig_p,
                              void *scal_i, void *scal_d, void *cuda_stream);
/* put_state in a narrow wire form (13 instead of 20 bytes per edge over PCIe): topo8 = int8 [P][NL][4] (up0, up1, down, pop),
 * mcount = uint8 [P][NL] migration events per edge, the pools mig_t / mig_p [P][CAP] holding the events of a pair in edge
 * order (segment starts are the prefix sums of mcount, which is how fetch_state delivers them); the other buffers as in
 * put_state.  Needs 2n-1 <= 127 edges and mig_capacity <= 255; otherwise IMA2P_E_ARG and put_state is the way. */
/* The same narrow form as ONE host block, so that a step's state crosses PCIe in a single transfer: sections at the byte
 * offsets ima2p_engine_state_block_layout returns for `total_events` migration events over all pairs --
 * out10 = {time f64[P][NL], scal_d f64[P][4], uvals f64[P][IMA2P_MAX_LINKED], tvals f64[nchains][nsplit], mig_t f64[events],
 * scal_i i32[P][2], mig_p i16[events], topo8 i8[P][NL][4], mcount u8[P][NL], total bytes}; the events of pair 0 first, then
 * pair 1, ..., each pair's in edge order (scal_i[p][1] of them). */
int ima2p_engine_state_block_layout (ima2p_engine * e, long long total_events, uint64_t * out10);
int ima2p_engine_put_state_block (ima2p_engine * e, const void *block, long long total_events, void *cuda_stream);
/* put_state_block in two halves for a caller that steps a stream of uploaded states: upload_block starts the transfer of a
 * block into one of two device staging slots on `copy_stream` and returns (IMA2P_E_ARG when both slots are taken);
 * adopt_block makes `cuda_stream` wait for the oldest transfer, widens that block into the resident state and re-evaluates
 * it.  With the next block uploaded before this step's results are read, the copy overlaps the step's kernels. */
int ima2p_engine_upload_block (ima2p_engine * e, const void *block, long long total_events, void *copy_stream);
int ima2p_engine_adopt_block (ima2p_engine * e, void *cuda_stream);
int ima2p_engine_put_state_packed (ima2p_engine * e, const void *topo8, const void *time, const void *mcount,
                                   const void *mig_t, const void *mig_p, const void *scal_i, const void *scal_d,
                                   const void *uvals, const double *tvals, void *cuda_stream);
/* per-pair summaries of the current genealogies: sd[P][4] = {roottime, length, tlength, pdg}, si[P][2] = {root, mignum},
 * wi[P][NI] = coalescence | migration counts (struct genealogy fields imamp.hpp:956-987); NULL pointers are skipped */
int ima2p_engine_fetch_pair_summaries (ima2p_engine * e, double *sd, int *si, int *wi, void *cuda_stream);
/* P(D|G) of every locus of one local chain, pdg[nloci] (C[ci]->G[li].pdg; what checkhighs output.cpp:207-240 reads) */
int ima2p_engine_fetch_chain_pdg (ima2p_engine * e, int ci, double *pdg);
/* per-chain summary after a run: out[ci] = {beta, probg, pdg, S} */
int ima2p_engine_fetch_chain_summary (ima2p_engine * e, double *out4 /* [nchains_local][4] */ , void *cuda_stream);

/* ---- L mode: replaces gsampinf + margincalc / marginp (surface_call_functions.cpp:25-173) and
 * jointp (jointfind.cpp:885-1047) over the .ti rows (ginfo.cpp:288-304) -------------------------- */
int ima2p_lmode_create (ima2p_lmode ** out, int device, int nq, int nm, int nsplit, const double *q_max,
                        const double *q_min, const double *m_max, const double *m_min, const double *m_mean,
                        int expoprior);
void ima2p_lmode_destroy (ima2p_lmode * l);
/* rows: host float [nrows][rowlen] (this rank's shard); nrows_total = rows over all ranks */
int ima2p_lmode_load (ima2p_lmode * l, const float *rows, int nrows, int rowlen, long long nrows_total);
/* sums[i] = sum over this rank's rows of the margincalc term of parameter `param` at x[i]
 * (surface_call_functions.cpp:139-162; INTEGERROUND counts).  dev_sums may be NULL (then host_sums is
 * filled) or device memory for an NCCL all-reduce by the caller. */
int ima2p_lmode_marginal_sums (ima2p_lmode * l, int param, const double *x, int nx, int first, int last,
                               int round_counts, double *host_sums, double *dev_sums, void *cuda_stream);
/* margincalc (:119-173) / marginp (:25-80) on a single GPU holding all rows */
int ima2p_lmode_margincalc (ima2p_lmode * l, int param, const double *x, int nx, double yadjust, int logi, double *out);
int ima2p_lmode_marginp (ima2p_lmode * l, int param, int firsttree, int lasttree, const double *x, int nx, double *out);
/* n independent one-point evaluations in one device pass, for searches that advance in lock step (the marginal peak searches
 * and the 95% bounds of findmarginpeaks, surface_call_functions.cpp:175-297, each of which calls marginp / margincalc once per
 * iterate in the reference): kind[q] = 0 -> marginp(param[q], first[q], last[q], x[q]); kind[q] = 1 -> log margincalc(x[q])
 * of param[q] over all rows minus yadjust[q].  Values are bit for bit those of the single calls. */
int ima2p_lmode_marginal_many (ima2p_lmode * l, int n, const int *kind, const int *param, const int *first, const int *last,
                               const double *x, const double *yadjust, double *out);
/* jointp for nvec parameter vectors x[nvec][nq+nm]; out_q[nvec] = -log joint density, out_ess[nvec] */
int ima2p_lmode_jointp (ima2p_lmode * l, const double *x, int nvec, int calc_ess, double *out_q, double *out_ess);
/* nowmodeltype (jointfind.cpp:1104-1133): 0 all parameters (two populations, the default), 1 population sizes only, 2 migration
 * rates only -- the two full models of a three-population search (:949-952, :973-980) */
int ima2p_lmode_set_joint_model (ima2p_lmode * l, int modeltype);

/* Sharded form (rows split over GPUs; the caller exchanges a few doubles per vector over NCCL):
 *   phase 1: p_g of every local row for nvec (<= 32) vectors; seed_before[v] = max of p over the rows held by
 *            lower ranks (NULL on rank 0); localmax_out[v] = max(seed, local rows)
 *   reseed : same prefixes recomputed with the seed once it is known (after the ranks exchanged their local maxima)
 *   phase 2: given the global maximum, records_out[v] = {inserted, kept, sum, sum of squares, smallest kept p,
 *            its scaled term} over the local rows; global_row0 = global index of local row 0
 *   finish : the closing arithmetic of jointp (:1011-1046) on the records summed over ranks */
int ima2p_lmode_joint_phase1 (ima2p_lmode * l, const double *x, int nvec, const double *seed_before, double *localmax_out);
int ima2p_lmode_joint_reseed (ima2p_lmode * l, int nvec, const double *seed_before, double *localmax_out);
int ima2p_lmode_joint_phase2 (ima2p_lmode * l, int nvec, const double *globalmax, long long global_row0,
                              double *records_out);
void ima2p_lmode_joint_finish (const double *rec6, double globalmax, long long nrows_total, int calc_ess, double *q,
                               double *ess);
/* The same phases with every intermediate left on the device, for ranks that exchange with device collectives (an NCCL
 * all-gather of the local maxima, then one of the records): nothing crosses PCIe between the phases and nothing
 * synchronises, so batch after batch queues on one stream.
 *   begin : nvec (<= 512) vectors over the local rows; dev_localmax_out[nvec] (device memory) = maxima over the local rows
 *   middle: dev_allmax[world][nvec] (device, the gathered maxima) -> dev_records_out[nvec][8] (device) = the six record
 *           fields of phase 2, the global maximum, 0.  Sum the first four fields over ranks, take fields 4-5 from the rank with
 *           the smallest field 4, and close with ima2p_lmode_joint_finish. */
/* measured FP64 peaks of the device the roofline fractions of the FP64-bound kernels are quoted against (SURVEY.md section
 * 8d): out2[0] = fused multiply-adds per second (x 2 = flop/s), out2[1] = exp evaluations per second */
int ima2p_debug_fp64_peaks (int device, double *out2);
void ima2p_lmode_joint_finish_gathered (const double *records8 /* [world][nvec][8] */, int world, int nvec, long long nrows_total,
                                        int calc_ess, double *q, double *ess);
int ima2p_lmode_joint_begin (ima2p_lmode * l, const double *x, int nvec, double *dev_localmax_out, void *cuda_stream);
int ima2p_lmode_joint_middle (ima2p_lmode * l, int nvec, const double *dev_allmax, int world, int rank, long long global_row0,
                              double *dev_records_out, void *cuda_stream);

/* ---- L mode, the other evaluators that stream over the rows ------------------------------
 * moments = print_means_variances_correlations (output.cpp:687-745) over calcx (output.cpp:14-134): means[np],
 * variances[np] (E[x^2] - mean^2), correlations[np][np] (entries p < q; may be NULL); np = nq + nm <= 32.  raw_sums
 * (may be NULL) receives the row sums themselves: [np] of calcx(.,p,0), [np] of calcx(.,p,1), [np][np] of the products. */
int ima2p_lmode_moments (ima2p_lmode * l, double *means, double *variances, double *correlations, double *raw_sums);
/* sharded rows: every rank calls moments (raw_sums) / popmig_sums on its rows, the caller all-reduces the sums and finishes:
 * moments_finish is the closing arithmetic of output.cpp:709-739 on the summed raw_sums; the summed popmig_sums divided by
 * the total number of rows are calc_popmig's mean (uniform migration prior only) */
void ima2p_lmode_moments_finish (int np, const double *raw_sums, long long nrows_total, double *means, double *variances,
                                 double *correlations);
int ima2p_lmode_popmig_sums (ima2p_lmode * l, int thetai, int mi, const double *x, int nx, int first, int last,
                             double *out_sums);
/* density of the product 2NM = theta_thetai * m_mi / 2 at x[nx]: calc_popmig (popmig.cpp:9-97) or, when the handle was
 * created with the exponential migration prior, calc_pop_expomig (:101-170); prob_or_like = 1 divides by the prior density */
int ima2p_lmode_popmig (ima2p_lmode * l, int thetai, int mi, const double *x, int nx, int prob_or_like, double *out);
/* marginpopmig (popmig.cpp:176-268) / marginpop_expomig (:272-357): minus the mean over rows [firsttree, lasttree) with
 * the reference's divisor; 1 (OFFSCALEVAL) outside the plotted range.  The function marginalopt_popmig minimises. */
int ima2p_lmode_marginpopmig (ima2p_lmode * l, int thetai, int mi, int firsttree, int lasttree, const double *x, int nx,
                              double *out);

/* gtpops / gtmig (gtint.cpp:128-330): probability that parameter i is greater than parameter j (kind 0: population sizes,
 * 1: migration rates), over the rows print_greater_than_tests uses (:341-351); *out = -1 where the reference prints "na" */
int ima2p_lmode_greater_than (ima2p_lmode * l, int kind, int i, int j, double *out);

/* ---- the .u input file (readata.cpp): host code, needs no device -------------------------------------------------
 * dataset_read = readdata (readata.cpp:1038-1123): top lines (:891-1036), locus header lines (parse_locus_info
 * :618-866), and the data with the reference's site handling: infinite-sites / joint loci keep the segregating,
 * two-state, all-acgt columns recoded 0/1 against the first gene (findsegsites :34-178, readseqIS :347-497); HKY loci
 * drop gapped columns and merge identical ones with multiplicities (readseqHKY :246-345, eliminategaps, sortseq);
 * stepwise parts give allele lengths and their range (readseqSW :500-588).  Malformed input is a negative return
 * code where the reference calls IM_err. */
typedef struct ima2p_dataset ima2p_dataset;
int ima2p_dataset_read (const char *path, ima2p_dataset ** out);
void ima2p_dataset_free (ima2p_dataset * d);
int ima2p_dataset_dims (const ima2p_dataset * d, int *npops, int *nloci, char *tree, int tree_len);
/* text the reference echoes into its report (readata.cpp:916-963): kind 0 = the title line (index 0) and the '#' lines under
 * it (index 1.., without the '#'), kind 1 = population names; IMA2P_E_ARG past the last one */
int ima2p_dataset_text (const ima2p_dataset * d, int kind, int index, char *buf, int buf_len);
/* info[8] = model, numgenes, numsites, totsites, numbases, nlinked, mutation rates given on the header line,
 *           flags: bit 0 the model letter carried a count (S2, J1: the reference's SW_M / IS+SW_M, readata.cpp:664-692),
 *           bit 1 an inheritance scalar stood on the header line (:729-734; echoed "%5.3lf", else "%lf" :826-831) */
int ima2p_dataset_locus (const ima2p_dataset * d, int locus, int *info, double *hval, int *samppop, char *name,
                         int name_len);
/* seq[numgenes][numsites]; mult[numsites] (HKY); A[nlinked][numgenes]; minA, maxA[nlinked]; pi[4] (HKY); urate[info[6]] */
int ima2p_dataset_locus_data (const ima2p_dataset * d, int locus, int *seq, int *mult, int *A, int *minA, int *maxA,
                              double *pi, double *urate);

/* What the host reads after a step (recording, ima_main_mpi.cpp:2891, 3035) in one kernel, one copy and one
 * synchronisation: chain4[nchains][4] = beta, probg, P(D|G), swap sum of every local chain; row = the cold chain's .ti
 * row (ima2p_engine_cold_row) when it lives on this rank (*present). */
int ima2p_engine_step_report (ima2p_engine * e, double *chain4, float *row, int *present, void *cuda_stream);
/* The same in two halves with two slots (0, 1): _begin queues the packing kernel and the copy and returns; _end waits for
 * that slot's copy only.  A host that queues step s+1 before it reads the results of step s (the recording of :2891 does not
 * feed back into the step) keeps the device busy while it reads. */
int ima2p_engine_step_report_begin (ima2p_engine * e, int slot, void *cuda_stream);
int ima2p_engine_step_report_end (ima2p_engine * e, int slot, double *chain4, float *row, int *present);

/* ---- the MCMC state file (.mcf): writemcf / readmcf, mcmcfile.cpp:203-442 --------------------------------------
 * write_mcf: the chains this engine holds, in the reference's record stream ("name type count values", doubles as
 * %.10lg); read_mcf: loads such a file into the engine's chains (read again from the top when it holds fewer, as the
 * reference does), uploads and evaluates (init_p, :130-193). */
int ima2p_engine_write_mcf (ima2p_engine * e, const char *path);
int ima2p_engine_read_mcf (ima2p_engine * e, const char *path);

/* ---- the .ti file of sampled genealogies: M mode writes it, L mode reads it back ---------------------------------
 * ti_create: header block ending in "VALUESSTART" (ima_main_mpi.cpp:2123-2141); ti_append: one line per row, every value
 * "%.6f\t" (savegenealogyfile, output.cpp:662-685); ti_load (loadgenealogyvalues, ima_main_mpi.cpp:3216-3440):
 * rows == NULL counts the genealogies, else up to max_rows rows of rowlen floats are read; *nrows_out = rows read. */
int ima2p_ti_create (const char *path, const char *header_text);
int ima2p_ti_append (const char *path, const float *rows, long long nrows, int rowlen);
int ima2p_ti_load (const char *path, int rowlen, float *rows, long long max_rows, long long *nrows_out);

#ifdef __cplusplus
}
#endif
#endif
