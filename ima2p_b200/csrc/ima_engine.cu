// Host side of the M-mode engine and its C ABI (include/ima2p_b200.h).
//
// All chains' genealogies live in HBM in two pair-major buffers (current / proposed).  A step is three
// kernel launches (propose over pairs, accept over chains, swap) captured once into a CUDA graph and
// replayed; the step counter that feeds the counter-based RNG lives on the device so the graph needs no
// parameter updates.
#include "ima_kernels.h"
#include "ima_fastpath.h"
#include "ima_updates.h"
#include "../../include/ima2p_b200.h"
#include <string>
#include <vector>
#include <new>
#if !IMA_CUDA
#include <map>
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
#endif

namespace ima {

#if !IMA_CUDA
thread_local EmuCtx g_emu;
#endif

#if !IMA_CUDA
static std::map<const void *, std::pair<std::string, size_t>> g_emu_tables;     // exchange tables of this process (host emulation)
#endif
#if IMA_CUDA
bool g_programmatic_launch = false;
#endif
static thread_local std::string g_last_error;
static int fail(int code, const std::string &msg) { g_last_error = msg; return code; }

struct HostLocus {
  DevLocus d;
  std::vector<uint32_t> sitemask;
  std::vector<unsigned char> seq;
  std::vector<int> mult;
  bool set = false;
};

constexpr int kMaxGroups = 16;

struct Engine {
  int device = 0;
  EngineDims d{};
  DevModel model{};
  bool model_set = false, finalized = false, graph_ready = false;
  int graph_swaptries = -1;
  unsigned long long seed = 0;
  std::vector<HostLocus> loci;
  // host staging (engine layout)
  std::vector<short4_t> h_topo;
  std::vector<double> h_time, h_mig_t, h_sd, h_uvals, h_kappa, h_pi, h_tvals, h_beta_table;
  std::vector<ushort2_t> h_mseg;
  std::vector<short> h_mig_p, h_A;
  std::vector<int> h_si, h_rank_of_chain, h_chain_of_rank;
  // device
  EngineView v{};
  SwapView sv{};
  std::vector<void *> allocs;
  double *d_logfact = nullptr;
  int *d_err = nullptr;
  signed char *d_topo8 = nullptr;        // staging of the packed wire form (ima2p_engine_put_state_packed), made on first use
  unsigned char *d_mcount = nullptr;
  // staging of the one-block wire form (ima2p_engine_put_state_block / upload_block + adopt_block): two slots, so that the
  // block of the next step can travel while the kernels of this one run
  unsigned char *d_block[2] = {nullptr, nullptr};
  long long block_events[2] = {0, 0};
  int block_pending = 0, block_next_up = 0, block_next_adopt = 0;
  int *d_block_moff = nullptr;
  size_t block_cap = 0;
#if IMA_CUDA
  cudaEvent_t block_up_ev[2] = {nullptr, nullptr}, block_done_ev[2] = {nullptr, nullptr};
  bool block_done_set[2] = {false, false};
#endif
  DevLocus *d_loci = nullptr;
  double *d_beta_table = nullptr;
  double *d_prop_dbg = nullptr;
  unsigned long long *d_swap_counts = nullptr;
  double *d_thermosum = nullptr;
  double *d_report = nullptr, *h_report = nullptr;      // step_report: device staging and its pinned host mirror
  double *d_report2[2] = {nullptr, nullptr}, *h_report2[2] = {nullptr, nullptr};    // step_report_begin / _end: two slots in flight
#if IMA_CUDA
  cudaEvent_t report_ev[2] = {nullptr, nullptr};
#endif
  bool report_pending[2] = {false, false};
  UpdateView uv{};              // split-time / mutation-scalar updates
  int t_updates = 0, u_every = 0;
  std::vector<int> h_ul_l, h_ul_a;
#if IMA_CUDA
  cudaGraphExec_t graph_exec = nullptr, graph_exec_deep = nullptr;   // one step; `depth` steps
  cudaGraphExec_t graph_exec_sh = nullptr, graph_exec_sh_deep = nullptr;   // the same for a shard (exchange of swap sums inside)
  int graph_sh_swaptries = -1;
  cudaStream_t own_stream = nullptr;
  cudaStream_t group_stream[kMaxGroups][2] = {};    // per chain group: proposals (low priority), decisions (high priority)
  std::vector<cudaEvent_t> pipe_events;
#endif
  int groups = 1, depth = 1, pipe_prio = 0;          // see capture_steps
  // programmatic dependent launch inside the step graph: a kernel is scheduled while the one before it on its stream drains
  // (measured 1.5 % of the step, same run; IMA2P_PDL=0 turns it off)
  int pdl = getenv("IMA2P_PDL") ? atoi(getenv("IMA2P_PDL")) : 1;
  bool fast_ok = false, fast = false;                // the two-kernel proposal path (ima_fastpath.h): possible / in use
  int ppw = 0;                                       // pairs per warp of k_move (0: chosen from the number of pairs)
  int redo_grid = 16;
  int maxng = 0;                                     // largest sample over the loci
  // multi-GPU exchange (struct Exchange): this rank's table, and whether peers are attached
  unsigned char *d_xch = nullptr;
  size_t xch_bytes = 0;
  bool xch_attached = false;
  int cold_len = 0;
  unsigned long long cold_seq = 0;                  // requests for the cold chain's record so far (the same on every rank)
  double *d_cold = nullptr, *h_cold = nullptr;
  char xch_shm[64] = {0};                           // host emulation: name of the shared-memory object of the table
  size_t pair_smem = 0, chain_smem = 0, accept_smem = 0;
  int spec = 3;        // speculative depth of the accept sweep (see k_accept)

  template <class T> T *alloc(size_t n) {
    T *p = (T *)dev_alloc(n * sizeof(T));
    if (p) allocs.push_back(p);
    return p;
  }
  ~Engine() {
#if IMA_CUDA
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
    if (graph_exec_deep) cudaGraphExecDestroy(graph_exec_deep);
    if (graph_exec_sh) cudaGraphExecDestroy(graph_exec_sh);
    if (graph_exec_sh_deep) cudaGraphExecDestroy(graph_exec_sh_deep);
    if (own_stream) cudaStreamDestroy(own_stream);
    for (auto &g : group_stream) for (auto &x : g) if (x) cudaStreamDestroy(x);
    for (auto &x : pipe_events) if (x) cudaEventDestroy(x);
    for (int k = 0; k < 2; k++) { if (block_up_ev[k]) cudaEventDestroy(block_up_ev[k]); if (block_done_ev[k]) cudaEventDestroy(block_done_ev[k]); }
#endif
    for (void *p : allocs) dev_free(p);
#if !IMA_CUDA
    if (xch_shm[0]) { if (d_xch) munmap(d_xch, xch_bytes); shm_unlink(xch_shm); }
#endif
  }
};

static bool use_device(Engine *e) {
#if IMA_CUDA
  return IMA_CUDA_OK(cudaSetDevice(e->device));
#else
  (void)e;
  return true;
#endif
}

static stream_t pick_stream(Engine *e, void *s) {
#if IMA_CUDA
  return s ? (cudaStream_t)s : e->own_stream;
#else
  (void)e; (void)s;
  return nullptr;
#endif
}

static int check_device_error(Engine *e, stream_t s) {
  int code = 0;
  if (!d2h(&code, e->d_err, sizeof(int), s) || !dev_sync(s)) return fail(IMA2P_E_CUDA, "device sync failed");
  if (code != 0) {
    char buf[160];
    snprintf(buf, sizeof buf, "device error word %d raised by a kernel (14 = LogDiff a<=b, 15 = incomplete gamma, 16 = logfact range)", code);
    return fail(IMA2P_E_DEVICE, buf);
  }
  return IMA2P_OK;
}

static int launch_eval(Engine *e, stream_t s) {
  const int gp = (e->d.P + kWarpsPerBlock - 1) / kWarpsPerBlock, gc = (e->d.nchains + kWarpsPerBlock - 1) / kWarpsPerBlock;
  IMA_LAUNCH(k_eval_pairs, gp, kWarpsPerBlock, e->pair_smem * kWarpsPerBlock, s, e->v);
  IMA_LAUNCH(k_eval_chains, gc, kWarpsPerBlock, e->chain_smem * kWarpsPerBlock, s, e->v);
#if IMA_CUDA
  if (!IMA_CUDA_OK(cudaGetLastError())) return fail(IMA2P_E_CUDA, "kernel launch failed (eval)");
#endif
  return IMA2P_OK;
}

static int accept_block_warps(int spec) { return IMA_CUDA ? spec * kTermWarps : 1; }

// the view a launch gets: which chains it covers and which step (relative to the device counter) it belongs to
static EngineView view_of(const Engine *e, int c_lo, int c_n, int step_off, int grp = 0) {
  EngineView v = e->v;
  v.c_lo = c_lo; v.c_n = c_n; v.step_off = step_off; v.grp = grp; v.redo_grid = e->redo_grid;
  v.xch.publisher = 0;
  return v;
}
static int pair_grid(const Engine *e, const EngineView &v) { return (v.c_n * e->d.nloci + kWarpsPerBlock - 1) / kWarpsPerBlock; }

// pairs per warp of k_move: few pairs -> few per warp (more warps in flight, less divergence); many -> fuller warps
static int move_ppw(const Engine *e, int npairs) {
#if IMA_CUDA
  if (e->ppw == 1 || e->ppw == 2 || e->ppw == 4 || e->ppw == 8 || e->ppw == 16) return e->ppw;
  return npairs <= 148 * 96 ? 4 : 8;
#else
  (void)e; (void)npairs;
  return 1;
#endif
}
static void launch_move(int ppw, int gm, size_t sm, stream_t s, const EngineView &v) {
#if IMA_CUDA
  if (ppw == 1) IMA_LAUNCH(k_move<1>, gm, kMoveWarps, sm, s, v);
  else if (ppw == 2) IMA_LAUNCH(k_move<2>, gm, kMoveWarps, sm, s, v);
  else if (ppw == 4) IMA_LAUNCH(k_move<4>, gm, kMoveWarps, sm, s, v);
  else if (ppw == 8) IMA_LAUNCH(k_move<8>, gm, kMoveWarps, sm, s, v);
  else IMA_LAUNCH(k_move<16>, gm, kMoveWarps, sm, s, v);
#else
  (void)ppw;
  IMA_LAUNCH(k_move<1>, gm, kMoveWarps, sm, s, v);
#endif
}
static void launch_propose(Engine *e, stream_t s, const EngineView &v) {
  if (!e->fast) {
    IMA_LAUNCH(k_propose, pair_grid(e, v), kWarpsPerBlock, e->pair_smem * kWarpsPerBlock, s, v);
    return;
  }
  const int npairs = v.c_n * e->d.nloci, ppw = move_ppw(e, npairs);
  const int gm = (npairs + ppw * kMoveWarps - 1) / (ppw * kMoveWarps);
  const size_t sm = move_smem_bytes_per_pair(e->d) * ppw * kMoveWarps;
  launch_move(ppw, gm, sm, s, v);
  IMA_LAUNCH(k_weigh, pair_grid(e, v), kWarpsPerBlock, weigh_smem_bytes(e->d) * kWarpsPerBlock, s, v);
  IMA_LAUNCH(k_propose_redo, e->redo_grid, kWarpsPerBlock, e->pair_smem * kWarpsPerBlock, s, v);
}
static void launch_accept(Engine *e, stream_t s, const EngineView &v) {
  const int nw = accept_block_warps(e->spec);
  if (e->spec >= 4) IMA_LAUNCH(k_accept<4>, v.c_n, nw, e->accept_smem, s, v);
  else if (e->spec == 3) IMA_LAUNCH(k_accept<3>, v.c_n, nw, e->accept_smem, s, v);
  else if (e->spec == 2) IMA_LAUNCH(k_accept<2>, v.c_n, nw, e->accept_smem, s, v);
  else IMA_LAUNCH(k_accept<1>, v.c_n, nw, e->accept_smem, s, v);
}
static size_t changeu_smem(const Engine *e) {
  const size_t need = changeu_smem_doubles(e->d.nloci) * sizeof(double);
  return need > e->pair_smem ? need : e->pair_smem;
}
static bool does_split_t(const Engine *e) { return e->t_updates && e->model.nsplit > 0; }
static bool does_changeu(const Engine *e) { return e->u_every > 0 && (e->uv.nurates > 1 || e->loci[0].d.model == kHKY); }
// the rest of qupdate's schedule (ima_main_mpi.cpp:1867-1945): a split-time update of every chain each step
// (TUPDATEINC 0), the mutation scalars every u_every-th step (UUPDATEINC 4); both are local to a chain.
// Every chain picks one of the two split-time updates (t_proposal); a warp whose chain picked the other one returns at once.
static void launch_split_t(Engine *e, stream_t s, const EngineView &v) {
  if (!e->fast) {
    IMA_LAUNCH(k_split_t, pair_grid(e, v), kWarpsPerBlock, e->pair_smem * kWarpsPerBlock, s, v, e->uv);
    return;
  }
  IMA_LAUNCH(k_split_t_fast, pair_grid(e, v), kWarpsPerBlock, split_smem_bytes(e->d) * kWarpsPerBlock, s, v, e->uv);
  IMA_LAUNCH(k_split_t_redo, e->redo_grid, kWarpsPerBlock, e->pair_smem * kWarpsPerBlock, s, v, e->uv);
}
static void launch_accept_t(Engine *e, stream_t s, const EngineView &v) {
  IMA_LAUNCH(k_accept_t, v.c_n, IMA_CUDA ? kTWarps : 1, accept_t_smem_bytes(e->d), s, v, e->uv);
}
// data whose likelihood is recomputed under a new scalar (HKY, stepwise, joint loci) and more than two scalars: the levelled walk
static bool changeu_by_levels(const Engine *e) {
  return (e->d.any_hky || e->d.any_sw) && e->uv.nurates > 2 && !e->uv.u_forced && !getenv("IMA2P_CHANGEU_IN_ORDER") &&
         changeu_levels_smem_bytes(e->d, e->uv.nurates) <= 200 * 1024;
}
static void launch_changeu(Engine *e, stream_t s, const EngineView &v) {
  UpdateView u = e->uv;
  u.u_every = e->u_every;
  if (changeu_by_levels(e)) {
    u.u_levels_in_order = getenv("IMA2P_CHANGEU_LEVELS_IN_ORDER") ? 1 : 0;
    IMA_LAUNCH(k_changeu_levels, v.c_n, IMA_CUDA ? kUWarps : 1, changeu_levels_smem_bytes(e->d, u.nurates), s, v, u);
    return;
  }
  IMA_LAUNCH(k_changeu, (v.c_n + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock, changeu_smem(e) * kWarpsPerBlock, s, v, u);
}
static void launch_param_updates(Engine *e, stream_t s, const EngineView &v) {
  if (does_split_t(e)) { launch_split_t(e, s, v); launch_accept_t(e, s, v); }
  if (does_changeu(e)) launch_changeu(e, s, v);
}
// one step of chains [c_lo, c_lo + c_n) up to, not including, the swaps
static void launch_update(Engine *e, stream_t s, const EngineView &v) {
  launch_propose(e, s, v);
  launch_accept(e, s, v);
  launch_param_updates(e, s, v);
}
static void launch_update(Engine *e, stream_t s) { launch_update(e, s, view_of(e, 0, e->d.nchains, 0)); }

// advance: what the launch adds to the device step counter (1 for a plain step, the number of steps at the end of a
// multi-step graph, 0 otherwise); step_off: the step it belongs to, relative to the counter
static void launch_swap(Engine *e, stream_t s, const double *S_global, int swaptries, int step_already_advanced = 0, int advance = 1, int step_off = 0) {
  SwapView sv = e->sv;
  sv.S_global = S_global;
  sv.sequential = getenv("IMA2P_SWAP_SEQUENTIAL") ? 1 : 0;
  sv.use_exchange = S_global == nullptr ? 1 : 0;             // a shard: the sums of all ranks come through the exchange tables
  sv.swaptries = swaptries;
  sv.advance_step = step_already_advanced ? 0 : advance;
  sv.step_bias = step_already_advanced ? 1 : 0;
  const int G = e->d.nchains_global;
  sv.smem_chains = G <= 4000 ? G : 0;                        // 24 bytes per chain, 96 KB opted in at finalize
  IMA_LAUNCH(k_swap, 1, 1, swap_smem_bytes(sv.smem_chains), s, view_of(e, 0, e->d.nchains, step_off), sv);
}

// which kernel ends a chain's step (and publishes its swap sum when the engine holds a shard)
static int last_kernel_of_step(const Engine *e) { return does_changeu(e) ? 3 : (does_split_t(e) ? 2 : 1); }

#if IMA_CUDA
// ---- the step as a CUDA graph --------------------------------------------------------------------------------------
// Chains only meet at the swaps, and the proposal half of a step does not read the temperatures.  The accept sweep is a
// long dependent chain per chain (loci in order) that leaves most of the GPU idle, so the chains are cut into `groups`
// groups, each on its own stream: the sweep of one group overlaps the proposals / split-time proposals of the others.
// A graph holds `depth` consecutive steps: group g's proposals of step j+1 start as soon as ITS step j is done (they
// need neither the other groups nor the swaps of step j); only its accept sweep waits for the swaps.  Every random
// stream is keyed by (chain, locus, step, purpose), so the run is bit for bit the plain one-stream run.
static cudaEvent_t pipe_event(Engine &e, size_t &next) {
  if (next == e.pipe_events.size()) {
    cudaEvent_t ev = nullptr;
    cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    e.pipe_events.push_back(ev);
  }
  return e.pipe_events[next++];
}
static bool capture_steps(Engine &e, int swaptries, int depth, cudaGraphExec_t *out, bool sharded = false) {
  struct Pdl { Pdl(bool on) { g_programmatic_launch = on; } ~Pdl() { g_programmatic_launch = false; } } pdl_scope(e.pdl != 0);
  int G = e.groups < 1 ? 1 : e.groups;
  if (G > e.d.nchains) G = e.d.nchains;
  if (G > kMaxGroups) G = kMaxGroups;
  for (int g = 0; g < G; g++)
    for (int k = 0; k < 2; k++)
      if (!e.group_stream[g][k]) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (!IMA_CUDA_OK(cudaStreamCreateWithPriority(&e.group_stream[g][k], cudaStreamNonBlocking, k ? hi : lo))) return false;
      }
  cudaStream_t s = e.own_stream;
  size_t nev = 0;
  cudaGraph_t graph = nullptr;
  if (!IMA_CUDA_OK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal))) return false;
  cudaEvent_t fork = pipe_event(e, nev);
  cudaEventRecord(fork, s);
  std::vector<cudaEvent_t> prev_swap(1, nullptr);
  cudaEvent_t swap_done = nullptr;
  for (int g = 0; g < G; g++) cudaStreamWaitEvent(e.group_stream[g][0], fork, 0);
  for (int j = 0; j < depth; j++) {
    std::vector<cudaEvent_t> done(G);
    for (int g = 0; g < G; g++) {
      const int c_lo = (int)((long long)e.d.nchains * g / G), c_hi = (int)((long long)e.d.nchains * (g + 1) / G);
      EngineView v = view_of(&e, c_lo, c_hi - c_lo, j, g);
      if (sharded) v.xch.publisher = last_kernel_of_step(&e);
      cudaStream_t sp = e.group_stream[g][0], sa = e.pipe_prio ? e.group_stream[g][1] : sp;   // proposals / decisions
      auto hop = [&](cudaStream_t from, cudaStream_t to) {
        if (from == to) return;
        cudaEvent_t ev = pipe_event(e, nev);
        cudaEventRecord(ev, from);
        cudaStreamWaitEvent(to, ev, 0);
      };
      launch_propose(&e, sp, v);
      hop(sp, sa);
      if (swap_done) cudaStreamWaitEvent(sa, swap_done, 0);          // the temperatures of this step
      launch_accept(&e, sa, v);
      if (does_split_t(&e)) {
        hop(sa, sp);
        launch_split_t(&e, sp, v);
        hop(sp, sa);
        launch_accept_t(&e, sa, v);
      }
      if (does_changeu(&e)) launch_changeu(&e, sa, v);
      done[g] = pipe_event(e, nev);
      cudaEventRecord(done[g], sa);
      if (sa != sp) cudaStreamWaitEvent(sp, done[g], 0);             // the group's next proposals follow its own step
    }
    for (int g = 0; g < G; g++) cudaStreamWaitEvent(s, done[g], 0);
    launch_swap(&e, s, sharded ? nullptr : e.v.swapsum, swaptries, 0, j == depth - 1 ? depth : 0, j);
    if (j + 1 < depth) { swap_done = pipe_event(e, nev); cudaEventRecord(swap_done, s); }
  }
  if (!IMA_CUDA_OK(cudaStreamEndCapture(s, &graph))) return false;
  const bool ok = IMA_CUDA_OK(cudaGraphInstantiate(out, graph, 0));
  cudaGraphDestroy(graph);
  return ok;
}
#endif

}  // namespace ima

using namespace ima;

struct ima2p_engine { Engine eng; };

extern "C" {

#if IMA_CUDA
const char *ima2p_version(void) { return "ima2p_b200 0.1 (sm_100a)"; }
#else
const char *ima2p_version(void) { return "ima2p_b200 0.1 (host emulation of the kernels: tests only, not a product build)"; }
#endif
const char *ima2p_last_error(void) { return g_last_error.c_str(); }
void ima2p_internal_set_error(const char *msg) { g_last_error = msg ? msg : ""; }

int ima2p_engine_create(ima2p_engine **out, int device, int nchains_local, int nchains_global, int chain0, int nloci,
                        int mig_capacity, uint64_t seed) {
  if (!out || nchains_local < 1 || nloci < 1 || nchains_global < nchains_local || chain0 < 0 ||
      chain0 + nchains_local > nchains_global || mig_capacity < 8 || mig_capacity > 8000)
    return fail(IMA2P_E_ARG, "ima2p_engine_create: bad argument");
#if IMA_CUDA
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    return fail(IMA2P_E_CUDA, "no CUDA device: ima2p_b200 has no CPU path");
  if (device < 0 || device >= ndev) return fail(IMA2P_E_ARG, "device index out of range");
#endif
  ima2p_engine *h = new (std::nothrow) ima2p_engine();
  if (!h) return fail(IMA2P_E_ARG, "out of host memory");
  Engine &e = h->eng;
  e.device = device;
  e.d.nchains = nchains_local; e.d.nchains_global = nchains_global; e.d.chain0 = chain0;
  e.d.nloci = nloci; e.d.P = nchains_local * nloci; e.d.CAP = mig_capacity;
  e.seed = seed;
  e.loci.resize(nloci);
  if (!use_device(&e)) { delete h; return fail(IMA2P_E_CUDA, "cudaSetDevice failed"); }
#if IMA_CUDA
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  if (!IMA_CUDA_OK(cudaStreamCreateWithPriority(&e.own_stream, cudaStreamNonBlocking, prio_hi))) { delete h; return fail(IMA2P_E_CUDA, "stream create failed"); }
#endif
  *out = h;
  return IMA2P_OK;
}

void ima2p_engine_destroy(ima2p_engine *h) {
  if (!h) return;
  use_device(&h->eng);
  delete h;
}

int ima2p_engine_set_model(ima2p_engine *h, int npops, int nsplit, const int *plist, const int *addpop, const int *droppops,
                           const int *pt_e, const int *pt_down, int rootpop, int nq, const int *q_off, const int *q_p,
                           const int *q_r, const double *q_max, const double *q_min, int nm, const int *m_off,
                           const int *m_p, const int *m_r, const int *m_c, const double *m_max, const double *m_min,
                           const double *m_mean, int nomig_n, const int *nomig_p, const int *nomig_r, const int *nomig_c,
                           int nomigration, int expoprior, int thermo, double gbeta) {
  if (!h) return fail(IMA2P_E_ARG, "null engine");
  Engine &e = h->eng;
  if (e.finalized) return fail(IMA2P_E_ARG, "set_model after finalize");
  if (npops < 1 || npops > kMaxPops || nsplit < 0 || nsplit > npops - 1 + (npops == 1) || nq < 1 || nq > kMaxParams || nm < 0 ||
      nm > kMaxParams || nomig_n < 0 || nomig_n > kMaxParams)
    return fail(IMA2P_E_ARG, "set_model: sizes out of range");
  DevModel &M = e.model;
  memset(&M, 0, sizeof M);
  M.npops = npops; M.nsplit = nsplit; M.ntreepops = 2 * npops - 1; M.rootpop = rootpop;
  M.nq = nq; M.nm = nm; M.nomigration = nomigration; M.expoprior = expoprior; M.thermo = thermo; M.gbeta = gbeta;
  for (int k = 0; k < npops; k++) for (int i = 0; i < npops; i++) M.plist[k][i] = (signed char)plist[k * npops + i];
  for (int k = 0; k <= nsplit; k++) {
    M.addpop[k] = (signed char)addpop[k];
    M.droppops[k][0] = (signed char)droppops[2 * k]; M.droppops[k][1] = (signed char)droppops[2 * k + 1];
  }
  for (int i = 0; i < M.ntreepops; i++) { M.pt_e[i] = (signed char)pt_e[i]; M.pt_down[i] = (signed char)pt_down[i]; }
  for (int q = 0; q < M.ntreepops; q++)          // q contributes to itself and to every ancestor
    for (int a = q, guard = 0; a >= 0 && guard < M.ntreepops; a = M.pt_down[a], guard++) M.desc_mask[a] |= 1 << q;
  M.cc_off[0] = M.mc_off[0] = 0;
  for (int k = 0; k <= nsplit; k++) {
    M.cc_off[k + 1] = (short)(M.cc_off[k] + (npops - k));
    M.mc_off[k + 1] = (short)(M.mc_off[k] + (npops - k) * (npops - k));
  }
  M.ncc = M.cc_off[nsplit + 1];
  M.nmc = M.mc_off[nsplit];
  for (int i = 0; i < nq; i++) {
    const int n = q_off[i + 1] - q_off[i];
    if (n < 0 || n > kMaxWp) return fail(IMA2P_E_ARG, "set_model: too many weight positions for a parameter");
    M.q_n[i] = (signed char)n;
    for (int j = 0; j < n; j++) M.q_idx[i][j] = (short)(M.cc_off[q_p[q_off[i] + j]] + q_r[q_off[i] + j]);
    M.q_max[i] = q_max[i]; M.q_min[i] = q_min[i];
  }
  for (int i = 0; i < nm; i++) {
    const int n = m_off[i + 1] - m_off[i];
    if (n < 0 || n > kMaxWp) return fail(IMA2P_E_ARG, "set_model: too many weight positions for a parameter");
    M.m_n[i] = (signed char)n;
    for (int j = 0; j < n; j++) {
      const int k = m_p[m_off[i] + j];
      M.m_idx[i][j] = (short)(M.mc_off[k] + m_r[m_off[i] + j] * (npops - k) + m_c[m_off[i] + j]);
    }
    M.m_max[i] = m_max[i]; M.m_min[i] = m_min[i]; M.m_mean[i] = m_mean ? m_mean[i] : 0.0;
  }
  M.nomig_n = nomig_n;
  for (int i = 0; i < nomig_n; i++) M.nomig_idx[i] = (short)(M.mc_off[nomig_p[i]] + nomig_r[i] * (npops - nomig_p[i]) + nomig_c[i]);
  e.d.NI = M.ncc + M.nmc;
  e.d.ND = 2 * M.ncc + M.nmc;
  e.d.NT = nq + nm;
  e.model_set = true;
  return IMA2P_OK;
}

int ima2p_engine_set_locus(ima2p_engine *h, int li, int model, int numgenes, int numsites, int totsites, double hval,
                           const int *samppop, const int *seq, const int *mult, int nlinked, const int *minA,
                           const int *maxA, double sumlogk) {
  if (!h) return fail(IMA2P_E_ARG, "null engine");
  Engine &e = h->eng;
  if (!e.model_set || e.finalized) return fail(IMA2P_E_ARG, "set_locus: call after set_model and before finalize");
  if (li < 0 || li >= e.d.nloci || numgenes < 2 || numgenes > 16000 || numsites < 0 || nlinked < 1 || nlinked > kMaxLinked)
    return fail(IMA2P_E_ARG, "set_locus: bad argument");
  if (model != kInfiniteSites && model != kHKY && model != kStepwise && model != kJointISSW) return fail(IMA2P_E_UNSUPPORTED, "set_locus: mutation model not supported");
  if (model == kJointISSW && nlinked < 2) return fail(IMA2P_E_ARG, "set_locus: a joint locus has an infinite-sites part and at least one stepwise part");
  HostLocus &L = e.loci[li];
  memset(&L.d, 0, sizeof L.d);
  L.d.model = model; L.d.ng = numgenes; L.d.nl = 2 * numgenes - 1; L.d.nsites = numsites; L.d.totsites = totsites;
  L.d.nwords = (numgenes + 31) / 32; L.d.nlinked = nlinked; L.d.hval = hval; L.d.sumlogk = sumlogk;
  L.d.hlog = log(hval); L.d.h2term = 1 / (2 * hval);
  int tot = 0;
  for (int i = 0; i < e.model.npops; i++) { L.d.samppop[i] = samppop[i]; tot += samppop[i]; }
  if (tot != numgenes) return fail(IMA2P_E_ARG, "set_locus: samppop does not sum to numgenes");
  for (int i = 0; i < nlinked; i++) { L.d.minA[i] = minA ? minA[i] : 0; L.d.maxA[i] = maxA ? maxA[i] : 0; }
  L.sitemask.clear(); L.seq.clear(); L.mult.clear();
  if (has_infinite_sites(model)) {
    if (numsites > 0 && !seq) return fail(IMA2P_E_ARG, "set_locus: seq required");
    L.sitemask.assign((size_t)numsites * L.d.nwords, 0u);
    for (int s = 0; s < numsites; s++) {
      int carriers = 0;
      for (int j = 0; j < numgenes; j++) {
        const int b = seq[(size_t)j * numsites + s];
        if (b != 0 && b != 1) return fail(IMA2P_E_ARG, "set_locus: infinite-sites data must be 0/1");
        if (b) { L.sitemask[(size_t)s * L.d.nwords + (j >> 5)] |= 1u << (j & 31); carriers++; }
      }
      // a monomorphic column is IMERR_INFINITESITESFAIL in the reference (calc_prob_data.cpp:823-826)
      if (carriers == 0 || carriers == numgenes) return fail(IMA2P_E_ARG, "set_locus: non-segregating infinite-sites column");
      // canonical key: a carrier set that contains gene 0 is stored as its complement (see build_tip_keys)
      if (L.sitemask[(size_t)s * L.d.nwords] & 1u)
        for (int w = 0; w < L.d.nwords; w++) {
          const uint32_t full = (w == L.d.nwords - 1 && (numgenes & 31)) ? ((1u << (numgenes & 31)) - 1u) : 0xffffffffu;
          L.sitemask[(size_t)s * L.d.nwords + w] = ~L.sitemask[(size_t)s * L.d.nwords + w] & full;
        }
    }
  } else if (model == kHKY) {
    if (!seq || !mult) return fail(IMA2P_E_ARG, "set_locus: seq and mult required for HKY");
    L.seq.resize((size_t)numgenes * numsites);
    for (size_t i = 0; i < L.seq.size(); i++) L.seq[i] = (unsigned char)seq[i];
    L.mult.assign(mult, mult + numsites);
  }
  L.set = true;
  return IMA2P_OK;
}

}  // extern "C"
// everything that depends on the migration capacity: table sizes, shared-memory footprints, launch attributes.  Called by
// finalize and again by ima2p_engine_grow_capacity.
static int size_for_capacity(Engine &e, int maxng) {
  EngineDims &d = e.d;
  int ev = (maxng - 1) + d.CAP + e.model.nsplit;
  d.EVP = 1; while (d.EVP < ev) d.EVP <<= 1;
  // the two-kernel proposal path: tables for FC migration events per genealogy (every shipped input stays far below; a pair
  // that does not fit takes the general path, whose tables hold CAP)
  d.FC = d.CAP < 64 ? d.CAP : 64;
  d.FP = 64;
  if (const char *x = getenv("IMA2P_FAST_EVENTS")) { const int v = atoi(x); if (v >= 8 && v <= d.CAP) d.FC = v; }
  if (const char *x = getenv("IMA2P_FAST_POOL")) { const int v = atoi(x); if (v >= 8 && v <= 4096) d.FP = v; }
  d.FS = d.FC;
  d.FEV = (maxng - 1) + d.FC + e.model.nsplit;
  e.fast_ok = !d.any_sw && d.NL <= 4096 && move_smem_bytes_per_pair(d) * 4 * kMoveWarps <= 200 * 1024 && weigh_smem_bytes(d) * kWarpsPerBlock <= 200 * 1024 && split_smem_bytes(d) * kWarpsPerBlock <= 200 * 1024;
  e.fast = e.fast_ok && !getenv("IMA2P_GENERAL_PATH");
  // measured on B200 (profiles/r2s1_pipeline_*.jsonl, r2s2_*, r2s6_paths.jsonl): two chain groups and four steps per graph; four groups
  // and two steps where the per-pair kernels run many waves (76,800 pairs: 2.065 against 2.109 ms a step)
  const bool many_waves = (long long)d.nchains * d.nloci >= 20000 && d.nchains >= 64;
  e.groups = many_waves ? 4 : (d.nchains >= 16 ? 2 : 1);
  e.depth = many_waves ? 2 : 4;
  e.pair_smem = pair_smem_bytes(d);
  e.chain_smem = chain_smem_bytes(d);
  e.accept_smem = accept_smem_bytes(d);
  // deep speculation needs a 15- or 20-warp block per chain (one per SM); with more chains than SMs two 10-warp blocks per SM win
  e.spec = d.nchains <= 148 ? 3 : 2;
#if IMA_CUDA
  if (e.pair_smem * kWarpsPerBlock > 227 * 1024) return fail(IMA2P_E_ARG, "finalize: pair does not fit in shared memory; lower mig_capacity");
  if (!IMA_CUDA_OK(cudaFuncSetAttribute(k_accept<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e.accept_smem)) ||
      !IMA_CUDA_OK(cudaFuncSetAttribute(k_accept<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e.accept_smem)) ||
      !IMA_CUDA_OK(cudaFuncSetAttribute(k_accept<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e.accept_smem)) ||
      !IMA_CUDA_OK(cudaFuncSetAttribute(k_accept<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e.accept_smem)) ||
      !IMA_CUDA_OK(cudaFuncSetAttribute(k_propose, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(e.pair_smem * kWarpsPerBlock))) ||
      !IMA_CUDA_OK(cudaFuncSetAttribute(k_propose_redo, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(e.pair_smem * kWarpsPerBlock))) ||
      !IMA_CUDA_OK(cudaFuncSetAttribute(k_eval_pairs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(e.pair_smem * kWarpsPerBlock))) ||
      !IMA_CUDA_OK(cudaFuncSetAttribute(k_split_t, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(e.pair_smem * kWarpsPerBlock))) ||
      !IMA_CUDA_OK(cudaFuncSetAttribute(k_swap, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)swap_smem_bytes(4000))) ||
      !IMA_CUDA_OK(cudaFuncSetAttribute(k_accept_t, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)accept_t_smem_bytes(d))) ||
      !IMA_CUDA_OK(cudaFuncSetAttribute(k_changeu_levels, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) ||
      !IMA_CUDA_OK(cudaFuncSetAttribute(k_changeu, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(changeu_smem(&e) * kWarpsPerBlock))))
    return fail(IMA2P_E_CUDA, "cudaFuncSetAttribute failed");
  if (e.fast_ok) {
    const int per = (int)(move_smem_bytes_per_pair(d) * kMoveWarps);
    const int cap = 227 * 1024;
    if (!IMA_CUDA_OK(cudaFuncSetAttribute(k_move<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, per < cap ? per : cap)) ||
        !IMA_CUDA_OK(cudaFuncSetAttribute(k_move<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, per * 2 < cap ? per * 2 : cap)) ||
        !IMA_CUDA_OK(cudaFuncSetAttribute(k_move<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, per * 4 < cap ? per * 4 : cap)) ||
        !IMA_CUDA_OK(cudaFuncSetAttribute(k_move<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, per * 8 < cap ? per * 8 : cap)) ||
        !IMA_CUDA_OK(cudaFuncSetAttribute(k_move<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, per * 16 < cap ? per * 16 : cap)) ||
        !IMA_CUDA_OK(cudaFuncSetAttribute(k_weigh, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(weigh_smem_bytes(d) * kWarpsPerBlock))) ||
        !IMA_CUDA_OK(cudaFuncSetAttribute(k_split_t_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(split_smem_bytes(d) * kWarpsPerBlock))) ||
        !IMA_CUDA_OK(cudaFuncSetAttribute(k_split_t_redo, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(e.pair_smem * kWarpsPerBlock))))
      return fail(IMA2P_E_CUDA, "cudaFuncSetAttribute failed (fast path)");
  }
  if (!IMA_CUDA_OK(cudaMemcpyToSymbol(c_model, &e.model, sizeof(DevModel)))) return fail(IMA2P_E_CUDA, "model upload failed");
#else
  c_model = e.model;
#endif
  return IMA2P_OK;
}

extern "C" {
int ima2p_engine_finalize(ima2p_engine *h) {
  if (!h) return fail(IMA2P_E_ARG, "null engine");
  Engine &e = h->eng;
  if (!e.model_set || e.finalized) return fail(IMA2P_E_ARG, "finalize: model not set or already finalized");
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  EngineDims &d = e.d;
  d.NL = 0; d.W = 1; d.S = 0; d.any_sw = 0; d.any_hky = 0;
  int maxng = 0;
  std::vector<uint32_t> sm; std::vector<unsigned char> sq; std::vector<int> mu; std::vector<DevLocus> dl(d.nloci);
  for (int li = 0; li < d.nloci; li++) {
    HostLocus &L = e.loci[li];
    if (!L.set) return fail(IMA2P_E_ARG, "finalize: a locus was not set");
    if (L.d.nl > d.NL) d.NL = L.d.nl;
    if (L.d.nwords > d.W) d.W = L.d.nwords;
    if (L.d.nsites > d.S) d.S = L.d.nsites;
    if (L.d.ng > maxng) maxng = L.d.ng;
    if (has_stepwise(L.d.model)) d.any_sw = 1;
    if (L.d.model == kHKY) d.any_hky = 1;
    L.d.sitemask_off = (long long)sm.size(); sm.insert(sm.end(), L.sitemask.begin(), L.sitemask.end());
    L.d.seq_off = (long long)sq.size(); sq.insert(sq.end(), L.seq.begin(), L.seq.end());
    L.d.mult_off = (long long)mu.size(); mu.insert(mu.end(), L.mult.begin(), L.mult.end());
    dl[li] = L.d;
  }
  if (d.NL > 32000) return fail(IMA2P_E_ARG, "finalize: sample too large for 16-bit edge indices");
  d.W64 = (e.model.ntreepops + 3) / 4;
  e.maxng = maxng;
  { const int rc = size_for_capacity(e, maxng); if (rc) return rc; }
  const size_t P = d.P, NL = d.NL, CAP = d.CAP, C = d.nchains, G = d.nchains_global;
  // logfact table: same running sum as setlogfact (utilities.cpp:1405-1414)
  const int nlf = 100 * 5000 + 1;
  std::vector<double> lf(nlf);
  lf[0] = 0;
  for (int i = 1; i < nlf; i++) lf[i] = lf[i - 1] + log((double)i);
  e.d_logfact = e.alloc<double>(nlf);
  e.d_err = e.alloc<int>(1);
  e.d_loci = e.alloc<DevLocus>(d.nloci);
  uint32_t *d_sm = e.alloc<uint32_t>(sm.size() + 1);
  unsigned char *d_sq = e.alloc<unsigned char>(sq.size() + 1);
  int *d_mu = e.alloc<int>(mu.size() + 1);
  {
    // the tables the event sweep indexes per lane (EngineDims::tab)
    std::vector<int> tab(kTabInts, 0);
    const DevModel &M = e.model;
    for (int i = 0; i < M.ntreepops; i++) tab[kTabDesc + i] = M.desc_mask[i];
    for (int k = 0; k <= M.nsplit + 1 && k < kMaxPeriods + 1; k++) { tab[kTabCcOff + k] = M.cc_off[k]; tab[kTabMcOff + k] = M.mc_off[k]; }
    for (int k = 0; k < M.npops; k++) for (int i = 0; i < M.npops; i++) tab[kTabPlist + k * kMaxPops + i] = M.plist[k][i];
    int *d_tab = e.alloc<int>(kTabInts);
    if (!d_tab || !h2d(d_tab, tab.data(), kTabInts * sizeof(int), pick_stream(&e, nullptr)) || !dev_sync(pick_stream(&e, nullptr)))
      return fail(IMA2P_E_CUDA, "table upload failed");
    d.tab = d_tab;
  }
  EngineView &v = e.v;
  v.d = d;
  for (int b = 0; b < 2; b++) {
    PairBuf &B = v.buf[b];
    B.topo = e.alloc<short4_t>(P * NL); B.time = e.alloc<double>(P * NL); B.mseg = e.alloc<ushort2_t>(P * NL);
    B.mig_t = e.alloc<double>(P * CAP); B.mig_p = e.alloc<short>(P * CAP);
    B.sd = e.alloc<double>(P * 4); B.si = e.alloc<int>(P * 2);
    B.gwi = e.alloc<int>(P * d.NI); B.gwd = e.alloc<double>(P * d.ND);
    if (d.any_sw) { B.A = e.alloc<short>(P * kMaxLinked * NL); B.dlikeA = e.alloc<double>(P * kMaxLinked * NL); B.pdg_a = e.alloc<double>(P * kMaxLinked); }
    else { B.A = nullptr; B.dlikeA = nullptr; B.pdg_a = nullptr; }
    B.hky_mask = nullptr;
  }
  if (d.any_hky) {
    int hs = 0, hg = 0;
    for (int li = 0; li < d.nloci; li++) if (e.loci[li].d.model == kHKY) { if (e.loci[li].d.nsites > hs) hs = e.loci[li].d.nsites; if (e.loci[li].d.ng > hg) hg = e.loci[li].d.ng; }
    v.d.hky_sites = d.hky_sites = hs;
    v.d.hky_mask_words = d.hky_mask_words = (hg - 1 + 31) / 32;
    v.d.hky_stride = d.hky_stride = (long long)(hg - 1) * 2 * 5 * hs;
    v.hky_frac = e.alloc<double>((size_t)P * d.hky_stride);
    v.prop_ids = e.alloc<short>((size_t)P * 2);
    for (int b = 0; b < 2; b++) v.buf[b].hky_mask = e.alloc<uint32_t>((size_t)P * d.hky_mask_words);
    if (!v.hky_frac || !v.prop_ids || !v.buf[1].hky_mask) return fail(IMA2P_E_CUDA, "device allocation failed (HKY partial likelihoods)");
  }
  v.cur = e.alloc<unsigned char>(P);
  v.uvals = e.alloc<double>(P * kMaxLinked); v.kappa = e.alloc<double>(P); v.pi = e.alloc<double>(P * 4);
  v.tvals = e.alloc<double>(C * kMaxPeriods); v.beta = e.alloc<double>(C);
  v.all_i = e.alloc<int>(C * d.NI); v.all_d = e.alloc<double>(C * d.ND);
  v.qint = e.alloc<double>(C * kMaxParams); v.mint = e.alloc<double>(C * kMaxParams);
  v.probg = e.alloc<double>(C); v.pdgsum = e.alloc<double>(C); v.swapsum = e.alloc<double>(C);
  v.prop_extra = e.alloc<double>(P); v.prop_flags = e.alloc<uint32_t>(P); v.prop_dbg = nullptr;   // see ima2p_engine_set_debug_records
  v.redo_count = e.alloc<int>(kMaxGroups * 4); v.redo_list = e.alloc<int>(2 * P);
  if (!v.redo_count || !v.redo_list) return fail(IMA2P_E_CUDA, "device allocation failed");
  v.acc = e.alloc<unsigned int>(P * 3); v.cold_acc = e.alloc<unsigned int>((size_t)d.nloci * 3);
  v.nsteps = e.alloc<unsigned long long>(1); v.overflow = e.alloc<unsigned long long>(1);
  v.seed = e.seed;
  v.c_lo = 0; v.c_n = d.nchains; v.step_off = 0; v.grp = 0; v.redo_grid = e.redo_grid;
  v.loci = e.d_loci; v.sitemask = d_sm; v.seq = d_sq; v.mult = d_mu;
  v.mc.logfact = e.d_logfact; v.mc.logfact_n = nlf; v.mc.err = e.d_err;
  e.sv.rank_of_chain = e.alloc<int>(G); e.sv.chain_of_rank = e.alloc<int>(G);
  e.d_beta_table = e.alloc<double>(G); e.sv.beta_table = e.d_beta_table;
  e.d_swap_counts = e.alloc<unsigned long long>(2); e.sv.swap_counts = e.d_swap_counts;
  e.sv.adj_counts = e.alloc<unsigned long long>((size_t)G * 2);
  e.d_thermosum = e.alloc<double>(G);
  {
    // mutation-rate scalars in the order of readata.cpp:832-834; update parameters start at the reference's defaults
    for (int li = 0; li < d.nloci; li++) for (int a = 0; a < e.loci[li].d.nlinked; a++) { e.h_ul_l.push_back(li); e.h_ul_a.push_back(a); }
    UpdateView &u = e.uv;
    u.nurates = (int)e.h_ul_l.size();
    int *ul_l = e.alloc<int>(u.nurates), *ul_a = e.alloc<int>(u.nurates);
    u.ul_l = ul_l; u.ul_a = ul_a;
    u.t_counts = e.alloc<int>(P * 4); u.t_out = e.alloc<double>(C * 4); u.u_out = e.alloc<double>(C * 4);
    u.stats = e.alloc<unsigned long long>(update_stats_len(u.nurates));
    if (!u.stats || !u.u_out) return fail(IMA2P_E_CUDA, "device allocation failed");
    stream_t s0 = pick_stream(&e, nullptr);
    if (!h2d(ul_l, e.h_ul_l.data(), u.nurates * sizeof(int), s0) || !h2d(ul_a, e.h_ul_a.data(), u.nurates * sizeof(int), s0) || !dev_sync(s0))
      return fail(IMA2P_E_CUDA, "table upload failed");
    for (int k = 0; k < kMaxPeriods; k++) { u.t_max[k] = kTimeMax; u.t_min[k] = 0.0; }
    const double umax = log(10000.0);                      // UMAX, imamp.hpp:152; initialize.cpp:1448-1450, update_mc_params.cpp:52-56
    u.u_win = umax / d.nloci; u.u_maxratio = 3.0 * umax;
    u.kappa_win = 2.0; u.kappa_max = 100.0;                // initialize.cpp:1499-1501
    u.t_forced = nullptr; u.t_forced_period = 0; u.t_force_accept = -1; u.t_forced_method = 0; u.t_methods = 1; u.u_forced = 0; u.u_every = 5;
  }
  if (!v.acc || !v.cold_acc || !e.sv.adj_counts || !v.buf[1].gwd || !e.d_swap_counts || !v.overflow) return fail(IMA2P_E_CUDA, "device allocation failed");
  stream_t s = pick_stream(&e, nullptr);
  bool ok = h2d(e.d_logfact, lf.data(), nlf * sizeof(double), s) && h2d(e.d_loci, dl.data(), dl.size() * sizeof(DevLocus), s);
  if (!sm.empty()) ok = ok && h2d(d_sm, sm.data(), sm.size() * sizeof(uint32_t), s);
  if (!sq.empty()) ok = ok && h2d(d_sq, sq.data(), sq.size(), s);
  if (!mu.empty()) ok = ok && h2d(d_mu, mu.data(), mu.size() * sizeof(int), s);
  if (!ok || !dev_sync(s)) return fail(IMA2P_E_CUDA, "table upload failed");
  // host staging
  e.h_topo.assign(P * NL, short4_t{-1, -1, -1, -1});
  e.h_time.assign(P * NL, 0.0); e.h_mseg.assign(P * NL, ushort2_t{0, 0});
  e.h_mig_t.assign(P * CAP, 0.0); e.h_mig_p.assign(P * CAP, 0);
  e.h_sd.assign(P * 4, 0.0); e.h_si.assign(P * 2, 0);
  e.h_uvals.assign(P * kMaxLinked, 1.0); e.h_kappa.assign(P, 2.0); e.h_pi.assign(P * 4, 0.25);
  e.h_tvals.assign(C * kMaxPeriods, kTimeMax);
  if (d.any_sw) e.h_A.assign(P * kMaxLinked * NL, 0);
  e.h_beta_table.assign(G, 1.0);
  e.h_rank_of_chain.resize(G); e.h_chain_of_rank.resize(G);
  for (size_t i = 0; i < G; i++) e.h_rank_of_chain[i] = e.h_chain_of_rank[i] = (int)i;
  e.finalized = true;
  return ima2p_engine_set_betas(h, e.h_beta_table.data());
}

// The migration capacity of a running engine: the reference grows an edge's list whenever it fills (checkmig,
// utilities.cpp:1365-1383, up to ABSMIGMAX 5000 per edge).  Here the pools are rows of `mig_capacity` events per genealogy; a
// caller that sees proposals dropped for capacity (ima2p_engine_counters, field 7) grows the rows at a step boundary with this
// call: new pools, the resident genealogies copied over, every table and launch footprint re-sized.  The chains are unchanged.
int ima2p_engine_grow_capacity(ima2p_engine *h, int new_capacity) {
  if (!h || !h->eng.finalized) return fail(IMA2P_E_ARG, "grow_capacity: not finalized");
  Engine &e = h->eng;
  if (new_capacity <= e.d.CAP) return IMA2P_OK;
  if (new_capacity > 8000) return fail(IMA2P_E_CAPACITY, "grow_capacity: more than 8000 migration events per genealogy");
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, nullptr);
#if IMA_CUDA
  if (!IMA_CUDA_OK(cudaDeviceSynchronize())) return fail(IMA2P_E_CUDA, "device synchronize failed");
#endif
  const size_t P = e.d.P, oldc = e.d.CAP, newc = (size_t)new_capacity;
  const EngineDims keep_d = e.d;
  const int keep_groups = e.groups, keep_depth = e.depth, keep_spec = e.spec, keep_ppw = e.ppw;
  const bool keep_fast = e.fast;
  e.d.CAP = new_capacity;
  int rc = size_for_capacity(e, e.maxng);
  if (rc) { e.d = keep_d; size_for_capacity(e, e.maxng); return fail(IMA2P_E_CAPACITY, "grow_capacity: a genealogy with that many migration events does not fit in shared memory"); }
  e.groups = keep_groups; e.depth = keep_depth; e.spec = keep_spec; e.ppw = keep_ppw; e.fast = keep_fast && e.fast_ok;
  for (int b = 0; b < 2; b++) {
    PairBuf &B = e.v.buf[b];
    double *nt = e.alloc<double>(P * newc);
    short *np_ = e.alloc<short>(P * newc);
    if (!nt || !np_) return fail(IMA2P_E_CUDA, "device allocation failed (grow_capacity)");
#if IMA_CUDA
    if (!IMA_CUDA_OK(cudaMemcpy2DAsync(nt, newc * 8, B.mig_t, oldc * 8, oldc * 8, P, cudaMemcpyDeviceToDevice, s)) ||
        !IMA_CUDA_OK(cudaMemcpy2DAsync(np_, newc * 2, B.mig_p, oldc * 2, oldc * 2, P, cudaMemcpyDeviceToDevice, s)))
      return fail(IMA2P_E_CUDA, "copy failed (grow_capacity)");
#else
    for (size_t p = 0; p < P; p++) { memcpy(nt + p * newc, B.mig_t + p * oldc, oldc * 8); memcpy(np_ + p * newc, B.mig_p + p * oldc, oldc * 2); }
#endif
    B.mig_t = nt; B.mig_p = np_;           // the old pools stay allocated until the engine is destroyed (they are small)
  }
  if (!dev_sync(s)) return fail(IMA2P_E_CUDA, "sync failed (grow_capacity)");
  e.v.d = e.d;
  // host staging rows (set_genealogy / upload) follow the new pitch
  std::vector<double> ht(P * newc, 0.0); std::vector<short> hp(P * newc, 0);
  for (size_t p = 0; p < P; p++) { memcpy(&ht[p * newc], &e.h_mig_t[p * oldc], oldc * 8); memcpy(&hp[p * newc], &e.h_mig_p[p * oldc], oldc * 2); }
  e.h_mig_t.swap(ht); e.h_mig_p.swap(hp);
  // the one-block upload staging is re-made at the new size on its next use (a block uploaded for the old size is dropped)
  e.block_cap = 0; e.d_block[0] = e.d_block[1] = nullptr; e.block_pending = 0; e.block_next_up = e.block_next_adopt = 0;
  e.graph_ready = false;
#if IMA_CUDA
  if (e.graph_exec_sh) { cudaGraphExecDestroy(e.graph_exec_sh); e.graph_exec_sh = nullptr; }
  if (e.graph_exec_sh_deep) { cudaGraphExecDestroy(e.graph_exec_sh_deep); e.graph_exec_sh_deep = nullptr; }
#endif
  return IMA2P_OK;
}

int ima2p_engine_set_betas(ima2p_engine *h, const double *betas_global) {
  if (!h || !h->eng.finalized || !betas_global) return fail(IMA2P_E_ARG, "set_betas: bad argument / not finalized");
  Engine &e = h->eng;
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  const int G = e.d.nchains_global;
  // beta_table[r] = r-th largest beta; chain i starts at the rank of its own beta (setheat: slot i <-> index i)
  std::vector<int> order(G);
  for (int i = 0; i < G; i++) order[i] = i;
  for (int i = 1; i < G; i++) { int k = order[i], j = i; while (j > 0 && betas_global[order[j - 1]] < betas_global[k]) { order[j] = order[j - 1]; j--; } order[j] = k; }
  for (int r = 0; r < G; r++) { e.h_beta_table[r] = betas_global[order[r]]; e.h_chain_of_rank[r] = order[r]; e.h_rank_of_chain[order[r]] = r; }
  std::vector<double> local(e.d.nchains);
  for (int c = 0; c < e.d.nchains; c++) local[c] = betas_global[e.d.chain0 + c];
  stream_t s = pick_stream(&e, nullptr);
  bool ok = h2d(e.d_beta_table, e.h_beta_table.data(), G * sizeof(double), s) &&
            h2d(e.sv.rank_of_chain, e.h_rank_of_chain.data(), G * sizeof(int), s) &&
            h2d(e.sv.chain_of_rank, e.h_chain_of_rank.data(), G * sizeof(int), s) &&
            h2d(e.v.beta, local.data(), local.size() * sizeof(double), s);
  if (!ok || !dev_sync(s)) return fail(IMA2P_E_CUDA, "beta upload failed");
  return IMA2P_OK;
}

int ima2p_engine_set_heating(ima2p_engine *h, int heatmode, double hval1, double hval2) {
  if (!h || !h->eng.finalized) return fail(IMA2P_E_ARG, "set_heating: not finalized");
  const int G = h->eng.d.nchains_global;
  std::vector<double> b(G, 1.0);
  // beta[] of setheat (swapchains.cpp:119-163), N = chains over all ranks, one table for everything
  // (the reference's allbetas[] inconsistencies are not copied)
  if (G > 1)
    for (int ci = 0; ci < G; ci++) {
      if (heatmode == 0) { const double h1 = hval1 < 0.05 ? 0.05 : hval1; b[ci] = 1.0 / (1.0 + h1 * ci); }
      else if (heatmode == 1) b[ci] = 1 - (1 - hval2) * (ci)*pow(hval1, (double)(G - 1 - (ci))) / (double)(G - 1);
      else if (heatmode == 2) b[ci] = 1.0 - ci * (1.0 / (G - 1));
      else return fail(IMA2P_E_ARG, "set_heating: heatmode must be 0, 1 or 2");
      const bool bad = h->eng.model.thermo ? (b[ci] < 0.0 || b[ci] > 1.0) : (b[ci] <= 0.0 || b[ci] > 1.0);
      if (bad) return fail(IMA2P_E_ARG, "set_heating: heating terms give a beta out of range (IMERR_COMMANDLINEHEATINGTERMS)");
    }
  return ima2p_engine_set_betas(h, b.data());
}

int ima2p_engine_set_chain(ima2p_engine *h, int ci, const double *tvals) {
  if (!h || !h->eng.finalized || ci < 0 || ci >= h->eng.d.nchains || !tvals) return fail(IMA2P_E_ARG, "set_chain: bad argument");
  Engine &e = h->eng;
  for (int k = 0; k < kMaxPeriods; k++) e.h_tvals[(size_t)ci * kMaxPeriods + k] = k < e.model.nsplit ? tvals[k] : kTimeMax;
  return IMA2P_OK;
}

int ima2p_engine_set_genealogy(ima2p_engine *h, int ci, int li, const int *up0, const int *up1, const int *down, const int *pop,
                               const double *time, const int *mig_off, const double *mig_t, const int *mig_p, int root,
                               double roottime, const double *uvals, double kappa, const double *pi, const int *A) {
  if (!h || !h->eng.finalized) return fail(IMA2P_E_ARG, "set_genealogy: not finalized");
  Engine &e = h->eng;
  if (ci < 0 || ci >= e.d.nchains || li < 0 || li >= e.d.nloci) return fail(IMA2P_E_ARG, "set_genealogy: index out of range");
  const DevLocus &L = e.loci[li].d;
  const size_t p = (size_t)ci * e.d.nloci + li, NL = e.d.NL, CAP = e.d.CAP;
  const int total = mig_off[L.nl] - mig_off[0];
  if (total > e.d.CAP) return fail(IMA2P_E_CAPACITY, "set_genealogy: more migration events than mig_capacity");
  if (root < 0 || root >= L.nl || down[root] != -1) return fail(IMA2P_E_ARG, "set_genealogy: bad root");
  // a state file is outside data: nothing read from it is used as an index unchecked
  if (root < L.ng) return fail(IMA2P_E_ARG, "set_genealogy: the root is a tip");
  for (int i = 0; i < L.nl; i++) {
    const bool tip = i < L.ng;
    if (down[i] < -1 || down[i] >= L.nl || (down[i] >= 0 && down[i] < L.ng)) return fail(IMA2P_E_ARG, "set_genealogy: edge link out of range");
    if (tip ? (up0[i] != -1 || up1[i] != -1) : (up0[i] < 0 || up0[i] >= L.nl || up1[i] < 0 || up1[i] >= L.nl || up0[i] == up1[i]))
      return fail(IMA2P_E_ARG, "set_genealogy: edge link out of range");
    if (pop[i] < 0 || pop[i] >= e.model.ntreepops) return fail(IMA2P_E_ARG, "set_genealogy: population out of range");
    if (mig_off[i + 1] < mig_off[i]) return fail(IMA2P_E_ARG, "set_genealogy: migration offsets must not decrease");
  }
  for (int j = 0; j < total; j++)
    if (mig_p[mig_off[0] + j] < 0 || mig_p[mig_off[0] + j] >= e.model.ntreepops) return fail(IMA2P_E_ARG, "set_genealogy: migration target out of range");
  for (int i = 0; i < L.nl; i++) {
    e.h_topo[p * NL + i] = short4_t{(short)up0[i], (short)up1[i], (short)down[i], (short)pop[i]};
    e.h_time[p * NL + i] = time[i];
    e.h_mseg[p * NL + i] = ushort2_t{(unsigned short)(mig_off[i] - mig_off[0]), (unsigned short)(mig_off[i + 1] - mig_off[i])};
  }
  for (int j = 0; j < total; j++) { e.h_mig_t[p * CAP + j] = mig_t[mig_off[0] + j]; e.h_mig_p[p * CAP + j] = (short)mig_p[mig_off[0] + j]; }
  e.h_si[p * 2] = root; e.h_si[p * 2 + 1] = total;
  e.h_sd[p * 4] = roottime;
  for (int a = 0; a < L.nlinked; a++) e.h_uvals[p * kMaxLinked + a] = uvals ? uvals[a] : 1.0;
  e.h_kappa[p] = kappa;
  if (pi) for (int k = 0; k < 4; k++) e.h_pi[p * 4 + k] = pi[k];
  if (A && e.d.any_sw) for (int a = 0; a < L.nlinked; a++) for (int i = 0; i < L.nl; i++) e.h_A[(p * kMaxLinked + a) * NL + i] = (short)A[(size_t)a * L.nl + i];
  return IMA2P_OK;
}

static int upload_all(Engine &e, stream_t s) {
  const size_t P = e.d.P, NL = e.d.NL, CAP = e.d.CAP;
  PairBuf &B = e.v.buf[0];
  bool ok = h2d(B.topo, e.h_topo.data(), P * NL * sizeof(short4_t), s) && h2d(B.time, e.h_time.data(), P * NL * sizeof(double), s) &&
            h2d(B.mseg, e.h_mseg.data(), P * NL * sizeof(ushort2_t), s) && h2d(B.mig_t, e.h_mig_t.data(), P * CAP * sizeof(double), s) &&
            h2d(B.mig_p, e.h_mig_p.data(), P * CAP * sizeof(short), s) && h2d(B.sd, e.h_sd.data(), P * 4 * sizeof(double), s) &&
            h2d(B.si, e.h_si.data(), P * 2 * sizeof(int), s) && h2d(e.v.uvals, e.h_uvals.data(), P * kMaxLinked * sizeof(double), s) &&
            h2d(e.v.kappa, e.h_kappa.data(), P * sizeof(double), s) && h2d(e.v.pi, e.h_pi.data(), P * 4 * sizeof(double), s) &&
            h2d(e.v.tvals, e.h_tvals.data(), e.h_tvals.size() * sizeof(double), s);
  if (e.d.any_sw) ok = ok && h2d(B.A, e.h_A.data(), e.h_A.size() * sizeof(short), s) && h2d(e.v.buf[1].A, e.h_A.data(), e.h_A.size() * sizeof(short), s);
#if IMA_CUDA
  ok = ok && IMA_CUDA_OK(cudaMemsetAsync(e.v.cur, 0, P, s));
#else
  memset(e.v.cur, 0, P);
#endif
  return ok ? IMA2P_OK : fail(IMA2P_E_CUDA, "state upload failed");
}

int ima2p_engine_upload(ima2p_engine *h) {
  if (!h || !h->eng.finalized) return fail(IMA2P_E_ARG, "upload: not finalized");
  Engine &e = h->eng;
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, nullptr);
  int rc = upload_all(e, s);
  if (rc) return rc;
  return dev_sync(s) ? IMA2P_OK : fail(IMA2P_E_CUDA, "sync failed");
}

int ima2p_engine_eval(ima2p_engine *h) {
  if (!h || !h->eng.finalized) return fail(IMA2P_E_ARG, "eval: not finalized");
  Engine &e = h->eng;
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, nullptr);
  int rc = launch_eval(&e, s);
  if (rc) return rc;
  rc = check_device_error(&e, s);
  if (rc) return rc;
  // a state the kernels could not evaluate is an input error, not a soft failure
  std::vector<uint32_t> fl(e.d.P);
  if (!d2h(fl.data(), e.v.prop_flags, fl.size() * sizeof(uint32_t), s) || !dev_sync(s)) return fail(IMA2P_E_CUDA, "flag download failed");
  for (size_t p = 0; p < fl.size(); p++) {
    if (fl[p] & kFlagBadTree) return fail(IMA2P_E_ARG, "eval: inconsistent genealogy (lineage counts do not close)");
    if (fl[p] & kFlagOverflow) return fail(IMA2P_E_CAPACITY, "eval: event table capacity exceeded");
  }
  return IMA2P_OK;
}

int ima2p_engine_dims(ima2p_engine *h, int *out) {
  if (!h || !h->eng.model_set || !out) return fail(IMA2P_E_ARG, "dims: bad argument");
  Engine &e = h->eng;
  out[0] = e.d.NI; out[1] = e.d.ND; out[2] = e.d.NL; out[3] = e.d.CAP;
  out[4] = 4 * e.model.nq + (e.model.nomigration ? 0 : 3 * e.model.nm) + e.model.nsplit + 2;   // calc_gsampinf_length ginfo.cpp:306-316
  return IMA2P_OK;
}

int ima2p_engine_get_pair(ima2p_engine *h, int ci, int li, int *wi, double *wd, double *out_d, int *out_i) {
  if (!h || !h->eng.finalized) return fail(IMA2P_E_ARG, "get_pair: not finalized");
  Engine &e = h->eng;
  if (ci < 0 || ci >= e.d.nchains || li < 0 || li >= e.d.nloci) return fail(IMA2P_E_ARG, "get_pair: index out of range");
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, nullptr);
  const size_t p = (size_t)ci * e.d.nloci + li;
  unsigned char cur = 0;
  if (!d2h(&cur, e.v.cur + p, 1, s) || !dev_sync(s)) return fail(IMA2P_E_CUDA, "download failed");
  const PairBuf &B = e.v.buf[cur];
  double sd[4]; int si[2];
  bool ok = d2h(sd, B.sd + p * 4, sizeof sd, s) && d2h(si, B.si + p * 2, sizeof si, s);
  if (wi) ok = ok && d2h(wi, B.gwi + p * e.d.NI, e.d.NI * sizeof(int), s);
  if (wd) ok = ok && d2h(wd, B.gwd + p * e.d.ND, e.d.ND * sizeof(double), s);
  if (!ok || !dev_sync(s)) return fail(IMA2P_E_CUDA, "download failed");
  if (out_d) { out_d[0] = sd[3]; out_d[1] = sd[1]; out_d[2] = sd[2]; out_d[3] = sd[0]; }
  if (out_i) { out_i[0] = si[1]; out_i[1] = si[0]; }
  return IMA2P_OK;
}

int ima2p_engine_get_chain(ima2p_engine *h, int ci, int *all_wi, double *all_wd, double *qintegrate, double *mintegrate, double *out_d) {
  if (!h || !h->eng.finalized) return fail(IMA2P_E_ARG, "get_chain: not finalized");
  Engine &e = h->eng;
  if (ci < 0 || ci >= e.d.nchains) return fail(IMA2P_E_ARG, "get_chain: index out of range");
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, nullptr);
  bool ok = true;
  if (all_wi) ok = ok && d2h(all_wi, e.v.all_i + (size_t)ci * e.d.NI, e.d.NI * sizeof(int), s);
  if (all_wd) ok = ok && d2h(all_wd, e.v.all_d + (size_t)ci * e.d.ND, e.d.ND * sizeof(double), s);
  if (qintegrate) ok = ok && d2h(qintegrate, e.v.qint + (size_t)ci * kMaxParams, e.model.nq * sizeof(double), s);
  if (mintegrate && e.model.nm) ok = ok && d2h(mintegrate, e.v.mint + (size_t)ci * kMaxParams, e.model.nm * sizeof(double), s);
  if (out_d) ok = ok && d2h(out_d, e.v.probg + ci, sizeof(double), s) && d2h(out_d + 1, e.v.pdgsum + ci, sizeof(double), s) && d2h(out_d + 2, e.v.beta + ci, sizeof(double), s);
  if (!ok || !dev_sync(s)) return fail(IMA2P_E_CUDA, "download failed");
  return IMA2P_OK;
}

int ima2p_engine_get_genealogy(ima2p_engine *h, int ci, int li, int which, int *up0, int *up1, int *down, int *pop, double *time, int *mig_off,
                               double *mig_t, int *mig_p, int mig_room, int *root, double *roottime) {
  if (!h || !h->eng.finalized) return fail(IMA2P_E_ARG, "get_genealogy: not finalized");
  Engine &e = h->eng;
  if (ci < 0 || ci >= e.d.nchains || li < 0 || li >= e.d.nloci) return fail(IMA2P_E_ARG, "get_genealogy: index out of range");
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, nullptr);
  const DevLocus &L = e.loci[li].d;
  const size_t p = (size_t)ci * e.d.nloci + li, NL = e.d.NL, CAP = e.d.CAP;
  unsigned char cur = 0;
  if (!d2h(&cur, e.v.cur + p, 1, s) || !dev_sync(s)) return fail(IMA2P_E_CUDA, "download failed");
  const PairBuf &B = e.v.buf[which ? (cur ^ 1) : cur];
  std::vector<short4_t> topo(L.nl); std::vector<ushort2_t> ms(L.nl); std::vector<double> mt(CAP); std::vector<short> mp(CAP);
  int si[2]; double sd[4];
  bool ok = d2h(topo.data(), B.topo + p * NL, L.nl * sizeof(short4_t), s) && d2h(time, B.time + p * NL, L.nl * sizeof(double), s) &&
            d2h(ms.data(), B.mseg + p * NL, L.nl * sizeof(ushort2_t), s) && d2h(mt.data(), B.mig_t + p * CAP, CAP * sizeof(double), s) &&
            d2h(mp.data(), B.mig_p + p * CAP, CAP * sizeof(short), s) && d2h(si, B.si + p * 2, sizeof si, s) && d2h(sd, B.sd + p * 4, sizeof sd, s);
  if (!ok || !dev_sync(s)) return fail(IMA2P_E_CUDA, "download failed");
  if (si[1] > mig_room) return fail(IMA2P_E_ARG, "get_genealogy: mig_room too small");
  int o = 0;
  for (int i = 0; i < L.nl; i++) {
    up0[i] = topo[i].x; up1[i] = topo[i].y; down[i] = topo[i].z; pop[i] = topo[i].w;
    mig_off[i] = o;
    for (int j = 0; j < ms[i].y; j++) { mig_t[o] = mt[ms[i].x + j]; mig_p[o] = mp[ms[i].x + j]; o++; }
  }
  mig_off[L.nl] = o;
  *root = si[0]; *roottime = sd[0];
  return IMA2P_OK;
}

int ima2p_engine_get_alleles(ima2p_engine *h, int ci, int li, int which, int *A, double *dlikeA, double *pdg_a) {
  if (!h || !h->eng.finalized) return fail(IMA2P_E_ARG, "get_alleles: not finalized");
  Engine &e = h->eng;
  if (ci < 0 || ci >= e.d.nchains || li < 0 || li >= e.d.nloci || !e.d.any_sw) return fail(IMA2P_E_ARG, "get_alleles: bad index or no stepwise locus");
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, nullptr);
  const DevLocus &L = e.loci[li].d;
  const size_t p = (size_t)ci * e.d.nloci + li, NL = e.d.NL;
  unsigned char cur = 0;
  if (!d2h(&cur, e.v.cur + p, 1, s) || !dev_sync(s)) return fail(IMA2P_E_CUDA, "download failed");
  const PairBuf &B = e.v.buf[which ? (cur ^ 1) : cur];
  std::vector<short> a((size_t)kMaxLinked * NL);
  std::vector<double> dl((size_t)kMaxLinked * NL), pa(kMaxLinked);
  bool ok = d2h(a.data(), B.A + p * kMaxLinked * NL, a.size() * sizeof(short), s) && d2h(dl.data(), B.dlikeA + p * kMaxLinked * NL, dl.size() * sizeof(double), s) &&
            d2h(pa.data(), B.pdg_a + p * kMaxLinked, kMaxLinked * sizeof(double), s);
  if (!ok || !dev_sync(s)) return fail(IMA2P_E_CUDA, "download failed");
  for (int ai = 0; ai < L.nlinked; ai++) {
    for (int i = 0; i < L.nl; i++) { if (A) A[(size_t)ai * L.nl + i] = a[(size_t)ai * NL + i]; if (dlikeA) dlikeA[(size_t)ai * L.nl + i] = dl[(size_t)ai * NL + i]; }
    if (pdg_a) pdg_a[ai] = pa[ai];
  }
  return IMA2P_OK;
}

static int ensure_steppable(Engine &e) {
  if (!e.finalized) return fail(IMA2P_E_ARG, "engine not finalized");
  return IMA2P_OK;
}

int ima2p_engine_run(ima2p_engine *h, int nsteps, int swaptries, void *cuda_stream) {
  if (!h || nsteps < 0 || swaptries < 0) return fail(IMA2P_E_ARG, "run: bad argument");
  Engine &e = h->eng;
  int rc = ensure_steppable(e);
  if (rc) return rc;
  if (e.d.nchains != e.d.nchains_global) return fail(IMA2P_E_ARG, "run: engine holds a shard; use update_genealogies + swap_replay");
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, cuda_stream);
#if IMA_CUDA
  if (!e.graph_ready || e.graph_swaptries != swaptries) {
    if (e.graph_exec) { cudaGraphExecDestroy(e.graph_exec); e.graph_exec = nullptr; }
    if (e.graph_exec_deep) { cudaGraphExecDestroy(e.graph_exec_deep); e.graph_exec_deep = nullptr; }
    if (e.graph_exec_sh) { cudaGraphExecDestroy(e.graph_exec_sh); e.graph_exec_sh = nullptr; }
    if (e.graph_exec_sh_deep) { cudaGraphExecDestroy(e.graph_exec_sh_deep); e.graph_exec_sh_deep = nullptr; }
    if (!capture_steps(e, swaptries, 1, &e.graph_exec)) return fail(IMA2P_E_CUDA, "graph capture failed");
    if (e.depth > 1 && !capture_steps(e, swaptries, e.depth, &e.graph_exec_deep)) return fail(IMA2P_E_CUDA, "graph capture failed (deep)");
    e.graph_ready = true; e.graph_swaptries = swaptries;
  }
  int i = 0;
  if (e.graph_exec_deep)
    for (; i + e.depth <= nsteps; i += e.depth)
      if (!IMA_CUDA_OK(cudaGraphLaunch(e.graph_exec_deep, s))) return fail(IMA2P_E_CUDA, "graph launch failed");
  for (; i < nsteps; i++)
    if (!IMA_CUDA_OK(cudaGraphLaunch(e.graph_exec, s))) return fail(IMA2P_E_CUDA, "graph launch failed");
#else
  for (int i = 0; i < nsteps; i++) { launch_update(&e, s); launch_swap(&e, s, e.v.swapsum, swaptries); }
#endif
  return IMA2P_OK;
}

// Same work as ima2p_engine_run, launched kernel by kernel with CUDA events recorded on the launching stream around each
// kernel of every step (no overlap between kernels); kernel_ms[IMA2P_TIMED_SLOTS] receives the summed device time over the
// nsteps of {0 proposal kernels together, 1 accept, 2 swap, 3 split-time proposals, 4 accept_t, 5 changeu, 6 k_move, 7 k_weigh,
// 8 k_propose_redo (6-8 are zero on the general path), 9-11 unused}.  bench.py's roofline numerators come from here.
int ima2p_engine_run_timed(ima2p_engine *h, int nsteps, int swaptries, void *cuda_stream, float *kernel_ms) {
  if (!h || nsteps < 0 || swaptries < 0 || !kernel_ms) return fail(IMA2P_E_ARG, "run_timed: bad argument");
  Engine &e = h->eng;
  int rc = ensure_steppable(e);
  if (rc) return rc;
  if (e.d.nchains != e.d.nchains_global) return fail(IMA2P_E_ARG, "run_timed: engine holds a shard");
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, cuda_stream);
  for (int k = 0; k < IMA2P_TIMED_SLOTS; k++) kernel_ms[k] = 0.f;
#if IMA_CUDA
  const int chunk = 128, per = 12;                       // events per step
  std::vector<cudaEvent_t> ev((size_t)chunk * per);
  std::vector<int> slot_of((size_t)chunk * per, -1);     // interval ending at event i belongs to this slot
  const bool do_t = does_split_t(&e), do_u = does_changeu(&e);
  const EngineView all = view_of(&e, 0, e.d.nchains, 0);
  for (auto &x : ev) if (!IMA_CUDA_OK(cudaEventCreate(&x))) return fail(IMA2P_E_CUDA, "event create failed");
  for (int s0 = 0; s0 < nsteps; s0 += chunk) {
    const int n = nsteps - s0 < chunk ? nsteps - s0 : chunk;
    size_t ne = 0;
    auto mark = [&](int slot) { slot_of[ne] = slot; cudaEventRecord(ev[ne++], s); };
    for (int i = 0; i < n; i++) {
      mark(-1);
      if (e.fast) {
        // the three launches of launch_propose one by one
        Engine &ee = e;
        const int npairs = all.c_n * ee.d.nloci, ppw = move_ppw(&ee, npairs);
        const int gm = (npairs + ppw * kMoveWarps - 1) / (ppw * kMoveWarps);
        const size_t sm = move_smem_bytes_per_pair(ee.d) * ppw * kMoveWarps;
        launch_move(ppw, gm, sm, s, all);
        mark(6);
        IMA_LAUNCH(k_weigh, pair_grid(&ee, all), kWarpsPerBlock, weigh_smem_bytes(ee.d) * kWarpsPerBlock, s, all);
        mark(7);
        IMA_LAUNCH(k_propose_redo, ee.redo_grid, kWarpsPerBlock, ee.pair_smem * kWarpsPerBlock, s, all);
        mark(8);
      } else {
        launch_propose(&e, s, all);
        mark(0);
      }
      launch_accept(&e, s, all);
      mark(1);
      if (do_t) { launch_split_t(&e, s, all); mark(3); launch_accept_t(&e, s, all); mark(4); }
      if (do_u) { launch_changeu(&e, s, all); mark(5); }
      launch_swap(&e, s, e.v.swapsum, swaptries);
      mark(2);
    }
    if (!IMA_CUDA_OK(cudaStreamSynchronize(s))) return fail(IMA2P_E_CUDA, "sync failed (run_timed)");
    for (size_t i = 1; i < ne; i++)
      if (slot_of[i] >= 0) { float ms = 0.f; cudaEventElapsedTime(&ms, ev[i - 1], ev[i]); kernel_ms[slot_of[i]] += ms; }
  }
  kernel_ms[0] += kernel_ms[6] + kernel_ms[7] + kernel_ms[8];
  for (auto &x : ev) cudaEventDestroy(x);
#else
  for (int i = 0; i < nsteps; i++) { launch_update(&e, s); launch_swap(&e, s, e.v.swapsum, swaptries); }
#endif
  return IMA2P_OK;
}

int ima2p_engine_update_genealogies(ima2p_engine *h, double *dev_S_local, void *cuda_stream) {
  if (!h) return fail(IMA2P_E_ARG, "null engine");
  Engine &e = h->eng;
  int rc = ensure_steppable(e);
  if (rc) return rc;
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, cuda_stream);
  launch_update(&e, s);
  if (dev_S_local) IMA_LAUNCH(k_copy_swapsum, (e.d.nchains + kWarpsPerBlock * IMA_WARP - 1) / (kWarpsPerBlock * IMA_WARP), kWarpsPerBlock, 0, s, e.v, dev_S_local, 0);
#if IMA_CUDA
  if (!IMA_CUDA_OK(cudaGetLastError())) return fail(IMA2P_E_CUDA, "kernel launch failed (update)");
#endif
  return IMA2P_OK;
}

// Split-phase form of the same step.  The proposals of step s+1 do not read the temperatures, so a caller can launch them
// while the all-gather and the swap replay of step s are still in flight on another stream:
//   propose (stream A) | wait for the swaps of the previous step | decide (A): accept sweep, split-time and scalar updates,
//   S to dev_S_local, step counter + 1 | all-gather + swap_replay_late (stream B) ...
// The swap draws stay keyed by the step they belong to, so the run is the one ima2p_engine_run makes.
int ima2p_engine_step_propose(ima2p_engine *h, void *cuda_stream) {
  if (!h) return fail(IMA2P_E_ARG, "null engine");
  Engine &e = h->eng;
  int rc = ensure_steppable(e);
  if (rc) return rc;
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, cuda_stream);
  launch_propose(&e, s, view_of(&e, 0, e.d.nchains, 0));
#if IMA_CUDA
  if (!IMA_CUDA_OK(cudaGetLastError())) return fail(IMA2P_E_CUDA, "kernel launch failed (propose)");
#endif
  return IMA2P_OK;
}

int ima2p_engine_step_decide(ima2p_engine *h, double *dev_S_local, void *cuda_stream) {
  if (!h || !dev_S_local) return fail(IMA2P_E_ARG, "step_decide: bad argument");
  Engine &e = h->eng;
  int rc = ensure_steppable(e);
  if (rc) return rc;
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, cuda_stream);
  launch_accept(&e, s, view_of(&e, 0, e.d.nchains, 0));
  launch_param_updates(&e, s, view_of(&e, 0, e.d.nchains, 0));
  IMA_LAUNCH(k_copy_swapsum, (e.d.nchains + kWarpsPerBlock * IMA_WARP - 1) / (kWarpsPerBlock * IMA_WARP), kWarpsPerBlock, 0, s, e.v, dev_S_local, 1);
#if IMA_CUDA
  if (!IMA_CUDA_OK(cudaGetLastError())) return fail(IMA2P_E_CUDA, "kernel launch failed (decide)");
#endif
  return IMA2P_OK;
}

int ima2p_engine_swap_replay_late(ima2p_engine *h, const double *dev_S_global, int swaptries, void *cuda_stream) {
  if (!h || !h->eng.finalized || !dev_S_global || swaptries < 0) return fail(IMA2P_E_ARG, "swap_replay_late: bad argument");
  Engine &e = h->eng;
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, cuda_stream);
  launch_swap(&e, s, dev_S_global, swaptries, 1);
#if IMA_CUDA
  if (!IMA_CUDA_OK(cudaGetLastError())) return fail(IMA2P_E_CUDA, "kernel launch failed (swap)");
#endif
  return IMA2P_OK;
}

int ima2p_engine_swap_replay(ima2p_engine *h, const double *dev_S_global, int swaptries, void *cuda_stream) {
  if (!h || !h->eng.finalized || !dev_S_global || swaptries < 0) return fail(IMA2P_E_ARG, "swap_replay: bad argument");
  Engine &e = h->eng;
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, cuda_stream);
  launch_swap(&e, s, dev_S_global, swaptries);
#if IMA_CUDA
  if (!IMA_CUDA_OK(cudaGetLastError())) return fail(IMA2P_E_CUDA, "kernel launch failed (swap)");
#endif
  return IMA2P_OK;
}

// ---- multi-GPU: chains sharded over ranks, swap sums exchanged through peer memory (struct Exchange, ima_model.h) ----------
// Replaces swapchains_bwprocesses' MPI messages (swapchains.cpp:192-523).  Set-up, once per run and on every rank:
//   exchange_create -> (this rank's table) -> hand it to the peers (same process: the pointer itself after
//   cudaDeviceEnablePeerAccess; other processes: ima2p_ipc_export / ima2p_ipc_import) -> exchange_attach(all tables).
// Every rank must attach before any rank steps, and all ranks step in lockstep (same nsteps, same swaptries).
int ima2p_engine_exchange_create(ima2p_engine *h, void **table, uint64_t *bytes) {
  if (!h || !h->eng.finalized || !table || !bytes) return fail(IMA2P_E_ARG, "exchange_create: bad argument");
  Engine &e = h->eng;
  if (e.d.nchains_global % e.d.nchains != 0 || e.d.chain0 % e.d.nchains != 0 || e.d.nchains_global / e.d.nchains > kMaxRanks)
    return fail(IMA2P_E_ARG, "exchange_create: chains must shard evenly over at most 16 ranks");
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  if (!e.d_xch) {
    int dm[5];
    ima2p_engine_dims(h, dm);
    e.cold_len = dm[4] + 2 + e.d.nloci;
    e.xch_bytes = (size_t)2 * e.d.nchains_global * sizeof(double) + 64 + 64 + (size_t)e.cold_len * sizeof(double);
#if IMA_CUDA
    // its own allocation (not a slice of a pool): cudaIpcGetMemHandle exports whole allocations
    void *p = nullptr;
    if (!IMA_CUDA_OK(cudaMalloc(&p, e.xch_bytes)) || !IMA_CUDA_OK(cudaMemset(p, 0, e.xch_bytes)) || !IMA_CUDA_OK(cudaDeviceSynchronize()))
      return fail(IMA2P_E_CUDA, "exchange_create: allocation failed");
    e.d_xch = (unsigned char *)p;
    e.allocs.push_back(p);
#else
    // host emulation (tests): the table lives in POSIX shared memory so that ranks can be separate processes here too
    snprintf(e.xch_shm, sizeof e.xch_shm, "/ima2p_xch_%d_%p", (int)getpid(), (void *)&e);
    const int fd = shm_open(e.xch_shm, O_CREAT | O_RDWR, 0600);
    if (fd < 0 || ftruncate(fd, (off_t)e.xch_bytes) != 0) return fail(IMA2P_E_CUDA, "exchange_create: shared memory failed");
    e.d_xch = (unsigned char *)mmap(nullptr, e.xch_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (e.d_xch == (unsigned char *)MAP_FAILED) { e.d_xch = nullptr; return fail(IMA2P_E_CUDA, "exchange_create: mmap failed"); }
    memset(e.d_xch, 0, e.xch_bytes);
#endif
    if (!e.d_xch) return fail(IMA2P_E_CUDA, "exchange_create: allocation failed");
  }
#if !IMA_CUDA
  g_emu_tables[e.d_xch] = std::make_pair(std::string(e.xch_shm), e.xch_bytes);
#endif
  *table = e.d_xch; *bytes = e.xch_bytes;
  return IMA2P_OK;
}

int ima2p_engine_exchange_attach(ima2p_engine *h, void *const *tables) {
  if (!h || !h->eng.finalized || !tables || !h->eng.d_xch) return fail(IMA2P_E_ARG, "exchange_attach: call exchange_create first");
  Engine &e = h->eng;
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  Exchange &X = e.v.xch;
  X.world = e.d.nchains_global / e.d.nchains;
  X.rank = e.d.chain0 / e.d.nchains;
  X.publisher = 0;
  for (int r = 0; r < X.world; r++) {
    unsigned char *t = r == X.rank ? e.d_xch : (unsigned char *)tables[r];
    if (!t) return fail(IMA2P_E_ARG, "exchange_attach: a peer's table is missing");
    X.peer_S[r] = (double *)t;
    X.peer_arrived[r] = (unsigned long long *)(t + (size_t)2 * e.d.nchains_global * sizeof(double));
    if (r == 0) {
      X.cold_seq0 = (unsigned long long *)(t + (size_t)2 * e.d.nchains_global * sizeof(double) + 64);
      X.cold_msg0 = (double *)(t + (size_t)2 * e.d.nchains_global * sizeof(double) + 128);
      X.cold_len = e.cold_len;
    }
  }
  stream_t s = pick_stream(&e, nullptr);
  unsigned long long st = 0;
  if (!d2h(&st, e.v.nsteps, sizeof st, s) || !dev_sync(s)) return fail(IMA2P_E_CUDA, "download failed");
  X.step0 = st;
  e.xch_attached = true;
  e.graph_ready = false;
#if IMA_CUDA
  if (e.graph_exec_sh) { cudaGraphExecDestroy(e.graph_exec_sh); e.graph_exec_sh = nullptr; }
  if (e.graph_exec_sh_deep) { cudaGraphExecDestroy(e.graph_exec_sh_deep); e.graph_exec_sh_deep = nullptr; }
#endif
  return IMA2P_OK;
}

// a device allocation of this process as 64 opaque bytes another process of the same node can open
int ima2p_ipc_export(const void *dev_ptr, unsigned char *handle64) {
  if (!dev_ptr || !handle64) return fail(IMA2P_E_ARG, "ipc_export: bad argument");
#if IMA_CUDA
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  cudaIpcMemHandle_t hd;
  if (!IMA_CUDA_OK(cudaIpcGetMemHandle(&hd, (void *)dev_ptr))) return fail(IMA2P_E_CUDA, "cudaIpcGetMemHandle failed");
  memcpy(handle64, &hd, 64);
  return IMA2P_OK;
#else
  // host emulation: the handle is the name of the shared-memory object the table lives in (exchange_create), looked up by address
  memset(handle64, 0, 64);
  for (auto &kv : g_emu_tables) if (kv.first == dev_ptr) { snprintf((char *)handle64, 64, "%s %zu", kv.second.first.c_str(), kv.second.second); return IMA2P_OK; }
  return fail(IMA2P_E_ARG, "ipc_export: not an exchange table");
#endif
}
int ima2p_ipc_import(int device, const unsigned char *handle64, void **dev_ptr) {
  if (!handle64 || !dev_ptr) return fail(IMA2P_E_ARG, "ipc_import: bad argument");
#if IMA_CUDA
  if (!IMA_CUDA_OK(cudaSetDevice(device))) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle64, 64);
  if (!IMA_CUDA_OK(cudaIpcOpenMemHandle(dev_ptr, hd, cudaIpcMemLazyEnablePeerAccess))) return fail(IMA2P_E_CUDA, "cudaIpcOpenMemHandle failed");
  return IMA2P_OK;
#else
  (void)device;
  char name[64]; size_t bytes = 0;
  if (sscanf((const char *)handle64, "%63s %zu", name, &bytes) != 2) return fail(IMA2P_E_ARG, "ipc_import: bad handle");
  const int fd = shm_open(name, O_RDWR, 0600);
  if (fd < 0) return fail(IMA2P_E_CUDA, "ipc_import: shared memory object not found");
  void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) return fail(IMA2P_E_CUDA, "ipc_import: mmap failed");
  *dev_ptr = p;
  return IMA2P_OK;
#endif
}

// The two halves of a shard's step, for callers that keep the ranks in lockstep themselves (one process driving several GPUs
// or the host emulation: every rank's update, then every rank's swap).  ima2p_engine_run_sharded issues both, as one graph.
int ima2p_engine_sharded_update(ima2p_engine *h, void *cuda_stream) {
  if (!h || !h->eng.xch_attached) return fail(IMA2P_E_ARG, "sharded_update: attach the exchange first");
  Engine &e = h->eng;
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, cuda_stream);
  EngineView v = view_of(&e, 0, e.d.nchains, 0);
  v.xch.publisher = last_kernel_of_step(&e);
  launch_update(&e, s, v);
#if IMA_CUDA
  if (!IMA_CUDA_OK(cudaGetLastError())) return fail(IMA2P_E_CUDA, "kernel launch failed (sharded update)");
#endif
  return IMA2P_OK;
}
int ima2p_engine_sharded_swap(ima2p_engine *h, int swaptries, void *cuda_stream) {
  if (!h || !h->eng.xch_attached || swaptries < 0) return fail(IMA2P_E_ARG, "sharded_swap: attach the exchange first");
  Engine &e = h->eng;
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, cuda_stream);
  launch_swap(&e, s, nullptr, swaptries);
#if IMA_CUDA
  if (!IMA_CUDA_OK(cudaGetLastError())) return fail(IMA2P_E_CUDA, "kernel launch failed (sharded swap)");
#endif
  return IMA2P_OK;
}

// The cold chain's record of a sharded job (all ranks call this at the same step boundary): on rank 0 out_msg[rowlen + 2 + nloci]
// = the .ti row (floats, as doubles), probg, P(D|G), P(D|G) of every locus -- whichever rank holds the chain at beta = 1 stored it
// into rank 0's table; on the other ranks nothing is returned.  savegsampinf ginfo.cpp:318-377 for the row.
int ima2p_engine_cold_message(ima2p_engine *h, double *out_msg, void *cuda_stream) {
  if (!h || !h->eng.xch_attached) return fail(IMA2P_E_ARG, "cold_message: attach the exchange first");
  Engine &e = h->eng;
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, cuda_stream);
  int dm[5];
  ima2p_engine_dims(h, dm);
  if (!e.d_cold) {
    e.d_cold = e.alloc<double>(e.cold_len);
#if IMA_CUDA
    if (!IMA_CUDA_OK(cudaMallocHost((void **)&e.h_cold, e.cold_len * sizeof(double)))) return fail(IMA2P_E_CUDA, "pinned allocation failed");
#else
    e.h_cold = (double *)malloc(e.cold_len * sizeof(double));
#endif
    if (!e.d_cold || !e.h_cold) return fail(IMA2P_E_CUDA, "allocation failed (cold message)");
  }
  e.cold_seq++;
  IMA_LAUNCH(k_cold_message, 1, 1, 0, s, view_of(&e, 0, e.d.nchains, 0), (const int *)e.sv.chain_of_rank, dm[4], e.cold_seq, e.d_cold);
  if (e.v.xch.rank == 0) {
    if (!out_msg) return fail(IMA2P_E_ARG, "cold_message: rank 0 needs the output buffer");
    if (!d2h(e.h_cold, e.d_cold, e.cold_len * sizeof(double), s) || !dev_sync(s)) return fail(IMA2P_E_CUDA, "download failed");
    memcpy(out_msg, e.h_cold, e.cold_len * sizeof(double));
    return check_device_error(&e, s);
  }
  return dev_sync(s) ? IMA2P_OK : fail(IMA2P_E_CUDA, "sync failed");
}

// nsteps whole steps of this rank's chains; the swap sums travel through the exchange tables inside the kernels, so there is
// nothing for the host to do between steps: the step is one CUDA graph (ima2p_engine_set_pipeline applies), replayed nsteps times
int ima2p_engine_run_sharded(ima2p_engine *h, int nsteps, int swaptries, void *cuda_stream) {
  if (!h || nsteps < 0 || swaptries < 0) return fail(IMA2P_E_ARG, "run_sharded: bad argument");
  Engine &e = h->eng;
  if (!e.xch_attached) return fail(IMA2P_E_ARG, "run_sharded: attach the exchange first");
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, cuda_stream);
#if IMA_CUDA
  if (!e.graph_exec_sh || e.graph_sh_swaptries != swaptries || !e.graph_ready) {
    if (e.graph_exec_sh) { cudaGraphExecDestroy(e.graph_exec_sh); e.graph_exec_sh = nullptr; }
    if (e.graph_exec_sh_deep) { cudaGraphExecDestroy(e.graph_exec_sh_deep); e.graph_exec_sh_deep = nullptr; }
    if (!capture_steps(e, swaptries, 1, &e.graph_exec_sh, true)) return fail(IMA2P_E_CUDA, "graph capture failed (sharded)");
    if (e.depth > 1 && !capture_steps(e, swaptries, e.depth, &e.graph_exec_sh_deep, true)) return fail(IMA2P_E_CUDA, "graph capture failed (sharded, deep)");
    e.graph_sh_swaptries = swaptries;
    // the single-rank graphs are rebuilt by ima2p_engine_run when it is next called
    if (e.graph_exec) { cudaGraphExecDestroy(e.graph_exec); e.graph_exec = nullptr; }
    if (e.graph_exec_deep) { cudaGraphExecDestroy(e.graph_exec_deep); e.graph_exec_deep = nullptr; }
    e.graph_swaptries = -1;
    e.graph_ready = true;
  }
  int i = 0;
  if (e.graph_exec_sh_deep)
    for (; i + e.depth <= nsteps; i += e.depth)
      if (!IMA_CUDA_OK(cudaGraphLaunch(e.graph_exec_sh_deep, s))) return fail(IMA2P_E_CUDA, "graph launch failed");
  for (; i < nsteps; i++)
    if (!IMA_CUDA_OK(cudaGraphLaunch(e.graph_exec_sh, s))) return fail(IMA2P_E_CUDA, "graph launch failed");
#else
  // host emulation: ranks are separate processes sharing their tables (the swap kernel polls, IMA2P_EMU_WAIT_MS)
  for (int i = 0; i < nsteps; i++) {
    EngineView v = view_of(&e, 0, e.d.nchains, 0);
    v.xch.publisher = last_kernel_of_step(&e);
    launch_update(&e, s, v);
    launch_swap(&e, s, nullptr, swaptries);
  }
#endif
  return IMA2P_OK;
}

int ima2p_engine_get_proposal(ima2p_engine *h, int ci, int li, double *out4, unsigned int *flags, int *is_current) {
  if (!h || !h->eng.finalized) return fail(IMA2P_E_ARG, "get_proposal: not finalized");
  Engine &e = h->eng;
  if (ci < 0 || ci >= e.d.nchains || li < 0 || li >= e.d.nloci) return fail(IMA2P_E_ARG, "get_proposal: index out of range");
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, nullptr);
  const size_t p = (size_t)ci * e.d.nloci + li;
  unsigned char cur = 0;
  uint32_t fl = 0;
  if (!e.v.prop_dbg) return fail(IMA2P_E_ARG, "get_proposal: call ima2p_engine_set_debug_records(e, 1) before the step");
  bool ok = d2h(out4, e.v.prop_dbg + p * 4, 4 * sizeof(double), s) && d2h(out4 + 4, e.v.prop_extra + p, sizeof(double), s) &&
            d2h(&fl, e.v.prop_flags + p, sizeof fl, s) && d2h(&cur, e.v.cur + p, 1, s);
  if (!ok || !dev_sync(s)) return fail(IMA2P_E_CUDA, "download failed");
  if (flags) *flags = fl;
  if (is_current) *is_current = cur;
  return IMA2P_OK;
}

// parity hook for the device numerics (a10): out[4*n] = uppergamma, lowergamma (scalar forms, one lane) and
// their warp-cooperative forms for every (a, x); lowergamma entries are 0 where a == 0
int ima2p_debug_gamma(int device, const int *a, const double *x, int n, double *out) {
  if (!a || !x || !out || n < 1) return fail(IMA2P_E_ARG, "debug_gamma: bad argument");
#if IMA_CUDA
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fail(IMA2P_E_CUDA, "no CUDA device: ima2p_b200 has no CPU path");
  if (!IMA_CUDA_OK(cudaSetDevice(device))) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
#else
  (void)device;
#endif
  const int nlf = 100 * 5000 + 1;
  std::vector<double> lf(nlf);
  lf[0] = 0;
  for (int i = 1; i < nlf; i++) lf[i] = lf[i - 1] + log((double)i);
  double *d_lf = (double *)dev_alloc(nlf * sizeof(double)), *d_x = (double *)dev_alloc(n * sizeof(double)), *d_out = (double *)dev_alloc((size_t)n * 4 * sizeof(double));
  int *d_a = (int *)dev_alloc(n * sizeof(int)), *d_err = (int *)dev_alloc(sizeof(int));
  bool ok = d_lf && d_x && d_out && d_a && d_err && h2d(d_lf, lf.data(), nlf * sizeof(double), 0) && h2d(d_x, x, n * sizeof(double), 0) && h2d(d_a, a, n * sizeof(int), 0);
  if (ok) {
    MathCtx mc; mc.logfact = d_lf; mc.logfact_n = nlf; mc.err = d_err;
    IMA_LAUNCH(k_debug_gamma, (n + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock, 0, 0, mc, d_a, d_x, n, d_out);
    ok = d2h(out, d_out, (size_t)n * 4 * sizeof(double), 0) && dev_sync(0);
  }
  dev_free(d_lf); dev_free(d_x); dev_free(d_out); dev_free(d_a); dev_free(d_err);
  return ok ? IMA2P_OK : fail(IMA2P_E_CUDA, "debug_gamma failed");
}

int ima2p_engine_set_speculation(ima2p_engine *h, int depth) {
  if (!h || depth < 1 || depth > kSpecMax) return fail(IMA2P_E_ARG, "set_speculation: 1..4");
  h->eng.spec = depth;
  h->eng.graph_ready = false;
  return IMA2P_OK;
}

// how ima2p_engine_run issues its steps (capture_steps): chain groups on their own streams, steps per graph, and whether a
// group's decision kernels go to a high-priority stream of their own.  The chains a run visits do not depend on it.
int ima2p_engine_set_pipeline(ima2p_engine *h, int groups, int depth, int decisions_first) {
  if (!h || groups < 1 || groups > kMaxGroups || depth < 1 || depth > 64) return fail(IMA2P_E_ARG, "set_pipeline: groups 1..16, depth 1..64");
  h->eng.groups = groups; h->eng.depth = depth; h->eng.pipe_prio = decisions_first ? 1 : 0;
  h->eng.graph_ready = false;
  return IMA2P_OK;
}

int ima2p_engine_launches_per_step(ima2p_engine *h, int swaptries) {
  if (!h || !h->eng.finalized) return 0;
  Engine &e = h->eng;
  int G = e.groups < 1 ? 1 : e.groups;
  if (G > e.d.nchains) G = e.d.nchains;
#if !IMA_CUDA
  G = 1;
#endif
  int per_group = (e.fast ? 3 : 1) + 1;                                    // proposals, accept sweep
  if (does_split_t(&e)) per_group += (e.fast ? 2 : 1) + 1;                 // split-time proposals, their decision
  if (does_changeu(&e)) per_group += 1;
  (void)swaptries;
  return G * per_group + 1;                                                // k_swap is launched every step: it also advances the step counter
}

// which proposal path ima2p_engine_run uses: fast != 0 the two kernels of ima_fastpath.h (with pairs_per_warp lanes of a
// k_move warp at work: 4, 8, 16, 32, or 0 = chosen from the number of pairs), fast == 0 the general one-warp-per-pair kernel
// for every pair.  The chain does not depend on it.
int ima2p_engine_set_proposal_path(ima2p_engine *h, int fast, int pairs_per_warp) {
  if (!h || !h->eng.finalized) return fail(IMA2P_E_ARG, "set_proposal_path: not finalized");
  Engine &e = h->eng;
  if (pairs_per_warp != 0 && pairs_per_warp != 1 && pairs_per_warp != 2 && pairs_per_warp != 4 && pairs_per_warp != 8 && pairs_per_warp != 16)
    return fail(IMA2P_E_ARG, "set_proposal_path: pairs_per_warp must be 0, 1, 2, 4, 8 or 16");
  if (fast && !e.fast_ok) return fail(IMA2P_E_UNSUPPORTED, "set_proposal_path: this data set takes the general path (stepwise loci or very large samples)");
#if IMA_CUDA
  if (fast && pairs_per_warp && move_smem_bytes_per_pair(e.d) * pairs_per_warp * kMoveWarps > 227 * 1024)
    return fail(IMA2P_E_ARG, "set_proposal_path: that many pairs per warp do not fit in shared memory");
#endif
  e.fast = fast != 0; e.ppw = pairs_per_warp;
  e.graph_ready = false;
  return IMA2P_OK;
}

// parity tests: keep the per-proposal record ima2p_engine_get_proposal reads (migration weight, slide weight, slide distance,
// edge moved); off by default -- it is 32 bytes written per pair and step that nothing else reads
int ima2p_engine_set_debug_records(ima2p_engine *h, int on) {
  if (!h || !h->eng.finalized) return fail(IMA2P_E_ARG, "set_debug_records: not finalized");
  Engine &e = h->eng;
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  if (on && !e.d_prop_dbg) e.d_prop_dbg = e.alloc<double>((size_t)e.d.P * 4);
  if (on && !e.d_prop_dbg) return fail(IMA2P_E_CUDA, "device allocation failed");
  e.v.prop_dbg = on ? e.d_prop_dbg : nullptr;
  e.graph_ready = false;
  return IMA2P_OK;
}

int ima2p_engine_sync(ima2p_engine *h) {
  if (!h) return fail(IMA2P_E_ARG, "null engine");
  Engine &e = h->eng;
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
#if IMA_CUDA
  if (!IMA_CUDA_OK(cudaDeviceSynchronize())) return fail(IMA2P_E_CUDA, "device synchronize failed");
#endif
  return e.finalized ? check_device_error(&e, pick_stream(&e, nullptr)) : IMA2P_OK;
}

int ima2p_engine_counters(ima2p_engine *h, uint64_t *out8) {
  if (!h || !h->eng.finalized || !out8) return fail(IMA2P_E_ARG, "counters: bad argument");
  Engine &e = h->eng;
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, nullptr);
  std::vector<unsigned int> acc((size_t)e.d.P * 3);
  unsigned long long steps = 0, ovf = 0, sw[2] = {0, 0};
  bool ok = d2h(acc.data(), e.v.acc, acc.size() * sizeof(unsigned int), s) && d2h(&steps, e.v.nsteps, 8, s) && d2h(&ovf, e.v.overflow, 8, s) &&
            d2h(sw, e.d_swap_counts, 16, s);
  if (!ok || !dev_sync(s)) return fail(IMA2P_E_CUDA, "download failed");
  uint64_t a = 0, t = 0, m = 0;
  for (size_t p = 0; p < (size_t)e.d.P; p++) { a += acc[p * 3]; t += acc[p * 3 + 1]; m += acc[p * 3 + 2]; }
  out8[0] = steps; out8[1] = steps * (uint64_t)e.d.P; out8[2] = a; out8[3] = t; out8[4] = m; out8[5] = sw[0]; out8[6] = sw[1]; out8[7] = ovf;
  return IMA2P_OK;
}

// ---- split-time and mutation-scalar updates (update_t_RY.cpp, update_mc_params.cpp) ------------------------------
int ima2p_engine_set_update_schedule(ima2p_engine *h, int t_updates, int u_every) {
  if (!h || !h->eng.finalized || t_updates < 0 || t_updates > 3 || u_every < 0) return fail(IMA2P_E_ARG, "set_update_schedule: bad argument");
  Engine &e = h->eng;
  if (e.t_updates != t_updates || e.u_every != u_every) e.graph_ready = false;
  e.t_updates = t_updates; e.u_every = u_every;
  e.uv.t_methods = t_updates;
  return IMA2P_OK;
}

int ima2p_engine_set_update_priors(ima2p_engine *h, const double *t_max, const double *t_min, double u_prior_max, double u_window,
                                   double kappa_window, double kappa_max) {
  if (!h || !h->eng.finalized) return fail(IMA2P_E_ARG, "set_update_priors: not finalized");
  Engine &e = h->eng;
  for (int k = 0; k < e.model.nsplit; k++) {
    if (t_max) e.uv.t_max[k] = t_max[k];
    if (t_min) e.uv.t_min[k] = t_min[k];
  }
  if (u_prior_max > 0) { e.uv.u_maxratio = 3.0 * u_prior_max; e.uv.u_win = u_window > 0 ? u_window : u_prior_max / e.d.nloci; }
  else if (u_window > 0) e.uv.u_win = u_window;
  if (kappa_window > 0) e.uv.kappa_win = kappa_window;
  if (kappa_max > 0) e.uv.kappa_max = kappa_max;
  e.graph_ready = false;
  return IMA2P_OK;
}

// parity hook: one changet_RY1 on every chain with the proposed split times given (newt[nchains]); force_accept
// -1 = draw, 0 = reject, 1 = accept; out[nchains][4] = period, proposed time, MH term, accepted
int ima2p_engine_debug_split_time(ima2p_engine *h, int method, int period, const double *newt, int force_accept, double *out) {
  if (!h || !out) return fail(IMA2P_E_ARG, "debug_split_time: bad argument");
  Engine &e = h->eng;
  int rc = ensure_steppable(e);
  if (rc) return rc;
  if (e.model.nsplit < 1 || period < 0 || period >= e.model.nsplit || method < 0 || method > 1) return fail(IMA2P_E_ARG, "debug_split_time: no such split time / update");
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, nullptr);
  const size_t C = e.d.nchains;
  double *d_newt = nullptr;
  UpdateView u = e.uv;
  if (newt) {
    d_newt = e.alloc<double>(C);
    if (!d_newt || !h2d(d_newt, newt, C * sizeof(double), s)) return fail(IMA2P_E_CUDA, "upload failed");
    u.t_forced = d_newt; u.t_forced_period = period; u.t_force_accept = force_accept; u.t_forced_method = method;
  } else u.t_methods = method ? 2 : 1;
  const EngineView all = view_of(&e, 0, e.d.nchains, 0);
  IMA_LAUNCH(k_split_t, pair_grid(&e, all), kWarpsPerBlock, e.pair_smem * kWarpsPerBlock, s, all, u);
  IMA_LAUNCH(k_accept_t, e.d.nchains, IMA_CUDA ? kTWarps : 1, accept_t_smem_bytes(e.d), s, all, u);
  if (!d2h(out, e.uv.t_out, C * 4 * sizeof(double), s) || !dev_sync(s)) return fail(IMA2P_E_CUDA, "download failed");
  return check_device_error(&e, s);
}

// parity hook: evaluates (never applies) one changeu proposal on one chain: scalars j and k, u_j *= d, u_k /= d, and
// the proposed kappas of their loci when those are HKY; out[4] = new P(D|G) of j's part, of k's part, MH term, 0
int ima2p_engine_debug_changeu(ima2p_engine *h, int chain, int j, int k, double d, double kappa_j, double kappa_k, double *out) {
  if (!h || !out) return fail(IMA2P_E_ARG, "debug_changeu: bad argument");
  Engine &e = h->eng;
  int rc = ensure_steppable(e);
  if (rc) return rc;
  const int nur = e.uv.nurates;
  if (chain < 0 || chain >= e.d.nchains || !(d > 0)) return fail(IMA2P_E_ARG, "debug_changeu: bad argument");
  if (nur > 1 && (j < 0 || j >= nur || k < 0 || k >= nur || j == k)) return fail(IMA2P_E_ARG, "debug_changeu: bad scalar index");
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, nullptr);
  UpdateView u = e.uv;
  u.u_forced = 1; u.u_chain = chain; u.u_j = j; u.u_k = k; u.u_d = d; u.u_kappa[0] = kappa_j; u.u_kappa[1] = kappa_k; u.u_every = 1;
  const int gc = (e.d.nchains + kWarpsPerBlock - 1) / kWarpsPerBlock;
  IMA_LAUNCH(k_changeu, gc, kWarpsPerBlock, changeu_smem(&e) * kWarpsPerBlock, s, view_of(&e, 0, e.d.nchains, 0), u);
  if (!d2h(out, e.uv.u_out + (size_t)chain * 4, 4 * sizeof(double), s) || !dev_sync(s)) return fail(IMA2P_E_CUDA, "download failed");
  return check_device_error(&e, s);
}

// tries / accepts of the split-time and the mutation-scalar updates since the engine was created
int ima2p_engine_update_counters(ima2p_engine *h, uint64_t *out4) {
  if (!h || !h->eng.finalized || !out4) return fail(IMA2P_E_ARG, "update_counters: bad argument");
  Engine &e = h->eng;
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, nullptr);
  unsigned long long v[4];
  if (!d2h(v, e.uv.stats, sizeof(v), s) || !dev_sync(s)) return fail(IMA2P_E_CUDA, "download failed");
  for (int i = 0; i < 4; i++) out4[i] = v[i];
  return IMA2P_OK;
}

// The reference's update-rate tables (callprintacceptancerates, ima_main_mpi.cpp:3473-3900) and its swap table
// (printchaininfo, swapchains.cpp:760-778) count the cold chain / adjacent temperatures only.
int ima2p_engine_cold_counters(ima2p_engine *h, uint64_t *genealogy, uint64_t *split, uint64_t *scalars, uint64_t *adjacent) {
  if (!h || !h->eng.finalized) return fail(IMA2P_E_ARG, "cold_counters: bad argument");
  Engine &e = h->eng;
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, nullptr);
  const int nur = e.uv.nurates, G = e.d.nchains_global, nsplit = e.model.nsplit;
  std::vector<unsigned int> g((size_t)e.d.nloci * 3);
  std::vector<unsigned long long> st(update_stats_len(nur)), adj((size_t)G * 2);
  if (!d2h(g.data(), e.v.cold_acc, g.size() * sizeof(unsigned int), s) || !d2h(st.data(), e.uv.stats, st.size() * 8, s) ||
      !d2h(adj.data(), e.sv.adj_counts, adj.size() * 8, s) || !dev_sync(s))
    return fail(IMA2P_E_CUDA, "download failed");
  if (genealogy) for (size_t i = 0; i < g.size(); i++) genealogy[i] = g[i];
  if (split) for (int k = 0; k < nsplit; k++) for (int j = 0; j < 4; j++) split[k * 4 + j] = st[kColdTStat + k * 4 + j];
  if (scalars) for (int j = 0; j < 2 * nur; j++) scalars[j] = st[kColdUStat + j];
  if (adjacent) for (int r = 0; r + 1 < G; r++) { adjacent[r * 2] = adj[r * 2]; adjacent[r * 2 + 1] = adj[r * 2 + 1]; }
  return IMA2P_OK;
}

// current split times of one chain: tvals[nsplit]
int ima2p_engine_get_split_times(ima2p_engine *h, int ci, double *tvals) {
  if (!h || !h->eng.finalized || !tvals) return fail(IMA2P_E_ARG, "get_split_times: bad argument");
  Engine &e = h->eng;
  if (ci < 0 || ci >= e.d.nchains) return fail(IMA2P_E_ARG, "get_split_times: index out of range");
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, nullptr);
  if (e.model.nsplit > 0 && (!d2h(tvals, e.v.tvals + (size_t)ci * kMaxPeriods, e.model.nsplit * sizeof(double), s) || !dev_sync(s)))
    return fail(IMA2P_E_CUDA, "download failed");
  return IMA2P_OK;
}

// every chain's split times (tvals[nchains][nsplit]) and every pair's scalars (uvals[P][IMA2P_MAX_LINKED], kappa[P]) in one call
int ima2p_engine_fetch_parameters(ima2p_engine *h, double *tvals, double *uvals, double *kappa) {
  if (!h || !h->eng.finalized) return fail(IMA2P_E_ARG, "fetch_parameters: not finalized");
  Engine &e = h->eng;
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, nullptr);
  const size_t C = e.d.nchains, P = e.d.P;
  std::vector<double> tv(C * kMaxPeriods);
  bool ok = true;
  if (tvals) ok = ok && d2h(tv.data(), e.v.tvals, tv.size() * sizeof(double), s);
  if (uvals) ok = ok && d2h(uvals, e.v.uvals, P * kMaxLinked * sizeof(double), s);
  if (kappa) ok = ok && d2h(kappa, e.v.kappa, P * sizeof(double), s);
  if (!ok || !dev_sync(s)) return fail(IMA2P_E_CUDA, "download failed");
  if (tvals) for (size_t c = 0; c < C; c++) for (int k = 0; k < e.model.nsplit; k++) tvals[c * e.model.nsplit + k] = tv[c * kMaxPeriods + k];
  return IMA2P_OK;
}

// current mutation-rate scalars and kappa of one (chain, locus): uvals[IMA2P_MAX_LINKED]
int ima2p_engine_get_scalars(ima2p_engine *h, int ci, int li, double *uvals, double *kappa) {
  if (!h || !h->eng.finalized || !uvals) return fail(IMA2P_E_ARG, "get_scalars: bad argument");
  Engine &e = h->eng;
  if (ci < 0 || ci >= e.d.nchains || li < 0 || li >= e.d.nloci) return fail(IMA2P_E_ARG, "get_scalars: bad index");
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, nullptr);
  const size_t p = (size_t)ci * e.d.nloci + li;
  double k = 0.0;
  if (!d2h(uvals, e.v.uvals + p * kMaxLinked, kMaxLinked * sizeof(double), s) || !d2h(&k, e.v.kappa + p, sizeof(double), s) || !dev_sync(s))
    return fail(IMA2P_E_CUDA, "download failed");
  if (kappa) *kappa = k;
  return IMA2P_OK;
}

// ---- thermodynamic integration (marglike.cpp) ------------------------------------------------------------------
int ima2p_engine_thermo_accumulate(ima2p_engine *h, void *cuda_stream) {
  if (!h || !h->eng.finalized) return fail(IMA2P_E_ARG, "thermo_accumulate: not finalized");
  Engine &e = h->eng;
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, cuda_stream);
  const int per = kWarpsPerBlock * IMA_WARP;
  IMA_LAUNCH(k_thermo_accumulate, (e.d.nchains + per - 1) / per, kWarpsPerBlock, 0, s, e.v, (const int *)e.sv.rank_of_chain, e.d_thermosum);
#if IMA_CUDA
  if (!IMA_CUDA_OK(cudaGetLastError())) return fail(IMA2P_E_CUDA, "kernel launch failed (thermo)");
#endif
  return IMA2P_OK;
}

// this rank's share of thermosum[] (zero where the chain at that temperature lives on another rank): sum over ranks
int ima2p_engine_thermo_sums(ima2p_engine *h, double *thermosum_global, int reset) {
  if (!h || !h->eng.finalized || !thermosum_global) return fail(IMA2P_E_ARG, "thermo_sums: bad argument");
  Engine &e = h->eng;
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, nullptr);
  const size_t G = e.d.nchains_global;
  if (!d2h(thermosum_global, e.d_thermosum, G * sizeof(double), s) || !dev_sync(s)) return fail(IMA2P_E_CUDA, "download failed");
  if (reset) {
    std::vector<double> z(G, 0.0);
    if (!h2d(e.d_thermosum, z.data(), G * sizeof(double), s) || !dev_sync(s)) return fail(IMA2P_E_CUDA, "upload failed");
  }
  return IMA2P_OK;
}

// thermomarginlikecalc (marglike.cpp:121-150): Simpson's rule over the evenly spaced betas, the hottest chain
// (beta = 0) contributing 0; host arithmetic on numchains doubles, needs no device
int ima2p_thermo_marginlike(const double *thermosum, int numchains, int k, double *out) {
  if (!thermosum || !out || numchains < 2 || k < 1) return fail(IMA2P_E_ARG, "thermo_marginlike: bad argument");
  const double width = 1.0 / (float)(numchains - 1);
  double sum = 0.0;
  for (int i = 0; i <= numchains - 1; i += 2)
    if (i != numchains - 1) sum += 4.0 * (thermosum[i] / k);
  for (int i = 1; i <= numchains - 2; i += 2) sum += 2.0 * (thermosum[i] / k);
  *out = width * sum / 3.0;
  return IMA2P_OK;
}

int ima2p_engine_get_betas(ima2p_engine *h, double *betas_global) {
  if (!h || !h->eng.finalized || !betas_global) return fail(IMA2P_E_ARG, "get_betas: bad argument");
  Engine &e = h->eng;
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, nullptr);
  const int G = e.d.nchains_global;
  std::vector<int> roc(G);
  if (!d2h(roc.data(), e.sv.rank_of_chain, G * sizeof(int), s) || !dev_sync(s)) return fail(IMA2P_E_CUDA, "download failed");
  for (int c = 0; c < G; c++) betas_global[c] = e.h_beta_table[roc[c]];
  return IMA2P_OK;
}

int ima2p_engine_cold_row(ima2p_engine *h, float *row, int *present) {
  if (!h || !h->eng.finalized || !row || !present) return fail(IMA2P_E_ARG, "cold_row: bad argument");
  Engine &e = h->eng;
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, nullptr);
  const DevModel &M = e.model;
  int cold = -1;
  if (!d2h(&cold, e.sv.chain_of_rank, sizeof(int), s) || !dev_sync(s)) return fail(IMA2P_E_CUDA, "download failed");
  const int c = cold - e.d.chain0;
  *present = (c >= 0 && c < e.d.nchains);
  if (!*present) return IMA2P_OK;
  std::vector<int> wi(e.d.NI); std::vector<double> wd(e.d.ND), q(kMaxParams), m(kMaxParams), tv(kMaxPeriods);
  double pg[2];
  bool ok = d2h(wi.data(), e.v.all_i + (size_t)c * e.d.NI, e.d.NI * sizeof(int), s) && d2h(wd.data(), e.v.all_d + (size_t)c * e.d.ND, e.d.ND * sizeof(double), s) &&
            d2h(q.data(), e.v.qint + (size_t)c * kMaxParams, kMaxParams * sizeof(double), s) && d2h(m.data(), e.v.mint + (size_t)c * kMaxParams, kMaxParams * sizeof(double), s) &&
            d2h(tv.data(), e.v.tvals + (size_t)c * kMaxPeriods, kMaxPeriods * sizeof(double), s) && d2h(pg, e.v.probg + c, 8, s) && d2h(pg + 1, e.v.pdgsum + c, 8, s);
  if (!ok || !dev_sync(s)) return fail(IMA2P_E_CUDA, "download failed");
  // savegsampinf ginfo.cpp:318-377 (sums accumulated in float, as there)
  const int nq = M.nq, nm = M.nomigration ? 0 : M.nm;
  const int fcp = nq, hccp = 2 * nq, mcp = 3 * nq, fmp = mcp + nm, qip = fmp + nm, mip = qip + nq, pdgp = mip + nm;
  for (int i = 0; i < nq; i++) {
    int cc = 0; float f = 0.f, hc = 0.f;
    for (int j = 0; j < M.q_n[i]; j++) { const int x = M.q_idx[i][j]; cc += wi[x]; f += (float)wd[x]; hc += (float)wd[M.ncc + x]; }
    row[i] = (float)cc; row[fcp + i] = f; row[hccp + i] = hc; row[qip + i] = (float)q[i];
  }
  for (int i = 0; i < nm; i++) {
    int cm = 0; float f = 0.f;
    for (int j = 0; j < M.m_n[i]; j++) { const int x = M.m_idx[i][j]; cm += wi[M.ncc + x]; f += (float)wd[2 * M.ncc + x]; }
    row[mcp + i] = (float)cm; row[fmp + i] = f; row[mip + i] = (float)m[i];
  }
  row[pdgp] = (float)pg[1]; row[pdgp + 1] = (float)pg[0];
  for (int i = 0; i < M.nsplit; i++) row[pdgp + 2 + i] = (float)tv[i];
  return IMA2P_OK;
}

int ima2p_engine_state_bytes(ima2p_engine *h, uint64_t *out8) {
  if (!h || !h->eng.finalized || !out8) return fail(IMA2P_E_ARG, "state_bytes: bad argument");
  const EngineDims &d = h->eng.d;
  const uint64_t P = d.P, NL = d.NL, CAP = d.CAP;
  out8[0] = P * NL * sizeof(short4_t); out8[1] = P * NL * sizeof(double); out8[2] = P * NL * sizeof(ushort2_t);
  out8[3] = P * CAP * sizeof(double); out8[4] = P * CAP * sizeof(short); out8[5] = P * 2 * sizeof(int);
  out8[6] = P * 4 * sizeof(double); out8[7] = P * kMaxLinked * sizeof(double);
  return IMA2P_OK;
}

int ima2p_engine_put_state(ima2p_engine *h, const void *topo, const void *time, const void *mseg, const void *mig_t, const void *mig_p,
                           const void *scal_i, const void *scal_d, const void *uvals, const double *tvals, void *cuda_stream) {
  if (!h || !h->eng.finalized) return fail(IMA2P_E_ARG, "put_state: not finalized");
  Engine &e = h->eng;
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, cuda_stream);
  uint64_t b[8];
  ima2p_engine_state_bytes(h, b);
  PairBuf &B = e.v.buf[0];
  bool ok = h2d(B.topo, topo, b[0], s) && h2d(B.time, time, b[1], s) && h2d(B.mseg, mseg, b[2], s) &&
            h2d(B.si, scal_i, b[5], s) && h2d(B.sd, scal_d, b[6], s) && h2d(e.v.uvals, uvals, b[7], s);
  // the migration pools are [P][CAP] with only the first mignum entries of a row in use: copy that many columns
  {
    const int *si = (const int *)scal_i;
    int maxmig = 0;
    for (size_t p = 0; p < (size_t)e.d.P; p++) if (si[2 * p + 1] > maxmig) maxmig = si[2 * p + 1];
    if (maxmig > e.d.CAP) return fail(IMA2P_E_CAPACITY, "put_state: more migration events than mig_capacity");
#if IMA_CUDA
    if (maxmig > 0)
      ok = ok && IMA_CUDA_OK(cudaMemcpy2DAsync(B.mig_t, (size_t)e.d.CAP * 8, mig_t, (size_t)e.d.CAP * 8, (size_t)maxmig * 8, e.d.P, cudaMemcpyHostToDevice, s)) &&
           IMA_CUDA_OK(cudaMemcpy2DAsync(B.mig_p, (size_t)e.d.CAP * 2, mig_p, (size_t)e.d.CAP * 2, (size_t)maxmig * 2, e.d.P, cudaMemcpyHostToDevice, s));
#else
    ok = ok && h2d(B.mig_t, mig_t, b[3], s) && h2d(B.mig_p, mig_p, b[4], s);
#endif
  }
  if (tvals) {
    for (int c = 0; c < e.d.nchains; c++) for (int k = 0; k < kMaxPeriods; k++) e.h_tvals[(size_t)c * kMaxPeriods + k] = k < e.model.nsplit ? tvals[(size_t)c * e.model.nsplit + k] : kTimeMax;
    ok = ok && h2d(e.v.tvals, e.h_tvals.data(), e.h_tvals.size() * sizeof(double), s);
  }
#if IMA_CUDA
  ok = ok && IMA_CUDA_OK(cudaMemsetAsync(e.v.cur, 0, e.d.P, s));
#else
  memset(e.v.cur, 0, e.d.P);
#endif
  if (!ok) return fail(IMA2P_E_CUDA, "put_state failed");
  return launch_eval(&e, s);
}

// The same state in its narrow wire form: 13 instead of 20 bytes per edge cross PCIe (int8 links and population, uint8
// migration counts, the pools in edge order so that the segment starts are a prefix sum); widened on the device.
int ima2p_engine_put_state_packed(ima2p_engine *h, const void *topo8, const void *time, const void *mcount, const void *mig_t, const void *mig_p,
                                  const void *scal_i, const void *scal_d, const void *uvals, const double *tvals, void *cuda_stream) {
  if (!h || !h->eng.finalized) return fail(IMA2P_E_ARG, "put_state_packed: not finalized");
  Engine &e = h->eng;
  if (e.d.NL > 127 || e.d.CAP > 255 || e.model.ntreepops > 127) return fail(IMA2P_E_ARG, "put_state_packed: the sample does not fit the 8-bit wire form; use ima2p_engine_put_state");
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, cuda_stream);
  const size_t P = e.d.P, NL = e.d.NL;
  if (!e.d_topo8) { e.d_topo8 = e.alloc<signed char>(P * NL * 4); e.d_mcount = e.alloc<unsigned char>(P * NL); }
  if (!e.d_topo8 || !e.d_mcount) return fail(IMA2P_E_CUDA, "device allocation failed");
  uint64_t b[8];
  ima2p_engine_state_bytes(h, b);
  PairBuf &B = e.v.buf[0];
  bool ok = h2d(e.d_topo8, topo8, P * NL * 4, s) && h2d(B.time, time, b[1], s) && h2d(e.d_mcount, mcount, P * NL, s) &&
            h2d(B.si, scal_i, b[5], s) && h2d(B.sd, scal_d, b[6], s) && h2d(e.v.uvals, uvals, b[7], s);
  {
    const int *si = (const int *)scal_i;
    int maxmig = 0;
    for (size_t p = 0; p < P; p++) if (si[2 * p + 1] > maxmig) maxmig = si[2 * p + 1];
    if (maxmig > e.d.CAP) return fail(IMA2P_E_CAPACITY, "put_state_packed: more migration events than mig_capacity");
#if IMA_CUDA
    if (maxmig > 0)
      ok = ok && IMA_CUDA_OK(cudaMemcpy2DAsync(B.mig_t, (size_t)e.d.CAP * 8, mig_t, (size_t)e.d.CAP * 8, (size_t)maxmig * 8, e.d.P, cudaMemcpyHostToDevice, s)) &&
           IMA_CUDA_OK(cudaMemcpy2DAsync(B.mig_p, (size_t)e.d.CAP * 2, mig_p, (size_t)e.d.CAP * 2, (size_t)maxmig * 2, e.d.P, cudaMemcpyHostToDevice, s));
#else
    ok = ok && h2d(B.mig_t, mig_t, b[3], s) && h2d(B.mig_p, mig_p, b[4], s);
#endif
  }
  if (tvals) {
    for (int c = 0; c < e.d.nchains; c++) for (int k = 0; k < kMaxPeriods; k++) e.h_tvals[(size_t)c * kMaxPeriods + k] = k < e.model.nsplit ? tvals[(size_t)c * e.model.nsplit + k] : kTimeMax;
    ok = ok && h2d(e.v.tvals, e.h_tvals.data(), e.h_tvals.size() * sizeof(double), s);
  }
#if IMA_CUDA
  ok = ok && IMA_CUDA_OK(cudaMemsetAsync(e.v.cur, 0, e.d.P, s));
#else
  memset(e.v.cur, 0, e.d.P);
#endif
  if (!ok) return fail(IMA2P_E_CUDA, "put_state_packed failed");
  IMA_LAUNCH(k_unpack_state, (e.d.P + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock, 0, s, e.v, e.d_topo8, e.d_mcount);
  return launch_eval(&e, s);
}

// One host block, one transfer: sections as state_block_layout (ima_kernels.h), offsets from ima2p_engine_state_block_layout.
int ima2p_engine_state_block_layout(ima2p_engine *h, long long total_events, uint64_t *out10) {
  if (!h || !h->eng.finalized || !out10 || total_events < 0) return fail(IMA2P_E_ARG, "state_block_layout: bad argument");
  const StateBlock b = state_block_layout(h->eng.d, h->eng.model.nsplit, total_events);
  const size_t v[10] = {b.time, b.sd, b.uvals, b.tvals, b.mig_t, b.si, b.mig_p, b.topo8, b.mcount, b.total};
  for (int i = 0; i < 10; i++) out10[i] = v[i];
  return IMA2P_OK;
}

// The one-block upload in two halves, so that a caller that steps a stream of uploaded states can keep the copy engine and the SMs
// busy at the same time: upload_block starts the transfer of a block into one of two staging slots on `copy_stream` and returns;
// adopt_block makes `cuda_stream` wait for the oldest transfer, widens the block into the resident state and re-evaluates it.
int ima2p_engine_upload_block(ima2p_engine *h, const void *block, long long total_events, void *copy_stream) {
  if (!h || !h->eng.finalized || !block || total_events < 0) return fail(IMA2P_E_ARG, "upload_block: bad argument");
  Engine &e = h->eng;
  if (e.d.NL > 127 || e.d.CAP > 255 || e.model.ntreepops > 127) return fail(IMA2P_E_ARG, "upload_block: the sample does not fit the 8-bit wire form; use ima2p_engine_put_state");
  if (total_events > (long long)e.d.P * e.d.CAP) return fail(IMA2P_E_CAPACITY, "upload_block: more migration events than mig_capacity");
  if (e.block_pending >= 2) return fail(IMA2P_E_ARG, "upload_block: both staging slots hold a block that has not been adopted");
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, copy_stream);
  const StateBlock L = state_block_layout(e.d, e.model.nsplit, total_events);
  if (L.total > e.block_cap || !e.d_block[0] || !e.d_block[1]) {
    if (e.block_pending) return fail(IMA2P_E_ARG, "upload_block: staging must grow while a block is pending");
    const size_t cap = state_block_layout(e.d, e.model.nsplit, (long long)e.d.P * e.d.CAP).total;
    e.d_block[0] = e.alloc<unsigned char>(cap);
    e.d_block[1] = e.alloc<unsigned char>(cap);
    if (!e.d_block_moff) e.d_block_moff = e.alloc<int>(e.d.P);
    e.block_cap = (e.d_block[0] && e.d_block[1]) ? cap : 0;
  }
  if (!e.d_block[0] || !e.d_block[1] || !e.d_block_moff) return fail(IMA2P_E_CUDA, "device allocation failed");
  const int k = e.block_next_up;
#if IMA_CUDA
  if (!e.block_up_ev[k]) {
    cudaEventCreateWithFlags(&e.block_up_ev[k], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&e.block_done_ev[k], cudaEventDisableTiming);
  }
  if (e.block_done_set[k]) cudaStreamWaitEvent(s, e.block_done_ev[k], 0);       // the slot's previous block has been widened
#endif
  if (!h2d(e.d_block[k], block, L.total, s)) return fail(IMA2P_E_CUDA, "upload_block failed");
#if IMA_CUDA
  if (!IMA_CUDA_OK(cudaEventRecord(e.block_up_ev[k], s))) return fail(IMA2P_E_CUDA, "upload_block failed");
#endif
  e.block_events[k] = total_events;
  e.block_next_up = k ^ 1;
  e.block_pending++;
  return IMA2P_OK;
}

int ima2p_engine_adopt_block(ima2p_engine *h, void *cuda_stream) {
  if (!h || !h->eng.finalized) return fail(IMA2P_E_ARG, "adopt_block: not finalized");
  Engine &e = h->eng;
  if (e.block_pending < 1) return fail(IMA2P_E_ARG, "adopt_block: no uploaded block is waiting");
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, cuda_stream);
  const int k = e.block_next_adopt;
  const StateBlock L = state_block_layout(e.d, e.model.nsplit, e.block_events[k]);
#if IMA_CUDA
  cudaStreamWaitEvent(s, e.block_up_ev[k], 0);
#endif
  IMA_LAUNCH(k_block_offsets, 1, kWarpsPerBlock, kWarpsPerBlock * sizeof(int), s, e.v, e.d_block[k], L, e.d_block_moff);
  IMA_LAUNCH(k_unpack_block, (e.d.P + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock, 0, s, e.v, e.d_block[k], L, e.d_block_moff, e.model.nsplit);
#if IMA_CUDA
  cudaEventRecord(e.block_done_ev[k], s);
  e.block_done_set[k] = true;
#endif
  e.block_next_adopt = k ^ 1;
  e.block_pending--;
  return launch_eval(&e, s);
}

int ima2p_engine_put_state_block(ima2p_engine *h, const void *block, long long total_events, void *cuda_stream) {
  if (h && h->eng.block_pending) return fail(IMA2P_E_ARG, "put_state_block: an uploaded block is waiting to be adopted");
  const int rc = ima2p_engine_upload_block(h, block, total_events, cuda_stream);
  return rc != IMA2P_OK ? rc : ima2p_engine_adopt_block(h, cuda_stream);
}

// gathers the CURRENT buffer of every pair (pairs flip independently) into host memory
int ima2p_engine_fetch_state(ima2p_engine *h, void *topo, void *time, void *mseg, void *mig_t, void *mig_p, void *scal_i, void *scal_d,
                             void *cuda_stream) {
  if (!h || !h->eng.finalized) return fail(IMA2P_E_ARG, "fetch_state: not finalized");
  Engine &e = h->eng;
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, cuda_stream);
  const size_t P = e.d.P, NL = e.d.NL, CAP = e.d.CAP;
  std::vector<unsigned char> cur(P);
  if (!d2h(cur.data(), e.v.cur, P, s) || !dev_sync(s)) return fail(IMA2P_E_CUDA, "download failed");
  bool ok = true;
  // runs of pairs living in the same buffer are copied together
  for (size_t p0 = 0; p0 < P && ok;) {
    size_t p1 = p0 + 1;
    while (p1 < P && cur[p1] == cur[p0]) p1++;
    const PairBuf &B = e.v.buf[cur[p0]];
    const size_t n = p1 - p0;
    ok = d2h((short4_t *)topo + p0 * NL, B.topo + p0 * NL, n * NL * sizeof(short4_t), s) && d2h((double *)time + p0 * NL, B.time + p0 * NL, n * NL * sizeof(double), s) &&
         d2h((ushort2_t *)mseg + p0 * NL, B.mseg + p0 * NL, n * NL * sizeof(ushort2_t), s) && d2h((double *)mig_t + p0 * CAP, B.mig_t + p0 * CAP, n * CAP * sizeof(double), s) &&
         d2h((short *)mig_p + p0 * CAP, B.mig_p + p0 * CAP, n * CAP * sizeof(short), s) && d2h((int *)scal_i + p0 * 2, B.si + p0 * 2, n * 2 * sizeof(int), s) &&
         d2h((double *)scal_d + p0 * 4, B.sd + p0 * 4, n * 4 * sizeof(double), s);
    p0 = p1;
  }
  if (!ok || !dev_sync(s)) return fail(IMA2P_E_CUDA, "download failed");
  return check_device_error(&e, s);
}

// per-pair summaries of the CURRENT genealogies in one call: sd[P][4] = roottime, length, tlength, pdg;
// si[P][2] = root, mignum; wi[P][NI] = cc | mc counts (any pointer may be NULL)
int ima2p_engine_fetch_pair_summaries(ima2p_engine *h, double *sd, int *si, int *wi, void *cuda_stream) {
  if (!h || !h->eng.finalized) return fail(IMA2P_E_ARG, "fetch_pair_summaries: not finalized");
  Engine &e = h->eng;
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, cuda_stream);
  const size_t P = e.d.P, NI = e.d.NI;
  std::vector<unsigned char> cur(P);
  if (!d2h(cur.data(), e.v.cur, P, s) || !dev_sync(s)) return fail(IMA2P_E_CUDA, "download failed");
  bool ok = true;
  for (size_t p0 = 0; p0 < P && ok;) {
    size_t p1 = p0 + 1;
    while (p1 < P && cur[p1] == cur[p0]) p1++;
    const PairBuf &B = e.v.buf[cur[p0]];
    const size_t n = p1 - p0;
    if (sd) ok = ok && d2h(sd + p0 * 4, B.sd + p0 * 4, n * 4 * sizeof(double), s);
    if (si) ok = ok && d2h(si + p0 * 2, B.si + p0 * 2, n * 2 * sizeof(int), s);
    if (wi) ok = ok && d2h(wi + p0 * NI, B.gwi + p0 * NI, n * NI * sizeof(int), s);
    p0 = p1;
  }
  if (!ok || !dev_sync(s)) return fail(IMA2P_E_CUDA, "download failed");
  return IMA2P_OK;
}

// P(D|G) of every locus of one chain (C[ci]->G[li].pdg): three small copies and one synchronisation whatever the number of
// chains -- what checkhighs (output.cpp:207-240) looks at for the cold chain at a recorded step
int ima2p_engine_fetch_chain_pdg(ima2p_engine *h, int ci, double *pdg) {
  if (!h || !h->eng.finalized || !pdg || ci < 0 || ci >= h->eng.d.nchains) return fail(IMA2P_E_ARG, "fetch_chain_pdg: bad argument");
  Engine &e = h->eng;
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, nullptr);
  const size_t L = e.d.nloci, p0 = (size_t)ci * L;
  std::vector<unsigned char> cur(L);
  std::vector<double> sd0(L * 4), sd1(L * 4);
  if (!d2h(cur.data(), e.v.cur + p0, L, s) || !d2h(sd0.data(), e.v.buf[0].sd + p0 * 4, L * 4 * sizeof(double), s) ||
      !d2h(sd1.data(), e.v.buf[1].sd + p0 * 4, L * 4 * sizeof(double), s) || !dev_sync(s))
    return fail(IMA2P_E_CUDA, "download failed");
  for (size_t li = 0; li < L; li++) pdg[li] = (cur[li] ? sd1 : sd0)[li * 4 + 3];
  return IMA2P_OK;
}

// The per-step read-back in one kernel, one copy and one synchronisation: chain4[nchains][4] = beta, probg, pdg, S;
// row[rowlen] = the cold chain's .ti row when it lives on this rank (*present), as ima2p_engine_cold_row
int ima2p_engine_step_report(ima2p_engine *h, double *chain4, float *row, int *present, void *cuda_stream) {
  if (!h || !h->eng.finalized || !chain4 || !row || !present) return fail(IMA2P_E_ARG, "step_report: bad argument");
  Engine &e = h->eng;
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, cuda_stream);
  int dims[5];
  ima2p_engine_dims(h, dims);
  const int rowlen = dims[4], C = e.d.nchains;
  const size_t n = 4 * (size_t)C + rowlen + 2;
  if (!e.d_report) {
    e.d_report = e.alloc<double>(n);
#if IMA_CUDA
    if (!IMA_CUDA_OK(cudaMallocHost((void **)&e.h_report, n * sizeof(double)))) return fail(IMA2P_E_CUDA, "pinned allocation failed");
#else
    e.h_report = (double *)malloc(n * sizeof(double));
#endif
    if (!e.d_report || !e.h_report) return fail(IMA2P_E_CUDA, "allocation failed (step report)");
  }
  IMA_LAUNCH(k_pack_report, 1, 1, 0, s, e.v, (const int *)e.sv.chain_of_rank, rowlen, e.d_report);
  if (!d2h(e.h_report, e.d_report, n * sizeof(double), s) || !dev_sync(s)) return fail(IMA2P_E_CUDA, "download failed");
  memcpy(chain4, e.h_report, 4 * (size_t)C * sizeof(double));
  const double *r = e.h_report + 4 * (size_t)C;
  *present = r[rowlen] != 0.0;
  if (*present) for (int i = 0; i < rowlen; i++) row[i] = (float)r[i];
  if (r[rowlen + 1] != 0.0) {
    char buf[120];
    snprintf(buf, sizeof buf, "device error word %d raised by a kernel", (int)r[rowlen + 1]);
    return fail(IMA2P_E_DEVICE, buf);
  }
  return IMA2P_OK;
}

// The same report in two halves, so that the host can queue step s+1 before it waits for the results of step s: _begin packs
// and starts the copy into slot (0 or 1) and returns at once; _end waits for that slot's copy only and hands the results out.
// The device never idles between steps while the host wakes up and reads: the two slots alternate.
int ima2p_engine_step_report_begin(ima2p_engine *h, int slot, void *cuda_stream) {
  if (!h || !h->eng.finalized || slot < 0 || slot > 1) return fail(IMA2P_E_ARG, "step_report_begin: bad argument");
  Engine &e = h->eng;
  if (e.report_pending[slot]) return fail(IMA2P_E_ARG, "step_report_begin: the slot holds a report that has not been read");
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, cuda_stream);
  int dims[5];
  ima2p_engine_dims(h, dims);
  const int rowlen = dims[4], C = e.d.nchains;
  const size_t n = 4 * (size_t)C + rowlen + 2;
  if (!e.d_report2[slot]) {
    e.d_report2[slot] = e.alloc<double>(n);
#if IMA_CUDA
    if (!IMA_CUDA_OK(cudaMallocHost((void **)&e.h_report2[slot], n * sizeof(double)))) return fail(IMA2P_E_CUDA, "pinned allocation failed");
    if (!IMA_CUDA_OK(cudaEventCreateWithFlags(&e.report_ev[slot], cudaEventDisableTiming))) return fail(IMA2P_E_CUDA, "event creation failed");
#else
    e.h_report2[slot] = (double *)malloc(n * sizeof(double));
#endif
    if (!e.d_report2[slot] || !e.h_report2[slot]) return fail(IMA2P_E_CUDA, "allocation failed (step report)");
  }
  IMA_LAUNCH(k_pack_report, 1, 1, 0, s, e.v, (const int *)e.sv.chain_of_rank, rowlen, e.d_report2[slot]);
  if (!d2h(e.h_report2[slot], e.d_report2[slot], n * sizeof(double), s)) return fail(IMA2P_E_CUDA, "download failed");
#if IMA_CUDA
  if (!IMA_CUDA_OK(cudaEventRecord(e.report_ev[slot], s))) return fail(IMA2P_E_CUDA, "event record failed");
#endif
  e.report_pending[slot] = true;
  return IMA2P_OK;
}

int ima2p_engine_step_report_end(ima2p_engine *h, int slot, double *chain4, float *row, int *present) {
  if (!h || !h->eng.finalized || slot < 0 || slot > 1 || !chain4 || !row || !present) return fail(IMA2P_E_ARG, "step_report_end: bad argument");
  Engine &e = h->eng;
  if (!e.report_pending[slot]) return fail(IMA2P_E_ARG, "step_report_end: no report was begun in this slot");
#if IMA_CUDA
  if (!use_device(&e) || !IMA_CUDA_OK(cudaEventSynchronize(e.report_ev[slot]))) return fail(IMA2P_E_CUDA, "waiting for the report failed");
#endif
  e.report_pending[slot] = false;
  int dims[5];
  ima2p_engine_dims(h, dims);
  const int rowlen = dims[4], C = e.d.nchains;
  memcpy(chain4, e.h_report2[slot], 4 * (size_t)C * sizeof(double));
  const double *r = e.h_report2[slot] + 4 * (size_t)C;
  *present = r[rowlen] != 0.0;
  if (*present) for (int i = 0; i < rowlen; i++) row[i] = (float)r[i];
  if (r[rowlen + 1] != 0.0) {
    char buf[120];
    snprintf(buf, sizeof buf, "device error word %d raised by a kernel", (int)r[rowlen + 1]);
    return fail(IMA2P_E_DEVICE, buf);
  }
  return IMA2P_OK;
}

int ima2p_engine_fetch_chain_summary(ima2p_engine *h, double *out4, void *cuda_stream) {
  if (!h || !h->eng.finalized || !out4) return fail(IMA2P_E_ARG, "fetch_chain_summary: bad argument");
  Engine &e = h->eng;
  if (!use_device(&e)) return fail(IMA2P_E_CUDA, "cudaSetDevice failed");
  stream_t s = pick_stream(&e, cuda_stream);
  const int C = e.d.nchains;
  std::vector<double> b(C), g(C), p(C), w(C);
  bool ok = d2h(b.data(), e.v.beta, C * 8, s) && d2h(g.data(), e.v.probg, C * 8, s) && d2h(p.data(), e.v.pdgsum, C * 8, s) && d2h(w.data(), e.v.swapsum, C * 8, s);
  if (!ok || !dev_sync(s)) return fail(IMA2P_E_CUDA, "download failed");
  for (int c = 0; c < C; c++) { out4[c * 4] = b[c]; out4[c * 4 + 1] = g[c]; out4[c * 4 + 2] = p[c]; out4[c * 4 + 3] = w[c]; }
  return check_device_error(&e, s);
}

}  // extern "C"

#include "ima_mcf.h"
