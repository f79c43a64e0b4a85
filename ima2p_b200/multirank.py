"""One process per GPU: chains shard by rank, loci stay together (they are coupled through the prior,
SURVEY.md fact 1).  The only exchange of an M-mode step is an all-gather of one double per chain --
S = sum_li pdg + probg (swapweight, swapchains.cpp:12-34) -- after which every rank replays the same swap
attempts with the same counter-based random stream and ends with the same beta permutation
(temperatures move, chains do not: swapbetasonly, ima_main_mpi.cpp:1958-1962).  torch.distributed is the
plumbing (NCCL over NVLink on GPUs; gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def attach_exchange(engine, device_index=0):
    """One rank per process: hand every rank's exchange table to every other rank (cudaIpc handles through
    torch.distributed's object all-gather, once) and attach them.  After this the ranks only call
    ``engine.run_sharded(nsteps)`` in lockstep: the swap sums cross NVLink inside the kernels
    (include/ima2p_b200.h, ima2p_engine_exchange_*), no collective is called per step."""
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = engine.exchange_handle()
    handles = [None] * world
    dist.all_gather_object(handles, mine)
    tables = [None if r == rank else engine.exchange_open(handles[r], device_index) for r in range(world)]
    engine.exchange_attach(tables)
    dist.barrier()                      # nobody steps before everybody is attached


class ShardedStepper:
    """Whole qupdate steps over chains sharded by rank.

    On GPUs the step runs in split phases (include/ima2p_b200.h, ima2p_engine_step_*): the all-gather and the swap replay
    of step s go to a side stream and hide behind the proposals of step s+1, which do not read the temperatures; only the
    accept sweep of step s+1 waits for them.  The result is the run `Engine.run` makes on one GPU with all the chains
    (tests/test_multirank_gloo.py), whatever the number of ranks."""

    def __init__(self, engine, device, stream=None, overlap=True, split_phase=True):
        self.eng = engine
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.S_local = torch.zeros(engine.nchains, dtype=torch.float64, device=device)
        self.S_global = torch.zeros(engine.nchains_global, dtype=torch.float64, device=device)
        self.stream = stream
        self.cuda = torch.cuda.is_available() and self.S_local.device.type == "cuda"
        self.overlap = bool(overlap) and bool(split_phase) and self.cuda
        self.split_phase = bool(split_phase)
        if self.overlap:
            self.main = torch.cuda.ExternalStream(stream) if stream else torch.cuda.current_stream()
            self.side = torch.cuda.Stream()
            self.ev_decided, self.ev_swapped = torch.cuda.Event(), torch.cuda.Event()
            self._pending = False
        assert engine.nchains * self.world == engine.nchains_global, "chains must shard evenly over ranks"

    def _gather(self):
        if self.world > 1:
            dist.all_gather_into_tensor(self.S_global, self.S_local)
        else:
            self.S_global.copy_(self.S_local)

    def step(self, swaptries):
        if self.overlap:
            return self._step_overlapped(swaptries)
        if self.split_phase:                                     # the same calls in order (CPU tests, one stream)
            self.eng.step_propose(self.stream)
            self.eng.step_decide(self.S_local.data_ptr(), self.stream)
            self._gather()
            self.eng.swap_replay_late(self.S_global.data_ptr(), swaptries, self.stream)
            return
        self.eng.update_genealogies(self.S_local.data_ptr(), self.stream)
        self._gather()
        self.eng.swap_replay(self.S_global.data_ptr(), swaptries, self.stream)

    def _step_overlapped(self, swaptries):
        eng, main, side = self.eng, self.main, self.side
        eng.step_propose(main.cuda_stream)                       # needs only the genealogies left by the previous decide
        if self._pending:
            main.wait_event(self.ev_swapped)                     # the temperatures of this step
        eng.step_decide(self.S_local.data_ptr(), main.cuda_stream)
        self.ev_decided.record(main)
        side.wait_event(self.ev_decided)
        with torch.cuda.stream(side):
            self._gather()
            eng.swap_replay_late(self.S_global.data_ptr(), swaptries, side.cuda_stream)
            self.ev_swapped.record(side)
        self._pending = True

    def finish(self):
        """The launching stream waits for the swaps still in flight (call before reading results or timing)."""
        if self.overlap and self._pending:
            self.main.wait_event(self.ev_swapped)
            self._pending = False

    def capture(self, swaptries):
        """One step (kernels + the NCCL all-gather) as a CUDA graph on the current torch stream: the step is a handful
        of short launches, so at 2-8 GPUs the launch gaps between them are what the graph removes.  Returns False and
        keeps the eager path when the capture is not possible (CPU/gloo, or a torch/NCCL that cannot capture)."""
        if not self.cuda:
            return False
        try:
            self.finish()
            self.overlap = self.split_phase = False
            for _ in range(3):                      # warm-up outside the capture (NCCL channels, lazy allocations)
                self.step(swaptries)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            cur = torch.cuda.current_stream()
            self.stream = cur.cuda_stream
            with torch.cuda.graph(g, stream=cur):
                self.step(swaptries)
            self._graph, self._graph_swaptries = g, swaptries
            return True
        except Exception:
            self._graph = None
            return False

    def run(self, nsteps, swaptries=None):
        st = self.eng.default_swaptries() if swaptries is None else swaptries
        g = getattr(self, "_graph", None)
        if g is not None and self._graph_swaptries == st:
            for _ in range(nsteps):
                g.replay()
            return
        for _ in range(nsteps):
            self.step(st)
        self.finish()


# ---- L mode: the sampled genealogies (.ti rows) shard by rank; README.md:117 of the reference: "a separate L mode run
# will run in serial" -- here every evaluation is local partial sums plus one tiny collective ----------------------

def sharded_margincalc(lm, x, yadjust, pi, logi, device="cpu"):
    """margincalc (surface_call_functions.cpp:119-173) over rows sharded across ranks: all-reduce of nx doubles."""
    import numpy as np
    sums = torch.from_numpy(lm.marginal_sums(pi, x, round_counts=1)).to(device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    s = sums.cpu().numpy() / lm.nrows_total
    if logi:
        s = np.where(s <= 0, -1e200, np.log(np.where(s <= 0, 1.0, s)))
    return s - yadjust


def sharded_jointp(lm, x, calc_ess=True, device="cpu"):
    """jointp (jointfind.cpp:885-1047) over rows sharded across ranks, <= 32 vectors per call.  Exchanges: the local
    maxima (all-gather, gives every rank the maximum of the rows before it and the global maximum), then six doubles
    per vector (sums all-reduced; the smallest kept term is the minimum over ranks)."""
    import numpy as np
    x = np.atleast_2d(x)
    nv = len(x)
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    local = torch.from_numpy(np.ascontiguousarray(lm.joint_phase1(x)).reshape(-1)).to(device)
    if world > 1:
        allmax = torch.zeros(world * nv, dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(allmax, local)
        allmax = allmax.cpu().numpy().reshape(world, nv)
    else:
        allmax = local.cpu().numpy().reshape(1, nv)
    gmax = allmax.max(axis=0)
    if rank > 0:
        lm.joint_reseed(nv, allmax[:rank].max(axis=0))
    rec = torch.from_numpy(np.ascontiguousarray(lm.joint_phase2(nv, gmax)).reshape(-1)).to(device)
    if world > 1:
        allrec = torch.zeros(world * nv * 6, dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(allrec, rec)
        allrec = allrec.cpu().numpy().reshape(world, nv, 6)
    else:
        allrec = rec.cpu().numpy().reshape(1, nv, 6)
    q, ess = np.zeros(nv), np.zeros(nv)
    for v in range(nv):
        tot = allrec[:, v, :].sum(axis=0)
        k = int(np.argmin(allrec[:, v, 4]))
        tot[4], tot[5] = allrec[k, v, 4], allrec[k, v, 5]
        q[v], ess[v] = lm.joint_finish(tot, gmax[v], calc_ess)
    return q, ess


def sharded_jointp_device(lm, x, calc_ess=True, device="cuda", batch=512):
    """sharded_jointp with nothing but device collectives between the phases: per batch of <= 512 vectors two NCCL all-gathers
    on device buffers (local maxima; records), every batch queued behind the previous one on the current stream, ONE
    device-to-host copy of all records at the end.  Any number of vectors per call."""
    import numpy as np
    x = np.atleast_2d(x)
    nv_all = len(x)
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    stream = torch.cuda.current_stream().cuda_stream
    recs = []
    for b0 in range(0, nv_all, batch):
        xb = x[b0:b0 + batch]
        nv = len(xb)
        local = torch.empty(nv, dtype=torch.float64, device=device)
        lm.joint_begin(xb, local.data_ptr(), stream)
        allmax = torch.empty(world * nv, dtype=torch.float64, device=device)
        if world > 1:
            dist.all_gather_into_tensor(allmax, local)
        else:
            allmax.copy_(local)
        rec = torch.empty(nv * 8, dtype=torch.float64, device=device)
        lm.joint_middle(nv, allmax.data_ptr(), world, rank, rec.data_ptr(), stream)
        allrec = torch.empty(world * nv * 8, dtype=torch.float64, device=device)
        if world > 1:
            dist.all_gather_into_tensor(allrec, rec)
        else:
            allrec.copy_(rec)
        recs.append((nv, allrec, local, allmax, rec))           # the tensors stay alive until the stream has used them
    q, ess = np.zeros(nv_all), np.zeros(nv_all)
    o = 0
    import ctypes as C
    for nv, allrec, _, _, _ in recs:
        a = np.ascontiguousarray(allrec.cpu().numpy())
        qb, eb = np.zeros(nv), np.zeros(nv)
        dp = lambda z: z.ctypes.data_as(C.POINTER(C.c_double))
        lm.lib.ima2p_lmode_joint_finish_gathered(dp(a), world, nv, lm.nrows_total, int(calc_ess), dp(qb), dp(eb))
        q[o:o + nv], ess[o:o + nv] = qb, eb
        o += nv
    return q, ess


def sharded_moments(lm, device="cpu"):
    """print_means_variances_correlations (output.cpp:687-745) over rows sharded across ranks: one all-reduce of the calcx row
    sums (2 np + np^2 doubles), then the reference's closing arithmetic on the totals."""
    raw = torch.from_numpy(lm.moments_raw()).to(device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(raw, op=dist.ReduceOp.SUM)
    return lm.moments_finish(raw.cpu().numpy(), lm.nrows_total)


def sharded_popmig(lm, thetai, mi, x, device="cpu"):
    """calc_popmig with prob_or_like = 0 (popmig.cpp:9-97, uniform migration prior) over rows sharded across ranks."""
    sums = torch.from_numpy(lm.popmig_sums(thetai, mi, x)).to(device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    return sums.cpu().numpy() / lm.nrows_total
