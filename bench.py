#!/usr/bin/env python
"""Benchmark of the IMa2p hot path on B200 (contract: see the task description / DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # our arm (rank 0 prints ONE JSON line)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's own CPU code

Metric (BASELINE.json): chain x locus genealogy updates/sec.  One "step" = one M-mode step of the whole job (qupdate,
ima_main_mpi.cpp:1788-2047): updategenealogy for every chain x locus, a split-time update of every chain, the mutation
scalars every 5th step, the step's MC3 swap attempts.  Workload at N = 1: BASELINE configs[1] -- Sim1_50loci-shaped loci (50
infinite-sites loci, 15+15 genes) with 128 Metropolis-coupled chains; with N GPUs every GPU holds 128 chains (weak scaling,
chains shard by rank and only the chains' swap sums cross GPUs, through peer memory inside the kernels).  With N >= 2 the line
also carries `config3`: BASELINE configs[2]'s shape, 300 loci x 256 chains per GPU (2,048 chains at N = 8), with the
reference's own figure on the same 300-locus input from all host cores of the same run.
"""
import argparse
import gzip
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")

WORKLOADS = {
    # name: (loci, genes pop0, genes pop1, chains per GPU, description, golden fixture holding the reference's own input)
    "sim50x128": (50, 15, 15, 128, "Sim1_50loci-shaped: 50 IS loci, 15+15 genes, 128 coupled chains per GPU", "state_sim50_hn3"),
    "sim300x256": (300, 15, 15, 256, "Sim1_300loci-shaped: 300 IS loci, 15+15 genes, 256 coupled chains per GPU", "state_sim300_hn1"),
    "sim5x4": (5, 10, 10, 4, "Sim1_5loci-shaped: 5 IS loci, 10+10 genes, 4 coupled chains", "state_sim5_hn4"),
}
PRIOR_Q, PRIOR_M, PRIOR_T = 10.0, 1.0, 3.0
T0 = 0.5 * PRIOR_T           # the reference starts every chain at (i+1)/(nsplit+1) of the -t prior (initialize.cpp:1959); both
                             # arms then run the whole qupdate step, split-time updates included, from there
HEAT = (1, 0.96, 0.9)        # -hfg -ha 0.96 -hb 0.9
BURN = 1000                  # whole steps before anything is timed, in BOTH arms (the migration load per genealogy settles)
L2_MB = 126.0


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def mark(self):
        """Samples from here on count (called when the timed region starts; nvidia-smi is already running)."""
        self.first = len(self.rows)

    def count(self):
        return len(self.rows) - getattr(self, "first", 0)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = self.rows[getattr(self, "first", 0):]
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            if len(r) >= 9:
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- workloads ------------------------------------------------------------------------------------------------------------
def fixture(wl):
    return json.load(gzip.open(os.path.join(ROOT, "tests", "golden", WORKLOADS[wl][5] + ".json.gz")))


def dataset(wl, data):
    """Loci in the form synth.make_dataset returns.  data == "real": the reference's own Simulations/*.u input as the golden
    fixture of the workload holds it (the fixtures were written by the reference after it read the file: same 0/1 columns in the
    same order; /root/reference does not exist on the GPU box)."""
    from ima2p_b200 import synth
    nloci, n0, n1 = WORKLOADS[wl][:3]
    if data == "synthetic":
        return synth.make_dataset(nloci, n0, n1, seed=11)
    d = fixture(wl)
    loci = []
    for li, L in enumerate(d["loci"]):
        seq = np.asarray(L["seq"], np.int32).reshape(L["numgenes"], L["numsites"])
        loci.append(dict(name="Locus%d" % (li + 1), n=L["numgenes"], samppop=list(L["samppop"]), numsites=L["numsites"], seq=seq))
    return loci


def state_from_fixture(d, nchains, NL, CAP):
    """put_state buffers for nchains chains from the genealogies of a fixture (its chains used in turn: valid genealogies of
    the reference's own run on the real input, with the split times they belong to)."""
    nloci = len(d["loci"])
    P = nchains * nloci
    topo = -np.ones((P, NL, 4), np.int16); time_ = np.zeros((P, NL)); mseg = np.zeros((P, NL, 2), np.uint16)
    mig_t = np.zeros((P, CAP)); mig_p = np.zeros((P, CAP), np.int16); si = np.zeros((P, 2), np.int32); sd = np.zeros((P, 4))
    uv = np.ones((P, 4)); tvals = np.zeros((nchains, 1))
    for c in range(nchains):
        ch = d["chains"][c % len(d["chains"])]
        tvals[c, 0] = ch["tvals"][0]
        for li, g in enumerate(ch["G"]):
            t, p = g["tree"], c * nloci + li
            nl = len(t["up0"])
            topo[p, :nl, 0], topo[p, :nl, 1], topo[p, :nl, 2], topo[p, :nl, 3] = t["up0"], t["up1"], t["down"], t["pop"]
            time_[p, :nl] = t["time"]
            o = 0
            for e, lst in enumerate(t["mig"]):
                k = len(lst) // 2
                mseg[p, e] = (o, k)
                mig_t[p, o:o + k], mig_p[p, o:o + k] = lst[0::2], lst[1::2]
                o += k
            si[p] = (t["root"], o)
            sd[p, 0] = t["roottime"]
            uv[p, 0] = g["uvals"][0]
    return dict(topo=topo, time=time_, mseg=mseg, mig_t=mig_t, mig_p=mig_p, scal_i=si, scal_d=sd, uvals=uv, tvals=tvals)


def build_engine(wl, rank, world, seed=2026, mig_capacity=64, data="synthetic"):
    from ima2p_b200 import Engine, synth
    nloci, n0, n1, cpg = WORKLOADS[wl][:4]
    loci = dataset(wl, data)
    if data == "real":
        # the product's own reader on the file, as a user's run would: write the input in the reference's .u format, read it back
        from ima2p_b200.readu import read_u
        tmp = tempfile.mkdtemp(prefix="ima2p_u_")
        synth.write_u(os.path.join(tmp, "real.u"), loci)
        rd = read_u(os.path.join(tmp, "real.u"))["loci"]
        assert len(rd) == nloci and all(np.array_equal(a["seq"], b["seq"]) for a, b in zip(rd, loci)), "the .u reader changed the data"
    model = synth.two_population_model(PRIOR_Q, PRIOR_M)
    eng = Engine(cpg, nloci, mig_capacity=mig_capacity, seed=seed, device=0 if world == 1 else int(os.environ.get("LOCAL_RANK", 0)),
                 nchains_global=cpg * world, chain0=cpg * rank)
    eng.set_model(**model)
    for li, L in enumerate(loci):
        eng.set_locus(li, 0, L["n"], L["numsites"], L["samppop"], seq=L["seq"])
    eng.finalize()
    if cpg * world >= 4:
        eng.set_heating(*HEAT)
    elif cpg * world > 1:
        eng.set_heating(0, 0.05, 0.0)
    if data == "real":
        st = state_from_fixture(fixture(wl), cpg, eng.NL, eng.CAP)
    else:
        st = synth.initial_state(loci, cpg, eng.NL, eng.CAP, t0=T0, seed=100 + rank)
    return eng, loci, st


STATE_KEYS = ["topo", "time", "mseg", "mig_t", "mig_p", "scal_i", "scal_d", "uvals"]


def algorithmic_bytes_per_update(n, mig_per_genealogy, p_acc, NI, ND):
    """B_IS = 24(2n-1) + 12 M + W_g + 24 + p_acc (72 + 12 M_e + W_g + 16)."""
    W_g = 4 * NI + 8 * ND
    M = mig_per_genealogy
    M_e = M / max(1.0, 2.0 * n - 1) * 2.0
    return 24.0 * (2 * n - 1) + 12.0 * M + W_g + 24.0 + p_acc * (72.0 + 12.0 * M_e + W_g + 16.0)


def state_mb_per_gpu(wl, cap=64):
    """Both state buffers of one GPU's chains: topo 8 + time 8 + mseg 4 bytes per edge, 10 per pool entry,
    40 + 108 of scalars and weights per pair."""
    nloci, n0, n1, cpg = WORKLOADS[wl][:4]
    nl = 2 * (n0 + n1) - 1
    return 2.0 * cpg * nloci * (20.0 * nl + 10.0 * cap + 148.0) / 1e6


def make_config(wl, world, schedule, data):
    """The same dictionary in both arms (the driver compares them)."""
    nloci, n0, n1, cpg, desc = WORKLOADS[wl][:5]
    mb = state_mb_per_gpu(wl)
    return {"workload": desc, "chains_total": cpg * max(world, 1), "loci": nloci, "genes_per_locus": n0 + n1,
            "priors": "-q %g -m %g -t %g" % (PRIOR_Q, PRIOR_M, PRIOR_T), "heating": "-hfg -ha 0.96 -hb 0.9",
            "parallelism": "chains sharded by rank x%d" % world, "input": data, "burn_in_steps": BURN,
            "l2": "no flush between steps: the resident state (%.0f MB per GPU, both buffers) is what every step re-reads by design; %s"
                  % (mb, "it fits the 126 MB L2" if mb < L2_MB else "it is larger than the 126 MB L2, every step streams it from HBM"),
            "schedule": ("qupdate: updategenealogy for every chain x locus, split-time update of every chain, mutation scalars every 5th step, swaps")
            if schedule == "full" else "updategenealogy for every chain x locus + swaps"}


# ---- the reference's own CPU code ---------------------------------------------------------------------------------------------
def run_reference_processes(ufile, total_chains, nproc, iters, chunks, burn, full, tmp, budget_s=20.0, gburn=0):
    """The reference's own updategenealogy()/qupdate() loop (oracle/_ref/ref_harness `bench` mode) in nproc
    independent serial processes, each holding total_chains/nproc chains (no MPI in this image: no cross-process
    swaps, which makes this an upper bound on the reference's MPI build)."""
    per = [total_chains // nproc + (1 if i < total_chains % nproc else 0) for i in range(nproc)]
    per = [c for c in per if c > 0]

    def launch(i, c, seed):
        out = os.path.join(tmp, "ref_%d_%d.json" % (i, seed))
        heat = ["-hfg", "-ha", "0.96", "-hb", "0.9"] if c >= 4 else (["-hfl", "-ha", "0.05"] if c > 1 else [])
        cmd = [HARNESS, "bench", out, "burn=%d" % burn, "iters=%d" % iters, "chunks=%d" % chunks, "full=%d" % full,
               "gburn=%d" % gburn, "seed=%d" % seed, "--", "-i", ufile, "-o", os.path.join(tmp, "ref_%d.out" % i), "-q", str(PRIOR_Q), "-m",
               str(PRIOR_M), "-t", str(PRIOR_T), "-b", "100", "-l", "100", "-p01", "-z", "100000000", "-hn", str(c)] + heat
        return subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=tmp), out

    # The reference itself occasionally spins forever for some RNG seeds (observed: seed 1 on this input, inside
    # its own proposal code); such a process is killed at the deadline and left out of the aggregate.
    # A process that is still running long after most of the others have finished is taken to be in that state: it is
    # killed and started again with another seed (twice at most), so that every host core contributes to the figure.
    t_start = time.time()
    deadline = t_start + max(120.0, 8.0 * budget_s)
    live = {i: launch(i, c, 1000 + 17 * i) + (0, time.time()) for i, c in enumerate(per)}
    res, durations = [], []
    while live and time.time() < deadline:
        time.sleep(0.2)
        for i in list(live):
            p, out, tries, t0 = live[i]
            if p.poll() is not None:
                del live[i]
                try:
                    res.append(json.load(open(out)))
                    durations.append(time.time() - t0)
                except (ValueError, OSError):
                    pass
            elif len(durations) >= max(1, len(per) // 2) and time.time() - t0 > 3.0 * float(np.median(durations)) + 5.0 and tries < 2:
                p.kill()
                live[i] = launch(i, per[i], 1000 + 17 * i + 7919 * (tries + 1)) + (tries + 1, time.time())
    for p, _, _, _ in live.values():
        p.kill()
    return res, len(res)


def reference_throughput(wl, world, steps, warmup, budget_s, full=1, data="synthetic", burn=BURN):
    """(value updates/s over all host cores, ms per step, cores, sample description); each step is one chunk."""
    from ima2p_b200 import synth
    if not os.path.exists(HARNESS):
        return None
    nloci, n0, n1, cpg = WORKLOADS[wl][:4]
    total_chains = cpg * world
    ncores = os.cpu_count() or 1
    nproc = max(1, min(ncores, total_chains))
    tmp = tempfile.mkdtemp(prefix="ima2p_ref_")
    ufile = os.path.join(tmp, "input.u")
    synth.write_u(ufile, dataset(wl, data))
    chains_pp = -(-total_chains // nproc)
    est = 30000.0 if not full else 17000.0                 # updates/s/core, survey probe
    # the untimed burn-in is bounded too: what a core does in about 20 s (whole steps, as in our arm)
    burn = int(max(5, min(burn, 20.0 * est / (chains_pp * nloci))))
    chunks = steps + warmup
    iters = max(1, int(budget_s * est / (chunks * chains_pp * nloci)))
    res, used = run_reference_processes(ufile, total_chains, nproc, iters, chunks, burn=burn, full=full, tmp=tmp, budget_s=budget_s + 20.0)
    if not res:
        return None
    chunk_s = np.array([r["chunk_seconds"] for r in res])            # [proc][chunk]
    upd = sum(r["updates_per_chunk"] for r in res)
    timed = chunk_s[:, warmup:]
    step_s = timed.max(axis=0).mean()
    sample = "%d serial reference processes x %d chains, %d loci (%s input); %d %s per step after %d whole steps of burn-in" % (
        used, chains_pp, nloci, data, iters, "whole qupdate() steps" if full else "updategenealogy sweeps", burn)
    return dict(value=upd / step_s, ms_per_step=step_s * 1e3, cores=used, sample=sample,
                accept=sum(r["accepted"] for r in res) / max(1, sum(r.get("accept_base", r["updates"]) for r in res)))


# ---- our arm ----------------------------------------------------------------------------------------------------------------
class Job:
    """One workload on this rank's GPU: engine, streams, stepping (one GPU: Engine.run; several: Engine.run_sharded after the
    ranks attached each other's exchange tables)."""

    def __init__(self, wl, args, rank, world, dev):
        import torch
        self.torch, self.wl, self.rank, self.world, self.dev = torch, wl, rank, world, dev
        self.nloci, self.n0, self.n1, self.cpg = WORKLOADS[wl][:4]
        self.eng, self.loci, self.st = build_engine(wl, rank, world, data=args.data)
        eng = self.eng
        eng.set_update_priors(t_max=[PRIOR_T])
        self.full = 1 if args.schedule == "full" else 0
        if self.full:
            eng.set_update_schedule(3, 5)
        if args.pipeline:
            eng.set_pipeline(*[int(x) for x in args.pipeline.split(",")])
        if args.proposal:
            eng.set_proposal_path(*[int(x) for x in args.proposal.split(",")])
        if os.environ.get("IMA_SPEC"):
            eng.set_speculation(int(os.environ["IMA_SPEC"]))
        self.work_stream = torch.cuda.Stream()         # a real (non-default) stream: kernels and the timing events all go here
        torch.cuda.set_stream(self.work_stream)
        self.stream = self.work_stream.cuda_stream
        self.swaptries = eng.default_swaptries()
        self.pinned = {k: torch.from_numpy(np.ascontiguousarray(self.st[k])).pin_memory() for k in STATE_KEYS}
        self.bufs = [self.pinned[k].data_ptr() for k in STATE_KEYS]
        eng.put_state(self.bufs, self.st["tvals"], self.stream)
        torch.cuda.synchronize()
        if world > 1:
            from ima2p_b200.multirank import attach_exchange
            attach_exchange(eng, torch.cuda.current_device())

    def run_steps(self, n):
        if self.world == 1:
            self.eng.run(n, self.swaptries, self.stream)
        else:
            self.eng.run_sharded(n, self.swaptries, self.stream)

    def barrier(self):
        import torch.distributed as dist
        self.torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
            self.torch.cuda.synchronize()

    def timed(self, steps):
        """ms for `steps` steps, device events on the launching stream, max over ranks."""
        import torch.distributed as dist
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        self.run_steps(steps)
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sim50x128", choices=sorted(WORKLOADS))
    ap.add_argument("--data", default="real", choices=["synthetic", "real"],
                    help="synthetic: Simulations-shaped loci drawn by ima2p_b200.synth; real: the reference's own Simulations/*.u loci "
                         "(as held by the golden fixtures), written as a .u file and read by the product's reader")
    ap.add_argument("--burn", type=int, default=BURN, help="untimed whole steps before warm-up (the same in both arms)")
    ap.add_argument("--pipeline", default="", help="groups,depth[,decisions_first] of Engine.set_pipeline (default: the engine's)")
    ap.add_argument("--proposal", default="", help="fast,pairs_per_warp of Engine.set_proposal_path (default: the engine's)")
    ap.add_argument("--schedule", default="full", choices=["full", "genealogy"],
                    help="full: qupdate's schedule (genealogies, split time every step, mutation scalars every 5th, swaps); "
                         "genealogy: updategenealogy + swaps only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-lmode", action="store_true")
    ap.add_argument("--no-config3", action="store_true")
    ap.add_argument("--no-models", action="store_true", help="skip the HKY / stepwise / joint per-model throughput section (N = 1)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    wl = args.workload
    nloci, n0, n1, cpg = WORKLOADS[wl][:4]
    W = max(3, args.warmup)
    metric, unit = "chain_x_locus_genealogy_updates_per_sec", "updates/s"
    config = make_config(wl, world, args.schedule, args.data)
    full = 1 if args.schedule == "full" else 0

    if args.impl == "reference":
        if rank != 0:
            return
        r = reference_throughput(wl, world, args.steps, W, budget_s=100.0, full=full, data=args.data, burn=args.burn)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_harness was not built (needs /root/reference at build time)"}))
            return
        print(json.dumps({"impl": "reference", "metric": metric, "value": r["value"], "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": W, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f64", "data": args.data, "config": config,
                          "cpu_baseline": {"value": r["value"], "unit": unit, "cores": r["cores"], "kind": "reference", "sample": r["sample"]},
                          "e2e": {"value": r["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    if world > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group("nccl")
    dev = torch.device("cuda", torch.cuda.current_device())
    job = Job(wl, args, rank, world, dev)
    eng, stream, swaptries, pinned = job.eng, job.stream, job.swaptries, job.pinned
    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()               # started early: nvidia-smi needs a few hundred ms before its first line
    job.run_steps(args.burn)      # burn-in: leave the artificial starting genealogies behind (untimed)
    job.run_steps(W)
    job.barrier()
    c0 = eng.counters()
    sampler.mark()
    ms = job.timed(args.steps)    # the public path: Engine.run / Engine.run_sharded (captured graphs)
    kernel_ms = None
    if world == 1:
        # same K steps again, kernel by kernel with CUDA events on the launching stream around each kernel
        # (no overlap between kernels): per-kernel durations for the roofline accounting
        kernel_ms = eng.run_timed(args.steps, swaptries, stream)
        torch.cuda.synchronize()
    c1 = eng.counters()
    # a short timed region can end before nvidia-smi (100 ms period) has reported three times: the same step keeps running,
    # untimed, until it has, so that the clocks are always read under this load
    clocks_extended = False
    t_ext = time.perf_counter()
    for _ in range(40):
        enough = 1.0 if (sampler.proc is None or sampler.count() >= 3 or time.perf_counter() - t_ext > 3.0) else 0.0
        if world > 1:             # every rank takes the same number of extra steps (they exchange swap sums)
            f = torch.tensor([enough], dtype=torch.float64, device=dev)
            dist.all_reduce(f, op=dist.ReduceOp.MIN)
            enough = float(f.item())
        if enough:
            break
        job.run_steps(100)
        torch.cuda.synchronize()
        clocks_extended = True
    clocks = sampler.stop()
    clocks["extended_past_timed_region"] = clocks_extended
    eng.sync()
    updates_all = cpg * world * nloci * args.steps
    value = updates_all / (ms * 1e-3)
    p_acc = (c1["accepted"] - c0["accepted"]) / max(1, c1["updates"] - c0["updates"])

    # the same K steps with the genealogy updates + swaps only (no split-time / scalar updates), for the record
    graph_ms = None
    if world == 1 and full:
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.set_update_schedule(False, 0)
        eng.run(8, swaptries, stream)
        torch.cuda.synchronize()
        g0.record(); eng.run(args.steps, swaptries, stream); g1.record()
        torch.cuda.synchronize()
        graph_ms = g0.elapsed_time(g1)
        eng.set_update_schedule(3, 5)

    # ---- end to end through the C ABI with HOST buffers: every step uploads the genealogies from pinned host
    # memory (H2D), evaluates them, runs one M-mode step and reads the per-chain results back (D2H)
    ke = min(args.steps, 50)
    sb = eng.state_bytes()
    # the burned-in genealogies come back to pinned host memory once (untimed); they are what every e2e step uploads
    eng.fetch_state([pinned[k].data_ptr() for k in STATE_KEYS[:7]], stream)
    torch.cuda.synchronize()
    tv_now, uv_now, _ = eng.fetch_parameters()          # the split times and scalars the burned-in genealogies belong to
    tv_now = np.ascontiguousarray(tv_now)
    pinned["uvals"].numpy()[:] = uv_now.reshape(pinned["uvals"].shape)
    split_time_mean = float(tv_now.mean())
    mig_mean = float(pinned["scal_i"].numpy()[:, 1].mean())
    mig_max = int(pinned["scal_i"].numpy()[:, 1].max())
    # the migration pools travel as their used columns only (ima2p_engine_put_state)
    h2d = int(sum(sb)) - int(sb[3] + sb[4]) + cpg * nloci * mig_max * 10 + tv_now.nbytes
    d2h = (cpg * 4 + eng.rowlen + 2) * 8          # the packed step report (ima2p_engine_step_report)
    # one-block wire form (ima2p_engine_put_state_block) when the sample fits it: the same genealogies with 8-bit links and
    # migration counts (13 instead of 20 bytes per edge) and the migration events stored ragged, as ONE pinned host block ->
    # one PCIe transfer per step; packed once here (untimed), uploaded from pinned memory every step
    put, wire = (lambda: eng.put_state(job.bufs, tv_now, stream)), "put_state"
    blk = None if os.environ.get("IMA_PLAIN_UPLOAD") else eng.pack_state_block([pinned[k].numpy() for k in STATE_KEYS], tv_now)
    if blk is not None:
        pinned["block"] = torch.from_numpy(blk[0]).pin_memory()
        bptr, nev = pinned["block"].data_ptr(), blk[1]
        put, wire = (lambda: eng.put_state_block(bptr, nev, stream)), "put_state_block"
        h2d = int(pinned["block"].numel())
    for _ in range(10):           # the first transfers from a freshly pinned buffer are slow on this (virtualised) PCIe path
        put(); job.run_steps(1); eng.step_report(stream)
    # Every step's state is uploaded from pinned host memory and every step's results are read back.  With the one-block form
    # the transfer of step s+1 is started (on a copy stream, into the second staging slot) before the results of step s are
    # waited for, so the copy engine works while the step's kernels run: upload_block / adopt_block instead of put_state_block.
    overlap = blk is not None and not os.environ.get("IMA_SERIAL_UPLOAD")
    copy_stream = torch.cuda.Stream() if overlap else None
    serial_report = bool(os.environ.get("IMA_SERIAL_REPORT"))

    def e2e_loop(n):
        if not overlap:
            for _ in range(n):
                put()
                job.run_steps(1)
                out = eng.step_report(stream)
            return out
        # ... and the results of step s are read (ima2p_engine_step_report_begin / _end, two slots) after step s+1 has been
        # queued, so the device does not wait for the host between steps; every step's report is still copied and read
        eng.upload_block(bptr, nev, copy_stream.cuda_stream)
        for i in range(n):
            eng.adopt_block(stream)
            if i + 1 < n:
                eng.upload_block(bptr, nev, copy_stream.cuda_stream)
            job.run_steps(1)
            if serial_report:
                out = eng.step_report(stream)
                continue
            eng.step_report_begin(i & 1, stream)
            if i:
                out = eng.step_report_end((i - 1) & 1)
        return out if serial_report else eng.step_report_end((n - 1) & 1)

    for _ in range(3):             # untimed: the (virtualised) PCIe path needs ~100 transfers from a freshly pinned block to settle
        e2e_loop(ke)
    # the figure moves with the state of the (virtualised) PCIe path: median of five repetitions of the ke-step loop
    reps = []
    for _ in range(5):
        job.barrier()
        t0 = time.perf_counter()
        summ, _row = e2e_loop(ke)
        job.barrier()
        e2e_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        reps.append(e2e_s)
    e2e_value = cpg * world * nloci * ke / float(np.median(reps))
    assert np.isfinite(summ).all()
    # where an end-to-end step goes (each part synchronised on its own; untimed for the metric)
    parts = {"upload_and_evaluate": 0.0, "step": 0.0, "read_back": 0.0}
    for _ in range(10):
        t0 = time.perf_counter(); put(); torch.cuda.synchronize()
        t1 = time.perf_counter(); job.run_steps(1); torch.cuda.synchronize()
        t2 = time.perf_counter(); eng.step_report(stream); torch.cuda.synchronize()
        t3 = time.perf_counter()
        parts["upload_and_evaluate"] += (t1 - t0) * 100; parts["step"] += (t2 - t1) * 100; parts["read_back"] += (t3 - t2) * 100

    lmode_multi = None
    if world > 1 and not args.no_lmode:
        try:
            lmode_multi = lmode_bench_sharded(eng, dev, rank, world, job)
        except Exception as ex:
            lmode_multi = {"error": str(ex)}
    upd_counters = {k: int(v) for k, v in eng.update_counters().items()}
    # per chain group: k_move, k_weigh, k_propose_redo, k_accept, k_split_t_fast, k_split_t_redo, k_accept_t, k_changeu; k_swap once
    launches_per_step = eng.launches_per_step()

    # ---- BASELINE configs[2]'s shape on the same GPUs: 300 loci x 256 chains per GPU ------------------------------------------
    config3 = None
    if world > 1 and not args.no_config3 and wl != "sim300x256":
        try:
            job.barrier()
            eng.close()
            torch.cuda.empty_cache()
            job3 = Job("sim300x256", args, rank, world, dev)
            job3.run_steps(min(args.burn, 600))
            job3.run_steps(W)
            k3 = max(10, args.steps // 4)
            ms3 = job3.timed(k3)
            l3, _, _, c3 = WORKLOADS["sim300x256"][:4]
            v3 = c3 * world * l3 * k3 / (ms3 * 1e-3)
            cnt3 = job3.eng.counters()
            config3 = {"config": make_config("sim300x256", world, args.schedule, args.data), "value": v3, "unit": unit, "ms_per_step": ms3 / k3,
                       "steps": k3, "accept_rate": cnt3["accepted"] / max(1, cnt3["updates"]), "dropped_for_capacity": cnt3["dropped"],
                       "gpu_launches": job3.eng.launches_per_step() * k3}
            if rank == 0 and not args.no_cpu_baseline:
                r3 = reference_throughput("sim300x256", world, 2, 1, budget_s=15.0, full=full, data=args.data)
                if r3 is not None:
                    config3["cpu_baseline"] = {"value": r3["value"], "unit": unit, "cores": r3["cores"], "kind": "reference", "sample": r3["sample"]}
                    config3["ratio_to_cpu_baseline"] = v3 / r3["value"]
            job3.barrier()
            job3.eng.close()
        except Exception as ex:
            config3 = {"error": repr(ex)}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk, pk_kind = peaks()
    roof = None
    if kernel_ms is not None:
        names = ["proposal_kernels", "k_accept", "k_swap", "k_split_t", "k_accept_t", "k_changeu", "k_move", "k_weigh", "k_propose_redo"]
        per = np.asarray(kernel_ms, dtype=np.float64)[:len(names)] / args.steps              # ms per launch
        P = cpg * nloci
        b_update = algorithmic_bytes_per_update(n0 + n1, mig_mean, p_acc, eng.NI, eng.ND)      # B_IS of SURVEY.md section 8(d)
        W_g = 4 * eng.NI + 8 * eng.ND
        b_pair = 24.0 * (2 * (n0 + n1) - 1) + 12.0 * mig_mean + W_g + 24.0                # one genealogy with its weights
        # algorithmic bytes per launch of every kernel: what it must read and write of the pairs' records
        alg = {"proposal_kernels": b_update * P, "k_move": 2.0 * (b_pair - W_g) * P, "k_weigh": (b_pair + W_g + 24.0) * P,
               "k_propose_redo": 0.0, "k_accept": (2 * W_g + 16 + 12 + 1) * P + (W_g + 8 * (eng.nq + eng.nm) + 40) * cpg,
               "k_swap": 16.0 * cpg, "k_split_t": 2.0 * b_pair * P, "k_accept_t": (W_g + 8 + 16 + 1) * P, "k_changeu": 48.0 * P}
        timed_names = [n for n in names if n != "proposal_kernels"] if per[6] > 0 else names[:6]
        dom = max(timed_names, key=lambda n: per[names.index(n)])
        dom_ms = per[names.index(dom)]
        # the roofline line of the contract: SURVEY 8(d)'s per-update figure x the updates of one launch over the dominant
        # kernel's launch time; the same bytes over the whole step, and every kernel against its own bytes, beside it
        achieved = b_update * P / (dom_ms * 1e-3) / 1e9
        traffic = None
        tf = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tf):
            traffic = json.load(open(tf)).get(dom)
        step_gbps = b_update * P / (ms / args.steps * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": pk["hbm_gbs"], "peak_kind": pk_kind, "unit": "GB/s",
                "frac": achieved / pk["hbm_gbs"], "traffic": traffic, "algorithmic_bytes_per_update": b_update,
                "algorithmic_bytes_per_launch": b_update * P,
                "whole_step": {"achieved": step_gbps, "frac": step_gbps / pk["hbm_gbs"]},
                "kernel_ms_per_launch": {n: float(v) for n, v in zip(names, per)},
                "per_kernel": {n: {"algorithmic_bytes_per_launch": alg[n],
                                   "GBps": (alg[n] / (per[names.index(n)] * 1e-3) / 1e9) if per[names.index(n)] > 0 else None,
                                   "frac": (alg[n] / (per[names.index(n)] * 1e-3) / 1e9 / pk["hbm_gbs"]) if per[names.index(n)] > 0 else None}
                               for n in names},
                "note": "latency/dependency-bound path: small FP64/integer work per pair, see DESIGN.md"}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        r = reference_throughput(wl, 1, 3, 1, budget_s=12.0, full=full, data=args.data, burn=args.burn)
        if r is not None:
            cpu = {"value": r["value"], "unit": unit, "cores": r["cores"], "kind": "reference", "sample": r["sample"], "accept_rate": r["accept"]}
    models = None
    if world == 1 and not args.no_models:
        try:
            models = models_bench(stream)
        except Exception as ex:
            models = {"error": repr(ex)}
    lmode = lmode_multi
    if not args.no_lmode and world == 1:
        try:
            lmode = lmode_bench(eng, dev)
        except Exception as ex:       # the L-mode line is supplementary; the M-mode metric must still be reported
            lmode = {"error": str(ex)}
    out = {"metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": W, "ms_per_step": ms / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": args.data, "config": config,
           "clocks": clocks, "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": ke, "parts_ms": parts,
                                     "upload": wire + (" as upload_block + adopt_block: the block of step s+1 travels while step s runs" if overlap else ""),
                                     "read_back": ("step_report_begin / _end (two slots): the report of step s is waited for after step s+1 has been queued"
                                                   if overlap and not serial_report else "step_report: one synchronisation per step"),
                                     "repetitions_s": reps, "statistic": "median of 5 repetitions"},
           "gpu_launches": launches_per_step * args.steps,
           "roofline": roof, "cpu_baseline": cpu, "accept_rate": p_acc, "mig_events_per_genealogy": mig_mean, "mig_events_max": mig_max,
           "genealogy_updates_only": ({"ms_per_step": graph_ms / args.steps, "value": updates_all / (graph_ms * 1e-3), "unit": unit}
                                      if graph_ms else None), "lmode": lmode, "config3": config3, "models": models,
           "multi_gpu_step": ("one CUDA graph per step on every rank; swap sums exchanged through peer memory inside the kernels "
                              "(ima2p_engine_run_sharded)" if world > 1 else None),
           "dropped_for_capacity": c1["dropped"], "split_time_mean": split_time_mean, "update_counters": upd_counters,
           "swap_rate": (c1["swaps"] - c0["swaps"]) / max(1, c1["swap_attempts"] - c0["swap_attempts"])}
    print(json.dumps(out, default=float))
    if world > 1:
        dist.destroy_process_group()


MODEL_FIXTURES = {
    # mutation model -> golden fixture written by the reference (tests/golden/generate.py): HKY = Sim1_5loci relabelled H,
    # stepwise / joint = the synthetic S1 / J1 inputs of SURVEY.md section 8(d) (BASELINE configs[4])
    "hky": "state_sim5_hky_hn2", "stepwise": "state_sim3_sw_hn2", "joint_is_sw": "state_sim3_joint_hn2",
}


def models_bench(stream, nloci=50, nchains=128, burn=300, steps=100):
    """updates/s of the whole step on loci of the other mutation models (BASELINE configs[4]): the loci and
    the genealogies of a reference-written fixture, repeated to nloci loci x nchains chains.  Per model: value, step time,
    per-kernel times, the algorithmic bytes of an update (B_HKY = B_IS + 120 S R for HKY, R = internal nodes recomputed per
    proposal) and the HBM fraction they amount to."""
    import torch
    from ima2p_b200 import Engine, synth
    pk, _ = peaks()
    out = {}
    for name, fx in MODEL_FIXTURES.items():
        d = json.load(gzip.open(os.path.join(ROOT, "tests", "golden", fx + ".json.gz")))
        fl, fc = d["loci"], d["chains"]
        eng = Engine(nchains, nloci, mig_capacity=64, seed=77)
        eng.set_model(**synth.two_population_model(PRIOR_Q, PRIOR_M))
        for li in range(nloci):
            L = fl[li % len(fl)]
            eng.set_locus(li, L["model"], L["numgenes"], L["numsites"], L["samppop"], seq=L["seq"] if L["seq"] else None, mult=L.get("mult"),
                          hval=L["hval"], totsites=L["totsites"], nlinked=L["nlinked"], minA=L["minA"], maxA=L["maxA"], sumlogk=L["sumlogk"])
        eng.finalize()
        eng.set_heating(*HEAT)
        for c in range(nchains):
            ch = fc[c % len(fc)]
            eng.set_chain(c, ch["tvals"])
            for li in range(nloci):
                g = ch["G"][li % len(fl)]
                t = g["tree"]
                off, mt, mp = [0], [], []
                for lst in t["mig"]:
                    mt += lst[0::2]; mp += lst[1::2]; off.append(len(mt))
                A = np.stack([np.asarray(a, np.int32) for a in t["A"]]) if t.get("A") else None
                eng.set_genealogy(c, li, t["up0"], t["up1"], t["down"], t["pop"], t["time"], off, mt, mp, t["root"], t["roottime"],
                                  uvals=g["uvals"], kappa=g["kappa"], pi=g["pi"], A=A)
        eng.upload()
        eng.eval()
        eng.set_update_priors(t_max=[PRIOR_T])
        eng.set_update_schedule(3, 5)
        sw = eng.default_swaptries()
        eng.run(burn, sw, stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0 = eng.counters()
        e0.record(); eng.run(steps, sw, stream); e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        km = np.asarray(eng.run_timed(steps, sw, stream), np.float64) / steps
        torch.cuda.synchronize()
        c1 = eng.counters()
        eng.sync()
        p_acc = (c1["accepted"] - c0["accepted"]) / max(1, c1["updates"] - c0["updates"])
        n = fl[0]["numgenes"]
        b_is = algorithmic_bytes_per_update(n, 2.0, p_acc, eng.NI, eng.ND)
        S = float(np.mean([L["numsites"] for L in fl]))
        # internal nodes on the union of the root paths of the touched edges: about twice the expected depth of a node
        R = 2.0 * np.log2(n) if name == "hky" else 0.0
        extra = 120.0 * S * R if name == "hky" else 12.0 * (2 * n - 1) * max(1, fl[0]["nlinked"] - (1 if name == "joint_is_sw" else 0))
        b_upd = b_is + extra
        P = nloci * nchains
        names = ["proposal_kernels", "k_accept", "k_swap", "k_split_t", "k_accept_t", "k_changeu", "k_move", "k_weigh", "k_propose_redo"]
        out[name] = {"fixture": fx, "loci": nloci, "chains": nchains, "genes_per_locus": n, "value": P / (ms * 1e-3), "unit": "updates/s", "ms_per_step": ms,
                     "accept_rate": p_acc, "kernel_ms_per_launch": {k: float(v) for k, v in zip(names, km)},
                     "algorithmic_bytes_per_update": b_upd, "formula": ("B_IS + 120 S R, S = %.1f patterns, R = %.1f nodes" % (S, R)) if name == "hky" else "B_IS + 12 bytes per edge and linked stepwise part",
                     "roofline": {"bound": "hbm", "achieved": b_upd * P / (ms * 1e-3) / 1e9, "frac": b_upd * P / (ms * 1e-3) / 1e9 / pk["hbm_gbs"], "unit": "GB/s"},
                     "proposal_path": "two kernels (k_move + k_weigh)" if km[6] > 0 else "general one-warp-per-pair kernel", "dropped_for_capacity": c1["dropped"]}
        eng.close()
    return out


def fp64_peaks(device=0):
    """(FP64 fused multiply-adds per second, FP64 exp evaluations per second) measured on this GPU by the library's own
    micro-kernels (ima2p_debug_fp64_peaks): what the FP64-bound L-mode kernels are quoted against."""
    import ctypes as C
    from ima2p_b200 import capi
    out = (C.c_double * 2)()
    capi.check(capi.lib(), capi.lib().ima2p_debug_fp64_peaks(device, out))
    return float(out[0]), float(out[1])


def lmode_bench(eng, dev, G=1000000):
    """L-mode evals/sec (BASELINE config 4 shape: 1e6 sampled genealogies, 21 floats each): margincalc on the
    1000-point grids of all parameters (histograms.cpp:81-99) and jointp for a differential-evolution population."""
    import torch
    from ima2p_b200 import LMode
    rows = []
    for _ in range(200):                              # real cold-chain rows of this run, then bootstrap to G
        eng.run(2)
        r = eng.cold_row()
        if r is not None:
            rows.append(r.copy())
    base = np.stack(rows)
    rng = np.random.default_rng(5)
    big = base[rng.integers(0, len(base), G)]
    lm = LMode(eng.nq, eng.nm, eng.nsplit, [PRIOR_Q] * 3, [0.0] * 3, [PRIOR_M] * 2, [0.0] * 2)
    lm.load(big)
    grid = [(np.arange(1000) + 0.5) / 1000 * (PRIOR_Q if p < 3 else PRIOR_M) for p in range(5)]
    lm.margincalc(grid[0], 0.0, 0, 0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for p in range(5):
        lm.margincalc(grid[p], 0.0, p, 0)
    t1 = time.perf_counter()
    NV = 512                                          # as the sharded measurement: a differential-evolution generation
    xs = np.column_stack([rng.uniform(0.05, 0.9, NV) * (PRIOR_Q if p < 3 else PRIOR_M) for p in range(5)])
    lm.jointp(xs[:64])
    t2 = time.perf_counter()
    lm.jointp(xs)
    t3 = time.perf_counter()
    # section 8 (f3) evaluators over the same rows: one unit = one genealogy contributing to one evaluation
    lm.moments()
    t4 = time.perf_counter()
    lm.moments()                                      # calcx of 5 parameters in 2 modes per row + the 10 products
    t5 = time.perf_counter()
    xg = (np.arange(100) + 0.5) / 100 * PRIOR_Q * PRIOR_M / 2
    lm.popmig(0, 0, xg[:8])
    t6 = time.perf_counter()
    for ti, mi in ((0, 0), (1, 1)):
        lm.popmig(ti, mi, xg)                         # the 100-bin scan of marginalopt_popmig, 2NM of both populations
    t7 = time.perf_counter()
    lm.greater_than(0, 0, 1)
    t8 = time.perf_counter()
    for a, b in ((0, 1), (1, 0), (0, 2), (2, 0)):
        lm.greater_than(0, a, b)                      # quadrature over 20,000 rows each (USETREESMAX)
    t9 = time.perf_counter()
    lm.close()
    # the marginal kernel serves 8 evaluation points per pass over the rows (4 columns for a size parameter, 3 for a migration
    # parameter, 4 bytes each): bytes it streams, against the HBM figure; the 84 MB of rows stay in the 126 MB L2 after
    # the first pass, so the kernel is bound by FP64 exp throughput (one exp per genealogy-eval), not by HBM
    passes = -(-1000 // 8)
    streamed = G * 4.0 * passes * (3 * 4 + 2 * 3)
    pk, _ = peaks()
    fma_s, exp_s = fp64_peaks(torch.cuda.current_device())
    return {"rows": G, "margincalc_geneval_per_sec": 5 * 1000 * G / (t1 - t0), "jointp_geneval_per_sec": NV * G / (t3 - t2), "jointp_vectors": NV,
            "fp64_fma_per_sec_measured": fma_s, "fp64_exp_per_sec_measured": exp_s,
            # one exp per genealogy-evaluation in margincalc; one exp (eexp) + 2 (3 nq + 2 nm) multiply-adds in jointp
            "fp64_frac": {"margincalc_vs_exp_peak": 5 * 1000 * G / (t1 - t0) / exp_s, "jointp_vs_exp_peak": NV * G / (t3 - t2) / exp_s},
            "unit": "genealogy evals/s", "timing": "host wall clock around the C-ABI calls (includes H2D of x and D2H of results)",
            "margincalc_streamed_GBps": streamed / (t1 - t0) / 1e9, "margincalc_streamed_frac_of_hbm_peak": streamed / (t1 - t0) / 1e9 / pk["hbm_gbs"],
            "margincalc_algorithmic_GBps_one_pass_per_x": 5 * 1000 * G * 15.2 / (t1 - t0) / 1e9,
            "moments_geneval_per_sec": 10 * G / (t5 - t4), "popmig_geneval_per_sec": 200 * G / (t7 - t6),
            "greater_than_rows_per_sec": 4 * 20000 / (t9 - t8),
            "f3_note": "calcx: 2 incomplete gammas per parameter and row; 2NM density: 2-4 per row and point; greater-than: a trapezoid quadrature of incomplete gammas per row -- FP64-bound"}


def lmode_bench_sharded(eng, dev, rank, world, job, G=1000000):
    """The same L-mode evaluations with the G rows sharded over the ranks (BASELINE config 4): every evaluation is local
    partial sums plus one small collective (ima2p_b200/multirank.py).  Whole-job evals/s, slowest rank."""
    import torch
    import torch.distributed as dist
    from ima2p_b200 import LMode
    from ima2p_b200.multirank import sharded_jointp_device, sharded_margincalc
    mine = []
    for _ in range(200):                              # cold-chain rows of this run, wherever the cold chain lives
        job.run_steps(2)
        r = eng.cold_row()
        if r is not None:
            mine.append(r.copy())
    every = [None] * world
    dist.all_gather_object(every, mine)
    base = np.stack([r for part in every for r in part])
    rng = np.random.default_rng(5)
    big = base[rng.integers(0, len(base), G)]
    lo, hi = G * rank // world, G * (rank + 1) // world
    lm = LMode(eng.nq, eng.nm, eng.nsplit, [PRIOR_Q] * 3, [0.0] * 3, [PRIOR_M] * 2, [0.0] * 2, device=torch.cuda.current_device())
    lm.load(big[lo:hi], nrows_total=G, row0=lo)
    grid = [(np.arange(1000) + 0.5) / 1000 * (PRIOR_Q if p < 3 else PRIOR_M) for p in range(5)]
    sharded_margincalc(lm, grid[0], 0.0, 0, 0, device=dev)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for p in range(5):
        sharded_margincalc(lm, grid[p], 0.0, p, 0, device=dev)
    torch.cuda.synchronize(); dist.barrier()
    t1 = time.perf_counter()
    # a differential-evolution generation of the joint search: 500 trial vectors (100 x the 5 parameters, jointfind.cpp:599-812),
    # in batches of 32 queued on one stream with two device all-gathers each and one copy back at the end
    NV = 512
    xs = np.column_stack([rng.uniform(0.05, 0.9, NV) * (PRIOR_Q if p < 3 else PRIOR_M) for p in range(5)])
    sharded_jointp_device(lm, xs[:64], device=dev)
    torch.cuda.synchronize(); dist.barrier()
    t2 = time.perf_counter()
    q1, _ = sharded_jointp_device(lm, xs, device=dev)
    torch.cuda.synchronize(); dist.barrier()
    t3 = time.perf_counter()
    lm.close()
    t = torch.tensor([t1 - t0, t3 - t2], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    fma_s, exp_s = fp64_peaks(torch.cuda.current_device())
    return {"rows": G, "rows_per_rank": hi - lo, "margincalc_geneval_per_sec": 5 * 1000 * G / float(t[0]), "jointp_geneval_per_sec": NV * G / float(t[1]),
            "jointp_vectors": NV, "fp64_exp_per_sec_measured_per_gpu": exp_s,
            "fp64_frac": {"margincalc_vs_exp_peak_all_gpus": 5 * 1000 * G / float(t[0]) / (exp_s * world), "jointp_vs_exp_peak_all_gpus": NV * G / float(t[1]) / (exp_s * world)},
            "unit": "genealogy evals/s", "checksum": float(np.sum(q1)),
            "timing": "host wall clock around the sharded calls incl. the collectives, max over ranks"}


if __name__ == "__main__":
    main()
