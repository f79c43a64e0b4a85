"""N > 1 path on CPU: two gloo ranks, each holding half of the chains in the tests-only host-emulation build of
the kernels, must reproduce the single-process run bit for bit (random streams are keyed by GLOBAL chain and
locus, swap attempts are replayed identically on every rank)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "hostemu", "libima2p_hostemu.so")
FIXTURE, NSTEPS, SWAPTRIES = "state_sim5_hn4", 25, 3


def _run_single():
    from ima2p_b200 import capi
    from support import engine_from_fixture, load_golden
    d = load_golden(FIXTURE)
    betas = [ch["beta"] for ch in d["chains"]]
    eng, _ = engine_from_fixture(d, lib=capi.bind(EMU), seed=99, betas=betas)
    eng.eval()
    eng.set_update_priors(t_max=[3.0])
    eng.set_update_schedule(3, 5)              # the whole qupdate step: split-time and scalar updates are local to a chain
    eng.run(NSTEPS, SWAPTRIES)
    out = np.array([[eng.chain(c)["probg"], eng.chain(c)["pdg"], eng.chain(c)["tvals"][0]] for c in range(eng.nchains)])
    return out, eng.betas(), eng.counters()


def _worker(rank, world, port, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ima2p_b200 import Engine, capi
    from ima2p_b200.multirank import ShardedStepper
    from support import FlatModel, FlatTree, load_golden
    d = load_golden(FIXTURE)
    fm = FlatModel(d["model"])
    nglob = len(d["chains"])
    per = nglob // world
    eng = Engine(per, len(d["loci"]), seed=99, lib=capi.bind(EMU), nchains_global=nglob, chain0=per * rank)
    eng.set_model_flat(*fm.create_args())
    for li, loc in enumerate(d["loci"]):
        eng.set_locus(li, loc["model"], loc["numgenes"], loc["numsites"], loc["samppop"], seq=loc["seq"], hval=loc["hval"],
                      sumlogk=loc["sumlogk"])
    eng.finalize()
    eng.set_betas([ch["beta"] for ch in d["chains"]])
    for k in range(per):
        ch = d["chains"][per * rank + k]
        eng.set_chain(k, ch["tvals"])
        for li, g in enumerate(ch["G"]):
            t = FlatTree(g["tree"])
            eng.set_genealogy(k, li, t.up0, t.up1, t.down, t.pop, t.time, t.mig_off, t.mig_t[:-1], t.mig_p[:-1], t.root,
                              t.roottime, uvals=g["uvals"])
    eng.upload()
    eng.eval()
    eng.set_update_priors(t_max=[3.0])
    eng.set_update_schedule(3, 5)
    ShardedStepper(eng, torch.device("cpu")).run(NSTEPS, SWAPTRIES)
    out = np.array([[eng.chain(c)["probg"], eng.chain(c)["pdg"], eng.chain(c)["tvals"][0]] for c in range(per)])
    q.put((rank, out, eng.betas(), eng.counters()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_reproduce_the_single_process_run():
    subprocess.run([os.path.join(HERE, "hostemu", "build.sh")], check=True)
    single, betas1, cnt1 = _run_single()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    both = np.concatenate([g[1] for g in got])
    assert np.array_equal(both, single)                    # bit-identical chains regardless of sharding
    for g in got:
        assert np.array_equal(g[2], betas1)                # every rank ends with the same beta permutation
        assert g[3]["swap_attempts"] == cnt1["swap_attempts"] and g[3]["swaps"] == cnt1["swaps"]
    assert sum(g[3]["accepted"] for g in got) == cnt1["accepted"]
    assert cnt1["swaps"] > 0 and len(set(single[:, 2])) > 1      # split times moved


def _lmode_worker(rank, world, port, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ima2p_b200 import LMode, capi
    from ima2p_b200.multirank import sharded_jointp, sharded_margincalc
    from support import FlatModel, load_golden
    d = load_golden("lmode_sim5_hn2")
    fm = FlatModel(d["model"])
    rows = np.ascontiguousarray(d["rows"], dtype=np.float32)
    G = len(rows)
    cut = [0, G // 3 + 11, G]          # uneven shards
    lm = LMode(fm.nq, fm.nm, fm.nsplit, fm.q_max, fm.q_min, fm.m_max, fm.m_min, fm.m_mean, fm.expoprior, lib=capi.bind(EMU))
    lm.load(rows[cut[rank]:cut[rank + 1]], nrows_total=G, row0=cut[rank])
    xs = np.array([j["x"] for j in d["jointp"]])[:32]
    qv, ess = sharded_jointp(lm, xs)
    tab = [t for t in d["margincalc"] if t[0] == 1]
    mc = sharded_margincalc(lm, np.array([t[1] for t in tab]), 0.25, 1, 1)
    q.put((rank, qv, ess, mc))
    dist.barrier()
    dist.destroy_process_group()


def test_lmode_sharded_over_two_ranks_matches_reference():
    from support import _num, load_golden, rel_close
    subprocess.run([os.path.join(HERE, "hostemu", "build.sh")], check=True)
    d = load_golden("lmode_sim5_hn2")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_lmode_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref_q = [j["q"] for j in d["jointp"]][:32]
    ref_e = [j["ess"] for j in d["jointp"]][:32]
    ref_mc = [_num(t[3]) for t in d["margincalc"] if t[0] == 1]
    for _, qv, ess, mc in got:          # every rank ends with the same, reference-matching values
        assert rel_close(qv, ref_q, 1e-10) and rel_close(ess, ref_e, 1e-8)
        assert rel_close(mc, ref_mc, 1e-10)


def _lmode_extra_worker(rank, world, port, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ima2p_b200 import LMode, capi
    from ima2p_b200.multirank import sharded_moments, sharded_popmig
    from support import FlatModel, load_golden
    d = load_golden("lmode_extra_sim5_hn2")
    fm = FlatModel(d["model"])
    rows = np.ascontiguousarray(d["rows"], dtype=np.float32)
    G = len(rows)
    cut = [0, G // 3 + 7, G]
    lm = LMode(fm.nq, fm.nm, fm.nsplit, fm.q_max, fm.q_min, fm.m_max, fm.m_min, fm.m_mean, fm.expoprior, lib=capi.bind(EMU))
    lm.load(rows[cut[rank]:cut[rank + 1]], nrows_total=G, row0=cut[rank])
    means, var, corr = sharded_moments(lm)
    sel = [t for t in d["popmig"] if t[0] == 1 and t[1] == 0]
    pm = sharded_popmig(lm, 1, 0, np.array([t[2] for t in sel]))
    q.put((rank, means, var, corr, pm))
    dist.barrier()
    dist.destroy_process_group()


def test_lmode_moments_and_popmig_sharded_over_two_ranks_match_reference():
    """section 8 (f3) evaluators with the rows split over two ranks: all-reduced row sums, the reference's values."""
    from support import _num, load_golden, rel_close
    subprocess.run([os.path.join(HERE, "hostemu", "build.sh")], check=True)
    d = load_golden("lmode_extra_sim5_hn2")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_lmode_extra_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    G = len(d["rows"])
    n = len(d["calcx"]["sum0"])
    rm = np.array([_num(v) for v in d["calcx"]["sum0"]]) / G
    rv = np.array([_num(v) for v in d["calcx"]["sum1"]]) / G - rm * rm
    rc = np.array([_num(v) for v in d["calcx"]["cross"]]).reshape(n, n)
    ref_pm = [_num(t[3]) for t in d["popmig"] if t[0] == 1 and t[1] == 0]
    for _, means, var, corr, pm in got:
        assert rel_close(means, rm, 1e-10) and rel_close(var, rv, 1e-7)
        for a in range(n - 1):
            for b in range(a + 1, n):
                assert abs(corr[a, b] - (rc[a, b] / G - rm[a] * rm[b]) / np.sqrt(rv[a] * rv[b])) < 1e-6
        assert rel_close(pm, ref_pm, 1e-10, 1e-300)


def _shard_engine(d, fm, lib, rank, world, seed=99):
    from ima2p_b200 import Engine
    from support import FlatTree
    nglob = len(d["chains"])
    per = nglob // world
    eng = Engine(per, len(d["loci"]), seed=seed, lib=lib, nchains_global=nglob, chain0=per * rank)
    eng.set_model_flat(*fm.create_args())
    for li, loc in enumerate(d["loci"]):
        eng.set_locus(li, loc["model"], loc["numgenes"], loc["numsites"], loc["samppop"], seq=loc["seq"], hval=loc["hval"],
                      sumlogk=loc["sumlogk"])
    eng.finalize()
    eng.set_betas([ch["beta"] for ch in d["chains"]])
    for k in range(per):
        ch = d["chains"][per * rank + k]
        eng.set_chain(k, ch["tvals"])
        for li, g in enumerate(ch["G"]):
            t = FlatTree(g["tree"])
            eng.set_genealogy(k, li, t.up0, t.up1, t.down, t.pop, t.time, t.mig_off, t.mig_t[:-1], t.mig_p[:-1], t.root,
                              t.roottime, uvals=g["uvals"])
    eng.upload()
    eng.eval()
    eng.set_update_priors(t_max=[3.0])
    eng.set_update_schedule(3, 5)
    return eng


@pytest.mark.parametrize("world", [2, 4])
def test_peer_memory_exchange_reproduces_the_single_process_run(world):
    """The exchange the product uses between GPUs (include/ima2p_b200.h, ima2p_engine_exchange_*): every rank's kernels store
    their chains' swap sums into every rank's table and the swap kernel reads its own table -- here `world` shard engines of one
    process, attached to each other's tables through the export / import calls and stepped in lockstep (every rank's update,
    then every rank's swap).  The chains must be the single-engine run's, bit for bit."""
    from ima2p_b200 import capi
    from support import FlatModel, load_golden
    subprocess.run([os.path.join(HERE, "hostemu", "build.sh")], check=True)
    single, betas1, cnt1 = _run_single()
    lib = capi.bind(EMU)
    d = load_golden(FIXTURE)
    fm = FlatModel(d["model"])
    engs = [_shard_engine(d, fm, lib, r, world) for r in range(world)]
    handles = [e.exchange_handle() for e in engs]
    for r, e in enumerate(engs):
        e.exchange_attach([None if k == r else e.exchange_open(handles[k]) for k in range(world)])
    for _ in range(NSTEPS):
        for e in engs:
            e.sharded_update()
        for e in engs:
            e.sharded_swap(SWAPTRIES)
    for e in engs:
        e.sync()                                           # raises if a swap kernel did not find every chain's sum
    both = np.concatenate([np.array([[e.chain(c)["probg"], e.chain(c)["pdg"], e.chain(c)["tvals"][0]] for c in range(e.nchains)]) for e in engs])
    assert np.array_equal(both, single)
    for e in engs:
        assert np.array_equal(e.betas(), betas1)
        assert e.counters()["swaps"] == cnt1["swaps"]
    assert sum(e.counters()["accepted"] for e in engs) == cnt1["accepted"]
    # a rank that steps alone finds the others' sums missing: the device error word says so (no hang, no silent garbage)
    engs[0].sharded_update()
    engs[0].sharded_swap(SWAPTRIES)
    with pytest.raises(capi.Ima2pError):
        engs[0].sync()


def test_device_resident_sharded_jointp_equals_the_single_rank_jointp():
    """ima2p_lmode_joint_begin / _middle (the phases of the sharded jointp with every intermediate left in device memory, as the
    NCCL path of ima2p_b200.multirank.sharded_jointp_device uses them): three uneven shards held by three LMode objects of one
    process, the all-gathers played by concatenation, against jointp over all rows on one object (jointfind.cpp:885-1047)."""
    from ima2p_b200 import LMode, capi
    from support import FlatModel, load_golden
    subprocess.run([os.path.join(HERE, "hostemu", "build.sh")], check=True)
    lib = capi.bind(EMU)
    d = load_golden("lmode_sim5_hn2")
    fm = FlatModel(d["model"])
    rows = np.ascontiguousarray(d["rows"], dtype=np.float32)
    G = len(rows)
    mk = lambda: LMode(fm.nq, fm.nm, fm.nsplit, fm.q_max, fm.q_min, fm.m_max, fm.m_min, fm.m_mean, fm.expoprior, lib=lib)
    whole = mk()
    whole.load(rows)
    xs = np.array([j["x"] for j in d["jointp"]])[:40]
    q1, e1 = whole.jointp(xs[:32])
    q1b, e1b = whole.jointp(xs[32:])
    want_q, want_e = np.r_[q1, q1b], np.r_[e1, e1b]
    cut = [0, G // 4 + 7, G // 2 + 3, G]
    world = 3
    lms = []
    for r in range(world):
        lm = mk()
        lm.load(rows[cut[r]:cut[r + 1]], nrows_total=G, row0=cut[r])
        lms.append(lm)
    got_q, got_e = [], []
    for b0 in range(0, len(xs), 256):                          # one call for all vectors (up to 512)
        xb = xs[b0:b0 + 256]
        nv = len(xb)
        local = [np.zeros(nv) for _ in range(world)]
        for r, lm in enumerate(lms):
            lm.joint_begin(xb, local[r].ctypes.data)
        allmax = np.ascontiguousarray(np.concatenate(local))
        rec = [np.zeros(nv * 8) for _ in range(world)]
        for r, lm in enumerate(lms):
            lm.joint_middle(nv, allmax.ctypes.data, world, r, rec[r].ctypes.data)
        a = np.stack(rec).reshape(world, nv, 8)
        for v in range(nv):
            tot = a[:, v, :6].sum(axis=0)
            k = int(np.argmin(a[:, v, 4]))
            tot[4], tot[5] = a[k, v, 4], a[k, v, 5]
            q, e = lms[0].joint_finish(tot, a[0, v, 6], True)
            got_q.append(q); got_e.append(e)
    assert np.allclose(got_q, want_q, rtol=1e-12, atol=0) and np.allclose(got_e, want_e, rtol=1e-9, atol=0), (got_q[:3], want_q[:3])
    # against the reference's own values of the fixture
    ref = np.array([j["q"] for j in d["jointp"]])[:40]
    assert np.allclose(got_q, ref, rtol=1e-9, atol=0)
