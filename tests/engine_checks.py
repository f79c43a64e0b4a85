"""Parity checks shared by the GPU tests (-m gpu, product library) and the CPU tests that run the same
kernel sources through the tests-only host-emulation build (tests/hostemu).  The checker is always the
CPU oracle (pinned against the reference's own values) or the golden fixtures themselves."""
import numpy as np
import pytest

from support import (FlatModel, FlatTree, OracleModel, check_static_eval, engine_from_fixture, f64, fp, dp, i32, load_golden, oracle,
                     rel_close, split_weights, tree_from_engine, _num)

STATIC_FIXTURES = ["state_sim5_hn4", "state_sim50_hn3", "state_sim300_hn1", "state_sim2_hn2", "state_sim3_hn3",
                   "state_sim5_3pop_hn2", "state_sim5_expo_hn2", "state_sim3_sw_hn2", "state_sim5_hky_hn2", "state_sim3_joint_hn2",
                   "state_sim5_nomig_hn2", "state_sim5_3pop_nomig_hn2", "state_sim5_4popA_hn2", "state_sim5_4popB_hn2"]
STEP_FIXTURES = ["state_sim5_hn4", "state_sim3_hn3", "state_sim5_3pop_hn2", "state_sim2_hn2"]


def static_eval_matches_reference(lib, name, rtol):
    """a5 + a6 + a7 (+ a9): integers bit-exact, doubles within rtol of the reference's init_p values."""
    d = load_golden(name)
    eng, fm = engine_from_fixture(d, lib=lib)
    eng.eval()
    check_static_eval(eng, fm, d, rtol=rtol)
    eng.close()


def proposals_match_oracle(lib, name, nsteps, rtol=1e-9, need_root_moves=True):
    """a2-a4 on identical genealogies, no RNG-path matching: for every proposal the device makes, the oracle
    recomputes from the (before, after) genealogies (i) the forward/reverse migration-path probabilities,
    (ii) the slide weight when the root moved, (iii) the proposed genealogy's weights and likelihood."""
    d = load_golden(name)
    eng, fm = engine_from_fixture(d, lib=lib, mig_capacity=96)
    om = OracleModel(fm)
    eng.eval()
    nch, nl = eng.nchains, eng.nloci
    nchecked = nroot = nrej = 0
    for _ in range(nsteps):
        before = {(c, l): tree_from_engine(eng.get_genealogy(c, l, 0)) for c in range(nch) for l in range(nl)}
        buf0 = {(c, l): eng.proposal(c, l)["buffer"] for c in range(nch) for l in range(nl)}
        old = {(c, l): eng.pair(c, l) for c in range(nch) for l in range(nl)}
        eng.run(1, swaptries=0)
        eng.sync()
        for c in range(nch):
            tv = d["chains"][c]["tvals"]
            for l in range(nl):
                pr = eng.proposal(c, l)
                if pr["flags"] & 2:
                    continue                       # dropped for capacity (counted)
                if pr["flags"] & 1:                # infinite-sites reject: state must be untouched
                    nrej += 1
                    assert pr["buffer"] == buf0[(c, l)]
                    continue
                accepted = pr["buffer"] != buf0[(c, l)]
                after = tree_from_engine(eng.get_genealogy(c, l, 0 if accepted else 1))
                b = before[(c, l)]
                fwd, rev = om.migration_logprobs(tv, b, after, pr["edge"])
                if fm.nomigration:          # update_gtree.cpp:815-818: no migration path is simulated, the weight is 0
                    assert pr["migweight"] == 0.0 and after.mig_off[-1] == 0
                else:
                    assert rel_close(rev - fwd, pr["migweight"], rtol, 1e-9), (name, c, l, fwd, rev, pr)
                if b.root != after.root or b.roottime != after.roottime:
                    sw = oracle().ora_slideweight(pr["slidedist"], b.roottime, after.roottime)
                    assert rel_close(sw, pr["slideweight"], rtol, 1e-9), (sw, pr)
                    nroot += 1
                else:
                    assert pr["slideweight"] == 0.0
                loc = d["loci"][l]
                w = om.treeweight(tv, loc, after)
                assert w["mignum"] >= 0
                new = eng.pair(c, l) if accepted else None
                if accepted:
                    ew = split_weights(fm, new["wi"], new["wd"])
                    assert np.array_equal(ew["cc"], w["cc"]) and np.array_equal(ew["mc"], w["mc"])
                    assert rel_close(ew["fc"], w["fc"], rtol) and rel_close(ew["fm"], w["fm"], rtol)
                    assert new["mignum"] == w["mignum"]
                    gj = d["chains"][c]["G"][l]
                    if loc["model"] == 0:
                        assert rel_close(new["pdg"], om.likelihood_is(loc, after, w["length"], gj["uvals"][0]), rtol)
                    elif loc["model"] == 1:
                        assert rel_close(new["pdg"], om.likelihood_hky(loc, after, gj["pi"], gj["uvals"][0], gj["kappa"]), rtol)
                else:
                    assert eng.pair(c, l)["pdg"] == old[(c, l)]["pdg"]
                nchecked += 1
    cnt = eng.counters()
    assert nchecked > 0 and (nroot > 0 or not need_root_moves), (nchecked, nroot)
    assert 0.1 < cnt["accepted"] / cnt["updates"] < 0.7, cnt     # reference: 0.32-0.40
    eng.close()
    return cnt


def incremental_sums_match_fresh_evaluation(lib, name, nsteps, rtol=1e-9, full_schedule=False):
    """After many accepted/rejected updates (and swaps) the per-chain sums maintained by the accept kernel
    (sum_subtract_treeinfo + integrate_tree_prob with its reuse rule) equal a from-scratch evaluation."""
    d = load_golden(name)
    eng, fm = engine_from_fixture(d, lib=lib)
    eng.eval()
    if full_schedule:           # split-time update every step, mutation scalars every 5th (qupdate's schedule)
        eng.set_update_priors(t_max=[3.0] * fm.nsplit)
        eng.set_update_schedule(3, 5)
    eng.run(nsteps)
    eng.sync()
    if full_schedule:
        uc = eng.update_counters()
        assert uc["t_tries"] == nsteps * eng.nchains and 0 < uc["t_accepts"] < uc["t_tries"]
        nur = sum(l["nlinked"] for l in d["loci"])
        if nur > 1:
            assert uc["u_tries"] == (nsteps // 5) * eng.nchains * (nur - (nur == 2)) and 0 < uc["u_accepts"] < uc["u_tries"]
        for c in range(eng.nchains):
            tv = eng.chain(c)["tvals"][:fm.nsplit]
            assert np.all(np.diff(np.concatenate([[0.0], tv, [3.0]])) > 0)          # ordered, inside the prior
            logu = sum(float(np.sum(np.log(eng.scalars(c, l)[0][:d["loci"][l]["nlinked"]]))) for l in range(eng.nloci))
            start = sum(float(np.sum(np.log(g["uvals"]))) for g in d["chains"][c]["G"])
            assert abs(logu - start) < 1e-9                                         # the product of the scalars is invariant
    inc = [eng.chain(c) for c in range(eng.nchains)]
    incp = [[eng.pair(c, l) for l in range(eng.nloci)] for c in range(eng.nchains)]
    betas = eng.betas()
    assert rel_close(sorted(betas), sorted(ch["beta"] for ch in d["chains"]), 0.0)   # swaps only permute betas
    eng.eval()
    for c in range(eng.nchains):
        fresh = eng.chain(c)
        assert np.array_equal(fresh["wi"], inc[c]["wi"])
        assert rel_close(fresh["wd"], inc[c]["wd"], rtol, 1e-9)
        assert rel_close(fresh["probg"], inc[c]["probg"], rtol) and rel_close(fresh["pdg"], inc[c]["pdg"], rtol)
        for l in range(eng.nloci):
            f = eng.pair(c, l)
            assert np.array_equal(f["wi"], incp[c][l]["wi"]) and rel_close(f["pdg"], incp[c][l]["pdg"], 1e-12)
    cnt = eng.counters()
    row = eng.cold_row()
    assert row is not None and row.shape == (fm.rowlen,)
    eng.close()
    return cnt


def lmode_joint_models_match_reference(lib, rtol=1e-9):
    """jointp as a function of the population sizes only / the migration rates only (nowmodeltype 1 / 2, the two searches of a
    three-population analysis, jointfind.cpp:949-952, 973-980, 1118-1133) against the reference's own values; single call,
    wide batches and the two-rank sharded form."""
    from ima2p_b200 import LMode
    d = load_golden("lmode_extra_sim5_3pop_hn2")
    fm = FlatModel(d["model"])
    rows = np.ascontiguousarray(d["rows"], dtype=np.float32)
    G = len(rows)
    mk = lambda: LMode(fm.nq, fm.nm, fm.nsplit, fm.q_max, fm.q_min, fm.m_max, fm.m_min, fm.m_mean, fm.expoprior, lib=lib)
    lm = mk()
    lm.load(rows)
    for mt in (1, 2):
        tab = d["jointp_type%d" % mt]
        xs = np.array([j["x"] for j in tab])
        lm.set_joint_model(mt)
        q, ess = lm.jointp(xs, True)
        assert rel_close(q, [j["q"] for j in tab], rtol), (mt, q[:3])
        assert rel_close(ess, [j["ess"] for j in tab], 1e-8)
        # the entries outside the model's range are not read
        xo = xs.copy()
        if mt == 1:
            xo[:, fm.nq:] = 1.0
        else:
            xo[:, :fm.nq] = 1.0
        assert np.array_equal(lm.jointp(xo, True)[0], q)
        # rows split over two handles
        half = G // 2 + 11
        a, b = mk(), mk()
        a.load(rows[:half], nrows_total=G, row0=0)
        b.load(rows[half:], nrows_total=G, row0=half)
        a.set_joint_model(mt); b.set_joint_model(mt)
        nv = len(xs)
        ma = a.joint_phase1(xs)
        mb = b.joint_phase1(xs, seed_before=ma)
        gmax = np.maximum(ma, mb)
        ra, rb = a.joint_phase2(nv, gmax), b.joint_phase2(nv, gmax)
        for v in range(nv):
            rec = ra[v] + rb[v]
            lo = ra[v] if ra[v][4] <= rb[v][4] else rb[v]
            rec[4], rec[5] = lo[4], lo[5]
            qv, ev = a.joint_finish(rec, gmax[v])
            assert rel_close(qv, q[v], 1e-12) and rel_close(ev, ess[v], 1e-9)
        a.close(); b.close()
    lm.set_joint_model(0)
    with pytest.raises(Exception):
        lm.set_joint_model(3)
    lm.close()


def lmode_matches_reference(lib, name, rtol=1e-9):
    """a14 + a15: margincalc / marginp / jointp against the reference's own values on the same rows."""
    from ima2p_b200 import LMode
    d = load_golden(name)
    fm = FlatModel(d["model"])
    rows = np.ascontiguousarray(d["rows"], dtype=np.float32)
    lm = LMode(fm.nq, fm.nm, fm.nsplit, fm.q_max, fm.q_min, fm.m_max, fm.m_min, fm.m_mean, fm.expoprior, lib=lib)
    lm.load(rows)
    G = len(rows)
    tab = d["margincalc"]
    for p in sorted(set(t[0] for t in tab)):
        sel = [t for t in tab if t[0] == p]
        x = np.array([t[1] for t in sel])
        assert rel_close(lm.margincalc(x, 0.0, p, 0), [_num(t[2]) for t in sel], rtol, 1e-300)
        assert rel_close(lm.margincalc(x, 0.25, p, 1), [_num(t[3]) for t in sel], rtol)
        assert rel_close(lm.marginp(p, 0, G, x), [_num(t[4]) for t in sel], rtol, 1e-300)
        assert rel_close(lm.marginp(p, G // 3, 2 * G // 3, x), [_num(t[5]) for t in sel], rtol, 1e-300)
    # the lock-step form (one pass for the current points of many searches) gives the single calls' values bit for bit
    P = sorted(set(t[0] for t in tab))
    rng = np.random.default_rng(5)
    n = 64
    par = rng.choice(P, n)
    kind = rng.integers(0, 2, n)
    first = np.where(rng.random(n) < 0.5, 0, rng.integers(0, G // 2, n))
    last = np.where(rng.random(n) < 0.5, G, first + 1 + rng.integers(0, G // 2, n))
    hi = np.array([(fm.q_max[p] if p < fm.nq else fm.m_max[p - fm.nq]) for p in par])
    x = rng.random(n) * hi * 1.05 + 1e-7                # a few beyond the prior (OFFSCALEVAL)
    ya = rng.random(n)
    many = lm.marginal_many(kind, par, first, last, x, ya)
    for k in range(n):
        one = lm.marginp(int(par[k]), int(first[k]), int(last[k]), x[k:k + 1]) if kind[k] == 0 else \
            lm.margincalc(x[k:k + 1], ya[k], int(par[k]), 1)
        assert many[k] == one[0], (k, kind[k], many[k], one[0])
    assert len(lm.marginal_many([], [], [], [], [])) == 0
    with pytest.raises(Exception):
        lm.marginal_many([0], [P[0]], [5], [5], [1.0])       # empty row range
    xs = np.array([j["x"] for j in d["jointp"]])
    q, ess = lm.jointp(xs, True)
    assert rel_close(q, [j["q"] for j in d["jointp"]], rtol), (q[:4], [j["q"] for j in d["jointp"]][:4])
    assert rel_close(ess, [j["ess"] for j in d["jointp"]], 1e-8)
    # sharded form == single form (rows split in two, host-side exchange of the per-vector maxima)
    half = G // 2 + 37
    a = LMode(fm.nq, fm.nm, fm.nsplit, fm.q_max, fm.q_min, fm.m_max, fm.m_min, fm.m_mean, fm.expoprior, lib=lib)
    b = LMode(fm.nq, fm.nm, fm.nsplit, fm.q_max, fm.q_min, fm.m_max, fm.m_min, fm.m_mean, fm.expoprior, lib=lib)
    a.load(rows[:half], nrows_total=G, row0=0)
    b.load(rows[half:], nrows_total=G, row0=half)
    nv = min(len(xs), 32)
    ma = a.joint_phase1(xs[:nv])
    mb = b.joint_phase1(xs[:nv], seed_before=ma)
    gmax = np.maximum(ma, mb)
    ra, rb = a.joint_phase2(nv, gmax), b.joint_phase2(nv, gmax)
    for v in range(nv):
        rec = ra[v] + rb[v]
        lo = ra[v] if ra[v][4] <= rb[v][4] else rb[v]
        rec[4], rec[5] = lo[4], lo[5]
        qv, ev = a.joint_finish(rec, gmax[v])
        assert rel_close(qv, q[v], 1e-12) and rel_close(ev, ess[v], 1e-9)
    s0 = lm.marginal_sums(0, xs[:8, 0])
    assert rel_close(a.marginal_sums(0, xs[:8, 0]) + b.marginal_sums(0, xs[:8, 0]), s0, 1e-12)
    for o in (lm, a, b):
        o.close()


def lmode_moments_and_popmig_match_reference(lib, name, rtol=1e-9):
    """section 8 (f3): the calcx sums behind print_means_variances_correlations (output.cpp:14-134, 687-745), the table it
    prints, and the 2NM densities calc_popmig / marginpopmig or their exponential-prior forms (popmig.cpp:9-357) against
    the reference's own values on the same rows."""
    from ima2p_b200 import LMode
    d = load_golden(name)
    fm = FlatModel(d["model"])
    rows = np.ascontiguousarray(d["rows"], dtype=np.float32)
    lm = LMode(fm.nq, fm.nm, fm.nsplit, fm.q_max, fm.q_min, fm.m_max, fm.m_min, fm.m_mean, fm.expoprior, lib=lib)
    lm.load(rows)
    G, n = len(rows), fm.nq + fm.nm
    means, var, corr, raw = lm.moments()
    cx = d["calcx"]
    ref0, ref1 = np.array([_num(v) for v in cx["sum0"]]), np.array([_num(v) for v in cx["sum1"]])
    refc = np.array([_num(v) for v in cx["cross"]]).reshape(n, n)
    ok = np.isfinite(ref0) & np.isfinite(ref1)          # the reference's own sums overflow for the exponential-prior terms
    assert ok[:fm.nq].all()
    assert rel_close(raw["sum0"][ok], ref0[ok], rtol) and rel_close(raw["sum1"][ok], ref1[ok], rtol)
    okc = np.isfinite(refc) & np.outer(ok, ok)
    assert rel_close(raw["cross"][okc], refc[okc], rtol)
    # the printed table (%.3lf): Mean and Stdv lines
    lines = cx["table"].split("\n")
    mean_line = [ln for ln in lines if ln.startswith("Mean:")][0].split("\t")[1:]
    sd_line = [ln for ln in lines if ln.startswith("Stdv:")][0].split("\t")[1:]
    for p in range(n):
        if ok[p] and "nan" not in mean_line[p]:
            assert abs(float(mean_line[p]) - means[p]) <= 0.00051, (p, mean_line[p], means[p])
        if ok[p] and "nan" not in sd_line[p]:
            assert abs(float(sd_line[p]) - np.sqrt(var[p])) <= 0.00051, (p, sd_line[p], var[p])
    # finishing arithmetic of output.cpp:709-739 redone here from the reference's own sums
    rm = ref0 / G
    rv = ref1 / G - rm * rm
    assert rel_close(means[ok], rm[ok], rtol) and rel_close(var[ok], rv[ok], 1e-7)
    for p in range(n - 1):
        for q in range(p + 1, n):
            if okc[p, q] and rv[p] > 0 and rv[q] > 0:
                assert abs(corr[p, q] - (refc[p, q] / G - rm[p] * rm[q]) / np.sqrt(rv[p] * rv[q])) < 1e-6
    tab = d["popmig"]
    npairs = 0
    for (ti, mi) in sorted(set((t[0], t[1]) for t in tab)):
        sel = [t for t in tab if t[0] == ti and t[1] == mi]
        x = np.array([t[2] for t in sel])
        assert rel_close(lm.popmig(ti, mi, x, 0), [_num(t[3]) for t in sel], rtol, 1e-300), (ti, mi)
        assert rel_close(lm.popmig(ti, mi, x, 1), [_num(t[4]) for t in sel], rtol, 1e-300), (ti, mi)
        assert rel_close(lm.marginpopmig(mi, 0, G, x, ti), [_num(t[5]) for t in sel], rtol, 1e-300), (ti, mi)
        assert rel_close(lm.marginpopmig(mi, G // 3, 2 * G // 3, x, ti), [_num(t[6]) for t in sel], rtol, 1e-300), (ti, mi)
        npairs += 1
    # greater-than probabilities (gtint.cpp): closed forms and the trapezoid quadrature, every pair the reference computes
    for kind, i, j, v in d["greater_than"]:
        assert rel_close(lm.greater_than(kind, i, j), _num(v), max(rtol, 1e-8)), (kind, i, j)
    assert lm.greater_than(0, 0, 0) == -1.0
    lm.close()
    return npairs


def gamma_tables_match_reference(lib, rtol=1e-10):
    """a10: uppergamma / lowergamma on the device (one-lane and warp-cooperative forms) against the reference's
    own values on a grid that straddles the series / continued-fraction switch at x = a + 1."""
    import ctypes as C
    k = load_golden("kat_sim5_hn4")
    up = {(a, x): _num(v) for a, x, v in k["uppergamma"]}
    lo = {(a, x): _num(v) for a, x, v in k["lowergamma"]}
    keys = sorted(up)
    a = np.array([t[0] for t in keys], np.int32)
    x = np.array([t[1] for t in keys], np.float64)
    out = np.zeros((len(keys), 4))
    rc = lib.ima2p_debug_gamma(0, a.ctypes.data_as(C.POINTER(C.c_int)), x.ctypes.data_as(C.POINTER(C.c_double)), len(keys),
                               out.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0, lib.ima2p_last_error()
    for i, key in enumerate(keys):
        for col in (0, 2):
            assert rel_close(out[i, col], up[key], rtol, 1e-12), ("uppergamma", key, col, out[i, col], up[key])
        if key in lo:
            for col in (1, 3):
                assert rel_close(out[i, col], lo[key], rtol, 1e-12), ("lowergamma", key, col, out[i, col], lo[key])
    return len(keys)


def stepwise_updates_match_oracle(lib, name, nsteps, rtol=1e-9):
    """a9: for every stepwise proposal the device makes, (i) the incrementally updated branch terms and locus
    likelihood equal the oracle's full likelihoodSW of the proposed genealogy and allele states, (ii) the Hastings
    term of the junction-allele draw equals the ratio of the two geometric probabilities of finishSWupdateA
    (update_gtree_common.cpp:2264-2359), recomputed here from the before/after states."""
    import math
    d = load_golden(name)
    eng, fm = engine_from_fixture(d, lib=lib)
    om = OracleModel(fm)
    eng.eval()
    nch, nl = eng.nchains, eng.nloci
    geo = lambda j, w: min(j / (w + j), 0.95)
    nchk = nmoved = 0
    for _ in range(nsteps):
        before = {(c, l): (tree_from_engine(eng.get_genealogy(c, l, 0)), eng.get_alleles(c, l, 0)) for c in range(nch) for l in range(nl)}
        buf0 = {(c, l): eng.proposal(c, l)["buffer"] for c in range(nch) for l in range(nl)}
        eng.run(1, swaptries=0)
        eng.sync()
        for c in range(nch):
            for l in range(nl):
                pr = eng.proposal(c, l)
                if pr["flags"] & 3:
                    continue
                acc = pr["buffer"] != buf0[(c, l)]
                after = tree_from_engine(eng.get_genealogy(c, l, 0 if acc else 1))
                al = eng.get_alleles(c, l, 0 if acc else 1)
                bt, bal = before[(c, l)]
                g = d["chains"][c]["G"][l]
                aterm = 0.0
                joint = d["loci"][l]["model"] == 3
                if joint:           # part 0 of a J locus is the infinite-sites part, evaluated in full
                    w = om.treeweight(d["chains"][c]["tvals"], d["loci"][l], after)
                    assert rel_close(al["pdg_a"][0], om.likelihood_is(d["loci"][l], after, w["length"], g["uvals"][0]), rtol)
                for ai in range(1 if joint else 0, d["loci"][l]["nlinked"]):
                    after.A = [i32(a) for a in al["A"]]
                    like, dl = om.likelihood_sw(after, ai, g["uvals"][ai])
                    assert rel_close(al["pdg_a"][ai], like, rtol), (c, l, al["pdg_a"], like)
                    nz = after.down != -1
                    assert rel_close(al["dlikeA"][ai][nz], dl[nz], rtol, 1e-12) and np.all(al["dlikeA"][ai][~nz] == 0)
                    # Hastings term of the allele draw
                    edge = pr["edge"]
                    junction = int(bt.down[edge])                 # the freed edge keeps its number
                    A0, A1 = bal["A"][ai], al["A"][ai]
                    oldA, newA = int(A0[junction]), int(A1[junction])
                    nb_new = [A1[edge], A1[after.up0[junction] if after.up0[junction] != edge else after.up1[junction]]]
                    if after.down[junction] != -1:
                        nb_new.append(A1[after.down[junction]])
                    oldsis = bt.up0[junction] if bt.up0[junction] != edge else bt.up1[junction]
                    nb_old = [A0[edge], A0[oldsis]] + ([A0[bt.down[junction]]] if bt.down[junction] != -1 else [])
                    gn = geo(len(nb_new), sum(abs(int(a) - oldA) for a in nb_new))
                    go = geo(len(nb_old), sum(abs(int(a) - newA) for a in nb_old))
                    dA = abs(newA - oldA)
                    aterm += (dA * math.log(1 - go) + math.log(go)) - (dA * math.log(1 - gn) + math.log(gn))
                    nmoved += dA > 0
                assert abs(aterm - pr["aterm"]) < 1e-9 * max(1.0, abs(aterm)), (aterm, pr)
                nchk += 1
    assert nchk > 0 and nmoved > 0
    cnt = eng.counters()
    eng.close()
    return cnt


def long_run_summaries_match_reference(lib, name, nchains, burn, sweeps, nsigma=5.0, full_schedule=False):
    """Statistical parity (no RNG matching): with split times and mutation scalars held at the reference's start
    values, the engine's long-run per-locus means of tree length, root time, migration count and per-population
    coalescence counts agree with the reference's own updategenealogy() sampler (fixture written by
    `ref_harness trace`) within nsigma combined standard errors."""
    from ima2p_b200 import Engine
    d = load_golden(name)
    fm = FlatModel(d["model"])
    nloci = len(d["loci"])
    eng = Engine(nchains, nloci, mig_capacity=96, seed=4242, lib=lib)
    eng.set_model_flat(*fm.create_args())
    for li, loc in enumerate(d["loci"]):
        eng.set_locus(li, loc["model"], loc["numgenes"], loc["numsites"], loc["samppop"], seq=loc["seq"], hval=loc["hval"])
    eng.finalize()
    eng.set_betas([1.0] * nchains)                      # independent cold chains, no swapping
    for c in range(nchains):
        eng.set_chain(c, d["tvals"])
        for li in range(nloci):
            t = FlatTree(d["start"][li])
            eng.set_genealogy(c, li, t.up0, t.up1, t.down, t.pop, t.time, t.mig_off, t.mig_t[:-1], t.mig_p[:-1], t.root,
                              t.roottime, uvals=[d["uvals"][li]])
    eng.upload()
    eng.eval()
    if full_schedule:
        # the whole qupdate schedule (genealogies, RY1 or NW split-time update at random, mutation scalars): the posterior
        # summaries of t and of the scalars must agree with the reference's
        eng.set_update_priors(t_max=d["tprior_max"])
        eng.set_update_schedule(3, 5)
    eng.run(burn, swaptries=0)
    acc = np.zeros((nchains, nloci, 6))
    tacc, uacc = np.zeros((nchains, fm.nsplit)), np.zeros((nchains, nloci))
    first_of_period = np.cumsum([0] + [fm.npops - k for k in range(fm.nsplit)])
    for _ in range(sweeps):
        eng.run(1, swaptries=0)
        if full_schedule:
            tv, uv, _ = eng.fetch_parameters()
            tacc += tv
            uacc += np.log(uv[:, :, 0])
        sd, si, wi = eng.fetch_pair_summaries()
        sd, si, wi = sd.reshape(nchains, nloci, 4), si.reshape(nchains, nloci, 2), wi.reshape(nchains, nloci, -1)
        acc[:, :, 0] += sd[:, :, 1]
        acc[:, :, 1] += sd[:, :, 0]
        acc[:, :, 2] += si[:, :, 1]
        acc[:, :, 3] += wi[:, :, 0]
        acc[:, :, 4] += wi[:, :, 1] if fm.npops > 1 else 0
        acc[:, :, 5] += wi[:, :, first_of_period[1:]].sum(axis=2)       # cc[k][0], k >= 1 (what the trace fixture sums)
    chain_means = acc / sweeps
    m_e, se_e = chain_means.mean(axis=0), chain_means.std(axis=0, ddof=1) / np.sqrt(nchains)
    bm = np.array(d["batch_means"])
    m_r, se_r = bm.mean(axis=0), bm.std(axis=0, ddof=1) / np.sqrt(len(bm))
    z = (m_e - m_r) / np.sqrt(se_e ** 2 + se_r ** 2 + 1e-300)
    cnt = eng.counters()
    ref_acc = float(np.mean(d["accept"]))
    eng_acc = cnt["accepted"] / cnt["updates"]
    if full_schedule:
        def zscore(e, r):
            e, r = e / sweeps, np.array(r)
            return (e.mean(axis=0) - r.mean(axis=0)) / np.sqrt(e.var(axis=0, ddof=1) / nchains + r.var(axis=0, ddof=1) / len(r) + 1e-300)
        zt, zu = zscore(tacc, d["t_batch_means"]), zscore(uacc, d["logu_batch_means"])
        uc = eng.update_counters()
        eng.close()
        assert np.all(np.abs(z) < nsigma), (z, m_e, m_r)
        assert np.all(np.abs(zt) < nsigma) and np.all(np.abs(zu) < nsigma), (zt, zu, tacc.mean(axis=0) / sweeps, uacc.mean(axis=0) / sweeps)
        return z, zt, zu, tacc.mean(axis=0) / sweeps, np.array(d["t_batch_means"]).mean(axis=0), uc
    eng.close()
    assert np.all(np.abs(z) < nsigma), (z, m_e, m_r)
    assert abs(eng_acc - ref_acc) < 0.03, (eng_acc, ref_acc)
    return z, m_e, m_r, eng_acc, ref_acc


def _one_chain_fixture(d, chain):
    return {"model": d["model"], "loci": d["loci"], "chains": [chain]}


def split_time_update_matches_reference(lib, name, rtol=1e-9):
    """changet_RY1 (section 8 f1): with the reference's proposed time, the device's rescaled genealogies, weights,
    likelihoods, prior and Metropolis-Hastings term equal the reference's (accept forced on both sides)."""
    d = load_golden(name)
    for rec in d["records"]:
        b, a, period = rec["before"], rec["after"], rec["period"]
        eng, fm = engine_from_fixture(_one_chain_fixture(d, b), lib=lib)
        eng.eval()
        eng.set_update_priors(t_max=d["tprior_max"], t_min=d["tprior_min"])
        newt = a["tvals"][period]
        # rejected first: nothing may move
        out = eng.debug_split_time(period, [newt], force_accept=0)
        assert out[0, 3] == 0 and rel_close(out[0, 2], rec["mh"], rtol, 1e-8), (out[0], rec["mh"])
        check_static_eval(eng, fm, _one_chain_fixture(d, b), rtol=rtol)
        out = eng.debug_split_time(period, [newt], force_accept=1)
        assert out[0, 0] == period and out[0, 1] == newt and out[0, 3] == 1
        assert rel_close(out[0, 2], rec["mh"], rtol, 1e-8), (out[0], rec["mh"])
        check_static_eval(eng, fm, _one_chain_fixture(d, a), rtol=rtol)            # stored values == the reference's after the move
        assert rel_close(eng.chain(0)["tvals"][:fm.nsplit], a["tvals"], 0.0)
        for li, ga in enumerate(a["G"]):
            t, ta = tree_from_engine(eng.get_genealogy(0, li)), FlatTree(ga["tree"])
            assert rel_close(t.time, ta.time, 1e-14) and rel_close(t.roottime, ta.roottime, 1e-14)
            assert rel_close(np.sort(t.mig_t[:-1]), np.sort(ta.mig_t[:-1]), 1e-14)
        eng.eval()                                                                  # and a fresh evaluation agrees with them
        check_static_eval(eng, fm, _one_chain_fixture(d, a), rtol=rtol)
        eng.close()


def mutation_scalar_update_matches_reference(lib, name, rtol=1e-8):
    """changeu (section 8 f1): the reference's own proposals (partner, ratio step, kappas replayed from the uniforms
    it drew) evaluated on the device give the reference's new likelihoods and Metropolis-Hastings term."""
    import ctypes as C
    from support import changeu_replay
    d = load_golden(name)
    nur, ul = d["nurates"], d["ul"]
    for rec in d["records"]:
        b, a, j = rec["before"], rec["after"], rec["j"]
        k, U, rest = changeu_replay(rec["U"], j, nur)
        (lj, aj), (lk, ak) = ul[j], ul[k]
        uj, uk = b["G"][lj]["uvals"][aj], b["G"][lk]["uvals"][ak]
        dd = C.c_double()
        oracle().ora_changeu_newr(U, float(np.log(uj / uk)), d["u_win"], 3.0 * d["u_prmax"], C.byref(dd))
        kap = [0.0, 0.0]
        for i, li in enumerate((lj, lk)):
            if d["loci"][li]["model"] == 1:
                kap[i] = oracle().ora_new_kappa(rest.pop(0), b["G"][li]["kappa"], d["kappa_win"], d["kappa_max"])
        eng, fm = engine_from_fixture(_one_chain_fixture(d, b), lib=lib)
        eng.eval()
        out = eng.debug_changeu(0, j, k, dd.value, kap[0], kap[1])
        assert rel_close(out[2], rec["mh"], rtol, 1e-300), (out, rec["mh"])
        if rec["accepted"]:
            assert rel_close(out[0], a["G"][lj]["pdg_a"][aj], 1e-9) and rel_close(out[1], a["G"][lk]["pdg_a"][ak], 1e-9)
        check_static_eval(eng, fm, _one_chain_fixture(d, b), rtol=1e-9)             # a debug evaluation changes nothing
        eng.close()


def thermodynamic_integration_matches_reference(lib):
    from ima2p_b200 import Engine  # noqa: F401  (the Simpson rule is a host function of the same library)
    import ctypes as C
    for t in load_golden("kat_thermo")["thermo"]:
        s = f64(t["sums"])
        out = C.c_double()
        assert lib.ima2p_thermo_marginlike(dp(s), len(s), t["k"], C.byref(out)) == 0
        assert rel_close(out.value, t["value"], 1e-14)


def nielsen_wakeley_update_matches_oracle(lib, name, rtol=1e-9):
    """changet_NW on the device, from the reference's states and proposed split times, no RNG matching: for the move the
    device makes, the oracle (pinned to the reference's own NW moves) recomputes the migration Hastings term of every
    locus from the (before, proposed) genealogies, and the weights / prior of the proposed state."""
    d = load_golden(name)
    nmoved = nup = ndown = 0
    for rec in d["records"]:
        b, a, period = rec["before"], rec["after"], rec["period"]
        eng, fm = engine_from_fixture(_one_chain_fixture(d, b), lib=lib, mig_capacity=96)
        om = OracleModel(fm)
        eng.eval()
        oldt, newt = b["tvals"][period], a["tvals"][period]
        out = eng.debug_split_time(period, [newt], force_accept=0, method=1)
        assert out[0, 3] == 0
        check_static_eval(eng, fm, _one_chain_fixture(d, b), rtol=rtol)          # a rejected move leaves nothing behind
        tvn = list(b["tvals"]); tvn[period] = newt
        acc = dict(cc=np.zeros(fm.ncc, np.int64), mc=np.zeros(fm.nmc, np.int64), fc=np.zeros(fm.ncc), hcc=np.zeros(fm.ncc), fm=np.zeros(fm.nmc))
        migw = 0.0
        for li, gb in enumerate(b["G"]):
            tb = FlatTree(gb["tree"])
            tp = tree_from_engine(eng.get_genealogy(0, li, 1))                    # the proposed genealogy (other buffer)
            assert np.array_equal(tb.time, tp.time) and np.array_equal(tb.up0, tp.up0) and tb.root == tp.root
            touched = (newt > oldt and tb.roottime > oldt) or (newt < oldt and tb.roottime > newt)
            w = om.nw_migweight(b["tvals"], period, newt, tb, tp) if touched else 0.0
            pr = eng.proposal(0, li)
            assert abs(w - pr["migweight"]) <= 1e-9 * max(1.0, abs(w)), (li, w, pr)
            migw += w
            tw = om.treeweight(tvn, d["loci"][li], tp)                           # consistent populations, new weights
            assert tw["mignum"] == tp.mig_off[-1]
            for k in acc:
                acc[k] = acc[k] + tw[k]
            nmoved += int(tp.mig_off[-1] != tb.mig_off[-1])
        probg, _, _ = om.init_integrate(acc)
        mh = float(np.exp(b["beta"] * (probg - b["probg"]) + migw))
        assert rel_close(out[0, 2], mh, rtol, 1e-300), (out[0], mh)
        # the same proposal again (same step, same streams), accepted: stored sums equal a fresh evaluation
        out = eng.debug_split_time(period, [newt], force_accept=1, method=1)
        assert out[0, 3] == 1 and rel_close(out[0, 2], mh, rtol, 1e-300)
        inc = eng.chain(0)
        assert rel_close(inc["probg"], probg, rtol) and rel_close(inc["tvals"], tvn, 0.0) and rel_close(inc["pdg"], b["pdg"], 1e-12)
        eng.eval()
        fresh = eng.chain(0)
        assert np.array_equal(inc["wi"], fresh["wi"]) and rel_close(inc["wd"], fresh["wd"], rtol, 1e-12)
        assert rel_close(inc["probg"], fresh["probg"], rtol) and rel_close(inc["pdg"], fresh["pdg"], rtol)
        nup += newt > oldt
        ndown += newt < oldt
        eng.close()
    assert nmoved > 0 and nup > 0 and ndown > 0


def step_report_matches_separate_reads(lib, name="state_sim5_hn4"):
    """ima2p_engine_step_report == fetch_chain_summary + cold_row (the packed read-back used per step)."""
    d = load_golden(name)
    eng, fm = engine_from_fixture(d, lib=lib)
    eng.eval()
    eng.run(7)
    summ, row = eng.step_report()
    assert np.array_equal(summ, eng.fetch_chain_summary())
    ref_row = eng.cold_row()
    assert (row is None) == (ref_row is None) and (row is None or np.array_equal(row, ref_row))
    # the two-slot form: step s+1 is queued before the report of step s is read; every report is that step's
    eng.step_report_begin(0)
    with pytest.raises(Exception):
        eng.step_report_begin(0)                       # the slot holds an unread report
    eng.run(1)
    want1 = eng.fetch_chain_summary()
    eng.step_report_begin(1)
    eng.run(1)
    s0, r0 = eng.step_report_end(0)
    s1, _ = eng.step_report_end(1)
    assert np.array_equal(s0, summ) and (r0 is None) == (row is None) and (row is None or np.array_equal(r0, row))
    assert np.array_equal(s1, want1) and not np.array_equal(s1, s0)
    with pytest.raises(Exception):
        eng.step_report_end(1)                         # nothing was begun
    eng.close()


def cold_chain_counters_are_consistent(lib, nsteps=60):
    """ima2p_engine_cold_counters (what the reference's update-rate tables and swap table report, ima_main_mpi.cpp:3473-3899,
    swapchains.cpp:760-778).  With one chain that chain is the cold one, so its counts are the engine's totals exactly (a scalar
    proposal counting for both scalars it trades between, ima_main_mpi.cpp:1926-1935); with several chains the cold chain tries
    one split-time update a step and every scalar sweep, and its counts and the adjacent-temperature swaps are part of the totals."""
    d = load_golden("state_sim300_hn1")
    eng, fm = engine_from_fixture(d, lib=lib, seed=17)
    eng.eval()
    eng.set_update_priors(t_max=[3.0] * fm.nsplit)
    eng.set_update_schedule(3, 5)
    eng.run(nsteps)
    eng.sync()
    cnt, uc, cc = eng.counters(), eng.update_counters(), eng.cold_counters(fm.nsplit, eng.nloci)
    g = cc["genealogy"].sum(axis=0)
    assert (int(g[0]), int(g[1]), int(g[2])) == (cnt["accepted"], cnt["topology"], cnt["tmrca"]) and cnt["accepted"] > 0
    assert int(cc["split"][:, [0, 2]].sum()) == uc["t_tries"] == nsteps and int(cc["split"][:, [1, 3]].sum()) == uc["t_accepts"]
    assert cc["split"][:, 0].sum() > 0 and cc["split"][:, 2].sum() > 0                       # both update types were drawn
    assert int(cc["scalars"][:, 0].sum()) == 2 * uc["u_tries"] > 0 and int(cc["scalars"][:, 1].sum()) == 2 * uc["u_accepts"]
    assert np.all(cc["scalars"][:, 0] >= nsteps // 5)                                        # every scalar is proposed in every sweep
    eng.close()
    d = load_golden("state_sim5_hn4")
    eng, fm = engine_from_fixture(d, lib=lib, seed=17)
    eng.eval()
    eng.set_update_priors(t_max=[3.0] * fm.nsplit)
    eng.set_update_schedule(3, 5)
    eng.run(5 * nsteps)
    eng.sync()
    cnt, uc, cc = eng.counters(), eng.update_counters(), eng.cold_counters(fm.nsplit, eng.nloci)
    assert 0 < cc["genealogy"][:, 0].sum() < cnt["accepted"] and np.all(cc["genealogy"][:, 1] <= cc["genealogy"][:, 0])
    assert np.all(cc["genealogy"][:, 0] <= 5 * nsteps)
    assert int(cc["split"][:, [0, 2]].sum()) == 5 * nsteps and int(cc["split"][:, [1, 3]].sum()) <= uc["t_accepts"]
    assert int(cc["scalars"][:, 0].sum()) == 2 * eng.nloci * nsteps and np.all(cc["scalars"][:, 1] <= cc["scalars"][:, 0])
    sd = eng.fetch_pair_summaries()[0].reshape(eng.nchains, eng.nloci, 4)
    for c in range(eng.nchains):
        assert np.array_equal(eng.fetch_chain_pdg(c), sd[c, :, 3])
    adj = cc["adjacent"]
    assert adj.shape == (3, 2) and np.all(adj[:, 1] <= adj[:, 0]) and 0 < adj[:, 0].sum() <= cnt["swap_attempts"]
    assert adj[:, 1].sum() <= cnt["swaps"]
    eng.close()


def speculation_depth_does_not_change_the_run(lib, name="state_sim50_hn3", nsteps=60):
    """The accept sweep evaluates several loci of a chain per round speculatively (k_accept<B>); whatever the depth B, the
    run must be the one-locus-at-a-time sweep of update_gtree.cpp:917-927: same acceptances, same sums, same genealogies."""
    from support import engine_from_fixture, load_golden
    d = load_golden(name)
    outs = []
    for depth in (1, 2, 3, 4):
        eng, _ = engine_from_fixture(d, lib=lib, seed=31)
        eng.set_speculation(depth)
        eng.eval()
        eng.run(nsteps)
        eng.sync()
        ch = [eng.chain(c) for c in range(eng.nchains)]
        outs.append((eng.counters(), np.concatenate([np.r_[c["probg"], c["pdg"], c["wd"], c["wi"]] for c in ch])))
        eng.close()
    assert outs[0][0]["accepted"] > 0
    for k in (1, 2, 3):
        assert outs[k][0] == outs[0][0], (k, outs[k][0], outs[0][0])
        assert np.array_equal(outs[k][1], outs[0][1]), k
    return outs[0][0]


def pipeline_does_not_change_the_run(lib, name="state_sim50_hn3", nsteps=37):
    """ima2p_engine_set_pipeline only changes how the step is issued (chain groups on their own streams, several steps per
    graph, cross-step overlap): with the whole qupdate schedule running, every setting must visit exactly the same chain."""
    from support import engine_from_fixture, load_golden
    d = load_golden(name)
    outs = []
    settings = [(1, 1, 0), (2, 1, 0), (3, 4, 0), (3, 5, 1), (16, 8, 1)]
    for groups, depth, first in settings:
        eng, fm = engine_from_fixture(d, lib=lib, seed=77)
        eng.set_update_priors(t_max=[3.0] * fm.nsplit)
        eng.set_update_schedule(3, 5)
        eng.set_pipeline(groups, depth, first)
        eng.eval()
        eng.run(nsteps)
        eng.run(3)                                   # a second call: the step counter carried over correctly
        eng.sync()
        ch = [eng.chain(c) for c in range(eng.nchains)]
        outs.append((eng.counters(), eng.update_counters(),
                     np.concatenate([np.r_[c["probg"], c["pdg"], c["beta"], c["tvals"], c["wd"], c["wi"]] for c in ch])))
        eng.close()
    assert outs[0][0]["accepted"] > 0 and outs[0][0]["steps"] == nsteps + 3
    for k in range(1, len(settings)):
        assert outs[k][0] == outs[0][0], (settings[k], outs[k][0], outs[0][0])
        assert outs[k][1] == outs[0][1], (settings[k], outs[k][1], outs[0][1])
        assert np.array_equal(outs[k][2], outs[0][2]), settings[k]
    return outs[0][0]


def fast_path_equals_general_path(lib, name="state_sim50_hn3", nsteps=40, ppws=(0,)):
    """The two-kernel proposal path (lane-per-pair k_move + warp-per-pair k_weigh, ima_fastpath.h) and the general
    one-warp-per-pair kernel make the same moves from the same random streams: whichever path the pairs take -- all fast,
    all general, or a mixture because the fast kernels' tables are too small for some pairs -- the run is the same chain."""
    import os
    from support import engine_from_fixture, load_golden
    d = load_golden(name)
    outs = []
    # (fast, pairs per warp, environment): tiny tables send most pairs through the redo list
    settings = [(0, 0, {})] + [(1, w, {}) for w in ppws] + [(1, ppws[-1], {"IMA2P_FAST_POOL": "12", "IMA2P_FAST_EVENTS": "8"})]
    for fast, ppw, env in settings:
        os.environ.update(env)
        try:
            eng, fm = engine_from_fixture(d, lib=lib, seed=123)
        finally:
            for k in env:
                del os.environ[k]
        eng.set_update_priors(t_max=[3.0] * fm.nsplit)
        eng.set_update_schedule(3, 5)
        eng.set_proposal_path(fast, ppw)
        eng.eval()
        eng.run(nsteps)
        eng.sync()
        ch = [eng.chain(c) for c in range(eng.nchains)]
        trees = [eng.get_genealogy(c, l) for c in range(eng.nchains) for l in range(eng.nloci)]
        outs.append((eng.counters(), np.concatenate([np.r_[c["probg"], c["pdg"], c["beta"], c["tvals"], c["wd"], c["wi"]] for c in ch]),
                     np.concatenate([np.r_[t["time"], t["mig_t"], t["up0"], t["pop"], t["mig_p"]] for t in trees])))
        eng.close()
    assert outs[0][0]["accepted"] > 0
    for k in range(1, len(settings)):
        assert outs[k][0] == outs[0][0], (settings[k], outs[k][0], outs[0][0])
        assert np.array_equal(outs[k][1], outs[0][1]), settings[k]
        assert np.array_equal(outs[k][2], outs[0][2]), settings[k]
    return outs[0][0]


def hky_partials_stay_consistent(lib, name="state_sim5_hky_hn2", nsteps=400):
    """HKY loci keep the partial likelihoods of every internal node between steps (two slots per node, a slot mask per
    genealogy buffer) and a proposal recomputes only the nodes above the edges it touched (makefrac's rule,
    calc_prob_data.cpp:137-164).  After many whole steps -- genealogy proposals accepted and rejected, Rannala-Yang and
    Nielsen-Wakeley split-time updates, mutation-scalar and kappa updates -- the likelihood carried along must equal a
    from-scratch pruning of the same genealogies, on both the fast and the general proposal path."""
    from support import engine_from_fixture, load_golden, rel_close
    d = load_golden(name)
    out = {}
    for fast in (1, 0):
        eng, fm = engine_from_fixture(d, lib=lib, seed=4711)
        eng.set_update_priors(t_max=[3.0] * fm.nsplit)
        eng.set_update_schedule(3, 5)
        eng.set_proposal_path(fast, 0)
        eng.eval()
        for _ in range(4):
            eng.run(nsteps // 4)
            eng.sync()
            kept = [(eng.chain(c)["pdg"], [eng.pair(c, l)["pdg"] for l in range(eng.nloci)]) for c in range(eng.nchains)]
            eng.eval()                                  # prunes every genealogy from the tips again
            for c in range(eng.nchains):
                assert rel_close(eng.chain(c)["pdg"], kept[c][0], 1e-9), (fast, c, eng.chain(c)["pdg"], kept[c][0])
                for l in range(eng.nloci):
                    assert rel_close(eng.pair(c, l)["pdg"], kept[c][1][l], 1e-9), (fast, c, l)
        cnt, uc = eng.counters(), eng.update_counters()
        assert cnt["accepted"] > nsteps and uc["t_accepts"] > 0 and uc["u_accepts"] > 0, (cnt, uc)
        out[fast] = (cnt, uc)
        eng.close()
    return out


def run_summary(lib, name="state_sim50_hn3", nsteps=120, seed=5):
    """Per-chain P(G), P(D|G), split times and the counters after nsteps whole steps from a fixture's state (one flat array)."""
    from support import engine_from_fixture, load_golden
    d = load_golden(name)
    eng, fm = engine_from_fixture(d, lib=lib, seed=seed)
    eng.set_update_priors(t_max=[3.0] * fm.nsplit)
    eng.set_update_schedule(3, 5)
    eng.eval()
    eng.run(nsteps)
    eng.sync()
    tv, uv, _ = eng.fetch_parameters()
    cnt = eng.counters()
    out = np.concatenate([np.concatenate([np.r_[eng.chain(c)["probg"], eng.chain(c)["pdg"]] for c in range(eng.nchains)]), tv.reshape(-1), uv.reshape(-1),
                          np.array([cnt["accepted"], cnt["swaps"]], dtype=np.float64)])
    eng.close()
    return out


def scalar_walk_by_levels_equals_the_walk_in_order(lib, names=("state_sim5_hky_hn2", "state_sim3_sw_hn2"), nsteps=60):
    """k_changeu_levels evaluates the proposals of a mutation-scalar sweep that share no locus in parallel (levels).  With every
    proposal given a level of its own (IMA2P_CHANGEU_LEVELS_IN_ORDER: the reference's walk in order, same draws) the chains must
    be the same to the last bit -- likelihoods, scalars, kappas, counters."""
    import os
    from support import engine_from_fixture, load_golden
    for name in names:
        d = load_golden(name)
        got = []
        for in_order in (False, True):
            if in_order:
                os.environ["IMA2P_CHANGEU_LEVELS_IN_ORDER"] = "1"
            try:
                eng, fm = engine_from_fixture(d, lib=lib, seed=77)
                eng.set_update_priors(t_max=[3.0] * fm.nsplit)
                eng.set_update_schedule(3, 2)                        # the scalars every second step
                eng.eval()
                eng.run(nsteps)
                eng.sync()
                tv, uv, kp = eng.fetch_parameters()
                uc = eng.update_counters()
                got.append((np.concatenate([np.r_[eng.chain(c)["probg"], eng.chain(c)["pdg"]] for c in range(eng.nchains)]), uv.copy(), np.asarray(kp).copy(),
                            uc["u_tries"], uc["u_accepts"]))
                eng.close()
            finally:
                os.environ.pop("IMA2P_CHANGEU_LEVELS_IN_ORDER", None)
        a, b = got
        assert a[3] == b[3] > 0 and a[4] == b[4] > 0, (name, a[3:], b[3:])
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]), name


def capacity_grows_without_changing_the_chain(lib, name="state_sim3_hn3", nsteps=60):
    """ima2p_engine_grow_capacity (checkmig, utilities.cpp:1365-1383): growing the migration pools between two steps re-houses
    the resident genealogies and nothing else -- a run that grows half way is the run that had the room from the start.  With a
    capacity the genealogies outgrow, proposals are dropped and COUNTED (never silently): the count is what a caller reacts to."""
    from support import engine_from_fixture, load_golden
    d = load_golden(name)

    def run(cap0, grow_to):
        eng, fm = engine_from_fixture(d, lib=lib, seed=9, mig_capacity=cap0)
        eng.set_update_priors(t_max=[3.0] * fm.nsplit)
        eng.set_update_schedule(3, 5)
        eng.eval()
        eng.run(nsteps // 2)
        if grow_to:
            eng.grow_capacity(grow_to)
        eng.run(nsteps - nsteps // 2)
        eng.sync()
        ch = [eng.chain(c) for c in range(eng.nchains)]
        trees = [eng.get_genealogy(c, l) for c in range(eng.nchains) for l in range(eng.nloci)]
        out = (eng.counters(), np.concatenate([np.r_[c["probg"], c["pdg"], c["tvals"], c["wd"], c["wi"]] for c in ch]),
               np.concatenate([np.r_[t["time"], t["mig_t"], t["mig_p"]] for t in trees]))
        eng.close()
        return out
    a, b = run(64, 160), run(160, 0)
    assert a[0] == b[0] and a[0]["dropped"] == 0 and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    # a capacity below what the start state holds is refused outright; one the run outgrows is counted
    mx = max(len(sum(g["tree"]["mig"], [])) // 2 for ch in d["chains"] for g in ch["G"])
    tight = run(max(8, mx + 1), 0)
    assert tight[0]["dropped"] >= 0
    return a[0], tight[0]["dropped"]


def swap_walk_by_levels_equals_the_walk_in_order(lib, nchains=320, nloci=3, nsteps=40):
    """k_swap decides the attempts of a step level by level (attempts that share no temperature rank commute and are decided
    by the lanes together); with many chains and many attempts a step (swaptries = chains / 10) the temperatures, the swap
    counts and the chains must be those of the one-lane walk in attempt order (swapchains.cpp:192-523)."""
    import os
    from ima2p_b200 import Engine, synth
    loci = synth.make_dataset(nloci, 6, 6, seed=3)
    outs = []
    for seq in (True, False):
        if seq:
            os.environ["IMA2P_SWAP_SEQUENTIAL"] = "1"
        else:
            os.environ.pop("IMA2P_SWAP_SEQUENTIAL", None)
        try:
            eng = Engine(nchains, nloci, seed=12, lib=lib)
            eng.set_model(**synth.two_population_model(10.0, 1.0))
            for li, L in enumerate(loci):
                eng.set_locus(li, 0, L["n"], L["numsites"], L["samppop"], seq=L["seq"])
            eng.finalize()
            eng.set_heating(1, 0.99, 0.3)
            st = synth.initial_state(loci, nchains, eng.NL, eng.CAP, t0=1.5, seed=100)
            eng.put_state([st[k] for k in ("topo", "time", "mseg", "mig_t", "mig_p", "scal_i", "scal_d", "uvals")], st["tvals"])
            eng.run(nsteps)
            eng.sync()
            outs.append((eng.counters(), eng.betas().copy(), np.array([eng.chain(c)["pdg"] for c in range(nchains)])))
            eng.close()
        finally:
            os.environ.pop("IMA2P_SWAP_SEQUENTIAL", None)
    assert outs[0][0] == outs[1][0] and outs[0][0]["swaps"] > 10 * nsteps, outs[0][0]
    assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2])
    return outs[0][0]


def full_size_workload_properties(lib, nloci, nchains, nsteps, noracle=48, seed=5):
    """BASELINE-sized runs (configs[1]: 50 loci x 128 chains; configs[2]'s per-GPU shard: 300 loci x 256 chains), checked through
    properties that do not need a stored answer: after `nsteps` whole qupdate steps (genealogies, split times, scalars, swaps)
      * every genealogy still has exactly n - 1 coalescences (integer counts summed over the chain: bit-exact),
      * the sums the accept sweep maintains incrementally equal a from-scratch evaluation (ints exact, doubles 1e-9),
      * the temperatures are a permutation of the ladder,
      * for a random sample of (chain, locus) pairs the oracle's treeweight / infinite-sites likelihood of the genealogy
        downloaded from the device agree with the device's own values (counts exact, doubles 1e-9)."""
    from ima2p_b200 import Engine, synth
    from support import OracleModel, tree_from_engine
    n0 = n1 = 15
    loci = synth.make_dataset(nloci, n0, n1, seed=11)
    eng = Engine(nchains, nloci, mig_capacity=64, seed=seed, lib=lib)
    eng.set_model(**synth.two_population_model(10.0, 1.0))
    for li, L in enumerate(loci):
        eng.set_locus(li, 0, L["n"], L["numsites"], L["samppop"], seq=L["seq"])
    eng.finalize()
    eng.set_heating(1, 0.96, 0.9)                     # HEAT_GEOMETRIC (-hfg -ha 0.96 -hb 0.9), as bench.py
    st = synth.initial_state(loci, nchains, eng.NL, eng.CAP, t0=1.5, seed=100)
    arrs = [np.ascontiguousarray(st[k]) for k in ["topo", "time", "mseg", "mig_t", "mig_p", "scal_i", "scal_d", "uvals"]]
    eng.put_state(arrs, st["tvals"])
    eng.sync()
    betas0 = sorted(eng.betas())
    eng.set_update_priors(t_max=[3.0])
    eng.set_update_schedule(3, 5)
    eng.run(nsteps)
    eng.sync()
    cnt = eng.counters()
    assert cnt["steps"] == nsteps and cnt["updates"] == nsteps * nchains * nloci and cnt["dropped"] == 0
    assert 0.15 < cnt["accepted"] / cnt["updates"] < 0.7
    assert rel_close(sorted(eng.betas()), betas0, 0.0)
    fm = FlatModel(load_golden("state_sim5_hn4")["model"])        # the same 2-population -q10 -m1 model, as the oracle takes it
    inc = [eng.chain(c) for c in range(nchains)]
    ncoal = nloci * (n0 + n1 - 1)
    for c in range(nchains):
        assert int(np.sum(inc[c]["wi"][:fm.ncc])) == ncoal, c
    rng = np.random.default_rng(seed)
    sample = [(int(rng.integers(nchains)), int(rng.integers(nloci))) for _ in range(noracle)]
    om = OracleModel(fm)
    for c, li in sample:
        t = tree_from_engine(eng.get_genealogy(c, li))
        r = eng.pair(c, li)
        loc = dict(samppop=loci[li]["samppop"], hval=1.0, seq=loci[li]["seq"], numsites=loci[li]["numsites"], sumlogk=0.0)
        w = om.treeweight(inc[c]["tvals"], loc, t)
        ew = split_weights(fm, r["wi"], r["wd"])
        assert np.array_equal(ew["cc"], w["cc"]) and np.array_equal(ew["mc"], w["mc"]) and r["mignum"] == w["mignum"]
        assert rel_close(ew["fc"], w["fc"], 1e-9, 1e-12) and rel_close(ew["fm"], w["fm"], 1e-9, 1e-12)
        assert rel_close(r["length"], w["length"], 1e-9)
        u = eng.scalars(c, li)[0][0]
        assert rel_close(r["pdg"], om.likelihood_is(loc, t, w["length"], u), 1e-9), (c, li)
    eng.eval()
    for c in range(nchains):
        fresh = eng.chain(c)
        assert np.array_equal(fresh["wi"], inc[c]["wi"])
        assert rel_close(fresh["wd"], inc[c]["wd"], 1e-9, 1e-9)
        assert rel_close(fresh["probg"], inc[c]["probg"], 1e-9) and rel_close(fresh["pdg"], inc[c]["pdg"], 1e-9)
    eng.close()
    return cnt


def packed_upload_equals_plain_upload(lib, nloci=10, nchains=5, nsteps=60):
    """ima2p_engine_put_state_packed / put_state_block (8-bit wire forms, widened on the device) load exactly the state put_state loads: the
    same evaluation, and the same run afterwards.  The state is one with migration events (taken after some steps)."""
    from ima2p_b200 import Engine, synth
    keys = ["topo", "time", "mseg", "mig_t", "mig_p", "scal_i", "scal_d", "uvals"]

    def make():
        loci = synth.make_dataset(nloci, 15, 15, seed=11)
        eng = Engine(nchains, nloci, mig_capacity=64, seed=9, lib=lib)
        eng.set_model(**synth.two_population_model(10.0, 1.0))
        for li, L in enumerate(loci):
            eng.set_locus(li, 0, L["n"], L["numsites"], L["samppop"], seq=L["seq"])
        eng.finalize()
        eng.set_heating(1, 0.96, 0.9)
        eng.set_update_priors(t_max=[3.0])
        eng.set_update_schedule(3, 5)
        return eng, loci
    eng, loci = make()
    st = synth.initial_state(loci, nchains, eng.NL, eng.CAP, t0=1.5, seed=100)
    arrs = [np.ascontiguousarray(st[k]) for k in keys]
    eng.put_state(arrs, st["tvals"])
    eng.run(nsteps)
    eng.sync()
    eng.fetch_state(arrs[:7])
    tv, uv, _ = eng.fetch_parameters()
    arrs[7][...] = uv.reshape(arrs[7].shape)
    assert arrs[5].reshape(-1, 2)[:, 1].max() > 0                 # there are migration events to carry
    packed = Engine.pack_state(arrs[0].reshape(nchains * nloci, eng.NL, 4), arrs[2].reshape(nchains * nloci, eng.NL, 2))
    assert packed is not None
    outs = []
    for mode in ("plain", "packed", "block", "two halves"):
        e2, _ = make()
        if mode == "plain":
            e2.put_state(arrs, tv)
        elif mode == "packed":
            e2.put_state_packed([packed[0], arrs[1], packed[1]] + arrs[3:], tv)
        elif mode == "block":
            blk, events = e2.pack_state_block(arrs, tv)
            assert events == int(arrs[5].reshape(-1, 2)[:, 1].sum())
            e2.put_state_block(blk, events)
        else:
            # upload_block + adopt_block with both staging slots in use: a block of another state first (the start state), adopted
            # and stepped; the wanted block travels meanwhile and is adopted after it.  A third upload before anything is adopted
            # is refused.
            blk0, ev0 = e2.pack_state_block([np.ascontiguousarray(st[k]) for k in keys], st["tvals"])
            blk, events = e2.pack_state_block(arrs, tv)
            e2.upload_block(blk0, ev0)
            e2.upload_block(blk, events)
            try:
                e2.upload_block(blk, events)
                raise AssertionError("a third pending upload was accepted")
            except Exception as ex:
                assert "staging" in str(ex)
            e2.adopt_block()
            e2.run(3)
            e2.adopt_block()
            try:
                e2.adopt_block()
                raise AssertionError("adopt_block without a pending upload was accepted")
            except Exception as ex:
                assert "no uploaded block" in str(ex)

        e2.sync()
        ev = [e2.chain(c) for c in range(nchains)]
        e2.run(20)
        e2.sync()
        after = [e2.chain(c) for c in range(nchains)]
        outs.append(np.concatenate([np.r_[c["probg"], c["pdg"], c["wd"], c["wi"]] for c in ev + after]))
        back = [np.zeros_like(a) for a in arrs[:7]]
        e2.fetch_state(back)
        outs.append(np.concatenate([b.reshape(-1).astype(np.float64) for b in back[:3]]))
        e2.close()
    assert np.array_equal(outs[0], outs[2]) and np.array_equal(outs[1], outs[3])
    assert np.array_equal(outs[0], outs[4]) and np.array_equal(outs[1], outs[5])
    # the two-halves engine ran three steps before the wanted state arrived: its step counter (which keys the draws) differs,
    # so only the evaluation on arrival is comparable
    n_ev = len(outs[0]) // 2
    assert np.array_equal(outs[0][:n_ev], outs[6][:n_ev]) and not np.array_equal(outs[0][n_ev:], outs[6][n_ev:])
    eng.close()


def lmode_f3_matches_oracle_on_bootstrapped_rows(lib, nrows, name="lmode_extra_sim5_hn2", seed=3, rtol=1e-9):
    """section 8 (f3) on more rows than the reference fixtures hold: rows drawn with replacement (seeded) from a fixture's
    rows, the device against the oracle on the same rows -- moment sums, the 2NM density row sums on a grid, and the
    greater-than probabilities (with more than 20,000 rows that includes the reference's row thinning, gtint.cpp:341-351)."""
    from ima2p_b200 import LMode
    from support import OracleModel, oracle, fp, dp
    d = load_golden(name)
    fm = FlatModel(d["model"])
    src = np.ascontiguousarray(d["rows"], dtype=np.float32)
    rng = np.random.default_rng(seed)
    rows = np.ascontiguousarray(src[rng.integers(0, len(src), size=nrows)])
    G, rl = rows.shape
    n = fm.nq + fm.nm
    lm = LMode(fm.nq, fm.nm, fm.nsplit, fm.q_max, fm.q_min, fm.m_max, fm.m_min, fm.m_mean, fm.expoprior, lib=lib)
    lm.load(rows)
    om, ora = OracleModel(fm), oracle()
    sums = np.zeros(2 * n + n * n)
    ora.ora_moment_sums(om.h, fp(rows), rl, G, dp(sums))
    assert rel_close(lm.moments_raw(), sums, rtol)
    x = fm.q_max[0] * fm.m_max[0] / 2.0 * (np.arange(9) + 0.5) / 9.0
    for ti, mi in ((0, 0), (fm.nq - 1, fm.nm - 1)):
        ref = np.array([ora.ora_popmig_sum(om.h, fp(rows), rl, 0, G, ti, mi, float(v)) for v in x])
        assert rel_close(lm.popmig_sums(ti, mi, x), ref, rtol, 1e-300)
        half = np.array([ora.ora_popmig_sum(om.h, fp(rows), rl, G // 4, G // 2, ti, mi, float(v)) for v in x])
        assert rel_close(lm.popmig_sums(ti, mi, x, G // 4, G // 2), half, rtol, 1e-300)
    for kind, i, j in ((0, 0, 1), (0, 2, 0), (1, 0, 1), (1, 1, 0)):
        assert rel_close(lm.greater_than(kind, i, j), ora.ora_greater_than(om.h, fp(rows), rl, G, kind, i, j), max(rtol, 1e-8)), (kind, i, j)
    lm.close()
