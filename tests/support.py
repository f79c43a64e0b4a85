"""Test support: golden-fixture loading, flat model/tree arrays and ctypes bindings to the CPU oracle.

TEST INFRASTRUCTURE ONLY.  Nothing here is imported by the ima2p_b200 package.  The oracle
(oracle/liboracle.so) is the checker for the CUDA path; it is never the thing measured or shipped.
"""
import ctypes as C
import gzip
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_DIR = os.path.join(ROOT, "oracle")

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)
c_flt_p = C.POINTER(C.c_float)


def load_golden(name):
    with gzip.open(os.path.join(GOLDEN, name + ".json.gz"), "rt") as f:
        d = json.load(f, parse_constant=float)
    m = d.get("model") if isinstance(d, dict) else None
    if m and m.get("nomigration"):
        # -m 0: the reference keeps no migration weights at all; here they exist and stay zero
        nmc = sum((m["npops"] - k) ** 2 for k in range(m["numsplittimes"]))

        def fill(x):
            if isinstance(x, dict):
                if "cc" in x and "mc" in x and not x["mc"]:
                    x["mc"], x["fm"] = [0] * nmc, [0.0] * nmc
                for v in x.values():
                    fill(v)
            elif isinstance(x, list):
                for v in x:
                    fill(v)
        fill(d)
    return d


def _num(v):
    if isinstance(v, str):
        return {"inf": np.inf, "-inf": -np.inf, "nan": np.nan}[v]
    return v


def ip(a):
    return a.ctypes.data_as(c_int_p)


def dp(a):
    return a.ctypes.data_as(c_dbl_p)


def fp(a):
    return a.ctypes.data_as(c_flt_p)


def i32(x):
    return np.ascontiguousarray(x, dtype=np.int32)


def f64(x):
    return np.ascontiguousarray(x, dtype=np.float64)


class FlatModel:
    """Chain-invariant model tables in the flat form both the oracle and the C-ABI take.

    Built from the "model" object of a golden fixture (itself a dump of the reference's poptree,
    plist, itheta/imig weight-position lists and priors, initialize.cpp:201-727)."""

    def __init__(self, mj):
        self.npops = mj["npops"]
        self.nsplit = mj["numsplittimes"]
        self.ntreepops = mj["numtreepops"]
        self.nq = mj["numpopsizeparams"]
        self.nm = mj["nummigrateparams"]
        self.rootpop = mj["rootpop"]
        self.nomigration = mj["nomigration"]
        self.expoprior = mj["expomigrationprior"]
        self.thermo = mj["calcmarginallikelihood"]
        self.gbeta = float(mj["gbeta"])
        self.rowlen = mj["gsampinflength"]
        n = self.npops
        pl = -np.ones((n, n), dtype=np.int32)
        for k, row in enumerate(mj["plist"]):
            pl[k, :len(row)] = row
        self.plist = pl
        self.addpop = i32(mj["addpop"])
        self.droppops = i32(mj["droppops"]).reshape(-1)
        self.pt_e = i32([p["e"] for p in mj["poptree"]])
        self.pt_down = i32([p["down"] for p in mj["poptree"]])
        self.q_off = i32(np.cumsum([0] + [len(t["p"]) for t in mj["itheta"]]))
        self.q_p = i32(sum((t["p"] for t in mj["itheta"]), []))
        self.q_r = i32(sum((t["r"] for t in mj["itheta"]), []))
        self.q_max = f64([t["max"] for t in mj["itheta"]])
        self.q_min = f64([t["min"] for t in mj["itheta"]])
        self.m_off = i32(np.cumsum([0] + [len(t["p"]) for t in mj["imig"]]))
        self.m_p = i32(sum((t["p"] for t in mj["imig"]), []))
        self.m_r = i32(sum((t["r"] for t in mj["imig"]), []))
        self.m_c = i32(sum((t["c"] for t in mj["imig"]), []))
        self.m_max = f64([t["max"] for t in mj["imig"]])
        self.m_min = f64([t["min"] for t in mj["imig"]])
        self.m_mean = f64([t["mean"] for t in mj["imig"]])
        nc = mj["nomigrationchecklist"]
        self.nomig_p, self.nomig_r, self.nomig_c = i32(nc["p"]), i32(nc["r"]), i32(nc["c"])
        self.ncc = sum(n - k for k in range(self.nsplit + 1))
        self.nmc = sum((n - k) ** 2 for k in range(self.nsplit))

    def create_args(self):
        """Argument tuple shared by ora_model_create and ima2p_engine_set_model."""
        return (self.npops, self.nsplit, ip(self.plist), ip(self.addpop), ip(self.droppops), ip(self.pt_e),
                ip(self.pt_down), self.rootpop, self.nq, ip(self.q_off), ip(self.q_p), ip(self.q_r), dp(self.q_max),
                dp(self.q_min), self.nm, ip(self.m_off), ip(self.m_p), ip(self.m_r), ip(self.m_c), dp(self.m_max),
                dp(self.m_min), dp(self.m_mean), len(self.nomig_p), ip(self.nomig_p), ip(self.nomig_r),
                ip(self.nomig_c), self.nomigration, self.expoprior, self.thermo, C.c_double(self.gbeta))


class FlatTree:
    """One genealogy as SoA arrays + CSR migration lists (from a fixture "tree" object)."""

    def __init__(self, tj):
        self.root = tj["root"]
        self.roottime = float(tj["roottime"])
        self.up0, self.up1 = i32(tj["up0"]), i32(tj["up1"])
        self.down, self.pop = i32(tj["down"]), i32(tj["pop"])
        self.time = f64(tj["time"])
        self.numlines = len(self.up0)
        self.numgenes = (self.numlines + 1) // 2
        off, mt, mp = [0], [], []
        for lst in tj["mig"]:
            mt += lst[0::2]
            mp += lst[1::2]
            off.append(len(mt))
        self.mig_off, self.mig_t, self.mig_p = i32(off), f64(mt + [0.0]), i32(mp + [0])
        self.A = [i32(a) for a in tj.get("A", [])]
        self.dlikeA = [f64(a) for a in tj.get("dlikeA", [])]

    def args(self):
        return (ip(self.up0), ip(self.up1), ip(self.down), ip(self.pop), dp(self.time), ip(self.mig_off),
                dp(self.mig_t), ip(self.mig_p))


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "oracle"], check=True)
    return os.path.join(ORACLE_DIR, "liboracle.so")


_ORACLE = None


def oracle():
    """ctypes handle to the CPU oracle with argument/return types set."""
    global _ORACLE
    if _ORACLE is not None:
        return _ORACLE
    lib = C.CDLL(build_oracle())
    d, i, v = C.c_double, C.c_int, C.c_void_p
    lib.ora_model_create.restype = v
    lib.ora_model_create.argtypes = [i, i, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p, i, i, c_int_p, c_int_p, c_int_p,
                                     c_dbl_p, c_dbl_p, i, c_int_p, c_int_p, c_int_p, c_int_p, c_dbl_p, c_dbl_p, c_dbl_p,
                                     i, c_int_p, c_int_p, c_int_p, i, i, i, d]
    lib.ora_model_destroy.argtypes = [v]
    for name, args in {
        "ora_logfact": [i], "ora_uppergamma": [i, d], "ora_lowergamma": [i, d], "ora_bessi": [i, d],
        "ora_mylogcosh": [d], "ora_mylogsinh": [d], "ora_calcmrate": [i, d],
        "ora_integrate_coalescent_term": [i, d, d, d, d], "ora_integrate_migration_term": [i, d, d, d],
        "ora_integrate_migration_term_expo_prior": [i, d, d], "ora_swapweight": [d, d, d, d],
        "ora_slideweight": [d, d, d],
    }.items():
        getattr(lib, name).restype = d
        getattr(lib, name).argtypes = args
    lib.ora_eexp.argtypes = [d, c_dbl_p, c_int_p]
    lib.ora_initialize_integrate_tree_prob.restype = d
    lib.ora_initialize_integrate_tree_prob.argtypes = [v, c_int_p, c_dbl_p, c_dbl_p, c_int_p, c_dbl_p, c_dbl_p, c_dbl_p]
    lib.ora_integrate_tree_prob.restype = d
    lib.ora_integrate_tree_prob.argtypes = [v, c_int_p, c_dbl_p, c_dbl_p, c_int_p, c_dbl_p, c_int_p, c_dbl_p, c_int_p,
                                            c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p]
    lib.ora_sum_subtract_treeinfo.argtypes = [v] + [c_int_p, c_dbl_p, c_dbl_p, c_int_p, c_dbl_p] * 3
    lib.ora_treeweight.restype = i
    lib.ora_treeweight.argtypes = [v, c_dbl_p, i, c_int_p, d, c_int_p, c_int_p, c_int_p, c_int_p, c_dbl_p, c_int_p,
                                   c_dbl_p, c_int_p, i, d, c_int_p, c_dbl_p, c_dbl_p, c_int_p, c_dbl_p, c_dbl_p]
    lib.ora_calc_sumlogk.restype = d
    lib.ora_calc_sumlogk.argtypes = [i, i, c_int_p, c_int_p, c_int_p, c_int_p]
    lib.ora_likelihoodIS.restype = d
    lib.ora_likelihoodIS.argtypes = [i, i, c_int_p, c_int_p, c_int_p, c_int_p, c_dbl_p, d, d, d]
    lib.ora_likelihoodHKY.restype = d
    lib.ora_likelihoodHKY.argtypes = [i, i, i, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p, c_dbl_p, i, c_dbl_p, d, d]
    lib.ora_likelihoodSW.restype = d
    lib.ora_likelihoodSW.argtypes = [i, c_int_p, c_int_p, c_dbl_p, c_int_p, d, d, c_dbl_p]
    tree = [c_int_p, c_int_p, c_int_p, c_int_p, c_dbl_p, c_int_p, c_dbl_p, c_int_p, i]
    lib.ora_migration_proposal_logprobs.restype = i
    lib.ora_migration_proposal_logprobs.argtypes = [v, c_dbl_p, i] + tree + tree + [i, c_dbl_p]
    lib.ora_setheat.argtypes = [i, d, d, i, c_dbl_p]
    lib.ora_savegsampinf.argtypes = [v, c_int_p, c_dbl_p, c_dbl_p, c_int_p, c_dbl_p, c_dbl_p, c_dbl_p, d, d, c_dbl_p,
                                     c_flt_p]
    lib.ora_marginp.restype = d
    lib.ora_marginp.argtypes = [v, c_flt_p, i, i, i, i, d]
    lib.ora_margincalc.restype = d
    lib.ora_margincalc.argtypes = [v, c_flt_p, i, i, d, d, i, i]
    lib.ora_jointp.restype = d
    lib.ora_jointp.argtypes = [v, c_flt_p, i, i, c_dbl_p, i, c_dbl_p]
    lib.ora_jointp_model.restype = d
    lib.ora_jointp_model.argtypes = [v, c_flt_p, i, i, c_dbl_p, i, i, c_dbl_p]
    lib.ora_calcx.restype = d
    lib.ora_calcx.argtypes = [v, c_flt_p, i, i, i, i]
    lib.ora_moment_sums.argtypes = [v, c_flt_p, i, i, c_dbl_p]
    lib.ora_popmig_sum.restype = d
    lib.ora_popmig_sum.argtypes = [v, c_flt_p, i, i, i, i, i, d]
    lib.ora_greater_than.restype = d
    lib.ora_greater_than.argtypes = [v, c_flt_p, i, i, i, i, i]
    lib.ora_getnewt.restype = d
    lib.ora_getnewt.argtypes = [d, i, i, i, d, d, d]
    lib.ora_ry1_rescale.argtypes = [i, c_int_p, c_dbl_p, i, c_dbl_p, c_dbl_p, i, i, d, d, d, d, c_int_p]
    lib.ora_ry1_hastings.restype = d
    lib.ora_ry1_hastings.argtypes = [i, i, d, d, d, d, i, i, i, i]
    lib.ora_nw_migweight.restype = d
    lib.ora_nw_migweight.argtypes = [v, c_dbl_p, i, d, i, i, c_int_p, c_int_p, c_int_p, c_dbl_p, c_int_p, c_int_p, c_dbl_p,
                                     c_int_p, c_int_p, c_int_p, c_dbl_p, c_int_p]
    lib.ora_changeu_newr.restype = d
    lib.ora_changeu_newr.argtypes = [d, d, d, d, c_dbl_p]
    lib.ora_new_kappa.restype = d
    lib.ora_new_kappa.argtypes = [d, d, d, d]
    lib.ora_thermomarginlike.restype = d
    lib.ora_thermomarginlike.argtypes = [c_dbl_p, i, i]
    lib.ora_last_error.restype = i
    _ORACLE = lib
    return lib


class OracleModel:
    def __init__(self, fm):
        self.fm = fm
        self.lib = oracle()
        self.h = self.lib.ora_model_create(*fm.create_args())
        assert self.h, "ora_model_create failed"

    def __del__(self):
        try:
            self.lib.ora_model_destroy(self.h)
        except Exception:
            pass

    # -- per-pair evaluations ------------------------------------------------------------------
    def treeweight(self, tvals, locus, tree):
        fm = self.fm
        cc, mc = np.zeros(fm.ncc, np.int32), np.zeros(max(fm.nmc, 1), np.int32)
        fc, hcc, fmw = np.zeros(fm.ncc), np.zeros(fm.ncc), np.zeros(max(fm.nmc, 1))
        out = np.zeros(2)
        tv = f64(tvals)
        mignum = self.lib.ora_treeweight(self.h, dp(tv), tree.numgenes, ip(i32(locus["samppop"])), float(locus["hval"]),
                                         *tree.args(), tree.root, tree.roottime, ip(cc), dp(fc), dp(hcc), ip(mc),
                                         dp(fmw), dp(out))
        return dict(mignum=mignum, cc=cc, fc=fc, hcc=hcc, mc=mc[:fm.nmc], fm=fmw[:fm.nmc], length=out[0], tlength=out[1])

    def likelihood_is(self, locus, tree, length, u, sumlogk=None):
        seq = i32(locus["seq"])
        if sumlogk is None:
            sumlogk = float(locus["sumlogk"])
        return self.lib.ora_likelihoodIS(tree.numgenes, locus["numsites"], ip(seq), ip(tree.up0), ip(tree.up1),
                                         ip(tree.down), dp(tree.time), length, u, sumlogk)

    def likelihood_hky(self, locus, tree, pi, u, kappa):
        seq = i32(locus["seq"])
        return self.lib.ora_likelihoodHKY(tree.numgenes, locus["numsites"], locus["totsites"], ip(seq),
                                          ip(i32(locus["mult"])), ip(tree.up0), ip(tree.up1), ip(tree.down),
                                          dp(tree.time), tree.root, dp(f64(pi)), u, kappa)

    def likelihood_sw(self, tree, ai, u):
        dl = np.zeros(tree.numlines)
        like = self.lib.ora_likelihoodSW(tree.numgenes, ip(tree.up0), ip(tree.down), dp(tree.time), ip(tree.A[ai]), u, 1.0,
                                         dp(dl))
        return like, dl

    def init_integrate(self, w):
        qint, mint = np.zeros(max(self.fm.nq, 1)), np.zeros(max(self.fm.nm, 1))
        mc, fmw = i32(np.resize(w["mc"], max(self.fm.nmc, 1))), f64(np.resize(w["fm"], max(self.fm.nmc, 1)))
        probg = self.lib.ora_initialize_integrate_tree_prob(self.h, ip(i32(w["cc"])), dp(f64(w["fc"])), dp(f64(w["hcc"])),
                                                            ip(mc), dp(fmw), dp(qint), dp(mint))
        return probg, qint[:self.fm.nq], mint[:self.fm.nm]

    def nw_migweight(self, tvals, period, newt, before, after):
        """log Hastings ratio of the migration events of one changet_NW() move (oracle replay of update_mig_tNW)."""
        tv = f64(tvals)
        return self.lib.ora_nw_migweight(self.h, dp(tv), period, newt, before.numgenes, before.root, ip(before.up0), ip(before.up1),
                                         ip(before.down), dp(before.time), ip(before.pop), ip(before.mig_off), dp(before.mig_t),
                                         ip(before.mig_p), ip(after.pop), ip(after.mig_off), dp(after.mig_t), ip(after.mig_p))

    def migration_logprobs(self, tvals, before, after, edge):
        out = np.zeros(2)
        tv = f64(tvals)
        self.lib.ora_migration_proposal_logprobs(self.h, dp(tv), before.numgenes, *before.args(), before.root,
                                                 *after.args(), after.root, edge, dp(out))
        return out[0], out[1]


def weights_from_json(wj):
    return dict(cc=i32(wj["cc"]), fc=f64(wj["fc"]), hcc=f64(wj["hcc"]), mc=i32(wj["mc"]), fm=f64(wj["fm"]))


def rel_close(a, b, rtol, atol=0.0):
    a, b = np.atleast_1d(np.asarray(a, dtype=np.float64)), np.atleast_1d(np.asarray(b, dtype=np.float64))
    same = (a == b)                      # also covers equal infinities
    fin = np.isfinite(a) & np.isfinite(b)
    ok = np.zeros(a.shape, dtype=bool)
    ok[fin] = np.abs(a[fin] - b[fin]) <= atol + rtol * np.maximum(np.abs(a[fin]), np.abs(b[fin]))
    return bool(np.all(same | ok))


# ---- engine loading helpers (shared by the hostemu CPU tests and the GPU parity tests) -------------------
def engine_from_fixture(d, lib=None, chains=None, mig_capacity=64, seed=1234, betas=None):
    """Build an ima2p_b200.Engine holding the chains of a "state" fixture (all loci)."""
    from ima2p_b200 import Engine
    fm = FlatModel(d["model"])
    chains = list(range(len(d["chains"]))) if chains is None else chains
    nloci = len(d["loci"])
    eng = Engine(len(chains), nloci, mig_capacity=mig_capacity, seed=seed, lib=lib)
    eng.set_model_flat(*fm.create_args())
    for li, loc in enumerate(d["loci"]):
        eng.set_locus(li, loc["model"], loc["numgenes"], loc["numsites"], loc["samppop"],
                      seq=loc["seq"] if loc["seq"] else None, mult=loc.get("mult"), hval=loc["hval"],
                      totsites=loc["totsites"], nlinked=loc["nlinked"], minA=loc["minA"], maxA=loc["maxA"],
                      sumlogk=loc["sumlogk"])
    eng.finalize()
    eng.set_debug_records(True)                 # the checks read the per-proposal record (Engine.proposal)
    eng.set_betas(betas if betas is not None else [d["chains"][c]["beta"] for c in chains])
    for k, c in enumerate(chains):
        ch = d["chains"][c]
        eng.set_chain(k, ch["tvals"])
        for li, g in enumerate(ch["G"]):
            t = FlatTree(g["tree"])
            A = np.stack(t.A) if t.A else None
            eng.set_genealogy(k, li, t.up0, t.up1, t.down, t.pop, t.time, t.mig_off, t.mig_t[:-1], t.mig_p[:-1], t.root,
                              t.roottime, uvals=g["uvals"], kappa=g["kappa"], pi=g["pi"], A=A)
    eng.upload()
    return eng, fm


def tree_from_engine(g):
    """Engine.get_genealogy() dict -> FlatTree-like object for the oracle."""
    t = FlatTree.__new__(FlatTree)
    t.root, t.roottime = g["root"], g["roottime"]
    t.up0, t.up1, t.down, t.pop = i32(g["up0"]), i32(g["up1"]), i32(g["down"]), i32(g["pop"])
    t.time = f64(g["time"])
    t.numlines = len(t.up0)
    t.numgenes = (t.numlines + 1) // 2
    t.mig_off = i32(g["mig_off"])
    t.mig_t, t.mig_p = f64(np.append(g["mig_t"], 0.0)), i32(np.append(g["mig_p"], 0))
    t.A, t.dlikeA = [], []
    return t


def split_weights(fm, wi, wd):
    """Engine weight records -> dict(cc, mc, fc, hcc, fm) in fixture order."""
    return dict(cc=wi[:fm.ncc], mc=wi[fm.ncc:], fc=wd[:fm.ncc], hcc=wd[fm.ncc:2 * fm.ncc], fm=wd[2 * fm.ncc:])


def check_static_eval(eng, fm, d, chains=None, rtol=1e-9):
    """Engine.eval() results against the fixture's values (the reference's init_p): ints exact, doubles rtol."""
    chains = list(range(len(d["chains"]))) if chains is None else chains
    for k, c in enumerate(chains):
        ch = d["chains"][c]
        for li, g in enumerate(ch["G"]):
            r = eng.pair(k, li)
            w = split_weights(fm, r["wi"], r["wd"])
            e = g["gweight"]
            assert r["mignum"] == g["mignum"], (c, li)
            assert np.array_equal(w["cc"], e["cc"]) and np.array_equal(w["mc"], e["mc"]), (c, li)
            assert rel_close(w["fc"], e["fc"], rtol) and rel_close(w["fm"], e["fm"], rtol), (c, li)
            assert rel_close(w["hcc"], e["hcc"], rtol, 1e-300), (c, li)
            assert rel_close(r["length"], g["length"], rtol) and rel_close(r["tlength"], g["tlength"], rtol), (c, li)
            assert rel_close(r["pdg"], g["pdg"], rtol), (c, li, r["pdg"], g["pdg"])
        cr = eng.chain(k)
        wa = split_weights(fm, cr["wi"], cr["wd"])
        ea = ch["allgweight"]
        assert np.array_equal(wa["cc"], ea["cc"]) and np.array_equal(wa["mc"], ea["mc"])
        assert rel_close(wa["fc"], ea["fc"], rtol) and rel_close(wa["fm"], ea["fm"], rtol)
        assert rel_close(cr["qintegrate"], ch["qintegrate"], rtol), (cr["qintegrate"], ch["qintegrate"])
        assert rel_close(cr["mintegrate"], ch["mintegrate"], rtol)
        assert rel_close(cr["probg"], ch["probg"], rtol) and rel_close(cr["pdg"], ch["pdg"], rtol)


# ---- split-time (Rannala-Yang) and mutation-scalar updates: shared pieces of the oracle-side restatement -------------
TIMEMAX = 1000000.0


def ry1_bounds(tvals, period, nsplit):
    """t_u, t_d of changet_RY1 (update_t_RY.cpp:242-246)."""
    t_u = 0.0 if period == 0 else tvals[period - 1]
    t_d = TIMEMAX if period == nsplit - 1 else tvals[period + 1]
    return t_u, t_d


def ry1_rescale_tree(tree, period, nsplit, oldt, newt, t_u, t_d, counts):
    """In-place oracle rescaling of one FlatTree; counts (int32[4]) accumulates."""
    rt = C.c_double(tree.roottime)
    nmig = int(tree.mig_off[-1])
    oracle().ora_ry1_rescale(tree.numlines, ip(tree.down), dp(tree.time), nmig, dp(tree.mig_t), C.byref(rt), period, nsplit,
                             oldt, newt, t_u, t_d, ip(counts))
    tree.roottime = rt.value


def changeu_replay(U, j, nurates):
    """The draws of one changeu() call in order (update_mc_params.cpp:78-90, 201-212): returns k and the ratio draw."""
    it = iter(U)
    if nurates > 2:
        while True:
            k = int(next(it) * nurates)
            if k != j and 0 <= k < nurates:
                break
    else:
        k = 1
    return k, next(it), list(it)
