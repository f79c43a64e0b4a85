"""Host-side mirror of the reference's function seam for the hot path.

The reference drives its hot path through free functions over process-global state:
``updategenealogy(ci, li)``, ``treeweight(ci, li)``, ``integrate_tree_prob``, ``likelihoodIS``,
``swapchains``, ``savegsampinf`` (imamp.hpp:1167-1322) and, in L mode, ``margincalc`` / ``marginp`` /
``jointp`` (imamp.hpp:1242-1248, 1420).  :class:`Engine` and :class:`LMode` expose the same operations,
batched over all chains x loci, on top of the C ABI (include/ima2p_b200.h).  Nothing here computes:
every method is one or two calls into the CUDA library.
"""
import ctypes as C

import numpy as np

from . import capi

HEAT_LINEAR, HEAT_GEOMETRIC, HEAT_EVEN = 0, 1, 2     # HLINEAR / HGEOMETRIC / HEVEN, swapchains.cpp:94-110
MODEL_IS, MODEL_HKY, MODEL_SW = 0, 1, 2
MAXLINKED = 4


def _ip(a):
    return a.ctypes.data_as(capi.c_int_p)


def _dp(a):
    return a.ctypes.data_as(capi.c_dbl_p)


def _i32(x):
    return np.ascontiguousarray(x, dtype=np.int32)


def _f64(x):
    return np.ascontiguousarray(x, dtype=np.float64)


class Engine:
    """All Metropolis-coupled chains x loci of one GPU, resident in HBM.

    ``nchains`` chains live on this GPU; with several GPUs ``nchains_global`` is the total and ``chain0``
    the global index of local chain 0 (chains shard by rank exactly like the reference's ``-hn`` chains
    per MPI process, README.md:106).
    """

    def __init__(self, nchains, nloci, mig_capacity=64, seed=1, device=0, nchains_global=None, chain0=0, lib=None):
        self.lib = lib if lib is not None else capi.lib()
        self.nchains, self.nloci = nchains, nloci
        self.nchains_global = nchains if nchains_global is None else nchains_global
        self.chain0 = chain0
        self._h = C.c_void_p()
        self._ck(self.lib.ima2p_engine_create(C.byref(self._h), device, nchains, self.nchains_global, chain0, nloci,
                                              mig_capacity, seed))
        self._loci = {}
        self.nsplit = None

    def _ck(self, rc):
        capi.check(self.lib, rc)

    def close(self):
        if self._h:
            self.lib.ima2p_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- set-up (setup(), initialize.cpp:2074) -------------------------------------------------------------
    def set_model(self, npops, nsplit, plist, addpop, droppops, pt_e, pt_down, rootpop, q_wp, q_max, q_min, m_wp,
                  m_max, m_min, m_mean=None, nomig=(), nomigration=0, expoprior=0, thermo=0, gbeta=1.0):
        """q_wp / m_wp: per parameter, the list of (period, row[, col]) weight positions (struct weightposition)."""
        pl = -np.ones((npops, npops), dtype=np.int32)
        for k, row in enumerate(plist):
            pl[k, :len(row)] = row
        q_off = _i32(np.cumsum([0] + [len(w) for w in q_wp]))
        q_p = _i32([t[0] for w in q_wp for t in w]); q_r = _i32([t[1] for w in q_wp for t in w])
        m_off = _i32(np.cumsum([0] + [len(w) for w in m_wp]))
        m_p = _i32([t[0] for w in m_wp for t in w]); m_r = _i32([t[1] for w in m_wp for t in w])
        m_c = _i32([t[2] for w in m_wp for t in w])
        nm = len(m_wp)
        m_mean = _f64(m_mean if m_mean is not None else np.zeros(max(nm, 1)))
        n_p = _i32([t[0] for t in nomig]); n_r = _i32([t[1] for t in nomig]); n_c = _i32([t[2] for t in nomig])
        self._ck(self.lib.ima2p_engine_set_model(
            self._h, npops, nsplit, _ip(pl), _ip(_i32(addpop)), _ip(_i32(droppops).reshape(-1)), _ip(_i32(pt_e)),
            _ip(_i32(pt_down)), rootpop, len(q_wp), _ip(q_off), _ip(q_p), _ip(q_r), _dp(_f64(q_max)), _dp(_f64(q_min)),
            nm, _ip(m_off), _ip(m_p), _ip(m_r), _ip(m_c), _dp(_f64(m_max)), _dp(_f64(m_min)), _dp(m_mean), len(nomig),
            _ip(n_p), _ip(n_r), _ip(n_c), nomigration, expoprior, thermo, float(gbeta)))
        self.npops, self.nsplit, self.nq, self.nm = npops, nsplit, len(q_wp), nm

    def set_model_flat(self, *create_args):
        """Same tables in the flat CSR form of ima2p_engine_set_model (used by the tests' fixture loader)."""
        self._ck(self.lib.ima2p_engine_set_model(self._h, *create_args))
        self.npops, self.nsplit, self.nq, self.nm = create_args[0], create_args[1], create_args[8], create_args[14]

    def set_model_from_tree(self, npops, tree, qmax, mmax, expo_prior=0, m_mean=0.0, thermo=0, gbeta=1.0):
        """Default model from the population tree string and the priors (setup_poptree + setup_iparams)."""
        h = C.c_void_p()
        self._ck(self.lib.ima2p_modelspec_create(C.byref(h), npops, tree.encode(), qmax, mmax, expo_prior, m_mean, thermo, gbeta))
        try:
            dims = (C.c_int * 6)()
            self._ck(self.lib.ima2p_modelspec_dims(h, dims))
            self._ck(self.lib.ima2p_engine_set_model_spec(self._h, h))
            self.adopt_model_dims(dims[0], dims[1], dims[3], dims[4])
        finally:
            self.lib.ima2p_modelspec_free(h)

    def adopt_model_dims(self, npops, nsplit, nq, nm):
        self.npops, self.nsplit, self.nq, self.nm = npops, nsplit, nq, nm

    def set_locus(self, li, model, numgenes, numsites, samppop, seq=None, mult=None, hval=1.0, totsites=0, nlinked=1,
                  minA=None, maxA=None, sumlogk=0.0):
        seq_a = _i32(seq) if seq is not None and numsites > 0 else None
        mult_a = _i32(mult) if mult is not None else None
        mina = _i32(minA if minA is not None else [0] * nlinked)
        maxa = _i32(maxA if maxA is not None else [0] * nlinked)
        self._ck(self.lib.ima2p_engine_set_locus(
            self._h, li, model, numgenes, numsites, totsites, float(hval), _ip(_i32(samppop)),
            _ip(seq_a) if seq_a is not None else None, _ip(mult_a) if mult_a is not None else None, nlinked, _ip(mina),
            _ip(maxa), float(sumlogk)))
        self._loci[li] = dict(numgenes=numgenes, numlines=2 * numgenes - 1, nlinked=nlinked, model=model)

    def finalize(self):
        self._ck(self.lib.ima2p_engine_finalize(self._h))
        d = _i32(np.zeros(5))
        self._ck(self.lib.ima2p_engine_dims(self._h, _ip(d)))
        self.NI, self.ND, self.NL, self.CAP, self.rowlen = (int(v) for v in d)

    def set_heating(self, heatmode, hval1, hval2=0.0):
        """setheat (swapchains.cpp:71-178)."""
        self._ck(self.lib.ima2p_engine_set_heating(self._h, heatmode, float(hval1), float(hval2)))

    def set_betas(self, betas_global):
        self._ck(self.lib.ima2p_engine_set_betas(self._h, _dp(_f64(betas_global))))

    # ---- state ---------------------------------------------------------------------------------------------
    def set_chain(self, ci, tvals):
        self._ck(self.lib.ima2p_engine_set_chain(self._h, ci, _dp(_f64(tvals))))

    def set_genealogy(self, ci, li, up0, up1, down, pop, time, mig_off, mig_t, mig_p, root, roottime, uvals=(1.0,),
                      kappa=2.0, pi=None, A=None):
        mt, mp = _f64(np.append(_f64(mig_t), 0.0)), _i32(np.append(_i32(mig_p), 0))
        a = _i32(A) if A is not None else None
        self._ck(self.lib.ima2p_engine_set_genealogy(
            self._h, ci, li, _ip(_i32(up0)), _ip(_i32(up1)), _ip(_i32(down)), _ip(_i32(pop)), _dp(_f64(time)),
            _ip(_i32(mig_off)), _dp(mt), _ip(mp), int(root), float(roottime), _dp(_f64(uvals)), float(kappa),
            _dp(_f64(pi)) if pi is not None else None, _ip(a) if a is not None else None))

    def upload(self):
        self._ck(self.lib.ima2p_engine_upload(self._h))

    def get_genealogy(self, ci, li, which=0):
        nl = self._loci[li]["numlines"]
        up0, up1, down, pop = (np.zeros(nl, np.int32) for _ in range(4))
        time, off = np.zeros(nl), np.zeros(nl + 1, np.int32)
        mt, mp = np.zeros(self.CAP), np.zeros(self.CAP, np.int32)
        root, rt = C.c_int(), C.c_double()
        self._ck(self.lib.ima2p_engine_get_genealogy(self._h, ci, li, which, _ip(up0), _ip(up1), _ip(down), _ip(pop),
                                                     _dp(time), _ip(off), _dp(mt), _ip(mp), self.CAP, C.byref(root),
                                                     C.byref(rt)))
        n = off[nl]
        return dict(up0=up0, up1=up1, down=down, pop=pop, time=time, mig_off=off, mig_t=mt[:n], mig_p=mp[:n],
                    root=root.value, roottime=rt.value)

    def get_alleles(self, ci, li, which=0):
        nl, k = self._loci[li]["numlines"], self._loci[li]["nlinked"]
        A, dl, pa = np.zeros((k, nl), np.int32), np.zeros((k, nl)), np.zeros(k)
        self._ck(self.lib.ima2p_engine_get_alleles(self._h, ci, li, which, _ip(A), _dp(dl), _dp(pa)))
        return dict(A=A, dlikeA=dl, pdg_a=pa)

    # ---- evaluation of the loaded state: init_p (mcmcfile.cpp:130-193) ---------------------------------------
    def eval(self):
        self._ck(self.lib.ima2p_engine_eval(self._h))

    def pair(self, ci, li):
        """treeweight + likelihood results of one (chain, locus): C[ci]->G[li].gweight, pdg, length, ..."""
        wi, wd = np.zeros(self.NI, np.int32), np.zeros(self.ND)
        od, oi = np.zeros(4), np.zeros(2, np.int32)
        self._ck(self.lib.ima2p_engine_get_pair(self._h, ci, li, _ip(wi), _dp(wd), _dp(od), _ip(oi)))
        return dict(wi=wi, wd=wd, pdg=od[0], length=od[1], tlength=od[2], roottime=od[3], mignum=int(oi[0]),
                    root=int(oi[1]))

    def chain(self, ci):
        """C[ci]->allgweight and allpcalc (struct probcalc)."""
        wi, wd = np.zeros(self.NI, np.int32), np.zeros(self.ND)
        q, m, od = np.zeros(max(self.nq, 1)), np.zeros(max(self.nm, 1)), np.zeros(3)
        self._ck(self.lib.ima2p_engine_get_chain(self._h, ci, _ip(wi), _dp(wd), _dp(q), _dp(m), _dp(od)))
        tv = np.zeros(max(self.nsplit, 1))
        self._ck(self.lib.ima2p_engine_get_split_times(self._h, ci, _dp(tv)))
        return dict(wi=wi, wd=wd, qintegrate=q[:self.nq], mintegrate=m[:self.nm], probg=od[0], pdg=od[1], beta=od[2],
                    tvals=tv[:self.nsplit])

    # ---- M mode --------------------------------------------------------------------------------------------
    def run(self, nsteps, swaptries=None, stream=None):
        """nsteps x [updategenealogy for every chain x locus; swapchains(swaptries)] (qupdate)."""
        if swaptries is None:
            swaptries = max(1, self.nchains_global // 10) if self.nchains_global > 1 else 0    # ima_main_mpi.cpp:1378
        self._ck(self.lib.ima2p_engine_run(self._h, nsteps, swaptries, stream))

    def set_pipeline(self, groups=1, depth=1, decisions_first=False):
        """Chain groups on their own streams and steps per CUDA graph of `run` (the chains a run visits do not depend on it)."""
        self._ck(self.lib.ima2p_engine_set_pipeline(self._h, groups, depth, 1 if decisions_first else 0))

    def launches_per_step(self, swaptries=None):
        """Kernel launches one step of `run` makes with the current settings."""
        if swaptries is None:
            swaptries = self.default_swaptries()
        return int(self.lib.ima2p_engine_launches_per_step(self._h, swaptries))

    # ---- chains sharded over GPUs: swap sums exchanged through peer memory by the kernels themselves ------------------
    def exchange_create(self):
        """This rank's exchange table: (device pointer, bytes)."""
        p, n = C.c_void_p(), C.c_uint64()
        self._ck(self.lib.ima2p_engine_exchange_create(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def exchange_handle(self):
        """64 bytes another process of this node can open with exchange_open."""
        ptr, _ = self.exchange_create()
        buf = C.create_string_buffer(64)
        self._ck(self.lib.ima2p_ipc_export(C.c_void_p(ptr), buf))
        return buf.raw

    def exchange_open(self, handle64, device=0):
        p = C.c_void_p()
        self._ck(self.lib.ima2p_ipc_import(device, C.c_char_p(handle64), C.byref(p)))
        return p.value

    def exchange_attach(self, tables):
        """tables[r] = rank r's table as this GPU addresses it (None for this rank)."""
        arr = (C.c_void_p * len(tables))(*[C.c_void_p(t) if t else None for t in tables])
        self._ck(self.lib.ima2p_engine_exchange_attach(self._h, arr))

    def run_sharded(self, nsteps, swaptries=None, stream=None):
        self._ck(self.lib.ima2p_engine_run_sharded(self._h, nsteps, self.default_swaptries() if swaptries is None else swaptries, stream))

    def sharded_update(self, stream=None):
        self._ck(self.lib.ima2p_engine_sharded_update(self._h, stream))

    def sharded_swap(self, swaptries=None, stream=None):
        self._ck(self.lib.ima2p_engine_sharded_swap(self._h, self.default_swaptries() if swaptries is None else swaptries, stream))

    def set_proposal_path(self, fast=True, pairs_per_warp=0):
        """Two-kernel proposal path (lane-per-pair move + warp-per-pair weights) or the general kernel for every pair."""
        self._ck(self.lib.ima2p_engine_set_proposal_path(self._h, 1 if fast else 0, pairs_per_warp))

    def grow_capacity(self, new_capacity):
        """More room for migration events per genealogy, between two steps (checkmig, utilities.cpp:1365-1383)."""
        self._ck(self.lib.ima2p_engine_grow_capacity(self._h, new_capacity))
        self.CAP = max(self.CAP, new_capacity)

    def set_debug_records(self, on=True):
        """Keep the per-proposal record `proposal()` reads (parity tests)."""
        self._ck(self.lib.ima2p_engine_set_debug_records(self._h, 1 if on else 0))

    def set_speculation(self, depth):
        """Loci evaluated per round of the accept sweep (1..4); the results do not depend on it."""
        self._ck(self.lib.ima2p_engine_set_speculation(self._h, depth))

    def default_swaptries(self):
        return max(1, self.nchains_global // 10) if self.nchains_global > 1 else 0             # ima_main_mpi.cpp:1378

    def run_timed(self, nsteps, swaptries=None, stream=None):
        """run() launched kernel by kernel; returns summed device ms of (proposal kernels together, accept, swap, split_t,
        accept_t, changeu, k_move, k_weigh, k_propose_redo, 3 unused)."""
        ms = np.zeros(12, np.float32)
        self._ck(self.lib.ima2p_engine_run_timed(self._h, nsteps, self.default_swaptries() if swaptries is None else swaptries,
                                                 stream, ms.ctypes.data_as(capi.c_flt_p)))
        return ms

    def update_genealogies(self, dev_S_local=None, stream=None):
        self._ck(self.lib.ima2p_engine_update_genealogies(self._h, dev_S_local, stream))

    def swap_replay(self, dev_S_global, swaptries, stream=None):
        self._ck(self.lib.ima2p_engine_swap_replay(self._h, dev_S_global, swaptries, stream))

    # split-phase step (multi-GPU): see include/ima2p_b200.h
    def step_propose(self, stream=None):
        self._ck(self.lib.ima2p_engine_step_propose(self._h, stream))

    def step_decide(self, dev_S_local, stream=None):
        self._ck(self.lib.ima2p_engine_step_decide(self._h, dev_S_local, stream))

    def swap_replay_late(self, dev_S_global, swaptries, stream=None):
        self._ck(self.lib.ima2p_engine_swap_replay_late(self._h, dev_S_global, swaptries, stream))

    def sync(self):
        self._ck(self.lib.ima2p_engine_sync(self._h))

    def proposal(self, ci, li):
        out, fl, buf = np.zeros(5), C.c_uint(), C.c_int()
        self._ck(self.lib.ima2p_engine_get_proposal(self._h, ci, li, _dp(out), C.byref(fl), C.byref(buf)))
        return dict(migweight=out[0], slideweight=out[1], slidedist=out[2], edge=int(out[3]), extra=out[4],
                    aterm=out[4] - out[0] - out[1], flags=fl.value, buffer=buf.value)

    def counters(self):
        out = np.zeros(8, np.uint64)
        self._ck(self.lib.ima2p_engine_counters(self._h, out.ctypes.data_as(capi.c_u64_p)))
        keys = ["steps", "updates", "accepted", "topology", "tmrca", "swap_attempts", "swaps", "dropped"]
        return dict(zip(keys, (int(v) for v in out)))

    def set_update_schedule(self, t_updates=3, u_every=5):
        """Split-time updates every step (1 Rannala-Yang, 2 Nielsen-Wakeley, 3 either at random as the reference does)
        and mutation-scalar updates every ``u_every``-th (ima_main_mpi.cpp:1784-1785, 1871-1872)."""
        self._ck(self.lib.ima2p_engine_set_update_schedule(self._h, int(t_updates), int(u_every)))

    def set_update_priors(self, t_max=None, t_min=None, u_prior_max=0.0, u_window=0.0, kappa_window=0.0, kappa_max=0.0):
        tm = None if t_max is None else _f64(np.atleast_1d(t_max))
        tn = None if t_min is None else _f64(np.atleast_1d(t_min))
        self._ck(self.lib.ima2p_engine_set_update_priors(self._h, None if tm is None else _dp(tm), None if tn is None else _dp(tn),
                                                         u_prior_max, u_window, kappa_window, kappa_max))

    def update_counters(self):
        out = (C.c_uint64 * 4)()
        self._ck(self.lib.ima2p_engine_update_counters(self._h, out))
        return dict(t_tries=out[0], t_accepts=out[1], u_tries=out[2], u_accepts=out[3])

    def fetch_chain_pdg(self, chain):
        out = np.zeros(self.nloci)
        self._ck(self.lib.ima2p_engine_fetch_chain_pdg(self._h, chain, _dp(out)))
        return out

    def cold_counters(self, nsplit, nurates):
        """Cold-chain update counts as the reference's update-rate tables report them, and adjacent-temperature swaps."""
        g = np.zeros((self.nloci, 3), dtype=np.uint64)
        t = np.zeros((max(nsplit, 1), 4), dtype=np.uint64)
        u = np.zeros((max(nurates, 1), 2), dtype=np.uint64)
        a = np.zeros((max(self.nchains_global - 1, 1), 2), dtype=np.uint64)
        p = lambda x: x.ctypes.data_as(C.POINTER(C.c_uint64))
        self._ck(self.lib.ima2p_engine_cold_counters(self._h, p(g), p(t), p(u), p(a)))
        return dict(genealogy=g, split=t[:nsplit], scalars=u[:nurates], adjacent=a[:self.nchains_global - 1])

    def scalars(self, chain, locus):
        u = np.zeros(capi.MAX_LINKED)
        k = C.c_double()
        self._ck(self.lib.ima2p_engine_get_scalars(self._h, chain, locus, _dp(u), C.byref(k)))
        return u, k.value

    def fetch_parameters(self):
        """(tvals[nchains][nsplit], uvals[nchains][nloci][MAX_LINKED], kappa[nchains][nloci]) of the current state."""
        tv = np.zeros((self.nchains, max(self.nsplit, 1)))
        u = np.zeros((self.nchains, self.nloci, capi.MAX_LINKED))
        k = np.zeros((self.nchains, self.nloci))
        self._ck(self.lib.ima2p_engine_fetch_parameters(self._h, _dp(tv) if self.nsplit else None, _dp(u), _dp(k)))
        return tv[:, :self.nsplit], u, k

    def debug_split_time(self, period, newt=None, force_accept=-1, method=0):
        """One changet_RY1 (method 0) / changet_NW (method 1) on every chain; rows of (period, newt, MH term, accepted)."""
        out = np.zeros((self.nchains, 4))
        nt = None if newt is None else _f64(newt)
        self._ck(self.lib.ima2p_engine_debug_split_time(self._h, method, period, None if nt is None else _dp(nt), force_accept, _dp(out)))
        return out

    def debug_changeu(self, chain, j, k, d, kappa_j=0.0, kappa_k=0.0):
        out = np.zeros(4)
        self._ck(self.lib.ima2p_engine_debug_changeu(self._h, chain, j, k, d, kappa_j, kappa_k, _dp(out)))
        return out

    def step_report(self, stream=None):
        """(chain summary [nchains][4] = beta, probg, pdg, S; cold-chain .ti row or None) with one device-to-host copy."""
        out = np.zeros((self.nchains, 4))
        row = np.zeros(self.rowlen, np.float32)
        present = C.c_int()
        self._ck(self.lib.ima2p_engine_step_report(self._h, _dp(out), row.ctypes.data_as(capi.c_flt_p), C.byref(present), stream))
        return out, (row if present.value else None)

    def step_report_begin(self, slot, stream=None):
        """First half of step_report: queue the packing kernel and the copy into slot 0 / 1, return at once."""
        self._ck(self.lib.ima2p_engine_step_report_begin(self._h, int(slot), stream))

    def step_report_end(self, slot):
        """Second half: wait for that slot's copy and return what step_report returns."""
        out = np.zeros((self.nchains, 4))
        row = np.zeros(self.rowlen, np.float32)
        present = C.c_int()
        self._ck(self.lib.ima2p_engine_step_report_end(self._h, int(slot), _dp(out), row.ctypes.data_as(capi.c_flt_p), C.byref(present)))
        return out, (row if present.value else None)

    def write_mcf(self, path):
        """writemcf (mcmcfile.cpp:203-296): the state of the local chains in the reference's .mcf format."""
        self._ck(self.lib.ima2p_engine_write_mcf(self._h, str(path).encode()))

    def read_mcf(self, path):
        """readmcf (mcmcfile.cpp:310-442) + init_p: load, upload and evaluate."""
        self._ck(self.lib.ima2p_engine_read_mcf(self._h, str(path).encode()))

    def thermo_accumulate(self, stream=None):
        """summarginlikecalc (marglike.cpp:51-87) for the local chains."""
        self._ck(self.lib.ima2p_engine_thermo_accumulate(self._h, stream))

    def thermo_sums(self, reset=False):
        out = np.zeros(self.nchains_global)
        self._ck(self.lib.ima2p_engine_thermo_sums(self._h, _dp(out), int(reset)))
        return out

    def thermo_marginlike(self, thermosum, k):
        """thermomarginlikecalc (marglike.cpp:121-150)."""
        t = _f64(thermosum)
        out = C.c_double()
        self._ck(self.lib.ima2p_thermo_marginlike(_dp(t), len(t), int(k), C.byref(out)))
        return out.value

    def betas(self):
        b = np.zeros(self.nchains_global)
        self._ck(self.lib.ima2p_engine_get_betas(self._h, _dp(b)))
        return b

    def cold_row(self):
        """savegsampinf (ginfo.cpp:318-377) of the chain with beta == 1, or None when it lives on another GPU."""
        row, present = np.zeros(self.rowlen, np.float32), C.c_int()
        self._ck(self.lib.ima2p_engine_cold_row(self._h, row.ctypes.data_as(capi.c_flt_p), C.byref(present)))
        return row if present.value else None

    # ---- bulk state I/O in the engine's packed layout (end-to-end timing path) -----------------------------
    def state_bytes(self):
        out = np.zeros(8, np.uint64)
        self._ck(self.lib.ima2p_engine_state_bytes(self._h, out.ctypes.data_as(capi.c_u64_p)))
        return [int(v) for v in out]

    def put_state(self, bufs, tvals, stream=None):
        """bufs: 8 host buffers (objects with .ctypes or integer addresses) in state_bytes() order."""
        ptrs = [b if isinstance(b, int) else b.ctypes.data for b in bufs]
        tv = _f64(tvals)
        self._ck(self.lib.ima2p_engine_put_state(self._h, *ptrs, _dp(tv), stream))

    @staticmethod
    def pack_state(topo, mseg):
        """(topo int16 [P][NL][4], mseg uint16 [P][NL][2]) -> (topo8 int8, mcount uint8) of put_state_packed, or None when
        the state does not fit that form (more than 127 edges, or pools that are not in edge order)."""
        topo, mseg = np.asarray(topo), np.asarray(mseg)
        if topo.shape[1] > 127 or topo.min() < -128 or topo.max() > 127 or mseg[..., 1].max() > 255:
            return None
        cnt = mseg[..., 1].astype(np.int64)
        start = np.cumsum(cnt, axis=1) - cnt
        if not np.array_equal(start[cnt > 0], mseg[..., 0].astype(np.int64)[cnt > 0]):
            return None
        return np.ascontiguousarray(topo.astype(np.int8)), np.ascontiguousarray(cnt.astype(np.uint8))

    def pack_state_block(self, arrs, tvals, out=None):
        """The 8 put_state arrays + split times -> one uint8 block for put_state_block (None when the state does not fit the
        8-bit wire form).  Returns (block, total_events); `out` may be a preallocated (e.g. pinned) uint8 array."""
        P, NL, CAP = self.nchains * self.nloci, self.NL, self.CAP
        topo, time, mseg, mig_t, mig_p, si, sd, uv = [np.asarray(a) for a in arrs]
        pk = self.pack_state(topo.reshape(P, NL, 4), mseg.reshape(P, NL, 2))
        if pk is None:
            return None
        nm = si.reshape(P, 2)[:, 1].astype(np.int64)
        keep = np.arange(CAP)[None, :] < nm[:, None]
        events = int(nm.sum())
        lay = np.zeros(10, np.uint64)
        self._ck(self.lib.ima2p_engine_state_block_layout(self._h, events, lay.ctypes.data_as(capi.c_u64_p)))
        lay = [int(v) for v in lay]
        blk = np.zeros(lay[9], np.uint8) if out is None else out[:lay[9]]
        parts = [np.ascontiguousarray(time, np.float64), np.ascontiguousarray(sd, np.float64), np.ascontiguousarray(uv, np.float64),
                 np.ascontiguousarray(tvals, np.float64), np.ascontiguousarray(mig_t.reshape(P, CAP)[keep], np.float64),
                 np.ascontiguousarray(si, np.int32), np.ascontiguousarray(mig_p.reshape(P, CAP)[keep], np.int16), pk[0], pk[1]]
        for off, a in zip(lay[:9], parts):
            b = a.reshape(-1).view(np.uint8)
            blk[off:off + b.size] = b
        return blk, events

    def put_state_block(self, block, events, stream=None):
        ptr = block if isinstance(block, int) else block.ctypes.data
        self._ck(self.lib.ima2p_engine_put_state_block(self._h, ptr, events, stream))

    def upload_block(self, block, events, copy_stream=None):
        """First half of put_state_block: the transfer only, into one of two staging slots (on copy_stream)."""
        ptr = block if isinstance(block, int) else block.ctypes.data
        self._ck(self.lib.ima2p_engine_upload_block(self._h, ptr, events, copy_stream))

    def adopt_block(self, stream=None):
        """Second half: `stream` waits for the oldest uploaded block, widens it into the resident state, re-evaluates it."""
        self._ck(self.lib.ima2p_engine_adopt_block(self._h, stream))

    def put_state_packed(self, bufs, tvals, stream=None):
        """bufs as put_state, with bufs[0] = topo8 and bufs[2] = mcount from pack_state (8-bit wire form)."""
        ptrs = [b if isinstance(b, int) else b.ctypes.data for b in bufs]
        tv = _f64(tvals)
        self._ck(self.lib.ima2p_engine_put_state_packed(self._h, *ptrs, _dp(tv), stream))

    def fetch_state(self, bufs, stream=None):
        ptrs = [b if isinstance(b, int) else b.ctypes.data for b in bufs]
        self._ck(self.lib.ima2p_engine_fetch_state(self._h, *ptrs, stream))

    def fetch_pair_summaries(self, stream=None):
        """(sd[P][4] = roottime, length, tlength, pdg; si[P][2] = root, mignum; wi[P][NI]) of the current genealogies."""
        P = self.nchains * self.nloci
        sd, si, wi = np.zeros((P, 4)), np.zeros((P, 2), np.int32), np.zeros((P, self.NI), np.int32)
        self._ck(self.lib.ima2p_engine_fetch_pair_summaries(self._h, _dp(sd), _ip(si), _ip(wi), stream))
        return sd, si, wi

    def fetch_chain_summary(self, stream=None):
        out = np.zeros((self.nchains, 4))
        self._ck(self.lib.ima2p_engine_fetch_chain_summary(self._h, _dp(out), stream))
        return out


class LMode:
    """L mode over the sampled-genealogy rows of a .ti file (``gsampinf``, ginfo.cpp:288-304)."""

    def __init__(self, nq, nm, nsplit, q_max, q_min, m_max, m_min, m_mean=None, expoprior=0, device=0, lib=None):
        self.lib = lib if lib is not None else capi.lib()
        self.nq, self.nm, self.nsplit = nq, nm, nsplit
        self._h = C.c_void_p()
        mm = _f64(m_mean if m_mean is not None else np.zeros(max(nm, 1)))
        capi.check(self.lib, self.lib.ima2p_lmode_create(C.byref(self._h), device, nq, nm, nsplit, _dp(_f64(q_max)),
                                                         _dp(_f64(q_min)), _dp(_f64(m_max)), _dp(_f64(m_min)), _dp(mm),
                                                         expoprior))
        self.nrows = 0

    def close(self):
        if self._h:
            self.lib.ima2p_lmode_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load(self, rows, nrows_total=None, row0=0):
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        self.nrows, self.rowlen = rows.shape
        self.nrows_total = self.nrows if nrows_total is None else nrows_total
        self.row0 = row0
        capi.check(self.lib, self.lib.ima2p_lmode_load(self._h, rows.ctypes.data_as(capi.c_flt_p), self.nrows,
                                                       self.rowlen, self.nrows_total))

    def marginal_sums(self, param, x, first=0, last=None, round_counts=1):
        x = _f64(np.atleast_1d(x))
        out = np.zeros(len(x))
        capi.check(self.lib, self.lib.ima2p_lmode_marginal_sums(self._h, param, _dp(x), len(x), first,
                                                                self.nrows if last is None else last, round_counts,
                                                                _dp(out), None, None))
        return out

    def margincalc(self, x, yadjust, pi, logi):
        """margincalc(x, yadjust, pi, logi) (surface_call_functions.cpp:119-173), vectorised over x."""
        x = _f64(np.atleast_1d(x))
        out = np.zeros(len(x))
        capi.check(self.lib, self.lib.ima2p_lmode_margincalc(self._h, pi, _dp(x), len(x), float(yadjust), int(logi),
                                                             _dp(out)))
        return out

    def marginp(self, param, firsttree, lasttree, x):
        """marginp(param, firsttree, lasttree, x) (surface_call_functions.cpp:25-80), vectorised over x."""
        x = _f64(np.atleast_1d(x))
        out = np.zeros(len(x))
        capi.check(self.lib, self.lib.ima2p_lmode_marginp(self._h, param, firsttree, lasttree, _dp(x), len(x), _dp(out)))
        return out

    def marginal_many(self, kind, param, first, last, x, yadjust=None):
        """The current points of many independent searches in one device pass (ima2p_lmode_marginal_many): request q is
        marginp(param[q], first[q], last[q], x[q]) for kind[q] = 0 and log margincalc(x[q]) - yadjust[q] for kind[q] = 1; the
        lock-step form of the calls marginalopt / margin95 make one at a time (surface_call_functions.cpp:175-297)."""
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        kind, param, first, last = i32(kind), i32(param), i32(first), i32(last)
        x = _f64(np.atleast_1d(x))
        ya = _f64(np.zeros(len(x)) if yadjust is None else yadjust)
        out = np.zeros(len(x))
        ip = lambda a: a.ctypes.data_as(capi.c_int_p)
        capi.check(self.lib, self.lib.ima2p_lmode_marginal_many(self._h, len(x), ip(kind), ip(param), ip(first), ip(last), _dp(x),
                                                                _dp(ya), _dp(out)))
        return out

    def moments(self):
        """print_means_variances_correlations (output.cpp:687-745): means, variances, correlations (p < q entries) of the
        parameters from the calcx sums over every row, plus the raw sums (sum0[np], sum1[np], cross[np][np])."""
        n = self.nq + self.nm
        means, var, corr, raw = np.zeros(n), np.zeros(n), np.zeros((n, n)), np.zeros(2 * n + n * n)
        capi.check(self.lib, self.lib.ima2p_lmode_moments(self._h, _dp(means), _dp(var), _dp(corr), _dp(raw)))
        return means, var, corr, {"sum0": raw[:n].copy(), "sum1": raw[n:2 * n].copy(), "cross": raw[2 * n:].reshape(n, n).copy()}

    def moments_raw(self):
        """Row sums behind moments() over this handle's rows (additive over ranks), flat: sum0[np], sum1[np], cross[np][np]."""
        n = self.nq + self.nm
        means, var, raw = np.zeros(n), np.zeros(n), np.zeros(2 * n + n * n)
        capi.check(self.lib, self.lib.ima2p_lmode_moments(self._h, _dp(means), _dp(var), None, _dp(raw)))
        return raw

    def moments_finish(self, raw, nrows_total):
        n = self.nq + self.nm
        raw = _f64(raw)
        means, var, corr = np.zeros(n), np.zeros(n), np.zeros((n, n))
        self.lib.ima2p_lmode_moments_finish(n, _dp(raw), int(nrows_total), _dp(means), _dp(var), _dp(corr))
        return means, var, corr

    def popmig_sums(self, thetai, mi, x, first=0, last=None):
        x = _f64(np.atleast_1d(x))
        out = np.zeros(len(x))
        capi.check(self.lib, self.lib.ima2p_lmode_popmig_sums(self._h, thetai, mi, _dp(x), len(x), first, self.nrows if last is None else last, _dp(out)))
        return out

    def popmig(self, thetai, mi, x, prob_or_like=0):
        """calc_popmig / calc_pop_expomig (popmig.cpp:9-170): density of 2NM at x, vectorised over x."""
        x = _f64(np.atleast_1d(x))
        out = np.zeros(len(x))
        capi.check(self.lib, self.lib.ima2p_lmode_popmig(self._h, thetai, mi, _dp(x), len(x), int(prob_or_like), _dp(out)))
        return out

    def marginpopmig(self, mi, firsttree, lasttree, x, thetai):
        """marginpopmig / marginpop_expomig (popmig.cpp:176-357), argument order of the reference, vectorised over x."""
        x = _f64(np.atleast_1d(x))
        out = np.zeros(len(x))
        capi.check(self.lib, self.lib.ima2p_lmode_marginpopmig(self._h, thetai, mi, firsttree, lasttree, _dp(x), len(x), _dp(out)))
        return out

    def greater_than(self, kind, i, j):
        """gtpops(i, j) (kind 0) / gtmig(i, j) (kind 1) (gtint.cpp:128-330): P(parameter i > parameter j); -1 = "na"."""
        out = C.c_double(0.0)
        capi.check(self.lib, self.lib.ima2p_lmode_greater_than(self._h, int(kind), int(i), int(j), C.byref(out)))
        return out.value

    def jointp(self, x, calc_ess=True):
        """jointp(x, calc_ess, &ess) (jointfind.cpp:885-1047) for a batch of parameter vectors x[nvec][nq+nm]."""
        x = _f64(np.atleast_2d(x))
        q, ess = np.zeros(len(x)), np.zeros(len(x))
        capi.check(self.lib, self.lib.ima2p_lmode_jointp(self._h, _dp(x), len(x), int(calc_ess), _dp(q), _dp(ess)))
        return q, ess

    def set_joint_model(self, modeltype):
        """nowmodeltype of findjointpeaks (jointfind.cpp:1104-1133): 0 all parameters (two populations), 1 the population sizes,
        2 the migration rates (the two searches of a three-population analysis); applies to the joint evaluations that follow."""
        capi.check(self.lib, self.lib.ima2p_lmode_set_joint_model(self._h, int(modeltype)))

    # sharded form: see ima2p_lmode_joint_phase1/2 in include/ima2p_b200.h
    def joint_phase1(self, x, seed_before=None):
        x = _f64(np.atleast_2d(x))
        out = np.zeros(len(x))
        sb = _f64(seed_before) if seed_before is not None else None
        capi.check(self.lib, self.lib.ima2p_lmode_joint_phase1(self._h, _dp(x), len(x), _dp(sb) if sb is not None else None,
                                                               _dp(out)))
        return out

    def joint_reseed(self, nvec, seed_before):
        out = np.zeros(nvec)
        sb = _f64(seed_before)
        capi.check(self.lib, self.lib.ima2p_lmode_joint_reseed(self._h, nvec, _dp(sb), _dp(out)))
        return out

    def joint_phase2(self, nvec, globalmax):
        rec = np.zeros((nvec, 6))
        capi.check(self.lib, self.lib.ima2p_lmode_joint_phase2(self._h, nvec, _dp(_f64(globalmax)), self.row0, _dp(rec)))
        return rec

    def joint_begin(self, x, dev_localmax, stream=None):
        """Device-resident phase 1: dev_localmax = device pointer to len(x) doubles (see ima2p_lmode_joint_begin)."""
        x = _f64(np.atleast_2d(x))
        capi.check(self.lib, self.lib.ima2p_lmode_joint_begin(self._h, _dp(x), len(x), dev_localmax, stream))

    def joint_middle(self, nvec, dev_allmax, world, rank, dev_records, stream=None):
        capi.check(self.lib, self.lib.ima2p_lmode_joint_middle(self._h, nvec, dev_allmax, world, rank, self.row0, dev_records, stream))

    def joint_finish(self, rec, globalmax, calc_ess=True):
        q, ess = C.c_double(), C.c_double()
        self.lib.ima2p_lmode_joint_finish(_dp(_f64(rec)), float(globalmax), self.nrows_total, int(calc_ess), C.byref(q),
                                          C.byref(ess))
        return q.value, ess.value
