"""Model tables built from the population tree string (C ABI ima2p_modelspec_*) against the reference's own tables
(setup_poptree + setup_iparams, dumped into the "model" object of every fixture): 2, 3 and 4 populations, balanced and
caterpillar trees, with and without migration, exponential migration prior.  Host code; no GPU needed."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from support import FlatModel, check_static_eval, engine_from_fixture, load_golden

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "hostemu", "libima2p_hostemu.so")
CASES = {"state_sim5_hn4": "(0,1):2", "state_sim5_3pop_hn2": "((0,1):3,2):4", "state_sim5_4popA_hn2": "((0,1):4,(2,3):5):6",
         "state_sim5_4popB_hn2": "(((0,1):4,2):5,3):6", "state_sim5_nomig_hn2": "(0,1):2", "state_sim5_3pop_nomig_hn2": "((0,1):3,2):4",
         "state_sim5_expo_hn2": "(0,1):2"}


@pytest.fixture(scope="module")
def lib():
    from ima2p_b200 import capi
    subprocess.run([os.path.join(HERE, "hostemu", "build.sh")], check=True)
    return capi.bind(EMU)


def _spec(lib, fm, tree, mj):
    from ima2p_b200 import capi
    h = C.c_void_p()
    mmax = 0.0 if fm.nomigration else (mj["imig"][0]["max"] if mj["imig"] else 0.0)
    mean = mj["imig"][0]["mean"] if (fm.expoprior and mj["imig"]) else 0.0
    capi.check(lib, lib.ima2p_modelspec_create(C.byref(h), fm.npops, tree.encode(), mj["itheta"][0]["max"], mmax, fm.expoprior, mean, fm.thermo, fm.gbeta))
    return h


@pytest.mark.parametrize("name", sorted(CASES))
def test_tables_match_the_reference(lib, name):
    from ima2p_b200 import capi
    d = load_golden(name)
    mj = d["model"]
    fm = FlatModel(mj)
    h = _spec(lib, fm, CASES[name], mj)
    dims = (C.c_int * 6)()
    capi.check(lib, lib.ima2p_modelspec_dims(h, dims))
    assert list(dims)[:5] == [fm.npops, fm.nsplit, fm.ntreepops, fm.nq, fm.nm]
    n, nt = fm.npops, fm.ntreepops
    z = lambda k: np.zeros(max(k, 1), np.int32)
    plist, addpop, drop, pb, pe, pd = z(n * n), z(n), z(2 * n), z(nt), z(nt), z(nt)
    qo, qp, qr = z(fm.nq + 1), z(len(fm.q_p)), z(len(fm.q_p))
    mo, mp, mr, mc = z(fm.nm + 1), z(dims[5]), z(dims[5]), z(dims[5])
    ip = lambda a: a.ctypes.data_as(capi.c_int_p)
    capi.check(lib, lib.ima2p_modelspec_tables(h, ip(plist), ip(addpop), ip(drop), ip(pb), ip(pe), ip(pd), ip(qo), ip(qp), ip(qr), ip(mo), ip(mp), ip(mr), ip(mc)))
    lib.ima2p_modelspec_free(h)
    assert np.array_equal(plist.reshape(n, n), fm.plist)
    assert np.array_equal(addpop[:fm.nsplit + 1], fm.addpop[:fm.nsplit + 1]) and np.array_equal(drop[:2 * (fm.nsplit + 1)], fm.droppops[:2 * (fm.nsplit + 1)])
    assert np.array_equal(pe[:nt], fm.pt_e) and np.array_equal(pd[:nt], fm.pt_down)
    assert [p["b"] for p in mj["poptree"]] == list(pb[:nt])
    assert np.array_equal(qo[:fm.nq + 1], fm.q_off) and np.array_equal(qp[:len(fm.q_p)], fm.q_p) and np.array_equal(qr[:len(fm.q_r)], fm.q_r)
    assert dims[5] == len(fm.m_p)
    assert np.array_equal(mo[:fm.nm + 1], fm.m_off) and np.array_equal(mp[:len(fm.m_p)], fm.m_p)
    assert np.array_equal(mr[:len(fm.m_r)], fm.m_r) and np.array_equal(mc[:len(fm.m_c)], fm.m_c)


@pytest.mark.parametrize("name", ["state_sim5_4popA_hn2", "state_sim5_4popB_hn2", "state_sim5_expo_hn2"])
def test_engine_set_up_from_a_tree_string_evaluates_to_reference_values(lib, name):
    """set_model_spec instead of the dumped tables: the reference's state then evaluates to the reference's values."""
    from ima2p_b200 import Engine, capi
    from support import FlatTree
    d = load_golden(name)
    fm = FlatModel(d["model"])
    eng = Engine(len(d["chains"]), len(d["loci"]), mig_capacity=64, seed=1, lib=lib)
    h = _spec(lib, fm, CASES[name], d["model"])
    capi.check(lib, lib.ima2p_engine_set_model_spec(eng._h, h))
    lib.ima2p_modelspec_free(h)
    eng.adopt_model_dims(fm.npops, fm.nsplit, fm.nq, fm.nm)
    for li, loc in enumerate(d["loci"]):
        eng.set_locus(li, loc["model"], loc["numgenes"], loc["numsites"], loc["samppop"], seq=loc["seq"], hval=loc["hval"], sumlogk=loc["sumlogk"])
    eng.finalize()
    eng.set_betas([ch["beta"] for ch in d["chains"]])
    for k, ch in enumerate(d["chains"]):
        eng.set_chain(k, ch["tvals"])
        for li, g in enumerate(ch["G"]):
            t = FlatTree(g["tree"])
            eng.set_genealogy(k, li, t.up0, t.up1, t.down, t.pop, t.time, t.mig_off, t.mig_t[:-1], t.mig_p[:-1], t.root, t.roottime, uvals=g["uvals"])
    eng.upload()
    eng.eval()
    check_static_eval(eng, fm, d, rtol=1e-10)
    eng.close()


def test_bad_tree_strings_are_rejected(lib):
    from ima2p_b200 import capi
    for npops, tree in ((3, "((0,1):4,2):3"), (3, "((0,1):3,2)"), (2, "(0,2):2"), (3, "((0,1):3,1):4"), (4, "((0,1):4,(2,3):4):6")):
        h = C.c_void_p()
        assert lib.ima2p_modelspec_create(C.byref(h), npops, tree.encode(), 10.0, 1.0, 0, 0.0, 0, 1.0) != 0, tree
