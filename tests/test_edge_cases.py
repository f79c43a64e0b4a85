"""Edge cases of the path's entry points, through the tests-only host-emulation build of the same sources: the smallest
engine, empty and single-row genealogy files, samples that do not fit the narrow upload form, states the packers must refuse,
argument errors.  (The GPU suite runs the parity checks proper; these are host logic and kernel logic on degenerate sizes.)"""
import os
import subprocess

import numpy as np
import pytest

from ima2p_b200 import Engine, LMode, capi, synth
from ima2p_b200.capi import Ima2pError
from ima2p_b200.readu import ti_append, ti_create, ti_load

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "hostemu", "libima2p_hostemu.so")
KEYS = ["topo", "time", "mseg", "mig_t", "mig_p", "scal_i", "scal_d", "uvals"]


@pytest.fixture(scope="module")
def emu():
    subprocess.run([os.path.join(HERE, "hostemu", "build.sh")], check=True)
    return capi.bind(EMU)


def _engine(lib, nloci, nchains, n0=15, n1=15, seed=4):
    loci = synth.make_dataset(nloci, n0, n1, seed=11)
    eng = Engine(nchains, nloci, mig_capacity=64, seed=seed, lib=lib)
    eng.set_model(**synth.two_population_model(10.0, 1.0))
    for li, L in enumerate(loci):
        eng.set_locus(li, 0, L["n"], L["numsites"], L["samppop"], seq=L["seq"])
    eng.finalize()
    if nchains > 1:
        eng.set_heating(0, 0.05, 0.0)
    st = synth.initial_state(loci, nchains, eng.NL, eng.CAP, t0=1.5, seed=100)
    arrs = [np.ascontiguousarray(st[k]) for k in KEYS]
    eng.put_state(arrs, st["tvals"])
    eng.set_update_priors(t_max=[3.0])
    eng.set_update_schedule(3, 5)
    return eng, loci, arrs, st["tvals"]


def test_one_chain_one_locus(emu):
    # no coupling at all: no swaps, a sweep of one locus, scalar updates with a single scalar (nothing to update)
    eng, _, _, _ = _engine(emu, 1, 1)
    eng.run(200, 0)
    eng.sync()
    cnt = eng.counters()
    assert cnt["steps"] == 200 and cnt["updates"] == 200 and cnt["swap_attempts"] == 0 and 0 < cnt["accepted"] < 200
    inc = eng.chain(0)
    eng.eval()
    fresh = eng.chain(0)
    assert np.array_equal(fresh["wi"], inc["wi"]) and abs(fresh["probg"] - inc["probg"]) <= 1e-9 * abs(fresh["probg"])
    assert int(np.sum(inc["wi"][:3])) == 29
    eng.close()


def test_narrow_upload_forms_refuse_what_does_not_fit(emu):
    # 70 + 70 genes: 279 edges do not fit 8-bit links; the full-width upload is the way and works
    eng, _, arrs, tv = _engine(emu, 2, 2, n0=70, n1=70)
    assert eng.NL == 279
    assert Engine.pack_state(arrs[0].reshape(4, eng.NL, 4), arrs[2].reshape(4, eng.NL, 2)) is None
    assert eng.pack_state_block(arrs, tv) is None
    with pytest.raises(Ima2pError):
        eng.put_state_packed(arrs, tv)
    with pytest.raises(Ima2pError):
        eng.put_state_block(np.zeros(64, np.uint8), 0)
    eng.run(5)
    eng.sync()
    assert eng.counters()["updates"] == 5 * 4
    eng.close()


def test_pools_out_of_edge_order_are_not_packed(emu):
    eng, _, arrs, tv = _engine(emu, 3, 2)
    eng.run(80)
    eng.sync()
    eng.fetch_state(arrs[:7])
    P = 6
    mseg = arrs[2].reshape(P, eng.NL, 2).copy()
    assert Engine.pack_state(arrs[0].reshape(P, eng.NL, 4), mseg) is not None
    has = np.argwhere(mseg[..., 1] > 0)
    assert len(has) > 0
    p, e = has[-1]
    mseg[p, e, 0] += 3                         # a segment that does not start where the prefix sum says
    assert Engine.pack_state(arrs[0].reshape(P, eng.NL, 4), mseg) is None
    eng.close()


def test_ti_files_empty_and_single_row(emu, tmp_path):
    rowlen = 21
    path = tmp_path / "empty.ti"
    ti_create(path, "header only", lib=emu)
    assert ti_load(path, rowlen, lib=emu).shape == (0, rowlen)
    row = np.arange(rowlen, dtype=np.float32)[None, :] * 0.25
    ti_append(path, row, lib=emu)
    back = ti_load(path, rowlen, lib=emu)
    assert back.shape == (1, rowlen) and np.allclose(back, row, atol=1e-6)
    with pytest.raises(Ima2pError):
        ti_load(tmp_path / "missing.ti", rowlen, lib=emu)
    with pytest.raises(Ima2pError):
        ti_load(path, rowlen + 1, lib=emu)     # a row length that is not the model's


def test_lmode_single_row_and_argument_errors(emu):
    from support import FlatModel, load_golden
    d = load_golden("lmode_extra_sim5_hn2")
    fm = FlatModel(d["model"])
    rows = np.ascontiguousarray(d["rows"], dtype=np.float32)
    lm = LMode(fm.nq, fm.nm, fm.nsplit, fm.q_max, fm.q_min, fm.m_max, fm.m_min, fm.m_mean, fm.expoprior, lib=emu)
    with pytest.raises(Ima2pError):
        lm.moments()                          # nothing loaded yet
    lm.load(rows[:1])
    means, var, corr, raw = lm.moments()
    assert np.all(np.isfinite(means)) and np.allclose(means, raw["sum0"])          # the mean over one row is that row's value
    x = np.array([0.5, 2.0])
    one = lm.margincalc(x, 0.0, 0, 0)
    assert np.all(one >= 0) and np.all(np.isfinite(one))
    assert 0.0 <= lm.greater_than(0, 0, 1) <= 1.0 and lm.greater_than(1, 1, 1) == -1.0
    with pytest.raises(Ima2pError):
        lm.greater_than(0, 0, fm.nq)          # parameter index out of range
    with pytest.raises(Ima2pError):
        lm.popmig(fm.nq, 0, x)
    with pytest.raises(Ima2pError):
        lm.marginpopmig(0, 1, 1, x, 0)        # empty row range
    with pytest.raises(Ima2pError):
        lm.load(rows)                         # rows are loaded once per handle
    lm.close()
    with pytest.raises(Ima2pError):
        LMode(fm.nq, fm.nm, fm.nsplit, fm.q_max, fm.q_min, fm.m_max, fm.m_min, fm.m_mean, fm.expoprior, lib=emu).load(rows[:, :-1])


def test_engine_argument_errors(emu):
    with pytest.raises(Ima2pError):
        Engine(0, 1, lib=emu)
    with pytest.raises(Ima2pError):
        Engine(2, 1, lib=emu, nchains_global=1)
    eng = Engine(1, 1, lib=emu)
    with pytest.raises(Ima2pError):
        eng.run(1)                            # no model, no loci, not finalized
    eng.close()
    eng, _, _, _ = _engine(emu, 2, 2)
    with pytest.raises(Ima2pError):
        eng.set_speculation(5)
    with pytest.raises(Ima2pError):
        eng.pair(0, 2)
    eng.close()


def test_malformed_inputs_are_error_codes_not_crashes(tmp_path):
    """The .u reader and the state loader are fed by files: a negative sequence length, a negative sample size, an HKY locus
    whose columns are all gaps, a genealogy whose links point outside the tree -- every one of them is an error code with a
    message, never an exception through the C boundary or an out-of-range index on the device."""
    import os
    import numpy as np
    from ima2p_b200 import Engine, capi, synth
    from ima2p_b200.readu import read_u
    here = os.path.dirname(os.path.abspath(__file__))
    lib = capi.bind(os.path.join(here, "hostemu", "libima2p_hostemu.so"))
    head = "bad input\n2\npop1 pop2\n(0,1):2\n1\n"
    cases = {"neglen": "loc 2 2 -5 H 1\n", "negsamp": "loc -1 5 4 I 1\n" + "".join("g%-9dACGT\n" % i for i in range(4)),
             "allgaps": "loc 2 2 3 H 1\n" + "".join("g%-9d---\n" % i for i in range(4))}
    for name, body in cases.items():
        p = tmp_path / (name + ".u")
        p.write_text(head + body)
        with pytest.raises(capi.Ima2pError):
            read_u(str(p), lib=lib)
    # a genealogy with a link outside the tree / a tip as root / a migration into a population that does not exist
    loci = synth.make_dataset(1, 3, 3, seed=2)
    eng = Engine(1, 1, lib=lib)
    eng.set_model(**synth.two_population_model())
    eng.set_locus(0, 0, loci[0]["n"], loci[0]["numsites"], loci[0]["samppop"], seq=loci[0]["seq"])
    eng.finalize()
    L = loci[0]
    nl = 2 * L["n"] - 1
    pop = np.r_[np.zeros(3, int), np.ones(3, int), np.full(nl - 6, 2)]
    time = np.where(L["down"] >= 0, 2.0 + L["height"][np.maximum(L["down"], 0)], 1e6)
    good = dict(up0=L["up0"].copy(), up1=L["up1"].copy(), down=L["down"].copy(), pop=pop.copy(), time=time, mig_off=np.zeros(nl + 1, int),
                mig_t=[], mig_p=[], root=nl - 1, roottime=2.0 + L["height"][nl - 1])
    eng.set_genealogy(0, 0, **good)
    for field, idx, val in (("up0", nl - 1, nl + 3), ("down", 0, -7), ("pop", 2, 9)):
        bad = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in good.items()}
        bad[field][idx] = val
        with pytest.raises(capi.Ima2pError):
            eng.set_genealogy(0, 0, **bad)
    bad = dict(good, root=0)
    with pytest.raises(capi.Ima2pError):
        eng.set_genealogy(0, 0, **bad)
    eng.close()
