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
