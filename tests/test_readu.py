"""The .u reader of the C ABI (BASELINE north_star: site-pattern compression bit-exact).

Inputs are corner-case files of our own making (tests/golden/inputs, written by tests/golden/generate.py); the expected
parse is what the unmodified reference made of the same files (readdata, dumped by oracle/_ref/ref_harness).  The
reader is host code of the product library; it needs no GPU, so these tests run through the host-emulation build too."""
import os
import subprocess

import numpy as np
import pytest

from support import load_golden

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "hostemu", "libima2p_hostemu.so")


@pytest.fixture(scope="module")
def lib():
    from ima2p_b200 import capi
    subprocess.run([os.path.join(HERE, "hostemu", "build.sh")], check=True)
    return capi.bind(EMU)


@pytest.mark.parametrize("name", ["parse_is_3pop", "parse_hky", "parse_sw_joint"])
def test_reader_matches_the_reference_parse(lib, name):
    from ima2p_b200.readu import read_u
    ref = load_golden(name)
    got = read_u(os.path.join(HERE, "golden", "inputs", name + ".u"), lib)
    assert got["npops"] == ref["model"]["npops"] and len(got["loci"]) == len(ref["loci"])
    ch = ref["chains"][0]
    for li, (g, r) in enumerate(zip(got["loci"], ref["loci"])):
        assert (g["model"], g["numgenes"], g["numsites"], g["nlinked"]) == (r["model"], r["numgenes"], r["numsites"], r["nlinked"]), li
        assert g["samppop"] == r["samppop"] and g["hval"] == r["hval"] and g["numbases"] == r["numbases"]
        if r["model"] in (0, 1, 3):          # compressed site patterns, bit for bit
            assert np.array_equal(g["seq"], np.array(r["seq"], dtype=np.int32).reshape(r["numgenes"], -1)), li
        if r["model"] == 1:
            assert g["totsites"] == r["totsites"] and np.array_equal(g["mult"], r["mult"])
            assert np.allclose(g["pi"], ch["G"][li]["pi"], rtol=1e-15, atol=0)
        if r["model"] in (2, 3):             # allele lengths: the tips of the reference's starting genealogy carry the data
            first = 1 if r["model"] == 3 else 0
            for ai in range(first, r["nlinked"]):
                assert np.array_equal(g["A"][ai], ch["G"][li]["tree"]["A"][ai][:r["numgenes"]]), (li, ai)


def test_reader_feeds_the_engine_inputs(lib):
    """Reader output -> set_locus: the static evaluation of the reference's starting state from the parsed data equals
    the reference's own values (the fixture was produced from the same file)."""
    from ima2p_b200.readu import read_u
    from support import FlatModel, check_static_eval, engine_from_fixture
    ref = load_golden("parse_is_3pop")
    got = read_u(os.path.join(HERE, "golden", "inputs", "parse_is_3pop.u"), lib)
    for g, r in zip(got["loci"], ref["loci"]):
        r["seq"] = g["seq"].reshape(-1).tolist()          # the engine is loaded from OUR parse
    eng, fm = engine_from_fixture(ref, lib=lib)
    eng.eval()
    check_static_eval(eng, fm, ref, rtol=1e-10)
    eng.close()


def test_reader_rejects_malformed_files(lib, tmp_path):
    from ima2p_b200.capi import Ima2pError
    from ima2p_b200.readu import read_u
    good = open(os.path.join(HERE, "golden", "inputs", "parse_sw_joint.u")).read().split("\n")
    cases = {
        "short.u": good[:12],                                         # file ends inside a locus
        "null_allele.u": [l.replace("s3        ", "s3        0 ") if l.startswith("s3 ") else l for l in good],
        "short_seq.u": [l[:-3] if l.startswith("p4 ") else l for l in good],
        "digit.u": [l[:12] + "7" + l[13:] if l.startswith("p2 ") else l for l in good],
    }
    for fn, lines in cases.items():
        p = tmp_path / fn
        p.write_text("\n".join(lines) + "\n")
        with pytest.raises(Ima2pError):
            read_u(p, lib)
    with pytest.raises(Ima2pError):
        read_u(tmp_path / "missing.u", lib)


def test_ti_file_round_trip_matches_reference_text(lib, tmp_path):
    """.ti rows: the loader reads what the reference's savegenealogyfile wrote; the writer reproduces that text byte for
    byte; L-mode values computed from the loaded file agree with the reference's (which used the in-memory rows)."""
    from ima2p_b200.readu import ti_append, ti_create, ti_load
    ref = load_golden("lmode_ti_sim3_hn2")
    rows = np.array(ref["rows"], dtype=np.float32)
    src = os.path.join(HERE, "golden", "inputs", "sample_sim3.ti")
    got = ti_load(src, rows.shape[1], lib=lib)
    assert got.shape == rows.shape
    assert np.array_equal(got, np.array([[np.float32("%.6f" % v) for v in r] for r in rows]))     # %.6f text of each float
    assert len(ti_load(src, rows.shape[1], max_rows=7, lib=lib)) == 7
    out = tmp_path / "mine.ti"
    ti_create(out, "a header", lib=lib)
    ti_append(out, rows[:25], lib=lib)
    ti_append(out, rows[25:], lib=lib)
    body = lambda p: open(p).read().split("VALUESSTART\n", 1)[1]
    assert body(out) == body(src)
    from ima2p_b200.capi import Ima2pError
    with pytest.raises(Ima2pError):
        ti_load(src, rows.shape[1] + 1, lib=lib)                  # too few values per genealogy
    with pytest.raises(Ima2pError):
        ti_load(src, rows.shape[1] - 1, lib=lib)                  # too many


def test_reader_keeps_the_text_the_report_echoes(lib):
    """ima2p_dataset_text and the header-line flags: what readdata echoes into the report (readata.cpp:916-963, 640-832) --
    the title and '#' lines, the population names, whether the model letter carried a count (SW_M / IS+SW_M) and whether an
    inheritance scalar stood on the line ("%5.3lf" against "%lf")."""
    import ctypes as C

    def texts(d, kind):
        out, buf, k = [], C.create_string_buffer(512), 0
        while lib.ima2p_dataset_text(d, kind, k, buf, 512) == 0:
            out.append(buf.value.decode())
            k += 1
        return out

    def flags(d, nloci):
        info = (C.c_int * 8)()
        res = []
        for li in range(nloci):
            assert lib.ima2p_dataset_locus(d, li, info, None, None, None, 0) == 0
            res.append(info[7])
        return res
    d = C.c_void_p()
    assert lib.ima2p_dataset_read(os.path.join(HERE, "golden", "inputs", "parse_is_3pop.u").encode(), C.byref(d)) == 0
    assert texts(d, 0) == ["parser corner cases, infinite sites", " a comment line", "another"]
    assert texts(d, 1) == ["popA", "popB", "popC"]
    assert flags(d, 4) == [0, 2, 2, 2]                      # loc0 has no inheritance scalar on its line
    lib.ima2p_dataset_free(d)
    d = C.c_void_p()
    assert lib.ima2p_dataset_read(os.path.join(HERE, "golden", "inputs", "parse_sw_joint.u").encode(), C.byref(d)) == 0
    assert flags(d, 3) == [3, 3, 2]                         # S2, J1, I -- all with a scalar
    assert len(texts(d, 1)) == 2 and texts(d, 0)[0]
    lib.ima2p_dataset_free(d)
