"""The .mcf state file through the C ABI, both directions against the reference:

* inputs/<name>.mcf.gz was written by the reference's writemcf; golden/<name>.json.gz is the reference's state after its
  own readmcf of that file (so both sides parse the same text).  The engine must evaluate the file to the same values.
* inputs/<name>_ours.mcf.gz was written by the engine from that state; golden/<name>_ours.json.gz is what the reference
  made of it (readmcf + init_p).  The engine must still write that exact text, and agree with the reference about it.
Runs on the host-emulation build (no GPU needed for file formats); the -m gpu suite repeats the read on the device."""
import gzip
import os
import subprocess

import pytest

from support import check_static_eval, engine_from_fixture, load_golden

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "hostemu", "libima2p_hostemu.so")
NAMES = ["mcf_sim5_hn2", "mcf_sim3_sw_hn2", "mcf_sim5_hky_hn2", "mcf_sim3_joint_hn2"]


def _unpack(name, tmp_path):
    dst = tmp_path / (name + ".mcf")
    with gzip.open(os.path.join(HERE, "golden", "inputs", name + ".mcf.gz"), "rb") as f:
        dst.write_bytes(f.read())
    return dst


def mcf_read_matches_reference(lib, name, tmp_path, rtol=1e-9):
    d = load_golden(name)
    eng, fm = engine_from_fixture(d, lib=lib)           # model and loci from the fixture; the state comes from the file
    eng.read_mcf(_unpack(name, tmp_path))
    check_static_eval(eng, fm, d, rtol=rtol)
    eng.close()


@pytest.fixture(scope="module")
def emu():
    from ima2p_b200 import capi
    subprocess.run([os.path.join(HERE, "hostemu", "build.sh")], check=True)
    return capi.bind(EMU)


@pytest.mark.parametrize("name", NAMES)
def test_reference_written_file_loads_to_the_reference_values(emu, name, tmp_path):
    mcf_read_matches_reference(emu, name, tmp_path)


@pytest.mark.parametrize("name", NAMES)
def test_engine_written_file_is_what_the_reference_read(emu, name, tmp_path):
    d = load_golden(name)
    eng, fm = engine_from_fixture(d, lib=emu)
    eng.eval()
    out = tmp_path / "ours.mcf"
    eng.write_mcf(out)
    assert out.read_bytes() == _unpack(name + "_ours", tmp_path).read_bytes()      # the text the reference was given
    eng.read_mcf(out)
    check_static_eval(eng, fm, load_golden(name + "_ours"), rtol=1e-9)             # and what it made of it
    again = tmp_path / "again.mcf"
    eng.write_mcf(again)
    assert again.read_bytes() == out.read_bytes()                                  # write -> read -> write is a fixed point
    eng.close()


def test_a_short_file_fills_the_remaining_chains_from_its_top(emu, tmp_path):
    # readmcf :415-433: fewer chains in the file than in the run
    d = load_golden("mcf_sim5_hn2")
    one = dict(d, chains=d["chains"][:1])
    eng1, _ = engine_from_fixture(one, lib=emu)
    eng1.eval()
    p = tmp_path / "one.mcf"
    eng1.write_mcf(p)
    eng2, fm = engine_from_fixture(d, lib=emu)
    eng2.read_mcf(p)
    a, b = eng2.chain(0), eng2.chain(1)
    assert a["probg"] == b["probg"] == eng1.chain(0)["probg"] and a["pdg"] == b["pdg"]
    from ima2p_b200.capi import Ima2pError
    bad = tmp_path / "bad.mcf"
    bad.write_text(p.read_text().replace("pop 0 1", "pup 0 1", 1))
    with pytest.raises(Ima2pError):
        eng2.read_mcf(bad)
    eng1.close(); eng2.close()
