"""The command-line front end (csrc/ima_frontend.cpp) over the C ABI, built here against the host-emulation library:
host logic only -- option scan, .u -> model -> starting genealogies -> schedule -> .ti / .mcf / report files."""
import os
import re
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(HERE, "hostemu", "IMa2p_hostemu")
INPUTS = os.path.join(HERE, "golden", "inputs")


@pytest.fixture(scope="module")
def exe():
    subprocess.run([os.path.join(HERE, "hostemu", "build.sh")], check=True)
    return EXE


def _run(exe, args, **kw):
    return subprocess.run([exe] + args, capture_output=True, text=True, timeout=600, **kw)


@pytest.mark.parametrize("name,genes", [("parse_is_3pop", [14] * 4), ("parse_sw_joint", [13] * 3), ("parse_hky", [9, 10, 11])])
def test_run_from_u_file_to_ti_file(exe, tmp_path, name, genes):
    from ima2p_b200 import capi
    from ima2p_b200.readu import read_u, ti_load
    lib = capi.bind(os.path.join(HERE, "hostemu", "libima2p_hostemu.so"))
    out = tmp_path / name
    r = _run(exe, ["-i", os.path.join(INPUTS, name + ".u"), "-o", str(out), "-q10", "-m", "1", "-t3", "-b", "150", "-l", "12", "-d", "5",
                   "-hn3", "-hfl", "-ha", "0.05", "-s", "11", "-r3"])
    assert r.returncode == 0, r.stderr
    d = read_u(os.path.join(INPUTS, name + ".u"), lib)
    npops = d["npops"]
    nq, nm = 2 * npops - 1, {2: 2, 3: 8}[npops]
    rowlen = 4 * nq + 3 * nm + 2 + (npops - 1)
    rows = ti_load(str(out) + ".ti", rowlen, lib=lib)
    assert rows.shape == (12, rowlen)
    cc = rows[:, :nq]
    assert np.all(cc == np.round(cc)) and np.all(cc >= 0)
    assert np.all(cc.sum(axis=1) == sum(g - 1 for g in genes))        # every coalescence is in exactly one population
    mc = rows[:, 3 * nq:3 * nq + nm]
    assert np.all(mc == np.round(mc)) and np.all(mc >= 0)
    t = rows[:, rowlen - (npops - 1):]
    assert np.all(t > 0) and np.all(t < 3) and np.all(np.diff(t, axis=1) > 0)
    rep = open(out).read()
    assert "genealogies saved 12" in rep and "proposals dropped for migration capacity 0" in rep
    # the M-mode report ends with the sections computed over the genealogies it saved (as the reference's printoutput does)
    assert "MEANS, VARIANCES and CORRELATIONS OF PARAMETERS" in rep or npops == 3 and "-j" in rep
    assert "Marginal Peak Locations and Probabilities" in rep and "HISTOGRAM GROUP 2" in rep and rep.rstrip().endswith("END OF OUTPUT")
    mean_line = [ln for ln in rep.split("\n") if ln.startswith("Mean:")][0].split("\t")[1:]
    assert len(mean_line) == nq + nm and all(0.0 < float(v) < 10.0 for v in mean_line[:nq])
    assert os.path.getsize(str(out) + ".mcf") > 1000
    # the state file it wrote restarts a run (-f)
    r2 = _run(exe, ["-i", os.path.join(INPUTS, name + ".u"), "-o", str(tmp_path / "again"), "-q10", "-m1", "-t3", "-b0", "-l3", "-d2", "-hn3", "-hfl",
                    "-ha0.05", "-f", str(out) + ".mcf"])
    assert r2.returncode == 0, r2.stderr
    assert len(ti_load(str(tmp_path / "again") + ".ti", rowlen, lib=lib)) == 3


@pytest.mark.parametrize("name", ["lmode_report_sim3", "lmode_report_3pop"])
def test_l_mode_report_sections_equal_the_reference_text(exe, tmp_path, name):
    """-r0 -v: the genealogies the reference saved (its own .ti file) go through the device evaluators; the greater-than
    tables (-p6), the means / variances / correlations table, the marginal peak table (peak search, 95% bounds, migration
    likelihood-ratio tests, with -p5 the 2NM terms), the histogram group of the size and migration parameters and with -p5
    the histogram group of the 2NM terms must be the reference's text,
    character for character (fixtures: the reference's own L-mode reports of the same files; 2 and 3 populations)."""
    import gzip
    import json
    import shutil
    ref = json.load(gzip.open(os.path.join(HERE, "golden", name + ".json.gz")))
    with gzip.open(os.path.join(INPUTS, name + ".ti.gz"), "rb") as f, open(tmp_path / "ref.ti", "wb") as g:
        shutil.copyfileobj(f, g)
    if name == "lmode_report_3pop":
        u = os.path.join(INPUTS, "parse_is_3pop.u")
    else:
        u = tmp_path / "Sim3.u"
        u.write_text(_sim3_u())
    r = _run(exe, ["-r0", "-v", str(tmp_path / "ref"), "-i", str(u), "-o", str(tmp_path / "l.out"), "-q10", "-m1", "-t3", "-p56", "-c2", "-s", "3"])
    assert r.returncode == 0, r.stderr
    rep = open(tmp_path / "l.out").read()
    for key in ("greater_than", "moments", "peaks", "t_histograms", "histograms", "popmig_histograms"):
        assert ref[key].strip("\n") in rep, key
    if "joint" in ref:
        _same_joint_table(ref["joint"], rep, 1 if name == "lmode_report_sim3" else 2)
    r = _run(exe, ["-r0", "-i", str(u), "-o", str(tmp_path / "l2.out"), "-q10", "-m1", "-t3"])
    assert r.returncode == 8 and "-v" in r.stderr            # IMERR_MISSINGCOMMANDINFO


def _same_joint_table(ref_text, rep, nrows):
    """The joint-posterior peak table (-c2 / -w): a stochastic search (differential evolution, every generation one batched
    device call) with this program's own random numbers, so the comparison is numerical: the peaks the reference found
    (two populations: the FULL model; three: all population sizes, then all migration rates, jointfind.cpp:1118-1133; each
    followed by the nested models of its type).  Model list, columns, #terms, df and brackets must be the reference's text."""
    def peak_rows(text):
        a = text.index("Joint Peak Locations")
        lines = text[a:].split("\n")
        k = [i for i, ln in enumerate(lines) if ln.startswith("Model#\tlog(P)")][0]
        rows = []
        for ln in lines[k + 1:]:
            if not ln.strip() or ln.startswith("    *"):
                break
            rows.append(ln.split("\t"))
        return text[a:a + text[a:].index("Model#\tlog(P)")], lines[k].split("\t"), rows, lines[k + 1 + len(rows)]
    (dr, hr, rr, tr), (do, ho, ro, to) = peak_rows(ref_text), peak_rows(rep)
    assert dr == do and hr == ho and len(rr) == len(ro) == nrows and tr == to     # the model list, the columns, the footnote
    num = lambda t: float(t.strip("[]"))
    for vr, vo in zip(rr, ro):
        assert vr[0] == vo[0] and vr[2:4] == vo[2:4]                      # model number; #terms, df
        assert abs(float(vr[1]) - float(vo[1])) <= 2e-3 * max(1.0, abs(float(vr[1])))      # log(P) at the peak (4 digits printed)
        if vr[4] != "-":
            assert abs(float(vr[4]) - float(vo[4])) <= 5e-3 * max(1.0, abs(float(vr[4])))  # 2LLR against the full model
        else:
            assert vo[4] == "-"
        assert abs(float(vr[5]) - float(vo[5])) <= 0.02 * float(vr[5])    # effective sample size there
        for a, b in zip(vr[6:], vo[6:]):
            assert (a == "-") == (b == "-") and a.startswith("[") == b.startswith("["), (vr, vo)     # outside the model; tied / fixed
            if a != "-":
                assert abs(num(a) - num(b)) <= 2e-3 * max(1.0, abs(num(a))), (vr, vo)


@pytest.mark.parametrize("name", ["lmode_report_sim3", "lmode_report_3pop"])
def test_nested_models_of_the_joint_search(exe, tmp_path, name):
    """-w <nested model file> (jointfind.cpp:40-135, setup_mapping :380-543): parameters tied together ('equal') or fixed
    ('constant', clamped into the prior, boundary footnote); every nested model is searched over its free parameters and
    compared with the full model of its type.  Fixtures: the reference's own tables for the committed .ti and model files."""
    import gzip
    import json
    import shutil
    ref = json.load(gzip.open(os.path.join(HERE, "golden", name + ".json.gz")))
    with gzip.open(os.path.join(INPUTS, name + ".ti.gz"), "rb") as f, open(tmp_path / "ref.ti", "wb") as g:
        shutil.copyfileobj(f, g)
    if name == "lmode_report_3pop":
        u, nest, nrows = os.path.join(INPUTS, "parse_is_3pop.u"), os.path.join(INPUTS, "nested_models_3pop.txt"), 4
    else:
        u, nest, nrows = tmp_path / "Sim3.u", os.path.join(INPUTS, "nested_models_2pop.txt"), 3
        u.write_text(_sim3_u())
    r = _run(exe, ["-r0", "-v", str(tmp_path / "ref"), "-i", str(u), "-o", str(tmp_path / "l.out"), "-q10", "-m1", "-t3", "-W", nest, "-s", "3"])
    assert r.returncode == 0, r.stderr
    rep = open(tmp_path / "l.out").read().replace(nest, "NESTEDFILE")
    _same_joint_table(ref["joint_nested"], rep, nrows)
    bad = tmp_path / "bad.txt"
    bad.write_text("1\nmodel nothing tied\n")
    r = _run(exe, ["-r0", "-v", str(tmp_path / "ref"), "-i", str(u), "-o", str(tmp_path / "l2.out"), "-q10", "-m1", "-t3", "-w", str(bad)])
    assert r.returncode == 16 and "does not include any" in r.stderr          # IMERR_NESTEDMODELLSPECIFYLFAIL


def _sim3_u():
    """A .u file with the shape of Simulations/Sim3.u (2 populations, 2 infinite-sites loci of 15 + 15 genes): L mode reads the
    data file only for the population tree and the number of loci."""
    rows = []
    for l, ns in enumerate((16, 21)):
        rows.append("locus%d 15 15 %d I 1" % (l, ns))
        for g in range(30):
            rows.append("%-10s%s" % ("g%d" % g, "".join("AC"[(g >> (s % 4)) & 1] if s < 4 else "A" if g else "C" for s in range(ns))))
    return "synthetic\n2\npop0 pop1\n(0,1):2\n2\n" + "\n".join(rows) + "\n"


def test_refuses_what_it_does_not_implement(exe, tmp_path):
    u = os.path.join(INPUTS, "parse_hky.u")
    for bad in (["-a1"], ["-j3"], ["-c0"]):
        r = _run(exe, ["-i", u, "-o", str(tmp_path / "x"), "-q10", "-m1", "-t3", "-b1", "-l1"] + bad)
        assert r.returncode != 0 and r.stderr.startswith("IMa2:")
    r = _run(exe, ["-i", str(tmp_path / "nothere.u"), "-o", str(tmp_path / "x"), "-q10", "-m1", "-t3", "-b1", "-l1"])
    assert r.returncode != 0 and "can't be opened" in r.stderr
    r = _run(exe, ["-i", u, "-q10", "-m1", "-t3", "-b1", "-l1"])
    assert r.returncode != 0 and "-o is required" in r.stderr


def frontend_posterior_matches_reference(exe, lib, tmp_path, nsigma=5.0):
    """A run of the front end from the .u file alone (its own model tables, starting genealogies and schedule) samples the
    posterior the reference samples with whole qupdate() steps on the same file: split times, coalescences per sampled
    population and migration events (fixture trace_full_parse_is_3pop, written by the reference)."""
    from ima2p_b200.readu import ti_load
    from support import load_golden
    out = tmp_path / "long"
    r = _run(exe, ["-i", os.path.join(INPUTS, "parse_is_3pop.u"), "-o", str(out), "-q10", "-m1", "-t3", "-b10000", "-l3000", "-d20", "-hn1", "-s5"])
    assert r.returncode == 0, r.stderr
    rows = ti_load(str(out) + ".ti", 4 * 5 + 3 * 8 + 2 + 2, lib=lib)
    d = load_golden("trace_full_parse_is_3pop")
    t, b = np.array(d["t_batch_means"]), np.array(d["batch_means"]).sum(axis=1)
    pairs = {"t0": (rows[:, -2], t[:, 0]), "t1": (rows[:, -1], t[:, 1]), "cc0": (rows[:, 0], b[:, 3]), "cc1": (rows[:, 1], b[:, 4]),
             "migrations": (rows[:, 15:23].sum(axis=1), b[:, 2])}
    for k, (mine, ref) in pairs.items():
        bm = mine[:3000].reshape(30, -1).mean(axis=1)                     # batch means of the correlated series
        z = (bm.mean() - ref.mean()) / np.hypot(bm.std(ddof=1) / np.sqrt(30), ref.std(ddof=1) / np.sqrt(len(ref)))
        assert abs(z) < nsigma, (k, bm.mean(), ref.mean(), z)


def test_front_end_samples_the_reference_posterior(exe, tmp_path):
    from ima2p_b200 import capi
    frontend_posterior_matches_reference(exe, capi.bind(os.path.join(HERE, "hostemu", "libima2p_hostemu.so")), tmp_path)


def _rate_tables(text):
    """{table title: {row name: [(tries, accepts, percent), ...]}} of the 'Update Rates -- ...' tables, and the swap table as
    [(temp1, temp2, swaps, attempts, rate), ...]."""
    tables, swaps, title, in_swaps = {}, [], None, False
    for ln in text.split("\n"):
        if ln.startswith("Update Rates -- "):
            title, in_swaps = ln[len("Update Rates -- "):], False
            tables[title] = {}
        elif ln.startswith("Temp1"):
            title, in_swaps = None, True
        elif in_swaps and ln.strip():
            w = ln.split()
            if len(w) == 5:
                swaps.append((float(w[0]), float(w[1]), int(w[2]), int(w[3]), float(w[4])))
        elif title and ln.startswith(" ") and "\t" in ln and not ln.strip().startswith("#"):
            w = ln.split("\t")
            vals = [(float(w[i]), float(w[i + 1]), float(w[i + 2])) for i in range(1, len(w) - 2, 3)]
            tables[title][w[0].strip()] = vals
    return tables, swaps


def frontend_report_head_matches_reference(exe, tmp_path, lib=None):
    """The opening sections of an M-mode report: the run information the reference echoes is the
    reference's text line for line; the update-rate tables of the cold chain (split times by update type, genealogies per
    locus by branch / topology / tmrca, mutation scalars) and the swap table between adjacent temperatures have the
    reference's layout, and their rates agree with the reference's own run of the same command (fixture
    report_head_3pop, written by the reference's unmodified main(); 60,000 steps after the burn-in)."""
    import gzip
    import json
    d = json.load(gzip.open(os.path.join(HERE, "golden", "report_head_3pop.json.gz")))
    outs = []
    for seed in (21, 22, 23):            # three runs: the rates are compared as means over runs (see check_report_rates)
        out = tmp_path / ("head%d.out" % seed)
        r = _run(exe, ["-i", os.path.join(INPUTS, "parse_is_3pop.u"), "-o", str(out)] + d["args"] + ["-s", str(seed)])
        assert r.returncode == 0, r.stderr
        outs.append(str(out))
    check_report_head(outs[0], d["head"], lib)
    check_report_rates(outs)


def check_report_rates(outs):
    """Update-rate tables and swap table of several runs of the front end against the reference's own runs of the same command
    from eight seeds (fixture report_rates_3pop: `ref_harness stock seed=K`).  Every cell is compared as mean over our runs
    against mean over the reference's runs.  Tolerances in percentage points: genealogies 2.5 and mutation scalars 1.0 (the
    reference's run-to-run standard deviation is at most 1.04 and 0.25 there, so the difference of the means has a standard
    deviation of 0.7 and 0.17); split times, which mix slowly (run-to-run standard deviation 1.2 to 3.4 points), four standard
    deviations of the difference of the means.  profiles/r2_split_time_replicates.md holds the 12-seed study behind this."""
    import gzip
    import json
    runs = json.load(gzip.open(os.path.join(HERE, "golden", "report_rates_3pop.json.gz")))["runs"]
    mine = [_rate_tables(open(o).read()[:open(o).read().index("\nENGINE INFORMATION")]) for o in outs]
    nr, nm = len(runs), len(mine)
    for title in ("Population Splitting Times", "Genealogies", "Mutation Rate Scalars"):
        assert list(mine[0][0][title]) == list(runs[0]["tables"][title])
        for row in runs[0]["tables"][title]:
            ref = np.array([[c[2] for c in r["tables"][title][row]] for r in runs])          # [run][update type] percent
            got = np.array([[c[2] for c in m[0][title][row]] for m in mine])
            sd = ref.std(axis=0, ddof=1) * np.sqrt(1.0 / nr + 1.0 / nm)
            tol = {"Genealogies": np.full(ref.shape[1], 2.5), "Mutation Rate Scalars": np.full(ref.shape[1], 1.0),
                   "Population Splitting Times": np.maximum(2.5, 4.0 * sd)}[title]
            dev = np.abs(got.mean(axis=0) - ref.mean(axis=0))
            assert np.all(dev < tol), (title, row, got.mean(axis=0), ref.mean(axis=0), tol)
    ref_sw = np.array([[s[4] for s in r["swaps"]] for r in runs])
    got_sw = np.array([[s[4] for s in m[1]] for m in mine])
    assert got_sw.shape[1] == ref_sw.shape[1] == 3
    # the swap rates follow the split times (slow mixing): three runs of the front end spread by 0.07 on the device
    assert np.all(np.abs(got_sw.mean(axis=0) - ref_sw.mean(axis=0)) < 0.08), (got_sw.mean(axis=0), ref_sw.mean(axis=0))


def check_report_head(out, ref, lib=None):
    """The text checks of frontend_report_head_matches_reference on a report file `out` (and its .ti file) already written;
    the rates are compared by check_report_rates."""
    import re
    rep = open(out).read()
    mine = rep[:rep.index("\nENGINE INFORMATION")]

    def between(text, a, b):
        i = text.index(a)
        return text[i:text.index(b, i)]
    # what readdata / setup echo: identical text (file names and the seed line differ, they are above this block)
    assert between(mine, "- Run Duration -", "All genealogy information") == between(ref, "- Run Duration -", "All genealogy information")
    assert between(mine, "\nText from input file", "\n\nMCMC INFORMATION") == between(ref, "\nText from input file", "\n\nMCMC INFORMATION")
    assert between(mine, "\nMCMC INFORMATION", "\nTime Elapsed") == between(ref, "\nMCMC INFORMATION", "\nTime Elapsed")
    # same skeleton from the highest likelihoods to the end of the swap table (digits blanked; the serial build's per-chain
    # swap table has no counterpart when only temperatures move)
    def skeleton(text):
        t = between(text + "\n\nEND", "Highest Sampled Joint", "\n\nEND")
        t = re.sub(r"Chain    #Swaps    Rate\n(?: +\d+ +\d+ +[\d.]+\n)+\n", "", t)
        return re.sub(r"\s+", " ", re.sub(r"-?\d[\d.e]*", "#", t)).strip()
    assert skeleton(mine) == skeleton(ref)
    (tm, sm), (tr, sr) = _rate_tables(mine), _rate_tables(ref)
    assert list(tm) == list(tr) == ["Population Splitting Times", "Genealogies", "Mutation Rate Scalars"]
    # same rows, same numbers of tries (the rates themselves: check_report_rates, means over several runs of both programs)
    for title in tr:
        assert list(tm[title]) == list(tr[title])
        for row in tr[title]:
            for (t1, a1, p1), (t2, a2, p2) in zip(tm[title][row], tr[title][row]):
                assert abs(t1 - t2) <= 0.06 * t2, (title, row, t1, t2)
    assert [g[0][0] for g in tm["Genealogies"].values()] == [6.0e4] * 4                 # every step tries every locus of the cold chain
    assert len(sm) == len(sr) == 3
    for a, b in zip(sm, sr):
        assert a[:2] == b[:2] and a[3] > 0 and abs(a[4] - b[4]) < 0.08, (a, b)
    # highest likelihoods: maxima of the recorded cold-chain values, i.e. of the P(D|G) and P(G) columns of the .ti rows.  (The
    # reference's own figures are not comparable run to run: it looks at the highs at its print intervals only, and its
    # infinite-sites constant, calc_sumlogk, depends on the random genealogy a run starts with; the front end applies the same
    # rule to its own starting genealogy.)
    from ima2p_b200 import capi
    from ima2p_b200.readu import ti_load
    rows = ti_load(str(out) + ".ti", 4 * 5 + 3 * 8 + 2 + 2, lib=capi.bind(os.path.join(HERE, "hostemu", "libima2p_hostemu.so")) if lib is None else lib)
    hm = [float(x) for x in re.findall(r"\(log\) :\s+(-?[\d.]+)", mine)]
    assert len(hm) == 2 and abs(hm[0] - rows[:, -3].max()) < 2e-3 and abs(hm[1] - rows[:, -4].max()) < 2e-3, (hm, rows[:, -3].max(), rows[:, -4].max())
    per_locus = [float(x) for x in re.findall(r"\n\t\d+\t(-?[\d.]+)", between(mine, "Highest P(D|G) (log) for each Locus", "Update Rates"))]
    assert len(per_locus) == 4 and all(v < 0 for v in per_locus) and sum(per_locus) >= hm[1] - 2e-3


def test_report_head_matches_reference(exe, tmp_path):
    frontend_report_head_matches_reference(exe, tmp_path)


def test_l_mode_ascii_curves_equal_the_reference_text(exe, tmp_path):
    """The 'ASCII Curves - Approximate Posterior Densities' section (asciicurve, output.cpp:428-524, over callasciicurves,
    ima_main_mpi.cpp:3905-4000) of an L-mode report: population sizes, migration rates and the split time, drawn from the
    device's marginal densities, against the reference's own L-mode report of the same .ti file.  Character for character,
    except the cell of a curve's single highest point: (int)(50 * y / ymax) at y == ymax is 50 or 49 depending on the last bit
    of y, and the densities agree with the reference's to 1e-12, not to the bit; that point may sit one row lower."""
    import gzip
    import json
    import shutil
    ref = json.load(gzip.open(os.path.join(HERE, "golden", "lmode_ascii_sim3.json.gz")))["curves"]
    with gzip.open(os.path.join(INPUTS, "lmode_report_sim3.ti.gz"), "rb") as f, open(tmp_path / "ref.ti", "wb") as g:
        shutil.copyfileobj(f, g)
    u = tmp_path / "Sim3.u"
    u.write_text(_sim3_u())
    r = _run(exe, ["-r0", "-v", str(tmp_path / "ref"), "-i", str(u), "-o", str(tmp_path / "l.out"), "-q10", "-m1", "-t3"])
    assert r.returncode == 0, r.stderr
    rep = open(tmp_path / "l.out").read()
    a = rep.index("ASCII Curves - Approximate Posterior Densities")
    mine = rep[a:rep.index("\nTime Elapsed", a)].rstrip("\n").split("\n")
    assert rep.rstrip().endswith("END OF OUTPUT")
    want = ref.rstrip("\n").split("\n")
    assert len(mine) == len(want) == 2 + 6 * 54 - 1
    differing = [i for i, (x, y) in enumerate(zip(mine, want)) if x != y]
    # only top rows of plots (line 1 of each 54-line block after the two header lines) may differ, and then only in the one cell
    for i in differing:
        assert (i - 2) % 54 == 1, (i, mine[i], want[i])
        assert mine[i][:11] == want[i][:11] and sorted((mine[i][11:].count("*"), want[i][11:].count("*"))) == [0, 1], (mine[i], want[i])
    assert len(differing) <= 3


def test_migration_pools_grow_and_nothing_is_lost_silently(exe, tmp_path):
    """A wide migration prior (-m 10) on Sim3-shaped data with pools that start far too small (-cap 8): the front end must notice
    every proposal that did not fit, say so on stderr, double the pools (ima2p_engine_grow_capacity; the reference's checkmig grows
    an edge's list the same way, utilities.cpp:1365-1383) and finish; started with room to spare (-cap 512) the same command
    reports no proposal lost."""
    from ima2p_b200 import synth
    u = tmp_path / "Sim3.u"
    synth.write_u(str(u), synth.make_dataset(2, 15, 15, seed=3))       # two infinite-sites loci of 15 + 15 genes, like Simulations/Sim3.u
    common = ["-i", str(u), "-q10", "-m10", "-t3", "-b3000", "-l60", "-d20", "-hn2", "-hfl", "-ha0.3", "-s7"]
    r = _run(exe, common + ["-o", str(tmp_path / "tight.out"), "-cap", "8"])
    assert r.returncode == 0, r.stderr
    assert "rejected unseen" in r.stderr and "the pools now hold" in r.stderr, r.stderr[-600:]
    lost = int(re.search(r"proposals dropped for migration capacity (\d+)", open(tmp_path / "tight.out").read()).group(1))
    assert lost > 0
    r = _run(exe, common + ["-o", str(tmp_path / "roomy.out"), "-cap", "512"])
    assert r.returncode == 0 and "rejected unseen" not in r.stderr, r.stderr[-600:]
    assert int(re.search(r"proposals dropped for migration capacity (\d+)", open(tmp_path / "roomy.out").read()).group(1)) == 0
    # many migration events were indeed sampled: the .ti rows' migration counts (columns mc0, mc1 of 21)
    from ima2p_b200 import capi
    from ima2p_b200.readu import ti_load
    rows = ti_load(str(tmp_path / "roomy.out") + ".ti", 21, lib=capi.bind(os.path.join(HERE, "hostemu", "libima2p_hostemu.so")))
    assert rows[:, 9:11].sum(axis=1).mean() > 8


def test_two_ranks_write_the_single_rank_ti_file(exe, tmp_path):
    """The front end started once per rank (RANK / WORLD_SIZE in the environment, -hn chains per rank, as `mpirun -np 2 IMa2p -hn 2`
    starts the reference, ima_main_mpi.cpp:4317-4560): the ranks find each other through the rendezvous files, exchange swap sums
    and the cold chain's record through each other's tables, and rank 0 writes the .ti file -- the file one rank with all four
    chains writes from the same seed, byte for byte; the update-rate tables add up to the same counts."""
    from ima2p_b200 import synth
    u = tmp_path / "two.u"
    synth.write_u(str(u), synth.make_dataset(3, 6, 6, seed=5))
    common = ["-i", str(u), "-q10", "-m1", "-t3", "-b400", "-l40", "-d10", "-hfg", "-ha0.9", "-hb0.8", "-s11"]
    r1 = _run(exe, common + ["-hn4", "-o", str(tmp_path / "one.out")])
    assert r1.returncode == 0, r1.stderr
    env = dict(os.environ, WORLD_SIZE="2", MASTER_PORT="29999", IMA2P_RENDEZVOUS_DIR=str(tmp_path), IMA2P_EMU_WAIT_MS="60000")
    procs = [subprocess.Popen([exe] + common + ["-hn2", "-o", str(tmp_path / "two.out")], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600) for p in procs]
    assert all(p.returncode == 0 for p in procs), [o[1][-400:] for o in outs]
    body = lambda path: open(path).read().split("VALUESSTART", 1)[1]
    assert body(tmp_path / "one.out.ti") == body(tmp_path / "two.out.ti")
    (t1, s1), (t2, s2) = _rate_tables(open(tmp_path / "one.out").read().split("\nENGINE INFORMATION")[0]), _rate_tables(open(tmp_path / "two.out").read().split("\nENGINE INFORMATION")[0])
    assert t1 == t2 and s1 == s2
    assert not [f for f in os.listdir(tmp_path) if f.startswith("ima2p_")]         # the rendezvous files are gone
