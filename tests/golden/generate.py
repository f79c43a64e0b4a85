#!/usr/bin/env python
"""Regenerate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference and `make -C oracle ref`):

    python tests/golden/generate.py

Every fixture is the output of oracle/_ref/ref_harness (the reference's own code, see
oracle/ref_harness/harness_main.cpp) on the reference's own Simulations/*.u inputs or on inputs
derived from them by the rules of SURVEY.md section 8c/8d (HKY by relabelling the model letter,
synthetic stepwise alleles, a 3-population split of the same samples).  Nothing under tests/,
bench.py or smoke() reads /root/reference at run time -- they read these files.
"""
import gzip
import os
import random
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("IMA2P_REFERENCE", "/root/reference")
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
TMP = "/tmp/ima2p_golden"

PRIORS = ["-q", "10", "-m", "1", "-t", "3"]
HEAT = ["-hfg", "-ha", "0.96", "-hb", "0.9"]       # geometric needs >= 4 chains (ima_main_mpi.cpp:1187)
HEAT_LINEAR = ["-hfl", "-ha", "0.05"]
COMMON = ["-b", "100", "-l", "100", "-p01", "-z", "100000000"]


ONLY = set(sys.argv[1:])        # optional fixture names: regenerate just those


def run(mode, name, ufile, hn, kv, extra=(), priors=None):
    if ONLY and name not in ONLY:
        return
    out = os.path.join(TMP, name + ".json")
    cmd = [HARNESS, mode, out] + ["%s=%s" % p for p in kv.items()] + ["--", "-i", ufile, "-o",
          os.path.join(TMP, name + ".out")] + (priors or PRIORS) + COMMON + ["-hn", str(hn)] + (HEAT if hn >= 4 else HEAT_LINEAR if hn > 1 else []) + list(extra)
    with open(os.path.join(TMP, name + ".log"), "w") as log:
        subprocess.run(cmd, check=True, stdout=log, stderr=subprocess.STDOUT, cwd=TMP)
    with open(out, "rb") as f, gzip.GzipFile(os.path.join(HERE, name + ".json.gz"), "wb", mtime=0) as g:
        shutil.copyfileobj(f, g)
    print("wrote", name + ".json.gz", os.path.getsize(os.path.join(HERE, name + ".json.gz")), "bytes")


def run_trace(name, ufile, seeds, kv, priors=None):
    """Long-run summary statistics of the reference's updategenealogy sampler (split times and mutation scalars held
    at their start values): several independently seeded runs merged into one fixture."""
    if ONLY and name not in ONLY:
        return
    import json
    merged = None
    for sd in seeds:
        out = os.path.join(TMP, "%s_%d.json" % (name, sd))
        cmd = [HARNESS, "trace", out, "seed=%d" % sd] + ["%s=%s" % p for p in kv.items()] + ["--", "-i", ufile, "-o",
              os.path.join(TMP, name + ".out")] + (priors or PRIORS) + COMMON + ["-hn", "1"]
        try:    # the reference itself occasionally spins forever for some seeds; such a run is dropped
            with open(os.path.join(TMP, name + ".log"), "w") as log:
                subprocess.run(cmd, check=True, stdout=log, stderr=subprocess.STDOUT, cwd=TMP, timeout=120)
        except subprocess.TimeoutExpired:
            print("seed", sd, "did not finish (reference hang); skipped")
            continue
        d = json.load(open(out))
        if merged is None:
            merged = d
            merged["accept"] = [d["accept"]]
        else:
            assert d["tvals"] == merged["tvals"] and d["uvals"] == merged["uvals"]
            merged["batch_means"] += d["batch_means"]
            merged["t_batch_means"] += d["t_batch_means"]
            merged["logu_batch_means"] += d["logu_batch_means"]
            merged["accept"].append(d["accept"])
    with gzip.GzipFile(os.path.join(HERE, name + ".json.gz"), "wb", mtime=0) as g:
        g.write(json.dumps(merged).encode())
    print("wrote", name + ".json.gz", os.path.getsize(os.path.join(HERE, name + ".json.gz")), "bytes")


def relabel(src, dst, fn):
    """Copy a .u file, passing each locus header line through fn(fields) -> fields."""
    lines = open(src).read().split("\n")
    out, i = [], 0
    out.append(lines[0]); i = 1
    while lines[i].startswith("#"):
        out.append(lines[i]); i += 1
    npops = int(lines[i].split()[0]); out.append(lines[i]); i += 1
    out.append(lines[i]); i += 1          # population names
    out.append(lines[i]); i += 1          # tree string
    nloci = int(lines[i].split()[0]); out.append(lines[i]); i += 1
    for _ in range(nloci):
        f = lines[i].split(); i += 1
        n = sum(int(x) for x in f[1:1 + npops])
        rows = lines[i:i + n]; i += n
        f, rows = fn(f, rows, npops)
        out.append(" ".join(f)); out.extend(rows)
    open(dst, "w").write("\n".join(out) + "\n")


def to_hky(f, rows, npops):
    f = list(f); f[2 + npops] = "H"; return f, rows


def to_sw(f, rows, npops, rng=random.Random(1)):
    # synthetic S1 locus: one allele column, uniform int in [10,16] (> MINSTRLENGTH 3, imamp.hpp:144)
    f = list(f); f[1 + npops] = "1"; f[2 + npops] = "S1"
    rows = ["%-10s%d" % (r[:10].strip(), rng.randint(10, 16)) for r in rows]
    return f, rows


def to_joint(f, rows, npops, rng=random.Random(2)):
    # J1 locus: the infinite-sites columns as they are plus one stepwise allele column in front (readata.cpp:93-97, 428-440)
    f = list(f); f[2 + npops] = "J1"
    rows = ["%-10s%d %s" % (r[:10].strip(), rng.randint(10, 16), r[10:].strip()) for r in rows]
    return f, rows


def three_pops(src, dst):
    """Same samples as `src`, re-divided into 3 populations with tree ((0,1):3,2):4."""
    lines = open(src).read().split("\n")
    out, i = [lines[0]], 1
    while lines[i].startswith("#"):
        out.append(lines[i]); i += 1
    i += 3
    out += ["3", "pop1 pop2 pop3", "((0,1):3,2):4"]
    nloci = int(lines[i].split()[0]); out.append(lines[i]); i += 1
    for _ in range(nloci):
        f = lines[i].split(); i += 1
        n0, n1 = int(f[1]), int(f[2])
        rows = lines[i:i + n0 + n1]; i += n0 + n1
        a = n0 // 2
        out.append(" ".join([f[0], str(a), str(n0 - a), str(n1)] + f[3:]))
        out.extend(rows)
    open(dst, "w").write("\n".join(out) + "\n")


def make_parse_inputs():
    """Small .u files of our own making that exercise the reader's corner cases (kept columns vs monomorphic, three-state
    and non-acgt columns; gaps and repeated columns under HKY, interleaved blocks; linked stepwise parts; a joint locus;
    the optional tree line; comment lines; blanks inside sequences).  The infinite-sites columns come from genealogies,
    so that the reference accepts the files; the expected parse is then the reference's own (state fixtures below)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    import numpy as np
    from ima2p_b200 import synth
    rng = np.random.default_rng(5)
    out = os.path.join(HERE, "inputs")
    os.makedirs(out, exist_ok=True)

    def is_rows(L, junk=True):
        n, cols = L["n"], []
        for s in range(L["numsites"]):
            anc, der = rng.choice(list("ACGT"), 2, replace=False)
            cols.append([der if L["seq"][j][s] else anc for j in range(n)])
            if junk and s % 3 == 0:                      # a monomorphic column
                cols.append([str(rng.choice(list("ACGT")))] * n)
            if junk and s % 5 == 1:                      # three states: dropped
                c = [str(x) for x in rng.choice(list("ACG"), n)]
                c[0], c[1], c[2] = "A", "C", "G"
                cols.append(c)
            if junk and s % 7 == 2:                      # an unknown base: dropped
                c = ["T"] * n
                c[int(rng.integers(1, n))] = "N"
                c[int(rng.integers(1, n))] = "C"
                cols.append(c)
        rows = ["".join(cols[s][j] for s in range(len(cols))) for j in range(n)]
        return rows, len(cols)

    def case(rows):                                      # lower case on odd rows, a blank in the middle of even rows
        return [r.lower() if j % 2 else r[:len(r) // 2] + " " + r[len(r) // 2:] for j, r in enumerate(rows)]

    # 1: infinite sites, 3 populations, tree line, comments, inheritance scalars, a mutation rate
    loci = synth.make_dataset(4, 6, 8, seed=21, min_sites=5)
    with open(os.path.join(out, "parse_is_3pop.u"), "w") as f:
        f.write("parser corner cases, infinite sites\n# a comment line\n#another\n3\npopA popB popC\n((0,1):3,2):4\n4\n")
        for k, L in enumerate(loci):
            rows, nb = is_rows(L)
            tail = ["", " 0.75", " 0.25 0.0000012", " 1"][k]
            f.write("loc%d 6 5 3 %d I%s\n" % (k, nb, tail))
            for j, r in enumerate(case(rows)):
                f.write("%-10s%s\n" % ("g%d_%d" % (k, j), r))
    # 2: HKY, 2 populations; gaps, repeated columns, u for t, an interleaved locus
    with open(os.path.join(out, "parse_hky.u"), "w") as f:
        f.write("parser corner cases, HKY\n2\npop0 pop1\n(0,1):2\n3\n")
        for k in range(3):
            n0, n1, nb = 5, 4 + k, 40 + 7 * k
            base = rng.choice(list("ACGT"), nb)
            rows = []
            for j in range(n0 + n1):
                r = base.copy()
                mut = rng.random(nb) < 0.08
                r[mut] = rng.choice(list("ACGT"), mut.sum())
                gap = rng.random(nb) < 0.02
                r[gap] = rng.choice(list("N-."), gap.sum())
                r = "".join(r)
                rows.append(r.replace("T", "U") if j == 2 else (r.lower() if j % 3 == 1 else r))
            f.write("hloc%d %d %d %d H %s\n" % (k, n0, n1, nb, ["1", "0.5", "1"][k]))
            if k == 1:                                   # two interleaved blocks
                cut = nb // 2
                for j, r in enumerate(rows):
                    f.write("%-10s%s\n" % ("h%d" % j, r[:cut]))
                for j, r in enumerate(rows):
                    f.write("%-10s%s\n" % ("h%d" % j, r[cut:]))
            else:
                for j, r in enumerate(rows):
                    f.write("%-10s%s %s\n" % ("h%d" % j, r[:10], r[10:]))
    # 3: stepwise with two linked parts, a joint locus, an infinite-sites locus; 2 populations with the tree line
    loci = synth.make_dataset(2, 7, 6, seed=23, min_sites=6)
    with open(os.path.join(out, "parse_sw_joint.u"), "w") as f:
        f.write("parser corner cases, stepwise and joint\n#c\n2\nwest east\n(0,1):2\n3\n")
        f.write("str2 7 6 2 S2 1\n")
        for j in range(13):
            f.write("%-10s%d %d\n" % ("s%d" % j, rng.integers(8, 15), rng.integers(20, 31)))
        rows, nb = is_rows(loci[0])
        f.write("joint 7 6 %d J1 0.5\n" % nb)
        for j, r in enumerate(rows):
            f.write("%-10s%d %s\n" % ("j%d" % j, rng.integers(9, 14), r))
        rows, nb = is_rows(loci[1], junk=False)
        f.write("plain 7 6 %d I 1\n" % nb)
        for j, r in enumerate(rows):
            f.write("%-10s%s\n" % ("p%d" % j, r))
    return [os.path.join(out, n) for n in ("parse_is_3pop.u", "parse_hky.u", "parse_sw_joint.u")]


def four_pops(src, dst, tree):
    """Same samples as `src` (two populations of 10), re-divided into 4 populations of 5 with the given tree."""
    lines = open(src).read().split("\n")
    out, i = [lines[0]], 1
    while lines[i].startswith("#"):
        out.append(lines[i]); i += 1
    i += 3
    out += ["4", "popA popB popC popD", tree]
    nloci = int(lines[i].split()[0]); out.append(lines[i]); i += 1
    for _ in range(nloci):
        f = lines[i].split(); i += 1
        n0, n1 = int(f[1]), int(f[2])
        rows = lines[i:i + n0 + n1]; i += n0 + n1
        out.append(" ".join([f[0], str(n0 // 2), str(n0 - n0 // 2), str(n1 // 2), str(n1 - n1 // 2)] + f[3:]))
        out.extend(rows)
    open(dst, "w").write("\n".join(out) + "\n")


def main():
    if not os.path.exists(HARNESS):
        sys.exit("build the reference harness first: make -C oracle ref")
    os.makedirs(TMP, exist_ok=True)
    sims = os.path.join(REF, "Simulations")
    s5, s50, s300 = (os.path.join(sims, "Sim1_%dloci.u" % k) for k in (5, 50, 300))
    s2, s3 = os.path.join(sims, "Sim2.u"), os.path.join(sims, "Sim3.u")
    hky5 = os.path.join(TMP, "Sim1_5loci_HKY.u"); relabel(s5, hky5, to_hky)
    sw3 = os.path.join(TMP, "Sim3_SW.u"); relabel(s3, sw3, to_sw)
    p3 = os.path.join(TMP, "Sim1_5loci_3pop.u"); three_pops(s5, p3)
    j3 = os.path.join(TMP, "Sim3_JOINT.u"); relabel(s3, j3, to_joint)

    # static-evaluation fixtures (a5, a6, a7, a8, a9, a13): states after a short burn-in
    run("state", "state_sim5_hn4", s5, 4, {"burn": 200})                 # BASELINE config 1
    run("state", "state_sim50_hn3", s50, 3, {"burn": 40})                # config 2 shape, 3 chains
    run("state", "state_sim300_hn1", s300, 1, {"burn": 10})              # config 3 shape, 1 chain
    run("state", "state_sim2_hn2", s2, 2, {"burn": 60})                  # n = 100 genes
    run("state", "state_sim3_hn3", s3, 3, {"burn": 300})
    run("state", "state_sim5_3pop_hn2", p3, 2, {"burn": 150})
    run("state", "state_sim5_expo_hn2", s5, 2, {"burn": 100}, extra=["-j7"])   # exponential m prior
    run("state", "state_sim5_hky_hn2", hky5, 2, {"burn": 60})
    run("state", "state_sim3_sw_hn2", sw3, 2, {"burn": 100})
    run("state", "state_sim3_joint_hn2", j3, 2, {"burn": 100})
    p4a = os.path.join(TMP, "Sim1_5loci_4popA.u"); four_pops(s5, p4a, "((0,1):4,(2,3):5):6")
    p4b = os.path.join(TMP, "Sim1_5loci_4popB.u"); four_pops(s5, p4b, "(((0,1):4,2):5,3):6")
    run("state", "state_sim5_4popA_hn2", p4a, 2, {"burn": 100})
    run("state", "state_sim5_4popB_hn2", p4b, 2, {"burn": 100})
    nomig = ["-q", "10", "-m", "0", "-t", "3"]                       # -m 0 sets NOMIGRATION (ima_main_mpi.cpp:822-823)
    run("state", "state_sim5_nomig_hn2", s5, 2, {"burn": 100}, priors=nomig)
    run("state", "state_sim5_3pop_nomig_hn2", p3, 2, {"burn": 100}, priors=nomig)
    # the .u reader (section 8 f2): our own corner-case inputs, parsed by the reference
    if not ONLY or any(n.startswith("parse_") for n in ONLY):
        u1, u2, u3 = make_parse_inputs()
        run("state", "parse_is_3pop", u1, 1, {"burn": 0})
        run("state", "parse_hky", u2, 1, {"burn": 0})
        run("state", "parse_sw_joint", u3, 1, {"burn": 0})
    # L-mode report (section 8 f2/f4): the reference's own main() runs M mode on Sim3.u, then L mode (-r0 -v, -p6) on the .ti it
    # wrote; the .ti file and the report sections that the front end reproduces are kept
    for rep_name, rep_u in (("lmode_report_sim3", s3), ("lmode_report_3pop", os.path.join(HERE, "inputs", "parse_is_3pop.u"))):
        if ONLY and rep_name not in ONLY:
            continue
        base = os.path.join(TMP, rep_name + "_m.out")
        common = ["-i", rep_u, "-q10", "-m1", "-t3"]
        subprocess.run([HARNESS, "stock", os.path.join(TMP, rep_name + "_m.json"), "--"] + common + ["-o", base, "-b2000", "-l300", "-d10", "-hn2", "-hfl", "-ha0.9"],
                       check=True, cwd=TMP, stdout=subprocess.DEVNULL, timeout=900)
        rep = os.path.join(TMP, rep_name + "_l.out")
        joint = ["-c2"] if rep_name == "lmode_report_sim3" else []          # the joint-posterior search (two populations: the FULL model)
        subprocess.run([HARNESS, "stock", os.path.join(TMP, rep_name + "_l.json"), "--"] + common + ["-o", rep, "-r0", "-v", base, "-p56"] + joint,
                       check=True, cwd=TMP, stdout=subprocess.DEVNULL, timeout=900)
        text = open(rep).read()

        def section(start, end):
            a = text.index(start)
            return text[a:text.index(end, a)]
        keep = {"greater_than": section("\nPARAMETER COMPARISONS", "\nMEANS, VARIANCES"),
                "moments": section("\nMEANS, VARIANCES", "\nMarginal Peak Locations"),
                "peaks": section("\nMarginal Peak Locations", "\nHISTOGRAMS\n"),
                "t_histograms": section("HISTOGRAM GROUP 1", " After\t"),
                "histograms": section("HISTOGRAM GROUP 2", " After\t"),
                "popmig_histograms": section("HISTOGRAM GROUP 3: POPULATION MIGRATION", " After\t")}
        if joint:
            keep["joint"] = section("Joint Peak Locations", "\nHISTOGRAMS\n")
        import json
        with gzip.GzipFile(os.path.join(HERE, rep_name + ".json.gz"), "wb", mtime=0) as g:
            g.write(json.dumps(keep).encode())
        with open(base + ".ti", "rb") as f, gzip.GzipFile(os.path.join(HERE, "inputs", rep_name + ".ti.gz"), "wb", mtime=0) as g:
            shutil.copyfileobj(f, g)
        print("wrote %s.json.gz and inputs/%s.ti.gz" % (rep_name, rep_name))
    # the joint-posterior search of a three-population analysis (-c2: all population sizes, then all migration rates,
    # jointfind.cpp:1118-1133): the reference's L mode on the committed .ti file of lmode_report_3pop; the table is added to
    # that fixture under "joint"
    if not ONLY or "lmode_report_3pop_joint" in ONLY:
        import json
        base = os.path.join(TMP, "j3pop")
        with gzip.open(os.path.join(HERE, "inputs", "lmode_report_3pop.ti.gz"), "rb") as f, open(base + ".ti", "wb") as g:
            shutil.copyfileobj(f, g)
        rep = os.path.join(TMP, "j3pop_l.out")
        subprocess.run([HARNESS, "stock", os.path.join(TMP, "j3pop_l.json"), "--", "-i", os.path.join(HERE, "inputs", "parse_is_3pop.u"), "-q10", "-m1", "-t3",
                        "-o", rep, "-r0", "-v", base, "-p56", "-c2"], check=True, cwd=TMP, stdout=subprocess.DEVNULL, timeout=900)
        text = open(rep).read()
        a = text.index("Joint Peak Locations")
        fx = os.path.join(HERE, "lmode_report_3pop.json.gz")
        keep = json.load(gzip.open(fx))
        keep["joint"] = text[a:text.index("\nHISTOGRAMS\n", a)]
        with gzip.GzipFile(fx, "wb", mtime=0) as g:
            g.write(json.dumps(keep).encode())
        print("added the joint table to lmode_report_3pop.json.gz")
    # nested models (-w file, jointfind.cpp:40-135, 380-543): the reference's L mode on the committed .ti files with the committed
    # nested-model files; its joint table is added to the report fixtures under "joint_nested"
    if not ONLY or "lmode_nested_joint" in ONLY:
        import json
        for rep_name, rep_u, nest in (("lmode_report_sim3", s3, "nested_models_2pop.txt"),
                                      ("lmode_report_3pop", os.path.join(HERE, "inputs", "parse_is_3pop.u"), "nested_models_3pop.txt")):
            base = os.path.join(TMP, "nest_" + rep_name)
            with gzip.open(os.path.join(HERE, "inputs", rep_name + ".ti.gz"), "rb") as f, open(base + ".ti", "wb") as g:
                shutil.copyfileobj(f, g)
            rep = base + "_l.out"
            nestf = os.path.join(TMP, nest)                      # a short path: the report echoes the file name
            shutil.copy(os.path.join(HERE, "inputs", nest), nestf)
            subprocess.run([HARNESS, "stock", base + "_l.json", "--", "-i", rep_u, "-q10", "-m1", "-t3", "-o", rep, "-r0", "-v", base, "-w", nestf],
                           check=True, cwd=TMP, stdout=subprocess.DEVNULL, timeout=900)
            text = open(rep).read()
            a = text.index("Joint Peak Locations")
            fx = os.path.join(HERE, rep_name + ".json.gz")
            keep = json.load(gzip.open(fx))
            keep["joint_nested"] = text[a:text.index("\nHISTOGRAMS\n", a)].replace(nestf, "NESTEDFILE")
            with gzip.GzipFile(fx, "wb", mtime=0) as g:
                g.write(json.dumps(keep).encode())
            print("added the nested-model joint table to %s.json.gz" % rep_name)
    # ASCII curves of an L-mode report (section 8 f4): the reference's L mode on the committed .ti file of lmode_report_sim3
    if not ONLY or "lmode_ascii_sim3" in ONLY:
        base = os.path.join(TMP, "ascii_ref")
        with gzip.open(os.path.join(HERE, "inputs", "lmode_report_sim3.ti.gz"), "rb") as f, open(base + ".ti", "wb") as g:
            shutil.copyfileobj(f, g)
        rep = os.path.join(TMP, "lmode_ascii_sim3.out")
        subprocess.run([HARNESS, "stock", os.path.join(TMP, "lmode_ascii_sim3.json"), "--", "-i", s3, "-q10", "-m1", "-t3", "-o", rep, "-r0", "-v", base],
                       check=True, cwd=TMP, stdout=subprocess.DEVNULL, timeout=900)
        text = open(rep).read()
        a = text.index("ASCII Curves - Approximate Posterior Densities")
        import json
        with gzip.GzipFile(os.path.join(HERE, "lmode_ascii_sim3.json.gz"), "wb", mtime=0) as g:
            g.write(json.dumps({"curves": text[a:text.index("ASCII Plots of Parameter Trends", a)]}).encode())
        print("wrote lmode_ascii_sim3.json.gz")
    # opening sections of an M-mode report (section 8 f4): the reference's own main() on a committed input; kept from the top of
    # the file to the means / variances table -- run information, highest likelihoods, update-rate tables of the cold chain,
    # swap table
    if not ONLY or "report_head_3pop" in ONLY:
        rep = os.path.join(TMP, "report_head_3pop.out")
        args = ["-i", os.path.join(HERE, "inputs", "parse_is_3pop.u"), "-q10", "-m1", "-t3", "-b5000", "-l6000", "-d10", "-hn4", "-hfg", "-ha0.96", "-hb0.9"]
        subprocess.run([HARNESS, "stock", os.path.join(TMP, "report_head_3pop.json"), "--"] + args + ["-o", rep], check=True, cwd=TMP,
                       stdout=subprocess.DEVNULL, timeout=900)
        text = open(rep).read()
        import json
        with gzip.GzipFile(os.path.join(HERE, "report_head_3pop.json.gz"), "wb", mtime=0) as g:
            g.write(json.dumps({"args": args[2:], "head": text[:text.index("\nMEANS, VARIANCES")]}).encode())
        print("wrote report_head_3pop.json.gz")
    # the same command from 8 seeds (ref_harness stock seed=K; the serial build itself always seeds with 0): the update-rate and
    # swap tables of every run, so that a single run of the engine can be held against the reference's run-to-run spread
    if not ONLY or "report_rates_3pop" in ONLY:
        import json
        from concurrent.futures import ThreadPoolExecutor
        sys.path.insert(0, os.path.dirname(HERE))
        from test_frontend import _rate_tables
        args = ["-i", os.path.join(HERE, "inputs", "parse_is_3pop.u"), "-q10", "-m1", "-t3", "-b5000", "-l6000", "-d10", "-hn4", "-hfg", "-ha0.96", "-hb0.9"]

        def one(seed):
            rep = os.path.join(TMP, "report_rates_3pop_%d.out" % seed)
            try:
                subprocess.run([HARNESS, "stock", rep + ".json", "seed=%d" % seed, "--"] + args + ["-o", rep], check=True, cwd=TMP,
                               stdout=subprocess.DEVNULL, timeout=600)
            except subprocess.TimeoutExpired:       # the reference occasionally spins forever for some seeds
                return None
            text = open(rep).read()
            tables, swaps = _rate_tables(text[:text.index("\nMEANS, VARIANCES")])
            return {"seed": seed, "tables": tables, "swaps": swaps}
        with ThreadPoolExecutor(8) as ex:
            runs = [r for r in ex.map(one, range(101, 111)) if r is not None][:8]
        with gzip.GzipFile(os.path.join(HERE, "report_rates_3pop.json.gz"), "wb", mtime=0) as g:
            g.write(json.dumps({"args": args[2:], "runs": runs}).encode())
        print("wrote report_rates_3pop.json.gz", len(runs), "runs")
    # the .mcf state file: written by the reference (inputs/*.mcf.gz) and by the engine (inputs/*_ours.mcf.gz), each
    # read back by the reference's readmcf and dumped
    for nm, uf, hn, burn in (("mcf_sim5_hn2", s5, 2, 60), ("mcf_sim3_sw_hn2", sw3, 2, 60), ("mcf_sim5_hky_hn2", hky5, 2, 30),
                             ("mcf_sim3_joint_hn2", j3, 2, 60)):
        if ONLY and nm not in ONLY:
            continue
        mcf = os.path.join(TMP, nm + ".mcf")
        run("mcf", nm, uf, hn, {"burn": burn, "mcf": mcf})
        with open(mcf, "rb") as f, gzip.GzipFile(os.path.join(HERE, "inputs", nm + ".mcf.gz"), "wb", mtime=0) as g:
            shutil.copyfileobj(f, g)
        # the same state written by the engine (host-emulation build), read by the reference
        sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
        from support import engine_from_fixture, load_golden
        from ima2p_b200 import capi
        emu = capi.bind(os.path.join(os.path.dirname(HERE), "hostemu", "libima2p_hostemu.so"))
        eng, _ = engine_from_fixture(load_golden(nm), lib=emu)
        eng.eval()
        ours = os.path.join(TMP, nm + "_ours.mcf")
        eng.write_mcf(ours)
        eng.close()
        ONLY.add(nm + "_ours") if ONLY else None
        run("mcf", nm + "_ours", uf, hn, {"burn": 0, "mcf": ours, "load": 1})
        with open(ours, "rb") as f, gzip.GzipFile(os.path.join(HERE, "inputs", nm + "_ours.mcf.gz"), "wb", mtime=0) as g:
            shutil.copyfileobj(f, g)
    # proposal known-answer fixtures (a2-a4): accepted updategenealogy() calls
    run("updates", "updates_sim5_hn2", s5, 2, {"burn": 50, "n": 250})
    run("updates", "updates_sim3_hn2", s3, 2, {"burn": 50, "n": 250})
    run("updates", "updates_sim5_3pop_hn2", p3, 2, {"burn": 50, "n": 250})
    # numerics tables (a10, a13) and L mode (a14, a15)
    run("kat", "kat_sim5_hn4", s5, 4, {"burn": 100})
    run("lmode", "lmode_sim5_hn2", s5, 2, {"burn": 200, "rows": 600, "every": 3})
    os.makedirs(os.path.join(HERE, "inputs"), exist_ok=True)
    run("lmode", "lmode_ti_sim3_hn2", s3, 2, {"burn": 100, "rows": 60, "every": 2, "ti": os.path.join(HERE, "inputs", "sample_sim3.ti")})
    run("lmode", "lmode_sim5_expo_hn2", s5, 2, {"burn": 200, "rows": 300, "every": 3}, extra=["-j7"])
    # section 8 (f3): calcx moments and the 2NM densities over the rows (uniform and exponential migration priors, 3 populations)
    run("lmode", "lmode_extra_sim5_hn2", s5, 2, {"burn": 200, "rows": 400, "every": 3, "extra": 1})
    run("lmode", "lmode_extra_sim5_expo_hn2", s5, 2, {"burn": 200, "rows": 300, "every": 3, "extra": 1}, extra=["-j7"])
    run("lmode", "lmode_extra_sim5_3pop_hn2", p3, 2, {"burn": 150, "rows": 200, "every": 3, "extra": 1})
    # split-time, mutation-scalar updates (section 8 f1) and thermodynamic integration (a16)
    run("tupdates", "tupdates_sim5_hn2", s5, 2, {"burn": 100, "n": 30, "between": 3})
    run("tupdates", "tupdates_sim5_3pop_hn2", p3, 2, {"burn": 100, "n": 24, "between": 3})
    run("tupdates", "tupdates_sim3_sw_hn2", sw3, 2, {"burn": 100, "n": 12, "between": 3})
    run("nwupdates", "nwupdates_sim5_hn2", s5, 2, {"burn": 100, "n": 30, "between": 3})
    run("nwupdates", "nwupdates_sim3_hn2", s3, 2, {"burn": 100, "n": 30, "between": 3})
    run("nwupdates", "nwupdates_sim5_3pop_hn2", p3, 2, {"burn": 100, "n": 40, "between": 3})
    run("uupdates", "uupdates_sim5_hn2", s5, 2, {"burn": 100, "n": 40, "between": 2})
    run("uupdates", "uupdates_sim5_hky_hn2", hky5, 2, {"burn": 60, "n": 20, "between": 2})
    run("uupdates", "uupdates_sim3_sw_hn2", sw3, 2, {"burn": 100, "n": 16, "between": 2})
    run("tupdates", "tupdates_sim3_joint_hn2", j3, 2, {"burn": 100, "n": 12, "between": 3})
    run("uupdates", "uupdates_sim3_joint_hn2", j3, 2, {"burn": 100, "n": 24, "between": 2})
    run("thermo", "kat_thermo", s5, 2, {})
    # statistical parity (north_star: posterior summaries from long runs agree with the reference)
    run_trace("trace_sim5", s5, [1, 2, 3, 4, 5, 6], {"gburn": 3000, "sweeps": 60000, "nbatch": 12})
    run_trace("trace_sim3", s3, [1, 2, 3, 4], {"gburn": 3000, "sweeps": 60000, "nbatch": 12})
    # the same with a recent split time (most of every genealogy lies in the ancestral population)
    run_trace("trace_sim3_recent", s3, [1, 2, 3, 4], {"gburn": 3000, "sweeps": 60000, "nbatch": 12}, priors=["-q", "10", "-m", "1", "-t", "0.5"])
    # no migration: the other slider (slider_nomigration)
    run_trace("trace_sim3_nomig", s3, [1, 2, 3, 4], {"gburn": 3000, "sweeps": 60000, "nbatch": 12}, priors=nomig)
    run_trace("trace_sim5_3pop_nomig", p3, [1, 2, 3, 4], {"gburn": 3000, "sweeps": 60000, "nbatch": 12}, priors=nomig)
    # whole qupdate steps: genealogies + split time (RY1 or NW) + mutation scalars; the posterior of t and of the scalars
    run_trace("trace_full_sim5", s5, [11, 12, 13, 14, 15, 16], {"gburn": 5000, "sweeps": 60000, "nbatch": 12, "full": 1})
    # the same on one of our own input files: what the command-line front end is compared with
    if not ONLY or "trace_full_parse_is_3pop" in ONLY:
        u1 = os.path.join(HERE, "inputs", "parse_is_3pop.u")
        run_trace("trace_full_parse_is_3pop", u1, [11, 12, 13, 14, 15, 16], {"gburn": 5000, "sweeps": 60000, "nbatch": 12, "full": 1})
    run_trace("trace_full_sim3", s3, [11, 12, 13, 14], {"gburn": 5000, "sweeps": 60000, "nbatch": 12, "full": 1})


if __name__ == "__main__":
    main()
