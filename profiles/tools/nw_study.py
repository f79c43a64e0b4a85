#!/usr/bin/env python
"""Replicate study of the split-time samplers on a 3-population input (task 2).

Reference side: `oracle/_ref/ref_harness trace` (the reference's own qupdate / changet_NW / changet_RY1) with several seeds.
Engine side: the same kernels through the host emulation (tests only), several independent cold chains with their own seeds.
Schedules: full (RY1 or NW at random + changeu), nw (NW only + changeu), ry (RY1 only, no changeu).
Each replicate (one reference seed / one engine chain) contributes its run means; z = difference of the replicate means over
the combined between-replicate standard error.  Writes a markdown table to stdout and a JSON dump next to it.

    python profiles/tools/nw_study.py [--reps 10] [--sweeps 60000] [--burn 5000] [--schedules full,nw,ry]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
UFILE = os.path.join(ROOT, "tests", "golden", "inputs", "parse_is_3pop.u")
PRIORS = ["-q", "10", "-m", "1", "-t", "3"]
COMMON = ["-b", "100", "-l", "100", "-p01", "-z", "100000000"]
FULL = {"full": 1, "nw": 3, "ry": 4}
SCHED = {"full": (3, 5), "nw": (2, 5), "ry": (1, 0)}


def ref_run(args):
    sched, seed, burn, sweeps, tmp = args
    out = os.path.join(tmp, "ref_%s_%d.json" % (sched, seed))
    cmd = [HARNESS, "trace", out, "seed=%d" % seed, "gburn=%d" % burn, "sweeps=%d" % sweeps, "nbatch=12", "full=%d" % FULL[sched], "--",
           "-i", UFILE, "-o", os.path.join(tmp, "r_%s_%d.out" % (sched, seed))] + PRIORS + COMMON + ["-hn", "1"]
    try:
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=tmp, timeout=900)
    except (subprocess.TimeoutExpired, subprocess.CalledProcessError):
        return None
    d = json.load(open(out))
    bm = np.array(d["batch_means"])                       # [batch][locus][6]
    stats = {"t": np.array(d["t_batch_means"]).mean(axis=0).tolist(), "mignum": float(bm[:, :, 2].mean(axis=0).sum()),
             "length": float(bm[:, :, 0].mean(axis=0).sum()), "roottime": float(bm[:, :, 1].mean()),
             "t_rates": d["t_rates"]}
    return stats, {k: d[k] for k in ("tvals", "uvals", "tprior_max", "start", "model", "loci")}


def eng_run(args):
    sched, seed, burn, sweeps, setup = args
    from ima2p_b200 import Engine, capi
    from support import FlatModel, FlatTree
    lib = capi.bind(os.environ.get("IMA_STUDY_LIB", os.path.join(ROOT, "tests", "hostemu", "libima2p_hostemu.so")))
    fm = FlatModel(setup["model"])
    nloci = len(setup["loci"])
    eng = Engine(1, nloci, mig_capacity=200, seed=seed, lib=lib)
    eng.set_model_flat(*fm.create_args())
    for li, loc in enumerate(setup["loci"]):
        eng.set_locus(li, loc["model"], loc["numgenes"], loc["numsites"], loc["samppop"], seq=loc["seq"], hval=loc["hval"])
    eng.finalize()
    eng.set_betas([1.0])
    eng.set_chain(0, setup["tvals"])
    for li in range(nloci):
        t = FlatTree(setup["start"][li])
        eng.set_genealogy(0, li, t.up0, t.up1, t.down, t.pop, t.time, t.mig_off, t.mig_t[:-1], t.mig_p[:-1], t.root, t.roottime,
                          uvals=[setup["uvals"][li]])
    eng.upload()
    eng.eval()
    eng.set_update_priors(t_max=setup["tprior_max"])
    eng.set_update_schedule(*SCHED[sched])
    eng.run(burn, swaptries=0)
    eng.sync()
    c0 = eng.cold_counters(fm.nsplit, nloci)
    tacc = np.zeros(fm.nsplit)
    acc = np.zeros(3)
    every = 5                                               # sample every 5th step (the statistics are strongly autocorrelated)
    n = 0
    for _ in range(sweeps // every):
        eng.run(every, swaptries=0)
        tv, _, _ = eng.fetch_parameters()
        sd, si, _ = eng.fetch_pair_summaries()
        tacc += tv[0]
        acc += [si.reshape(nloci, 2)[:, 1].sum(), sd.reshape(nloci, 4)[:, 1].sum(), sd.reshape(nloci, 4)[:, 0].mean()]
        n += 1
    c1 = eng.cold_counters(fm.nsplit, nloci)
    dropped = eng.counters()["dropped"]
    eng.close()
    split = (np.array(c1["split"], dtype=np.int64) - np.array(c0["split"], dtype=np.int64)).reshape(fm.nsplit, 4)
    return {"t": (tacc / n).tolist(), "mignum": float(acc[0] / n), "length": float(acc[1] / n), "roottime": float(acc[2] / n),
            "t_rates": split.tolist(), "dropped": int(dropped)}


def summarise(reps):
    """replicate means -> dict name -> (mean, standard error, n)"""
    out = {}
    def put(name, vals):
        v = np.array(vals, dtype=float)
        out[name] = (float(v.mean()), float(v.std(ddof=1) / np.sqrt(len(v))) if len(v) > 1 else float("nan"), len(v))
    nsplit = len(reps[0]["t"])
    for k in range(nsplit):
        put("t%d" % k, [r["t"][k] for r in reps])
    put("migration events (all loci)", [r["mignum"] for r in reps])
    put("tree length (all loci)", [r["length"] for r in reps])
    put("root time (mean over loci)", [r["roottime"] for r in reps])
    for k in range(nsplit):
        for name, a, b in (("RY1", 0, 1), ("NW", 2, 3)):
            rates = [100.0 * r["t_rates"][k][b] / r["t_rates"][k][a] for r in reps if r["t_rates"][k][a] > 0]
            if rates:
                put("accept %% t%d %s" % (k, name), rates)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--sweeps", type=int, default=60000)
    ap.add_argument("--burn", type=int, default=5000)
    ap.add_argument("--schedules", default="full,nw,ry")
    ap.add_argument("--workers", type=int, default=os.cpu_count() or 4)
    ap.add_argument("--json", default="")
    args = ap.parse_args()
    tmp = tempfile.mkdtemp(prefix="nwstudy_")
    dump = {}
    print("# Split-time samplers on `tests/golden/inputs/parse_is_3pop.u` (3 populations, 4 loci, one cold chain)\n")
    print("%d replicates per side, %d steps after %d of burn-in; reference = `ref_harness trace` (its own `qupdate` / `changet_NW` / "
          "`changet_RY1`), engine = the same kernels through the host emulation.  z = difference of replicate means / combined "
          "between-replicate standard error.\n" % (args.reps, args.sweeps, args.burn))
    with ProcessPoolExecutor(args.workers) as ex:
        for sched in args.schedules.split(","):
            refs = list(ex.map(ref_run, [(sched, 100 + s, args.burn, args.sweeps, tmp) for s in range(args.reps + 3)]))
            refs = [r for r in refs if r is not None][:args.reps]
            setup = refs[0][1]
            engs = list(ex.map(eng_run, [(sched, 9000 + 13 * s, args.burn, args.sweeps, setup) for s in range(args.reps)]))
            R, E = summarise([r[0] for r in refs]), summarise(engs)
            dump[sched] = {"reference": [r[0] for r in refs], "engine": engs}
            print("## schedule `%s` (%s)\n" % (sched, {"full": "RY1 or NW at random, changeu every 5th step", "nw": "Nielsen-Wakeley only, changeu every 5th step",
                                                       "ry": "Rannala-Yang only, no changeu"}[sched]))
            print("| statistic | reference (mean ± se, n) | engine (mean ± se, n) | z |")
            print("|---|---|---|---|")
            for name in R:
                if name not in E:
                    continue
                (mr, sr, nr), (me, se, ne) = R[name], E[name]
                z = (me - mr) / np.sqrt(sr * sr + se * se + 1e-300)
                print("| %s | %.4f ± %.4f (%d) | %.4f ± %.4f (%d) | %+.2f |" % (name, mr, sr, nr, me, se, ne, z))
            print("\nengine proposals dropped for capacity: %d\n" % sum(e["dropped"] for e in engs), flush=True)
    if args.json:
        json.dump(dump, open(args.json, "w"))


if __name__ == "__main__":
    main()
