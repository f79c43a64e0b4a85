#!/usr/bin/env python
"""Step time of Engine.run for a list of settings (one GPU).  Prints one JSON line per setting.
usage: python profiles/tools/pipe_sweep.py WORKLOAD STEPS "g,d,f[,fast,ppw[,spec]] ..."   (set_pipeline groups, depth, decisions_first;
set_proposal_path fast, pairs_per_warp).  IMA_TIMED=1 adds the per-kernel times of run_timed for every setting."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    wl, steps = sys.argv[1], int(sys.argv[2])
    settings = [tuple(int(x) for x in s.split(",")) for s in sys.argv[3].split()]
    burn = int(os.environ.get("IMA_BURN", 1500))
    eng, loci, st = bench.build_engine(wl, 0, 1)
    eng.set_update_priors(t_max=[bench.PRIOR_T])
    eng.set_update_schedule(3, 5)
    ws = torch.cuda.Stream()
    torch.cuda.set_stream(ws)
    stream = ws.cuda_stream
    sw = eng.default_swaptries()
    pinned = {k: torch.from_numpy(np.ascontiguousarray(st[k])).pin_memory() for k in bench.STATE_KEYS}
    eng.put_state([pinned[k].data_ptr() for k in bench.STATE_KEYS], st["tvals"], stream)
    torch.cuda.synchronize()
    eng.run(burn, sw, stream)
    torch.cuda.synchronize()
    nloci, _, _, cpg = bench.WORKLOADS[wl][:4]
    names = ["proposal", "k_accept", "k_swap", "k_split_t", "k_accept_t", "k_changeu", "k_move", "k_weigh", "k_propose_redo"]
    for s in settings:
        eng.set_pipeline(*s[:3])
        if len(s) >= 5:
            eng.set_proposal_path(s[3], s[4])
        if len(s) >= 6:
            eng.set_speculation(s[5])
        eng.run(2 * max(1, s[1]) + 8, sw, stream)
        torch.cuda.synchronize()
        best = None
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.run(steps, sw, stream)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            best = ms if best is None else min(best, ms)
        rec = {"workload": wl, "setting": s, "ms_per_step": best, "updates_per_s": cpg * nloci / (best * 1e-3)}
        if os.environ.get("IMA_TIMED"):
            km = eng.run_timed(steps, sw, stream)
            torch.cuda.synchronize()
            rec["kernel_us"] = {n: round(float(v) / steps * 1e3, 2) for n, v in zip(names, km)}
        print(json.dumps(rec), flush=True)
    c = eng.counters()
    print(json.dumps({"counters": {k: int(v) for k, v in c.items()}}), flush=True)


if __name__ == "__main__":
    main()
