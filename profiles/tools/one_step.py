#!/usr/bin/env python
"""A few steps launched kernel by kernel after a short burn-in: the target of `ncu -k regex:... -s N -c M`.
usage: python profiles/tools/one_step.py WORKLOAD BURN NSTEPS [fast ppw]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    wl, burn, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    eng, loci, st = bench.build_engine(wl, 0, 1)
    eng.set_update_priors(t_max=[bench.PRIOR_T])
    eng.set_update_schedule(3, 5)
    if len(sys.argv) > 5:
        eng.set_proposal_path(int(sys.argv[4]), int(sys.argv[5]))
    if os.environ.get("IMA_SPEC"):
        eng.set_speculation(int(os.environ["IMA_SPEC"]))
    ws = torch.cuda.Stream()
    torch.cuda.set_stream(ws)
    stream = ws.cuda_stream
    sw = eng.default_swaptries()
    pinned = {k: torch.from_numpy(np.ascontiguousarray(st[k])).pin_memory() for k in bench.STATE_KEYS}
    eng.put_state([pinned[k].data_ptr() for k in bench.STATE_KEYS], st["tvals"], stream)
    torch.cuda.synchronize()
    eng.run(burn, sw, stream)
    torch.cuda.synchronize()
    km = eng.run_timed(n, sw, stream)
    torch.cuda.synchronize()
    print("kernel ms over %d steps:" % n, [round(float(x), 4) for x in km])


if __name__ == "__main__":
    main()
