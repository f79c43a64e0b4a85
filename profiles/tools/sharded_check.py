#!/usr/bin/env python
"""Two or more ranks (torchrun, one GPU each): chains of a golden fixture sharded by rank, stepped with
ima2p_engine_run_sharded (swap sums through peer memory inside the kernels); rank 0 also runs all chains on its own GPU with
ima2p_engine_run and the two runs must agree bit for bit.  Then times the sharded step.
usage: torchrun --nproc-per-node N profiles/tools/sharded_check.py [fixture] [nsteps]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from ima2p_b200 import Engine, capi
    from ima2p_b200.multirank import attach_exchange
    from support import FlatModel, FlatTree, load_golden
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    name = sys.argv[1] if len(sys.argv) > 1 else "state_sim5_hn4"
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    d = load_golden(name)
    fm = FlatModel(d["model"])
    nglob = len(d["chains"]) // world * world
    per = nglob // world
    betas = [d["chains"][c]["beta"] for c in range(nglob)]

    def build(nch, c0, device):
        eng = Engine(nch, len(d["loci"]), seed=99, lib=capi.lib(), nchains_global=nglob, chain0=c0, device=device)
        eng.set_model_flat(*fm.create_args())
        for li, loc in enumerate(d["loci"]):
            eng.set_locus(li, loc["model"], loc["numgenes"], loc["numsites"], loc["samppop"], seq=loc["seq"], hval=loc["hval"], sumlogk=loc["sumlogk"])
        eng.finalize()
        eng.set_betas(betas)
        for k in range(nch):
            ch = d["chains"][c0 + k]
            eng.set_chain(k, ch["tvals"])
            for li, g in enumerate(ch["G"]):
                t = FlatTree(g["tree"])
                eng.set_genealogy(k, li, t.up0, t.up1, t.down, t.pop, t.time, t.mig_off, t.mig_t[:-1], t.mig_p[:-1], t.root, t.roottime, uvals=g["uvals"])
        eng.upload()
        eng.eval()
        eng.set_update_priors(t_max=[3.0] * fm.nsplit)
        eng.set_update_schedule(3, 5)
        return eng

    eng = build(per, per * rank, local)
    attach_exchange(eng, local)
    sw = max(1, nglob // 10)
    eng.run_sharded(nsteps, sw)
    eng.sync()
    mine = np.array([[eng.chain(c)["probg"], eng.chain(c)["pdg"], eng.chain(c)["tvals"][0], eng.chain(c)["beta"]] for c in range(per)])
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    ok = True
    if rank == 0:
        full = build(nglob, 0, local)
        full.run(nsteps, sw)
        full.sync()
        want = np.array([[full.chain(c)["probg"], full.chain(c)["pdg"], full.chain(c)["tvals"][0], full.chain(c)["beta"]] for c in range(nglob)])
        got = np.concatenate(parts)
        ok = bool(np.array_equal(got, want))
        print("sharded run over %d ranks equals the one-GPU run bit for bit: %s (swaps %d)" % (world, ok, full.counters()["swaps"]), flush=True)
    # timing of the sharded step (device events, max over ranks)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s = torch.cuda.current_stream().cuda_stream
    eng.run_sharded(50, sw, s)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0.record(); eng.run_sharded(400, sw, s); e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 400], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("sharded step %.4f ms (max over ranks)" % float(t.item()), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
