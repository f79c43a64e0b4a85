#!/usr/bin/env python
"""A few whole steps of small engines (infinite sites; HKY; stepwise) for compute-sanitizer:
   compute-sanitizer --tool memcheck|racecheck|synccheck python profiles/tools/sanitize.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from support import engine_from_fixture, load_golden  # noqa: E402


def main():
    nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    for name in ("state_sim5_hn4", "state_sim5_hky_hn2", "state_sim3_sw_hn2"):
        d = load_golden(name)
        eng, fm = engine_from_fixture(d, seed=11)
        eng.set_update_priors(t_max=[3.0] * fm.nsplit)
        eng.set_update_schedule(3, 2)
        eng.eval()
        eng.run(nsteps)
        eng.sync()
        c = eng.counters()
        print(name, "steps", c["steps"], "accepted", c["accepted"], "dropped", c["dropped"], flush=True)
        eng.close()


if __name__ == "__main__":
    main()
