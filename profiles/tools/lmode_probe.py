"""L-mode probe for ncu / timing (round 2, session 8): G bootstrapped .ti rows of the two-population model of bench.py's
L-mode section, then jointp for NV vectors in one device pass and one lock-step round of the peak searches
(ima2p_lmode_marginal_many: 3 row sets x 5 parameters x 2 brackets = 30 points).
usage: python profiles/tools/lmode_probe.py [G] [NV] [all-model-types]"""
import os
import sys
import time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402,F401  (device runtime)
from ima2p_b200 import LMode  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
NV = int(sys.argv[2]) if len(sys.argv) > 2 else 512
rng = np.random.default_rng(5)
nq, nm, nsplit = 3, 2, 1
rowlen = 3 * nq + 2 * nm + nq + nm + 2 + nsplit      # cc fc hcc | mc fm | qint mint | pdg probg | t
base = np.zeros((400, rowlen), np.float32)
base[:, 0:nq] = rng.integers(5, 40, (400, nq))                       # coalescent counts
base[:, nq:2 * nq] = rng.uniform(1, 20, (400, nq))                   # fc
base[:, 2 * nq:3 * nq] = rng.uniform(0, 30, (400, nq))               # hcc
base[:, 3 * nq:3 * nq + nm] = rng.integers(0, 6, (400, nm))          # migration counts
base[:, 3 * nq + nm:3 * nq + 2 * nm] = rng.uniform(0.5, 10, (400, nm))
base[:, 3 * nq + 2 * nm:3 * nq + 2 * nm + nq + nm] = rng.uniform(-5, 5, (400, nq + nm))
base[:, -3] = rng.uniform(-900, -850, 400)
base[:, -2] = base[:, 3 * nq + 2 * nm:3 * nq + 2 * nm + nq + nm].sum(axis=1)
base[:, -1] = rng.uniform(0.1, 2.9, 400)
rows = base[rng.integers(0, 400, G)]
lm = LMode(nq, nm, nsplit, [10.0] * nq, [0.0] * nq, [1.0] * nm, [0.0] * nm)
lm.load(rows)
xs = np.column_stack([rng.uniform(0.05, 0.9, NV) * (10.0 if p < nq else 1.0) for p in range(nq + nm)])
if len(sys.argv) > 3:                                  # every model type once (compute-sanitizer runs)
    for mt in (1, 2, 0):
        lm.set_joint_model(mt)
        lm.jointp(xs)
lm.jointp(xs[:64])
t0 = time.perf_counter()
for _ in range(3):
    q, _ = lm.jointp(xs)
t1 = time.perf_counter()
n = 30
kind = np.zeros(n, np.int32)
par = np.tile(np.repeat(np.arange(5), 2), 3)
first = np.repeat([0, G // 2 + 1, 0], 10)
last = np.repeat([G // 2, G, G], 10)
x = rng.uniform(0.05, 0.9, n) * np.where(par < nq, 10.0, 1.0)
lm.marginal_many(kind, par, first, last, x)
t2 = time.perf_counter()
for _ in range(10):
    m = lm.marginal_many(kind, par, first, last, x)
t3 = time.perf_counter()
one = np.array([lm.marginp(int(par[k]), int(first[k]), int(last[k]), x[k:k + 1])[0] for k in range(n)])
t4 = time.perf_counter()
assert np.array_equal(one, m)
print({"rows": G, "vectors": NV, "jointp_ms_per_pass": (t1 - t0) / 3 * 1e3, "jointp_geneval_per_sec": NV * G * 3 / (t1 - t0),
       "lockstep_round_ms": (t3 - t2) / 10 * 1e3, "same_30_points_one_call_each_ms": (t4 - t3) * 1e3, "checksum": float(q.sum())})
lm.close()
