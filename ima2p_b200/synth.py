"""Synthetic "Simulations/*.u-shaped" inputs (BASELINE.json north_star: throughput is reported on
Simulations/*.u-shaped synthetic loci) and initial states in the engine's packed host layout.

The reference's shipped inputs are all infinite-sites, two populations, tree (0,1):2:
Sim1_5loci (10+10 genes), Sim1_50loci / Sim1_300loci (15+15 genes, 10-73 segregating sites per locus),
Sim2 (50+50), Sim3 (15+15).  ``make_dataset`` draws loci of that shape: a coalescent tree per locus with
mutations dropped on its branches (so the data satisfy the infinite-sites model by construction).
``write_u`` writes them in the reference's .u format so that the reference binary can be timed on
exactly the same data; ``initial_state`` builds a valid starting genealogy per (chain, locus) -- the
generating topology with every coalescence placed below the split time, no migration -- like the
reference's makeIS start (build_gtree.cpp:328-491) it is only a starting point for burn-in.
"""
import numpy as np

TIMEMAX = 1000000.0


def two_population_model(qmax=10.0, mmax=1.0):
    """Model tables of the 2-population IM model, tree (0,1):2 -- what setup_poptree / setup_iparams build
    (build_poptree.cpp:628-709, initialize.cpp:201-727) for -q qmax -m mmax."""
    return dict(npops=2, nsplit=1, plist=[[0, 1], [2]], addpop=[-1, 2], droppops=[[-1, -1], [0, 1]], pt_e=[1, 1, -1],
                pt_down=[2, 2, -1], rootpop=2, q_wp=[[(0, 0)], [(0, 1)], [(1, 0)]], q_max=[qmax] * 3, q_min=[0.0] * 3,
                m_wp=[[(0, 0, 1)], [(0, 1, 0)]], m_max=[mmax] * 2, m_min=[0.0] * 2)


def _coalescent_tree(n, rng):
    """Random-joining coalescent tree in the reference's edge encoding: (up0, up1, down, node_height)."""
    nl = 2 * n - 1
    up0, up1, down = -np.ones(nl, int), -np.ones(nl, int), -np.ones(nl, int)
    height = np.zeros(nl)
    active = list(range(n))
    t = 0.0
    for k in range(n, nl):
        m = len(active)
        t += rng.exponential(2.0 / (m * (m - 1)))
        i, j = rng.choice(m, 2, replace=False)
        a, b = active[i], active[j]
        up0[k], up1[k] = a, b
        down[a] = down[b] = k
        height[k] = t
        active = [e for e in active if e not in (a, b)] + [k]
    return up0, up1, down, height


def _im_tree(n0, n1, rng, split, mig):
    """Genealogy of n0 + n1 genes under the isolation-with-migration model the shipped inputs were simulated from
    (their headers: ms ... -I 2 n0 n1 ... split time 0.1): two populations of equal size exchanging migrants until
    they merge, backwards in time, at `split`.  Time in units of 2N generations (pairwise coalescence rate 1);
    `mig` = 4Nm, i.e. rate mig/4 per lineage in these units.  Tips 0..n0-1 are population 0."""
    n = n0 + n1
    nl = 2 * n - 1
    up0, up1, down = -np.ones(nl, int), -np.ones(nl, int), -np.ones(nl, int)
    height = np.zeros(nl)
    pops = [list(range(n0)), list(range(n0, n))]
    t, k = 0.0, n
    while k < nl:
        if t < split:
            rc = [len(p) * (len(p) - 1) / 2.0 for p in pops]
            rm = [len(p) * mig / 4.0 for p in pops]
            tot = sum(rc) + sum(rm)
            dt = rng.exponential(1.0 / tot) if tot > 0 else np.inf
            if t + dt >= split:
                t = split
                pops = [pops[0] + pops[1], []]
                continue
            t += dt
            x = rng.uniform(0, tot)
            if x < rc[0] + rc[1]:
                p = pops[0] if x < rc[0] else pops[1]
            else:
                src = 0 if x < rc[0] + rc[1] + rm[0] else 1
                e = pops[src].pop(int(rng.integers(len(pops[src]))))
                pops[1 - src].append(e)
                continue
        else:
            p = pops[0]
            m = len(p)
            t += rng.exponential(2.0 / (m * (m - 1)))
        i, j = rng.choice(len(p), 2, replace=False)
        a, b = p[i], p[j]
        up0[k], up1[k] = a, b
        down[a] = down[b] = k
        height[k] = t
        p[:] = [e for e in p if e not in (a, b)] + [k]
        k += 1
    return up0, up1, down, height


def make_dataset(nloci, n0, n1, seed=1, theta=5.0, min_sites=8, split=0.2, mig=1.0, structured=True):
    """nloci infinite-sites loci with n0 + n1 genes; returns a list of dicts (seq is [n][S] 0/1).

    structured: isolation-with-migration genealogies (ms -t 5 -I 2 n0 n1, split at 0.1 x 4N generations like
    Simulations/Sim1_*.u); otherwise one panmictic population with genes dealt to the two samples at random."""
    rng = np.random.default_rng(seed)
    n = n0 + n1
    loci = []
    while len(loci) < nloci:
        up0, up1, down, height = _im_tree(n0, n1, rng, split, mig) if structured else _coalescent_tree(n, rng)
        nl = 2 * n - 1
        tips = [None] * nl
        for e in range(nl):
            tips[e] = {e} if e < n else tips[up0[e]] | tips[up1[e]]
        cols = []
        for e in range(nl - 1):
            blen = height[down[e]] - height[e]
            for _ in range(rng.poisson(theta / 2.0 * blen)):
                col = np.zeros(n, np.int8)
                col[list(tips[e])] = 1
                cols.append(col)
        if len(cols) < min_sites:
            continue
        seq = np.stack(cols, axis=1)
        seq = seq[:, rng.permutation(seq.shape[1])]
        # the reference codes the base of the first sequence as 0 at every segregating site (readseqIS)
        flip = seq[0] == 1
        seq[:, flip] = 1 - seq[:, flip]
        # panmictic case: random assignment of genes to populations (tips 0..n0-1 are population 0)
        perm = np.arange(n) if structured else rng.permutation(n)
        seq = seq[perm]
        flip = seq[0] == 1
        seq[:, flip] = 1 - seq[:, flip]
        inv = np.argsort(perm)          # old tip -> new tip index
        remap = np.concatenate([inv, np.arange(n, nl)])
        u0, u1 = remap[up0[n:]], remap[up1[n:]]
        nu0, nu1, nd = -np.ones(nl, int), -np.ones(nl, int), -np.ones(nl, int)
        nu0[n:], nu1[n:] = u0, u1
        for k in range(n, nl):
            nd[nu0[k]] = nd[nu1[k]] = k
        loci.append(dict(name="Locus%d" % (len(loci) + 1), n=n, samppop=[n0, n1], numsites=seq.shape[1],
                         seq=seq.astype(np.int32), up0=nu0, up1=nu1, down=nd, height=height.copy()))
    return loci


def write_u(path, loci, title="synthetic Simulations-shaped data for ima2p_b200"):
    """Reference .u format (readata.cpp:1037-1056, 617-866): ancestral allele A, derived G at every column."""
    with open(path, "w") as f:
        f.write(title + "\n# generated by ima2p_b200.synth\n2\npop1\tpop2\n(0,1):2\n%d\n" % len(loci))
        for L in loci:
            f.write("%s %d %d %d I0 1\n" % (L["name"], L["samppop"][0], L["samppop"][1], L["numsites"]))
            for j in range(L["n"]):
                f.write("%-10s%s\n" % ("Indiv%d" % j, "".join("AG"[b] for b in L["seq"][j])))


def initial_state(loci, nchains, NL, CAP, t0=0.15, seed=7, u=1.0):
    """Packed host buffers (the layout of ima2p_engine_put_state) for nchains x len(loci) genealogies.

    Every chain starts from the generating topology with its node heights shifted below the split time t0
    (tips in their sampled populations, internal edges in the ancestral population, no migration)."""
    rng = np.random.default_rng(seed)
    nloci = len(loci)
    P = nchains * nloci
    topo = -np.ones((P, NL, 4), np.int16)
    time = np.zeros((P, NL))
    mseg = np.zeros((P, NL, 2), np.uint16)
    mig_t = np.zeros((P, CAP))
    mig_p = np.zeros((P, CAP), np.int16)
    si = np.zeros((P, 2), np.int32)
    sd = np.zeros((P, 4))
    uv = np.ones((P, 4)) * u
    tvals = np.full((nchains, 1), t0)
    for li, L in enumerate(loci):
        n, nl = L["n"], 2 * L["n"] - 1
        pop = np.full(nl, 2, np.int16)
        pop[:L["samppop"][0]] = 0
        pop[L["samppop"][0]:n] = 1
        root = nl - 1
        for c in range(nchains):
            p = c * nloci + li
            scale = 0.5 + rng.random()
            h = t0 * 1.05 + L["height"] * scale        # node heights, all below (older than) t0
            topo[p, :nl, 0], topo[p, :nl, 1], topo[p, :nl, 2], topo[p, :nl, 3] = L["up0"], L["up1"], L["down"], pop
            tm = np.where(L["down"] >= 0, h[np.maximum(L["down"], 0)], TIMEMAX)
            time[p, :nl] = tm
            si[p] = (root, 0)
            sd[p, 0] = h[root]
    return dict(topo=topo, time=time, mseg=mseg, mig_t=mig_t, mig_p=mig_p, scal_i=si, scal_d=sd, uvals=uv, tvals=tvals)
