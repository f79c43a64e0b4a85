"""Random streams: every Philox stream is keyed by (seed, stream id, step, purpose tag) (ima_kernels.h rng_for).  Two uses that are meant
to be independent must never produce the same (stream id, purpose) in one step -- round 1's Nielsen-Wakeley update did (its second
pass used edge NL + e of pair p: the id of edge e of pair p + 1 in the first pass, same tag).  This is a static check of the kernel
sources: every call site that opens a stream is listed here with its purpose tag and its id expression; a new call site, or an old one
with another id, fails the test until it is entered -- and the entries are checked for overlap on a concrete engine shape."""
import itertools
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "ima2p_b200", "csrc")

# purpose tag -> {id expression: what draws from it}.  Expressions with the same tag must have disjoint values; the two marked
# EXCLUSIVE alternatives of kRngScalars belong to different kernels of which an engine runs exactly one (launch_changeu).
EXPECTED = {
    "kRngPropose": {"(E.d.chain0 + c) * E.d.nloci + li": "the move of a pair (k_move and the general kernel: the same move by design)"},
    "kRngAlleles": {"(E.d.chain0 + c) * E.d.nloci + li": "stepwise allele states of a pair"},
    "kRngAccept": {"(E.d.chain0 + c) * E.d.nloci + l": "accept draw of a pair"},
    "kRngSplitTime": {"E.d.chain0 + c": "period, method and new split time of a chain",
                      "E.d.nchains_global + E.d.chain0 + c": "accept draw of the chain's split-time update"},
    "kRngSplitMig": {"pair_global * nl_max + edge": "Nielsen-Wakeley, first pass: lower-end population of an edge"},
    "kRngSplitMigSim": {"pair_global * nl_max + edge": "Nielsen-Wakeley, second pass: the re-simulated path of an edge"},
    "kRngScalars": {"E.d.chain0 + c": "EXCLUSIVE the one-warp scalar walk of a chain",
                    "(E.d.chain0 + c) * nloci + li": "EXCLUSIVE scalar proposal li of a chain (all loci infinite sites)",
                    "(E.d.chain0 + c) * nur + j": "EXCLUSIVE scalar proposal j of a chain (levelled walk)"},
    "kRngSwap": {"0xffffffffu": "the swap attempts of a step"},
}


def _call_sites():
    sites = []
    for fn in sorted(os.listdir(CSRC)):
        if not fn.endswith(".h"):
            continue
        for n, line in enumerate(open(os.path.join(CSRC, fn)), 1):
            code = line.split("//")[0]
            m = re.search(r"rng_for\(\w+, E, \(uint32_t\)\((.*)\), (\w+)\);", code)
            if m and "IMA_DEV" not in code:
                sites.append((fn, n, m.group(2), m.group(1).strip()))
                continue
            m = re.search(r"\.init\(E\.seed, (?:\(uint32_t\)\()?(.*?)\)?, \(uint32_t\)step, (kRng\w+)", code)
            if m:
                sites.append((fn, n, m.group(2), m.group(1).strip()))
                continue
            m = re.search(r"nw_edge_rng\(rng, E, pair_global, ([^,]+), ([^,]+), (kRng\w+)\)", code)
            if m:
                # the helper's own rng_for is the listed expression; here: the pitch must be NL and the edge a plain edge index
                assert m.group(1).strip() == "E.d.NL" and re.fullmatch(r"\w+", m.group(2).strip()), (fn, n, line)
                sites.append((fn, n, m.group(3), "pair_global * nl_max + edge"))
    return sites


def test_every_stream_call_site_is_a_known_one():
    sites = _call_sites()
    assert len(sites) >= 13, sites
    for fn, n, tag, expr in sites:
        if tag == "purpose":                       # the helper nw_edge_rng itself: its callers are listed
            continue
        assert tag in EXPECTED and expr in EXPECTED[tag], "%s:%d opens a stream (%s, %s) that tests/test_rng_streams.py does not know" % (fn, n, tag, expr)
    seen = {(tag, expr) for _, _, tag, expr in sites}
    for tag, exprs in EXPECTED.items():
        for expr in exprs:
            assert (tag, expr) in seen, "no call site left for (%s, %s): remove the entry" % (tag, expr)


def test_streams_of_one_purpose_do_not_overlap():
    nchains_global, chain0, nchains, nloci, NL, nur = 24, 8, 8, 7, 59, 9

    def ids(expr):
        out = set()
        e = expr.replace("E.d.chain0", str(chain0)).replace("E.d.nloci", str(nloci)).replace("E.d.nchains_global", str(nchains_global))
        e = e.replace("0xffffffffu", str(0xffffffff))
        edges = range(NL) if "edge" in expr else [0]
        js = range(nur) if re.search(r"\bj\b", expr) else [0]
        for c, li, edge, j in itertools.product(range(nchains), range(nloci), edges, js):
            env = {"c": c, "li": li, "l": li, "j": j, "edge": edge, "nloci": nloci, "nur": nur, "nl_max": NL, "pair_global": (chain0 + c) * nloci + li}
            out.add(eval(e, {}, env))
        return out
    for tag, exprs in EXPECTED.items():
        sets = {expr: ids(expr) for expr in exprs}
        for (a, sa), (b, sb) in itertools.combinations(sets.items(), 2):
            if exprs[a].startswith("EXCLUSIVE") and exprs[b].startswith("EXCLUSIVE"):
                continue
            assert not (sa & sb), (tag, a, b)
    # an id is one (pair, edge): no two pairs share an edge's stream (the overlap ADVICE.md found)
    nw = {}
    for c, li, edge in itertools.product(range(nchains), range(nloci), range(NL)):
        key = ((chain0 + c) * nloci + li) * NL + edge
        assert key not in nw, (c, li, edge, nw[key])
        nw[key] = (c, li, edge)
