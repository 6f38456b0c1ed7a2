"""Dependency structure of a merge sequence (CPU oracle; design probe for K7).

For the C2 frame: S, E, M, |b| and touched-edge distributions, the critical path of the merge DAG
(merge j depends on the last earlier merge that touched any region of its closed neighbourhood),
and what a look-ahead window over the sorted head edges would commit per step.
"""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fast-3d-pointcloud-segmentation_b200"))
from oracle.oracle_py import Oracle
from f3ps import synth

def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 20020
    flags = sys.argv[2] if len(sys.argv) > 2 else "c2"
    pts = synth.make_frame(seed)
    o = Oracle()
    o.set_vccs_params()
    if flags == "c2": o.set_merge_params(color_mode=0, geom_mode=1, merge_mode=1, merge_impl=1)
    else: o.set_merge_params(color_mode=0, geom_mode=0, merge_mode=2, bins=200, merge_impl=1)
    o.set_input(pts)
    t = time.time(); o.run(0, 0.2); print("oracle s", time.time() - t)
    ab = o.array("edges_ab"); w = o.array("edges_w"); mab = o.array("merges_ab"); mw = o.array("merges_w")
    svl = o.array("sv_label"); svc = o.array("sv_count")
    S, E, M = len(svl), len(ab), len(mab)
    print("S", S, "E", E, "M", M)
    size = dict(zip(svl.tolist(), svc.tolist()))
    adj = {int(l): set() for l in svl}
    for a, b in ab.tolist(): adj[a].add(b); adj[b].add(a)
    last = {int(l): 0 for l in svl}          # depth of the last merge that changed region / its incident edges
    depth = np.zeros(M, np.int64); cost = np.zeros(M); T = np.zeros(M, np.int64); nb = np.zeros(M, np.int64)
    lastc = {int(l): 0.0 for l in svl}
    for i, (a, b) in enumerate(mab.tolist()):
        nbh = (adj[a] | adj[b]) - {a, b}
        T[i] = len(adj[a]) + len(adj[b]) - 2
        nb[i] = size[b]
        # this merge reads a, b and every neighbour's stats; writes a and edges to every neighbour
        d = max([last[a], last[b]] + [last[x] for x in nbh])
        c = max([lastc[a], lastc[b]] + [lastc[x] for x in nbh])
        depth[i] = d + 1
        cost[i] = c + 2500 + 12 * size[b]      # cycles: fixed chain + fold
        # writers: a changes (all neighbours read it later); neighbours' edges change -> neighbours' incident sets
        last[a] = depth[i]; lastc[a] = cost[i]
        for x in nbh: last[x] = max(last[x], depth[i]); lastc[x] = max(lastc[x], cost[i])   # conservative: edge (a,x) re-weighted
        for x in adj[b]:
            if x != a: adj[x].discard(b); adj[x].add(a)
        adj[a] = nbh; del adj[b]
        size[a] += size[b]
    print("sum|b|", nb.sum(), "mean|b|", nb.mean(), "max|b|", nb.max())
    print("T mean", T.mean(), "max", T.max(), "pct", np.percentile(T, [50, 90, 99]))
    print("DAG critical path (merges, conservative)", depth.max(), " cost-weighted cycles", cost.max(), "=> ms", cost.max() / 1.965e6)
    print("serial cost model ms", (2500 * M + 12 * nb.sum()) / 1.965e6)
    nm = int((mw[1:] < mw[:-1]).sum()); print("non-monotone head steps", nm)
    np.savez("/tmp/w/c2_merge_%d.npz" % seed, ab=ab, w=w, mab=mab, mw=mw, T=T, nb=nb, depth=depth)

main()
