#!/usr/bin/env python
"""K7 on one frame: merge-log parity against the oracle, kernel time, per-phase cycle counters.
Usage: python tools/k7_probe.py [small|vga|<seed>] [cvx_al|eq|rgb_ml] [reps]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "fast-3d-pointcloud-segmentation_b200"))
import oracle_py
import f3ps
from f3ps import synth

def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "vga"
    mode = sys.argv[2] if len(sys.argv) > 2 else "cvx_al"
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    pts = synth.make_frame(seed=11, width=160, height=120) if which == "small" else synth.make_frame(seed=20020 if which == "vga" else int(which))
    mp = dict(color_mode=0, geom_mode=1, merge_mode=1, lam=0.5, bins=500)
    if mode == "eq": mp = dict(color_mode=0, geom_mode=0, merge_mode=2, lam=0.5, bins=200)
    if mode == "rgb_ml": mp = dict(color_mode=1, geom_mode=0, merge_mode=0, lam=0.5, bins=500)
    thr = 0.2
    o = oracle_py.Oracle(); o.set_vccs_params(); o.set_merge_params(merge_impl=1, **mp); o.set_input(pts); o.run(0, thr)
    g = f3ps.Segmenter(); g.set_vccs_params(); g.set_merge_params(**mp); g.set_input(pts)
    g.extract(); g.graph(); g.sync()
    ok = True
    for kern in (0, 3, 2):
        g.set_merge_kernel(kern)
        ms = []
        for r in range(reps):
            g.merge(thr); g.sync(); ms.append(g.stage_ms().get("merge_kernel", float("nan")))
        c = g.counts()
        bad = [n for n in ("merges_ab", "merges_w", "merges_left", "final_ab", "final_w", "out_label", "out_voxel")
               if not np.array_equal(g.array(n), o.array(n))]
        ok = ok and not bad
        print("kernel %d path %d: M=%d S=%d E=%d max_T=%d  merge_kernel ms min %.3f med %.3f  us/merge %.3f  parity %s" % (
            kern, c.merge_path, c.n_merges, c.n_supervoxels, c.n_edges, c.max_touched, min(ms), sorted(ms)[len(ms) // 2],
            1e3 * min(ms) / max(1, c.n_merges), "OK" if not bad else "DIFF " + ",".join(bad)))
        if kern == 4:
            p = g.merge_profile(); M = max(1, c.n_merges)
            for role in ("worker", "mean", "cov"):
                print("   %-6s cycles/merge:" % role, {k: round(v / M) for k, v in p[role].items()})
            print("   by touched-edge class:", p["by_touched"])
            print("   guess_misses %d  ciede_evals %d  sum_T %d  fold_steps %d" % (p["guess_misses"], p["ciede_evals"], p["sum_T"], c.fold_steps))
    g.set_merge_kernel(0)
    for i in range(3):
        g.set_input(pts); t = time.time(); g.run(thr); g.sync(); print("full run wall %.2f ms" % ((time.time() - t) * 1e3), {k: round(v, 3) for k, v in g.stage_ms().items()})
    return 0 if ok else 1

if __name__ == "__main__":
    sys.exit(main())
