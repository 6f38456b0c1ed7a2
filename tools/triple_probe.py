"""Frames where three seed cells elect one voxel (two phantom holders + the owner): GPU against the literal oracle."""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fast-3d-pointcloud-segmentation_b200"))
from oracle.oracle_py import Oracle
import f3ps
from f3ps import synth
for seed in (22030, 22049, 20020):
    pts = synth.make_frame(seed=seed)
    o = Oracle(); o.set_vccs_params(); o.set_merge_params(color_mode=0, geom_mode=1, merge_mode=1, merge_impl=1); o.set_input(pts); o.run(0, 0.2)
    g = f3ps.Segmenter(); g.set_vccs_params(); g.set_merge_params(color_mode=0, geom_mode=1, merge_mode=1); g.set_input(pts); g.run(0.2)
    s = o.array("seeds"); u, c = np.unique(s, return_counts=True)
    bad = [n for n in ("seeds", "labels", "dist", "sv_label", "sv_count", "sv_xyz", "sv_rgb", "sv_normal", "edges_ab", "edges_w", "merges_ab", "out_label", "out_voxel")
           if not np.array_equal(g.array(n), o.array(n), equal_nan=g.array(n).dtype.kind == 'f')]
    print(seed, "max election multiplicity", c.max(), "S", len(o.array("sv_label")), "M", len(o.array("merges_ab")), "DIFF:", bad)
    for itr in (2,):
        o2 = Oracle(); o2.set_vccs_params(); o2.set_merge_params(color_mode=0, geom_mode=1, merge_mode=1, merge_impl=1); o2.set_input(pts)
        for st in (1, 2, 3, 4, 5): o2.run(st)
        o2.refine(itr); o2.run(6); o2.run(7, 0.2)
        g.set_input(pts); g.extract(); g.refine(itr); g.graph(); g.merge(0.2)
        bad = [n for n in ("normals", "labels", "dist", "sv_label", "sv_count", "merges_ab") if not np.array_equal(g.array(n), o2.array(n), equal_nan=g.array(n).dtype.kind == 'f')]
        print("   refine", itr, "DIFF:", bad)
