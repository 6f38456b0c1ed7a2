"""f3ps_refine against the oracle's refine on a few frames (which arrays differ, where)."""
import sys, os, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fast-3d-pointcloud-segmentation_b200"))
from oracle.oracle_py import Oracle
import f3ps
from f3ps import synth
for seed, w, h, itr in ((11, 160, 120, 3), (52, 160, 120, 1), (52, 160, 120, 2), (52, 160, 120, 3), (20020, 640, 480, 3)):
    pts = synth.make_frame(seed=seed, width=w, height=h)
    o = Oracle(); o.set_vccs_params(); o.set_merge_params(color_mode=0, geom_mode=1, merge_mode=1, merge_impl=1); o.set_input(pts)
    for st in (1, 2, 3, 4, 5): o.run(st)
    t = time.time(); o.refine(itr); to = time.time() - t
    o.run(6); o.run(7, 0.2)
    g = f3ps.Segmenter(); g.set_vccs_params(); g.set_merge_params(color_mode=0, geom_mode=1, merge_mode=1); g.set_input(pts)
    g.extract(); g.sync(); t = time.time(); g.refine(itr); g.sync(); tg = time.time() - t
    g.graph(); g.merge(0.2)
    bad = [n for n in ("normals", "curvature", "labels", "dist", "sv_label", "sv_count", "sv_xyz", "sv_rgb", "sv_normal", "edges_ab", "edges_w", "merges_ab", "out_label", "seeds")
           if not np.array_equal(g.array(n), o.array(n), equal_nan=g.array(n).dtype.kind == 'f')]
    print(seed, "itr", itr, "S", len(o.array("sv_label")), "M", len(o.array("merges_ab")), "oracle %.2fs gpu %.4fs" % (to, tg), "DIFF:", bad)
    if "dist" in bad:
        dg, do = g.array("dist"), o.array("dist"); idx = np.nonzero(dg != do)[0]
        lab = o.array("labels"); cnt = dict(zip(o.array("sv_label").tolist(), o.array("sv_count").tolist()))
        for v in idx[:10]:
            print("   voxel", v, "gpu", dg[v], "oracle", do[v], "label", lab[v], "helper size", cnt.get(int(lab[v])), "nbr_count", o.array("nbr_count")[v])
