#!/usr/bin/env python
"""Stage-by-stage GPU-vs-oracle comparison on one frame (development probe; the asserting
version lives in tests/test_gpu_parity.py).  Usage: python tools/gpu_parity_probe.py [small|vga] [flags]"""
import os, sys, time, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "fast-3d-pointcloud-segmentation_b200"))
import oracle_py
import f3ps
from f3ps import synth

def cmp(name, a, b, exact=True):
    a = np.asarray(a); b = np.asarray(b)
    if a.shape != b.shape:
        print("  %-12s SHAPE MISMATCH gpu %s oracle %s" % (name, a.shape, b.shape)); return False
    if a.dtype.kind == 'f':
        same = (a == b) | (np.isnan(a) & np.isnan(b))
    else:
        same = a == b
    bad = int((~same).sum())
    extra = ""
    if bad and a.dtype.kind == 'f':
        with np.errstate(invalid='ignore', divide='ignore'):
            rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-30)
        extra = " max_rel %.3g" % np.nanmax(rel[~same])
    print("  %-12s %s  mismatches %d / %d%s" % (name, "OK " if bad == 0 else "DIFF", bad, a.size, extra))
    return bad == 0

def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    mode = sys.argv[2] if len(sys.argv) > 2 else "cvx_al"
    if which == "small": pts = synth.make_frame(seed=11, width=160, height=120)
    elif which == "vga": pts = synth.make_frame(seed=20020)
    else: pts = synth.make_frame(seed=int(which))
    mp = dict(color_mode=0, geom_mode=1, merge_mode=1, lam=0.5, bins=500)
    if mode == "eq": mp = dict(color_mode=0, geom_mode=0, merge_mode=2, lam=0.5, bins=200)
    if mode == "rgb_ml": mp = dict(color_mode=1, geom_mode=0, merge_mode=0, lam=0.5, bins=500)
    thr = 0.2
    o = oracle_py.Oracle(); o.set_vccs_params(); o.set_merge_params(merge_impl=0, **mp); o.set_input(pts)
    t = time.time(); o.run(0, thr); print("oracle %.1f ms" % ((time.time() - t) * 1e3), o.array("stage_ms"))
    g = f3ps.Segmenter(); g.set_vccs_params(); g.set_merge_params(**mp); g.set_input(pts)
    if os.environ.get("F3PS_MERGE_KERNEL"): g.set_merge_kernel(int(os.environ["F3PS_MERGE_KERNEL"]))
    stages = [("voxelize", ["keys", "voxel_xyz", "voxel_rgb", "voxel_rgba", "voxel_count", "point_voxel"]),
              ("neighbors", ["nbr_count", "nbr"]), ("normals", ["normals", "curvature"]), ("seeds", ["seeds"]),
              ("expand", ["labels", "dist", "sv_label", "sv_xyz", "sv_rgb", "sv_normal", "sv_count"]),
              ("graph", ["adj", "edges_ab", "edges_dc", "edges_dg", "edges_w"])]
    for st, names in stages:
        try:
            t = time.time(); getattr(g, st)(); g.sync(); dt = (time.time() - t) * 1e3
            print("stage %s (%.2f ms wall)" % (st, dt))
            for n in names: cmp(n, g.array(n), o.array(n))
        except Exception:
            traceback.print_exc(); return 1
    c = g.counts(); print("  lambda gpu %.9g oracle %.9g" % (c.lambda_, o.scalars()["lambda"]))
    if mode == "eq":
        cmp("cdf_c", g.array("cdf_c"), o.array("cdf_c")); cmp("cdf_g", g.array("cdf_g"), o.array("cdf_g"))
    try:
        t = time.time(); g.merge(thr); dt = (time.time() - t) * 1e3
        print("stage merge (%.2f ms wall)" % dt)
        for n in ["merges_ab", "merges_w", "merges_left", "final_ab", "final_w", "out_label", "out_voxel", "out_xyz"]:
            cmp(n, g.array(n), o.array(n))
    except Exception:
        traceback.print_exc(); return 1
    c = g.counts()
    print("counts:", {k: getattr(c, k) for k, _ in c._fields_})
    # timed full runs
    for i in range(3):
        g.set_input(pts); t = time.time(); g.run(thr); print("full run wall %.2f ms" % ((time.time() - t) * 1e3), g.stage_ms())
    print("launches", g.launch_count())
    print("merge profile (cycles)", g.merge_profile())
    print("expand profile (ns)", g.expand_profile())
    return 0

if __name__ == "__main__":
    sys.exit(main())
