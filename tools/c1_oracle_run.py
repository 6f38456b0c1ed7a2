"""BASELINE config 1 on the CPU oracle: the bundled cloud, defaults (-v 0.008 -s 0.08, L*a*b*, adaptive lambda), automatic
threshold (41 thresholds 0.8 .. 1 against the ground truth; the file has no label field -> one ground-truth segment).
Runs only where /root/reference exists (this container); prints the figures DESIGN.md quotes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "fast-3d-pointcloud-segmentation_b200")]
import numpy as np
import oracle_py, oracle_testing as ot
from f3ps import pcd
path = "/root/reference/pcd/milk_cartoon_all_small_clorox.pcd"
pts = pcd.read_pcd(path)
if isinstance(pts, tuple):
    pts = pts[0]
print("points", len(pts), "finite", int(np.isfinite(pts["x"]).sum()), "all finite z < 0:", bool((pts["z"][np.isfinite(pts["z"])] < 0).all()))
def make():
    o = oracle_py.Oracle(); o.set_vccs_params(fold_negative_z=True); o.set_merge_params(color_mode=0, geom_mode=0, merge_mode=1, merge_impl=1); o.set_input(pts)
    return o
o = make(); t0 = time.time(); o.run(0, 0.2); dt = time.time() - t0
s = o.scalars()
print("V", o.array("keys").shape[0], "S", o.array("sv_label").shape[0], "E", o.array("edges_ab").shape[0], "lambda", s.get("lambda"),
      "M(t=0.2)", o.array("merges_ab").shape[0], "segments", len(np.unique(o.array("out_label"))), "oracle s", round(dt, 2), "stage_ms", o.array("stage_ms"))
V = o.array("voxel_xyz").shape[0]
t0 = time.time()
res = ot.all_thresh(make, o.array("voxel_xyz"), np.zeros(V, np.uint32))
bt, bp = ot.best_thresh(res)
print("auto threshold:", bt, bp, "sweep s", round(time.time() - t0, 1))
