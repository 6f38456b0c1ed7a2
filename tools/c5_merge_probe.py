"""K7 on the large scenes -- c5: 10 M-point room scan (S ~ 14 k, E ~ 50 k, M ~ 14 k); c4: the 10 M-point dense scene (S ~ 21 k,
E ~ 134 k, M ~ 19 k, 3,430 merges with more than 928 adjacency entries): the resident kernel with its tables in L2
(f3ps_set_merge_kernel 0 / 3) against the general kernel (2): ms, us per merge, identical merge logs; then the phase counters of
the same kernel compiled with them (5)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fast-3d-pointcloud-segmentation_b200"))
import numpy as np
import f3ps
from f3ps import synth
which = sys.argv[1] if len(sys.argv) > 1 else "c5"
if which == "c5":
    pts = synth.make_room_scan(n_points=int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000)
    vccs = dict(voxel_res=0.01, seed_res=0.1); mp = dict(color_mode=0, geom_mode=1, merge_mode=1)
else:
    pts = synth.make_dense_scene(seed=40000)
    vccs = dict(voxel_res=0.004, seed_res=0.04); mp = dict(color_mode=1, geom_mode=0, merge_mode=0, lam=0.5)
g = f3ps.Segmenter(); g.set_vccs_params(**vccs); g.set_merge_params(**mp); g.set_input(pts)
g.extract(); g.graph()
ref = None
for kern in (2, 0):
    g.set_merge_kernel(kern)
    for rep in range(2):
        g.merge(0.2); g.sync()
    c = g.counts(); ms = g.stage_ms()
    arr = {n: g.array(n).copy() for n in ("merges_ab", "merges_w", "merges_left", "final_ab", "final_w", "out_label")}
    if ref is None: ref = arr
    bad = [n for n in arr if not np.array_equal(arr[n], ref[n], equal_nan=arr[n].dtype.kind == "f")]
    print("kernel %d path %d: S %d E %d M %d max_T %d fold_steps %d  merge_kernel %.2f ms = %.2f us/merge  vs general: %s" % (
        kern, c.merge_path, c.n_supervoxels, c.n_edges, c.n_merges, c.max_touched, c.fold_steps, ms["merge_kernel"],
        1e3 * ms["merge_kernel"] / max(1, c.n_merges), "identical" if not bad else "DIFF " + ",".join(bad)), flush=True)
print({k: round(v, 3) for k, v in g.stage_ms().items()})
g.set_merge_kernel(5); g.merge(0.2); g.sync()            # the same kernel compiled with phase counters
import json
print("phase counters (kernel 5, %.2f ms):" % g.stage_ms()["merge_kernel"], json.dumps(g.merge_profile()))
