"""Phase profile of the merge stage on a C5-style graph (10 M-point room scan: S ~ 14 k, E ~ 50 k, M ~ 14 k)."""
import sys, time
sys.path.insert(0, "fast-3d-pointcloud-segmentation_b200")
import numpy as np
import f3ps
from f3ps import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
pts = synth.make_room_scan(n_points=n)
g = f3ps.Segmenter()
g.set_vccs_params(voxel_res=0.01, seed_res=0.1)
g.set_merge_params(color_mode=0, geom_mode=1, merge_mode=1)
g.set_input(pts)
g.run(0.2)
for rep in range(2):
    g.merge(0.2)
    g.sync()
    c = g.counts()
    ms = g.stage_ms()
    prof = g.merge_profile()
    print("path", c.merge_path, "S", c.n_supervoxels, "E", c.n_edges, "M", c.n_merges, "max_T", c.max_touched, "fold_steps", c.fold_steps,
          "merge_ms", round(ms["merge"], 2), "kernel_ms", round(ms["merge_kernel"], 2), "us/merge", round(1e3 * ms["merge_kernel"] / max(1, c.n_merges), 2))
    if isinstance(prof, dict):
        tot = sum(v for v in prof.values() if isinstance(v, int)) or 1
        print({k: (round(v / 1.965e3 / max(1, c.n_merges), 2) if isinstance(v, int) else v) for k, v in prof.items()}, "(us per merge)", "avg T", round(prof.get("sum_T", 0) / max(1, c.n_merges), 1))
print(g.stage_ms())
