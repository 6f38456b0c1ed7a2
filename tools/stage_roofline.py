#!/usr/bin/env python
"""Per-stage time and achieved algorithmic GB/s on the large configurations (C4-style 10 M-point frame), where the
front-end kernels are bandwidth-relevant.  Byte formulas: DESIGN.md section 3 / SURVEY.md 8(d).  Writes a markdown table."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-3d-pointcloud-segmentation_b200"))
import f3ps
from f3ps import synth

peak = 6536.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
which = sys.argv[1] if len(sys.argv) > 1 else "c4"
if which == "c4":
    pts = synth.make_dense_scene(seed=40000); vr, sr = 0.004, 0.04; mp = dict(color_mode=1, geom_mode=0, merge_mode=0, lam=0.5); name = "C4: 10 M-point dense frame, -v 0.004 -s 0.04 --RGB --ML 0.5"
else:
    pts = synth.make_frame(seed=20020); vr, sr = 0.008, 0.08; mp = dict(color_mode=0, geom_mode=1, merge_mode=1); name = "C2: VGA frame"
g = f3ps.Segmenter(); g.set_vccs_params(voxel_res=vr, seed_res=sr); g.set_merge_params(**mp)
for rep in range(2):
    g.set_input(pts); g.extract()
c = g.counts(); ms = g.stage_ms(); xp = g.expand_profile()
N, V, S0, S, E = c.n_points, c.n_voxels, c.n_seeds, c.n_supervoxels, c.n_edges
sweeps, rounds = c.sweeps, c.rounds
bytes_ = {
    "voxelize": 16 * N + 32 * V,
    "neighbors": 8 * V + 4 * 28 * V,
    "normals": 16 * V + 112 * V + 20 * V,
    "seeds": 16 * V + 4 * S0,
    "expand": sweeps * (156 + 12) * V + rounds * (44 * V + 40 * S0),
    "graph": (4 + 112) * V + 16 * V + 28 * E,
}
print("# %s\n" % name)
print("N=%d V=%d S0=%d S=%d E=%d, %d expansion rounds / %d sweeps; peak %.0f GB/s (measured copy bandwidth)\n" % (N, V, S0, S, E, rounds, sweeps, peak))
print("| stage | ms | algorithmic MB | GB/s | fraction of peak |\n|---|---|---|---|---|")
for k in ("voxelize", "neighbors", "normals", "seeds", "expand", "graph"):
    gbs = bytes_[k] / (ms[k] * 1e-3) / 1e9
    print("| %s | %.3f | %.1f | %.1f | %.3f |" % (k, ms[k], bytes_[k] / 1e6, gbs, gbs / peak))
tot = sum(ms[k] for k in ("voxelize", "neighbors", "normals", "seeds", "expand", "graph"))
print("\nK1..K6 total %.3f ms = %.1f Mpoints/s (single frame, one stream, host round trips included)" % (tot, N / tot / 1e3))
print("expand phases (ns):", xp)
if len(sys.argv) > 2 and sys.argv[2] == "merge":
    t0 = time.time(); g.merge(0.2); dt = time.time() - t0
    c = g.counts()
    print("\nmerge: %d merges in %.1f ms (kernel path %d, max touched %d, fold steps %d): %.2f us per merge" % (
        c.n_merges, dt * 1e3, c.merge_path, c.max_touched, c.fold_steps, dt * 1e6 / max(1, c.n_merges)))
