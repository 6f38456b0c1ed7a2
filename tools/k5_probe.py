"""K5 probe: the expansion stage of one VGA frame with the cooperative grid and with one cluster of 1..16 CTAs
(f3ps_set_expand_kernel): ms per launch, ns per phase, parity of labels / dist between the kernels."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-3d-pointcloud-segmentation_b200"))
import f3ps
from f3ps import synth

def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 20020
    pts = synth.make_frame(seed=seed)
    g = f3ps.Segmenter(); g.set_vccs_params(); g.set_merge_params(color_mode=0, geom_mode=1, merge_mode=1); g.set_input(pts)
    g.voxelize(); g.neighbors(); g.normals(); g.seeds()
    ref = None
    for which, ctas in ((1, 0), (2, 16), (2, 12), (2, 8), (2, 4), (2, 2), (2, 1), (0, 0)):
        g.set_expand_kernel(which, ctas)
        ms = []
        for r in range(5):
            g.expand(); g.sync(); ms.append(g.stage_ms()["expand"])
        lab, dist = g.array("labels").copy(), g.array("dist").copy()
        if ref is None: ref = (lab, dist)
        ok = np.array_equal(lab, ref[0]) and np.array_equal(dist, ref[1])
        c = g.counts()
        print("kernel %d ctas %2d: V=%d S=%d  expand ms min %.3f med %.3f  parity %s  phases_us %s" % (
            which, ctas, c.n_voxels, c.n_supervoxels, min(ms), sorted(ms)[len(ms) // 2], "OK" if ok else "DIFF",
            {k: round(v / 1e3, 1) for k, v in g.expand_profile().items()}))
    g.set_expand_kernel(0, 0)
    for r in range(3):
        g.run(0.2); g.sync(); print("full run", {k: round(v, 3) for k, v in g.stage_ms().items()})

main()
