import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fast-3d-pointcloud-segmentation_b200"))
import f3ps
from f3ps import synth
pts = synth.make_frame(seed=20020)
g = f3ps.Segmenter(); g.set_vccs_params(); g.set_merge_params(color_mode=0, geom_mode=1, merge_mode=1); g.set_input(pts)
g.voxelize(); g.neighbors(); g.normals(); g.seeds()
g.set_expand_kernel(2, int(sys.argv[1]) if len(sys.argv) > 1 else 1)
for r in range(2):
    g.expand(); g.sync()
