#!/usr/bin/env python
"""Two C2 frames through f3ps_run (first = warm-up), for ncu captures.  Usage: python tools/profile_frame.py [n_frames]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-3d-pointcloud-segmentation_b200"))
import f3ps
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
pts = f3ps.synth.make_frame(seed=20020)
g = f3ps.Segmenter(); g.set_vccs_params(); g.set_merge_params(color_mode=0, geom_mode=1, merge_mode=1); g.set_input(pts)
for i in range(n):
    g.set_input(pts); g.run(0.2)
print(g.stage_ms(), "launches", g.launch_count())
