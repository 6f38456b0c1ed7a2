#!/usr/bin/env python
"""Which stage limits frames-in-flight scaling: per-stage concurrency probe (development aid)."""
import os, sys, time, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-3d-pointcloud-segmentation_b200"))
import torch
import f3ps
from f3ps import synth
frame = synth.make_frame(seed=20020)
npts = len(frame)
mp = dict(color_mode=0, geom_mode=1, merge_mode=1)
def mk():
    s = f3ps.Segmenter(); s.set_vccs_params(); s.set_merge_params(**mp); s.set_input(frame); s.run(0.2); return s
def timed(segs, fn, reps):
    def w(s):
        for _ in range(reps): fn(s)
    th = [threading.Thread(target=w, args=(s,)) for s in segs]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / (reps * len(segs)) * 1e3
segs = [mk() for _ in range(32)]
def front(s):
    s.set_input(frame); s.voxelize(); s.neighbors(); s.normals(); s.seeds()
def front_expand(s):
    front(s); s.expand()
def extract(s):
    front(s); s.expand(); s.graph()
for F in (1, 4, 16, 32):
    sub = segs[:F]
    print("in flight %2d: merge only %.3f ms/frame | K1-K4 %.3f | K1-K5 %.3f | K1-K6 %.3f | full %.3f" % (
        F, timed(sub, lambda s: s.merge(0.2), 4), timed(sub, front, 4), timed(sub, front_expand, 4), timed(sub, extract, 4),
        timed(sub, lambda s: (s.set_input(frame), s.run(0.2)), 4)), flush=True)
