#!/usr/bin/env python
"""Timeline of individual merges of the resident K7 kernel (development probe): SM clock at fixed points of every role,
printed relative to the moment worker thread 0 knows the head.  Usage: python tools/k7_trace.py [first_merge] [seed]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-3d-pointcloud-segmentation_b200"))
import f3ps
from f3ps import synth

first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 20020
pts = synth.make_frame(seed=seed)
g = f3ps.Segmenter(); g.set_vccs_params(); g.set_merge_params(color_mode=0, geom_mode=1, merge_mode=1); g.set_input(pts)
g.extract(); g.graph(); g.set_merge_kernel(4)
for rep in range(2):
    g.merge_trace(first); g.merge(0.2); g.sync()
tr = g.merge_trace().astype(np.int64)
names = {0: "w.top", 1: "w.rescan", 2: "w.head", 3: "w.read", 4: "w.dc", 5: "w.F", 6: "w.D", 7: "w.W4",
         12: "c.head", 13: "c.fetch", 14: "c.fold", 15: "c.eig", 16: "m.head", 17: "m.guess", 18: "m.fetch", 19: "m.fold", 20: "m.lab"}
order = [0, 1, 2, 12, 16, 17, 3, 13, 18, 14, 19, 20, 15, 4, 5, 6, 7]
print("merge    T   |b|  " + " ".join("%8s" % names[k] for k in order) + "    total")
rows = []
for i in range(255):
    if tr[i, 2] == 0 or tr[i + 1, 0] == 0: continue
    base = tr[i, 2]
    rel = [(int(tr[i, k]) - int(base)) & 0xffffffff for k in order]
    rel = [r - (1 << 32) if r > (1 << 31) else r for r in rel]
    total = (int(tr[i + 1, 2]) - int(base)) & 0xffffffff
    rows.append((int(tr[i, 8]), int(tr[i, 21]), rel, total))
    if i < 48: print("%5d %4d %5d  " % (first + i, tr[i, 8], tr[i, 21]) + " ".join("%8d" % r for r in rel) + " %8d" % total)
small = [r for r in rows if r[0] <= 32 and r[1] <= 64]
if small:
    med = np.median(np.array([r[2] + [r[3]] for r in small]), axis=0)
    print("median over %d merges with T<=32, |b|<=64:" % len(small)); print("                   " + " ".join("%8d" % v for v in med))
big = [r for r in rows if r[0] > 128]
if big:
    med = np.median(np.array([r[2] + [r[3]] for r in big]), axis=0)
    print("median over %d merges with T>128:" % len(big)); print("                   " + " ".join("%8d" % v for v in med))
