#!/usr/bin/env python
"""Throughput of one GPU vs the number of frames in flight (development probe)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-3d-pointcloud-segmentation_b200"))
import torch
import f3ps
from f3ps import synth, sweep
frames = [synth.make_frame(seed=20020 + i) for i in range(4)]
npts = len(frames[0])
dev = [torch.from_numpy(f.view(np.uint8).reshape(-1, 32).copy()).cuda() for f in frames]
mp = dict(color_mode=0, geom_mode=1, merge_mode=1)
for F in (1, 2, 4, 8, 16, 32, 64):
    pool = sweep.FramePool(F, merge=mp)
    K = max(2 * F, 16)
    ptrs = [dev[i % 4].data_ptr() for i in range(K)]
    pool.run(ptrs[:F], on_device=True, npts=npts)          # warm-up (buffer allocation)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); pool.run(ptrs, on_device=True, npts=npts); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    host = [frames[i % 4] for i in range(K)]
    t0 = time.perf_counter(); pool.run(host, collect=lambda s, k: (s.array("out_label").shape, s.array("merges_w").shape)); dt2 = time.perf_counter() - t0
    print("in flight %3d: resident %.2f ms/frame %.1f Mpoints/s | host buffers + read-back %.2f ms/frame %.1f Mpoints/s" % (
        F, dt / K * 1e3, npts * K / dt / 1e6, dt2 / K * 1e3, npts * K / dt2 / 1e6), flush=True)
    pool.close()
