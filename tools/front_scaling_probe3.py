"""frames/s of each front stage vs host threads, K5 as cooperative grid or as one cluster (which stage caps a sweep?)"""
import sys, time, threading, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fast-3d-pointcloud-segmentation_b200"))
import numpy as np, torch
import f3ps
from f3ps import synth
pts = synth.make_frame(seed=20020)
d = torch.from_numpy(pts.view(np.uint8).reshape(-1, 32).copy()).cuda()
n = len(pts)
FL = dict(color_mode=0, geom_mode=1, merge_mode=1)
STAGES = {"k1": lambda s: (s.set_input_device(d.data_ptr(), n, 32), s.voxelize()), "k2": lambda s: s.neighbors(), "k3": lambda s: s.normals(),
          "k4": lambda s: s.seeds(), "k5coop24": lambda s: s.expand(), "k5cl16": lambda s: s.expand(), "k5cl4": lambda s: s.expand(), "k5cl1": lambda s: s.expand(), "k6": lambda s: s.graph(),
          "all": lambda s: (s.set_input_device(d.data_ptr(), n, 32), s.extract(), s.graph())}
for what in sys.argv[1].split(","):
    for T in [int(x) for x in sys.argv[2].split(",")]:
        segs = [f3ps.Segmenter() for _ in range(T)]
        for s in segs:
            s.set_vccs_params(); s.set_merge_params(**FL); s.set_blocking_wait(True)
            if what.startswith("k5cl"): s.set_expand_kernel(2, int(what[4:]))
            if what == "k5coop24": s.set_expand_sharing(24, 0)
            if what == "all": s.set_expand_kernel(2, 8)
            s.set_input_device(d.data_ptr(), n, 32); s.extract(); s.graph()
        reps = 16
        def work(s):
            for _ in range(reps): STAGES[what](s)
        th = [threading.Thread(target=work, args=(s,)) for s in segs]
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("%-9s threads=%2d  %6.0f frames/s  %.3f ms/frame  per-thread latency %.2f ms" % (what, T, T * reps / dt, dt / (T * reps) * 1e3, dt / reps * 1e3), flush=True)
        for s in segs: s.close()
