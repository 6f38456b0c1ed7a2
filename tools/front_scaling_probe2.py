"""Which front-end stage is the serialised resource of a sweep?  frames/s of (a) K1..K4 only, (b) K5 only, vs host threads."""
import sys, time, threading
sys.path.insert(0, "fast-3d-pointcloud-segmentation_b200")
import numpy as np, torch
import f3ps
from f3ps import synth
pts = synth.make_frame(seed=20020)
d = torch.from_numpy(pts.view(np.uint8).reshape(-1, 32).copy()).cuda()
n = len(pts)
FL = dict(color_mode=0, geom_mode=1, merge_mode=1)
for what in ("k1-k4", "k5", "k6"):
    for T in (1, 4, 16):
        segs = [f3ps.Segmenter() for _ in range(T)]
        for s in segs:
            s.set_vccs_params(); s.set_merge_params(**FL); s.set_blocking_wait(True)
            s.set_input_device(d.data_ptr(), n, 32); s.extract(); s.graph()
        reps = 24
        def work(s):
            for _ in range(reps):
                if what == "k1-k4":
                    s.set_input_device(d.data_ptr(), n, 32); s.voxelize(); s.neighbors(); s.normals(); s.seeds()
                elif what == "k5":
                    s.expand()
                else:
                    s.graph()
        if what == "k5":
            for s in segs:
                s.set_input_device(d.data_ptr(), n, 32); s.voxelize(); s.neighbors(); s.normals(); s.seeds()
        th = [threading.Thread(target=work, args=(s,)) for s in segs]
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("%s threads=%2d  %.0f frames/s  %.3f ms/frame  per-thread latency %.2f ms" % (what, T, T * reps / dt, dt / (T * reps) * 1e3, dt / reps * 1e3))
        for s in segs: s.close()
