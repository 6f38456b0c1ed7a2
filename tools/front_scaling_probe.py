"""Does the front end (K1..K6) of a sweep scale with host threads?  frames/s of extract+graph vs threads, spin vs blocking waits."""
import sys, time, threading
sys.path.insert(0, "fast-3d-pointcloud-segmentation_b200")
import numpy as np, torch
import f3ps
from f3ps import synth
pts = synth.make_frame(seed=20020)
d = torch.from_numpy(pts.view(np.uint8).reshape(-1, 32).copy()).cuda()
n = len(pts)
FL = dict(color_mode=0, geom_mode=1, merge_mode=1)
for blocking in (True, False):
    for T in (1, 2, 4, 8, 16, 32):
        if not blocking and T > 16:
            continue
        segs = [f3ps.Segmenter() for _ in range(T)]
        for s in segs:
            s.set_vccs_params(); s.set_merge_params(**FL); s.set_blocking_wait(blocking)
            s.set_input_device(d.data_ptr(), n, 32); s.extract(); s.graph()
        reps = 24
        def work(s):
            for _ in range(reps):
                s.set_input_device(d.data_ptr(), n, 32); s.extract(); s.graph()
        th = [threading.Thread(target=work, args=(s,)) for s in segs]
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        st = segs[0].stage_ms()
        print("blocking=%d threads=%2d  %.0f frames/s  %.3f ms/frame  per-thread latency %.2f ms  launches/frame %d  stage(last) %s" % (
            blocking, T, T * reps / dt, dt / (T * reps) * 1e3, dt / reps * 1e3, segs[0].launch_count() // (reps + 1),
            {k: round(v, 2) for k, v in st.items() if k in ("voxelize", "neighbors", "normals", "seeds", "expand", "graph")}))
        for s in segs: s.close()
