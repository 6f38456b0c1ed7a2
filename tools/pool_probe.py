"""BatchPool.run wall time per frame, without the bench around it."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-3d-pointcloud-segmentation_b200"))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
import f3ps
from f3ps import synth, sweep
F = int(sys.argv[1]); R = int(sys.argv[2]); cl = int(sys.argv[3]); ctas = int(sys.argv[4]) if len(sys.argv) > 4 else 0
workers = int(sys.argv[5]) if len(sys.argv) > 5 else None
e2e = int(sys.argv[6]) if len(sys.argv) > 6 else 0       # 1: pinned host buffers in, results read back (the bench's e2e leg)
frames = [synth.make_frame(seed=20020 + i) for i in range(8)]
npts = len(frames[0])
d = [torch.from_numpy(frames[i % 8].view(np.uint8).reshape(-1, 32).copy()).cuda() for i in range(F)]
ptrs = [d[i % F].data_ptr() for i in range(F * R)]
pinned = [torch.from_numpy(frames[i % 8].view(np.uint8).reshape(-1, 32).copy()).pin_memory() for i in range(F)]
host_views = [pinned[i % F].numpy().view(synth.POINT_DTYPE).reshape(-1) for i in range(F * R)]
def collect(seg, k):
    r = seg.fetch_result()
    return int(r['out_label'].shape[0])
pool = sweep.BatchPool(batch=F, workers=workers, device=0, merge=dict(color_mode=0, geom_mode=1, merge_mode=1), threshold=0.2, expand_ctas=-ctas, expand_cluster=cl)
for rep in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if e2e: pool.run(host_views, collect=collect)
    else: pool.run(ptrs, on_device=True, npts=npts)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("e2e %d workers %s F %d R %d cluster %d ctas %d rep %d: wall %.1f ms = %.3f ms/frame -> %.1f Mpoints/s; stage %s" % (
        e2e, workers, F, R, cl, ctas, rep, (t1 - t0) * 1e3, (t1 - t0) * 1e3 / (F * R), npts * F * R / (t1 - t0) / 1e6,
        {k: round(v, 2) for k, v in pool.segs[0].stage_ms().items()}), flush=True)
print("timeline (ms):", [(w, g, round(a * 1e3, 1), round(b * 1e3, 1)) for w, g, a, b in sorted(pool.timeline, key=lambda x: x[2])])
