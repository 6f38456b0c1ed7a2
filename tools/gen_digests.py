"""Oracle digests of the full-size configurations (BASELINE configs[3] and a 10 M-point configs[4] cloud) for
tests/test_gpu_fullsize.py: the CPU oracle (stamp-based merge, identical results to the literal replay) takes minutes at these
sizes, so its arrays are reduced HERE to SHA-256 digests (exact arrays) and float64 sums (arrays whose last bit depends on libm)
and committed as tests/golden/fullsize_digests.json.  Run in this container: python tools/gen_digests.py [c4] [c5]"""
import hashlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fast-3d-pointcloud-segmentation_b200"))
from oracle.oracle_py import Oracle
from f3ps import synth

EXACT = ["keys", "voxel_count", "nbr_count", "seeds", "labels", "dist", "sv_label", "sv_count", "edges_ab", "merges_ab", "merges_left", "final_ab",
         "out_label", "out_voxel"]
SUMS = ["voxel_xyz", "normals", "edges_dg", "edges_w", "merges_w", "final_w"]
CONFIGS = {
    "c4": dict(scene="make_dense_scene(seed=40000)", vccs=dict(voxel_res=0.004, seed_res=0.04), merge=dict(color_mode=1, geom_mode=0, merge_mode=0, lam=0.5), thr=0.2),
    "c5": dict(scene="make_room_scan(seed=50000, n_points=10_000_000)", vccs=dict(voxel_res=0.01, seed_res=0.1), merge=dict(color_mode=0, geom_mode=1, merge_mode=1), thr=0.2),
}


def sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def fsum(a):
    a = np.asarray(a, np.float64)
    return float(np.nansum(a)), int(np.isnan(a).sum())


def main():
    out_path = os.path.join(ROOT, "tests", "golden", "fullsize_digests.json")
    out = json.load(open(out_path)) if os.path.exists(out_path) else {}
    for name in (sys.argv[1:] or list(CONFIGS)):
        cfg = CONFIGS[name]
        pts = eval("synth." + cfg["scene"])
        o = Oracle(); o.set_vccs_params(**cfg["vccs"]); o.set_merge_params(merge_impl=1, **cfg["merge"]); o.set_input(pts)
        t0 = time.time(); o.run(0, cfg["thr"]); dt = time.time() - t0
        rec = {"scene": cfg["scene"], "vccs": cfg["vccs"], "merge": cfg["merge"], "threshold": cfg["thr"], "oracle_seconds": round(dt, 1),
               "n_points": int(len(pts)), "shape": {n: list(o.array(n).shape) for n in EXACT + SUMS},
               "sha256": {n: sha(o.array(n)) for n in EXACT}, "sum": {n: fsum(o.array(n)) for n in SUMS},
               "scalars": {k: (int(v) if float(v).is_integer() else float(v)) for k, v in o.scalars().items()}}
        print(name, "oracle", dt, "s", rec["shape"], flush=True)
        out[name] = rec
        json.dump(out, open(out_path, "w"), indent=1, sort_keys=True)


main()
