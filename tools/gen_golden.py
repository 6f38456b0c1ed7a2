#!/usr/bin/env python
"""Generate tests/golden/small_frame.npz: a 160x120 synthetic frame and the CPU oracle's outputs for it
(literal std::multimap merge), for --CVX --AL -t 0.2 and --EQ 200 -t 0.6.  The GPU parity tests and the
oracle regression test compare against this file.  Run from the repo root after building the oracle."""
import hashlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "fast-3d-pointcloud-segmentation_b200"))
import oracle_py
from f3ps import synth

def digest(a):
    """sha256 of an array with every NaN replaced by one canonical NaN (payload/sign of a NaN is not part of parity)."""
    a = np.ascontiguousarray(a)
    if a.dtype.kind == "f":
        a = np.where(np.isnan(a), np.float32(np.nan), a).astype(np.float32)
        a = a.view(np.uint32).copy()
        a[a == 0x80000000] = 0          # -0.0 == +0.0
    return hashlib.sha256(a.tobytes()).digest()

def bits(a): return np.ascontiguousarray(a, np.float32).view(np.uint32)

def main():
    xyz, rgba = synth.make_frame(seed=11, width=160, height=120, as_struct=False)
    pts = synth.pack_points(xyz, rgba)
    out = {"xyz": xyz, "rgba": rgba}
    for tag, mp, thr in (("al", dict(color_mode=0, geom_mode=1, merge_mode=1), 0.2), ("eq", dict(color_mode=0, geom_mode=0, merge_mode=2, bins=200), 0.6)):
        o = oracle_py.Oracle(); o.set_vccs_params(); o.set_merge_params(merge_impl=0, **mp); o.set_input(pts); o.run(0, thr)
        if tag == "al":
            for n in ("keys", "voxel_count", "nbr_count", "seeds", "labels", "sv_label", "sv_count"):
                out[n] = o.array(n)
            for n in ("voxel_xyz", "voxel_rgb", "normals", "nbr", "dist"):       # digests of the big arrays
                out[n + "_sha256"] = np.frombuffer(digest(o.array(n)), np.uint8)
        out[tag + "_edges_ab"] = o.array("edges_ab"); out[tag + "_edges_w_bits"] = bits(o.array("edges_w"))
        out[tag + "_merges_ab"] = o.array("merges_ab"); out[tag + "_merges_w_bits"] = bits(o.array("merges_w"))
        out[tag + "_out_label"] = o.array("out_label"); out[tag + "_out_voxel"] = o.array("out_voxel")
        out[tag + "_lambda_bits"] = bits(np.array([o.scalars()["lambda"]], np.float32))
        out[tag + "_threshold"] = np.array([thr], np.float32)
        print(tag, "V", len(o.array("keys")), "S", len(o.array("sv_label")), "E", len(o.array("edges_w")), "M", len(o.array("merges_w")), "nan", o.scalars()["nan_weights"])
    path = os.path.join(ROOT, "tests", "golden", "small_frame.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")

if __name__ == "__main__":
    main()
