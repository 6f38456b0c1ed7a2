"""ctypes wrapper over the CPU ORACLE (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LUT_PATH = os.path.join(ROOT, "fast-3d-pointcloud-segmentation_b200", "data", "lab_lut_s16.bin")

_lib = None

_DTYPES = {
    "keys": np.uint32, "morton": np.uint64, "voxel_xyz": np.float32, "voxel_rgb": np.float32,
    "voxel_rgba": np.uint32, "voxel_count": np.int32, "point_voxel": np.int32, "nbr": np.int32,
    "nbr_count": np.int32, "normals": np.float32, "curvature": np.float32, "seed_cells_nn": np.int32,
    "seeds": np.int32, "labels": np.uint32, "dist": np.float32, "steals": np.int32,
    "sv_label": np.uint32, "sv_xyz": np.float32, "sv_rgb": np.float32, "sv_normal": np.float32,
    "sv_count": np.int32, "adj": np.uint32, "cdf_c": np.float32, "cdf_g": np.float32,
    "out_xyz": np.float32, "out_label": np.uint32, "out_voxel": np.uint32,
    "edges_ab": np.uint32, "edges_dc": np.float32, "edges_dg": np.float32, "edges_w": np.float32,
    "merges_ab": np.uint32, "merges_w": np.float32, "merges_left": np.uint32,
    "final_ab": np.uint32, "final_w": np.float32, "stage_ms": np.float64, "scalars": np.float64,
}
_SHAPES = {"keys": 3, "voxel_xyz": 3, "voxel_rgb": 3, "nbr": 27, "normals": 4, "sv_xyz": 3, "sv_rgb": 3,
           "sv_normal": 4, "adj": 2, "out_xyz": 3, "edges_ab": 2, "merges_ab": 2, "merges_left": 2, "final_ab": 2}
SCALARS = ["depth", "bmin_x", "bmin_y", "bmin_z", "bmax_x", "bmax_y", "bmax_z", "seed_depth",
           "seed_min_x", "seed_min_y", "seed_min_z", "rounds", "lambda", "nan_weights", "n_segments"]


def build(force=False):
    so = os.path.join(HERE, "liboracle.so")
    srcs = [os.path.join(HERE, f) for f in ("oracle_vccs.cpp", "oracle_merge.cpp", "oracle_capi.cpp", "oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return so


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_void_p]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_last_error.restype = C.c_char_p
        L.orc_last_error.argtypes = [C.c_void_p]
        L.orc_set_vccs_params.argtypes = [C.c_void_p] + [C.c_float] * 5 + [C.c_int, C.c_int]
        L.orc_set_merge_params.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int]
        L.orc_set_expand_impl.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_switches.argtypes = [C.c_void_p] + [C.c_int] * 4
        L.orc_set_input.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_int]
        L.orc_run.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.orc_run.restype = C.c_int
        L.orc_array.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]
        L.orc_array.restype = C.c_long
        L.orc_set_graph.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p]
        L.orc_set_graph.restype = C.c_int
        L.orc_rgb2lab.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_lab_ciede00.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_lab_ciede00.restype = C.c_float
        L.orc_rgb2lab_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]
        L.orc_lab_ciede00_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]
        L.orc_rgb_eucl.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_rgb_eucl.restype = C.c_float
        L.orc_normals_diff.argtypes = [C.c_void_p] * 4
        L.orc_normals_diff.restype = C.c_float
        L.orc_is_convex.argtypes = [C.c_void_p] * 4
        L.orc_is_convex.restype = C.c_int
        L.orc_plane_from_accu.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_cr_logf.argtypes = [C.c_float]
        L.orc_cr_logf.restype = C.c_float
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Mirrors the product binding (f3ps.Segmenter) so parity tests read symmetrically."""

    def __init__(self):
        self.L = lib()
        self._lut = np.fromfile(LUT_PATH, dtype="<i2")
        assert self._lut.size == 33 * 33 * 33 * 3
        self.h = self.L.orc_create(_p(self._lut))

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    def set_vccs_params(self, voxel_res=0.008, seed_res=0.08, color=0.2, spatial=0.4, normal=1.0,
                        use_transform=True, fold_negative_z=True):
        self.L.orc_set_vccs_params(self.h, voxel_res, seed_res, color, spatial, normal,
                                   int(use_transform), int(fold_negative_z))

    def set_merge_params(self, color_mode=0, geom_mode=0, merge_mode=1, lam=0.5, bins=500, merge_impl=0):
        self.L.orc_set_merge_params(self.h, color_mode, geom_mode, merge_mode, lam, bins, merge_impl)

    def set_expand_impl(self, impl):
        self.L.orc_set_expand_impl(self.h, impl)

    def set_switches(self, leaf_desc=0, keybits_floor=0, init_seed_voxel=0, shifted_cov=0):
        self.L.orc_set_switches(self.h, leaf_desc, keybits_floor, init_seed_voxel, shifted_cov)

    def set_input(self, pts):
        pts = np.ascontiguousarray(pts)
        self._pts = pts
        self.L.orc_set_input(self.h, _p(pts), pts.shape[0], pts.dtype.itemsize)

    def run(self, stage=0, threshold=0.2):
        rc = self.L.orc_run(self.h, stage, threshold)
        if rc:
            raise RuntimeError(self.L.orc_last_error(self.h).decode())

    def refine(self, num_itr=3):
        """pcl::SupervoxelClustering::refineSupervoxels on the result of run(5); run(6) afterwards rebuilds the graph"""
        self.L.orc_refine.argtypes = [C.c_void_p, C.c_int]
        rc = self.L.orc_refine(self.h, num_itr)
        if rc:
            raise RuntimeError(self.L.orc_last_error(self.h).decode())

    def set_graph(self, vxyz, vrgba, labels, vox_lists, centroids, normals, adj_pairs):
        vxyz = np.ascontiguousarray(vxyz, np.float32)
        vrgba = np.ascontiguousarray(vrgba, np.uint32)
        labels = np.ascontiguousarray(labels, np.uint32)
        off = np.zeros(len(vox_lists) + 1, np.int64)
        off[1:] = np.cumsum([len(v) for v in vox_lists])
        idx = np.ascontiguousarray(np.concatenate(vox_lists) if len(vox_lists) else np.zeros(0), np.int32)
        centroids = np.ascontiguousarray(centroids, np.float32)
        normals = np.ascontiguousarray(normals, np.float32)
        adj_pairs = np.ascontiguousarray(adj_pairs, np.uint32).reshape(-1, 2)
        rc = self.L.orc_set_graph(self.h, vxyz.shape[0], _p(vxyz), _p(vrgba), len(labels), _p(labels), _p(off),
                                  _p(idx), _p(centroids), _p(normals), adj_pairs.shape[0], _p(adj_pairs))
        if rc:
            raise RuntimeError(self.L.orc_last_error(self.h).decode())

    def array(self, name):
        ptr = C.c_void_p()
        n = self.L.orc_array(self.h, name.encode(), C.byref(ptr))
        if n < 0:
            raise KeyError(name)
        dt = np.dtype(_DTYPES[name])
        if n == 0:
            out = np.zeros(0, dt)
        else:
            buf = (C.c_char * (n * dt.itemsize)).from_address(ptr.value)
            out = np.frombuffer(buf, dtype=dt).copy()
        k = _SHAPES.get(name)
        return out.reshape(-1, k) if k else out

    def scalars(self):
        return dict(zip(SCALARS, self.array("scalars")))

    # metric kernels
    def rgb2lab(self, rgb255):
        a = np.ascontiguousarray(rgb255, np.float32)
        o = np.zeros(3, np.float32)
        self.L.orc_rgb2lab(self.h, _p(a), _p(o))
        return o

    def lab_ciede00(self, l1, l2):
        a = np.ascontiguousarray(l1, np.float32)
        b = np.ascontiguousarray(l2, np.float32)
        return float(self.L.orc_lab_ciede00(_p(a), _p(b)))

    def rgb2lab_batch(self, rgb255):
        a = np.ascontiguousarray(rgb255, np.float32).reshape(-1, 3)
        o = np.zeros_like(a)
        self.L.orc_rgb2lab_batch(self.h, _p(a), _p(o), C.c_long(len(a)))
        return o

    def lab_ciede00_batch(self, l1, l2):
        a = np.ascontiguousarray(l1, np.float32).reshape(-1, 3)
        b = np.ascontiguousarray(l2, np.float32).reshape(-1, 3)
        o = np.zeros(len(a), np.float32)
        self.L.orc_lab_ciede00_batch(_p(a), _p(b), _p(o), C.c_long(len(a)))
        return o

    def rgb_eucl(self, c1, c2):
        a = np.ascontiguousarray(c1, np.float32)
        b = np.ascontiguousarray(c2, np.float32)
        return float(self.L.orc_rgb_eucl(_p(a), _p(b)))
