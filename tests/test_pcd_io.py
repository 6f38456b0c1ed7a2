"""PCD readers (next row f2): the CLI's C++ reader (host/pcd_io.cpp, stands in for pcl::io::loadPCDFile,
/root/reference/src/supervoxel_clustering.cpp:313) against the Python reader on ascii / binary / binary_compressed files,
and both on the reference's bundled cloud (a real PCL-written binary_compressed file, tests/fixtures/)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "fast-3d-pointcloud-segmentation_b200", "host")
FIXTURE = os.path.join(ROOT, "tests", "fixtures", "milk_cartoon_all_small_clorox.pcd")


def fnv(pts, label):
    """the checksum pcd_dump prints: FNV-1a over the little-endian bytes of x, y, z (NaNs canonical), rgba, label per point"""
    n = len(pts)
    rec = np.zeros((n, 5), np.uint32)
    for k, nm in enumerate(("x", "y", "z")):
        v = np.ascontiguousarray(pts[nm], np.float32)
        rec[:, k] = np.where(np.isnan(v), np.uint32(0x7fc00000), v.view(np.uint32))
    rec[:, 3] = pts["rgba"]
    rec[:, 4] = 0 if label is None else label
    h = 1469598103934665603
    for b in rec.tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


@pytest.fixture(scope="module")
def pcd_dump():
    subprocess.check_call(["make", "-C", HOST, "-s", "pcd_dump"])
    return os.path.join(HOST, "pcd_dump")


def dump(exe, path):
    out = subprocess.run([exe, path], capture_output=True, text=True)
    return out.returncode, out.stdout.split()


def small_cloud(with_nan=True):
    from f3ps import synth
    pts = synth.make_frame(seed=7, width=40, height=30)
    rng = np.random.default_rng(5)
    label = rng.integers(0, 9, len(pts)).astype(np.uint32)
    assert np.isnan(pts["x"]).any() or not with_nan
    return pts, label


@pytest.mark.parametrize("mode", ["ascii", "binary", "binary_compressed"])
@pytest.mark.parametrize("with_label", [False, True], ids=["xyzrgba", "xyzrgba_label"])
def test_cpp_reader_matches_python_reader(pcd_dump, tmp_path, mode, with_label):
    from f3ps import pcd
    pts, label = small_cloud()
    path = str(tmp_path / ("c_%s.pcd" % mode))
    pcd.write_pcd(path, pts, label if with_label else None, mode=mode)
    back, blabel, hdr = pcd.read_pcd(path)
    assert hdr["DATA"][0] == mode
    for nm in ("x", "y", "z"):
        assert np.array_equal(np.isnan(back[nm]), np.isnan(pts[nm]))
        ok = ~np.isnan(pts[nm])
        assert np.array_equal(back[nm][ok], pts[nm][ok]), nm        # repr(float) round-trips float32 exactly in ascii
    assert np.array_equal(back["rgba"], pts["rgba"])
    assert (blabel is None) == (not with_label) and (blabel is None or np.array_equal(blabel, label))
    rc, out = dump(pcd_dump, path)
    assert rc == 0 and int(out[0]) == len(pts) and int(out[3]) == int(np.isfinite(pts["x"]).sum())
    assert out[4] == fnv(pts, label if with_label else None)


def test_cpp_reader_ascii_rgb_float_field_as_pcl_writes_it(pcd_dump, tmp_path):
    """PCL >= 1.8 prints an `rgb` field of TYPE F in ascii as the uint32 reinterpretation of the packed colour."""
    from f3ps import pcd
    pts, _ = small_cloud()
    path = str(tmp_path / "rgbf.pcd")
    pcd.write_pcd(path, pts, None, mode="ascii", rgb_as_float=True)
    assert "4 4 4 4" in open(path).read(400) and "FIELDS x y z rgb\n" in open(path).read(400)
    rc, out = dump(pcd_dump, path)
    assert rc == 0 and out[4] == fnv(pts, None)


def test_cpp_reader_rejects_truncated_and_oversized_files(pcd_dump, tmp_path):
    from f3ps import pcd
    pts, _ = small_cloud()
    path = str(tmp_path / "t.pcd")
    pcd.write_pcd(path, pts, None, mode="binary")
    raw = open(path, "rb").read()
    open(path, "wb").write(raw[:len(raw) // 2])
    assert dump(pcd_dump, path)[0] == 1
    open(path, "wb").write(raw.replace(b"POINTS %d" % len(pts), b"POINTS 4000000000"))
    assert dump(pcd_dump, path)[0] == 1
    assert dump(pcd_dump, str(tmp_path / "missing.pcd"))[0] == 1


def test_bundled_cloud_both_readers(pcd_dump):
    """The reference's sample (640x480 organised, binary_compressed, FIELDS x y z rgba)."""
    from f3ps import pcd
    pts, label, hdr = pcd.read_pcd(FIXTURE)
    assert hdr["DATA"][0] == "binary_compressed" and hdr["FIELDS"] == ["x", "y", "z", "rgba"] and label is None
    assert len(pts) == 307200 and int(np.isfinite(pts["z"]).sum()) == 241407
    z = pts["z"][np.isfinite(pts["z"])]
    assert (z < 0).all() and abs(z.min() + 2.063) < 1e-3 and abs(z.max() + 0.501) < 1e-3
    rc, out = dump(pcd_dump, FIXTURE)
    assert rc == 0 and out[:4] == ["307200", "640", "480", "241407"] and out[4] == fnv(pts, None)
