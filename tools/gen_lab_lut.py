#!/usr/bin/env python
"""Generate the 33^3 x 3 int16 lattice that OpenCV's float RGB->Lab path interpolates.

The reference converts region mean colours with cv::cvtColor(CV_32FC3, COLOR_RGB2Lab)
(/root/reference/src/color_utilities.cpp:52-69,151-160).  OpenCV 4.x evaluates that
call with a 33x33x33 lattice of fixed-point Lab values and integer trilinear
interpolation.  The lattice values are recovered exactly by
converting the lattice points themselves: at a lattice point all interpolation
weight falls on one corner.

Writes  fast-3d-pointcloud-segmentation_b200/data/lab_lut_s16.bin
        (33*33*33*3 little-endian int16, index [r][g][b][L,a,b])
and a few known-answer colours into tests/golden/lab_kat.json.
Needs cv2 (present in the build image: 4.13.0).  Run from the repo root.
"""
import json, os, sys
import numpy as np
import cv2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def main():
    g = np.arange(33, dtype=np.float32) / np.float32(32.0)
    r, gg, b = np.meshgrid(g, g, g, indexing="ij")
    rgb = np.stack([r, gg, b], axis=-1).reshape(1, -1, 3).astype(np.float32)
    lab = cv2.cvtColor(rgb, cv2.COLOR_RGB2Lab).reshape(-1, 3).astype(np.float64)
    L = lab[:, 0] / 100.0 * 16384.0
    a = (lab[:, 1] + 128.0) / 256.0 * 16384.0
    bb = (lab[:, 2] + 128.0) / 256.0 * 16384.0
    q = np.stack([L, a, bb], axis=-1)
    qi = np.rint(q)
    resid = np.abs(q - qi).max()
    assert resid == 0.0, "lattice values are not exact fixed-point numbers: %g" % resid
    assert qi.min() >= -32768 and qi.max() <= 32767
    lut = qi.astype("<i2")
    out = os.path.join(ROOT, "fast-3d-pointcloud-segmentation_b200", "data", "lab_lut_s16.bin")
    lut.tofile(out)
    print("wrote", out, lut.nbytes, "bytes; cv2", cv2.__version__)

    # known-answer vectors (inputs are 0..255 floats as the reference passes them)
    rng = np.random.default_rng(7)
    cols = np.concatenate([
        np.array([[123, 10, 200], [0, 0, 0], [255, 255, 255], [255, 255, 0]], np.float32),
        (rng.random((60, 3)) * 255).astype(np.float32)])
    scaled = (cols / np.float32(255)).astype(np.float32)
    labk = cv2.cvtColor(scaled.reshape(1, -1, 3), cv2.COLOR_RGB2Lab).reshape(-1, 3)
    kat = {"cv2_version": cv2.__version__,
           "rgb255_f32_hex": [[float(v).hex() for v in c] for c in cols],
           "lab_f32_hex": [[float(v).hex() for v in c] for c in labk]}
    with open(os.path.join(ROOT, "tests", "golden", "lab_kat.json"), "w") as f:
        json.dump(kat, f, indent=0)
    print("wrote tests/golden/lab_kat.json", len(cols), "vectors")

if __name__ == "__main__":
    sys.exit(main())
