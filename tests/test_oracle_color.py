"""Pin the oracle's colour kernels against the reference's own known-answer vectors
(src/color_utilities.cpp:324-460) and against cv2 (tests/golden/lab_kat.json)."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))

# the 34 Sharma-Wu-Dalal vectors of ColorUtilities::lab_test, src/color_utilities.cpp:357-458
CIEDE_KAT = [
    (50.0000, 2.6772, -79.7751, 50.0000, 0.0000, -82.7485, 2.0425),
    (50.0000, 3.1571, -77.2803, 50.0000, 0.0000, -82.7485, 2.8615),
    (50.0000, 2.8361, -74.0200, 50.0000, 0.0000, -82.7485, 3.4412),
    (50.0000, -1.3802, -84.2814, 50.0000, 0.0000, -82.7485, 1.0000),
    (50.0000, -1.1848, -84.8006, 50.0000, 0.0000, -82.7485, 1.0000),
    (50.0000, -0.9009, -85.5211, 50.0000, 0.0000, -82.7485, 1.0000),
    (50.0000, 0.0000, 0.0000, 50.0000, -1.0000, 2.0000, 2.3669),
    (50.0000, -1.0000, 2.0000, 50.0000, 0.0000, 0.0000, 2.3669),
    (50.0000, 2.4900, -0.0010, 50.0000, -2.4900, 0.0009, 7.1792),
    (50.0000, 2.4900, -0.0010, 50.0000, -2.4900, 0.0010, 7.1792),
    (50.0000, 2.4900, -0.0010, 50.0000, -2.4900, 0.0011, 7.2195),
    (50.0000, 2.4900, -0.0010, 50.0000, -2.4900, 0.0012, 7.2195),
    (50.0000, -0.0010, 2.4900, 50.0000, 0.0009, -2.4900, 4.8045),
    (50.0000, -0.0010, 2.4900, 50.0000, 0.0010, -2.4900, 4.8045),
    (50.0000, -0.0010, 2.4900, 50.0000, 0.0011, -2.4900, 4.7461),
    (50.0000, 2.5000, 0.0000, 50.0000, 0.0000, -2.5000, 4.3065),
    (50.0000, 2.5000, 0.0000, 73.0000, 25.0000, -18.0000, 27.1492),
    (50.0000, 2.5000, 0.0000, 61.0000, -5.0000, 29.0000, 22.8977),
    (50.0000, 2.5000, 0.0000, 56.0000, -27.0000, -3.0000, 31.9030),
    (50.0000, 2.5000, 0.0000, 58.0000, 24.0000, 15.0000, 19.4535),
    (50.0000, 2.5000, 0.0000, 50.0000, 3.1736, 0.5854, 1.0000),
    (50.0000, 2.5000, 0.0000, 50.0000, 3.2972, 0.0000, 1.0000),
    (50.0000, 2.5000, 0.0000, 50.0000, 1.8634, 0.5757, 1.0000),
    (50.0000, 2.5000, 0.0000, 50.0000, 3.2592, 0.3350, 1.0000),
    (60.2574, -34.0099, 36.2677, 60.4626, -34.1751, 39.4387, 1.2644),
    (63.0109, -31.0961, -5.8663, 62.8187, -29.7946, -4.0864, 1.2630),
    (61.2901, 3.7196, -5.3901, 61.4292, 2.2480, -4.9620, 1.8731),
    (35.0831, -44.1164, 3.7933, 35.0232, -40.0716, 1.5901, 1.8645),
    (22.7233, 20.0904, -46.6940, 23.0331, 14.9730, -42.5619, 2.0373),
    (36.4612, 47.8580, 18.3852, 36.2715, 50.5065, 21.2231, 1.4146),
    (90.8027, -2.0831, 1.4410, 91.1528, -1.6435, 0.0447, 1.4441),
    (90.9257, -0.5406, -0.9208, 88.6381, -0.8985, -0.7239, 1.5381),
    (6.7747, -0.2908, -2.4247, 5.8714, -0.0985, -2.2286, 0.6377),
    (2.0776, 0.0795, -1.1350, 0.9033, -0.0636, -0.5514, 0.9082),
]

# ColorUtilities::rgb_test, src/color_utilities.cpp:326-348
RGB_KAT = [
    ((0, 0, 0), (0, 0, 0), 0.0),
    ((0, 0, 0), (255, 255, 255), 441.672943),
    ((255, 255, 255), (255, 255, 255), 0.0),
    ((0, 0, 0), (255, 0, 0), 255.0),
    ((0, 255, 0), (0, 0, 0), 255.0),
    ((0, 255, 0), (255, 0, 255), 441.672943),
    ((100, 20, 35), (104, 20, 32), 5.0),
]


def test_ciede2000_reference_vectors(oracle_mod):
    o = oracle_mod.Oracle()
    worst = 0.0
    for L1, a1, b1, L2, a2, b2, exp in CIEDE_KAT:
        got = o.lab_ciede00([L1, a1, b1], [L2, a2, b2])
        worst = max(worst, abs(got - exp))
    assert worst < 1e-4, worst          # the table is rounded to 4 decimals


def test_rgb_eucl_reference_vectors(oracle_mod):
    o = oracle_mod.Oracle()
    for c1, c2, exp in RGB_KAT:
        got = o.rgb_eucl(c1, c2)
        assert got == np.float32(exp), (c1, c2, got, exp)


def test_rgb2lab_golden_vectors(oracle_mod):
    """Bit-exact against cv2.cvtColor(float32, COLOR_RGB2Lab) outputs recorded by tools/gen_lab_lut.py."""
    with open(os.path.join(HERE, "golden", "lab_kat.json")) as f:
        kat = json.load(f)
    o = oracle_mod.Oracle()
    for rgb_hex, lab_hex in zip(kat["rgb255_f32_hex"], kat["lab_f32_hex"]):
        rgb = np.array([float.fromhex(h) for h in rgb_hex], np.float32)
        lab = np.array([float.fromhex(h) for h in lab_hex], np.float32)
        got = o.rgb2lab(rgb)
        assert np.array_equal(got, lab), (rgb, got, lab)


def test_rgb2lab_against_cv2_live(oracle_mod):
    cv2 = pytest.importorskip("cv2")
    o = oracle_mod.Oracle()
    rng = np.random.default_rng(3)
    cols = (rng.random((4000, 3)) * 255).astype(np.float32)
    ref = cv2.cvtColor((cols / np.float32(255)).reshape(1, -1, 3), cv2.COLOR_RGB2Lab).reshape(-1, 3)
    for c, r in zip(cols, ref):
        assert np.array_equal(o.rgb2lab(c), r)
    # convert_test's four colours (src/color_utilities.cpp:468-492)
    assert np.allclose(o.rgb2lab([255, 255, 255]), [100, 0, 0])
    assert np.array_equal(o.rgb2lab([0, 0, 0]), np.zeros(3, np.float32))
