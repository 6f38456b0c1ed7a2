"""The evaluation restatement (oracle/oracle_testing.py) against hand-computed cases of the reference's formulas
(src/testing.cpp:239-362).  The reference ships no fixture for this module."""
import math

import numpy as np

import oracle_testing as ot


def _cloud(n):
    return np.stack([np.arange(n, dtype=np.float32), np.zeros(n, np.float32), np.ones(n, np.float32)], 1)


def test_perfect_segmentation_scores():
    xyz = _cloud(10)
    truth = np.array([5] * 4 + [9] * 6)
    seg = np.array([0] * 4 + [1] * 6)
    s = ot.scores(xyz, seg, xyz, truth)
    assert abs(s["precision"] - 1) < 1e-6 and abs(s["recall"] - 1) < 1e-6 and abs(s["fscore"] - 1) < 1e-6
    assert abs(s["wov"] - 1) < 1e-6 and s["fpr"] == 0 and s["fnr"] == 0 and abs(s["voi"]) < 1e-6


def test_hand_computed_case():
    # truth: A = 6 points, B = 4 points; segmentation: s0 = first 5, s1 = last 5  -> inter = [[5, 0], [1, 4]]
    xyz = _cloud(10)
    truth = np.array([0] * 6 + [1] * 4)
    seg = np.array([0] * 5 + [1] * 5)
    s = ot.scores(xyz, seg, xyz, truth)
    # matches: A (larger) -> s0 (5), B -> s1 (4)
    p = (5 * 6 / 5 + 4 * 4 / 5) / 10
    r = (5 + 4) / 10
    assert abs(s["precision"] - p) < 1e-6 and abs(s["recall"] - r) < 1e-6
    assert abs(s["fscore"] - 2 * p * r / (p + r)) < 1e-6
    assert abs(s["fpr"] - (0 + 1) / 10) < 1e-6 and abs(s["fnr"] - (1 + 0) / 10) < 1e-6
    assert abs(s["wov"] - (5 * 6 / 6 + 4 * 4 / 5) / 10) < 1e-6
    hs = -2 * 0.5 * math.log(0.5); ht = -(0.6 * math.log(0.6) + 0.4 * math.log(0.4))
    mi = 0.5 * math.log(10 * 5 / (5 * 6)) + 0.1 * math.log(10 * 1 / (5 * 6)) + 0.4 * math.log(10 * 4 / (5 * 4))
    assert abs(s["voi"] - (hs + ht - 2 * mi)) < 1e-5


def test_equal_sized_truth_segments_quirk():
    # std::map<size_t, uint32_t>::insert keeps only the FIRST truth segment of a given size (testing.cpp:97-110):
    # the second 5-point truth segment is never matched and counts as false negatives.
    xyz = _cloud(10)
    truth = np.array([0] * 5 + [1] * 5)
    seg = truth.copy()
    s = ot.scores(xyz, seg, xyz, truth)
    assert abs(s["recall"] - 0.5) < 1e-6 and abs(s["fnr"] - 0.5) < 1e-6 and abs(s["precision"] - 0.5) < 1e-6


def test_partial_segmentation_and_label_gaps():
    # the segmentation covers 8 of 10 truth points (unowned voxels are absent), labels are renumbered densely
    xyz = _cloud(10)
    truth = np.array([3] * 7 + [8] * 3)
    keep = np.array([0, 1, 2, 3, 4, 5, 7, 8])
    seg = np.array([10, 10, 10, 10, 10, 10, 40, 40])
    s = ot.scores(xyz[keep], seg, xyz, truth)
    assert abs(s["recall"] - (6 + 2) / 10) < 1e-6 and abs(s["fnr"] - (1 + 1) / 10) < 1e-6 and s["fpr"] == 0


def test_sweep_thresholds_float_steps():
    thr = [np.float32(0.8)]
    t = np.float32(np.float32(0.8) + np.float32(0.005))
    while t <= np.float32(1.0):
        thr.append(t); t = np.float32(t + np.float32(0.005))
    assert len(thr) == 41 and abs(float(thr[-1]) - 0.99999982) < 1e-6
