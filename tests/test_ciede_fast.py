"""CPU checks of the product's branch-free FP64 CIEDE2000 (csrc/ciede_fast.h, compiled for the host):
elementary functions against libm within 2 ulp on the ranges the formula produces, and the whole formula
against the oracle's literal restatement of ColorUtilities::lab_ciede00 (src/color_utilities.cpp:190-294)
and the reference's own 34 known-answer vectors (:357-458)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "ciede_fast_host.cpp")
LIB = os.path.join(HERE, "native", "libciede_fast_host.so")


@pytest.fixture(scope="module")
def cf():
    hdr = os.path.join(HERE, "..", "fast-3d-pointcloud-segmentation_b200", "csrc", "ciede_fast.h")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-mfma", "-o", LIB, SRC])
    L = C.CDLL(LIB)
    for name in ("cf_ciede_batch", "cf_sincos_batch", "cf_atan2_batch", "cf_exp_batch"):
        getattr(L, name).restype = None
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def ulp_err(a, ref):
    return np.abs(a - ref) / np.maximum(np.spacing(np.abs(ref)), 5e-324)


def test_sincos_within_2ulp(cf):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-30, 30, 400000), rng.uniform(-1e-3, 1e-3, 50000), np.arange(-40, 41) * (np.pi / 4), [0.0]])
    s = np.empty_like(x); c = np.empty_like(x)
    cf.cf_sincos_batch(_p(x), _p(s), _p(c), C.c_long(len(x)))
    # absolute error relative to ulp(1) near zeros of the function (libm is correctly rounded there, a reduction is not)
    assert np.max(np.minimum(ulp_err(s, np.sin(x)), np.abs(s - np.sin(x)) / 2.3e-16)) <= 2.0
    assert np.max(np.minimum(ulp_err(c, np.cos(x)), np.abs(c - np.cos(x)) / 2.3e-16)) <= 2.0
    assert s[-1] == 0.0 and c[-1] == 1.0


def test_atan2_within_2ulp(cf):
    rng = np.random.default_rng(2)
    y = np.concatenate([rng.uniform(-130, 130, 400000), np.zeros(10), rng.uniform(-1e-3, 1e-3, 20000)])
    x = np.concatenate([rng.uniform(-200, 200, 400000), rng.uniform(-5, 5, 10), rng.uniform(-130, 130, 20000)])
    x[x == 0] = 1.0
    o = np.empty_like(x)
    cf.cf_atan2_batch(_p(y), _p(x), _p(o), C.c_long(len(x)))
    assert np.max(ulp_err(o, np.arctan2(y, x))) <= 2.0


def test_exp_within_2ulp(cf):
    x = -np.random.default_rng(3).uniform(0, 130, 300000)
    o = np.empty_like(x)
    cf.cf_exp_batch(_p(x), _p(o), C.c_long(len(x)))
    assert np.max(ulp_err(o, np.exp(x))) <= 2.0


def lab_pairs(oracle_mod, n, rng, near):
    o = oracle_mod.Oracle()
    rgb1 = rng.uniform(0, 255, (n, 3)).astype(np.float32)
    rgb2 = (rgb1 + rng.normal(0, 3.0, (n, 3)).astype(np.float32)).clip(0, 255).astype(np.float32) if near else rng.uniform(0, 255, (n, 3)).astype(np.float32)
    return o.rgb2lab_batch(rgb1), o.rgb2lab_batch(rgb2), o


def test_formula_matches_oracle(cf, oracle_mod):
    rng = np.random.default_rng(4)
    total = bad = 0
    for near in (False, True):
        l1, l2, o = lab_pairs(oracle_mod, 150000, rng, near)
        want = o.lab_ciede00_batch(l1, l2)
        got = np.empty(len(l1), np.float32)
        cf.cf_ciede_batch(_p(np.ascontiguousarray(l1)), _p(np.ascontiguousarray(l2)), _p(got), C.c_long(len(l1)))
        ne = got != want
        total += len(l1); bad += int(ne.sum())
        # any disagreement is a last-bit double difference surviving the cast to float: one float ulp at most
        assert np.all(np.abs(got[ne] - want[ne]) <= np.spacing(np.abs(want[ne])))
    assert bad <= total * 1e-5, (bad, total)
    # identical colours, greys (C' == 0 branch) and the black/white extremes
    same = np.array([[50, 10, -20], [0, 0, 0], [100, 0, 0], [35.5, 0, 0]], np.float32)
    got = np.empty(4, np.float32)
    cf.cf_ciede_batch(_p(same), _p(same.copy()), _p(got), C.c_long(4))
    assert np.all(got == 0)
    g1 = np.array([[20, 0, 0], [0, 0, 0]], np.float32); g2 = np.array([[80, 0, 0], [100, 0, 0]], np.float32)
    got = np.empty(2, np.float32)
    cf.cf_ciede_batch(_p(g1), _p(g2), _p(got), C.c_long(2))
    assert np.array_equal(got, oracle_mod.Oracle().lab_ciede00_batch(g1, g2))


def test_reference_known_answers(cf):
    from test_oracle_color import CIEDE_KAT
    vec = np.array(CIEDE_KAT, np.float64)
    l1 = np.ascontiguousarray(vec[:, 0:3], np.float32); l2 = np.ascontiguousarray(vec[:, 3:6], np.float32)
    want = vec[:, 6].astype(np.float32)
    got = np.empty(len(vec), np.float32)
    cf.cf_ciede_batch(_p(l1), _p(l2), _p(got), C.c_long(len(vec)))
    assert np.max(np.abs(got - want)) < 1e-4      # the table is rounded to 4 decimals (src/color_utilities.cpp:357-458)
