"""CPU tests of the oracle itself: golden fixture regression, literal-vs-stamp merge equivalence,
structural invariants, the bundled frame of the reference when it is reachable."""
import hashlib
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "small_frame.npz")
AL = dict(color_mode=0, geom_mode=1, merge_mode=1)
EQ = dict(color_mode=0, geom_mode=0, merge_mode=2, bins=200)


def digest(a):
    """sha256 of an array with every NaN replaced by one canonical NaN (payload/sign of a NaN is not part of parity)."""
    a = np.ascontiguousarray(a)
    if a.dtype.kind == "f":
        a = np.where(np.isnan(a), np.float32(np.nan), a).astype(np.float32)
        a = a.view(np.uint32).copy()
        a[a == 0x80000000] = 0          # -0.0 == +0.0
    return hashlib.sha256(a.tobytes()).digest()


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def golden_points():
    from f3ps import synth
    g = np.load(GOLD)
    return g, synth.pack_points(g["xyz"], g["rgba"])


def run_oracle(oracle_mod, pts, mp, thr, merge_impl=0, stage=0, **vccs):
    o = oracle_mod.Oracle()
    o.set_vccs_params(**vccs)
    o.set_merge_params(merge_impl=merge_impl, **mp)
    o.set_input(pts)
    o.run(stage, thr)
    return o


def test_golden_fixture_regression(oracle_mod):
    g, pts = golden_points()
    o = run_oracle(oracle_mod, pts, AL, 0.2)
    for n in ("keys", "voxel_count", "nbr_count", "seeds", "labels", "sv_label", "sv_count"):
        assert np.array_equal(o.array(n), g[n]), n
    for n in ("voxel_xyz", "voxel_rgb", "normals", "nbr", "dist"):
        d = np.frombuffer(digest(o.array(n)), np.uint8)
        assert np.array_equal(d, g[n + "_sha256"]), n
    assert np.array_equal(o.array("edges_ab"), g["al_edges_ab"])
    assert np.array_equal(bits(o.array("edges_w")), g["al_edges_w_bits"])
    assert np.array_equal(o.array("merges_ab"), g["al_merges_ab"])
    assert np.array_equal(bits(o.array("merges_w")), g["al_merges_w_bits"])
    assert np.array_equal(o.array("out_label"), g["al_out_label"])


@pytest.mark.parametrize("mp,thr", [(AL, 0.2), (EQ, 0.6), (dict(color_mode=1, geom_mode=0, merge_mode=0, lam=0.3), 0.15)])
def test_stamp_merge_equals_literal_multimap(oracle_mod, mp, thr):
    """stamp tie keys + prefix-continued statistics replay the literal multimap loop."""
    g, pts = golden_points()
    lit = run_oracle(oracle_mod, pts, mp, thr, merge_impl=0)
    fast = run_oracle(oracle_mod, pts, mp, thr, merge_impl=1)
    assert len(lit.array("merges_w")) > 100
    for n in ("merges_ab", "merges_left", "final_ab", "out_label", "out_voxel"):
        assert np.array_equal(lit.array(n), fast.array(n)), n
    assert np.array_equal(bits(lit.array("merges_w")), bits(fast.array("merges_w")))
    assert np.array_equal(bits(lit.array("final_w")), bits(fast.array("final_w")))


def test_stamp_merge_random_tied_graphs(oracle_mod):
    """Heavy exact ties (EQ with few bins) on random graphs injected through set_graph."""
    rng = np.random.default_rng(5)
    for trial in range(25):
        S = int(rng.integers(5, 40))
        sizes = rng.integers(3, 12, S)
        V = int(sizes.sum())
        vxyz = rng.normal(0, 1, (V, 3)).astype(np.float32) + np.float32(3)
        vrgba = (rng.integers(0, 4, V).astype(np.uint32) * 60) * np.uint32(0x010101)
        labels = np.sort(rng.choice(np.arange(1, 200), S, replace=False)).astype(np.uint32)
        off = np.concatenate([[0], np.cumsum(sizes)])
        lists = [np.arange(off[i], off[i + 1]) for i in range(S)]
        cen = np.stack([vxyz[l].mean(0) for l in lists]).astype(np.float32)
        nrm = rng.normal(0, 1, (S, 3)).astype(np.float32)
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        pairs = set()
        for i in range(S - 1):
            pairs.add((i, i + 1))
        for _ in range(2 * S):
            i, j = sorted(rng.choice(S, 2, replace=False))
            pairs.add((int(i), int(j)))
        adj = []
        for i, j in sorted(pairs):
            adj += [(labels[i], labels[j]), (labels[j], labels[i])]
        adj = np.array(sorted(adj), np.uint32)
        res = []
        for impl in (0, 1):
            o = oracle_mod.Oracle()
            o.set_merge_params(color_mode=1, geom_mode=trial % 2, merge_mode=2, bins=int(rng.integers(2, 9)) if impl == 0 else res[0][2], merge_impl=impl)
            if impl == 0:
                b = int(rng.integers(2, 9))
                o.set_merge_params(color_mode=1, geom_mode=trial % 2, merge_mode=2, bins=b, merge_impl=0)
            o.set_graph(vxyz, vrgba, labels, lists, cen, nrm, adj)
            o.run(7, 0.9)
            res.append((o.array("merges_ab"), bits(o.array("merges_w")), b))
        assert np.array_equal(res[0][0], res[1][0]), trial
        assert np.array_equal(res[0][1], res[1][1]), trial


def test_structural_invariants(oracle_mod):
    g, pts = golden_points()
    o = run_oracle(oracle_mod, pts, AL, 0.2, stage=8)
    keys = o.array("keys"); morton = o.array("morton"); s = o.scalars()
    depth = int(s["depth"])
    assert keys.max() < (1 << depth)
    assert np.all(np.diff(morton.astype(np.int64)) > 0)                  # DFS leaf order = ascending x-major Morton
    assert int(o.array("voxel_count").sum()) == int(np.isfinite(g["xyz"]).all(axis=1).sum())
    nbr, cnt = o.array("nbr"), o.array("nbr_count")
    V = len(cnt)
    for v in range(0, V, 97):                                            # symmetry + self inclusion
        row = nbr[v, :cnt[v]]
        assert v in row
        for u in row:
            assert v in nbr[u, :cnt[u]]
    nrm = o.array("normals")
    ok = np.isfinite(nrm).all(axis=1)
    assert np.allclose(np.linalg.norm(nrm[ok, :3], axis=1), 1.0, atol=1e-5)
    labels = o.array("labels")
    assert labels.max() <= len(o.array("seeds"))
    ab = o.array("edges_ab")
    assert np.all(ab[:, 0] < ab[:, 1])
    assert np.all(np.lexsort((ab[:, 1], ab[:, 0])) == np.arange(len(ab)))  # lexicographic edge order


def test_switches_change_results(oracle_mod):
    """Every version-dependent PCL behaviour sits behind a named switch that really is wired."""
    g, pts = golden_points()
    base = run_oracle(oracle_mod, pts, AL, 0.2, stage=8)
    o = oracle_mod.Oracle(); o.set_vccs_params(); o.set_switches(init_seed_voxel=1); o.set_input(pts); o.run(8)
    assert not np.array_equal(o.array("labels"), base.array("labels"))
    o = oracle_mod.Oracle(); o.set_vccs_params(); o.set_switches(leaf_desc=1); o.set_input(pts); o.run(1)
    assert np.array_equal(o.array("keys")[::-1], base.array("keys"))


def test_no_transform_and_negative_z(oracle_mod):
    g, pts = golden_points()
    p2 = pts.copy(); p2["z"] = -p2["z"]                                  # the bundled cloud has z < 0 everywhere
    a = run_oracle(oracle_mod, pts, AL, 0.2, stage=1)
    b = run_oracle(oracle_mod, p2, AL, 0.2, stage=1)
    assert np.array_equal(a.array("keys"), b.array("keys"))              # main() folds z<0 to |z| (:317-321)
    c = run_oracle(oracle_mod, pts, AL, 0.2, stage=1, use_transform=False)
    assert len(c.array("keys")) != len(a.array("keys"))


def test_empty_and_nan_inputs(oracle_mod):
    from f3ps import synth
    for pts in (np.zeros(0, synth.POINT_DTYPE), synth.pack_points(np.full((50, 3), np.nan, np.float32), np.zeros(50, np.uint32))):
        o = run_oracle(oracle_mod, pts, AL, 0.2)
        assert len(o.array("keys")) == 0 and len(o.array("merges_w")) == 0 and len(o.array("out_label")) == 0


def test_float_libm_model_matches_glibc_mostly(oracle_mod):
    """The oracle defines logf as the correctly rounded value; glibc's own logf may differ in rare last-bit cases."""
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.logf.restype = ctypes.c_float
    libm.logf.argtypes = [ctypes.c_float]
    rng = np.random.default_rng(0)
    x = rng.uniform(0.3, 5.0, 20000).astype(np.float32)
    L = oracle_mod.lib()
    ours = np.array([L.orc_cr_logf(float(v)) for v in x], np.float32)
    glibc = np.array([libm.logf(float(v)) for v in x], np.float32)
    assert np.mean(ours != glibc) < 0.01
    ulp = np.spacing(np.maximum(np.abs(ours), np.float32(1e-30))).astype(np.float64)
    assert np.max(np.abs(ours.astype(np.float64) - glibc.astype(np.float64)) / ulp) <= 1.0


def test_bundled_frame_config1(oracle_mod):
    """C1: the reference's sample cloud (tests/fixtures/, a copy of /root/reference/pcd/) at defaults; sizes recorded by
    the survey probes."""
    from f3ps import pcd
    pts, label, hdr = pcd.read_pcd(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures", "milk_cartoon_all_small_clorox.pcd"))
    assert len(pts) == 307200 and int(np.isfinite(pts["z"]).sum()) == 241407
    o = run_oracle(oracle_mod, pts, dict(color_mode=0, geom_mode=1, merge_mode=1), 0.2, merge_impl=1)
    assert len(o.array("keys")) == 34211 and int(o.scalars()["depth"]) == 8
    assert abs(o.array("nbr_count").mean() - 12.87) < 0.01
    assert len(o.array("seeds")) == 604
    assert list(o.array("steals")[:4]) == [4, 1989, 3749, 3377]
    assert len(o.array("sv_label")) == 597 and len(o.array("edges_w")) == 1556
    assert abs(o.scalars()["lambda"] - 0.637) < 0.01
    assert 560 <= len(o.array("merges_w")) <= 600


@pytest.mark.parametrize("seed,w,h", [(11, 160, 120), (16, 160, 120), (52, 160, 120), (127, 160, 120), (30000, 320, 240), (30005, 320, 240)])
def test_expansion_fixed_point_equals_literal_sequential(oracle_mod, seed, w, h):
    """the per-voxel fold + steal-table fixed point (what the CUDA kernel runs) reproduces PCL's
    sequential helper loop exactly, including 'phantom' seed leaves (two seed cells electing one voxel; seeds 16, 52
    and 127 keep such a leaf to the end, so it shows up in voxels_, the adjacency and the labelled cloud)."""
    from f3ps import synth
    pts = synth.make_frame(seed=seed, width=w, height=h)
    res = []
    for impl in (0, 1):
        o = oracle_mod.Oracle(); o.set_vccs_params(); o.set_merge_params(merge_impl=1, **AL); o.set_expand_impl(impl)
        o.set_input(pts); o.run(0, 0.2)
        res.append(o)
    lit, fp = res
    for n in ("labels", "sv_label", "sv_count", "adj", "edges_ab", "merges_ab", "out_label", "out_voxel"):
        assert np.array_equal(lit.array(n), fp.array(n)), n
    for n in ("dist", "sv_xyz", "sv_rgb", "sv_normal", "edges_w", "merges_w"):
        assert np.array_equal(bits(lit.array(n)), bits(fp.array(n))), n
    if seed in (16, 52, 127):
        assert int(lit.array("sv_count").sum()) > int((lit.array("labels") > 0).sum())      # a phantom leaf survived


def test_refine_supervoxels_oracle_invariants(oracle_mod):
    """Oracle::refine (pcl refineSupervoxels restated): labels stay a partition into the surviving helpers, every helper's
    leaf list is its label's voxels (+ at most one phantom leaf), the helper count never grows, zero iterations change nothing."""
    from f3ps import synth
    pts = synth.make_frame(seed=11, width=160, height=120)
    o = oracle_mod.Oracle(); o.set_vccs_params(); o.set_merge_params(merge_impl=1, **AL); o.set_input(pts)
    for st in (1, 2, 3, 4, 5):
        o.run(st)
    l0 = o.array("labels").copy(); s0 = o.array("sv_label").copy(); n0 = o.array("normals").copy()
    o.refine(0)
    assert np.array_equal(o.array("labels"), l0) and np.array_equal(o.array("sv_label"), s0)
    o.refine(2)
    l1 = o.array("labels"); s1 = o.array("sv_label"); c1 = o.array("sv_count")
    assert set(s1.tolist()) <= set(s0.tolist()) and len(s1) <= len(s0)
    assert set(np.unique(l1[l1 > 0]).tolist()) == set(s1.tolist())
    owned = np.bincount(l1, minlength=int(s0.max()) + 1)[s1]
    assert np.all((c1 == owned) | (c1 == owned + 1))
    n1 = o.array("normals")
    assert not np.array_equal(n0, n1, equal_nan=True)
    nn = np.linalg.norm(n1[:, :3][np.isfinite(n1[:, 0])], axis=1)
    assert np.allclose(nn, 1.0, atol=1e-4)
    o.run(6); o.run(7, 0.2)
    assert len(o.array("merges_ab")) > 0
