"""Auto-threshold sweep on the device (f3ps_eval_thresholds = Clustering::all_thresh + Testing::eval_performance,
src/clustering.cpp:691-774, src/testing.cpp:239-362) against the literal restatement in oracle/oracle_testing.py, which
re-clusters per threshold and intersects point sets by xyz like the reference does.  Scores are float32 sums of logs and
ratios: tolerance 1e-5 absolute (libm logf vs numpy log); segment counts, prefix lengths and the chosen threshold exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

AL = dict(color_mode=0, geom_mode=1, merge_mode=1)


@pytest.fixture(scope="module")
def gpu():
    import f3ps
    return f3ps


def _truth_from(o_truth, V):
    """a ground truth with a handful of segments: a coarser segmentation of the same voxels; unowned voxels get label 777"""
    t = np.full(V, 777, np.uint32)
    t[o_truth.array("out_voxel")] = o_truth.array("out_label") % 7 + 1      # fold to 7 classes: several regions per class
    return t


@pytest.mark.parametrize("sweep", [(0.8, 1.0, 0.005), (0.05, 0.6, 0.05)])
def test_all_thresh_matches_reference_restatement(gpu, oracle_mod, small_frame, sweep):
    import oracle_testing as ot

    def make_oracle():
        o = oracle_mod.Oracle(); o.set_vccs_params(); o.set_merge_params(merge_impl=1, **AL); o.set_input(small_frame)
        return o
    ot_o = make_oracle(); ot_o.run(0, 0.35)
    V = ot_o.array("voxel_xyz").shape[0]
    truth = _truth_from(ot_o, V)
    want = ot.all_thresh(make_oracle, ot_o.array("voxel_xyz"), truth, *sweep)
    g = gpu.Segmenter(); g.set_vccs_params(); g.set_merge_params(**AL); g.set_input(small_frame); g.extract(); g.graph()
    got = g.all_thresh(truth, *sweep)
    assert sorted(got) == sorted(want)
    for t in sorted(want):
        for f in gpu.Segmenter.PERF_FIELDS:
            assert abs(got[t][f] - want[t][f]) < 1e-5, (t, f, got[t][f], want[t][f])
    # prefix lengths: merges done at each threshold = the oracle's own count when clustering to that threshold
    o = make_oracle()
    for k, t in enumerate(g.last_sweep["thresholds"]):
        if k % 8 == 0:
            o.run(0, float(t))
            assert int(g.last_sweep["n_merges"][k]) == o.array("merges_ab").shape[0]
            assert int(g.last_sweep["n_segments"][k]) == len(np.unique(o.array("out_label")))
    bt, bp = g.best_thresh(truth, *sweep)
    wt, wp = ot.best_thresh(want)
    assert bt == wt and abs(bp["fscore"] - wp["fscore"]) < 1e-5
    # the handle is left clustered at the last threshold (main() then re-clusters at the best one, :443)
    assert g.counts().n_merges == int(g.last_sweep["n_merges"][-1])


def test_single_truth_segment_like_the_bundled_file(gpu, small_frame):
    """A cloud without a label field has ONE truth segment (all labels 0): recall = share of the largest segment."""
    g = gpu.Segmenter(); g.set_vccs_params(); g.set_merge_params(**AL); g.set_input(small_frame); g.extract(); g.graph()
    V = g.counts().n_voxels
    res = g.all_thresh(np.zeros(V, np.uint32), 0.8, 1.0, 0.005)
    assert len(res) == 41
    sizes = np.bincount(g.array("out_label"))
    last = res[max(res)]
    assert abs(last["recall"] - sizes.max() / V) < 1e-6 and abs(last["fpr"]) < 1e-6


def test_eval_errors(gpu, small_frame):
    g = gpu.Segmenter(); g.set_vccs_params(); g.set_merge_params(**AL); g.set_input(small_frame)
    with pytest.raises(gpu.LogicError):
        g.all_thresh(np.zeros(10, np.uint32))                   # no initial state yet (src/clustering.cpp:671-673)
    g.extract(); g.graph()
    with pytest.raises(ValueError):
        g.all_thresh(np.zeros(10, np.uint32))                   # one label per voxel
    with pytest.raises(IndexError):
        g.all_thresh(np.zeros(g.counts().n_voxels, np.uint32), 0.8, 1.5, 0.005)     # std::out_of_range (:694-698)


def test_eval_label_pairs_equals_reference_testing_class_golden(gpu):
    """f3ps_eval_label_pairs (what host/facade.cpp's Testing calls) against the scores of the REFERENCE's own Testing class on the
    committed cloud pairs (tests/golden/testing_ref.json <- /root/reference/src/testing.cpp compiled against oracle/ref_shim,
    tools/gen_testing_golden.py): float32 sums of ratios and logs, 1e-5 absolute (device logf vs libm)."""
    import json
    import os
    d = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "testing_ref.json")))
    g = gpu.Segmenter()
    for k, c in enumerate(d["cases"]):
        want = [float.fromhex(h) for h in c["scores_f32_hex"]]
        got = g.eval_clouds(np.array(c["seg_xyz"], np.float32), c["seg_label"], np.array(c["truth_xyz"], np.float32), c["truth_label"])
        for n, w in zip(d["order"], want):
            assert abs(got[n] - w) < 1e-5, (k, n, got[n], w)


@pytest.mark.parametrize("name", ["sweep_hub150_rgb_cvx_ml", "sweep_hub400_lab_cvx_al", "sweep_frame_cvx_al", "sweep_c1_defaults"])
def test_all_thresh_equals_reference_clustering_golden(gpu, name):
    """f3ps_eval_thresholds (one merge replay on the device) against Clustering::all_thresh / best_thresh of the REFERENCE's own classes
    (compiled from /root/reference/src against oracle/ref_shim, tools/gen_clustering_golden.py -> tests/golden/clustering_ref.npz):
    the same thresholds, every score within 1e-5, the same chosen threshold."""
    import os
    from test_oracle_reference_clustering import load_sweep
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "clustering_ref.npz"))
    graph, flags, (t0, t1, dt), truth, thr, perf, best = load_sweep(z, name)
    g = gpu.Segmenter(); g.set_merge_params(**flags)
    g.set_graph(*graph)
    got = g.all_thresh(truth[g._graph_order], t0, t1, dt)
    assert np.array_equal(np.array(sorted(got), np.float32), thr)
    for k, t in enumerate(thr):
        for j, n in enumerate(gpu.Segmenter.PERF_FIELDS):
            assert abs(got[float(t)][n] - perf[k, j]) < 1e-5, (name, float(t), n, got[float(t)][n], perf[k, j])
    bt, bp = g.best_thresh(truth[g._graph_order], t0, t1, dt)
    assert np.float32(bt) == best[0] and abs(bp["fscore"] - best[4]) < 1e-5
