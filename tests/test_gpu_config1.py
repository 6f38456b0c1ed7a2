"""BASELINE config 1 on the GPU: the reference's bundled cloud (tests/fixtures/milk_cartoon_all_small_clorox.pcd, a copy of
/root/reference/pcd/; loaded as src/supervoxel_clustering.cpp:313-342 does, z<0 folded) through the CUDA path, every
stage array and the full merge replay against the oracle; the launch file's flags (--CVX --AL -t 0.2,
launch/supervoxel_clustering.launch:4), the CLI defaults, and the CLI's automatic threshold (:428-443)."""
import os
import subprocess

import numpy as np
import pytest

from test_gpu_parity import AL, MERGE_ARRAYS, STAGE_ARRAYS, assert_parity, run_both

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "fast-3d-pointcloud-segmentation_b200", "host")
FIXTURE = os.path.join(ROOT, "tests", "fixtures", "milk_cartoon_all_small_clorox.pcd")
DEFAULTS = dict(color_mode=0, geom_mode=0, merge_mode=1)            # L*a*b*, plain normals difference, adaptive lambda


@pytest.fixture(scope="module")
def gpu():
    import f3ps
    f3ps.build()
    return f3ps


@pytest.fixture(scope="module")
def bundled():
    from f3ps import pcd
    pts, label, hdr = pcd.read_pcd(FIXTURE)
    assert label is None and len(pts) == 307200
    return pts


@pytest.mark.parametrize("mp,thr", [(AL, 0.2), (DEFAULTS, 0.2), (AL, 1.0)], ids=["launch_file_flags", "cli_defaults", "full_replay"])
def test_config1_all_stages_and_merge_replay(gpu, oracle_mod, bundled, mp, thr):
    g, o = run_both(gpu, oracle_mod, bundled, mp, thr, merge_impl=1, fold_negative_z=True)
    assert_parity(g, o, STAGE_ARRAYS + MERGE_ARRAYS)
    c = g.counts()
    # the survey's independent probe of this file
    assert c.n_points == 307200 and c.n_voxels == 34211 and len(g.array("seeds")) == 604
    assert c.n_supervoxels == 597 and c.n_edges == 1556
    assert c.n_merges == len(o.array("merges_w")) and (thr < 1.0 or c.n_merges >= 590)
    assert np.mean(g.array("labels") == o.array("labels")) == 1.0
    assert c.merge_path == 1                                    # the resident kernel


def test_config1_cli_launch_file_and_auto_threshold(gpu, oracle_mod, bundled, tmp_path):
    """`supervoxel_clustering -p <bundled> --CVX --AL -t 0.2 --V` prints the reference's per-merge debug line
    (src/clustering.cpp:390-392) for exactly the oracle's sequence; without -t the 41-threshold sweep picks its threshold
    against the single ground-truth segment (the file has no label field) and re-clusters there."""
    subprocess.check_call(["make", "-C", HOST, "-s"])
    cli = os.path.join(HOST, "supervoxel_clustering")
    run = subprocess.run([cli, "-p", FIXTURE, "--CVX", "--AL", "-t", "0.2", "--V"], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout + run.stderr
    o = oracle_mod.Oracle(); o.set_vccs_params(fold_negative_z=True); o.set_merge_params(merge_impl=1, **AL)
    o.set_input(bundled); o.run(0, 0.2)
    trace = [l for l in run.stdout.splitlines() if l.startswith("left: ")]
    ab, left = o.array("merges_ab"), o.array("merges_left")
    assert len(trace) == len(ab) > 500
    for line, (a, b), (el, rl) in zip(trace, ab, left):
        assert line.startswith("left: %de/%dp" % (el, rl)) and line.endswith("[%d, %d]...OK" % (a, b))
    auto = subprocess.run([cli, "-p", FIXTURE], capture_output=True, text=True, timeout=600)
    assert auto.returncode == 0, auto.stdout + auto.stderr
    lines = auto.stdout.splitlines()
    sweep = [l for l in lines if l.startswith("<T, Fscore, voi, wov>")]
    assert len(sweep) == 41 and sweep[0].startswith("<T, Fscore, voi, wov> = <0.800000,")
    g = gpu.Segmenter(); g.set_vccs_params(); g.set_merge_params(**DEFAULTS); g.set_input(bundled); g.extract(); g.graph()
    bt, bp = g.best_thresh(np.zeros(g.counts().n_voxels, np.uint32))
    best = [l for l in lines if l.startswith("Using best threshold:")]
    assert len(best) == 1 and best[0].startswith("Using best threshold: %f (F-score %f" % (bt, bp["fscore"]))
    assert abs(bt - 0.8) < 1e-6                                 # tools/c1_oracle_run.py: one truth segment -> the first threshold wins
