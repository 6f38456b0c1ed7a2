"""The reference-shaped C++ classes (SupervoxelClustering -> Clustering) and the supervoxel_clustering CLI on
a real GPU: same merge sequence as the fused C-ABI path, the reference's exception types, CLI flags."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "fast-3d-pointcloud-segmentation_b200", "host")


@pytest.fixture(scope="module")
def pcd_file(tmp_path_factory, small_frame):
    import f3ps
    from f3ps import pcd
    f3ps.build()
    subprocess.check_call(["make", "-C", HOST, "-s"])
    p = str(tmp_path_factory.mktemp("pcd") / "small.pcd")
    pcd.write_pcd_binary(p, small_frame)
    return p


def test_facade_matches_fused_path(pcd_file):
    out = subprocess.run([os.path.join(HOST, "facade_selftest"), pcd_file, "0.2"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "FACADE OK" in out.stdout


def test_cli_flags_and_merge_trace(pcd_file, oracle_mod, small_frame, tmp_path):
    cli = os.path.join(HOST, "supervoxel_clustering")
    assert subprocess.run([cli], capture_output=True).returncode == 1                        # argc < 3 -> usage, exit 1
    assert subprocess.run([cli, "-p", pcd_file, "--facade"], capture_output=True).returncode == 1   # the sweep runs on the direct path
    bad = subprocess.run([cli, "-p", pcd_file, "-t", "0.2", "--ML", "--AL"], capture_output=True, text=True)
    assert bad.returncode == 1 and "Only one parameter" in bad.stderr
    outp = str(tmp_path / "labels.pcd")
    run = subprocess.run([cli, "-p", pcd_file, "-t", "0.2", "--CVX", "--AL", "--V", "-o", outp], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout + run.stderr
    o = oracle_mod.Oracle(); o.set_vccs_params(fold_negative_z=True); o.set_merge_params(color_mode=0, geom_mode=1, merge_mode=1, merge_impl=1)
    o.set_input(small_frame); o.run(0, 0.2)
    trace = [l for l in run.stdout.splitlines() if l.startswith("left: ")]
    ab, left = o.array("merges_ab"), o.array("merges_left")
    assert len(trace) == len(ab)
    for line, (a, b), (el, rl) in zip(trace, ab, left):
        assert line.startswith("left: %de/%dp" % (el, rl)) and line.endswith("[%d, %d]...OK" % (a, b))
    rows = [l.split() for l in open(outp).read().splitlines()[11:]]
    assert len(rows) == len(o.array("out_label"))
    assert np.array_equal(np.array([int(r[3]) for r in rows], np.uint32), o.array("out_label"))
    # --facade goes through the classes and must print the same trace
    run2 = subprocess.run([cli, "-p", pcd_file, "-t", "0.2", "--CVX", "--AL", "--V", "--facade"], capture_output=True, text=True, timeout=300)
    assert run2.returncode == 0, run2.stdout + run2.stderr
    assert [l for l in run2.stdout.splitlines() if l.startswith("left: ")] == trace


def test_cli_directory_sweep_frames_in_flight(tmp_path, oracle_mod):
    """-d sweep: several frames in flight per GPU (one handle / stream / host thread each); reports come back in file
    order and every file merges exactly as it does alone."""
    import f3ps
    from f3ps import pcd, synth
    subprocess.check_call(["make", "-C", HOST, "-s"])
    d = tmp_path / "sweep"; d.mkdir()
    want = {}
    for i in range(7):
        pts = synth.make_frame(seed=300 + i, width=160, height=120)
        path = str(d / ("f%02d.pcd" % i))
        pcd.write_pcd_binary(path, pts)
        o = oracle_mod.Oracle(); o.set_vccs_params(fold_negative_z=True); o.set_merge_params(color_mode=0, geom_mode=0, merge_mode=2, bins=200, merge_impl=1)
        o.set_input(pts); o.run(0, 0.3)
        want[path] = (len(o.array("merges_ab")), len(o.array("out_label")))
    cli = os.path.join(HOST, "supervoxel_clustering")
    run = subprocess.run([cli, "-d", str(d), "-t", "0.3", "--EQ", "200", "--inflight", "3"], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout + run.stderr
    lines = run.stdout.splitlines()
    assert lines[0] == "Found 7 files"
    files = [l.split("'")[1] for l in lines if l.startswith("Loading pointcloud")]
    done = [l for l in lines if l.startswith("Clustering complete")]
    assert sorted(files) == sorted(want) and len(done) == 7
    for f, l in zip(files, done):
        tok = l.split()
        assert int(tok[5]) == want[f][0] and int(tok[11]) == want[f][1], (f, l, want[f])


def test_cli_auto_threshold_like_the_reference_default(pcd_file, small_frame):
    """No -t: main() sweeps 41 thresholds 0.8 .. 1 (src/supervoxel_clustering.cpp:428-438) against the ground truth and
    re-clusters at the best F-score.  The file has no label field -> one truth segment, as the reference's bundled cloud."""
    import f3ps
    cli = os.path.join(HOST, "supervoxel_clustering")
    run = subprocess.run([cli, "-p", pcd_file, "--CVX", "--AL"], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout + run.stderr
    lines = run.stdout.splitlines()
    sweep = [l for l in lines if l.startswith("<T, Fscore, voi, wov>")]
    assert len(sweep) == 41 and sweep[0].startswith("<T, Fscore, voi, wov> = <0.800000,")
    best = [l for l in lines if l.startswith("Using best threshold:")]
    assert len(best) == 1
    g = f3ps.Segmenter(); g.set_vccs_params(); g.set_merge_params(color_mode=0, geom_mode=1, merge_mode=1)
    g.set_input(small_frame); g.extract(); g.graph()
    bt, bp = g.best_thresh(np.zeros(g.counts().n_voxels, np.uint32))
    assert best[0].startswith("Using best threshold: %f (F-score %f" % (bt, bp["fscore"]))
    g.merge(bt)
    done = [l for l in lines if l.startswith("Clustering complete")][0].split()
    assert int(done[5]) == g.counts().n_merges and int(done[8]) == g.counts().n_segments
