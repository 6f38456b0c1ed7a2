"""Test-infrastructure restatement of the reference's evaluation module, used ONLY by tests/ as the checker of
f3ps_eval_thresholds.  Literal: Testing::label_map / compute_intersections / eval_* (/root/reference/src/testing.cpp:62-146,
175-219, 239-362) on two labelled point clouds, with the intersections by exact xyz (count_intersect's sort +
set_intersection), float32 arithmetic in the reference's order, and Clustering::all_thresh / best_thresh
(/root/reference/src/clustering.cpp:691-774) driven by re-running the oracle's cluster() per threshold.
Pinned by the reference's OWN Testing class: /root/reference/src/testing.cpp compiles where it lies against the container stand-ins
of oracle/ref_shim/ (oracle/Makefile `ref`); its scores on 38 labelled cloud pairs are committed as tests/golden/testing_ref.json
(tools/gen_testing_golden.py) and tests/test_oracle_testing.py compares this module with them (and with the compiled class live where
the reference tree exists), next to hand-computed cases."""
import numpy as np

F = np.float32


def label_map(labels):
    """dense renumbering in ascending label order -> (dense label per point, list of original labels)"""
    uniq = np.unique(labels)
    return np.searchsorted(uniq, labels), uniq


def _xyz_keys(xyz):
    a = np.ascontiguousarray(xyz, np.float32)
    a = np.where(a == 0, np.float32(0), a)                     # -0 == +0 under compareXYZ
    return [tuple(r) for r in a.view(np.uint32).reshape(-1, 3)]


def scores(seg_xyz, seg_label, truth_xyz, truth_label):
    """performanceSet of Testing(segm, truth).eval_performance()"""
    sd, su = label_map(np.asarray(seg_label))
    td, tu = label_map(np.asarray(truth_label))
    n_seg, n_truth = len(su), len(tu)
    # point sets per segment (std::set semantics after the sort: duplicates inside one set stay duplicates; none occur)
    skeys, tkeys = _xyz_keys(seg_xyz), _xyz_keys(truth_xyz)
    tsets = [dict() for _ in range(n_truth)]
    for k, j in zip(tkeys, td):
        tsets[j][k] = tsets[j].get(k, 0) + 1
    inter = np.zeros((n_seg, n_truth), np.int64)
    ssz = np.bincount(sd, minlength=n_seg).astype(np.int64)
    g = np.bincount(td, minlength=n_truth).astype(np.int64)
    for k, i in zip(skeys, sd):
        for j in range(n_truth):
            if k in tsets[j]:
                inter[i, j] += 1
    t_sizes = {}
    for j in range(n_truth):                                    # std::map::insert keeps the first truth segment of a size
        t_sizes.setdefault(int(g[j]), j)
    matches = -np.ones(n_truth, np.int64)
    for size in sorted(t_sizes, reverse=True):
        j = t_sizes[size]
        col = inter[:, j].copy()
        row = int(np.argmax(col))                               # first maximum, as Eigen's maxCoeff
        while (matches == row).any():
            col[row] = 0
            if col.any():
                row = int(np.argmax(col))
            else:
                row = -1
                break
        matches[j] = row
    n = F(len(truth_label))
    h_s = F(0); h_t = F(0); mi = F(0)
    for i in range(n_seg):
        p = F(ssz[i])
        h_s = F(h_s - F(F(np.log(F(p / n))) * p) / n)
        for j in range(n_truth):
            q = F(g[j])
            if i == 0:
                h_t = F(h_t - F(F(np.log(F(q / n))) * q) / n)
            r = F(inter[i, j])
            if r != 0:
                mi = F(mi + F(F(np.log(F(F(n * r) / F(p * q)))) * r) / n)
    voi = F(F(h_s + h_t) - F(F(2) * mi))
    p_ = F(0); r_ = F(0); fp = F(0); fn = F(0)
    for j in range(n_truth):
        i = matches[j]
        if i != -1:
            it = F(inter[i, j]); s = F(ssz[i]); gg = F(g[j])
            p_ = F(p_ + F(F(it * gg) / s)); r_ = F(r_ + it); fp = F(fp + F(s - it)); fn = F(fn + F(gg - it))
        else:
            fn = F(fn + F(g[j]))
    precision, recall, fpr, fnr = F(p_ / n), F(r_ / n), F(fp / n), F(fn / n)
    fscore = F(0) if (precision == 0 and recall == 0) else F(F(F(2) * F(precision * recall)) / F(precision + recall))
    w = F(0)
    for j in range(n_truth):
        i = matches[j]
        if i != -1:
            it = F(inter[i, j]); un = F(ssz[i] + g[j] - inter[i, j])
            w = F(w + F(F(it * F(g[j])) / un))
    wov = F(w / n)
    return dict(voi=float(voi), precision=float(precision), recall=float(recall), fscore=float(fscore), wov=float(wov),
                fpr=float(fpr), fnr=float(fnr))


def all_thresh(make_oracle, truth_xyz, truth_label, start=0.8, end=1.0, step=0.005):
    """Clustering::all_thresh: make_oracle() returns an oracle with input and parameters set; cluster(t) per threshold
    (continuing from the current state equals restarting)."""
    start, end, step = F(start), F(end), F(step)
    thr = [start]
    t = F(start + step)
    while t <= end:
        thr.append(t)
        t = F(t + step)
    out = {}
    o = make_oracle()
    for t in thr:
        o.run(0, float(t))
        out[float(t)] = scores(o.array("out_xyz"), o.array("out_label"), truth_xyz, truth_label)
    return out


def best_thresh(res):
    best_t, best = 0.0, dict(fscore=0.0)
    for t in sorted(res):
        if res[t]["fscore"] > best["fscore"]:
            best_t, best = t, res[t]
    return best_t, best
