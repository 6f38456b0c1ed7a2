// ref_testing_wrap.cpp -- plain-C entry to the REFERENCE's own Testing class (/root/reference/src/testing.cpp, compiled where it lies
// against the container stand-ins of oracle/ref_shim/): Testing(segm, truth).eval_performance()
// (/root/reference/include/supervoxel_clustering/testing.h:77-134).  Test infrastructure: built into oracle/_ref/libref_clustering.so by
// oracle/Makefile when /root/reference exists; only tests/ and tools/gen_testing_golden.py load it.
#include <cstdint>
#include <exception>
#include "supervoxel_clustering/testing.h"

static PointLCloudT::Ptr make_cloud(const float* xyz, const uint32_t* label, int64_t n) {
    PointLCloudT::Ptr c = boost::make_shared<PointLCloudT>();
    for (int64_t i = 0; i < n; ++i) { PointLT p; p.x = xyz[3 * i]; p.y = xyz[3 * i + 1]; p.z = xyz[3 * i + 2]; p.label = label[i]; c->push_back(p); }
    return c;
}
// out[7] = voi, precision, recall, fscore, wov, fpr, fnr; returns 0, or 1 when the reference throws (an empty cloud)
extern "C" int ref_testing_eval(const float* seg_xyz, const uint32_t* seg_label, int64_t n_seg, const float* truth_xyz, const uint32_t* truth_label,
                                int64_t n_truth, float* out) {
    try {
        Testing t(make_cloud(seg_xyz, seg_label, n_seg), make_cloud(truth_xyz, truth_label, n_truth));
        const performanceSet p = t.eval_performance();
        out[0] = p.voi; out[1] = p.precision; out[2] = p.recall; out[3] = p.fscore; out[4] = p.wov; out[5] = p.fpr; out[6] = p.fnr;
        return 0;
    } catch (const std::exception&) { return 1; }
}
