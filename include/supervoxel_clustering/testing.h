/* testing.h -- performanceSet of the reference's evaluation module (include/supervoxel_clustering/testing.h:68-75).
 * The scores themselves (Testing::eval_performance, src/testing.cpp:239-362) are computed by f3ps_eval_thresholds from
 * one merge replay on the device; Clustering::all_thresh / best_thresh expose them with the reference's signatures. */
#ifndef F3PS_TESTING_H_
#define F3PS_TESTING_H_

struct performanceSet {
    performanceSet() : voi(0), precision(0), recall(0), fscore(0), wov(0), fpr(0), fnr(0) {}
    float voi, precision, recall, fscore, wov, fpr, fnr;
};

#endif
