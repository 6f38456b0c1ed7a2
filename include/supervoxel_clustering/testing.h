/* testing.h -- the reference's evaluation class with its public surface (include/supervoxel_clustering/testing.h:40-134,
 * src/testing.cpp): precision, recall, F-score, VoI, wOv, FPR and FNR of a labelled cloud against the ground truth.
 * Points of the two clouds are paired by exact xyz (compareXYZ, count_intersect :175-193); the label-pair contingency table
 * is then counted on the device and scored in the reference's float order (f3ps_eval_label_pairs).  The same scores for all
 * thresholds of a sweep come from Clustering::all_thresh in one merge replay. */
#ifndef F3PS_TESTING_H_
#define F3PS_TESTING_H_

#include <map>
#include <vector>

#include "pcl_shim.h"

typedef pcl::PointXYZL PointLT;
typedef pcl::PointCloud<PointLT> PointLCloudT;
typedef std::map<uint32_t, PointLCloudT::Ptr> labelMapT;

struct compareXYZ {
    bool operator()(PointLT const& p1, PointLT const& p2) const {
        if (p1.x != p2.x) return p1.x < p2.x;
        if (p1.y != p2.y) return p1.y < p2.y;
        return p1.z < p2.z;
    }
};

struct performanceSet {
    performanceSet() : voi(0), precision(0), recall(0), fscore(0), wov(0), fpr(0), fnr(0) {}
    float voi, precision, recall, fscore, wov, fpr, fnr;
};

class Testing {
    PointLCloudT::Ptr segm, truth;
    labelMapT segm_labels, truth_labels;
    float precision, recall, fscore, voi, wov, fpr, fnr;
    bool is_set_segm, is_set_truth;

    void init_performance();
    labelMapT label_map(PointLCloudT::Ptr in);
    void compute_intersections();          /* pairs the clouds and scores them on the device */

    Testing() { init_performance(); }

public:
    Testing(PointLCloudT::Ptr s, PointLCloudT::Ptr t);

    float eval_precision();
    float eval_recall();
    float eval_fscore();
    float eval_voi();
    float eval_wov();
    float eval_fpr();
    float eval_fnr();
    performanceSet eval_performance();

    PointLCloudT::Ptr get_segm() const { return segm; }
    PointLCloudT::Ptr get_truth() const { return truth; }

    void set_segm(PointLCloudT::Ptr s);
    void set_truth(PointLCloudT::Ptr t);
};

#endif
