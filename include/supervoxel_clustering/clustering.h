/* clustering.h -- the Clustering class with the reference's public surface
 * (reference: include/supervoxel_clustering/clustering.h:51-212, src/clustering.cpp), executing on the
 * GPU through the f3ps C ABI.  set_initialstate uploads the supervoxel graph once (f3ps_set_graph);
 * cluster(t) replays the reference's serial min-edge merge order in the persistent merge kernel and
 * mirrors the resulting state back into `state`.
 *
 * Error behaviour (src/clustering.cpp:574-597, 670-673): same exception types, same messages. */
#ifndef F3PS_CLUSTERING_H_
#define F3PS_CLUSTERING_H_

#include <map>
#include <memory>
#include <set>
#include <vector>

#include "pcl_shim.h"
#include "color_utilities.h"
#include "clustering_state.h"
#include "testing.h"

typedef pcl::Normal Normal;
typedef pcl::PointXYZL PointLT;
typedef pcl::PointXYZRGBL PointLCT;
typedef pcl::PointCloud<PointT> PointCloudT;
typedef pcl::PointCloud<PointLT> PointLCloudT;
typedef std::multimap<uint32_t, uint32_t> AdjacencyMapT;
typedef std::multiset<float> DeltasDistribT;

enum ColorDistance { LAB_CIEDE00, RGB_EUCL };
enum GeometricDistance { NORMALS_DIFF, CONVEX_NORMALS_DIFF };
enum MergingCriterion { MANUAL_LAMBDA, ADAPTIVE_LAMBDA, EQUALIZATION };

struct MergeStep { uint32_t a, b; float w; uint32_t edges_left, regions_left; };

class Clustering {
    ColorDistance delta_c_type;
    GeometricDistance delta_g_type;
    MergingCriterion merging_type;
    float lambda;
    short bins_num;
    bool set_initial_state, init_initial_weights;
    ClusteringState initial_state, state;
    std::shared_ptr<f3ps::Handle> h_;
    std::vector<PointT> flat_voxels_;          /* voxels of every initial supervoxel, label order */
    std::vector<Normal> flat_normals_;
    std::vector<MergeStep> merge_log_;

    void push_params();
    void pull_state(bool merged);

public:
    Clustering();
    Clustering(ColorDistance c, GeometricDistance g, MergingCriterion m);

    void set_delta_c(ColorDistance d) { delta_c_type = d; }
    void set_delta_g(GeometricDistance d) { delta_g_type = d; }
    void set_merging(MergingCriterion m);
    void set_lambda(float l);
    void set_bins_num(short b);
    void set_initialstate(ClusteringT segm, AdjacencyMapT adj);

    ColorDistance get_delta_c() const { return delta_c_type; }
    GeometricDistance get_delta_g() const { return delta_g_type; }
    MergingCriterion get_merging() const { return merging_type; }
    float get_lambda() const { return lambda; }
    short get_bins_num() const { return bins_num; }

    std::pair<ClusteringT, AdjacencyMapT> get_currentstate() const;
    PointCloudT::Ptr get_colored_cloud() const;
    PointLCloudT::Ptr get_labeled_cloud() const;

    void cluster(float threshold);

    /* threshold sweep of src/clustering.cpp:691-774: every threshold is a prefix of ONE merge replay on the device;
     * ground_truth = labelled voxel cloud (points matched to the segmentation's voxels by exact xyz, as Testing does) */
    std::map<float, performanceSet> all_thresh(PointLCloudT::Ptr ground_truth, float start_thresh, float end_thresh, float step_thresh);
    std::pair<float, performanceSet> best_thresh(PointLCloudT::Ptr ground_truth, float start_thresh, float end_thresh, float step_thresh);
    std::pair<float, performanceSet> best_thresh(std::map<float, performanceSet> all_thresh);

    /* the reference's --V trace "left: %de/%dp - w: %f - [%d, %d]" (src/clustering.cpp:390-392), as data */
    const std::vector<MergeStep>& get_merge_log() const { return merge_log_; }

    /* ColorUtilities' print-only self checks (src/clustering.cpp:779-783) */
    void test_all() const;

    static PointCloudT::Ptr label2color(PointLCloudT::Ptr label_cloud);
    static PointLCloudT::Ptr color2label(PointCloudT::Ptr colored_cloud);
};

#endif
