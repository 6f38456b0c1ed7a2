/* clustering_state.h -- ClusteringState with the reference's public surface
 * (reference: include/supervoxel_clustering/clustering_state.h:47-123, src/clustering_state.cpp:49-52).
 * A value type: region map + weight multimap ordered by weight, equal weights in insertion order.
 * In this build it is a host-side snapshot of the device state kept behind f3ps_ctx. */
#ifndef F3PS_CLUSTERINGSTATE_H_
#define F3PS_CLUSTERINGSTATE_H_

#include <map>
#include <utility>
#include "pcl_shim.h"

typedef pcl::PointXYZRGBA PointT;
typedef pcl::Supervoxel<PointT> SupervoxelT;
typedef std::map<uint32_t, SupervoxelT::Ptr> ClusteringT;
typedef std::multimap<float, std::pair<uint32_t, uint32_t> > WeightMapT;
typedef std::pair<float, std::pair<uint32_t, uint32_t> > WeightedPairT;

class ClusteringState {
    friend class Clustering;
    ClusteringT segments;
    WeightMapT weight_map;

public:
    ClusteringState() {}
    ClusteringState(ClusteringT s, WeightMapT w);

    ClusteringT get_segments() const { return segments; }
    void set_segments(ClusteringT s) { segments = s; }
    WeightMapT get_weight_map() const { return weight_map; }
    void set_weight_map(WeightMapT w) { weight_map = w; }
    /* smallest-weight edge; the caller checks for emptiness first, as in the reference */
    WeightedPairT get_first_weight() const { return *(weight_map.begin()); }
};

#endif
