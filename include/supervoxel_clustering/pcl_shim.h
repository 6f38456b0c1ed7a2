/* pcl_shim.h -- the few PCL types the reference's hot-path API is written against, for builds
 * without PCL (this image has none).  Layout-compatible with PCL 1.10: 32-byte, 16-byte aligned
 * points, colour bytes b,g,r,a.  With PCL installed, define F3PS_USE_REAL_PCL and the real headers
 * are used instead; pcl::SupervoxelClustering is then still replaced by f3ps::SupervoxelClustering.
 *
 * pcl::SupervoxelClustering as consumed by the reference: src/supervoxel_clustering.cpp:348-384.
 */
#ifndef F3PS_PCL_SHIM_H_
#define F3PS_PCL_SHIM_H_

#include <cstdint>
#include <cstring>
#include <map>
#include <set>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../f3ps.h"

#ifdef F3PS_USE_REAL_PCL
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/segmentation/supervoxel_clustering.h>
#else
namespace pcl {

struct alignas(16) PointXYZRGBA {
    float x = 0, y = 0, z = 0, data_w = 1.0f;
    union { struct { uint8_t b, g, r, a; }; uint32_t rgba; float rgb; };
    uint32_t pad_[3] = {0, 0, 0};
    PointXYZRGBA() : rgba(0) {}
};
struct alignas(16) PointXYZL {
    float x = 0, y = 0, z = 0, data_w = 1.0f;
    uint32_t label = 0;
    uint32_t pad_[3] = {0, 0, 0};
};
struct alignas(16) PointXYZRGBL {
    float x = 0, y = 0, z = 0, data_w = 1.0f;
    union { struct { uint8_t b, g, r, a; }; uint32_t rgba; float rgb; };
    uint32_t label = 0;
    uint32_t pad_[2] = {0, 0};
    PointXYZRGBL() : rgba(0) {}
};
struct alignas(16) Normal {
    float normal_x = 0, normal_y = 0, normal_z = 0, data_n_w = 0;
    float curvature = 0;
    uint32_t pad_[3] = {0, 0, 0};
};
struct alignas(16) PointNormal {
    float x = 0, y = 0, z = 0, data_w = 1.0f;
    float normal_x = 0, normal_y = 0, normal_z = 0, data_n_w = 0;
    float curvature = 0;
    uint32_t pad_[3] = {0, 0, 0};
};
static_assert(sizeof(PointXYZRGBA) == 32 && sizeof(PointXYZL) == 32 && sizeof(PointXYZRGBL) == 32 && sizeof(Normal) == 32,
              "PCL point layouts are 32 bytes");

template <typename PointT>
class PointCloud {
public:
    typedef std::shared_ptr<PointCloud<PointT>> Ptr;
    typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
    typedef typename std::vector<PointT>::iterator iterator;
    typedef typename std::vector<PointT>::const_iterator const_iterator;
    std::vector<PointT> points;
    uint32_t width = 0, height = 0;
    bool is_dense = true;

    size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    void clear() { points.clear(); width = height = 0; }
    void resize(size_t n) { points.resize(n); width = (uint32_t)n; height = 1; }
    void push_back(const PointT& p) { points.push_back(p); width = (uint32_t)points.size(); height = 1; }
    iterator begin() { return points.begin(); }
    iterator end() { return points.end(); }
    const_iterator begin() const { return points.begin(); }
    const_iterator end() const { return points.end(); }
    PointT& operator[](size_t i) { return points[i]; }
    const PointT& operator[](size_t i) const { return points[i]; }
    PointCloud& operator+=(const PointCloud& rhs) {        // concatenation, lhs first
        points.insert(points.end(), rhs.points.begin(), rhs.points.end());
        width = (uint32_t)points.size(); height = 1; is_dense = is_dense && rhs.is_dense;
        return *this;
    }
    PointCloud operator+(const PointCloud& rhs) const { return (PointCloud(*this) += rhs); }
};

template <typename PointT>
class Supervoxel {
public:
    typedef std::shared_ptr<Supervoxel<PointT>> Ptr;
    Supervoxel() : voxels_(new PointCloud<PointT>()), normals_(new PointCloud<Normal>()) {}
    Normal normal_;
    PointXYZRGBA centroid_;
    typename PointCloud<PointT>::Ptr voxels_;
    typename PointCloud<Normal>::Ptr normals_;
};

template <typename A, typename B>
void copyPointCloud(const PointCloud<A>& in, PointCloud<B>& out);   // specialisations in the host library

} // namespace pcl
#endif /* F3PS_USE_REAL_PCL */

namespace f3ps {

/* RAII wrapper of one f3ps handle; converts status codes back into the reference's exception types */
class Handle {
public:
    explicit Handle(int device = 0);
    ~Handle();
    Handle(const Handle&) = delete;
    Handle& operator=(const Handle&) = delete;
    f3ps_ctx* get() const { return ctx_; }
    void check(int status) const;       /* F3PS_ERR_LOGIC -> std::logic_error, INVALID_ARGUMENT -> std::invalid_argument, else runtime_error */
private:
    f3ps_ctx* ctx_ = nullptr;
};

/* what pcl::SupervoxelClustering::getSupervoxelAdjacencyList fills (there a boost::adjacency_list<setS, setS, undirectedS, uint32_t, float>) */
struct VoxelAdjacencyList {
    std::set<uint32_t> vertices;                                  /* supervoxel labels */
    std::map<std::pair<uint32_t, uint32_t>, float> edges;         /* (a < b) -> distance between the two centroids */
};

/* Drop-in for pcl::SupervoxelClustering<pcl::PointXYZRGBA> over the CUDA path (K1..K5 + supervoxel tables).
 * Same method names and semantics as the calls at src/supervoxel_clustering.cpp:348-367. */
template <typename PointT>
class SupervoxelClustering {
public:
    SupervoxelClustering(float voxel_resolution, float seed_resolution, int device = 0);
    void setUseSingleCameraTransform(bool val) { use_transform_ = val; }
    void setInputCloud(const typename pcl::PointCloud<PointT>::ConstPtr& cloud) { input_ = cloud; }
    void setColorImportance(float val) { color_importance_ = val; }
    void setSpatialImportance(float val) { spatial_importance_ = val; }
    void setNormalImportance(float val) { normal_importance_ = val; }
    void extract(std::map<uint32_t, typename pcl::Supervoxel<PointT>::Ptr>& supervoxel_clusters);
    typename pcl::PointCloud<PointT>::Ptr getVoxelCentroidCloud() const;
    pcl::PointCloud<pcl::PointXYZL>::Ptr getLabeledCloud() const;           /* input points with their supervoxel label */
    pcl::PointCloud<pcl::PointXYZL>::Ptr getLabeledVoxelCloud() const;
    void getSupervoxelAdjacency(std::multimap<uint32_t, uint32_t>& label_adjacency) const;
    /* refineSupervoxels(num_itr, clusters), src/supervoxel_clustering.cpp:369-371: the labelled clouds / adjacency returned
     * afterwards are the refined ones, as in PCL */
    void refineSupervoxels(int num_itr, std::map<uint32_t, typename pcl::Supervoxel<PointT>::Ptr>& supervoxel_clusters);
    /* getSupervoxelAdjacencyList, :376-384: labels as vertices, centroid distance as edge weight.  The library fills the plain
     * struct above; any other graph type (the reference passes a boost::adjacency_list<setS, setS, undirectedS, uint32_t, float>)
     * goes through the caller's f3ps_copy_adjacency_list(const f3ps::VoxelAdjacencyList&, GraphT&) -- INTEGRATION.md shows the BGL one. */
    void getSupervoxelAdjacencyList(VoxelAdjacencyList& adjacency_list_arg) const;
    template <typename GraphT>
    void getSupervoxelAdjacencyList(GraphT& adjacency_list_arg) const {
        VoxelAdjacencyList plain;
        getSupervoxelAdjacencyList(plain);
        f3ps_copy_adjacency_list(plain, adjacency_list_arg);
    }
    static pcl::PointCloud<pcl::PointNormal>::Ptr makeSupervoxelNormalCloud(
        std::map<uint32_t, typename pcl::Supervoxel<PointT>::Ptr>& supervoxel_clusters);
    float getVoxelResolution() const { return resolution_; }
    float getSeedResolution() const { return seed_resolution_; }
    Handle& handle() { return *h_; }
private:
    void collect(std::map<uint32_t, typename pcl::Supervoxel<PointT>::Ptr>& supervoxel_clusters);   /* makeSupervoxels from the device tables */
    std::shared_ptr<Handle> h_;
    float resolution_, seed_resolution_;
    float color_importance_ = 0.1f, spatial_importance_ = 0.4f, normal_importance_ = 1.0f;   /* PCL defaults */
    bool use_transform_ = true;
    typename pcl::PointCloud<PointT>::ConstPtr input_;
    bool extracted_ = false;
};

} // namespace f3ps

#ifndef F3PS_USE_REAL_PCL
namespace pcl { using f3ps::SupervoxelClustering; }
#endif

#endif /* F3PS_PCL_SHIM_H_ */
