// Stand-in for pcl/point_cloud.h: a vector of points with ::Ptr (oracle/ref_shim/README.md).  Test infrastructure only.
#pragma once
#include <cmath>      // (the real PCL / Eigen headers bring these in; src/testing.cpp relies on that)
#include <cstdint>
#include <iterator>
#include <stdexcept>
#include <vector>
#include "../Eigen/Core"
#include "../boost/make_shared.hpp"
namespace pcl {
template <class PointT> class PointCloud {
public:
    typedef boost::shared_ptr<PointCloud<PointT> > Ptr;
    typedef std::vector<PointT, Eigen::aligned_allocator<PointT> > VectorType;
    typedef typename VectorType::iterator iterator;
    typedef typename VectorType::const_iterator const_iterator;
    VectorType points; uint32_t width = 0, height = 1;
    iterator begin() { return points.begin(); } iterator end() { return points.end(); }
    const_iterator begin() const { return points.begin(); } const_iterator end() const { return points.end(); }
    std::size_t size() const { return points.size(); } bool empty() const { return points.empty(); }
    void push_back(const PointT& p) { points.push_back(p); width = (uint32_t)points.size(); }
};
}  // namespace pcl
