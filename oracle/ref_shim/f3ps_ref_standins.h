// f3ps_ref_standins.h -- container / math stand-ins that let the reference's OWN sources compile where they lie (oracle/Makefile,
// target `ref`; oracle/ref_shim/README.md): src/testing.cpp, src/clustering.cpp, src/clustering_state.cpp, src/color_utilities.cpp.
// TEST INFRASTRUCTURE ONLY.  What is stood in: containers (pcl::PointCloud, pcl::Supervoxel, point structs, boost::make_shared),
// the handful of Eigen operations those files use (3-vector / 4-vector arithmetic in Eigen's evaluation order as the oracle
// restates it: a0 + (a1 + a2) and (a0 + a1) + (a2 + a3)), PCL's computeCentroid / computePointNormal / flipNormalTowardsViewpoint
// (= the oracle's restatement of PCL 1.10, oracle_vccs.cpp -- PCL itself stays parity-unpinned), cv::cvtColor for one pixel
// (= the oracle's 33^3 LUT interpolation, pinned bit-exactly against cv2 4.13.0), PCL's Glasbey table (any table), and the
// console printers (silent; print_debug("left: ...") is captured: it is the reference's own per-merge log line).
// Everything else -- the merge loop, the weight multimap and its order, contains(), init_weights, the adaptive lambda, the CDFs,
// t_c / t_g, mean_color, CIEDE2000, RGB distance, the evaluation scores -- is the reference's code, executed.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <cstdint>
#include <fstream>
#include <iostream>
#include <iterator>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace Eigen {
const int Dynamic = -1;
template <class T> struct aligned_allocator : std::allocator<T> {
    template <class U> struct rebind { typedef aligned_allocator<U> other; };
    aligned_allocator() {}
    template <class U> aligned_allocator(const aligned_allocator<U>&) {}
};
template <class S> struct AnyView {                         // result of `== scalar` and of `.array()`: only any() is asked
    std::vector<S> v;
    bool any() const { for (const S& x : v) if (x) return true; return false; }
};
template <class S, int R, int C> class Matrix {             // column-major, like Eigen's default
public:
    std::size_t rows_ = 0, cols_ = 0; std::vector<S> d;
    Matrix() {}
    Matrix(std::size_t r, std::size_t c) : rows_(r), cols_(c), d(r * c, S(0)) {}
    static Matrix Zero(std::size_t r, std::size_t c) { return Matrix(r, c); }
    S& operator()(std::size_t i, std::size_t j) { return d[j * rows_ + i]; }
    const S& operator()(std::size_t i, std::size_t j) const { return d[j * rows_ + i]; }
    S& operator()(std::size_t i) { return d[i]; }
    const S& operator()(std::size_t i) const { return d[i]; }
    Matrix<S, Dynamic, 1> col(std::size_t j) const {
        Matrix<S, Dynamic, 1> c(rows_, 1);
        for (std::size_t i = 0; i < rows_; ++i) c.d[i] = d[j * rows_ + i];
        return c;
    }
    // Eigen's visitor keeps the FIRST coefficient that attains the maximum (it only replaces on a strictly larger value)
    template <class IndexT> S maxCoeff(IndexT* index) const {
        std::size_t best = 0;
        for (std::size_t i = 1; i < d.size(); ++i) if (d[i] > d[best]) best = i;
        *index = (IndexT)best;
        return d[best];
    }
    AnyView<S> array() const { AnyView<S> a; a.v = d; return a; }
};
template <class S, int R, int C> class Array {
public:
    std::vector<S> d;
    static Array Zero(std::size_t r, std::size_t c) { Array a; a.d.assign(r * c, S(0)); return a; }
    Array operator-(S s) const { Array a = *this; for (S& x : a.d) x -= s; return a; }
    S& operator()(std::size_t i) { return d[i]; }
    const S& operator()(std::size_t i) const { return d[i]; }
    template <class T> AnyView<bool> operator==(T s) const { AnyView<bool> a; for (const S& x : d) a.v.push_back(x == (S)s); return a; }
};
struct Vector3f {                                            // fixed-size float vector: redux order a0 + (a1 + a2)
    float v[3];
    Vector3f() { v[0] = v[1] = v[2] = 0; }
    Vector3f(float x, float y, float z) { v[0] = x; v[1] = y; v[2] = z; }
    float& operator[](int i) { return v[i]; } float operator[](int i) const { return v[i]; }
    Vector3f operator-(const Vector3f& o) const { return Vector3f(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
    Vector3f& operator/=(float s) { v[0] /= s; v[1] /= s; v[2] /= s; return *this; }
    float dot(const Vector3f& o) const { return v[0] * o.v[0] + (v[1] * o.v[1] + v[2] * o.v[2]); }
    float norm() const { return std::sqrt(v[0] * v[0] + (v[1] * v[1] + v[2] * v[2])); }
    Vector3f cross(const Vector3f& o) const { return Vector3f(v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]); }
};
struct Vector4f {                                            // redux order (a0 + a1) + (a2 + a3)
    float v[4];
    Vector4f() { v[0] = v[1] = v[2] = v[3] = 0; }
    float& operator[](int i) { return v[i]; } float operator[](int i) const { return v[i]; }
    void normalize() {                                       // Eigen: z = squaredNorm(); if (z > 0) *this /= sqrt(z)
        const float z = (v[0] * v[0] + v[1] * v[1]) + (v[2] * v[2] + v[3] * v[3]);
        if (z > 0.0f) { const float s = std::sqrt(z); v[0] /= s; v[1] /= s; v[2] /= s; v[3] /= s; }
    }
};
}  // namespace Eigen

namespace boost { using std::shared_ptr; using std::make_shared; }

namespace pcl {
struct RGB { union { struct { uint8_t b, g, r, a; }; uint32_t rgba; }; };
struct PointXYZL { float x, y, z; uint32_t label; };
struct PointXYZRGBA {
    float x, y, z; union { struct { uint8_t b, g, r, a; }; uint32_t rgba; };
    PointXYZRGBA() : x(0), y(0), z(0), rgba(0) {}
    Eigen::Vector3f getVector3fMap() const { return Eigen::Vector3f(x, y, z); }
};
struct PointXYZRGBL {
    float x, y, z; union { struct { uint8_t b, g, r, a; }; uint32_t rgba; float rgb; }; uint32_t label;
    PointXYZRGBL() : x(0), y(0), z(0), rgba(0), label(0) {}
};
struct Normal {
    float normal_x, normal_y, normal_z, curvature;
    Normal() : normal_x(0), normal_y(0), normal_z(0), curvature(0) {}
    Eigen::Vector3f getNormalVector3fMap() const { return Eigen::Vector3f(normal_x, normal_y, normal_z); }
};
template <class PointT> class PointCloud {
public:
    typedef boost::shared_ptr<PointCloud<PointT> > Ptr;
    typedef boost::shared_ptr<const PointCloud<PointT> > ConstPtr;
    typedef std::vector<PointT, Eigen::aligned_allocator<PointT> > VectorType;
    typedef typename VectorType::iterator iterator;
    typedef typename VectorType::const_iterator const_iterator;
    VectorType points; uint32_t width = 0, height = 1;
    iterator begin() { return points.begin(); } iterator end() { return points.end(); }
    const_iterator begin() const { return points.begin(); } const_iterator end() const { return points.end(); }
    std::size_t size() const { return points.size(); } bool empty() const { return points.empty(); }
    void push_back(const PointT& p) { points.push_back(p); width = (uint32_t)points.size(); }
    PointT& at(std::size_t i) { return points.at(i); } const PointT& at(std::size_t i) const { return points.at(i); }
    PointCloud& operator+=(const PointCloud& o) { points.insert(points.end(), o.points.begin(), o.points.end()); width = (uint32_t)points.size(); return *this; }
    PointCloud operator+(const PointCloud& o) const { PointCloud c = *this; c += o; return c; }     // voxels_ = a ++ b (clustering.cpp:408)
};
inline void copy_fields(const PointXYZL& a, PointXYZRGBL& b) { b.x = a.x; b.y = a.y; b.z = a.z; b.label = a.label; }
inline void copy_fields(const PointXYZRGBL& a, PointXYZRGBA& b) { b.x = a.x; b.y = a.y; b.z = a.z; b.rgba = a.rgba; }
inline void copy_fields(const PointXYZRGBA& a, PointXYZRGBL& b) { b.x = a.x; b.y = a.y; b.z = a.z; b.rgba = a.rgba; }
inline void copy_fields(const PointXYZRGBL& a, PointXYZL& b) { b.x = a.x; b.y = a.y; b.z = a.z; b.label = a.label; }
template <class A, class B> void copyPointCloud(const PointCloud<A>& in, PointCloud<B>& out) {
    out.points.clear();
    for (const A& p : in.points) { B q = B(); copy_fields(p, q); out.points.push_back(q); }
    out.width = (uint32_t)out.points.size(); out.height = 1;
}
template <class PointT> class Supervoxel {
public:
    typedef boost::shared_ptr<Supervoxel<PointT> > Ptr;
    Supervoxel() : voxels_(new PointCloud<PointT>()), normals_(new PointCloud<Normal>()) {}
    Normal normal_; PointXYZRGBA centroid_;
    typename PointCloud<PointT>::Ptr voxels_; typename PointCloud<Normal>::Ptr normals_;
};
// PCL 1.10 as the oracle restates it (oracle_vccs.cpp: plane_from_accu + SURVEY.md A.3 / A.7); defined in ref_standins.cpp
void computeCentroid(const PointCloud<PointXYZRGBA>& cloud, PointXYZRGBA& centroid);
void computePointNormal(const PointCloud<PointXYZRGBA>& cloud, Eigen::Vector4f& plane, float& curvature);
void flipNormalTowardsViewpoint(const PointXYZRGBA& p, float vx, float vy, float vz, Eigen::Vector4f& n);
struct GlasbeyLUT { static RGB at(std::size_t i); static std::size_t size(); };
namespace console {
void print_debug(const char* fmt, ...);       // captures the reference's per-merge line "left: %de/%dp - w: %f - [%d, %d]..." (clustering.cpp:390-392)
inline void print_info(const char*, ...) {} inline void print_warn(const char*, ...) {} inline void print_error(const char*, ...) {}
inline void print_highlight(const char*, ...) {}
}
}  // namespace pcl

namespace cv {
const int COLOR_RGB2Lab = 45, COLOR_Lab2RGB = 57;
struct Scalar { double v[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { v[0] = a; v[1] = b; v[2] = c; v[3] = d; } };
struct Vec3f { float v[3]; float& operator[](int i) { return v[i]; } float operator[](int i) const { return v[i]; } };
struct Mat {                                               // one CV_32FC3 pixel is all color_conversion builds (color_utilities.cpp:57-62)
    Vec3f px;
    Mat(int, int, int, const Scalar& s) { px.v[0] = (float)s.v[0]; px.v[1] = (float)s.v[1]; px.v[2] = (float)s.v[2]; }
    template <class T> T& at(int, int) { return px; }
};
void cvtColor(const Mat& in, Mat& out, int code);           // RGB2Lab = the oracle's LUT path (pinned against cv2 4.13.0); Lab2RGB unsupported
}  // namespace cv
#define CV_32FC3 21
