// ref_standins.cpp -- the non-inline part of oracle/ref_shim/f3ps_ref_standins.h (test infrastructure; oracle/ref_shim/README.md):
// PCL's centroid / plane fit / normal flip exactly as the oracle restates PCL 1.10 (oracle_vccs.cpp, oracle_merge.cpp:
// region_geometry), cv::cvtColor through the oracle's LUT path, a Glasbey-like table, and the capture of the reference's per-merge
// debug line.
#include "ref_shim/f3ps_ref_standins.h"

#include <cstring>
#include <limits>

#include "oracle.h"
#include "ref_capture.h"

namespace f3ps_ref {
const int16_t* lab_lut = nullptr;
std::vector<MergeLine>* sink = nullptr;
}

namespace pcl {
void computeCentroid(const PointCloud<PointXYZRGBA>& cloud, PointXYZRGBA& c) {            // CentroidPoint: float sums in cloud order / n
    float sx = 0, sy = 0, sz = 0;
    for (const PointXYZRGBA& p : cloud.points) { sx += p.x; sy += p.y; sz += p.z; }
    const float n = (float)cloud.size();
    c = PointXYZRGBA(); c.x = sx / n; c.y = sy / n; c.z = sz / n;
}
void computePointNormal(const PointCloud<PointXYZRGBA>& cloud, Eigen::Vector4f& plane, float& curvature) {
    float n4[4];
    if (cloud.size() < 3) { n4[0] = n4[1] = n4[2] = n4[3] = std::numeric_limits<float>::quiet_NaN(); curvature = n4[0]; }
    else {
        float accu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (const PointXYZRGBA& p : cloud.points) {
            const float x = p.x, y = p.y, z = p.z;
            accu[0] += x * x; accu[1] += x * y; accu[2] += x * z; accu[3] += y * y; accu[4] += y * z; accu[5] += z * z;
            accu[6] += x; accu[7] += y; accu[8] += z;
        }
        f3ps_oracle::plane_from_accu(accu, (int)cloud.size(), n4, &curvature);
    }
    for (int k = 0; k < 4; ++k) plane[k] = n4[k];
}
void flipNormalTowardsViewpoint(const PointXYZRGBA& p, float vx, float vy, float vz, Eigen::Vector4f& n) {
    const float cos_theta = ((vx - p.x) * n[0] + (vy - p.y) * n[1]) + ((vz - p.z) * n[2] + 0.0f * n[3]);
    if (cos_theta < 0) { n[0] *= -1; n[1] *= -1; n[2] *= -1; }       // (PCL also rewrites n[3]; the caller overwrites it, clustering.cpp:419)
}
RGB GlasbeyLUT::at(std::size_t i) { RGB c; c.rgba = 0; c.r = (uint8_t)(37 * i + 11); c.g = (uint8_t)(91 * i + 53); c.b = (uint8_t)(173 * i + 7); return c; }
std::size_t GlasbeyLUT::size() { return 256; }
namespace console {
void print_debug(const char* fmt, ...) {
    if (!f3ps_ref::sink || std::strncmp(fmt, "left:", 5) != 0) return;
    va_list ap; va_start(ap, fmt);
    f3ps_ref::MergeLine m;
    m.edges_left = (uint32_t)va_arg(ap, std::size_t); m.regions_left = (uint32_t)va_arg(ap, std::size_t);     // weight_map.size(), segments.size()
    m.w = (float)va_arg(ap, double);
    m.a = va_arg(ap, uint32_t); m.b = va_arg(ap, uint32_t);
    va_end(ap);
    f3ps_ref::sink->push_back(m);
}
}
}  // namespace pcl

namespace cv {
void cvtColor(const Mat& in, Mat& out, int code) {
    if (code != COLOR_RGB2Lab || !f3ps_ref::lab_lut) throw std::logic_error("cv::cvtColor stand-in: only COLOR_RGB2Lab with a LUT");
    f3ps_oracle::rgb_unit2lab(f3ps_ref::lab_lut, in.px.v, out.px.v);
}
}
