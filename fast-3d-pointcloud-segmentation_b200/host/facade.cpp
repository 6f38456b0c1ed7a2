// facade.cpp -- host side of the drop-in: f3ps::Handle, f3ps::SupervoxelClustering (pcl::SupervoxelClustering
// as the reference consumes it), Clustering, ClusteringState, ColorUtilities -- all thin layers over the C ABI
// (include/f3ps.h).  Nothing here computes on the CPU what the reference computes in its hot path; the host
// only marshals PCL-shaped containers in and out.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <limits>
#include <stdexcept>

#include "supervoxel_clustering/clustering.h"

namespace f3ps {

Handle::Handle(int device) {
    int rc = f3ps_create(device, nullptr, &ctx_);
    if (rc != F3PS_OK) throw std::runtime_error("f3ps_create failed: no usable CUDA device (there is no CPU fallback)");
}
Handle::~Handle() { if (ctx_) f3ps_destroy(ctx_); }
void Handle::check(int status) const {
    if (status == F3PS_OK) return;
    const std::string msg = f3ps_last_error(ctx_);
    if (status == F3PS_ERR_LOGIC) throw std::logic_error(msg);
    if (status == F3PS_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
}

template <typename PointT>
SupervoxelClustering<PointT>::SupervoxelClustering(float voxel_resolution, float seed_resolution, int device)
    : h_(new Handle(device)), resolution_(voxel_resolution), seed_resolution_(seed_resolution) {}

template <typename PointT>
void SupervoxelClustering<PointT>::extract(std::map<uint32_t, typename pcl::Supervoxel<PointT>::Ptr>& out) {
    out.clear();
    if (!input_ || input_->points.empty()) return;           // prepareForSegmentation returns false on an empty cloud
    f3ps_ctx* c = h_->get();
    // fold_negative_z = 0: main() already did that clean-up before setInputCloud
    h_->check(f3ps_set_vccs_params(c, resolution_, seed_resolution_, color_importance_, spatial_importance_, normal_importance_,
                                   use_transform_ ? 1 : 0, 0));
    h_->check(f3ps_set_input(c, input_->points.data(), (int64_t)input_->points.size(), (int)sizeof(PointT), 0));
    h_->check(f3ps_extract(c));
    extracted_ = true;
    collect(out);
}

// pcl::SupervoxelClustering::refineSupervoxels (src/supervoxel_clustering.cpp:369-371): num_itr rounds of refineNormals /
// reseedSupervoxels / expandSupervoxels on the device (f3ps_refine), then makeSupervoxels into `out`.  As in PCL, the labelled
// clouds and the adjacency the object returns afterwards are those of the refined supervoxels.
template <typename PointT>
void SupervoxelClustering<PointT>::refineSupervoxels(int num_itr, std::map<uint32_t, typename pcl::Supervoxel<PointT>::Ptr>& out) {
    if (!extracted_) { fprintf(stderr, "Supervoxels must be extracted before they can be refined\n"); return; }   // PCL_FATAL + return
    f3ps_ctx* c = h_->get();
    h_->check(f3ps_refine(c, num_itr));
    h_->check(f3ps_graph(c));
    collect(out);
}

// getSupervoxelAdjacencyList: vertices = supervoxel labels, one undirected edge per adjacent pair, edge weight = distance between
// the two centroids (PCL builds a boost::adjacency_list<setS, setS, undirectedS, uint32_t, float>; Boost is not a dependency here).
template <typename PointT>
void SupervoxelClustering<PointT>::getSupervoxelAdjacencyList(VoxelAdjacencyList& g) const {
    g.vertices.clear(); g.edges.clear();
    if (!extracted_) return;
    f3ps_ctx* c = h_->get();
    f3ps_counts n; h_->check(f3ps_get_counts(c, &n));
    const size_t S = (size_t)n.n_supervoxels;
    std::vector<uint32_t> label(S); std::vector<float> cen(3 * S);
    h_->check(f3ps_get_supervoxels(c, label.data(), cen.data(), nullptr, nullptr, nullptr, (int64_t)S));
    std::map<uint32_t, size_t> at;
    for (size_t s = 0; s < S; ++s) { g.vertices.insert(label[s]); at[label[s]] = s; }
    std::multimap<uint32_t, uint32_t> adj; getSupervoxelAdjacency(adj);
    for (auto& kv : adj) {
        if (kv.first >= kv.second) continue;
        const float* a = &cen[3 * at[kv.first]]; const float* b = &cen[3 * at[kv.second]];
        const float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
        g.edges[std::make_pair(kv.first, kv.second)] = std::sqrt(dx * dx + dy * dy + dz * dz);
    }
}

template <typename PointT>
void SupervoxelClustering<PointT>::collect(std::map<uint32_t, typename pcl::Supervoxel<PointT>::Ptr>& out) {
    out.clear();
    f3ps_ctx* c = h_->get();
    f3ps_counts n; h_->check(f3ps_get_counts(c, &n));
    const size_t V = (size_t)n.n_voxels, S = (size_t)n.n_supervoxels;
    std::vector<float> vxyz(3 * V), vn(4 * V), vcurv(V);
    std::vector<uint32_t> vrgba(V);
    h_->check(f3ps_get_voxel_centroids(c, vxyz.data(), nullptr, vrgba.data(), nullptr, (int64_t)V));
    h_->check(f3ps_get_voxel_normals(c, vn.data(), vcurv.data(), (int64_t)V));
    std::vector<uint32_t> label(S); std::vector<float> cen(3 * S), rgb(3 * S), nrm(4 * S); std::vector<int32_t> cnt(S);
    h_->check(f3ps_get_supervoxels(c, label.data(), cen.data(), rgb.data(), nrm.data(), cnt.data(), (int64_t)S));
    std::vector<int32_t> idx(V + S); std::vector<int64_t> off(S + 1);     // + S: a helper may also list one phantom leaf
    h_->check(f3ps_get_supervoxel_voxels(c, idx.data(), off.data(), (int64_t)(V + S), (int64_t)S));
    for (size_t s = 0; s < S; ++s) {
        typename pcl::Supervoxel<PointT>::Ptr sv(new pcl::Supervoxel<PointT>());
        sv->centroid_.x = cen[3 * s]; sv->centroid_.y = cen[3 * s + 1]; sv->centroid_.z = cen[3 * s + 2];
        sv->centroid_.rgba = ((uint32_t)rgb[3 * s] << 16) | ((uint32_t)rgb[3 * s + 1] << 8) | (uint32_t)rgb[3 * s + 2];
        sv->normal_.normal_x = nrm[4 * s]; sv->normal_.normal_y = nrm[4 * s + 1]; sv->normal_.normal_z = nrm[4 * s + 2];
        sv->normal_.curvature = nrm[4 * s + 3];
        const size_t m = (size_t)(off[s + 1] - off[s]);
        sv->voxels_->resize(m); sv->normals_->resize(m);
        for (size_t k = 0; k < m; ++k) {
            const int v = idx[off[s] + k];
            PointT& p = sv->voxels_->points[k];
            p.x = vxyz[3 * v]; p.y = vxyz[3 * v + 1]; p.z = vxyz[3 * v + 2]; p.rgba = vrgba[v];
            pcl::Normal& q = sv->normals_->points[k];
            q.normal_x = vn[4 * v]; q.normal_y = vn[4 * v + 1]; q.normal_z = vn[4 * v + 2]; q.curvature = vcurv[v];
        }
        out[label[s]] = sv;
    }
}

template <typename PointT>
typename pcl::PointCloud<PointT>::Ptr SupervoxelClustering<PointT>::getVoxelCentroidCloud() const {
    typename pcl::PointCloud<PointT>::Ptr cloud(new pcl::PointCloud<PointT>());
    if (!extracted_) return cloud;
    f3ps_counts n; h_->check(f3ps_get_counts(h_->get(), &n));
    const size_t V = (size_t)n.n_voxels;
    std::vector<float> vxyz(3 * V); std::vector<uint32_t> vrgba(V);
    h_->check(f3ps_get_voxel_centroids(h_->get(), vxyz.data(), nullptr, vrgba.data(), nullptr, (int64_t)V));
    cloud->resize(V);
    for (size_t v = 0; v < V; ++v) { PointT& p = cloud->points[v]; p.x = vxyz[3 * v]; p.y = vxyz[3 * v + 1]; p.z = vxyz[3 * v + 2]; p.rgba = vrgba[v]; }
    return cloud;
}

template <typename PointT>
pcl::PointCloud<pcl::PointXYZL>::Ptr SupervoxelClustering<PointT>::getLabeledVoxelCloud() const {
    pcl::PointCloud<pcl::PointXYZL>::Ptr cloud(new pcl::PointCloud<pcl::PointXYZL>());
    if (!extracted_) return cloud;
    f3ps_counts n; h_->check(f3ps_get_counts(h_->get(), &n));
    const size_t V = (size_t)n.n_voxels;
    std::vector<float> vxyz(3 * V); std::vector<uint32_t> lab(V);
    h_->check(f3ps_get_voxel_centroids(h_->get(), vxyz.data(), nullptr, nullptr, nullptr, (int64_t)V));
    h_->check(f3ps_get_voxel_labels(h_->get(), lab.data(), nullptr, (int64_t)V));
    cloud->resize(V);
    for (size_t v = 0; v < V; ++v) { pcl::PointXYZL& p = cloud->points[v]; p.x = vxyz[3 * v]; p.y = vxyz[3 * v + 1]; p.z = vxyz[3 * v + 2]; p.label = lab[v]; }
    return cloud;
}

template <typename PointT>
pcl::PointCloud<pcl::PointXYZL>::Ptr SupervoxelClustering<PointT>::getLabeledCloud() const {
    pcl::PointCloud<pcl::PointXYZL>::Ptr cloud(new pcl::PointCloud<pcl::PointXYZL>());
    if (!extracted_ || !input_) return cloud;
    f3ps_counts n; h_->check(f3ps_get_counts(h_->get(), &n));
    std::vector<int32_t> pv((size_t)n.n_points); std::vector<uint32_t> lab((size_t)n.n_voxels);
    h_->check(f3ps_get_point_voxel(h_->get(), pv.data(), n.n_points));
    h_->check(f3ps_get_voxel_labels(h_->get(), lab.data(), nullptr, n.n_voxels));
    cloud->resize(input_->points.size());
    for (size_t i = 0; i < input_->points.size(); ++i) {
        pcl::PointXYZL& p = cloud->points[i];
        p.x = input_->points[i].x; p.y = input_->points[i].y; p.z = input_->points[i].z;
        p.label = pv[i] >= 0 ? lab[pv[i]] : 0;
    }
    return cloud;
}

template <typename PointT>
void SupervoxelClustering<PointT>::getSupervoxelAdjacency(std::multimap<uint32_t, uint32_t>& label_adjacency) const {
    label_adjacency.clear();
    if (!extracted_) return;
    f3ps_counts n; h_->check(f3ps_get_counts(h_->get(), &n));
    std::vector<uint32_t> pairs(4 * (size_t)n.n_edges);
    h_->check(f3ps_get_adjacency(h_->get(), pairs.data(), 2 * (int64_t)n.n_edges));
    for (size_t i = 0; i < 2 * (size_t)n.n_edges; ++i) label_adjacency.insert(std::make_pair(pairs[2 * i], pairs[2 * i + 1]));
}

template <typename PointT>
pcl::PointCloud<pcl::PointNormal>::Ptr SupervoxelClustering<PointT>::makeSupervoxelNormalCloud(
        std::map<uint32_t, typename pcl::Supervoxel<PointT>::Ptr>& supervoxel_clusters) {
    pcl::PointCloud<pcl::PointNormal>::Ptr cloud(new pcl::PointCloud<pcl::PointNormal>());
    cloud->resize(supervoxel_clusters.size());
    size_t i = 0;
    for (auto& kv : supervoxel_clusters) {
        pcl::PointNormal& p = cloud->points[i++];
        p.x = kv.second->centroid_.x; p.y = kv.second->centroid_.y; p.z = kv.second->centroid_.z;
        p.normal_x = kv.second->normal_.normal_x; p.normal_y = kv.second->normal_.normal_y; p.normal_z = kv.second->normal_.normal_z;
        p.curvature = kv.second->normal_.curvature;
    }
    return cloud;
}

template class SupervoxelClustering<pcl::PointXYZRGBA>;

static Handle& shared_handle() { static Handle h(0); return h; }

} // namespace f3ps

#ifndef F3PS_USE_REAL_PCL
namespace pcl {
template <> void copyPointCloud(const PointCloud<PointXYZRGBL>& in, PointCloud<PointXYZRGBA>& out) {
    out.points.resize(in.size()); out.width = in.width; out.height = in.height; out.is_dense = in.is_dense;
    for (size_t i = 0; i < in.size(); ++i) { out[i].x = in[i].x; out[i].y = in[i].y; out[i].z = in[i].z; out[i].rgba = in[i].rgba; }
}
template <> void copyPointCloud(const PointCloud<PointXYZRGBL>& in, PointCloud<PointXYZL>& out) {
    out.points.resize(in.size()); out.width = in.width; out.height = in.height; out.is_dense = in.is_dense;
    for (size_t i = 0; i < in.size(); ++i) { out[i].x = in[i].x; out[i].y = in[i].y; out[i].z = in[i].z; out[i].label = in[i].label; }
}
template <> void copyPointCloud(const PointCloud<PointXYZL>& in, PointCloud<PointXYZRGBL>& out) {
    out.points.resize(in.size()); out.width = in.width; out.height = in.height; out.is_dense = in.is_dense;
    for (size_t i = 0; i < in.size(); ++i) { out[i].x = in[i].x; out[i].y = in[i].y; out[i].z = in[i].z; out[i].label = in[i].label; }
}
template <> void copyPointCloud(const PointCloud<PointXYZRGBA>& in, PointCloud<PointXYZRGBL>& out) {
    out.points.resize(in.size()); out.width = in.width; out.height = in.height; out.is_dense = in.is_dense;
    for (size_t i = 0; i < in.size(); ++i) { out[i].x = in[i].x; out[i].y = in[i].y; out[i].z = in[i].z; out[i].rgba = in[i].rgba; }
}
}
#endif

// ---------------------------------------------------------------------------------------------------
ClusteringState::ClusteringState(ClusteringT s, WeightMapT w) { set_segments(s); set_weight_map(w); }

// ---------------------------------------------------------------------------------------------------
// 256 maximally distinct colours by the construction of Glasbey et al. (greedy maximisation of the minimum CIE76 distance over
// an 18^3 sRGB lattice, 15 < L* < 95; generated offline).  pcl::GlasbeyLUT's own table is not part of /root/reference.
static const uint8_t kGlasbey[256 * 3] = {75,180,75,0,0,255,255,0,90,30,135,210,255,165,0,75,45,15,255,120,255,255,225,210,0,255,255,45,0,120,225,255,0,0,255,30,30,120,105,165,90,135,180,60,0,240,240,135,30,45,75,135,120,0,0,105,255,210,0,255,225,195,255,255,135,135,210,0,135,30,75,0,255,0,0,150,210,165,135,0,45,150,135,105,0,255,180,135,210,240,150,105,210,150,180,0,255,180,120,90,0,195,15,45,30,165,255,105,75,15,75,135,135,165,240,210,15,165,0,165,165,90,60,255,135,210,30,60,135,0,255,120,45,195,0,120,150,75,0,135,165,180,90,255,255,105,0,255,0,210,195,120,30,90,45,60,180,255,165,0,135,75,210,150,165,255,90,75,225,75,120,210,195,135,90,90,90,150,255,0,0,180,180,0,195,150,135,0,105,0,135,255,150,165,255,210,180,60,60,120,0,105,60,195,105,30,0,105,90,30,165,180,180,45,75,255,180,255,240,105,45,135,180,0,30,225,240,255,0,90,135,180,135,75,255,135,75,0,180,240,120,255,210,195,135,210,210,240,90,120,105,165,135,30,255,165,75,90,195,75,165,150,105,105,255,180,165,150,195,90,75,105,60,120,30,75,135,210,0,225,240,195,60,90,195,0,75,90,165,0,90,210,45,225,135,75,0,105,240,135,150,195,255,0,225,255,210,90,75,90,165,0,225,120,150,240,150,255,255,30,135,255,90,195,135,165,135,240,0,45,120,90,120,105,135,135,90,105,15,255,165,210,105,210,135,255,210,120,0,75,45,45,60,15,75,165,135,135,120,255,225,195,210,255,180,75,90,60,120,195,90,210,0,45,195,195,225,150,0,45,150,225,60,15,120,150,30,105,255,75,90,15,30,225,135,105,180,180,90,165,150,210,255,180,255,180,195,225,150,75,150,150,15,195,210,45,75,180,135,0,105,0,150,240,135,30,45,45,45,120,60,45,195,195,165,135,15,0,105,225,225,30,180,210,0,30,225,120,90,60,195,150,135,90,210,75,255,225,90,0,135,0,195,75,135,120,120,210,255,30,255,180,150,255,210,0,165,180,240,120,255,195,45,150,45,150,195,135,180,195,210,0,90,165,90,45,30,75,255,210,165,195,225,210,255,90,165,90,150,195,150,135,135,195,255,195,195,255,75,60,90,75,195,120,255,135,195,180,180,90,30,15,105,30,255,0,165,195,120,120,165,45,240,0,105,195,165,165,30,45,30,105,120,75,240,165,45,45,105,165,180,255,240,165,180,150,60,0,180,105,30,0,150,255,240,0,195,210,90,135,135,75,210,90,105,150,105,0,210,0,90,75,75,120,15,105,60,75,60,45,120,60,105,135,255,180,90,90,60,15,255,150,75,60,75,150,180,120,255,165,180,120,255,120,90,60,0,210,165,120,0,210,195,255,90,105,75,105,120,0,165,15,105,0,15,255,90,255,60,150,150,120,225,195,240,75,210,90,120,180,165,120,90,75,60,150,210,135,75,240,240,75,75,135,90,45,75,210,60,30,30,135,75,90,105,225,0,150,60,195,45,225,75,210,90,240,90,165,255,195,180,0,105,75,75,180,240,255,75,120,45,165,165,120,15,135,45,0,45,60,225,90,45,180,150,180,150,165,210,15,75,135,135,210,75,75,0,105,180,15,0,255,195,240,255,120,90,150,105,180,180,30,75,105,90,210,0,255,225,0,210,105,15,225,15,150,60,30};

uint8_t* ColorUtilities::get_glasbey(uint32_t label) {
    uint8_t* out = new uint8_t[3];
    const uint8_t* c = kGlasbey + 3 * (label % 256u);
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2];
    return out;
}
float* ColorUtilities::mean_color(SupervoxelT::Ptr s) {
    // one supervoxel as a one-node graph: the device computes the running mean exactly as the merge loop does
    f3ps::Handle& h = f3ps::shared_handle();
    const size_t n = s->voxels_->size();
    std::vector<float> xyz(3 * n); std::vector<uint32_t> rgba(n);
    for (size_t i = 0; i < n; ++i) { const PointT& p = s->voxels_->points[i]; xyz[3 * i] = p.x; xyz[3 * i + 1] = p.y; xyz[3 * i + 2] = p.z; rgba[i] = p.rgba; }
    const uint32_t label = 1; const int64_t off[2] = {0, (int64_t)n}; const float cen[3] = {0, 0, 0}, nrm[3] = {0, 0, 1};
    h.check(f3ps_set_graph(h.get(), (int64_t)n, xyz.data(), rgba.data(), 1, &label, off, cen, nrm, 0, nullptr));
    float* rgb = new float[3]();
    h.check(f3ps_get_region_mean_color(h.get(), 0, rgb));
    return rgb;
}
// cv::cvtColor(CV_32FC3, COLOR_Lab2RGB) stand-in (src/color_utilities.cpp:52-69, code CV_Lab2RGB): the analytic CIE L*a*b* ->
// linear -> sRGB chain in float.  Only convert_test calls it (print-only in the reference); within 0.5 / 255 of OpenCV 4.13.
static void lab2rgb_host(const float lab[3], float rgb[3]) {
    const float fy = (lab[0] + 16.0f) / 116.0f, fx = fy + lab[1] / 500.0f, fz = fy - lab[2] / 200.0f;
    auto finv = [](float t) { return t > 6.0f / 29.0f ? t * t * t : (t - 16.0f / 116.0f) / 7.787f; };
    const float X = 0.950456f * finv(fx), Y = finv(fy), Z = 1.088754f * finv(fz);
    const float lin[3] = {3.240479f * X - 1.53715f * Y - 0.498535f * Z, -0.969256f * X + 1.875991f * Y + 0.041556f * Z,
                          0.055648f * X - 0.204043f * Y + 1.057311f * Z};
    for (int k = 0; k < 3; ++k) {
        float v = std::min(1.0f, std::max(0.0f, lin[k]));
        rgb[k] = v <= 0.0031308f ? 12.92f * v : 1.055f * std::pow(v, 1.0f / 2.4f) - 0.055f;
    }
}
float* ColorUtilities::color_conversion(float in[3], int code) {
    float* out = new float[3]();
    if (code == 0) {
        f3ps::Handle& h = f3ps::shared_handle();
        h.check(f3ps_test_rgb2lab(h.get(), in, out, 1));            // OpenCV's LUT + trilinear interpolation, bit-exact on the device
    } else lab2rgb_host(in, out);
    return out;
}
float* ColorUtilities::rgb2lab(float rgb[3]) { return color_conversion(rgb, 0); }            // the kernel takes 0..255 as the reference's caller passes it (:151-160)
float* ColorUtilities::lab2rgb(float lab[3]) {
    float* out = color_conversion(lab, 1);
    for (int k = 0; k < 3; ++k) out[k] *= 255.0f;                                             // :169-178
    return out;
}
// lab_ciede00 with weights other than 1 (nothing in the reference passes any): the formula of src/color_utilities.cpp:190-294 on the host
static float ciede00_weighted(const float l1[3], const float l2[3], double kL, double kC, double kH) {
    const double PI = 3.14159265358979323846;
    const double Cab = ((double)std::sqrt(l1[1] * l1[1] + l1[2] * l1[2]) + (double)std::sqrt(l2[1] * l2[1] + l2[2] * l2[2])) / 2.0;
    const double G = 0.5 * (1.0 - std::sqrt(std::pow(Cab, 7.0) / (std::pow(Cab, 7.0) + std::pow(25.0, 7.0))));
    const double ap1 = (1.0 + G) * l1[1], ap2 = (1.0 + G) * l2[1];
    const double Cp1 = std::sqrt(ap1 * ap1 + (double)(l1[2] * l1[2])), Cp2 = std::sqrt(ap2 * ap2 + (double)(l2[2] * l2[2]));
    const double Cpp = Cp1 * Cp2;
    double hp1 = std::atan2((double)l1[2], ap1); if (hp1 < 0) hp1 += 2 * PI; if (std::fabs(ap1) + std::fabs(l1[2]) == 0) hp1 = 0;
    double hp2 = std::atan2((double)l2[2], ap2); if (hp2 < 0) hp2 += 2 * PI; if (std::fabs(ap2) + std::fabs(l2[2]) == 0) hp2 = 0;
    const double dL = (double)(l2[0] - l1[0]), dC = Cp2 - Cp1;
    double dhp = hp2 - hp1; if (dhp > PI) dhp -= 2 * PI; if (dhp < -PI) dhp += 2 * PI; if (Cpp == 0) dhp = 0;
    const double dH = 2.0 * std::sqrt(Cpp) * std::sin(dhp / 2.0);
    const double Lp = (double)(l1[0] + l2[0]) / 2.0, Cp = (Cp1 + Cp2) / 2.0;
    double hp = (hp1 + hp2) / 2.0; if (std::fabs(hp1 - hp2) > PI) hp -= PI; if (hp < 0) hp += 2 * PI; if (Cpp == 0) hp = hp1 + hp2;
    const double Lpm = (Lp - 50.0) * (Lp - 50.0);
    const double SL = 1.0 + 0.015 * Lpm / std::sqrt(20.0 + Lpm), SC = 1.0 + 0.045 * Cp;
    const double T = 1.0 - 0.17 * std::cos(hp - PI / 6.0) + 0.24 * std::cos(2.0 * hp) + 0.32 * std::cos(3.0 * hp + PI / 30.0) - 0.20 * std::cos(4.0 * hp - 63.0 * PI / 180.0);
    const double SH = 1.0 + 0.015 * Cp * T;
    const double dth = (30.0 * PI / 180.0) * std::exp(-std::pow((180.0 / PI * hp - 275.0) / 25.0, 2.0));
    const double RC = 2.0 * std::sqrt(std::pow(Cp, 7.0) / (std::pow(Cp, 7.0) + std::pow(25.0, 7.0)));
    const double RT = -std::sin(2.0 * dth) * RC;
    const double tL = dL / (kL * SL), tC = dC / (kC * SC), tH = dH / (kH * SH);
    return (float)std::sqrt(tL * tL + tC * tC + tH * tH + RT * tC * tH);
}
float ColorUtilities::lab_ciede00(float lab1[3], float lab2[3], double kL, double kC, double kH) {
    if (kL != 1.0 || kC != 1.0 || kH != 1.0) return ciede00_weighted(lab1, lab2, kL, kC, kH);
    float d; f3ps::Handle& h = f3ps::shared_handle();
    h.check(f3ps_test_lab_ciede00(h.get(), lab1, lab2, &d, 1));
    return d;
}
float ColorUtilities::rgb_eucl(float rgb1[3], float rgb2[3]) {
    float d; f3ps::Handle& h = f3ps::shared_handle();
    h.check(f3ps_test_rgb_eucl(h.get(), rgb1, rgb2, &d, 1));
    return d;
}
void ColorUtilities::rgb_test() {                                    // src/color_utilities.cpp:324-349: prints value and expectation
    float c[8][3] = {{0, 0, 0}, {255, 255, 255}, {255, 0, 0}, {0, 255, 0}, {0, 255, 0}, {255, 0, 255}, {100, 20, 35}, {104, 20, 32}};
    const int pairs[7][2] = {{0, 0}, {0, 1}, {1, 1}, {0, 2}, {3, 0}, {4, 5}, {6, 7}};
    const float expect[7] = {0, 441.672943f, 0, 255, 255, 441.672943f, 5};
    for (int i = 0; i < 7; ++i) std::printf("RGB distance test %d: %f (should be %f)\n", i + 1, rgb_eucl(c[pairs[i][0]], c[pairs[i][1]]), expect[i]);
}
float ColorUtilities::ciede00_test(float L1, float a1, float b1, float L2, float a2, float b2, float result) {
    float l1[3] = {L1, a1, b1}, l2[3] = {L2, a2, b2};
    const float d = lab_ciede00(l1, l2);
    std::printf("CIEDE00 test: %f (should be %f)\n", d, result);
    return std::fabs(d - result);
}
void ColorUtilities::lab_test() {                                    // :354-460, the Sharma-Wu-Dalal table (34 pairs)
    static const float v[34][7] = {
        {50.0000f, 2.6772f, -79.7751f, 50.0000f, 0.0000f, -82.7485f, 2.0425f}, {50.0000f, 3.1571f, -77.2803f, 50.0000f, 0.0000f, -82.7485f, 2.8615f},
        {50.0000f, 2.8361f, -74.0200f, 50.0000f, 0.0000f, -82.7485f, 3.4412f}, {50.0000f, -1.3802f, -84.2814f, 50.0000f, 0.0000f, -82.7485f, 1.0000f},
        {50.0000f, -1.1848f, -84.8006f, 50.0000f, 0.0000f, -82.7485f, 1.0000f}, {50.0000f, -0.9009f, -85.5211f, 50.0000f, 0.0000f, -82.7485f, 1.0000f},
        {50.0000f, 0.0000f, 0.0000f, 50.0000f, -1.0000f, 2.0000f, 2.3669f}, {50.0000f, -1.0000f, 2.0000f, 50.0000f, 0.0000f, 0.0000f, 2.3669f},
        {50.0000f, 2.4900f, -0.0010f, 50.0000f, -2.4900f, 0.0009f, 7.1792f}, {50.0000f, 2.4900f, -0.0010f, 50.0000f, -2.4900f, 0.0010f, 7.1792f},
        {50.0000f, 2.4900f, -0.0010f, 50.0000f, -2.4900f, 0.0011f, 7.2195f}, {50.0000f, 2.4900f, -0.0010f, 50.0000f, -2.4900f, 0.0012f, 7.2195f},
        {50.0000f, -0.0010f, 2.4900f, 50.0000f, 0.0009f, -2.4900f, 4.8045f}, {50.0000f, -0.0010f, 2.4900f, 50.0000f, 0.0010f, -2.4900f, 4.8045f},
        {50.0000f, -0.0010f, 2.4900f, 50.0000f, 0.0011f, -2.4900f, 4.7461f}, {50.0000f, 2.5000f, 0.0000f, 50.0000f, 0.0000f, -2.5000f, 4.3065f},
        {50.0000f, 2.5000f, 0.0000f, 73.0000f, 25.0000f, -18.0000f, 27.1492f}, {50.0000f, 2.5000f, 0.0000f, 61.0000f, -5.0000f, 29.0000f, 22.8977f},
        {50.0000f, 2.5000f, 0.0000f, 56.0000f, -27.0000f, -3.0000f, 31.9030f}, {50.0000f, 2.5000f, 0.0000f, 58.0000f, 24.0000f, 15.0000f, 19.4535f},
        {50.0000f, 2.5000f, 0.0000f, 50.0000f, 3.1736f, 0.5854f, 1.0000f}, {50.0000f, 2.5000f, 0.0000f, 50.0000f, 3.2972f, 0.0000f, 1.0000f},
        {50.0000f, 2.5000f, 0.0000f, 50.0000f, 1.8634f, 0.5757f, 1.0000f}, {50.0000f, 2.5000f, 0.0000f, 50.0000f, 3.2592f, 0.3350f, 1.0000f},
        {60.2574f, -34.0099f, 36.2677f, 60.4626f, -34.1751f, 39.4387f, 1.2644f}, {63.0109f, -31.0961f, -5.8663f, 62.8187f, -29.7946f, -4.0864f, 1.2630f},
        {61.2901f, 3.7196f, -5.3901f, 61.4292f, 2.2480f, -4.9620f, 1.8731f}, {35.0831f, -44.1164f, 3.7933f, 35.0232f, -40.0716f, 1.5901f, 1.8645f},
        {22.7233f, 20.0904f, -46.6940f, 23.0331f, 14.9730f, -42.5619f, 2.0373f}, {36.4612f, 47.8580f, 18.3852f, 36.2715f, 50.5065f, 21.2231f, 1.4146f},
        {90.8027f, -2.0831f, 1.4410f, 91.1528f, -1.6435f, 0.0447f, 1.4441f}, {90.9257f, -0.5406f, -0.9208f, 88.6381f, -0.8985f, -0.7239f, 1.5381f},
        {6.7747f, -0.2908f, -2.4247f, 5.8714f, -0.0985f, -2.2286f, 0.6377f}, {2.0776f, 0.0795f, -1.1350f, 0.9033f, -0.0636f, -0.5514f, 0.9082f}};
    float err = 0;
    for (int i = 0; i < 34; ++i) err = std::max(err, ciede00_test(v[i][0], v[i][1], v[i][2], v[i][3], v[i][4], v[i][5], v[i][6]));
    std::printf("CIEDE00 max error: %f\n", err);
}
void ColorUtilities::convert_test() {                                // :465-499: RGB -> Lab -> RGB round trip on four colours, printed
    float cols[4][3] = {{123, 10, 200}, {0, 0, 0}, {255, 255, 255}, {255, 255, 0}};
    for (int i = 0; i < 4; ++i) {
        float* lab = rgb2lab(cols[i]);
        float* back = lab2rgb(lab);
        std::printf("RGB (%g, %g, %g) -> Lab (%f, %f, %f) -> RGB (%f, %f, %f)\n", cols[i][0], cols[i][1], cols[i][2], lab[0], lab[1], lab[2], back[0], back[1], back[2]);
        delete[] lab; delete[] back;
    }
}

// ---------------------------------------------------------------------------------------------------
// Testing (src/testing.cpp): pairs by exact xyz on the host, contingency table + scores through f3ps_eval_label_pairs
void Testing::init_performance() { precision = recall = fscore = voi = wov = fpr = fnr = -1; }
labelMapT Testing::label_map(PointLCloudT::Ptr in) {                 // :62-82: one sub-cloud per label, renumbered 0..L-1 in ascending label order
    std::map<uint32_t, PointLCloudT::Ptr> by_label;
    for (const PointLT& p : in->points) {
        auto it = by_label.find(p.label);
        if (it == by_label.end()) it = by_label.insert(std::make_pair(p.label, PointLCloudT::Ptr(new PointLCloudT()))).first;
        it->second->push_back(p);
    }
    labelMapT out; uint32_t k = 0;
    for (auto& kv : by_label) out[k++] = kv.second;
    return out;
}
void Testing::compute_intersections() {
    // dense labels in ascending label order for both clouds; truth points indexed by xyz (compareXYZ: -0 == +0, as operator<)
    std::map<uint32_t, uint32_t> sdense, tdense;
    for (const PointLT& p : segm->points) sdense[p.label] = 0;
    for (const PointLT& p : truth->points) tdense[p.label] = 0;
    uint32_t ns = 0, nt = 0;
    for (auto& kv : sdense) kv.second = ns++;
    for (auto& kv : tdense) kv.second = nt++;
    std::map<PointLT, uint32_t, compareXYZ> where;
    std::vector<uint64_t> tsizes(nt, 0);
    for (const PointLT& p : truth->points) { const uint32_t j = tdense[p.label]; where.insert(std::make_pair(p, j)); tsizes[j]++; }
    std::vector<uint32_t> sl(segm->size()), tl(segm->size());
    for (size_t i = 0; i < segm->size(); ++i) {
        const PointLT& p = segm->points[i];
        sl[i] = sdense[p.label];
        auto it = where.find(p);
        tl[i] = it == where.end() ? nt : it->second;
    }
    f3ps::Handle& h = f3ps::shared_handle();
    f3ps_performance pf;
    h.check(f3ps_eval_label_pairs(h.get(), sl.data(), tl.data(), (int64_t)sl.size(), (int32_t)ns, (int32_t)nt, tsizes.data(), (int64_t)truth->size(), &pf));
    precision = pf.precision; recall = pf.recall; fscore = pf.fscore; voi = pf.voi; wov = pf.wov; fpr = pf.fpr; fnr = pf.fnr;
}
Testing::Testing(PointLCloudT::Ptr s, PointLCloudT::Ptr t) { is_set_segm = false; is_set_truth = false; init_performance(); set_segm(s); set_truth(t); }
void Testing::set_segm(PointLCloudT::Ptr s) {
    if (s->empty()) throw std::invalid_argument("The pointcloud to be set as 'segm' cannot be empty");
    segm = s; init_performance(); segm_labels = label_map(s); is_set_segm = true;
    if (is_set_truth) compute_intersections();
}
void Testing::set_truth(PointLCloudT::Ptr t) {
    if (t->empty()) throw std::invalid_argument("The pointcloud to be set as 'truth' cannot be empty");
    truth = t; init_performance(); truth_labels = label_map(t); is_set_truth = true;
    if (is_set_segm) compute_intersections();
}
float Testing::eval_precision() { return precision; }
float Testing::eval_recall() { return recall; }
float Testing::eval_fscore() { return fscore; }
float Testing::eval_voi() { return voi; }
float Testing::eval_wov() { return wov; }
float Testing::eval_fpr() { return fpr; }
float Testing::eval_fnr() { return fnr; }
performanceSet Testing::eval_performance() {
    performanceSet p; p.voi = voi; p.precision = precision; p.recall = recall; p.fscore = fscore; p.wov = wov; p.fpr = fpr; p.fnr = fnr;
    return p;
}

// ---------------------------------------------------------------------------------------------------
Clustering::Clustering() : h_(new f3ps::Handle(0)) {
    set_delta_c(LAB_CIEDE00); set_delta_g(NORMALS_DIFF); set_merging(ADAPTIVE_LAMBDA);
    set_initial_state = false; init_initial_weights = false;
}
Clustering::Clustering(ColorDistance c, GeometricDistance g, MergingCriterion m) : h_(new f3ps::Handle(0)) {
    set_delta_c(c); set_delta_g(g); set_merging(m);
    set_initial_state = false; init_initial_weights = false;
}
void Clustering::set_merging(MergingCriterion m) { merging_type = m; lambda = 0.5f; bins_num = 500; init_initial_weights = false; }
void Clustering::set_lambda(float l) {
    if (merging_type != MANUAL_LAMBDA)
        throw std::logic_error("Lambda can be set only if the merging criterion is set to MANUAL_LAMBDA");
    if (l < 0 || l > 1) throw std::invalid_argument("Argument outside range [0, 1]");
    lambda = l; init_initial_weights = false;
}
void Clustering::set_bins_num(short b) {
    if (merging_type != EQUALIZATION)
        throw std::logic_error("Bins number can be set only if the merging criterion is set to EQUALIZATION");
    if (b < 0) throw std::invalid_argument("Argument lower than 0");
    bins_num = b; init_initial_weights = false;
}
void Clustering::push_params() {
    h_->check(f3ps_set_merge_params(h_->get(), (int)delta_c_type, (int)delta_g_type, (int)merging_type, lambda, bins_num));
}

void Clustering::set_initialstate(ClusteringT segm, AdjacencyMapT adj) {
    // flatten the supervoxels (std::map order = ascending label) and hand the graph to the device
    std::vector<uint32_t> labels; std::vector<int64_t> off(1, 0); std::vector<float> xyz, cen, nrm; std::vector<uint32_t> rgba;
    flat_voxels_.clear(); flat_normals_.clear();
    for (auto& kv : segm) {
        labels.push_back(kv.first);
        for (const PointT& p : kv.second->voxels_->points) { xyz.push_back(p.x); xyz.push_back(p.y); xyz.push_back(p.z); rgba.push_back(p.rgba); flat_voxels_.push_back(p); }
        for (size_t i = 0; i < kv.second->voxels_->size(); ++i)
            flat_normals_.push_back(i < kv.second->normals_->size() ? kv.second->normals_->points[i] : Normal());
        off.push_back((int64_t)rgba.size());
        cen.push_back(kv.second->centroid_.x); cen.push_back(kv.second->centroid_.y); cen.push_back(kv.second->centroid_.z);
        nrm.push_back(kv.second->normal_.normal_x); nrm.push_back(kv.second->normal_.normal_y); nrm.push_back(kv.second->normal_.normal_z);
    }
    std::vector<uint32_t> pairs; WeightMapT w0;
    for (auto& e : adj) {
        pairs.push_back(e.first); pairs.push_back(e.second);
        if (e.first <= e.second) w0.insert(WeightedPairT(-1.0f, e));        // clear_adjacency + adj2weight
    }
    push_params();
    h_->check(f3ps_set_graph(h_->get(), (int64_t)rgba.size(), xyz.data(), rgba.data(), (int32_t)labels.size(), labels.data(), off.data(),
                             cen.data(), nrm.data(), (int64_t)(pairs.size() / 2), pairs.data()));
    ClusteringState init_state(segm, w0);
    initial_state = init_state; state = init_state;
    set_initial_state = true; init_initial_weights = false;
    merge_log_.clear();
}

void Clustering::pull_state(bool merged) {
    f3ps_ctx* c = h_->get();
    f3ps_counts n; h_->check(f3ps_get_counts(c, &n));
    // initial weights (init_weights result)
    {
        const size_t E = (size_t)n.n_edges;
        std::vector<uint32_t> ab(2 * E); std::vector<float> w(E);
        h_->check(f3ps_get_edges(c, ab.data(), nullptr, nullptr, w.data(), (int64_t)E));
        WeightMapT wm;
        std::vector<size_t> ord(E);
        for (size_t i = 0; i < E; ++i) ord[i] = i;
        // NaN weights (degenerate regions) order as +inf, as on the device: `<` alone is not a strict weak order with NaNs
        auto wkey = [&](size_t i) { return std::isnan(w[i]) ? std::numeric_limits<float>::infinity() : w[i]; };
        std::stable_sort(ord.begin(), ord.end(), [&](size_t x, size_t y) { return wkey(x) < wkey(y); });
        for (size_t i : ord) wm.insert(wm.end(), WeightedPairT(w[i], std::make_pair(ab[2 * i], ab[2 * i + 1])));
        initial_state.set_weight_map(wm);
        lambda = n.lambda;
    }
    if (!merged) return;
    const size_t K = (size_t)n.n_segments, L = (size_t)n.n_labeled, El = (size_t)n.n_edges_left, M = (size_t)n.n_merges;
    std::vector<uint32_t> lab(K); std::vector<float> cen(3 * K), nrm(3 * K); std::vector<int32_t> cnt(K);
    h_->check(f3ps_get_state_regions(c, lab.data(), cen.data(), nrm.data(), cnt.data(), (int64_t)K));
    std::vector<float> oxyz(3 * L); std::vector<uint32_t> olab(L), ovox(L);
    h_->check(f3ps_get_labeled_cloud(c, oxyz.data(), olab.data(), ovox.data(), (int64_t)L));
    ClusteringT seg;
    size_t pos = 0;
    for (size_t k = 0; k < K; ++k) {
        SupervoxelT::Ptr sv(new SupervoxelT());
        sv->centroid_.x = cen[3 * k]; sv->centroid_.y = cen[3 * k + 1]; sv->centroid_.z = cen[3 * k + 2];
        sv->normal_.normal_x = nrm[3 * k]; sv->normal_.normal_y = nrm[3 * k + 1]; sv->normal_.normal_z = nrm[3 * k + 2];
        for (int i = 0; i < cnt[k]; ++i, ++pos) { sv->voxels_->push_back(flat_voxels_[ovox[pos]]); sv->normals_->push_back(flat_normals_[ovox[pos]]); }
        seg[lab[k]] = sv;
    }
    std::vector<uint32_t> eab(2 * El); std::vector<float> ew(El);
    h_->check(f3ps_get_state_edges(c, eab.data(), ew.data(), (int64_t)El));
    WeightMapT wm;
    for (size_t i = 0; i < El; ++i) wm.insert(wm.end(), WeightedPairT(ew[i], std::make_pair(eab[2 * i], eab[2 * i + 1])));
    state.set_segments(seg); state.set_weight_map(wm);
    std::vector<uint32_t> mab(2 * M), mleft(2 * M); std::vector<float> mw(M);
    h_->check(f3ps_get_merge_log(c, mab.data(), mw.data(), mleft.data(), (int64_t)M));
    merge_log_.resize(M);
    for (size_t m = 0; m < M; ++m) merge_log_[m] = MergeStep{mab[2 * m], mab[2 * m + 1], mw[m], mleft[2 * m], mleft[2 * m + 1]};
}

void Clustering::cluster(float threshold) {
    if (!set_initial_state)
        throw std::logic_error("Cannot call 'cluster' before setting an initial state with 'set_initialstate'");
    if (!init_initial_weights) { push_params(); h_->check(f3ps_graph(h_->get())); init_initial_weights = true; }
    h_->check(f3ps_merge(h_->get(), threshold));
    pull_state(true);
}

std::pair<ClusteringT, AdjacencyMapT> Clustering::get_currentstate() const {
    std::pair<ClusteringT, AdjacencyMapT> ret;
    ret.first = state.segments;
    for (const auto& e : state.weight_map) ret.second.insert(e.second);     // weight2adj
    return ret;
}

// ---- Clustering::all_thresh / best_thresh (src/clustering.cpp:691-774) ---------------------------------------------
std::map<float, performanceSet> Clustering::all_thresh(PointLCloudT::Ptr ground_truth, float start_thresh, float end_thresh, float step_thresh) {
    if (start_thresh < 0 || start_thresh > 1 || end_thresh < 0 || end_thresh > 1 || step_thresh < 0 || step_thresh > 1)
        throw std::out_of_range("start_thresh, end_thresh and/or step_thresh outside of range [0, 1]");
    if (start_thresh > end_thresh) std::swap(start_thresh, end_thresh);            // "inverting" (:699-704)
    if (!set_initial_state)
        throw std::logic_error("Cannot call 'cluster' before setting an initial state with 'set_initialstate'");
    if (!ground_truth || ground_truth->empty())
        throw std::invalid_argument("The pointcloud to be set as 'truth' cannot be empty");     // Testing::set_truth (testing.cpp:430-433)
    if (!init_initial_weights) { push_params(); h_->check(f3ps_graph(h_->get())); init_initial_weights = true; }
    std::vector<float> thr(1, start_thresh);
    for (float t = start_thresh + step_thresh; t <= end_thresh; t += step_thresh) thr.push_back(t);
    // Testing::count_intersect matches points by exact xyz (compareXYZ): ground-truth label of every voxel of the graph
    struct Key { uint32_t k[3]; bool operator<(const Key& o) const { return std::lexicographical_compare(k, k + 3, o.k, o.k + 3); } };
    auto key_of = [](float x, float y, float z) { Key q; float v[3] = {x == 0 ? 0.0f : x, y == 0 ? 0.0f : y, z == 0 ? 0.0f : z}; memcpy(q.k, v, 12); return q; };
    std::map<Key, size_t> where;
    for (size_t i = 0; i < ground_truth->size(); ++i) { const PointLT& p = ground_truth->points[i]; where.insert(std::make_pair(key_of(p.x, p.y, p.z), i)); }
    std::vector<char> used(ground_truth->size(), 0);
    std::vector<uint32_t> truth(flat_voxels_.size(), 0xffffffffu), extra;
    for (size_t v = 0; v < flat_voxels_.size(); ++v) {
        auto it = where.find(key_of(flat_voxels_[v].x, flat_voxels_[v].y, flat_voxels_[v].z));
        if (it != where.end()) { truth[v] = ground_truth->points[it->second].label; used[it->second] = 1; }
    }
    for (size_t i = 0; i < ground_truth->size(); ++i) if (!used[i]) extra.push_back(ground_truth->points[i].label);
    std::vector<f3ps_performance> perf(thr.size());
    h_->check(f3ps_eval_thresholds(h_->get(), truth.data(), (int64_t)truth.size(), extra.empty() ? nullptr : extra.data(), (int64_t)extra.size(),
                                   thr.data(), (int)thr.size(), perf.data(), nullptr, nullptr));
    pull_state(true);                                                               // `state` = clustering at the last threshold
    std::map<float, performanceSet> out;
    for (size_t k = 0; k < thr.size(); ++k) {
        performanceSet p;
        p.voi = perf[k].voi; p.precision = perf[k].precision; p.recall = perf[k].recall; p.fscore = perf[k].fscore;
        p.wov = perf[k].wov; p.fpr = perf[k].fpr; p.fnr = perf[k].fnr;
        out.insert(std::make_pair(thr[k], p));
    }
    return out;
}

std::pair<float, performanceSet> Clustering::best_thresh(PointLCloudT::Ptr ground_truth, float start_thresh, float end_thresh, float step_thresh) {
    return best_thresh(all_thresh(ground_truth, start_thresh, end_thresh, step_thresh));
}

std::pair<float, performanceSet> Clustering::best_thresh(std::map<float, performanceSet> all_thresh) {
    float best_t = 0;
    performanceSet best_performance;
    for (auto it = all_thresh.begin(); it != all_thresh.end(); ++it)
        if (it->second.fscore > best_performance.fscore) { best_performance = it->second; best_t = it->first; }
    return std::pair<float, performanceSet>(best_t, best_performance);
}

PointLCloudT::Ptr Clustering::get_labeled_cloud() const {
    PointLCloudT::Ptr out(new PointLCloudT());
    uint32_t dense = 0;
    for (const auto& kv : state.segments) {
        for (const PointT& p : kv.second->voxels_->points) { PointLT q; q.x = p.x; q.y = p.y; q.z = p.z; q.label = dense; out->push_back(q); }
        ++dense;
    }
    return out;
}
PointCloudT::Ptr Clustering::get_colored_cloud() const { return label2color(get_labeled_cloud()); }
void Clustering::test_all() const { ColorUtilities::rgb_test(); ColorUtilities::lab_test(); ColorUtilities::convert_test(); }

PointCloudT::Ptr Clustering::label2color(PointLCloudT::Ptr label_cloud) {
    PointCloudT::Ptr out(new PointCloudT());
    out->resize(label_cloud->size());
    for (size_t i = 0; i < label_cloud->size(); ++i) {
        const PointLT& p = label_cloud->points[i];
        uint8_t* c = ColorUtilities::get_glasbey(p.label);
        PointT& q = out->points[i];
        q.x = p.x; q.y = p.y; q.z = p.z; q.rgba = 0; q.r = c[0]; q.g = c[1]; q.b = c[2];
        delete[] c;
    }
    return out;
}
PointLCloudT::Ptr Clustering::color2label(PointCloudT::Ptr colored_cloud) {
    PointLCloudT::Ptr out(new PointLCloudT());
    std::map<uint32_t, uint32_t> seen;
    out->resize(colored_cloud->size());
    for (size_t i = 0; i < colored_cloud->size(); ++i) {
        const PointT& p = colored_cloud->points[i];
        const uint32_t key = p.rgba & 0x00ffffffu;
        auto it = seen.find(key);
        if (it == seen.end()) it = seen.insert(std::make_pair(key, (uint32_t)seen.size())).first;
        PointLT& q = out->points[i];
        q.x = p.x; q.y = p.y; q.z = p.z; q.label = it->second;
    }
    return out;
}
