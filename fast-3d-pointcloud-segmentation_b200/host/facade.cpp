// facade.cpp -- host side of the drop-in: f3ps::Handle, f3ps::SupervoxelClustering (pcl::SupervoxelClustering
// as the reference consumes it), Clustering, ClusteringState, ColorUtilities -- all thin layers over the C ABI
// (include/f3ps.h).  Nothing here computes on the CPU what the reference computes in its hot path; the host
// only marshals PCL-shaped containers in and out.
#include <algorithm>
#include <cmath>
#include <cstdio>
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
        sv->normal_.curvature = 0.0f;
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
std::array<uint8_t, 3> ColorUtilities::get_glasbey(uint32_t label) {
    // deterministic distinct-colour palette (golden-angle hue walk); pcl::GlasbeyLUT's table is not reproduced
    const uint32_t k = label % 256u;
    const float h = std::fmod(k * 0.61803398875f, 1.0f), s = 0.55f + 0.45f * ((k * 7u) % 5u) / 4.0f, v = 0.6f + 0.4f * ((k * 3u) % 4u) / 3.0f;
    const float c = v * s, hp = h * 6.0f, x = c * (1.0f - std::fabs(std::fmod(hp, 2.0f) - 1.0f)), m = v - c;
    float r = 0, g = 0, b = 0;
    if (hp < 1) { r = c; g = x; } else if (hp < 2) { r = x; g = c; } else if (hp < 3) { g = c; b = x; }
    else if (hp < 4) { g = x; b = c; } else if (hp < 5) { r = x; b = c; } else { r = c; b = x; }
    return {(uint8_t)((r + m) * 255.0f), (uint8_t)((g + m) * 255.0f), (uint8_t)((b + m) * 255.0f)};
}
std::array<float, 3> ColorUtilities::mean_color(SupervoxelT::Ptr s) {
    // one supervoxel as a one-node graph: the device computes the running mean exactly as the merge loop does
    f3ps::Handle& h = f3ps::shared_handle();
    const size_t n = s->voxels_->size();
    std::vector<float> xyz(3 * n); std::vector<uint32_t> rgba(n);
    for (size_t i = 0; i < n; ++i) { const PointT& p = s->voxels_->points[i]; xyz[3 * i] = p.x; xyz[3 * i + 1] = p.y; xyz[3 * i + 2] = p.z; rgba[i] = p.rgba; }
    const uint32_t label = 1; const int64_t off[2] = {0, (int64_t)n}; const float cen[3] = {0, 0, 0}, nrm[3] = {0, 0, 1};
    h.check(f3ps_set_graph(h.get(), (int64_t)n, xyz.data(), rgba.data(), 1, &label, off, cen, nrm, 0, nullptr));
    float rgb[3] = {0, 0, 0};
    h.check(f3ps_get_region_mean_color(h.get(), 0, rgb));
    return {rgb[0], rgb[1], rgb[2]};
}
std::array<float, 3> ColorUtilities::rgb2lab(const float rgb[3]) {
    float lab[3]; f3ps::Handle& h = f3ps::shared_handle();
    h.check(f3ps_test_rgb2lab(h.get(), rgb, lab, 1));
    return {lab[0], lab[1], lab[2]};
}
float ColorUtilities::lab_ciede00(const float lab1[3], const float lab2[3]) {
    float d; f3ps::Handle& h = f3ps::shared_handle();
    h.check(f3ps_test_lab_ciede00(h.get(), lab1, lab2, &d, 1));
    return d;
}
float ColorUtilities::rgb_eucl(const float rgb1[3], const float rgb2[3]) {
    float d; f3ps::Handle& h = f3ps::shared_handle();
    h.check(f3ps_test_rgb_eucl(h.get(), rgb1, rgb2, &d, 1));
    return d;
}
float ColorUtilities::rgb_test() {
    const float c[8][3] = {{0, 0, 0}, {255, 255, 255}, {255, 0, 0}, {0, 255, 0}, {0, 255, 0}, {255, 0, 255}, {100, 20, 35}, {104, 20, 32}};
    const int pairs[7][2] = {{0, 0}, {0, 1}, {1, 1}, {0, 2}, {3, 0}, {4, 5}, {6, 7}};
    const float expect[7] = {0, 441.672943f, 0, 255, 255, 441.672943f, 5};
    float err = 0;
    for (int i = 0; i < 7; ++i) err = std::max(err, std::fabs(rgb_eucl(c[pairs[i][0]], c[pairs[i][1]]) - expect[i]));
    return err;
}
float ColorUtilities::lab_test() {
    // a few rows of the Sharma-Wu-Dalal table (the full table is exercised by tests/)
    const float v[4][7] = {{50.0000f, 2.6772f, -79.7751f, 50.0000f, 0.0000f, -82.7485f, 2.0425f},
                           {50.0000f, 2.5000f, 0.0000f, 73.0000f, 25.0000f, -18.0000f, 27.1492f},
                           {60.2574f, -34.0099f, 36.2677f, 60.4626f, -34.1751f, 39.4387f, 1.2644f},
                           {2.0776f, 0.0795f, -1.1350f, 0.9033f, -0.0636f, -0.5514f, 0.9082f}};
    float err = 0;
    for (int i = 0; i < 4; ++i) err = std::max(err, std::fabs(lab_ciede00(v[i], v[i] + 3) - v[i][6]));
    return err;
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
        std::stable_sort(ord.begin(), ord.end(), [&](size_t x, size_t y) { return w[x] < w[y]; });
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

PointCloudT::Ptr Clustering::label2color(PointLCloudT::Ptr label_cloud) {
    PointCloudT::Ptr out(new PointCloudT());
    out->resize(label_cloud->size());
    for (size_t i = 0; i < label_cloud->size(); ++i) {
        const PointLT& p = label_cloud->points[i];
        const std::array<uint8_t, 3> c = ColorUtilities::get_glasbey(p.label);
        PointT& q = out->points[i];
        q.x = p.x; q.y = p.y; q.z = p.z; q.rgba = 0; q.r = c[0]; q.g = c[1]; q.b = c[2];
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
