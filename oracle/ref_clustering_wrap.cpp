// ref_clustering_wrap.cpp -- plain-C entry to the REFERENCE's own Clustering class (/root/reference/src/clustering.cpp,
// clustering_state.cpp, color_utilities.cpp compiled where they lie against oracle/ref_shim/): set_initialstate + cluster(threshold)
// on caller-supplied supervoxels (what main() does at /root/reference/src/supervoxel_clustering.cpp:408-443), returning the
// reference's per-merge debug lines (:390-392), the remaining edges, the regions and the labelled cloud.
// Test infrastructure: built into oracle/_ref/libref_clustering.so by oracle/Makefile when /root/reference exists.
#include <cstdint>
#include <exception>

#include "ref_capture.h"
#include "supervoxel_clustering/clustering.h"

namespace {
ClusteringT build_segments(int32_t n_sv, const uint32_t* labels, const int64_t* vox_off, const float* vox_xyz, const uint32_t* vox_rgba,
                           const float* centroid_xyz, const float* normal_xyz) {
    ClusteringT segm;
    for (int32_t s = 0; s < n_sv; ++s) {
        SupervoxelT::Ptr sv = boost::make_shared<SupervoxelT>();
        for (int64_t v = vox_off[s]; v < vox_off[s + 1]; ++v) {
            PointT p; p.x = vox_xyz[3 * v]; p.y = vox_xyz[3 * v + 1]; p.z = vox_xyz[3 * v + 2]; p.rgba = vox_rgba[v];
            sv->voxels_->push_back(p);
        }
        sv->centroid_.x = centroid_xyz[3 * s]; sv->centroid_.y = centroid_xyz[3 * s + 1]; sv->centroid_.z = centroid_xyz[3 * s + 2];
        sv->normal_.normal_x = normal_xyz[3 * s]; sv->normal_.normal_y = normal_xyz[3 * s + 1]; sv->normal_.normal_z = normal_xyz[3 * s + 2];
        segm.insert(std::make_pair(labels[s], sv));
    }
    return segm;
}
}  // namespace

// Clustering::all_thresh + best_thresh (/root/reference/src/clustering.cpp:691-774) as main() uses them for the automatic threshold
// (/root/reference/src/supervoxel_clustering.cpp:428-438): truth_label[v] = ground-truth label of voxel v (same order as vox_xyz).
// out_t[k], out_perf[k][7] = (voi, precision, recall, fscore, wov, fpr, fnr) per threshold in map order; best[8] = threshold + its scores.
extern "C" int ref_all_thresh(const int16_t* lab_lut, int32_t n_sv, const uint32_t* labels, const int64_t* vox_off, const float* vox_xyz,
                              const uint32_t* vox_rgba, const float* centroid_xyz, const float* normal_xyz, int64_t n_adj, const uint32_t* adj_pairs,
                              int color, int geom, int merging, float lambda, int bins, const uint32_t* truth_label,
                              float start, float end, float step, int32_t cap, float* out_t, float* out_perf, int32_t* n_out, float* best) {
    try {
        f3ps_ref::lab_lut = lab_lut;
        ClusteringT segm = build_segments(n_sv, labels, vox_off, vox_xyz, vox_rgba, centroid_xyz, normal_xyz);
        AdjacencyMapT adj;
        for (int64_t k = 0; k < n_adj; ++k) adj.insert(std::make_pair(adj_pairs[2 * k], adj_pairs[2 * k + 1]));
        PointLCloudT::Ptr truth = boost::make_shared<PointLCloudT>();
        for (int64_t v = 0; v < vox_off[n_sv]; ++v) { PointLT p; p.x = vox_xyz[3 * v]; p.y = vox_xyz[3 * v + 1]; p.z = vox_xyz[3 * v + 2]; p.label = truth_label[v]; truth->push_back(p); }
        Clustering c((ColorDistance)color, (GeometricDistance)geom, (MergingCriterion)merging);
        if (merging == MANUAL_LAMBDA) c.set_lambda(lambda);
        if (merging == EQUALIZATION) c.set_bins_num((short)bins);
        c.set_initialstate(segm, adj);
        const std::map<float, performanceSet> all = c.all_thresh(truth, start, end, step);
        *n_out = (int32_t)all.size();
        int32_t k = 0;
        for (auto& kv : all) {
            if (k < cap) {
                out_t[k] = kv.first; const performanceSet& p = kv.second;
                float* o = out_perf + 7 * k; o[0] = p.voi; o[1] = p.precision; o[2] = p.recall; o[3] = p.fscore; o[4] = p.wov; o[5] = p.fpr; o[6] = p.fnr;
            }
            ++k;
        }
        const std::pair<float, performanceSet> b = c.best_thresh(all);
        best[0] = b.first; best[1] = b.second.voi; best[2] = b.second.precision; best[3] = b.second.recall; best[4] = b.second.fscore;
        best[5] = b.second.wov; best[6] = b.second.fpr; best[7] = b.second.fnr;
        return 0;
    } catch (const std::exception&) { return 1; }
}

extern "C" int ref_cluster(const int16_t* lab_lut, int32_t n_sv, const uint32_t* labels, const int64_t* vox_off, const float* vox_xyz,
                           const uint32_t* vox_rgba, const float* centroid_xyz, const float* normal_xyz, int64_t n_adj, const uint32_t* adj_pairs,
                           int color, int geom, int merging, float lambda, int bins, float threshold,
                           int64_t cap_merges, uint32_t* m_ab, float* m_w, uint32_t* m_left, int64_t* n_merges,
                           int64_t cap_edges, uint32_t* f_ab, int64_t* n_edges,
                           uint32_t* r_label, int32_t* r_size, float* r_centroid, float* r_normal4, int32_t* n_regions,
                           int64_t cap_points, uint32_t* out_label, float* out_xyz, int64_t* n_points, float* lambda_out) {
    try {
        f3ps_ref::lab_lut = lab_lut;
        ClusteringT segm = build_segments(n_sv, labels, vox_off, vox_xyz, vox_rgba, centroid_xyz, normal_xyz);
        AdjacencyMapT adj;
        for (int64_t k = 0; k < n_adj; ++k) adj.insert(std::make_pair(adj_pairs[2 * k], adj_pairs[2 * k + 1]));
        Clustering c((ColorDistance)color, (GeometricDistance)geom, (MergingCriterion)merging);
        if (merging == MANUAL_LAMBDA) c.set_lambda(lambda);
        if (merging == EQUALIZATION) c.set_bins_num((short)bins);
        c.set_initialstate(segm, adj);
        std::vector<f3ps_ref::MergeLine> log;
        f3ps_ref::sink = &log;
        c.cluster(threshold);
        f3ps_ref::sink = nullptr;
        *n_merges = (int64_t)log.size();
        for (int64_t m = 0; m < (int64_t)log.size() && m < cap_merges; ++m) {
            m_ab[2 * m] = log[m].a; m_ab[2 * m + 1] = log[m].b; m_w[m] = log[m].w; m_left[2 * m] = log[m].edges_left; m_left[2 * m + 1] = log[m].regions_left;
        }
        const std::pair<ClusteringT, AdjacencyMapT> st = c.get_currentstate();
        *n_edges = (int64_t)st.second.size();
        int64_t e = 0;
        for (auto& kv : st.second) { if (e < cap_edges) { f_ab[2 * e] = kv.first; f_ab[2 * e + 1] = kv.second; } ++e; }
        *n_regions = (int32_t)st.first.size();
        int32_t r = 0;
        for (auto& kv : st.first) {
            r_label[r] = kv.first; r_size[r] = (int32_t)kv.second->voxels_->size();
            r_centroid[3 * r] = kv.second->centroid_.x; r_centroid[3 * r + 1] = kv.second->centroid_.y; r_centroid[3 * r + 2] = kv.second->centroid_.z;
            r_normal4[4 * r] = kv.second->normal_.normal_x; r_normal4[4 * r + 1] = kv.second->normal_.normal_y; r_normal4[4 * r + 2] = kv.second->normal_.normal_z;
            r_normal4[4 * r + 3] = kv.second->normal_.curvature;
            ++r;
        }
        PointLCloudT::Ptr lc = c.get_labeled_cloud();
        *n_points = (int64_t)lc->size();
        for (int64_t i = 0; i < (int64_t)lc->size() && i < cap_points; ++i) {
            out_label[i] = lc->points[(size_t)i].label; out_xyz[3 * i] = lc->points[(size_t)i].x; out_xyz[3 * i + 1] = lc->points[(size_t)i].y; out_xyz[3 * i + 2] = lc->points[(size_t)i].z;
        }
        *lambda_out = c.get_lambda();
        return 0;
    } catch (const std::exception&) { f3ps_ref::sink = nullptr; return 1; }
}

// ColorUtilities::lab_ciede00 / rgb_eucl (/root/reference/src/color_utilities.cpp:190-319) on n pairs; the reference returns plain floats
extern "C" void ref_color_distances(int64_t n, const float* c1, const float* c2, float* ciede00, float* eucl) {
    for (int64_t i = 0; i < n; ++i) {
        float a[3] = {c1[3 * i], c1[3 * i + 1], c1[3 * i + 2]}, b[3] = {c2[3 * i], c2[3 * i + 1], c2[3 * i + 2]};
        if (ciede00) ciede00[i] = ColorUtilities::lab_ciede00(a, b);
        if (eucl) eucl[i] = ColorUtilities::rgb_eucl(a, b);
    }
}

// The reference's exception behaviour, case by case (the facade and the C ABI's status codes mirror it): returns 0 = no exception,
// 1 = std::logic_error, 2 = std::invalid_argument, 3 = std::out_of_range, 4 = anything else.
extern "C" int ref_exception_case(int which) {
    try {
        PointLCloudT::Ptr one = boost::make_shared<PointLCloudT>(), none = boost::make_shared<PointLCloudT>();
        PointLT p; p.x = p.y = p.z = 0; p.label = 0; one->push_back(p);
        switch (which) {
            case 0: { Clustering c(LAB_CIEDE00, NORMALS_DIFF, MANUAL_LAMBDA); c.set_lambda(1.5f); break; }          // clustering.cpp:578-579
            case 1: { Clustering c(LAB_CIEDE00, NORMALS_DIFF, ADAPTIVE_LAMBDA); c.set_lambda(0.5f); break; }        // :575-577
            case 2: { Clustering c(LAB_CIEDE00, NORMALS_DIFF, EQUALIZATION); c.set_bins_num(-3); break; }           // :593-594
            case 3: { Clustering c(LAB_CIEDE00, NORMALS_DIFF, ADAPTIVE_LAMBDA); c.set_bins_num(10); break; }        // :590-592
            case 4: { Clustering c; c.cluster(0.2f); break; }                                                       // :671-673
            case 5: { Clustering c; c.all_thresh(one, 1.5f, 1.0f, 0.005f); break; }                                 // :694-698
            case 6: { Testing t(none, one); break; }                                                                // testing.cpp:420-423
            case 7: { Testing t(one, none); break; }                                                                // :430-433
            case 8: { Clustering c(LAB_CIEDE00, NORMALS_DIFF, MANUAL_LAMBDA); c.set_lambda(1.0f); c.set_lambda(0.0f); break; }   // the closed range is accepted
            default: return 4;
        }
        return 0;
    } catch (const std::out_of_range&) { return 3; }
    catch (const std::invalid_argument&) { return 2; }
    catch (const std::logic_error&) { return 1; }
    catch (...) { return 4; }
}
