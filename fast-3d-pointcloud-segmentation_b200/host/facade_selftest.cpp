// facade_selftest <file.pcd> [threshold]: runs one cloud through (a) the fused C-ABI path and (b) the
// reference-shaped classes (SupervoxelClustering -> Clustering), and checks that both give the same merge
// sequence and labelled cloud, plus the reference's error behaviour and colour known answers.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "pcd_io.h"
#include "supervoxel_clustering/clustering.h"

#define CHECK(cond, msg) do { if (!(cond)) { fprintf(stderr, "FACADE FAIL: %s (%s:%d)\n", msg, __FILE__, __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s file.pcd [threshold]\n", argv[0]); return 2; }
    const float thr = argc > 2 ? (float)atof(argv[2]) : 0.2f;
    pcl::PointCloud<pcl::PointXYZRGBL> input;
    CHECK(f3ps::loadPCDFile(argv[1], input) == 0, "cannot read the PCD file");
    for (auto& p : input.points) if (p.z < 0) p.z = std::abs(p.z);
    pcl::PointCloud<pcl::PointXYZRGBA>::Ptr cloud(new pcl::PointCloud<pcl::PointXYZRGBA>());
    pcl::copyPointCloud(input, *cloud);

    // error behaviour of the reference (src/clustering.cpp:574-597, 670-673)
    {
        Clustering c;
        bool threw = false;
        try { c.cluster(0.2f); } catch (const std::logic_error&) { threw = true; }
        CHECK(threw, "cluster before set_initialstate must throw std::logic_error");
        threw = false;
        try { c.set_lambda(0.3f); } catch (const std::logic_error&) { threw = true; }
        CHECK(threw, "set_lambda under ADAPTIVE_LAMBDA must throw std::logic_error");
        c.set_merging(MANUAL_LAMBDA);
        threw = false;
        try { c.set_lambda(1.5f); } catch (const std::invalid_argument&) { threw = true; }
        CHECK(threw, "set_lambda(1.5) must throw std::invalid_argument");
        c.set_merging(EQUALIZATION);
        CHECK(c.get_bins_num() == 500 && c.get_lambda() == 0.5f, "set_merging resets lambda=0.5, bins=500");
        threw = false;
        try { c.set_bins_num(-1); } catch (const std::invalid_argument&) { threw = true; }
        CHECK(threw, "set_bins_num(-1) must throw std::invalid_argument");
    }
    {   // ColorUtilities with the reference's signatures (float* results are new[]-allocated, as there)
        float c1[3] = {0, 0, 0}, c2[3] = {255, 255, 255}, c3[3] = {100, 20, 35}, c4[3] = {104, 20, 32};
        CHECK(ColorUtilities::rgb_eucl(c1, c2) == 441.672943f && ColorUtilities::rgb_eucl(c3, c4) == 5.0f, "rgb_eucl known answers");
        float l1[3] = {50.0000f, 2.6772f, -79.7751f}, l2[3] = {50.0000f, 0.0000f, -82.7485f};
        CHECK(std::fabs(ColorUtilities::lab_ciede00(l1, l2) - 2.0425f) < 1e-4f, "CIEDE2000 known answer");
        CHECK(std::fabs(ColorUtilities::lab_ciede00(l1, l2, 2.0, 1.0, 1.0) - ColorUtilities::lab_ciede00(l1, l2)) < 1e-4f, "kL only scales the (zero) lightness term here");
        float rgb[3] = {123, 10, 200};
        float* lab = ColorUtilities::rgb2lab(rgb);
        CHECK(lab[0] == 35.113525390625f && lab[1] == 69.984375f && lab[2] == -71.28125f, "rgb2lab = OpenCV's LUT value (SURVEY Appendix F)");
        float* back = ColorUtilities::lab2rgb(lab);
        CHECK(std::fabs(back[0] - 123) < 1.0f && std::fabs(back[1] - 10) < 1.0f && std::fabs(back[2] - 200) < 1.0f, "lab2rgb round trip");
        delete[] lab; delete[] back;
        uint8_t* g0 = ColorUtilities::get_glasbey(7); uint8_t* g1 = ColorUtilities::get_glasbey(7 + 256);
        CHECK(g0[0] == g1[0] && g0[1] == g1[1] && g0[2] == g1[2], "glasbey wraps at 256");
        delete[] g0; delete[] g1;
        Clustering().test_all();                                      // prints, as the reference's does
    }

    // (a) fused path
    f3ps::Handle h(0);
    h.check(f3ps_set_vccs_params(h.get(), 0.008f, 0.08f, 0.2f, 0.4f, 1.0f, 1, 0));
    h.check(f3ps_set_merge_params(h.get(), F3PS_LAB_CIEDE00, F3PS_CONVEX_NORMALS_DIFF, F3PS_ADAPTIVE_LAMBDA, 0.5f, 500));
    h.check(f3ps_set_input(h.get(), cloud->points.data(), (int64_t)cloud->size(), 32, 0));
    h.check(f3ps_run(h.get(), thr));
    f3ps_counts n; h.check(f3ps_get_counts(h.get(), &n));
    std::vector<uint32_t> ab(2 * (size_t)n.n_merges), left(2 * (size_t)n.n_merges); std::vector<float> w(n.n_merges);
    h.check(f3ps_get_merge_log(h.get(), ab.data(), w.data(), left.data(), n.n_merges));
    std::vector<float> xyz(3 * (size_t)n.n_labeled); std::vector<uint32_t> lab(n.n_labeled), vox(n.n_labeled);
    h.check(f3ps_get_labeled_cloud(h.get(), xyz.data(), lab.data(), vox.data(), n.n_labeled));

    // (b) the reference's call sequence (src/supervoxel_clustering.cpp:348-367, 408-449)
    pcl::SupervoxelClustering<pcl::PointXYZRGBA> super(0.008f, 0.08f);
    super.setUseSingleCameraTransform(true);
    super.setInputCloud(cloud);
    super.setColorImportance(0.2f); super.setSpatialImportance(0.4f); super.setNormalImportance(1.0f);
    std::map<uint32_t, pcl::Supervoxel<pcl::PointXYZRGBA>::Ptr> supervoxel_clusters;
    super.extract(supervoxel_clusters);
    std::multimap<uint32_t, uint32_t> label_adjacency;
    super.getSupervoxelAdjacency(label_adjacency);
    CHECK((int)supervoxel_clusters.size() == n.n_supervoxels, "supervoxel count");
    CHECK((int)label_adjacency.size() == 2 * n.n_edges, "adjacency size");
    CHECK((int64_t)super.getVoxelCentroidCloud()->size() == n.n_voxels, "voxel centroid cloud size");
    CHECK(super.getLabeledCloud()->size() == cloud->size(), "labelled input cloud size");
    Clustering segmentation;
    segmentation.set_delta_g(CONVEX_NORMALS_DIFF);
    segmentation.set_initialstate(supervoxel_clusters, label_adjacency);
    CHECK(segmentation.get_lambda() == 0.5f, "lambda reads 0.5 before the weights are initialised (quirk D.6)");
    segmentation.cluster(thr);
    const std::vector<MergeStep>& log = segmentation.get_merge_log();
    CHECK((int)log.size() == n.n_merges, "merge count differs between the fused path and the class path");
    for (size_t m = 0; m < log.size(); ++m)
        CHECK(log[m].a == ab[2 * m] && log[m].b == ab[2 * m + 1] && log[m].w == w[m] && log[m].edges_left == left[2 * m], "merge sequence differs");
    CHECK(std::fabs(segmentation.get_lambda() - n.lambda) == 0.0f, "adaptive lambda differs");
    PointLCloudT::Ptr labeled = segmentation.get_labeled_cloud();
    CHECK((int)labeled->size() == n.n_labeled, "labelled cloud size");
    for (size_t i = 0; i < labeled->size(); ++i)
        CHECK(labeled->points[i].label == lab[i] && labeled->points[i].x == xyz[3 * i] && labeled->points[i].z == xyz[3 * i + 2], "labelled cloud differs");
    std::pair<ClusteringT, AdjacencyMapT> st = segmentation.get_currentstate();
    CHECK((int)st.first.size() == n.n_segments && (int)st.second.size() == n.n_edges_left, "current state size");
    // restart from the initial state with a lower threshold: a prefix of the same sequence
    segmentation.cluster(thr * 0.5f);
    const std::vector<MergeStep>& log2 = segmentation.get_merge_log();
    CHECK(log2.size() <= (size_t)n.n_merges, "prefix length");
    for (size_t m = 0; m < log2.size(); ++m) CHECK(log2[m].a == ab[2 * m] && log2[m].b == ab[2 * m + 1], "prefix differs");
    // threshold sweep: Clustering::all_thresh on a ground-truth voxel cloud == f3ps_eval_thresholds on the fused handle
    {
        PointLCloudT::Ptr truth(new PointLCloudT());
        pcl::PointCloud<pcl::PointXYZRGBA>::Ptr vc = super.getVoxelCentroidCloud();
        std::vector<uint32_t> tl(vc->size());
        for (size_t v = 0; v < vc->size(); ++v) {                       // three bands along x as a stand-in ground truth
            PointLT p; p.x = vc->points[v].x; p.y = vc->points[v].y; p.z = vc->points[v].z;
            p.label = tl[v] = vc->points[v].x < -0.3f ? 4u : (vc->points[v].x < 0.4f ? 9u : 2u);
            truth->push_back(p);
        }
        std::map<float, performanceSet> all = segmentation.all_thresh(truth, 0.05f, 0.5f, 0.05f);
        std::vector<float> tv; for (auto& kv : all) tv.push_back(kv.first);
        std::vector<f3ps_performance> perf(tv.size());
        h.check(f3ps_eval_thresholds(h.get(), tl.data(), (int64_t)tl.size(), nullptr, 0, tv.data(), (int)tv.size(), perf.data(), nullptr, nullptr));
        size_t k = 0;
        for (auto& kv : all) {
            CHECK(kv.second.fscore == perf[k].fscore && kv.second.voi == perf[k].voi && kv.second.wov == perf[k].wov && kv.second.fnr == perf[k].fnr,
                  "all_thresh differs between the class path and the fused path");
            ++k;
        }
        std::pair<float, performanceSet> best = segmentation.best_thresh(all);
        for (auto& kv : all) CHECK(kv.second.fscore <= best.second.fscore, "best_thresh is not the maximum F-score");
        CHECK(best.second.fscore > 0.0f && all.size() == tv.size() && tv.size() >= 9, "sweep size / positive F-score");
        bool threw = false;
        try { segmentation.all_thresh(truth, 0.5f, 1.5f, 0.1f); } catch (const std::out_of_range&) { threw = true; }
        CHECK(threw, "all_thresh outside [0,1] must throw std::out_of_range (src/clustering.cpp:694-698)");
    }
    {   // refineSupervoxels + getSupervoxelAdjacencyList (src/supervoxel_clustering.cpp:367-384) == f3ps_refine on the fused handle
        pcl::SupervoxelClustering<pcl::PointXYZRGBA> sup2(0.008f, 0.08f);
        sup2.setUseSingleCameraTransform(true); sup2.setInputCloud(cloud);
        sup2.setColorImportance(0.2f); sup2.setSpatialImportance(0.4f); sup2.setNormalImportance(1.0f);
        std::map<uint32_t, pcl::Supervoxel<pcl::PointXYZRGBA>::Ptr> first, refined;
        sup2.extract(first);
        f3ps::VoxelAdjacencyList gl; sup2.getSupervoxelAdjacencyList(gl);
        CHECK(gl.vertices.size() == first.size() && (int)gl.edges.size() == n.n_edges, "adjacency list size");
        for (auto& e : gl.edges) CHECK(e.first.first < e.first.second && e.second > 0.0f && gl.vertices.count(e.first.first) && gl.vertices.count(e.first.second), "adjacency list edge");
        sup2.refineSupervoxels(3, refined);
        f3ps::Handle h2(0);
        h2.check(f3ps_set_vccs_params(h2.get(), 0.008f, 0.08f, 0.2f, 0.4f, 1.0f, 1, 0));
        h2.check(f3ps_set_input(h2.get(), cloud->points.data(), (int64_t)cloud->points.size(), (int)sizeof(pcl::PointXYZRGBA), 0));
        h2.check(f3ps_extract(h2.get())); h2.check(f3ps_refine(h2.get(), 3)); h2.check(f3ps_graph(h2.get()));
        f3ps_counts n2; h2.check(f3ps_get_counts(h2.get(), &n2));
        CHECK((int)refined.size() == n2.n_supervoxels && refined.size() > 0 && refined.size() <= first.size(), "refined supervoxel count");
        std::vector<uint32_t> l2((size_t)n2.n_supervoxels); std::vector<float> c2(3 * (size_t)n2.n_supervoxels); std::vector<int32_t> k2((size_t)n2.n_supervoxels);
        h2.check(f3ps_get_supervoxels(h2.get(), l2.data(), c2.data(), nullptr, nullptr, k2.data(), (int64_t)n2.n_supervoxels));
        size_t s2 = 0;
        for (auto& kv : refined) {
            CHECK(kv.first == l2[s2] && kv.second->centroid_.x == c2[3 * s2] && kv.second->centroid_.z == c2[3 * s2 + 2] &&
                  (int)kv.second->voxels_->size() == k2[s2], "refined supervoxel differs from f3ps_refine");
            ++s2;
        }
        std::multimap<uint32_t, uint32_t> adj2; sup2.getSupervoxelAdjacency(adj2);
        CHECK((int)adj2.size() == 2 * n2.n_edges, "adjacency after refinement");
        pcl::SupervoxelClustering<pcl::PointXYZRGBA> sup3(0.008f, 0.08f);
        std::map<uint32_t, pcl::Supervoxel<pcl::PointXYZRGBA>::Ptr> none;
        sup3.refineSupervoxels(3, none);                                 // before extract(): PCL_FATAL + return, nothing happens
        CHECK(none.empty(), "refine before extract");
    }
    {   // Testing(segm, truth) == the fused sweep's score at the same threshold; label2color / color2label round trip
        segmentation.cluster(thr);
        PointLCloudT::Ptr seg = segmentation.get_labeled_cloud();
        PointLCloudT::Ptr truth(new PointLCloudT());
        pcl::PointCloud<pcl::PointXYZRGBA>::Ptr vc = super.getVoxelCentroidCloud();
        std::vector<uint32_t> tl(vc->size());
        for (size_t v = 0; v < vc->size(); ++v) {
            PointLT p; p.x = vc->points[v].x; p.y = vc->points[v].y; p.z = vc->points[v].z;
            p.label = tl[v] = vc->points[v].x < -0.3f ? 4u : (vc->points[v].x < 0.4f ? 9u : 2u);
            truth->push_back(p);
        }
        Testing test(seg, truth);
        performanceSet ps = test.eval_performance();
        f3ps_performance pf; const float t1 = thr;
        h.check(f3ps_eval_thresholds(h.get(), tl.data(), (int64_t)tl.size(), nullptr, 0, &t1, 1, &pf, nullptr, nullptr));
        CHECK(ps.fscore == pf.fscore && ps.voi == pf.voi && ps.wov == pf.wov && ps.precision == pf.precision && ps.recall == pf.recall &&
              ps.fpr == pf.fpr && ps.fnr == pf.fnr, "Testing differs from the sweep's scores");
        CHECK(test.eval_fscore() == ps.fscore && test.eval_recall() == ps.recall, "eval_* getters");
        bool threw = false;
        try { Testing bad(PointLCloudT::Ptr(new PointLCloudT()), truth); } catch (const std::invalid_argument&) { threw = true; }
        CHECK(threw, "an empty segmentation must throw std::invalid_argument (src/testing.cpp:413-416)");
        PointLCloudT::Ptr back = Clustering::color2label(Clustering::label2color(seg));
        std::map<uint32_t, uint32_t> fwd;
        bool consistent = back->size() == seg->size();
        for (size_t i = 0; consistent && i < seg->size(); ++i) {
            auto it = fwd.find(seg->points[i].label);
            if (it == fwd.end()) fwd[seg->points[i].label] = back->points[i].label;
            else consistent = it->second == back->points[i].label;
        }
        std::map<uint32_t, uint32_t> inv; for (auto& kv : fwd) inv[kv.second] = kv.first;
        CHECK(consistent && (inv.size() == fwd.size() || fwd.size() > 256), "label2color / color2label is not a bijection on the labels");
    }
    printf("FACADE OK: N=%zu V=%lld S=%d E=%d M=%d segments=%d\n", cloud->size(), (long long)n.n_voxels, n.n_supervoxels, n.n_edges, n.n_merges, n.n_segments);
    return 0;
}
