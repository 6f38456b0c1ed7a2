// oracle.h -- CPU ORACLE for the supervoxel-plus-merging hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build, load or call it.  The product library (libf3ps.so) never links
// or falls back to this code.
//
// What it is: a single-threaded C++ restatement of the reference's algorithm,
//   back half  (in /root/reference, fully pinned by source):
//       src/clustering.cpp:53-162,172-346,356-528,605-679
//       src/color_utilities.cpp:52-69,117-319
//       include/supervoxel_clustering/clustering_state.h:47-123
//   front half (PCL 1.10 semantics, third-party, NOT in /root/reference):
//       call sites src/supervoxel_clustering.cpp:315-337,348-367
//       behaviour restated from SURVEY.md Appendix A.
//
// PARITY STATUS
//   * CIEDE2000 (lab_ciede00) and rgb_eucl are pinned by the reference's own
//     known-answer vectors (src/color_utilities.cpp:324-460), see tests/.
//   * RGB->Lab is pinned against cv2 4.13.0 in this image (tools/gen_lab_lut.py,
//     tests/golden/lab_kat.json); the reference does not pin OpenCV's version.
//   * The whole back half -- set_initialstate, init_weights, adaptive lambda, CDFs,
//     mean_color, the multimap merge loop and its order, get_labeled_cloud, and the
//     evaluation scores -- is pinned by the reference's OWN classes: clustering.cpp,
//     clustering_state.cpp, color_utilities.cpp and testing.cpp compile where they lie
//     against the stand-ins of oracle/ref_shim/ (Makefile target `ref` -> oracle/_ref/);
//     their merge sequences / scores are committed as tests/golden/clustering_ref.npz and
//     testing_ref.json and this oracle reproduces them bit for bit
//     (tests/test_oracle_reference_clustering.py, tests/test_oracle_testing.py).
//   * The VCCS front end (PCL) is **parity unpinned**: PCL is absent from this
//     image and from /root/reference, the reference has no test or fixture for
//     it, so every PCL/Eigen detail below is restated from knowledge of PCL
//     1.10 / Eigen 3.3 and carries a named switch where it is version dependent.
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off, no fast-math).
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace f3ps_oracle {

// ---- switches for version-dependent third-party behaviour ----
struct Switches {
    int leaf_order_descending = 0;   // PCL <=1.8.1 visits octree children 7->0
    int keybits_floor = 0;           // PCL 1.8.1: floor((max-min)/res) instead of ceil(.. - eps)
    int init_centroid_seed_voxel = 0;// 0 = literal (helper centroid starts at zero)
    int shifted_covariance = 0;      // PCL >=1.11 subtracts the first sample
};

struct Params {
    float voxel_res = 0.008f;        // -v  src/supervoxel_clustering.cpp:247
    float seed_res = 0.08f;          // -s  :252
    float color_imp = 0.2f;          // -c  :257
    float spatial_imp = 0.4f;        // -z  :261
    float normal_imp = 1.0f;         // -n  :265
    int use_transform = 1;           // !--NT :192,349
    int fold_negative_z = 1;         // main's clean-up :317-321
    int color_mode = 0;              // 0 LAB_CIEDE00, 1 RGB_EUCL   clustering.h:62
    int geom_mode = 0;               // 0 NORMALS_DIFF, 1 CONVEX_NORMALS_DIFF :66
    int merge_mode = 1;              // 0 MANUAL_LAMBDA, 1 ADAPTIVE_LAMBDA, 2 EQUALIZATION :70
    float lambda = 0.5f;             // clustering.cpp:564
    int bins = 500;                  // clustering.cpp:565
    int merge_impl = 0;              // 0 literal std::multimap replay, 1 stamp/prefix "fixed" variant
    int expand_impl = 0;             // 0 literal sequential helpers, 1 data-parallel fixed-point formulation (what the GPU runs)
    Switches sw;
};

struct Region {                      // pcl::Supervoxel as consumed by Clustering
    std::vector<int> voxels;         // voxels_ in order (indices into the voxel arrays)
    float cx = 0, cy = 0, cz = 0;    // centroid_
    float nx = 0, ny = 0, nz = 0;    // normal_
    float curvature = 0;
};

struct Edge { uint32_t a, b; float dc, dg, w; };
struct MergeRec { uint32_t a, b; float w; uint32_t edges_left, regions_left; };

struct Oracle {
    Params P;
    const int16_t* lab_lut = nullptr;     // 33*33*33*3, loaded by the caller

    // ---- input after main()'s clean-up ----
    std::vector<float> px, py, pz;
    std::vector<uint32_t> prgba;

    // ---- K1: voxels ----
    int depth = 0;
    double bmin[3] = {0, 0, 0}, bmax[3] = {0, 0, 0};
    std::vector<uint32_t> keys;           // 3V (x,y,z)
    std::vector<uint64_t> morton;         // V
    std::vector<float> vxyz, vrgb;        // 3V each (float means)
    std::vector<uint32_t> vrgba;          // V truncated colours (centroid cloud)
    std::vector<int> vcount;              // V
    std::vector<int> point_voxel;         // N -> voxel (-1 if skipped)
    // ---- K2 ----
    std::vector<int> nbr;                 // 27V, list order, -1 padded
    std::vector<int> nbr_count;           // V
    // ---- K3 ----
    std::vector<float> normals;           // 4V
    std::vector<float> curvature;         // V
    // ---- K4 ----
    int seed_depth = 0;
    double seed_min[3] = {0, 0, 0};
    std::vector<int> seed_cells_nn;       // nearest voxel per occupied seed cell (ordered)
    std::vector<int> seeds;               // kept seed voxel indices (label = i+1)
    // ---- K5 ----
    int rounds = 0;
    std::vector<uint32_t> labels;         // V, 0 = unowned
    std::vector<float> dist;              // V stored distance_
    std::vector<int> steals_per_round;
    // ---- K6a ----
    std::vector<uint32_t> sv_label;       // S surviving labels ascending
    std::vector<float> sv_xyz, sv_rgb, sv_normal; // 3S,3S,4S helper centroids
    std::vector<int> sv_count;
    std::vector<std::vector<int>> sv_leaves; // per surviving helper: its leaf set in idx order (may hold a phantom leaf)
    std::vector<uint32_t> adj;            // 2*(2E) directed pairs, sorted
    // ---- K6b / K7 ----
    std::map<uint32_t, Region> initial_segments;
    std::vector<Edge> edges;              // initial edges, lexicographic (a<b)
    float lambda_used = 0.5f;
    std::vector<float> cdf_c, cdf_g;
    std::vector<MergeRec> merges;
    std::map<uint32_t, Region> segments;  // current state
    std::vector<Edge> final_edges;        // current weight map in order
    long nan_weights = 0;
    std::vector<float> out_xyz;           // labelled voxel cloud (get_labeled_cloud)
    std::vector<uint32_t> out_label;
    std::vector<uint32_t> out_voxel;      // voxel index of each output point

    double stage_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0}; // K1..K7, [7]=total

    // pipeline
    void set_input(const uint8_t* pts, long n, int stride);
    void voxelize();            // K1  (A.1, A.2)
    void neighbors();           // K2  (A.2)
    void voxel_normals();       // K3  (A.3)
    void select_seeds();        // K4  (A.4)
    void expand();              // K5  (A.5)
    void expand_fixed_point();  // K5, restated as the per-voxel fold + steal-table fixed point
    int sweeps_total = 0;
    void refine(int num_itr);   // refineSupervoxels (supervoxel_clustering.cpp:369-371) on the result of expand()
    void make_supervoxels();    // K6a (A.6)
    void set_initialstate();    // clustering.cpp:605-612 on the VCCS output
    void init_weights();        // clustering.cpp:212-251
    void cluster(float thr);    // clustering.cpp:670-679 (always restarts from initial_state)
    void labeled_cloud();       // clustering.cpp:640-663
    void run_all(float thr);
};

// ---- metric kernels exposed for known-answer tests ----
void rgb2lab(const int16_t* lut, const float rgb255[3], float lab[3]);     // color_utilities.cpp:151-160 + cv::cvtColor
void rgb_unit2lab(const int16_t* lut, const float unit[3], float lab[3]);  // cv::cvtColor(COLOR_RGB2Lab) alone, channels in [0, 1]
float lab_ciede00(const float lab1[3], const float lab2[3]);               // color_utilities.cpp:190-294
float rgb_eucl(const float rgb1[3], const float rgb2[3]);                  // color_utilities.cpp:304-319
float normals_diff(const float n1[3], const float c1[3], const float n2[3], const float c2[3]); // clustering.cpp:79-96
bool is_convex(const float n1[3], const float c1[3], const float n2[3], const float c2[3]);     // clustering.cpp:53-67
// PCL computePointNormal pieces (A.3): covariance accumulators -> normal, curvature
void plane_from_accu(const float accu9_sum[9], int n, float normal4[4], float* curvature);
void eigen33_smallest(const float cov[6], float* eigenvalue, float evec[3]);

// correctly-rounded float libm model used wherever the reference calls float libm
float cr_logf(float x);
float cr_atan2f(float y, float x);
float cr_cosf(float x);
float cr_sinf(float x);

} // namespace f3ps_oracle
