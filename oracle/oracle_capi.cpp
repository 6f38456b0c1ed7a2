// oracle_capi.cpp -- C entry points over the CPU ORACLE for ctypes (tests/, smoke(),
// bench.py cpu_baseline only).  TEST INFRASTRUCTURE ONLY (see oracle.h).
#include "oracle.h"

#include <cstring>
#include <string>

using namespace f3ps_oracle;

namespace {
struct Handle {
    Oracle O;
    std::vector<int16_t> lut;
    std::string err;
    std::vector<uint32_t> edges_ab, merges_ab, merges_left, final_ab;
    std::vector<float> edges_dc, edges_dg, edges_w, merges_w, final_w;
    std::vector<double> scalars;
    std::vector<int> steals;
};
}

extern "C" {

void* orc_create(const int16_t* lab_lut) {
    Handle* h = new Handle();
    h->lut.assign(lab_lut, lab_lut + 33 * 33 * 33 * 3);
    h->O.lab_lut = h->lut.data();
    return h;
}
void orc_destroy(void* hv) { delete (Handle*)hv; }
const char* orc_last_error(void* hv) { return ((Handle*)hv)->err.c_str(); }

void orc_set_vccs_params(void* hv, float rv, float rs, float wc, float ws, float wn, int use_transform, int fold_negative_z) {
    Params& P = ((Handle*)hv)->O.P;
    P.voxel_res = rv; P.seed_res = rs; P.color_imp = wc; P.spatial_imp = ws; P.normal_imp = wn;
    P.use_transform = use_transform; P.fold_negative_z = fold_negative_z;
}
void orc_set_merge_params(void* hv, int color_mode, int geom_mode, int merge_mode, float lambda, int bins, int merge_impl) {
    Params& P = ((Handle*)hv)->O.P;
    P.color_mode = color_mode; P.geom_mode = geom_mode; P.merge_mode = merge_mode; P.lambda = lambda; P.bins = bins;
    P.merge_impl = merge_impl;
}
void orc_set_expand_impl(void* hv, int impl) { ((Handle*)hv)->O.P.expand_impl = impl; }
void orc_set_switches(void* hv, int leaf_desc, int keybits_floor, int init_seed_voxel, int shifted_cov) {
    Switches& s = ((Handle*)hv)->O.P.sw;
    s.leaf_order_descending = leaf_desc; s.keybits_floor = keybits_floor;
    s.init_centroid_seed_voxel = init_seed_voxel; s.shifted_covariance = shifted_cov;
}
void orc_set_input(void* hv, const void* pts, long n, int stride) { ((Handle*)hv)->O.set_input((const uint8_t*)pts, n, stride); }

// stage: 0 all, 1..6 = K1..K6a (+set_initialstate/init_weights with 6), 7 = cluster(thr) only, 8 = K1..K6 without merge
int orc_run(void* hv, int stage, float thr) {
    Handle* h = (Handle*)hv; Oracle& O = h->O;
    try {
        switch (stage) {
            case 0: O.run_all(thr); break;
            case 1: O.voxelize(); break;
            case 2: O.neighbors(); break;
            case 3: O.voxel_normals(); break;
            case 4: O.select_seeds(); break;
            case 5: O.expand(); break;
            case 6: O.make_supervoxels(); O.set_initialstate(); O.init_weights(); break;
            case 7: O.cluster(thr); O.labeled_cloud(); break;
            case 8: O.voxelize(); O.neighbors(); O.voxel_normals(); O.select_seeds(); O.expand();
                    O.make_supervoxels(); O.set_initialstate(); O.init_weights(); break;
            default: h->err = "bad stage"; return 1;
        }
    } catch (const std::exception& e) { h->err = e.what(); return 2; }
    return 0;
}

// refineSupervoxels(num_itr) on the result of stage 5 (K5); stage 6 afterwards rebuilds the graph from the refined supervoxels
int orc_refine(void* hv, int num_itr) {
    Handle* h = (Handle*)hv;
    try { h->O.refine(num_itr); } catch (const std::exception& e) { h->err = e.what(); return 2; }
    return 0;
}

// Inject a graph directly (Clustering facade tests): regions given as voxel lists over
// caller-provided voxel arrays.
int orc_set_graph(void* hv, long V, const float* vxyz, const uint32_t* vrgba,
                  long S, const uint32_t* labels, const long* vox_off, const int* vox_idx,
                  const float* centroids, const float* normals, long n_adj, const uint32_t* adj_pairs) {
    Handle* h = (Handle*)hv; Oracle& O = h->O;
    O.vxyz.assign(vxyz, vxyz + 3 * V); O.vrgba.assign(vrgba, vrgba + V);
    O.initial_segments.clear();
    for (long s = 0; s < S; ++s) {
        Region r;
        r.voxels.assign(vox_idx + vox_off[s], vox_idx + vox_off[s + 1]);
        r.cx = centroids[3 * s]; r.cy = centroids[3 * s + 1]; r.cz = centroids[3 * s + 2];
        r.nx = normals[3 * s]; r.ny = normals[3 * s + 1]; r.nz = normals[3 * s + 2];
        O.initial_segments[labels[s]] = r;
    }
    O.adj.assign(adj_pairs, adj_pairs + 2 * n_adj);
    try { O.set_initialstate(); O.init_weights(); } catch (const std::exception& e) { h->err = e.what(); return 2; }
    return 0;
}

// named array access: returns element count, *ptr points at oracle-owned storage
long orc_array(void* hv, const char* name, const void** ptr) {
    Handle* h = (Handle*)hv; Oracle& O = h->O;
    std::string n(name);
#define RET(vec) do { *ptr = (vec).data(); return (long)(vec).size(); } while (0)
    if (n == "keys") RET(O.keys);
    if (n == "morton") RET(O.morton);
    if (n == "voxel_xyz") RET(O.vxyz);
    if (n == "voxel_rgb") RET(O.vrgb);
    if (n == "voxel_rgba") RET(O.vrgba);
    if (n == "voxel_count") RET(O.vcount);
    if (n == "point_voxel") RET(O.point_voxel);
    if (n == "nbr") RET(O.nbr);
    if (n == "nbr_count") RET(O.nbr_count);
    if (n == "normals") RET(O.normals);
    if (n == "curvature") RET(O.curvature);
    if (n == "seed_cells_nn") RET(O.seed_cells_nn);
    if (n == "seeds") RET(O.seeds);
    if (n == "labels") RET(O.labels);
    if (n == "dist") RET(O.dist);
    if (n == "steals") { h->steals = O.steals_per_round; RET(h->steals); }
    if (n == "sv_label") RET(O.sv_label);
    if (n == "sv_xyz") RET(O.sv_xyz);
    if (n == "sv_rgb") RET(O.sv_rgb);
    if (n == "sv_normal") RET(O.sv_normal);
    if (n == "sv_count") RET(O.sv_count);
    if (n == "adj") RET(O.adj);
    if (n == "cdf_c") RET(O.cdf_c);
    if (n == "cdf_g") RET(O.cdf_g);
    if (n == "out_xyz") RET(O.out_xyz);
    if (n == "out_label") RET(O.out_label);
    if (n == "out_voxel") RET(O.out_voxel);
    if (n == "edges_ab") { h->edges_ab.clear(); for (auto& e : O.edges) { h->edges_ab.push_back(e.a); h->edges_ab.push_back(e.b); } RET(h->edges_ab); }
    if (n == "edges_dc") { h->edges_dc.clear(); for (auto& e : O.edges) h->edges_dc.push_back(e.dc); RET(h->edges_dc); }
    if (n == "edges_dg") { h->edges_dg.clear(); for (auto& e : O.edges) h->edges_dg.push_back(e.dg); RET(h->edges_dg); }
    if (n == "edges_w") { h->edges_w.clear(); for (auto& e : O.edges) h->edges_w.push_back(e.w); RET(h->edges_w); }
    if (n == "merges_ab") { h->merges_ab.clear(); for (auto& m : O.merges) { h->merges_ab.push_back(m.a); h->merges_ab.push_back(m.b); } RET(h->merges_ab); }
    if (n == "merges_w") { h->merges_w.clear(); for (auto& m : O.merges) h->merges_w.push_back(m.w); RET(h->merges_w); }
    if (n == "merges_left") { h->merges_left.clear(); for (auto& m : O.merges) { h->merges_left.push_back(m.edges_left); h->merges_left.push_back(m.regions_left); } RET(h->merges_left); }
    if (n == "final_ab") { h->final_ab.clear(); for (auto& e : O.final_edges) { h->final_ab.push_back(e.a); h->final_ab.push_back(e.b); } RET(h->final_ab); }
    if (n == "final_w") { h->final_w.clear(); for (auto& e : O.final_edges) h->final_w.push_back(e.w); RET(h->final_w); }
    if (n == "stage_ms") { *ptr = O.stage_ms; return 8; }
    if (n == "scalars") {
        h->scalars = {(double)O.depth, O.bmin[0], O.bmin[1], O.bmin[2], O.bmax[0], O.bmax[1], O.bmax[2],
                      (double)O.seed_depth, O.seed_min[0], O.seed_min[1], O.seed_min[2], (double)O.rounds,
                      (double)O.lambda_used, (double)O.nan_weights, (double)O.segments.size()};
        RET(h->scalars);
    }
#undef RET
    *ptr = nullptr; return -1;
}

// metric kernels for known-answer tests
void orc_rgb2lab(void* hv, const float* rgb255, float* lab) { rgb2lab(((Handle*)hv)->O.lab_lut, rgb255, lab); }
float orc_lab_ciede00(const float* lab1, const float* lab2) { return lab_ciede00(lab1, lab2); }
void orc_rgb2lab_batch(void* hv, const float* rgb255, float* lab, long n) { for (long i = 0; i < n; ++i) rgb2lab(((Handle*)hv)->O.lab_lut, rgb255 + 3 * i, lab + 3 * i); }
void orc_lab_ciede00_batch(const float* lab1, const float* lab2, float* out, long n) { for (long i = 0; i < n; ++i) out[i] = lab_ciede00(lab1 + 3 * i, lab2 + 3 * i); }
float orc_rgb_eucl(const float* a, const float* b) { return rgb_eucl(a, b); }
float orc_normals_diff(const float* n1, const float* c1, const float* n2, const float* c2) { return normals_diff(n1, c1, n2, c2); }
int orc_is_convex(const float* n1, const float* c1, const float* n2, const float* c2) { return is_convex(n1, c1, n2, c2) ? 1 : 0; }
void orc_plane_from_accu(const float* accu9, int n, float* normal4, float* curv) { plane_from_accu(accu9, n, normal4, curv); }
float orc_cr_logf(float x) { return cr_logf(x); }

} // extern "C"
