// oracle_merge.cpp -- CPU ORACLE, back half: ColorUtilities + Clustering/ClusteringState
// as written in /root/reference (file:line cited per function).  TEST INFRASTRUCTURE
// ONLY (see oracle.h).
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <limits>
#include <set>
#include <stdexcept>
#include <tuple>
#include <unordered_map>

namespace f3ps_oracle {

double now_ms();

static const float RGB_RANGE = 441.672943;   // include/supervoxel_clustering/color_utilities.h:62
static const float LAB_RANGE = 137.3607;     // :63

static inline float sum3(float a0, float a1, float a2) { return a0 + (a1 + a2); }          // Eigen 3-vector redux
static inline float sum4(float a0, float a1, float a2, float a3) { return (a0 + a1) + (a2 + a3); }

// ---------------------------------------------------------------------------------
// ColorUtilities::rgb2lab, src/color_utilities.cpp:151-160 -> color_conversion :52-69
// -> cv::cvtColor(CV_32FC3, COLOR_RGB2Lab): OpenCV 4 evaluates it with a 33^3 fixed
// point lattice + integer trilinear interpolation (pinned
// bit-exact against cv2 4.13.0 by tests/test_oracle_color.py).
void rgb2lab(const int16_t* lut, const float rgb[3], float lab[3]) {
    const float unit[3] = {rgb[0] / 255, rgb[1] / 255, rgb[2] / 255};   // :153-155
    rgb_unit2lab(lut, unit, lab);
}
// cv::cvtColor(CV_32FC3, COLOR_RGB2Lab) on one pixel with channels in [0, 1] (what color_conversion hands it, :57-60)
void rgb_unit2lab(const int16_t* lut, const float unit[3], float lab[3]) {
    int c[3];
    for (int k = 0; k < 3; ++k) {
        float v = unit[k];
        v = std::min(std::max(v, 0.0f), 1.0f);
        c[k] = (int)std::nearbyint(v * 16384.0f);            // round-half-even
    }
    int t[3], f[3];
    for (int k = 0; k < 3; ++k) { t[k] = c[k] >> 9; f[k] = (c[k] >> 5) & 15; }
    int out[3] = {0, 0, 0};
    for (int dr = 0; dr < 2; ++dr) for (int dg = 0; dg < 2; ++dg) for (int db = 0; db < 2; ++db) {
        int w = (dr ? f[0] : 16 - f[0]) * (dg ? f[1] : 16 - f[1]) * (db ? f[2] : 16 - f[2]);
        int ir = std::min(t[0] + dr, 32), ig = std::min(t[1] + dg, 32), ib = std::min(t[2] + db, 32);
        const int16_t* e = lut + ((ir * 33 + ig) * 33 + ib) * 3;
        out[0] += w * e[0]; out[1] += w * e[1]; out[2] += w * e[2];
    }
    for (int k = 0; k < 3; ++k) out[k] = (out[k] + 2048) >> 12;
    lab[0] = ((float)out[0] / 16384.0f) * 100.0f;
    lab[1] = ((float)out[1] / 16384.0f) * 256.0f - 128.0f;
    lab[2] = ((float)out[2] / 16384.0f) * 256.0f - 128.0f;
}

// ColorUtilities::lab_ciede00, src/color_utilities.cpp:190-294 (kL=kC=kH=1).  The
// float/double mix is kept expression by expression.
float lab_ciede00(const float lab1[3], const float lab2[3]) {
    const double kL = 1.0, kC = 1.0, kH = 1.0;
    float L1 = lab1[0], a1 = lab1[1], b1 = lab1[2];
    float L2 = lab2[0], a2 = lab2[1], b2 = lab2[2];
    double Cab1 = std::sqrt(a1 * a1 + b1 * b1);              // float sqrt  :200
    double Cab2 = std::sqrt(a2 * a2 + b2 * b2);
    double Cab = (Cab1 + Cab2) / 2.0;
    double G = 0.5 * (1.0 - std::sqrt(std::pow(Cab, 7.0) / (std::pow(Cab, 7.0) + std::pow(25.0, 7.0))));
    double ap1 = (1.0 + G) * a1;
    double ap2 = (1.0 + G) * a2;
    double Cp1 = std::sqrt(ap1 * ap1 + b1 * b1);             // b1*b1 is a float product  :211
    double Cp2 = std::sqrt(ap2 * ap2 + b2 * b2);
    double Cp_prod = (Cp2 * Cp1);
    double hp1 = 0;
    if ((std::abs(ap1) + std::abs(b1)) != 0.0) {
        hp1 = std::atan2((double)b1, ap1);
        if (hp1 < 0) hp1 += 2.0 * M_PI;
    }
    double hp2 = 0;
    if ((std::abs(ap2) + std::abs(b2)) != 0.0) {
        hp2 = std::atan2((double)b2, ap2);
        if (hp2 < 0) hp2 += 2.0 * M_PI;
    }
    double dL = (L2 - L1);                                   // float subtraction :233
    double dC = (Cp2 - Cp1);
    double dhp = (hp2 - hp1);
    if (dhp > M_PI) dhp -= 2.0 * M_PI;
    else if (dhp < -M_PI) dhp += 2.0 * M_PI;
    if (Cp_prod == 0.0) dhp = 0.0;
    double dH = 2.0 * std::sqrt(Cp_prod) * std::sin(dhp / 2.0);
    double Lp = (L2 + L1) / 2.0;                             // float addition :254
    double Cp = (Cp1 + Cp2) / 2.0;
    double hp = (hp1 + hp2) / 2.0;
    if (std::abs(hp1 - hp2) > M_PI) hp -= M_PI;
    if (hp < 0) hp += 2.0 * M_PI;
    if (Cp_prod == 0.0) hp = hp1 + hp2;
    double Lpm502 = (Lp - 50.0) * (Lp - 50.0);
    double T = 1.0 - 0.17 * std::cos(hp - M_PI / 6.0) + 0.24 * std::cos(2.0 * hp)
             + 0.32 * std::cos(3.0 * hp + M_PI / 30.0) - 0.20 * std::cos(4.0 * hp - 63.0 * M_PI / 180.0);
    double dheta_rad = (30.0 * M_PI / 180.0) * std::exp(-std::pow(((180.0 / M_PI * hp - 275.0) / 25.0), 2.0));
    double Rc = 2.0 * std::sqrt(std::pow(Cp, 7.0) / (std::pow(Cp, 7.0) + std::pow(25.0, 7.0)));
    double kLSL = kL * (1.0 + 0.015 * Lpm502 / std::sqrt(20.0 + Lpm502));
    double kLSC = kC * (1.0 + 0.045 * Cp);
    double kHSH = kH * (1.0 + 0.015 * Cp * T);
    double RT = -std::sin(2.0 * dheta_rad) * Rc;
    float delta_e = std::sqrt(std::pow((dL / kLSL), 2.0) + std::pow((dC / kLSC), 2.0)
                              + std::pow((dH / kHSH), 2.0) + RT * (dC / kLSC) * (dH / kHSH));
    return delta_e;
}

// ColorUtilities::rgb_eucl, src/color_utilities.cpp:304-319.  std::pow(float,int)
// promotes to double; the square of a float is exact in double, so storing it to a
// float equals the rounded float product.
float rgb_eucl(const float rgb1[3], const float rgb2[3]) {
    float rd = (float)std::pow((double)(rgb1[0] - rgb2[0]), 2.0);
    float gd = (float)std::pow((double)(rgb1[1] - rgb2[1]), 2.0);
    float bd = (float)std::pow((double)(rgb1[2] - rgb2[2]), 2.0);
    return std::sqrt(rd + gd + bd);
}

// Clustering::normals_diff, src/clustering.cpp:79-96 (Eigen 3-vector evaluation order)
float normals_diff(const float n1[3], const float c1[3], const float n2[3], const float c2[3]) {
    float C[3] = {c1[0] - c2[0], c1[1] - c2[1], c1[2] - c2[2]};
    float nrm = std::sqrt(sum3(C[0] * C[0], C[1] * C[1], C[2] * C[2]));
    C[0] /= nrm; C[1] /= nrm; C[2] /= nrm;
    float x[3] = {n1[1] * n2[2] - n1[2] * n2[1], n1[2] * n2[0] - n1[0] * n2[2], n1[0] * n2[1] - n1[1] * n2[0]};
    float N1xN2 = std::sqrt(sum3(x[0] * x[0], x[1] * x[1], x[2] * x[2]));
    float N1_C = std::abs(sum3(n1[0] * C[0], n1[1] * C[1], n1[2] * C[2]));
    float N2_C = std::abs(sum3(n2[0] * C[0], n2[1] * C[1], n2[2] * C[2]));
    return (N1xN2 + N1_C + N2_C) / 3;
}

// Clustering::is_convex, src/clustering.cpp:53-67
bool is_convex(const float n1[3], const float c1[3], const float n2[3], const float c2[3]) {
    float C[3] = {c1[0] - c2[0], c1[1] - c2[1], c1[2] - c2[2]};
    float nrm = std::sqrt(sum3(C[0] * C[0], C[1] * C[1], C[2] * C[2]));
    C[0] /= nrm; C[1] /= nrm; C[2] /= nrm;
    float cos1 = sum3(n1[0] * C[0], n1[1] * C[1], n1[2] * C[2]);
    float cos2 = sum3(n2[0] * C[0], n2[1] * C[1], n2[2] * C[2]);
    return cos1 >= cos2;
}

namespace {

// ColorUtilities::mean_color, src/color_utilities.cpp:117-142: running mean over the
// truncated uint8 colours of voxels_ in order.
struct MeanState { float count = 0, r = 0, g = 0, b = 0; };
inline void mean_step(MeanState& m, uint32_t rgba) {
    float r = (float)((rgba >> 16) & 255u), g = (float)((rgba >> 8) & 255u), b = (float)(rgba & 255u);
    m.count++;
    m.r = m.r + (1 / m.count) * (r - m.r);
    m.g = m.g + (1 / m.count) * (g - m.g);
    m.b = m.b + (1 / m.count) * (b - m.b);
}

struct Merger {
    Oracle& O;
    explicit Merger(Oracle& o) : O(o) {}

    void mean_color(const Region& s, float rgb[3]) const {
        MeanState m;
        for (int v : s.voxels) mean_step(m, O.vrgba[v]);
        rgb[0] = m.r; rgb[1] = m.g; rgb[2] = m.b;
    }
    // Clustering::delta_c_g, src/clustering.cpp:107-142, given the two mean colours
    std::pair<float, float> delta_c_g_rgb(const float rgb1[3], const float rgb2[3], const Region& s1, const Region& s2) const {
        float delta_c = 0;
        if (O.P.color_mode == 0) {
            float lab1[3], lab2[3];
            rgb2lab(O.lab_lut, rgb1, lab1); rgb2lab(O.lab_lut, rgb2, lab2);
            delta_c = lab_ciede00(lab1, lab2);
            delta_c /= LAB_RANGE;
        } else {
            delta_c = rgb_eucl(rgb1, rgb2);
            delta_c /= RGB_RANGE;
        }
        const float n1[3] = {s1.nx, s1.ny, s1.nz}, n2[3] = {s2.nx, s2.ny, s2.nz};
        const float c1[3] = {s1.cx, s1.cy, s1.cz}, c2[3] = {s2.cx, s2.cy, s2.cz};
        float delta_g = normals_diff(n1, c1, n2, c2);
        if (O.P.geom_mode == 1 && is_convex(n1, c1, n2, c2)) delta_g *= 0.5;
        return {delta_c, delta_g};
    }
    std::pair<float, float> delta_c_g(const Region& s1, const Region& s2) const {
        float rgb1[3], rgb2[3];
        mean_color(s1, rgb1); mean_color(s2, rgb2);
        return delta_c_g_rgb(rgb1, rgb2, s1, s2);
    }
    // Clustering::t_c / t_g, src/clustering.cpp:324-376
    float t_c(float delta_c) const {
        if (O.P.merge_mode == 2) {
            short bin = (short)std::floor(delta_c * O.P.bins);
            if (bin == O.P.bins) bin--;
            return O.cdf_c.at(bin) / 2;
        }
        return O.lambda_used * delta_c;
    }
    float t_g(float delta_g) const {
        if (O.P.merge_mode == 2) {
            short bin = (short)std::floor(delta_g * O.P.bins);
            return O.cdf_g.at(bin) / 2;                      // no clamp (quirk)
        }
        return (1 - O.lambda_used) * delta_g;
    }
    float delta(const Region& s1, const Region& s2) const {
        auto d = delta_c_g(s1, s2);
        return t_c(d.first) + t_g(d.second);
    }
};

// Clustering::deltas_mean, src/clustering.cpp:515-528 (ascending multiset order)
float deltas_mean(const std::multiset<float>& d) {
    float count = 0, mean_d = 0;
    for (float delta : d) { count++; mean_d = mean_d + (1 / count) * (delta - mean_d); }
    return mean_d;
}
// Clustering::compute_cdf, src/clustering.cpp:289-314
std::vector<float> compute_cdf(const std::multiset<float>& dist, int bins_num) {
    std::vector<int> bins(bins_num, 0);
    int n = (int)dist.size();
    for (float d : dist) {
        short bin = (short)std::floor(d * bins_num);
        if (bin == bins_num) bin--;
        bins.at(bin)++;
    }
    std::vector<float> cdf(bins_num);
    for (int i = 0; i < bins_num; ++i) {
        float v = 0;
        for (int j = 0; j <= i; ++j) v += bins[j];
        v /= n;
        cdf[i] = v;
    }
    return cdf;
}

// merged-region geometry, src/clustering.cpp:411-424 + SURVEY.md A.7
void region_geometry(Oracle& O, Region& r) {
    float sx = 0, sy = 0, sz = 0;
    for (int v : r.voxels) { sx += O.vxyz[3 * v]; sy += O.vxyz[3 * v + 1]; sz += O.vxyz[3 * v + 2]; }
    float n = (float)r.voxels.size();
    r.cx = sx / n; r.cy = sy / n; r.cz = sz / n;            // CentroidPoint
    float n4[4]; float curv;
    if (r.voxels.size() < 3) {
        n4[0] = n4[1] = n4[2] = n4[3] = std::numeric_limits<float>::quiet_NaN(); curv = n4[0];
    } else {
        float accu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int v : r.voxels) {
            float x = O.vxyz[3 * v], y = O.vxyz[3 * v + 1], z = O.vxyz[3 * v + 2];
            accu[0] += x * x; accu[1] += x * y; accu[2] += x * z; accu[3] += y * y; accu[4] += y * z; accu[5] += z * z;
            accu[6] += x; accu[7] += y; accu[8] += z;
        }
        plane_from_accu(accu, (int)r.voxels.size(), n4, &curv);
    }
    // flipNormalTowardsViewpoint(centroid_, 0,0,0, n) ; n[3]=0 ; normalize
    float cos_theta = sum4((0.0f - r.cx) * n4[0], (0.0f - r.cy) * n4[1], (0.0f - r.cz) * n4[2], 0.0f * n4[3]);
    if (cos_theta < 0) { n4[0] *= -1; n4[1] *= -1; n4[2] *= -1; }
    n4[3] = 0.0f;
    float z = sum4(n4[0] * n4[0], n4[1] * n4[1], n4[2] * n4[2], 0.0f);
    if (z > 0.0f) { float s = std::sqrt(z); n4[0] /= s; n4[1] /= s; n4[2] /= s; }
    r.nx = n4[0]; r.ny = n4[1]; r.nz = n4[2]; r.curvature = curv;
}

typedef std::multimap<float, std::pair<uint32_t, uint32_t>> WeightMapT;   // clustering_state.h:50

// Clustering::contains, src/clustering.cpp:497-506.  The reference takes the map BY
// VALUE (a full copy per call); the copy is reproduced only when asked (cost model).
bool contains(const WeightMapT& w, uint32_t i1, uint32_t i2, bool literal_copy) {
    if (literal_copy) {
        WeightMapT c(w);
        for (auto& e : c) if (e.second.first == i1 && e.second.second == i2) return true;
        return false;
    }
    for (auto& e : w) if (e.second.first == i1 && e.second.second == i2) return true;
    return false;
}

} // namespace

// ---------------------------------------------------------------------------------
// Clustering::set_initialstate, src/clustering.cpp:605-612: clear_adjacency (:476-486)
// keeps first <= second; adj2weight (:193-207) gives every edge weight -1.
void Oracle::set_initialstate() {
    edges.clear();
    for (size_t i = 0; i + 1 < adj.size(); i += 2)
        if (adj[i] <= adj[i + 1]) edges.push_back(Edge{adj[i], adj[i + 1], -1.0f, -1.0f, -1.0f});
    segments = initial_segments;
    merges.clear();
}

// Clustering::init_weights, src/clustering.cpp:212-251 + init_merging_parameters :260-280
void Oracle::init_weights() {
    double t0 = now_ms();
    Merger M(*this);
    std::multiset<float> deltas_c, deltas_g;
    for (auto& e : edges) {
        auto d = M.delta_c_g(initial_segments.at(e.a), initial_segments.at(e.b));
        e.dc = d.first; e.dg = d.second;
        deltas_c.insert(d.first); deltas_g.insert(d.second);
    }
    lambda_used = P.lambda; cdf_c.clear(); cdf_g.clear();
    if (P.merge_mode == 1) {
        float mean_c = deltas_mean(deltas_c);
        float mean_g = deltas_mean(deltas_g);
        lambda_used = mean_g / (mean_c + mean_g);
    } else if (P.merge_mode == 2) {
        cdf_c = compute_cdf(deltas_c, P.bins);
        cdf_g = compute_cdf(deltas_g, P.bins);
    }
    for (auto& e : edges) e.w = M.t_c(e.dc) + M.t_g(e.dg);
    stage_ms[5] += now_ms() - t0;
}

// ---------------------------------------------------------------------------------
// Clustering::cluster(float) :670-679 -> cluster(ClusteringState, float) :384-396 ->
// merge :403-469.  merge_impl 0: literal std::multimap replay.  merge_impl 1: the
// same sequence through stamp tie-keys and prefix-continued region statistics
// (SURVEY.md Appendix C.2/C.3), for inputs where the literal form is too slow.
static void cluster_literal(Oracle& O, float threshold, bool literal_copy) {
    Merger M(O);
    WeightMapT wm;
    for (auto& e : O.edges) {
        if (std::isnan(e.w)) O.nan_weights++;
        wm.insert({e.w, {e.a, e.b}});                       // lexicographic insertion order, :237-246
    }
    O.segments = O.initial_segments;
    while (!wm.empty() && wm.begin()->first < threshold) {
        auto next = *wm.begin();
        const uint32_t a = next.second.first, b = next.second.second;
        O.merges.push_back(MergeRec{a, b, next.first, (uint32_t)wm.size(), (uint32_t)O.segments.size()});
        Region nr;
        nr.voxels = O.segments.at(a).voxels;                // operator+ : lhs first
        const auto& vb = O.segments.at(b).voxels;
        nr.voxels.insert(nr.voxels.end(), vb.begin(), vb.end());
        region_geometry(O, nr);
        O.segments.erase(a); O.segments.erase(b);
        O.segments[a] = nr;
        WeightMapT new_map;
        auto it = wm.begin(); ++it;
        for (; it != wm.end(); ++it) {
            std::pair<uint32_t, uint32_t> ids = it->second;
            bool touched = true;
            if (ids.first == a || ids.second == a) {
            } else if (ids.first == b) {
                ids.first = a;
            } else if (ids.second == b) {
                if (ids.first < a) ids.second = a;
                else { ids.second = ids.first; ids.first = a; }
            } else touched = false;
            if (!touched) { new_map.insert(*it); continue; }
            if (!contains(new_map, ids.first, ids.second, literal_copy)) {
                float w = M.delta(O.segments.at(ids.first), O.segments.at(ids.second));
                if (std::isnan(w)) O.nan_weights++;
                new_map.insert({w, ids});
            }
        }
        wm.swap(new_map);
    }
    O.final_edges.clear();
    for (auto& e : wm) O.final_edges.push_back(Edge{e.second.first, e.second.second, 0, 0, e.first});
}

namespace {
struct FastRegion {
    int n = 0;
    MeanState mean;
    float accu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<uint32_t> rope;          // initial supervoxel labels in concatenation order
    std::set<uint32_t> nb;               // adjacent region labels
    Region geo;                          // centroid/normal only (voxels unused)
};
struct EKey { float w; long long stamp; };
inline bool key_less(const EKey& x, const EKey& y) { return x.w < y.w || (x.w == y.w && x.stamp < y.stamp); }
}

static void cluster_fast(Oracle& O, float threshold) {
    Merger M(O);
    std::map<uint32_t, FastRegion> R;
    for (auto& kv : O.initial_segments) {
        FastRegion fr; fr.geo = kv.second; fr.geo.voxels.clear(); fr.rope.push_back(kv.first);
        for (int v : kv.second.voxels) {
            mean_step(fr.mean, O.vrgba[v]);
            float x = O.vxyz[3 * v], y = O.vxyz[3 * v + 1], z = O.vxyz[3 * v + 2];
            fr.accu[0] += x * x; fr.accu[1] += x * y; fr.accu[2] += x * z; fr.accu[3] += y * y; fr.accu[4] += y * z; fr.accu[5] += z * z;
            fr.accu[6] += x; fr.accu[7] += y; fr.accu[8] += z; fr.n++;
        }
        R[kv.first] = fr;
    }
    typedef std::pair<uint32_t, uint32_t> PairT;
    std::map<PairT, EKey> ek;
    auto cmp = [](const std::tuple<float, long long, uint32_t, uint32_t>& x, const std::tuple<float, long long, uint32_t, uint32_t>& y) {
        if (std::get<0>(x) != std::get<0>(y)) return std::get<0>(x) < std::get<0>(y);
        return std::get<1>(x) < std::get<1>(y); };
    std::set<std::tuple<float, long long, uint32_t, uint32_t>, decltype(cmp)> pq(cmp);
    long long counter = 0;
    for (auto& e : O.edges) {
        if (std::isnan(e.w)) { O.nan_weights++; }
        float w = std::isnan(e.w) ? std::numeric_limits<float>::infinity() : e.w;
        ek[{e.a, e.b}] = EKey{w, counter};
        pq.insert(std::make_tuple(w, counter, e.a, e.b));
        R[e.a].nb.insert(e.b); R[e.b].nb.insert(e.a);
        ++counter;
    }
    auto wdelta = [&](uint32_t x, uint32_t y) {
        const FastRegion& rx = R.at(x); const FastRegion& ry = R.at(y);
        float c1[3] = {rx.mean.r, rx.mean.g, rx.mean.b}, c2[3] = {ry.mean.r, ry.mean.g, ry.mean.b};
        auto d = M.delta_c_g_rgb(c1, c2, rx.geo, ry.geo);
        float w = M.t_c(d.first) + M.t_g(d.second);
        if (std::isnan(w)) { O.nan_weights++; w = std::numeric_limits<float>::infinity(); }
        return w; };
    size_t nregions = R.size();
    while (!pq.empty() && std::get<0>(*pq.begin()) < threshold) {
        auto head = *pq.begin();
        const uint32_t a = std::get<2>(head), b = std::get<3>(head);
        O.merges.push_back(MergeRec{a, b, std::get<0>(head), (uint32_t)pq.size(), (uint32_t)nregions});
        FastRegion& ra = R.at(a); FastRegion& rb = R.at(b);
        // old keys of every touched edge, before anything changes
        struct Touched { PairT old_pair; PairT new_pair; EKey old; };
        std::vector<Touched> tv;
        for (uint32_t x : ra.nb) if (x != b) { PairT p = {std::min(a, x), std::max(a, x)}; tv.push_back({p, p, ek.at(p)}); }
        for (uint32_t x : rb.nb) if (x != a) { PairT p = {std::min(b, x), std::max(b, x)}; tv.push_back({p, {std::min(a, x), std::max(a, x)}, ek.at(p)}); }
        std::sort(tv.begin(), tv.end(), [](const Touched& x, const Touched& y) { return key_less(x.old, y.old); });
        // fold b onto a: prefix continuation (C.3)
        for (uint32_t lbl : rb.rope) for (int v : O.initial_segments.at(lbl).voxels) {
            mean_step(ra.mean, O.vrgba[v]);
            float x = O.vxyz[3 * v], y = O.vxyz[3 * v + 1], z = O.vxyz[3 * v + 2];
            ra.accu[0] += x * x; ra.accu[1] += x * y; ra.accu[2] += x * z; ra.accu[3] += y * y; ra.accu[4] += y * z; ra.accu[5] += z * z;
            ra.accu[6] += x; ra.accu[7] += y; ra.accu[8] += z; ra.n++;
        }
        ra.rope.insert(ra.rope.end(), rb.rope.begin(), rb.rope.end());
        {
            float n = (float)ra.n;
            Region& g = ra.geo;
            g.cx = ra.accu[6] / n; g.cy = ra.accu[7] / n; g.cz = ra.accu[8] / n;
            float n4[4]; float curv;
            if (ra.n < 3) { n4[0] = n4[1] = n4[2] = n4[3] = std::numeric_limits<float>::quiet_NaN(); curv = n4[0]; }
            else plane_from_accu(ra.accu, ra.n, n4, &curv);
            float cos_theta = sum4((0.0f - g.cx) * n4[0], (0.0f - g.cy) * n4[1], (0.0f - g.cz) * n4[2], 0.0f * n4[3]);
            if (cos_theta < 0) { n4[0] *= -1; n4[1] *= -1; n4[2] *= -1; }
            n4[3] = 0.0f;
            float z = sum4(n4[0] * n4[0], n4[1] * n4[1], n4[2] * n4[2], 0.0f);
            if (z > 0.0f) { float s = std::sqrt(z); n4[0] /= s; n4[1] /= s; n4[2] /= s; }
            g.nx = n4[0]; g.ny = n4[1]; g.nz = n4[2]; g.curvature = curv;
        }
        // remove head + all touched edges from the structures
        pq.erase(pq.begin()); ek.erase({a, b});
        for (auto& t : tv) { pq.erase(std::make_tuple(t.old.w, t.old.stamp, t.old_pair.first, t.old_pair.second)); ek.erase(t.old_pair); }
        // graph relabel b -> a
        for (uint32_t x : rb.nb) if (x != a) { R.at(x).nb.erase(b); R.at(x).nb.insert(a); ra.nb.insert(x); }
        ra.nb.erase(b);
        R.erase(b); --nregions;
        // emit survivors in old order (dedupe keeps the earlier position)
        std::set<PairT> emitted;
        struct Arr { PairT p; float w; EKey old; };
        std::vector<Arr> fronts, backs;
        for (auto& t : tv) {
            if (!emitted.insert(t.new_pair).second) continue;
            float w = wdelta(t.new_pair.first, t.new_pair.second);
            if (w == t.old.w) { ek[t.new_pair] = EKey{w, t.old.stamp}; pq.insert(std::make_tuple(w, t.old.stamp, t.new_pair.first, t.new_pair.second)); }
            else if (w > t.old.w) fronts.push_back({t.new_pair, w, t.old});
            else backs.push_back({t.new_pair, w, t.old});
        }
        for (auto& x : backs) { long long s = counter++; ek[x.p] = EKey{x.w, s}; pq.insert(std::make_tuple(x.w, s, x.p.first, x.p.second)); }
        for (auto it = fronts.rbegin(); it != fronts.rend(); ++it) { long long s = -(counter++); ek[it->p] = EKey{it->w, s}; pq.insert(std::make_tuple(it->w, s, it->p.first, it->p.second)); }
    }
    // materialise the state in the literal representation
    O.segments.clear();
    for (auto& kv : R) {
        Region r = kv.second.geo;
        for (uint32_t lbl : kv.second.rope) { const auto& vv = O.initial_segments.at(lbl).voxels; r.voxels.insert(r.voxels.end(), vv.begin(), vv.end()); }
        O.segments[kv.first] = r;
    }
    O.final_edges.clear();
    for (auto& e : pq) O.final_edges.push_back(Edge{std::get<2>(e), std::get<3>(e), 0, 0, std::get<0>(e)});
}

void Oracle::cluster(float threshold) {
    double t0 = now_ms();
    merges.clear(); nan_weights = 0;
    if (P.merge_impl == 1) cluster_fast(*this, threshold);
    else cluster_literal(*this, threshold, P.merge_impl == 2);
    stage_ms[6] = now_ms() - t0;
}

// Clustering::get_labeled_cloud, src/clustering.cpp:640-663
void Oracle::labeled_cloud() {
    out_xyz.clear(); out_label.clear(); out_voxel.clear();
    uint32_t current_l = 0;
    for (auto& kv : segments) {
        for (int v : kv.second.voxels) {
            out_xyz.push_back(vxyz[3 * v]); out_xyz.push_back(vxyz[3 * v + 1]); out_xyz.push_back(vxyz[3 * v + 2]);
            out_label.push_back(current_l); out_voxel.push_back((uint32_t)v);
        }
        current_l++;
    }
}

void Oracle::run_all(float thr) {
    double t0 = now_ms();
    voxelize(); neighbors(); voxel_normals(); select_seeds(); expand(); make_supervoxels();
    set_initialstate(); init_weights(); cluster(thr); labeled_cloud();
    stage_ms[7] = now_ms() - t0;
}

} // namespace f3ps_oracle
