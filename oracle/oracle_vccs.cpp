// oracle_vccs.cpp -- CPU ORACLE, front half: PCL SupervoxelClustering (VCCS) as the
// reference drives it (src/supervoxel_clustering.cpp:315-367).  TEST INFRASTRUCTURE
// ONLY (see oracle.h).  PARITY UNPINNED: PCL is third-party and absent here; every
// function restates PCL 1.10 / Eigen 3.3 behaviour listed in SURVEY.md Appendix A.
#include "oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <set>
#include <stdexcept>
#include <unordered_map>

namespace f3ps_oracle {

// ---------------------------------------------------------------------------------
// float libm model: the reference calls float libm (std::log, atan2, cos, sin on
// float arguments inside PCL).  glibc's float functions are not bit-stable across
// versions; the oracle defines them as the correctly rounded result, obtained by
// evaluating in double and rounding once.
float cr_logf(float x) { return (float)std::log((double)x); }
float cr_atan2f(float y, float x) { return (float)std::atan2((double)y, (double)x); }
float cr_cosf(float x) { return (float)std::cos((double)x); }
float cr_sinf(float x) { return (float)std::sin((double)x); }

static inline bool finite3(float x, float y, float z) {
    return std::isfinite(x) && std::isfinite(y) && std::isfinite(z);
}

// Eigen 3.3 fixed-size reductions (Redux.h, redux_novec_unroller): a 3-vector sum is
// a0 + (a1 + a2); a Vector4f sum is vectorised, SSE3 hadd: (a0 + a1) + (a2 + a3).
static inline float sum3(float a0, float a1, float a2) { return a0 + (a1 + a2); }
static inline float sum4(float a0, float a1, float a2, float a3) { return (a0 + a1) + (a2 + a3); }
static inline float dot3(const float* a, const float* b) { return sum3(a[0] * b[0], a[1] * b[1], a[2] * b[2]); }

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---------------------------------------------------------------------------------
// main()'s input clean-up, src/supervoxel_clustering.cpp:315-337: z<0 -> |z|, every
// point (NaNs included) is copied to an unorganised cloud.
void Oracle::set_input(const uint8_t* pts, long n, int stride) {
    px.resize(n); py.resize(n); pz.resize(n); prgba.resize(n);
    const int rgba_off = (stride >= 32) ? 16 : 12;
    for (long i = 0; i < n; ++i) {
        const uint8_t* p = pts + (size_t)i * stride;
        float x, y, z; uint32_t c;
        std::memcpy(&x, p, 4); std::memcpy(&y, p + 4, 4); std::memcpy(&z, p + 8, 4);
        std::memcpy(&c, p + rgba_off, 4);
        if (P.fold_negative_z && z < 0) z = std::abs(z);
        px[i] = x; py[i] = y; pz[i] = z; prgba[i] = c;
    }
}

static inline uint64_t morton_xmajor(uint32_t x, uint32_t y, uint32_t z, int depth) {
    // child index per level = (xbit<<2)|(ybit<<1)|zbit  (OctreeKey::getChildIdxWithDepthMask)
    uint64_t m = 0;
    for (int b = 0; b < depth; ++b) {
        uint64_t c = (((x >> b) & 1u) << 2) | (((y >> b) & 1u) << 1) | ((z >> b) & 1u);
        m |= c << (3 * b);
    }
    return m;
}

// PCL OctreePointCloud::getKeyBitSize for an EMPTY tree (A.1): depth + centred cube.
static void key_bit_size(double mn[3], double mx[3], double res, int keybits_floor, int* depth_out) {
    const float minValue = std::numeric_limits<float>::epsilon();
    unsigned mk[3];
    for (int a = 0; a < 3; ++a) {
        if (keybits_floor) mk[a] = (unsigned)std::floor((mx[a] - mn[a]) / res);
        else mk[a] = (unsigned)std::ceil((mx[a] - mn[a] - minValue) / res);
    }
    unsigned max_voxels = std::max(std::max(std::max(mk[0], mk[1]), mk[2]), 2u);
    unsigned d = (unsigned)std::ceil(std::log2((double)max_voxels) - minValue);
    d = std::min(32u, d);
    double side = (double)(1ull << d) * res;
    for (int a = 0; a < 3; ++a) {
        double over = (side - (mx[a] - mn[a])) / 2.0;
        if (over > minValue) { mn[a] -= over; mx[a] += over; }
    }
    *depth_out = (int)d;
}

// ---------------------------------------------------------------------------------
// K1: OctreePointCloudAdjacency::addPointsFromInputCloud + VoxelData::addPoint /
// computeData (A.1) and the DFS leaf order (A.2).
void Oracle::voxelize() {
    double t0 = now_ms();
    const long n = (long)px.size();
    const double res = (double)P.voxel_res;
    float mnf[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
    float mxf[3] = {-std::numeric_limits<float>::max(), -std::numeric_limits<float>::max(), -std::numeric_limits<float>::max()};
    auto transform = [&](float& x, float& y, float& z) {
        if (P.use_transform) { x = x / z; y = y / z; z = cr_logf(z); }   // SupervoxelClustering::transformFunction
    };
    bool any = false;
    for (long i = 0; i < n; ++i) {
        float x = px[i], y = py[i], z = pz[i];
        transform(x, y, z);
        if (!finite3(x, y, z)) continue;
        any = true;
        if (x < mnf[0]) mnf[0] = x;
        if (y < mnf[1]) mnf[1] = y;
        if (z < mnf[2]) mnf[2] = z;
        if (x > mxf[0]) mxf[0] = x;
        if (y > mxf[1]) mxf[1] = y;
        if (z > mxf[2]) mxf[2] = z;
    }
    keys.clear(); morton.clear(); vxyz.clear(); vrgb.clear(); vrgba.clear(); vcount.clear();
    point_voxel.assign(n, -1);
    depth = 0;
    if (!any) { stage_ms[0] = now_ms() - t0; return; }
    for (int a = 0; a < 3; ++a) { bmin[a] = mnf[a]; bmax[a] = mxf[a]; }
    key_bit_size(bmin, bmax, res, P.sw.keybits_floor, &depth);
    if (depth > 21) throw std::runtime_error("oracle: octree depth > 21 not supported");

    // per point: key in double arithmetic (genOctreeKeyforPoint); points whose ORIGINAL
    // coordinates are not finite are skipped; a finite point whose transform is not
    // finite gets the default key (0,0,0).
    struct PK { uint64_t m; long i; uint32_t k[3]; };
    std::vector<PK> pk; pk.reserve(n);
    for (long i = 0; i < n; ++i) {
        if (!finite3(px[i], py[i], pz[i])) continue;
        float x = px[i], y = py[i], z = pz[i];
        transform(x, y, z);
        PK e; e.i = i;
        if (finite3(x, y, z)) {
            e.k[0] = (unsigned)(((double)x - bmin[0]) / res);
            e.k[1] = (unsigned)(((double)y - bmin[1]) / res);
            e.k[2] = (unsigned)(((double)z - bmin[2]) / res);
        } else { e.k[0] = e.k[1] = e.k[2] = 0; }
        e.m = morton_xmajor(e.k[0], e.k[1], e.k[2], depth);
        pk.push_back(e);
    }
    std::stable_sort(pk.begin(), pk.end(), [&](const PK& a, const PK& b) {
        return P.sw.leaf_order_descending ? a.m > b.m : a.m < b.m; });
    // leaves in DFS order; sums accumulate in INPUT order (stable sort keeps it)
    size_t s = 0;
    while (s < pk.size()) {
        size_t e = s;
        float sx = 0, sy = 0, sz = 0, sr = 0, sg = 0, sb = 0; int cnt = 0;
        while (e < pk.size() && pk[e].m == pk[s].m) {
            long i = pk[e].i;
            sx += px[i]; sy += py[i]; sz += pz[i];                    // VoxelData::addPoint
            uint32_t c = prgba[i];
            sr += (float)((c >> 16) & 255u); sg += (float)((c >> 8) & 255u); sb += (float)(c & 255u);
            ++cnt; point_voxel[i] = (int)morton.size(); ++e;
        }
        float fc = (float)cnt;                                        // computeData
        float r = sr / fc, g = sg / fc, b = sb / fc;
        keys.push_back(pk[s].k[0]); keys.push_back(pk[s].k[1]); keys.push_back(pk[s].k[2]);
        morton.push_back(pk[s].m);
        vxyz.push_back(sx / fc); vxyz.push_back(sy / fc); vxyz.push_back(sz / fc);
        vrgb.push_back(r); vrgb.push_back(g); vrgb.push_back(b);
        // VoxelData::getPoint: float -> uint32 truncation, alpha 0 (A.3)
        vrgba.push_back(((uint32_t)r << 16) | ((uint32_t)g << 8) | (uint32_t)b);
        vcount.push_back(cnt);
        s = e;
    }
    stage_ms[0] = now_ms() - t0;
}

// ---------------------------------------------------------------------------------
// K2: OctreePointCloudAdjacency::computeNeighbors (A.2): dx,dy,dz nested, clipped at
// the cube faces, self included, list order kept.
void Oracle::neighbors() {
    double t0 = now_ms();
    const int V = (int)morton.size();
    nbr.assign((size_t)V * 27, -1); nbr_count.assign(V, 0);
    std::unordered_map<uint64_t, int> lut; lut.reserve((size_t)V * 2);
    for (int v = 0; v < V; ++v) lut[morton[v]] = v;
    const uint32_t maxk = (depth >= 32) ? 0xffffffffu : ((1u << depth) - 1u);
    for (int v = 0; v < V; ++v) {
        uint32_t kx = keys[3 * v], ky = keys[3 * v + 1], kz = keys[3 * v + 2];
        if (kx > maxk || ky > maxk || kz > maxk) continue;
        int dxm = kx > 0 ? -1 : 0, dym = ky > 0 ? -1 : 0, dzm = kz > 0 ? -1 : 0;
        int dxM = kx == maxk ? 0 : 1, dyM = ky == maxk ? 0 : 1, dzM = kz == maxk ? 0 : 1;
        int c = 0;
        for (int dx = dxm; dx <= dxM; ++dx) for (int dy = dym; dy <= dyM; ++dy) for (int dz = dzm; dz <= dzM; ++dz) {
            auto it = lut.find(morton_xmajor(kx + dx, ky + dy, kz + dz, depth));
            if (it != lut.end()) nbr[(size_t)v * 27 + c++] = it->second;
        }
        nbr_count[v] = c;
    }
    stage_ms[1] = now_ms() - t0;
}

// ---------------------------------------------------------------------------------
// pcl::computeRoots2 / computeRoots / eigen33 (smallest eigenpair), Scalar = float (A.3)
static void compute_roots2(float b, float c, float roots[3]) {
    roots[0] = 0.0f;
    float d = (float)((double)(b * b) - 4.0 * (double)c);
    if (d < 0.0f) d = 0.0f;
    float sd = std::sqrt(d);
    roots[2] = 0.5f * (b + sd);
    roots[1] = 0.5f * (b - sd);
}

static void compute_roots(const float m[9], float roots[3]) {
    // m row-major symmetric 3x3
    const float m00 = m[0], m01 = m[1], m02 = m[2], m11 = m[4], m12 = m[5], m22 = m[8];
    float c0 = m00 * m11 * m22 + 2.0f * m01 * m02 * m12 - m00 * m12 * m12 - m11 * m02 * m02 - m22 * m01 * m01;
    float c1 = m00 * m11 - m01 * m01 + m00 * m22 - m02 * m02 + m11 * m22 - m12 * m12;
    float c2 = m00 + m11 + m22;
    if (std::abs(c0) < std::numeric_limits<float>::epsilon()) { compute_roots2(c2, c1, roots); return; }
    const float s_inv3 = (float)(1.0 / 3.0);
    const float s_sqrt3 = std::sqrt(3.0f);
    float c2_over_3 = c2 * s_inv3;
    float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
    if (a_over_3 > 0.0f) a_over_3 = 0.0f;
    float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
    float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
    if (q > 0.0f) q = 0.0f;
    float rho = std::sqrt(-a_over_3);
    float theta = cr_atan2f(std::sqrt(-q), half_b) * s_inv3;
    float cos_theta = cr_cosf(theta);
    float sin_theta = cr_sinf(theta);
    roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
    roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
    roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
    if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
    if (roots[1] >= roots[2]) {
        std::swap(roots[1], roots[2]);
        if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
    }
    if (roots[0] <= 0.0f) compute_roots2(c2, c1, roots);
}

void eigen33_smallest(const float cov[6], float* eigenvalue, float evec[3]) {
    // cov = xx,xy,xz,yy,yz,zz
    float mat[9] = {cov[0], cov[1], cov[2], cov[1], cov[3], cov[4], cov[2], cov[4], cov[5]};
    float scale = 0.0f;
    for (int i = 0; i < 9; ++i) scale = std::max(scale, std::abs(mat[i]));
    if (scale <= std::numeric_limits<float>::min()) scale = 1.0f;
    float sm[9];
    for (int i = 0; i < 9; ++i) sm[i] = mat[i] / scale;
    float roots[3];
    compute_roots(sm, roots);
    *eigenvalue = roots[0] * scale;
    sm[0] -= roots[0]; sm[4] -= roots[0]; sm[8] -= roots[0];
    const float* r0 = sm; const float* r1 = sm + 3; const float* r2 = sm + 6;
    auto cross = [](const float* a, const float* b, float* o) {
        o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; };
    float v1[3], v2[3], v3[3];
    cross(r0, r1, v1); cross(r0, r2, v2); cross(r1, r2, v3);
    float l1 = dot3(v1, v1), l2 = dot3(v2, v2), l3 = dot3(v3, v3);
    const float* v; float l;
    if (l1 >= l2 && l1 >= l3) { v = v1; l = l1; }
    else if (l2 >= l1 && l2 >= l3) { v = v2; l = l2; }
    else { v = v3; l = l3; }
    float s = std::sqrt(l);
    evec[0] = v[0] / s; evec[1] = v[1] / s; evec[2] = v[2] / s;
}

// computeMeanAndCovarianceMatrix tail + solvePlaneParameters (A.3).  accu = raw sums
// xx,xy,xz,yy,yz,zz,x,y,z over n samples.  normal4[3] receives the Hessian d term.
void plane_from_accu(const float accu_sum[9], int n, float normal4[4], float* curv) {
    float accu[9];
    for (int i = 0; i < 9; ++i) accu[i] = accu_sum[i] / (float)n;
    float cov[6];
    cov[0] = accu[0] - accu[6] * accu[6];
    cov[1] = accu[1] - accu[6] * accu[7];
    cov[2] = accu[2] - accu[6] * accu[8];
    cov[3] = accu[3] - accu[7] * accu[7];
    cov[4] = accu[4] - accu[7] * accu[8];
    cov[5] = accu[5] - accu[8] * accu[8];
    float ev, vec[3];
    eigen33_smallest(cov, &ev, vec);
    normal4[0] = vec[0]; normal4[1] = vec[1]; normal4[2] = vec[2];
    float eig_sum = cov[0] + cov[3] + cov[5];
    *curv = (eig_sum != 0) ? std::abs(ev / eig_sum) : 0.0f;
    // plane_parameters[3] = -1 * plane_parameters.dot(centroid) with [3]=0, centroid[3]=1
    normal4[3] = -1 * sum4(vec[0] * accu[6], vec[1] * accu[7], vec[2] * accu[8], 0.0f * 1.0f);
}

// flipNormalTowardsViewpoint(point, 0,0,0, normal4) then normal[3]=0; normalize() (A.3)
static void flip_and_normalize(const float p[3], float n4[4]) {
    float vp[4] = {0.0f - p[0], 0.0f - p[1], 0.0f - p[2], 0.0f};
    float cos_theta = sum4(vp[0] * n4[0], vp[1] * n4[1], vp[2] * n4[2], vp[3] * n4[3]);
    if (cos_theta < 0) { n4[0] *= -1; n4[1] *= -1; n4[2] *= -1; }
    n4[3] = 0.0f;
    float z = sum4(n4[0] * n4[0], n4[1] * n4[1], n4[2] * n4[2], n4[3] * n4[3]);
    if (z > 0.0f) { float s = std::sqrt(z); n4[0] /= s; n4[1] /= s; n4[2] /= s; n4[3] /= s; }
}

// K3: SupervoxelClustering::computeVoxelData (A.3)
void Oracle::voxel_normals() {
    double t0 = now_ms();
    const int V = (int)morton.size();
    normals.assign((size_t)V * 4, 0.0f); curvature.assign(V, 0.0f);
    const float qnan = std::numeric_limits<float>::quiet_NaN();
    for (int v = 0; v < V; ++v) {
        float accu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        int cnt = 0;
        auto add = [&](int i) {
            float x = vxyz[3 * i], y = vxyz[3 * i + 1], z = vxyz[3 * i + 2];
            if (P.sw.shifted_covariance) { x -= vxyz[3 * v]; y -= vxyz[3 * v + 1]; z -= vxyz[3 * v + 2]; }
            accu[0] += x * x; accu[1] += x * y; accu[2] += x * z;
            accu[3] += y * y; accu[4] += y * z; accu[5] += z * z;
            accu[6] += x; accu[7] += y; accu[8] += z; ++cnt; };
        add(v);
        for (int a = 0; a < nbr_count[v]; ++a) {
            int nb = nbr[(size_t)v * 27 + a];
            add(nb);
            for (int b = 0; b < nbr_count[nb]; ++b) add(nbr[(size_t)nb * 27 + b]);
        }
        float n4[4]; float curv;
        if (cnt < 3) { n4[0] = n4[1] = n4[2] = n4[3] = qnan; curv = qnan; }
        else plane_from_accu(accu, cnt, n4, &curv);
        flip_and_normalize(&vxyz[3 * v], n4);
        for (int k = 0; k < 4; ++k) normals[(size_t)v * 4 + k] = n4[k];
        curvature[v] = curv;
    }
    stage_ms[2] = now_ms() - t0;
}

// ---------------------------------------------------------------------------------
// K4: selectInitialSupervoxelSeeds (A.4): seed octree grown point by point
// (OctreePointCloud::adoptBoundingBoxToPoint), occupied-cell centres in DFS order,
// exact 1-NN voxel per centre, radius filter.
void Oracle::select_seeds() {
    double t0 = now_ms();
    seeds.clear(); seed_cells_nn.clear();
    const int V = (int)morton.size();
    if (V == 0) { stage_ms[3] = now_ms() - t0; return; }
    const double res = (double)P.seed_res;
    const float minValue = std::numeric_limits<float>::epsilon();
    double mn[3], mx[3]; int d = 0; bool defined = false;
    long long off[3] = {0, 0, 0};                      // cells added below the original origin
    std::vector<long long> ckey((size_t)V * 3);
    for (int v = 0; v < V; ++v) {
        const float p[3] = {vxyz[3 * v], vxyz[3 * v + 1], vxyz[3 * v + 2]};
        while (true) {
            bool lo[3], hi[3]; bool viol = false;
            for (int a = 0; a < 3; ++a) {
                lo[a] = defined && ((double)p[a] < mn[a]); hi[a] = defined && ((double)p[a] >= mx[a]);
                viol = viol || lo[a] || hi[a];
            }
            if (!viol && defined) break;
            if (defined) {
                double side = (double)(1ull << d) * res;
                for (int a = 0; a < 3; ++a) if (!hi[a]) { mn[a] -= side; off[a] += (1ll << d); }
                ++d;
                side = (double)(1ull << d) * res - minValue;
                for (int a = 0; a < 3; ++a) mx[a] = mn[a] + side;
            } else {
                for (int a = 0; a < 3; ++a) { mn[a] = (double)p[a] - res / 2; mx[a] = (double)p[a] + res / 2; }
                key_bit_size(mn, mx, res, P.sw.keybits_floor, &d);
                defined = true;
            }
        }
        for (int a = 0; a < 3; ++a)            // key at insertion time, kept relative to the first origin
            ckey[(size_t)v * 3 + a] = (long long)(unsigned)(((double)p[a] - mn[a]) / res) - off[a];
    }
    if (d > 21) throw std::runtime_error("oracle: seed octree depth > 21 not supported");
    seed_depth = d; for (int a = 0; a < 3; ++a) seed_min[a] = mn[a];
    // final keys and buckets
    std::map<uint64_t, std::vector<int>> cells;    // morton -> voxels (idx order)
    std::vector<uint32_t> fk((size_t)V * 3);
    for (int v = 0; v < V; ++v) {
        for (int a = 0; a < 3; ++a) fk[(size_t)v * 3 + a] = (uint32_t)(ckey[(size_t)v * 3 + a] + off[a]);
        cells[morton_xmajor(fk[3 * v], fk[3 * v + 1], fk[3 * v + 2], d)].push_back(v);
    }
    auto sqd = [&](const float* a, const float* b) {   // flann::L2_Simple<float>
        float r = 0.0f; for (int k = 0; k < 3; ++k) { float df = a[k] - b[k]; r += df * df; } return r; };
    auto for_cells_around = [&](const uint32_t k[3], auto&& fn) {
        for (int dx = -1; dx <= 1; ++dx) for (int dy = -1; dy <= 1; ++dy) for (int dz = -1; dz <= 1; ++dz) {
            long long x = (long long)k[0] + dx, y = (long long)k[1] + dy, z = (long long)k[2] + dz;
            if (x < 0 || y < 0 || z < 0 || x >= (1ll << d) || y >= (1ll << d) || z >= (1ll << d)) continue;
            auto it = cells.find(morton_xmajor((uint32_t)x, (uint32_t)y, (uint32_t)z, d));
            if (it != cells.end()) for (int u : it->second) fn(u);
        } };
    const float search_radius = 0.5f * P.seed_res;
    const float min_points = 0.05f * (search_radius) * (search_radius) * 3.1415926536f / (P.voxel_res * P.voxel_res);
    const float r2 = (float)((double)search_radius * (double)search_radius);
    for (auto& kv : cells) {                        // ascending x-major Morton = DFS order 0..7
        int v0 = kv.second[0];
        const uint32_t k[3] = {fk[3 * v0], fk[3 * v0 + 1], fk[3 * v0 + 2]};
        float c[3];
        for (int a = 0; a < 3; ++a) c[a] = (float)(((double)k[a] + 0.5f) * res + mn[a]);   // genLeafNodeCenterFromOctreeKey
        int best = -1; float bd = 0;
        for_cells_around(k, [&](int u) { float dd = sqd(c, &vxyz[3 * u]); if (best < 0 || dd < bd || (dd == bd && u < best)) { best = u; bd = dd; } });
        seed_cells_nn.push_back(best);
        const uint32_t kb[3] = {fk[3 * best], fk[3 * best + 1], fk[3 * best + 2]};
        int num = 0;
        for_cells_around(kb, [&](int u) { if (sqd(&vxyz[3 * best], &vxyz[3 * u]) < r2) ++num; });
        if ((float)num > min_points) seeds.push_back(best);
    }
    stage_ms[3] = now_ms() - t0;
}

// ---------------------------------------------------------------------------------
// K5: createSupervoxelHelpers + expandSupervoxels (A.5), literal sequential form.
namespace {
struct Helper {
    uint32_t label;
    std::set<int> leaves;                 // ordered by idx_ (compareLeaves)
    float xyz[3] = {0, 0, 0}, rgb[3] = {0, 0, 0}, nrm[4] = {0, 0, 0, 0};
    bool alive = true;
};
}

void Oracle::expand() {
    if (P.expand_impl == 1) { expand_fixed_point(); return; }
    double t0 = now_ms();
    const int V = (int)morton.size();
    labels.assign(V, 0); dist.assign(V, std::numeric_limits<float>::max());
    steals_per_round.clear();
    std::vector<Helper> H(seeds.size());
    std::vector<int> owner(V, -1);            // helper index
    for (size_t i = 0; i < seeds.size(); ++i) {
        H[i].label = (uint32_t)i + 1;
        H[i].leaves.insert(seeds[i]);         // addLeaf: owner_ = this (last helper wins a shared seed voxel)
        owner[seeds[i]] = (int)i;
    }
    auto update_centroid = [&](Helper& h) {   // SupervoxelHelper::updateCentroid
        float n[4] = {0, 0, 0, 0}, x[3] = {0, 0, 0}, c[3] = {0, 0, 0};
        for (int u : h.leaves) {
            for (int k = 0; k < 4; ++k) n[k] += normals[(size_t)u * 4 + k];
            for (int k = 0; k < 3; ++k) { x[k] += vxyz[3 * u + k]; c[k] += vrgb[3 * u + k]; }
        }
        float z = sum4(n[0] * n[0], n[1] * n[1], n[2] * n[2], n[3] * n[3]);
        if (z > 0.0f) { float s = std::sqrt(z); for (int k = 0; k < 4; ++k) n[k] /= s; }
        float cnt = (float)h.leaves.size();
        for (int k = 0; k < 3; ++k) { h.xyz[k] = x[k] / cnt; h.rgb[k] = c[k] / cnt; }
        for (int k = 0; k < 4; ++k) h.nrm[k] = n[k];
    };
    if (P.sw.init_centroid_seed_voxel) for (auto& h : H) update_centroid(h);
    auto vdist = [&](const Helper& h, int u) {  // voxelDataDistance(centroid_, voxel)
        float dx[3], dc[3];
        for (int k = 0; k < 3; ++k) { dx[k] = h.xyz[k] - vxyz[3 * u + k]; dc[k] = h.rgb[k] - vrgb[3 * u + k]; }
        float spatial = std::sqrt(dot3(dx, dx)) / P.seed_res;
        float color = std::sqrt(dot3(dc, dc)) / 255.0f;
        const float* m = &normals[(size_t)u * 4];
        float cosang = 1.0f - std::abs(sum4(h.nrm[0] * m[0], h.nrm[1] * m[1], h.nrm[2] * m[2], h.nrm[3] * m[3]));
        return cosang * P.normal_imp + color * P.color_imp + spatial * P.spatial_imp;
    };
    int max_depth = (int)(1.8f * P.seed_res / P.voxel_res);
    rounds = 0;
    for (int it = 1; it < max_depth; ++it) {
        int steals = 0;
        for (size_t hi = 0; hi < H.size(); ++hi) {
            Helper& h = H[hi];
            if (!h.alive) continue;
            std::vector<int> new_owned;
            for (int u : h.leaves) {
                for (int a = 0; a < nbr_count[u]; ++a) {
                    int nb = nbr[(size_t)u * 27 + a];
                    if (owner[nb] == (int)hi) continue;
                    float dd = vdist(h, nb);
                    if (dd < dist[nb]) {
                        dist[nb] = dd;
                        if (owner[nb] >= 0) { H[owner[nb]].leaves.erase(nb); ++steals; }
                        owner[nb] = (int)hi;
                        new_owned.push_back(nb);
                    }
                }
            }
            for (int u : new_owned) h.leaves.insert(u);
        }
        for (auto& h : H) {
            if (!h.alive) continue;
            if (h.leaves.empty()) h.alive = false; else update_centroid(h);
        }
        steals_per_round.push_back(steals);
        ++rounds;
    }
    for (int v = 0; v < V; ++v) labels[v] = owner[v] >= 0 ? H[owner[v]].label : 0;
    // K6a part 1: helper centroids of survivors (makeSupervoxels)
    sv_label.clear(); sv_xyz.clear(); sv_rgb.clear(); sv_normal.clear(); sv_count.clear(); sv_leaves.clear();
    for (auto& h : H) {
        if (!h.alive) continue;
        sv_leaves.push_back(std::vector<int>(h.leaves.begin(), h.leaves.end()));
        // a helper that never ran updateCentroid (max_depth<=1) keeps its initial centroid
        sv_label.push_back(h.label);
        for (int k = 0; k < 3; ++k) { sv_xyz.push_back(h.xyz[k]); sv_rgb.push_back(h.rgb[k]); }
        for (int k = 0; k < 4; ++k) sv_normal.push_back(h.nrm[k]);
        sv_count.push_back((int)h.leaves.size());
    }
    stage_ms[4] = now_ms() - t0;
}

// ---------------------------------------------------------------------------------
// K5 restated as a data-parallel fixed point -- the formulation the CUDA kernels run,
// kept here so that its equivalence with the literal sequential form above is testable without a GPU.
//   per round: start-of-round owner0/D0; ST[u] = label of the first helper that steals u this round.
//   sweep: every voxel n folds, in ascending label order, the helpers that hold a leaf adjacent to n at
//   their turn: regular leaves u != n with owner0[u] == h and ST[u] > h, plus "phantom" leaves.
//   Phantom leaf: two seed cells elected the same voxel u; createSupervoxelHelpers puts u into both
//   helpers' leaf sets but only the later helper owns it.  The earlier helper keeps expanding from u
//   (also onto u itself), counts u in its centroid, and lists u in its voxels_, until it steals u.
void Oracle::expand_fixed_point() {
    double t0 = now_ms();
    const int V = (int)morton.size();
    const int S0 = (int)seeds.size();
    const uint32_t NONE = 0xffffffffu;
    labels.assign(V, 0); dist.assign(V, std::numeric_limits<float>::max());
    steals_per_round.clear(); sweeps_total = 0;
    std::vector<uint32_t> owner0(V, 0), owner1(V, 0), phantom(V, 0);     // phantom[u] = label holding u without owning it
    std::vector<float> D0(V, std::numeric_limits<float>::max()), D1(V);
    std::vector<int> phantom_leaf(S0 + 1, -1);                           // per label
    struct Cen { float xyz[3] = {0, 0, 0}, rgb[3] = {0, 0, 0}, nrm[4] = {0, 0, 0, 0}; bool alive = true; };
    std::vector<Cen> cen(S0 + 1);
    long triple = 0;
    for (int i = 0; i < S0; ++i) {                                       // createSupervoxelHelpers: last helper owns
        const int u = seeds[i]; const uint32_t l = (uint32_t)i + 1;
        if (owner0[u]) { if (phantom[u]) ++triple; phantom[u] = owner0[u]; phantom_leaf[owner0[u]] = u; }
        owner0[u] = l;
    }
    if (triple) throw std::runtime_error("oracle(fixed point): three helpers on one seed voxel is not modelled");
    auto vdist = [&](const Cen& h, int u) {
        float dx[3], dc[3];
        for (int k = 0; k < 3; ++k) { dx[k] = h.xyz[k] - vxyz[3 * u + k]; dc[k] = h.rgb[k] - vrgb[3 * u + k]; }
        float spatial = std::sqrt(dot3(dx, dx)) / P.seed_res;
        float color = std::sqrt(dot3(dc, dc)) / 255.0f;
        const float* m = &normals[(size_t)u * 4];
        float cosang = 1.0f - std::abs(sum4(h.nrm[0] * m[0], h.nrm[1] * m[1], h.nrm[2] * m[2], h.nrm[3] * m[3]));
        return cosang * P.normal_imp + color * P.color_imp + spatial * P.spatial_imp;
    };
    auto fold_centroids = [&](const std::vector<uint32_t>& owner) {
        // voxels grouped by owner in idx order (stable), phantom leaf merged in at its idx position
        std::vector<std::vector<int>> lists(S0 + 1);
        for (int v = 0; v < V; ++v) if (owner[v]) lists[owner[v]].push_back(v);
        for (int l = 1; l <= S0; ++l) {
            Cen& c = cen[l];
            if (!c.alive) continue;
            std::vector<int>& L = lists[l];
            if (phantom_leaf[l] >= 0) L.insert(std::lower_bound(L.begin(), L.end(), phantom_leaf[l]), phantom_leaf[l]);
            if (L.empty()) { c.alive = false; continue; }
            float n[4] = {0, 0, 0, 0}, x[3] = {0, 0, 0}, col[3] = {0, 0, 0};
            for (int u : L) {
                for (int k = 0; k < 4; ++k) n[k] += normals[(size_t)u * 4 + k];
                for (int k = 0; k < 3; ++k) { x[k] += vxyz[3 * u + k]; col[k] += vrgb[3 * u + k]; }
            }
            float z = sum4(n[0] * n[0], n[1] * n[1], n[2] * n[2], n[3] * n[3]);
            if (z > 0.0f) { float s = std::sqrt(z); for (int k = 0; k < 4; ++k) n[k] /= s; }
            float cnt = (float)L.size();
            for (int k = 0; k < 3; ++k) { c.xyz[k] = x[k] / cnt; c.rgb[k] = col[k] / cnt; }
            for (int k = 0; k < 4; ++k) c.nrm[k] = n[k];
        }
    };
    if (P.sw.init_centroid_seed_voxel) fold_centroids(owner0);
    const int max_depth = (int)(1.8f * P.seed_res / P.voxel_res);
    rounds = 0;
    std::vector<uint32_t> st_in(V), st_out(V);
    std::vector<uint8_t> ph_won(V, 0);                                   // the phantom holder stole its own leaf this round
    for (int it = 1; it < max_depth; ++it) {
        std::fill(st_in.begin(), st_in.end(), NONE);
        int sweeps = 0;
        while (true) {
            bool changed = false;
            for (int n = 0; n < V; ++n) {
                uint32_t cur = owner0[n]; float D = D0[n];
                std::vector<uint32_t> cand;
                for (int a = 0; a < nbr_count[n]; ++a) {
                    const int u = nbr[(size_t)n * 27 + a];
                    if (u != n) { const uint32_t h = owner0[u]; if (h && st_in[u] > h) cand.push_back(h); }
                    if (phantom[u]) cand.push_back(phantom[u]);            // phantom leaves support every neighbour, u itself included
                }
                std::sort(cand.begin(), cand.end());
                cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
                uint32_t first = NONE;
                bool won = false;
                for (uint32_t h : cand) {
                    if (h == cur || !cen[h].alive) continue;
                    const float d = vdist(cen[h], n);
                    if (d < D) { if (first == NONE) first = h; D = d; cur = h; if (h == phantom[n]) won = true; }
                }
                owner1[n] = cur; D1[n] = D; st_out[n] = first; ph_won[n] = won ? 1 : 0;
                if (first != st_in[n]) changed = true;
            }
            ++sweeps;
            st_in.swap(st_out);
            if (!changed) break;
            if (sweeps > 64) throw std::runtime_error("oracle(fixed point): no convergence");
        }
        sweeps_total += sweeps;
        int steals = 0;
        for (int v = 0; v < V; ++v) if (owner0[v] && owner1[v] != owner0[v]) ++steals;
        steals_per_round.push_back(steals);
        owner0 = owner1; D0 = D1;
        // a holder that stole its phantom leaf owns it regularly from then on (and loses it for good if a later helper
        // steals it again in the same round: removeLeaf erases it from the holder's set)
        for (int v = 0; v < V; ++v) if (phantom[v] && ph_won[v]) { phantom_leaf[phantom[v]] = -1; phantom[v] = 0; }
        fold_centroids(owner0);
        ++rounds;
    }
    for (int v = 0; v < V; ++v) { labels[v] = owner0[v]; dist[v] = D0[v]; }
    sv_label.clear(); sv_xyz.clear(); sv_rgb.clear(); sv_normal.clear(); sv_count.clear(); sv_leaves.clear();
    std::vector<int> counts(S0 + 1, 0);
    for (int v = 0; v < V; ++v) counts[owner0[v]]++;
    std::vector<std::vector<int>> final_lists(S0 + 1);
    for (int v = 0; v < V; ++v) if (owner0[v]) final_lists[owner0[v]].push_back(v);
    for (int l = 1; l <= S0; ++l) if (phantom_leaf[l] >= 0) {
        std::vector<int>& L = final_lists[l];
        L.insert(std::lower_bound(L.begin(), L.end(), phantom_leaf[l]), phantom_leaf[l]);
    }
    for (int l = 1; l <= S0; ++l) {
        if (rounds == 0) { if (!counts[l] && phantom_leaf[l] < 0) continue; }
        else if (!cen[l].alive) continue;
        sv_label.push_back((uint32_t)l);
        for (int k = 0; k < 3; ++k) { sv_xyz.push_back(cen[l].xyz[k]); sv_rgb.push_back(cen[l].rgb[k]); }
        for (int k = 0; k < 4; ++k) sv_normal.push_back(cen[l].nrm[k]);
        sv_count.push_back(counts[l] + (phantom_leaf[l] >= 0 ? 1 : 0));
        sv_leaves.push_back(final_lists[l]);
    }
    stage_ms[4] = now_ms() - t0;
}

// ---------------------------------------------------------------------------------
// pcl::SupervoxelClustering::refineSupervoxels(num_itr, clusters) (call site /root/reference/src/supervoxel_clustering.cpp:369-371),
// PCL 1.10 semantics restated (supervoxel_clustering.hpp): per iteration
//   SupervoxelHelper::refineNormals   every leaf's normal / curvature from computePointNormal over the index list
//                                     [neighbour u + u's neighbours, both only when owned by this helper] (duplicates kept,
//                                     the leaf itself comes in as its own neighbour), flipped towards the origin, normalised
//   reseedSupervoxels                 every helper drops its leaves (owner_ = 0, distance_ = max), then takes the voxel nearest
//                                     to its centroid (kd-tree 1-NN) as its only leaf -- centroids are NOT reset
//   expandSupervoxels(max_depth)      the rounds of extract(), starting from those centroids
// and makeSupervoxels at the end.  Works on the surviving helpers of expand() (sv_* arrays) and leaves the same arrays behind.
void Oracle::refine(int num_itr) {
    const int V = (int)morton.size();
    const int S = (int)sv_label.size();
    std::vector<Helper> H(S);
    std::vector<int> owner(V, -1);
    for (int s = 0; s < S; ++s) {
        H[s].label = sv_label[s];
        H[s].leaves.insert(sv_leaves[s].begin(), sv_leaves[s].end());
        for (int k = 0; k < 3; ++k) { H[s].xyz[k] = sv_xyz[3 * s + k]; H[s].rgb[k] = sv_rgb[3 * s + k]; }
        for (int k = 0; k < 4; ++k) H[s].nrm[k] = sv_normal[4 * s + k];
    }
    std::map<uint32_t, int> helper_of;
    for (int s = 0; s < S; ++s) helper_of[sv_label[s]] = s;
    for (int v = 0; v < V; ++v) if (labels[v]) owner[v] = helper_of[labels[v]];
    const float qnan = std::numeric_limits<float>::quiet_NaN();
    auto update_centroid = [&](Helper& h) {
        float n[4] = {0, 0, 0, 0}, x[3] = {0, 0, 0}, c[3] = {0, 0, 0};
        for (int u : h.leaves) {
            for (int k = 0; k < 4; ++k) n[k] += normals[(size_t)u * 4 + k];
            for (int k = 0; k < 3; ++k) { x[k] += vxyz[3 * u + k]; c[k] += vrgb[3 * u + k]; }
        }
        float z = sum4(n[0] * n[0], n[1] * n[1], n[2] * n[2], n[3] * n[3]);
        if (z > 0.0f) { float sq = std::sqrt(z); for (int k = 0; k < 4; ++k) n[k] /= sq; }
        float cnt = (float)h.leaves.size();
        for (int k = 0; k < 3; ++k) { h.xyz[k] = x[k] / cnt; h.rgb[k] = c[k] / cnt; }
        for (int k = 0; k < 4; ++k) h.nrm[k] = n[k];
    };
    auto vdist = [&](const Helper& h, int u) {
        float dx[3], dc[3];
        for (int k = 0; k < 3; ++k) { dx[k] = h.xyz[k] - vxyz[3 * u + k]; dc[k] = h.rgb[k] - vrgb[3 * u + k]; }
        float spatial = std::sqrt(dot3(dx, dx)) / P.seed_res;
        float color = std::sqrt(dot3(dc, dc)) / 255.0f;
        const float* m = &normals[(size_t)u * 4];
        float cosang = 1.0f - std::abs(sum4(h.nrm[0] * m[0], h.nrm[1] * m[1], h.nrm[2] * m[2], h.nrm[3] * m[3]));
        return cosang * P.normal_imp + color * P.color_imp + spatial * P.spatial_imp;
    };
    const int max_depth = (int)(1.8f * P.seed_res / P.voxel_res);
    for (int itr = 0; itr < num_itr; ++itr) {
        // ---- refineNormals (helpers in list order; reads only centroids and owners, so the last writer of a voxel decides) ----
        std::vector<float> nn = normals, cc = curvature;
        for (int hi = 0; hi < (int)H.size(); ++hi) {
            if (!H[hi].alive) continue;
            for (int v : H[hi].leaves) {
                float accu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; int cnt = 0;
                auto add = [&](int i) {
                    const float x = vxyz[3 * i], y = vxyz[3 * i + 1], z = vxyz[3 * i + 2];
                    accu[0] += x * x; accu[1] += x * y; accu[2] += x * z; accu[3] += y * y; accu[4] += y * z; accu[5] += z * z;
                    accu[6] += x; accu[7] += y; accu[8] += z; ++cnt; };
                for (int a = 0; a < nbr_count[v]; ++a) {
                    const int u = nbr[(size_t)v * 27 + a];
                    if (owner[u] != hi) continue;
                    add(u);
                    for (int b = 0; b < nbr_count[u]; ++b) { const int w = nbr[(size_t)u * 27 + b]; if (owner[w] == hi) add(w); }
                }
                float n4[4]; float curv;
                if (cnt < 3) { n4[0] = n4[1] = n4[2] = n4[3] = qnan; curv = qnan; }
                else plane_from_accu(accu, cnt, n4, &curv);
                flip_and_normalize(&vxyz[3 * v], n4);
                for (int k = 0; k < 4; ++k) nn[(size_t)v * 4 + k] = n4[k];
                cc[v] = curv;
            }
        }
        normals.swap(nn); curvature.swap(cc);
        // ---- reseedSupervoxels ----
        for (auto& h : H) { for (int u : h.leaves) { owner[u] = -1; dist[u] = std::numeric_limits<float>::max(); } h.leaves.clear(); }
        for (int hi = 0; hi < (int)H.size(); ++hi) {
            if (!H[hi].alive) continue;
            int best = -1; float bd = 0;
            for (int u = 0; u < V; ++u) {                       // exact 1-NN (flann::L2_Simple), ties to the lowest index
                float r = 0.0f; for (int k = 0; k < 3; ++k) { const float df = H[hi].xyz[k] - vxyz[3 * u + k]; r += df * df; }
                if (best < 0 || r < bd) { best = u; bd = r; }
            }
            if (best >= 0) { H[hi].leaves.insert(best); owner[best] = hi; }
        }
        // ---- expandSupervoxels ----
        for (int it = 1; it < max_depth; ++it) {
            for (int hi = 0; hi < (int)H.size(); ++hi) {
                Helper& h = H[hi];
                if (!h.alive) continue;
                std::vector<int> new_owned;
                for (int u : h.leaves)
                    for (int a = 0; a < nbr_count[u]; ++a) {
                        const int nb = nbr[(size_t)u * 27 + a];
                        if (owner[nb] == hi) continue;
                        const float dd = vdist(h, nb);
                        if (dd < dist[nb]) {
                            dist[nb] = dd;
                            if (owner[nb] >= 0) H[owner[nb]].leaves.erase(nb);
                            owner[nb] = hi;
                            new_owned.push_back(nb);
                        }
                    }
                for (int u : new_owned) h.leaves.insert(u);
            }
            for (auto& h : H) { if (!h.alive) continue; if (h.leaves.empty()) h.alive = false; else update_centroid(h); }
        }
    }
    for (int v = 0; v < V; ++v) labels[v] = owner[v] >= 0 ? H[owner[v]].label : 0;
    sv_label.clear(); sv_xyz.clear(); sv_rgb.clear(); sv_normal.clear(); sv_count.clear(); sv_leaves.clear();
    for (auto& h : H) {
        if (!h.alive) continue;
        sv_leaves.push_back(std::vector<int>(h.leaves.begin(), h.leaves.end()));
        sv_label.push_back(h.label);
        for (int k = 0; k < 3; ++k) { sv_xyz.push_back(h.xyz[k]); sv_rgb.push_back(h.rgb[k]); }
        for (int k = 0; k < 4; ++k) sv_normal.push_back(h.nrm[k]);
        sv_count.push_back((int)h.leaves.size());
    }
}

// K6a: makeSupervoxels + getSupervoxelAdjacency (A.6): both walk the helpers' LEAF SETS (getVoxels,
// getNeighborLabels), so a phantom leaf (see expand_fixed_point) is listed by its holder as well.
void Oracle::make_supervoxels() {
    double t0 = now_ms();
    initial_segments.clear(); adj.clear();
    std::set<std::pair<uint32_t, uint32_t>> pairs;
    for (size_t s = 0; s < sv_label.size(); ++s) {
        Region r;
        r.cx = sv_xyz[3 * s]; r.cy = sv_xyz[3 * s + 1]; r.cz = sv_xyz[3 * s + 2];
        r.nx = sv_normal[4 * s]; r.ny = sv_normal[4 * s + 1]; r.nz = sv_normal[4 * s + 2];
        r.curvature = 0.0f;
        r.voxels = sv_leaves[s];
        const uint32_t l = sv_label[s];
        for (int v : r.voxels)
            for (int a = 0; a < nbr_count[v]; ++a) {
                uint32_t m = labels[nbr[(size_t)v * 27 + a]];
                if (m && m != l) pairs.insert({l, m});
            }
        initial_segments[l] = r;
    }
    // adjacency is symmetric in PCL only through both helpers' walks; a phantom leaf adds (holder -> owner) one way
    for (auto& p : pairs) { adj.push_back(p.first); adj.push_back(p.second); }
    stage_ms[5] = now_ms() - t0;
}

} // namespace f3ps_oracle
