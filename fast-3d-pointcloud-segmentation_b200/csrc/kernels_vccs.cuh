// kernels_vccs.cuh -- K1..K5: the VCCS front end (pcl::SupervoxelClustering as driven by
// /root/reference/src/supervoxel_clustering.cpp:348-367) as data-parallel sm_100a kernels.
// Behavioural spec: SURVEY.md Appendix A (PCL 1.10 semantics).  All float arithmetic is
// the scalar IEEE sequence (compiled -fmad=false); ordered sums stay ordered.
#pragma once
#include "common.cuh"

namespace f3ps {

// ---- input record: pcl::PointXYZRGBA (stride 32) or packed {x,y,z,bgra} (stride 16) -------
struct PointLoader {
    const uint8_t* base; int stride; int fold_z;
    __device__ __forceinline__ float4 xyzw(int64_t i) const {   // w carries rgba bits for stride 16
        float4 p = *reinterpret_cast<const float4*>(base + (size_t)i * stride);
        if (fold_z && p.z < 0) p.z = fabsf(p.z);                 // main(): z<0 -> |z|  (:317-321)
        return p;
    }
    __device__ __forceinline__ uint32_t rgba(int64_t i, const float4& p) const {
        return stride >= 32 ? *reinterpret_cast<const uint32_t*>(base + (size_t)i * stride + 16) : __float_as_uint(p.w);
    }
};

__device__ __forceinline__ void vccs_transform(int use_transform, float& x, float& y, float& z) {
    if (use_transform) { x = x / z; y = y / z; z = cr_logf(z); }   // SupervoxelClustering::transformFunction
}

// ---- K1a: bounding box of the transformed finite points (A.1) ---------------------------
__global__ void __launch_bounds__(256) bbox_kernel(PointLoader pl, int64_t n, int use_transform, FrameParams* fp) {
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    int any = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float4 p = pl.xyzw(i);
        float x = p.x, y = p.y, z = p.z;
        vccs_transform(use_transform, x, y, z);
        if (!finite3(x, y, z)) continue;
        any = 1;
        mn[0] = fminf(mn[0], x); mn[1] = fminf(mn[1], y); mn[2] = fminf(mn[2], z);
        mx[0] = fmaxf(mx[0], x); mx[1] = fmaxf(mx[1], y); mx[2] = fmaxf(mx[2], z);
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(kFull, mn[a], off));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(kFull, mx[a], off));
        }
        any |= __shfl_xor_sync(kFull, any, off);
    }
    if ((threadIdx.x & 31) == 0 && any) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { atomicMin(&fp->ord_min[a], f2ord(mn[a])); atomicMax(&fp->ord_max[a], f2ord(mx[a])); }
        fp->any_finite = 1;
    }
}

// ---- K1b: OctreePointCloud::defineBoundingBox + getKeyBitSize on an empty tree (A.1) ----
__device__ inline int key_bit_size(double mn[3], double mx[3], double res) {
    const float minValue = FLT_EPSILON;
    unsigned mk[3];
    for (int a = 0; a < 3; ++a) mk[a] = (unsigned)ceil((mx[a] - mn[a] - minValue) / res);
    unsigned max_voxels = max(max(max(mk[0], mk[1]), mk[2]), 2u);
    unsigned d = (unsigned)ceil(log2((double)max_voxels) - minValue);
    d = min(32u, d);
    double side = (double)(1ull << d) * res;
    for (int a = 0; a < 3; ++a) {
        double over = (side - (mx[a] - mn[a])) / 2.0;
        if (over > minValue) { mn[a] -= over; mx[a] += over; }
    }
    return (int)d;
}

__global__ void frame_setup_kernel(FrameParams* fp, float voxel_res) {
    if (threadIdx.x || blockIdx.x) return;
    fp->res = (double)voxel_res;
    if (!fp->any_finite) { fp->depth = 0; return; }
    double mn[3], mx[3];
    for (int a = 0; a < 3; ++a) { mn[a] = (double)ord2f(fp->ord_min[a]); mx[a] = (double)ord2f(fp->ord_max[a]); }
    fp->depth = key_bit_size(mn, mx, fp->res);
    for (int a = 0; a < 3; ++a) fp->bmin[a] = mn[a];
}

// ---- K1c: per-point Morton key (genOctreeKeyforPoint) fused with the compaction of the
// points whose ORIGINAL coordinates are finite.  Plugs into compact_kernel (radix_sort.cuh).
template <typename KeyT>
struct KeygenOp {
    PointLoader pl; int use_transform; const FrameParams* fp; KeyT* keys; unsigned* vals;
    typedef KeyT Payload;
    __device__ __forceinline__ bool test(int64_t i, Payload& key) const {
        float4 p = pl.xyzw(i);
        if (!finite3(p.x, p.y, p.z)) return false;
        float x = p.x, y = p.y, z = p.z;
        vccs_transform(use_transform, x, y, z);
        uint32_t kx = 0, ky = 0, kz = 0;
        if (finite3(x, y, z)) {
            const double res = fp->res;
            kx = (unsigned)(((double)x - fp->bmin[0]) / res);
            ky = (unsigned)(((double)y - fp->bmin[1]) / res);
            kz = (unsigned)(((double)z - fp->bmin[2]) / res);
        }
        key = (KeyT)morton_xmajor(kx, ky, kz);
        return true;
    }
    __device__ __forceinline__ void emit(unsigned pos, int64_t i, const Payload& key) const { keys[pos] = key; vals[pos] = (unsigned)i; }
};

// heads of equal-key runs in a sorted key array
template <typename KeyT>
struct HeadOp {
    const KeyT* keys; unsigned* starts;
    typedef int Payload;
    __device__ __forceinline__ bool test(int64_t i, Payload&) const { return i == 0 || keys[i] != keys[i - 1]; }
    __device__ __forceinline__ void emit(unsigned pos, int64_t i, const Payload&) const { starts[pos] = (unsigned)i; }
};

// ---- K1d: VoxelData::addPoint / computeData: ordered per-voxel sums (A.1) -----------------
template <typename KeyT>
__global__ void __launch_bounds__(128) voxel_accumulate_kernel(PointLoader pl, const KeyT* __restrict__ sorted_keys,
        const unsigned* __restrict__ sorted_idx, const unsigned* __restrict__ starts, const unsigned* __restrict__ n_vox_ptr,
        const unsigned* __restrict__ n_valid_ptr, float4* __restrict__ vox_xyz, float4* __restrict__ vox_rgb,
        uint64_t* __restrict__ vox_key, int* __restrict__ point_voxel) {
    const unsigned V = *n_vox_ptr;
    const unsigned n_valid = *n_valid_ptr;
    for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
        unsigned s = starts[v], e = (v + 1 < V) ? starts[v + 1] : n_valid;
        float sx = 0, sy = 0, sz = 0, sr = 0, sg = 0, sb = 0;
        for (unsigned j = s; j < e; ++j) {
            unsigned i = sorted_idx[j];
            float4 p = pl.xyzw(i);
            uint32_t c = pl.rgba(i, p);
            sx += p.x; sy += p.y; sz += p.z;
            sr += (float)((c >> 16) & 255u); sg += (float)((c >> 8) & 255u); sb += (float)(c & 255u);
            point_voxel[i] = (int)v;
        }
        float fc = (float)(e - s);
        float r = sr / fc, g = sg / fc, b = sb / fc;
        uint32_t rgba = ((uint32_t)r << 16) | ((uint32_t)g << 8) | (uint32_t)b;   // VoxelData::getPoint truncation (A.3)
        vox_xyz[v] = make_float4(sx / fc, sy / fc, sz / fc, __uint_as_float(rgba));
        vox_rgb[v] = make_float4(r, g, b, fc);
        vox_key[v] = (uint64_t)sorted_keys[s];
    }
}

// ---- K2: voxel hash table + 26(+self) neighbourhood (A.2) ---------------------------------
__device__ __forceinline__ uint32_t hash64(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (uint32_t)k;
}
__global__ void __launch_bounds__(256) hash_build_kernel(const uint64_t* __restrict__ vox_key, const unsigned* __restrict__ n_vox_ptr,
        unsigned long long* __restrict__ slots, unsigned* __restrict__ slot_val, unsigned mask) {
    const unsigned V = *n_vox_ptr;
    for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
        unsigned long long k = vox_key[v] + 1ull;
        unsigned h = hash64(k) & mask;
        while (true) {
            unsigned long long prev = atomicCAS(&slots[h], 0ull, k);
            if (prev == 0ull || prev == k) { slot_val[h] = v; break; }
            h = (h + 1) & mask;
        }
    }
}
__device__ __forceinline__ int hash_find(const unsigned long long* __restrict__ slots, const unsigned* __restrict__ slot_val,
                                         unsigned mask, uint64_t key) {
    unsigned long long k = key + 1ull;
    unsigned h = hash64(k) & mask;
    while (true) {
        unsigned long long s = slots[h];
        if (s == k) return (int)slot_val[h];
        if (s == 0ull) return -1;
        h = (h + 1) & mask;
    }
}

constexpr int kNbrStride = 28;   // row-major rows: 27 entries + count

// computeNeighbors: dx,dy,dz nested (lexicographic), clipped at the cube faces, self included.
// Writes the list twice: row-major [V][28] (2-ring walks) and column-major [27][V] (per-voxel sweeps).
__global__ void __launch_bounds__(128) neighbors_kernel(const uint64_t* __restrict__ vox_key, const unsigned* __restrict__ n_vox_ptr,
        const FrameParams* __restrict__ fp, const unsigned long long* __restrict__ slots, const unsigned* __restrict__ slot_val,
        unsigned mask, int* __restrict__ nbr_row, int* __restrict__ nbr_col, unsigned V_cap) {
    const unsigned V = *n_vox_ptr;
    const int depth = fp->depth;
    const uint32_t maxk = depth >= 32 ? 0xffffffffu : ((1u << depth) - 1u);
    for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
        uint32_t kx, ky, kz;
        morton_decode(vox_key[v], kx, ky, kz);
        int cnt = 0;
        int* row = nbr_row + (size_t)v * kNbrStride;
        if (!(kx > maxk || ky > maxk || kz > maxk)) {
            const int dxm = kx > 0 ? -1 : 0, dym = ky > 0 ? -1 : 0, dzm = kz > 0 ? -1 : 0;
            const int dxM = kx == maxk ? 0 : 1, dyM = ky == maxk ? 0 : 1, dzM = kz == maxk ? 0 : 1;
            for (int dx = dxm; dx <= dxM; ++dx) for (int dy = dym; dy <= dyM; ++dy) for (int dz = dzm; dz <= dzM; ++dz) {
                int u = (dx == 0 && dy == 0 && dz == 0) ? (int)v
                        : hash_find(slots, slot_val, mask, morton_xmajor(kx + dx, ky + dy, kz + dz));
                if (u >= 0) { row[cnt] = u; nbr_col[(size_t)cnt * V_cap + v] = u; ++cnt; }
            }
        }
        for (int r = cnt; r < 27; ++r) row[r] = -1;
        row[27] = cnt;
    }
}

// ---- K3: per-voxel normals: computePointNormal over the 2-ring index list with duplicates,
// closed-form eigen33, flip towards the origin, normalise (A.3) -------------------------------
__device__ inline void compute_roots2(float b, float c, float roots[3]) {
    roots[0] = 0.0f;
    float d = (float)((double)(b * b) - 4.0 * (double)c);
    if (d < 0.0f) d = 0.0f;
    float sd = sqrtf(d);
    roots[2] = 0.5f * (b + sd);
    roots[1] = 0.5f * (b - sd);
}
__device__ inline void swapf(float& a, float& b) { float t = a; a = b; b = t; }
__device__ inline void compute_roots(const float m[9], float roots[3]) {
    const float m00 = m[0], m01 = m[1], m02 = m[2], m11 = m[4], m12 = m[5], m22 = m[8];
    float c0 = m00 * m11 * m22 + 2.0f * m01 * m02 * m12 - m00 * m12 * m12 - m11 * m02 * m02 - m22 * m01 * m01;
    float c1 = m00 * m11 - m01 * m01 + m00 * m22 - m02 * m02 + m11 * m22 - m12 * m12;
    float c2 = m00 + m11 + m22;
    if (fabsf(c0) < FLT_EPSILON) { compute_roots2(c2, c1, roots); return; }
    const float s_inv3 = (float)(1.0 / 3.0);
    const float s_sqrt3 = sqrtf(3.0f);
    float c2_over_3 = c2 * s_inv3;
    float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
    if (a_over_3 > 0.0f) a_over_3 = 0.0f;
    float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
    float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
    if (q > 0.0f) q = 0.0f;
    float rho = sqrtf(-a_over_3);
    float theta = cr_atan2f(sqrtf(-q), half_b) * s_inv3;
    float cos_theta = cr_cosf(theta);
    float sin_theta = cr_sinf(theta);
    roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
    roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
    roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
    if (roots[0] >= roots[1]) swapf(roots[0], roots[1]);
    if (roots[1] >= roots[2]) {
        swapf(roots[1], roots[2]);
        if (roots[0] >= roots[1]) swapf(roots[0], roots[1]);
    }
    if (roots[0] <= 0.0f) compute_roots2(c2, c1, roots);
}
__device__ inline void cross3(const float* a, const float* b, float* o) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
// pcl::eigen33 (smallest eigenpair) on cov = xx,xy,xz,yy,yz,zz
__device__ inline void eigen33_smallest(const float cov[6], float& eigenvalue, float evec[3]) {
    float sm[9] = {cov[0], cov[1], cov[2], cov[1], cov[3], cov[4], cov[2], cov[4], cov[5]};
    float scale = 0.0f;
#pragma unroll
    for (int i = 0; i < 9; ++i) scale = fmaxf(scale, fabsf(sm[i]));
    if (scale <= FLT_MIN) scale = 1.0f;
#pragma unroll
    for (int i = 0; i < 9; ++i) sm[i] = sm[i] / scale;
    float roots[3];
    compute_roots(sm, roots);
    eigenvalue = roots[0] * scale;
    sm[0] -= roots[0]; sm[4] -= roots[0]; sm[8] -= roots[0];
    float v1[3], v2[3], v3[3];
    cross3(sm, sm + 3, v1); cross3(sm, sm + 6, v2); cross3(sm + 3, sm + 6, v3);
    float l1 = sum3(v1[0] * v1[0], v1[1] * v1[1], v1[2] * v1[2]);
    float l2 = sum3(v2[0] * v2[0], v2[1] * v2[1], v2[2] * v2[2]);
    float l3 = sum3(v3[0] * v3[0], v3[1] * v3[1], v3[2] * v3[2]);
    const float* v; float l;
    if (l1 >= l2 && l1 >= l3) { v = v1; l = l1; }
    else if (l2 >= l1 && l2 >= l3) { v = v2; l = l2; }
    else { v = v3; l = l3; }
    float s = sqrtf(l);
    evec[0] = v[0] / s; evec[1] = v[1] / s; evec[2] = v[2] / s;
}
// computeMeanAndCovarianceMatrix tail + solvePlaneParameters; accu = raw sums over n samples
__device__ inline void plane_from_accu(const float accu_sum[9], int n, float normal[3], float& curv) {
    float accu[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) accu[i] = accu_sum[i] / (float)n;
    float cov[6];
    cov[0] = accu[0] - accu[6] * accu[6];
    cov[1] = accu[1] - accu[6] * accu[7];
    cov[2] = accu[2] - accu[6] * accu[8];
    cov[3] = accu[3] - accu[7] * accu[7];
    cov[4] = accu[4] - accu[7] * accu[8];
    cov[5] = accu[5] - accu[8] * accu[8];
    float ev;
    eigen33_smallest(cov, ev, normal);
    float eig_sum = cov[0] + cov[3] + cov[5];
    curv = (eig_sum != 0) ? fabsf(ev / eig_sum) : 0.0f;
}
// flipNormalTowardsViewpoint(p, 0,0,0, n) ; n[3]=0 ; normalize()
__device__ inline void flip_and_normalize(float px, float py, float pz, float n[3]) {
    float cos_theta = sum4((0.0f - px) * n[0], (0.0f - py) * n[1], (0.0f - pz) * n[2], 0.0f);
    if (cos_theta < 0) { n[0] *= -1; n[1] *= -1; n[2] *= -1; }
    float z = sum4(n[0] * n[0], n[1] * n[1], n[2] * n[2], 0.0f);
    if (z > 0.0f) { float s = sqrtf(z); n[0] /= s; n[1] /= s; n[2] /= s; }
}

struct Accu9 {
    float a[9];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < 9; ++i) a[i] = 0.0f;
    }
    __device__ __forceinline__ void add(float x, float y, float z) {
        a[0] += x * x; a[1] += x * y; a[2] += x * z; a[3] += y * y; a[4] += y * z; a[5] += z * z;
        a[6] += x; a[7] += y; a[8] += z;
    }
};

__global__ void __launch_bounds__(128) voxel_normals_kernel(const float4* __restrict__ vox_xyz, const int* __restrict__ nbr_row,
        const unsigned* __restrict__ n_vox_ptr, float4* __restrict__ vox_normal, float* __restrict__ vox_curv,
        unsigned v_begin, unsigned v_end) {          // [v_begin, v_end): the voxels this handle owns (slab mode), else everything
    const unsigned V = min(*n_vox_ptr, v_end);
    for (unsigned v = v_begin + blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
        Accu9 A; A.clear();
        int total = 1;
        const float4 pv = vox_xyz[v];
        A.add(pv.x, pv.y, pv.z);
        const int* row = nbr_row + (size_t)v * kNbrStride;
        const int cnt = row[27];
        for (int a = 0; a < cnt; ++a) {
            const int nb = row[a];
            const float4 pn = vox_xyz[nb];
            A.add(pn.x, pn.y, pn.z);
            const int4* rown = reinterpret_cast<const int4*>(nbr_row + (size_t)nb * kNbrStride);
            int r[28];
#pragma unroll
            for (int q = 0; q < 7; ++q) { int4 t = rown[q]; r[4 * q] = t.x; r[4 * q + 1] = t.y; r[4 * q + 2] = t.z; r[4 * q + 3] = t.w; }
            const int cn = r[27];
            total += 1 + cn;
#pragma unroll
            for (int b = 0; b < 27; ++b) {
                if (b < cn) { const float4 p2 = vox_xyz[r[b]]; A.add(p2.x, p2.y, p2.z); }
            }
        }
        float n[3]; float curv;
        if (total < 3) { n[0] = n[1] = n[2] = nanf(""); curv = n[0]; }
        else plane_from_accu(A.a, total, n, curv);
        flip_and_normalize(pv.x, pv.y, pv.z, n);
        vox_normal[v] = make_float4(n[0], n[1], n[2], 0.0f);
        vox_curv[v] = curv;
    }
}

// ---- refineSupervoxels, part 1: SupervoxelHelper::refineNormals (pcl supervoxel_clustering.hpp; call site
// /root/reference/src/supervoxel_clustering.cpp:369-371).  A voxel's normal is recomputed from the index list
// [u + N(u) for u in N(v)], every entry only when its owner is v's owner (N includes the voxel itself; duplicates kept).
// Reads centroids and owners only, so the voxels are independent; unowned voxels keep their normal.  A phantom leaf (held by
// helpers `phantom[3 v + s]` without being owned by them) is recomputed by its owner and by every holder in PCL, each with its own
// filter; the helpers run in label order, so the later of the two decides -- the holder when a third helper with a smaller
// label has stolen the voxel from its first owner.
__global__ void __launch_bounds__(128) refine_normals_kernel(const float4* __restrict__ vox_xyz, const int* __restrict__ nbr_row,
        const unsigned* __restrict__ n_vox_ptr, const unsigned* __restrict__ label, const unsigned* __restrict__ phantom,
        float4* __restrict__ vox_normal, float* __restrict__ vox_curv) {
    const unsigned V = *n_vox_ptr;
    for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
        unsigned L = label[v];
        if (L == 0u) continue;
        for (int s = 0; s < 3; ++s) { const unsigned ph = phantom[(size_t)3 * v + s]; if (ph > L) L = ph; }   // kPhSlots holders (kernels_expand.cuh)
        Accu9 A; A.clear();
        int total = 0;
        const float4 pv = vox_xyz[v];
        const int* row = nbr_row + (size_t)v * kNbrStride;
        const int cnt = row[27];
        for (int a = 0; a < cnt; ++a) {
            const int nb = row[a];
            if (label[nb] != L) continue;
            const float4 pn = vox_xyz[nb];
            A.add(pn.x, pn.y, pn.z); ++total;
            const int* rown = nbr_row + (size_t)nb * kNbrStride;
            const int cn = rown[27];
            for (int b = 0; b < cn; ++b) {
                const int w = rown[b];
                if (label[w] == L) { const float4 p2 = vox_xyz[w]; A.add(p2.x, p2.y, p2.z); ++total; }
            }
        }
        float n[3]; float curv;
        if (total < 3) { n[0] = n[1] = n[2] = nanf(""); curv = n[0]; }
        else plane_from_accu(A.a, total, n, curv);
        flip_and_normalize(pv.x, pv.y, pv.z, n);
        vox_normal[v] = make_float4(n[0], n[1], n[2], 0.0f);
        vox_curv[v] = curv;
    }
}

// ---- refineSupervoxels, part 2: reseedSupervoxels.  Every surviving helper (centroid count > 0) takes the voxel nearest to its
// centroid as its only leaf (voxel_kdtree_->nearestKSearch(point, 1): exact 1-NN under flann::L2_Simple, ties to the lowest
// index); an erased helper gets -1.  One block per helper, brute force over the voxel centroids (S x V distance evaluations).
__global__ void __launch_bounds__(256) reseed_kernel(const float4* __restrict__ cen_xyz, unsigned S0, const float4* __restrict__ vox_xyz,
        const unsigned* __restrict__ n_vox_ptr, int* __restrict__ seeds) {
    __shared__ unsigned long long s_best[8];
    const unsigned V = *n_vox_ptr;
    for (unsigned i = blockIdx.x; i < S0; i += gridDim.x) {
        const float4 c = cen_xyz[i + 1];
        if (!(c.w > 0.0f)) { if (threadIdx.x == 0) seeds[i] = -1; continue; }
        unsigned long long best = ~0ull;                                   // (distance bits : voxel index), distances are >= 0
        for (unsigned u = threadIdx.x; u < V; u += blockDim.x) {
            const float4 p = vox_xyz[u];
            float r = 0.0f;
            { const float d0 = c.x - p.x; r += d0 * d0; const float d1 = c.y - p.y; r += d1 * d1; const float d2 = c.z - p.z; r += d2 * d2; }
            const unsigned long long k = ((unsigned long long)__float_as_uint(r) << 32) | u;
            if (!(r != r) && k < best) best = k;
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) { const unsigned long long o = __shfl_xor_sync(kFull, best, off); if (o < best) best = o; }
        if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = best;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; ++w) if (s_best[w] < best) best = s_best[w];
            seeds[i] = best == ~0ull ? -1 : (int)(unsigned)best;
        }
        __syncthreads();
    }
}

// ---- K4: seed selection (A.4) ---------------------------------------------------------------
// Seed octree box growth (OctreePointCloud::adoptBoundingBoxToPoint) over the voxel centroids in
// idx order.  One block; the cursor only moves forward, so the centroids are read once.
__global__ void __launch_bounds__(1024) seed_box_kernel(const float4* __restrict__ vox_xyz, const unsigned* __restrict__ n_vox_ptr,
                                                        float seed_res, SeedBox* __restrict__ sb) {
    __shared__ double s_mn[3], s_mx[3];
    __shared__ unsigned s_first;
    __shared__ int s_depth;
    __shared__ long long s_off[3];
    const unsigned V = *n_vox_ptr;
    const double res = (double)seed_res;
    const float minValue = FLT_EPSILON;
    if (threadIdx.x == 0) {
        sb->n_events = 0; sb->res = res; sb->depth = 0;
        for (int a = 0; a < 3; ++a) { s_off[a] = 0; sb->off_final[a] = 0; sb->mn_final[a] = 0; }
        if (V > 0) {
            const float4 p0 = vox_xyz[0];
            const float p[3] = {p0.x, p0.y, p0.z};
            double mn[3], mx[3];
            for (int a = 0; a < 3; ++a) { mn[a] = (double)p[a] - res / 2; mx[a] = (double)p[a] + res / 2; }
            s_depth = key_bit_size(mn, mx, res);
            for (int a = 0; a < 3; ++a) { s_mn[a] = mn[a]; s_mx[a] = mx[a]; }
            SeedEvent& e = sb->ev[0];
            for (int a = 0; a < 3; ++a) { e.mn[a] = mn[a]; e.off[a] = 0; }
            e.first = 0; e.depth = s_depth; sb->n_events = 1;
        }
    }
    __syncthreads();
    if (V == 0) return;
    unsigned cursor = 1;
    while (cursor < V) {
        // first voxel >= cursor outside the current box, searched in chunks of 8 * blockDim
        if (threadIdx.x == 0) s_first = 0xffffffffu;
        __syncthreads();
        const double mn0 = s_mn[0], mn1 = s_mn[1], mn2 = s_mn[2], mx0 = s_mx[0], mx1 = s_mx[1], mx2 = s_mx[2];
        unsigned chunk_end = cursor;
        while (chunk_end < V) {
            unsigned mine = 0xffffffffu;
            const unsigned lim = min(V, chunk_end + 8u * blockDim.x);
            for (unsigned i = chunk_end + threadIdx.x; i < lim; i += blockDim.x) {
                const float4 p = vox_xyz[i];
                bool out = (double)p.x < mn0 || (double)p.x >= mx0 || (double)p.y < mn1 || (double)p.y >= mx1 ||
                           (double)p.z < mn2 || (double)p.z >= mx2;
                if (out) { mine = i; break; }
            }
            if (mine != 0xffffffffu) atomicMin(&s_first, mine);
            __syncthreads();
            chunk_end = lim;
            if (s_first != 0xffffffffu) break;
            __syncthreads();
        }
        __syncthreads();
        const unsigned first = s_first;
        __syncthreads();
        if (first == 0xffffffffu) break;
        if (threadIdx.x == 0) {
            const float4 p4 = vox_xyz[first];
            const float p[3] = {p4.x, p4.y, p4.z};
            int d = s_depth;
            double mn[3] = {s_mn[0], s_mn[1], s_mn[2]}, mx[3] = {s_mx[0], s_mx[1], s_mx[2]};
            while (true) {
                bool lo[3], hi[3], viol = false;
                for (int a = 0; a < 3; ++a) { lo[a] = (double)p[a] < mn[a]; hi[a] = (double)p[a] >= mx[a]; viol = viol || lo[a] || hi[a]; }
                if (!viol) break;
                double side = (double)(1ull << d) * res;
                for (int a = 0; a < 3; ++a) if (!hi[a]) { mn[a] -= side; s_off[a] += (1ll << d); }
                ++d;
                side = (double)(1ull << d) * res - minValue;
                for (int a = 0; a < 3; ++a) mx[a] = mn[a] + side;
            }
            s_depth = d;
            for (int a = 0; a < 3; ++a) { s_mn[a] = mn[a]; s_mx[a] = mx[a]; }
            int k = sb->n_events;
            if (k < kMaxSeedEvents) {
                SeedEvent& e = sb->ev[k];
                for (int a = 0; a < 3; ++a) { e.mn[a] = mn[a]; e.off[a] = s_off[a]; }
                e.first = (int)first; e.depth = d; sb->n_events = k + 1;
            } else sb->n_events = kMaxSeedEvents + 1;   // overflow marker
        }
        cursor = first + 1;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        sb->depth = s_depth;
        for (int a = 0; a < 3; ++a) { sb->off_final[a] = s_off[a]; sb->mn_final[a] = s_mn[a]; }
    }
}

// per-voxel seed cell code: key at insertion time under the origin of its epoch, moved into the final frame
__global__ void __launch_bounds__(256) seed_cell_kernel(const float4* __restrict__ vox_xyz, const unsigned* __restrict__ n_vox_ptr,
        const SeedBox* __restrict__ sb, uint64_t* __restrict__ cell_code) {
    __shared__ SeedBox s_sb;
    for (int i = threadIdx.x; i < (int)(sizeof(SeedBox) / 4); i += blockDim.x) ((int*)&s_sb)[i] = ((const int*)sb)[i];
    __syncthreads();
    const unsigned V = *n_vox_ptr;
    const int ne = min(s_sb.n_events, kMaxSeedEvents);
    for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
        int e = 0;
        for (int k = 1; k < ne; ++k) if ((unsigned)s_sb.ev[k].first <= v) e = k;
        const float4 p4 = vox_xyz[v];
        const float p[3] = {p4.x, p4.y, p4.z};
        uint32_t k3[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            long long kin = (long long)(unsigned)(((double)p[a] - s_sb.ev[e].mn[a]) / s_sb.res);
            k3[a] = (uint32_t)(kin - s_sb.ev[e].off[a] + s_sb.off_final[a]);
        }
        cell_code[v] = morton_xmajor(k3[0], k3[1], k3[2]);
    }
}

__device__ __forceinline__ float l2_simple(float ax, float ay, float az, float bx, float by, float bz) {
    float r = 0.0f, d;                       // flann::L2_Simple<float>
    d = ax - bx; r += d * d;
    d = ay - by; r += d * d;
    d = az - bz; r += d * d;
    return r;
}
__device__ __forceinline__ int find_cell(const uint64_t* __restrict__ codes, int n, uint64_t code) {
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (codes[mid] < code) lo = mid + 1; else hi = mid; }
    return (lo < n && codes[lo] == code) ? lo : -1;
}

// gather the unique cell codes at the run heads
__global__ void __launch_bounds__(256) gather_cell_codes_kernel(const uint64_t* __restrict__ sorted_code, const unsigned* __restrict__ cell_start,
        const unsigned* __restrict__ n_cells_ptr, uint64_t* __restrict__ cell_codes) {
    const unsigned C = *n_cells_ptr;
    for (unsigned c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) cell_codes[c] = sorted_code[cell_start[c]];
}

// one warp per occupied seed cell: exact 1-NN voxel of the cell centre, then the radius filter
__global__ void __launch_bounds__(256) seed_select_kernel(const float4* __restrict__ vox_xyz, const uint64_t* __restrict__ vox_cell,
        const unsigned* __restrict__ sorted_vox, const unsigned* __restrict__ cell_start, const uint64_t* __restrict__ cell_codes,
        const unsigned* __restrict__ n_cells_ptr, const unsigned* __restrict__ n_vox_ptr, const SeedBox* __restrict__ sb,
        float seed_res, float voxel_res, int* __restrict__ cell_nn, unsigned* __restrict__ cell_keep) {
    const int C = (int)*n_cells_ptr;
    const unsigned V = *n_vox_ptr;
    const int lane = threadIdx.x & 31;
    const int warps_total = (gridDim.x * blockDim.x) >> 5;
    const int depth = sb->depth;
    const long long kmax = 1ll << depth;
    const double res = sb->res;
    const float search_radius = 0.5f * seed_res;
    const float min_points = 0.05f * (search_radius) * (search_radius) * 3.1415926536f / (voxel_res * voxel_res);
    const float r2 = (float)((double)search_radius * (double)search_radius);
    for (int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < C; c += warps_total) {
        uint32_t k[3];
        morton_decode(cell_codes[c], k[0], k[1], k[2]);
        float ctr[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) ctr[a] = (float)(((double)k[a] + 0.5f) * res + sb->mn_final[a]);   // genLeafNodeCenterFromOctreeKey
        float bd = FLT_MAX; unsigned best = 0xffffffffu;
        // the 27 cell look-ups (a binary search each) run side by side, one per lane, instead of one after the other in every lane
        int my_cc = -1;
        if (lane < 27) {
            const long long x = (long long)k[0] + (lane / 9 - 1), y = (long long)k[1] + ((lane / 3) % 3 - 1), z = (long long)k[2] + (lane % 3 - 1);
            if (!(x < 0 || y < 0 || z < 0 || x >= kmax || y >= kmax || z >= kmax)) my_cc = find_cell(cell_codes, C, morton_xmajor((uint32_t)x, (uint32_t)y, (uint32_t)z));
        }
        for (int q = 0; q < 27; ++q) {
            const int cc = __shfl_sync(kFull, my_cc, q);
            if (cc < 0) continue;
            unsigned s = cell_start[cc], e = (cc + 1 < C) ? cell_start[cc + 1] : V;
            for (unsigned j = s + lane; j < e; j += 32) {
                unsigned u = sorted_vox[j];
                float4 p = vox_xyz[u];
                float dd = l2_simple(ctr[0], ctr[1], ctr[2], p.x, p.y, p.z);
                if (dd < bd || (dd == bd && u < best)) { bd = dd; best = u; }
            }
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) {
            float od = __shfl_xor_sync(kFull, bd, off); unsigned ou = __shfl_xor_sync(kFull, best, off);
            if (ou != 0xffffffffu && (best == 0xffffffffu || od < bd || (od == bd && ou < best))) { bd = od; best = ou; }
        }
        // radius search around the candidate voxel
        const float4 pb = vox_xyz[best];
        uint32_t kb[3];
        morton_decode(vox_cell[best], kb[0], kb[1], kb[2]);
        int num = 0;
        my_cc = -1;
        if (lane < 27) {
            const long long x = (long long)kb[0] + (lane / 9 - 1), y = (long long)kb[1] + ((lane / 3) % 3 - 1), z = (long long)kb[2] + (lane % 3 - 1);
            if (!(x < 0 || y < 0 || z < 0 || x >= kmax || y >= kmax || z >= kmax)) my_cc = find_cell(cell_codes, C, morton_xmajor((uint32_t)x, (uint32_t)y, (uint32_t)z));
        }
        for (int q = 0; q < 27; ++q) {
            const int cc = __shfl_sync(kFull, my_cc, q);
            if (cc < 0) continue;
            unsigned s = cell_start[cc], e = (cc + 1 < C) ? cell_start[cc + 1] : V;
            for (unsigned j = s + lane; j < e; j += 32) {
                float4 p = vox_xyz[sorted_vox[j]];
                if (l2_simple(pb.x, pb.y, pb.z, p.x, p.y, p.z) < r2) ++num;
            }
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) num += __shfl_xor_sync(kFull, num, off);
        if (lane == 0) { cell_nn[c] = (int)best; cell_keep[c] = ((float)num > min_points) ? 1u : 0u; }
    }
}

struct KeepOp {
    const unsigned* keep; const int* cell_nn; int* seeds;
    typedef int Payload;
    __device__ __forceinline__ bool test(int64_t i, Payload&) const { return keep[i] != 0; }
    __device__ __forceinline__ void emit(unsigned pos, int64_t i, const Payload&) const { seeds[pos] = cell_nn[i]; }
};

// ---- K5 helpers shared with kernels_expand.cuh (A.5) ---------------------------------------
struct Centroids {            // per label (index = label, 0 unused)
    float4* xyz;              // mean xyz, w = voxel count
    float4* rgb;              // mean rgb
    float4* nrm;              // normalised normal sum (4-vector, w stays 0)
};

constexpr unsigned kNoSteal = 0xffffffffu;

__device__ __forceinline__ float voxel_data_distance(const float4& cx, const float4& cc, const float4& cn, const float4& vx,
                                                     const float4& vc, const float4& vn, const VccsParams& P) {
    float dx0 = cx.x - vx.x, dx1 = cx.y - vx.y, dx2 = cx.z - vx.z;
    float dc0 = cc.x - vc.x, dc1 = cc.y - vc.y, dc2 = cc.z - vc.z;
    float spatial = sqrtf(sum3(dx0 * dx0, dx1 * dx1, dx2 * dx2)) / P.seed_res;
    float color = sqrtf(sum3(dc0 * dc0, dc1 * dc1, dc2 * dc2)) / 255.0f;
    float cosang = 1.0f - fabsf(sum4(cn.x * vn.x, cn.y * vn.y, cn.z * vn.z, cn.w * vn.w));
    return cosang * P.normal_imp + color * P.color_imp + spatial * P.spatial_imp;
}

} // namespace f3ps
