// common.cuh -- shared device helpers for the f3ps kernels (sm_100a).
//
// Numerical contract (DESIGN.md "Bit-exactness"): every kernel that feeds the merge
// order is compiled with -fmad=false and evaluates the reference's float expressions
// in the reference's association order, so results are the same IEEE-754 sequence a
// scalar CPU evaluation produces.  Float libm calls of the reference (logf, atan2f,
// cosf, sinf inside PCL) are evaluated in double and rounded once ("correctly
// rounded model").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>
#include <math.h>
#include "ciede_fast.h"

namespace f3ps {

constexpr int kSMs = 148;                 // B200: 2 dies x 74 SMs
constexpr unsigned kFull = 0xffffffffu;

#define F3PS_CUDA_OK(expr)                                                              \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) return ctx_fail_cuda(ctx, _e, #expr, __FILE__, __LINE__); \
    } while (0)

// Eigen 3.3 fixed-size reductions: 3-vector a0+(a1+a2); Vector4f (SSE3 hadd) (a0+a1)+(a2+a3)
__device__ __forceinline__ float sum3(float a0, float a1, float a2) { return a0 + (a1 + a2); }
__device__ __forceinline__ float sum4(float a0, float a1, float a2, float a3) { return (a0 + a1) + (a2 + a3); }

// correctly rounded float libm model
__device__ __forceinline__ float cr_logf(float x) { return (float)log((double)x); }
// (branch-free FP64 kernels of ciede_fast.h, <= 2 ulp(double): the rounded float is the same except within ~2^-28 of a tie)
__device__ __forceinline__ float cr_atan2f(float y, float x) {
    if (y == 0.0f && x == 0.0f) return 0.0f;                 // atan2(+0, +0)
    return (float)f3ps_fastmath::atan2_fast((double)y, (double)x);
}
__device__ __forceinline__ float cr_cosf(float x) { return (float)f3ps_fastmath::cos_fast((double)x); }
__device__ __forceinline__ float cr_sinf(float x) { return (float)f3ps_fastmath::sin_fast((double)x); }

__device__ __forceinline__ bool finite3(float x, float y, float z) { return isfinite(x) && isfinite(y) && isfinite(z); }

// order-preserving float <-> uint encoding for atomicMin/Max
__device__ __forceinline__ unsigned f2ord(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(unsigned u) {
    unsigned v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(v);
#else
    float f; memcpy(&f, &v, 4); return f;
#endif
}

// Morton code with x as the most significant interleaved bit (PCL child index = x<<2|y<<1|z)
__host__ __device__ __forceinline__ uint64_t spread3(uint32_t v) {   // 21 bits -> every 3rd bit
    uint64_t x = v & 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
__host__ __device__ __forceinline__ uint32_t compact3(uint64_t x) {
    x &= 0x1249249249249249ull;
    x = (x ^ (x >> 2)) & 0x10c30c30c30c30c3ull;
    x = (x ^ (x >> 4)) & 0x100f00f00f00f00full;
    x = (x ^ (x >> 8)) & 0x1f0000ff0000ffull;
    x = (x ^ (x >> 16)) & 0x1f00000000ffffull;
    x = (x ^ (x >> 32)) & 0x1fffffull;
    return (uint32_t)x;
}
__host__ __device__ __forceinline__ uint64_t morton_xmajor(uint32_t x, uint32_t y, uint32_t z) {
    return (spread3(x) << 2) | (spread3(y) << 1) | spread3(z);
}
__host__ __device__ __forceinline__ void morton_decode(uint64_t m, uint32_t& x, uint32_t& y, uint32_t& z) {
    x = compact3(m >> 2); y = compact3(m >> 1); z = compact3(m);
}

constexpr uint64_t kInvalidKey = ~0ull;

// frame-level parameters produced on the device by the bounding-box pass
struct FrameParams {
    double bmin[3];        // centred cube origin (transformed space)
    double res;            // (double)(float)Rv
    int depth;             // adjacency octree depth
    int any_finite;        // 0 -> empty frame
    unsigned ord_min[3];   // encoded float min/max accumulators
    unsigned ord_max[3];
    int pad[2];
};

struct SeedEvent {         // one growth step of the seed octree (adoptBoundingBoxToPoint)
    double mn[3];          // box origin valid from voxel `first` on
    long long off[3];      // cells added below the very first origin so far
    int first;             // first voxel index inserted under this origin
    int depth;
};
constexpr int kMaxSeedEvents = 48;
struct SeedBox {
    SeedEvent ev[kMaxSeedEvents];
    int n_events;
    int depth;             // final depth
    double res;
    long long off_final[3];
    double mn_final[3];
};

struct VccsParams {
    float voxel_res, seed_res, color_imp, spatial_imp, normal_imp;
    int use_transform, fold_negative_z;
};
struct MergeParams {
    int color_mode, geom_mode, merge_mode;
    float lambda;
    int bins;
};

} // namespace f3ps
