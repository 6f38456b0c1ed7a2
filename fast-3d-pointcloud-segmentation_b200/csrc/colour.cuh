// colour.cuh -- device versions of the reference's metric kernels:
//   ColorUtilities::rgb2lab  (src/color_utilities.cpp:151-160 -> cv::cvtColor float RGB2Lab)
//   ColorUtilities::lab_ciede00 (:190-294), rgb_eucl (:304-319)
//   Clustering::normals_diff (src/clustering.cpp:79-96), is_convex (:53-67)
//   Clustering::t_c / t_g (:324-376)
// Same float/double mix per expression as the reference; compiled with -fmad=false.
#pragma once
#include "common.cuh"
#include "ciede_fast.h"

namespace f3ps {

#define F3PS_RGB_RANGE 441.672943f   // include/supervoxel_clustering/color_utilities.h:62
#define F3PS_LAB_RANGE 137.3607f     // :63
#define F3PS_PI 3.14159265358979323846   // M_PI

// OpenCV's float sRGB->Lab: 33^3 fixed-point lattice + integer trilinear interpolation
// (SURVEY.md Appendix B).  lut = [33][33][33][3] int16, L2-resident (215,622 B).
__device__ inline void rgb2lab(const short* __restrict__ lut, float r255, float g255, float b255, float lab[3]) {
    const float in[3] = {r255, g255, b255};
    int t[3], f[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float v = in[k] / 255;
        v = fminf(fmaxf(v, 0.0f), 1.0f);
        int c = (int)rintf(v * 16384.0f);
        t[k] = c >> 9; f[k] = (c >> 5) & 15;
    }
    int out[3] = {0, 0, 0};
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int dr = q >> 2, dg = (q >> 1) & 1, db = q & 1;
        const int w = (dr ? f[0] : 16 - f[0]) * (dg ? f[1] : 16 - f[1]) * (db ? f[2] : 16 - f[2]);
        const int ir = min(t[0] + dr, 32), ig = min(t[1] + dg, 32), ib = min(t[2] + db, 32);
        const short* e = lut + ((ir * 33 + ig) * 33 + ib) * 3;
        out[0] += w * (int)e[0]; out[1] += w * (int)e[1]; out[2] += w * (int)e[2];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) out[k] = (out[k] + 2048) >> 12;
    lab[0] = ((float)out[0] / 16384.0f) * 100.0f;
    lab[1] = ((float)out[1] / 16384.0f) * 256.0f - 128.0f;
    lab[2] = ((float)out[2] / 16384.0f) * 256.0f - 128.0f;
}

// CIEDE2000 (src/color_utilities.cpp:190-294): branch-free FP64 evaluation shared with the host-side accuracy tests
__device__ __forceinline__ float lab_ciede00(const float lab1[3], const float lab2[3]) { return f3ps_fastmath::ciede00(lab1, lab2); }

__device__ inline float rgb_eucl(const float c1[3], const float c2[3]) {
    const float d0 = c1[0] - c2[0], d1 = c1[1] - c2[1], d2 = c1[2] - c2[2];
    const float rd = d0 * d0, gd = d1 * d1, bd = d2 * d2;   // (float)pow((double)d, 2): exact square, rounded once
    return sqrtf(rd + gd + bd);
}

__device__ inline float normals_diff(const float n1[3], const float c1[3], const float n2[3], const float c2[3]) {
    float C[3] = {c1[0] - c2[0], c1[1] - c2[1], c1[2] - c2[2]};
    const float nrm = sqrtf(sum3(C[0] * C[0], C[1] * C[1], C[2] * C[2]));
    C[0] /= nrm; C[1] /= nrm; C[2] /= nrm;
    const float x0 = n1[1] * n2[2] - n1[2] * n2[1], x1 = n1[2] * n2[0] - n1[0] * n2[2], x2 = n1[0] * n2[1] - n1[1] * n2[0];
    const float N1xN2 = sqrtf(sum3(x0 * x0, x1 * x1, x2 * x2));
    const float N1_C = fabsf(sum3(n1[0] * C[0], n1[1] * C[1], n1[2] * C[2]));
    const float N2_C = fabsf(sum3(n2[0] * C[0], n2[1] * C[1], n2[2] * C[2]));
    return (N1xN2 + N1_C + N2_C) / 3;
}
__device__ inline bool is_convex(const float n1[3], const float c1[3], const float n2[3], const float c2[3]) {
    float C[3] = {c1[0] - c2[0], c1[1] - c2[1], c1[2] - c2[2]};
    const float nrm = sqrtf(sum3(C[0] * C[0], C[1] * C[1], C[2] * C[2]));
    C[0] /= nrm; C[1] /= nrm; C[2] /= nrm;
    const float cos1 = sum3(n1[0] * C[0], n1[1] * C[1], n1[2] * C[2]);
    const float cos2 = sum3(n2[0] * C[0], n2[1] * C[1], n2[2] * C[2]);
    return cos1 >= cos2;
}

// region statistics kept per graph node: every sequential accumulator of the
// reference over a region's voxel list, resumable when another list is appended.
struct RegionStats {
    float cnt, r, g, b;      // ColorUtilities::mean_color running mean (count kept as float, as there)
    float accu[9];           // computeMeanAndCovarianceMatrix raw sums; accu[6..8] == CentroidPoint xyz sums
    int n;
};
__device__ __forceinline__ void stats_step(RegionStats& s, float x, float y, float z, uint32_t rgba) {
    const float r = (float)((rgba >> 16) & 255u), g = (float)((rgba >> 8) & 255u), b = (float)(rgba & 255u);
    s.cnt = s.cnt + 1.0f;
    const float inv = 1 / s.cnt;
    s.r = s.r + inv * (r - s.r);
    s.g = s.g + inv * (g - s.g);
    s.b = s.b + inv * (b - s.b);
    s.accu[0] += x * x; s.accu[1] += x * y; s.accu[2] += x * z; s.accu[3] += y * y; s.accu[4] += y * z; s.accu[5] += z * z;
    s.accu[6] += x; s.accu[7] += y; s.accu[8] += z;
    s.n++;
}

struct EdgeParams {          // frozen at init_weights
    int color_mode, geom_mode, merge_mode, bins;
    float lambda;
    const float* cdf_c; const float* cdf_g;
    const short* lab_lut;
};

// Clustering::delta_c_g (src/clustering.cpp:107-142) on two regions' cached statistics
__device__ inline void delta_c_g(const EdgeParams& ep, const float rgb1[3], const float rgb2[3], const float n1[3], const float c1[3],
                                 const float n2[3], const float c2[3], float& delta_c, float& delta_g) {
    if (ep.color_mode == 0) {
        float lab1[3], lab2[3];
        rgb2lab(ep.lab_lut, rgb1[0], rgb1[1], rgb1[2], lab1);
        rgb2lab(ep.lab_lut, rgb2[0], rgb2[1], rgb2[2], lab2);
        delta_c = lab_ciede00(lab1, lab2);
        delta_c /= F3PS_LAB_RANGE;
    } else {
        delta_c = rgb_eucl(rgb1, rgb2);
        delta_c /= F3PS_RGB_RANGE;
    }
    delta_g = normals_diff(n1, c1, n2, c2);
    if (ep.geom_mode == 1 && is_convex(n1, c1, n2, c2)) delta_g *= 0.5f;
}
// The same on the regions' cached colour vectors (cvec = Lab of the mean colour under LAB_CIEDE00, the mean colour
// itself under RGB_EUCL): rgb2lab is a pure function of the mean, so it is evaluated once per region state.
__device__ inline void delta_cached(const EdgeParams& ep, const float cv1[3], const float cv2[3], const float n1[3], const float c1[3],
                                    const float n2[3], const float c2[3], float& delta_c, float& delta_g) {
    if (ep.color_mode == 0) { delta_c = lab_ciede00(cv1, cv2); delta_c /= F3PS_LAB_RANGE; }
    else { delta_c = rgb_eucl(cv1, cv2); delta_c /= F3PS_RGB_RANGE; }
    delta_g = normals_diff(n1, c1, n2, c2);
    if (ep.geom_mode == 1 && is_convex(n1, c1, n2, c2)) delta_g *= 0.5f;
}
__device__ inline void colour_vector(const EdgeParams& ep, float r, float g, float b, float cv[3]) {
    if (ep.color_mode == 0) rgb2lab(ep.lab_lut, r, g, b, cv);
    else { cv[0] = r; cv[1] = g; cv[2] = b; }
}
// Clustering::t_c + t_g (:324-376).  Out-of-range bins (NaN deltas; the reference throws from
// map::at there) are clamped and flagged by the caller through the NaN weight they produce.
__device__ inline float unify(const EdgeParams& ep, float delta_c, float delta_g) {
    if (ep.merge_mode == 2) {
        short bc = (short)floorf(delta_c * ep.bins);
        if (bc == ep.bins) bc--;
        short bg = (short)floorf(delta_g * ep.bins);
        if (bc < 0 || bc >= ep.bins || bg < 0 || bg >= ep.bins) return nanf("");
        return ep.cdf_c[bc] / 2 + ep.cdf_g[bg] / 2;
    }
    return ep.lambda * delta_c + (1 - ep.lambda) * delta_g;
}

} // namespace f3ps
