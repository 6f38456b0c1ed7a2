// colour.cuh -- device versions of the reference's metric kernels:
//   ColorUtilities::rgb2lab  (src/color_utilities.cpp:151-160 -> cv::cvtColor float RGB2Lab)
//   ColorUtilities::lab_ciede00 (:190-294), rgb_eucl (:304-319)
//   Clustering::normals_diff (src/clustering.cpp:79-96), is_convex (:53-67)
//   Clustering::t_c / t_g (:324-376)
// Same float/double mix per expression as the reference; compiled with -fmad=false.
#pragma once
#include "common.cuh"

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

__device__ inline float lab_ciede00(const float lab1[3], const float lab2[3]) {
    const double kL = 1.0, kC = 1.0, kH = 1.0;
    const float L1 = lab1[0], a1 = lab1[1], b1 = lab1[2];
    const float L2 = lab2[0], a2 = lab2[1], b2 = lab2[2];
    const double Cab1 = (double)sqrtf(a1 * a1 + b1 * b1);
    const double Cab2 = (double)sqrtf(a2 * a2 + b2 * b2);
    const double Cab = (Cab1 + Cab2) / 2.0;
    const double p25_7 = 6103515625.0;                       // pow(25.0, 7.0), exact
    const double Cab7 = pow(Cab, 7.0);
    const double G = 0.5 * (1.0 - sqrt(Cab7 / (Cab7 + p25_7)));
    const double ap1 = (1.0 + G) * (double)a1;
    const double ap2 = (1.0 + G) * (double)a2;
    const double Cp1 = sqrt(ap1 * ap1 + (double)(b1 * b1));
    const double Cp2 = sqrt(ap2 * ap2 + (double)(b2 * b2));
    const double Cp_prod = (Cp2 * Cp1);
    double hp1 = 0;
    if ((fabs(ap1) + (double)fabsf(b1)) != 0.0) {
        hp1 = atan2((double)b1, ap1);
        if (hp1 < 0) hp1 += 2.0 * F3PS_PI;
    }
    double hp2 = 0;
    if ((fabs(ap2) + (double)fabsf(b2)) != 0.0) {
        hp2 = atan2((double)b2, ap2);
        if (hp2 < 0) hp2 += 2.0 * F3PS_PI;
    }
    const double dL = (double)(L2 - L1);
    const double dC = (Cp2 - Cp1);
    double dhp = (hp2 - hp1);
    if (dhp > F3PS_PI) dhp -= 2.0 * F3PS_PI;
    else if (dhp < -F3PS_PI) dhp += 2.0 * F3PS_PI;
    if (Cp_prod == 0.0) dhp = 0.0;
    const double dH = 2.0 * sqrt(Cp_prod) * sin(dhp / 2.0);
    const double Lp = (double)(L2 + L1) / 2.0;
    const double Cp = (Cp1 + Cp2) / 2.0;
    double hp = (hp1 + hp2) / 2.0;
    if (fabs(hp1 - hp2) > F3PS_PI) hp -= F3PS_PI;
    if (hp < 0) hp += 2.0 * F3PS_PI;
    if (Cp_prod == 0.0) hp = hp1 + hp2;
    const double Lpm502 = (Lp - 50.0) * (Lp - 50.0);
    const double T = 1.0 - 0.17 * cos(hp - F3PS_PI / 6.0) + 0.24 * cos(2.0 * hp)
                   + 0.32 * cos(3.0 * hp + F3PS_PI / 30.0) - 0.20 * cos(4.0 * hp - 63.0 * F3PS_PI / 180.0);
    const double hq = ((180.0 / F3PS_PI * hp - 275.0) / 25.0);
    const double dheta_rad = (30.0 * F3PS_PI / 180.0) * exp(-(hq * hq));       // pow(x, 2.0) == x*x rounded
    const double Cp7 = pow(Cp, 7.0);
    const double Rc = 2.0 * sqrt(Cp7 / (Cp7 + p25_7));
    const double kLSL = kL * (1.0 + 0.015 * Lpm502 / sqrt(20.0 + Lpm502));
    const double kLSC = kC * (1.0 + 0.045 * Cp);
    const double kHSH = kH * (1.0 + 0.015 * Cp * T);
    const double RT = -sin(2.0 * dheta_rad) * Rc;
    const double tL = dL / kLSL, tC = dC / kLSC, tH = dH / kHSH;
    return (float)sqrt(tL * tL + tC * tC + tH * tH + RT * tC * tH);
}

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
