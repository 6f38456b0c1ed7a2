// Host build of the product's branch-free CIEDE2000 (csrc/ciede_fast.h) for the CPU-side accuracy tests.
#include "../../fast-3d-pointcloud-segmentation_b200/csrc/ciede_fast.h"
extern "C" {
void cf_ciede_batch(const float* lab1, const float* lab2, float* out, long n) {
    for (long i = 0; i < n; ++i) out[i] = f3ps_fastmath::ciede00(lab1 + 3 * i, lab2 + 3 * i);
}
void cf_sincos_batch(const double* x, double* s, double* c, long n) { for (long i = 0; i < n; ++i) f3ps_fastmath::sincos_fast(x[i], s[i], c[i]); }
void cf_atan2_batch(const double* y, const double* x, double* o, long n) { for (long i = 0; i < n; ++i) o[i] = f3ps_fastmath::atan2_fast(y[i], x[i]); }
void cf_exp_batch(const double* x, double* o, long n) { for (long i = 0; i < n; ++i) o[i] = f3ps_fastmath::exp_fast(x[i]); }
}
