// Micro-probe: FP64 latency / throughput on one SM and the cost of one CIEDE2000 evaluation (development aid).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../fast-3d-pointcloud-segmentation_b200/csrc/ciede_fast.h"

__global__ void dep_dfma(double* out, long long* cyc, int n) {
    double a = out[0], b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) a = fma(a, b, c);
    long long t1 = clock64();
    out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void ind_dfma(double* out, long long* cyc, int n) {
    double a0 = out[0], a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7, b = 1.0000001, c = 1e-9;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c); }
    __syncthreads();
    long long t1 = clock64();
    out[threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void dep_ffma(float* out, long long* cyc, int n) {
    float a = out[0], b = 1.0000001f, c = 1e-9f;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) a = fmaf(a, b, c);
    long long t1 = clock64();
    out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void ciede_lat(const float* lab, float* out, long long* cyc, int reps) {
    float l1[3] = {lab[0], lab[1], lab[2]}, l2[3] = {lab[3], lab[4], lab[5]};
    float acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < reps; ++i) { const float e = acc * 1e-9f; l1[0] += e; l1[1] += e; l1[2] -= e; l2[0] -= e; l2[1] += e; l2[2] += e; acc += f3ps_fastmath::ciede00(l1, l2); }
    long long t1 = clock64();
    out[threadIdx.x] = acc; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void ddiv_lat(double* out, long long* cyc, int n) {
    double a = out[0] + 3.0, b = 1.0000001;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) a = a / b;
    long long t1 = clock64();
    double s = a;
    long long t2 = clock64();
    for (int i = 0; i < n; ++i) s = sqrt(s) + 1.0;
    long long t3 = clock64();
    out[threadIdx.x] = a + s; if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t3 - t2; }
}
int main() {
    double* d; long long* c; float* f; float* lab;
    cudaMalloc(&d, 8192 * 8); cudaMalloc(&c, 64); cudaMalloc(&f, 8192 * 4); cudaMalloc(&lab, 24);
    cudaMemset(d, 0, 8192 * 8); cudaMemset(f, 0, 8192 * 4);
    float hl[6] = {52.3f, 14.25f, -33.5f, 48.1f, 10.0f, -30.25f};
    cudaMemcpy(lab, hl, 24, cudaMemcpyHostToDevice);
    long long h[2]; const int n = 4096;
    for (int threads : {32, 128, 256, 1024}) {
        dep_dfma<<<1, threads>>>(d, c, n); cudaMemcpy(h, c, 16, cudaMemcpyDeviceToHost);
        printf("dependent DFMA  threads=%4d: %.2f cycles/op\n", threads, (double)h[0] / n);
        ind_dfma<<<1, threads>>>(d, c, n); cudaMemcpy(h, c, 16, cudaMemcpyDeviceToHost);
        printf("8 indep DFMA    threads=%4d: %.2f cycles per 8 ops -> %.1f DFMA lanes/clk/SM\n", threads, (double)h[0] / n, 8.0 * threads * n / (double)h[0]);
    }
    dep_ffma<<<1, 32>>>(f, c, n); cudaMemcpy(h, c, 16, cudaMemcpyDeviceToHost);
    printf("dependent FFMA  threads=  32: %.2f cycles/op\n", (double)h[0] / n);
    ddiv_lat<<<1, 32>>>(d, c, n); cudaMemcpy(h, c, 16, cudaMemcpyDeviceToHost);
    printf("dependent DDIV %.1f cycles, DSQRT+DADD %.1f cycles\n", (double)h[0] / n, (double)h[1] / n);
    for (int threads : {1, 32, 128, 512, 1024}) {
        ciede_lat<<<1, threads>>>(lab, f, c, 64); cudaMemcpy(h, c, 16, cudaMemcpyDeviceToHost);
        printf("ciede00 threads=%4d: %.0f cycles per evaluation (dependent chain of 64)\n", threads, (double)h[0] / 64);
    }
    cudaError_t e = cudaDeviceSynchronize(); printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
