// Micro-probe: latencies that bound the per-merge critical path of the resident merge kernel (development aid).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void chase(const unsigned* next, int n, long long* out, unsigned* sink) {
    unsigned i = 0; long long t0 = clock64();
    for (int k = 0; k < n; ++k) i = __ldcg(next + i);
    long long t1 = clock64(); out[0] = (t1 - t0) / n; *sink = i;
}
__global__ void bulk_lat(const float4* src, long long* out, int bytes, int reps, float* sink) {
    __shared__ __align__(128) float4 stage[512];
    __shared__ unsigned long long mbar;
    const unsigned mb = smem_addr(&mbar), st = smem_addr(stage);
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb)); asm volatile("fence.proxy.async.shared::cta;"); }
    __syncthreads();
    unsigned parity = 0; long long tot = 0; float acc = 0;
    for (int r = 0; r < reps; ++r) {
        long long t0 = clock64();
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(st), "l"(src + (r * 64) % 4096), "r"(bytes), "r"(mb) : "memory");
        }
        unsigned ok;
        do { asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(mb), "r"(parity) : "memory"); } while (!ok);
        parity ^= 1;
        long long t1 = clock64(); tot += t1 - t0; acc += stage[threadIdx.x & 15].x;
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = tot / reps; *sink = acc; }
}
__global__ void bar_lat(long long* out, int reps) {
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0) / reps;
}
__global__ void redux_lat(long long* out, int reps, unsigned* sink) {
    unsigned v = threadIdx.x * 2654435761u; long long t0 = clock64();
    for (int r = 0; r < reps; ++r) v = __reduce_min_sync(0xffffffffu, v + threadIdx.x) + r;
    long long t1 = clock64(); unsigned w = v; long long t2 = clock64();
    for (int r = 0; r < reps; ++r) w = __shfl_sync(0xffffffffu, w + 1, r & 31);
    long long t3 = clock64(); unsigned b = w; long long t4 = clock64();
    for (int r = 0; r < reps; ++r) b = __ballot_sync(0xffffffffu, (b + threadIdx.x) & 1) + r;
    long long t5 = clock64();
    __shared__ unsigned s[64]; s[threadIdx.x & 63] = 0; __syncthreads();
    long long t6 = clock64(); unsigned q = 0;
    for (int r = 0; r < reps; ++r) q = s[(q + r) & 63] + 1;
    long long t7 = clock64();
    __shared__ unsigned short h[64]; h[threadIdx.x & 63] = 0xffff; __syncthreads();
    long long t8 = clock64();
    for (int r = 0; r < reps; ++r) q += atomicCAS(&h[(threadIdx.x * 2) & 63], (unsigned short)0xffff, (unsigned short)r);
    long long t9 = clock64();
    if (threadIdx.x == 0) { out[0] = (t1 - t0) / reps; out[1] = (t3 - t2) / reps; out[2] = (t5 - t4) / reps; out[3] = (t7 - t6) / reps; out[4] = (t9 - t8) / reps; *sink = v + w + b + q; }
}
__device__ __forceinline__ int warp_argmin(unsigned hi, unsigned lo, unsigned& m_hi, unsigned& m_lo) {
    m_hi = __reduce_min_sync(0xffffffffu, hi);
    m_lo = __reduce_min_sync(0xffffffffu, hi == m_hi ? lo : 0xffffffffu);
    return __ffs(__ballot_sync(0xffffffffu, hi == m_hi && lo == m_lo)) - 1;
}
// the argmin skeleton of the merge kernel: 25 warps publish a partial minimum, one CTA barrier, every warp derives the head
__global__ void head_loop(long long* out, int reps, unsigned* sink) {
    __shared__ unsigned wm_hi[32], wm_lo[32], wm_e[32], wm_ab[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned l_hi = threadIdx.x * 2654435761u >> 4, l_lo = threadIdx.x, acc = 0;
    if (threadIdx.x < 32) { wm_hi[lane] = wm_lo[lane] = 0xffffffffu; wm_e[lane] = wm_ab[lane] = 0; }
    __syncthreads();
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (warp >= 7) {
            unsigned m_hi, m_lo; const int win = warp_argmin(l_hi, l_lo, m_hi, m_lo);
            if (lane == win) { wm_hi[warp - 7] = m_hi; wm_lo[warp - 7] = m_lo; wm_e[warp - 7] = threadIdx.x; wm_ab[warp - 7] = l_hi ^ l_lo; }
        }
        __syncthreads();
        unsigned h_hi, h_lo;
        const unsigned hi = lane < 25 ? wm_hi[lane] : 0xffffffffu, lo = lane < 25 ? wm_lo[lane] : 0xffffffffu;
        const int win = warp_argmin(hi, lo, h_hi, h_lo);
        const unsigned e = wm_e[win], ab = wm_ab[win];
        acc += e + ab;
        if (threadIdx.x == e) l_hi += 977u;          // the head changes every round
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0) / reps;
    sink[0] = acc;
}
// rope walk + staging of R runs of 20 voxels, as the merge loader does it (one warp)
__global__ void loader_loop(const float4* pos_data, long long* out, int R, int reps, int mode, float* sink) {
    __shared__ __align__(128) float4 stage[2048];
    __shared__ unsigned rs[512], re[512]; __shared__ unsigned short nx[512];
    __shared__ unsigned long long mbar;
    const int lane = threadIdx.x;
    for (int i = lane; i < 512; i += 32) { rs[i] = (unsigned)((i * 37) % 500) * 20u; re[i] = rs[i] + 20u; nx[i] = (unsigned short)((i * 7 + 3) % 512); }
    const unsigned mb = smem_addr(&mbar), st = smem_addr(stage);
    if (lane == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb)); asm volatile("fence.proxy.async.shared::cta;"); }
    __syncwarp();
    unsigned parity = 0; float acc = 0; long long tot = 0; unsigned run0 = 1;
    for (int r = 0; r < reps; ++r) {
        long long t0 = clock64();
        unsigned run = run0; unsigned pos = rs[run]; int off = 0; const int cn = R * 20;
        if (mode == 0) {
            if (lane == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(cn * 16) : "memory");
                while (off < cn) {
                    const unsigned end = re[run]; const int take = min((int)(end - pos), cn - off);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(st + off * 16), "l"(pos_data + pos), "r"(take * 16), "r"(mb) : "memory");
                    off += take; pos += take;
                    if (pos == end) { run = nx[run]; pos = rs[run]; }
                }
            }
        } else {
            while (off < cn) {
                const unsigned end = re[run]; const int take = min((int)(end - pos), cn - off);
                for (int i = lane; i < take; i += 32) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(st + (off + i) * 16), "l"(pos_data + pos + i) : "memory");
                off += take; pos += take;
                if (pos == end) { run = nx[run]; pos = rs[run]; }
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mb) : "memory");
        }
        unsigned ok;
        do { asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(mb), "r"(parity) : "memory"); } while (!ok);
        parity ^= 1;
        long long t1 = clock64(); tot += t1 - t0; acc += stage[lane].x; run0 = run;
        __syncwarp();
    }
    if (lane == 0) { out[0] = tot / reps; *sink = acc; }
}
int main() {
    const int N = 1 << 16;   // 256 KB chain: L2-resident, larger than nothing else
    unsigned* h = new unsigned[N]; for (int i = 0; i < N; ++i) h[i] = (unsigned)((i * 40503u + 12345u) % N);
    unsigned* d; long long* o; unsigned* sink; float4* src; float* fs;
    cudaMalloc(&d, N * 4); cudaMalloc(&o, 64); cudaMalloc(&sink, 4); cudaMalloc(&src, 16384 * 16); cudaMalloc(&fs, 4);
    cudaMemcpy(d, h, N * 4, cudaMemcpyHostToDevice); cudaMemset(src, 0, 16384 * 16);
    long long r[8];
    chase<<<1, 1>>>(d, 2000, o, sink); chase<<<1, 1>>>(d, 4000, o, sink); cudaMemcpy(r, o, 8, cudaMemcpyDeviceToHost);
    printf("ld.global.cg dependent chase (L2): %lld cycles\n", r[0]);
    for (int bytes : {16, 256, 1024, 8192}) { bulk_lat<<<1, 32>>>(src, o, bytes, 200, fs); cudaMemcpy(r, o, 8, cudaMemcpyDeviceToHost); printf("cp.async.bulk %5d B issue->mbarrier observed: %lld cycles\n", bytes, r[0]); }
    for (int t : {64, 256, 1024}) { bar_lat<<<1, t>>>(o, 1000); cudaMemcpy(r, o, 8, cudaMemcpyDeviceToHost); printf("__syncthreads %4d threads: %lld cycles\n", t, r[0]); }
    redux_lat<<<1, 32>>>(o, 1000, sink); cudaMemcpy(r, o, 40, cudaMemcpyDeviceToHost);
    printf("redux.min %lld, shfl %lld, ballot %lld, lds %lld, smem atomicCAS16 %lld cycles (dependent)\n", r[0], r[1], r[2], r[3], r[4]);
    for (int mode = 0; mode < 2; ++mode) for (int R : {1, 6, 24}) { loader_loop<<<1, 32>>>(src, o, R, 200, mode, fs); cudaMemcpy(r, o, 8, cudaMemcpyDeviceToHost); printf("loader %s, %2d runs of 20 voxels: %lld cycles\n", mode ? "cp.async per lane" : "bulk copy per run", R, r[0]); }
    head_loop<<<1, 1024>>>(o, 2000, sink); cudaMemcpy(r, o, 8, cudaMemcpyDeviceToHost); printf("argmin skeleton (publish + barrier + head in all 32 warps): %lld cycles per round\n", r[0]);
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
