// kernels_graph.cuh -- K6: supervoxel table, supervoxel adjacency, initial region statistics and
// the initial edge weights.
//   pcl makeSupervoxels / getSupervoxelAdjacency            
//   Clustering::set_initialstate / init_weights / init_merging_parameters / compute_cdf
//                                                           (src/clustering.cpp:212-314, 605-612)
#pragma once
#include "colour.cuh"
#include "kernels_vccs.cuh"

namespace f3ps {

// graph nodes are indexed by RANK s in [0,S): ascending label order (a < b  <=>  rank a < rank b)
struct RegionArrays {
    float4* mean;      // cnt, r, g, b
    float4* accu0;     // a0..a3
    float4* accu1;     // a4..a7
    float4* accu2;     // a8, -, -, -
    float4* centroid;  // cx, cy, cz, -
    float4* normal;    // nx, ny, nz, curvature
    float4* cvec;      // colour vector the edge weights use: Lab of the mean colour (LAB_CIEDE00) or the mean colour (RGB_EUCL)
    int* n;            // voxel count, 0 = erased
    int* head;         // rope: first / last run (run id = rank of an initial supervoxel), next pointer per run
    int* tail;
    int* next_run;
};

__device__ __forceinline__ void load_stats(const RegionArrays& R, int s, RegionStats& st) {
    const float4 m = R.mean[s], a0 = R.accu0[s], a1 = R.accu1[s], a2 = R.accu2[s];
    st.cnt = m.x; st.r = m.y; st.g = m.z; st.b = m.w;
    st.accu[0] = a0.x; st.accu[1] = a0.y; st.accu[2] = a0.z; st.accu[3] = a0.w;
    st.accu[4] = a1.x; st.accu[5] = a1.y; st.accu[6] = a1.z; st.accu[7] = a1.w; st.accu[8] = a2.x;
    st.n = R.n[s];
}
__device__ __forceinline__ void store_stats(const RegionArrays& R, int s, const RegionStats& st) {
    R.mean[s] = make_float4(st.cnt, st.r, st.g, st.b);
    R.accu0[s] = make_float4(st.accu[0], st.accu[1], st.accu[2], st.accu[3]);
    R.accu1[s] = make_float4(st.accu[4], st.accu[5], st.accu[6], st.accu[7]);
    R.accu2[s] = make_float4(st.accu[8], 0, 0, 0);
    R.n[s] = st.n;
}

// Warp-cooperative ordered fold of the voxels order[s..e) into `st` (replicated in every lane):
// the lanes fetch 32 voxels at a time, then every lane replays them in order.
__device__ __forceinline__ void fold_run(RegionStats& st, const unsigned* __restrict__ order, unsigned s, unsigned e,
                                         const float4* __restrict__ vox_xyz, int lane) {
    for (unsigned base = s; base < e; base += 32) {
        const unsigned m = min(32u, e - base);
        float4 mine = make_float4(0, 0, 0, 0);
        if (base + lane < e) mine = vox_xyz[order[base + lane]];
        for (unsigned j = 0; j < m; ++j) {
            const float x = __shfl_sync(kFull, mine.x, j), y = __shfl_sync(kFull, mine.y, j), z = __shfl_sync(kFull, mine.z, j);
            const uint32_t c = __float_as_uint(__shfl_sync(kFull, mine.w, j));
            stats_step(st, x, y, z, c);
        }
    }
}

// alive labels -> ranks
struct AliveOp {
    const unsigned* seg_start; const unsigned* seg_end; unsigned* sv_label; unsigned* rank_of_label;
    typedef int Payload;
    __device__ __forceinline__ bool test(int64_t i, Payload&) const { unsigned l = (unsigned)i + 1; return seg_end[l] > seg_start[l]; }
    __device__ __forceinline__ void emit(unsigned pos, int64_t i, const Payload&) const { sv_label[pos] = (unsigned)i + 1; rank_of_label[i + 1] = pos; }
};

// initial regions: one warp per supervoxel; statistics over its own voxels, geometry from the
// helper centroid (pcl::Supervoxel::centroid_ / normal_, NOT the covariance normal -- C.3)
__global__ void __launch_bounds__(256) region_init_kernel(const unsigned* __restrict__ sv_label, const unsigned* __restrict__ n_sv_ptr,
        const unsigned* __restrict__ seg_start, const unsigned* __restrict__ seg_end, const unsigned* __restrict__ sorted_vox,
        const float4* __restrict__ vox_xyz, Centroids cen, RegionArrays R, unsigned* __restrict__ run_start, unsigned* __restrict__ run_end) {
    const unsigned S = *n_sv_ptr;
    const int lane = threadIdx.x & 31;
    const unsigned warps_total = (gridDim.x * blockDim.x) >> 5;
    for (unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < S; s += warps_total) {
        const unsigned l = sv_label[s];
        const unsigned b = seg_start[l], e = seg_end[l];
        RegionStats st; st.cnt = st.r = st.g = st.b = 0.0f; st.n = 0;
#pragma unroll
        for (int k = 0; k < 9; ++k) st.accu[k] = 0.0f;
        fold_run(st, sorted_vox, b, e, vox_xyz, lane);
        if (lane == 0) {
            store_stats(R, s, st);
            const float4 c = cen.xyz[l], nn = cen.nrm[l];
            R.centroid[s] = make_float4(c.x, c.y, c.z, 0.0f);
            R.normal[s] = make_float4(nn.x, nn.y, nn.z, 0.0f);
            R.head[s] = (int)s; R.tail[s] = (int)s; R.next_run[s] = -1;
            run_start[s] = b; run_end[s] = e;
        }
    }
}

// f3ps_set_graph path: statistics over caller-supplied voxel ranges (order = identity)
__global__ void __launch_bounds__(256) region_init_ranges_kernel(unsigned S, const unsigned* __restrict__ run_start,
        const unsigned* __restrict__ run_end, const unsigned* __restrict__ order, const float4* __restrict__ vox_xyz, RegionArrays R) {
    const int lane = threadIdx.x & 31;
    const unsigned warps_total = (gridDim.x * blockDim.x) >> 5;
    for (unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < S; s += warps_total) {
        RegionStats st; st.cnt = st.r = st.g = st.b = 0.0f; st.n = 0;
#pragma unroll
        for (int k = 0; k < 9; ++k) st.accu[k] = 0.0f;
        fold_run(st, order, run_start[s], run_end[s], vox_xyz, lane);
        if (lane == 0) { store_stats(R, s, st); R.head[s] = (int)s; R.tail[s] = (int)s; R.next_run[s] = -1; }
    }
}

// getSupervoxelAdjacency + clear_adjacency: unique (a<b) label pairs through a hash set.  Walks the helpers'
// LEAF lists (position i: helper pos_label[i] holds voxel pos_vox[i]), as getNeighborLabels does, so a phantom
// leaf contributes (holder -> owners around it) as well; only pairs with first < second survive clear_adjacency.
__global__ void __launch_bounds__(256) edge_collect_kernel(const int* __restrict__ nbr_col, unsigned V_cap, const int* __restrict__ nbr_row,
        unsigned n_pos, const unsigned* __restrict__ pos_label, const unsigned* __restrict__ pos_vox, const unsigned* __restrict__ owner,
        const unsigned* __restrict__ rank_of_label, unsigned long long* __restrict__ set_slots, unsigned mask, unsigned* __restrict__ overflow) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pos; i += gridDim.x * blockDim.x) {
        const unsigned lv = pos_label[i];
        const unsigned v = pos_vox[i];
        const int cnt = nbr_row[(size_t)v * kNbrStride + 27];
        unsigned last = 0;
        for (int r = 0; r < cnt; ++r) {
            const unsigned lu = owner[nbr_col[(size_t)r * V_cap + v]];
            if (lu == 0 || lu <= lv || lu == last) continue;     // keep a < b only (clear_adjacency)
            last = lu;
            const unsigned long long k = (((unsigned long long)rank_of_label[lv]) << 32 | rank_of_label[lu]) + 1ull;
            unsigned h = hash64(k) & mask;
            unsigned probes = 0;
            while (true) {
                const unsigned long long prev = atomicCAS(&set_slots[h], 0ull, k);
                if (prev == 0ull || prev == k) break;
                h = (h + 1) & mask;
                if (++probes > mask) { *overflow = 1; break; }
            }
        }
    }
}
struct EdgeSlotOp {      // hash-set slots -> dense key list (rank_a << 32 | rank_b), unsorted
    const unsigned long long* slots; unsigned long long* keys; unsigned* vals;
    typedef int Payload;
    __device__ __forceinline__ bool test(int64_t i, Payload&) const { return slots[i] != 0ull; }
    __device__ __forceinline__ void emit(unsigned pos, int64_t i, const Payload&) const { keys[pos] = slots[i] - 1ull; vals[pos] = pos; }
};

struct EdgeArrays {
    unsigned* a; unsigned* b;        // ranks, a < b
    float* dc; float* dg; float* w;
    long long* stamp;                // tie key; dead edges carry kDeadStamp
};
constexpr long long kDeadStamp = 0x7fffffffffffffffll;

__device__ __forceinline__ void region_inputs(const RegionArrays& R, int s, float cv[3], float n[3], float c[3]) {
    const float4 m = R.cvec[s], cc = R.centroid[s], nn = R.normal[s];
    cv[0] = m.x; cv[1] = m.y; cv[2] = m.z;
    n[0] = nn.x; n[1] = nn.y; n[2] = nn.z;
    c[0] = cc.x; c[1] = cc.y; c[2] = cc.z;
}

// colour vectors of the initial regions (depends on the colour distance in force, so it belongs to init_weights)
__global__ void __launch_bounds__(256) region_cvec_kernel(const unsigned* __restrict__ n_sv_ptr, RegionArrays R, EdgeParams ep) {
    const unsigned S = *n_sv_ptr;
    for (unsigned s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
        const float4 m = R.mean[s];
        float cv[3]; colour_vector(ep, m.y, m.z, m.w, cv);
        R.cvec[s] = make_float4(cv[0], cv[1], cv[2], 0.0f);
    }
}

// init_weights first loop: (delta_c, delta_g) of every initial edge, lexicographic order
__global__ void __launch_bounds__(128) edge_delta_kernel(const unsigned long long* __restrict__ sorted_keys, const unsigned* __restrict__ n_edges_ptr,
        RegionArrays R, EdgeParams ep, EdgeArrays E, unsigned* __restrict__ dc_bits, unsigned* __restrict__ dg_bits) {
    const unsigned n = *n_edges_ptr;
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const unsigned long long k = sorted_keys[e];
        const unsigned a = (unsigned)(k >> 32), b = (unsigned)k;
        float rgb1[3], n1[3], c1[3], rgb2[3], n2[3], c2[3];
        region_inputs(R, a, rgb1, n1, c1); region_inputs(R, b, rgb2, n2, c2);
        float dc, dg;
        delta_cached(ep, rgb1, rgb2, n1, c1, n2, c2, dc, dg);
        E.a[e] = a; E.b[e] = b; E.dc[e] = dc; E.dg[e] = dg;
        dc_bits[e] = __float_as_uint(dc); dg_bits[e] = __float_as_uint(dg);   // sortable: deltas are >= 0
    }
}

// Clustering::deltas_mean over the ascending-sorted deltas (src/clustering.cpp:515-528): one warp per distribution;
// the warp stages 32 deltas and their reciprocals 1/count in shared memory (next chunk prefetched meanwhile) and
// lane 0 runs the dependent chain mean += (1/count)(delta - mean).  Then lambda = mean_g / (mean_c + mean_g)
// (init_merging_parameters, ADAPTIVE_LAMBDA).
__global__ void __launch_bounds__(64) adaptive_lambda_kernel(const unsigned* __restrict__ sorted_dc, const unsigned* __restrict__ sorted_dg,
                                                             const unsigned* __restrict__ n_edges_ptr, float* __restrict__ lambda_out) {
    __shared__ float s_val[2][2][32], s_inv[2][2][32];
    __shared__ float s_mean[2];
    const unsigned n = *n_edges_ptr;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned* src = w == 0 ? sorted_dc : sorted_dg;
    float mean_d = 0.0f;
    float v = lane < n ? __uint_as_float(src[lane]) : 0.0f;
    int buf = 0;
    for (unsigned base = 0; base < n; base += 32, buf ^= 1) {
        s_val[w][buf][lane] = v; s_inv[w][buf][lane] = 1 / (float)(base + lane + 1);     // count = count + 1.0f is exact here
        const unsigned nxt = base + 32 + lane;
        v = nxt < n ? __uint_as_float(src[nxt]) : 0.0f;
        __syncwarp();
        if (lane == 0) {
            const int m = (int)min(32u, n - base);
#pragma unroll 8
            for (int j = 0; j < m; ++j) mean_d = mean_d + s_inv[w][buf][j] * (s_val[w][buf][j] - mean_d);
        }
        __syncwarp();
    }
    if (lane == 0) s_mean[w] = mean_d;
    __syncthreads();
    if (threadIdx.x == 0) *lambda_out = s_mean[1] / (s_mean[0] + s_mean[1]);
}

// The same for graphs whose deltas fit shared memory (E <= 16384): block 0 sorts delta_c, block 1 delta_g (bitonic, in
// place), runs the chain, and the block that finishes last forms lambda -- one launch instead of two radix sorts.
__global__ void __launch_bounds__(1024) adaptive_lambda_smem_kernel(const float* __restrict__ dc, const float* __restrict__ dg,
        const unsigned* __restrict__ n_edges_ptr, unsigned n_pow2, float* __restrict__ means, unsigned* __restrict__ done, float* __restrict__ lambda_out) {
    extern __shared__ unsigned s_sort[];
    __shared__ float s_inv2[32];
    const unsigned n = *n_edges_ptr;
    const float* src = blockIdx.x == 0 ? dc : dg;
    for (unsigned i = threadIdx.x; i < n_pow2; i += blockDim.x) s_sort[i] = i < n ? __float_as_uint(src[i]) : 0xffffffffu;   // deltas are >= 0: bit order = value order
    __syncthreads();
    for (unsigned k = 2; k <= n_pow2; k <<= 1)
        for (unsigned j = k >> 1; j > 0; j >>= 1) {
            for (unsigned i = threadIdx.x; i < n_pow2; i += blockDim.x) {
                const unsigned l = i ^ j;
                if (l > i) {
                    const unsigned x = s_sort[i], y = s_sort[l];
                    if (((i & k) == 0) == (x > y)) { s_sort[i] = y; s_sort[l] = x; }
                }
            }
            __syncthreads();
        }
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        float mean_d = 0.0f;
        for (unsigned base = 0; base < n; base += 32) {
            s_inv2[lane] = 1 / (float)(base + lane + 1);                                  // count = count + 1.0f is exact here
            __syncwarp();
            if (lane == 0) {
                const int m = (int)min(32u, n - base);
#pragma unroll 8
                for (int j = 0; j < m; ++j) mean_d = mean_d + s_inv2[j] * (__uint_as_float(s_sort[base + j]) - mean_d);
            }
            __syncwarp();
        }
        if (lane == 0) {
            means[blockIdx.x] = mean_d;
            __threadfence();
            if (atomicAdd(done, 1u) == 1u) {                                              // the other distribution is finished too
                __threadfence();
                const float mc = *(volatile float*)&means[0], mg = *(volatile float*)&means[1];
                *lambda_out = mg / (mc + mg);
                *done = 0u;
            }
        }
    }
}

// Clustering::compute_cdf for both distributions: histogram (shared-memory atomics when the bins
// fit, else global), inclusive scan, cdf[i] = float(cum_i) / float(n).  One block per distribution.
__global__ void __launch_bounds__(1024) cdf_kernel(const float* __restrict__ dc, const float* __restrict__ dg, const unsigned* __restrict__ n_edges_ptr,
        int bins, float* __restrict__ cdf_c, float* __restrict__ cdf_g, unsigned* __restrict__ ghist, unsigned* __restrict__ bad_bin) {
    extern __shared__ unsigned s_bins[];
    const unsigned n = *n_edges_ptr;
    const float* src = blockIdx.x == 0 ? dc : dg;
    float* cdf = blockIdx.x == 0 ? cdf_c : cdf_g;
    const bool use_smem = bins <= 8192;
    unsigned* hist = use_smem ? s_bins : ghist + (size_t)blockIdx.x * bins;
    for (int i = threadIdx.x; i < bins; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
        short bin = (short)floorf(src[i] * bins);
        if (bin == bins) bin--;
        if (bin < 0 || bin >= bins) { *bad_bin = 1; continue; }   // the reference indexes out of bounds here (UB)
        atomicAdd(&hist[bin], 1u);
    }
    __syncthreads();
    // inclusive scan in chunks of blockDim with a running carry
    __shared__ unsigned s_warp[32];
    __shared__ unsigned s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < bins; base += blockDim.x) {
        const int i = base + threadIdx.x;
        unsigned v = i < bins ? hist[i] : 0u;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { unsigned t = __shfl_up_sync(kFull, v, off); if (lane >= off) v += t; }
        if (lane == 31) s_warp[warp] = v;
        __syncthreads();
        unsigned wb = 0;
        for (int w = 0; w < warp; ++w) wb += s_warp[w];
        const unsigned cum = s_carry + wb + v;
        if (i < bins) cdf[i] = (float)cum / (float)(int)n;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = cum;
        __syncthreads();
    }
}

// init_weights second loop: w = t_c(dc) + t_g(dg); stamp = lexicographic rank (multimap insertion order)
__global__ void __launch_bounds__(256) edge_weight_kernel(const unsigned* __restrict__ n_edges_ptr, EdgeParams ep, const float* __restrict__ lambda_dev,
                                                          EdgeArrays E, unsigned* __restrict__ nan_count) {
    const unsigned n = *n_edges_ptr;
    if (lambda_dev) ep.lambda = *lambda_dev;
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const float w = unify(ep, E.dc[e], E.dg[e]);
        if (isnan(w)) atomicAdd(nan_count, 1u);
        E.w[e] = w;
        E.stamp[e] = (long long)e;
    }
}

} // namespace f3ps
