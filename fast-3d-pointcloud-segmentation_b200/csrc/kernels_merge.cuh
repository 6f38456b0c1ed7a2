// kernels_merge.cuh -- K7: the greedy hierarchical merge, Clustering::cluster / merge / contains
// (/root/reference/src/clustering.cpp:384-469, 497-506), as ONE persistent thread block that
// replays the reference's serial min-edge order exactly:
//   * multimap order  ==  argmin over (weight, stamp)   -- stamp rule C.2
//   * region statistics are continued, not recomputed     -- prefix property C.3
// Per merge: block-wide argmin -> [warp 0: fold region b onto a, eigen-solve the new normal]
// in parallel with [warps 1..31: collect the edges incident to a or b] -> order them by their old
// key, drop duplicates (the earlier survives), recompute the survivors' weights in parallel
// (CIEDE2000 in FP64), assign tie stamps.
#pragma once
#include "kernels_graph.cuh"

namespace f3ps {

constexpr int kMergeThreads = 1024;
struct MergeScratch { int* e[2]; float* w[2]; long long* st[2]; unsigned* x[2]; unsigned char* cls; };

struct MergeLog { unsigned* a; unsigned* b; float* w; unsigned* edges_left; unsigned* regions_left; };
struct MergeCtl {
    unsigned n_merges, edges_alive, regions_alive, error;
    unsigned nan_weights, max_touched, pad0, pad1;
    long long counter;
    unsigned long long fold_steps;
    unsigned long long phase_cycles[32];  // clock64 deltas / event counts per phase (see f3ps_merge_profile)
};

__device__ __forceinline__ bool key_less(float w1, long long s1, float w2, long long s2) {
    return w1 < w2 || (w1 == w2 && s1 < s2);
}

__global__ void __launch_bounds__(kMergeThreads, 1) merge_kernel(RegionArrays R, EdgeArrays E, const unsigned* __restrict__ n_edges_ptr,
        const unsigned* __restrict__ n_sv_ptr, EdgeParams ep, const float* __restrict__ lambda_dev, float threshold,
        const unsigned* __restrict__ run_start, const unsigned* __restrict__ run_end, const unsigned* __restrict__ order,
        const float4* __restrict__ vox_xyz, const unsigned* __restrict__ sv_label, MergeLog mlog, unsigned log_cap, MergeCtl* ctl, MergeScratch scr) {
    __shared__ float s_rw[32]; __shared__ long long s_rs[32]; __shared__ int s_ri[32];
    __shared__ int s_head; __shared__ float s_head_w;
    __shared__ int s_tcount;
    int* const s_e[2] = {scr.e[0], scr.e[1]}; float* const s_w[2] = {scr.w[0], scr.w[1]}; long long* const s_st[2] = {scr.st[0], scr.st[1]};
    unsigned* const s_x[2] = {scr.x[0], scr.x[1]}; unsigned char* const s_class = scr.cls;     // global scratch, capacity = E
    __shared__ unsigned s_nm, s_ealive, s_ralive; __shared__ long long s_counter;
    __shared__ unsigned long long s_fold;
    unsigned long long pc[6] = {0, 0, 0, 0, 0, 0};
    long long t_prev = clock64();
#define PHASE(i) do { if (tid == 0) { long long t_now = clock64(); pc[i] += (unsigned long long)(t_now - t_prev); t_prev = t_now; } } while (0)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned nE = *n_edges_ptr;
    if (lambda_dev) ep.lambda = *lambda_dev;
    if (tid == 0) { s_nm = 0; s_ealive = nE; s_ralive = *n_sv_ptr; s_counter = (long long)nE; s_fold = 0; }
    __syncthreads();
    const float INF = __int_as_float(0x7f800000);
    enum { C_KEEP = 0, C_FRONT = 1, C_BACK = 2, C_DUP = 3 };

    while (true) {
        // ---- A: head of the weight map = argmin (w, stamp) -------------------------------
        float bw = INF; long long bs = kDeadStamp; int bi = -1;
        for (unsigned e = tid; e < nE; e += kMergeThreads) {
            const long long st = E.stamp[e];
            if (st == kDeadStamp) continue;
            const float w = E.w[e];
            if (bi < 0 || key_less(w, st, bw, bs)) { bw = w; bs = st; bi = (int)e; }
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) {
            const float ow = __shfl_xor_sync(kFull, bw, off); const long long os = __shfl_xor_sync(kFull, bs, off);
            const int oi = __shfl_xor_sync(kFull, bi, off);
            if (oi >= 0 && (bi < 0 || key_less(ow, os, bw, bs))) { bw = ow; bs = os; bi = oi; }
        }
        if (lane == 0) { s_rw[warp] = bw; s_rs[warp] = bs; s_ri[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            bw = s_rw[lane]; bs = s_rs[lane]; bi = s_ri[lane];
#pragma unroll
            for (int off = 16; off; off >>= 1) {
                const float ow = __shfl_xor_sync(kFull, bw, off); const long long os = __shfl_xor_sync(kFull, bs, off);
                const int oi = __shfl_xor_sync(kFull, bi, off);
                if (oi >= 0 && (bi < 0 || key_less(ow, os, bw, bs))) { bw = ow; bs = os; bi = oi; }
            }
            if (lane == 0) { s_head = bi; s_head_w = bw; s_tcount = 0; }
        }
        __syncthreads();
        const int head = s_head;
        PHASE(0);
        if (head < 0 || !(s_head_w < threshold)) break;          // strict <, src/clustering.cpp:388-389
        const unsigned a = E.a[head], b = E.b[head];
        __syncthreads();
        if (tid == 0) {
            const unsigned m = s_nm;
            if (m < log_cap) {                                    // debug line of :390-392
                mlog.a[m] = sv_label[a]; mlog.b[m] = sv_label[b]; mlog.w[m] = s_head_w;
                mlog.edges_left[m] = s_ealive; mlog.regions_left[m] = s_ralive;
            }
            E.stamp[head] = kDeadStamp;
        }
        // ---- B: fold (warp 0)  ||  touched-edge scan (warps 1..31) -----------------------
        if (warp == 0) {
            RegionStats st; load_stats(R, (int)a, st);
            unsigned long long steps = 0;
            for (int run = R.head[b]; run >= 0; run = R.next_run[run]) {
                const unsigned rs = run_start[run], re = run_end[run];
                fold_run(st, order, rs, re, vox_xyz, lane);       // voxels_ = a ++ b  (:408)
                steps += re - rs;
            }
            if (lane == 0) {
                store_stats(R, (int)a, st);
                R.next_run[R.tail[a]] = R.head[b]; R.tail[a] = R.tail[b];
                R.n[b] = 0;
                const float fn = (float)st.n;
                const float cx = st.accu[6] / fn, cy = st.accu[7] / fn, cz = st.accu[8] / fn;   // computeCentroid (:411-413)
                float n[3]; float curv;
                if (st.n < 3) { n[0] = n[1] = n[2] = nanf(""); curv = n[0]; }
                else plane_from_accu(st.accu, st.n, n, curv);                                     // computePointNormal (:415-417)
                flip_and_normalize(cx, cy, cz, n);                                                // :418-420
                R.centroid[a] = make_float4(cx, cy, cz, 0.0f);
                R.normal[a] = make_float4(n[0], n[1], n[2], curv);
                float cv[3]; colour_vector(ep, st.r, st.g, st.b, cv);
                R.cvec[a] = make_float4(cv[0], cv[1], cv[2], 0.0f);
                s_fold += steps;
            }
        } else {
            for (unsigned e = tid - 32; e < nE; e += kMergeThreads - 32) {
                if (E.stamp[e] == kDeadStamp || (int)e == head) continue;
                const unsigned ea = E.a[e], eb = E.b[e];
                if (ea == a || eb == a || ea == b || eb == b) {
                    const int slot = atomicAdd(&s_tcount, 1);
                    s_e[0][slot] = (int)e;
                }
            }
        }
        __threadfence();
        __syncthreads();
        PHASE(1);
        const int T = s_tcount;
        // ---- C: order the touched edges by their old key ---------------------------------
        for (int i = tid; i < T; i += kMergeThreads) {
            const int e = s_e[0][i];
            s_w[0][i] = E.w[e]; s_st[0][i] = E.stamp[e];
            const unsigned ea = E.a[e], eb = E.b[e];
            s_x[0][i] = (ea == a || ea == b) ? eb : ea;
        }
        __threadfence();
        __syncthreads();
        for (int i = tid; i < T; i += kMergeThreads) {
            const float w = s_w[0][i]; const long long st = s_st[0][i];
            int r = 0;
            for (int j = 0; j < T; ++j) r += key_less(s_w[0][j], s_st[0][j], w, st) ? 1 : 0;
            s_e[1][r] = s_e[0][i]; s_w[1][r] = w; s_st[1][r] = st; s_x[1][r] = s_x[0][i];
        }
        __threadfence();
        __syncthreads();
        PHASE(2);
        // ---- D: dedupe (earlier survives), recompute, classify ---------------------------
        for (int i = tid; i < T; i += kMergeThreads) {
            const unsigned x = s_x[1][i];
            bool dup = false;
            for (int q = 0; q < i; ++q) dup = dup || (s_x[1][q] == x);
            if (dup) s_class[i] = C_DUP;
            else {
                const unsigned lo = min(a, x), hi = max(a, x);
                float rgb1[3], n1[3], c1[3], rgb2[3], n2[3], c2[3];
                region_inputs(R, (int)lo, rgb1, n1, c1); region_inputs(R, (int)hi, rgb2, n2, c2);
                float dc, dg;
                delta_cached(ep, rgb1, rgb2, n1, c1, n2, c2, dc, dg);
                float w_new = unify(ep, dc, dg);
                if (isnan(w_new)) { atomicAdd(&ctl->nan_weights, 1u); w_new = INF; }
                const float w_old = s_w[1][i];
                s_class[i] = (w_new == w_old) ? C_KEEP : (w_new > w_old ? C_FRONT : C_BACK);
                s_w[0][i] = w_new;                                 // slot 0 of the scratch is free again: new weights
            }
        }
        __threadfence();
        __syncthreads();
        PHASE(3);
        // ---- E: tie stamps (C.2) and write back -------------------------------------------
        for (int i = tid; i < T; i += kMergeThreads) {
            const int cls = s_class[i];
            const int e = s_e[1][i];
            if (cls == C_DUP) E.stamp[e] = kDeadStamp;
            else {
                int nb = 0, nf = 0, rb = 0, rf = 0;
                for (int q = 0; q < T; ++q) {
                    const int c = s_class[q];
                    nb += (c == C_BACK); nf += (c == C_FRONT);
                    if (q < i) { rb += (c == C_BACK); rf += (c == C_FRONT); }
                }
                long long st = s_st[1][i];
                if (cls == C_BACK) st = s_counter + rb;
                else if (cls == C_FRONT) st = -(s_counter + nb + (nf - 1 - rf));
                const unsigned x = s_x[1][i];
                E.a[e] = min(a, x); E.b[e] = max(a, x); E.w[e] = s_w[0][i]; E.stamp[e] = st;
            }
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            int nb = 0, nf = 0, nd = 0;
            for (int q = 0; q < T; ++q) { const int c = s_class[q]; nb += (c == C_BACK); nf += (c == C_FRONT); nd += (c == C_DUP); }
            s_counter += nb + nf;
            s_ealive -= 1 + nd; s_ralive -= 1; s_nm += 1;
            if ((unsigned)T > ctl->max_touched) ctl->max_touched = (unsigned)T;
        }
        __threadfence();
        __syncthreads();
        PHASE(4);
    }
    if (tid == 0) {
        for (int i = 0; i < 6; ++i) ctl->phase_cycles[i] = pc[i];
        ctl->n_merges = s_nm; ctl->edges_alive = s_ealive; ctl->regions_alive = s_ralive; ctl->counter = s_counter;
        ctl->fold_steps = s_fold;
    }
}

// ---- result extraction: Clustering::get_labeled_cloud (:640-663) ---------------------------------
// dense labels 0..K-1 in ascending region label order; output offset of every run of every rope.
// One block: two exclusive scans over the regions (alive flag -> dense label, voxel count -> output offset), then one
// thread per surviving region walks its rope.
__global__ void __launch_bounds__(1024) dense_label_kernel(RegionArrays R, const unsigned* __restrict__ n_sv_ptr, const unsigned* __restrict__ run_start,
        const unsigned* __restrict__ run_end, unsigned* __restrict__ run_out_off, unsigned* __restrict__ run_dense,
        unsigned* __restrict__ region_dense, unsigned* __restrict__ n_out) {
    __shared__ unsigned s_wa[32], s_wn[32];
    __shared__ unsigned s_ca, s_cn;
    const unsigned S = *n_sv_ptr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_ca = 0; s_cn = 0; }
    __syncthreads();
    for (unsigned base = 0; base < S; base += blockDim.x) {
        const unsigned s = base + threadIdx.x;
        const int nv = s < S ? R.n[s] : 0;
        const unsigned alive = nv > 0 ? 1u : 0u, cnt = nv > 0 ? (unsigned)nv : 0u;
        unsigned ia = alive, in = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned ta = __shfl_up_sync(kFull, ia, o), tn = __shfl_up_sync(kFull, in, o);
            if (lane >= o) { ia += ta; in += tn; }
        }
        if (lane == 31) { s_wa[warp] = ia; s_wn[warp] = in; }
        __syncthreads();
        unsigned wa = 0, wn = 0;
        for (int w = 0; w < warp; ++w) { wa += s_wa[w]; wn += s_wn[w]; }
        const unsigned ca = s_ca, cn = s_cn;
        if (s < S) {
            if (!alive) region_dense[s] = 0xffffffffu;
            else {
                const unsigned dense = ca + wa + ia - 1u;
                unsigned off = cn + wn + in - cnt;
                region_dense[s] = dense;
                for (int run = R.head[s]; run >= 0; run = R.next_run[run]) {
                    run_out_off[run] = off; run_dense[run] = dense;
                    off += run_end[run] - run_start[run];
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) { s_ca = ca + wa + ia; s_cn = cn + wn + in; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = s_cn;
}
__global__ void __launch_bounds__(256) labeled_cloud_kernel(const unsigned* __restrict__ pos_run, unsigned n_pos, const unsigned* __restrict__ order,
        const unsigned* __restrict__ run_start, const unsigned* __restrict__ run_out_off, const unsigned* __restrict__ run_dense,
        const float4* __restrict__ vox_xyz, float* __restrict__ out_xyz, unsigned* __restrict__ out_label, unsigned* __restrict__ out_voxel,
        unsigned* __restrict__ vox_segment, const unsigned* __restrict__ pos_label, const unsigned* __restrict__ owner) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pos; i += gridDim.x * blockDim.x) {
        const unsigned run = pos_run[i];
        if (run == 0xffffffffu) continue;                 // unowned voxel: absent from every region
        const unsigned o = run_out_off[run] + (i - run_start[run]);
        const unsigned v = order[i];
        const float4 p = vox_xyz[v];
        out_xyz[3 * (size_t)o] = p.x; out_xyz[3 * (size_t)o + 1] = p.y; out_xyz[3 * (size_t)o + 2] = p.z;
        out_label[o] = run_dense[run]; out_voxel[o] = v;
        if (!owner || owner[v] == pos_label[i]) vox_segment[v] = run_dense[run];   // a phantom leaf does not relabel its voxel
    }
}

} // namespace f3ps
