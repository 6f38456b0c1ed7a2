// kernels_merge_cluster.cuh -- K7 on a thread-block CLUSTER of four CTAs (four SMs), one role per SM:
//
//   rank 0, 1   owners   the weight map in registers, 64 warps, edge e -> warp e / (32 SLOTS)
//   rank 2      delta    duplicates, colour / geometry deltas, weights, tie stamps for the touched edges
//   rank 3      fold     ordered folds of the merged region's statistics (covariance warp, colour-mean warp, TMA loader)
//
// Why four SMs for a serial loop: the instruction cache of an SM holds 32 KB (2 k SASS instructions) and one merge
// walks ~6 k instructions of role code, so the single-CTA kernel (kernels_merge_fast.cuh) refetches its code from L2 on
// every merge (sm__icc_request_hit_rate 75 %, phases 3-4x slower than the same code in a hot loop).  Here every SM
// loops over one role that fits its cache.  The roles hand data over through distributed shared memory: the producer
// stores into the consumer's shared memory (st.shared::cluster) and arrives on the consumer's mbarrier with release
// semantics at cluster scope; the consumer waits with acquire.  Per merge:
//
//   owners  --head_part, mb_head-->  all four CTAs        (each owner CTA publishes its minimum; everybody takes the smaller)
//   owners  --te_*, mb_touched---->  delta                 (edges incident to a or b; slots from a remote atomic counter)
//   fold    --newgeo, mb_folded--->  delta                 (colour vector, centroid, normal of the merged region)
//   delta   --res_*, mb_results--->  owners                (new keys, pushed into the owning CTA)
//
// Same replay rules, memoised / speculated colour deltas and tie stamps as kernels_merge_fast.cuh.
#pragma once
#include "kernels_merge_fast.cuh"

namespace f3ps {

constexpr int kClThreads = 1024;
constexpr int kClOwnerWarps = 64;                 // two CTAs x 32 warps
constexpr int kClMaxTouched = 1024;

struct ClusterSmem {
    // ---- every CTA (targets of remote stores use the same offsets everywhere) ----
    unsigned long long *mb_head, *mb_touched, *mb_folded, *mb_results, *mb_stage;
    unsigned* head_part;                     // [2][4]: hi, lo, e, ab of each owner CTA's minimum
    int* err;
    // ---- owner view ----
    unsigned *wm_hi, *wm_lo, *wm_e, *wm_ab;  // per-warp partial minima
    unsigned *res_hi, *res_lo, *res_ab; float* res_dc;
    float* dc; unsigned* wmask;
    // ---- delta view ----
    int* misc; float* newgeo; unsigned* prof;
    unsigned *te_hi, *te_lo, *te_x; float* te_dc;
    float4 *te_ce, *te_nr; unsigned *hkey, *hcnt;
    unsigned short *partner, *mark; unsigned char* cls; int* dn; unsigned* wnew;
    // ---- fold view ----
    float4* stage; unsigned long long* rope;   // per run: next (16) | length (16) << 16 | first position << 32 -- one load per hop
    int* n; unsigned short *head, *tail;
    size_t bytes;
    __host__ __device__ ClusterSmem(char* base, unsigned S, unsigned E_half) {
        size_t o = 0;
        auto take = [&](size_t b) { char* p = base + o; o += (b + 15) & ~(size_t)15; return p; };
        mb_head = (unsigned long long*)take(8); mb_touched = (unsigned long long*)take(8); mb_folded = (unsigned long long*)take(8);
        mb_results = (unsigned long long*)take(8); mb_stage = (unsigned long long*)take(8);
        head_part = (unsigned*)take(8 * 4); err = (int*)take(4);
        const size_t role = o;
        // owner
        wm_hi = (unsigned*)take(32 * 4); wm_lo = (unsigned*)take(32 * 4); wm_e = (unsigned*)take(32 * 4); wm_ab = (unsigned*)take(32 * 4);
        res_hi = (unsigned*)take(kClMaxTouched * 4); res_lo = (unsigned*)take(kClMaxTouched * 4); res_ab = (unsigned*)take(kClMaxTouched * 4);
        res_dc = (float*)take(kClMaxTouched * 4);
        dc = (float*)take((size_t)E_half * 4); wmask = (unsigned*)take((size_t)S * 4);
        size_t hi_water = o;
        // delta
        o = role;
        misc = (int*)take(16 * 4); newgeo = (float*)take(16 * 4); prof = (unsigned*)take(16 * 4);
        te_hi = (unsigned*)take(kClMaxTouched * 4); te_lo = (unsigned*)take(kClMaxTouched * 4); te_x = (unsigned*)take(kClMaxTouched * 4);
        te_dc = (float*)take(kClMaxTouched * 4);
        te_ce = (float4*)take(kClMaxTouched * 16); te_nr = (float4*)take(kClMaxTouched * 16);
        hkey = (unsigned*)take(kFastHash * 4); hcnt = (unsigned*)take(kFastHash * 4);
        partner = (unsigned short*)take(kClMaxTouched * 2); mark = (unsigned short*)take((size_t)S * 2);
        cls = (unsigned char*)take(kClMaxTouched); dn = (int*)take((size_t)S * 4); wnew = (unsigned*)take(kClMaxTouched * 4);
        if (o > hi_water) hi_water = o;
        // fold
        o = role;
        stage = (float4*)take((size_t)kFastStage * 16);
        rope = (unsigned long long*)take((size_t)S * 8); n = (int*)take((size_t)S * 4);
        head = (unsigned short*)take((size_t)S * 2); tail = (unsigned short*)take((size_t)S * 2);
        if (o > hi_water) hi_water = o;
        bytes = hi_water;
    }
};

// ---- cluster PTX helpers ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned map_remote(const void* p, unsigned rank) {           // shared::cluster address of p in CTA `rank`
    unsigned r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr(p)), "r"(rank)); return r;
}
__device__ __forceinline__ void st_remote(unsigned raddr, unsigned v) { asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(raddr), "r"(v) : "memory"); }
__device__ __forceinline__ void st_remote_f(unsigned raddr, float v) { asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(raddr), "f"(v) : "memory"); }
__device__ __forceinline__ unsigned atom_add_remote(unsigned raddr, unsigned v) {
    unsigned old; asm volatile("atom.shared::cluster.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(raddr), "r"(v) : "memory"); return old;
}
__device__ __forceinline__ void mbar_arrive_remote(unsigned raddr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(unsigned mbar, unsigned parity) {     // acquire at cluster scope
    unsigned ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

struct ClHead { unsigned hi, lo, e, ab; };
__device__ __forceinline__ ClHead cl_head(const ClusterSmem& sm) {                         // smaller of the two owner CTAs' minima
    const unsigned h0 = sm.head_part[0], l0 = sm.head_part[1], h1 = sm.head_part[4], l1 = sm.head_part[5];
    const int w = key_less32(h1, l1, h0, l0) ? 1 : 0;
    ClHead h; h.hi = w ? h1 : h0; h.lo = w ? l1 : l0; h.e = sm.head_part[4 * w + 2] | ((unsigned)w << 31); h.ab = sm.head_part[4 * w + 3];
    return h;
}

template <int SLOTS>
__global__ void __launch_bounds__(kClThreads, 1) merge_cluster_kernel(FastArgs A) {
    extern __shared__ __align__(128) char smem_raw[];
    const ClusterSmem sm(smem_raw, A.S_cap, A.E_cap / 2);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned rank = cluster_rank();
    const unsigned nE = *A.n_edges_ptr, S = *A.n_sv_ptr;
    const RegionArrays R = A.R;
    const unsigned mb_head = smem_addr(sm.mb_head), mb_touched = smem_addr(sm.mb_touched), mb_folded = smem_addr(sm.mb_folded),
                   mb_results = smem_addr(sm.mb_results), mb_stage = smem_addr(sm.mb_stage);
    int* const newgeo_i = reinterpret_cast<int*>(sm.newgeo);

    // ---- set-up --------------------------------------------------------------------------------------------------
    if (tid == 0) {
        mbar_init(mb_head, 2u); mbar_init(mb_touched, (unsigned)kClOwnerWarps); mbar_init(mb_folded, 2u); mbar_init(mb_results, 1u); mbar_init(mb_stage, 1u);
        *sm.err = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (rank < 2) {
        for (unsigned s = tid; s < S; s += kClThreads) sm.wmask[s] = 0u;
    } else if (rank == 2) {
        for (unsigned s = tid; s < S; s += kClThreads) { sm.mark[s] = (unsigned short)kNil16; sm.dn[s] = R.n[s]; }
        for (int i = tid; i < kFastHash; i += kClThreads) { sm.hkey[i] = kDeadKey; sm.hcnt[i] = 0u; }
        for (int i = tid; i < kClMaxTouched; i += kClThreads) sm.partner[i] = (unsigned short)kNil16;
        if (tid == 0) {
            for (int i = 0; i < 16; ++i) { sm.misc[i] = 0; sm.prof[i] = 0u; }
            sm.misc[FM_EALIVE] = (int)nE; sm.misc[FM_RALIVE] = (int)S; sm.misc[FM_COUNTER] = (int)nE;
        }
    } else {
        for (unsigned s = tid; s < S; s += kClThreads) {
            sm.n[s] = R.n[s];
            const int h = R.head[s], t = R.tail[s], nx = R.next_run[s];
            sm.head[s] = (unsigned short)(h < 0 ? kNil16 : (unsigned)h); sm.tail[s] = (unsigned short)(t < 0 ? kNil16 : (unsigned)t);
            const unsigned rs = A.run_start[s], len = A.run_end[s] - rs;
            sm.rope[s] = (unsigned long long)(nx < 0 ? kNil16 : (unsigned)nx) | ((unsigned long long)len << 16) | ((unsigned long long)rs << 32);
        }
    }
    __syncthreads();
    cluster_sync_all();                                        // every CTA's mbarriers and tables exist before any remote access
    unsigned ph_head = 0, n_merges = 0;

    if (rank < 2) {
        // ================= owners =================
        const int gw = (int)rank * 32 + warp;                                               // 0..63
        const unsigned ebase = (unsigned)gw * 32u * SLOTS + (unsigned)lane;
        const unsigned dc_base = (unsigned)warp * 32u * SLOTS + (unsigned)lane;              // index into this CTA's dc[]
        unsigned khi[SLOTS], klo[SLOTS], kab[SLOTS];
        unsigned pending = 0;
#pragma unroll
        for (int j = 0; j < SLOTS; ++j) {
            khi[j] = kDeadKey; klo[j] = kDeadKey; kab[j] = 0xffffffffu;
            const unsigned e = ebase + 32u * j;
            if (e < nE && A.E.stamp[e] != kDeadStamp) {
                const float w = A.E.w[e];
                khi[j] = isnan(w) ? 0x7f800000u : __float_as_uint(w);
                klo[j] = (unsigned)(int)A.E.stamp[e] ^ 0x80000000u;
                const unsigned ea = A.E.a[e], eb = A.E.b[e];
                kab[j] = (ea << 16) | eb;
                sm.dc[dc_base + 32u * j] = A.E.dc[e];
                atomicOr(&sm.wmask[ea], 1u << warp); atomicOr(&sm.wmask[eb], 1u << warp);
            }
        }
        bool dirty = true;
        unsigned l_hi = kDeadKey, l_lo = kDeadKey, l_ab = 0xffffffffu; int l_slot = 0;
        unsigned ph_results = 0;
        unsigned last_a = 0xffffu, last_b = 0xffffu;
        FPROF_DECL;
#define OPHASE(i) FPROF(rank == 0 && tid == 0, i)
        // remote addresses in the delta CTA
        const unsigned r_tcount = map_remote(&sm.misc[FM_TCOUNT], 2u), r_te_hi = map_remote(sm.te_hi, 2u), r_te_lo = map_remote(sm.te_lo, 2u),
                       r_te_x = map_remote(sm.te_x, 2u), r_te_dc = map_remote(sm.te_dc, 2u), r_mb_touched = map_remote(sm.mb_touched, 2u);
        __syncthreads();
        while (true) {
            if (n_merges) { mbar_wait_cluster(mb_results, ph_results); ph_results ^= 1u; }     // the previous merge's keys are in res_*
            OPHASE(4);
            if (tid == 0 && last_a != 0xffffu) sm.wmask[last_a] |= sm.wmask[last_b];           // b's edges now name a
            if (pending) {
#pragma unroll
                for (int j = 0; j < SLOTS; ++j)
                    if (pending & (1u << j)) {
                        const unsigned p = klo[j];
                        khi[j] = sm.res_hi[p]; klo[j] = sm.res_lo[p]; kab[j] = khi[j] == kDeadKey ? 0xffffffffu : sm.res_ab[p];
                        sm.dc[dc_base + 32u * j] = sm.res_dc[p];
                        if (j == l_slot) dirty = true;
                        else if (key_less32(khi[j], klo[j], l_hi, l_lo)) { l_hi = khi[j]; l_lo = klo[j]; l_ab = kab[j]; l_slot = j; }
                    }
                pending = 0;
            }
            if (dirty) {
                l_hi = khi[0]; l_lo = klo[0]; l_ab = kab[0]; l_slot = 0;
#pragma unroll
                for (int j = 1; j < SLOTS; ++j)
                    if (key_less32(khi[j], klo[j], l_hi, l_lo)) { l_hi = khi[j]; l_lo = klo[j]; l_ab = kab[j]; l_slot = j; }
                dirty = false;
            }
            {
                unsigned m_hi, m_lo;
                const int win = warp_argmin(l_hi, l_lo, m_hi, m_lo);
                if (lane == win) { sm.wm_hi[warp] = m_hi; sm.wm_lo[warp] = m_lo; sm.wm_e[warp] = ((unsigned)l_slot << 16) | (unsigned)tid; sm.wm_ab[warp] = l_ab; }
            }
            OPHASE(0);
            __syncthreads();
            OPHASE(1);
            if (warp == 0) {                                                                   // this CTA's minimum -> all four CTAs
                unsigned c_hi, c_lo;
                const int win = warp_argmin(sm.wm_hi[lane], sm.wm_lo[lane], c_hi, c_lo);
                const unsigned c_e = sm.wm_e[win], c_ab = sm.wm_ab[win];
                if (lane < 4) {
                    const unsigned dst = map_remote(sm.head_part + 4 * rank, (unsigned)lane);
                    st_remote(dst, c_hi); st_remote(dst + 4, c_lo); st_remote(dst + 8, c_e); st_remote(dst + 12, c_ab);
                    mbar_arrive_remote(map_remote(sm.mb_head, (unsigned)lane));
                }
            }
            mbar_wait_cluster(mb_head, ph_head); ph_head ^= 1u;
            const ClHead hd = cl_head(sm);
            OPHASE(2);
            if (*sm.err || hd.hi == kDeadKey || !(__uint_as_float(hd.hi) < A.threshold)) break;  // strict <, src/clustering.cpp:388-389
            const unsigned a = hd.ab >> 16, b = hd.ab & 0xffffu;
            last_a = a; last_b = b;
            if (((sm.wmask[a] | sm.wmask[b]) >> warp) & 1u) {
                if ((hd.e >> 31) == rank && (hd.e & 0xffffu) == (unsigned)tid) {               // the head edge leaves the map
                    const int hs = (int)((hd.e >> 16) & 0x7fffu);
#pragma unroll
                    for (int j = 0; j < SLOTS; ++j) if (j == hs) { khi[j] = kDeadKey; klo[j] = kDeadKey; kab[j] = 0xffffffffu; }
                    dirty = true;
                }
                const unsigned aa = a * 0x10001u, bb = b * 0x10001u;
                unsigned hits = 0;
#pragma unroll
                for (int j = 0; j < SLOTS; ++j) hits |= ((__vcmpeq2(kab[j], aa) | __vcmpeq2(kab[j], bb)) != 0u ? 1u : 0u) << j;
                if (__any_sync(kFull, hits != 0u)) {
                    const int cnt = __popc(hits);
                    int incl = cnt;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(kFull, incl, off); if (lane >= off) incl += t; }
                    unsigned base = 0;
                    if (lane == 31) base = atom_add_remote(r_tcount, (unsigned)incl);
                    int p = (int)__shfl_sync(kFull, base, 31) + incl - cnt;
#pragma unroll
                    for (int j = 0; j < SLOTS; ++j) {
                        if (!((hits >> j) & 1u)) continue;
                        if (p < kClMaxTouched) {
                            const unsigned ea = kab[j] >> 16, eb = kab[j] & 0xffffu;
                            const bool on_a = ea == a || eb == a;
                            const unsigned x = (ea == a || ea == b) ? eb : ea;
                            st_remote(r_te_hi + 4u * p, khi[j]); st_remote(r_te_lo + 4u * p, klo[j]);
                            st_remote(r_te_x + 4u * p, x | (on_a ? 0x10000u : 0u) | (rank << 17));
                            st_remote_f(r_te_dc + 4u * p, sm.dc[dc_base + 32u * j]);
                            klo[j] = (unsigned)p; pending |= 1u << j;
                        }
                        ++p;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(r_mb_touched);
            OPHASE(3);
            ++n_merges;
        }
        if (rank == 0 && tid == 0) { for (int i = 0; i < 4; ++i) A.ctl->phase_cycles[20 + i] = pc[i]; A.ctl->phase_cycles[28] = pc[4]; }
#undef OPHASE
#pragma unroll
        for (int j = 0; j < SLOTS; ++j) {
            const unsigned e = ebase + 32u * j;
            if (e >= nE) continue;
            if (khi[j] == kDeadKey && klo[j] == kDeadKey) A.E.stamp[e] = kDeadStamp;
            else {
                A.E.a[e] = kab[j] >> 16; A.E.b[e] = kab[j] & 0xffffu; A.E.w[e] = __uint_as_float(khi[j]);
                A.E.stamp[e] = (long long)(int)(klo[j] ^ 0x80000000u); A.E.dc[e] = sm.dc[dc_base + 32u * j];
            }
        }
    } else if (rank == 2) {
        // ================= delta =================
        EdgeParams ep = A.ep;
        if (A.lambda_dev) ep.lambda = *A.lambda_dev;
        unsigned ph_touched = 0, ph_folded = 0;
        const unsigned r_mb_res0 = map_remote(sm.mb_results, 0u), r_mb_res1 = map_remote(sm.mb_results, 1u);
        unsigned r_res_hi[2], r_res_lo[2], r_res_ab[2], r_res_dc[2], r_err[4];
#pragma unroll
        for (int q = 0; q < 2; ++q) { r_res_hi[q] = map_remote(sm.res_hi, q); r_res_lo[q] = map_remote(sm.res_lo, q); r_res_ab[q] = map_remote(sm.res_ab, q); r_res_dc[q] = map_remote(sm.res_dc, q); }
#pragma unroll
        for (int q = 0; q < 4; ++q) r_err[q] = map_remote(sm.err, q);
        FPROF_DECL;
#define CPHASE(i) FPROF(tid == 0, i)
        while (true) {
            mbar_wait_cluster(mb_head, ph_head); ph_head ^= 1u;
            const ClHead hd = cl_head(sm);
            CPHASE(0);
            if (*sm.err || hd.hi == kDeadKey || !(__uint_as_float(hd.hi) < A.threshold)) break;
            const unsigned a = hd.ab >> 16, b = hd.ab & 0xffffu;
            const int na = sm.dn[a], nb = sm.dn[b];
            const bool big_is_a = na >= nb;
            const bool speculate = max(na, nb) >= 4 * min(na, nb);                              // see kernels_merge_fast.cuh
            const float4 guess = __ldcg(R.cvec + (big_is_a ? a : b));
            mbar_wait_cluster(mb_touched, ph_touched); ph_touched ^= 1u;                       // owners published the touched edges
            const int T = sm.misc[FM_TCOUNT];
            CPHASE(1);
            const int counter = sm.misc[FM_COUNTER];
            const unsigned ealive0 = (unsigned)sm.misc[FM_EALIVE], ralive0 = (unsigned)sm.misc[FM_RALIVE];
            if (T > kClMaxTouched) {                                                           // the host falls back to merge_kernel
                if (tid < 4) st_remote(r_err[tid], kFastErrTouched);
                mbar_wait_cluster(mb_folded, ph_folded); ph_folded ^= 1u;
                __syncthreads();
                if (tid == 0) { mbar_arrive_remote(r_mb_res0); mbar_arrive_remote(r_mb_res1); }
                continue;
            }
            if (T <= 32) {
                // ---- few touched edges: one warp, registers only ----
                if (warp == 0) {
                    const bool act = lane < T;
                    unsigned my_hi = kDeadKey, my_lo = kDeadKey, x = 0xffff0000u | (unsigned)lane, side_a = 0, orank = 0;
                    float dc = 0.0f;
                    if (act) { my_hi = sm.te_hi[lane]; my_lo = sm.te_lo[lane]; const unsigned tx = sm.te_x[lane]; x = tx & 0xffffu; side_a = (tx >> 16) & 1u; orank = (tx >> 17) & 1u; dc = sm.te_dc[lane]; }
                    unsigned lessmask = 0; bool dup = false;
                    for (int q = 0; q < T; ++q) {
                        const unsigned qh = sm.te_hi[q], ql = sm.te_lo[q], qx = sm.te_x[q] & 0xffffu;
                        const bool less = key_less32(qh, ql, my_hi, my_lo);
                        if (less) lessmask |= 1u << q;
                        dup = dup || (less && qx == x);                                        // the earlier of (a,x), (b,x) survives
                    }
                    const bool live = act && !dup;
                    float4 xcv = make_float4(0, 0, 0, 0), c4 = xcv, n4 = xcv;
                    if (live) { xcv = __ldcg(R.cvec + x); c4 = __ldcg(R.centroid + x); n4 = __ldcg(R.normal + x); }
                    const bool reuse = side_a ? big_is_a : (!big_is_a && ((b < x) == (a < x)));
                    const bool need = live && !reuse;
                    CPHASE(2);
                    if (speculate && __any_sync(kFull, need)) { if (need) dc = a < x ? colour_delta(ep.color_mode, guess, xcv) : colour_delta(ep.color_mode, xcv, guess); }
                    CPHASE(3);
                    mbar_wait_cluster(mb_folded, ph_folded);                                   // region a's new colour vector / centroid / normal
                    CPHASE(4);
                    const bool hit = newgeo_i[9] != 0;
                    const bool redo = hit ? (need && !speculate) : live;                      // wrong guess: every survivor; no guess made: the ones that cannot reuse
                    if (__any_sync(kFull, redo)) {
                        const float4 acv = make_float4(sm.newgeo[0], sm.newgeo[1], sm.newgeo[2], 0.0f);
                        if (redo) dc = a < x ? colour_delta(ep.color_mode, acv, xcv) : colour_delta(ep.color_mode, xcv, acv);
                    }
                    CPHASE(5);
                    int cls = FC_DUP; unsigned wbits = kDeadKey, ab = 0u;
                    if (live) {
                        const float4 ace = make_float4(sm.newgeo[3], sm.newgeo[4], sm.newgeo[5], 0.0f), anr = make_float4(sm.newgeo[6], sm.newgeo[7], sm.newgeo[8], 0.0f);
                        const bool a_first = a < x;
                        const float dg = geom_delta(ep.geom_mode, a_first ? anr : n4, a_first ? ace : c4, a_first ? n4 : anr, a_first ? c4 : ace);
                        ab = a_first ? (a << 16) | x : (x << 16) | a;
                        float w_new = unify(ep, dc, dg);
                        if (isnan(w_new)) { atomicAdd(&sm.misc[FM_NANW], 1); w_new = __int_as_float(0x7f800000); }
                        wbits = __float_as_uint(w_new);
                        cls = wbits == my_hi ? FC_KEEP : (wbits > my_hi ? FC_FRONT : FC_BACK);
                    }
                    const unsigned backm = __ballot_sync(kFull, act && cls == FC_BACK), frontm = __ballot_sync(kFull, act && cls == FC_FRONT);
                    const unsigned dupm = __ballot_sync(kFull, act && cls == FC_DUP), needm = __ballot_sync(kFull, need);
                    const int nbk = __popc(backm), nfr = __popc(frontm);
                    if (act) {
                        unsigned lo = my_lo;
                        if (cls == FC_BACK) lo = (unsigned)(counter + __popc(lessmask & backm)) ^ 0x80000000u;
                        else if (cls == FC_FRONT) lo = (unsigned)(-(counter + nbk + (nfr - 1 - __popc(lessmask & frontm)))) ^ 0x80000000u;
                        else if (cls == FC_DUP) lo = kDeadKey;
                        st_remote(r_res_hi[orank] + 4u * lane, wbits); st_remote(r_res_lo[orank] + 4u * lane, lo);
                        st_remote(r_res_ab[orank] + 4u * lane, ab); st_remote_f(r_res_dc[orank] + 4u * lane, dc);
                    }
                    if (lane == 0) {
                        sm.misc[FM_COUNTER] = counter + nbk + nfr;
                        sm.misc[FM_EALIVE] -= 1 + __popc(dupm); sm.misc[FM_RALIVE] -= 1;
                        if (T > sm.misc[FM_MAXT]) sm.misc[FM_MAXT] = T;
                        sm.misc[FM_SUMT] += T; sm.misc[FM_TCOUNT] = 0;
                        sm.misc[FM_EVALS] += __popc(needm) + (hit ? 0 : T - __popc(dupm)); sm.misc[FM_MISS] += hit ? 0 : 1;
                    }
                }
                if (warp != 0) mbar_wait_cluster(mb_folded, ph_folded);
                ph_folded ^= 1u;
            } else {
                // ---- many touched edges: the whole CTA, one entry per thread ----
                const int p = tid;
                const bool act = p < T;
                unsigned x = 0, side_a = 0, orank = 0;
                if (act) {
                    const unsigned tx = sm.te_x[p]; x = tx & 0xffffu; side_a = (tx >> 16) & 1u; orank = (tx >> 17) & 1u;
                    const unsigned short old = atomicCAS(&sm.mark[x], (unsigned short)kNil16, (unsigned short)p);
                    if (old != (unsigned short)kNil16) { sm.partner[p] = old; sm.partner[old] = (unsigned short)p; }
                }
                __syncthreads();
                bool live = false, need = false;
                float4 xcv = make_float4(0, 0, 0, 0), c4 = xcv, n4 = xcv;
                float dc = 0.0f;
                if (act) {
                    const unsigned q = sm.partner[p];
                    const bool dup = q != kNil16 && key_less32(sm.te_hi[q], sm.te_lo[q], sm.te_hi[p], sm.te_lo[p]);
                    sm.mark[x] = (unsigned short)kNil16;
                    live = !dup;
                    if (live) {
                        xcv = __ldcg(R.cvec + x); c4 = __ldcg(R.centroid + x); n4 = __ldcg(R.normal + x);
                        const bool reuse = side_a ? big_is_a : (!big_is_a && ((b < x) == (a < x)));
                        need = !reuse; dc = sm.te_dc[p];
                    }
                }
                CPHASE(7);
                if (need && speculate) dc = a < x ? colour_delta(ep.color_mode, guess, xcv) : colour_delta(ep.color_mode, xcv, guess);
                CPHASE(3);
                mbar_wait_cluster(mb_folded, ph_folded); ph_folded ^= 1u;
                CPHASE(4);
                const bool hit = newgeo_i[9] != 0;
                if (hit ? (need && !speculate) : live) {
                    const float4 acv = make_float4(sm.newgeo[0], sm.newgeo[1], sm.newgeo[2], 0.0f);
                    dc = a < x ? colour_delta(ep.color_mode, acv, xcv) : colour_delta(ep.color_mode, xcv, acv);
                }
                CPHASE(5);
                int cls = FC_DUP; unsigned wbits = kDeadKey, ab = 0u; int hslot = -1;
                if (act) sm.partner[p] = (unsigned short)kNil16;
                if (live) {
                    const float4 ace = make_float4(sm.newgeo[3], sm.newgeo[4], sm.newgeo[5], 0.0f), anr = make_float4(sm.newgeo[6], sm.newgeo[7], sm.newgeo[8], 0.0f);
                    const bool a_first = a < x;
                    const float dg = geom_delta(ep.geom_mode, a_first ? anr : n4, a_first ? ace : c4, a_first ? n4 : anr, a_first ? c4 : ace);
                    ab = a_first ? (a << 16) | x : (x << 16) | a;
                    float w_new = unify(ep, dc, dg);
                    if (isnan(w_new)) { atomicAdd(&sm.misc[FM_NANW], 1); w_new = __int_as_float(0x7f800000); }
                    wbits = __float_as_uint(w_new);
                    const unsigned old_hi = sm.te_hi[p];
                    cls = wbits == old_hi ? FC_KEEP : (wbits > old_hi ? FC_FRONT : FC_BACK);
                    if (cls != FC_KEEP) {                                                      // tie groups: same new weight, same side
                        const unsigned key = wbits | (cls == FC_FRONT ? 0x80000000u : 0u);
                        unsigned h = (key * 2654435761u) >> 21;
                        while (true) {
                            const unsigned prev = atomicCAS(&sm.hkey[h], kDeadKey, key);
                            if (prev == kDeadKey || prev == key) break;
                            h = (h + 1) & (kFastHash - 1);
                        }
                        atomicAdd(&sm.hcnt[h], 1u);
                        hslot = (int)h;
                    }
                }
                if (act) { sm.cls[p] = (unsigned char)cls; sm.wnew[p] = wbits; }
                if (act && !live) atomicAdd(&sm.misc[FM_ND], 1);
                __syncthreads();
                unsigned lo = act ? sm.te_lo[p] : 0u;
                if (hslot >= 0) {
                    const unsigned gsz = sm.hcnt[hslot];
                    unsigned rnk = 0;
                    if (gsz > 1) {
                        const unsigned ph = sm.te_hi[p], pl = sm.te_lo[p];
                        for (int q = 0; q < T; ++q)
                            if (sm.cls[q] == cls && sm.wnew[q] == wbits && key_less32(sm.te_hi[q], sm.te_lo[q], ph, pl)) ++rnk;
                    }
                    const int st = cls == FC_BACK ? counter + (int)rnk : -(counter + (int)(gsz - 1 - rnk));
                    lo = (unsigned)st ^ 0x80000000u;
                }
                if (act) {
                    if (!live) lo = kDeadKey;
                    st_remote(r_res_hi[orank] + 4u * p, wbits); st_remote(r_res_lo[orank] + 4u * p, lo);
                    st_remote(r_res_ab[orank] + 4u * p, ab); st_remote_f(r_res_dc[orank] + 4u * p, dc);
                }
                const unsigned needn = __syncthreads_count(need), liven = __syncthreads_count(live);
                if (hslot >= 0) { sm.hkey[hslot] = kDeadKey; sm.hcnt[hslot] = 0u; }
                if (tid == 0) {
                    sm.misc[FM_COUNTER] = counter + T;
                    sm.misc[FM_EALIVE] -= 1 + sm.misc[FM_ND]; sm.misc[FM_ND] = 0; sm.misc[FM_RALIVE] -= 1;
                    if (T > sm.misc[FM_MAXT]) sm.misc[FM_MAXT] = T;
                    sm.misc[FM_SUMT] += T; sm.misc[FM_BIGT] += 1; sm.misc[FM_TCOUNT] = 0;
                    sm.misc[FM_EVALS] += (int)needn + (hit ? 0 : (int)liven); sm.misc[FM_MISS] += hit ? 0 : 1;
                }
            }
            CPHASE(6);
            if (tid == 0) {
                sm.dn[a] = na + nb; sm.dn[b] = 0;
                if (n_merges < A.log_cap) {                                                    // debug line of :390-392 (ranks; labels at the end)
                    A.mlog.a[n_merges] = a; A.mlog.b[n_merges] = b; A.mlog.w[n_merges] = __uint_as_float(hd.hi);
                    A.mlog.edges_left[n_merges] = ealive0; A.mlog.regions_left[n_merges] = ralive0;
                }
            }
            __syncthreads();                                                                   // every result store of this CTA is issued
            if (tid == 0) { mbar_arrive_remote(r_mb_res0); mbar_arrive_remote(r_mb_res1); }
            ++n_merges;
        }
        FPROF_STORE(tid == 0, 0, 8);
#undef CPHASE
        __syncthreads();
        for (unsigned m = tid; m < n_merges && m < A.log_cap; m += kClThreads) { A.mlog.a[m] = A.sv_label[A.mlog.a[m]]; A.mlog.b[m] = A.sv_label[A.mlog.b[m]]; }
        if (tid == 0) {
            MergeCtl* ctl = A.ctl;
            ctl->phase_cycles[24] = (unsigned long long)sm.misc[FM_MISS]; ctl->phase_cycles[25] = (unsigned long long)sm.misc[FM_EVALS];
            ctl->phase_cycles[26] = (unsigned long long)sm.misc[FM_BIGT]; ctl->phase_cycles[27] = (unsigned long long)sm.misc[FM_SUMT];
            ctl->n_merges = n_merges; ctl->edges_alive = (unsigned)sm.misc[FM_EALIVE]; ctl->regions_alive = (unsigned)sm.misc[FM_RALIVE];
            ctl->counter = (long long)sm.misc[FM_COUNTER];
            ctl->max_touched = (unsigned)sm.misc[FM_MAXT]; ctl->nan_weights = (unsigned)sm.misc[FM_NANW]; ctl->error = (unsigned)*sm.err;
        }
    } else {
        // ================= fold =================
        const unsigned r_newgeo = map_remote(sm.newgeo, 2u), r_mb_folded = map_remote(sm.mb_folded, 2u);
        if (warp == 0) {
            // ---- covariance sums + xyz sums: lanes 0..8 each continue one accumulator of region a over b's voxels ----
            const float* stage_f = reinterpret_cast<const float*>(sm.stage);
            const int pi = lane < 3 ? 0 : (lane < 5 ? 1 : (lane == 5 ? 2 : (lane < 9 ? lane - 6 : 0)));
            const int qi = lane < 3 ? lane : (lane < 5 ? lane - 2 : 2);
            const bool prod = lane < 6;
            unsigned parity = 0;
            FPROF_DECL;
            while (true) {
                mbar_wait_cluster(mb_head, ph_head); ph_head ^= 1u;
                const ClHead hd = cl_head(sm);
                FPROF(lane == 0, 3);
                if (*sm.err || hd.hi == kDeadKey || !(__uint_as_float(hd.hi) < A.threshold)) break;
                const unsigned a = hd.ab >> 16, b = hd.ab & 0xffffu;
                const int na = sm.n[a], nb = sm.n[b];
                float acc = 0.0f;
                if (lane < 9) {
                    const float* src = lane < 4 ? reinterpret_cast<const float*>(R.accu0 + a) + lane
                                     : (lane < 8 ? reinterpret_cast<const float*>(R.accu1 + a) + (lane - 4) : reinterpret_cast<const float*>(R.accu2 + a));
                    acc = __ldcg(src);
                }
                for (int done = 0; done < nb; done += kFastStage) {
                    const int cn = min(nb - done, kFastStage);
                    mbar_wait(mb_stage, parity); parity ^= 1u;
                    FPROF(lane == 0, 0);
                    int j = 0;
                    for (; j + 4 <= cn; j += 4) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float p = stage_f[4 * (j + u) + pi];
                            const float q = prod ? stage_f[4 * (j + u) + qi] : 1.0f;
                            acc = acc + p * q;
                        }
                    }
                    for (; j < cn; ++j) {
                        const float p = stage_f[4 * j + pi];
                        const float q = prod ? stage_f[4 * j + qi] : 1.0f;
                        acc = acc + p * q;
                    }
                    __syncwarp(); named_bar(BAR_STAGE, 96);
                }
                FPROF(lane == 0, 1);
                float ac[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) ac[k] = __shfl_sync(kFull, acc, k);
                if (lane == 0) {
                    const int nn = na + nb;
                    const float fn = (float)nn;
                    const float cx = ac[6] / fn, cy = ac[7] / fn, cz = ac[8] / fn;              // computeCentroid (:411-413)
                    float nv[3]; float curv;
                    if (nn < 3) { nv[0] = nv[1] = nv[2] = nanf(""); curv = nv[0]; }
                    else plane_from_accu(ac, nn, nv, curv);                                     // computePointNormal (:415-417)
                    flip_and_normalize(cx, cy, cz, nv);                                         // :418-420
                    st_remote_f(r_newgeo + 12, cx); st_remote_f(r_newgeo + 16, cy); st_remote_f(r_newgeo + 20, cz);
                    st_remote_f(r_newgeo + 24, nv[0]); st_remote_f(r_newgeo + 28, nv[1]); st_remote_f(r_newgeo + 32, nv[2]);
                    R.centroid[a] = make_float4(cx, cy, cz, 0.0f);
                    R.normal[a] = make_float4(nv[0], nv[1], nv[2], curv);
                    R.accu0[a] = make_float4(ac[0], ac[1], ac[2], ac[3]);
                    R.accu1[a] = make_float4(ac[4], ac[5], ac[6], ac[7]);
                    R.accu2[a] = make_float4(ac[8], 0.0f, 0.0f, 0.0f);
                    mbar_arrive_remote(r_mb_folded);
                }
                __syncwarp();
                FPROF(lane == 0, 2);
                named_bar(BAR_STAGE, 96);                      // the loader may splice: nobody reads n[] / the ropes of this merge any more
                named_bar(BAR_STAGE, 96);                      // splice done
            }
            FPROF_STORE(lane == 0, 16, 4);
        } else if (warp == 1) {
            // ---- ColorUtilities::mean_color continued: lanes 0..2 carry r, g, b; every lane prepares one reciprocal ----
            const unsigned* stage_u = reinterpret_cast<const unsigned*>(sm.stage);
            const int shift = lane < 3 ? 16 - 8 * lane : 0;
            const EdgeParams ep = A.ep;
            unsigned parity = 0;
            unsigned long long fold_steps = 0;
            FPROF_DECL;
            while (true) {
                mbar_wait_cluster(mb_head, ph_head); ph_head ^= 1u;
                const ClHead hd = cl_head(sm);
                FPROF(lane == 0, 3);
                if (*sm.err || hd.hi == kDeadKey || !(__uint_as_float(hd.hi) < A.threshold)) break;
                const unsigned a = hd.ab >> 16, b = hd.ab & 0xffffu;
                const int na = sm.n[a], nb = sm.n[b];
                float m = 0.0f;
                if (lane < 3) m = __ldcg(reinterpret_cast<const float*>(R.mean + a) + 1 + lane);
                const float4 guess = __ldcg(R.cvec + (na >= nb ? a : b));                      // the delta SM's guess for the new colour vector
                const float cnt0 = (float)na;
                for (int done = 0; done < nb; done += kFastStage) {
                    const int cn = min(nb - done, kFastStage);
                    float inv_next = 1 / (cnt0 + (float)(done + lane + 1));
                    mbar_wait(mb_stage, parity); parity ^= 1u;
                    FPROF(lane == 0, 0);
                    for (int base = 0; base < cn; base += 32) {
                        const float inv_mine = inv_next;
                        inv_next = 1 / (cnt0 + (float)(done + base + 32 + lane + 1));
                        const int mcount = min(32, cn - base);
#pragma unroll 8
                        for (int j = 0; j < mcount; ++j) {
                            const float inv = __shfl_sync(kFull, inv_mine, j);
                            const float x = (float)((stage_u[4 * (base + j) + 3] >> shift) & 255u);
                            m = m + inv * (x - m);
                        }
                    }
                    __syncwarp(); named_bar(BAR_STAGE, 96);
                }
                FPROF(lane == 0, 1);
                const float mr = __shfl_sync(kFull, m, 0), mg = __shfl_sync(kFull, m, 1), mb = __shfl_sync(kFull, m, 2);
                float cv[3];
                if (ep.color_mode == 0) rgb2lab_lanes(ep.lab_lut, mr, mg, mb, lane, cv);
                else { cv[0] = mr; cv[1] = mg; cv[2] = mb; }
                if (lane == 0) {
                    st_remote_f(r_newgeo, cv[0]); st_remote_f(r_newgeo + 4, cv[1]); st_remote_f(r_newgeo + 8, cv[2]);
                    st_remote(r_newgeo + 36, (__float_as_uint(cv[0]) == __float_as_uint(guess.x) && __float_as_uint(cv[1]) == __float_as_uint(guess.y) &&
                                              __float_as_uint(cv[2]) == __float_as_uint(guess.z)) ? 1u : 0u);
                    R.mean[a] = make_float4((float)(na + nb), mr, mg, mb);
                    R.cvec[a] = make_float4(cv[0], cv[1], cv[2], 0.0f);
                    mbar_arrive_remote(r_mb_folded);
                }
                __syncwarp();
                FPROF(lane == 0, 2);
                named_bar(BAR_STAGE, 96);
                named_bar(BAR_STAGE, 96);
                fold_steps += (unsigned long long)nb; ++n_merges;
            }
            if (lane == 0) A.ctl->fold_steps = fold_steps;
            FPROF_STORE(lane == 0, 12, 4);
        } else if (warp == 2) {
            // ---- loader: walk b's rope, one bulk copy per run (or part of a run) into the stage; splice the ropes ----
            const unsigned stage_addr = smem_addr(sm.stage);
            FPROF_DECL;
            unsigned n_runs = 0;
            while (true) {
                mbar_wait_cluster(mb_head, ph_head); ph_head ^= 1u;
                const ClHead hd = cl_head(sm);
                FPROF(lane == 0, 2);
                if (*sm.err || hd.hi == kDeadKey || !(__uint_as_float(hd.hi) < A.threshold)) break;
                const unsigned a = hd.ab >> 16, b = hd.ab & 0xffffu;
                const int na = sm.n[a], nb = sm.n[b];
                unsigned run = sm.head[b];
                unsigned pos = 0, left = 0, nxt = kNil16;      // current run: next position, voxels left, successor
                if (run != kNil16) { const unsigned long long d = sm.rope[run]; nxt = (unsigned)d & 0xffffu; left = ((unsigned)d >> 16) & 0xffffu; pos = (unsigned)(d >> 32); }
                for (int done = 0; done < nb; done += kFastStage) {
                    const int cn = min(nb - done, kFastStage);
                    {                                           // the whole warp walks the rope in lockstep (broadcast reads)
                        int off = 0;
                        while (off < cn) {
                            const int take = min((int)left, cn - off);
                            for (int i = lane; i < take; i += 32) cp_async16(stage_addr + (unsigned)(off + i) * 16u, A.pos_data + pos + i);
                            off += take; pos += (unsigned)take; left -= (unsigned)take; ++n_runs;
                            if (left == 0) {
                                run = nxt;
                                if (run == kNil16) break;
                                const unsigned long long d = sm.rope[run]; nxt = (unsigned)d & 0xffffu; left = ((unsigned)d >> 16) & 0xffffu; pos = (unsigned)(d >> 32);
                            }
                        }
                        cp_async_wait_all();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_local(mb_stage);
                    }
                    FPROF(lane == 0, 0);
                    __syncwarp(); named_bar(BAR_STAGE, 96);    // the fold warps are done with this stage
                    FPROF(lane == 0, 1);
                }
                __syncwarp();
                named_bar(BAR_STAGE, 96);                      // ... and with n[] / the ropes of this merge
                if (lane == 0) {                               // voxels_ = a ++ b (:406-409, :426-429)
                    reinterpret_cast<unsigned short*>(sm.rope + sm.tail[a])[0] = sm.head[b]; sm.tail[a] = sm.tail[b];   // next field = low 16 bits
                    sm.n[a] = na + nb; sm.n[b] = 0;
                }
                __syncwarp();
                named_bar(BAR_STAGE, 96);                      // splice visible before the next merge's readers
                FPROF(lane == 0, 1);
            }
            if (lane == 0) { A.ctl->phase_cycles[29] = pc[0]; A.ctl->phase_cycles[30] = pc[1]; A.ctl->phase_cycles[31] = n_runs; A.ctl->phase_cycles[8] = pc[2]; }
        }
        // warps 3..31 of the fold CTA have no role
        __syncthreads();
        for (unsigned s = tid; s < S; s += kClThreads) {
            R.n[s] = sm.n[s];
            R.head[s] = sm.head[s] == kNil16 ? -1 : (int)sm.head[s]; R.tail[s] = sm.tail[s] == kNil16 ? -1 : (int)sm.tail[s];
            const unsigned nx = (unsigned)sm.rope[s] & 0xffffu;
            R.next_run[s] = nx == kNil16 ? -1 : (int)nx;
        }
    }
    cluster_sync_all();                                        // no CTA leaves while its shared memory may still be addressed
}

} // namespace f3ps
