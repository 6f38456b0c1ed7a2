// kernels_merge_fast.cuh -- K7, the resident variant: Clustering::cluster / merge / contains
// (/root/reference/src/clustering.cpp:384-469, 497-506) replayed by ONE persistent CTA whose whole
// working set lives on the SM (SURVEY.md Appendix C for the replay rules):
//
//   edges         registers of 25 "owner" warps: order key (weight bits, 32-bit tie stamp) + packed (a,b);
//                 owner warp w holds edges [w*32*SLOTS, (w+1)*32*SLOTS): slot j, lane l = edge base + 32 j + l
//   ropes / sizes shared memory (per region: head / tail / next run, run bounds, voxel count)
//   voxels        position-ordered float4 (x,y,z,rgba) array in HBM/L2; the runs of region b are brought
//                 into a shared-memory stage by cp.async.bulk (one elected thread, mbarrier completion)
//   region stats  HBM/L2 (one read of a's accumulators per merge, issued before they are needed)
//
// The CTA is warp-specialised; every role runs its own loop (so the register allocation of one role does not
// carry another role's live state) and they meet at one CTA-wide barrier per merge plus producer/consumer named
// barriers (bar.arrive on the producing side, bar.sync on the consuming side):
//   B1 (all)            partial minima published -> every warp derives the head (a, b, w) by itself
//   warp 0              ordered fold of the 9 covariance sums (lanes 0..8, one chain each), centroid, eigen-solve
//   warp 1              ordered fold of the running colour mean (lanes 0..2), Lab of the new mean (8 lanes, one LUT corner each)
//   warp 2              rope walk + bulk copies, rope splice
//   warps 7..31 owners  find the edges incident to a or b, publish them               -> arrive bar 1
//   warps 3..6  delta   sync bar 1; order the touched edges by their old key, drop duplicates, prefetch the other end's
//                       geometry; sync bar 3 (fold warps arrive); survivors' weights (CIEDE2000 in FP64), tie stamps
//                                                                                        -> arrive bar 4
//   owners              sync bar 4; take the new keys, local minima                    -> B1
// The fold is the critical path: 3 dependent FP32 operations per voxel of b (m += (1/k)(x-m)) and cannot be
// re-associated without changing the reference's rounding.
//
// Limits of this variant (the host falls back to merge_kernel of kernels_merge.cuh beyond them):
// S < 65535 and the per-region tables must fit in shared memory, E <= 800 * SLOTS, at most 1024 edges touched by one merge.
#pragma once
#include "kernels_merge.cuh"

namespace f3ps {

constexpr int kFastThreads = 1024;
constexpr int kFastDeltaWarp0 = 3;                                  // warps 3..6 re-weight the touched edges
constexpr int kFastDeltaWarps = 4;
constexpr int kFastDeltaThreads = 32 * kFastDeltaWarps;             // 128
constexpr int kFastRoleWarps = kFastDeltaWarp0 + kFastDeltaWarps;   // 7 warps own no edges
constexpr int kFastOwners = kFastThreads - 32 * kFastRoleWarps;     // 800
constexpr int kFastOwnerWarps = kFastOwners / 32;                   // 25
constexpr int kFastFoldedCount = 64 + 32 + kFastDeltaThreads;          // fold warps arrive; loader + delta warps wait
constexpr int kFastMaxPer = 8;                                      // touched edges per delta thread (1024 / 128)
constexpr int kFastStage = 2048;                                    // voxels per staging round
constexpr int kFastMaxTouched = 1024;
constexpr int kFastHash = 2048;
constexpr unsigned kDeadKey = 0xffffffffu;
constexpr unsigned kNil16 = 0xffffu;
constexpr unsigned kFastErrTouched = 4u;      // == F3PS_MERGE_ERR_TOUCHED
constexpr unsigned kFastErrStamp = 8u;

struct FastArgs {
    RegionArrays R; EdgeArrays E;
    const unsigned* n_edges_ptr; const unsigned* n_sv_ptr;
    EdgeParams ep; const float* lambda_dev; float threshold;
    const unsigned* run_start; const unsigned* run_end;
    const float4* pos_data;                  // voxel (x,y,z,rgba) by position of the label-ordered list
    const unsigned* sv_label;
    MergeLog mlog; unsigned log_cap;
    MergeCtl* ctl;
    unsigned S_cap, E_cap;                   // table capacities the shared-memory layout was sized for
};

// shared-memory layout, shared by host (size) and device (pointers)
struct FastSmem {
    float4* stage; unsigned long long* mbar;
    unsigned *te_hi, *te_lo, *te_x, *res_hi, *res_lo, *res_ab, *hkey, *hcnt;
    float *te_dc; float4 *te_ce, *te_nr; unsigned short* need;     // colour delta per touched edge, x's centroid / normal, CIEDE work list
    float* dc;                               // colour delta of every edge (owner-managed), valid for the current colour vectors of both ends
    unsigned *wm_hi, *wm_lo, *wm_e, *wm_ab;
    float* newgeo;                           // cvec[3], centroid[3], normal[3]
    unsigned* prof;                          // owner-side cycle counters (one lane)
    int* misc;                               // tcount, ealive, ralive, counter, nd, nanw, error, maxt
    unsigned *rs, *re; int* n;
    unsigned* wmask;                         // per region: owner warps that hold an edge incident to it (superset)
    unsigned short *head, *tail, *next, *mark, *partner; unsigned char* cls;
    size_t bytes;
    __host__ __device__ FastSmem(char* base, unsigned S, unsigned E_cap) {
        size_t o = 0;
        auto take = [&](size_t b) { char* p = base + o; o += (b + 15) & ~(size_t)15; return p; };
        stage = (float4*)take((size_t)kFastStage * 16); mbar = (unsigned long long*)take(16);
        te_hi = (unsigned*)take(kFastMaxTouched * 4); te_lo = (unsigned*)take(kFastMaxTouched * 4); te_x = (unsigned*)take(kFastMaxTouched * 4);
        res_hi = (unsigned*)take(kFastMaxTouched * 4); res_lo = (unsigned*)take(kFastMaxTouched * 4); res_ab = (unsigned*)take(kFastMaxTouched * 4);
        hkey = (unsigned*)take(kFastHash * 4); hcnt = (unsigned*)take(kFastHash * 4);
        te_dc = (float*)take(kFastMaxTouched * 4); te_ce = (float4*)take(kFastMaxTouched * 16); te_nr = (float4*)take(kFastMaxTouched * 16);
        need = (unsigned short*)take(kFastMaxTouched * 2); dc = (float*)take((size_t)E_cap * 4);
        wm_hi = (unsigned*)take(32 * 4); wm_lo = (unsigned*)take(32 * 4); wm_e = (unsigned*)take(32 * 4); wm_ab = (unsigned*)take(32 * 4);
        newgeo = (float*)take(16 * 4); misc = (int*)take(16 * 4); prof = (unsigned*)take(16 * 4);
        rs = (unsigned*)take((size_t)S * 4); re = (unsigned*)take((size_t)S * 4); n = (int*)take((size_t)S * 4);
        wmask = (unsigned*)take((size_t)S * 4);
        head = (unsigned short*)take((size_t)S * 2); tail = (unsigned short*)take((size_t)S * 2); next = (unsigned short*)take((size_t)S * 2);
        mark = (unsigned short*)take((size_t)S * 2); partner = (unsigned short*)take(kFastMaxTouched * 2);
        cls = (unsigned char*)take(kFastMaxTouched);
        bytes = o;
    }
};
enum { FM_TCOUNT = 0, FM_EALIVE, FM_RALIVE, FM_COUNTER, FM_ND, FM_NANW, FM_ERROR, FM_MAXT, FM_BIGT, FM_SUMT, FM_NEED, FM_MISS, FM_EVALS };

// ---- PTX helpers: mbarrier + 1-D bulk copy (TMA engine, no tensor map), named barriers ------------------------
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned mbar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
// per-lane 16-byte asynchronous copy (LDGSTS) and a plain CTA-scope arrive (used by the cluster variant's loader)
__device__ __forceinline__ void cp_async16(unsigned dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_local(unsigned mbar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__device__ __forceinline__ bool key_less32(unsigned h1, unsigned l1, unsigned h2, unsigned l2) { return h1 < h2 || (h1 == h2 && l1 < l2); }

// warp-wide argmin of (hi, lo); returns the winning lane (keys are unique among live edges)
__device__ __forceinline__ int warp_argmin(unsigned hi, unsigned lo, unsigned& m_hi, unsigned& m_lo) {
    m_hi = __reduce_min_sync(kFull, hi);
    m_lo = __reduce_min_sync(kFull, hi == m_hi ? lo : 0xffffffffu);
    const unsigned who = __ballot_sync(kFull, hi == m_hi && lo == m_lo);
    return __ffs(who) - 1;
}

// OpenCV's LUT interpolation spread over 8 lanes (one lattice corner each); result valid in every lane of the warp
__device__ __forceinline__ void rgb2lab_lanes(const short* __restrict__ lut, float r255, float g255, float b255, int lane, float lab[3]) {
    const float in[3] = {r255, g255, b255};
    int t[3], f[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float v = in[k] / 255;
        v = fminf(fmaxf(v, 0.0f), 1.0f);
        const int c = (int)rintf(v * 16384.0f);
        t[k] = c >> 9; f[k] = (c >> 5) & 15;
    }
    int out[3] = {0, 0, 0};
    if (lane < 8) {
        const int dr = lane >> 2, dg = (lane >> 1) & 1, db = lane & 1;
        const int w = (dr ? f[0] : 16 - f[0]) * (dg ? f[1] : 16 - f[1]) * (db ? f[2] : 16 - f[2]);
        const int ir = min(t[0] + dr, 32), ig = min(t[1] + dg, 32), ib = min(t[2] + db, 32);
        const short* e = lut + ((ir * 33 + ig) * 33 + ib) * 3;
        out[0] = w * (int)e[0]; out[1] = w * (int)e[1]; out[2] = w * (int)e[2];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int v = out[k];
        v += __shfl_xor_sync(kFull, v, 4); v += __shfl_xor_sync(kFull, v, 2); v += __shfl_xor_sync(kFull, v, 1);
        out[k] = (__shfl_sync(kFull, v, 0) + 2048) >> 12;
    }
    lab[0] = ((float)out[0] / 16384.0f) * 100.0f;
    lab[1] = ((float)out[1] / 16384.0f) * 256.0f - 128.0f;
    lab[2] = ((float)out[2] / 16384.0f) * 256.0f - 128.0f;
}


struct FastHead { unsigned hi, lo, e, ab; };
#define FPROF_DECL unsigned pc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; unsigned t_prev = (unsigned)clock()
#define FPROF(cond, i) do { if (cond) { const unsigned t_now = (unsigned)clock(); pc[i] += t_now - t_prev; t_prev = t_now; } } while (0)
#define FPROF_STORE(cond, base, n) do { if (cond) for (int i_ = 0; i_ < (n); ++i_) A.ctl->phase_cycles[(base) + i_] = pc[i_]; } while (0)
// every warp derives the head of the weight map from the owner warps' partial minima (after B1)
__device__ __forceinline__ FastHead fast_head(const FastSmem& sm, int lane) {
    FastHead h;
    const unsigned hi = lane < kFastOwnerWarps ? sm.wm_hi[lane] : kDeadKey, lo = lane < kFastOwnerWarps ? sm.wm_lo[lane] : kDeadKey;
    const int win = warp_argmin(hi, lo, h.hi, h.lo);
    h.e = sm.wm_e[win]; h.ab = sm.wm_ab[win];
    return h;
}

enum { FC_KEEP = 0, FC_FRONT = 1, FC_BACK = 2, FC_DUP = 3, FC_REUSE = 0x10 };
enum { BAR_TOUCHED = 1, BAR_STAGE = 2, BAR_FOLDED = 3, BAR_RESULTS = 4, BAR_DELTA = 5 };
// producer side of a named barrier: the barrier itself orders the producer's earlier shared / global writes before the
// consumers' reads after their bar.sync (PTX: barrier instructions order prior accesses among the participants)
__device__ __forceinline__ void bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// delta_c of Clustering::delta_c_g (src/clustering.cpp:113-122) on two colour vectors, first argument = smaller label.
// One copy of the FP64 code in the kernel (the instruction cache, not the FP64 pipe, bounds a replicated version).
__device__ __noinline__ float colour_delta(int color_mode, float4 lo, float4 hi) {
    const float c1[3] = {lo.x, lo.y, lo.z}, c2[3] = {hi.x, hi.y, hi.z};
    float dc;
    if (color_mode == 0) { dc = lab_ciede00(c1, c2); dc /= F3PS_LAB_RANGE; }
    else { dc = rgb_eucl(c1, c2); dc /= F3PS_RGB_RANGE; }
    return dc;
}

// delta_g of Clustering::delta_c_g (src/clustering.cpp:126-138), first argument = smaller label (one copy, see above)
__device__ __noinline__ float geom_delta(int geom_mode, float4 n_lo, float4 c_lo, float4 n_hi, float4 c_hi) {
    const float n1[3] = {n_lo.x, n_lo.y, n_lo.z}, c1[3] = {c_lo.x, c_lo.y, c_lo.z}, n2[3] = {n_hi.x, n_hi.y, n_hi.z}, c2[3] = {c_hi.x, c_hi.y, c_hi.z};
    float dg = normals_diff(n1, c1, n2, c2);
    if (geom_mode == 1 && is_convex(n1, c1, n2, c2)) dg *= 0.5f;
    return dg;
}

template <int SLOTS>
__device__ __forceinline__ void merge_fast_body(const FastArgs& A) {
    extern __shared__ __align__(128) char smem_raw[];
    const FastSmem sm(smem_raw, A.S_cap, A.E_cap);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned nE = *A.n_edges_ptr, S = *A.n_sv_ptr;
    const RegionArrays R = A.R;
    const unsigned mbar = smem_addr(sm.mbar);
    int* const newgeo_i = reinterpret_cast<int*>(sm.newgeo);

    // ---- set-up: ropes -> shared memory ------------------------------------------------------------------------
    for (unsigned s = tid; s < S; s += kFastThreads) {
        sm.rs[s] = A.run_start[s]; sm.re[s] = A.run_end[s]; sm.n[s] = R.n[s];
        const int h = R.head[s], t = R.tail[s], nx = R.next_run[s];
        sm.head[s] = (unsigned short)(h < 0 ? kNil16 : (unsigned)h); sm.tail[s] = (unsigned short)(t < 0 ? kNil16 : (unsigned)t);
        sm.next[s] = (unsigned short)(nx < 0 ? kNil16 : (unsigned)nx); sm.mark[s] = (unsigned short)kNil16; sm.wmask[s] = 0u;
    }
    for (int i = tid; i < kFastHash; i += kFastThreads) { sm.hkey[i] = kDeadKey; sm.hcnt[i] = 0u; }
    if (tid == 0) {
        for (int i = 0; i < 16; ++i) { sm.misc[i] = 0; sm.prof[i] = 0u; }
        sm.misc[FM_EALIVE] = (int)nE; sm.misc[FM_RALIVE] = (int)S; sm.misc[FM_COUNTER] = (int)nE;
        mbar_init(mbar, 1u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    unsigned n_merges = 0;
    __syncthreads();                                          // tables initialised (the owners add to wmask[] next)

    if (warp == 0) {
        // =========== covariance sums + xyz sums: lanes 0..8 each continue one accumulator of region a over b's voxels =====
        const float* stage_f = reinterpret_cast<const float*>(sm.stage);
        const int pi = lane < 3 ? 0 : (lane < 5 ? 1 : (lane == 5 ? 2 : (lane < 9 ? lane - 6 : 0)));
        const int qi = lane < 3 ? lane : (lane < 5 ? lane - 2 : 2);
        const bool prod = lane < 6;
        unsigned parity = 0;
        FPROF_DECL;
        __syncthreads();
        while (true) {
            __syncthreads();                                                                   // B1
            const FastHead hd = fast_head(sm, lane);
            FPROF(lane == 0, 3);
            if (sm.misc[FM_ERROR] || hd.hi == kDeadKey || !(__uint_as_float(hd.hi) < A.threshold)) break;
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
                mbar_wait(mbar, parity); parity ^= 1u;
                FPROF(lane == 0, 0);
                int j = 0;
                for (; j + 8 <= cn; j += 8) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
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
                if (done + cn < nb) { __syncwarp(); named_bar(BAR_STAGE, 96); }
            }
            FPROF(lane == 0, 1);
            __syncwarp();
            float ac[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) ac[k] = __shfl_sync(kFull, acc, k);
            if (lane == 0) {
                const int nn = na + nb;
                const float fn = (float)nn;
                const float cx = ac[6] / fn, cy = ac[7] / fn, cz = ac[8] / fn;                  // computeCentroid (:411-413)
                float nv[3]; float curv;
                if (nn < 3) { nv[0] = nv[1] = nv[2] = nanf(""); curv = nv[0]; }
                else plane_from_accu(ac, nn, nv, curv);                                         // computePointNormal (:415-417)
                flip_and_normalize(cx, cy, cz, nv);                                             // :418-420
                sm.newgeo[3] = cx; sm.newgeo[4] = cy; sm.newgeo[5] = cz;
                sm.newgeo[6] = nv[0]; sm.newgeo[7] = nv[1]; sm.newgeo[8] = nv[2];
                R.centroid[a] = make_float4(cx, cy, cz, 0.0f);
                R.normal[a] = make_float4(nv[0], nv[1], nv[2], curv);
                R.accu0[a] = make_float4(ac[0], ac[1], ac[2], ac[3]);
                R.accu1[a] = make_float4(ac[4], ac[5], ac[6], ac[7]);
                R.accu2[a] = make_float4(ac[8], 0.0f, 0.0f, 0.0f);
            }
            __syncwarp();
            bar_arrive(BAR_FOLDED, kFastFoldedCount);
            FPROF(lane == 0, 2);
        }
        FPROF_STORE(lane == 0, 16, 4);
    } else if (warp == 1) {
        // =========== ColorUtilities::mean_color continued: lanes 0..2 carry r, g, b; every lane prepares one reciprocal =====
        const unsigned* stage_u = reinterpret_cast<const unsigned*>(sm.stage);
        const int shift = lane < 3 ? 16 - 8 * lane : 0;
        const EdgeParams ep = A.ep;
        unsigned parity = 0;
        unsigned long long fold_steps = 0;
        FPROF_DECL;
        __syncthreads();
        while (true) {
            __syncthreads();                                                                   // B1
            const FastHead hd = fast_head(sm, lane);
            FPROF(lane == 0, 3);
            if (sm.misc[FM_ERROR] || hd.hi == kDeadKey || !(__uint_as_float(hd.hi) < A.threshold)) break;
            const unsigned a = hd.ab >> 16, b = hd.ab & 0xffffu;
            const int na = sm.n[a], nb = sm.n[b];
            float m = 0.0f;
            if (lane < 3) m = __ldcg(reinterpret_cast<const float*>(R.mean + a) + 1 + lane);
            const float4 guess = __ldcg(R.cvec + (na >= nb ? a : b));                          // the delta warps' guess for the new colour vector
            const float cnt0 = (float)na;
            for (int done = 0; done < nb; done += kFastStage) {
                const int cn = min(nb - done, kFastStage);
                float inv_next = 1 / (cnt0 + (float)(done + lane + 1));
                mbar_wait(mbar, parity); parity ^= 1u;
                FPROF(lane == 0, 0);
                __syncwarp();                                  // converged warp: the shuffles below take their fast path
                for (int base = 0; base < cn; base += 32) {
                    const float inv_mine = inv_next;
                    inv_next = 1 / (cnt0 + (float)(done + base + 32 + lane + 1));
                    const int mcount = min(32, cn - base);
                    if (mcount == 32) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float inv = __shfl_sync(kFull, inv_mine, j);
                            const float x = (float)((stage_u[4 * (base + j) + 3] >> shift) & 255u);
                            m = m + inv * (x - m);
                        }
                    } else {
                        for (int j = 0; j < mcount; ++j) {
                            const float inv = __shfl_sync(kFull, inv_mine, j);
                            const float x = (float)((stage_u[4 * (base + j) + 3] >> shift) & 255u);
                            m = m + inv * (x - m);
                        }
                    }
                }
                if (done + cn < nb) { __syncwarp(); named_bar(BAR_STAGE, 96); }
            }
            FPROF(lane == 0, 1);
            __syncwarp();
            const float mr = __shfl_sync(kFull, m, 0), mg = __shfl_sync(kFull, m, 1), mb = __shfl_sync(kFull, m, 2);
            float cv[3];
            if (ep.color_mode == 0) rgb2lab_lanes(ep.lab_lut, mr, mg, mb, lane, cv);
            else { cv[0] = mr; cv[1] = mg; cv[2] = mb; }
            if (lane == 0) {
                sm.newgeo[0] = cv[0]; sm.newgeo[1] = cv[1]; sm.newgeo[2] = cv[2];
                newgeo_i[9] = (__float_as_uint(cv[0]) == __float_as_uint(guess.x) && __float_as_uint(cv[1]) == __float_as_uint(guess.y) &&
                               __float_as_uint(cv[2]) == __float_as_uint(guess.z)) ? 1 : 0;
                R.mean[a] = make_float4((float)(na + nb), mr, mg, mb);
                R.cvec[a] = make_float4(cv[0], cv[1], cv[2], 0.0f);
            }
            __syncwarp();
            bar_arrive(BAR_FOLDED, kFastFoldedCount);
            FPROF(lane == 0, 2);
            fold_steps += (unsigned long long)nb; ++n_merges;
        }
        if (lane == 0) A.ctl->fold_steps = fold_steps;
        FPROF_STORE(lane == 0, 12, 4);
    } else if (warp == 2) {
        // =========== loader: walk b's rope, one bulk copy per run (or part of a run) into the stage; splice the ropes =====
        const unsigned stage_addr = smem_addr(sm.stage);
        __syncthreads();
        while (true) {
            __syncthreads();                                                                   // B1
            const FastHead hd = fast_head(sm, lane);
            if (sm.misc[FM_ERROR] || hd.hi == kDeadKey || !(__uint_as_float(hd.hi) < A.threshold)) break;
            const unsigned a = hd.ab >> 16, b = hd.ab & 0xffffu;
            const int na = sm.n[a], nb = sm.n[b];
            unsigned run = sm.head[b];
            unsigned pos = run != kNil16 ? sm.rs[run] : 0u;
            for (int done = 0; done < nb; done += kFastStage) {
                const int cn = min(nb - done, kFastStage);
                if (lane == 0) {
                    mbar_arrive_expect_tx(mbar, (unsigned)cn * 16u);
                    int off = 0;
                    while (off < cn) {
                        const unsigned end = sm.re[run];
                        const int take = min((int)(end - pos), cn - off);
                        if (take > 0) bulk_g2s(stage_addr + (unsigned)off * 16u, A.pos_data + pos, (unsigned)take * 16u, mbar);
                        off += take; pos += (unsigned)take;
                        if (pos == end) { run = sm.next[run]; if (run != kNil16) pos = sm.rs[run]; else break; }
                    }
                }
                if (done + cn < nb) { __syncwarp(); named_bar(BAR_STAGE, 96); }
            }
            // rope splice and sizes (voxels_ = a ++ b, :406-409, :426-429) once nobody reads the old ones any more
            __syncwarp();
            named_bar(BAR_FOLDED, kFastFoldedCount);
            if (lane == 0) {
                sm.next[sm.tail[a]] = sm.head[b]; sm.tail[a] = sm.tail[b];
                sm.n[a] = na + nb; sm.n[b] = 0;
                sm.wmask[a] |= sm.wmask[b];                    // b's edges now name a (every owner read the masks before BAR_TOUCHED)
            }
        }
    } else if (warp < kFastRoleWarps) {
        // =========== delta warps: touched edges -> dedupe, colour / geometry deltas, weights, tie stamps =====
        // Colour deltas are memoised per edge and speculated: while the fold runs, the edges that cannot reuse their stored
        // delta get CIEDE2000 against a GUESS of the merged region's colour vector (that of the larger side; the Lab lattice
        // quantises the mean colour, so absorbing a small region usually leaves it bit-identical).  A wrong guess re-evaluates.
        EdgeParams ep = A.ep;
        if (A.lambda_dev) ep.lambda = *A.lambda_dev;
        const int d = tid - 32 * kFastDeltaWarp0;                                              // 0..127
        FPROF_DECL;
#define FPHASE(i) FPROF(d == 0, i)
        __syncthreads();
        while (true) {
            __syncthreads();                                                                   // B1
            const FastHead hd = fast_head(sm, lane);
            FPHASE(0);
            if (sm.misc[FM_ERROR] || hd.hi == kDeadKey || !(__uint_as_float(hd.hi) < A.threshold)) break;
            const unsigned a = hd.ab >> 16, b = hd.ab & 0xffffu;
            const int na_ = sm.n[a], nb_ = sm.n[b];
            const bool big_is_a = na_ >= nb_;
            // guess the merged colour vector only when one side dominates (4x): two comparable regions move the mean for sure,
            // and a wasted CIEDE2000 evaluation sits on the critical path of the short folds
            const bool speculate = max(na_, nb_) >= 4 * min(na_, nb_);
            const float4 guess = __ldcg(R.cvec + (big_is_a ? a : b));
            named_bar(BAR_TOUCHED, kFastOwners + kFastDeltaThreads);                           // owners published the touched edges
            const int T = sm.misc[FM_TCOUNT];
            FPHASE(1);
            if (T > kFastMaxTouched) {
                if (d == 0) sm.misc[FM_ERROR] = (int)kFastErrTouched;                          // the host falls back to merge_kernel
                named_bar(BAR_FOLDED, kFastFoldedCount);
                bar_arrive(BAR_RESULTS, kFastOwners + kFastDeltaThreads);
                continue;
            }
            const int counter = sm.misc[FM_COUNTER];
            if (T <= 32) {
                // ---- few touched edges: one warp, registers and shuffles only (no mark, no hash, no barrier) ----
                if (warp == kFastDeltaWarp0) {
                    const bool act = lane < T;
                    unsigned my_hi = kDeadKey, my_lo = kDeadKey, x = 0xffff0000u | (unsigned)lane, side_a = 0;
                    float dc = 0.0f;
                    if (act) { my_hi = sm.te_hi[lane]; my_lo = sm.te_lo[lane]; const unsigned tx = sm.te_x[lane]; x = tx & 0xffffu; side_a = (tx >> 16) & 1u; dc = sm.te_dc[lane]; }
                    unsigned lessmask = 0; bool dup = false;
                    for (int q = 0; q < T; ++q) {                                              // (broadcast reads: cheaper than three shuffles per step)
                        const unsigned qh = sm.te_hi[q], ql = sm.te_lo[q], qx = sm.te_x[q] & 0xffffu;
                        const bool less = key_less32(qh, ql, my_hi, my_lo);
                        if (less) lessmask |= 1u << q;
                        dup = dup || (less && qx == x);                                        // the earlier of (a,x), (b,x) survives
                    }
                    const bool live = act && !dup;
                    FPHASE(7);
                    float4 xcv = make_float4(0, 0, 0, 0), c4 = xcv, n4 = xcv;
                    if (live) { xcv = __ldcg(R.cvec + x); c4 = __ldcg(R.centroid + x); n4 = __ldcg(R.normal + x); }
                    const bool reuse = side_a ? big_is_a : (!big_is_a && ((b < x) == (a < x)));
                    const bool need = live && !reuse;
                    FPHASE(2);
                    __syncwarp();
                    if (speculate && __any_sync(kFull, need)) { if (need) dc = a < x ? colour_delta(ep.color_mode, guess, xcv) : colour_delta(ep.color_mode, xcv, guess); }
                    FPHASE(3);
                    named_bar(BAR_FOLDED, kFastFoldedCount);                                   // region a's new colour vector / centroid / normal
                    FPHASE(4);
                    const bool hit = newgeo_i[9] != 0;
                    const bool redo = hit ? (need && !speculate) : live;                      // wrong guess: every survivor; no guess made: the ones that cannot reuse
                    if (__any_sync(kFull, redo)) {
                        const float4 acv = make_float4(sm.newgeo[0], sm.newgeo[1], sm.newgeo[2], 0.0f);
                        if (redo) dc = a < x ? colour_delta(ep.color_mode, acv, xcv) : colour_delta(ep.color_mode, xcv, acv);
                    }
                    FPHASE(5);
                    __syncwarp();
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
                        sm.res_hi[lane] = wbits; sm.res_lo[lane] = lo; sm.res_ab[lane] = ab; sm.te_dc[lane] = dc;
                    }
                    if (lane == 0) {
                        sm.misc[FM_COUNTER] = counter + nbk + nfr;
                        sm.misc[FM_EALIVE] -= 1 + __popc(dupm); sm.misc[FM_RALIVE] -= 1;
                        if (T > sm.misc[FM_MAXT]) sm.misc[FM_MAXT] = T;
                        sm.misc[FM_SUMT] += T; sm.misc[FM_TCOUNT] = 0;
                        sm.misc[FM_EVALS] += __popc(needm) + (hit ? 0 : T - __popc(dupm)); sm.misc[FM_MISS] += hit ? 0 : 1;
                    }
                } else {
                    named_bar(BAR_FOLDED, kFastFoldedCount);
                }
            } else {
                // ---- A: duplicates (a,x)/(b,x) meet through a per-region mark ----
                for (int p = d; p < T; p += kFastDeltaThreads) {
                    const unsigned x = sm.te_x[p] & 0xffffu;
                    const unsigned short old = atomicCAS(&sm.mark[x], (unsigned short)kNil16, (unsigned short)p);
                    if (old != (unsigned short)kNil16) { sm.partner[p] = old; sm.partner[old] = (unsigned short)p; }
                }
                named_bar(BAR_DELTA, kFastDeltaThreads);
                // ---- B: the earlier of two duplicates survives; survivors fetch x's geometry and queue for CIEDE unless their
                //         stored colour delta stays valid under the guess ----
                for (int p0 = 0; p0 < T; p0 += kFastDeltaThreads) {
                    const int p = p0 + d;
                    bool need = false;
                    if (p < T) {
                        const unsigned tx = sm.te_x[p], x = tx & 0xffffu, side_a = (tx >> 16) & 1u;
                        const unsigned q = sm.partner[p];
                        const bool dup = q != kNil16 && key_less32(sm.te_hi[q], sm.te_lo[q], sm.te_hi[p], sm.te_lo[p]);
                        sm.mark[x] = (unsigned short)kNil16;
                        if (dup) sm.cls[p] = FC_DUP;
                        else {
                            sm.te_ce[p] = __ldcg(R.centroid + x); sm.te_nr[p] = __ldcg(R.normal + x);
                            const bool reuse = side_a ? big_is_a : (!big_is_a && ((b < x) == (a < x)));   // same ends' colours, same argument order
                            sm.cls[p] = reuse ? FC_REUSE : FC_KEEP;
                            need = !reuse;
                        }
                    }
                    const unsigned m = __ballot_sync(kFull, need);
                    if (m) {
                        int base = 0;
                        if (lane == __ffs(m) - 1) base = atomicAdd(&sm.misc[FM_NEED], __popc(m));
                        base = __shfl_sync(kFull, base, __ffs(m) - 1);
                        if (need) sm.need[base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)p;
                    }
                }
                named_bar(BAR_DELTA, kFastDeltaThreads);
                FPHASE(2);
                // ---- C: speculative colour deltas against the guess (overlaps the fold) ----
                const int n_need = sm.misc[FM_NEED];
    #pragma unroll 1
                for (int i = d; i < n_need; i += kFastDeltaThreads) {
                    const int p = sm.need[i];
                    const unsigned x = sm.te_x[p] & 0xffffu;
                    const float4 xcv = __ldcg(R.cvec + x);
                    sm.te_dc[p] = a < x ? colour_delta(ep.color_mode, guess, xcv) : colour_delta(ep.color_mode, xcv, guess);
                }
                FPHASE(3);
                named_bar(BAR_FOLDED, kFastFoldedCount);                                           // region a's new colour vector / centroid / normal
                FPHASE(4);
                const float4 acv = make_float4(sm.newgeo[0], sm.newgeo[1], sm.newgeo[2], 0.0f);
                if (!newgeo_i[9]) {                                                                // wrong guess: every survivor re-evaluates
    #pragma unroll 1
                    for (int p = d; p < T; p += kFastDeltaThreads) {
                        if ((sm.cls[p] & 3) == FC_DUP) continue;
                        const unsigned x = sm.te_x[p] & 0xffffu;
                        const float4 xcv = __ldcg(R.cvec + x);
                        sm.te_dc[p] = a < x ? colour_delta(ep.color_mode, acv, xcv) : colour_delta(ep.color_mode, xcv, acv);
                    }
                }
                named_bar(BAR_DELTA, kFastDeltaThreads);
                FPHASE(5);
                // ---- D: geometry delta, weight, classification against the old weight; tie groups (same new weight, same side) ----
                const float4 ace = make_float4(sm.newgeo[3], sm.newgeo[4], sm.newgeo[5], 0.0f), anr = make_float4(sm.newgeo[6], sm.newgeo[7], sm.newgeo[8], 0.0f);
    #pragma unroll 1
                for (int p = d; p < T; p += kFastDeltaThreads) {
                    sm.partner[p] = (unsigned short)kNil16;
                    if ((sm.cls[p] & 3) == FC_DUP) { sm.cls[p] = FC_DUP; sm.res_hi[p] = kDeadKey; sm.res_lo[p] = kDeadKey; sm.res_ab[p] = 0u; atomicAdd(&sm.misc[FM_ND], 1); continue; }
                    const unsigned x = sm.te_x[p] & 0xffffu;
                    const float4 c4 = sm.te_ce[p], n4 = sm.te_nr[p];
                    const bool a_first = a < x;
                    const float dg = geom_delta(ep.geom_mode, a_first ? anr : n4, a_first ? ace : c4, a_first ? n4 : anr, a_first ? c4 : ace);
                    float w_new = unify(ep, sm.te_dc[p], dg);
                    if (isnan(w_new)) { atomicAdd(&sm.misc[FM_NANW], 1); w_new = __int_as_float(0x7f800000); }
                    const unsigned wbits = __float_as_uint(w_new), old_hi = sm.te_hi[p];
                    const int cls = wbits == old_hi ? FC_KEEP : (wbits > old_hi ? FC_FRONT : FC_BACK);
                    sm.cls[p] = (unsigned char)cls; sm.res_hi[p] = wbits; sm.res_ab[p] = a < x ? (a << 16) | x : (x << 16) | a; sm.res_lo[p] = sm.te_lo[p];
                    if (cls != FC_KEEP) {
                        const unsigned key = wbits | (cls == FC_FRONT ? 0x80000000u : 0u);
                        unsigned h = (key * 2654435761u) >> 21;
                        while (true) {
                            const unsigned prev = atomicCAS(&sm.hkey[h], kDeadKey, key);
                            if (prev == kDeadKey || prev == key) break;
                            h = (h + 1) & (kFastHash - 1);
                        }
                        atomicAdd(&sm.hcnt[h], 1u);
                        sm.partner[p] = (unsigned short)h;                                        // (the duplicate pass is done with partner[])
                    }
                }
                named_bar(BAR_DELTA, kFastDeltaThreads);
                // ---- E: tie stamps: new arrivals keep their old relative order inside a tie group ----
    #pragma unroll 1
                for (int p = d; p < T; p += kFastDeltaThreads) {
                    const unsigned h = sm.partner[p];
                    if (h == kNil16) continue;
                    const int cls = sm.cls[p];
                    const unsigned gsz = sm.hcnt[h];
                    unsigned rank = 0;
                    if (gsz > 1) {
                        const unsigned wbits = sm.res_hi[p], ph = sm.te_hi[p], pl = sm.te_lo[p];
                        for (int q = 0; q < T; ++q)
                            if (sm.cls[q] == cls && sm.res_hi[q] == wbits && key_less32(sm.te_hi[q], sm.te_lo[q], ph, pl)) ++rank;
                    }
                    const int st = cls == FC_BACK ? counter + (int)rank : -(counter + (int)(gsz - 1 - rank));
                    sm.res_lo[p] = (unsigned)st ^ 0x80000000u;
                }
                named_bar(BAR_DELTA, kFastDeltaThreads);
    #pragma unroll 1
                for (int p = d; p < T; p += kFastDeltaThreads) { const unsigned h = sm.partner[p]; if (h != kNil16) { sm.hkey[h] = kDeadKey; sm.hcnt[h] = 0u; } }
                if (d == 0) {
                    sm.misc[FM_COUNTER] = counter + T;
                    sm.misc[FM_EALIVE] -= 1 + sm.misc[FM_ND]; sm.misc[FM_ND] = 0; sm.misc[FM_RALIVE] -= 1;
                    if (T > sm.misc[FM_MAXT]) sm.misc[FM_MAXT] = T;
                    sm.misc[FM_SUMT] += T; sm.misc[FM_TCOUNT] = 0;
                    sm.misc[FM_EVALS] += n_need + (newgeo_i[9] ? 0 : T); sm.misc[FM_MISS] += newgeo_i[9] ? 0 : 1; sm.misc[FM_NEED] = 0;
                }

            }
            __syncwarp();
            FPHASE(6);
            bar_arrive(BAR_RESULTS, kFastOwners + kFastDeltaThreads);
            ++n_merges;
        }
        FPROF_STORE(d == 0, 0, 8);
#undef FPHASE
    } else {
        // =========== owners: the weight map itself, in registers =====
        // Blocked assignment: owner warp w holds edges [w * 32 * SLOTS, (w + 1) * 32 * SLOTS), slot j / lane l = edge
        // base + 32 j + l.  The initial edges are sorted by (a, b), so the edges of a region sit in one or two warps (plus
        // the warps of its lower-labelled neighbours); wmask[] lets every other warp skip the scan of a merge entirely.
        const int ow = warp - kFastRoleWarps;                                               // 0..24
        const unsigned ebase = (unsigned)ow * 32u * SLOTS + (unsigned)lane;
        unsigned khi[SLOTS], klo[SLOTS], kab[SLOTS];
        unsigned pending = 0;                                 // slots whose new key is waiting in res_*[klo[slot]]
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
                sm.dc[e] = A.E.dc[e];
                atomicOr(&sm.wmask[ea], 1u << ow); atomicOr(&sm.wmask[eb], 1u << ow);
            }
        }
        bool dirty = true;
        unsigned l_hi = kDeadKey, l_lo = kDeadKey, l_ab = 0xffffffffu; int l_slot = 0;
        const bool probe = ow == 0 && lane == 0;
#define OPROF(i) do { if (probe) { const unsigned t_ = (unsigned)clock(); sm.prof[i] += t_ - sm.prof[15]; sm.prof[15] = t_; } } while (0)
        __syncthreads();
        if (probe) sm.prof[15] = (unsigned)clock();
        while (true) {
            // ---- take the previous merge's results, local minimum, warp minimum ----
            if (pending) {                                     // a new key below the cached minimum replaces it; the minimum's own slot
#pragma unroll                                                 // moving up (or dying) asks for a rescan
                for (int j = 0; j < SLOTS; ++j)
                    if (pending & (1u << j)) {
                        const unsigned p = klo[j];
                        khi[j] = sm.res_hi[p]; klo[j] = sm.res_lo[p]; kab[j] = khi[j] == kDeadKey ? 0xffffffffu : sm.res_ab[p];
                        sm.dc[ebase + 32u * j] = sm.te_dc[p];
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
                OPROF(0);
                if (lane == win) { sm.wm_hi[ow] = m_hi; sm.wm_lo[ow] = m_lo; sm.wm_e[ow] = ((unsigned)l_slot << 16) | (unsigned)(ow * 32 + lane); sm.wm_ab[ow] = l_ab; }
            }
            __syncthreads();                                                                   // B1
            OPROF(1);
            const FastHead hd = fast_head(sm, lane);
            if (sm.misc[FM_ERROR] || hd.hi == kDeadKey || !(__uint_as_float(hd.hi) < A.threshold)) break;   // strict <, src/clustering.cpp:388-389
            const unsigned a = hd.ab >> 16, b = hd.ab & 0xffffu;
            OPROF(2);
            if (((sm.wmask[a] | sm.wmask[b]) >> ow) & 1u) {                                    // warp-uniform: does this warp hold an incident edge?
                if ((hd.e & 0xffffu) == (unsigned)(ow * 32 + lane)) {                          // the head edge leaves the map
                    const int hs = (int)(hd.e >> 16);
#pragma unroll
                    for (int j = 0; j < SLOTS; ++j) if (j == hs) { khi[j] = kDeadKey; klo[j] = kDeadKey; kab[j] = 0xffffffffu; }
                    dirty = true;
                }
                // ---- edges incident to a or b: both ends of a slot compared at once (dead slots hold 0xffff:0xffff) ----
                const unsigned aa = a * 0x10001u, bb = b * 0x10001u;
                unsigned hits = 0;
#pragma unroll
                for (int j = 0; j < SLOTS; ++j) hits |= ((__vcmpeq2(kab[j], aa) | __vcmpeq2(kab[j], bb)) != 0u ? 1u : 0u) << j;
                if (__any_sync(kFull, hits != 0u)) {
                    const int cnt = __popc(hits);
                    int incl = cnt;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(kFull, incl, off); if (lane >= off) incl += t; }
                    int base = 0;
                    if (lane == 31) base = atomicAdd(&sm.misc[FM_TCOUNT], incl);
                    int p = __shfl_sync(kFull, base, 31) + incl - cnt;
#pragma unroll
                    for (int j = 0; j < SLOTS; ++j) {
                        if (!((hits >> j) & 1u)) continue;
                        if (p < kFastMaxTouched) {
                            const unsigned ea = kab[j] >> 16, eb = kab[j] & 0xffffu;
                            const bool on_a = ea == a || eb == a;
                            const unsigned x = (ea == a || ea == b) ? eb : ea;
                            sm.te_hi[p] = khi[j]; sm.te_lo[p] = klo[j]; sm.te_x[p] = x | (on_a ? 0x10000u : 0u);
                            sm.te_dc[p] = sm.dc[ebase + 32u * j];
                            sm.partner[p] = (unsigned short)kNil16;
                            klo[j] = (unsigned)p; pending |= 1u << j;
                        }
                        ++p;
                    }
                }
            }
            __syncwarp();
            OPROF(3);
            bar_arrive(BAR_TOUCHED, kFastOwners + kFastDeltaThreads);
            if (ow == 0 && lane == 0 && n_merges < A.log_cap) {                                // debug line of :390-392 (ranks; labels at the end)
                A.mlog.a[n_merges] = a; A.mlog.b[n_merges] = b; A.mlog.w[n_merges] = __uint_as_float(hd.hi);
                A.mlog.edges_left[n_merges] = (unsigned)sm.misc[FM_EALIVE]; A.mlog.regions_left[n_merges] = (unsigned)sm.misc[FM_RALIVE];
            }
            named_bar(BAR_RESULTS, kFastOwners + kFastDeltaThreads);                           // new keys are in res_*
            OPROF(4);
            ++n_merges;
        }
#undef OPROF
        // ---- write the weight map back ----
#pragma unroll
        for (int j = 0; j < SLOTS; ++j) {
            const unsigned e = ebase + 32u * j;
            if (e >= nE) continue;
            if (khi[j] == kDeadKey && klo[j] == kDeadKey) A.E.stamp[e] = kDeadStamp;
            else {
                A.E.a[e] = kab[j] >> 16; A.E.b[e] = kab[j] & 0xffffu; A.E.w[e] = __uint_as_float(khi[j]);
                A.E.stamp[e] = (long long)(int)(klo[j] ^ 0x80000000u); A.E.dc[e] = sm.dc[e];
            }
        }
    }

    // ---- write back: ropes, log labels, counters ---------------------------------------------------------------------
    __syncthreads();
    if (warp == 1 && lane == 0) sm.misc[15] = (int)n_merges;
    __syncthreads();
    n_merges = (unsigned)sm.misc[15];
    for (unsigned s = tid; s < S; s += kFastThreads) {
        R.n[s] = sm.n[s];
        R.head[s] = sm.head[s] == kNil16 ? -1 : (int)sm.head[s]; R.tail[s] = sm.tail[s] == kNil16 ? -1 : (int)sm.tail[s];
        R.next_run[s] = sm.next[s] == kNil16 ? -1 : (int)sm.next[s];
    }
    for (unsigned m = tid; m < n_merges && m < A.log_cap; m += kFastThreads) { A.mlog.a[m] = A.sv_label[A.mlog.a[m]]; A.mlog.b[m] = A.sv_label[A.mlog.b[m]]; }
    if (tid == 0) {
        MergeCtl* ctl = A.ctl;
        for (int i = 0; i < 4; ++i) ctl->phase_cycles[20 + i] = sm.prof[i];
        ctl->phase_cycles[28] = sm.prof[4];
        ctl->phase_cycles[24] = (unsigned long long)sm.misc[FM_MISS]; ctl->phase_cycles[25] = (unsigned long long)sm.misc[FM_EVALS];
        ctl->phase_cycles[26] = (unsigned long long)sm.misc[FM_BIGT]; ctl->phase_cycles[27] = (unsigned long long)sm.misc[FM_SUMT];
        ctl->n_merges = n_merges; ctl->edges_alive = (unsigned)sm.misc[FM_EALIVE]; ctl->regions_alive = (unsigned)sm.misc[FM_RALIVE];
        ctl->counter = (long long)sm.misc[FM_COUNTER];
        ctl->max_touched = (unsigned)sm.misc[FM_MAXT]; ctl->nan_weights = (unsigned)sm.misc[FM_NANW]; ctl->error = (unsigned)sm.misc[FM_ERROR];
    }
}

// one frame: one CTA
template <int SLOTS>
__global__ void __launch_bounds__(kFastThreads, 1) merge_fast_kernel(const __grid_constant__ FastArgs A) { merge_fast_body<SLOTS>(A); }

// A batch of frames in ONE launch, CTA i replays frame i (f3ps_merge_batch).  Independent streams share at most 32 hardware
// queues (CUDA_DEVICE_MAX_CONNECTIONS), so at most 32 single-CTA merge kernels ever overlap; one grid has no such limit.
// The per-frame arguments travel in the kernel parameter space (<= 32,764 bytes on sm_70+ with CUDA >= 12.1).
constexpr int kFastBatchMax = 32764 / (int)sizeof(FastArgs) < 96 ? 32764 / (int)sizeof(FastArgs) : 96;
struct FastBatch { FastArgs a[kFastBatchMax]; };
static_assert(sizeof(FastBatch) <= 32764, "kernel parameter space");
template <int SLOTS>
__global__ void __launch_bounds__(kFastThreads, 1) merge_fast_batch_kernel(const __grid_constant__ FastBatch B) { merge_fast_body<SLOTS>(B.a[blockIdx.x]); }

} // namespace f3ps
