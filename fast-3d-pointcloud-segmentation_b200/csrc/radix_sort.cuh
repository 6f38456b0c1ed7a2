// radix_sort.cuh -- hand-written stable LSD radix sort (key,value pairs) and ordered
// stream compaction for sm_100a.
//
// Sort = one histogram kernel for every digit position + one "onesweep" kernel per
// 8-bit digit: each tile ranks its keys with warp match.any (stable), publishes its
// digit counts and resolves its global offsets by decoupled look-back over the tiles
// before it, so keys and values are read and written exactly once per pass
// (algorithmic traffic per pass = 2 * n * (sizeof(Key) + 4) bytes).
// Tile ids come from an atomic ticket so a tile only ever waits on tiles that are
// already running.
//
// Used by K1 (Morton keys of points), K4 (seed cells), K5 (voxels by supervoxel
// label, every round) and K6 (edge list).
#pragma once
#include "common.cuh"

namespace f3ps {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr unsigned kFlagAgg = 1u << 30;
constexpr unsigned kFlagPrefix = 2u << 30;
constexpr unsigned kFlagMask = 3u << 30;
constexpr unsigned kValueMask = ~kFlagMask;

template <typename KeyT>
__global__ void __launch_bounds__(256) radix_histogram_kernel(const KeyT* __restrict__ keys, const unsigned* __restrict__ n_ptr,
                                                              int64_t n_cap, int passes, unsigned* __restrict__ ghist) {
    extern __shared__ unsigned s_hist[];   // passes * 256
    const int64_t n = n_ptr ? min((int64_t)*n_ptr, n_cap) : n_cap;
    for (int i = threadIdx.x; i < passes * kRadix; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        KeyT k = keys[i];
        for (int p = 0; p < passes; ++p) atomicAdd(&s_hist[p * kRadix + (int)((k >> (p * kRadixBits)) & (kRadix - 1))], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * kRadix; i += blockDim.x) {
        unsigned c = s_hist[i];
        if (c) atomicAdd(&ghist[i], c);
    }
}

// exclusive scan of each pass's 256-bin histogram, in place (one block per pass)
__global__ void __launch_bounds__(256) radix_scan_hist_kernel(unsigned* __restrict__ ghist) {
    __shared__ unsigned s[kRadix];
    unsigned* h = ghist + blockIdx.x * kRadix;
    unsigned v = h[threadIdx.x];
    s[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < kRadix; off <<= 1) {
        unsigned t = threadIdx.x >= off ? s[threadIdx.x - off] : 0u;
        __syncthreads();
        s[threadIdx.x] += t;
        __syncthreads();
    }
    h[threadIdx.x] = s[threadIdx.x] - v;
}

template <typename KeyT, int ITEMS>
__global__ void __launch_bounds__(kSortThreads) radix_onesweep_kernel(
        const KeyT* __restrict__ keys_in, const unsigned* __restrict__ vals_in, KeyT* __restrict__ keys_out,
        unsigned* __restrict__ vals_out, const unsigned* __restrict__ n_ptr, int64_t n_cap, int shift,
        const unsigned* __restrict__ ghist_pass, volatile unsigned* tile_state, unsigned* ticket) {
    constexpr int TILE = kSortThreads * ITEMS;
    __shared__ unsigned s_warp_hist[kSortWarps][kRadix];
    __shared__ unsigned s_digit_base[kRadix];
    __shared__ unsigned s_tile;
    const int64_t n = n_ptr ? min((int64_t)*n_ptr, n_cap) : n_cap;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < kSortWarps * kRadix; i += kSortThreads) (&s_warp_hist[0][0])[i] = 0;
    __syncthreads();
    const unsigned tile = s_tile;
    const int64_t tile_base = (int64_t)tile * TILE;
    if (tile_base >= n) return;                       // whole tile past the end (grid sized for n_cap)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int64_t warp_base = tile_base + (int64_t)warp * ITEMS * 32;

    KeyT key[ITEMS];
    unsigned val[ITEMS];
    unsigned short rank[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        int64_t idx = warp_base + j * 32 + lane;
        bool ok = idx < n;
        key[j] = ok ? keys_in[idx] : (KeyT)~(KeyT)0;
        val[j] = ok ? (vals_in ? vals_in[idx] : (unsigned)idx) : 0u;   // vals_in == nullptr: identity payload
    }
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        int64_t idx = warp_base + j * 32 + lane;
        bool ok = idx < n;
        unsigned d = (unsigned)((key[j] >> shift) & (kRadix - 1));
        unsigned dx = ok ? d : kRadix;                // out-of-range lanes form their own group
        unsigned peers = __match_any_sync(kFull, dx);
        int leader = __ffs(peers) - 1;
        unsigned old = 0;
        if (lane == leader && ok) {
            old = s_warp_hist[warp][d];
            s_warp_hist[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(kFull, old, leader);
        rank[j] = (unsigned short)(old + __popc(peers & lt_mask));
        __syncwarp();
    }
    __syncthreads();
    // one thread per digit: exclusive scan over warps, publish, look back
    {
        const int d = threadIdx.x;
        unsigned run = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) { unsigned c = s_warp_hist[w][d]; s_warp_hist[w][d] = run; run += c; }
        volatile unsigned* st = tile_state + (size_t)tile * kRadix + d;
        unsigned excl = 0;
        if (tile == 0) {
            *st = kFlagPrefix | run;
        } else {
            *st = kFlagAgg | run;
            for (int64_t t = (int64_t)tile - 1; t >= 0; --t) {
                volatile unsigned* ps = tile_state + (size_t)t * kRadix + d;
                unsigned v;
                do { v = *ps; } while ((v & kFlagMask) == 0);
                excl += v & kValueMask;
                if ((v & kFlagMask) == kFlagPrefix) break;
            }
            *st = kFlagPrefix | (excl + run);
        }
        s_digit_base[d] = ghist_pass[d] + excl;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        int64_t idx = warp_base + j * 32 + lane;
        if (idx < n) {
            unsigned d = (unsigned)((key[j] >> shift) & (kRadix - 1));
            unsigned pos = s_digit_base[d] + s_warp_hist[warp][d] + rank[j];
            keys_out[pos] = key[j];
            vals_out[pos] = val[j];
        }
    }
}

// ---------------------------------------------------------------------------------
// Ordered stream compaction: for the k-th index i in [0,n) with op.test(i, payload),
// op.emit(k, i, payload); the total goes to *count.  Single pass, decoupled look-back
// (same scheme as above).  Op carries its own arrays (see KeygenOp / HeadOp / KeepOp).
constexpr int kSelThreads = 256;
constexpr int kSelItems = 8;

template <typename Op>
__global__ void __launch_bounds__(kSelThreads) compact_kernel(Op op, const unsigned* __restrict__ n_ptr, int64_t n_cap,
                                                              unsigned* __restrict__ count, volatile unsigned* tile_state,
                                                              unsigned* ticket) {
    constexpr int TILE = kSelThreads * kSelItems;
    __shared__ unsigned s_warp[kSelThreads / 32];
    __shared__ unsigned s_tile, s_excl;
    const int64_t n = n_ptr ? min((int64_t)*n_ptr, n_cap) : n_cap;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    const int64_t base = (int64_t)tile * TILE + (int64_t)threadIdx.x * kSelItems;   // blocked: keeps index order
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    typename Op::Payload pay[kSelItems];
    unsigned flags = 0; int cnt = 0;
    if ((int64_t)tile * TILE < n) {
#pragma unroll
        for (int j = 0; j < kSelItems; ++j) {
            int64_t i = base + j;
            if (i < n && op.test(i, pay[j])) { flags |= 1u << j; ++cnt; }
        }
    }
    unsigned incl = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { unsigned t = __shfl_up_sync(kFull, incl, off); if (lane >= off) incl += t; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kSelThreads / 32; ++w) { unsigned c = s_warp[w]; if (w < warp) wbase += c; total += c; }
    if (threadIdx.x == 0) {
        volatile unsigned* st = tile_state + tile;
        unsigned excl = 0;
        if (tile == 0) *st = kFlagPrefix | total;
        else {
            *st = kFlagAgg | total;
            for (int64_t t = (int64_t)tile - 1; t >= 0; --t) {
                unsigned v;
                do { v = tile_state[t]; } while ((v & kFlagMask) == 0);
                excl += v & kValueMask;
                if ((v & kFlagMask) == kFlagPrefix) break;
            }
            *st = kFlagPrefix | (excl + total);
        }
        s_excl = excl;
        // the tile that covers the last element knows the grand total
        if ((int64_t)tile * TILE < n && (int64_t)(tile + 1) * TILE >= n) *count = excl + total;
        if (n == 0 && tile == 0) *count = 0;
    }
    __syncthreads();
    unsigned pos = s_excl + wbase + (incl - cnt);
#pragma unroll
    for (int j = 0; j < kSelItems; ++j) if (flags & (1u << j)) op.emit(pos++, base + j, pay[j]);
}

} // namespace f3ps
