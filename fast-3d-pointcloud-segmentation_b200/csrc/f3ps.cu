// f3ps.cu -- host orchestration and the C ABI (include/f3ps.h) of the B200-native
// supervoxel-plus-merging path.  One handle = one device + one stream; every stage is a short
// chain of kernels on that stream; the host only waits where an array size must be known.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <condition_variable>
#include <map>
#include <mutex>
#include <new>
#include <set>
#include <utility>

#include "context.cuh"

using namespace f3ps;

extern "C" const unsigned char f3ps_lab_lut_begin[];   // lab_lut.S (.incbin of data/lab_lut_s16.bin)
extern "C" const unsigned char f3ps_lab_lut_end[];

namespace {

int ctx_fail(f3ps_ctx* ctx, int code, const std::string& msg) { ctx->err = msg; return code; }
int ctx_fail_cuda(f3ps_ctx* ctx, cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    ctx->err = buf;
    return F3PS_ERR_CUDA;
}

inline int grid_for(int64_t n, int threads, int max_blocks = kSMs * 16) {
    int64_t b = (n + threads - 1) / threads;
    return (int)std::max<int64_t>(1, std::min<int64_t>(b, max_blocks));
}
inline unsigned next_pow2(unsigned v) { unsigned p = 1; while (p < v) p <<= 1; return p; }
inline int bits_for(unsigned maxval) { int b = 1; while ((maxval >> b) != 0 && b < 32) ++b; return b; }

#define LAUNCH(ctx, kernel, grid, block, smem, ...)                                         \
    do {                                                                                    \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                    \
        (ctx)->launches++;                                                                  \
        F3PS_CUDA_OK(cudaPeekAtLastError());                                                \
    } while (0)

// Wait for the handle's stream.  Spinning (cudaStreamSynchronize) has the lowest latency for one frame; a sweep with more
// frames in flight than host cores must sleep instead (blocking event), or the spinning waiters starve the launching threads.
cudaError_t wait_stream(f3ps_ctx* ctx) {
    if (!ctx->blocking_wait) return cudaStreamSynchronize(ctx->stream);
    cudaError_t e = cudaEventRecord(ctx->ev_wait, ctx->stream);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(ctx->ev_wait);
}

int pull_scalars(f3ps_ctx* ctx) {
    F3PS_CUDA_OK(cudaMemcpyAsync(ctx->h_sc, ctx->d_sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, ctx->stream));
    F3PS_CUDA_OK(wait_stream(ctx));
    return F3PS_OK;
}
#define SC(field) (&ctx->d_sc->field)

int mark(f3ps_ctx* ctx, int i) {
    F3PS_CUDA_OK(cudaEventRecord(ctx->ev[i], ctx->stream));
    ctx->ev_valid[i] = true;
    return F3PS_OK;
}

// ---- radix sort driver ------------------------------------------------------------------------
// pass 0 reads (src_k, src_v) -- src_v == nullptr means the payload is the element index -- and the
// passes ping-pong between scratch pairs A and B (A may alias the source when it can be clobbered).
template <typename KeyT>
int sort_pairs(f3ps_ctx* ctx, const KeyT* src_k, const unsigned* src_v, KeyT* a_k, unsigned* a_v, KeyT* b_k, unsigned* b_v,
               const unsigned* n_ptr, int64_t n_cap, int bits, KeyT** keys_out, unsigned** vals_out) {
    const int passes = std::max(1, (bits + kRadixBits - 1) / kRadixBits);
    const bool small = n_cap < (1 << 20);
    const int tile = kSortThreads * (small ? 4 : 16);
    const int64_t tiles = std::max<int64_t>(1, (n_cap + tile - 1) / tile);
    const size_t hist_words = (size_t)passes * kRadix;
    const size_t state_words = (size_t)passes * tiles * kRadix;
    const size_t words = hist_words + state_words + passes;
    F3PS_CUDA_OK(ctx->sort_scratch.ensure(words * 4));
    unsigned* ghist = ctx->sort_scratch.as<unsigned>();
    unsigned* state = ghist + hist_words;
    unsigned* tickets = state + state_words;
    F3PS_CUDA_OK(cudaMemsetAsync(ghist, 0, words * 4, ctx->stream));
    LAUNCH(ctx, radix_histogram_kernel<KeyT>, grid_for(n_cap, 256 * 8, kSMs * 4), 256, hist_words * 4, src_k, n_ptr, n_cap, passes, ghist);
    LAUNCH(ctx, radix_scan_hist_kernel, passes, kRadix, 0, ghist);
    const KeyT* ki = src_k; const unsigned* vi = src_v;
    KeyT* ko = a_k; unsigned* vo = a_v;
    for (int p = 0; p < passes; ++p) {
        if (small)
            LAUNCH(ctx, (radix_onesweep_kernel<KeyT, 4>), (int)tiles, kSortThreads, 0, ki, vi, ko, vo, n_ptr, n_cap, p * kRadixBits,
                   ghist + (size_t)p * kRadix, state + (size_t)p * tiles * kRadix, tickets + p);
        else
            LAUNCH(ctx, (radix_onesweep_kernel<KeyT, 16>), (int)tiles, kSortThreads, 0, ki, vi, ko, vo, n_ptr, n_cap, p * kRadixBits,
                   ghist + (size_t)p * kRadix, state + (size_t)p * tiles * kRadix, tickets + p);
        ki = ko; vi = vo;
        if (ko == a_k) { ko = b_k; vo = b_v; } else { ko = a_k; vo = a_v; }
    }
    *keys_out = const_cast<KeyT*>(ki); *vals_out = const_cast<unsigned*>(vi);
    return F3PS_OK;
}

template <typename Op>
int run_compact(f3ps_ctx* ctx, Op op, const unsigned* n_ptr, int64_t n_cap, unsigned* count) {
    const int tile = kSelThreads * kSelItems;
    const int64_t tiles = std::max<int64_t>(1, (n_cap + tile - 1) / tile);
    F3PS_CUDA_OK(ctx->compact_scratch.ensure((tiles + 1) * 4));
    unsigned* state = ctx->compact_scratch.as<unsigned>();
    F3PS_CUDA_OK(cudaMemsetAsync(state, 0, (tiles + 1) * 4, ctx->stream));
    LAUNCH(ctx, compact_kernel<Op>, (int)tiles, kSelThreads, 0, op, n_ptr, n_cap, count, state, state + tiles);
    return F3PS_OK;
}

__global__ void iota_kernel(unsigned* p, unsigned n) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = i;
}
__global__ void pos_run_kernel(const unsigned* __restrict__ sorted_label, const unsigned* __restrict__ rank_of_label, unsigned n, unsigned* __restrict__ pos_run) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned l = sorted_label[i];
        pos_run[i] = l ? rank_of_label[l] : 0xffffffffu;
    }
}
__global__ void pos_gather_kernel(const unsigned* __restrict__ order, const float4* __restrict__ vox_xyz, unsigned n, float4* __restrict__ pos_data) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) pos_data[i] = vox_xyz[order[i]];
}
__global__ void decode_edge_keys_kernel(const unsigned long long* __restrict__ compact, unsigned n, int kb, unsigned long long* __restrict__ wide) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long k = compact[i];
        wide[i] = ((k >> kb) << 32) | (k & ((1ull << kb) - 1ull));
    }
}
struct EdgeSlotCompactOp {      // hash-set slots -> compact keys (rank_a << kb | rank_b), unsorted
    const unsigned long long* slots; unsigned long long* keys; unsigned* vals; int kb;
    typedef int Payload;
    __device__ __forceinline__ bool test(int64_t i, Payload&) const { return slots[i] != 0ull; }
    __device__ __forceinline__ void emit(unsigned pos, int64_t i, const Payload&) const {
        const unsigned long long k = slots[i] - 1ull;
        keys[pos] = ((k >> 32) << kb) | (k & 0xffffffffull); vals[pos] = pos;
    }
};
__global__ void lab_test_kernel(const short* lut, const float* rgb, float* lab, int64_t n) {
    for (int64_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float o[3]; rgb2lab(lut, rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], o);
        lab[3 * i] = o[0]; lab[3 * i + 1] = o[1]; lab[3 * i + 2] = o[2];
    }
}
__global__ void ciede_test_kernel(const float* l1, const float* l2, float* out, int64_t n) {
    for (int64_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = lab_ciede00(l1 + 3 * i, l2 + 3 * i);
}
__global__ void rgb_eucl_test_kernel(const float* c1, const float* c2, float* out, int64_t n) {
    for (int64_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = rgb_eucl(c1 + 3 * i, c2 + 3 * i);
}
RegionArrays carve_regions(void* base, size_t S) {
    RegionArrays R; char* p = (char*)base;
    auto take = [&](size_t bytes) { char* q = p; p += (bytes + 255) & ~(size_t)255; return q; };
    R.mean = (float4*)take(S * 16); R.accu0 = (float4*)take(S * 16); R.accu1 = (float4*)take(S * 16); R.accu2 = (float4*)take(S * 16);
    R.centroid = (float4*)take(S * 16); R.normal = (float4*)take(S * 16); R.cvec = (float4*)take(S * 16);
    R.n = (int*)take(S * 4); R.head = (int*)take(S * 4); R.tail = (int*)take(S * 4); R.next_run = (int*)take(S * 4);
    return R;
}
size_t region_bytes(size_t S) { return 7 * ((S * 16 + 255) & ~(size_t)255) + 4 * ((S * 4 + 255) & ~(size_t)255); }
EdgeArrays carve_edges(void* base, size_t E) {
    EdgeArrays A; char* p = (char*)base;
    auto take = [&](size_t bytes) { char* q = p; p += (bytes + 255) & ~(size_t)255; return q; };
    A.stamp = (long long*)take(E * 8);
    A.a = (unsigned*)take(E * 4); A.b = (unsigned*)take(E * 4); A.dc = (float*)take(E * 4); A.dg = (float*)take(E * 4); A.w = (float*)take(E * 4);
    return A;
}
size_t edge_bytes(size_t E) { return ((E * 8 + 255) & ~(size_t)255) + 5 * ((E * 4 + 255) & ~(size_t)255); }

EdgeParams edge_params(const f3ps_ctx* ctx) {
    EdgeParams ep;
    ep.color_mode = ctx->mp.color_mode; ep.geom_mode = ctx->mp.geom_mode; ep.merge_mode = ctx->mp.merge_mode;
    ep.bins = ctx->mp.bins; ep.lambda = ctx->mp.lambda;
    ep.cdf_c = ctx->cdf_c.as<float>(); ep.cdf_g = ctx->cdf_g.as<float>(); ep.lab_lut = ctx->d_lab_lut;
    return ep;
}

int need(f3ps_ctx* ctx, int level, const char* what) {
    if (ctx->progress < level) return ctx_fail(ctx, F3PS_ERR_LOGIC, std::string(what) + ": earlier stage has not run");
    return F3PS_OK;
}

// ---- K1 -------------------------------------------------------------------------------------------
template <typename KeyT>
int voxelize_typed(f3ps_ctx* ctx, PointLoader pl) {
    const int64_t N = ctx->n_points;
    F3PS_CUDA_OK(ctx->keys_a.ensure(N * sizeof(KeyT))); F3PS_CUDA_OK(ctx->keys_b.ensure(N * sizeof(KeyT)));
    F3PS_CUDA_OK(ctx->vals_a.ensure(N * 4)); F3PS_CUDA_OK(ctx->vals_b.ensure(N * 4));
    F3PS_CUDA_OK(ctx->starts.ensure((N + 1) * 4));
    KeygenOp<KeyT> kop{pl, ctx->vp.use_transform, SC(fp), ctx->keys_a.as<KeyT>(), ctx->vals_a.as<unsigned>()};
    int rc = run_compact(ctx, kop, nullptr, N, SC(n_valid));
    if (rc) return rc;
    KeyT* sk; unsigned* sv;
    rc = sort_pairs<KeyT>(ctx, ctx->keys_a.as<KeyT>(), ctx->vals_a.as<unsigned>(), ctx->keys_b.as<KeyT>(), ctx->vals_b.as<unsigned>(),
                          ctx->keys_a.as<KeyT>(), ctx->vals_a.as<unsigned>(), SC(n_valid), N, 3 * ctx->depth, &sk, &sv);
    if (rc) return rc;
    ctx->sorted_keys = sk; ctx->sorted_idx = sv;
    HeadOp<KeyT> hop{sk, ctx->starts.as<unsigned>()};
    rc = run_compact(ctx, hop, SC(n_valid), N, SC(n_voxels));
    if (rc) return rc;
    rc = pull_scalars(ctx);                                   // host needs V to size the voxel arrays
    if (rc) return rc;
    ctx->V = ctx->h_sc->n_voxels; ctx->n_valid = ctx->h_sc->n_valid;
    const size_t V = std::max(1u, ctx->V);
    F3PS_CUDA_OK(ctx->vox_xyz.ensure(V * 16)); F3PS_CUDA_OK(ctx->vox_rgb.ensure(V * 16)); F3PS_CUDA_OK(ctx->vox_key.ensure(V * 8));
    if (ctx->V)
        LAUNCH(ctx, voxel_accumulate_kernel<KeyT>, grid_for(ctx->V, 128), 128, 0, pl, sk, sv, ctx->starts.as<unsigned>(), SC(n_voxels), SC(n_valid),
               ctx->vox_xyz.as<float4>(), ctx->vox_rgb.as<float4>(), ctx->vox_key.as<uint64_t>(), ctx->point_voxel.as<int>());
    return F3PS_OK;
}

} // namespace

// =====================================================================================================
extern "C" {

const char* f3ps_version(void) { return "f3ps-b200 0.1 (sm_100a)"; }

int f3ps_create(int device, void* stream, f3ps_ctx** out) {
    if (!out) return F3PS_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    // Frames of a sweep run on independent handles / streams; the default of 8 hardware work queues would serialise
    // their persistent merge kernels 4-8 at a time.  Only effective when set before the CUDA context exists.
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return F3PS_ERR_CUDA;   // no CPU fallback
    f3ps_ctx* ctx = new (std::nothrow) f3ps_ctx();
    if (!ctx) return F3PS_ERR_CUDA;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return F3PS_ERR_CUDA; }
    if (stream) ctx->stream = (cudaStream_t)stream;
    else { if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return F3PS_ERR_CUDA; } ctx->private_stream = ctx->stream; }
    bool ok = cudaMalloc(&ctx->d_sc, sizeof(DevScalars)) == cudaSuccess &&
              cudaMallocHost(&ctx->h_sc, sizeof(DevScalars)) == cudaSuccess;
    const size_t lut_bytes = (size_t)(f3ps_lab_lut_end - f3ps_lab_lut_begin);
    ok = ok && lut_bytes == 33 * 33 * 33 * 3 * 2 && cudaMalloc(&ctx->d_lab_lut, lut_bytes) == cudaSuccess &&
         cudaMemcpy(ctx->d_lab_lut, f3ps_lab_lut_begin, lut_bytes, cudaMemcpyHostToDevice) == cudaSuccess;
    for (int i = 0; i < f3ps_ctx::kEvents && ok; ++i) ok = cudaEventCreate(&ctx->ev[i]) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_wait, cudaEventBlockingSync | cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_batch, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { f3ps_destroy(ctx); return F3PS_ERR_CUDA; }
    // the handle's stream is non-blocking: it does not wait for the legacy stream the two copies above ran on
    cudaMemsetAsync(ctx->d_sc, 0, sizeof(DevScalars), ctx->stream);
    if (cudaDeviceSynchronize() != cudaSuccess) { f3ps_destroy(ctx); return F3PS_ERR_CUDA; }
    memset(ctx->h_sc, 0, sizeof(DevScalars));
    *out = ctx;
    return F3PS_OK;
}

void f3ps_destroy(f3ps_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    // (device buffers are DevBuf members: `delete ctx` below releases every one of them)
    if (ctx->d_sc) cudaFree(ctx->d_sc);
    if (ctx->h_sc) cudaFreeHost(ctx->h_sc);
    if (ctx->d_lab_lut) cudaFree(ctx->d_lab_lut);
    for (int i = 0; i < f3ps_ctx::kEvents; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->ev_wait) cudaEventDestroy(ctx->ev_wait);
    if (ctx->ev_batch) cudaEventDestroy(ctx->ev_batch);
    if (ctx->private_stream) cudaStreamDestroy(ctx->private_stream);
    delete ctx;
}

const char* f3ps_last_error(const f3ps_ctx* ctx) { return ctx ? ctx->err.c_str() : "null handle"; }
int64_t f3ps_launch_count(const f3ps_ctx* ctx) { return ctx ? ctx->launches : 0; }

int f3ps_set_vccs_params(f3ps_ctx* ctx, float rv, float rs, float wc, float ws, float wn, int use_transform, int fold_negative_z) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    if (!(rv > 0) || !(rs > 0)) return ctx_fail(ctx, F3PS_ERR_INVALID_ARGUMENT, "resolutions must be positive");
    ctx->vp = VccsParams{rv, rs, wc, ws, wn, use_transform ? 1 : 0, fold_negative_z ? 1 : 0};
    ctx->slab_frame = false; ctx->slab_range = false;
    ctx->progress = std::min(ctx->progress, (int)P_INPUT);
    return F3PS_OK;
}

// Clustering::set_merging / set_lambda / set_bins_num, src/clustering.cpp:562-597
int f3ps_set_merge_params(f3ps_ctx* ctx, int color, int geom, int merging, float lambda, int bins) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    if (color < 0 || color > 1 || geom < 0 || geom > 1 || merging < 0 || merging > 2)
        return ctx_fail(ctx, F3PS_ERR_INVALID_ARGUMENT, "unknown distance / merging enum value");
    if (merging == F3PS_MANUAL_LAMBDA && (lambda < 0 || lambda > 1 || lambda != lambda))
        return ctx_fail(ctx, F3PS_ERR_INVALID_ARGUMENT, "Argument outside range [0, 1]");
    if (merging == F3PS_EQUALIZATION && (bins <= 0 || bins > 32767))
        return ctx_fail(ctx, F3PS_ERR_INVALID_ARGUMENT, "Argument lower than 0");
    ctx->mp = MergeParams{color, geom, merging, merging == F3PS_MANUAL_LAMBDA ? lambda : 0.5f, merging == F3PS_EQUALIZATION ? bins : 500};
    if (ctx->progress >= P_GRAPH) ctx->progress = ctx->graph_from_host ? (int)P_GRAPH - 1 : (int)P_EXPANDED;   // weights are stale
    return F3PS_OK;
}

int f3ps_set_input(f3ps_ctx* ctx, const void* points, int64_t n, int stride, int on_device) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    if (n < 0 || (stride != 16 && stride != 32) || (n > 0 && !points)) return ctx_fail(ctx, F3PS_ERR_INVALID_ARGUMENT, "bad input description");
    if (n >= (1ll << 31)) return ctx_fail(ctx, F3PS_ERR_CAPACITY, "more than 2^31 points per handle");
    cudaSetDevice(ctx->device);
    if (on_device) ctx->d_points = (const uint8_t*)points;
    else {
        F3PS_CUDA_OK(ctx->in_buf.ensure(std::max<int64_t>(n, 1) * stride));
        if (n) F3PS_CUDA_OK(cudaMemcpyAsync(ctx->in_buf.p, points, (size_t)n * stride, cudaMemcpyHostToDevice, ctx->stream));
        ctx->d_points = ctx->in_buf.as<uint8_t>();
    }
    ctx->n_points = n; ctx->stride = stride;
    ctx->progress = P_INPUT; ctx->graph_from_host = false;
    return F3PS_OK;
}

int f3ps_sync(f3ps_ctx* ctx) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    F3PS_CUDA_OK(wait_stream(ctx));
    return F3PS_OK;
}

// ---- K1 -------------------------------------------------------------------------------------------
int f3ps_voxelize(f3ps_ctx* ctx) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_INPUT, "f3ps_voxelize"); if (rc) return rc;
    cudaSetDevice(ctx->device);
    rc = mark(ctx, 0); if (rc) return rc;
    const int64_t N = ctx->n_points;
    PointLoader pl{ctx->d_points, ctx->stride, ctx->vp.fold_negative_z};
    DevScalars init; memset(&init, 0, sizeof init);
    for (int a = 0; a < 3; ++a) { init.fp.ord_min[a] = 0xffffffffu; init.fp.ord_max[a] = 0u; }
    if (ctx->slab_frame) init.fp = ctx->slab_fp;
    *ctx->h_sc = init;
    F3PS_CUDA_OK(cudaMemcpyAsync(ctx->d_sc, ctx->h_sc, sizeof(DevScalars), cudaMemcpyHostToDevice, ctx->stream));
    F3PS_CUDA_OK(ctx->point_voxel.ensure(std::max<int64_t>(N, 1) * 4));
    F3PS_CUDA_OK(cudaMemsetAsync(ctx->point_voxel.p, 0xff, std::max<int64_t>(N, 1) * 4, ctx->stream));
    ctx->V = 0; ctx->n_valid = 0; ctx->depth = 0;
    ctx->slab_range = false;
    if (ctx->slab_frame) {                                    // slab mode: the frame of the WHOLE cloud, agreed on by all ranks
        ctx->depth = ctx->slab_fp.depth;
        if (N > 0 && ctx->slab_fp.any_finite) {
            rc = ctx->key64 ? voxelize_typed<uint64_t>(ctx, pl) : voxelize_typed<uint32_t>(ctx, pl);
            if (rc) return rc;
        }
    } else if (N > 0) {
        LAUNCH(ctx, bbox_kernel, grid_for(N, 256 * 4, kSMs * 8), 256, 0, pl, N, ctx->vp.use_transform, SC(fp));
        LAUNCH(ctx, frame_setup_kernel, 1, 1, 0, SC(fp), ctx->vp.voxel_res);
        rc = pull_scalars(ctx); if (rc) return rc;            // host needs the depth to pick the key width / pass count
        ctx->depth = ctx->h_sc->fp.depth;
        if (ctx->h_sc->fp.any_finite) {
            if (ctx->depth > 21) return ctx_fail(ctx, F3PS_ERR_CAPACITY, "adjacency octree depth > 21");
            ctx->key64 = 3 * ctx->depth > 32;
            rc = ctx->key64 ? voxelize_typed<uint64_t>(ctx, pl) : voxelize_typed<uint32_t>(ctx, pl);
            if (rc) return rc;
        }
    }
    ctx->progress = P_VOXELS;
    return mark(ctx, 1);
}

// ---- K2 -------------------------------------------------------------------------------------------
int f3ps_neighbors(f3ps_ctx* ctx) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_VOXELS, "f3ps_neighbors"); if (rc) return rc;
    cudaSetDevice(ctx->device);
    const unsigned V = ctx->V; const size_t Vc = std::max(1u, V);
    const unsigned cap = next_pow2(std::max(64u, 2 * V));
    ctx->hash_mask = cap - 1;
    F3PS_CUDA_OK(ctx->hash_slots.ensure((size_t)cap * 8)); F3PS_CUDA_OK(ctx->hash_vals.ensure((size_t)cap * 4));
    F3PS_CUDA_OK(ctx->nbr_row.ensure(Vc * kNbrStride * 4)); F3PS_CUDA_OK(ctx->nbr_col.ensure(Vc * 27 * 4));
    F3PS_CUDA_OK(cudaMemsetAsync(ctx->hash_slots.p, 0, (size_t)cap * 8, ctx->stream));
    if (V) {
        LAUNCH(ctx, hash_build_kernel, grid_for(V, 256), 256, 0, ctx->vox_key.as<uint64_t>(), SC(n_voxels),
               ctx->hash_slots.as<unsigned long long>(), ctx->hash_vals.as<unsigned>(), ctx->hash_mask);
        LAUNCH(ctx, neighbors_kernel, grid_for(V, 128), 128, 0, ctx->vox_key.as<uint64_t>(), SC(n_voxels), SC(fp),
               ctx->hash_slots.as<unsigned long long>(), ctx->hash_vals.as<unsigned>(), ctx->hash_mask, ctx->nbr_row.as<int>(),
               ctx->nbr_col.as<int>(), V);
    }
    ctx->progress = P_NEIGHBORS;
    return mark(ctx, 2);
}

// ---- K3 -------------------------------------------------------------------------------------------
int f3ps_normals(f3ps_ctx* ctx) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_NEIGHBORS, "f3ps_normals"); if (rc) return rc;
    cudaSetDevice(ctx->device);
    const unsigned V = ctx->V; const size_t Vc = std::max(1u, V);
    F3PS_CUDA_OK(ctx->vox_normal.ensure(Vc * 16)); F3PS_CUDA_OK(ctx->vox_curv.ensure(Vc * 4));
    const unsigned vb = ctx->slab_range ? ctx->own_begin : 0u, ve = ctx->slab_range ? ctx->own_end : 0xffffffffu;
    if (V && ve > vb)
        LAUNCH(ctx, voxel_normals_kernel, grid_for(std::min(V, ve) - vb, 128), 128, 0, ctx->vox_xyz.as<float4>(), ctx->nbr_row.as<int>(), SC(n_voxels),
               ctx->vox_normal.as<float4>(), ctx->vox_curv.as<float>(), vb, ve);
    ctx->progress = P_NORMALS;
    return mark(ctx, 3);
}

// ---- K4 -------------------------------------------------------------------------------------------
int f3ps_seeds(f3ps_ctx* ctx) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_NORMALS, "f3ps_seeds"); if (rc) return rc;
    cudaSetDevice(ctx->device);
    const unsigned V = ctx->V; const size_t Vc = std::max(1u, V);
    ctx->n_cells = 0; ctx->S0 = 0;
    DevBuf* b8[] = {&ctx->cell_code, &ctx->cell_code_b, &ctx->vox_cell, &ctx->cell_codes};
    for (DevBuf* b : b8) F3PS_CUDA_OK(b->ensure(Vc * 8));
    DevBuf* b4[] = {&ctx->cell_vox, &ctx->cell_vox_b, &ctx->cell_start, &ctx->cell_nn, &ctx->cell_keep, &ctx->seeds};
    for (DevBuf* b : b4) F3PS_CUDA_OK(b->ensure((Vc + 1) * 4));
    if (V) {
        LAUNCH(ctx, seed_box_kernel, 1, 1024, 0, ctx->vox_xyz.as<float4>(), SC(n_voxels), ctx->vp.seed_res, SC(sb));
        LAUNCH(ctx, seed_cell_kernel, grid_for(V, 256), 256, 0, ctx->vox_xyz.as<float4>(), SC(n_voxels), SC(sb), ctx->vox_cell.as<uint64_t>());
        rc = pull_scalars(ctx); if (rc) return rc;            // seed octree depth decides the sort width
        const SeedBox& sb = ctx->h_sc->sb;
        if (sb.n_events > kMaxSeedEvents) return ctx_fail(ctx, F3PS_ERR_CAPACITY, "seed octree grew more than 48 times");
        if (sb.depth > 21) return ctx_fail(ctx, F3PS_ERR_CAPACITY, "seed octree depth > 21");
        uint64_t* sc; unsigned* sv;
        rc = sort_pairs<uint64_t>(ctx, ctx->vox_cell.as<uint64_t>(), nullptr, ctx->cell_code.as<uint64_t>(), ctx->cell_vox.as<unsigned>(),
                                  ctx->cell_code_b.as<uint64_t>(), ctx->cell_vox_b.as<unsigned>(), SC(n_voxels), V, 3 * sb.depth, &sc, &sv);
        if (rc) return rc;
        HeadOp<uint64_t> hop{sc, ctx->cell_start.as<unsigned>()};
        rc = run_compact(ctx, hop, SC(n_voxels), V, SC(n_cells)); if (rc) return rc;
        LAUNCH(ctx, gather_cell_codes_kernel, grid_for(V, 256), 256, 0, sc, ctx->cell_start.as<unsigned>(), SC(n_cells), ctx->cell_codes.as<uint64_t>());
        LAUNCH(ctx, seed_select_kernel, grid_for((int64_t)V * 4, 256), 256, 0, ctx->vox_xyz.as<float4>(), ctx->vox_cell.as<uint64_t>(), sv,
               ctx->cell_start.as<unsigned>(), ctx->cell_codes.as<uint64_t>(), SC(n_cells), SC(n_voxels), SC(sb), ctx->vp.seed_res,
               ctx->vp.voxel_res, ctx->cell_nn.as<int>(), ctx->cell_keep.as<unsigned>());
        KeepOp kop{ctx->cell_keep.as<unsigned>(), ctx->cell_nn.as<int>(), ctx->seeds.as<int>()};
        rc = run_compact(ctx, kop, SC(n_cells), V, SC(n_seeds)); if (rc) return rc;
        rc = pull_scalars(ctx); if (rc) return rc;            // S0 sizes the per-label tables
        ctx->n_cells = ctx->h_sc->n_cells; ctx->S0 = ctx->h_sc->n_seeds;
    }
    ctx->progress = P_SEEDS;
    return mark(ctx, 4);
}

// ---- K5 -------------------------------------------------------------------------------------------
// One cooperative launch runs createSupervoxelHelpers, every expansion round and makeSupervoxels' lists
// (kernels_expand.cuh); the host only learns the number of surviving helpers afterwards.
// Sweeps: cooperative launches run ONE AT A TIME on a device (measured: K5 of concurrent frames serialises at its solo
// duration, 0.74 ms per frame, while K1..K4 and K6 scale with the streams), so a sweep may instead launch the expansion
// kernel as an ordinary small grid (f3ps_set_expand_sharing).  Its software grid barrier needs all CTAs of a launch resident:
// that holds as long as the CTAs of all expansion kernels in flight fit the SMs no long-running kernel occupies, which the
// per-device counting semaphore below enforces (the caller states the bound).
namespace {
struct ExpandGate { std::mutex m; std::condition_variable cv; int in_flight = 0; int limit = 1; };
ExpandGate g_expand_gate[16];
struct ExpandTicket {
    ExpandGate* g = nullptr;
    explicit ExpandTicket(ExpandGate* gate) : g(gate) { std::unique_lock<std::mutex> l(g->m); g->cv.wait(l, [&] { return g->in_flight < g->limit; }); ++g->in_flight; }
    ~ExpandTicket() { { std::lock_guard<std::mutex> l(g->m); --g->in_flight; } g->cv.notify_one(); }
};
}

static int expand_prepare(f3ps_ctx* ctx, ExpandArgs& A) {
    const unsigned V = ctx->V, S0 = ctx->S0; const size_t Vc = std::max(1u, V), Sc = (size_t)S0 + 2, Lc = Vc + Sc;
    DevBuf* bv[] = {&ctx->own_a, &ctx->own_b, &ctx->dst_a, &ctx->dst_b, &ctx->st0, &ctx->st1, &ctx->owner0, &ctx->dist0};
    for (DevBuf* b : bv) F3PS_CUDA_OK(b->ensure(Vc * 4));
    F3PS_CUDA_OK(ctx->phantom.ensure(Vc * 4 * kPhSlots));
    DevBuf* bl[] = {&ctx->lab_keys_a, &ctx->lab_vals_a, &ctx->lab_vals_b};
    for (DevBuf* b : bl) F3PS_CUDA_OK(b->ensure(Lc * 4));
    F3PS_CUDA_OK(ctx->chg_a.ensure(Vc)); F3PS_CUDA_OK(ctx->chg_b.ensure(Vc));
    DevBuf* bs[] = {&ctx->seg_start, &ctx->seg_end, &ctx->lab_count, &ctx->lab_count2, &ctx->lab_fill, &ctx->phantom_leaf, &ctx->sv_label, &ctx->rank_of_label};
    for (DevBuf* b : bs) F3PS_CUDA_OK(b->ensure(Sc * 4));
    F3PS_CUDA_OK(ctx->cen_xyz.ensure(Sc * 16)); F3PS_CUDA_OK(ctx->cen_rgb.ensure(Sc * 16)); F3PS_CUDA_OK(ctx->cen_nrm.ensure(Sc * 16));
    const int max_depth = (int)(1.8f * ctx->vp.seed_res / ctx->vp.voxel_res);     // SupervoxelClustering::extract
    ctx->rounds = std::max(0, max_depth - 1);
    F3PS_CUDA_OK(cudaMemsetAsync(SC(xctl), 0, sizeof(ExpandCtl), ctx->stream));
    ctx->sorted_label = ctx->lab_keys_a.as<unsigned>(); ctx->sorted_vox = ctx->lab_vals_b.as<unsigned>();
    A.nbr_col = ctx->nbr_col.as<int>(); A.nbr_row = ctx->nbr_row.as<int>(); A.V_cap = V;
    A.vox_xyz = ctx->vox_xyz.as<float4>(); A.vox_rgb = ctx->vox_rgb.as<float4>(); A.vox_nrm = ctx->vox_normal.as<float4>();
    A.seeds = ctx->seeds.as<int>(); A.V = V; A.S0 = S0; A.rounds = ctx->rounds; A.P = ctx->vp; A.keep_centroids = 0;
    A.owner[0] = ctx->own_a.as<unsigned>(); A.owner[1] = ctx->own_b.as<unsigned>();
    A.dist[0] = ctx->dst_a.as<float>(); A.dist[1] = ctx->dst_b.as<float>();
    A.st[0] = ctx->st0.as<unsigned>(); A.st[1] = ctx->st1.as<unsigned>();
    A.chg[0] = ctx->chg_a.as<unsigned char>(); A.chg[1] = ctx->chg_b.as<unsigned char>();
    A.phantom = ctx->phantom.as<unsigned>(); A.phantom_leaf = ctx->phantom_leaf.as<int>();
    A.cen = Centroids{ctx->cen_xyz.as<float4>(), ctx->cen_rgb.as<float4>(), ctx->cen_nrm.as<float4>()};
    A.count[0] = ctx->lab_count.as<unsigned>(); A.count[1] = ctx->lab_count2.as<unsigned>(); A.fill = ctx->lab_fill.as<unsigned>(); A.off = ctx->seg_start.as<unsigned>();
    A.list_raw = ctx->lab_vals_a.as<unsigned>(); A.list_sorted = ctx->lab_vals_b.as<unsigned>(); A.pos_label = ctx->lab_keys_a.as<unsigned>();
    A.labels_out = ctx->owner0.as<unsigned>(); A.dist_out = ctx->dist0.as<float>();
    A.seg_end = ctx->seg_end.as<unsigned>(); A.sv_label = ctx->sv_label.as<unsigned>(); A.rank_of_label = ctx->rank_of_label.as<unsigned>();
    A.ctl = SC(xctl);
    return F3PS_OK;
}

static int expand_finish(f3ps_ctx* ctx) {
    int rc = pull_scalars(ctx); if (rc) return rc;
    const ExpandCtl& x = ctx->h_sc->xctl;
    if (x.error & EXPAND_ERR_TRIPLE) return ctx_fail(ctx, F3PS_ERR_CAPACITY, "more than four seed cells elected the same voxel (not modelled)");
    if (x.error & EXPAND_ERR_CAND) return ctx_fail(ctx, F3PS_ERR_CAPACITY, "more than 32 candidate helpers around one voxel");
    if (x.error & EXPAND_ERR_SWEEPS) return ctx_fail(ctx, F3PS_ERR_NOT_CONVERGED, "expansion fixed point not reached within 32 sweeps");
    ctx->S = ctx->V ? x.n_sv : 0; ctx->n_pos = ctx->V ? x.n_pos : 0;
    ctx->progress = P_EXPANDED;
    return mark(ctx, 5);
}

static int expand_launch(f3ps_ctx* ctx, bool keep_centroids) {
    const unsigned V = ctx->V;
    ExpandArgs A;
    int rc = expand_prepare(ctx, A); if (rc) return rc;
    A.keep_centroids = keep_centroids ? 1 : 0;
    if (keep_centroids) A.seeds = ctx->seeds_refine.as<int>();      // the seed voxels of f3ps_seeds stay what f3ps_get_seeds returns
    const bool cluster = V && ctx->expand_kernel_choice == 2;      // sweeps ask for it (f3ps_set_expand_kernel); one frame alone is fastest on the cooperative grid
    if (cluster) {
        // one thread-block cluster per frame (kernels_expand.cuh): ~2 voxels per thread, at most 16 CTAs
        int ncta = (int)std::min<int64_t>(kExpandClusterMax, ((int64_t)V + 2 * kExpandClusterThreads - 1) / (2 * kExpandClusterThreads));
        if (ctx->expand_cluster_ctas > 0) ncta = std::min(ncta, ctx->expand_cluster_ctas);
        ncta = std::max(1, ncta);
        if (!ctx->expand_cluster_attr_set) {
            F3PS_CUDA_OK(cudaFuncSetAttribute((const void*)expand_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            ctx->expand_cluster_attr_set = true;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(ncta); cfg.blockDim = dim3(kExpandClusterThreads); cfg.dynamicSmemBytes = 0; cfg.stream = ctx->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = (unsigned)ncta; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        F3PS_CUDA_OK(cudaLaunchKernelEx(&cfg, expand_cluster_kernel, A));
        ctx->launches++;
        ctx->expand_path = 2;
        return expand_finish(ctx);
    }
    ctx->expand_path = 1;
    if (V && ctx->expand_ctas > 0 && !ctx->expand_coop_cap) {
        // shared mode: a small ordinary grid, bounded concurrency (see ExpandGate); held until the kernel has finished
        ExpandTicket ticket(&g_expand_gate[ctx->device & 15]);
        const int64_t want = ((int64_t)V + kExpandThreads - 1) / kExpandThreads;
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, ctx->expand_ctas));
        expand_persistent_kernel<<<grid, kExpandThreads, 0, ctx->stream>>>(A);
        ctx->launches++;
        F3PS_CUDA_OK(cudaPeekAtLastError());
        return expand_finish(ctx);
    }
    if (V) {
        if (!ctx->expand_blocks_per_sm) {
            int nb = 0;
            F3PS_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, expand_persistent_kernel, kExpandThreads, 0));
            int sms = 0;
            F3PS_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
            ctx->expand_blocks_per_sm = std::max(1, nb); ctx->sm_count = std::max(1, sms);
        }
        const int64_t want = ((int64_t)V + kExpandThreads - 1) / kExpandThreads;
        int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)std::min(ctx->expand_blocks_per_sm, 2) * ctx->sm_count));
        if (ctx->expand_coop_cap && ctx->expand_ctas > 0) grid = std::min(grid, ctx->expand_ctas);   // sweeps: fewer CTAs per frame, more frames side by side
        void* args[] = {&A};
        F3PS_CUDA_OK(cudaLaunchCooperativeKernel((const void*)expand_persistent_kernel, dim3(grid), dim3(kExpandThreads), args, 0, ctx->stream));
        ctx->launches++;
    }
    return expand_finish(ctx);
}

int f3ps_expand(f3ps_ctx* ctx) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_SEEDS, "f3ps_expand"); if (rc) return rc;
    cudaSetDevice(ctx->device);
    return expand_launch(ctx, false);
}

// pcl::SupervoxelClustering::refineSupervoxels(num_itr, ...): per iteration refineNormals, reseedSupervoxels and the expansion
// rounds again, starting from the helpers' current centroids (kernels_vccs.cuh, kernels_expand.cuh).  Leaves the handle where
// f3ps_expand leaves it: labels, distances, voxel normals and the supervoxels are the refined ones, the graph has to be rebuilt.
int f3ps_refine(f3ps_ctx* ctx, int num_itr) {
    if (!ctx || num_itr < 0) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_EXPANDED, "f3ps_refine"); if (rc) return rc;          // PCL: "Supervoxels must be extracted before they can be refined"
    if (ctx->graph_from_host) return ctx_fail(ctx, F3PS_ERR_LOGIC, "f3ps_refine needs the voxels of f3ps_extract, not a graph set with f3ps_set_graph");
    if (ctx->slab_range) return ctx_fail(ctx, F3PS_ERR_LOGIC, "f3ps_refine is not available in slab mode");
    cudaSetDevice(ctx->device);
    const unsigned V = ctx->V, S0 = ctx->S0;
    F3PS_CUDA_OK(ctx->seeds_refine.ensure(((size_t)S0 + 1) * 4));
    for (int it = 0; it < num_itr && V && S0; ++it) {
        LAUNCH(ctx, refine_normals_kernel, grid_for(V, 128), 128, 0, ctx->vox_xyz.as<float4>(), ctx->nbr_row.as<int>(), SC(n_voxels),
               ctx->owner0.as<unsigned>(), ctx->phantom.as<unsigned>(), ctx->vox_normal.as<float4>(), ctx->vox_curv.as<float>());
        LAUNCH(ctx, reseed_kernel, (int)std::min<unsigned>(S0, 4u * kSMs), 256, 0, ctx->cen_xyz.as<float4>(), S0, ctx->vox_xyz.as<float4>(), SC(n_voxels),
               ctx->seeds_refine.as<int>());
        ctx->progress = P_SEEDS;
        rc = expand_launch(ctx, true); if (rc) return rc;
    }
    ctx->progress = P_EXPANDED;
    return F3PS_OK;
}

// ---- K6 -------------------------------------------------------------------------------------------
static int init_weights(f3ps_ctx* ctx, unsigned E_cap) {     // Clustering::init_weights on R0 / E0 + the edge keys
    EdgeParams ep = edge_params(ctx);
    const float* lambda_dev = nullptr;
    F3PS_CUDA_OK(ctx->dbits_a.ensure((size_t)E_cap * 4)); F3PS_CUDA_OK(ctx->dbits_b.ensure((size_t)E_cap * 4));
    F3PS_CUDA_OK(ctx->dbits_c.ensure((size_t)E_cap * 4)); F3PS_CUDA_OK(ctx->dbits_d.ensure((size_t)E_cap * 4));
    F3PS_CUDA_OK(ctx->edge_vals_a.ensure((size_t)E_cap * 4)); F3PS_CUDA_OK(ctx->edge_vals_b.ensure((size_t)E_cap * 4));
    F3PS_CUDA_OK(ctx->cdf_c.ensure((size_t)std::max(1, ctx->mp.bins) * 4)); F3PS_CUDA_OK(ctx->cdf_g.ensure((size_t)std::max(1, ctx->mp.bins) * 4));
    ep.cdf_c = ctx->cdf_c.as<float>(); ep.cdf_g = ctx->cdf_g.as<float>();
    LAUNCH(ctx, region_cvec_kernel, grid_for(std::max(1u, ctx->S), 256), 256, 0, SC(xctl.n_sv), ctx->R0, ep);
    LAUNCH(ctx, edge_delta_kernel, grid_for(E_cap, 128), 128, 0, ctx->sorted_edge_keys, SC(n_edges), ctx->R0, ep, ctx->E0,
           ctx->dbits_a.as<unsigned>(), ctx->dbits_c.as<unsigned>());
    if (ctx->mp.merge_mode == F3PS_ADAPTIVE_LAMBDA && E_cap <= 16384u) {
        const unsigned n_pow2 = next_pow2(std::max(2u, E_cap));
        if (!ctx->lambda_attr_set) {
            F3PS_CUDA_OK(cudaFuncSetAttribute(adaptive_lambda_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 4));
            ctx->lambda_attr_set = true;
        }
        LAUNCH(ctx, adaptive_lambda_smem_kernel, 2, 1024, (size_t)n_pow2 * 4, ctx->E0.dc, ctx->E0.dg, SC(n_edges), n_pow2, SC(padf[0]), SC(pad0), SC(lambda));
        lambda_dev = SC(lambda);
    } else if (ctx->mp.merge_mode == F3PS_ADAPTIVE_LAMBDA) {
        unsigned *sc, *sg, *dummy;
        int rc = sort_pairs<unsigned>(ctx, ctx->dbits_a.as<unsigned>(), nullptr, ctx->dbits_b.as<unsigned>(), ctx->edge_vals_a.as<unsigned>(),
                                      ctx->dbits_a.as<unsigned>(), ctx->edge_vals_b.as<unsigned>(), SC(n_edges), E_cap, 32, &sc, &dummy);
        if (rc) return rc;
        rc = sort_pairs<unsigned>(ctx, ctx->dbits_c.as<unsigned>(), nullptr, ctx->dbits_d.as<unsigned>(), ctx->edge_vals_a.as<unsigned>(),
                                  ctx->dbits_c.as<unsigned>(), ctx->edge_vals_b.as<unsigned>(), SC(n_edges), E_cap, 32, &sg, &dummy);
        if (rc) return rc;
        LAUNCH(ctx, adaptive_lambda_kernel, 1, 64, 0, sc, sg, SC(n_edges), SC(lambda));
        lambda_dev = SC(lambda);
    } else if (ctx->mp.merge_mode == F3PS_EQUALIZATION) {
        const int bins = ctx->mp.bins;
        F3PS_CUDA_OK(ctx->cdf_hist.ensure((size_t)2 * bins * 4));
        LAUNCH(ctx, cdf_kernel, 2, 1024, bins <= 8192 ? (size_t)bins * 4 : 0, ctx->E0.dc, ctx->E0.dg, SC(n_edges), bins, ctx->cdf_c.as<float>(),
               ctx->cdf_g.as<float>(), ctx->cdf_hist.as<unsigned>(), SC(bad_bin));
    }
    if (ctx->mp.merge_mode != F3PS_ADAPTIVE_LAMBDA) {
        const float lam = ctx->mp.lambda;
        F3PS_CUDA_OK(cudaMemcpyAsync(SC(lambda), &lam, 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    LAUNCH(ctx, edge_weight_kernel, grid_for(E_cap, 256), 256, 0, SC(n_edges), ep, lambda_dev, ctx->E0, SC(nan_weights_init));
    return F3PS_OK;
}

int f3ps_graph(f3ps_ctx* ctx) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    if (ctx->graph_from_host) {                                // weights only (parameters changed after f3ps_set_graph)
        int rc = need(ctx, P_GRAPH - 1, "f3ps_graph"); if (rc) return rc;
        F3PS_CUDA_OK(cudaMemsetAsync(SC(nan_weights_init), 0, 4, ctx->stream));
        rc = init_weights(ctx, std::max(1u, ctx->E)); if (rc) return rc;
        ctx->progress = P_GRAPH;
        return F3PS_OK;
    }
    int rc = need(ctx, P_EXPANDED, "f3ps_graph"); if (rc) return rc;
    const unsigned V = ctx->V, S = ctx->S; const size_t Sc = std::max(1u, S), Vc = std::max(1u, V);
    F3PS_CUDA_OK(ctx->reg_init.ensure(region_bytes(Sc))); F3PS_CUDA_OK(ctx->reg_work.ensure(region_bytes(Sc)));
    ctx->R0 = carve_regions(ctx->reg_init.p, Sc); ctx->R1 = carve_regions(ctx->reg_work.p, Sc);
    F3PS_CUDA_OK(ctx->run_start.ensure(Sc * 4)); F3PS_CUDA_OK(ctx->run_end.ensure(Sc * 4)); F3PS_CUDA_OK(ctx->pos_run.ensure((Vc + Sc) * 4));
    Centroids cen{ctx->cen_xyz.as<float4>(), ctx->cen_rgb.as<float4>(), ctx->cen_nrm.as<float4>()};
    ctx->order = ctx->sorted_vox; ctx->gxyz = ctx->vox_xyz.as<float4>();
    const unsigned P = ctx->n_pos;
    const unsigned set_cap = next_pow2(std::max(1024u, 32 * S));
    ctx->edge_set_mask = set_cap - 1; ctx->edge_kb = bits_for(std::max(1u, S));
    F3PS_CUDA_OK(ctx->edge_set.ensure((size_t)set_cap * 8));
    F3PS_CUDA_OK(ctx->edge_keys_a.ensure((size_t)set_cap * 8)); F3PS_CUDA_OK(ctx->edge_keys_b.ensure((size_t)set_cap * 8));
    F3PS_CUDA_OK(ctx->edge_vals_a.ensure((size_t)set_cap * 4)); F3PS_CUDA_OK(ctx->edge_vals_b.ensure((size_t)set_cap * 4));
    F3PS_CUDA_OK(ctx->edge_init.ensure(edge_bytes(set_cap))); F3PS_CUDA_OK(ctx->edge_work.ensure(edge_bytes(set_cap)));
    ctx->E0 = carve_edges(ctx->edge_init.p, set_cap); ctx->E1 = carve_edges(ctx->edge_work.p, set_cap);
    F3PS_CUDA_OK(cudaMemsetAsync(ctx->edge_set.p, 0, (size_t)set_cap * 8, ctx->stream));
    F3PS_CUDA_OK(cudaMemsetAsync(SC(n_edges), 0, 4, ctx->stream));
    F3PS_CUDA_OK(cudaMemsetAsync(SC(edge_overflow), 0, 16, ctx->stream));   // edge_overflow, bad_bin, nan_weights_init, pad
    if (V && S) {
        LAUNCH(ctx, pos_run_kernel, grid_for(P, 256), 256, 0, ctx->sorted_label, ctx->rank_of_label.as<unsigned>(), P, ctx->pos_run.as<unsigned>());
        F3PS_CUDA_OK(ctx->pos_data_buf.ensure((size_t)std::max(1u, P) * 16));
        LAUNCH(ctx, pos_gather_kernel, grid_for(P, 256), 256, 0, ctx->sorted_vox, ctx->vox_xyz.as<float4>(), P, ctx->pos_data_buf.as<float4>());
        ctx->pos_data = ctx->pos_data_buf.as<float4>();
        LAUNCH(ctx, region_init_kernel, grid_for((int64_t)S * 32, 256), 256, 0, ctx->sv_label.as<unsigned>(), SC(xctl.n_sv), ctx->seg_start.as<unsigned>(),
               ctx->seg_end.as<unsigned>(), ctx->sorted_vox, ctx->vox_xyz.as<float4>(), cen, ctx->R0, ctx->run_start.as<unsigned>(),
               ctx->run_end.as<unsigned>());
        LAUNCH(ctx, edge_collect_kernel, grid_for(P, 256), 256, 0, ctx->nbr_col.as<int>(), V, ctx->nbr_row.as<int>(), P, ctx->sorted_label, ctx->sorted_vox,
               ctx->owner0.as<unsigned>(), ctx->rank_of_label.as<unsigned>(), ctx->edge_set.as<unsigned long long>(), ctx->edge_set_mask, SC(edge_overflow));
        EdgeSlotCompactOp sop{ctx->edge_set.as<unsigned long long>(), ctx->edge_keys_a.as<unsigned long long>(), ctx->edge_vals_a.as<unsigned>(), ctx->edge_kb};
        rc = run_compact(ctx, sop, nullptr, set_cap, SC(n_edges)); if (rc) return rc;
        unsigned long long* sk; unsigned* sv;
        rc = sort_pairs<unsigned long long>(ctx, ctx->edge_keys_a.as<unsigned long long>(), nullptr, ctx->edge_keys_b.as<unsigned long long>(),
                                            ctx->edge_vals_b.as<unsigned>(), ctx->edge_keys_a.as<unsigned long long>(), ctx->edge_vals_a.as<unsigned>(),
                                            SC(n_edges), set_cap, 2 * ctx->edge_kb, &sk, &sv);
        if (rc) return rc;
        unsigned long long* wide = (sk == ctx->edge_keys_a.as<unsigned long long>()) ? ctx->edge_keys_b.as<unsigned long long>()
                                                                                      : ctx->edge_keys_a.as<unsigned long long>();
        rc = pull_scalars(ctx); if (rc) return rc;             // E (and the overflow flag) before the weights
        if (ctx->h_sc->edge_overflow) return ctx_fail(ctx, F3PS_ERR_CAPACITY, "supervoxel edge set overflow");
        ctx->E = ctx->h_sc->n_edges;
        if (ctx->E) {
            LAUNCH(ctx, decode_edge_keys_kernel, grid_for(ctx->E, 256), 256, 0, sk, ctx->E, ctx->edge_kb, wide);
            ctx->sorted_edge_keys = wide;
            rc = init_weights(ctx, ctx->E); if (rc) return rc;
        }
    } else ctx->E = 0;
    ctx->progress = P_GRAPH;
    return mark(ctx, 6);
}

int f3ps_set_graph(f3ps_ctx* ctx, int64_t n_voxels, const float* voxel_xyz, const uint32_t* voxel_rgba, int32_t n_sv, const uint32_t* labels,
                   const int64_t* voxel_offsets, const float* centroids_xyz, const float* normals_xyz, int64_t n_adj, const uint32_t* adj_pairs) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    if (n_voxels < 0 || n_sv < 0 || n_adj < 0) return ctx_fail(ctx, F3PS_ERR_INVALID_ARGUMENT, "negative size");
    cudaSetDevice(ctx->device);
    const unsigned V = (unsigned)n_voxels, S = (unsigned)n_sv;
    // supervoxels by ascending label (std::map order), voxel ranges regrouped accordingly
    std::vector<int> perm(S);
    for (unsigned s = 0; s < S; ++s) perm[s] = (int)s;
    std::sort(perm.begin(), perm.end(), [&](int x, int y) { return labels[x] < labels[y]; });
    std::map<uint32_t, unsigned> rank;
    std::vector<float4> h_xyz(std::max(1u, V));
    std::vector<unsigned> h_start(std::max(1u, S)), h_end(std::max(1u, S)), h_label(std::max(1u, S)), h_posrun(std::max(1u, V));
    std::vector<float4> h_cen(std::max(1u, S)), h_nrm(std::max(1u, S));
    unsigned pos = 0;
    for (unsigned r = 0; r < S; ++r) {
        const int s = perm[r];
        if (rank.count(labels[s])) return ctx_fail(ctx, F3PS_ERR_INVALID_ARGUMENT, "duplicate supervoxel label");
        rank[labels[s]] = r; h_label[r] = labels[s];
        h_start[r] = pos;
        for (int64_t i = voxel_offsets[s]; i < voxel_offsets[s + 1]; ++i, ++pos) {
            float w; uint32_t c = voxel_rgba[i] & 0x00ffffffu; memcpy(&w, &c, 4);
            h_xyz[pos] = make_float4(voxel_xyz[3 * i], voxel_xyz[3 * i + 1], voxel_xyz[3 * i + 2], w);
            h_posrun[pos] = r;
        }
        h_end[r] = pos;
        h_cen[r] = make_float4(centroids_xyz[3 * s], centroids_xyz[3 * s + 1], centroids_xyz[3 * s + 2], 0.0f);
        h_nrm[r] = make_float4(normals_xyz[3 * s], normals_xyz[3 * s + 1], normals_xyz[3 * s + 2], 0.0f);
    }
    if (pos != V) return ctx_fail(ctx, F3PS_ERR_INVALID_ARGUMENT, "voxel offsets do not cover the voxel arrays");
    // clear_adjacency keeps first <= second (src/clustering.cpp:476-486); iteration order = insertion order of the weights
    std::vector<unsigned long long> h_keys;
    for (int64_t i = 0; i < n_adj; ++i) {
        const uint32_t a = adj_pairs[2 * i], b = adj_pairs[2 * i + 1];
        if (a > b) continue;
        if (a == b) return ctx_fail(ctx, F3PS_ERR_INVALID_ARGUMENT, "self-adjacent supervoxel");
        auto ia = rank.find(a), ib = rank.find(b);
        if (ia == rank.end() || ib == rank.end()) return ctx_fail(ctx, F3PS_ERR_INVALID_ARGUMENT, "adjacency names an unknown label");   // map::at
        h_keys.push_back(((unsigned long long)ia->second << 32) | ib->second);
    }
    const unsigned E = (unsigned)h_keys.size();
    const size_t Sc = std::max(1u, S), Vc = std::max(1u, V), Ec = std::max(1u, E);
    F3PS_CUDA_OK(ctx->vox_xyz.ensure(Vc * 16)); F3PS_CUDA_OK(ctx->lab_vals_a.ensure(Vc * 4)); F3PS_CUDA_OK(ctx->pos_run.ensure(Vc * 4));
    F3PS_CUDA_OK(ctx->run_start.ensure(Sc * 4)); F3PS_CUDA_OK(ctx->run_end.ensure(Sc * 4)); F3PS_CUDA_OK(ctx->sv_label.ensure(Sc * 4));
    F3PS_CUDA_OK(ctx->reg_init.ensure(region_bytes(Sc))); F3PS_CUDA_OK(ctx->reg_work.ensure(region_bytes(Sc)));
    ctx->R0 = carve_regions(ctx->reg_init.p, Sc); ctx->R1 = carve_regions(ctx->reg_work.p, Sc);
    F3PS_CUDA_OK(ctx->edge_init.ensure(edge_bytes(Ec))); F3PS_CUDA_OK(ctx->edge_work.ensure(edge_bytes(Ec)));
    ctx->E0 = carve_edges(ctx->edge_init.p, Ec); ctx->E1 = carve_edges(ctx->edge_work.p, Ec);
    F3PS_CUDA_OK(ctx->edge_keys_a.ensure(Ec * 8));
    auto up = [&](void* d, const void* h, size_t bytes) { return bytes ? cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->stream) : cudaSuccess; };
    F3PS_CUDA_OK(up(ctx->vox_xyz.p, h_xyz.data(), (size_t)V * 16)); F3PS_CUDA_OK(up(ctx->pos_run.p, h_posrun.data(), (size_t)V * 4));
    F3PS_CUDA_OK(up(ctx->run_start.p, h_start.data(), (size_t)S * 4)); F3PS_CUDA_OK(up(ctx->run_end.p, h_end.data(), (size_t)S * 4));
    F3PS_CUDA_OK(up(ctx->sv_label.p, h_label.data(), (size_t)S * 4));
    F3PS_CUDA_OK(up(ctx->R0.centroid, h_cen.data(), (size_t)S * 16)); F3PS_CUDA_OK(up(ctx->R0.normal, h_nrm.data(), (size_t)S * 16));
    F3PS_CUDA_OK(up(ctx->edge_keys_a.p, h_keys.data(), (size_t)E * 8));
    DevScalars init; memset(&init, 0, sizeof init);
    init.n_voxels = V; init.xctl.n_sv = S; init.n_edges = E;
    *ctx->h_sc = init;
    F3PS_CUDA_OK(cudaMemcpyAsync(ctx->d_sc, ctx->h_sc, sizeof(DevScalars), cudaMemcpyHostToDevice, ctx->stream));
    if (V) LAUNCH(ctx, iota_kernel, grid_for(V, 256), 256, 0, ctx->lab_vals_a.as<unsigned>(), V);
    if (S) LAUNCH(ctx, region_init_ranges_kernel, grid_for((int64_t)S * 32, 256), 256, 0, S, ctx->run_start.as<unsigned>(), ctx->run_end.as<unsigned>(),
                  ctx->lab_vals_a.as<unsigned>(), ctx->vox_xyz.as<float4>(), ctx->R0);
    ctx->V = V; ctx->S = S; ctx->S0 = S; ctx->E = E; ctx->n_pos = V;
    ctx->order = ctx->lab_vals_a.as<unsigned>(); ctx->sorted_vox = ctx->lab_vals_a.as<unsigned>(); ctx->gxyz = ctx->vox_xyz.as<float4>();
    ctx->pos_data = ctx->vox_xyz.as<float4>();               // identity order
    ctx->sorted_edge_keys = ctx->edge_keys_a.as<unsigned long long>();
    ctx->graph_from_host = true;
    ctx->progress = P_GRAPH - 1;
    F3PS_CUDA_OK(cudaStreamSynchronize(ctx->stream));        // host staging vectors go out of scope
    return f3ps_graph(ctx);
}

// ---- K7 -------------------------------------------------------------------------------------------
extern "C++" {
namespace {
// edge capacity of the resident kernel's tables: E rounded up to whole blocks of 32 edges (0 = more than 32 blocks per worker warp)
unsigned lean_ecap_for(unsigned E, int threads = kFastThreads) {
    const unsigned owners = (unsigned)(threads - 32 * kLeanRoleWarps);
    const unsigned nb = std::max(1u, (E + owners - 1u) / owners);
    return nb <= 32u ? std::max(32u, (E + 31u) & ~31u) : 0u;
}
// one frame alone runs the 768-thread build (80 registers) when its edges fit 21 worker warps
void (*lean_kernel_for(bool prof, bool solo))(FastArgs) {
    if (solo) return prof ? merge_fast_kernel<true, kFastThreadsSolo> : merge_fast_kernel<false, kFastThreadsSolo>;
    return prof ? merge_fast_kernel<true, kFastThreads> : merge_fast_kernel<false, kFastThreads>;
}
// adjacency pool of the resident kernel (edge ids, 2 bytes each): the initial lists plus room for the lists that outgrow their block
constexpr size_t kLeanPoolSlack = 1u << 20;
int lean_pool(f3ps_ctx* ctx, unsigned E, FastArgs& A, bool big = false) {      // (edge ids: 2 bytes, BIG variant 4)
    const size_t entries = 2 * (size_t)E + (big ? 8 : 1) * kLeanPoolSlack;
    F3PS_CUDA_OK(ctx->adj_pool.ensure(entries * (big ? 4 : 2)));
    A.adj_pool = ctx->adj_pool.as<unsigned short>(); A.pool_cap = (unsigned)entries;
    return F3PS_OK;
}
// The opt-in to > 48 KB of dynamic shared memory is per function AND per device: remember it per (device, function).
int lean_attr(f3ps_ctx* ctx, const void* kern) {
    static std::mutex m;
    static std::set<std::pair<int, const void*>> done;
    std::lock_guard<std::mutex> l(m);
    if (done.count({ctx->device, kern})) return F3PS_OK;
    F3PS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    done.insert({ctx->device, kern});
    return F3PS_OK;
}
}
}
int f3ps_merge(f3ps_ctx* ctx, float threshold) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    if (ctx->progress < P_EXPANDED)                            // std::logic_error of Clustering::cluster (src/clustering.cpp:671-673)
        return ctx_fail(ctx, F3PS_ERR_LOGIC, "Cannot call 'cluster' before setting an initial state with 'set_initialstate'");
    cudaSetDevice(ctx->device);
    int rc;
    if (ctx->progress < P_GRAPH) { rc = f3ps_graph(ctx); if (rc) return rc; }   // if (!init_initial_weights) init_weights()  (:675-676)
    rc = mark(ctx, 7); if (rc) return rc;
    const unsigned S = ctx->S, E = ctx->E, P = ctx->n_pos; const size_t Sc = std::max(1u, S), Ec = std::max(1u, E), Pc = std::max(1u, P);
    // cluster(float) always restarts from initial_state (:678) -- unless f3ps_merge_batch hands over a replay its grid stopped in front
    // of a merge with too many adjacency entries (the working state and the counters are then those after n merges)
    const bool take_over = ctx->merge_take_over;
    ctx->merge_take_over = false;
    const size_t eb = std::min(ctx->edge_init.cap, ctx->edge_work.cap);
    if (!take_over) {
        F3PS_CUDA_OK(cudaMemcpyAsync(ctx->reg_work.p, ctx->reg_init.p, region_bytes(Sc), cudaMemcpyDeviceToDevice, ctx->stream));
        F3PS_CUDA_OK(cudaMemcpyAsync(ctx->edge_work.p, ctx->edge_init.p, eb, cudaMemcpyDeviceToDevice, ctx->stream));
        F3PS_CUDA_OK(cudaMemsetAsync(SC(mctl), 0, sizeof(MergeCtl), ctx->stream));
    }
    F3PS_CUDA_OK(ctx->mlog.ensure(Sc * 20));
    char* lp = (char*)ctx->mlog.p;
    ctx->ML.a = (unsigned*)lp; ctx->ML.b = (unsigned*)(lp + Sc * 4); ctx->ML.w = (float*)(lp + Sc * 8);
    ctx->ML.edges_left = (unsigned*)(lp + Sc * 12); ctx->ML.regions_left = (unsigned*)(lp + Sc * 16);
    F3PS_CUDA_OK(ctx->run_out_off.ensure(Sc * 4)); F3PS_CUDA_OK(ctx->run_dense.ensure(Sc * 4)); F3PS_CUDA_OK(ctx->region_dense.ensure(Sc * 4));
    F3PS_CUDA_OK(ctx->out_xyz.ensure(Pc * 12)); F3PS_CUDA_OK(ctx->out_label.ensure(Pc * 4)); F3PS_CUDA_OK(ctx->out_voxel.ensure(Pc * 4));
    F3PS_CUDA_OK(ctx->vox_segment.ensure(Pc * 4));
    F3PS_CUDA_OK(cudaMemsetAsync(ctx->vox_segment.p, 0xff, Pc * 4, ctx->stream));
    F3PS_CUDA_OK(cudaMemsetAsync(SC(n_out), 0, 4, ctx->stream));
    if (S) {
        EdgeParams ep = edge_params(ctx);
        // resident kernel when the graph fits one SM (kernels_merge_lean.cuh), else the general one
        const unsigned S_cap = (S + 7u) & ~7u;
        const unsigned E_cap = std::max(32u, lean_ecap_for(E));
        const size_t fast_bytes = FastSmem(nullptr, S_cap, E_cap).bytes;
        const bool single_ok = lean_ecap_for(E) && S <= 4096u && fast_bytes <= 227u * 1024u && P > 0 && ctx->merge_kernel_choice != 3 && ctx->merge_kernel_choice != 5;
        // graphs that do not fit an SM: the same kernel with its big tables in global memory (16-bit region ids, 32-bit edge ids;
        // 16 bytes of shared memory per block of 32 edges bound E near 380,000)
        const unsigned E_big = (E + 31u) & ~31u;
        const bool big_ok = !single_ok && P > 0 && S < 65535u && E_big <= (1u << 20) && ctx->merge_kernel_choice != 1 &&
                            FastSmem(nullptr, S_cap, E_big, (char*)16).bytes <= 227u * 1024u;
        bool fast = !ctx->force_general_merge && (single_ok || big_ok);
        // a graph that fits an SM but has a merge with more adjacency entries than worker threads continues on the L2 variant (which
        // loops over the entries) instead of handing single merges to the general kernel
        const bool can_switch = single_ok && P > 0 && ctx->merge_kernel_choice != 1 && FastSmem(nullptr, S_cap, E_big, (char*)16).bytes <= 227u * 1024u;
        bool use_big = !single_ok;
        // general kernel (any graph); resume: continue from the state the resident kernel left; stop_after: hand back after that many merges
        auto launch_general = [&](bool resume, unsigned stop_after) -> int {
            const size_t per = ((size_t)Ec * 4 + 255) & ~(size_t)255, per8 = ((size_t)Ec * 8 + 255) & ~(size_t)255;
            unsigned n2 = 2048; while (n2 < Ec) n2 <<= 1;
            const size_t perS = ((size_t)Sc * 4 + 255) & ~(size_t)255, sortb = (size_t)n2 * sizeof(MergeSortRec);
            F3PS_CUDA_OK(ctx->merge_scratch.ensure(6 * per + 2 * per8 + per + per8 + perS + sortb));
            char* sp = (char*)ctx->merge_scratch.p;
            MergeScratch scr;
            scr.st[0] = (long long*)sp; sp += per8; scr.st[1] = (long long*)sp; sp += per8;
            scr.sortbuf = (MergeSortRec*)sp; sp += sortb;
            scr.e[0] = (int*)sp; sp += per; scr.e[1] = (int*)sp; sp += per; scr.w[0] = (float*)sp; sp += per; scr.w[1] = (float*)sp; sp += per;
            scr.x[0] = (unsigned*)sp; sp += per; scr.x[1] = (unsigned*)sp; sp += per; scr.cls = (unsigned char*)sp; sp += per;
            scr.mark = (unsigned*)sp; sp += perS;
            const size_t dyn = (size_t)kMergeSortSmem * sizeof(MergeSortRec);
            if (!ctx->general_attr_set) {
                F3PS_CUDA_OK(cudaFuncSetAttribute(merge_kernel<unsigned>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
                F3PS_CUDA_OK(cudaFuncSetAttribute(merge_kernel<unsigned long long>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
                ctx->general_attr_set = true;
            }
            if (S < 65536u)
                LAUNCH(ctx, merge_kernel<unsigned>, 1, kMergeThreads, dyn, ctx->R1, ctx->E1, SC(n_edges), SC(xctl.n_sv), ep, SC(lambda), threshold, ctx->run_start.as<unsigned>(),
                       ctx->run_end.as<unsigned>(), ctx->order, ctx->gxyz, ctx->sv_label.as<unsigned>(), ctx->ML, (unsigned)Sc, SC(mctl), scr, (unsigned*)sp, ctx->pos_data, resume ? 1 : 0, stop_after);
            else
                LAUNCH(ctx, merge_kernel<unsigned long long>, 1, kMergeThreads, dyn, ctx->R1, ctx->E1, SC(n_edges), SC(xctl.n_sv), ep, SC(lambda), threshold, ctx->run_start.as<unsigned>(),
                       ctx->run_end.as<unsigned>(), ctx->order, ctx->gxyz, ctx->sv_label.as<unsigned>(), ctx->ML, (unsigned)Sc, SC(mctl), scr, (unsigned long long*)sp, ctx->pos_data, resume ? 1 : 0, stop_after);
            return F3PS_OK;
        };
        // resident kernel (shared-memory tables, or the BIG variant with its tables in L2)
        auto launch_resident = [&](bool resume) -> int {
            FastArgs A;
            A.R = ctx->R1; A.E = ctx->E1; A.n_edges_ptr = SC(n_edges); A.n_sv_ptr = SC(xctl.n_sv); A.ep = ep; A.lambda_dev = SC(lambda);
            A.threshold = threshold; A.run_start = ctx->run_start.as<unsigned>(); A.run_end = ctx->run_end.as<unsigned>();
            A.pos_data = ctx->pos_data; A.sv_label = ctx->sv_label.as<unsigned>(); A.mlog = ctx->ML; A.log_cap = (unsigned)Sc;
            A.ctl = SC(mctl); A.S_cap = S_cap;
            A.E_cap = !use_big ? E_cap : E_big;
            A.big = nullptr; A.big_cursor = nullptr; A.resume = resume ? 1 : 0;
            const bool solo = !use_big && lean_ecap_for(E, kFastThreadsSolo) != 0u && can_switch;      // (a wider merge than 672 entries continues on the L2 variant)
            void (*kern)(FastArgs) = !use_big ? lean_kernel_for(ctx->merge_kernel_choice == 4, solo) : (ctx->merge_kernel_choice == 5 ? merge_fast_big_kernel<true> : merge_fast_big_kernel<false>);
            size_t launch_bytes = fast_bytes;
            if (use_big) {
                const size_t bb = (FastSmem::big_bytes(S_cap, E_big) + 255) & ~(size_t)255;
                F3PS_CUDA_OK(ctx->lean_big.ensure(bb + lean_cursor_bytes(S_cap) + LeanWideScratch::bytes));
                A.big = (char*)ctx->lean_big.p; A.big_cursor = (unsigned*)((char*)ctx->lean_big.p + bb);
                launch_bytes = FastSmem(nullptr, S_cap, E_big, A.big).bytes;
            }
            int r = lean_attr(ctx, (const void*)kern); if (r) return r;
            r = lean_pool(ctx, E, A, use_big); if (r) return r;
            A.trace = nullptr; A.trace_first = ctx->merge_trace_first;
            if (ctx->merge_kernel_choice == 4 && !use_big && !resume) {
                F3PS_CUDA_OK(ctx->merge_trace.ensure(256 * 32 * 4));
                F3PS_CUDA_OK(cudaMemsetAsync(ctx->merge_trace.p, 0, 256 * 32 * 4, ctx->stream));
                A.trace = ctx->merge_trace.as<unsigned>();
            }
            kern<<<1, solo ? kFastThreadsSolo : kFastThreads, launch_bytes, ctx->stream>>>(A);
            ctx->launches++;
            F3PS_CUDA_OK(cudaPeekAtLastError());
            return F3PS_OK;
        };
        rc = mark(ctx, 9); if (rc) return rc;
        if (!fast) { rc = launch_general(false, 0); if (rc) return rc; ctx->merge_path = 2; }
        else {
            // The resident kernel stops IN FRONT of a merge whose two adjacency lists hold more entries than it has worker threads
            // (nothing of that merge has happened; its state is that after n merges).  The general kernel then replays exactly that
            // merge from the same state and hands back; after kHandOvers such hand-overs (a scene full of hubs) it keeps the rest.
            constexpr int kHandOvers = 16;
            ctx->merge_path = single_ok ? 1 : 3;
            bool resume = false;
            for (int hand = 0;; ++hand) {
                unsigned err = kFastErrTouched;                  // (take-over: the batch grid already stopped in front of a wide merge)
                if (!(take_over && hand == 0)) {
                    rc = launch_resident(resume); if (rc) return rc;
                    rc = pull_scalars(ctx); if (rc) return rc;   // did the resident kernel finish?
                    err = ctx->h_sc->mctl.error;
                }
                if (err == 0) break;
                if (err == kFastErrTouched) {
                    F3PS_CUDA_OK(cudaMemsetAsync(SC(mctl.error), 0, 4, ctx->stream));
                    if (!use_big && can_switch) { use_big = true; resume = true; ctx->merge_path = 6; continue; }   // same state, tables in L2 from here on
                    ctx->merge_path = single_ok ? 4 : 5;
                    rc = launch_general(true, hand < kHandOvers ? 1u : 0u); if (rc) return rc;
                    if (hand >= kHandOvers) break;
                    resume = true;
                    continue;
                }
                // adjacency pool / stamp range exhausted: start over on the general kernel
                F3PS_CUDA_OK(cudaMemcpyAsync(ctx->reg_work.p, ctx->reg_init.p, region_bytes(Sc), cudaMemcpyDeviceToDevice, ctx->stream));
                F3PS_CUDA_OK(cudaMemcpyAsync(ctx->edge_work.p, ctx->edge_init.p, eb, cudaMemcpyDeviceToDevice, ctx->stream));
                F3PS_CUDA_OK(cudaMemsetAsync(SC(mctl), 0, sizeof(MergeCtl), ctx->stream));
                rc = launch_general(false, 0); if (rc) return rc;
                ctx->merge_path = 2;
                break;
            }
        }
        rc = mark(ctx, 10); if (rc) return rc;
        LAUNCH(ctx, dense_label_kernel, 1, 1024, 0, ctx->R1, SC(xctl.n_sv), ctx->run_start.as<unsigned>(), ctx->run_end.as<unsigned>(), ctx->run_out_off.as<unsigned>(),
               ctx->run_dense.as<unsigned>(), ctx->region_dense.as<unsigned>(), SC(n_out));
        if (P) LAUNCH(ctx, labeled_cloud_kernel, grid_for(P, 256), 256, 0, ctx->pos_run.as<unsigned>(), P, ctx->order, ctx->run_start.as<unsigned>(),
                      ctx->run_out_off.as<unsigned>(), ctx->run_dense.as<unsigned>(), ctx->gxyz, ctx->out_xyz.as<float>(), ctx->out_label.as<unsigned>(),
                      ctx->out_voxel.as<unsigned>(), ctx->vox_segment.as<unsigned>(), ctx->graph_from_host ? nullptr : ctx->sorted_label,
                      ctx->graph_from_host ? nullptr : ctx->owner0.as<unsigned>());
    }
    rc = mark(ctx, 8); if (rc) return rc;
    rc = pull_scalars(ctx); if (rc) return rc;
    if (ctx->h_sc->mctl.error) return ctx_fail(ctx, F3PS_ERR_CAPACITY, "merge kernel reported an internal capacity error");
    ctx->n_out = ctx->h_sc->n_out;
    ctx->progress = P_MERGED;
    return F3PS_OK;
}

// Clustering::cluster(threshold) for MANY frames with ONE launch of the resident merge kernel: CTA i of the grid replays frame i.
// A sweep that keeps frames in flight on independent streams is capped by the 32 hardware queues of a context (at most 32
// single-CTA merge kernels overlap); one grid of n CTAs runs n frames side by side.  Every handle must be on the same device,
// with its graph built (f3ps_graph); handles whose graph does not fit the resident kernel run f3ps_merge on their own.
// Results per handle are exactly those of f3ps_merge(handle, threshold).
int f3ps_merge_batch(f3ps_ctx** ctxs, int n, float threshold) {
    if (!ctxs || n < 1) return F3PS_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < n; ++i) {
        if (!ctxs[i]) return F3PS_ERR_INVALID_ARGUMENT;
        if (ctxs[i]->device != ctxs[0]->device) return ctx_fail(ctxs[i], F3PS_ERR_INVALID_ARGUMENT, "f3ps_merge_batch: handles on different devices");
        if (ctxs[i]->progress < P_EXPANDED)
            return ctx_fail(ctxs[i], F3PS_ERR_LOGIC, "Cannot call 'cluster' before setting an initial state with 'set_initialstate'");
    }
    cudaSetDevice(ctxs[0]->device);
    int rc;
    std::vector<int> batch, solo;
    std::vector<int> slots_of(n, 0);                         // per handle: edge capacity of its tables
    for (int i = 0; i < n; ++i) {
        f3ps_ctx* ctx = ctxs[i];
        if (ctx->progress < P_GRAPH) { rc = f3ps_graph(ctx); if (rc) return rc; }
        const unsigned S = ctx->S, E = ctx->E, P = ctx->n_pos;
        const unsigned ecap = lean_ecap_for(E);
        const bool ok = S > 0 && ecap && S <= 4096u && P > 0 && !ctx->force_general_merge;
        slots_of[i] = (int)ecap;
        (ok ? batch : solo).push_back(i);
    }
    {   // every frame brings its own table capacities (the layout is per CTA); the launch asks for the largest footprint
        std::vector<int> keep;
        for (int i : batch) {
            const unsigned S_cap = (ctxs[i]->S + 7u) & ~7u;
            if (FastSmem(nullptr, S_cap, (unsigned)slots_of[i]).bytes <= 227u * 1024u) keep.push_back(i); else solo.push_back(i);
        }
        batch.swap(keep);
    }
    for (size_t g0 = 0; g0 < batch.size(); g0 += kFastBatchMax) {
        const size_t g1 = std::min(batch.size(), g0 + (size_t)kFastBatchMax);
        static thread_local FastBatch B;                     // 32 KB: not on the stack
        size_t bytes = 0;
        f3ps_ctx* lead = ctxs[batch[g0]];
        for (size_t k = g0; k < g1; ++k) {
            f3ps_ctx* ctx = ctxs[batch[k]];
            rc = mark(ctx, 7); if (rc) return rc;
            const unsigned S = ctx->S, P = ctx->n_pos; const size_t Sc = std::max(1u, S), Pc = std::max(1u, P);
            F3PS_CUDA_OK(cudaMemcpyAsync(ctx->reg_work.p, ctx->reg_init.p, region_bytes(Sc), cudaMemcpyDeviceToDevice, ctx->stream));   // cluster(float) restarts (:678)
            const size_t eb = std::min(ctx->edge_init.cap, ctx->edge_work.cap);
            F3PS_CUDA_OK(cudaMemcpyAsync(ctx->edge_work.p, ctx->edge_init.p, eb, cudaMemcpyDeviceToDevice, ctx->stream));
            F3PS_CUDA_OK(cudaMemsetAsync(SC(mctl), 0, sizeof(MergeCtl), ctx->stream));
            F3PS_CUDA_OK(ctx->mlog.ensure(Sc * 20));
            char* lp = (char*)ctx->mlog.p;
            ctx->ML.a = (unsigned*)lp; ctx->ML.b = (unsigned*)(lp + Sc * 4); ctx->ML.w = (float*)(lp + Sc * 8);
            ctx->ML.edges_left = (unsigned*)(lp + Sc * 12); ctx->ML.regions_left = (unsigned*)(lp + Sc * 16);
            F3PS_CUDA_OK(ctx->run_out_off.ensure(Sc * 4)); F3PS_CUDA_OK(ctx->run_dense.ensure(Sc * 4)); F3PS_CUDA_OK(ctx->region_dense.ensure(Sc * 4));
            F3PS_CUDA_OK(ctx->out_xyz.ensure(Pc * 12)); F3PS_CUDA_OK(ctx->out_label.ensure(Pc * 4)); F3PS_CUDA_OK(ctx->out_voxel.ensure(Pc * 4));
            F3PS_CUDA_OK(ctx->vox_segment.ensure(Pc * 4));
            F3PS_CUDA_OK(cudaMemsetAsync(ctx->vox_segment.p, 0xff, Pc * 4, ctx->stream));
            F3PS_CUDA_OK(cudaMemsetAsync(SC(n_out), 0, 4, ctx->stream));
            FastArgs& A = B.a[k - g0];
            A.R = ctx->R1; A.E = ctx->E1; A.n_edges_ptr = SC(n_edges); A.n_sv_ptr = SC(xctl.n_sv); A.ep = edge_params(ctx); A.lambda_dev = SC(lambda);
            A.threshold = threshold; A.run_start = ctx->run_start.as<unsigned>(); A.run_end = ctx->run_end.as<unsigned>();
            A.pos_data = ctx->pos_data; A.sv_label = ctx->sv_label.as<unsigned>(); A.mlog = ctx->ML; A.log_cap = (unsigned)Sc;
            A.ctl = SC(mctl); A.S_cap = (S + 7u) & ~7u; A.E_cap = (unsigned)slots_of[batch[k]];
            rc = lean_pool(ctx, ctx->E, A); if (rc) return rc;
            A.trace = nullptr; A.trace_first = 0; A.big = nullptr; A.big_cursor = nullptr; A.resume = 0;
            bytes = std::max(bytes, FastSmem(nullptr, A.S_cap, A.E_cap).bytes);
            rc = mark(ctx, 9); if (rc) return rc;
            if (ctx->stream != lead->stream) {               // the lead's stream runs the grid: it waits for everybody's set-up
                F3PS_CUDA_OK(cudaEventRecord(ctx->ev_batch, ctx->stream));
                F3PS_CUDA_OK(cudaStreamWaitEvent(lead->stream, ctx->ev_batch, 0));
            }
        }
        {
            f3ps_ctx* ctx = lead;
            void (*kern)(const FastBatch) = merge_fast_batch_kernel;
            rc = lean_attr(ctx, (const void*)kern); if (rc) return rc;
            kern<<<(unsigned)(g1 - g0), kFastThreads, bytes, ctx->stream>>>(B);
            ctx->launches++;
            F3PS_CUDA_OK(cudaPeekAtLastError());
            F3PS_CUDA_OK(cudaEventRecord(ctx->ev_batch, ctx->stream));
        }
        for (size_t k = g0; k < g1; ++k) {                   // every handle continues on its own stream after the grid
            f3ps_ctx* ctx = ctxs[batch[k]];
            if (ctx->stream != lead->stream) F3PS_CUDA_OK(cudaStreamWaitEvent(ctx->stream, lead->ev_batch, 0));
            ctx->merge_path = 1;
            rc = mark(ctx, 10); if (rc) return rc;
            const unsigned P = ctx->n_pos;
            LAUNCH(ctx, dense_label_kernel, 1, 1024, 0, ctx->R1, SC(xctl.n_sv), ctx->run_start.as<unsigned>(), ctx->run_end.as<unsigned>(), ctx->run_out_off.as<unsigned>(),
                   ctx->run_dense.as<unsigned>(), ctx->region_dense.as<unsigned>(), SC(n_out));
            if (P) LAUNCH(ctx, labeled_cloud_kernel, grid_for(P, 256), 256, 0, ctx->pos_run.as<unsigned>(), P, ctx->order, ctx->run_start.as<unsigned>(),
                          ctx->run_out_off.as<unsigned>(), ctx->run_dense.as<unsigned>(), ctx->gxyz, ctx->out_xyz.as<float>(), ctx->out_label.as<unsigned>(),
                          ctx->out_voxel.as<unsigned>(), ctx->vox_segment.as<unsigned>(), ctx->graph_from_host ? nullptr : ctx->sorted_label,
                          ctx->graph_from_host ? nullptr : ctx->owner0.as<unsigned>());
            rc = mark(ctx, 8); if (rc) return rc;
        }
    }
    // graphs that do not fit the resident kernel: one by one (general kernel), next to the grids already in flight
    for (int i : solo) { rc = f3ps_merge(ctxs[i], threshold); if (rc) return rc; }
    // results only after EVERY chunk's grid is in flight (a batch larger than one parameter block must not run its chunks back to back)
    {
        for (size_t k = 0; k < batch.size(); ++k) {
            f3ps_ctx* ctx = ctxs[batch[k]];
            rc = pull_scalars(ctx); if (rc) return rc;
            if (ctx->h_sc->mctl.error == kFastErrTouched) {  // the frame's CTA stopped in front of a merge with more adjacency entries than worker
                ctx->merge_take_over = true;                 // threads: this frame alone continues from that state (general kernel for that merge, then resident again)
                rc = f3ps_merge(ctx, threshold); if (rc) return rc;
                continue;
            }
            if (ctx->h_sc->mctl.error) {                     // adjacency pool / stamp range exhausted: this frame alone, general kernel, from the start
                const bool was = ctx->force_general_merge;
                ctx->force_general_merge = true;
                rc = f3ps_merge(ctx, threshold);
                ctx->force_general_merge = was;
                if (rc) return rc;
                continue;
            }
            ctx->n_out = ctx->h_sc->n_out;
            ctx->progress = P_MERGED;
        }
    }
    return F3PS_OK;
}

// ctas_per_frame = 0: K5 is one cooperative launch over the whole GPU (lowest latency for a single frame, the default).
// ctas_per_frame > 0: K5 is an ordinary grid of that many CTAs and at most max_concurrent such kernels are in flight on the
// device (process-wide); the caller guarantees ctas_per_frame * max_concurrent <= SMs not held by long-running kernels.
int f3ps_set_expand_sharing(f3ps_ctx* ctx, int ctas_per_frame, int max_concurrent) {
    if (!ctx || ctas_per_frame < 0 || ctas_per_frame > 2 * kSMs || max_concurrent < 0 ||
        (ctas_per_frame > 0 && max_concurrent > 0 && (int64_t)ctas_per_frame * max_concurrent > kSMs))
        return F3PS_ERR_INVALID_ARGUMENT;
    ctx->expand_ctas = ctas_per_frame;
    ctx->expand_coop_cap = ctas_per_frame > 0 && max_concurrent == 0;       // still a cooperative launch, only a smaller grid
    if (ctas_per_frame > 0 && max_concurrent > 0) {
        ExpandGate& g = g_expand_gate[ctx->device & 15];
        { std::lock_guard<std::mutex> l(g.m); g.limit = max_concurrent; }
        g.cv.notify_all();
    }
    return F3PS_OK;
}

int f3ps_set_expand_kernel(f3ps_ctx* ctx, int which, int cluster_ctas) {
    if (!ctx || which < 0 || which > 2 || cluster_ctas < 0 || cluster_ctas > kExpandClusterMax) return F3PS_ERR_INVALID_ARGUMENT;
    ctx->expand_kernel_choice = which;
    ctx->expand_cluster_ctas = cluster_ctas;
    return F3PS_OK;
}

int f3ps_set_blocking_wait(f3ps_ctx* ctx, int blocking) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    ctx->blocking_wait = blocking != 0;
    return F3PS_OK;
}

int f3ps_set_merge_kernel(f3ps_ctx* ctx, int which) {
    if (!ctx || which < 0 || which > 5) return F3PS_ERR_INVALID_ARGUMENT;
    ctx->force_general_merge = which == 2;
    ctx->merge_kernel_choice = which;
    return F3PS_OK;
}

int f3ps_extract(f3ps_ctx* ctx) {
    int rc;
    if ((rc = f3ps_voxelize(ctx))) return rc;
    if ((rc = f3ps_neighbors(ctx))) return rc;
    if ((rc = f3ps_normals(ctx))) return rc;
    if ((rc = f3ps_seeds(ctx))) return rc;
    if ((rc = f3ps_expand(ctx))) return rc;
    return f3ps_graph(ctx);
}
int f3ps_run(f3ps_ctx* ctx, float threshold) {
    int rc = f3ps_extract(ctx);
    if (rc) return rc;
    return f3ps_merge(ctx, threshold);
}


// =====================================================================================================
// Slab mode (BASELINE config 5): one cloud, one slab of the x-major Morton key space per
// GPU.  The per-rank pieces live here; the host driver (f3ps/slab.py) issues the exchanges between them on this
// handle's stream: all-reduce of the bounding box and the key histogram, all-to-all of the points, all-gather of the
// voxel slices, the normal slices and -- per expansion sweep -- the steal-table slices.
extern "C++" {
namespace {
template <typename KeyT>
int slab_keys_typed(f3ps_ctx* ctx, PointLoader pl, int shift, unsigned bins, unsigned* d_hist) {
    const int64_t N = ctx->n_points;
    F3PS_CUDA_OK(ctx->keys_a.ensure(N * sizeof(KeyT))); F3PS_CUDA_OK(ctx->keys_b.ensure(N * sizeof(KeyT)));
    F3PS_CUDA_OK(ctx->vals_a.ensure(N * 4)); F3PS_CUDA_OK(ctx->vals_b.ensure(N * 4));
    KeygenOp<KeyT> kop{pl, ctx->vp.use_transform, SC(fp), ctx->keys_a.as<KeyT>(), ctx->vals_a.as<unsigned>()};
    int rc = run_compact(ctx, kop, nullptr, N, SC(n_valid)); if (rc) return rc;
    LAUNCH(ctx, slab_key_hist_kernel<KeyT>, grid_for(N, 256 * 4, kSMs * 8), 256, 0, ctx->keys_a.as<KeyT>(), SC(n_valid), shift, bins, d_hist);
    return F3PS_OK;
}
template <typename KeyT>
int slab_route_typed(f3ps_ctx* ctx, PointLoader pl, const SlabSplitters& sp, int world, float4* d_send) {
    const int64_t N = ctx->n_points;
    F3PS_CUDA_OK(ctx->slab_dest.ensure(N * 4)); F3PS_CUDA_OK(ctx->starts.ensure((N + 1) * 4));
    LAUNCH(ctx, slab_dest_kernel<KeyT>, grid_for(N, 256 * 4, kSMs * 8), 256, 0, ctx->keys_a.as<KeyT>(), SC(n_valid), sp,
           ctx->slab_dest.as<unsigned>(), ctx->slab_tot.as<unsigned>());
    unsigned* sk; unsigned* sv;        // stable: input order survives inside a destination
    int rc = sort_pairs<unsigned>(ctx, ctx->slab_dest.as<unsigned>(), ctx->vals_a.as<unsigned>(), ctx->starts.as<unsigned>(), ctx->vals_b.as<unsigned>(),
                                  ctx->slab_dest.as<unsigned>(), ctx->vals_a.as<unsigned>(), SC(n_valid), N, bits_for((unsigned)std::max(1, world - 1)), &sk, &sv);
    if (rc) return rc;
    LAUNCH(ctx, slab_pack_kernel, grid_for(N, 256 * 2, kSMs * 8), 256, 0, pl, sv, SC(n_valid), d_send);
    return F3PS_OK;
}
} // namespace
} // extern "C++"

int f3ps_slab_reset(f3ps_ctx* ctx) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    ctx->slab_frame = false; ctx->slab_range = false; ctx->slab_expanding = false;
    ctx->progress = std::min(ctx->progress, (int)P_INPUT);
    return F3PS_OK;
}

// bounding box of this rank's points (transformed space): d_box8 = ord_min[3], ord_max[3], any_finite, 0 in the
// order-preserving unsigned encoding (all-reduce MIN / MAX / MAX as unsigned)
int f3ps_slab_bbox(f3ps_ctx* ctx, uint32_t* d_box8) {
    if (!ctx || !d_box8) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_INPUT, "f3ps_slab_bbox"); if (rc) return rc;
    cudaSetDevice(ctx->device);
    DevScalars init; memset(&init, 0, sizeof init);
    for (int a = 0; a < 3; ++a) { init.fp.ord_min[a] = 0xffffffffu; init.fp.ord_max[a] = 0u; }
    *ctx->h_sc = init;
    F3PS_CUDA_OK(cudaMemcpyAsync(ctx->d_sc, ctx->h_sc, sizeof(DevScalars), cudaMemcpyHostToDevice, ctx->stream));
    PointLoader pl{ctx->d_points, ctx->stride, ctx->vp.fold_negative_z};
    if (ctx->n_points > 0)
        LAUNCH(ctx, bbox_kernel, grid_for(ctx->n_points, 256 * 4, kSMs * 8), 256, 0, pl, ctx->n_points, ctx->vp.use_transform, SC(fp));
    F3PS_CUDA_OK(cudaMemcpyAsync(d_box8, SC(fp.ord_min), 24, cudaMemcpyDeviceToDevice, ctx->stream));
    F3PS_CUDA_OK(cudaMemcpyAsync(d_box8 + 6, SC(fp.any_finite), 4, cudaMemcpyDeviceToDevice, ctx->stream));
    F3PS_CUDA_OK(cudaMemsetAsync(d_box8 + 7, 0, 4, ctx->stream));
    ctx->slab_frame = false;
    return F3PS_OK;
}

// the reduced box of the whole cloud -> OctreePointCloud::defineBoundingBox / getKeyBitSize, identical on every rank
int f3ps_slab_set_frame(f3ps_ctx* ctx, const uint32_t* d_box8) {
    if (!ctx || !d_box8) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_INPUT, "f3ps_slab_set_frame"); if (rc) return rc;
    cudaSetDevice(ctx->device);
    F3PS_CUDA_OK(cudaMemcpyAsync(SC(fp.ord_min), d_box8, 24, cudaMemcpyDeviceToDevice, ctx->stream));
    F3PS_CUDA_OK(cudaMemcpyAsync(SC(fp.any_finite), d_box8 + 6, 4, cudaMemcpyDeviceToDevice, ctx->stream));
    LAUNCH(ctx, frame_setup_kernel, 1, 1, 0, SC(fp), ctx->vp.voxel_res);
    rc = pull_scalars(ctx); if (rc) return rc;
    ctx->slab_fp = ctx->h_sc->fp; ctx->slab_frame = true;
    ctx->depth = ctx->slab_fp.depth;
    if (ctx->depth > 21) return ctx_fail(ctx, F3PS_ERR_CAPACITY, "adjacency octree depth > 21");
    ctx->key64 = 3 * ctx->depth > 32;
    return F3PS_OK;
}

// Morton keys of this rank's valid points + histogram of their top `top_bits` bits (d_hist: 1 << top_bits counters, device);
// *used_bits / *shift describe the binning: bin = key >> shift, min(top_bits, 3 * depth) bits
int f3ps_slab_keys(f3ps_ctx* ctx, int top_bits, uint32_t* d_hist, int* used_bits, int* shift_out) {
    if (!ctx || !d_hist || top_bits < 1 || top_bits > kSlabHistBitsMax) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_INPUT, "f3ps_slab_keys"); if (rc) return rc;
    if (!ctx->slab_frame) return ctx_fail(ctx, F3PS_ERR_LOGIC, "f3ps_slab_keys: f3ps_slab_set_frame has not run");
    cudaSetDevice(ctx->device);
    const int kb = 3 * ctx->depth, tb = std::min(top_bits, std::max(1, kb)), shift = std::max(0, kb - tb);
    if (used_bits) *used_bits = tb;
    if (shift_out) *shift_out = shift;
    F3PS_CUDA_OK(cudaMemsetAsync(d_hist, 0, (size_t)4 << top_bits, ctx->stream));
    PointLoader pl{ctx->d_points, ctx->stride, ctx->vp.fold_negative_z};
    if (ctx->n_points == 0) { F3PS_CUDA_OK(cudaMemsetAsync(SC(n_valid), 0, 4, ctx->stream)); return F3PS_OK; }
    return ctx->key64 ? slab_keys_typed<uint64_t>(ctx, pl, shift, 1u << tb, d_hist) : slab_keys_typed<uint32_t>(ctx, pl, shift, 1u << tb, d_hist);
}

// splitters[world - 1]: ascending Morton keys, rank r owns [splitters[r-1], splitters[r]).  Packs this rank's valid points as
// 16-byte records {x, y, z, rgba} into d_send (capacity: n points), grouped by destination rank, input order kept inside
// a group; counts[world] = records per destination.
int f3ps_slab_route(f3ps_ctx* ctx, int world, const uint64_t* splitters, void* d_send, int64_t* counts) {
    if (!ctx || world < 1 || world > kSlabMaxWorld || (world > 1 && !splitters) || !counts) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_INPUT, "f3ps_slab_route"); if (rc) return rc;
    if (!ctx->slab_frame) return ctx_fail(ctx, F3PS_ERR_LOGIC, "f3ps_slab_route: f3ps_slab_set_frame / f3ps_slab_keys have not run");
    cudaSetDevice(ctx->device);
    for (int r = 0; r < world; ++r) counts[r] = 0;
    if (ctx->n_points == 0) return F3PS_OK;
    SlabSplitters sp; memset(&sp, 0, sizeof sp); sp.n = world - 1;
    for (int r = 0; r + 1 < world; ++r) { sp.key[r] = splitters[r]; if (r && splitters[r] < splitters[r - 1]) return ctx_fail(ctx, F3PS_ERR_INVALID_ARGUMENT, "splitters must ascend"); }
    F3PS_CUDA_OK(ctx->slab_tot.ensure(kSlabMaxWorld * 4));
    F3PS_CUDA_OK(cudaMemsetAsync(ctx->slab_tot.p, 0, kSlabMaxWorld * 4, ctx->stream));
    PointLoader pl{ctx->d_points, ctx->stride, ctx->vp.fold_negative_z};
    rc = ctx->key64 ? slab_route_typed<uint64_t>(ctx, pl, sp, world, (float4*)d_send) : slab_route_typed<uint32_t>(ctx, pl, sp, world, (float4*)d_send);
    if (rc) return rc;
    unsigned tot[kSlabMaxWorld];
    F3PS_CUDA_OK(cudaMemcpyAsync(tot, ctx->slab_tot.p, sizeof tot, cudaMemcpyDeviceToHost, ctx->stream));
    F3PS_CUDA_OK(wait_stream(ctx));
    for (int r = 0; r < world; ++r) counts[r] = tot[r];
    return F3PS_OK;
}

// device arrays the driver exchanges in place (pointer, elements, bytes per element)
int f3ps_slab_array(f3ps_ctx* ctx, int which, void** ptr, int64_t* n, int* elem_bytes) {
    if (!ctx || !ptr || !n || !elem_bytes) return F3PS_ERR_INVALID_ARGUMENT;
    const int64_t V = ctx->V, L = (int64_t)ctx->S0 + 2;
    const ExpandArgs& A = ctx->slab_A;
    switch (which) {
    case F3PS_SLAB_VOX_XYZ: *ptr = ctx->vox_xyz.p; *n = V; *elem_bytes = 16; break;
    case F3PS_SLAB_VOX_RGB: *ptr = ctx->vox_rgb.p; *n = V; *elem_bytes = 16; break;
    case F3PS_SLAB_VOX_KEY: *ptr = ctx->vox_key.p; *n = V; *elem_bytes = 8; break;
    case F3PS_SLAB_VOX_NORMAL: *ptr = ctx->vox_normal.p; *n = V; *elem_bytes = 16; break;
    case F3PS_SLAB_VOX_CURV: *ptr = ctx->vox_curv.p; *n = V; *elem_bytes = 4; break;
    case F3PS_SLAB_STEAL: if (!ctx->slab_expanding) return ctx_fail(ctx, F3PS_ERR_LOGIC, "no expansion in progress");
        *ptr = A.st[ctx->slab_k & 1]; *n = V; *elem_bytes = 4; break;
    case F3PS_SLAB_OWNER_NEXT: if (!ctx->slab_expanding) return ctx_fail(ctx, F3PS_ERR_LOGIC, "no expansion in progress");
        *ptr = A.owner[ctx->slab_cur ^ 1]; *n = V; *elem_bytes = 4; break;
    case F3PS_SLAB_DIST: if (!ctx->slab_expanding) return ctx_fail(ctx, F3PS_ERR_LOGIC, "no expansion in progress");
        *ptr = A.dist[ctx->slab_cur]; *n = V; *elem_bytes = 4; break;
    case F3PS_SLAB_COUNT: if (!ctx->slab_expanding) return ctx_fail(ctx, F3PS_ERR_LOGIC, "no expansion in progress");
        *ptr = A.count[(ctx->slab_k + 1) & 1]; *n = L; *elem_bytes = 4; break;     // tallies of the last sweep
    default: return ctx_fail(ctx, F3PS_ERR_INVALID_ARGUMENT, "unknown slab array");
    }
    return F3PS_OK;
}

// the voxel table of the WHOLE cloud (every rank's slice, leaf order) and the slice [own_begin, own_end) this rank computes
int f3ps_slab_set_voxels(f3ps_ctx* ctx, const void* d_xyz, const void* d_rgb, const void* d_key, int64_t n_voxels, int64_t own_begin, int64_t own_end) {
    if (!ctx || n_voxels < 0 || own_begin < 0 || own_end < own_begin || own_end > n_voxels) return F3PS_ERR_INVALID_ARGUMENT;
    if (n_voxels >= (1ll << 30)) return ctx_fail(ctx, F3PS_ERR_CAPACITY, "more than 2^30 voxels");
    if (!ctx->slab_frame) return ctx_fail(ctx, F3PS_ERR_LOGIC, "f3ps_slab_set_voxels: f3ps_slab_set_frame has not run");
    cudaSetDevice(ctx->device);
    const size_t V = (size_t)n_voxels, Vc = std::max<size_t>(1, V);
    if (d_xyz != ctx->vox_xyz.p) {
        F3PS_CUDA_OK(wait_stream(ctx));
        F3PS_CUDA_OK(ctx->vox_xyz.ensure(Vc * 16)); F3PS_CUDA_OK(ctx->vox_rgb.ensure(Vc * 16)); F3PS_CUDA_OK(ctx->vox_key.ensure(Vc * 8));
        if (V) {
            F3PS_CUDA_OK(cudaMemcpyAsync(ctx->vox_xyz.p, d_xyz, V * 16, cudaMemcpyDeviceToDevice, ctx->stream));
            F3PS_CUDA_OK(cudaMemcpyAsync(ctx->vox_rgb.p, d_rgb, V * 16, cudaMemcpyDeviceToDevice, ctx->stream));
            F3PS_CUDA_OK(cudaMemcpyAsync(ctx->vox_key.p, d_key, V * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        }
    }
    ctx->V = (unsigned)V;
    const unsigned v32 = (unsigned)V;
    F3PS_CUDA_OK(cudaMemcpyAsync(SC(n_voxels), &v32, 4, cudaMemcpyHostToDevice, ctx->stream));
    F3PS_CUDA_OK(wait_stream(ctx));
    ctx->own_begin = (unsigned)own_begin; ctx->own_end = (unsigned)own_end; ctx->slab_range = true;
    ctx->progress = P_VOXELS;
    return mark(ctx, 1);
}

// ---- K5 in slab mode: createSupervoxelHelpers, then per round {sweeps over the owned slice, round end}, then the tail ----
int f3ps_slab_expand_begin(f3ps_ctx* ctx) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_SEEDS, "f3ps_slab_expand_begin"); if (rc) return rc;
    if (!ctx->slab_range) return ctx_fail(ctx, F3PS_ERR_LOGIC, "f3ps_slab_expand_begin: not in slab mode");
    cudaSetDevice(ctx->device);
    rc = expand_prepare(ctx, ctx->slab_A); if (rc) return rc;
    const ExpandArgs& A = ctx->slab_A;
    ctx->slab_cur = 0; ctx->slab_k = 0; ctx->slab_sweeps = 0; ctx->slab_round = 0; ctx->slab_expanding = true;
    if (ctx->V) {
        LAUNCH(ctx, slab_expand_init_kernel, grid_for(ctx->V, 256), 256, 0, A);
        if (ctx->S0) {
            LAUNCH(ctx, slab_expand_seed_kernel, grid_for(ctx->S0, 256), 256, 0, A);
            LAUNCH(ctx, slab_expand_phantom_kernel, grid_for(ctx->S0, 256), 256, 0, A);
        }
    }
    return F3PS_OK;
}

// one sweep over [own_begin, own_end); *d_changed (device) = 1 if a steal-table entry of the slice moved.  Afterwards
// F3PS_SLAB_STEAL is the table this sweep wrote: exchange its slices, all-reduce the flag (MAX), repeat until it is 0.
int f3ps_slab_expand_sweep(f3ps_ctx* ctx, uint32_t* d_changed) {
    if (!ctx || !d_changed) return F3PS_ERR_INVALID_ARGUMENT;
    if (!ctx->slab_expanding) return ctx_fail(ctx, F3PS_ERR_LOGIC, "f3ps_slab_expand_sweep: f3ps_slab_expand_begin has not run");
    cudaSetDevice(ctx->device);
    const ExpandArgs& A = ctx->slab_A;
    F3PS_CUDA_OK(cudaMemsetAsync(d_changed, 0, 4, ctx->stream));
    F3PS_CUDA_OK(cudaMemsetAsync(A.count[ctx->slab_k & 1], 0, ((size_t)ctx->S0 + 2) * 4, ctx->stream));
    const unsigned nb = ctx->own_end - ctx->own_begin;
    if (nb)
        LAUNCH(ctx, slab_expand_sweep_kernel, grid_for(nb, kExpandThreads, kSMs * 2), kExpandThreads, 0, A, ctx->own_begin, ctx->own_end, ctx->slab_cur,
               ctx->slab_k, d_changed);
    ctx->slab_k++; ctx->slab_sweeps++;
    return F3PS_OK;
}

// after the converged sweep of a round, with F3PS_SLAB_OWNER_NEXT slices exchanged and F3PS_SLAB_COUNT all-reduced (SUM):
// helper lists, SupervoxelHelper::updateCentroid (every rank, whole cloud).  With zero expansion rounds call it once, without sweeps.
int f3ps_slab_expand_round_end(f3ps_ctx* ctx) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    if (!ctx->slab_expanding) return ctx_fail(ctx, F3PS_ERR_LOGIC, "f3ps_slab_expand_round_end: f3ps_slab_expand_begin has not run");
    cudaSetDevice(ctx->device);
    const ExpandArgs& A = ctx->slab_A;
    const unsigned V = ctx->V, S0 = ctx->S0;
    unsigned* cnt_final;
    if (ctx->rounds > 0) { cnt_final = A.count[(ctx->slab_k + 1) & 1]; ctx->slab_cur ^= 1; }
    else {
        cnt_final = A.count[0];
        if (V) LAUNCH(ctx, slab_expand_count0_kernel, grid_for(V, 256), 256, 0, A, ctx->slab_cur, cnt_final);
    }
    F3PS_CUDA_OK(cudaMemsetAsync(SC(xctl.cursor), 0, 4, ctx->stream));
    if (V && S0) {
        LAUNCH(ctx, slab_expand_alloc_kernel, grid_for(S0, 256), 256, 0, A, cnt_final);
        LAUNCH(ctx, slab_expand_fill_kernel, grid_for(V, 256), 256, 0, A, ctx->slab_cur, ctx->slab_k);
        LAUNCH(ctx, slab_expand_fold_kernel, grid_for((int64_t)S0 * 32, kExpandThreads, kSMs * 2), kExpandThreads, 0, A, cnt_final);
    }
    ctx->slab_round++;
    return F3PS_OK;
}

// after the last round (F3PS_SLAB_DIST slices exchanged): clean labels, surviving helpers (makeSupervoxels)
int f3ps_slab_expand_end(f3ps_ctx* ctx) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    if (!ctx->slab_expanding) return ctx_fail(ctx, F3PS_ERR_LOGIC, "f3ps_slab_expand_end: f3ps_slab_expand_begin has not run");
    cudaSetDevice(ctx->device);
    const ExpandArgs& A = ctx->slab_A;
    const unsigned* cnt_final = ctx->rounds > 0 ? A.count[(ctx->slab_k + 1) & 1] : A.count[0];
    if (ctx->V) LAUNCH(ctx, slab_expand_tail_kernel, 1 + grid_for(ctx->V, 1024, kSMs * 2), 1024, 0, A, ctx->slab_cur, cnt_final);
    F3PS_CUDA_OK(cudaMemcpyAsync(SC(xctl.sweeps_total), &ctx->slab_sweeps, 4, cudaMemcpyHostToDevice, ctx->stream));
    ctx->slab_expanding = false;
    return expand_finish(ctx);
}


int f3ps_merge_profile(f3ps_ctx* ctx, uint64_t cycles[32]) {
    if (!ctx || !cycles) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_MERGED, "f3ps_merge_profile"); if (rc) return rc;
    for (int i = 0; i < 32; ++i) cycles[i] = ctx->h_sc->mctl.phase_cycles[i];
    return F3PS_OK;
}

int f3ps_merge_trace(f3ps_ctx* ctx, int64_t first_merge, uint32_t* out, int64_t capacity_words) {
    if (!ctx || first_merge < 0) return F3PS_ERR_INVALID_ARGUMENT;
    ctx->merge_trace_first = (unsigned)first_merge;          // takes effect at the next f3ps_merge with kernel choice 4
    if (!out) return F3PS_OK;
    if (capacity_words < 256 * 32 || !ctx->merge_trace.p) return ctx_fail(ctx, F3PS_ERR_INVALID_ARGUMENT, "f3ps_merge_trace: no trace recorded or buffer below 8192 words");
    cudaSetDevice(ctx->device);
    F3PS_CUDA_OK(cudaMemcpyAsync(out, ctx->merge_trace.p, 256 * 32 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    F3PS_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return F3PS_OK;
}

int f3ps_expand_profile(f3ps_ctx* ctx, uint64_t ns[8]) {
    if (!ctx || !ns) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_EXPANDED, "f3ps_expand_profile"); if (rc) return rc;
    if ((rc = pull_scalars(ctx))) return rc;
    for (int i = 0; i < 8; ++i) ns[i] = ctx->h_sc->xctl.t_phase[i];
    return F3PS_OK;
}

int f3ps_stage_ms(f3ps_ctx* ctx, int stage, float* ms) {
    if (!ctx || !ms || stage < 0 || stage > 8) return F3PS_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    F3PS_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    int i0, i1;
    if (stage == F3PS_STAGE_TOTAL) { i0 = 0; i1 = 8; }
    else if (stage == F3PS_STAGE_MERGE_KERNEL) { i0 = 9; i1 = 10; }
    else if (stage == F3PS_STAGE_MERGE) { i0 = 7; i1 = 8; }
    else { i0 = stage; i1 = stage + 1; }
    if (!ctx->ev_valid[i0] || !ctx->ev_valid[i1]) return ctx_fail(ctx, F3PS_ERR_LOGIC, "stage has not run");
    F3PS_CUDA_OK(cudaEventElapsedTime(ms, ctx->ev[i0], ctx->ev[i1]));
    return F3PS_OK;
}

} // extern "C"

// =====================================================================================================
// results
namespace {
int d2h(f3ps_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!bytes || !dst) return F3PS_OK;
    F3PS_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return F3PS_OK;
}
int fin(f3ps_ctx* ctx) { F3PS_CUDA_OK(wait_stream(ctx)); return F3PS_OK; }
int cap_check(f3ps_ctx* ctx, int64_t need_n, int64_t capacity) {
    if (capacity < need_n) return ctx_fail(ctx, F3PS_ERR_INVALID_ARGUMENT, "output capacity too small");
    return F3PS_OK;
}
}

extern "C" {

int f3ps_get_counts(f3ps_ctx* ctx, f3ps_counts* out) {
    if (!ctx || !out) return F3PS_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    int rc = pull_scalars(ctx); if (rc) return rc;
    const DevScalars& h = *ctx->h_sc;
    memset(out, 0, sizeof *out);
    out->n_points = ctx->n_points; out->n_valid = ctx->n_valid; out->n_voxels = ctx->V; out->depth = ctx->depth;
    out->seed_depth = h.sb.depth; out->n_seed_cells = (int)ctx->n_cells; out->n_seeds = (int)ctx->S0;
    out->n_supervoxels = (int)ctx->S; out->n_edges = (int)ctx->E;
    out->n_merges = (int)h.mctl.n_merges; out->n_segments = (int)h.mctl.regions_alive; out->n_edges_left = (int)h.mctl.edges_alive;
    out->rounds = ctx->rounds; out->sweeps = (int)h.xctl.sweeps_total; out->n_labeled = (int)ctx->n_out;
    out->lambda = h.lambda;
    out->max_touched = (int)h.mctl.max_touched; out->fold_steps = (int64_t)h.mctl.fold_steps;
    out->nan_weights = (int)(h.mctl.nan_weights + h.nan_weights_init);
    out->merge_path = ctx->merge_path;
    if (ctx->progress < P_MERGED) { out->n_segments = (int)ctx->S; out->n_edges_left = (int)ctx->E; }
    return F3PS_OK;
}

int f3ps_get_voxel_keys(f3ps_ctx* ctx, uint32_t* keys, int64_t capacity) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_VOXELS, "f3ps_get_voxel_keys"); if (rc) return rc;
    if ((rc = cap_check(ctx, ctx->V, capacity))) return rc;
    std::vector<uint64_t> m(ctx->V);
    if ((rc = d2h(ctx, m.data(), ctx->vox_key.p, (size_t)ctx->V * 8))) return rc;
    if ((rc = fin(ctx))) return rc;
    for (unsigned v = 0; v < ctx->V; ++v) morton_decode(m[v], keys[3 * v], keys[3 * v + 1], keys[3 * v + 2]);
    return F3PS_OK;
}

int f3ps_get_voxel_centroids(f3ps_ctx* ctx, float* xyz, float* rgb, uint32_t* rgba, int32_t* count, int64_t capacity) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_VOXELS, "f3ps_get_voxel_centroids"); if (rc) return rc;
    if ((rc = cap_check(ctx, ctx->V, capacity))) return rc;
    const unsigned V = ctx->V;
    std::vector<float4> a(V), b(V);
    if ((rc = d2h(ctx, a.data(), ctx->vox_xyz.p, (size_t)V * 16))) return rc;
    if ((rc = d2h(ctx, b.data(), ctx->vox_rgb.p, (size_t)V * 16))) return rc;
    if ((rc = fin(ctx))) return rc;
    for (unsigned v = 0; v < V; ++v) {
        if (xyz) { xyz[3 * v] = a[v].x; xyz[3 * v + 1] = a[v].y; xyz[3 * v + 2] = a[v].z; }
        if (rgb) { rgb[3 * v] = b[v].x; rgb[3 * v + 1] = b[v].y; rgb[3 * v + 2] = b[v].z; }
        if (rgba) memcpy(&rgba[v], &a[v].w, 4);
        if (count) count[v] = (int)b[v].w;
    }
    return F3PS_OK;
}

int f3ps_get_point_voxel(f3ps_ctx* ctx, int32_t* pv, int64_t capacity) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_VOXELS, "f3ps_get_point_voxel"); if (rc) return rc;
    if ((rc = cap_check(ctx, ctx->n_points, capacity))) return rc;
    if ((rc = d2h(ctx, pv, ctx->point_voxel.p, (size_t)ctx->n_points * 4))) return rc;
    return fin(ctx);
}

int f3ps_get_voxel_neighbors(f3ps_ctx* ctx, int32_t* nbr, int32_t* nbr_count, int64_t capacity) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_NEIGHBORS, "f3ps_get_voxel_neighbors"); if (rc) return rc;
    if ((rc = cap_check(ctx, ctx->V, capacity))) return rc;
    const unsigned V = ctx->V;
    std::vector<int> rows((size_t)V * kNbrStride);
    if ((rc = d2h(ctx, rows.data(), ctx->nbr_row.p, rows.size() * 4))) return rc;
    if ((rc = fin(ctx))) return rc;
    for (unsigned v = 0; v < V; ++v) {
        if (nbr) for (int r = 0; r < 27; ++r) nbr[(size_t)v * 27 + r] = rows[(size_t)v * kNbrStride + r];
        if (nbr_count) nbr_count[v] = rows[(size_t)v * kNbrStride + 27];
    }
    return F3PS_OK;
}

int f3ps_get_voxel_normals(f3ps_ctx* ctx, float* normal4, float* curvature, int64_t capacity) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_NORMALS, "f3ps_get_voxel_normals"); if (rc) return rc;
    if ((rc = cap_check(ctx, ctx->V, capacity))) return rc;
    if ((rc = d2h(ctx, normal4, ctx->vox_normal.p, (size_t)ctx->V * 16))) return rc;
    if ((rc = d2h(ctx, curvature, ctx->vox_curv.p, (size_t)ctx->V * 4))) return rc;
    return fin(ctx);
}

int f3ps_get_seeds(f3ps_ctx* ctx, int32_t* seed_voxel, int64_t capacity) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_SEEDS, "f3ps_get_seeds"); if (rc) return rc;
    if ((rc = cap_check(ctx, ctx->S0, capacity))) return rc;
    if ((rc = d2h(ctx, seed_voxel, ctx->seeds.p, (size_t)ctx->S0 * 4))) return rc;
    return fin(ctx);
}

int f3ps_get_voxel_labels(f3ps_ctx* ctx, uint32_t* label, float* distance, int64_t capacity) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_EXPANDED, "f3ps_get_voxel_labels"); if (rc) return rc;
    if (ctx->graph_from_host) return ctx_fail(ctx, F3PS_ERR_LOGIC, "no voxel labels: graph was supplied by the caller");
    if ((rc = cap_check(ctx, ctx->V, capacity))) return rc;
    if ((rc = d2h(ctx, label, ctx->owner0.p, (size_t)ctx->V * 4))) return rc;
    if ((rc = d2h(ctx, distance, ctx->dist0.p, (size_t)ctx->V * 4))) return rc;
    return fin(ctx);
}

int f3ps_get_supervoxels(f3ps_ctx* ctx, uint32_t* label, float* centroid, float* mean_rgb, float* normal4, int32_t* n_voxels, int64_t capacity) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_EXPANDED, "f3ps_get_supervoxels"); if (rc) return rc;
    if (ctx->graph_from_host) return ctx_fail(ctx, F3PS_ERR_LOGIC, "supervoxels were supplied by the caller");
    if ((rc = cap_check(ctx, ctx->S, capacity))) return rc;
    const unsigned S = ctx->S, Sc = ctx->S0 + 2;
    std::vector<unsigned> lab(S);
    std::vector<float4> x(Sc), c(Sc), n(Sc);
    if ((rc = d2h(ctx, lab.data(), ctx->sv_label.p, (size_t)S * 4))) return rc;
    if ((rc = d2h(ctx, x.data(), ctx->cen_xyz.p, (size_t)Sc * 16))) return rc;
    if ((rc = d2h(ctx, c.data(), ctx->cen_rgb.p, (size_t)Sc * 16))) return rc;
    if ((rc = d2h(ctx, n.data(), ctx->cen_nrm.p, (size_t)Sc * 16))) return rc;
    if ((rc = fin(ctx))) return rc;
    for (unsigned s = 0; s < S; ++s) {
        const unsigned l = lab[s];
        if (label) label[s] = l;
        if (centroid) { centroid[3 * s] = x[l].x; centroid[3 * s + 1] = x[l].y; centroid[3 * s + 2] = x[l].z; }
        if (mean_rgb) { mean_rgb[3 * s] = c[l].x; mean_rgb[3 * s + 1] = c[l].y; mean_rgb[3 * s + 2] = c[l].z; }
        if (normal4) { normal4[4 * s] = n[l].x; normal4[4 * s + 1] = n[l].y; normal4[4 * s + 2] = n[l].z; normal4[4 * s + 3] = n[l].w; }
        if (n_voxels) n_voxels[s] = (int)x[l].w;
    }
    return F3PS_OK;
}

int f3ps_get_supervoxel_voxels(f3ps_ctx* ctx, int32_t* voxel_index, int64_t* offsets, int64_t cap_voxels, int64_t cap_sv) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_GRAPH, "f3ps_get_supervoxel_voxels"); if (rc) return rc;
    const unsigned S = ctx->S, P = ctx->n_pos;
    if ((rc = cap_check(ctx, S, cap_sv))) return rc;
    std::vector<unsigned> rs(S), re(S), ord(P);
    if ((rc = d2h(ctx, rs.data(), ctx->run_start.p, (size_t)S * 4))) return rc;
    if ((rc = d2h(ctx, re.data(), ctx->run_end.p, (size_t)S * 4))) return rc;
    if ((rc = d2h(ctx, ord.data(), ctx->order, (size_t)P * 4))) return rc;
    if ((rc = fin(ctx))) return rc;
    int64_t tot = 0;
    for (unsigned s = 0; s < S; ++s) tot += re[s] - rs[s];
    if ((rc = cap_check(ctx, tot, cap_voxels))) return rc;
    int64_t o = 0;
    for (unsigned s = 0; s < S; ++s) {
        if (offsets) offsets[s] = o;
        for (unsigned i = rs[s]; i < re[s]; ++i) { if (voxel_index) voxel_index[o] = (int)ord[i]; ++o; }
    }
    if (offsets) offsets[S] = o;
    return F3PS_OK;
}

static int fetch_edges(f3ps_ctx* ctx, const EdgeArrays& EA, unsigned n, std::vector<unsigned>& a, std::vector<unsigned>& b, std::vector<float>& dc,
                       std::vector<float>& dg, std::vector<float>& w, std::vector<long long>& st, std::vector<unsigned>& lab) {
    a.resize(n); b.resize(n); dc.resize(n); dg.resize(n); w.resize(n); st.resize(n); lab.resize(ctx->S);
    int rc;
    if ((rc = d2h(ctx, a.data(), EA.a, (size_t)n * 4))) return rc;
    if ((rc = d2h(ctx, b.data(), EA.b, (size_t)n * 4))) return rc;
    if ((rc = d2h(ctx, dc.data(), EA.dc, (size_t)n * 4))) return rc;
    if ((rc = d2h(ctx, dg.data(), EA.dg, (size_t)n * 4))) return rc;
    if ((rc = d2h(ctx, w.data(), EA.w, (size_t)n * 4))) return rc;
    if ((rc = d2h(ctx, st.data(), EA.stamp, (size_t)n * 8))) return rc;
    if ((rc = d2h(ctx, lab.data(), ctx->sv_label.p, (size_t)ctx->S * 4))) return rc;
    return fin(ctx);
}

int f3ps_get_edges(f3ps_ctx* ctx, uint32_t* ab, float* delta_c, float* delta_g, float* weight, int64_t capacity) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_GRAPH, "f3ps_get_edges"); if (rc) return rc;
    if ((rc = cap_check(ctx, ctx->E, capacity))) return rc;
    std::vector<unsigned> a, b, lab; std::vector<float> dc, dg, w; std::vector<long long> st;
    if ((rc = fetch_edges(ctx, ctx->E0, ctx->E, a, b, dc, dg, w, st, lab))) return rc;
    for (unsigned e = 0; e < ctx->E; ++e) {
        if (ab) { ab[2 * e] = lab[a[e]]; ab[2 * e + 1] = lab[b[e]]; }
        if (delta_c) delta_c[e] = dc[e];
        if (delta_g) delta_g[e] = dg[e];
        if (weight) weight[e] = w[e];
    }
    return F3PS_OK;
}

int f3ps_get_adjacency(f3ps_ctx* ctx, uint32_t* pairs, int64_t capacity) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_GRAPH, "f3ps_get_adjacency"); if (rc) return rc;
    if ((rc = cap_check(ctx, 2 * (int64_t)ctx->E, capacity))) return rc;
    std::vector<unsigned> a, b, lab; std::vector<float> dc, dg, w; std::vector<long long> st;
    if ((rc = fetch_edges(ctx, ctx->E0, ctx->E, a, b, dc, dg, w, st, lab))) return rc;
    std::vector<std::pair<unsigned, unsigned>> p;
    for (unsigned e = 0; e < ctx->E; ++e) { p.push_back({lab[a[e]], lab[b[e]]}); p.push_back({lab[b[e]], lab[a[e]]}); }
    std::sort(p.begin(), p.end());
    for (size_t i = 0; i < p.size(); ++i) { pairs[2 * i] = p[i].first; pairs[2 * i + 1] = p[i].second; }
    return F3PS_OK;
}

int f3ps_get_cdf(f3ps_ctx* ctx, float* cdf_c, float* cdf_g, int64_t capacity) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_GRAPH, "f3ps_get_cdf"); if (rc) return rc;
    if (ctx->mp.merge_mode != F3PS_EQUALIZATION) return ctx_fail(ctx, F3PS_ERR_LOGIC, "no CDF unless the merging criterion is EQUALIZATION");
    if ((rc = cap_check(ctx, ctx->mp.bins, capacity))) return rc;
    if ((rc = d2h(ctx, cdf_c, ctx->cdf_c.p, (size_t)ctx->mp.bins * 4))) return rc;
    if ((rc = d2h(ctx, cdf_g, ctx->cdf_g.p, (size_t)ctx->mp.bins * 4))) return rc;
    return fin(ctx);
}

int f3ps_get_merge_log(f3ps_ctx* ctx, uint32_t* ab, float* weight, uint32_t* left, int64_t capacity) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_MERGED, "f3ps_get_merge_log"); if (rc) return rc;
    const unsigned M = ctx->h_sc->mctl.n_merges;
    if ((rc = cap_check(ctx, M, capacity))) return rc;
    std::vector<unsigned> a(M), b(M), el(M), rl(M);
    if ((rc = d2h(ctx, a.data(), ctx->ML.a, (size_t)M * 4))) return rc;
    if ((rc = d2h(ctx, b.data(), ctx->ML.b, (size_t)M * 4))) return rc;
    if ((rc = d2h(ctx, el.data(), ctx->ML.edges_left, (size_t)M * 4))) return rc;
    if ((rc = d2h(ctx, rl.data(), ctx->ML.regions_left, (size_t)M * 4))) return rc;
    if ((rc = d2h(ctx, weight, ctx->ML.w, (size_t)M * 4))) return rc;
    if ((rc = fin(ctx))) return rc;
    for (unsigned m = 0; m < M; ++m) {
        if (ab) { ab[2 * m] = a[m]; ab[2 * m + 1] = b[m]; }
        if (left) { left[2 * m] = el[m]; left[2 * m + 1] = rl[m]; }
    }
    return F3PS_OK;
}

int f3ps_get_state_regions(f3ps_ctx* ctx, uint32_t* label, float* centroid, float* normal, int32_t* n_voxels, int64_t capacity) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_GRAPH, "f3ps_get_state_regions"); if (rc) return rc;
    const RegionArrays& R = ctx->progress >= P_MERGED ? ctx->R1 : ctx->R0;
    const unsigned S = ctx->S;
    std::vector<int> n(S); std::vector<float4> c(S), nn(S); std::vector<unsigned> lab(S);
    if ((rc = d2h(ctx, n.data(), R.n, (size_t)S * 4))) return rc;
    if ((rc = d2h(ctx, c.data(), R.centroid, (size_t)S * 16))) return rc;
    if ((rc = d2h(ctx, nn.data(), R.normal, (size_t)S * 16))) return rc;
    if ((rc = d2h(ctx, lab.data(), ctx->sv_label.p, (size_t)S * 4))) return rc;
    if ((rc = fin(ctx))) return rc;
    int64_t k = 0;
    for (unsigned s = 0; s < S; ++s) if (n[s] > 0) ++k;
    if ((rc = cap_check(ctx, k, capacity))) return rc;
    k = 0;
    for (unsigned s = 0; s < S; ++s) {
        if (n[s] <= 0) continue;
        if (label) label[k] = lab[s];
        if (centroid) { centroid[3 * k] = c[s].x; centroid[3 * k + 1] = c[s].y; centroid[3 * k + 2] = c[s].z; }
        if (normal) { normal[3 * k] = nn[s].x; normal[3 * k + 1] = nn[s].y; normal[3 * k + 2] = nn[s].z; }
        if (n_voxels) n_voxels[k] = n[s];
        ++k;
    }
    return F3PS_OK;
}

int f3ps_get_region_mean_color(f3ps_ctx* ctx, int32_t rank, float rgb[3]) {
    if (!ctx || !rgb) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_GRAPH, "f3ps_get_region_mean_color"); if (rc) return rc;
    if (rank < 0 || (unsigned)rank >= ctx->S) return ctx_fail(ctx, F3PS_ERR_INVALID_ARGUMENT, "region rank out of range");
    float4 m;
    if ((rc = d2h(ctx, &m, ctx->R0.mean + rank, 16))) return rc;
    if ((rc = fin(ctx))) return rc;
    rgb[0] = m.y; rgb[1] = m.z; rgb[2] = m.w;
    return F3PS_OK;
}

int f3ps_get_state_edges(f3ps_ctx* ctx, uint32_t* ab, float* weight, int64_t capacity) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_GRAPH, "f3ps_get_state_edges"); if (rc) return rc;
    const EdgeArrays& EA = ctx->progress >= P_MERGED ? ctx->E1 : ctx->E0;
    std::vector<unsigned> a, b, lab; std::vector<float> dc, dg, w; std::vector<long long> st;
    if ((rc = fetch_edges(ctx, EA, ctx->E, a, b, dc, dg, w, st, lab))) return rc;
    std::vector<unsigned> idx;
    for (unsigned e = 0; e < ctx->E; ++e) if (st[e] != kDeadStamp) idx.push_back(e);
    if ((rc = cap_check(ctx, (int64_t)idx.size(), capacity))) return rc;
    std::sort(idx.begin(), idx.end(), [&](unsigned x, unsigned y) { return w[x] < w[y] || (w[x] == w[y] && st[x] < st[y]); });   // multimap order
    for (size_t i = 0; i < idx.size(); ++i) {
        const unsigned e = idx[i];
        if (ab) { ab[2 * i] = lab[a[e]]; ab[2 * i + 1] = lab[b[e]]; }
        if (weight) weight[i] = w[e];
    }
    return F3PS_OK;
}

int f3ps_get_labeled_cloud(f3ps_ctx* ctx, float* xyz, uint32_t* label, uint32_t* voxel_index, int64_t capacity) {
    if (!ctx) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_MERGED, "f3ps_get_labeled_cloud"); if (rc) return rc;
    if ((rc = cap_check(ctx, ctx->n_out, capacity))) return rc;
    if ((rc = d2h(ctx, xyz, ctx->out_xyz.p, (size_t)ctx->n_out * 12))) return rc;
    if ((rc = d2h(ctx, label, ctx->out_label.p, (size_t)ctx->n_out * 4))) return rc;
    if ((rc = d2h(ctx, voxel_index, ctx->out_voxel.p, (size_t)ctx->n_out * 4))) return rc;
    return fin(ctx);
}

int f3ps_get_voxel_segments_device(f3ps_ctx* ctx, const uint32_t** device_ptr, int64_t* n) {
    if (!ctx || !device_ptr || !n) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_MERGED, "f3ps_get_voxel_segments_device"); if (rc) return rc;
    *device_ptr = ctx->vox_segment.as<uint32_t>(); *n = ctx->n_pos;
    return F3PS_OK;
}

// ---- device self tests ------------------------------------------------------------------------------
// =====================================================================================================
// "next" row f1: Clustering::all_thresh (src/clustering.cpp:691-729) + Testing::eval_performance (src/testing.cpp:239-362).
extern "C++" {
namespace {
// Testing::compute_intersections' best matches + the seven scores, in the reference's float order.
// table[n_seg][n_truth] = inter_matrix; g[j] = truth segment sizes; N = truth->size().
f3ps_performance testing_scores(const std::vector<unsigned>& table, unsigned n_seg, unsigned n_truth, const std::vector<size_t>& g, size_t N) {
    // table has n_truth + 1 columns: the last one counts segmentation points without a ground-truth point at the same xyz
    auto inter = [&](unsigned i, unsigned j) -> size_t { return table[(size_t)i * (n_truth + 1) + j]; };
    std::vector<size_t> ssz(n_seg, 0);
    for (unsigned i = 0; i < n_seg; ++i) for (unsigned j = 0; j <= n_truth; ++j) ssz[i] += inter(i, j);
    // t_sizes: std::map<size_t, uint32_t>, insert keeps the FIRST truth segment of every size (testing.cpp:97-110)
    std::map<size_t, unsigned> t_sizes;
    if (n_seg) for (unsigned j = 0; j < n_truth; ++j) t_sizes.insert(std::make_pair(g[j], j));
    std::vector<long long> matches(n_truth, -1);
    for (auto it = t_sizes.rbegin(); it != t_sizes.rend(); ++it) {           // largest truth segment first (:112-137)
        const unsigned j = it->second;
        std::vector<size_t> col(n_seg);
        for (unsigned i = 0; i < n_seg; ++i) col[i] = inter(i, j);
        auto argmax = [&]() { long long r = 0; for (unsigned i = 1; i < n_seg; ++i) if (col[i] > col[(size_t)r]) r = i; return r; };   // Eigen maxCoeff: first maximum
        long long row = argmax();
        auto taken = [&](long long r) { for (long long m : matches) if (m == r) return true; return false; };
        while (taken(row)) {
            col[(size_t)row] = 0;
            bool any = false; for (size_t c : col) any = any || c != 0;
            if (any) row = argmax(); else { row = -1; break; }
        }
        matches[j] = row;
    }
    f3ps_performance pf;
    const float n = (float)N;
    {   // eval_voi (:305-335)
        float h_s = 0, h_t = 0, mi = 0;
        for (unsigned i = 0; i < n_seg; ++i) {
            const float p = (float)ssz[i];
            h_s -= std::log(p / n) * p / n;
            for (unsigned j = 0; j < n_truth; ++j) {
                const float q = (float)g[j];
                if (i == 0) h_t -= std::log(q / n) * q / n;
                const float r = (float)inter(i, j);
                if (r != 0) mi += std::log(((n * r) / (p * q))) * r / n;
            }
        }
        pf.voi = h_s + h_t - 2 * mi;
    }
    {   // eval_precision (:239-268)
        float p = 0, r = 0, fp = 0, fn = 0;
        for (unsigned j = 0; j < n_truth; ++j) {
            const long long i = matches[j];
            if (i != -1) {
                const float in = (float)inter((unsigned)i, j), sz = (float)ssz[(size_t)i], gg = (float)g[j];
                p += in * gg / sz; r += in; fp += (sz - in); fn += (gg - in);
            } else fn += (float)g[j];
        }
        pf.precision = p / n; pf.recall = r / n; pf.fpr = fp / n; pf.fnr = fn / n;
    }
    if (pf.precision == 0 && pf.recall == 0) pf.fscore = 0;                 // eval_fscore (:287-299)
    else pf.fscore = 2 * (pf.precision * pf.recall) / (pf.precision + pf.recall);
    {   // eval_wov (:342-357): count_union of two point sets = |s| + |t| - |s ^ t|
        float w = 0;
        for (unsigned j = 0; j < n_truth; ++j) {
            const long long i = matches[j];
            if (i != -1) {
                const float in = (float)inter((unsigned)i, j);
                const float un = (float)(ssz[(size_t)i] + g[j] - inter((unsigned)i, j));
                w += in * (float)g[j] / un;
            }
        }
        pf.wov = w / n;
    }
    return pf;
}
} // namespace
} // extern "C++"

int f3ps_eval_thresholds(f3ps_ctx* ctx, const uint32_t* truth_label, int64_t n_voxels, const uint32_t* extra_truth_label, int64_t n_extra,
                         const float* thresholds, int n_thresholds, f3ps_performance* perf, int32_t* n_segments, int32_t* n_merges_at) {
    if (!ctx || !truth_label || !thresholds || !perf || n_thresholds < 1 || n_extra < 0 || (n_extra > 0 && !extra_truth_label)) return F3PS_ERR_INVALID_ARGUMENT;
    int rc = need(ctx, P_GRAPH, "f3ps_eval_thresholds"); if (rc) return rc;
    if (n_voxels != (int64_t)ctx->V) return ctx_fail(ctx, F3PS_ERR_INVALID_ARGUMENT, "one ground-truth label per voxel is required");
    for (int k = 1; k < n_thresholds; ++k) if (!(thresholds[k] >= thresholds[k - 1])) return ctx_fail(ctx, F3PS_ERR_INVALID_ARGUMENT, "thresholds must ascend");
    cudaSetDevice(ctx->device);
    rc = f3ps_merge(ctx, thresholds[n_thresholds - 1]); if (rc) return rc;      // ONE replay; every threshold is a prefix of it
    const unsigned S = ctx->S, V = ctx->V, P = ctx->n_pos;
    const unsigned M = ctx->h_sc->mctl.n_merges;
    if (ctx->n_out == 0) return ctx_fail(ctx, F3PS_ERR_INVALID_ARGUMENT, "The pointcloud to be set as 'segm' cannot be empty");  // testing.cpp:413-416
    // merge forest on the host (M, S are small): label -> rank, parent / time per initial supervoxel
    std::vector<unsigned> la(M), lb(M), svl(S); std::vector<float> lw(M);
    if (M) {
        if ((rc = d2h(ctx, la.data(), ctx->ML.a, (size_t)M * 4))) return rc;
        if ((rc = d2h(ctx, lb.data(), ctx->ML.b, (size_t)M * 4))) return rc;
        if ((rc = d2h(ctx, lw.data(), ctx->ML.w, (size_t)M * 4))) return rc;
    }
    if ((rc = d2h(ctx, svl.data(), ctx->sv_label.p, (size_t)S * 4))) return rc;
    if ((rc = fin(ctx))) return rc;
    std::map<unsigned, unsigned> rank_of;
    for (unsigned s = 0; s < S; ++s) rank_of[svl[s]] = s;
    std::vector<unsigned> parent(S), when(S, 0xffffffffu);
    for (unsigned s = 0; s < S; ++s) parent[s] = s;
    for (unsigned i = 0; i < M; ++i) { const unsigned ra = rank_of[la[i]], rb = rank_of[lb[i]]; parent[rb] = ra; when[rb] = i; }
    // Testing::label_map on the truth: dense labels in ascending label order, segment sizes
    std::map<unsigned, unsigned> tmap;
    size_t n_truth_points = 0;
    for (int64_t v = 0; v < n_voxels; ++v) if (truth_label[v] != 0xffffffffu) { tmap[truth_label[v]] = 0; ++n_truth_points; }
    for (int64_t v = 0; v < n_extra; ++v) { tmap[extra_truth_label[v]] = 0; ++n_truth_points; }
    if (n_truth_points == 0) return ctx_fail(ctx, F3PS_ERR_INVALID_ARGUMENT, "The pointcloud to be set as 'truth' cannot be empty");
    unsigned Kt = 0; for (auto& kv : tmap) kv.second = Kt++;
    std::vector<unsigned> tdense((size_t)V); std::vector<size_t> g(Kt, 0);
    for (unsigned v = 0; v < V; ++v) {
        if (truth_label[v] == 0xffffffffu) { tdense[v] = Kt; continue; }     // no ground-truth point at this voxel's xyz
        tdense[v] = tmap[truth_label[v]]; g[tdense[v]]++;
    }
    for (int64_t v = 0; v < n_extra; ++v) g[tmap[extra_truth_label[v]]]++;    // truth points outside the segmentation's voxels
    const unsigned Kc = Kt + 1;
    if ((size_t)S * Kc > ((size_t)1 << 28)) return ctx_fail(ctx, F3PS_ERR_CAPACITY, "contingency table above 2^28 cells");
    const size_t Sc = std::max(1u, S);
    F3PS_CUDA_OK(ctx->ev_parent.ensure(Sc * 4)); F3PS_CUDA_OK(ctx->ev_when.ensure(Sc * 4)); F3PS_CUDA_OK(ctx->ev_dense.ensure(Sc * 4 + 4));
    F3PS_CUDA_OK(ctx->ev_truth.ensure((size_t)V * 4)); F3PS_CUDA_OK(ctx->ev_table.ensure(std::max<size_t>(1, (size_t)S * Kc) * 4));
    F3PS_CUDA_OK(cudaMemcpyAsync(ctx->ev_parent.p, parent.data(), (size_t)S * 4, cudaMemcpyHostToDevice, ctx->stream));
    F3PS_CUDA_OK(cudaMemcpyAsync(ctx->ev_when.p, when.data(), (size_t)S * 4, cudaMemcpyHostToDevice, ctx->stream));
    F3PS_CUDA_OK(cudaMemcpyAsync(ctx->ev_truth.p, tdense.data(), (size_t)V * 4, cudaMemcpyHostToDevice, ctx->stream));
    unsigned* d_nseg = ctx->ev_dense.as<unsigned>() + Sc;
    std::vector<unsigned> table;
    unsigned m_prev = 0xffffffffu;
    for (int k = 0; k < n_thresholds; ++k) {
        unsigned m = M;                                                       // first step whose head weight is >= t (strict <, :388-389)
        for (unsigned i = 0; i < M; ++i) if (!(lw[i] < thresholds[k])) { m = i; break; }
        if (n_merges_at) n_merges_at[k] = (int32_t)m;
        if (m == m_prev) { perf[k] = perf[k - 1]; if (n_segments) n_segments[k] = n_segments[k - 1]; continue; }
        m_prev = m;
        LAUNCH(ctx, eval_dense_kernel, 1, 1024, 0, ctx->ev_when.as<unsigned>(), S, m, ctx->ev_dense.as<unsigned>(), d_nseg);
        F3PS_CUDA_OK(cudaMemsetAsync(ctx->ev_table.p, 0, std::max<size_t>(1, (size_t)S * Kc) * 4, ctx->stream));
        if (P) LAUNCH(ctx, eval_table_kernel, grid_for(P, 256), 256, 0, ctx->pos_run.as<unsigned>(), ctx->order, P, ctx->ev_parent.as<unsigned>(),
                      ctx->ev_when.as<unsigned>(), ctx->ev_dense.as<unsigned>(), m, ctx->ev_truth.as<unsigned>(), Kc, ctx->ev_table.as<unsigned>());
        unsigned n_seg = 0;
        if ((rc = d2h(ctx, &n_seg, d_nseg, 4))) return rc;
        if ((rc = fin(ctx))) return rc;
        table.resize((size_t)n_seg * Kc);
        if (n_seg) { if ((rc = d2h(ctx, table.data(), ctx->ev_table.p, (size_t)n_seg * Kc * 4))) return rc; if ((rc = fin(ctx))) return rc; }
        perf[k] = testing_scores(table, n_seg, Kt, g, n_truth_points);
        if (n_segments) n_segments[k] = (int32_t)n_seg;
    }
    return F3PS_OK;
}


int f3ps_eval_label_pairs(f3ps_ctx* ctx, const uint32_t* seg, const uint32_t* truth, int64_t n_pairs, int32_t n_seg, int32_t n_truth,
                          const uint64_t* truth_sizes, int64_t n_truth_points, f3ps_performance* perf) {
    if (!ctx || !perf || n_pairs < 0 || n_seg < 0 || n_truth < 0 || (n_pairs > 0 && (!seg || !truth)) || (n_truth > 0 && !truth_sizes)) return F3PS_ERR_INVALID_ARGUMENT;
    if (n_pairs == 0) return ctx_fail(ctx, F3PS_ERR_INVALID_ARGUMENT, "The pointcloud to be set as 'segm' cannot be empty");       // testing.cpp:413-416
    if (n_truth_points <= 0) return ctx_fail(ctx, F3PS_ERR_INVALID_ARGUMENT, "The pointcloud to be set as 'truth' cannot be empty"); // :430-433
    for (int64_t i = 0; i < n_pairs; ++i)
        if (seg[i] >= (uint32_t)n_seg || truth[i] > (uint32_t)n_truth) return ctx_fail(ctx, F3PS_ERR_INVALID_ARGUMENT, "f3ps_eval_label_pairs: label out of range");
    const unsigned Kc = (unsigned)n_truth + 1u;
    if ((size_t)n_seg * Kc > ((size_t)1 << 28)) return ctx_fail(ctx, F3PS_ERR_CAPACITY, "contingency table above 2^28 cells");
    cudaSetDevice(ctx->device);
    const size_t cells = std::max<size_t>(1, (size_t)n_seg * Kc);
    F3PS_CUDA_OK(ctx->ev_dense.ensure((size_t)n_pairs * 4)); F3PS_CUDA_OK(ctx->ev_truth.ensure((size_t)n_pairs * 4)); F3PS_CUDA_OK(ctx->ev_table.ensure(cells * 4));
    F3PS_CUDA_OK(cudaMemcpyAsync(ctx->ev_dense.p, seg, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, ctx->stream));
    F3PS_CUDA_OK(cudaMemcpyAsync(ctx->ev_truth.p, truth, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, ctx->stream));
    F3PS_CUDA_OK(cudaMemsetAsync(ctx->ev_table.p, 0, cells * 4, ctx->stream));
    LAUNCH(ctx, eval_pairs_kernel, grid_for(n_pairs, 256), 256, 0, ctx->ev_dense.as<unsigned>(), ctx->ev_truth.as<unsigned>(), (long long)n_pairs, Kc, ctx->ev_table.as<unsigned>());
    std::vector<unsigned> table(cells);
    int rc = d2h(ctx, table.data(), ctx->ev_table.p, cells * 4); if (rc) return rc;
    if ((rc = fin(ctx))) return rc;
    std::vector<size_t> g((size_t)n_truth);
    for (int32_t j = 0; j < n_truth; ++j) g[(size_t)j] = (size_t)truth_sizes[j];
    *perf = testing_scores(table, (unsigned)n_seg, (unsigned)n_truth, g, (size_t)n_truth_points);
    return F3PS_OK;
}

static int test_map3(f3ps_ctx* ctx, int which, const float* in1, const float* in2, float* out, int64_t n) {
    if (!ctx || n < 0) return F3PS_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    DevBuf a, b, o;
    const size_t nb = (size_t)std::max<int64_t>(n, 1) * 12;
    int rc = F3PS_OK;
    do {
        if (a.ensure(nb) != cudaSuccess || b.ensure(nb) != cudaSuccess || o.ensure(nb) != cudaSuccess) { rc = ctx_fail(ctx, F3PS_ERR_CUDA, "cudaMalloc failed"); break; }
        cudaMemcpyAsync(a.p, in1, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->stream);
        if (in2) cudaMemcpyAsync(b.p, in2, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->stream);
        const int g = grid_for(n, 128);
        if (which == 0) lab_test_kernel<<<g, 128, 0, ctx->stream>>>(ctx->d_lab_lut, a.as<float>(), o.as<float>(), n);
        else if (which == 1) ciede_test_kernel<<<g, 128, 0, ctx->stream>>>(a.as<float>(), b.as<float>(), o.as<float>(), n);
        else rgb_eucl_test_kernel<<<g, 128, 0, ctx->stream>>>(a.as<float>(), b.as<float>(), o.as<float>(), n);
        ctx->launches++;
        cudaMemcpyAsync(out, o.p, (size_t)n * (which == 0 ? 12 : 4), cudaMemcpyDeviceToHost, ctx->stream);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = ctx_fail_cuda(ctx, e, "self test", __FILE__, __LINE__);
    } while (0);
    a.release(); b.release(); o.release();
    return rc;
}
int f3ps_test_rgb2lab(f3ps_ctx* ctx, const float* rgb255, float* lab, int64_t n) { return test_map3(ctx, 0, rgb255, nullptr, lab, n); }
int f3ps_test_lab_ciede00(f3ps_ctx* ctx, const float* l1, const float* l2, float* out, int64_t n) { return test_map3(ctx, 1, l1, l2, out, n); }
int f3ps_test_rgb_eucl(f3ps_ctx* ctx, const float* c1, const float* c2, float* out, int64_t n) { return test_map3(ctx, 2, c1, c2, out, n); }

int f3ps_test_sort_pairs(f3ps_ctx* ctx, uint64_t* keys, uint32_t* values, int64_t n, int key_bits) {
    if (!ctx || n < 0 || key_bits < 1 || key_bits > 64) return F3PS_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    DevBuf ka, kb, va, vb;
    const size_t nk = (size_t)std::max<int64_t>(n, 1);
    int rc = F3PS_OK;
    do {
        if (ka.ensure(nk * 8) != cudaSuccess || kb.ensure(nk * 8) != cudaSuccess || va.ensure(nk * 4) != cudaSuccess || vb.ensure(nk * 4) != cudaSuccess) {
            rc = ctx_fail(ctx, F3PS_ERR_CUDA, "cudaMalloc failed"); break; }
        cudaMemcpyAsync(ka.p, keys, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(va.p, values, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream);
        unsigned long long* sk = ka.as<unsigned long long>(); unsigned* sv = va.as<unsigned>();
        if (n) {
            rc = sort_pairs<unsigned long long>(ctx, ka.as<unsigned long long>(), va.as<unsigned>(), kb.as<unsigned long long>(), vb.as<unsigned>(),
                                                ka.as<unsigned long long>(), va.as<unsigned>(), nullptr, n, key_bits, &sk, &sv);
            if (rc) break;
        }
        cudaMemcpyAsync(keys, sk, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(values, sv, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = ctx_fail_cuda(ctx, e, "sort self test", __FILE__, __LINE__);
    } while (0);
    ka.release(); kb.release(); va.release(); vb.release();
    return rc;
}

} // extern "C"
