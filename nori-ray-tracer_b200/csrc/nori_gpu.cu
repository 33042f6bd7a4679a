// nori_gpu.cu -- the extern "C" boundary of include/nori_gpu.h over the kernels in kernels.cuh.
// One context = one CUDA device + one stream.  No torch, no exceptions across the ABI, no CPU
// fallback: every entry point that computes does so on the device or fails with an error string.
#define NORI_WITH_PERLIN 1      // see traverse.cuh
#include "film_kernels.cuh"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#define NORI_DEFAULT_RESULTS_MB 8192
static std::string g_init_error;

#define NORI_MAX_WAVEFRONTS 4
#ifndef NORI_SHADE_GRID_PER_SM
#define NORI_SHADE_GRID_PER_SM 16         // CTAs of k_shade per SM (two rounds of the 8 resident ones)
#endif
struct nori_gpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    bool has_scene = false, has_perlin = false;

    DScene ds{};
    std::vector<void *> scene_allocs;  // unused by the arena path; kept for freeAll symmetry
    // all scene arrays live in ONE device allocation that is reused by the next upload when it is large enough:
    // re-uploading a scene (the e2e loop, progressive rendering) then costs no cudaMalloc / cudaFree at all
    // (their occasional 100+ ms driver stalls showed up in the end-to-end numbers)
    char *arena = nullptr; size_t arena_cap = 0, arena_used = 0;
    size_t film_cap = 0;
    nori_gpu_filter filter{};
    int W = 0, H = 0, border = 0;
    uint32_t bsdf_mask = 0;            // which BSDF types occur in the scene
    uint32_t n_bsdfs = 0;

    float4 *film = nullptr;
    float4 *vsum = nullptr, *vsum2 = nullptr; uint32_t var_passes = 0;   // variance statistic (option "variance")
    int64_t opt_variance = 0;
    Pool pool{};
    std::vector<void *> pool_allocs;
    float4 *results = nullptr; size_t results_cap = 0;      // in float4 elements
    Counters *ctr = nullptr; Counters *h_ctr = nullptr;      // device / pinned host
    // concurrent wavefronts (option "wavefronts", traceBatch): wavefront 0 uses `stream` / `ctr` / `h_ctr`, the others these
    cudaStream_t wf_stream[NORI_MAX_WAVEFRONTS - 1] = {}; Counters *wf_ctr[NORI_MAX_WAVEFRONTS - 1] = {}, *wf_h_ctr[NORI_MAX_WAVEFRONTS - 1] = {};
    cudaEvent_t wf_done[NORI_MAX_WAVEFRONTS - 1] = {};
    int wf_used = 1;                                        // wavefronts of the last batch (foldStats adds their counters)
    int64_t opt_wavefronts = 2;
    float4 *flush_buf = nullptr; size_t flush_n = 0;
    void *scratch = nullptr; size_t scratch_cap = 0;       // staging for resolve / variance / trace / probes / pcg32

    // options
    int64_t opt_pool = 1 << 20, opt_results_mb = NORI_DEFAULT_RESULTS_MB, opt_stats = 0, opt_megakernel = 0, opt_poll = 8;
    int64_t opt_emitter_sort = 1;      // path_mis with emitters of several types: (material, emitter type)-sorted shading queues
    uint32_t n_emitter_types = 0, emitter_type_mask = 0; bool has_envmap = false;
    int64_t opt_area_only = 1;         // scenes lit by area lights only: shade kernels compiled without the other emitter types
    int64_t opt_film_sep = 1;          // radius-2 filters: film kernel with per-sample tabulated weights (0: generic kernel)
    int64_t opt_film_tma = 1;          // ... whose sample tiles are staged by the TMA unit (0: per-thread loads)
    int64_t opt_drain = 1 << 15;       // finish the batch with k_drain once at most this many paths are alive (0: never)
    int64_t opt_drain_mode = 0;        // 0: one warp per remaining path (k_drain_warp), 1: one thread per path (k_drain)
    int64_t opt_shadow_pass = 0;       // 0 auto (own pass with the state-machine traversal), 1 always, 2 never (inside k_shade)
    // 0 reference child order (counters equal the reference's), 1 near child first, 2 auto: near child first when
    // rendering scenes with deep trees, reference order for small scenes and for the nori_gpu_trace test hook
    int64_t opt_order = 2;
    int64_t opt_wide = 1;              // 1: large-scene kernels walk the 4-wide layout (with the near-first order)
    int64_t opt_traversal = 0;         // 0 auto (by primitive count), 1 plain per-lane loops, 2 warp state machine
    int64_t opt_trace_kernel = 0;      // nori_gpu_trace: 0 k_trace (per-lane loops over the reference nodes), 2 the large-scene render kernels
    int64_t opt_l2_window = 0;         // MiB of the 4-wide records (top of the tree first) held by an L2 access-policy window; 0 = none
    size_t nodes4_bytes = 0; bool window_set = false;

    nori_gpu_stats stats{};
    nori_gpu_kernel_stats kstats[NORI_K_COUNT]{};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // optional per-launch timing (option "kernel_timing")
    int64_t opt_kernel_timing = 0;
    std::vector<cudaEvent_t> kev; std::vector<int> kev_kind; size_t kev_used = 0;
    bool last_wave = false, last_defer = false;

    // multi-GPU context (nori_gpu_init_multi): this is the context of devices[0]; `peers` are complete contexts of
    // the other devices (scene replicated, sample indices sharded in nori_gpu_render, films summed onto this one)
    std::vector<nori_gpu_ctx *> peers;
    std::vector<char> peer_direct;     // peers[i]'s film is readable from this device over NVLink (peer access enabled)
    float4 *peer_stage = nullptr; size_t peer_stage_cap = 0;   // staging for peers without direct access
    float multi_reduce_ms = 0.f;
};

// bracket one kernel launch: counts it and, with kernel_timing on, records a CUDA event pair on the stream
static inline void launchBegin(nori_gpu_ctx *ctx, int kind) {
    ctx->kstats[kind].launches++; ctx->stats.kernel_launches++;
    if (!ctx->opt_kernel_timing) return;
    if (ctx->kev_used + 2 > ctx->kev.size()) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); ctx->kev.push_back(a); ctx->kev.push_back(b); }
    ctx->kev_kind.push_back(kind);
    cudaEventRecord(ctx->kev[ctx->kev_used], ctx->stream);
}
static inline void launchEnd(nori_gpu_ctx *ctx) {
    if (!ctx->opt_kernel_timing) return;
    cudaEventRecord(ctx->kev[ctx->kev_used + 1], ctx->stream);
    ctx->kev_used += 2;
}
#define LAUNCH(kind, ...) do { launchBegin(ctx, kind); __VA_ARGS__; launchEnd(ctx); } while (0)

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return 1; } } while (0)
#define REQUIRE(cond, msg) do { if (!(cond)) { ctx->err = (msg); return 1; } } while (0)

template <typename T> static int devUpload(nori_gpu_ctx *ctx, std::vector<void *> &, const T *src, size_t n, const T **out) {
    *out = nullptr;
    if (!n) return 0;
    const size_t off = (ctx->arena_used + 255) & ~(size_t) 255, bytes = n * sizeof(T);
    if (off + bytes > ctx->arena_cap) { ctx->err = "upload_scene: scene arena too small (internal size estimate is wrong)"; return 1; }
    void *d = ctx->arena + off;
    ctx->arena_used = off + bytes;
    CK(cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    *out = (const T *) d;
    return 0;
}

// upper bound of the device bytes nori_gpu_upload_scene needs for `s` (every array padded to 256 bytes)
static size_t sceneArenaBytes(const nori_gpu_scene *s) {
    size_t total = 0, arrays = 16;
    auto add = [&](size_t bytes) { total += bytes; ++arrays; };
    for (uint32_t i = 0; i < s->n_images; ++i) add((size_t) s->images[i].width * s->images[i].height * 3);
    for (uint32_t i = 0; i < s->n_shapes; ++i) {
        const nori_gpu_shape &h = s->shapes[i];
        if (h.type != NORI_SHAPE_MESH) continue;
        add(12 * (size_t) h.n_vertices); add(12 * (size_t) h.n_vertices); add(8 * (size_t) h.n_vertices);
        add(12 * (size_t) h.n_triangles); add(4 * ((size_t) h.n_triangles + 1));
    }
    for (uint32_t i = 0; i < s->n_emitters; ++i) {
        const nori_gpu_emitter &e = s->emitters[i];
        if (e.type != NORI_EMITTER_ENVMAP) continue;
        const size_t R = e.env_rows > 0 ? e.env_rows : 0, C = e.env_cols > 0 ? e.env_cols : 0;
        add(R * C * 12); add(R * C * 4); add(R * (C + 1) * 4); add(R * 4); add((R + 1) * 4);
    }
    add(32 * (size_t) s->n_nodes); add(64 * (size_t) s->n_nodes); add(128 * (((size_t) s->n_nodes + 1) / 2)); add(48 * (size_t) s->n_indices);
    add(sizeof(DShape) * (size_t) s->n_shapes); add(sizeof(nori_gpu_bsdf) * (size_t) s->n_bsdfs);
    add(sizeof(DEmitter) * (size_t) s->n_emitters); add(sizeof(DImage) * (size_t) s->n_images);
    return total + 256 * arrays;
}

static void freeAll(std::vector<void *> &v) { for (void *p : v) cudaFree(p); v.clear(); }

extern "C" {

int nori_gpu_init(int device, nori_gpu_ctx **out) {
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_init_error = std::string("nori_gpu_init: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback";
        return 1;
    }
    if (device < 0 || device >= n) { g_init_error = "nori_gpu_init: device index out of range"; return 1; }
    nori_gpu_ctx *ctx = new nori_gpu_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess
        || cudaMalloc((void **) &ctx->ctr, sizeof(Counters)) != cudaSuccess
        || cudaMallocHost((void **) &ctx->h_ctr, sizeof(Counters)) != cudaSuccess
        || cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) {
        g_init_error = std::string("nori_gpu_init: ") + cudaGetErrorString(cudaGetLastError());
        delete ctx; return 1;
    }
    cudaMemsetAsync(ctx->ctr, 0, sizeof(Counters), ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    *out = ctx;
    return 0;
}

/* one context over several devices of one node.  The host calls the same entry points; upload_scene
 * replicates the scene, render shards the sample indices and sums the films (k_film_reduce over peer memory). */
int nori_gpu_init_multi(const int *devices, int n, nori_gpu_ctx **out) {
    *out = nullptr;
    if (!devices || n < 1) { g_init_error = "nori_gpu_init_multi: empty device list"; return 1; }
    for (int i = 0; i < n; ++i) for (int j = 0; j < i; ++j)
        if (devices[i] == devices[j]) { g_init_error = "nori_gpu_init_multi: device listed twice"; return 1; }
    nori_gpu_ctx *root = nullptr;
    if (nori_gpu_init(devices[0], &root)) return 1;
    for (int i = 1; i < n; ++i) {
        nori_gpu_ctx *p = nullptr;
        if (nori_gpu_init(devices[i], &p)) { nori_gpu_destroy(root); return 1; }
        root->peers.push_back(p);
        int can = 0;
        cudaSetDevice(root->device);
        if (cudaDeviceCanAccessPeer(&can, root->device, devices[i]) == cudaSuccess && can) {
            cudaError_t e = cudaDeviceEnablePeerAccess(devices[i], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
            can = e == cudaSuccess;
            if (!can) cudaGetLastError();
        }
        root->peer_direct.push_back((char) can);
    }
    *out = root;
    return 0;
}

int nori_gpu_device_count(const nori_gpu_ctx *ctx) { return ctx ? 1 + (int) ctx->peers.size() : 0; }

void nori_gpu_destroy(nori_gpu_ctx *ctx) {
    if (!ctx) return;
    for (nori_gpu_ctx *p : ctx->peers) nori_gpu_destroy(p);
    ctx->peers.clear();
    cudaSetDevice(ctx->device);
    cudaFree(ctx->peer_stage);
    cudaStreamSynchronize(ctx->stream);
    freeAll(ctx->scene_allocs); freeAll(ctx->pool_allocs);
    cudaFree(ctx->arena);
    cudaFree(ctx->film); cudaFree(ctx->vsum); cudaFree(ctx->vsum2); cudaFree(ctx->results); cudaFree(ctx->ctr); cudaFree(ctx->flush_buf); cudaFree(ctx->scratch);
    cudaFreeHost(ctx->h_ctr);
    for (int j = 0; j < NORI_MAX_WAVEFRONTS - 1; ++j) {
        if (ctx->wf_stream[j]) cudaStreamDestroy(ctx->wf_stream[j]);
        if (ctx->wf_done[j]) cudaEventDestroy(ctx->wf_done[j]);
        cudaFree(ctx->wf_ctr[j]); if (ctx->wf_h_ctr[j]) cudaFreeHost(ctx->wf_h_ctr[j]);
    }
    cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1);
    for (cudaEvent_t e : ctx->kev) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *nori_gpu_last_error(const nori_gpu_ctx *ctx) { return ctx ? ctx->err.c_str() : g_init_error.c_str(); }

int nori_gpu_abi_sizes(uint32_t *out, int n) {
    const uint32_t s[] = {sizeof(nori_gpu_bvh_node), sizeof(nori_gpu_shape), sizeof(nori_gpu_bsdf), sizeof(nori_gpu_emitter),
                          sizeof(nori_gpu_camera), sizeof(nori_gpu_filter), sizeof(nori_gpu_medium), sizeof(nori_gpu_scene),
                          sizeof(nori_gpu_ray), sizeof(nori_gpu_hit), sizeof(nori_gpu_stats), sizeof(nori_gpu_image)};
    int m = (int) (sizeof(s) / sizeof(s[0]));
    for (int i = 0; i < n && i < m; ++i) out[i] = s[i];
    return m;
}

int nori_gpu_set_option(nori_gpu_ctx *ctx, const char *name, int64_t value) {
    REQUIRE(ctx && name, "set_option: null argument");
    std::string k(name);
    if (!ctx->peers.empty()) {
        REQUIRE(!(k == "variance" && value), "set_option(variance): the running-mean statistic (render.cpp:238-247) depends on the order of "
                                             "all passes and does not decompose over devices: use a single-device context");
        for (nori_gpu_ctx *p : ctx->peers)
            if (nori_gpu_set_option(p, name, value)) { ctx->err = p->err; return 1; }
    }
    if (k == "pool") { REQUIRE(value >= 1024 && value <= (1ll << 26), "pool must be in [1024, 2^26]"); ctx->opt_pool = value; freeAll(ctx->pool_allocs); ctx->pool = Pool{}; }
    else if (k == "reset_options") {
        // every scheduling option back to its default (tests: a finalizer calls this so that no test leaks its settings)
        if (ctx->opt_pool != (1 << 20)) { ctx->opt_pool = 1 << 20; freeAll(ctx->pool_allocs); ctx->pool = Pool{}; }
        ctx->opt_results_mb = NORI_DEFAULT_RESULTS_MB; ctx->opt_stats = 0; ctx->opt_megakernel = 0; ctx->opt_poll = 8; ctx->opt_emitter_sort = 1;
        ctx->opt_area_only = 1; ctx->opt_film_sep = 1; ctx->opt_film_tma = 1; ctx->opt_drain = 1 << 15; ctx->opt_shadow_pass = 0; ctx->opt_order = 2;
        ctx->opt_wide = 1; ctx->opt_traversal = 0; ctx->opt_trace_kernel = 0; ctx->opt_kernel_timing = 0; ctx->opt_l2_window = 0; ctx->opt_drain_mode = 0;
        ctx->opt_wavefronts = 2;
    }
    else if (k == "results_mb") { REQUIRE(value >= 16, "results_mb must be >= 16"); ctx->opt_results_mb = value; }
    else if (k == "stats") ctx->opt_stats = value != 0;
    else if (k == "megakernel") ctx->opt_megakernel = value != 0;
    else if (k == "kernel_timing") ctx->opt_kernel_timing = value != 0;
    else if (k == "variance") {
        REQUIRE(ctx->has_scene, "set_option(variance): upload a scene first");
        CK(cudaSetDevice(ctx->device));
        ctx->opt_variance = value != 0;
        size_t nf = (size_t) (ctx->W + 2 * ctx->border) * (ctx->H + 2 * ctx->border);
        if (ctx->opt_variance && !ctx->vsum) {
            CK(cudaMalloc((void **) &ctx->vsum, nf * sizeof(float4))); CK(cudaMalloc((void **) &ctx->vsum2, nf * sizeof(float4)));
            CK(cudaMemsetAsync(ctx->vsum, 0, nf * sizeof(float4), ctx->stream)); CK(cudaMemsetAsync(ctx->vsum2, 0, nf * sizeof(float4), ctx->stream));
            ctx->var_passes = 0;
        }
    }
    else if (k == "traversal") { REQUIRE(value >= 0 && value <= 2, "traversal must be 0, 1 or 2"); ctx->opt_traversal = value; }
    else if (k == "trace_kernel") { REQUIRE(value == 0 || value == 2, "trace_kernel must be 0 (k_trace) or 2 (k_extend_sm / k_shadow_sm)"); ctx->opt_trace_kernel = value; }
    else if (k == "l2_window") { REQUIRE(value >= 0 && value <= 1024, "l2_window must be in [0, 1024] MiB"); ctx->opt_l2_window = value; }
    else if (k == "wide") { REQUIRE(value == 0 || value == 1, "wide must be 0 or 1"); ctx->opt_wide = value; }
    else if (k == "order") { REQUIRE(value >= 0 && value <= 2, "order must be 0 (reference child order), 1 (near child first) or 2 (auto)"); ctx->opt_order = value; }
    else if (k == "area_only") ctx->opt_area_only = value != 0;
    else if (k == "emitter_sort") { REQUIRE(value >= 0 && value <= 2, "emitter_sort must be 0 (off), 1 (auto) or 2 (always)"); ctx->opt_emitter_sort = value; }
    else if (k == "film_sep") ctx->opt_film_sep = value != 0;
    else if (k == "film_tma") ctx->opt_film_tma = value != 0;
    else if (k == "drain_mode") { REQUIRE(value == 0 || value == 1, "drain_mode must be 0 (one warp per path) or 1 (one thread per path)"); ctx->opt_drain_mode = value; }
    else if (k == "drain") { REQUIRE(value >= 0, "drain must be >= 0"); ctx->opt_drain = value; }
    else if (k == "shadow_pass") { REQUIRE(value >= 0 && value <= 2, "shadow_pass must be 0 (auto), 1 (own pass) or 2 (inside k_shade)"); ctx->opt_shadow_pass = value; }
    else if (k == "wavefronts") { REQUIRE(value >= 1 && value <= NORI_MAX_WAVEFRONTS, "wavefronts must be in [1,4]"); ctx->opt_wavefronts = value; }
    else if (k == "poll") { REQUIRE(value >= 1 && value <= 1024, "poll must be in [1,1024]"); ctx->opt_poll = value; }
    else if (k == "flush_l2") {
        // bench helper: overwrite a buffer larger than L2 (value = MiB)
        CK(cudaSetDevice(ctx->device));
        size_t n = (size_t) value * (1 << 20) / sizeof(float4);
        if (n > ctx->flush_n) { cudaFree(ctx->flush_buf); ctx->flush_buf = nullptr; CK(cudaMalloc((void **) &ctx->flush_buf, n * sizeof(float4))); ctx->flush_n = n; }
        k_flush<<<148 * 8, 256, 0, ctx->stream>>>(ctx->flush_buf, n);
        CK(cudaGetLastError());
    } else { ctx->err = "set_option: unknown option '" + k + "'"; return 1; }
    return 0;
}

int nori_gpu_upload_scene(nori_gpu_ctx *ctx, const nori_gpu_scene *s) {
    REQUIRE(ctx && s, "upload_scene: null argument");
    for (nori_gpu_ctx *p : ctx->peers)                       // the scene is replicated on every device
        if (nori_gpu_upload_scene(p, s)) { ctx->err = p->err; return 1; }
    REQUIRE(s->abi_version == NORI_GPU_ABI_VERSION, "upload_scene: ABI version mismatch");
    REQUIRE(s->integrator >= 0 && s->integrator <= NORI_INTEGRATOR_VOLUMETRIC, "upload_scene: unknown integrator");
    REQUIRE(s->camera.width > 0 && s->camera.height > 0, "upload_scene: empty film");
    REQUIRE(s->camera.type >= 0 && s->camera.type <= NORI_CAMERA_ADVANCED, "upload_scene: unknown camera type");
    REQUIRE(s->n_images == 0 || s->images, "upload_scene: null image table");
    REQUIRE(s->filter.radius > 0.f && s->filter.radius <= 8.f, "upload_scene: filter radius must be in (0, 8]");
    REQUIRE(s->n_shapes == 0 || (s->nodes && s->indices && s->shape_offset && s->shapes && s->bsdfs), "upload_scene: null scene arrays");
    REQUIRE(s->integrator != NORI_INTEGRATOR_VOLUMETRIC || s->medium.present, "upload_scene: volumetric integrator needs a medium");
    bool needLights = s->integrator == NORI_INTEGRATOR_PATH_MIS || s->integrator == NORI_INTEGRATOR_VOLUMETRIC;
    REQUIRE(!needLights || s->n_emitters > 0, "upload_scene: this integrator needs at least one emitter (Scene::getRandomEmitter)");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    freeAll(ctx->scene_allocs);
    ctx->has_scene = false;
    const size_t need = sceneArenaBytes(s);
    if (need > ctx->arena_cap) {
        cudaFree(ctx->arena); ctx->arena = nullptr; ctx->arena_cap = 0;
        CK(cudaMalloc((void **) &ctx->arena, need));
        ctx->arena_cap = need;
    }
    ctx->arena_used = 0;
    DScene ds{};
    ds.n_nodes = s->n_nodes; ds.n_prims = s->n_indices; ds.n_shapes = s->n_shapes; ds.n_emitters = s->n_emitters;
    ds.ordered = 0;
    ds.integrator = s->integrator; ds.av_length = s->av_length; ds.camera = s->camera; ds.medium = s->medium;

    // ---- image textures / normal maps
    std::vector<DImage> images(s->n_images);
    for (uint32_t i = 0; i < s->n_images; ++i) {
        const nori_gpu_image &h = s->images[i];
        REQUIRE(h.width > 0 && h.height > 0 && h.rgb, "upload_scene: empty image");
        REQUIRE(h.wrap == NORI_WRAP_REPEAT || h.wrap == NORI_WRAP_CLAMP, "upload_scene: unknown image wrap mode");
        images[i].width = h.width; images[i].height = h.height; images[i].wrap = h.wrap; images[i].pad = 0;
        if (devUpload(ctx, ctx->scene_allocs, h.rgb, (size_t) h.width * h.height * 3, &images[i].rgb)) return 1;
    }
    for (uint32_t i = 0; i < s->n_bsdfs; ++i) {
        const nori_gpu_bsdf &b = s->bsdfs[i];
        REQUIRE(b.albedo_texture >= NORI_TEXTURE_CONSTANT && b.albedo_texture <= NORI_TEXTURE_IMAGE, "upload_scene: unknown texture type");
        REQUIRE(b.albedo_texture != NORI_TEXTURE_IMAGE || (b.albedo_image >= 0 && (uint32_t) b.albedo_image < s->n_images),
                "upload_scene: BSDF references a missing image");
    }
    // ---- per-shape arrays + shape table
    std::vector<DShape> shapes(s->n_shapes);
    uint32_t mask = 0;
    bool hasPerlin = false;
    for (uint32_t i = 0; i < s->n_shapes; ++i) {
        const nori_gpu_shape &h = s->shapes[i];
        DShape &d = shapes[i]; memset(&d, 0, sizeof(d));
        REQUIRE(h.bsdf >= 0 && (uint32_t) h.bsdf < s->n_bsdfs, "upload_scene: shape references a missing BSDF");
        REQUIRE(h.emitter < (int32_t) s->n_emitters, "upload_scene: shape references a missing emitter");
        d.type = h.type; d.bsdf = h.bsdf; d.emitter = h.emitter; d.bsdf_type = s->bsdfs[h.bsdf].type;
        REQUIRE(d.bsdf_type >= 0 && d.bsdf_type < NORI_BSDF_COUNT, "upload_scene: unknown BSDF type");
        mask |= 1u << d.bsdf_type;
        REQUIRE(h.normal_map >= 0 && (uint32_t) h.normal_map <= s->n_images, "upload_scene: shape references a missing normal map");
        d.normal_map = h.normal_map;
        d.n_triangles = h.n_triangles; d.area_normalization = h.area_normalization;
        d.cx = h.center[0]; d.cy = h.center[1]; d.cz = h.center[2]; d.radius = h.radius;
        if (h.type == NORI_SHAPE_MESH) {
            REQUIRE(h.V && h.F, "upload_scene: mesh without vertices/faces");
            if (devUpload(ctx, ctx->scene_allocs, h.V, 3 * (size_t) h.n_vertices, &d.V)) return 1;
            if (devUpload(ctx, ctx->scene_allocs, h.N, h.N ? 3 * (size_t) h.n_vertices : 0, &d.N)) return 1;
            if (devUpload(ctx, ctx->scene_allocs, h.UV, h.UV ? 2 * (size_t) h.n_vertices : 0, &d.UV)) return 1;
            if (devUpload(ctx, ctx->scene_allocs, h.F, 3 * (size_t) h.n_triangles, &d.F)) return 1;
            if (devUpload(ctx, ctx->scene_allocs, h.area_cdf, h.area_cdf ? (size_t) h.n_triangles + 1 : 0, &d.cdf)) return 1;
            d.has_n = h.N != nullptr; d.has_uv = h.UV != nullptr;
            REQUIRE(h.emitter < 0 || h.area_cdf, "upload_scene: emitter mesh without an area CDF");
        } else {
            REQUIRE(h.type == NORI_SHAPE_SPHERE || h.type == NORI_SHAPE_PERLIN, "upload_scene: unknown shape type");
            REQUIRE(h.type != NORI_SHAPE_PERLIN || h.perlin_height != 0.f, "upload_scene: perlin sphere with height 0");
            d.perlin_height = h.perlin_height; d.perlin_scale = h.perlin_scale;
            hasPerlin = hasPerlin || h.type == NORI_SHAPE_PERLIN;
            // sphere.cpp:99: std::pow(1.f / r, 2) [double] * 0.25f * INV_PI [float]
            double ir = (double) (1.f / h.radius);
            d.sphere_pdf = (float) (ir * ir * (double) (0.25f * NORI_INV_PI));
        }
    }
    // ---- primitive records in leaf order (48 B each): same arithmetic as mesh.cpp:88 for the edges
    std::vector<float4> prims(3 * (size_t) s->n_indices);
    for (uint32_t i = 0; i < s->n_indices; ++i) {
        uint32_t idx = s->indices[i];
        REQUIRE(idx < s->shape_offset[s->n_shapes], "upload_scene: primitive index out of range");
        const uint32_t *it = std::lower_bound(s->shape_offset, s->shape_offset + s->n_shapes + 1, idx + 1) - 1;   // bvh.h:105-109
        uint32_t shape = (uint32_t) (it - s->shape_offset); idx -= *it;
        const nori_gpu_shape &h = s->shapes[shape];
        float4 *r = &prims[3 * (size_t) i];
        uint32_t tag;
        if (h.type == NORI_SHAPE_MESH) {
            REQUIRE(idx < h.n_triangles, "upload_scene: triangle index out of range");
            const uint32_t i0 = h.F[3 * idx], i1 = h.F[3 * idx + 1], i2 = h.F[3 * idx + 2];
            REQUIRE(i0 < h.n_vertices && i1 < h.n_vertices && i2 < h.n_vertices, "upload_scene: vertex index out of range");
            const float *p0 = &h.V[3 * i0], *p1 = &h.V[3 * i1], *p2 = &h.V[3 * i2];
            r[0] = make_float4(p0[0], p0[1], p0[2], 0.f);
            r[1] = make_float4(p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2], 0.f);
            r[2] = make_float4(p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2], 0.f);
            tag = 0u;
        } else {
            r[0] = make_float4(h.center[0], h.center[1], h.center[2], 0.f);
            r[1] = make_float4(h.radius, h.perlin_height, h.perlin_scale, 0.f);
            r[2] = make_float4(0.f, 0.f, 0.f, 0.f);
            tag = h.type == NORI_SHAPE_PERLIN ? 2u : 1u;
        }
        memcpy(&r[0].w, &idx, 4); memcpy(&r[1].w, &shape, 4); memcpy(&r[2].w, &tag, 4);
    }
    // ---- emitters
    std::vector<DEmitter> ems(s->n_emitters);
    uint32_t emitterTypeMask = 0;
    for (uint32_t i = 0; i < s->n_emitters; ++i) {
        ems[i].pod = s->emitters[i];
        if (s->emitters[i].type >= 0 && s->emitters[i].type < 4) emitterTypeMask |= 1u << s->emitters[i].type;
        nori_gpu_emitter &e = ems[i].pod;
        REQUIRE(e.type >= 0 && e.type <= NORI_EMITTER_ENVMAP, "upload_scene: unknown emitter type");
        if (e.type == NORI_EMITTER_AREA || e.type == NORI_EMITTER_ENVMAP)
            REQUIRE(e.shape >= 0 && (uint32_t) e.shape < s->n_shapes, "upload_scene: area/envmap emitter without a shape");
        if (e.type == NORI_EMITTER_ENVMAP) {
            const nori_gpu_emitter &h = s->emitters[i];
            REQUIRE(h.env_rows > 1 && h.env_cols > 1 && h.env_image && h.env_pdf && h.env_cdf && h.env_pmarginal && h.env_cmarginal,
                    "upload_scene: incomplete environment map tables");
            size_t R = h.env_rows, C = h.env_cols;
            if (devUpload(ctx, ctx->scene_allocs, h.env_image, R * C * 3, &e.env_image)) return 1;
            if (devUpload(ctx, ctx->scene_allocs, h.env_pdf, R * C, &e.env_pdf)) return 1;
            if (devUpload(ctx, ctx->scene_allocs, h.env_cdf, R * (C + 1), &e.env_cdf)) return 1;
            if (devUpload(ctx, ctx->scene_allocs, h.env_pmarginal, R, &e.env_pmarginal)) return 1;
            if (devUpload(ctx, ctx->scene_allocs, h.env_cmarginal, R + 1, &e.env_cmarginal)) return 1;
        } else { e.env_image = e.env_pdf = e.env_cdf = e.env_pmarginal = e.env_cmarginal = nullptr; }
    }
    // ---- the tree must be one the traversal can walk: children and leaf ranges in range, and no deeper than the
    // 64-entry traversal stack (the reference's own limit, bvh.cpp:405)
    if (s->n_nodes) {
        const uint32_t *w = (const uint32_t *) s->nodes;
        std::vector<std::pair<uint32_t, uint32_t>> st; st.reserve(128);
        st.push_back({0u, 1u});
        uint32_t maxDepth = 0; uint64_t visited = 0;
        while (!st.empty()) {
            const uint32_t i = st.back().first, depth = st.back().second; st.pop_back();
            ++visited;
            REQUIRE(visited <= s->n_nodes, "upload_scene: BVH nodes do not form a tree");
            maxDepth = std::max(maxDepth, depth);
            const uint32_t w0 = w[8 * (size_t) i], w1 = w[8 * (size_t) i + 1];
            if (w0 & 1u) REQUIRE((uint64_t) w1 + (w0 >> 1) <= s->n_indices, "upload_scene: BVH leaf range out of bounds");
            else {
                REQUIRE(i + 1 < s->n_nodes && w1 > i && w1 < s->n_nodes, "upload_scene: BVH child index out of range");
                st.push_back({w1, depth + 1}); st.push_back({i + 1, depth + 1});
            }
        }
        REQUIRE(maxDepth <= 64, "upload_scene: BVH deeper than the 64-entry traversal stack (bvh.cpp:405)");
    }
    // ---- layouts for the large-scene kernels (wave_extend.cu), derived from the reference nodes on the host
    // (host_layout.h): child-box pairs and 4-wide records.  Built only when every leaf fits the reference encoding.
    std::vector<uint32_t> nodes2, nodes4;
    if (noriBuildPairLayout((const uint32_t *) s->nodes, s->n_nodes, s->n_indices, nodes2, ds.root_ref)) {
        const uint32_t *w = (const uint32_t *) s->nodes;
        memcpy(ds.root_min, &w[2], 12); memcpy(ds.root_max, &w[5], 12);
        if (noriBuildWideLayout(w, s->n_nodes, s->n_indices, NORI_STACK2_MAX, nodes4)) ds.root_ref4 = 0u;   // record 0
    }
    if (devUpload(ctx, ctx->scene_allocs, (const uint4 *) nodes2.data(), nodes2.size() / 4, &ds.nodes2)) return 1;
    if (devUpload(ctx, ctx->scene_allocs, (const uint4 *) nodes4.data(), nodes4.size() / 4, &ds.nodes4)) return 1;
    ctx->nodes4_bytes = nodes4.size() * sizeof(uint32_t);
    static_assert(sizeof(nori_gpu_bvh_node) == 2 * sizeof(uint4), "node layout");
    if (devUpload(ctx, ctx->scene_allocs, (const uint4 *) s->nodes, 2 * (size_t) s->n_nodes, &ds.nodes)) return 1;
    if (devUpload(ctx, ctx->scene_allocs, prims.data(), prims.size(), &ds.prims)) return 1;
    if (devUpload(ctx, ctx->scene_allocs, shapes.data(), shapes.size(), &ds.shapes)) return 1;
    if (devUpload(ctx, ctx->scene_allocs, s->bsdfs, (size_t) s->n_bsdfs, &ds.bsdfs)) return 1;
    if (devUpload(ctx, ctx->scene_allocs, ems.data(), ems.size(), &ds.emitters)) return 1;
    if (devUpload(ctx, ctx->scene_allocs, images.data(), images.size(), &ds.images)) return 1;
    CK(cudaStreamSynchronize(ctx->stream));         // host staging vectors die at return

    ctx->ds = ds; ctx->filter = s->filter; ctx->bsdf_mask = mask; ctx->n_bsdfs = s->n_bsdfs;
    ctx->W = s->camera.width; ctx->H = s->camera.height;
    ctx->border = (int) std::ceil(s->filter.radius - 0.5f);              // block.cpp:57
    if (ctx->vsum) { cudaFree(ctx->vsum); cudaFree(ctx->vsum2); ctx->vsum = ctx->vsum2 = nullptr; }
    ctx->opt_variance = 0; ctx->var_passes = 0;
    size_t nf = (size_t) (ctx->W + 2 * ctx->border) * (ctx->H + 2 * ctx->border);
    if (nf > ctx->film_cap) {
        cudaFree(ctx->film); ctx->film = nullptr; ctx->film_cap = 0;
        CK(cudaMalloc((void **) &ctx->film, nf * sizeof(float4)));
        ctx->film_cap = nf;
    }
    CK(cudaMemsetAsync(ctx->film, 0, nf * sizeof(float4), ctx->stream));
    ctx->has_scene = true; ctx->has_perlin = hasPerlin; ctx->n_emitter_types = (uint32_t) __builtin_popcount(emitterTypeMask);
    ctx->has_envmap = (emitterTypeMask >> NORI_EMITTER_ENVMAP) & 1u; ctx->emitter_type_mask = emitterTypeMask;
    return 0;
}

} // extern "C"

static int ensurePool(nori_gpu_ctx *ctx, bool deferShadow, bool esort) {
    const nori_gpu_camera &cam = ctx->ds.camera;
    const bool chroma = cam.type == NORI_CAMERA_ADVANCED && !(cam.chromatic[0] == 0.f && cam.chromatic[1] == 0.f && cam.chromatic[2] == 0.f);
    if (ctx->pool.P == (uint32_t) ctx->opt_pool && !ctx->pool_allocs.empty() && (!deferShadow || ctx->pool.shD) && (!chroma || ctx->pool.acc)
        && (!esort || ctx->pool.equeue)) return 0;
    freeAll(ctx->pool_allocs);
    Pool p{}; p.P = (uint32_t) ctx->opt_pool;
    auto alloc = [&](size_t bytes) -> void * { void *d = nullptr; if (cudaMalloc(&d, bytes) != cudaSuccess) return nullptr; ctx->pool_allocs.push_back(d); return d; };
    float4 **f4[] = {&p.rayO, &p.rayD, &p.hit, &p.thr, &p.rad, &p.shD, &p.shC, &p.acc};
    for (auto pp : f4) {
        if (!deferShadow && (pp == &p.shD || pp == &p.shC)) continue;
        if (!chroma && pp == &p.acc) continue; *pp = (float4 *) alloc(p.P * sizeof(float4)); REQUIRE(*pp, "out of device memory (pool)"); }
    p.rng = (uint64_t *) alloc(p.P * sizeof(uint64_t)); p.sid = (uint32_t *) alloc(p.P * 4); p.flags = (uint32_t *) alloc(p.P * 4);
    REQUIRE(p.rng && p.sid && p.flags, "out of device memory (pool)");
    if (esort) { p.equeue = (uint32_t *) alloc((size_t) NORI_NEQ * p.P * 4); REQUIRE(p.equeue, "out of device memory (emitter-sorted queues)"); }
    for (int t = 0; t < NORI_NQ; ++t) { p.queue[t] = (uint32_t *) alloc(p.P * 4); REQUIRE(p.queue[t], "out of device memory (queues)"); }
    k_fill_u32<<<(p.P + 255) / 256, 256, 0, ctx->stream>>>(p.sid, NORI_FREE_SLOT, p.P);
    CK(cudaMemsetAsync(p.flags, 0, p.P * 4, ctx->stream));
    CK(cudaGetLastError());
    ctx->pool = p;
    return 0;
}

static int ensureResults(nori_gpu_ctx *ctx, size_t n) {
    if (n <= ctx->results_cap) return 0;
    cudaFree(ctx->results); ctx->results = nullptr; ctx->results_cap = 0;
    CK(cudaMalloc((void **) &ctx->results, n * sizeof(float4)));
    ctx->results_cap = n;
    return 0;
}

static const WaveKernels kWave = {noriPickExtend, noriLaunchShadeMats, noriLaunchShadeMisDeferred, noriLaunchShadowSm, noriShadowSmOccupancy,
                                  noriLaunchShadeMis, noriLaunchShadeVol, noriLaunchRebin, noriLaunchDrain};
static const WaveKernels kWavePerlin = {noriPickExtendPerlin, noriLaunchShadeMatsPerlin, noriLaunchShadeMisDeferredPerlin, noriLaunchShadowSmPerlin,
                                        noriShadowSmOccupancyPerlin, noriLaunchShadeMisPerlin, noriLaunchShadeVolPerlin, noriLaunchRebinPerlin,
                                        noriLaunchDrainPerlin};

// Tensor map of the sample buffer for the film kernel's TMA loads: a {4 W, H, layers} float tensor read in boxes of
// {4 S, S, 1} (one layer of a (32 + 2 halo)^2 pixel tile), zeros outside.  cuTensorMapEncodeTiled is a driver entry point;
// the library links the runtime only and fetches it here.  Returns false when the driver does not offer it.
static bool filmTensorMap(CUtensorMap *map, const float4 *results, int W, int H, uint32_t layers, int S) {
    typedef CUresult (*Encode)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                               const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static Encode encode = nullptr; static bool looked = false;
    if (!looked) {
        looked = true;
        void *fn = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) encode = (Encode) fn;
        else cudaGetLastError();
    }
    if (!encode || 4 * S > 256) return false;
    const cuuint64_t dims[3] = {(cuuint64_t) 4 * W, (cuuint64_t) H, (cuuint64_t) layers};
    const cuuint64_t strides[2] = {(cuuint64_t) W * 16, (cuuint64_t) W * H * 16};
    const cuuint32_t box[3] = {(cuuint32_t) (4 * S), (cuuint32_t) S, 1u}, estr[3] = {1u, 1u, 1u};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *) results, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Trace all camera paths of one batch; on return (stream-ordered) bt.results holds every sample.
static int traceBatch(nori_gpu_ctx *ctx, const Batch &bt, uint32_t nLayers) {
    const unsigned long long total = (unsigned long long) nLayers * bt.wh;
    const bool count = ctx->opt_stats != 0;
    const int integ = ctx->ds.integrator;
    ctx->wf_used = 1;                                       // (a batch that failed half-way must not leave its wavefronts to the next foldStats)
    ctx->ds.ordered = ctx->opt_order == 1 || (ctx->opt_order == 2 && ctx->ds.n_prims > 4096);
    ctx->ds.wide = ctx->opt_wide ? 1 : 0;
    const bool wave = (integ == NORI_INTEGRATOR_PATH_MIS || integ == NORI_INTEGRATOR_PATH_MATS || integ == NORI_INTEGRATOR_VOLUMETRIC)
                      && !ctx->opt_megakernel;
    // scenes with a Perlin-noise sphere run the Perlin-aware set of the same kernels (kernels.cuh)
    const WaveKernels &wk = ctx->has_perlin ? kWavePerlin : kWave;
    if (!wave) {
        const unsigned grid = (unsigned) ((total + 127) / 128);
        LAUNCH(NORI_K_SINGLE, noriLaunchMega(count, grid, ctx->stream, ctx->ds, bt, ctx->ctr, total));
        CK(cudaGetLastError());
        ctx->stats.iterations += 1; ctx->last_wave = false;
        return 0;
    }
    // plain per-lane loops for tiny scenes, the warp state machine once trees are deep (see wave_extend.cu)
    const bool sm = ctx->opt_traversal == 2 || (ctx->opt_traversal == 0 && ctx->ds.n_prims > 4096);
    const int mode = integ == NORI_INTEGRATOR_PATH_MIS ? MODE_MIS : integ == NORI_INTEGRATOR_PATH_MATS ? MODE_MATS : MODE_VOL;
    // NEE shadow rays: traced inside k_shade on small scenes, by their own state-machine pass on deep trees
    const bool defer = mode == MODE_MIS && (ctx->opt_shadow_pass == 1 || (ctx->opt_shadow_pass == 0 && sm));
    // emitter-sorted queues pay off when lanes would otherwise diverge between a cheap and an expensive light
    // (environment map: two binary searches + trigonometry); option 2 forces them for any mix of types
    const bool esort = mode == MODE_MIS && ctx->n_emitter_types > 1 && (ctx->opt_emitter_sort == 2 || (ctx->opt_emitter_sort == 1 && ctx->has_envmap));
    ctx->ds.esort = esort ? 1 : 0;
    ctx->ds.area_only = (ctx->opt_area_only && ctx->emitter_type_mask == (1u << NORI_EMITTER_AREA)) ? 1 : 0;
    if (ensurePool(ctx, defer, esort)) return 1;
    // L2 access-policy window over the head of the 4-wide records (numbered breadth-first from the root, host_bvh.cpp): the
    // part of a large tree that most rays walk stays resident while leaf-level records and primitives stream through
    {
        const bool want = ctx->opt_l2_window > 0 && sm && noriSmLayout(ctx->ds) == 2 && ctx->nodes4_bytes > 0;
        if (want || ctx->window_set) {
            cudaStreamAttrValue av{};
            if (want) {
                int maxWin = 0, maxPersist = 0;
                cudaDeviceGetAttribute(&maxWin, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
                cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
                size_t bytes = std::min<size_t>((size_t) ctx->opt_l2_window << 20, ctx->nodes4_bytes);
                bytes = std::min<size_t>(bytes, (size_t) std::max(maxWin, 0));
                cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min<size_t>(bytes, (size_t) std::max(maxPersist, 0)));
                av.accessPolicyWindow.base_ptr = (void *) ctx->ds.nodes4;
                av.accessPolicyWindow.num_bytes = bytes;
                av.accessPolicyWindow.hitRatio = maxPersist > 0 ? std::min(1.0f, (float) maxPersist / (float) std::max<size_t>(bytes, 1)) : 0.f;
                av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            } else { av.accessPolicyWindow.num_bytes = 0; cudaCtxResetPersistingL2Cache(); }
            cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &av);
            cudaGetLastError();                              // a device without the feature renders as before
            ctx->window_set = want;
        }
    }
    // ---- concurrent wavefronts.  Every kernel launch of the loop below ends with warps finishing unevenly and is followed
    // by a launch gap and a ramp-up, ~37 us per iteration in which most of the device idles.  The batch's sample layers are
    // therefore split between up to `wavefronts` independent wavefronts -- each with its own part of the path pool and of the
    // queues, its own counters, its own stream, and a whole number of layers (its Batch differs in `results`, `spp_first`
    // only, so the kernels are unchanged and every sample is computed exactly as before) -- and the block scheduler fills the
    // tail of one wavefront's kernel with the CTAs of the other's.  Per-launch event timing needs the launches serialised:
    // `kernel_timing` renders with one wavefront.
    int W = (int) std::min<int64_t>(std::min<int64_t>(ctx->opt_wavefronts, NORI_MAX_WAVEFRONTS), nLayers);
    if (ctx->opt_kernel_timing) W = 1;
    while (W > 1 && ((ctx->pool.P / (uint32_t) W) & ~255u) < 4096u) --W;
    struct Wavefront { Pool pool; Batch bt; Counters *ctr, *h; cudaStream_t st; unsigned long long total; uint32_t it; bool finished; };
    Wavefront wf[NORI_MAX_WAVEFRONTS];
    for (int j = 0; j < W; ++j) {
        Wavefront &w = wf[j];
        if (j > 0 && !ctx->wf_stream[j - 1]) {
            CK(cudaStreamCreateWithFlags(&ctx->wf_stream[j - 1], cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&ctx->wf_done[j - 1], cudaEventDisableTiming));
            CK(cudaMalloc((void **) &ctx->wf_ctr[j - 1], sizeof(Counters)));
            CK(cudaMallocHost((void **) &ctx->wf_h_ctr[j - 1], sizeof(Counters)));
        }
        w.st = j ? ctx->wf_stream[j - 1] : ctx->stream; w.ctr = j ? ctx->wf_ctr[j - 1] : ctx->ctr; w.h = j ? ctx->wf_h_ctr[j - 1] : ctx->h_ctr;
        const uint32_t first = (uint32_t) ((unsigned long long) nLayers * j / W), last = (uint32_t) ((unsigned long long) nLayers * (j + 1) / W);
        w.bt = bt; w.bt.results = bt.results + (size_t) first * bt.wh; w.bt.spp_first = bt.spp_first + first;
        w.bt.capacity = bt.capacity - (uint32_t) std::min<size_t>((size_t) first * bt.wh, bt.capacity);
        w.total = (unsigned long long) (last - first) * bt.wh; w.it = 0; w.finished = false;
        w.pool = ctx->pool;
        if (W > 1) {                                         // this wavefront's slice of every pool array
            const uint32_t Pj = (ctx->pool.P / (uint32_t) W) & ~255u;
            const size_t off = (size_t) j * Pj;
            Pool &q = w.pool; q.P = Pj;
            float4 **f4[] = {&q.rayO, &q.rayD, &q.hit, &q.thr, &q.rad, &q.acc, &q.shD, &q.shC};
            for (auto pp : f4) if (*pp) *pp += off;
            q.rng += off; q.sid += off; q.flags += off;
            for (int t = 0; t < NORI_NQ; ++t) q.queue[t] += off;
            if (q.equeue) q.equeue += (size_t) NORI_NEQ * off;
        }
        // per-batch counters (the cumulative ones are folded into ctx->stats by the caller)
        Counters zero{}; zero.total_samples = w.total;
        *w.h = zero;
        CK(cudaMemcpyAsync(w.ctr, w.h, sizeof(Counters), cudaMemcpyHostToDevice, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));                 // everything queued before this batch has finished: the other streams may start
    ctx->wf_used = W;
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const ExtendKernel kext = wk.pickExtend(sm, count, mode == MODE_VOL, noriSmLayout(ctx->ds));
    int occE = 8;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occE, kext, 128, 0);
    const int gridE = sms * std::max(1, occE), gridSh = sms * NORI_SHADE_GRID_PER_SM;
    const int gridShadow = defer ? sms * std::max(1, wk.shadowSmOccupancy(count, noriSmLayout(ctx->ds))) : 0;
    ctx->last_wave = true; ctx->last_defer = defer;
    // `poll` iterations of one wavefront, then its counters on their way to the host
    auto enqueue = [&](Wavefront &w) -> int {
        for (int i = 0; i < ctx->opt_poll; ++i, ++w.it) {
            LAUNCH(NORI_K_EXTEND, (kext<<<gridE, 128, 0, w.st>>>(ctx->ds, w.pool, w.bt, w.ctr, w.it)));
            if (esort) LAUNCH(NORI_K_GENERATE, wk.rebin((int) ((w.pool.P + 1023u) / 1024u), w.st, ctx->ds, w.pool, w.bt, w.ctr, w.it));
            if (defer) {
                LAUNCH(NORI_K_SHADE, wk.shadeMisDeferred(count, gridSh, w.st, ctx->ds, w.pool, w.bt, w.ctr, w.it));
                LAUNCH(NORI_K_SHADOW, wk.shadowSm(count, gridShadow, w.st, ctx->ds, w.pool, w.bt, w.ctr, w.it));
            } else if (mode == MODE_MIS) LAUNCH(NORI_K_SHADE, wk.shadeMis(count, gridSh, w.st, ctx->ds, w.pool, w.bt, w.ctr, w.it));
            else if (mode == MODE_MATS) LAUNCH(NORI_K_SHADE, wk.shadeMats(count, gridSh, w.st, ctx->ds, w.pool, w.bt, w.ctr, w.it));
            else LAUNCH(NORI_K_SHADE, wk.shadeVol(count, gridSh, w.st, ctx->ds, w.pool, w.bt, w.ctr, w.it));
            ctx->stats.iterations += 1;
        }
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(w.h, w.ctr, sizeof(Counters), cudaMemcpyDeviceToHost, w.st));
        return 0;
    };
    for (int j = 0; j < W; ++j) if (enqueue(wf[j])) return 1;
    // wait for one wavefront's counters while the others' launches keep the device busy; refill it before waiting for the next
    for (int left = W; left > 0;) {
        for (int j = 0; j < W; ++j) {
            Wavefront &w = wf[j];
            if (w.finished) continue;
            CK(cudaStreamSynchronize(w.st));
            if (w.h->done >= w.total) { w.finished = true; --left; continue; }
            // drain: no camera path left to start and only a few paths alive => finish them in one launch (mega.cu)
            const unsigned long long live = w.total - w.h->done;
            if (mode != MODE_VOL && ctx->opt_drain > 0 && w.h->next_sample >= w.total && live <= (unsigned long long) ctx->opt_drain) {
                // warp mode: 16 CTAs per SM = 64 warps, each scanning its share of the pool in 32-slot segments
                LAUNCH(NORI_K_SINGLE, wk.drain(mode == MODE_MIS, count, ctx->opt_drain_mode == 0 ? -(sms * 16) : sms * 8, w.st, ctx->ds, w.pool, w.bt, w.ctr));
                CK(cudaGetLastError());
                ctx->stats.iterations += 1;
                w.finished = true; --left;
                continue;
            }
            if (enqueue(w)) return 1;
        }
    }
    for (int j = 1; j < W; ++j) {                           // what follows on the context's stream (film pass, counters) waits for every wavefront
        CK(cudaEventRecord(ctx->wf_done[j - 1], wf[j].st));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->wf_done[j - 1], 0));
    }
    return 0;
}

static int foldStats(nori_gpu_ctx *ctx, unsigned long long samples) {
    CK(cudaMemcpyAsync(ctx->h_ctr, ctx->ctr, sizeof(Counters), cudaMemcpyDeviceToHost, ctx->stream));
    for (int j = 1; j < ctx->wf_used; ++j) CK(cudaMemcpyAsync(ctx->wf_h_ctr[j - 1], ctx->wf_ctr[j - 1], sizeof(Counters), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    Counters c = *ctx->h_ctr;
    for (int j = 1; j < ctx->wf_used; ++j) {                 // the other wavefronts of the batch (traceBatch)
        const Counters &o = *ctx->wf_h_ctr[j - 1];
        c.rays_ext += o.rays_ext; c.rays_sh += o.rays_sh; c.rays_sh_closest += o.rays_sh_closest; c.nodes_ext += o.nodes_ext; c.nodes_sh += o.nodes_sh;
        c.prims_ext += o.prims_ext; c.prims_sh += o.prims_sh; c.invalid += o.invalid; c.guard_redo += o.guard_redo; c.max_stack = std::max(c.max_stack, o.max_stack);
    }
    ctx->wf_used = 1;
    ctx->stats.samples += samples;
    ctx->stats.rays += c.rays_ext + c.rays_sh + c.rays_sh_closest; ctx->stats.shadow_rays += c.rays_sh;
    ctx->stats.nodes_visited += c.nodes_ext + c.nodes_sh; ctx->stats.prims_tested += c.prims_ext + c.prims_sh;
    ctx->stats.invalid_samples += c.invalid;
    ctx->stats.max_stack_depth = std::max<uint64_t>(ctx->stats.max_stack_depth, c.max_stack);
    ctx->stats.guard_retraces += c.guard_redo;
    if (ctx->last_wave) {
        // shadow rays are traced inside k_shade unless the deferred pass ran
        nori_gpu_kernel_stats &e = ctx->kstats[NORI_K_EXTEND], &s = ctx->kstats[ctx->last_defer ? NORI_K_SHADOW : NORI_K_SHADE];
        e.rays += c.rays_ext; e.nodes_visited += c.nodes_ext; e.prims_tested += c.prims_ext;
        s.rays += c.rays_sh + c.rays_sh_closest; s.nodes_visited += c.nodes_sh; s.prims_tested += c.prims_sh;
    } else {
        nori_gpu_kernel_stats &m = ctx->kstats[NORI_K_SINGLE];
        m.rays += c.rays_ext + c.rays_sh; m.nodes_visited += c.nodes_ext + c.nodes_sh; m.prims_tested += c.prims_ext + c.prims_sh;
    }
    for (size_t i = 0; i < ctx->kev_used; i += 2) {          // the stream is idle here: every event has completed
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->kev[i], ctx->kev[i + 1]) == cudaSuccess) ctx->kstats[ctx->kev_kind[i / 2]].ms += ms;
    }
    ctx->kev_used = 0; ctx->kev_kind.clear();
    CK(cudaMemsetAsync(ctx->ctr, 0, sizeof(Counters), ctx->stream));
    return 0;
}

static int renderImpl(nori_gpu_ctx *ctx, uint32_t spp_begin, uint32_t spp_count, uint64_t seed, float *out_rgba) {
    REQUIRE(ctx && ctx->has_scene, "render: no scene uploaded");
    CK(cudaSetDevice(ctx->device));
    if (spp_count == 0) return 0;
    const uint32_t wh = (uint32_t) ctx->W * ctx->H;
    REQUIRE((unsigned long long) wh * 1 < 0xffffffffull, "render: image too large");
    size_t maxLayers = std::max<size_t>(1, ((size_t) ctx->opt_results_mb << 20) / ((size_t) wh * sizeof(float4)));
    maxLayers = std::min<size_t>(maxLayers, 0xfffffff0ull / wh);       // sample ids are 32-bit inside a batch
    REQUIRE(maxLayers >= 1, "render: image too large for a single-layer batch");
    {   // the per-sample buffer takes what the device can spare: at most results_mb, at most 3/4 of the free memory (plus
        // what the buffer already holds), and half as many layers again whenever the allocation still fails
        size_t freeB = 0, totalB = 0;
        const bool fits = std::min<size_t>(maxLayers, spp_count) * wh <= ctx->results_cap;   // (the query costs up to a millisecond: only when the buffer must grow)
        if (!fits && cudaMemGetInfo(&freeB, &totalB) == cudaSuccess) {
            const size_t spare = freeB / 4 * 3 + ctx->results_cap * sizeof(float4);
            maxLayers = std::max<size_t>(1, std::min<size_t>(maxLayers, spare / ((size_t) wh * sizeof(float4))));
        }
        maxLayers = std::min<size_t>(maxLayers, spp_count);
        while (ensureResults(ctx, maxLayers * wh)) {
            if (maxLayers == 1) return 1;                   // not even one layer fits: the error of ensureResults stands
            cudaGetLastError(); ctx->err.clear();
            maxLayers = (maxLayers + 1) / 2;
        }
    }
    CK(cudaMemsetAsync(ctx->ctr, 0, sizeof(Counters), ctx->stream));
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    FilmParams fp{};
    fp.film = ctx->film; fp.W = ctx->W; fp.H = ctx->H; fp.border = ctx->border; fp.halo = ctx->border;
    fp.radius = ctx->filter.radius; fp.lookupFactor = NORI_FILTER_RESOLUTION / ctx->filter.radius;      // block.cpp:64
    memcpy(fp.table, ctx->filter.table, sizeof(fp.table));
    for (uint32_t done = 0; done < spp_count;) {
        uint32_t n = (uint32_t) std::min<size_t>(maxLayers, spp_count - done);
        Batch bt{}; bt.results = ctx->results; bt.seed = seed; bt.spp_first = spp_begin + done; bt.wh = wh; bt.capacity = (uint32_t) std::min<size_t>(ctx->results_cap, 0xffffffffu);
        if (traceBatch(ctx, bt, n)) return 1;
        if (out_rgba) {
            CK(cudaMemcpyAsync(out_rgba + (size_t) done * wh * 4, ctx->results, (size_t) n * wh * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
        } else {
            const int S = 32 + 2 * fp.halo;
            dim3 grid((ctx->W + 2 * ctx->border + 31) / 32, (ctx->H + 2 * ctx->border + 31) / 32);
            if (ctx->opt_variance) { fp.vsum = ctx->vsum; fp.vsum2 = ctx->vsum2; ctx->var_passes += n; }
            if (fp.halo == 2 && ctx->opt_film_sep) {                 // default Gaussian (radius 2): separable weights tabulated per sample
                CUtensorMap tmap{};
                const bool tma = ctx->opt_film_tma && filmTensorMap(&tmap, ctx->results, ctx->W, ctx->H, n, S);
                const size_t smem = (size_t) S * S * ((tma ? 2 : 1) * sizeof(float4) + 2 * 5 * sizeof(float));
                const dim3 blk(32, 32 / NORI_FILM_NPIX);
#define NORI_FILM_LAUNCH(VAR, TMA) do { \
                    CK(cudaFuncSetAttribute(k_film_sep<VAR, 2, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); \
                    LAUNCH(NORI_K_FILM, (k_film_sep<VAR, 2, TMA><<<grid, blk, smem, ctx->stream>>>(fp, bt, n, tmap))); } while (0)
                if (ctx->opt_variance) { if (tma) NORI_FILM_LAUNCH(true, true); else NORI_FILM_LAUNCH(true, false); }
                else { if (tma) NORI_FILM_LAUNCH(false, true); else NORI_FILM_LAUNCH(false, false); }
#undef NORI_FILM_LAUNCH
            } else {
                size_t smem = (size_t) S * S * (sizeof(float4) + sizeof(float2));
                if (ctx->opt_variance) {
                    CK(cudaFuncSetAttribute(k_film<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
                    LAUNCH(NORI_K_FILM, (k_film<true><<<grid, dim3(32, 32), smem, ctx->stream>>>(fp, bt, n)));
                } else {
                    CK(cudaFuncSetAttribute(k_film<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
                    LAUNCH(NORI_K_FILM, (k_film<false><<<grid, dim3(32, 32), smem, ctx->stream>>>(fp, bt, n)));
                }
            }
            CK(cudaGetLastError());
        }
        if (foldStats(ctx, (unsigned long long) n * wh)) return 1;
        done += n;
    }
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaEventSynchronize(ctx->ev1));
    float ms = 0.f; CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->stats.render_ms = ms;
    return 0;
}

// Reusable device scratch for the calls that stage a host buffer (resolve, variance, trace, probes, pcg32): grown on
// demand, never shrunk, freed with the context -- no cudaMalloc / cudaFree on any per-call path.
static int scratch(nori_gpu_ctx *ctx, size_t bytes, void **out) {
    if (bytes > ctx->scratch_cap) {
        cudaFree(ctx->scratch); ctx->scratch = nullptr; ctx->scratch_cap = 0;
        CK(cudaMalloc(&ctx->scratch, bytes));
        ctx->scratch_cap = bytes;
    }
    *out = ctx->scratch;
    return 0;
}

// nori_gpu_trace with option "trace_kernel" = 2: the rays go through the kernels that render large scenes
// (k_extend_sm / k_shadow_sm of wave_extend.cu, with the configured child order, node layout and counter variant),
// one pool-sized chunk at a time.  See k_trace_load / k_trace_store (film_kernels.cuh).
static int traceThroughRenderKernels(nori_gpu_ctx *ctx, const nori_gpu_ray *rays, uint64_t n, int shadow, nori_gpu_hit *out) {
    const bool count = ctx->opt_stats != 0;
    ctx->ds.ordered = ctx->opt_order == 1 || (ctx->opt_order == 2 && ctx->ds.n_prims > 4096);
    ctx->ds.wide = ctx->opt_wide ? 1 : 0;
    if (ensurePool(ctx, true, false)) return 1;
    const uint32_t P = ctx->pool.P;
    if (ensureResults(ctx, P)) return 1;
    nori_gpu_ray *dr = nullptr;
    if (scratch(ctx, (size_t) P * (sizeof(nori_gpu_ray) + sizeof(nori_gpu_hit)), (void **) &dr)) return 1;
    nori_gpu_hit *dh = (nori_gpu_hit *) (dr + P);
    DScene ds = ctx->ds;
    ds.camera.type = NORI_CAMERA_PERSPECTIVE;              // a miss ends the path here: no per-channel restarts (kernels.cuh: endOfPath)
    const int lay = noriSmLayout(ds);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const WaveKernels &wk = ctx->has_perlin ? kWavePerlin : kWave;
    const ExtendKernel kext = wk.pickExtend(true, count, false, lay);
    int occE = 8;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occE, kext, 128, 0);
    const int gridE = sms * std::max(1, occE), gridShadow = sms * std::max(1, wk.shadowSmOccupancy(count, lay));
    Batch bt{}; bt.results = ctx->results; bt.seed = 0; bt.spp_first = 0; bt.wh = P; bt.capacity = (uint32_t) std::min<size_t>(ctx->results_cap, 0xffffffffu);
    float msTotal = 0.f;
    ctx->last_wave = true; ctx->last_defer = true;
    for (uint64_t done = 0; done < n; done += P) {
        const uint32_t nb = (uint32_t) std::min<uint64_t>(P, n - done);
        CK(cudaMemsetAsync(ctx->ctr, 0, sizeof(Counters), ctx->stream));      // total_samples = 0: nothing is regenerated
        CK(cudaMemcpyAsync(dr, rays + done, (size_t) nb * sizeof(nori_gpu_ray), cudaMemcpyHostToDevice, ctx->stream));
        const unsigned gridP = (P + 255) / 256;
        if (shadow) k_trace_load<true><<<gridP, 256, 0, ctx->stream>>>(ctx->pool, dr, nb);
        else k_trace_load<false><<<gridP, 256, 0, ctx->stream>>>(ctx->pool, dr, nb);
        CK(cudaEventRecord(ctx->ev0, ctx->stream));
        if (shadow) LAUNCH(NORI_K_SHADOW, wk.shadowSm(count, gridShadow, ctx->stream, ds, ctx->pool, bt, ctx->ctr, 0));
        else LAUNCH(NORI_K_EXTEND, (kext<<<gridE, 128, 0, ctx->stream>>>(ds, ctx->pool, bt, ctx->ctr, 0)));
        CK(cudaEventRecord(ctx->ev1, ctx->stream));
        if (shadow) k_trace_store<true><<<(nb + 255) / 256, 256, 0, ctx->stream>>>(ds, ctx->pool, nb, dh);
        else k_trace_store<false><<<(nb + 255) / 256, 256, 0, ctx->stream>>>(ds, ctx->pool, nb, dh);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(out + done, dh, (size_t) nb * sizeof(nori_gpu_hit), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        float ms = 0.f; CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1)); msTotal += ms;
        if (foldStats(ctx, 0)) return 1;                   // rays / boxes / primitive tests into the kernel classes' counters
    }
    ctx->stats.trace_ms = msTotal;
    return 0;
}

// ---- multi-GPU inside render: device g renders the sample indices [begin + g*count/G, begin +
// (g+1)*count/G) -- disjoint pcg32 initstate ranges, scene replicated -- each from its own host thread (the wavefront
// loop polls its device), and the films are summed onto devices[0] by ONE kernel that reads the peers' accumulation
// buffers in place over NVLink (peer access) and hands them back zeroed, so that repeated render calls keep
// accumulating.  Peers without direct access are staged with cudaMemcpyPeerAsync first.
#define NORI_MAX_PEERS 15
struct PeerFilms { float4 *src[NORI_MAX_PEERS]; int n; };
__global__ void k_film_reduce(float4 *dst, PeerFilms peers, size_t n) {
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float4 a = dst[i];
        for (int p = 0; p < peers.n; ++p) {                  // fixed order: the sum does not depend on scheduling
            const float4 b = peers.src[p][i];
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            peers.src[p][i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        dst[i] = a;
    }
}

static int renderMulti(nori_gpu_ctx *ctx, uint32_t spp_begin, uint32_t spp_count, uint64_t seed) {
    REQUIRE(ctx->has_scene, "render: no scene uploaded");
    const int G = 1 + (int) ctx->peers.size();
    REQUIRE(G - 1 <= NORI_MAX_PEERS, "render: too many devices");
    std::vector<int> rc(G, 0);
    std::vector<std::thread> th;
    auto share = [&](int g) { return (uint32_t) (((uint64_t) spp_count * g) / G); };
    for (int g = 1; g < G; ++g)
        th.emplace_back([&, g] { rc[g] = renderImpl(ctx->peers[g - 1], spp_begin + share(g), share(g + 1) - share(g), seed, nullptr); });
    rc[0] = renderImpl(ctx, spp_begin + share(0), share(1) - share(0), seed, nullptr);
    for (auto &t : th) t.join();
    for (int g = 1; g < G; ++g) if (rc[g]) { ctx->err = "device " + std::to_string(ctx->peers[g - 1]->device) + ": " + ctx->peers[g - 1]->err; return 1; }
    if (rc[0]) return 1;
    // ---- the one exchange step of the path: sum of the accumulation buffers
    CK(cudaSetDevice(ctx->device));
    const size_t nf = (size_t) (ctx->W + 2 * ctx->border) * (ctx->H + 2 * ctx->border);
    PeerFilms pf{}; pf.n = G - 1;
    size_t staged = 0;
    for (int g = 1; g < G; ++g) if (!ctx->peer_direct[g - 1]) ++staged;
    if (staged * nf > ctx->peer_stage_cap) {
        cudaFree(ctx->peer_stage); ctx->peer_stage = nullptr; ctx->peer_stage_cap = 0;
        CK(cudaMalloc((void **) &ctx->peer_stage, staged * nf * sizeof(float4)));
        ctx->peer_stage_cap = staged * nf;
    }
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    staged = 0;
    for (int g = 1; g < G; ++g) {
        nori_gpu_ctx *p = ctx->peers[g - 1];
        if (ctx->peer_direct[g - 1]) pf.src[g - 1] = p->film;
        else {
            float4 *st = ctx->peer_stage + (staged++) * nf;
            CK(cudaMemcpyPeerAsync(st, ctx->device, p->film, p->device, nf * sizeof(float4), ctx->stream));
            pf.src[g - 1] = st;
        }
    }
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    LAUNCH(NORI_K_FILM, (k_film_reduce<<<sms * 4, 256, 0, ctx->stream>>>(ctx->film, pf, nf)));
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaEventSynchronize(ctx->ev1));
    float red = 0.f; CK(cudaEventElapsedTime(&red, ctx->ev0, ctx->ev1));
    ctx->multi_reduce_ms = red;
    double worst = ctx->stats.render_ms;
    for (int g = 1; g < G; ++g) {
        nori_gpu_ctx *p = ctx->peers[g - 1];
        if (!ctx->peer_direct[g - 1]) {                      // the staged copy was summed: clear the peer's own buffer
            CK(cudaSetDevice(p->device));
            CK(cudaMemsetAsync(p->film, 0, nf * sizeof(float4), p->stream));
            CK(cudaStreamSynchronize(p->stream));
        }
        const nori_gpu_stats &q = p->stats;
        worst = std::max(worst, q.render_ms);
        ctx->stats.samples += q.samples; ctx->stats.rays += q.rays; ctx->stats.shadow_rays += q.shadow_rays;
        ctx->stats.nodes_visited += q.nodes_visited; ctx->stats.prims_tested += q.prims_tested;
        ctx->stats.invalid_samples += q.invalid_samples; ctx->stats.kernel_launches += q.kernel_launches;
        ctx->stats.iterations += q.iterations; ctx->stats.guard_retraces += q.guard_retraces;
        ctx->stats.max_stack_depth = std::max(ctx->stats.max_stack_depth, q.max_stack_depth);
        for (int i = 0; i < NORI_K_COUNT; ++i) {             // work counters add up; times stay those of devices[0]
            ctx->kstats[i].rays += p->kstats[i].rays; ctx->kstats[i].nodes_visited += p->kstats[i].nodes_visited;
            ctx->kstats[i].prims_tested += p->kstats[i].prims_tested;
        }
        nori_gpu_reset_stats(p);
    }
    CK(cudaSetDevice(ctx->device));
    ctx->stats.render_ms = worst + red;                      // the devices run side by side: slowest device + the reduce
    ctx->stats.reduce_ms = red;
    return 0;
}

extern "C" {

int nori_gpu_render(nori_gpu_ctx *ctx, uint32_t spp_begin, uint32_t spp_count, uint64_t seed) {
    if (ctx && !ctx->peers.empty()) return renderMulti(ctx, spp_begin, spp_count, seed);
    return renderImpl(ctx, spp_begin, spp_count, seed, nullptr);
}

int nori_gpu_render_samples(nori_gpu_ctx *ctx, uint32_t spp_begin, uint32_t spp_count, uint64_t seed, float *out_rgba) {
    REQUIRE(ctx && out_rgba, "render_samples: null output");
    return renderImpl(ctx, spp_begin, spp_count, seed, out_rgba);
}

int nori_gpu_clear_film(nori_gpu_ctx *ctx) {
    REQUIRE(ctx && ctx->has_scene, "clear_film: no scene uploaded");
    for (nori_gpu_ctx *p : ctx->peers) if (nori_gpu_clear_film(p)) { ctx->err = p->err; return 1; }
    CK(cudaSetDevice(ctx->device));
    size_t nf = (size_t) (ctx->W + 2 * ctx->border) * (ctx->H + 2 * ctx->border);
    CK(cudaMemsetAsync(ctx->film, 0, nf * sizeof(float4), ctx->stream));
    if (ctx->vsum) { CK(cudaMemsetAsync(ctx->vsum, 0, nf * sizeof(float4), ctx->stream)); CK(cudaMemsetAsync(ctx->vsum2, 0, nf * sizeof(float4), ctx->stream)); }
    ctx->var_passes = 0;
    return 0;
}

int nori_gpu_download_variance(nori_gpu_ctx *ctx, float *rgb) {
    REQUIRE(ctx && ctx->has_scene && rgb, "download_variance: no scene / null buffer");
    REQUIRE(ctx->vsum && ctx->var_passes > 0, "download_variance: enable option \"variance\" before rendering");
    CK(cudaSetDevice(ctx->device));
    float *d = nullptr; size_t n = (size_t) ctx->W * ctx->H * 3;
    if (scratch(ctx, n * sizeof(float), (void **) &d)) return 1;
    dim3 blk(32, 8), grid((ctx->W + 31) / 32, (ctx->H + 7) / 8);
    k_variance<<<grid, blk, 0, ctx->stream>>>(ctx->vsum, ctx->vsum2, d, ctx->W, ctx->H, ctx->border, (float) ctx->var_passes);
    cudaError_t e = cudaMemcpyAsync(rgb, d, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { ctx->err = std::string("download_variance: ") + cudaGetErrorString(e); return 1; }
    return 0;
}

int nori_gpu_download_film(nori_gpu_ctx *ctx, float *rgbaw) {
    REQUIRE(ctx && ctx->has_scene && rgbaw, "download_film: no scene / null buffer");
    CK(cudaSetDevice(ctx->device));
    size_t nf = (size_t) (ctx->W + 2 * ctx->border) * (ctx->H + 2 * ctx->border);
    CK(cudaMemcpyAsync(rgbaw, ctx->film, nf * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int nori_gpu_upload_film(nori_gpu_ctx *ctx, const float *rgbaw) {
    REQUIRE(ctx && ctx->has_scene && rgbaw, "upload_film: no scene / null buffer");
    CK(cudaSetDevice(ctx->device));
    size_t nf = (size_t) (ctx->W + 2 * ctx->border) * (ctx->H + 2 * ctx->border);
    CK(cudaMemcpyAsync(ctx->film, rgbaw, nf * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int nori_gpu_film_device_ptr(nori_gpu_ctx *ctx, void **dptr, uint64_t *n_floats) {
    REQUIRE(ctx && ctx->has_scene && dptr && n_floats, "film_device_ptr: no scene / null argument");
    *dptr = ctx->film;
    *n_floats = (uint64_t) (ctx->W + 2 * ctx->border) * (ctx->H + 2 * ctx->border) * 4;
    return 0;
}

int nori_gpu_film_dims(const nori_gpu_ctx *ctx, int32_t *rows, int32_t *cols, int32_t *border) {
    if (!ctx || !ctx->has_scene) return 1;
    *rows = ctx->H + 2 * ctx->border; *cols = ctx->W + 2 * ctx->border; *border = ctx->border;
    return 0;
}

int nori_gpu_resolve(nori_gpu_ctx *ctx, float *rgb) {
    REQUIRE(ctx && ctx->has_scene && rgb, "resolve: no scene / null buffer");
    CK(cudaSetDevice(ctx->device));
    float *d = nullptr; size_t n = (size_t) ctx->W * ctx->H * 3;
    if (scratch(ctx, n * sizeof(float), (void **) &d)) return 1;
    dim3 blk(32, 8), grid((ctx->W + 31) / 32, (ctx->H + 7) / 8);
    k_resolve<<<grid, blk, 0, ctx->stream>>>(ctx->film, d, ctx->W, ctx->H, ctx->border);
    cudaError_t e = cudaMemcpyAsync(rgb, d, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { ctx->err = std::string("resolve: ") + cudaGetErrorString(e); return 1; }
    return 0;
}

int nori_gpu_trace(nori_gpu_ctx *ctx, const nori_gpu_ray *rays, uint64_t n, int shadow, nori_gpu_hit *out) {
    REQUIRE(ctx && ctx->has_scene, "trace: no scene uploaded");
    if (n == 0) return 0;                                  // empty batch is a no-op
    REQUIRE(rays && out, "trace: null buffer");
    CK(cudaSetDevice(ctx->device));
    if (ctx->opt_trace_kernel == 2) return traceThroughRenderKernels(ctx, rays, n, shadow, out);
    nori_gpu_ray *dr = nullptr; nori_gpu_hit *dh = nullptr;
    if (scratch(ctx, n * (sizeof(nori_gpu_ray) + sizeof(nori_gpu_hit)), (void **) &dr)) return 1;
    dh = (nori_gpu_hit *) (dr + n);
    cudaError_t e = cudaMemcpyAsync(dr, rays, n * sizeof(nori_gpu_ray), cudaMemcpyHostToDevice, ctx->stream);
    cudaEventRecord(ctx->ev0, ctx->stream);
    const unsigned grid = (unsigned) ((n + 127) / 128);
    ctx->ds.ordered = ctx->opt_order == 1;
    if (shadow) k_trace<true><<<grid, 128, 0, ctx->stream>>>(ctx->ds, dr, n, dh);
    else k_trace<false><<<grid, 128, 0, ctx->stream>>>(ctx->ds, dr, n, dh);
    ctx->stats.kernel_launches++;
    cudaEventRecord(ctx->ev1, ctx->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dh, n * sizeof(nori_gpu_hit), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    float ms = 0.f; if (e == cudaSuccess) cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->stats.trace_ms = ms;
    if (e != cudaSuccess) { ctx->err = std::string("trace: ") + cudaGetErrorString(e); return 1; }
    return 0;
}

} // extern "C"

static int probeImpl(nori_gpu_ctx *ctx, bool bsdf, uint32_t index, uint64_t n, const float *in, float *out) {
    REQUIRE(ctx && ctx->has_scene, "probe: no scene uploaded");
    if (n == 0) return 0;
    REQUIRE(in && out, "probe: null buffer");
    REQUIRE(index < (bsdf ? (uint32_t) ctx->n_bsdfs : ctx->ds.n_emitters), "probe: index out of range");
    CK(cudaSetDevice(ctx->device));
    const size_t ni = (bsdf ? 10 : 5) * n, no = (bsdf ? 12 : 15) * n;
    float *di = nullptr, *dout = nullptr;
    if (scratch(ctx, (ni + no) * 4, (void **) &di)) return 1;
    dout = di + ni;
    cudaError_t e = cudaMemcpyAsync(di, in, ni * 4, cudaMemcpyHostToDevice, ctx->stream);
    const unsigned grid = (unsigned) ((n + 127) / 128);
    if (bsdf) k_probe_bsdf<<<grid, 128, 0, ctx->stream>>>(ctx->ds, index, n, di, dout);
    else k_probe_emitter<<<grid, 128, 0, ctx->stream>>>(ctx->ds, index, n, di, dout);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout, no * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { ctx->err = std::string("probe: ") + cudaGetErrorString(e); return 1; }
    return 0;
}

static int pcgImpl(nori_gpu_ctx *ctx, uint64_t initstate, uint64_t initseq, uint64_t n, float *outf, uint32_t *outu) {
    REQUIRE(ctx, "pcg32: null context");
    if (n == 0) return 0;
    REQUIRE(outf || outu, "pcg32: null buffer");
    CK(cudaSetDevice(ctx->device));
    void *d = nullptr;
    if (scratch(ctx, n * 4, &d)) return 1;
    k_pcg32<<<1, 1, 0, ctx->stream>>>(initstate, initseq, n, outf ? (float *) d : nullptr, outf ? nullptr : (uint32_t *) d);
    cudaError_t e = cudaMemcpyAsync(outf ? (void *) outf : (void *) outu, d, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { ctx->err = std::string("pcg32: ") + cudaGetErrorString(e); return 1; }
    return 0;
}
extern "C" {

int nori_gpu_probe_bsdf(nori_gpu_ctx *ctx, uint32_t bsdf, uint64_t n, const float *in, float *out) { return probeImpl(ctx, true, bsdf, n, in, out); }
int nori_gpu_probe_emitter(nori_gpu_ctx *ctx, uint32_t emitter, uint64_t n, const float *in, float *out) { return probeImpl(ctx, false, emitter, n, in, out); }
int nori_gpu_pcg32(nori_gpu_ctx *ctx, uint64_t initstate, uint64_t initseq, uint64_t n, float *out) { return pcgImpl(ctx, initstate, initseq, n, out, nullptr); }
int nori_gpu_pcg32_uint(nori_gpu_ctx *ctx, uint64_t initstate, uint64_t initseq, uint64_t n, uint32_t *out) { return pcgImpl(ctx, initstate, initseq, n, nullptr, out); }

int nori_gpu_selftest(nori_gpu_ctx *ctx, uint64_t n, uint64_t *mismatch) {
    REQUIRE(ctx && mismatch, "selftest: null argument");
    CK(cudaSetDevice(ctx->device));
    unsigned long long *d = nullptr;
    if (scratch(ctx, 3 * sizeof(unsigned long long), (void **) &d)) return 1;
    CK(cudaMemsetAsync(d, 0, 3 * sizeof(unsigned long long), ctx->stream));
    if (n) k_selftest<<<(unsigned) ((n + 255) / 256), 256, 0, ctx->stream>>>(n, d);
    CK(cudaGetLastError());
    unsigned long long h[3];
    CK(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 3; ++i) mismatch[i] = h[i];
    return 0;
}

int nori_gpu_get_stats(nori_gpu_ctx *ctx, nori_gpu_stats *out) {
    REQUIRE(ctx && out, "get_stats: null argument");
    *out = ctx->stats; out->devices = 1 + ctx->peers.size();
    return 0;
}
int nori_gpu_get_kernel_stats(nori_gpu_ctx *ctx, nori_gpu_kernel_stats *out) {
    REQUIRE(ctx && out, "get_kernel_stats: null argument");
    for (int i = 0; i < NORI_K_COUNT; ++i) out[i] = ctx->kstats[i];
    return 0;
}
int nori_gpu_reset_stats(nori_gpu_ctx *ctx) {
    REQUIRE(ctx, "reset_stats: null context");
    ctx->stats = nori_gpu_stats{};
    for (int i = 0; i < NORI_K_COUNT; ++i) ctx->kstats[i] = nori_gpu_kernel_stats{};
    return 0;
}
int nori_gpu_synchronize(nori_gpu_ctx *ctx) {
    REQUIRE(ctx, "synchronize: null context");
    for (nori_gpu_ctx *p : ctx->peers) if (nori_gpu_synchronize(p)) { ctx->err = p->err; return 1; }
    CK(cudaSetDevice(ctx->device)); CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

} // extern "C"
