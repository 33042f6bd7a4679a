// kernels.cuh -- the __global__ kernels of the wavefront path tracer.
//
//   k_extend     persistent-thread closest-hit traversal over the pool (bvh.cpp:404-462) fused with raygen /
//                path regeneration: free slots claim the next sample index, seed their pcg32 stream, draw
//                the film/aperture samples and build the camera ray (render.cpp:98-124,
//                perspective.cpp:90-112, thinlens.cpp:126-171); misses are finalised and regenerated in
//                place; hits are binned by BSDF type into material queues with warp-aggregated atomics
//   k_shade      all material queues in one launch (material-coherent warps): hit info, emission (+MIS weight), NEE sample + BSDF
//                eval/pdf, the any-hit query of the NEE shadow ray (bvh.cpp:441-442), Russian roulette,
//                BSDF sample (path_mis.cpp:32-97 / path_mats.cpp:23-55); finalises paths the roulette ended
//   k_film       reconstruction-filter accumulation of a batch of finished samples into the film:
//                one CTA per 32x32 film tile, samples of the tile + halo staged in shared memory,
//                each film pixel owned by exactly one thread (gather, no atomics), one coalesced
//                float4 read-modify-write per film pixel per batch (block.cpp:93-133)
//   k_resolve    ImageBlock::toBitmap (block.cpp:76-82)
//   k_mega       one thread per sample for the short integrators (normals, av, direct*, volumetric)
//   k_trace / k_pcg32*   the ABI's test hooks
#pragma once
#ifndef NORI_PERLIN_VARIANT
#define NORI_PERLIN_VARIANT 0
#endif
#include "integrators.cuh"
#include "host_layout.h"        // NORI_STACK2_MAX

#define NORI_FREE_SLOT 0xffffffffu

struct Pool {
    float4 *rayO, *rayD;      // (o.xyz, mint) (d.xyz, maxt): the 32-byte ray record
    float4 *hit;              // (t, u, v, leafpos): the 16-byte hit record
    float4 *thr;              // (throughput rgb, pdf_mat)
    float4 *rad;              // (radiance rgb, -)
    float4 *acc;              // chromatic aberration only: value accumulated over the colour channels already traced
    float4 *shD, *shC;        // deferred NEE shadow ray (large scenes only): (direction, maxt) -- origin = rayO, mint = Epsilon --
                              // and its pending contribution; NULL when shadow rays are traced inside k_shade
    uint64_t *rng;            // pcg32 state (inc is a function of the pixel)
    uint32_t *sid;            // sample id inside the batch, NORI_FREE_SLOT when the slot is free
    uint32_t *flags;          // PF_*
    uint32_t *queue[NORI_NQ];
    uint32_t *equeue;         // emitter-sorted mode: NORI_NEQ sub-queues of P entries each (key = bsdf type * 4 + emitter type), or NULL
    uint32_t P;
};

struct Counters {
    unsigned long long next_sample, total_samples, done;
    unsigned long long rays_ext, rays_sh, nodes_ext, prims_ext, nodes_sh, prims_sh, invalid;
    unsigned long long guard_redo;           // with counters on: near-first queries answered again in reference order (traverse.cuh: order guard)
    unsigned long long rays_sh_closest;      // volumetric.cpp:63: the medium vertex's NEE query is a closest-hit one
    // per-iteration scheduling state, double-buffered by iteration parity: k_extend(it) uses [it & 1]
    // and zeroes [(it + 1) & 1], whose last readers (the kernels of iteration it - 1) have finished
    uint32_t qcount[2][NORI_NQ];
    uint32_t eqcount[2][NORI_NEQ];            // emitter-sorted mode: entries per (bsdf type, emitter type)
    uint32_t work_extend[2], work_shadow[2];
    uint32_t max_stack;                      // with counters on: deepest per-ray stack of the large-scene kernels
};

struct Batch {
    float4 *results;          // [spp_local][H][W] (r,g,b,valid)
    uint64_t seed;            // initstate of sample k is seed + k
    uint32_t spp_first;       // first absolute sample index of this batch
    uint32_t wh;              // W*H
    uint32_t capacity;        // samples the results buffer holds (assert-enabled build)
};

__device__ __forceinline__ void warpAdd(unsigned long long *dst, uint32_t v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(dst, (unsigned long long) v);
}

__device__ __forceinline__ void warpMax(uint32_t *dst, uint32_t v) {
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_down_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0 && v) atomicMax(dst, v);
}

__device__ __forceinline__ void finalizePath(const Batch &bt, Counters *ctr, uint32_t sid, V3 rad) {
    bool ok = validColor(rad);
    NORI_CHECK(sid < bt.capacity);
    bt.results[sid] = ok ? make_float4(rad.x, rad.y, rad.z, 1.f) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (!ok) atomicAdd(&ctr->invalid, 1ull);
}

// ------------------------------------------------------------------------------ raygen (device function)
// One iteration of renderBlock's loop head (render.cpp:98-124): seed the path's pcg32 stream, draw the
// film and aperture samples, build the camera ray (of colour channel `channel` when the camera has
// chromatic aberration, render.cpp:106-121).
template <bool NR = true>
__device__ __forceinline__ void generatePath(const DScene &sc, const Batch &bt, uint32_t sid, int channel, Ray &ray, uint64_t &rngState) {
    const uint32_t k = sid / bt.wh, pix = sid - k * bt.wh;
    const int W = sc.camera.width;
    const int py = pix / W, px = pix - py * W;
    Pcg32 rng; rng.seed(bt.seed + bt.spp_first + k, (uint64_t) pix);
    P2 a = rng.next2D();
    P2 ps; ps.x = (float) px + a.x; ps.y = (float) py + a.y;
    P2 ap = rng.next2D();
    V3 weight;
    ray = cameraRay<NR>(sc.camera, ps, ap, channel, weight);
    rngState = rng.state;
}

// A camera path has ended with radiance `rad`.  Normally that is the sample's value.  With chromatic
// aberration one sample is THREE paths, one per colour channel, traced one after the other on the same
// random stream and summed with the camera's per-channel weights (render.cpp:106-121,
// advancedCamera.cpp:176-183): the slot is restarted in place with the next channel's camera ray
// (same film / aperture sample, re-derived from the stream's seed) until the third path has ended.
template <bool NR = true>
__device__ __forceinline__ void endOfPath(const DScene &sc, const Pool &pool, const Batch &bt, Counters *ctr, uint32_t slot,
                                          uint32_t sid, V3 rad, uint64_t rngState, uint32_t flags, uint32_t &nDone) {
    if (hasChromaticAberrations(sc.camera)) {
        const int ch = (int) ((flags & PF_CH_MASK) >> PF_CH_SHIFT);
        const V3 w = mk(ch == 0 ? 1.f : 0.f, ch == 1 ? 1.f : 0.f, ch == 2 ? 1.f : 0.f);
        V3 acc = w * rad;                                          // value_ch = sampleRay(...) * Li
        if (ch > 0) { const float4 a = pool.acc[slot]; acc = mk(a.x, a.y, a.z) + acc; }
        if (ch < 2) {
            Ray ray; uint64_t unused;
            generatePath<NR>(sc, bt, sid, ch + 1, ray, unused);
            pool.acc[slot] = make_float4(acc.x, acc.y, acc.z, 0.f);
            pool.rayO[slot] = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.mint);
            pool.rayD[slot] = make_float4(ray.d.x, ray.d.y, ray.d.z, ray.maxt);
            pool.thr[slot] = make_float4(1.f, 1.f, 1.f, 0.f);
            pool.rad[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
            pool.rng[slot] = rngState;                              // the sampler carries on where this path stopped
            pool.flags[slot] = PF_ALIVE | PF_FIRST | ((uint32_t) (ch + 1) << PF_CH_SHIFT);
            return;
        }
        rad = acc;
    }
    finalizePath(bt, ctr, sid, rad);
    pool.sid[slot] = NORI_FREE_SLOT; pool.flags[slot] = 0u; ++nDone;
}

enum { MODE_MATS = 0, MODE_MIS = 1, MODE_VOL = 2 };

// ---- host-side launchers; each group of kernels lives in its own translation unit so that the
// library builds in parallel (wave_extend.cu, wave_shade.cu x3 modes, mega.cu, nori_gpu.cu)
//
// The wavefront kernels exist TWICE: the regular set knows triangles and spheres; the set compiled with
// -DNORI_PERLIN_VARIANT=1 (wave_*_p.o: same sources, NORI_WITH_PERLIN on, every kernel and launcher renamed below) also
// knows the Perlin-noise sphere (perlinnoise.cpp) and renders the scenes that contain one.  Two sets because the
// shape's code inside the traversal loops -- even as an out-of-line call that is never taken -- costs the scenes
// WITHOUT such a shape 30 % (traverse.cuh).
#if NORI_PERLIN_VARIANT
#define k_extend k_extend_perlin
#define k_extend_sm k_extend_sm_perlin
#define k_shadow_sm k_shadow_sm_perlin
#define k_shade k_shade_perlin
#define k_rebin k_rebin_perlin
#define k_drain k_drain_perlin
#define k_drain_warp k_drain_warp_perlin
#define noriPickExtend noriPickExtendPerlin
#define noriLaunchShadeMats noriLaunchShadeMatsPerlin
#define noriLaunchShadeMisDeferred noriLaunchShadeMisDeferredPerlin
#define noriLaunchShadowSm noriLaunchShadowSmPerlin
#define noriShadowSmOccupancy noriShadowSmOccupancyPerlin
#define noriLaunchShadeMis noriLaunchShadeMisPerlin
#define noriLaunchShadeVol noriLaunchShadeVolPerlin
#define noriLaunchRebin noriLaunchRebinPerlin
#define noriLaunchDrain noriLaunchDrainPerlin
#endif
typedef void (*ExtendKernel)(DScene, Pool, Batch, Counters *, uint32_t);
ExtendKernel noriPickExtend(bool stateMachine, bool count, bool vol, int layout);   // layout: 0 reference nodes, 1 child-box pairs, 2 4-wide
void noriLaunchShadeMats(bool count, int grid, cudaStream_t st, const DScene &sc, const Pool &pool, const Batch &bt, Counters *ctr, uint32_t it);
void noriLaunchShadeMisDeferred(bool count, int grid, cudaStream_t st, const DScene &sc, const Pool &pool, const Batch &bt, Counters *ctr, uint32_t it);
void noriLaunchShadowSm(bool count, int grid, cudaStream_t st, const DScene &sc, const Pool &pool, const Batch &bt, Counters *ctr, uint32_t it);
int noriShadowSmOccupancy(bool count, int layout);
static inline int noriSmLayout(const DScene &sc) { return sc.ordered ? (sc.wide && sc.nodes4 ? 2 : sc.nodes2 ? 1 : 0) : 0; }
void noriLaunchShadeMis(bool count, int grid, cudaStream_t st, const DScene &sc, const Pool &pool, const Batch &bt, Counters *ctr, uint32_t it);
void noriLaunchShadeVol(bool count, int grid, cudaStream_t st, const DScene &sc, const Pool &pool, const Batch &bt, Counters *ctr, uint32_t it);
void noriLaunchRebin(int grid, cudaStream_t st, const DScene &sc, const Pool &pool, const Batch &bt, Counters *ctr, uint32_t it);
void noriLaunchDrain(bool mis, bool count, int grid, cudaStream_t st, const DScene &sc, const Pool &pool, const Batch &bt, Counters *ctr);
void noriLaunchMega(bool count, unsigned grid, cudaStream_t st, const DScene &sc, const Batch &bt, Counters *ctr, unsigned long long total);
#if !NORI_PERLIN_VARIANT
// the Perlin-aware set (see above)
ExtendKernel noriPickExtendPerlin(bool stateMachine, bool count, bool vol, int layout);
void noriLaunchShadeMatsPerlin(bool count, int grid, cudaStream_t st, const DScene &sc, const Pool &pool, const Batch &bt, Counters *ctr, uint32_t it);
void noriLaunchShadeMisDeferredPerlin(bool count, int grid, cudaStream_t st, const DScene &sc, const Pool &pool, const Batch &bt, Counters *ctr, uint32_t it);
void noriLaunchShadowSmPerlin(bool count, int grid, cudaStream_t st, const DScene &sc, const Pool &pool, const Batch &bt, Counters *ctr, uint32_t it);
int noriShadowSmOccupancyPerlin(bool count, int layout);
void noriLaunchShadeMisPerlin(bool count, int grid, cudaStream_t st, const DScene &sc, const Pool &pool, const Batch &bt, Counters *ctr, uint32_t it);
void noriLaunchShadeVolPerlin(bool count, int grid, cudaStream_t st, const DScene &sc, const Pool &pool, const Batch &bt, Counters *ctr, uint32_t it);
void noriLaunchRebinPerlin(int grid, cudaStream_t st, const DScene &sc, const Pool &pool, const Batch &bt, Counters *ctr, uint32_t it);
void noriLaunchDrainPerlin(bool mis, bool count, int grid, cudaStream_t st, const DScene &sc, const Pool &pool, const Batch &bt, Counters *ctr);
#endif
// one set of wavefront entry points (regular / Perlin-aware), picked per scene by nori_gpu.cu
struct WaveKernels {
    ExtendKernel (*pickExtend)(bool, bool, bool, int);
    void (*shadeMats)(bool, int, cudaStream_t, const DScene &, const Pool &, const Batch &, Counters *, uint32_t);
    void (*shadeMisDeferred)(bool, int, cudaStream_t, const DScene &, const Pool &, const Batch &, Counters *, uint32_t);
    void (*shadowSm)(bool, int, cudaStream_t, const DScene &, const Pool &, const Batch &, Counters *, uint32_t);
    int (*shadowSmOccupancy)(bool, int);
    void (*shadeMis)(bool, int, cudaStream_t, const DScene &, const Pool &, const Batch &, Counters *, uint32_t);
    void (*shadeVol)(bool, int, cudaStream_t, const DScene &, const Pool &, const Batch &, Counters *, uint32_t);
    void (*rebin)(int, cudaStream_t, const DScene &, const Pool &, const Batch &, Counters *, uint32_t);
    void (*drain)(bool, bool, int, cudaStream_t, const DScene &, const Pool &, const Batch &, Counters *);
};
