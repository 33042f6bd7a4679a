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
#include "integrators.cuh"

#define NORI_FREE_SLOT 0xffffffffu
#define NORI_Q_MISS NORI_BSDF_COUNT          // volumetric only: rays that left the scene may still scatter in the medium
#define NORI_NQ (NORI_BSDF_COUNT + 1)

struct Pool {
    float4 *rayO, *rayD;      // (o.xyz, mint) (d.xyz, maxt): the 32-byte ray record
    float4 *hit;              // (t, u, v, leafpos): the 16-byte hit record
    float4 *thr;              // (throughput rgb, pdf_mat)
    float4 *rad;              // (radiance rgb, -)
    uint64_t *rng;            // pcg32 state (inc is a function of the pixel)
    uint32_t *sid;            // sample id inside the batch, NORI_FREE_SLOT when the slot is free
    uint32_t *flags;          // PF_*
    uint32_t *queue[NORI_NQ];
    uint32_t P;
};

struct Counters {
    unsigned long long next_sample, total_samples, done;
    unsigned long long rays_ext, rays_sh, nodes_ext, prims_ext, nodes_sh, prims_sh, invalid;
    unsigned long long rays_sh_closest;      // volumetric.cpp:63: the medium vertex's NEE query is a closest-hit one
    // per-iteration scheduling state, double-buffered by iteration parity: k_extend(it) uses [it & 1]
    // and zeroes [(it + 1) & 1], whose last readers (the kernels of iteration it - 1) have finished
    uint32_t qcount[2][NORI_NQ];
    uint32_t work_extend[2], pad[2];
};

struct Batch {
    float4 *results;          // [spp_local][H][W] (r,g,b,valid)
    uint64_t seed;            // initstate of sample k is seed + k
    uint32_t spp_first;       // first absolute sample index of this batch
    uint32_t wh;              // W*H
};

__device__ __forceinline__ void warpAdd(unsigned long long *dst, uint32_t v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(dst, (unsigned long long) v);
}

__device__ __forceinline__ void finalizePath(const Batch &bt, Counters *ctr, uint32_t sid, V3 rad) {
    bool ok = validColor(rad);
    bt.results[sid] = ok ? make_float4(rad.x, rad.y, rad.z, 1.f) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (!ok) atomicAdd(&ctr->invalid, 1ull);
}

// ------------------------------------------------------------------------------ raygen (device function)
// One iteration of renderBlock's loop head (render.cpp:98-124): seed the path's pcg32 stream, draw the
// film and aperture samples, build the camera ray.
__device__ __forceinline__ void generatePath(const DScene &sc, const Batch &bt, uint32_t sid, Ray &ray, uint64_t &rngState) {
    const uint32_t k = sid / bt.wh, pix = sid - k * bt.wh;
    const int W = sc.camera.width;
    const int py = pix / W, px = pix - py * W;
    Pcg32 rng; rng.seed(bt.seed + bt.spp_first + k, (uint64_t) pix);
    P2 a = rng.next2D();
    P2 ps; ps.x = (float) px + a.x; ps.y = (float) py + a.y;
    P2 ap = rng.next2D();
    ray = cameraRay(sc.camera, ps, ap);
    rngState = rng.state;
}

#define NORI_FETCH 256u      // pool slots claimed per warp per atomic (8 rounds of 32)

// ------------------------------------------------------------------------------ extend (+ regeneration)
// Persistent warps claim NORI_FETCH consecutive pool slots at a time and run three phases on them:
//   1. regeneration, compacted: the warp gathers its free slots into a shared-memory list, claims that
//      many sample indices with ONE atomic, and then every lane generates one camera path per step
//      (render.cpp:98-124) -- full SIMT width although only ~1/3 of the slots are free per iteration;
//   2. closest-hit traversal of every live slot (bvh.cpp:404-462), 32 slots per step;
//   3. binning of the hits by BSDF type into the material queues (one atomic per warp and material).
// A path that escapes the scene is finalised here and its slot handed to the next iteration.
// The miss rule (shared by both extend kernels).  path_mis.cpp:28-29 / :84-85: a ray that leaves the scene ends
// the path.  volumetric.cpp:34-38,147-151: it ends only if it also misses the medium's box
// (medium.cpp:62-66 returns hitObject without drawing a number); otherwise the free-flight sample may
// still scatter it, so the slot goes to the miss queue with t = inf (the reference's its.t after a miss).
template <bool VOL>
__device__ __forceinline__ int missRule(const DScene &sc, const Pool &pool, const Batch &bt, Counters *ctr, uint32_t slot, V3 o, V3 d, uint32_t &nDone) {
    if (VOL) {
        float nearT, farT;
        if (boundsHit(sc.medium, o, d, nearT, farT)) {
            pool.hit[slot] = make_float4(__int_as_float(0x7f800000), 0.f, 0.f, __uint_as_float(NORI_NO_HIT));
            return NORI_Q_MISS;
        }
    }
    const float4 r = pool.rad[slot];
    finalizePath(bt, ctr, pool.sid[slot], mk(r.x, r.y, r.z));
    pool.sid[slot] = NORI_FREE_SLOT; pool.flags[slot] = 0u; ++nDone;
    return -1;
}

template <bool COUNT, bool VOL>
__global__ void __launch_bounds__(128) k_extend(DScene sc, Pool pool, Batch bt, Counters *ctr, uint32_t it) {
    __shared__ uint32_t s_free[4][NORI_FETCH];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, par = it & 1u;
    uint32_t *freeList = s_free[warp];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int i = 0; i < NORI_NQ; ++i) ctr->qcount[par ^ 1u][i] = 0;
        ctr->work_extend[par ^ 1u] = 0;
    }
    const unsigned long long total = ctr->total_samples;
    uint32_t nRays = 0, nDone = 0; TraceCounters cnt; cnt.nodes = 0; cnt.prims = 0;
    while (true) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&ctr->work_extend[par], NORI_FETCH);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= pool.P) break;
        // ---- phase 1: compacted regeneration
        uint32_t nFree = 0;
        for (uint32_t round = 0; round < NORI_FETCH / 32u; ++round) {
            const uint32_t slot = base + round * 32u + lane;
            const bool isFree = slot < pool.P && pool.sid[slot] == NORI_FREE_SLOT;
            const uint32_t m = __ballot_sync(0xffffffffu, isFree);
            if (isFree) freeList[nFree + __popc(m & ((1u << lane) - 1u))] = slot;
            nFree += __popc(m);
        }
        if (nFree) {
            unsigned long long first = 0;
            if (lane == 0) first = atomicAdd(&ctr->next_sample, (unsigned long long) nFree);
            first = __shfl_sync(0xffffffffu, first, 0);
            __syncwarp();
            for (uint32_t j = lane; j < nFree; j += 32u) {
                const unsigned long long id = first + j;
                if (id >= total) break;                          // batch exhausted: the slot stays free
                const uint32_t slot = freeList[j];
                Ray ray; uint64_t rs;
                generatePath(sc, bt, (uint32_t) id, ray, rs);
                pool.rayO[slot] = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.mint);
                pool.rayD[slot] = make_float4(ray.d.x, ray.d.y, ray.d.z, ray.maxt);
                pool.thr[slot] = make_float4(1.f, 1.f, 1.f, 0.f);
                pool.rad[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
                pool.rng[slot] = rs; pool.sid[slot] = (uint32_t) id;
                pool.flags[slot] = PF_ALIVE | PF_FIRST;
            }
            __syncwarp();
        }
        // ---- phase 2 + 3: trace and bin
        for (uint32_t round = 0; round < NORI_FETCH / 32u; ++round) {
            const uint32_t slot = base + round * 32u + lane;
            int type = -1;
            if (slot < pool.P && (pool.flags[slot] & PF_ALIVE)) {
                const float4 ro = pool.rayO[slot], rd = pool.rayD[slot];
                Hit h; ++nRays;
                if (traverse<false, COUNT>(sc, mk(ro.x, ro.y, ro.z), mk(rd.x, rd.y, rd.z), ro.w, rd.w, h, cnt)) {
                    pool.hit[slot] = make_float4(h.t, h.u, h.v, __uint_as_float(h.leafpos));
                    type = sc.shapes[__float_as_uint(__ldg(&sc.prims[3 * h.leafpos + 1]).w)].bsdf_type;
                } else type = missRule<VOL>(sc, pool, bt, ctr, slot, mk(ro.x, ro.y, ro.z), mk(rd.x, rd.y, rd.z), nDone);
            }
#pragma unroll
            for (int t = 0; t < (VOL ? NORI_NQ : NORI_BSDF_COUNT); ++t) {
                const uint32_t m = __ballot_sync(0xffffffffu, type == t);
                if (!m) continue;
                uint32_t qb = 0; const int leader = __ffs(m) - 1;
                if ((int) lane == leader) qb = atomicAdd(&ctr->qcount[par][t], (uint32_t) __popc(m));
                qb = __shfl_sync(0xffffffffu, qb, leader);
                if (type == t) pool.queue[t][qb + __popc(m & ((1u << lane) - 1u))] = slot;
            }
        }
    }
    warpAdd(&ctr->rays_ext, nRays); warpAdd(&ctr->done, nDone);
    if (COUNT) { warpAdd(&ctr->nodes_ext, cnt.nodes); warpAdd(&ctr->prims_ext, cnt.prims); }
}

// ------------------------------------------------------------------------------ shade
// One queue entry: the whole loop body of PathMisIntegrator::Li for one path vertex, INCLUDING the
// any-hit query of the NEE shadow ray (path_mis.cpp:48).  Tracing the shadow ray here, in the thread
// that just built it, keeps the ray, its pending contribution and the roulette decision in registers:
// no shadow-ray record is written to the pool and no separate pass re-reads the path state
// (measured: shade + shadow went from 266 ms to the fused number in DESIGN.md on the Cornell box).
template <int BSDF, bool MIS, bool COUNT>
__device__ __forceinline__ void shadeSlot(const DScene &sc, const Pool &pool, const Batch &bt, Counters *ctr, uint32_t slot,
                                          uint32_t &nDone, uint32_t &nShadow, TraceCounters &cnt) {
    const float4 ro = pool.rayO[slot], rd = pool.rayD[slot], hh = pool.hit[slot], th = pool.thr[slot], ra = pool.rad[slot];
    const uint32_t sid = pool.sid[slot];
    PathState st;
    st.o = mk(ro.x, ro.y, ro.z); st.d = mk(rd.x, rd.y, rd.z);
    st.thr = mk(th.x, th.y, th.z); st.pdf_mat = th.w; st.rad = mk(ra.x, ra.y, ra.z);
    st.flags = pool.flags[slot];
    st.rng.state = pool.rng[slot]; st.rng.inc = ((uint64_t) (sid % bt.wh) << 1u) | 1u;
    Hit h; h.t = hh.x; h.u = hh.y; h.v = hh.z; h.leafpos = __float_as_uint(hh.w);
    VertexOut out;
    pathVertex<BSDF, MIS>(sc, h, st, out);
    if (MIS) {                                                  // scene->rayIntersect(eRec.shadowRay), path_mis.cpp:48
        Hit sh; ++nShadow;
        if (!traverse<true, COUNT>(sc, out.shadow.o, out.shadow.d, out.shadow.mint, out.shadow.maxt, sh, cnt))
            st.rad = st.rad + out.contrib;
    }
    if (st.flags & PF_ALIVE) {
        pool.rayO[slot] = make_float4(out.next.o.x, out.next.o.y, out.next.o.z, out.next.mint);
        pool.rayD[slot] = make_float4(out.next.d.x, out.next.d.y, out.next.d.z, out.next.maxt);
        pool.thr[slot] = make_float4(st.thr.x, st.thr.y, st.thr.z, st.pdf_mat);
        pool.rng[slot] = st.rng.state;
        if (st.rad.x != ra.x || st.rad.y != ra.y || st.rad.z != ra.z) pool.rad[slot] = make_float4(st.rad.x, st.rad.y, st.rad.z, 0.f);
        pool.flags[slot] = st.flags & (PF_ALIVE | PF_DISCRETE);
    } else {                                                    // Russian roulette ended the path
        finalizePath(bt, ctr, sid, st.rad);
        pool.sid[slot] = NORI_FREE_SLOT; pool.flags[slot] = 0u; ++nDone;
    }
}

// All material queues in ONE launch: the queues are concatenated (diffuse | mirror | dielectric |
// microfacet | disney) and work item i belongs to the queue whose range contains it, so warps are
// material-coherent except where a boundary falls inside one.
#ifndef NORI_SHADE_MINBLOCKS
#define NORI_SHADE_MINBLOCKS 6
#endif
template <bool MIS, bool COUNT>
__global__ void __launch_bounds__(128, NORI_SHADE_MINBLOCKS) k_shade(DScene sc, Pool pool, Batch bt, Counters *ctr, uint32_t it) {
    uint32_t off[NORI_BSDF_COUNT + 1]; off[0] = 0;
#pragma unroll
    for (int t = 0; t < NORI_BSDF_COUNT; ++t) off[t + 1] = off[t] + ctr->qcount[it & 1u][t];
    const uint32_t n = off[NORI_BSDF_COUNT];
    const uint32_t stride = gridDim.x * blockDim.x;
    uint32_t nDone = 0, nShadow = 0; TraceCounters cnt; cnt.nodes = 0; cnt.prims = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (i < off[1]) shadeSlot<NORI_BSDF_DIFFUSE, MIS, COUNT>(sc, pool, bt, ctr, pool.queue[0][i], nDone, nShadow, cnt);
        else if (i < off[2]) shadeSlot<NORI_BSDF_MIRROR, MIS, COUNT>(sc, pool, bt, ctr, pool.queue[1][i - off[1]], nDone, nShadow, cnt);
        else if (i < off[3]) shadeSlot<NORI_BSDF_DIELECTRIC, MIS, COUNT>(sc, pool, bt, ctr, pool.queue[2][i - off[2]], nDone, nShadow, cnt);
        else if (i < off[4]) shadeSlot<NORI_BSDF_MICROFACET, MIS, COUNT>(sc, pool, bt, ctr, pool.queue[3][i - off[3]], nDone, nShadow, cnt);
        else shadeSlot<NORI_BSDF_DISNEY, MIS, COUNT>(sc, pool, bt, ctr, pool.queue[4][i - off[4]], nDone, nShadow, cnt);
    }
    warpAdd(&ctr->done, nDone);
    if (MIS) warpAdd(&ctr->rays_sh, nShadow);
    if (COUNT) { warpAdd(&ctr->nodes_sh, cnt.nodes); warpAdd(&ctr->prims_sh, cnt.prims); }
}

// ------------------------------------------------------------------------------ large-scene trace kernels
// On scenes with deep trees the rays of one warp need very different numbers of node visits (10 M
// triangles: 158 on average, long-tailed), and a leaf costs several times an inner node.  Run as plain
// per-lane loops that leaves a warp at ~5 of 32 active lanes (ncu, profiles/).  These variants keep the
// SAME per-ray visiting order (so results and counters stay the reference's) but schedule the warp as a
// small state machine:
//   * every lane is IDLE, at a NODE (one box test pending) or in a LEAF (one primitive test pending);
//   * each warp step runs EITHER the node code for all NODE lanes OR the primitive code for all LEAF
//     lanes -- the primitive phase is entered once NORI_LEAF_MIN lanes wait in a leaf (or nothing else
//     is runnable), so both code paths execute with many lanes active;
//   * IDLE lanes are refilled from the warp's slot chunk as soon as NORI_REFILL_MIN lanes are idle, so
//     short rays do not wait for the longest ray of the warp.
#define NORI_LEAF_MIN 16
#define NORI_REFILL_MIN 8
#ifndef NORI_NODE_BURST
#define NORI_NODE_BURST 1
#endif
enum { ST_IDLE = 0, ST_NODE = 1, ST_LEAF = 2, ST_DONE = 3 };

struct LaneTrav {
    RayTrav r;
    uint32_t st, leafI, leafEnd, slot;
};

// node phase for one lane: returns true when the ray is finished
template <bool COUNT>
__device__ __forceinline__ bool smNode(const DScene &sc, LaneTrav &L, uint32_t *stack, TraceCounters &cnt) {
    RayTrav &r = L.r;
    const uint4 n0 = __ldg(&sc.nodes[2 * r.node]);
    const uint4 n1 = __ldg(&sc.nodes[2 * r.node + 1]);
    if (COUNT) ++cnt.nodes;
    float nearT = __int_as_float(0xff800000), farT = __int_as_float(0x7f800000);
    const bool in = slab(r.o.x, r.d.x, r.rcp.x, __uint_as_float(n0.z), __uint_as_float(n1.y), nearT, farT)
                 && slab(r.o.y, r.d.y, r.rcp.y, __uint_as_float(n0.w), __uint_as_float(n1.z), nearT, farT)
                 && slab(r.o.z, r.d.z, r.rcp.z, __uint_as_float(n1.x), __uint_as_float(n1.w), nearT, farT)
                 && (r.mint <= farT && nearT <= r.maxt);
    if (in) {
        if (!(n0.x & 1u)) { stack[r.sp++] = n0.y; ++r.node; return false; }
        const uint32_t size = n0.x >> 1;
        if (size) { L.st = ST_LEAF; L.leafI = n0.y; L.leafEnd = n0.y + size; return false; }
    }
    if (r.sp == 0) return true;
    r.node = stack[--r.sp];
    return false;
}

// primitive phase for one lane: one primitive test; returns true when the ray is finished
template <bool SHADOW, bool COUNT>
__device__ __forceinline__ bool smPrim(const DScene &sc, LaneTrav &L, uint32_t *stack, TraceCounters &cnt) {
    RayTrav &r = L.r;
    const uint32_t i = L.leafI;
    const float4 r0 = __ldg(&sc.prims[3 * i]);
    const float4 r1 = __ldg(&sc.prims[3 * i + 1]);
    const float4 r2 = __ldg(&sc.prims[3 * i + 2]);
    if (COUNT) ++cnt.prims;
    float u = 0.f, v = 0.f, t;
    bool h;
    if (__float_as_uint(r2.w) == 0u)
        h = triTest(mk(r0.x, r0.y, r0.z), mk(r1.x, r1.y, r1.z), mk(r2.x, r2.y, r2.z), r.o, r.d, r.mint, r.maxt, u, v, t);
    else
        h = sphereTest(mk(r0.x, r0.y, r0.z), r1.x, r.o, r.d, r.mint, r.maxt, t);
    if (h) {
        r.found = true;
        if (SHADOW) { r.hit.t = 0.f; return true; }
        r.maxt = t; r.hit.t = t; r.hit.u = u; r.hit.v = v; r.hit.leafpos = i;
    }
    if (++L.leafI < L.leafEnd) return false;
    L.st = ST_NODE;
    if (r.sp == 0) return true;
    r.node = stack[--r.sp];
    return false;
}

template <bool COUNT, bool VOL>
__global__ void __launch_bounds__(128) k_extend_sm(DScene sc, Pool pool, Batch bt, Counters *ctr, uint32_t it) {
    __shared__ uint32_t s_free[4][NORI_FETCH];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, par = it & 1u;
    const uint32_t ltMask = (1u << lane) - 1u;
    uint32_t *freeList = s_free[warp];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int i = 0; i < NORI_NQ; ++i) ctr->qcount[par ^ 1u][i] = 0;
        ctr->work_extend[par ^ 1u] = 0;
    }
    const unsigned long long total = ctr->total_samples;
    uint32_t nRays = 0, nDone = 0; TraceCounters cnt; cnt.nodes = 0; cnt.prims = 0;
    uint32_t stack[64];
    LaneTrav L; L.st = ST_IDLE; L.slot = 0; L.leafI = L.leafEnd = 0;
    uint32_t chunkBase = 0, chunkNext = NORI_FETCH;
    bool moreChunks = true;
    while (true) {
        // ---- refill
        uint32_t idle = __ballot_sync(0xffffffffu, L.st == ST_IDLE);
        if (idle && (moreChunks || chunkNext < NORI_FETCH) && (__popc(idle) >= NORI_REFILL_MIN || idle == 0xffffffffu)) {
            while (idle) {
                if (chunkNext >= NORI_FETCH) {
                    if (!moreChunks) break;
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(&ctr->work_extend[par], NORI_FETCH);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (base >= pool.P) { moreChunks = false; break; }
                    chunkBase = base; chunkNext = 0;
                    uint32_t nFree = 0;                          // compacted regeneration (render.cpp:98-124)
                    for (uint32_t round = 0; round < NORI_FETCH / 32u; ++round) {
                        const uint32_t s = base + round * 32u + lane;
                        const bool isFree = s < pool.P && pool.sid[s] == NORI_FREE_SLOT;
                        const uint32_t m = __ballot_sync(0xffffffffu, isFree);
                        if (isFree) freeList[nFree + __popc(m & ltMask)] = s;
                        nFree += __popc(m);
                    }
                    if (nFree) {
                        unsigned long long first = 0;
                        if (lane == 0) first = atomicAdd(&ctr->next_sample, (unsigned long long) nFree);
                        first = __shfl_sync(0xffffffffu, first, 0);
                        __syncwarp();
                        for (uint32_t j = lane; j < nFree; j += 32u) {
                            const unsigned long long id = first + j;
                            if (id >= total) break;
                            const uint32_t s = freeList[j];
                            Ray ray; uint64_t rs;
                            generatePath(sc, bt, (uint32_t) id, ray, rs);
                            pool.rayO[s] = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.mint);
                            pool.rayD[s] = make_float4(ray.d.x, ray.d.y, ray.d.z, ray.maxt);
                            pool.thr[s] = make_float4(1.f, 1.f, 1.f, 0.f);
                            pool.rad[s] = make_float4(0.f, 0.f, 0.f, 0.f);
                            pool.rng[s] = rs; pool.sid[s] = (uint32_t) id;
                            pool.flags[s] = PF_ALIVE | PF_FIRST;
                        }
                        __syncwarp();
                    }
                }
                const uint32_t idx = chunkNext + __popc(idle & ltMask);
                const bool take = L.st == ST_IDLE && idx < NORI_FETCH;
                chunkNext = min(chunkNext + (uint32_t) __popc(idle), NORI_FETCH);
                if (take) {
                    const uint32_t s = chunkBase + idx;
                    if (s < pool.P && (pool.flags[s] & PF_ALIVE)) {
                        const float4 ro = pool.rayO[s], rd = pool.rayD[s];
                        L.slot = s; ++nRays;
                        if (travInit(sc, L.r, mk(ro.x, ro.y, ro.z), mk(rd.x, rd.y, rd.z), ro.w, rd.w)) L.st = ST_NODE;
                        else { L.r.found = false; L.st = ST_DONE; }  // decided before the first node: a miss
                    }
                }
                idle = __ballot_sync(0xffffffffu, L.st == ST_IDLE);
            }
        }
        // ---- pick the phase
        const uint32_t mNode = __ballot_sync(0xffffffffu, L.st == ST_NODE);
        const uint32_t mLeaf = __ballot_sync(0xffffffffu, L.st == ST_LEAF);
        if (!(mNode | mLeaf)) { if (!moreChunks && chunkNext >= NORI_FETCH) break; else continue; }
        bool finished = false;
        if (mLeaf && (__popc(mLeaf) >= NORI_LEAF_MIN || !mNode)) {
            if (L.st == ST_LEAF) finished = smPrim<false, COUNT>(sc, L, stack, cnt);
        } else {
            // several node visits per scheduling decision: lanes that reach a leaf or finish sit out the rest
#pragma unroll 1
            for (int k = 0; k < NORI_NODE_BURST; ++k)
                if (L.st == ST_NODE && !finished) finished = smNode<COUNT>(sc, L, stack, cnt);
        }
        // ---- publish finished rays, bin hits by material (one atomic per warp and material)
        finished = finished || L.st == ST_DONE;
        if (__any_sync(0xffffffffu, finished)) {
            int type = -1;
            if (finished) {
                L.st = ST_IDLE;
                if (L.r.found) {
                    pool.hit[L.slot] = make_float4(L.r.hit.t, L.r.hit.u, L.r.hit.v, __uint_as_float(L.r.hit.leafpos));
                    type = sc.shapes[__float_as_uint(__ldg(&sc.prims[3 * L.r.hit.leafpos + 1]).w)].bsdf_type;
                } else type = missRule<VOL>(sc, pool, bt, ctr, L.slot, L.r.o, L.r.d, nDone);
            }
#pragma unroll
            for (int t = 0; t < (VOL ? NORI_NQ : NORI_BSDF_COUNT); ++t) {
                const uint32_t m = __ballot_sync(0xffffffffu, type == t);
                if (!m) continue;
                uint32_t qb = 0; const int leader = __ffs(m) - 1;
                if ((int) lane == leader) qb = atomicAdd(&ctr->qcount[par][t], (uint32_t) __popc(m));
                qb = __shfl_sync(0xffffffffu, qb, leader);
                if (type == t) pool.queue[t][qb + __popc(m & ltMask)] = L.slot;
            }
        }
    }
    warpAdd(&ctr->rays_ext, nRays); warpAdd(&ctr->done, nDone);
    if (COUNT) { warpAdd(&ctr->nodes_ext, cnt.nodes); warpAdd(&ctr->prims_ext, cnt.prims); }
}

// ------------------------------------------------------------------------------ short integrators
template <bool COUNT>
__global__ void __launch_bounds__(128) k_mega(DScene sc, Batch bt, Counters *ctr, unsigned long long total) {
    const unsigned long long id = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
    RayStats rs; rs.rays = 0; rs.shadow = 0; rs.cnt.nodes = 0; rs.cnt.prims = 0;
    if (id < total) {
        const uint32_t sid = (uint32_t) id;
        const uint32_t k = sid / bt.wh, pix = sid - k * bt.wh;
        const int W = sc.camera.width;
        const int py = pix / W, px = pix - py * W;
        Pcg32 rng; rng.seed(bt.seed + bt.spp_first + k, (uint64_t) pix);
        P2 a = rng.next2D();
        P2 ps; ps.x = (float) px + a.x; ps.y = (float) py + a.y;
        P2 ap = rng.next2D();
        Ray ray = cameraRay(sc.camera, ps, ap);
        V3 L = liDispatch<COUNT>(sc, rng, ray, rs);
        finalizePath(bt, ctr, sid, L);
    }
    warpAdd(&ctr->rays_ext, rs.rays - rs.shadow); warpAdd(&ctr->rays_sh, rs.shadow);
    if (COUNT) { warpAdd(&ctr->nodes_ext, rs.cnt.nodes); warpAdd(&ctr->prims_ext, rs.cnt.prims); }
}

// ------------------------------------------------------------------------------ film
struct FilmParams {
    float4 *film;             // (H+2b) x (W+2b) row-major (r,g,b,w): the memory image of ImageBlock m_block
    int W, H, border, halo;   // halo == border: source pixels that can reach a film pixel
    float radius, lookupFactor;
    float table[NORI_FILTER_RESOLUTION + 1];
    float4 *vsum, *vsum2;     // per-pixel sums of the running mean and of its square (render.cpp:238-247), or NULL
};

// smem per layer: (32+2*halo)^2 x { float4 value, float2 pos }
// VARIANCE: also accumulate, after every spp layer (= one pass of the reference's spp-major loop), the
// running mean m_k = rgb/w of the pixel and its square -- the reference's `*_variance.exr` statistic
// (render.cpp:225,238-247,263-278).  The accumulation then starts from the film's current
// value so that m_k covers every pass rendered so far, across batches and render() calls.
template <bool VARIANCE>
__global__ void __launch_bounds__(1024) k_film(FilmParams fp, Batch bt, uint32_t nLayers) {
    extern __shared__ float4 s_mem[];
    const int T = 32, halo = fp.halo, S = T + 2 * halo, nS = S * S;
    float4 *s_val = s_mem;
    float2 *s_pos = (float2 *) (s_mem + nS);
    __shared__ float s_table[NORI_FILTER_RESOLUTION + 1];
    const int tid = threadIdx.y * T + threadIdx.x;
    if (tid <= NORI_FILTER_RESOLUTION) s_table[tid] = fp.table[tid];
    const int b = fp.border;
    const int fx = blockIdx.x * T + threadIdx.x, fy = blockIdx.y * T + threadIdx.y;   // film pixel owned by this thread
    const int sx0 = blockIdx.x * T - b - halo, sy0 = blockIdx.y * T - b - halo;       // image coords of the staged region
    const int fcols = fp.W + 2 * b, frows = fp.H + 2 * b;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float3 vs = make_float3(0.f, 0.f, 0.f), vs2 = make_float3(0.f, 0.f, 0.f);
    const bool owner = fx < fcols && fy < frows;
    if (VARIANCE && owner) acc = fp.film[(size_t) fy * fcols + fx];
    for (uint32_t k = 0; k < nLayers; ++k) {
        __syncthreads();
        for (int i = tid; i < nS; i += T * T) {
            const int ly = i / S, lx = i - ly * S;
            const int sx = sx0 + lx, sy = sy0 + ly;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f); float2 p = make_float2(0.f, 0.f);
            if (sx >= 0 && sx < fp.W && sy >= 0 && sy < fp.H) {
                const uint32_t pix = (uint32_t) sy * fp.W + sx;
                v = bt.results[(size_t) k * bt.wh + pix];
                Pcg32 rng; rng.seed(bt.seed + bt.spp_first + k, (uint64_t) pix);
                P2 a = rng.next2D();
                const float psx = (float) sx + a.x, psy = (float) sy + a.y;
                // block.cpp:101-104 with the offset of the 32x32 block that rendered the sample
                const int ox = sx & ~(NORI_BLOCK_SIZE - 1), oy = sy & ~(NORI_BLOCK_SIZE - 1);
                p.x = __fsub_rn(__fsub_rn(psx, 0.5f), (float) (ox - b));
                p.y = __fsub_rn(__fsub_rn(psy, 0.5f), (float) (oy - b));
            }
            s_val[i] = v; s_pos[i] = p;
        }
        __syncthreads();
        if (owner) {
            // staged-region coordinates of the source pixels that can reach (fx, fy)
            const int cx = threadIdx.x + halo, cy = threadIdx.y + halo;   // own source pixel (image x = fx - b)
            for (int dy = -halo; dy <= halo; ++dy) {
                const int sy = sy0 + cy + dy;
                if (sy < 0 || sy >= fp.H) continue;
                const int oy = sy & ~(NORI_BLOCK_SIZE - 1);
                const float yb = (float) (fy - oy);                        // pixel row in that block's coordinates
                for (int dx = -halo; dx <= halo; ++dx) {
                    const int sx = sx0 + cx + dx;
                    if (sx < 0 || sx >= fp.W) continue;
                    const int ox = sx & ~(NORI_BLOCK_SIZE - 1);
                    const float xb = (float) (fx - ox);
                    const int i = (cy + dy) * S + (cx + dx);
                    const float2 p = s_pos[i];
                    // window [ceil(p-r), floor(p+r)] of block.cpp:107-110
                    if (xb < __fsub_rn(p.x, fp.radius) || xb > __fadd_rn(p.x, fp.radius)) continue;
                    if (yb < __fsub_rn(p.y, fp.radius) || yb > __fadd_rn(p.y, fp.radius)) continue;
                    const float wx = s_table[(int) __fmul_rn(fabsf(__fsub_rn(xb, p.x)), fp.lookupFactor)];
                    const float wy = s_table[(int) __fmul_rn(fabsf(__fsub_rn(yb, p.y)), fp.lookupFactor)];
                    const float4 v = s_val[i];
                    // Color4f(value) * wx * wy (block.cpp:121)
                    acc.x = __fadd_rn(acc.x, __fmul_rn(__fmul_rn(v.x, wx), wy));
                    acc.y = __fadd_rn(acc.y, __fmul_rn(__fmul_rn(v.y, wx), wy));
                    acc.z = __fadd_rn(acc.z, __fmul_rn(__fmul_rn(v.z, wx), wy));
                    acc.w = __fadd_rn(acc.w, __fmul_rn(__fmul_rn(v.w, wx), wy));
                }
            }
            if (VARIANCE) {                                      // Color4f::divideByFilterWeight (color.h:84-89)
                const float mx = acc.w != 0.f ? acc.x / acc.w : 0.f, my = acc.w != 0.f ? acc.y / acc.w : 0.f, mz = acc.w != 0.f ? acc.z / acc.w : 0.f;
                vs.x += mx; vs.y += my; vs.z += mz;
                vs2.x += mx * mx; vs2.y += my * my; vs2.z += mz * mz;
            }
        }
    }
    if (owner) {
        float4 *dst = &fp.film[(size_t) fy * fcols + fx];
        if (VARIANCE) {
            *dst = acc;
            float4 a = fp.vsum[(size_t) fy * fcols + fx], b = fp.vsum2[(size_t) fy * fcols + fx];
            a.x += vs.x; a.y += vs.y; a.z += vs.z; b.x += vs2.x; b.y += vs2.y; b.z += vs2.z;
            fp.vsum[(size_t) fy * fcols + fx] = a; fp.vsum2[(size_t) fy * fcols + fx] = b;
        } else {
            float4 f = *dst;
            f.x += acc.x; f.y += acc.y; f.z += acc.z; f.w += acc.w;
            *dst = f;
        }
    }
}

// var = sum2/N - (sum/N)^2 per channel (render.cpp:268-275)
__global__ void k_variance(const float4 *vsum, const float4 *vsum2, float *rgb, int W, int H, int b, float n) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const float4 s = vsum[(size_t) (y + b) * (W + 2 * b) + (x + b)], s2 = vsum2[(size_t) (y + b) * (W + 2 * b) + (x + b)];
    float *o = &rgb[((size_t) y * W + x) * 3];
    const float mx = s.x / n, my = s.y / n, mz = s.z / n;
    o[0] = s2.x / n - mx * mx; o[1] = s2.y / n - my * my; o[2] = s2.z / n - mz * mz;
}

__global__ void k_resolve(const float4 *film, float *rgb, int W, int H, int b) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const float4 c = film[(size_t) (y + b) * (W + 2 * b) + (x + b)];
    float *o = &rgb[((size_t) y * W + x) * 3];
    if (c.w != 0.f) { o[0] = c.x / c.w; o[1] = c.y / c.w; o[2] = c.z / c.w; } else { o[0] = o[1] = o[2] = 0.f; }
}

// ------------------------------------------------------------------------------ test hooks
template <bool SHADOW>
__global__ void __launch_bounds__(128) k_trace(DScene sc, const nori_gpu_ray *rays, unsigned long long n, nori_gpu_hit *out) {
    const unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const nori_gpu_ray r = rays[i];
    Hit h; TraceCounters cnt; cnt.nodes = 0; cnt.prims = 0;
    bool found = traverse<SHADOW, true>(sc, mk(r.o[0], r.o[1], r.o[2]), mk(r.d[0], r.d[1], r.d[2]), r.mint, r.maxt, h, cnt);
    nori_gpu_hit o;
    o.t = h.t; o.u = h.u; o.v = h.v; o.shape = NORI_NO_HIT; o.prim = NORI_NO_HIT;
    o.nodes_visited = cnt.nodes; o.prims_tested = cnt.prims; o.reserved = 0;
    if (found && !SHADOW) {
        o.prim = __float_as_uint(sc.prims[3 * h.leafpos].w);
        o.shape = __float_as_uint(sc.prims[3 * h.leafpos + 1].w);
    }
    out[i] = o;
}

__global__ void k_pcg32(uint64_t initstate, uint64_t initseq, unsigned long long n, float *outf, uint32_t *outu) {
    if (blockIdx.x || threadIdx.x) return;
    Pcg32 r; r.seed(initstate, initseq);
    for (unsigned long long i = 0; i < n; ++i) { if (outf) outf[i] = r.nextFloat(); else outu[i] = r.nextUInt(); }
}

// per-function probes (rows as in oracle/ref_tools/nori_export.cpp --probe):
//   bsdf    in (wi.xyz, wo.xyz, uv.xy, sample.xy)   out (eval.rgb, pdf, weight.rgb, wo.xyz, measure, pdf(sampled))
//   emitter in (ref.xyz, sample.xy)                 out (Li.rgb, wi.xyz, pdf, mint, maxt, p.xyz, eval.rgb)
__global__ void k_probe_bsdf(DScene sc, uint32_t bsdf, unsigned long long n, const float *in, float *out) {
    const unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const nori_gpu_bsdf &b = sc.bsdfs[bsdf];
    const float *q = &in[10 * i]; float *o = &out[12 * i];
    BRec e; e.wi = mk(q[0], q[1], q[2]); e.wo = mk(q[3], q[4], q[5]); e.measure = M_SOLID_ANGLE; e.uv.x = q[6]; e.uv.y = q[7];
    V3 ev = bsdfEvalDyn(b, e); float pdf = bsdfPdfDyn(b, e);
    BRec r; r.wi = e.wi; r.measure = M_UNKNOWN; r.uv = e.uv; P2 s; s.x = q[8]; s.y = q[9];
    V3 w = bsdfSampleDyn(b, r, s); float pdf2 = bsdfPdfDyn(b, r);
    o[0] = ev.x; o[1] = ev.y; o[2] = ev.z; o[3] = pdf; o[4] = w.x; o[5] = w.y; o[6] = w.z;
    o[7] = r.wo.x; o[8] = r.wo.y; o[9] = r.wo.z; o[10] = (float) r.measure; o[11] = pdf2;
}
__global__ void k_probe_emitter(DScene sc, uint32_t emitter, unsigned long long n, const float *in, float *out) {
    const unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const nori_gpu_emitter &em = sc.emitters[emitter].pod;
    const float *q = &in[5 * i]; float *o = &out[15 * i];
    ERec e = makeERec(mk(q[0], q[1], q[2])); P2 s; s.x = q[3]; s.y = q[4];
    e.shadow = mkray(e.ref, mk(0.f));
    V3 Li = emitterSample(sc, em, e, s); float pdf = emitterPdf(sc, em, e); V3 ev = emitterEval(sc, em, e);
    o[0] = Li.x; o[1] = Li.y; o[2] = Li.z; o[3] = e.wi.x; o[4] = e.wi.y; o[5] = e.wi.z; o[6] = pdf;
    o[7] = e.shadow.mint; o[8] = e.shadow.maxt; o[9] = e.p.x; o[10] = e.p.y; o[11] = e.p.z;
    o[12] = ev.x; o[13] = ev.y; o[14] = ev.z;
}

__global__ void k_fill_u32(uint32_t *p, uint32_t v, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void k_flush(float4 *buf, size_t n) {     // bench helper: evict L2 between timed steps
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t) gridDim.x * blockDim.x;
    for (; i < n; i += stride) buf[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}
