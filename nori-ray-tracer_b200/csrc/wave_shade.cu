// wave_shade.cu -- k_shade, compiled once per integrator mode (-DNORI_SHADE_MODE=0|1|2: path_mats,
// path_mis, volumetric) so the three variants build in parallel.
#include "kernels.cuh"

// ------------------------------------------------------------------------------ shade
// One queue entry: the whole loop body of PathMisIntegrator::Li for one path vertex, INCLUDING the
// any-hit query of the NEE shadow ray (path_mis.cpp:48).  Tracing the shadow ray here, in the thread
// that just built it, keeps the ray, its pending contribution and the roulette decision in registers:
// no shadow-ray record is written to the pool and no separate pass re-reads the path state
// (measured: shade + shadow went from 266 ms to the fused number in DESIGN.md on the Cornell box).

// DEFER (path_mis on large scenes): the shadow ray is not traced here but written to the pool (direction
// + far end; it starts at the next ray's origin with mint = Epsilon, arealight.cpp:56) together with its
// pending contribution, and k_shadow_sm traces it with the warp state machine -- on deep trees a
// plain per-lane loop inside this kernel leaves most lanes idle.
template <int BSDF, int MODE, bool COUNT, bool DEFER, bool AO>
__device__ __forceinline__ void shadeSlot(const DScene &sc, const Pool &pool, const Batch &bt, Counters *ctr, uint32_t slot,
                                          uint32_t &nDone, uint32_t &nShadow, uint32_t &nClosest, TraceCounters &cnt) {
    NORI_CHECK(slot < pool.P);
    const float4 ro = pool.rayO[slot], rd = pool.rayD[slot], hh = pool.hit[slot], th = pool.thr[slot], ra = pool.rad[slot];
    const uint32_t sid = pool.sid[slot];
    NORI_CHECK(sid != NORI_FREE_SLOT);
    PathState st;
    st.o = mk(ro.x, ro.y, ro.z); st.d = mk(rd.x, rd.y, rd.z);
    st.thr = mk(th.x, th.y, th.z); st.pdf_mat = th.w; st.rad = mk(ra.x, ra.y, ra.z);
    st.flags = pool.flags[slot];
    const uint32_t chBits = st.flags & PF_CH_MASK;
    st.rng.state = pool.rng[slot]; st.rng.inc = ((uint64_t) (sid % bt.wh) << 1u) | 1u;
    Hit h; h.t = hh.x; h.u = hh.y; h.v = hh.z; h.leafpos = __float_as_uint(hh.w);
    Ray next;
    if constexpr (MODE == MODE_VOL) {
        volVertex<COUNT>(sc, h, st, next, nClosest, nShadow, cnt);
    } else {
        VertexOut out;
        pathVertex<BSDF, MODE == MODE_MIS, AO>(sc, h, st, out);
        // a contribution of exactly zero needs no shadow ray (integrators.cuh: nullContribution)
        const bool needShadow = COUNT || !nullContribution(out.contrib.x, out.contrib.y, out.contrib.z);
        if (DEFER) {
            if (needShadow) {
                pool.shD[slot] = make_float4(out.shadow.d.x, out.shadow.d.y, out.shadow.d.z, out.shadow.maxt);
                pool.shC[slot] = make_float4(out.contrib.x, out.contrib.y, out.contrib.z, 0.f);
                if (!(st.flags & PF_ALIVE)) {                       // the roulette ended the path: k_shadow_sm finalises it
                    pool.rayO[slot] = make_float4(out.shadow.o.x, out.shadow.o.y, out.shadow.o.z, NORI_EPS);
                    if (st.rad.x != ra.x || st.rad.y != ra.y || st.rad.z != ra.z) pool.rad[slot] = make_float4(st.rad.x, st.rad.y, st.rad.z, 0.f);
                    pool.rng[slot] = st.rng.state;
                    pool.flags[slot] = PF_SHADOW | PF_TERMINATE | chBits;
                    return;
                }
            } else st.flags &= ~(uint32_t) PF_SHADOW;               // nothing for k_shadow_sm; an ended path is finalised below
        } else if (MODE == MODE_MIS && needShadow) {                // scene->rayIntersect(eRec.shadowRay), path_mis.cpp:48
            Hit sh; ++nShadow;
            if (!traverse<true, COUNT>(sc, out.shadow.o, out.shadow.d, out.shadow.mint, out.shadow.maxt, sh, cnt))
                st.rad = st.rad + out.contrib;
        }
        next = out.next;
    }
    if (st.flags & PF_ALIVE) {
        pool.rayO[slot] = make_float4(next.o.x, next.o.y, next.o.z, next.mint);
        pool.rayD[slot] = make_float4(next.d.x, next.d.y, next.d.z, next.maxt);
        pool.thr[slot] = make_float4(st.thr.x, st.thr.y, st.thr.z, st.pdf_mat);
        pool.rng[slot] = st.rng.state;
        if (st.rad.x != ra.x || st.rad.y != ra.y || st.rad.z != ra.z) pool.rad[slot] = make_float4(st.rad.x, st.rad.y, st.rad.z, 0.f);
        pool.flags[slot] = (st.flags & (PF_ALIVE | PF_DISCRETE | (DEFER ? PF_SHADOW : 0u))) | chBits;
    } else endOfPath(sc, pool, bt, ctr, slot, sid, st.rad, st.rng.state, chBits, nDone);   // Russian roulette ended the path
}

// All material queues in ONE launch: the queues are concatenated (diffuse | mirror | dielectric |
// microfacet | disney | [volumetric: misses]) and work item i belongs to the queue whose range contains
// it, so warps are material-coherent except where a boundary falls inside one.
#ifndef NORI_SHADE_TEMPLATED
#define NORI_SHADE_TEMPLATED 0
#endif
#ifndef NORI_SHADE_THREADS
#define NORI_SHADE_THREADS 128
#endif
#ifndef NORI_SHADE_MINBLOCKS
#define NORI_SHADE_MINBLOCKS 8
#endif
template <int MODE, bool COUNT, bool DEFER, bool ESORT, bool AO>
__global__ void __launch_bounds__(NORI_SHADE_THREADS, NORI_SHADE_MINBLOCKS) k_shade(DScene sc, Pool pool, Batch bt, Counters *ctr, uint32_t it) {
    // work list = the concatenated queues: by material (NQ queues), or -- ESORT: path_mis with emitters of several
    // types, after k_rebin -- by (material, emitter type), so that the lanes of a warp also sample the same kind of light
    constexpr int NQ = ESORT ? NORI_NEQ : NORI_NQ;
    uint32_t off[NQ + 1]; off[0] = 0;
#pragma unroll
    for (int t = 0; t < NQ; ++t) off[t + 1] = off[t] + (ESORT ? ctr->eqcount[it & 1u][t] : ctr->qcount[it & 1u][t]);
    const uint32_t n = off[NQ];
    const uint32_t stride = gridDim.x * blockDim.x;
    uint32_t nDone = 0, nShadow = 0, nClosest = 0; TraceCounters cnt;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
#if NORI_SHADE_TEMPLATED && NORI_SHADE_MODE != 2
        if constexpr (ESORT) {
            int k = 0;
#pragma unroll
            for (int t = 1; t < NQ; ++t) k += i >= off[t];
            const uint32_t slot = pool.equeue[(size_t) k * pool.P + (i - off[k])];
            const int q = k >> 2;
            if (q == NORI_BSDF_DIFFUSE) shadeSlot<NORI_BSDF_DIFFUSE, MODE, COUNT, DEFER, AO>(sc, pool, bt, ctr, slot, nDone, nShadow, nClosest, cnt);
            else if (q == NORI_BSDF_MIRROR) shadeSlot<NORI_BSDF_MIRROR, MODE, COUNT, DEFER, AO>(sc, pool, bt, ctr, slot, nDone, nShadow, nClosest, cnt);
            else if (q == NORI_BSDF_DIELECTRIC) shadeSlot<NORI_BSDF_DIELECTRIC, MODE, COUNT, DEFER, AO>(sc, pool, bt, ctr, slot, nDone, nShadow, nClosest, cnt);
            else if (q == NORI_BSDF_MICROFACET) shadeSlot<NORI_BSDF_MICROFACET, MODE, COUNT, DEFER, AO>(sc, pool, bt, ctr, slot, nDone, nShadow, nClosest, cnt);
            else shadeSlot<NORI_BSDF_DISNEY, MODE, COUNT, DEFER, AO>(sc, pool, bt, ctr, slot, nDone, nShadow, nClosest, cnt);
        } else {
            if (i < off[1]) shadeSlot<NORI_BSDF_DIFFUSE, MODE, COUNT, DEFER, AO>(sc, pool, bt, ctr, pool.queue[0][i], nDone, nShadow, nClosest, cnt);
            else if (i < off[2]) shadeSlot<NORI_BSDF_MIRROR, MODE, COUNT, DEFER, AO>(sc, pool, bt, ctr, pool.queue[1][i - off[1]], nDone, nShadow, nClosest, cnt);
            else if (i < off[3]) shadeSlot<NORI_BSDF_DIELECTRIC, MODE, COUNT, DEFER, AO>(sc, pool, bt, ctr, pool.queue[2][i - off[2]], nDone, nShadow, nClosest, cnt);
            else if (i < off[4]) shadeSlot<NORI_BSDF_MICROFACET, MODE, COUNT, DEFER, AO>(sc, pool, bt, ctr, pool.queue[3][i - off[3]], nDone, nShadow, nClosest, cnt);
            else if (i < off[5]) shadeSlot<NORI_BSDF_DISNEY, MODE, COUNT, DEFER, AO>(sc, pool, bt, ctr, pool.queue[4][i - off[4]], nDone, nShadow, nClosest, cnt);
        }
#else
        // ONE copy of the vertex code for every material: the BSDF's eval / pdf / sample are reached through a
        // switch on the BSDF type, which is warp-uniform because the queues are sorted by type (volumetric mode:
        // a copy per BSDF type and vertex kind made the kernel instruction-fetch bound)
        int q = 0;
#pragma unroll
        for (int t = 1; t < NQ; ++t) q += i >= off[t];
        const uint32_t slot = ESORT ? pool.equeue[(size_t) q * pool.P + (i - off[q])] : pool.queue[q][i - off[q]];
        shadeSlot<-1, MODE, COUNT, DEFER, AO>(sc, pool, bt, ctr, slot, nDone, nShadow, nClosest, cnt);
#endif
    }
    warpAdd(&ctr->done, nDone);
    if (MODE != MODE_MATS && !DEFER) warpAdd(&ctr->rays_sh, nShadow);
    if (MODE == MODE_VOL) warpAdd(&ctr->rays_sh_closest, nClosest);
    if (COUNT) { warpAdd(&ctr->nodes_sh, cnt.nodes); warpAdd(&ctr->prims_sh, cnt.prims); }
}

#if NORI_SHADE_MODE == 0
#define LAUNCHER noriLaunchShadeMats
#elif NORI_SHADE_MODE == 1
#define LAUNCHER noriLaunchShadeMis
#else
#define LAUNCHER noriLaunchShadeVol
#endif
void LAUNCHER(bool count, int grid, cudaStream_t st, const DScene &sc, const Pool &pool, const Batch &bt, Counters *ctr, uint32_t it) {
#if NORI_SHADE_MODE == 1
    if (sc.area_only) {
        if (count) k_shade<MODE_MIS, true, false, false, true><<<grid * 128 / NORI_SHADE_THREADS, NORI_SHADE_THREADS, 0, st>>>(sc, pool, bt, ctr, it);
        else k_shade<MODE_MIS, false, false, false, true><<<grid * 128 / NORI_SHADE_THREADS, NORI_SHADE_THREADS, 0, st>>>(sc, pool, bt, ctr, it);
        return;
    }
    if (sc.esort) {
        if (count) k_shade<MODE_MIS, true, false, true, false><<<grid * 128 / NORI_SHADE_THREADS, NORI_SHADE_THREADS, 0, st>>>(sc, pool, bt, ctr, it);
        else k_shade<MODE_MIS, false, false, true, false><<<grid * 128 / NORI_SHADE_THREADS, NORI_SHADE_THREADS, 0, st>>>(sc, pool, bt, ctr, it);
        return;
    }
#endif
    if (count) k_shade<NORI_SHADE_MODE, true, false, false, false><<<grid * 128 / NORI_SHADE_THREADS, NORI_SHADE_THREADS, 0, st>>>(sc, pool, bt, ctr, it);
    else k_shade<NORI_SHADE_MODE, false, false, false, false><<<grid * 128 / NORI_SHADE_THREADS, NORI_SHADE_THREADS, 0, st>>>(sc, pool, bt, ctr, it);
}
#if NORI_SHADE_MODE == 1
// Emitter-sorted shading queues.  The light a path_mis vertex samples is picked by the FIRST random number of the
// vertex (path_mis.cpp:42), so it can be peeked from the path's stream before shading.  k_rebin splits every
// material queue k_extend built into (material, emitter type) sub-queues, one warp-aggregated atomic per
// distinct key (__match_any_sync).  Only launched when the scene has emitters of several types.
__global__ void __launch_bounds__(1024) k_rebin(DScene sc, Pool pool, Batch bt, Counters *ctr, uint32_t it) {
    // one work item per thread; counts are aggregated per warp (match_any) and per CTA (shared memory) so that
    // each CTA issues one global atomic per key: with one atomic per warp the ~20 hot counters serialised the pass
    __shared__ uint32_t s_cnt[NORI_NEQ], s_base[NORI_NEQ];
    const uint32_t par = it & 1u, lane = threadIdx.x & 31;
    uint32_t off[NORI_BSDF_COUNT + 1]; off[0] = 0;
#pragma unroll
    for (int t = 0; t < NORI_BSDF_COUNT; ++t) off[t + 1] = off[t] + ctr->qcount[par][t];
    const uint32_t n = off[NORI_BSDF_COUNT];
    if (blockIdx.x * blockDim.x >= n) return;                             // whole CTA past the end
    if (threadIdx.x < NORI_NEQ) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t key = 0xffffffffu, slot = 0;
    if (i < n) {
        int q = 0;
#pragma unroll
        for (int t = 1; t < NORI_BSDF_COUNT; ++t) q += i >= off[t];
        slot = pool.queue[q][i - off[q]];
        Pcg32 peek; peek.state = pool.rng[slot]; peek.inc = ((uint64_t) (pool.sid[slot] % bt.wh) << 1u) | 1u;
        key = (uint32_t) q * 4u + (uint32_t) sc.emitters[randomEmitter(sc, peek.next1D())].pod.type;
    }
    const uint32_t peers = __match_any_sync(0xffffffffu, key);
    const int leader = __ffs(peers) - 1;
    uint32_t pos = 0;
    if ((int) lane == leader && key != 0xffffffffu) pos = atomicAdd(&s_cnt[key], (uint32_t) __popc(peers));
    pos = __shfl_sync(0xffffffffu, pos, leader) + __popc(peers & ((1u << lane) - 1u));
    __syncthreads();
    if (threadIdx.x < NORI_NEQ && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(&ctr->eqcount[par][threadIdx.x], s_cnt[threadIdx.x]);
    __syncthreads();
    if (key != 0xffffffffu) pool.equeue[(size_t) key * pool.P + s_base[key] + pos] = slot;
}
void noriLaunchRebin(int grid, cudaStream_t st, const DScene &sc, const Pool &pool, const Batch &bt, Counters *ctr, uint32_t it) {
    k_rebin<<<grid, 1024, 0, st>>>(sc, pool, bt, ctr, it);
}
void noriLaunchShadeMisDeferred(bool count, int grid, cudaStream_t st, const DScene &sc, const Pool &pool, const Batch &bt, Counters *ctr, uint32_t it) {
    if (sc.area_only) {
        if (count) k_shade<MODE_MIS, true, true, false, true><<<grid * 128 / NORI_SHADE_THREADS, NORI_SHADE_THREADS, 0, st>>>(sc, pool, bt, ctr, it);
        else k_shade<MODE_MIS, false, true, false, true><<<grid * 128 / NORI_SHADE_THREADS, NORI_SHADE_THREADS, 0, st>>>(sc, pool, bt, ctr, it);
        return;
    }
    if (sc.esort) {
        if (count) k_shade<MODE_MIS, true, true, true, false><<<grid * 128 / NORI_SHADE_THREADS, NORI_SHADE_THREADS, 0, st>>>(sc, pool, bt, ctr, it);
        else k_shade<MODE_MIS, false, true, true, false><<<grid * 128 / NORI_SHADE_THREADS, NORI_SHADE_THREADS, 0, st>>>(sc, pool, bt, ctr, it);
        return;
    }
    if (count) k_shade<MODE_MIS, true, true, false, false><<<grid * 128 / NORI_SHADE_THREADS, NORI_SHADE_THREADS, 0, st>>>(sc, pool, bt, ctr, it);
    else k_shade<MODE_MIS, false, true, false, false><<<grid * 128 / NORI_SHADE_THREADS, NORI_SHADE_THREADS, 0, st>>>(sc, pool, bt, ctr, it);
}
#endif
