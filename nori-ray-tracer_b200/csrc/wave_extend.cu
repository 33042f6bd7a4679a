// wave_extend.cu -- k_extend / k_extend_sm: closest-hit traversal of the path pool fused with camera-path
// regeneration and material binning (see kernels.cuh for the overview).
#include "kernels.cuh"

#ifndef NORI_EXTEND_PIPELINE
#define NORI_EXTEND_PIPELINE 1   // k_extend requests the next round's flags and ray before it traverses the current one
#endif
#ifndef NORI_FETCH
#define NORI_FETCH 128u      // pool slots claimed per warp per atomic (4 rounds of 32); with two concurrent wavefronts, Cornell box:
#endif                       // 64: 214.2, 128: 205.7, 256: 207.2, 512: 209.3 ms; 128 + the pipeline: 204.5 ms (10 M triangles 400.5 -> 397.8)

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
template <bool VOL, bool NR = true>
__device__ __forceinline__ int missRule(const DScene &sc, const Pool &pool, const Batch &bt, Counters *ctr, uint32_t slot, V3 o, V3 d, uint32_t &nDone) {
    if (VOL) {
        float nearT, farT;
        if (boundsHit(sc.medium, o, d, nearT, farT)) {
            pool.hit[slot] = make_float4(__int_as_float(0x7f800000), 0.f, 0.f, __uint_as_float(NORI_NO_HIT));
            return NORI_Q_MISS;
        }
    }
    const float4 r = pool.rad[slot];
    endOfPath<NR>(sc, pool, bt, ctr, slot, pool.sid[slot], mk(r.x, r.y, r.z), pool.rng[slot], pool.flags[slot], nDone);
    return -1;
}

#ifndef NORI_EXTEND_MINBLOCKS
#define NORI_EXTEND_MINBLOCKS 8
#endif
template <bool COUNT, bool VOL>
__global__ void __launch_bounds__(128, NORI_EXTEND_MINBLOCKS) k_extend(DScene sc, Pool pool, Batch bt, Counters *ctr, uint32_t it) {
    __shared__ uint32_t s_free[4][NORI_FETCH];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, par = it & 1u;
    uint32_t *freeList = s_free[warp];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int i = 0; i < NORI_NQ; ++i) ctr->qcount[par ^ 1u][i] = 0;
        for (int i = 0; i < NORI_NEQ; ++i) ctr->eqcount[par ^ 1u][i] = 0;
        ctr->work_extend[par ^ 1u] = 0; ctr->work_shadow[par ^ 1u] = 0;
    }
    const unsigned long long total = ctr->total_samples;
    uint32_t nRays = 0, nDone = 0; TraceCounters cnt;
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
                generatePath(sc, bt, (uint32_t) id, 0, ray, rs);
                pool.rayO[slot] = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.mint);
                pool.rayD[slot] = make_float4(ray.d.x, ray.d.y, ray.d.z, ray.maxt);
                pool.thr[slot] = make_float4(1.f, 1.f, 1.f, 0.f);
                pool.rad[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
                pool.rng[slot] = rs; pool.sid[slot] = (uint32_t) id;
                pool.flags[slot] = PF_ALIVE | PF_FIRST;
            }
            __syncwarp();
        }
        // ---- phase 2: trace; the material of every hit is parked in the warp's shared list (free again after phase 1)
#if NORI_EXTEND_PIPELINE
        // software pipeline: the next round's flags and ray are requested before this round is traversed
        uint32_t nfl = (base + lane) < pool.P ? pool.flags[base + lane] : 0u;
        float4 nro = make_float4(0.f, 0.f, 0.f, 0.f), nrd = nro;
        if (nfl & PF_ALIVE) { nro = pool.rayO[base + lane]; nrd = pool.rayD[base + lane]; }
#endif
        for (uint32_t round = 0; round < NORI_FETCH / 32u; ++round) {
            const uint32_t slot = base + round * 32u + lane;
            int type = -1;
#if NORI_EXTEND_PIPELINE
            const uint32_t fl = nfl; const float4 ro = nro, rd = nrd;
            if (round + 1 < NORI_FETCH / 32u) {
                const uint32_t ns = slot + 32u;
                nfl = ns < pool.P ? pool.flags[ns] : 0u;
                if (nfl & PF_ALIVE) { nro = pool.rayO[ns]; nrd = pool.rayD[ns]; }
            }
            if (fl & PF_ALIVE) {
#else
            if (slot < pool.P && (pool.flags[slot] & PF_ALIVE)) {
                const float4 ro = pool.rayO[slot], rd = pool.rayD[slot];
#endif
                Hit h; ++nRays;
                if (traverse<false, COUNT>(sc, mk(ro.x, ro.y, ro.z), mk(rd.x, rd.y, rd.z), ro.w, rd.w, h, cnt)) {
                    pool.hit[slot] = make_float4(h.t, h.u, h.v, __uint_as_float(h.leafpos));
                    type = sc.shapes[__float_as_uint(__ldg(&sc.prims[3 * h.leafpos + 1]).w)].bsdf_type;
                } else type = missRule<VOL>(sc, pool, bt, ctr, slot, mk(ro.x, ro.y, ro.z), mk(rd.x, rd.y, rd.z), nDone);
            }
            freeList[round * 32u + lane] = (uint32_t) type;
        }
        __syncwarp();
        // ---- phase 3: bin the chunk's hits by material with ONE atomic per material and chunk (one per round made
        // the handful of queue counters the hottest instruction of the kernel: ~11 % of its stall samples)
#pragma unroll
        for (int t = 0; t < (VOL ? NORI_NQ : NORI_BSDF_COUNT); ++t) {
            uint32_t masks = 0, total = 0, before = 0;                   // lane r (< NORI_FETCH / 32) keeps round r's ballot
            for (uint32_t round = 0; round < NORI_FETCH / 32u; ++round) {
                const uint32_t m = __ballot_sync(0xffffffffu, freeList[round * 32u + lane] == (uint32_t) t);
                if (lane == round) { masks = m; before = total; }
                total += __popc(m);
            }
            if (!total) continue;
            uint32_t qb = 0;
            if (lane == 0) qb = atomicAdd(&ctr->qcount[par][t], total);
            qb = __shfl_sync(0xffffffffu, qb, 0);
            for (uint32_t round = 0; round < NORI_FETCH / 32u; ++round) {
                const uint32_t m = __shfl_sync(0xffffffffu, masks, round), b = __shfl_sync(0xffffffffu, before, round);
                if ((m >> lane) & 1u) {
                    NORI_CHECK(qb + b + __popc(m & ((1u << lane) - 1u)) < pool.P && base + round * 32u + lane < pool.P);
                    pool.queue[t][qb + b + __popc(m & ((1u << lane) - 1u))] = base + round * 32u + lane;
                }
            }
        }
        __syncwarp();
    }
    warpAdd(&ctr->rays_ext, nRays); warpAdd(&ctr->done, nDone);
    if (COUNT) { warpAdd(&ctr->nodes_ext, cnt.nodes); warpAdd(&ctr->prims_ext, cnt.prims); }
}

// ------------------------------------------------------------------------------ large-scene trace kernels
// On scenes with deep trees the rays of one warp need very different numbers of node visits (10 M
// triangles: ~100 on average, long-tailed), and a leaf costs several times an inner node.  Run as plain
// per-lane loops that leaves a warp at ~5 of 32 active lanes (ncu, profiles/).  These variants keep the
// SAME per-ray visiting order (results, and in reference order the counters, are unchanged) but schedule
// the warp as a small state machine:
//   * every lane is IDLE, at a NODE (one box test pending), in a LEAF (one primitive test pending) or
//     DONE (its answer waits in registers to be published);
//   * each warp step runs EITHER the node code for all NODE lanes OR the primitive code for all LEAF
//     lanes -- the primitive phase is entered once NORI_LEAF_MIN lanes wait in a leaf (or nothing else
//     is runnable), so both code paths execute with many lanes active;
//     (measured on the 10M-triangle scene with the 4-wide layout: 12 is best, 8 and 16 cost 1-2 %, 4 costs 15 %);
//   * the stepping loop (smRun) costs two ballots per step; publishing answers and refilling lanes from
//     the warp's slot chunk happens only once NORI_REFILL_MIN lanes are out of work, so short rays do
//     not wait for the longest ray of the warp and long rays do not pay for the bookkeeping.
#ifndef NORI_PREFETCH_CHILDREN
#define NORI_PREFETCH_CHILDREN 0
#endif
#ifndef NORI_LEAF_BURST
#define NORI_LEAF_BURST 4
#endif
#ifndef NORI_LEAF_MIN
#define NORI_LEAF_MIN 12
#endif
#ifndef NORI_REFILL_MIN
#define NORI_REFILL_MIN 12       // with 128-slot chunks (10 M triangles, 16 spp, two wavefronts): 8: 407.4, 12: 397.7, 16: 394.8, 20: 398.1, 24: 416.8 ms;
                                 // 16 is slower with one wavefront (k_extend_sm alone 1.29 vs 1.31 Grays/s): 12 kept
#endif
// resident CTAs per SM asked of the compiler (register cap = 65536 / (128 * blocks)); LAY: 0 reference nodes,
// 1 child-box pairs, 2 4-wide records
#ifndef NORI_EXT_SM_BLOCKS4
#define NORI_EXT_SM_BLOCKS4 8          // 64 registers, no spills (the 4-wide node code has no slow path)
#endif
#ifndef NORI_SHADOW_SM_BLOCKS4
#define NORI_SHADOW_SM_BLOCKS4 9
#endif
#ifndef NORI_EXT_SM_BLOCKS2
#define NORI_EXT_SM_BLOCKS2 8
#endif
#ifndef NORI_SHADOW_SM_BLOCKS2
#define NORI_SHADOW_SM_BLOCKS2 10
#endif
#define NORI_EXT_SM_BLOCKS(LAY) ((LAY) == 2 ? NORI_EXT_SM_BLOCKS4 : (LAY) == 1 ? NORI_EXT_SM_BLOCKS2 : 8)
#define NORI_SHADOW_SM_BLOCKS(LAY) ((LAY) == 2 ? NORI_SHADOW_SM_BLOCKS4 : (LAY) == 1 ? NORI_SHADOW_SM_BLOCKS2 : 10)
enum { ST_IDLE = 0, ST_NODE = 1, ST_LEAF = 2, ST_DONE = 3 };

struct LaneTrav {
    RayTrav r;
    uint32_t st, leafI, leafEnd, slot;
    uint32_t cur;          // child-box layout only: reference of the current inner node (index << 2 | split axis)
    uint32_t neg;          // bit a set: the ray runs towards -a and the near-child-first order is on (descend(), traverse.cuh)
};

// Traversal stack of one lane: the first NORI_SM_STACK entries live in shared memory (column `tid` of a
// [entry][thread] array: conflict-free), deeper ones spill to a local array.  With the whole stack in local
// memory (ncu, 10M-triangle scene) every pop was an L1 lookup with a 47 % miss rate -- an L2 round trip in
// front of the node fetch that depends on it -- and the stack lines evicted node data from L1.
#ifndef NORI_SM_STACK
#define NORI_SM_STACK 32
#endif
struct LaneStack {
    uint32_t *sh;                       // &s_stack[0][tid]; entry e at sh[e * 128]
    uint32_t ovf[64 - NORI_SM_STACK];
    __device__ __forceinline__ void push(uint32_t sp, uint32_t v) { NORI_CHECK(sp < 64); if (sp < NORI_SM_STACK) sh[sp * 128u] = v; else ovf[sp - NORI_SM_STACK] = v; }
    __device__ __forceinline__ uint32_t pop(uint32_t sp) const { return sp < NORI_SM_STACK ? sh[sp * 128u] : ovf[sp - NORI_SM_STACK]; }
};

// Child-box layout (DScene::nodes2; built at upload from the reference's nodes, see nori_gpu.cu): one
// 64-byte record per INNER node holding the boxes and references of both children, so that
//   * one step answers two box tests from one record (4 x LDG.128) -- half the dependent fetches;
//   * leaf nodes are never fetched: a leaf child's reference carries its primitive range;
//   * the far child is pushed together with its entry distance and culled at pop time, without a fetch,
//     when a closer hit has been found meanwhile.
// The box arithmetic is the reference's slab test on the reference's boxes, so the set of primitives a
// ray can reach is unchanged; used with the near-child-first order only (the node counters then count
// boxes tested and no longer equal the reference's).
//   child reference: bit 31 = leaf; leaf: size in bits 30..25, first primitive in bits 24..0;
//                    inner: record index in bits 30..2, split axis in bits 1..0
#ifndef NORI_SM_STACK2
#define NORI_SM_STACK2 16
#endif
// deepest stack (NORI_STACK2_MAX, kernels.cuh): child-box pairs push one entry per level (<= 64 levels, checked
// at upload); the 4-wide layout pushes up to three per record level (checked when the records are built)
struct LaneStack2 {
    uint2 *sh;                          // &s_stack2[0][tid]; entry e at sh[e * 128]
    uint2 ovf[NORI_STACK2_MAX - NORI_SM_STACK2];
    __device__ __forceinline__ void push(uint32_t sp, uint32_t ref, float nearT) {
        NORI_CHECK(sp < NORI_STACK2_MAX);
        const uint2 v = make_uint2(ref, __float_as_uint(nearT));
        if (sp < NORI_SM_STACK2) sh[sp * 128u] = v; else ovf[sp - NORI_SM_STACK2] = v;
    }
    __device__ __forceinline__ uint2 pop(uint32_t sp) const { return sp < NORI_SM_STACK2 ? sh[sp * 128u] : ovf[sp - NORI_SM_STACK2]; }
};

// bbox.h:336-363 on one child box, plus the interval test of bvh.cpp:423
template <bool PLAIN = false>
__device__ __forceinline__ bool boxTest(const RayTrav &r, float3 mn, float3 mx, float &nearT) {
    if (PLAIN) return boxPlain(r.o, r.rcp, r.mint, r.cull, mn.x, mn.y, mn.z, mx.x, mx.y, mx.z, nearT);
    nearT = __int_as_float(0xff800000); float farT = __int_as_float(0x7f800000);
    return slab(r.o.x, r.d.x, r.rcp.x, mn.x, mx.x, nearT, farT) && slab(r.o.y, r.d.y, r.rcp.y, mn.y, mx.y, nearT, farT)
        && slab(r.o.z, r.d.z, r.rcp.z, mn.z, mx.z, nearT, farT) && (r.mint <= farT && nearT <= r.cull);
}

// The order guard of traverse.cuh for one lane.  A second hit candidate within the margin of the current best ends the
// near-first query and the lane starts over in REFERENCE MODE (bit NORI_REFMODE of LaneTrav::neg): left child first --
// in the 4-wide records: hit slots in slot order, which host_bvh.cpp keeps depth-first -- with the exact distance
// cull and the reference's `t <= maxt` acceptance, i.e. bvh.cpp:404-462 step for step (the merged nodes' own box
// tests are implied by their children's: nested boxes, nested intervals).  Still inside the state machine: the other
// lanes of the warp do not wait for it.
#define NORI_REFMODE 0x80000000u
template <int LAY, bool COUNT>
__device__ __forceinline__ void smCandidate(const DScene &sc, const float4 *rayD, LaneTrav &L, const float4 &r1, const float4 &r2, float t, float u, float v,
                                            uint32_t i, TraceCounters &cnt) {
    RayTrav &r = L.r;
    float m = 0.f;
#if NORI_ORDER_GUARD
    if (sc.ordered && !(L.neg & NORI_REFMODE)) {
        m = guardMargin(r1, r2, r.d);
        if (r.found && (t >= 2.0f * r.hit.t - r.cull || t >= __fmul_rn(r.hit.t, 1.0f - m))) {      // a second candidate: start over in reference mode
            if (COUNT) ++cnt.redo;
            L.neg = NORI_REFMODE; L.st = ST_NODE;
            L.cur = LAY == 2 ? sc.root_ref4 : sc.root_ref; r.node = 0; r.sp = 0;  // the root's own box was passed at the start
            r.found = false; r.cull = rayD[L.slot].w;                              // the query's own far end (not kept in a register)
            r.hit.t = __int_as_float(0x7f800000); r.hit.u = 0.f; r.hit.v = 0.f; r.hit.leafpos = NORI_NO_HIT;
            return;
        }
    }
#else                                                        // experiment: the leaf-position tie rule alone
    if (!(!r.found || t < r.hit.t || i > r.hit.leafpos)) return;
#endif
    r.found = true; r.cull = __fmaf_rn(t, m, t);            // reference mode / reference order: m = 0, the exact cull
    r.hit.t = t; r.hit.u = u; r.hit.v = v; r.hit.leafpos = i;
}

// continue with child `ref`: an inner child becomes the current node, a leaf child the current primitive range
__device__ __forceinline__ bool smEnter(LaneTrav &L, uint32_t ref) {
    if (ref & 0x80000000u) {
        const uint32_t size = (ref >> 25) & 63u, start = ref & 0x1ffffffu;
        if (!size) return false;                                 // empty leaf (bvh.cpp:437): nothing to do
        L.st = ST_LEAF; L.leafI = start; L.leafEnd = start + size;
    } else { L.st = ST_NODE; L.cur = ref; }
    return true;
}
// pop until an entry survives the distance cull; ST_DONE when the stack runs dry
__device__ __forceinline__ void smPop2(LaneTrav &L, const LaneStack2 &stack) {
    RayTrav &r = L.r;
    while (r.sp) {
        const uint2 e = stack.pop(--r.sp);
        if (__uint_as_float(e.y) <= r.cull && smEnter(L, e.x)) return;
    }
    L.st = ST_DONE;
}

template <bool COUNT>
__device__ __forceinline__ void smNode2(const DScene &sc, LaneTrav &L, LaneStack2 &stack, TraceCounters &cnt) {
    RayTrav &r = L.r;
    const uint4 *rec = &sc.nodes2[4 * (size_t) (L.cur >> 2)];
    uint4 a, b, c, d; ldgPair(rec, a, b); ldgPair(rec + 2, c, d);
    if (COUNT) cnt.nodes += 2;
    float nearL, nearR;
    bool hitL, hitR;
#define NORI_BOX(P, a, b, n) boxTest<P>(r, make_float3(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z)), \
                                        make_float3(__uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z)), n)
    if (r.plain) { hitL = NORI_BOX(true, a, b, nearL); hitR = NORI_BOX(true, c, d, nearR); }
    else { hitL = NORI_BOX(false, a, b, nearL); hitR = NORI_BOX(false, c, d, nearR); }
    const bool swap = (L.neg >> (L.cur & 3u)) & 1u;                  // the right child is the near one
    const uint32_t refN = swap ? b.w : a.w, refF = swap ? a.w : b.w;
    const bool hitN = swap ? hitR : hitL, hitF = swap ? hitL : hitR;
    const float nearF = swap ? nearL : nearR;
    if (hitN) {
        if (hitF) { stack.push(r.sp++, refF, nearF); if (COUNT) cnt.maxsp = max(cnt.maxsp, r.sp); }
        if (smEnter(L, refN)) return;
    } else if (hitF && smEnter(L, refF)) return;
    smPop2(L, stack);
}

// 4-wide layout (DScene::nodes4; built at upload from the reference's nodes, see nori_gpu.cu): one 128-byte
// record (one cache line) holds the boxes and references of up to four descendants of a binary inner node, so a
// visit answers four box tests and the chain of dependent fetches is less than half as long as with pairs.  The
// boxes are the reference's own and nested boxes give nested slab intervals (round-to-nearest is monotonic), so
// skipping the merged nodes' own box tests does not change the set of primitives a ray can reach -- for rays
// that cannot meet a NaN in the slab test (rayPlain(), traverse.cuh); the others never enter this layout (smStart).
//   record: slot k = quads 2k (min.xyz, reference of the descendant) and 2k+1 (max.xyz, -); an unused slot holds
//           the empty-leaf reference 0x80000000 and is skipped
//   child reference: bit 31 = leaf (as above); inner: record index
// Visiting order (closest-hit): hit slots by entry distance (a 5-exchange sorting network on (distance,
// reference)); the nearest is entered, the others wait on the stack, farthest lowest, and are culled at pop time
// as above.  Any-hit queries take the slots as they come.
template <bool SHADOW, bool COUNT>
__device__ __forceinline__ void smNode4(const DScene &sc, LaneTrav &L, LaneStack2 &stack, TraceCounters &cnt) {
    RayTrav &r = L.r;
    const uint4 *rec = &sc.nodes4[8 * (size_t) L.cur];
    uint4 a0, b0, a1, b1, a2, b2, a3, b3;
    ldgPair(rec, a0, b0); ldgPair(rec + 2, a1, b1); ldgPair(rec + 4, a2, b2); ldgPair(rec + 6, a3, b3);
#if NORI_PREFETCH_CHILDREN
    // experiment: request every inner child's record (one line each) into L2 before the box tests decide which are entered
    if (!(a0.w & 0x80000000u)) asm volatile("prefetch.global.L2 [%0];" :: "l"(&sc.nodes4[8 * (size_t) a0.w]));
    if (!(a1.w & 0x80000000u)) asm volatile("prefetch.global.L2 [%0];" :: "l"(&sc.nodes4[8 * (size_t) a1.w]));
    if (!(a2.w & 0x80000000u)) asm volatile("prefetch.global.L2 [%0];" :: "l"(&sc.nodes4[8 * (size_t) a2.w]));
    if (!(a3.w & 0x80000000u)) asm volatile("prefetch.global.L2 [%0];" :: "l"(&sc.nodes4[8 * (size_t) a3.w]));
#endif
    if (COUNT) cnt.nodes += (a0.w != 0x80000000u) + (a1.w != 0x80000000u) + (a2.w != 0x80000000u) + (a3.w != 0x80000000u);   // boxes tested
    float n0, n1, n2, n3;
    bool h0, h1, h2, h3;
    h0 = NORI_BOX(true, a0, b0, n0); h1 = NORI_BOX(true, a1, b1, n1); h2 = NORI_BOX(true, a2, b2, n2); h3 = NORI_BOX(true, a3, b3, n3);   // only rayPlain() rays get here (smStart)
    if (SHADOW) {                                                // any-hit: the order does not matter, nothing to cull by
        h0 = h0 && a0.w != 0x80000000u; h1 = h1 && a1.w != 0x80000000u; h2 = h2 && a2.w != 0x80000000u; h3 = h3 && a3.w != 0x80000000u;
        uint32_t ref = 0; bool any = false;
        if (h3) { ref = a3.w; any = true; }
        if (h2) { if (any) stack.push(r.sp++, ref, 0.f); ref = a2.w; any = true; }
        if (h1) { if (any) stack.push(r.sp++, ref, 0.f); ref = a1.w; any = true; }
        if (h0) { if (any) stack.push(r.sp++, ref, 0.f); ref = a0.w; any = true; }
        if (COUNT) cnt.maxsp = max(cnt.maxsp, r.sp);
        if (any && smEnter(L, ref)) return;
        smPop2(L, stack);
        return;
    }
#if NORI_ORDER_GUARD
    if (L.neg & NORI_REFMODE) {
        // reference mode (rare): the hit slots in the reference's depth-first order -- their rank is the spare word of
        // each slot -- pushed last to first with their entry distances, so that the pops walk them first to last
#pragma unroll
        for (uint32_t rank = 4; rank-- > 0;) {
            if (h0 && a0.w != 0x80000000u && b0.w == rank) stack.push(r.sp++, a0.w, n0);
            if (h1 && a1.w != 0x80000000u && b1.w == rank) stack.push(r.sp++, a1.w, n1);
            if (h2 && a2.w != 0x80000000u && b2.w == rank) stack.push(r.sp++, a2.w, n2);
            if (h3 && a3.w != 0x80000000u && b3.w == rank) stack.push(r.sp++, a3.w, n3);
        }
        if (COUNT) cnt.maxsp = max(cnt.maxsp, r.sp);
        smPop2(L, stack);
        return;
    }
#endif
    // sort key: the entry distance (capped below the "missed" key +inf; a pushed distance is only used to cull)
    const float inf = __int_as_float(0x7f800000), big = __int_as_float(0x7f7fffff);
    float k0 = (h0 && a0.w != 0x80000000u) ? fminf(n0, big) : inf, k1 = (h1 && a1.w != 0x80000000u) ? fminf(n1, big) : inf;
    float k2 = (h2 && a2.w != 0x80000000u) ? fminf(n2, big) : inf, k3 = (h3 && a3.w != 0x80000000u) ? fminf(n3, big) : inf;
    uint32_t r0 = a0.w, r1 = a1.w, r2 = a2.w, r3 = a3.w;
#define NORI_CSWAP(ka, ra, kb, rb) { const bool s_ = kb < ka; const float kt = s_ ? kb : ka; kb = s_ ? ka : kb; ka = kt; \
                                      const uint32_t rt = s_ ? rb : ra; rb = s_ ? ra : rb; ra = rt; }
    NORI_CSWAP(k0, r0, k1, r1) NORI_CSWAP(k2, r2, k3, r3) NORI_CSWAP(k0, r0, k2, r2) NORI_CSWAP(k1, r1, k3, r3) NORI_CSWAP(k1, r1, k2, r2)
#undef NORI_CSWAP
    if (k3 < inf) stack.push(r.sp++, r3, k3);
    if (k2 < inf) stack.push(r.sp++, r2, k2);
    if (k1 < inf) stack.push(r.sp++, r1, k1);
    if (COUNT) cnt.maxsp = max(cnt.maxsp, r.sp);
    if (k0 < inf && smEnter(L, r0)) return;
    smPop2(L, stack);
}

template <int LAY, bool SHADOW, bool COUNT>
__device__ __forceinline__ void smPrim2(const DScene &sc, const float4 *rayD, LaneTrav &L, LaneStack2 &stack, TraceCounters &cnt) {
    RayTrav &r = L.r;
    const uint32_t i = L.leafI;
    NORI_CHECK(i < sc.n_prims);
    const float4 r0 = __ldg(&sc.prims[3 * i]);
    const float4 r1 = __ldg(&sc.prims[3 * i + 1]);
    const float4 r2 = __ldg(&sc.prims[3 * i + 2]);
    if (COUNT) ++cnt.prims;
    float u = 0.f, v = 0.f, t;
    bool h;
    if (__float_as_uint(r2.w) == 0u)
        h = triTest(mk(r0.x, r0.y, r0.z), mk(r1.x, r1.y, r1.z), mk(r2.x, r2.y, r2.z), r.o, r.d, r.mint, r.cull, u, v, t);
    else
        h = roundTest(r0, r1, r2, r.o, r.d, r.mint, r.cull, t);
    if (h) {
        if (SHADOW) { r.found = true; r.hit.t = 0.f; L.st = ST_DONE; return; }
        smCandidate<LAY, COUNT>(sc, rayD, L, r1, r2, t, u, v, i, cnt);
        if (L.st != ST_LEAF) return;                        // the order guard restarted the lane
    }
    if (++L.leafI < L.leafEnd) return;
    smPop2(L, stack);
}

template <bool SHADOW, bool COUNT, bool WIDE>
__device__ __forceinline__ void smRun2(const DScene &sc, const float4 *rayD, LaneTrav &L, LaneStack2 &stack, TraceCounters &cnt, bool canRefill) {
    while (true) {
        const uint32_t mNode = __ballot_sync(0xffffffffu, L.st == ST_NODE);
        const uint32_t mLeaf = __ballot_sync(0xffffffffu, L.st == ST_LEAF);
        const uint32_t mWork = mNode | mLeaf;
        if (!mWork || (canRefill && __popc(mWork) <= 32 - NORI_REFILL_MIN)) return;
        if (mLeaf && (__popc(mLeaf) >= NORI_LEAF_MIN || !mNode)) {
            if (L.st == ST_LEAF) {
                const uint32_t leafEnd = L.leafEnd;                  // up to NORI_LEAF_BURST primitives of THIS leaf per step
                smPrim2<WIDE ? 2 : 1, SHADOW, COUNT>(sc, rayD, L, stack, cnt);
#pragma unroll 1
                for (int k = 1; k < NORI_LEAF_BURST; ++k)
                    if (L.st == ST_LEAF && L.leafEnd == leafEnd && L.leafI > 0) smPrim2<WIDE ? 2 : 1, SHADOW, COUNT>(sc, rayD, L, stack, cnt);
            }
        } else {
            if (L.st == ST_NODE) { if (WIDE) smNode4<SHADOW, COUNT>(sc, L, stack, cnt); else smNode2<COUNT>(sc, L, stack, cnt); }
        }
    }
}

template <int LAY, bool SHADOW, bool COUNT>
__device__ __forceinline__ void smStart(const DScene &sc, LaneTrav &L, V3 o, V3 d, float mint, float maxt, TraceCounters &cnt) {
    L.neg = sc.ordered ? ((d.x < 0.f ? 1u : 0u) | (d.y < 0.f ? 2u : 0u) | (d.z < 0.f ? 4u : 0u)) : 0u;
    if (!travInit(sc, L.r, o, d, mint, maxt)) { L.r.found = false; L.st = ST_DONE; return; }   // decided before the first node: a miss
    L.st = ST_NODE;
    if (LAY == 2 && !L.r.plain) {
        // The 4-wide records skip the box tests of the merged nodes, which is only equivalent while nested boxes give
        // nested slab intervals.  A ray outside rayPlain() can meet a NaN there (origin on a bounding plane, 1/d
        // infinite: bbox.h:347-348 then rejects THAT box), so it walks the reference's own nodes in the reference's
        // order, here and now.
        L.r.found = traverse<SHADOW, COUNT, false, false>(sc, o, d, mint, maxt, L.r.hit, cnt);
        L.st = ST_DONE;
        return;
    }
    if (LAY) {                                                   // the root's own box (bvh.cpp:421-423 on node 0)
        L.cur = LAY == 2 ? sc.root_ref4 : sc.root_ref;                // (the root record's index, 0 / the root's index and axis)
        float nearT;
        if (!boxTest(L.r, make_float3(sc.root_min[0], sc.root_min[1], sc.root_min[2]), make_float3(sc.root_max[0], sc.root_max[1], sc.root_max[2]), nearT)) L.st = ST_DONE;
    }
}

// node phase for one lane
template <bool COUNT>
__device__ __forceinline__ void smNode(const DScene &sc, LaneTrav &L, LaneStack &stack, TraceCounters &cnt) {
    RayTrav &r = L.r;
    uint4 n0, n1; ldgPair(&sc.nodes[2 * r.node], n0, n1);
    if (COUNT) ++cnt.nodes;
    if (nodeBox(r.plain, r.o, r.d, r.rcp, r.mint, r.cull, n0, n1)) {
        if (!(n0.x & 1u)) {                                      // inner: descend() of traverse.cuh with the sign mask
            const bool swap = (L.neg >> (n0.x >> 1)) & 1u;
            const uint32_t farC = swap ? r.node + 1 : n0.y;
            stack.push(r.sp++, farC);
            if (COUNT) cnt.maxsp = max(cnt.maxsp, r.sp);
            r.node = swap ? n0.y : r.node + 1;
            return;
        }
        const uint32_t size = n0.x >> 1;
        if (size) {
            L.st = ST_LEAF; L.leafI = n0.y; L.leafEnd = n0.y + size;
            return;
        }
    }
    if (r.sp == 0) { L.st = ST_DONE; return; }
    r.node = stack.pop(--r.sp);
}

// primitive phase for one lane: one primitive test
template <bool SHADOW, bool COUNT>
__device__ __forceinline__ void smPrim(const DScene &sc, const float4 *rayD, LaneTrav &L, LaneStack &stack, TraceCounters &cnt) {
    RayTrav &r = L.r;
    const uint32_t i = L.leafI;
    NORI_CHECK(i < sc.n_prims);
    const float4 r0 = __ldg(&sc.prims[3 * i]);
    const float4 r1 = __ldg(&sc.prims[3 * i + 1]);
    const float4 r2 = __ldg(&sc.prims[3 * i + 2]);
    if (COUNT) ++cnt.prims;
    float u = 0.f, v = 0.f, t;
    bool h;
    if (__float_as_uint(r2.w) == 0u)
        h = triTest(mk(r0.x, r0.y, r0.z), mk(r1.x, r1.y, r1.z), mk(r2.x, r2.y, r2.z), r.o, r.d, r.mint, r.cull, u, v, t);
    else
        h = roundTest(r0, r1, r2, r.o, r.d, r.mint, r.cull, t);
    if (h) {
        if (SHADOW) { r.found = true; r.hit.t = 0.f; L.st = ST_DONE; return; }
        smCandidate<0, COUNT>(sc, rayD, L, r1, r2, t, u, v, i, cnt);
        if (L.st != ST_LEAF) return;                        // the order guard restarted the lane
    }
    if (++L.leafI < L.leafEnd) return;
    if (r.sp == 0) { L.st = ST_DONE; return; }
    L.st = ST_NODE;
    r.node = stack.pop(--r.sp);
}

// Step the warp until NORI_REFILL_MIN lanes are out of work (and `canRefill` says new rays exist) or no
// lane has work left.
template <bool SHADOW, bool COUNT>
__device__ __forceinline__ void smRun(const DScene &sc, const float4 *rayD, LaneTrav &L, LaneStack &stack, TraceCounters &cnt, bool canRefill) {
    while (true) {
        const uint32_t mNode = __ballot_sync(0xffffffffu, L.st == ST_NODE);
        const uint32_t mLeaf = __ballot_sync(0xffffffffu, L.st == ST_LEAF);
        const uint32_t mWork = mNode | mLeaf;
        if (!mWork || (canRefill && __popc(mWork) <= 32 - NORI_REFILL_MIN)) return;
        if (mLeaf && (__popc(mLeaf) >= NORI_LEAF_MIN || !mNode)) {
            if (L.st == ST_LEAF) smPrim<SHADOW, COUNT>(sc, rayD, L, stack, cnt);
        } else {
            if (L.st == ST_NODE) smNode<COUNT>(sc, L, stack, cnt);
        }
    }
}

// Claim the warp's next NORI_FETCH pool slots; returns false when the pool is exhausted.
__device__ __forceinline__ bool smNextChunk(uint32_t *work, uint32_t P, uint32_t lane, uint32_t &chunkBase) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(work, NORI_FETCH);
    base = __shfl_sync(0xffffffffu, base, 0);
    chunkBase = base;
    return base < P;
}

template <bool COUNT, bool VOL, int LAY>
__global__ void __launch_bounds__(128, NORI_EXT_SM_BLOCKS(LAY)) k_extend_sm(DScene sc, Pool pool, Batch bt, Counters *ctr, uint32_t it) {
    __shared__ uint32_t s_free[4][NORI_FETCH];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, par = it & 1u;
    const uint32_t ltMask = (1u << lane) - 1u;
    uint32_t *freeList = s_free[warp];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int i = 0; i < NORI_NQ; ++i) ctr->qcount[par ^ 1u][i] = 0;
        for (int i = 0; i < NORI_NEQ; ++i) ctr->eqcount[par ^ 1u][i] = 0;
        ctr->work_extend[par ^ 1u] = 0; ctr->work_shadow[par ^ 1u] = 0;
    }
    const unsigned long long total = ctr->total_samples;
    uint32_t nRays = 0, nDone = 0; TraceCounters cnt;
    __shared__ uint2 s_stack_mem[(LAY ? NORI_SM_STACK2 * 2 : NORI_SM_STACK) * 128 / 2];
    LaneStack stack; stack.sh = (uint32_t *) s_stack_mem + threadIdx.x;
    LaneStack2 stack2; stack2.sh = s_stack_mem + threadIdx.x;
    LaneTrav L; L.st = ST_IDLE; L.slot = 0; L.leafI = L.leafEnd = 0; L.neg = 0; L.cur = 0;
    uint32_t chunkBase = 0, chunkNext = NORI_FETCH;
    bool moreChunks = true;
    while (true) {
        // ---- publish finished rays, bin hits by material (one atomic per warp and material)
        if (__any_sync(0xffffffffu, L.st == ST_DONE)) {
            int type = -1;
            if (L.st == ST_DONE) {
                L.st = ST_IDLE;
                if (L.r.found) {
                    pool.hit[L.slot] = make_float4(L.r.hit.t, L.r.hit.u, L.r.hit.v, __uint_as_float(L.r.hit.leafpos));
                    type = sc.shapes[__float_as_uint(__ldg(&sc.prims[3 * L.r.hit.leafpos + 1]).w)].bsdf_type;
                } else type = missRule<VOL, false>(sc, pool, bt, ctr, L.slot, L.r.o, L.r.d, nDone);
            }
#pragma unroll
            for (int t = 0; t < (VOL ? NORI_NQ : NORI_BSDF_COUNT); ++t) {
                const uint32_t m = __ballot_sync(0xffffffffu, type == t);
                if (!m) continue;
                uint32_t qb = 0; const int leader = __ffs(m) - 1;
                if ((int) lane == leader) qb = atomicAdd(&ctr->qcount[par][t], (uint32_t) __popc(m));
                qb = __shfl_sync(0xffffffffu, qb, leader);
                if (type == t) { NORI_CHECK(qb + __popc(m & ltMask) < pool.P && L.slot < pool.P); pool.queue[t][qb + __popc(m & ltMask)] = L.slot; }
            }
        }
        // ---- refill idle lanes from the warp's chunk of pool slots
        uint32_t idle = __ballot_sync(0xffffffffu, L.st == ST_IDLE);
        while (idle) {
            if (chunkNext >= NORI_FETCH) {
                if (!moreChunks) break;
                if (!smNextChunk(&ctr->work_extend[par], pool.P, lane, chunkBase)) { moreChunks = false; break; }
                chunkNext = 0;
                uint32_t nFree = 0;                              // compacted regeneration (render.cpp:98-124)
                for (uint32_t round = 0; round < NORI_FETCH / 32u; ++round) {
                    const uint32_t s = chunkBase + round * 32u + lane;
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
                        generatePath<false>(sc, bt, (uint32_t) id, 0, ray, rs);
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
                    smStart<LAY, false, COUNT>(sc, L, mk(ro.x, ro.y, ro.z), mk(rd.x, rd.y, rd.z), ro.w, rd.w, cnt);
                }
            }
            idle = __ballot_sync(0xffffffffu, L.st == ST_IDLE);
        }
        const bool canRefill = moreChunks || chunkNext < NORI_FETCH;
        if (!canRefill && !__any_sync(0xffffffffu, L.st != ST_IDLE)) break;
        if (LAY) smRun2<false, COUNT, LAY == 2>(sc, pool.rayD, L, stack2, cnt, canRefill); else smRun<false, COUNT>(sc, pool.rayD, L, stack, cnt, canRefill);
    }
    warpAdd(&ctr->rays_ext, nRays); warpAdd(&ctr->done, nDone);
    if (COUNT) { warpAdd(&ctr->nodes_ext, cnt.nodes); warpAdd(&ctr->prims_ext, cnt.prims); warpMax(&ctr->max_stack, cnt.maxsp); warpAdd(&ctr->guard_redo, cnt.redo); }
}

// ------------------------------------------------------------------------------ deferred shadow rays
// Any-hit traversal (bvh.cpp:441-442) of the NEE rays k_shade<.., DEFER> left in the pool, same warp
// state machine as k_extend_sm.  Adds the pending contribution when the ray is unoccluded
// (path_mis.cpp:48-61) and finalises the paths the roulette ended.
template <bool COUNT, int LAY>
__global__ void __launch_bounds__(128, NORI_SHADOW_SM_BLOCKS(LAY)) k_shadow_sm(DScene sc, Pool pool, Batch bt, Counters *ctr, uint32_t it) {
    const uint32_t lane = threadIdx.x & 31, par = it & 1u;
    const uint32_t ltMask = (1u << lane) - 1u;
    uint32_t nRays = 0, nDone = 0; TraceCounters cnt;
    __shared__ uint2 s_stack_mem[(LAY ? NORI_SM_STACK2 * 2 : NORI_SM_STACK) * 128 / 2];
    LaneStack stack; stack.sh = (uint32_t *) s_stack_mem + threadIdx.x;
    LaneStack2 stack2; stack2.sh = s_stack_mem + threadIdx.x;
    LaneTrav L; L.st = ST_IDLE; L.slot = 0; L.leafI = L.leafEnd = 0; L.neg = 0; L.cur = 0;
    uint32_t flags = 0;
    uint32_t chunkBase = 0, chunkNext = NORI_FETCH;
    bool moreChunks = true;
    while (true) {
        if (L.st == ST_DONE) {                                   // publish
            L.st = ST_IDLE;
            float4 ra = pool.rad[L.slot];
            if (!L.r.found) { const float4 c = pool.shC[L.slot]; ra.x = __fadd_rn(ra.x, c.x); ra.y = __fadd_rn(ra.y, c.y); ra.z = __fadd_rn(ra.z, c.z); }
            if (flags & PF_TERMINATE) endOfPath<false>(sc, pool, bt, ctr, L.slot, pool.sid[L.slot], mk(ra.x, ra.y, ra.z), pool.rng[L.slot], flags, nDone);
            else {
                if (!L.r.found) pool.rad[L.slot] = ra;
                pool.flags[L.slot] = flags & ~PF_SHADOW;
            }
        }
        uint32_t idle = __ballot_sync(0xffffffffu, L.st == ST_IDLE);
        while (idle) {
            if (chunkNext >= NORI_FETCH) {
                if (!moreChunks) break;
                if (!smNextChunk(&ctr->work_shadow[par], pool.P, lane, chunkBase)) { moreChunks = false; break; }
                chunkNext = 0;
            }
            const uint32_t idx = chunkNext + __popc(idle & ltMask);
            const bool take = L.st == ST_IDLE && idx < NORI_FETCH;
            chunkNext = min(chunkNext + (uint32_t) __popc(idle), NORI_FETCH);
            if (take) {
                const uint32_t s = chunkBase + idx;
                if (s < pool.P) {
                    const uint32_t f = pool.flags[s];
                    if (f & PF_SHADOW) {
                        const float4 so = pool.rayO[s], sd = pool.shD[s];
                        L.slot = s; flags = f; ++nRays;
                        smStart<LAY, true, COUNT>(sc, L, mk(so.x, so.y, so.z), mk(sd.x, sd.y, sd.z), NORI_EPS, sd.w, cnt);
                    }
                }
            }
            idle = __ballot_sync(0xffffffffu, L.st == ST_IDLE);
        }
        const bool canRefill = moreChunks || chunkNext < NORI_FETCH;
        if (!canRefill && !__any_sync(0xffffffffu, L.st != ST_IDLE)) break;
        if (LAY) smRun2<true, COUNT, LAY == 2>(sc, nullptr, L, stack2, cnt, canRefill); else smRun<true, COUNT>(sc, nullptr, L, stack, cnt, canRefill);
    }
    warpAdd(&ctr->rays_sh, nRays); warpAdd(&ctr->done, nDone);
    if (COUNT) { warpAdd(&ctr->nodes_sh, cnt.nodes); warpAdd(&ctr->prims_sh, cnt.prims); warpMax(&ctr->max_stack, cnt.maxsp); }
}

template <bool COUNT, int LAY> static void launchShadowSm(int grid, cudaStream_t st, const DScene &sc, const Pool &pool, const Batch &bt, Counters *ctr, uint32_t it) {
    k_shadow_sm<COUNT, LAY><<<grid, 128, 0, st>>>(sc, pool, bt, ctr, it);
}
template <bool COUNT> static void launchShadowSmLay(int lay, int grid, cudaStream_t st, const DScene &sc, const Pool &pool, const Batch &bt, Counters *ctr, uint32_t it) {
    if (lay == 2) launchShadowSm<COUNT, 2>(grid, st, sc, pool, bt, ctr, it);
    else if (lay == 1) launchShadowSm<COUNT, 1>(grid, st, sc, pool, bt, ctr, it);
    else launchShadowSm<COUNT, 0>(grid, st, sc, pool, bt, ctr, it);
}
void noriLaunchShadowSm(bool count, int grid, cudaStream_t st, const DScene &sc, const Pool &pool, const Batch &bt, Counters *ctr, uint32_t it) {
    if (count) launchShadowSmLay<true>(noriSmLayout(sc), grid, st, sc, pool, bt, ctr, it);
    else launchShadowSmLay<false>(noriSmLayout(sc), grid, st, sc, pool, bt, ctr, it);
}
template <bool COUNT> static int shadowSmOcc(int lay) {
    int occ = 8;
    if (lay == 2) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_shadow_sm<COUNT, 2>, 128, 0);
    else if (lay == 1) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_shadow_sm<COUNT, 1>, 128, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_shadow_sm<COUNT, 0>, 128, 0);
    return occ;
}
int noriShadowSmOccupancy(bool count, int lay) { return count ? shadowSmOcc<true>(lay) : shadowSmOcc<false>(lay); }

template <bool COUNT, bool VOL> static ExtendKernel pickExtendSm(int lay) {
    return lay == 2 ? k_extend_sm<COUNT, VOL, 2> : lay == 1 ? k_extend_sm<COUNT, VOL, 1> : k_extend_sm<COUNT, VOL, 0>;
}
ExtendKernel noriPickExtend(bool sm, bool count, bool vol, int lay) {
    if (sm) return count ? (vol ? pickExtendSm<true, true>(lay) : pickExtendSm<true, false>(lay)) : (vol ? pickExtendSm<false, true>(lay) : pickExtendSm<false, false>(lay));
    return count ? (vol ? k_extend<true, true> : k_extend<true, false>) : (vol ? k_extend<false, true> : k_extend<false, false>);
}
