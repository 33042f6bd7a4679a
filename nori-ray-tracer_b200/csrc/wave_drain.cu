// wave_drain.cu -- k_drain: finishes the last few paths of a wavefront batch, one thread per path.
#include "kernels.cuh"

// ------------------------------------------------------------------------------ drain
// Once the batch has no camera path left to start, the pool empties geometrically -- except for the few
// paths that keep a throughput of ~1 (chains of mirror / dielectric bounces survive the roulette with
// p = 0.99, path_mis.cpp:64), which kept the Cornell box iterating for ~770 more extend + shade launch
// pairs over a 4 Mi-slot pool that was almost empty (12 % of the step).  k_drain finishes them instead: one
// grid-stride scan of the pool, every live slot is run to its end by its thread with the same per-vertex
// code (bit-identical results), including the per-channel restarts of chromatic aberration.
template <bool MIS, bool COUNT>
__global__ void __launch_bounds__(128) k_drain(DScene sc, Pool pool, Batch bt, Counters *ctr) {
    uint32_t nRays = 0, nShadow = 0, nDone = 0; TraceCounters cnt;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < pool.P; slot += stride) {
        while (pool.flags[slot] & PF_ALIVE) {                       // one pass per colour channel (one pass normally)
            const float4 ro = pool.rayO[slot], rd = pool.rayD[slot], th = pool.thr[slot], ra = pool.rad[slot];
            const uint32_t sid = pool.sid[slot];
            PathState st;
            st.o = mk(ro.x, ro.y, ro.z); st.d = mk(rd.x, rd.y, rd.z);
            st.thr = mk(th.x, th.y, th.z); st.pdf_mat = th.w; st.rad = mk(ra.x, ra.y, ra.z);
            st.flags = pool.flags[slot];
            const uint32_t chBits = st.flags & PF_CH_MASK;
            st.rng.state = pool.rng[slot]; st.rng.inc = ((uint64_t) (sid % bt.wh) << 1u) | 1u;
            float mint = ro.w, maxt = rd.w;
            while (true) {
                Hit h; ++nRays;
                if (!traverse<false, COUNT>(sc, st.o, st.d, mint, maxt, h, cnt)) break;
                VertexOut out;
                pathVertex<-1, MIS>(sc, h, st, out);
                if (MIS && (COUNT || !nullContribution(out.contrib.x, out.contrib.y, out.contrib.z))) {
                    Hit sh; ++nShadow;
                    if (!traverse<true, COUNT>(sc, out.shadow.o, out.shadow.d, out.shadow.mint, out.shadow.maxt, sh, cnt))
                        st.rad = st.rad + out.contrib;
                }
                if (!(st.flags & PF_ALIVE)) break;
                mint = out.next.mint; maxt = out.next.maxt;         // st.o / st.d already hold the next ray
            }
            endOfPath(sc, pool, bt, ctr, slot, sid, st.rad, st.rng.state, chBits, nDone);
        }
    }
    warpAdd(&ctr->rays_ext, nRays); warpAdd(&ctr->done, nDone);
    if (MIS) warpAdd(&ctr->rays_sh, nShadow);
    if (COUNT) { warpAdd(&ctr->nodes_ext, cnt.nodes); warpAdd(&ctr->prims_ext, cnt.prims); }
}

// The same tail with ONE WARP PER PATH.  What the GPU waits for at the end of a batch is the longest surviving path (a
// chain of specular bounces: roulette p = 0.99 keeps one of 10^7 alive for a thousand vertices), i.e. the latency of one
// vertex times its length -- and with one thread per path a vertex is ~24 primitive tests one after the other (Cornell
// box: 14 primitives, closest-hit + any-hit).  Here every lane of the warp carries the same path (same state, same
// random numbers, same shading arithmetic: uniform control flow, no extra time) and the two traversals test the
// primitives of a leaf 32 at a time (traverseWarp): same answers, same counters, a third of the latency per vertex.
// Lane 0 alone writes.  Pool state is read with ld.global.cg: the per-channel restart of chromatic aberration re-reads
// what lane 0 just wrote.
template <bool MIS, bool COUNT>
__global__ void __launch_bounds__(128) k_drain_warp(DScene sc, Pool pool, Batch bt, Counters *ctr) {
    uint32_t nRays = 0, nShadow = 0, nDone = 0; TraceCounters cnt;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warpsInGrid = gridDim.x * (blockDim.x >> 5), wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (uint32_t base = wid * 32u; base < pool.P; base += warpsInGrid * 32u) {
        const uint32_t mine = base + lane;
        uint32_t live = __ballot_sync(0xffffffffu, mine < pool.P && (__ldcg(&pool.flags[mine]) & PF_ALIVE));
        while (live) {
            const uint32_t slot = base + (uint32_t) __ffs(live) - 1u; live &= live - 1u;
            while (__ldcg(&pool.flags[slot]) & PF_ALIVE) {           // one pass per colour channel (one pass normally)
                const float4 ro = __ldcg(&pool.rayO[slot]), rd = __ldcg(&pool.rayD[slot]), th = __ldcg(&pool.thr[slot]), ra = __ldcg(&pool.rad[slot]);
                const uint32_t sid = __ldcg(&pool.sid[slot]);
                PathState st;
                st.o = mk(ro.x, ro.y, ro.z); st.d = mk(rd.x, rd.y, rd.z);
                st.thr = mk(th.x, th.y, th.z); st.pdf_mat = th.w; st.rad = mk(ra.x, ra.y, ra.z);
                st.flags = __ldcg(&pool.flags[slot]);
                const uint32_t chBits = st.flags & PF_CH_MASK;
                st.rng.state = __ldcg(&pool.rng[slot]); st.rng.inc = ((uint64_t) (sid % bt.wh) << 1u) | 1u;
                float mint = ro.w, maxt = rd.w;
                while (true) {
                    Hit h; if (lane == 0) ++nRays;
                    if (!traverseWarp<false, COUNT>(sc, st.o, st.d, mint, maxt, h, cnt)) break;
                    VertexOut out;
                    pathVertex<-1, MIS>(sc, h, st, out);
                    if (MIS && (COUNT || !nullContribution(out.contrib.x, out.contrib.y, out.contrib.z))) {   // uniform over the warp
                        Hit sh; if (lane == 0) ++nShadow;
                        if (!traverseWarp<true, COUNT>(sc, out.shadow.o, out.shadow.d, out.shadow.mint, out.shadow.maxt, sh, cnt))
                            st.rad = st.rad + out.contrib;
                    }
                    if (!(st.flags & PF_ALIVE)) break;
                    mint = out.next.mint; maxt = out.next.maxt;     // st.o / st.d already hold the next ray
                }
                if (lane == 0) endOfPath(sc, pool, bt, ctr, slot, sid, st.rad, st.rng.state, chBits, nDone);
                __threadfence();
                __syncwarp();
            }
        }
    }
    warpAdd(&ctr->rays_ext, nRays); warpAdd(&ctr->done, nDone);
    if (MIS) warpAdd(&ctr->rays_sh, nShadow);
    if (COUNT) { warpAdd(&ctr->nodes_ext, cnt.nodes); warpAdd(&ctr->prims_ext, cnt.prims); }
}

void noriLaunchDrain(bool mis, bool count, int grid, cudaStream_t st, const DScene &sc, const Pool &pool, const Batch &bt, Counters *ctr) {
    if (grid < 0) {                                         // one warp per path (option drain_mode = 0, the default)
        grid = -grid;
        if (mis) { if (count) k_drain_warp<true, true><<<grid, 128, 0, st>>>(sc, pool, bt, ctr); else k_drain_warp<true, false><<<grid, 128, 0, st>>>(sc, pool, bt, ctr); }
        else { if (count) k_drain_warp<false, true><<<grid, 128, 0, st>>>(sc, pool, bt, ctr); else k_drain_warp<false, false><<<grid, 128, 0, st>>>(sc, pool, bt, ctr); }
        return;
    }
    if (mis) { if (count) k_drain<true, true><<<grid, 128, 0, st>>>(sc, pool, bt, ctr); else k_drain<true, false><<<grid, 128, 0, st>>>(sc, pool, bt, ctr); }
    else { if (count) k_drain<false, true><<<grid, 128, 0, st>>>(sc, pool, bt, ctr); else k_drain<false, false><<<grid, 128, 0, st>>>(sc, pool, bt, ctr); }
}

