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
                if (MIS) {
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

void noriLaunchDrain(bool mis, bool count, int grid, cudaStream_t st, const DScene &sc, const Pool &pool, const Batch &bt, Counters *ctr) {
    if (mis) { if (count) k_drain<true, true><<<grid, 128, 0, st>>>(sc, pool, bt, ctr); else k_drain<true, false><<<grid, 128, 0, st>>>(sc, pool, bt, ctr); }
    else { if (count) k_drain<false, true><<<grid, 128, 0, st>>>(sc, pool, bt, ctr); else k_drain<false, false><<<grid, 128, 0, st>>>(sc, pool, bt, ctr); }
}

