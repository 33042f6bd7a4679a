// mega.cu -- k_mega: one thread per camera sample for the short integrators (normals, av, direct*), for scenes with a
// Perlin-noise sphere, and as the bit-exact cross-check of the wavefront scheduler (option "megakernel").
#define NORI_WITH_PERLIN 1      // see traverse.cuh
#include "kernels.cuh"

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
        V3 L, weight;
        if (hasChromaticAberrations(sc.camera)) {                  // render.cpp:106-121: one path per colour channel
            L = mk(0.f);
            for (int ch = 0; ch < 3; ++ch) {
                Ray ray = cameraRay(sc.camera, ps, ap, ch, weight);
                V3 v = weight * liDispatch<COUNT>(sc, rng, ray, rs);
                L = ch == 0 ? v : L + v;
            }
        } else {
            Ray ray = cameraRay(sc.camera, ps, ap, -1, weight);
            L = liDispatch<COUNT>(sc, rng, ray, rs);
        }
        finalizePath(bt, ctr, sid, L);
    }
    warpAdd(&ctr->rays_ext, rs.rays - rs.shadow); warpAdd(&ctr->rays_sh, rs.shadow);
    if (COUNT) { warpAdd(&ctr->nodes_ext, rs.cnt.nodes); warpAdd(&ctr->prims_ext, rs.cnt.prims); }
}

void noriLaunchMega(bool count, unsigned grid, cudaStream_t st, const DScene &sc, const Batch &bt, Counters *ctr, unsigned long long total) {
    if (count) k_mega<true><<<grid, 128, 0, st>>>(sc, bt, ctr, total);
    else k_mega<false><<<grid, 128, 0, st>>>(sc, bt, ctr, total);
}
