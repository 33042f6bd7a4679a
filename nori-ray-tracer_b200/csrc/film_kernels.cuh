// film_kernels.cuh -- film accumulation / resolve kernels and the ABI's test-hook kernels (included by nori_gpu.cu only)
#pragma once
#include "kernels.cuh"
#include <cuda.h>               // CUtensorMap (the type only: the encoder is fetched through cudaGetDriverEntryPoint)

// ---- TMA plumbing of the film kernel: one cp.async.bulk.tensor.3d per spp layer brings the (32 + 2 halo)^2 tile of the
// [layers][H][W] float4 sample tensor into shared memory (out-of-bounds elements arrive as zeros: the image border needs
// no branches), completion is signalled on an mbarrier; two stages, so layer k + 1 is in flight while layer k is
// weighted and gathered.
__device__ __forceinline__ uint32_t smemAddr(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smemAddr(bar)), "r"(count));
}
__device__ __forceinline__ void mbarExpectTx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra LAB_WAIT;\nDONE:\n}"
                 :: "r"(smemAddr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmaLoad3D(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(smemAddr(dst)), "l"((uint64_t) map), "r"(smemAddr(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
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

// Same accumulation with the separable filter weights tabulated ONCE per staged sample instead of once per
// (sample, film pixel) pair: the thread that stages a sample also evaluates its 2*HALO+1 column weights and
// 2*HALO+1 row weights (window test of block.cpp:107-110 and table lookup of :114-117, in the reference's
// block-relative arithmetic; 0 outside the window), and a film pixel then costs two shared-memory weight
// reads and the reference's `value * wx * wy` products per candidate sample.  A weight of 0 adds +0 to the
// sums, which leaves them bit-identical to skipping the sample (invalid samples are staged as zeros already).
// Shared memory per CTA: (32+2*HALO)^2 x (16 + 8*(2*HALO+1)) bytes = 72.6 KB for the default Gaussian.
// Round 2: NPIX film pixels per thread (a vertical run; 32 x 32/NPIX threads per 32 x 32 tile).  Vertical neighbours share
// four of their five sample rows, so a thread reads (NPIX + 4) x 5 staged values and column weights for NPIX x 25
// contributions instead of NPIX x 25 (the kernel is bound by the shared-memory pipe and by issue slots), and fewer threads
// stage the 36 x 36 samples in more passes at 84 % occupancy instead of 1024 threads in two at 63 %.  Each pixel still adds
// its contributions row by row, left to right: same bits.
#ifndef NORI_FILM_NPIX
#define NORI_FILM_NPIX 2
#endif
// TMA = true: the sample tile of every layer is brought in by the TMA unit (see above) into one of two stages and read
// from there by the weighting pass and by the gather -- no per-thread global loads, no copy of the values; `tmap`
// describes bt.results as a {4 W, H, layers} float tensor with a {4 S, S, 1} box.
template <bool VARIANCE, int HALO, bool TMA>
__global__ void __launch_bounds__(1024 / NORI_FILM_NPIX) k_film_sep(FilmParams fp, Batch bt, uint32_t nLayers, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) float4 s_mem[];
    constexpr int T = 32, S = T + 2 * HALO, nS = S * S, NW = 2 * HALO + 1, NPIX = NORI_FILM_NPIX, NT = T * T / NPIX;
    float4 *s_val = s_mem;                                             // TMA: two stages of nS values; else one
    float *s_wx = (float *) (s_mem + (TMA ? 2 : 1) * nS), *s_wy = s_wx + nS * NW;
    __shared__ float s_table[NORI_FILTER_RESOLUTION + 1];
    __shared__ uint64_t s_bar[2];
    const int tid = threadIdx.y * T + threadIdx.x;
    if (tid <= NORI_FILTER_RESOLUTION) s_table[tid] = fp.table[tid];
    if (TMA && tid == 0) {
        mbarInit(&s_bar[0], 1); mbarInit(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int b = fp.border;                                           // == HALO
    const int fx = blockIdx.x * T + threadIdx.x, fy0 = blockIdx.y * T + NPIX * threadIdx.y;   // film pixels (fx, fy0 .. fy0 + NPIX - 1)
    const int sx0 = blockIdx.x * T - b - HALO, sy0 = blockIdx.y * T - b - HALO;
    const int fcols = fp.W + 2 * b, frows = fp.H + 2 * b;
    float4 acc[NPIX];
    float3 vs[NPIX], vs2[NPIX];
#pragma unroll
    for (int p = 0; p < NPIX; ++p) {
        acc[p] = make_float4(0.f, 0.f, 0.f, 0.f); vs[p] = make_float3(0.f, 0.f, 0.f); vs2[p] = vs[p];
        if (VARIANCE && fx < fcols && fy0 + p < frows) acc[p] = fp.film[(size_t) (fy0 + p) * fcols + fx];
    }
    const bool own = fx < fcols && fy0 < frows;                        // at least the first pixel is inside the film
    if (TMA) {
        __syncthreads();                                               // the barriers are initialised
        if (tid == 0 && nLayers > 0) { mbarExpectTx(&s_bar[0], nS * 16); tmaLoad3D(s_val, &tmap, &s_bar[0], 4 * sx0, sy0, 0); }
    }
    for (uint32_t k = 0; k < nLayers; ++k) {
        __syncthreads();                                               // everybody is done with layer k - 1: its stage and the weights are free
        const float4 *tile = s_val + (TMA ? (k & 1u) * nS : 0);
        if (TMA) {
            if (tid == 0 && k + 1 < nLayers) {
                mbarExpectTx(&s_bar[(k + 1) & 1u], nS * 16);
                tmaLoad3D(s_val + ((k + 1) & 1u) * nS, &tmap, &s_bar[(k + 1) & 1u], 4 * sx0, sy0, (int) (k + 1));
            }
            mbarWait(&s_bar[k & 1u], (k >> 1) & 1u);
        }
        for (int i = tid; i < nS; i += NT) {
            const int ly = i / S, lx = i - ly * S;
            const int sx = sx0 + lx, sy = sy0 + ly;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            float wx[NW], wy[NW];
#pragma unroll
            for (int j = 0; j < NW; ++j) { wx[j] = 0.f; wy[j] = 0.f; }
            if (sx >= 0 && sx < fp.W && sy >= 0 && sy < fp.H) {
                const uint32_t pix = (uint32_t) sy * fp.W + sx;
                if (!TMA) v = bt.results[(size_t) k * bt.wh + pix];
                Pcg32 rng; rng.seed(bt.seed + bt.spp_first + k, (uint64_t) pix);
                P2 a = rng.next2D();
                const float psx = (float) sx + a.x, psy = (float) sy + a.y;
                // block.cpp:101-104 with the offset of the 32x32 block that rendered the sample
                const int ox = sx & ~(NORI_BLOCK_SIZE - 1), oy = sy & ~(NORI_BLOCK_SIZE - 1);
                const float px = __fsub_rn(__fsub_rn(psx, 0.5f), (float) (ox - b));
                const float py = __fsub_rn(__fsub_rn(psy, 0.5f), (float) (oy - b));
                const float xlo = __fsub_rn(px, fp.radius), xhi = __fadd_rn(px, fp.radius);
                const float ylo = __fsub_rn(py, fp.radius), yhi = __fadd_rn(py, fp.radius);
#pragma unroll
                for (int j = 0; j < NW; ++j) {                          // film column / row (sx + b + j - HALO) in block coordinates
                    const float xb = (float) (sx - ox + b + j - HALO), yb = (float) (sy - oy + b + j - HALO);
                    if (!(xb < xlo || xb > xhi)) wx[j] = s_table[(int) __fmul_rn(fabsf(__fsub_rn(xb, px)), fp.lookupFactor)];
                    if (!(yb < ylo || yb > yhi)) wy[j] = s_table[(int) __fmul_rn(fabsf(__fsub_rn(yb, py)), fp.lookupFactor)];
                }
            }
            if (!TMA) s_val[i] = v;
#pragma unroll
            for (int j = 0; j < NW; ++j) { s_wx[i * NW + j] = wx[j]; s_wy[i * NW + j] = wy[j]; }
        }
        __syncthreads();
        if (own) {
            const int cx = threadIdx.x + HALO, cy = NPIX * threadIdx.y + HALO;    // own source pixel of (fx, fy0)
            // sample rows cy - HALO .. cy + HALO + NPIX - 1: row e contributes to pixel p with dy = e - p when |e - p| <= HALO;
            // each pixel sees its rows in ascending order, its columns left to right
#pragma unroll
            for (int e = -HALO; e <= HALO + NPIX - 1; ++e) {
#pragma unroll
                for (int dx = -HALO; dx <= HALO; ++dx) {
                    const int i = (cy + e) * S + (cx + dx);
                    const float wx = s_wx[i * NW + (HALO - dx)];
                    const float4 v = tile[i];
#pragma unroll
                    for (int p = 0; p < NPIX; ++p) {
                        if (e - p >= -HALO && e - p <= HALO) {
                            const float wy = s_wy[i * NW + (HALO - (e - p))];
                            acc[p].x = __fadd_rn(acc[p].x, __fmul_rn(__fmul_rn(v.x, wx), wy));
                            acc[p].y = __fadd_rn(acc[p].y, __fmul_rn(__fmul_rn(v.y, wx), wy));
                            acc[p].z = __fadd_rn(acc[p].z, __fmul_rn(__fmul_rn(v.z, wx), wy));
                            acc[p].w = __fadd_rn(acc[p].w, __fmul_rn(__fmul_rn(v.w, wx), wy));
                        }
                    }
                }
            }
            if (VARIANCE) {                                      // Color4f::divideByFilterWeight (color.h:84-89)
#pragma unroll
                for (int p = 0; p < NPIX; ++p) {
                    const float4 a = acc[p];
                    const float mx = a.w != 0.f ? a.x / a.w : 0.f, my = a.w != 0.f ? a.y / a.w : 0.f, mz = a.w != 0.f ? a.z / a.w : 0.f;
                    vs[p].x += mx; vs[p].y += my; vs[p].z += mz; vs2[p].x += mx * mx; vs2[p].y += my * my; vs2[p].z += mz * mz;
                }
            }
        }
    }
#pragma unroll
    for (int p = 0; p < NPIX; ++p) {
        const int fy = fy0 + p;
        if (!(fx < fcols && fy < frows)) continue;
        float4 *dst = &fp.film[(size_t) fy * fcols + fx];
        if (VARIANCE) {
            *dst = acc[p];
            float4 a = fp.vsum[(size_t) fy * fcols + fx], b2 = fp.vsum2[(size_t) fy * fcols + fx];
            a.x += vs[p].x; a.y += vs[p].y; a.z += vs[p].z; b2.x += vs2[p].x; b2.y += vs2[p].y; b2.z += vs2[p].z;
            fp.vsum[(size_t) fy * fcols + fx] = a; fp.vsum2[(size_t) fy * fcols + fx] = b2;
        } else {
            float4 f = *dst;
            f.x += acc[p].x; f.y += acc[p].y; f.z += acc[p].z; f.w += acc[p].w;
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
    Hit h; TraceCounters cnt;
    bool found = traverse<SHADOW, true, true>(sc, mk(r.o[0], r.o[1], r.o[2]), mk(r.d[0], r.d[1], r.d[2]), r.mint, r.maxt, h, cnt, sc.ordered != 0);
    nori_gpu_hit o;
    o.t = h.t; o.u = h.u; o.v = h.v; o.shape = NORI_NO_HIT; o.prim = NORI_NO_HIT;
    o.nodes_visited = cnt.nodes; o.prims_tested = cnt.prims; o.reserved = 0;
    if (found && !SHADOW) {
        o.prim = __float_as_uint(sc.prims[3 * h.leafpos].w);
        o.shape = __float_as_uint(sc.prims[3 * h.leafpos + 1].w);
    }
    out[i] = o;
}

// nori_gpu_trace through the LARGE-SCENE RENDER KERNELS (option "trace_kernel" = 2): the caller's rays are loaded into
// path-pool slots exactly as k_shade leaves them (closest-hit: an alive path with its next ray; any-hit: a deferred
// NEE ray with a pending contribution of 1), k_extend_sm / k_shadow_sm run one iteration over the pool with the
// configured child order and node layout, and the answers are read back from the slots.  The batch has no camera
// samples (total_samples = 0), so the kernels regenerate nothing.
template <bool SHADOW>
__global__ void k_trace_load(Pool pool, const nori_gpu_ray *rays, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pool.P) return;
    if (i >= n) { pool.sid[i] = NORI_FREE_SLOT; pool.flags[i] = 0u; return; }
    const nori_gpu_ray r = rays[i];
    pool.rayO[i] = make_float4(r.o[0], r.o[1], r.o[2], r.mint);
    pool.hit[i] = make_float4(__int_as_float(0x7f800000), 0.f, 0.f, __uint_as_float(NORI_NO_HIT));
    pool.thr[i] = make_float4(1.f, 1.f, 1.f, 0.f);
    pool.rad[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    pool.rng[i] = 0ull; pool.sid[i] = i;
    if (SHADOW) {                                           // k_shadow_sm: origin = rayO, mint = Epsilon, direction / far end = shD
        pool.shD[i] = make_float4(r.d[0], r.d[1], r.d[2], r.maxt);
        pool.shC[i] = make_float4(1.f, 1.f, 1.f, 0.f);
        pool.flags[i] = PF_ALIVE | PF_SHADOW;
    } else {
        pool.rayD[i] = make_float4(r.d[0], r.d[1], r.d[2], r.maxt);
        pool.flags[i] = PF_ALIVE | PF_FIRST;
    }
}
template <bool SHADOW>
__global__ void k_trace_store(DScene sc, Pool pool, uint32_t n, nori_gpu_hit *out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    nori_gpu_hit o;
    o.t = __int_as_float(0x7f800000); o.u = 0.f; o.v = 0.f; o.shape = NORI_NO_HIT; o.prim = NORI_NO_HIT;
    o.nodes_visited = 0; o.prims_tested = 0; o.reserved = 0;   // per-ray counters do not exist here: see get_kernel_stats
    if (SHADOW) { if (pool.rad[i].x == 0.f) o.t = 0.f; }        // the pending contribution was added iff the ray is free
    else if (pool.flags[i] & PF_ALIVE) {                        // a miss ended the path (missRule); a hit left its record
        const float4 h = pool.hit[i];
        const uint32_t leafpos = __float_as_uint(h.w);
        o.t = h.x; o.u = h.y; o.v = h.z;
        o.prim = __float_as_uint(sc.prims[3 * leafpos].w);
        o.shape = __float_as_uint(sc.prims[3 * leafpos + 1].w);
    }
    out[i] = o;
    pool.sid[i] = NORI_FREE_SLOT; pool.flags[i] = 0u;           // hand the pool back empty
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
    P2 uv; uv.x = q[6]; uv.y = q[7];
    BRec e = mkBRec(sc, b, mk(q[0], q[1], q[2]), M_SOLID_ANGLE, uv); e.wo = mk(q[3], q[4], q[5]);
    V3 ev = bsdfEvalDyn(b, e); float pdf = bsdfPdfDyn(b, e);
    BRec r = mkBRec(sc, b, e.wi, M_UNKNOWN, uv); P2 s; s.x = q[8]; s.y = q[9];
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

// nori_gpu_selftest: the slow-path-free IEEE sequences of device_common.cuh / traverse.cuh against the compiler's own
// __fdiv_rn / __fsqrt_rn / __frcp_rn on pseudo-random operands with exponents in [-60, 60] (pcg32 bit patterns).
__global__ void k_selftest(unsigned long long n, unsigned long long *mismatch) {
    const unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Pcg32 r; r.seed(i, 12345u);
    auto operand = [&]() {
        const uint32_t m = r.nextUInt(), e = 127u - 60u + r.nextUInt() % 121u;
        return __uint_as_float((m & 0x807fffffu) | (e << 23));
    };
    const float a = operand(), b = operand();
    if (__float_as_uint(xdiv_nr(a, b)) != __float_as_uint(__fdiv_rn(a, b))) atomicAdd(&mismatch[0], 1ull);
    if (__float_as_uint(xdiv_nr(0.0f, b)) != __float_as_uint(__fdiv_rn(0.0f, b))) atomicAdd(&mismatch[0], 1ull);
    const V3 v = xdivs_nr(mk(a, 1.0f, -a), b);
    if (__float_as_uint(v.x) != __float_as_uint(__fdiv_rn(a, b)) || __float_as_uint(v.y) != __float_as_uint(__fdiv_rn(1.0f, b))
        || __float_as_uint(v.z) != __float_as_uint(__fdiv_rn(-a, b))) atomicAdd(&mismatch[0], 1ull);
    if (__float_as_uint(xsqrt_nr(fabsf(a))) != __float_as_uint(__fsqrt_rn(fabsf(a)))) atomicAdd(&mismatch[1], 1ull);
    if (__float_as_uint(rcpNormalRange(b)) != __float_as_uint(__frcp_rn(b))) atomicAdd(&mismatch[2], 1ull);
}

__global__ void k_fill_u32(uint32_t *p, uint32_t v, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void k_flush(float4 *buf, size_t n) {     // bench helper: evict L2 between timed steps
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t) gridDim.x * blockDim.x;
    for (; i < n; i += stride) buf[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}
