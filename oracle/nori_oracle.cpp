/* nori_oracle.cpp -- CPU restatement of the reference's rendering hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the shipped product (the CUDA library behind
 * include/nori_gpu.h) links, loads or calls this file; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may.
 *
 * What it is: a plain scalar C++ re-statement (no Eigen, no TBB, no virtual dispatch) of the
 * algorithm the reference (francois141/nori-ray-tracer) runs per camera sample, consuming the same
 * flat scene description (include/nori_gpu.h) the GPU path consumes.  Every function cites the
 * reference file:line it follows.  Float arithmetic follows the reference build (x86-64 SSE2, no
 * FMA: compile with -ffp-contract=off) including Eigen 3.2.90's evaluation order for 3-vectors:
 * dot(a,b) = a0*b0 + (a1*b1 + a2*b2)  (Redux.h redux_novec_unroller), cross as OrthoMethods.h:36-38,
 * normalized() = v / sqrt(squaredNorm) with true division (Dot.h:114-120).
 *
 * Pinning (see tests/test_oracle_golden.py, tests/golden/):
 *   - traversal: bit-exact (t,u,v,shape,prim,#nodes,#prims) against ray batches answered by the real
 *     reference (oracle/_ref/nori_export), incl. the special cases of the slab test (--special);
 *   - pcg32: ext/pcg32/pcg32-demo.out known answers;
 *   - whole renders: RNG mode 1 below replays the reference's block-sequential sampler mapping
 *     (one pcg32 per 32x32 block, seeded with the block offset, independent.cpp:48-53,
 *     render.cpp:96-99) so an oracle render is comparable PIXEL BY PIXEL with a render of the real
 *     reference binary (oracle/_ref/nori_ref);
 *   - the reference's t-test known answers (scenes/pa4/tests/*.xml, scenes/pa3/tests/*.xml).
 * RNG mode 0 is the GPU path's mapping (one stream per camera path), which makes oracle and GPU
 * comparable SAMPLE BY SAMPLE.
 *
 * One function is NOT a restatement of the reference: rayIntersectOrdered (nori_oracle_trace_ordered) walks the
 * same tree with the children of every inner node in other orders, to test on the CPU that the reference-order
 * answer does not depend on the visiting order (what the GPU kernels rely on for large scenes).  Nothing else
 * uses it; in particular no render and no pin above goes through it.
 */
#include "nori_gpu.h"
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <atomic>
#include <functional>
#include <string>
#include <thread>
#include <vector>

namespace {

constexpr float kEps = 1e-4f;                         /* common.h:52 */
constexpr float kPi = 3.14159265358979323846f;        /* common.h:57 -- M_PI is a FLOAT literal in the reference */
constexpr float kInvPi = 0.31830988618379067154f;     /* common.h:58 */
constexpr float kInvFourPi = 0.07957747154594766788f; /* common.h:60 */
constexpr float kInf = std::numeric_limits<float>::infinity();


/* minimal work-sharing loop (std::thread; no OpenMP/TBB dependency). NORI_ORACLE_THREADS overrides. */
inline int oracleThreads() {
    const char *e = getenv("NORI_ORACLE_THREADS");
    int n = e ? atoi(e) : (int) std::thread::hardware_concurrency();
    return n < 1 ? 1 : n;
}
template <typename F> void parallelFor(int64_t n, int64_t grain, F fn) {
    int nt = oracleThreads();
    if (nt == 1 || n <= grain) { for (int64_t i = 0; i < n; ++i) fn(i); return; }
    std::atomic<int64_t> next(0);
    std::vector<std::thread> ts;
    for (int t = 0; t < nt; ++t) ts.emplace_back([&]() {
        for (;;) { int64_t b = next.fetch_add(grain); if (b >= n) break; for (int64_t i = b; i < std::min(n, b + grain); ++i) fn(i); }
    });
    for (auto &t : ts) t.join();
}

/* ------------------------------------------------------------------ 3-vectors, Eigen order --- */
struct V3 { float x = 0, y = 0, z = 0; V3() {} V3(float a, float b, float c) : x(a), y(b), z(c) {} explicit V3(float a) : x(a), y(a), z(a) {}
            float operator[](int i) const { return i == 0 ? x : i == 1 ? y : z; } };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3 &operator+=(V3 &a, V3 b) { a = a + b; return a; }
inline float dot(V3 a, V3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
inline float sqnorm(V3 a) { return dot(a, a); }
inline float norm(V3 a) { return std::sqrt(sqnorm(a)); }
inline V3 normalized(V3 a) { return a / norm(a); }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline V3 load3(const float *p) { return {p[0], p[1], p[2]}; }
/* Mesh::getInterpolatedNormal / setHitInformation normalise a DYNAMIC-size Eigen expression (sums of
 * MatrixXf columns, mesh.cpp:63-73,147-160): Eigen's non-unrolled redux adds left to right,
 * (x*x + y*y) + z*z -- unlike the fixed-size order used everywhere else.  Measured with the probes. */
inline V3 normalizedDyn(V3 a) { float n = std::sqrt((a.x * a.x + a.y * a.y) + a.z * a.z); return a / n; }

struct P2 { float x = 0, y = 0; };

/* ------------------------------------------------------------------ pcg32 (ext/pcg32/pcg32.h:38-110) */
struct Pcg32 {
    uint64_t state, inc;
    void seed(uint64_t initstate, uint64_t initseq) {         /* pcg32.h:51-57 */
        state = 0; inc = (initseq << 1) | 1u; nextUInt(); state += initstate; nextUInt();
    }
    uint32_t nextUInt() {                                      /* pcg32.h:60-66 */
        uint64_t old = state;
        state = old * 0x5851f42d4c957f2dULL + inc;
        uint32_t xs = (uint32_t) (((old >> 18) ^ old) >> 27), rot = (uint32_t) (old >> 59);
        return (xs >> rot) | (xs << ((~rot + 1u) & 31));
    }
    float nextFloat() {                                        /* pcg32.h:101-110 */
        uint32_t u = (nextUInt() >> 9) | 0x3f800000u; float f; memcpy(&f, &u, 4); return f - 1.0f;
    }
    float next1D() { return nextFloat(); }                     /* independent.cpp:58-60 */
    /* independent.cpp:62-67 builds Point2f(nextFloat(), nextFloat()); the order of the two calls is
     * unspecified in C++ and GCC (the reference's and this build's compiler) evaluates constructor
     * arguments right to left: the FIRST draw lands in y.  Measured with nori_export --seq. */
    P2 next2D() { P2 p; p.y = nextFloat(); p.x = nextFloat(); return p; }
};

/* ------------------------------------------------------------------ scene copy ---------------- */
struct Shape {
    nori_gpu_shape pod;
    std::vector<float> V, N, UV, cdf; std::vector<uint32_t> F;
};
struct Emitter {
    nori_gpu_emitter pod;
    std::vector<float> image, pdf, cdf, pm, cm;
};
struct Image { int w = 0, h = 0, wrap = 0; std::vector<uint8_t> rgb; };
struct Scene {
    nori_gpu_scene pod;
    std::vector<Image> images;
    std::vector<nori_gpu_bvh_node> nodes; std::vector<uint32_t> indices, shapeOffset;
    std::vector<Shape> shapes; std::vector<nori_gpu_bsdf> bsdfs; std::vector<Emitter> emitters;
    int border = 0; float lookupFactor = 0;
    /* counters (single-threaded use or relaxed adds; only used for reporting) */
    uint64_t rays = 0, shadowRays = 0, nodesVisited = 0, primsTested = 0, samples = 0, invalid = 0;
};

struct Ray {                                /* ray.h:38-100 */
    V3 o, d, dRcp; float mint = kEps, maxt = kInf;
    Ray() {}
    Ray(V3 o_, V3 d_) : o(o_), d(d_) { update(); }
    Ray(V3 o_, V3 d_, float a, float b) : o(o_), d(d_), mint(a), maxt(b) { update(); }
    void update() { dRcp = V3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z); }     /* cwiseInverse, ray.h:73-75 */
    V3 at(float t) const { return o + t * d; }
};

struct Frame { V3 s, t, n; };               /* frame.h:33-60 */
struct Its {                                /* shape.h:38-67 */
    V3 p; float t = kInf; P2 uv; Frame sh, geo; int shape = -1; uint32_t prim = 0;
    float baryU = 0, baryV = 0; uint32_t nodes = 0, prims = 0;
};

/* common.cpp:274-283 */
inline void coordinateSystem(V3 a, V3 &b, V3 &c) {
    if (std::abs(a.x) > std::abs(a.y)) {
        float invLen = 1.0f / std::sqrt(a.x * a.x + a.z * a.z);
        c = V3(a.z * invLen, 0.0f, -a.x * invLen);
    } else {
        float invLen = 1.0f / std::sqrt(a.y * a.y + a.z * a.z);
        c = V3(0.0f, a.z * invLen, -a.y * invLen);
    }
    b = cross(c, a);
}
inline Frame makeFrame(V3 n) { Frame f; f.n = n; coordinateSystem(n, f.s, f.t); return f; }   /* frame.h:49-51 */
inline V3 toLocal(const Frame &f, V3 v) { return {dot(v, f.s), dot(v, f.t), dot(v, f.n)}; }   /* frame.h:54-58 */
inline V3 toWorld(const Frame &f, V3 v) { return f.s * v.x + f.t * v.y + f.n * v.z; }          /* frame.h:61-63 */

/* ------------------------------------------------------------------ intersection kernels ------ */
/* bbox.h:336-363 */
inline bool boxHit(const nori_gpu_bvh_node &nd, const Ray &r) {
    float nearT = -kInf, farT = kInf;
    for (int i = 0; i < 3; i++) {
        float origin = r.o[i], minVal = nd.bmin[i], maxVal = nd.bmax[i];
        if (r.d[i] == 0) { if (origin < minVal || origin > maxVal) return false; }
        else {
            float t1 = (minVal - origin) * r.dRcp[i], t2 = (maxVal - origin) * r.dRcp[i];
            if (t1 > t2) std::swap(t1, t2);
            nearT = std::max(t1, nearT); farT = std::min(t2, farT);
            if (!(nearT <= farT)) return false;
        }
    }
    return r.mint <= farT && nearT <= r.maxt;
}

/* mesh.cpp:83-120 */
inline bool triHit(const Shape &m, uint32_t index, const Ray &ray, float &u, float &v, float &t) {
    uint32_t i0 = m.F[3 * index], i1 = m.F[3 * index + 1], i2 = m.F[3 * index + 2];
    V3 p0 = load3(&m.V[3 * i0]), p1 = load3(&m.V[3 * i1]), p2 = load3(&m.V[3 * i2]);
    V3 edge1 = p1 - p0, edge2 = p2 - p0;
    V3 pvec = cross(ray.d, edge2);
    float det = dot(edge1, pvec);
    if (det > -1e-8f && det < 1e-8f) return false;
    float inv_det = 1.0f / det;
    V3 tvec = ray.o - p0;
    u = dot(tvec, pvec) * inv_det;
    if (u < 0.0 || u > 1.0) return false;
    V3 qvec = cross(tvec, edge1);
    v = dot(ray.d, qvec) * inv_det;
    if (v < 0.0 || u + v > 1.0) return false;
    t = dot(edge2, qvec) * inv_det;
    return t >= ray.mint && t <= ray.maxt;
}

/* sphere.cpp:43-76 (b is formed in double: 2.0 * float, exact) */
inline bool sphereHit(const Shape &s, const Ray &ray, float &t) {
    V3 oc = ray.o - load3(s.pod.center);
    float a = dot(ray.d, ray.d);
    float b = (float) (2.0 * dot(oc, ray.d));
    float c = dot(oc, oc) - s.pod.radius * s.pod.radius;
    float discriminant = (b * b - 4 * a * c);
    if (!(discriminant > 0)) return false;
    float delta = std::sqrt(b * b - 4 * a * c);
    float t1 = (-b - delta) / (2 * a), t2 = (-b + delta) / (2 * a);
    if (ray.mint <= t1 && t1 <= ray.maxt) { t = t1; return true; }
    if (ray.mint <= t2 && t2 <= ray.maxt) { t = t2; return true; }
    return false;
}

/* ---- PerlinSphere (perlinnoise.cpp:25-203): a sphere whose radius is perturbed by 2D value noise of the hit
 * point.  The integer hash overflows `int` in the reference (wraps on x86); unsigned arithmetic gives the same
 * bits.  Interpolation runs in double (cos(double), literals 0.5 / 1073741824.0), octave weights are
 * pow(2, i) (double) and pow(2.0f, (float) i) (float) -- both exact powers of two. */
inline float perlinNoise(int x, int y) {                                       /* :200-204 */
    uint32_t n = (uint32_t) x + (uint32_t) y * 57u;
    n = (n << 13) ^ n;
    int32_t v = (int32_t) ((n * ((n * n * 15731u) + 789221u) + 1376312589u) & 0x7fffffffu);
    return (float) (1.0f - v / 1073741824.0);
}
inline float perlinBilinear(int x, int y) {                                    /* :193-197 */
    return ((perlinNoise(x - 1, y) + perlinNoise(x + 1, y) + perlinNoise(x, y - 1) + perlinNoise(x, y + 1)) / 8.0f) +
           ((perlinNoise(x - 1, y - 1) + perlinNoise(x + 1, y - 1) + perlinNoise(x - 1, y + 1) + perlinNoise(x + 1, y + 1)) / 16.0f) +
           (perlinNoise(x, y) / 4.0f);
}
inline float perlinCosine(float a, float b, float x) {                        /* :186-190 */
    double ft = x * kPi;
    double f = (1 - std::cos(ft)) * 0.5;
    return (float) (a * (1 - f) + b * f);
}
inline float perlinInterpolated(float x, float y) {                           /* :167-183 */
    int x_ = (int) x, y_ = (int) y;
    float dx = x - x_, dy = y - y_;
    float v0 = perlinBilinear(x_, y_), v1 = perlinBilinear(x_ + 1, y_), v2 = perlinBilinear(x_, y_ + 1), v3 = perlinBilinear(x_ + 1, y_ + 1);
    float i0 = perlinCosine(v0, v1, dx), i1 = perlinCosine(v2, v3, dx);
    return perlinCosine(i0, i1, dy);
}
inline float perlinNoisedRadius(const nori_gpu_shape &s, V3 p) {              /* :143-164 */
    float res = 0.0f, freq = 1.0f / s.perlin_height, amp = 1.0f, lac = 2.0f;
    for (int i = 0; i < 9; ++i) {
        res += perlinInterpolated(p.x * freq, p.y * freq) * amp;
        freq = (float) std::pow(2, i);
        amp = std::pow(lac, static_cast<float>(i));
    }
    float r_scale = res / 256.0f;
    return s.radius + s.perlin_scale * std::min(std::max(0.0f, r_scale), 1.0f);
}
inline size_t solveQuadratic(float a, float b, float c, float *t0, float *t1) { /* common.h:251-268 */
    float delta = (b * b) - (4 * a * c);
    if (delta < 0) return 0;
    if (delta == 0) { *t0 = (-b) / (2 * a); return 1; }
    *t0 = ((-b) + std::sqrt(delta)) / (2 * a);
    *t1 = ((-b) - std::sqrt(delta)) / (2 * a);
    return 2;
}
/* the root selection shared by rayIntersect (:36-58) and perlinRayIntersect (:118-139); note `t < maxt` */
inline bool perlinPick(size_t n, float t0, float t1, const Ray &ray, float &t) {
    if (n == 0) return false;
    if (n == 1) { t = t0; return ray.mint <= t && t < ray.maxt; }
    t = std::min(t0, t1);
    if (ray.mint <= t && t < ray.maxt) return true;
    t = std::max(t0, t1);
    return ray.mint <= t && t < ray.maxt;
}
inline bool perlinHit(const Shape &s, const Ray &ray, float &t) {             /* :25-59 + :106-140 */
    V3 c = load3(s.pod.center);
    V3 oc = ray.o - c;
    float a = dot(ray.d, ray.d);
    float b = (float) (2.0 * dot(oc, ray.d));
    float cc = dot(oc, oc) - s.pod.radius * s.pod.radius;
    float t0 = 0, t1 = 0;
    size_t n = solveQuadratic(a, b, cc, &t0, &t1);
    if (!perlinPick(n, t0, t1, ray, t)) return false;        /* n == 1 with an invalid root also returns false (:41-47) */
    V3 its_p = ray.at(t);
    float r = perlinNoisedRadius(s.pod, its_p);
    float c2 = dot(oc, oc) - r * r;
    n = solveQuadratic(a, b, c2, &t0, &t1);
    return perlinPick(n, t0, t1, ray, t);
}

/* bvh.h:105-109 */
inline uint32_t findShape(const Scene &sc, uint32_t &idx) {
    auto it = std::lower_bound(sc.shapeOffset.begin(), sc.shapeOffset.end(), idx + 1) - 1;
    idx -= *it;
    return (uint32_t) (it - sc.shapeOffset.begin());
}

/* ImageTexture::getData / NormalMap::getData (imagetexture.cpp:98-116, normalmap.cpp:98-120): the texel at
 * the truncated coordinates, wrapped with C's % (Repeat) or clamped.  A negative remainder indexes in
 * front of the array in the reference (undefined behaviour); here it is wrapped into range instead. */
inline V3 imageTexel(const Image &im, float xf, float yf) {
    int x, y;
    if (im.wrap == NORI_WRAP_REPEAT) {
        x = (int) xf % im.w; y = (int) yf % im.h;
        if (x < 0) x += im.w;
        if (y < 0) y += im.h;
    } else {
        x = std::min(std::max((int) xf, 0), im.w - 1); y = std::min(std::max((int) yf, 0), im.h - 1);
    }
    const uint8_t *t = &im.rgb[((size_t) x + (size_t) im.w * y) * 3];
    return V3((float) t[0] / 255, (float) t[1] / 255, (float) t[2] / 255);
}
/* ImageTexture::eval / NormalMap::eval (imagetexture.cpp:118-136, normalmap.cpp:122-139).  `x` is a float,
 * not a floor, so dstdx = uv.x * w - x is exactly 0 and the "bilinear" blend is 1*1*v00 + 1*0*v01 +
 * 0*1*v10 + 0*0*v11: the value of the texel at the truncated coordinates.  The blend is evaluated as
 * written so that the result is bit-identical (0 * finite terms, added in the same order). */
template <typename F> inline V3 imageEval(const Image &im, P2 uv, F map) {
    float x = uv.x * im.w, y = uv.y * im.h;
    V3 v00 = map(imageTexel(im, x, y)), v01 = map(imageTexel(im, x, y + 1.0f)), v10 = map(imageTexel(im, x + 1.0f, y)), v11 = map(imageTexel(im, x + 1.0f, y + 1.0f));
    float dstdx = uv.x * im.w - x, dstdy = uv.y * im.h - y;
    return (((1.0f - dstdx) * (1.0f - dstdy) * v00 + (1.0f - dstdx) * dstdy * v01) + dstdx * (1.0f - dstdy) * v10) + dstdx * dstdy * v11;
}

/* mesh.cpp:122-170, sphere.cpp:78-93 */
void setHitInformation(const Scene &sc, const Ray &ray, Its &its) {
    const Shape &m = sc.shapes[its.shape];
    if (m.pod.type == NORI_SHAPE_MESH) {
        float b1 = its.baryU, b2 = its.baryV, b0 = 1 - (b1 + b2);
        uint32_t i0 = m.F[3 * its.prim], i1 = m.F[3 * its.prim + 1], i2 = m.F[3 * its.prim + 2];
        V3 p0 = load3(&m.V[3 * i0]), p1 = load3(&m.V[3 * i1]), p2 = load3(&m.V[3 * i2]);
        its.p = (b0 * p0 + b1 * p1) + b2 * p2;
        its.uv.x = b1; its.uv.y = b2;
        if (!m.UV.empty()) {
            its.uv.x = (b0 * m.UV[2 * i0] + b1 * m.UV[2 * i1]) + b2 * m.UV[2 * i2];
            its.uv.y = (b0 * m.UV[2 * i0 + 1] + b1 * m.UV[2 * i1 + 1]) + b2 * m.UV[2 * i2 + 1];
        }
        its.geo = makeFrame(normalized(cross(p1 - p0, p2 - p0)));
        if (!m.N.empty()) {
            V3 n = (b0 * load3(&m.N[3 * i0]) + b1 * load3(&m.N[3 * i1])) + b2 * load3(&m.N[3 * i2]);
            its.sh = makeFrame(normalizedDyn(n));
            if (m.pod.normal_map > 0) {                                        /* mesh.cpp:149-154 */
                V3 nm = imageEval(sc.images[m.pod.normal_map - 1], its.uv,
                                  [](V3 c) { return V3(2.0f * c.x - 1.0f, 2.0f * c.y - 1.0f, 2.0f * c.z - 1.0f); });   /* normalmap.cpp:115-119 */
                its.sh = makeFrame(toWorld(its.sh, normalized(nm)));
            }
        } else its.sh = its.geo;
    } else if (m.pod.type == NORI_SHAPE_PERLIN) {                              /* perlinnoise.cpp:61-75 */
        V3 c = load3(m.pod.center);
        its.p = ray.o + its.t * ray.d;
        V3 n = normalized(its.p - c);
        its.sh = its.geo = makeFrame(n);
        float th = std::acos(n.z), ph = std::atan2(n.y, n.x);
        if (ph < 0) ph += 2 * kPi;
        its.uv.x = (float) (0.5 + th * (0.15915494309189533577f));             /* INV_TWOPI, common.h:61 */
        its.uv.y = ph * kInvPi;
    } else {
        /* by now ray.maxt == its.t (bvh.cpp:444) */
        V3 c = load3(m.pod.center);
        its.p = ray.o + its.t * ray.d;
        V3 n = normalized(its.p - c);
        its.sh = its.geo = makeFrame(n);
        /* sphericalCoordinates, common.cpp:264-272 */
        float th = std::acos(n.z), ph = std::atan2(n.y, n.x);
        if (ph < 0) ph += 2 * kPi;
        its.uv.x = (float) (0.5 + th / (2 * kPi));
        its.uv.y = ph / kPi;
    }
}

/* bvh.cpp:404-462 */
bool rayIntersect(Scene &sc, const Ray &_ray, Its &its, bool shadowRay, bool count = true) {
    uint32_t node_idx = 0, stack_idx = 0, stack[64];
    its.t = kInf; its.shape = -1; its.nodes = its.prims = 0;
    Ray ray(_ray);
    if (count) { __atomic_fetch_add(&sc.rays, 1, __ATOMIC_RELAXED); if (shadowRay) __atomic_fetch_add(&sc.shadowRays, 1, __ATOMIC_RELAXED); }
    if (ray.mint == kEps)
        ray.mint = std::max(ray.mint, ray.mint * std::max(std::abs(ray.o.x), std::max(std::abs(ray.o.y), std::abs(ray.o.z))));
    if (sc.nodes.empty() || ray.maxt < ray.mint) return false;
    bool found = false;
    while (true) {
        const nori_gpu_bvh_node &node = sc.nodes[node_idx];
        ++its.nodes;
        if (!boxHit(node, ray)) {
            if (stack_idx == 0) break;
            node_idx = stack[--stack_idx];
            continue;
        }
        if (!(node.data[0] & 1u)) {                       /* inner: left child first, always */
            stack[stack_idx++] = node.data[1];
            node_idx++;
        } else {
            uint32_t start = node.data[1], end = start + (node.data[0] >> 1);
            for (uint32_t i = start; i < end; ++i) {
                uint32_t idx = sc.indices[i];
                uint32_t s = findShape(sc, idx);
                float u = 0, v = 0, t = 0;
                ++its.prims;
                const Shape &shp = sc.shapes[s];
                bool hit = shp.pod.type == NORI_SHAPE_MESH ? triHit(shp, idx, ray, u, v, t)
                         : shp.pod.type == NORI_SHAPE_PERLIN ? perlinHit(shp, ray, t) : sphereHit(shp, ray, t);
                if (hit) {
                    if (shadowRay) { its.t = 0; goto done_shadow; }
                    found = true;
                    ray.maxt = its.t = t;
                    its.baryU = shp.pod.type == NORI_SHAPE_MESH ? u : 0.f;
                    its.baryV = shp.pod.type == NORI_SHAPE_MESH ? v : 0.f;
                    its.shape = (int) s; its.prim = idx;
                }
            }
            if (stack_idx == 0) break;
            node_idx = stack[--stack_idx];
        }
    }
    if (count) { __atomic_fetch_add(&sc.nodesVisited, its.nodes, __ATOMIC_RELAXED); __atomic_fetch_add(&sc.primsTested, its.prims, __ATOMIC_RELAXED); }
    if (found) setHitInformation(sc, ray, its);
    return found;
done_shadow:
    if (count) { __atomic_fetch_add(&sc.nodesVisited, its.nodes, __ATOMIC_RELAXED); __atomic_fetch_add(&sc.primsTested, its.prims, __ATOMIC_RELAXED); }
    return true;
}
inline bool occluded(Scene &sc, const Ray &r) { Its tmp; return rayIntersect(sc, r, tmp, true); }   /* scene.h:112-115 */

/* NOT the reference's algorithm: the same closest-hit query with the children of every inner node taken in
 * another order -- what the GPU kernels do on large scenes (near child first) and two adversarial orders -- to
 * test on the CPU, at scale, that the answer does not depend on the visiting order once exact ties are resolved
 * the way the reference's order resolves them (equal t: the higher leaf position wins).
 * order 1: the child on the ray's side of the node's recorded split axis first (traverse.cuh: descend());
 * order 2: that child LAST; order 3: always the right child first. */
bool rayIntersectOrdered(Scene &sc, const Ray &_ray, Its &its, int order) {
    uint32_t node_idx = 0, stack_idx = 0, stack[64];
    its.t = kInf; its.shape = -1; its.nodes = its.prims = 0;
    Ray ray(_ray);
    if (ray.mint == kEps)
        ray.mint = std::max(ray.mint, ray.mint * std::max(std::abs(ray.o.x), std::max(std::abs(ray.o.y), std::abs(ray.o.z))));
    if (sc.nodes.empty() || ray.maxt < ray.mint) return false;
    bool found = false; uint32_t bestPos = 0;
    while (true) {
        const nori_gpu_bvh_node &node = sc.nodes[node_idx];
        ++its.nodes;
        bool descend = false;
        if (boxHit(node, ray)) {
            if (!(node.data[0] & 1u)) {
                const uint32_t axis = (node.data[0] >> 1) & 3u, left = node_idx + 1, right = node.data[1];
                const float da = axis == 0 ? ray.d.x : axis == 1 ? ray.d.y : ray.d.z;
                bool rightFirst = order == 3 ? true : (da < 0.0f);
                if (order == 2) rightFirst = !rightFirst;
                stack[stack_idx++] = rightFirst ? left : right;
                node_idx = rightFirst ? right : left;
                descend = true;
            } else {
                uint32_t start = node.data[1], end = start + (node.data[0] >> 1);
                for (uint32_t i = start; i < end; ++i) {
                    uint32_t idx = sc.indices[i];
                    uint32_t s = findShape(sc, idx);
                    float u = 0, v = 0, t = 0;
                    ++its.prims;
                    const Shape &shp = sc.shapes[s];
                    bool hit = shp.pod.type == NORI_SHAPE_MESH ? triHit(shp, idx, ray, u, v, t)
                             : shp.pod.type == NORI_SHAPE_PERLIN ? perlinHit(shp, ray, t) : sphereHit(shp, ray, t);
                    if (hit && (!found || t < ray.maxt || i > bestPos)) {      /* the kernels' tie rule */
                        found = true; bestPos = i;
                        ray.maxt = its.t = t;
                        its.baryU = shp.pod.type == NORI_SHAPE_MESH ? u : 0.f;
                        its.baryV = shp.pod.type == NORI_SHAPE_MESH ? v : 0.f;
                        its.shape = (int) s; its.prim = idx;
                    }
                }
            }
        }
        if (descend) continue;
        if (stack_idx == 0) break;
        node_idx = stack[--stack_idx];
    }
    return found;
}

/* ------------------------------------------------------------------ warps (src/warp.cpp) ------- */
inline V3 squareToUniformSphere(P2 s) {                       /* warp.cpp:86-91 */
    float theta = std::acos(1 - 2 * (1 - s.x));
    float phi = 2.f * kPi * s.y;
    return V3(std::sin(theta) * std::cos(phi), std::sin(theta) * std::sin(phi), std::cos(theta));
}
inline V3 squareToCosineHemisphere(P2 s) {                    /* warp.cpp:110-115 */
    float theta = std::acos(std::sqrt(1 - (1 - s.x)));
    float phi = 2.f * kPi * s.y;
    return V3(std::sin(theta) * std::cos(phi), std::sin(theta) * std::sin(phi), std::cos(theta));
}
inline V3 squareToBeckmann(P2 s, float alpha) {               /* warp.cpp:122-127 */
    float theta = (float) std::atan(std::sqrt(-std::pow((double) alpha, 2) * std::log(1 - s.x)));
    float phi = 2 * kPi * s.y;
    return V3(std::sin(theta) * std::cos(phi), std::sin(theta) * std::sin(phi), std::cos(theta));
}
inline V3 squareToUniformTriangle(P2 s) {                     /* warp.cpp:135-140 */
    float su1 = sqrtf(s.x); float u = 1.f - su1, v = s.y * su1;
    return V3(u, v, 1.f - u - v);
}
inline P2 squareToConcentricDisk(P2 s) {                      /* warp.cpp:143-162 */
    float ox = 2.f * s.x - 1.f, oy = 2.f * s.y - 1.f; P2 r;
    if (ox == 0 && oy == 0) return r;
    float theta, rad;
    if (std::abs(ox) > std::abs(oy)) { rad = ox; theta = kPi * 0.25f * (oy / ox); }
    else { rad = oy; theta = kPi * 0.5f - kPi * 0.25f * (ox / oy); }
    r.x = rad * std::cos(theta); r.y = rad * std::sin(theta); return r;
}
inline V3 squareToGTR2(P2 s, float alpha) {                   /* warp.cpp:180-185 */
    float a2 = (float) std::pow((double) alpha, 2);
    float theta = std::acos(std::sqrt((1.0f - s.x) / (1.0f + (a2 - 1.0f) * s.x)));
    float phi = 2 * kPi * s.y;
    return V3(std::sin(theta) * std::cos(phi), std::sin(theta) * std::sin(phi), std::cos(theta));
}
inline float squareToGTR2Pdf(V3 m, float alpha) {             /* warp.cpp:187-193 */
    float a2 = (float) std::pow((double) alpha, 2);
    float cosTheta = m.z;
    float pdf = (float) (a2 * cosTheta * kInvPi / std::pow(1 + (a2 - 1.0f) * std::pow((double) cosTheta, 2), 2));
    return (cosTheta >= 0 && std::abs(sqnorm(m) - 1.0f) < 1.0f) ? pdf : 0.0f;
}

/* common.cpp:285-314 */
float fresnel(float cosThetaI, float extIOR, float intIOR) {
    float etaI = extIOR, etaT = intIOR;
    if (extIOR == intIOR) return 0.0f;
    if (cosThetaI < 0.0f) { std::swap(etaI, etaT); cosThetaI = -cosThetaI; }
    float eta = etaI / etaT, sinThetaTSqr = eta * eta * (1 - cosThetaI * cosThetaI);
    if (sinThetaTSqr > 1.0f) return 1.0f;
    float cosThetaT = std::sqrt(1.0f - sinThetaTSqr);
    float Rs = (etaI * cosThetaI - etaT * cosThetaT) / (etaI * cosThetaI + etaT * cosThetaT);
    float Rp = (etaT * cosThetaI - etaI * cosThetaT) / (etaT * cosThetaI + etaI * cosThetaT);
    return (Rs * Rs + Rp * Rp) / 2.0f;
}

inline float tanTheta(V3 v) { float temp = 1 - v.z * v.z; if (temp <= 0.0f) return 0.0f; return std::sqrt(temp) / v.z; }  /* frame.h:81-86 */

/* ------------------------------------------------------------------ BSDFs ---------------------- */
enum Measure { EUnknown = 0, ESolidAngle = 1, EDiscrete = 2 };
struct BRec { V3 wi, wo; float eta = 0; int measure = EUnknown; P2 uv; const Scene *scene = nullptr; };     /* bsdf.h:30-58 (+ the texture table) */

inline V3 albedoAt(const nori_gpu_bsdf &b, const BRec &r) {
    const P2 uv = r.uv;
    if (b.albedo_texture == NORI_TEXTURE_IMAGE)                                /* imagetexture.cpp:118-136 */
        return imageEval(r.scene->images[b.albedo_image], uv, [](V3 c) { return c; });
    if (b.albedo_texture == NORI_TEXTURE_CHECKERBOARD) {                      /* checkerboard.cpp:31-37 */
        int x = (int) std::abs(std::floor(uv.x / b.tex_scale[0] - b.tex_delta[0]));
        int y = (int) std::abs(std::floor(uv.y / b.tex_scale[1] - b.tex_delta[1]));
        return x % 2 == y % 2 ? load3(b.albedo) : load3(b.albedo2);
    }
    return load3(b.albedo);                                                   /* consttexture.cpp:30-32 */
}

/* microfacet.cpp:52-82 */
inline float evalBeckmann(float alpha, V3 m) {
    float temp = tanTheta(m) / alpha, ct = m.z, ct2 = ct * ct;
    return std::exp(-temp * temp) / (kPi * alpha * alpha * ct2 * ct2);
}
inline float smithBeckmannG1(float alpha, V3 v, V3 m) {
    float tt = tanTheta(v);
    if (tt == 0.0f) return 1.0f;
    if (dot(m, v) * v.z <= 0) return 0.0f;
    float a = 1.0f / (alpha * tt);
    if (a >= 1.6f) return 1.0f;
    float a2 = a * a;
    return (3.535f * a + 2.181f * a2) / (1.0f + 2.276f * a + 2.577f * a2);
}

/* disney.cpp:26-44 */
inline float schlickFresnel(float u) { float m = std::min(1.0f, std::max(0.0f, 1 - u)); return (float) std::pow((double) m, 5); }
inline float ggx(float NdotV, float alphaG) { float a = alphaG * alphaG, b = NdotV * NdotV; return 1 / (NdotV + std::sqrt(a + b - a * b)); }
inline V3 lerp3(float t, V3 a, V3 b) { return (1.0f - t) * a + t * b; }
inline float luminance(V3 c) { return c.x * 0.212671f + c.y * 0.715160f + c.z * 0.072169f; }   /* common.cpp:233-235 */

V3 bsdfEval(const nori_gpu_bsdf &b, const BRec &r) {
    switch (b.type) {
    case NORI_BSDF_DIFFUSE:                                                   /* diffuse.cpp:72-82 */
        if (r.measure != ESolidAngle || r.wi.z <= 0 || r.wo.z <= 0) return V3(0.f);
        return albedoAt(b, r) * kInvPi;
    case NORI_BSDF_MICROFACET: {                                              /* microfacet.cpp:84-94 */
        V3 n = normalized(r.wi + r.wo);
        float D = evalBeckmann(b.alpha, n);
        float F = fresnel(dot(n, r.wi), b.extIOR, b.intIOR);
        float G = smithBeckmannG1(b.alpha, r.wi, n) * smithBeckmannG1(b.alpha, r.wo, n);
        float denom = 4.0f * r.wi.z * r.wo.z;
        float spec = b.ks * D * F * G / denom;
        V3 kd = load3(b.kd) * kInvPi;
        return V3(kd.x + spec, kd.y + spec, kd.z + spec);
    }
    case NORI_BSDF_DISNEY: {                                                  /* disney.cpp:63-105 */
        float NdotV = r.wi.z, NdotL = r.wo.z;
        if (NdotV < 0 || NdotL < 0) return V3(0.f);
        V3 wh = normalized(r.wi + r.wo);
        float LdotH = dot(r.wo, wh), VdotH = dot(r.wi, wh);
        V3 base = load3(b.baseColor), white(1.f);
        float lum = luminance(base);
        V3 Ctint = lum > 0.f ? V3(base.x / lum, base.y / lum, base.z / lum) : V3(1.0f);
        V3 CtintMix = (float) (b.specular * 0.08) * lerp3(b.specularTint, white, Ctint);
        V3 Cspec = lerp3(b.metallic, CtintMix, base);
        float fd90 = (float) (0.5 + 2 * b.roughness * std::pow((double) VdotH, 2));
        float fl = schlickFresnel(NdotL), fv = schlickFresnel(NdotV);
        V3 diffuse = base * kInvPi * (1.f + (fd90 - 1.f) * fl) * (1.f + (fd90 - 1.f) * fv);
        float alpha = std::max(0.001f, b.roughness * b.roughness);
        float Ds = squareToGTR2Pdf(wh, alpha);
        float FH = schlickFresnel(LdotH);
        V3 Fs = lerp3(FH, Cspec, white);
        float Gs = ggx(NdotL, alpha) * ggx(NdotV, alpha);
        V3 specular = Gs * Fs * Ds;
        V3 Fsheen = FH * b.sheen * lerp3(b.sheenTint, white, Ctint);
        return (1 - b.metallic) * (diffuse + Fsheen) + specular;
    }
    default: return V3(0.f);                                                  /* mirror.cpp:29-32, dielectric.cpp:35-38 */
    }
}

float bsdfPdf(const nori_gpu_bsdf &b, const BRec &r) {
    switch (b.type) {
    case NORI_BSDF_DIFFUSE:                                                   /* diffuse.cpp:85-101 */
        if (r.measure != ESolidAngle || r.wi.z <= 0 || r.wo.z <= 0) return 0.0f;
        return kInvPi * r.wo.z;
    case NORI_BSDF_MICROFACET: {                                              /* microfacet.cpp:97-111 */
        float c = r.wo.z; if (c <= 0.0f) return 0.0f;
        V3 n = normalized(r.wi + r.wo);
        float metallicTerm = evalBeckmann(b.alpha, n) * n.z / (4.0f * std::abs(dot(n, r.wo)));
        return b.ks * metallicTerm + (1 - b.ks) * (c * kInvPi);
    }
    case NORI_BSDF_DISNEY: {                                                  /* disney.cpp:108-121 */
        float c = r.wo.z; if (c <= 0.0f) return 0.0f;
        V3 n = normalized(r.wi + r.wo);
        float metallicTerm = squareToGTR2Pdf(n, b.alpha) * n.z / (4.0f * std::abs(dot(n, r.wo)));
        return (1 - b.metallic) * (c * kInvPi) + b.metallic * metallicTerm;
    }
    default: return 0.0f;
    }
}

V3 bsdfSample(const nori_gpu_bsdf &b, BRec &r, P2 s) {
    switch (b.type) {
    case NORI_BSDF_DIFFUSE:                                                   /* diffuse.cpp:104-120 */
        if (r.wi.z <= 0) return V3(0.f);
        r.measure = ESolidAngle; r.wo = squareToCosineHemisphere(s); r.eta = 1.0f;
        return albedoAt(b, r);
    case NORI_BSDF_MIRROR:                                                    /* mirror.cpp:39-55 */
        if (r.wi.z <= 0) return V3(0.f);
        r.wo = V3(-r.wi.x, -r.wi.y, r.wi.z); r.measure = EDiscrete; r.eta = 1.0f;
        return V3(1.f);
    case NORI_BSDF_DIELECTRIC: {                                              /* dielectric.cpp:45-73 */
        float theta = r.wi.z; V3 nv(0, 0, 1.0f);
        if (fresnel(theta, b.extIOR, b.intIOR) > s.x) { r.eta = 1.0f; r.wo = V3(-r.wi.x, -r.wi.y, r.wi.z); }
        else {
            float factor = b.extIOR / b.intIOR;
            if (theta < 0.0f) { factor = 1 / factor; nv.z *= -1; }
            float win = dot(r.wi, nv);
            V3 part1 = -factor * (r.wi - win * nv);
            V3 part2 = -nv * (float) std::sqrt(1 - std::pow((double) factor, 2) * (1 - std::pow((double) win, 2)));
            r.wo = normalized(part1 + part2);
            r.eta = b.extIOR / b.intIOR;
        }
        r.measure = EDiscrete;
        return V3(1.f);
    }
    case NORI_BSDF_MICROFACET: {                                              /* microfacet.cpp:114-137 */
        if (r.wi.z <= 0.0f) return V3(0.f);
        if (s.x < b.ks) {
            P2 ns; ns.x = s.x / b.ks; ns.y = s.y;
            V3 n = squareToBeckmann(ns, b.alpha);
            r.wo = normalized((2.0f * dot(r.wi, n) * n) - r.wi);
        } else {
            P2 ns; ns.x = (s.x - b.ks) / (1.f - b.ks); ns.y = s.y;
            r.wo = squareToCosineHemisphere(ns);
        }
        float c = r.wo.z; if (c <= 0.f) return V3(0.f);
        return bsdfEval(b, r) * c / bsdfPdf(b, r);
    }
    case NORI_BSDF_DISNEY: {                                                  /* disney.cpp:124-145 */
        if (r.wi.z <= 0.0f) return V3(0.f);
        if (s.x <= b.metallic) {
            P2 ns; ns.x = s.x / b.metallic; ns.y = s.y;
            V3 n = squareToGTR2(ns, b.alpha);
            r.wo = normalized((2.0f * dot(r.wi, n) * n) - r.wi);
        } else {
            P2 ns; ns.x = (s.x - b.metallic) / (1 - b.metallic); ns.y = s.y;
            r.wo = squareToCosineHemisphere(ns);
        }
        float c = r.wo.z; if (c <= 0.0f) return V3(0.f);
        return bsdfEval(b, r) * c / bsdfPdf(b, r);
    }
    }
    return V3(0.f);
}

/* ------------------------------------------------------------------ emitters ------------------- */
struct ERec { V3 ref, p, n, wi; float pdf = 0; Ray shadowRay; };               /* emitter.h:31-59 */
inline ERec makeERec(V3 ref, V3 p, V3 n) { ERec e; e.ref = ref; e.p = p; e.n = n; e.wi = normalized(p - ref); return e; }

/* dpdf.h:119-157 */
inline size_t cdfSampleReuse(const std::vector<float> &cdf, float &s) {
    auto entry = std::lower_bound(cdf.begin(), cdf.end(), s);
    size_t index = (size_t) std::max((ptrdiff_t) 0, entry - cdf.begin() - 1);
    index = std::min(index, cdf.size() - 2);
    s = (s - cdf[index]) / (cdf[index + 1] - cdf[index]);
    return index;
}

/* mesh.cpp:40-61, sphere.cpp:95-105 */
void sampleSurface(const Shape &m, P2 s, V3 &p, V3 &n, float &pdf) {
    if (m.pod.type == NORI_SHAPE_MESH) {
        size_t idT = cdfSampleReuse(m.cdf, s.x);
        V3 bc = squareToUniformTriangle(s);
        uint32_t i0 = m.F[3 * idT], i1 = m.F[3 * idT + 1], i2 = m.F[3 * idT + 2];
        V3 p0 = load3(&m.V[3 * i0]), p1 = load3(&m.V[3 * i1]), p2 = load3(&m.V[3 * i2]);
        p = (bc.x * p0 + bc.y * p1) + bc.z * p2;
        if (!m.N.empty()) n = normalizedDyn((bc.x * load3(&m.N[3 * i0]) + bc.y * load3(&m.N[3 * i1])) + bc.z * load3(&m.N[3 * i2]));
        else n = normalized(cross(p1 - p0, p2 - p0));
        pdf = m.pod.area_normalization;
    } else if (m.pod.type == NORI_SHAPE_PERLIN) {                              /* perlinnoise.cpp:77-86 */
        V3 q = squareToUniformSphere(s);
        p = load3(m.pod.center) + m.pod.radius * q;
        float r = perlinNoisedRadius(m.pod, p);
        p = load3(m.pod.center) + r * q; n = q;
        pdf = (float) (std::pow((double) (1.f / r), 2) * (0.25f * kInvPi));
    } else {
        V3 q = squareToUniformSphere(s);
        p = load3(m.pod.center) + m.pod.radius * q; n = q;
        pdf = (float) (std::pow((double) (1.f / m.pod.radius), 2) * (0.25f * kInvPi));
    }
}
inline float pdfSurface(const Shape &m, V3 p) {
    if (m.pod.type == NORI_SHAPE_MESH) return m.pod.area_normalization;
    float r = m.pod.type == NORI_SHAPE_PERLIN ? perlinNoisedRadius(m.pod, p) : m.pod.radius;   /* perlinnoise.cpp:88-91 */
    return (float) (std::pow((double) (1.f / r), 2) * (0.25f * kInvPi));
}

/* envmap.cpp:60-88 */
inline P2 envMapIntersect(const Emitter &e, V3 vec) {
    float th = std::acos(vec.z), ph = std::atan2(vec.y, vec.x);
    if (ph < 0) ph += 2 * kPi;
    P2 r; r.x = th * (e.pod.env_rows - 1) * kInvPi; r.y = (float) (ph * 0.5 * (e.pod.env_cols - 1) * kInvPi);
    if (std::isnan(r.x) || std::isnan(r.y)) { r.x = r.y = 0; }
    return r;
}
inline int clampi(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }

V3 emitterEval(const Scene &sc, const Emitter &e, const ERec &l) {
    switch (e.pod.type) {
    case NORI_EMITTER_AREA:                                                   /* arealight.cpp:39-44 */
        return dot(l.n, -l.wi) > 0.0f ? load3(e.pod.radiance) : V3(0.f);
    case NORI_EMITTER_POINT:                                                  /* pointlight.cpp:26-29 */
        return load3(e.pod.radiance) / (4.f * kPi * sqnorm(load3(e.pod.position) - l.ref));
    case NORI_EMITTER_SPOT: {                                                 /* spotlight.cpp:38-42 */
        V3 c = load3(e.pod.radiance) / (4.f * kPi);
        return c * 2 * kPi * (float) (1 - 0.5 * (e.pod.cosFalloffStart + e.pod.cosTotalWidth));
    }
    case NORI_EMITTER_ENVMAP: {                                               /* envmap.cpp:124-156 */
        P2 uv = envMapIntersect(e, normalized(l.wi));
        int W = e.pod.env_rows, H = e.pod.env_cols;
        int u = clampi((int) uv.x, 0, W - 1), v = clampi((int) uv.y, 0, H - 1);
        int us = (u + 1) % W, vs = (v + 1) % H;
        auto px = [&](int i, int j) { return load3(&e.image[((size_t) i * H + j) * 3]); };
        V3 BL = px(u, v), UL = px(u, vs), BR = px(us, v), UR = px(us, vs);
        int dusu = us - u, dvsv = vs - v;
        float dusum = us - uv.x, dumu = uv.x - u, dvmv = uv.y - v, dvsvm = vs - uv.y;
        float k = (float) (1.0 / (dusu * dvsv));
        return e.pod.weight * (k * ((BL * dusum * dvsvm) + (BR * dumu * dvsvm) + (UL * dusum * dvmv) + (UR * dumu * dvmv)));
    }
    }
    return V3(0.f);
}

float emitterPdf(const Scene &sc, const Emitter &e, const ERec &l) {
    switch (e.pod.type) {
    case NORI_EMITTER_AREA:                                                   /* arealight.cpp:64-76 */
        return dot(l.n, -l.wi) > 0.0f ? pdfSurface(sc.shapes[e.pod.shape], l.p) : 0.0f;
    case NORI_EMITTER_POINT: return 1.0f;                                     /* pointlight.cpp:30-33 */
    case NORI_EMITTER_SPOT: return l.pdf;                                     /* spotlight.cpp:44-47 */
    case NORI_EMITTER_ENVMAP: {                                               /* envmap.cpp:184-192 */
        P2 its = envMapIntersect(e, normalized(l.wi));
        int i = clampi((int) its.x, 0, e.pod.env_rows - 1), j = clampi((int) its.y, 0, e.pod.env_cols - 1);
        return e.pm[i] * e.pdf[(size_t) i * e.pod.env_cols + j];
    }
    }
    return 0.f;
}

/* envmap.cpp:112-121; reads one past the table in the reference when no interval matches -- here the
 * search stops at the last valid interval instead (documented deviation) */
inline void envSample1D(const float *pfRow, const float *PfRow, int nPf, float s, float &x, float &prob) {
    int i;
    for (i = 0; i < nPf - 2; i++)
        if (PfRow[i] <= s && s < PfRow[i + 1]) break;
    float t = (PfRow[i + 1] - s) / (PfRow[i + 1] - PfRow[i]);
    x = (1 - t) * i + t * (i + 1);
    prob = pfRow[i];
}

V3 emitterSample(const Scene &sc, const Emitter &e, ERec &l, P2 s) {
    switch (e.pod.type) {
    case NORI_EMITTER_AREA: {                                                 /* arealight.cpp:46-62 */
        float spdf;
        sampleSurface(sc.shapes[e.pod.shape], s, l.p, l.n, spdf);
        l.wi = normalized(l.p - l.ref);
        l.shadowRay = Ray(l.ref, l.wi, kEps, norm(l.p - l.ref) - kEps);
        l.pdf = emitterPdf(sc, e, l);
        float att = dot(l.n, -l.wi) / sqnorm(l.p - l.ref);
        return l.pdf > 0.0f ? emitterEval(sc, e, l) * att / emitterPdf(sc, e, l) : V3(0.f);
    }
    case NORI_EMITTER_POINT: {                                                /* pointlight.cpp:15-24 */
        V3 pos = load3(e.pod.position);
        l.wi = normalized(pos - l.ref); l.p = pos; l.pdf = 1.0f;
        l.shadowRay = Ray(l.ref, l.wi, kEps, norm(pos - l.ref) - kEps);
        return load3(e.pod.radiance) / (4.f * kPi * sqnorm(pos - l.ref));
    }
    case NORI_EMITTER_SPOT: {                                                 /* spotlight.cpp:19-36 */
        V3 pos = load3(e.pod.position), dir = load3(e.pod.direction);
        l.wi = normalized(pos - l.ref); l.p = pos; l.pdf = 1.0f; l.n = dir;
        l.shadowRay = Ray(l.ref, l.wi, kEps, norm(pos - l.ref) - kEps);
        float cosTheta = dot(dir, normalized(-l.wi)), fall;
        if (cosTheta < e.pod.cosTotalWidth) fall = 0;
        else if (cosTheta > e.pod.cosFalloffStart) fall = 1;
        else fall = (std::acos(e.pod.cosTotalWidth) - std::acos(cosTheta)) / (std::acos(e.pod.cosTotalWidth) - std::acos(e.pod.cosFalloffStart));
        return load3(e.pod.radiance) * fall / (4.f * kPi * sqnorm(l.ref - l.p));
    }
    case NORI_EMITTER_ENVMAP: {                                               /* envmap.cpp:158-181 */
        int W = e.pod.env_rows, H = e.pod.env_cols;
        float st2 = 1.0f - l.wi.z * l.wi.z, sinTheta = st2 <= 0.0f ? 0.0f : std::sqrt(st2);   /* lRec.wi still 0 => 1 */
        float jacobian = (float) ((H - 1) * (W - 1) / (2 * std::pow((double) kPi, 2) * sinTheta));
        float u, v, up, vp;
        envSample1D(e.pm.data(), e.cm.data(), W + 1, s.x, u, up);
        int row = clampi((int) u, 0, W - 1);                                  /* reference indexes row (int)u unclamped */
        envSample1D(&e.pdf[(size_t) row * H], &e.cdf[(size_t) row * (H + 1)], H + 1, s.y, v, vp);
        float theta = u * kPi / (W - 1), phi = v * 2 * kPi / (H - 1);
        l.wi = normalized(V3(std::sin(theta) * std::cos(phi), std::sin(theta) * std::sin(phi), std::cos(theta)));
        l.shadowRay = Ray(l.ref, l.wi, kEps, 100000.f);
        vp = emitterPdf(sc, e, l) * jacobian;
        return emitterEval(sc, e, l) / vp;
    }
    }
    return V3(0.f);
}

/* ------------------------------------------------------------------ cameras -------------------- */
inline V3 xfPoint(const float *m, V3 p) {                                     /* transform.h:78-81 */
    float r[4];
    for (int i = 0; i < 4; ++i) r[i] = ((m[4 * i] * p.x + m[4 * i + 1] * p.y) + m[4 * i + 2] * p.z) + m[4 * i + 3] * 1.0f;
    return V3(r[0] / r[3], r[1] / r[3], r[2] / r[3]);
}
inline V3 xfVector(const float *m, V3 v) {                                    /* transform.h:68-70 */
    return V3((m[0] * v.x + m[1] * v.y) + m[2] * v.z, (m[4] * v.x + m[5] * v.y) + m[6] * v.z, (m[8] * v.x + m[9] * v.y) + m[10] * v.z);
}

/* warp.cpp:53-58 (M_PI is the reference's float literal; unqualified cos/sin/sqrt pick the float overloads,
 * as in squareToConcentricDisk above, which is pinned bit-exact) */
inline P2 squareToUniformDisk(P2 sample) {
    float angle = 2 * sample.x * kPi;
    float size = std::sqrt(sample.y);
    P2 r; r.x = std::cos(angle) * size; r.y = std::sin(angle) * size; return r;
}

/* perspective.cpp:90-112, thinlens.cpp:126-171, advancedCamera.cpp:133-228.  Returns the camera weight:
 * Color3f(1), or the unit vector of `channel` when chromatic aberration is on (advancedCamera.cpp:176-183). */
V3 sampleRay(const nori_gpu_camera &c, Ray &ray, P2 ps, P2 as, int channel = -1) {
    V3 nearP = xfPoint(c.sampleToCamera, V3(ps.x * c.invOutputSize[0], ps.y * c.invOutputSize[1], 0.0f));
    V3 d = normalized(nearP);
    V3 weight(1.0f);
    if (c.type == NORI_CAMERA_ADVANCED) {
        const bool distort = !(c.distortion[0] == 0.f && c.distortion[1] == 0.f);
        const bool chroma = !(c.chromatic[0] == 0.f && c.chromatic[1] == 0.f && c.chromatic[2] == 0.f);
        if (distort) {                                                        /* advancedCamera.cpp:145-170 */
            float qx = nearP.x / nearP.z, qy = nearP.y / nearP.z;
            float y = std::sqrt(qx * qx + qy * qy);
            float r = y, r2, f, df; int i = 0;
            while (true) {
                r2 = r * r;
                f = r * (1 + (c.distortion[0] * r2) + c.distortion[1] * (r2 * r2)) - y;
                df = 1 + (3 * c.distortion[0] * r2) + (5 * c.distortion[1] * r2 * r2);
                r = r - f / df;
                if ((double) std::abs(f) < 1e-6 || i++ > 4) break;             /* F_EPSILON is a double literal */
            }
            float distortionFactor = r / y;
            nearP.x *= distortionFactor; nearP.y *= distortionFactor;
            d = normalized(nearP);
        }
        float w = 0.0f;
        if (chroma) { w = c.chromatic[channel]; weight = V3(channel == 0 ? 1.f : 0.f, channel == 1 ? 1.f : 0.f, channel == 2 ? 1.f : 0.f); }
        float invZ = 1.0f / d.z;
        if (c.lensRadius > 0.0f || chroma) {                                  /* advancedCamera.cpp:192-216 */
            P2 disk = squareToUniformDisk(as);
            float lx = c.lensRadius * disk.x, ly = c.lensRadius * disk.y;
            float ft = c.focalDistance / d.z;
            V3 pFocus = V3(0.f) + ft * d;
            float spx = ps.x - (0.5f * c.width), spy = ps.y - (0.5f * c.height);
            float mx = (float) std::max(c.width, c.height);
            spx /= mx; spy /= mx;
            float sq = spx * spx + spy * spy;
            float dx = spx * sq * w, dy = spy * sq * w;
            pFocus = pFocus + V3(-dx, dy, 0.0f);
            V3 o(lx, ly, 0.0f);
            V3 dir = normalized(pFocus - o);
            ray.o = xfPoint(c.cameraToWorld, o); ray.d = xfVector(c.cameraToWorld, dir);
        } else {
            ray.o = xfPoint(c.cameraToWorld, V3(0, 0, 0)); ray.d = xfVector(c.cameraToWorld, d);
        }
        ray.mint = c.nearClip * invZ; ray.maxt = c.farClip * invZ;
        ray.update();
        return weight;
    }
    float invZ = 1.0f / d.z;
    if (c.type == NORI_CAMERA_THINLENS && c.lensRadius > 0.0f) {
        P2 disk = squareToConcentricDisk(as);
        float lx = c.lensRadius * disk.x, ly = c.lensRadius * disk.y;
        float ft = c.focalDistance / d.z;
        V3 pFocus = V3(0.f) + ft * d;
        V3 o(lx, ly, 0.0f);
        V3 dir = normalized(pFocus - o);
        ray.o = xfPoint(c.cameraToWorld, o); ray.d = xfVector(c.cameraToWorld, dir);
    } else {
        ray.o = xfPoint(c.cameraToWorld, V3(0, 0, 0)); ray.d = xfVector(c.cameraToWorld, d);
    }
    ray.mint = c.nearClip * invZ; ray.maxt = c.farClip * invZ;
    ray.update();
    return weight;
}
inline bool hasChromaticAberrations(const nori_gpu_camera &c) {              /* advancedCamera.cpp:230-232 */
    return c.type == NORI_CAMERA_ADVANCED && !(c.chromatic[0] == 0.f && c.chromatic[1] == 0.f && c.chromatic[2] == 0.f);
}

/* ------------------------------------------------------------------ integrators ---------------- */
inline const Emitter &randomEmitter(const Scene &sc, float rnd) {            /* scene.h:68-74 */
    size_t n = sc.emitters.size();
    size_t index = std::min(static_cast<size_t>(std::floor(n * rnd)), n - 1);
    return sc.emitters[index];
}
inline const nori_gpu_bsdf &bsdfOf(const Scene &sc, const Its &its) { return sc.bsdfs[sc.shapes[its.shape].pod.bsdf]; }
inline int emitterOf(const Scene &sc, const Its &its) { return sc.shapes[its.shape].pod.emitter; }

V3 LiPathMis(Scene &sc, Pcg32 &rng, const Ray &ray) {                         /* path_mis.cpp:17-101 */
    V3 color(0.f), att(1.f); Ray cur = ray; float w_mats = 1.0f; Its its;
    if (!rayIntersect(sc, cur, its, false)) return color;
    const size_t nLights = sc.emitters.size();
    while (true) {
        if (emitterOf(sc, its) >= 0) {
            ERec e = makeERec(cur.o, its.p, its.sh.n);
            color += att * w_mats * emitterEval(sc, sc.emitters[emitterOf(sc, its)], e);
        }
        const Emitter &light = randomEmitter(sc, rng.next1D());
        ERec e; e.ref = its.p;
        V3 Li = emitterSample(sc, light, e, rng.next2D()) * (float) nLights;
        float pdf_em = emitterPdf(sc, light, e);
        if (!occluded(sc, e.shadowRay)) {
            float theta = std::max(0.0f, toLocal(its.sh, e.wi).z);
            BRec b; b.scene = &sc; b.wi = toLocal(its.sh, -cur.d); b.wo = toLocal(its.sh, e.wi); b.measure = ESolidAngle; b.uv = its.uv;
            V3 f = bsdfEval(bsdfOf(sc, its), b);
            float pdf_mat = bsdfPdf(bsdfOf(sc, its), b);
            float w_ems = (pdf_mat + pdf_em) > 0.0f ? pdf_em / (pdf_mat + pdf_em) : pdf_em;
            color += att * w_ems * f * theta * Li;
        }
        float p = std::min(att.x, 0.99f);
        if (rng.next1D() > p) return color;
        att = att / p;
        BRec b; b.scene = &sc; b.wi = toLocal(its.sh, -cur.d); b.uv = its.uv;
        V3 w = bsdfSample(bsdfOf(sc, its), b, rng.next2D());
        att = att * w;
        cur = Ray(its.p, toWorld(its.sh, b.wo));
        float pdf_mat = bsdfPdf(bsdfOf(sc, its), b);
        V3 origin = its.p;
        if (!rayIntersect(sc, cur, its, false)) return color;
        if (emitterOf(sc, its) >= 0) {
            ERec e2 = makeERec(origin, its.p, its.sh.n);
            float pdf_em2 = emitterPdf(sc, sc.emitters[emitterOf(sc, its)], e2);
            w_mats = pdf_mat + pdf_em2 > 0.f ? pdf_mat / (pdf_mat + pdf_em2) : pdf_mat;
        }
        if (b.measure == EDiscrete) w_mats = 1.0f;
    }
}

V3 LiPathMats(Scene &sc, Pcg32 &rng, const Ray &ray) {                        /* path_mats.cpp:18-60 */
    V3 color(0.f), att(1.f); Ray cur = ray;
    while (true) {
        Its its;
        if (!rayIntersect(sc, cur, its, false)) return color;
        if (emitterOf(sc, its) >= 0) {
            ERec e = makeERec(cur.o, its.p, its.sh.n);
            color += att * emitterEval(sc, sc.emitters[emitterOf(sc, its)], e);
        }
        float p = std::min(att.x, 0.99f);
        if (rng.next1D() > p) return color;
        att = att / p;
        BRec b; b.scene = &sc; b.wi = toLocal(its.sh, -cur.d); b.uv = its.uv;
        V3 w = bsdfSample(bsdfOf(sc, its), b, rng.next2D());
        att = att * w;
        cur = Ray(its.p, toWorld(its.sh, b.wo));
    }
}

V3 LiDirect(Scene &sc, Pcg32 &rng, const Ray &ray, int kind) {
    Its its;
    if (!rayIntersect(sc, ray, its, false)) return V3(0.f);
    V3 color(0.f);
    if (kind != NORI_INTEGRATOR_DIRECT && emitterOf(sc, its) >= 0) {          /* direct_ems.cpp:27-30 etc. */
        ERec e = makeERec(ray.o, its.p, its.sh.n);
        color += emitterEval(sc, sc.emitters[emitterOf(sc, its)], e);
    }
    if (kind == NORI_INTEGRATOR_DIRECT || kind == NORI_INTEGRATOR_DIRECT_EMS || kind == NORI_INTEGRATOR_DIRECT_MIS) {
        for (const Emitter &light : sc.emitters) {                            /* direct.cpp:29-47, direct_ems.cpp:33-50, direct_mis.cpp:34-60 */
            ERec e; e.ref = its.p;
            P2 s; if (kind != NORI_INTEGRATOR_DIRECT) s = rng.next2D();       /* direct.cpp:27: zero-filled Vector2f */
            V3 traced = emitterSample(sc, light, e, s);
            float pdf_em = kind == NORI_INTEGRATOR_DIRECT_MIS ? emitterPdf(sc, light, e) : 0.f;
            if (!occluded(sc, e.shadowRay)) {
                V3 wi = toLocal(its.sh, e.wi), d = toLocal(its.sh, -ray.d);
                BRec b; b.scene = &sc; b.measure = ESolidAngle; b.uv = its.uv;
                if (kind == NORI_INTEGRATOR_DIRECT) { b.wi = wi; b.wo = d; } else { b.wi = d; b.wo = wi; }
                V3 f = bsdfEval(bsdfOf(sc, its), b);
                if (kind == NORI_INTEGRATOR_DIRECT_MIS) {
                    float pdf_mat = bsdfPdf(bsdfOf(sc, its), b);
                    float w_em = pdf_mat + pdf_em > 0.f ? pdf_em / (pdf_mat + pdf_em) : pdf_em;
                    color += w_em * f * traced * wi.z;
                } else color += f * wi.z * traced;
            }
        }
    }
    if (kind == NORI_INTEGRATOR_DIRECT_MATS || kind == NORI_INTEGRATOR_DIRECT_MIS) {   /* direct_mats.cpp:33-43, direct_mis.cpp:62-83 */
        BRec b; b.scene = &sc; b.wi = toLocal(its.sh, -ray.d); b.uv = its.uv;
        V3 w = bsdfSample(bsdfOf(sc, its), b, rng.next2D());
        float pdf_mat = kind == NORI_INTEGRATOR_DIRECT_MIS ? bsdfPdf(bsdfOf(sc, its), b) : 0.f;
        Ray nr(its.p, toWorld(its.sh, b.wo)); Its its2;
        if (rayIntersect(sc, nr, its2, false) && emitterOf(sc, its2) >= 0) {
            ERec e = makeERec(its.p, its2.p, its2.sh.n);
            const Emitter &em = sc.emitters[emitterOf(sc, its2)];
            V3 Le = emitterEval(sc, em, e);
            if (kind == NORI_INTEGRATOR_DIRECT_MIS) {
                float pdf_em = emitterPdf(sc, em, e);
                float w_mat = pdf_mat + pdf_em > 0.f ? pdf_mat / (pdf_mat + pdf_em) : 0.0f;
                color += w_mat * w * Le;
            } else color += w * Le;
        }
    }
    return color;
}

V3 LiNormals(Scene &sc, const Ray &ray) {                                     /* normals.cpp:15-23 */
    Its its; if (!rayIntersect(sc, ray, its, false)) return V3(0.f);
    return V3(std::abs(its.sh.n.x), std::abs(its.sh.n.y), std::abs(its.sh.n.z));
}

V3 LiAv(Scene &sc, Pcg32 &rng, const Ray &ray) {                              /* averagevisibility.cpp:16-25, warp.cpp:25-42 */
    Its its; if (!rayIntersect(sc, ray, its, false)) return V3(1.f);
    V3 v;
    do { v.x = 1.f - 2.f * rng.next1D(); v.y = 1.f - 2.f * rng.next1D(); v.z = 1.f - 2.f * rng.next1D(); } while (sqnorm(v) > 1.f);
    if (dot(v, its.sh.n) < 0.f) v = -v;
    v = v / norm(v);
    Ray nr(its.p, v, kEps, sc.pod.av_length);
    return occluded(sc, nr) ? V3(0.f) : V3(1.f);
}

/* ---- homogeneous medium (medium.cpp:22-94) ---- */
inline bool boundsHit(const nori_gpu_medium &m, const Ray &r, float &nearT, float &farT) {   /* bbox.h:366-392 */
    nearT = -kInf; farT = kInf;
    for (int i = 0; i < 3; i++) {
        float origin = r.o[i], minVal = m.bounds_min[i], maxVal = m.bounds_max[i];
        if (r.d[i] == 0) { if (origin < minVal || origin > maxVal) return false; }
        else {
            float t1 = (minVal - origin) * r.dRcp[i], t2 = (maxVal - origin) * r.dRcp[i];
            if (t1 > t2) std::swap(t1, t2);
            nearT = std::max(t1, nearT); farT = std::min(t2, farT);
            if (!(nearT <= farT)) return false;
        }
    }
    return true;
}
inline bool boundsContain(const nori_gpu_medium &m, V3 p) {                   /* bbox.h:115-123 */
    for (int i = 0; i < 3; ++i) if (!(p[i] >= m.bounds_min[i] && p[i] <= m.bounds_max[i])) return false;
    return true;
}
V3 mediumTr(const nori_gpu_medium &m, V3 src, V3 dst) {                       /* medium.cpp:22-57 */
    float nearT, farT;
    Ray ray(src, normalized(dst - src));
    if (!boundsHit(m, ray, nearT, farT)) return V3(1.0f);
    V3 sp = boundsContain(m, src) ? src : src + normalized(ray.d) * nearT;
    V3 ep = boundsContain(m, dst) ? dst : src + normalized(ray.d) * farT;
    float len = norm(ep - sp);
    V3 ext = load3(m.sigma_a) + load3(m.sigma_s);
    return V3(std::exp(-ext.x * len), std::exp(-ext.y * len), std::exp(-ext.z * len));
}
V3 mediumSample(const nori_gpu_medium &m, const Ray &ray, Pcg32 &rng, float tMax, bool &hitObject, V3 &p) {   /* medium.cpp:59-90 */
    float nearT, farT;
    if (!boundsHit(m, ray, nearT, farT)) { hitObject = true; return V3(1.0f); }
    V3 sp = boundsContain(m, ray.o) ? ray.o : ray.o + normalized(ray.d) * nearT;
    V3 ext = load3(m.sigma_a) + load3(m.sigma_s);
    float invTr = -1.0f * std::log(1 - rng.next1D()) / std::max(ext.x, std::max(ext.y, ext.z));   /* medium.cpp:92-94 */
    float distance = norm(sp - ray.o) + invTr;
    V3 albedo(m.sigma_s[0] / ext.x, m.sigma_s[1] / ext.y, m.sigma_s[2] / ext.z);
    if (distance >= tMax) hitObject = true; else { p = ray.at(distance); hitObject = false; }
    return albedo;
}

V3 LiVolumetric(Scene &sc, Pcg32 &rng, const Ray &ray) {                      /* volumetric.cpp:18-156 */
    const nori_gpu_medium &med = sc.pod.medium;
    V3 color(0.f), att(1.f); Ray cur = ray; float w_mats = 1.0f; Its its;
    const size_t nLights = sc.emitters.size();
    bool intersection = rayIntersect(sc, cur, its, false);
    while (true) {
        float tmax = intersection ? norm(its.p - cur.o) : its.t;
        bool hitObject; V3 mp;
        V3 sampled = mediumSample(med, cur, rng, tmax, hitObject, mp);
        if (!hitObject) {
            V3 wo = squareToUniformSphere(rng.next2D()); float pdf_mat = kInvFourPi;   /* phasefunction.cpp:13-16 */
            const Emitter &light = randomEmitter(sc, rng.next1D());
            ERec e; e.ref = mp;
            V3 Li = emitterSample(sc, light, e, rng.next2D()) * (float) nLights;
            att = att * sampled;
            Its tmp;
            if (!rayIntersect(sc, e.shadowRay, tmp, false))                    /* closest-hit query, volumetric.cpp:63 */
                color += att * mediumTr(med, mp, e.p) * Li * pdf_mat;
            float p = std::min(att.x, 0.80f);
            if (rng.next1D() > p) return color;
            att = att / p;
            cur = Ray(mp, normalized(wo));
            intersection = rayIntersect(sc, cur, its, false);
            if (intersection && emitterOf(sc, its) >= 0) {
                ERec l = makeERec(cur.o, its.p, its.sh.n);
                float pdf_em = emitterPdf(sc, sc.emitters[emitterOf(sc, its)], l);
                w_mats = pdf_mat + pdf_em > 0.f ? pdf_mat / (pdf_mat + pdf_em) : pdf_mat;
            }
        } else if (intersection) {
            if (emitterOf(sc, its) >= 0) {
                ERec e = makeERec(cur.o, its.p, its.sh.n);
                color += att * w_mats * emitterEval(sc, sc.emitters[emitterOf(sc, its)], e) * mediumTr(med, its.p, e.p);
            }
            const Emitter &light = randomEmitter(sc, rng.next1D());
            ERec e; e.ref = its.p;
            V3 Li = emitterSample(sc, light, e, rng.next2D()) * (float) nLights;
            if (!occluded(sc, e.shadowRay)) {
                float pdf_em = emitterPdf(sc, light, e);
                float theta = std::max(0.0f, toLocal(its.sh, e.wi).z);
                BRec b; b.scene = &sc; b.wi = toLocal(its.sh, -cur.d); b.wo = toLocal(its.sh, e.wi); b.measure = ESolidAngle;   /* uv left default, :103 */
                V3 f = bsdfEval(bsdfOf(sc, its), b);
                float pdf_mat = bsdfPdf(bsdfOf(sc, its), b);
                float w_ems = (pdf_mat + pdf_em) > 0.0f ? pdf_em / (pdf_mat + pdf_em) : pdf_em;
                color += att * w_ems * f * theta * Li * mediumTr(med, its.p, e.p);
            }
            float p = std::min(att.x, 0.80f);
            if (rng.next1D() > p) return color;
            att = att / p;
            BRec b; b.scene = &sc; b.wi = toLocal(its.sh, -cur.d);
            V3 w = bsdfSample(bsdfOf(sc, its), b, rng.next2D());
            att = att * w;
            float pdf_mat = bsdfPdf(bsdfOf(sc, its), b);
            cur = Ray(its.p, toWorld(its.sh, b.wo));
            intersection = rayIntersect(sc, cur, its, false);
            if (intersection) {
                if (emitterOf(sc, its) >= 0) {
                    ERec l = makeERec(cur.o, its.p, its.sh.n);
                    float pdf_em = emitterPdf(sc, sc.emitters[emitterOf(sc, its)], l);
                    w_mats = pdf_mat + pdf_em > 0.f ? pdf_mat / (pdf_mat + pdf_em) : pdf_mat;
                }
                if (b.measure == EDiscrete) w_mats = 1.0f;
            }
        } else break;
    }
    return color;
}

V3 Li(Scene &sc, Pcg32 &rng, const Ray &ray) {
    switch (sc.pod.integrator) {
    case NORI_INTEGRATOR_NORMALS: return LiNormals(sc, ray);
    case NORI_INTEGRATOR_PATH_MIS: return LiPathMis(sc, rng, ray);
    case NORI_INTEGRATOR_PATH_MATS: return LiPathMats(sc, rng, ray);
    case NORI_INTEGRATOR_AV: return LiAv(sc, rng, ray);
    case NORI_INTEGRATOR_VOLUMETRIC: return LiVolumetric(sc, rng, ray);
    default: return LiDirect(sc, rng, ray, sc.pod.integrator);
    }
}

/* one iteration of renderBlock's inner loop (render.cpp:98-126) */
inline V3 cameraSample(Scene &sc, Pcg32 &rng, int px, int py, P2 &pixelSample) {
    P2 a = rng.next2D(); pixelSample.x = (float) px + a.x; pixelSample.y = (float) py + a.y;
    P2 aperture = rng.next2D();
    if (hasChromaticAberrations(sc.pod.camera)) {     /* render.cpp:106-121: one path per colour channel, one sampler */
        V3 value(0.f);
        for (int ch = 0; ch < 3; ++ch) {
            Ray rc; V3 wgt = sampleRay(sc.pod.camera, rc, pixelSample, aperture, ch);
            value += wgt * Li(sc, rng, rc);
        }
        return value;
    }
    Ray ray; V3 wgt = sampleRay(sc.pod.camera, ray, pixelSample, aperture);
    return wgt * Li(sc, rng, ray);
}

inline bool validColor(V3 c) {                                                /* common.cpp:224-231 */
    for (int i = 0; i < 3; ++i) { float v = c[i]; if (v < 0 || !std::isfinite(v)) return false; }
    return true;
}

/* ImageBlock::put(pos, value), block.cpp:93-122, for a block at offset (ox,oy) of size (bw,bh) */
void blockPut(const Scene &sc, float *blk, int ox, int oy, int bw, int bh, P2 pos_, V3 value) {
    if (!validColor(value)) return;
    const int b = sc.border; const float r = sc.pod.filter.radius;
    int cols = bw + 2 * b, rows = bh + 2 * b;
    float px = pos_.x - 0.5f - (ox - b), py = pos_.y - 0.5f - (oy - b);
    int x0 = std::max(0, (int) std::ceil(px - r)), y0 = std::max(0, (int) std::ceil(py - r));
    int x1 = std::min(cols - 1, (int) std::floor(px + r)), y1 = std::min(rows - 1, (int) std::floor(py + r));
    float wx[64], wy[64];
    for (int x = x0, i = 0; x <= x1; ++x) wx[i++] = sc.pod.filter.table[(int) (std::abs(x - px) * sc.lookupFactor)];
    for (int y = y0, i = 0; y <= y1; ++y) wy[i++] = sc.pod.filter.table[(int) (std::abs(y - py) * sc.lookupFactor)];
    for (int y = y0, yr = 0; y <= y1; ++y, ++yr)
        for (int x = x0, xr = 0; x <= x1; ++x, ++xr) {
            float *c = &blk[((size_t) y * cols + x) * 4];
            c[0] += value.x * wx[xr] * wy[yr]; c[1] += value.y * wx[xr] * wy[yr];
            c[2] += value.z * wx[xr] * wy[yr]; c[3] += 1.0f * wx[xr] * wy[yr];
        }
}

} // namespace

/* =================================================================== C interface (ctypes) ===== */
extern "C" {

void *nori_oracle_create(const nori_gpu_scene *s) {
    Scene *sc = new Scene();
    sc->pod = *s;
    sc->nodes.assign(s->nodes, s->nodes + s->n_nodes);
    sc->indices.assign(s->indices, s->indices + s->n_indices);
    sc->shapeOffset.assign(s->shape_offset, s->shape_offset + s->n_shapes + 1);
    sc->bsdfs.assign(s->bsdfs, s->bsdfs + s->n_bsdfs);
    for (uint32_t i = 0; i < s->n_shapes; ++i) {
        Shape sh; sh.pod = s->shapes[i];
        if (sh.pod.type == NORI_SHAPE_MESH) {
            sh.V.assign(sh.pod.V, sh.pod.V + 3 * (size_t) sh.pod.n_vertices);
            if (sh.pod.N) sh.N.assign(sh.pod.N, sh.pod.N + 3 * (size_t) sh.pod.n_vertices);
            if (sh.pod.UV) sh.UV.assign(sh.pod.UV, sh.pod.UV + 2 * (size_t) sh.pod.n_vertices);
            sh.F.assign(sh.pod.F, sh.pod.F + 3 * (size_t) sh.pod.n_triangles);
            if (sh.pod.area_cdf) sh.cdf.assign(sh.pod.area_cdf, sh.pod.area_cdf + sh.pod.n_triangles + 1);
        }
        sc->shapes.push_back(std::move(sh));
    }
    for (uint32_t i = 0; i < s->n_emitters; ++i) {
        Emitter e; e.pod = s->emitters[i];
        if (e.pod.type == NORI_EMITTER_ENVMAP) {
            size_t R = e.pod.env_rows, C = e.pod.env_cols;
            e.image.assign(e.pod.env_image, e.pod.env_image + R * C * 3);
            e.pdf.assign(e.pod.env_pdf, e.pod.env_pdf + R * C);
            e.cdf.assign(e.pod.env_cdf, e.pod.env_cdf + R * (C + 1));
            e.pm.assign(e.pod.env_pmarginal, e.pod.env_pmarginal + R);
            e.cm.assign(e.pod.env_cmarginal, e.pod.env_cmarginal + R + 1);
        }
        sc->emitters.push_back(std::move(e));
    }
    for (uint32_t i = 0; i < s->n_images; ++i) {
        Image im; im.w = s->images[i].width; im.h = s->images[i].height; im.wrap = s->images[i].wrap;
        im.rgb.assign(s->images[i].rgb, s->images[i].rgb + (size_t) im.w * im.h * 3);
        sc->images.push_back(std::move(im));
    }
    sc->border = (int) std::ceil(s->filter.radius - 0.5f);                    /* block.cpp:57 */
    sc->lookupFactor = NORI_FILTER_RESOLUTION / s->filter.radius;             /* block.cpp:64 */
    return sc;
}
void nori_oracle_destroy(void *h) { delete (Scene *) h; }

void nori_oracle_film_dims(void *h, int32_t *rows, int32_t *cols, int32_t *border) {
    Scene *sc = (Scene *) h;
    *rows = sc->pod.camera.height + 2 * sc->border; *cols = sc->pod.camera.width + 2 * sc->border; *border = sc->border;
}

/* optional outputs p/uv/n/ng (3,2,3,3 floats per ray) may be NULL */
void nori_oracle_trace(void *h, const nori_gpu_ray *rays, uint64_t n, int shadow, nori_gpu_hit *out,
                       float *p, float *uv, float *nsh, float *ngeo) {
    Scene *sc = (Scene *) h;
    parallelFor((int64_t) n, 1024, [&](int64_t i) {
        Ray r(load3(rays[i].o), load3(rays[i].d), rays[i].mint, rays[i].maxt);
        Its its; bool hit = rayIntersect(*sc, r, its, shadow != 0);
        nori_gpu_hit &o = out[i]; memset(&o, 0, sizeof(o));
        o.t = its.t; o.u = its.baryU; o.v = its.baryV; o.shape = hit && !shadow ? (uint32_t) its.shape : 0xffffffffu;
        o.prim = hit && !shadow ? its.prim : 0xffffffffu; o.nodes_visited = its.nodes; o.prims_tested = its.prims;
        if (hit && !shadow) {
            if (p) { p[3 * i] = its.p.x; p[3 * i + 1] = its.p.y; p[3 * i + 2] = its.p.z; }
            if (uv) { uv[2 * i] = its.uv.x; uv[2 * i + 1] = its.uv.y; }
            if (nsh) { nsh[3 * i] = its.sh.n.x; nsh[3 * i + 1] = its.sh.n.y; nsh[3 * i + 2] = its.sh.n.z; }
            if (ngeo) { ngeo[3 * i] = its.geo.n.x; ngeo[3 * i + 1] = its.geo.n.y; ngeo[3 * i + 2] = its.geo.n.z; }
        }
    });
}

/* test-only: closest hits in another visiting order (rayIntersectOrdered); fills t,u,v,shape,prim */
void nori_oracle_trace_ordered(void *h, const nori_gpu_ray *rays, uint64_t n, int order, nori_gpu_hit *out) {
    Scene *sc = (Scene *) h;
    parallelFor((int64_t) n, 1024, [&](int64_t i) {
        Ray r(load3(rays[i].o), load3(rays[i].d), rays[i].mint, rays[i].maxt);
        Its its; bool hit = rayIntersectOrdered(*sc, r, its, order);
        nori_gpu_hit &o = out[i]; memset(&o, 0, sizeof(o));
        o.t = its.t; o.u = its.baryU; o.v = its.baryV; o.shape = hit ? (uint32_t) its.shape : 0xffffffffu;
        o.prim = hit ? its.prim : 0xffffffffu; o.nodes_visited = its.nodes; o.prims_tested = its.prims;
    });
}

void nori_oracle_pcg32(uint64_t initstate, uint64_t initseq, uint64_t n, float *out) {
    Pcg32 r; r.seed(initstate, initseq); for (uint64_t i = 0; i < n; ++i) out[i] = r.nextFloat();
}
void nori_oracle_pcg32_uint(uint64_t initstate, uint64_t initseq, uint64_t n, uint32_t *out) {
    Pcg32 r; r.seed(initstate, initseq); for (uint64_t i = 0; i < n; ++i) out[i] = r.nextUInt();
}

/* RNG mode 0 (the GPU mapping): path (x,y,k) uses pcg32.seed(seed + k, y*W + x).
 * out_rgba[(k*H + y)*W + x] = (r,g,b,valid) */
void nori_oracle_render_samples(void *h, uint32_t spp_begin, uint32_t spp_count, uint64_t seed, float *out) {
    Scene *sc = (Scene *) h; const int W = sc->pod.camera.width, H = sc->pod.camera.height;
    parallelFor((int64_t) spp_count * H, 4, [&](int64_t ky) {
            int64_t k = ky / H; int y = (int) (ky % H);
            for (int x = 0; x < W; ++x) {
                Pcg32 rng; rng.seed(seed + spp_begin + (uint64_t) k, (uint64_t) y * W + x);
                P2 ps; V3 v = cameraSample(*sc, rng, x, y, ps);
                float *o = &out[(((size_t) k * H + y) * W + x) * 4];
                bool ok = validColor(v);
                o[0] = ok ? v.x : 0.f; o[1] = ok ? v.y : 0.f; o[2] = ok ? v.z : 0.f; o[3] = ok ? 1.f : 0.f;
            }
    });
    __atomic_fetch_add(&sc->samples, (uint64_t) spp_count * W * H, __ATOMIC_RELAXED);
}

/* Accumulate into film ((H+2b)*(W+2b)*4 floats).
 * mode 0: per-path streams (as above); mode 1: the reference's mapping -- one pcg32 per 32x32 block,
 * seeded (offset.x, offset.y) when spp_begin == 0 and carried across passes through `block_rng`
 * (2 x uint64 per block, caller-owned, may be NULL if a single call renders everything). */
static void renderVarImpl(void *h, uint32_t spp_begin, uint32_t spp_count, uint64_t seed, int mode,
                          float *film, uint64_t *block_rng, float *vsum, float *vsum2, const float *given) {
    Scene *sc = (Scene *) h; const int W = sc->pod.camera.width, H = sc->pod.camera.height, b = sc->border;
    const int fcols = W + 2 * b, BS = NORI_BLOCK_SIZE;
    const int nbx = (W + BS - 1) / BS, nby = (H + BS - 1) / BS;
    std::vector<Pcg32> rngs((size_t) nbx * nby);
    for (int by = 0; by < nby; ++by) for (int bx = 0; bx < nbx; ++bx) {
        Pcg32 &r = rngs[(size_t) by * nbx + bx];
        if (block_rng && spp_begin != 0) { r.state = block_rng[2 * ((size_t) by * nbx + bx)]; r.inc = block_rng[2 * ((size_t) by * nbx + bx) + 1]; }
        else r.seed((uint64_t) bx * BS, (uint64_t) by * BS);               /* independent.cpp:48-53 */
    }
    for (uint32_t k = 0; k < spp_count; ++k) {
        std::vector<std::vector<float>> blocks((size_t) nbx * nby);
        parallelFor(nbx * nby, 1, [&](int64_t bi) {
            int bx = bi % nbx, by = bi / nbx, ox = bx * BS, oy = by * BS;
            int bw = std::min(BS, W - ox), bh = std::min(BS, H - oy);
            /* the reference's per-thread block is always (32+2b)^2 (render.cpp:203) */
            std::vector<float> &blk = blocks[bi]; blk.assign((size_t) (BS + 2 * b) * (BS + 2 * b) * 4, 0.f);
            for (int y = 0; y < bh; ++y) for (int x = 0; x < bw; ++x) {
                Pcg32 path; Pcg32 *rng = &rngs[bi];
                if (mode == 0) { path.seed(seed + spp_begin + k, (uint64_t) (y + oy) * W + (x + ox)); rng = &path; }
                P2 ps; V3 v;
                if (given) {                                    /* radiance supplied by the caller: only the film position is drawn */
                    P2 a = rng->next2D(); ps.x = (float) (x + ox) + a.x; ps.y = (float) (y + oy) + a.y;
                    const float *g = &given[(((size_t) k * H + (y + oy)) * W + (x + ox)) * 4];
                    v = g[3] != 0.f ? V3(g[0], g[1], g[2]) : V3(-1.f);              /* dropped samples stay dropped (block.cpp:94-98) */
                } else v = cameraSample(*sc, *rng, x + ox, y + oy, ps);
                if (!validColor(v)) __atomic_fetch_add(&sc->invalid, 1, __ATOMIC_RELAXED);
                blockPut(*sc, blk.data(), ox, oy, BS, BS, ps, v);
            }
        });
        /* ImageBlock::put(block), block.cpp:124-133, in block-id order (deterministic) */
        for (int bi = 0; bi < nbx * nby; ++bi) {
            int bx = bi % nbx, by = bi / nbx, ox = bx * BS, oy = by * BS;
            int bw = std::min(BS, W - ox), bh = std::min(BS, H - oy), bc = BS + 2 * b;
            for (int y = 0; y < bh + 2 * b; ++y) for (int x = 0; x < bw + 2 * b; ++x)
                for (int c = 0; c < 4; ++c)
                    film[((size_t) (y + oy) * fcols + (x + ox)) * 4 + c] += blocks[bi][((size_t) y * bc + x) * 4 + c];
        }
        /* per-pass running mean and its square (render.cpp:238-247); vsum / vsum2 are H*W*3, caller-zeroed */
        if (vsum && vsum2)
            for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
                const float *c = &film[((size_t) (y + b) * fcols + (x + b)) * 4];
                for (int ch = 0; ch < 3; ++ch) {
                    float m = c[3] != 0 ? c[ch] / c[3] : 0.f;
                    vsum[((size_t) y * W + x) * 3 + ch] += m; vsum2[((size_t) y * W + x) * 3 + ch] += m * m;
                }
            }
    }
    if (block_rng) for (size_t i = 0; i < rngs.size(); ++i) { block_rng[2 * i] = rngs[i].state; block_rng[2 * i + 1] = rngs[i].inc; }
    __atomic_fetch_add(&sc->samples, (uint64_t) spp_count * W * H, __ATOMIC_RELAXED);
}

/* The first n iterations of renderBlock (render.cpp:96-126) for block (0,0) with the reference's sampler
 * mapping; out rows = (pixelSample.x, pixelSample.y, r, g, b).  Compared with nori_export --seq. */
void nori_oracle_block_sequence(void *h, uint64_t n, float *out) {
    Scene *sc = (Scene *) h; const int W = sc->pod.camera.width, H = sc->pod.camera.height;
    int bw = std::min(NORI_BLOCK_SIZE, W), bh = std::min(NORI_BLOCK_SIZE, H);
    Pcg32 rng; rng.seed(0, 0);
    uint64_t done = 0;
    while (done < n)
        for (int y = 0; y < bh && done < n; ++y)
            for (int x = 0; x < bw && done < n; ++x, ++done) {
                P2 ps; V3 v = cameraSample(*sc, rng, x, y, ps);
                float *o = &out[5 * done]; o[0] = ps.x; o[1] = ps.y; o[2] = v.x; o[3] = v.y; o[4] = v.z;
            }
}

/* Per-function probes, same row layout as nori_export --probe (see oracle/ref_tools/nori_export.cpp). */
void nori_oracle_bsdf_probe(void *h, uint32_t bsdf, uint64_t n, const float *in, float *out) {
    Scene *sc = (Scene *) h; const nori_gpu_bsdf &b = sc->bsdfs[bsdf];
    for (uint64_t i = 0; i < n; ++i) {
        const float *q = &in[10 * i]; float *o = &out[12 * i];
        BRec e; e.scene = sc; e.wi = load3(q); e.wo = load3(q + 3); e.measure = ESolidAngle; e.uv.x = q[6]; e.uv.y = q[7];
        V3 ev = bsdfEval(b, e); float pdf = bsdfPdf(b, e);
        BRec r; r.scene = sc; r.wi = load3(q); r.uv = e.uv; P2 s; s.x = q[8]; s.y = q[9];
        V3 w = bsdfSample(b, r, s); float pdf2 = bsdfPdf(b, r);
        o[0] = ev.x; o[1] = ev.y; o[2] = ev.z; o[3] = pdf; o[4] = w.x; o[5] = w.y; o[6] = w.z;
        o[7] = r.wo.x; o[8] = r.wo.y; o[9] = r.wo.z; o[10] = (float) r.measure; o[11] = pdf2;
    }
}
void nori_oracle_emitter_probe(void *h, uint32_t emitter, uint64_t n, const float *in, float *out) {
    Scene *sc = (Scene *) h; const Emitter &em = sc->emitters[emitter];
    for (uint64_t i = 0; i < n; ++i) {
        const float *q = &in[5 * i]; float *o = &out[15 * i];
        ERec e; e.ref = load3(q); P2 s; s.x = q[3]; s.y = q[4];
        V3 Li = emitterSample(*sc, em, e, s); float pdf = emitterPdf(*sc, em, e); V3 ev = emitterEval(*sc, em, e);
        o[0] = Li.x; o[1] = Li.y; o[2] = Li.z; o[3] = e.wi.x; o[4] = e.wi.y; o[5] = e.wi.z; o[6] = pdf;
        o[7] = e.shadowRay.mint; o[8] = e.shadowRay.maxt; o[9] = e.p.x; o[10] = e.p.y; o[11] = e.p.z;
        o[12] = ev.x; o[13] = ev.y; o[14] = ev.z;
    }
}

void nori_oracle_render_var(void *h, uint32_t spp_begin, uint32_t spp_count, uint64_t seed, int mode,
                            float *film, uint64_t *block_rng, float *vsum, float *vsum2) {
    renderVarImpl(h, spp_begin, spp_count, seed, mode, film, block_rng, vsum, vsum2, nullptr);
}
/* ImageBlock::put / put(block) / the variance statistic applied to radiance values supplied by the caller
 * (samples[(k*H + y)*W + x] = (r,g,b,valid), the layout of nori_oracle_render_samples) at the film positions of the
 * per-path streams (RNG mode 0): the film semantics of block.cpp:93-133 and render.cpp:238-247 on their own, whatever
 * produced the radiance.  vsum / vsum2 may be NULL. */
void nori_oracle_splat(void *h, uint32_t spp_begin, uint32_t spp_count, uint64_t seed, const float *samples,
                       float *film, float *vsum, float *vsum2) {
    renderVarImpl(h, spp_begin, spp_count, seed, 0, film, nullptr, vsum, vsum2, samples);
}
void nori_oracle_render(void *h, uint32_t spp_begin, uint32_t spp_count, uint64_t seed, int mode,
                        float *film, uint64_t *block_rng) {
    nori_oracle_render_var(h, spp_begin, spp_count, seed, mode, film, block_rng, nullptr, nullptr);
}

/* ImageBlock::toBitmap, block.cpp:76-82 + color.h:84-89 */
void nori_oracle_resolve(void *h, const float *film, float *rgb) {
    Scene *sc = (Scene *) h; const int W = sc->pod.camera.width, H = sc->pod.camera.height, b = sc->border, fcols = W + 2 * b;
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
        const float *c = &film[((size_t) (y + b) * fcols + (x + b)) * 4];
        for (int k = 0; k < 3; ++k) rgb[((size_t) y * W + x) * 3 + k] = c[3] != 0 ? c[k] / c[3] : 0.f;
    }
}

void nori_oracle_stats(void *h, nori_gpu_stats *out) {
    Scene *sc = (Scene *) h; memset(out, 0, sizeof(*out));
    out->samples = sc->samples; out->rays = sc->rays; out->shadow_rays = sc->shadowRays;
    out->nodes_visited = sc->nodesVisited; out->prims_tested = sc->primsTested; out->invalid_samples = sc->invalid;
}
void nori_oracle_reset_stats(void *h) {
    Scene *sc = (Scene *) h; sc->samples = sc->rays = sc->shadowRays = sc->nodesVisited = sc->primsTested = sc->invalid = 0;
}

} /* extern "C" */
