// integrators.cuh -- the control flow of the reference's Integrator::Li implementations, split so
// that the same per-vertex code serves the wavefront kernels (path_mis / path_mats: one call per
// queue entry) and the single-thread-per-sample kernel used for the short integrators.
#pragma once
#include "shading.cuh"

#define NORI_Q_MISS NORI_BSDF_COUNT          // volumetric only: rays that left the scene may still scatter in the medium
#define NORI_NQ (NORI_BSDF_COUNT + 1)
#define NORI_NEQ (NORI_BSDF_COUNT * 4)     // (bsdf type, emitter type) keys of the emitter-sorted queues

enum { PF_ALIVE = 1u, PF_SHADOW = 2u, PF_TERMINATE = 4u, PF_FIRST = 8u, PF_DISCRETE = 16u,
       PF_CH_SHIFT = 8u, PF_CH_MASK = 3u << 8 };     // colour channel of the current camera path (chromatic aberration)

struct PathState {
    V3 o, d;            // current ray (origin is the previous vertex: path_mis.cpp:83 `origin`)
    V3 thr, rad;        // `attenuation`, `color`
    float pdf_mat;      // bsdf->pdf of the last sampled direction (path_mis.cpp:81)
    uint32_t flags;
    Pcg32 rng;
};
// An NEE contribution of exactly (0,0,0) -- a discrete BSDF (mirror / dielectric: eval = 0), a light sample below the
// surface's horizon, a back-facing emitter -- cannot change the radiance whatever its shadow ray hits (rad + 0 = rad), so
// the kernels do not trace that ray unless the traversal counters are on (then every query the reference issues is
// traced, and the counters equal the reference's).  NaN compares unequal: such a contribution is traced and added.
__device__ __forceinline__ bool nullContribution(float x, float y, float z) { return x == 0.f && y == 0.f && z == 0.f; }
struct VertexOut {
    Ray shadow; V3 contrib;       // NEE: added to rad iff the shadow ray is unoccluded
    Ray next;                     // extension ray when the path survives
};

template <int BSDF> __device__ __forceinline__ V3 evalT(const nori_gpu_bsdf &b, const BRec &r) {
    if constexpr (BSDF < 0) return bsdfEvalDyn(b, r); else return bsdfEval<BSDF>(b, r);
}
template <int BSDF> __device__ __forceinline__ float pdfT(const nori_gpu_bsdf &b, const BRec &r) {
    if constexpr (BSDF < 0) return bsdfPdfDyn(b, r); else return bsdfPdf<BSDF>(b, r);
}
template <int BSDF> __device__ __forceinline__ V3 sampleT(const nori_gpu_bsdf &b, BRec &r, P2 s) {
    if constexpr (BSDF < 0) return bsdfSampleDyn(b, r, s); else return bsdfSample<BSDF>(b, r, s);
}

// One iteration of the while(true) body of PathMisIntegrator::Li (path_mis.cpp:32-97; MIS=true) or
// PathMatsIntegrator::Li (path_mats.cpp:23-55; MIS=false) at a surface hit.  On return
// st.flags has PF_TERMINATE (Russian roulette ended the path) or PF_ALIVE (out.next is the new
// ray); PF_SHADOW is set when out.shadow / out.contrib are valid.
template <int BSDF, bool MIS, bool AO = false>
__device__ __forceinline__ void pathVertex(const DScene &sc, const Hit &hit, PathState &st, VertexOut &out) {
    Its its; hitInfo(sc, st.o, st.d, hit, its);
    const DShape &shp = sc.shapes[its.shape];
    NORI_CHECK(shp.bsdf >= 0 && shp.emitter < (int32_t) sc.n_emitters);
    const nori_gpu_bsdf &bsdf = sc.bsdfs[shp.bsdf];
    const uint32_t inFlags = st.flags;
    uint32_t flags = 0;

    if (shp.emitter >= 0) {                                        // path_mis.cpp:35-39 / :87-97, path_mats.cpp:32-36
        const nori_gpu_emitter &em = sc.emitters[shp.emitter].pod;
        ERec e = makeERec(st.o, its.p, its.sh.n);
        float w_mats = 1.0f;
        if (MIS && !(inFlags & (PF_FIRST | PF_DISCRETE))) {
            float pdf_em = emitterPdf<AO>(sc, em, e);
            w_mats = st.pdf_mat + pdf_em > 0.f ? fdiv(st.pdf_mat, st.pdf_mat + pdf_em) : st.pdf_mat;
        }
        V3 Le = emitterEval<AO>(sc, em, e);
        st.rad = st.rad + (MIS ? st.thr * w_mats * Le : st.thr * Le);
    }

    const V3 wiLocal = toLocal(its.sh, -st.d);
    if (MIS) {                                                     // path_mis.cpp:42-61
        const nori_gpu_emitter &light = sc.emitters[randomEmitter(sc, st.rng.next1D())].pod;
        ERec e = makeERec(its.p);
        V3 Li = emitterSample<AO>(sc, light, e, st.rng.next2D()) * (float) sc.n_emitters;
        float pdf_em = emitterPdf<AO>(sc, light, e);
        V3 woLocal = toLocal(its.sh, e.wi);
        float theta = fmaxf(0.0f, woLocal.z);
        BRec b = mkBRec(sc, bsdf, wiLocal, M_SOLID_ANGLE, its.uv); b.wo = woLocal;
        V3 f = evalT<BSDF>(bsdf, b);
        float pdf_mat = pdfT<BSDF>(bsdf, b);
        float w_ems = (pdf_mat + pdf_em) > 0.0f ? fdiv(pdf_em, pdf_mat + pdf_em) : pdf_em;
        out.shadow = e.shadow;
        out.contrib = st.thr * w_ems * f * theta * Li;
        flags |= PF_SHADOW;
    }

    float p = fminf(st.thr.x, 0.99f);                              // path_mis.cpp:64-69: roulette on the RED channel
    if (st.rng.next1D() > p) { st.flags = flags | PF_TERMINATE; return; }
    st.thr = st.thr / p;

    BRec b = mkBRec(sc, bsdf, wiLocal, M_UNKNOWN, its.uv);          // path_mis.cpp:72-81
    V3 w = sampleT<BSDF>(bsdf, b, st.rng.next2D());
    st.thr = st.thr * w;
    out.next = mkray(its.p, toWorld(its.sh, b.wo));
    if (MIS) st.pdf_mat = pdfT<BSDF>(bsdf, b);
    st.o = out.next.o; st.d = out.next.d;
    st.flags = flags | PF_ALIVE | (b.measure == M_DISCRETE ? PF_DISCRETE : 0u);
}

// ---------------------------------------------------------------------------------------------
// single-thread-per-sample integrators (normals, av, direct*, volumetric; also path_* as a
// cross-check of the wavefront scheduler).  Returns the radiance of one camera ray.
// ---------------------------------------------------------------------------------------------
struct RayStats { uint32_t rays, shadow; TraceCounters cnt; };

template <bool COUNT>
__device__ __forceinline__ bool closestHit(const DScene &sc, const Ray &r, Hit &h, RayStats &rs) {
    ++rs.rays; return traverse<false, COUNT, true>(sc, r.o, r.d, r.mint, r.maxt, h, rs.cnt, sc.ordered != 0);
}
template <bool COUNT>
__device__ __forceinline__ bool anyHit(const DScene &sc, const Ray &r, RayStats &rs) {
    Hit h; ++rs.rays; ++rs.shadow; return traverse<true, COUNT, true>(sc, r.o, r.d, r.mint, r.maxt, h, rs.cnt, sc.ordered != 0);
}

template <bool COUNT, bool MIS>
__device__ V3 liPath(const DScene &sc, Pcg32 &rng, Ray ray, RayStats &rs) {
    PathState st; st.o = ray.o; st.d = ray.d; st.thr = mk(1.f); st.rad = mk(0.f); st.pdf_mat = 0.f; st.flags = PF_FIRST; st.rng = rng;
    Ray cur = ray;
    while (true) {
        Hit h;
        if (!closestHit<COUNT>(sc, cur, h, rs)) break;
        VertexOut out;
        pathVertex<-1, MIS>(sc, h, st, out);
        if ((st.flags & PF_SHADOW) && (COUNT || !nullContribution(out.contrib.x, out.contrib.y, out.contrib.z))
            && !anyHit<COUNT>(sc, out.shadow, rs)) st.rad = st.rad + out.contrib;
        if (st.flags & PF_TERMINATE) break;
        cur = out.next;
    }
    rng = st.rng;
    return st.rad;
}

// direct.cpp:18-51, direct_ems.cpp:17-54, direct_mats.cpp:17-46, direct_mis.cpp:17-87
template <bool COUNT>
__device__ V3 liDirect(const DScene &sc, Pcg32 &rng, const Ray &ray, int kind, RayStats &rs) {
    Hit h;
    if (!closestHit<COUNT>(sc, ray, h, rs)) return mk(0.f);
    Its its; hitInfo(sc, ray.o, ray.d, h, its);
    const DShape &shp = sc.shapes[its.shape];
    const nori_gpu_bsdf &bsdf = sc.bsdfs[shp.bsdf];
    V3 color = mk(0.f);
    if (kind != NORI_INTEGRATOR_DIRECT && shp.emitter >= 0) {
        ERec e = makeERec(ray.o, its.p, its.sh.n);
        color = color + emitterEval(sc, sc.emitters[shp.emitter].pod, e);
    }
    const V3 dLocal = toLocal(its.sh, -ray.d);
    if (kind == NORI_INTEGRATOR_DIRECT || kind == NORI_INTEGRATOR_DIRECT_EMS || kind == NORI_INTEGRATOR_DIRECT_MIS) {
        for (uint32_t li = 0; li < sc.n_emitters; ++li) {
            const nori_gpu_emitter &light = sc.emitters[li].pod;
            ERec e = makeERec(its.p);
            P2 s; s.x = 0.f; s.y = 0.f;                             // direct.cpp:27 passes a zero-filled Vector2f
            if (kind != NORI_INTEGRATOR_DIRECT) s = rng.next2D();
            V3 traced = emitterSample(sc, light, e, s);
            float pdf_em = kind == NORI_INTEGRATOR_DIRECT_MIS ? emitterPdf(sc, light, e) : 0.f;
            if (!anyHit<COUNT>(sc, e.shadow, rs)) {
                V3 wi = toLocal(its.sh, e.wi);
                BRec b = mkBRec(sc, bsdf, wi, M_SOLID_ANGLE, its.uv);
                if (kind == NORI_INTEGRATOR_DIRECT) { b.wi = wi; b.wo = dLocal; } else { b.wi = dLocal; b.wo = wi; }
                V3 f = bsdfEvalDyn(bsdf, b);
                if (kind == NORI_INTEGRATOR_DIRECT_MIS) {
                    float pdf_mat = bsdfPdfDyn(bsdf, b);
                    float w_em = pdf_mat + pdf_em > 0.f ? fdiv(pdf_em, pdf_mat + pdf_em) : pdf_em;
                    color = color + w_em * f * traced * wi.z;
                } else color = color + f * wi.z * traced;
            }
        }
    }
    if (kind == NORI_INTEGRATOR_DIRECT_MATS || kind == NORI_INTEGRATOR_DIRECT_MIS) {
        BRec b = mkBRec(sc, bsdf, dLocal, M_UNKNOWN, its.uv);
        V3 w = bsdfSampleDyn(bsdf, b, rng.next2D());
        float pdf_mat = kind == NORI_INTEGRATOR_DIRECT_MIS ? bsdfPdfDyn(bsdf, b) : 0.f;
        Ray nr = mkray(its.p, toWorld(its.sh, b.wo));
        Hit h2;
        if (closestHit<COUNT>(sc, nr, h2, rs)) {
            Its its2; hitInfo(sc, nr.o, nr.d, h2, its2);
            const DShape &shp2 = sc.shapes[its2.shape];
            if (shp2.emitter >= 0) {
                const nori_gpu_emitter &em = sc.emitters[shp2.emitter].pod;
                ERec e = makeERec(its.p, its2.p, its2.sh.n);
                V3 Le = emitterEval(sc, em, e);
                if (kind == NORI_INTEGRATOR_DIRECT_MIS) {
                    float pdf_em = emitterPdf(sc, em, e);
                    float w_mat = pdf_mat + pdf_em > 0.f ? fdiv(pdf_mat, pdf_mat + pdf_em) : 0.0f;
                    color = color + w_mat * w * Le;
                } else color = color + w * Le;
            }
        }
    }
    return color;
}

// normals.cpp:15-23
template <bool COUNT>
__device__ V3 liNormals(const DScene &sc, const Ray &ray, RayStats &rs) {
    Hit h; if (!closestHit<COUNT>(sc, ray, h, rs)) return mk(0.f);
    Its its; hitInfo(sc, ray.o, ray.d, h, its);
    return mk(fabsf(its.sh.n.x), fabsf(its.sh.n.y), fabsf(its.sh.n.z));
}

// averagevisibility.cpp:16-25 + warp.cpp:25-42
template <bool COUNT>
__device__ V3 liAv(const DScene &sc, Pcg32 &rng, const Ray &ray, RayStats &rs) {
    Hit h; if (!closestHit<COUNT>(sc, ray, h, rs)) return mk(1.f);
    Its its; hitInfo(sc, ray.o, ray.d, h, its);
    V3 v;
    do { v.x = 1.f - 2.f * rng.next1D(); v.y = 1.f - 2.f * rng.next1D(); v.z = 1.f - 2.f * rng.next1D(); } while (sqnorm(v) > 1.f);
    if (dot(v, its.sh.n) < 0.f) v = -v;
    v = v / norm(v);
    Ray nr = mkray(its.p, v, NORI_EPS, sc.av_length);
    return anyHit<COUNT>(sc, nr, rs) ? mk(0.f) : mk(1.f);
}

// ---- homogeneous medium (medium.cpp:22-94) --------------------------------------------------
__device__ __forceinline__ bool boundsHit(const nori_gpu_medium &m, V3 o, V3 d, float &nearT, float &farT) {   // bbox.h:366-392
    nearT = __int_as_float(0xff800000); farT = __int_as_float(0x7f800000);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        float origin = comp(o, i), dd = comp(d, i), minVal = m.bounds_min[i], maxVal = m.bounds_max[i];
        if (dd == 0) { if (origin < minVal || origin > maxVal) return false; }
        else {
            float rcp = frcp(dd);
            float t1 = (minVal - origin) * rcp, t2 = (maxVal - origin) * rcp;
            if (t1 > t2) { float t = t1; t1 = t2; t2 = t; }
            nearT = std_max(t1, nearT); farT = std_min(t2, farT);
            if (!(nearT <= farT)) return false;
        }
    }
    return true;
}
__device__ __forceinline__ bool boundsContain(const nori_gpu_medium &m, V3 p) {          // bbox.h:115-123
    return p.x >= m.bounds_min[0] && p.x <= m.bounds_max[0] && p.y >= m.bounds_min[1] && p.y <= m.bounds_max[1]
        && p.z >= m.bounds_min[2] && p.z <= m.bounds_max[2];
}
static __device__ V3 mediumTr(const nori_gpu_medium &m, V3 src, V3 dst) {                       // medium.cpp:22-57
    float nearT, farT;
    V3 d = normalized(dst - src);
    if (!boundsHit(m, src, d, nearT, farT)) return mk(1.0f);
    V3 sp = boundsContain(m, src) ? src : src + normalized(d) * nearT;
    V3 ep = boundsContain(m, dst) ? dst : src + normalized(d) * farT;
    float len = norm(ep - sp);
    V3 ext = arr3(m.sigma_a) + arr3(m.sigma_s);
    return mk(expf(-ext.x * len), expf(-ext.y * len), expf(-ext.z * len));
}
static __device__ V3 mediumSample(const nori_gpu_medium &m, const Ray &ray, Pcg32 &rng, float tMax, bool &hitObject, V3 &p) {   // medium.cpp:59-90
    float nearT, farT;
    if (!boundsHit(m, ray.o, ray.d, nearT, farT)) { hitObject = true; return mk(1.0f); }
    V3 sp = boundsContain(m, ray.o) ? ray.o : ray.o + normalized(ray.d) * nearT;
    V3 ext = arr3(m.sigma_a) + arr3(m.sigma_s);
    float invTr = fdiv(-1.0f * logf(1 - rng.next1D()), fmaxf(ext.x, fmaxf(ext.y, ext.z)));     // medium.cpp:92-94
    float distance = norm(sp - ray.o) + invTr;
    V3 albedo = mk(fdiv(m.sigma_s[0], ext.x), fdiv(m.sigma_s[1], ext.y), fdiv(m.sigma_s[2], ext.z));
    if (distance >= tMax) hitObject = true; else { p = ray.o + distance * ray.d; hitObject = false; }
    return albedo;
}


// One iteration of the while(true) body of VolumetricIntegrator::Li (volumetric.cpp:31-151) for the
// wavefront: the closest-hit query of the current ray has been answered by k_extend (`hit`; BSDF ==
// NORI_Q_MISS when it left the scene).  Free-flight sampling decides between a medium vertex
// (volumetric.cpp:47-87) and a surface vertex (:89-145); both NEE queries are traced here.  `w_mats`
// of the reference (:76-82, :133-142) is a function of the pdf of the direction that produced the
// current ray, of whether it came from a discrete BSDF, and of the emitter that was hit -- it is
// recomputed here from st.pdf_mat / PF_DISCRETE / PF_FIRST with the same expression.
template <bool COUNT>
__device__ __forceinline__ void volVertex(const DScene &sc, const Hit &hit, PathState &st, Ray &next,
                                          uint32_t &nClosest, uint32_t &nShadow, TraceCounters &cnt) {
    // ONE copy of the emitter sampling, of the NEE traversal and of the transmittance serves both vertex
    // kinds, and the BSDF is reached through the (warp-uniform) type switch: a copy of this body per BSDF type
    // and per vertex kind made the volumetric k_shade instruction-fetch bound (ncu: 28.7 warps stalled on
    // `no instruction` per issue, 16 % issue-active).  Random numbers are drawn in the reference's order.
    const nori_gpu_medium &med = sc.medium;
    const bool intersection = hit.leafpos != NORI_NO_HIT;
    const uint32_t inFlags = st.flags;
    Its its;
    float tmax = hit.t;
    if (intersection) { hitInfo(sc, st.o, st.d, hit, its); tmax = norm(its.p - st.o); }
    bool hitObject; V3 mp = mk(0.f);
    const Ray cur = mkray(st.o, st.d);
    const V3 sampled = mediumSample(med, cur, st.rng, tmax, hitObject, mp);
    if (hitObject && !intersection) { st.flags = PF_TERMINATE; return; }   // volumetric.cpp:147-151
    const bool medium = !hitObject;                                 // medium vertex (:47-87) or surface vertex (:89-145)
    V3 wo = mk(0.f);
    const DShape *shp = nullptr; const nori_gpu_bsdf *bsdf = nullptr;
    if (medium) wo = squareToUniformSphere(st.rng.next2D());        // phasefunction.cpp:13-16, pdf = 1/(4 pi)
    else {
        shp = &sc.shapes[its.shape]; bsdf = &sc.bsdfs[shp->bsdf];
        if (shp->emitter >= 0) {
            const nori_gpu_emitter &em = sc.emitters[shp->emitter].pod;
            ERec l = makeERec(st.o, its.p, its.sh.n);
            float w_mats = 1.0f;
            if (!(inFlags & (PF_FIRST | PF_DISCRETE))) {
                float pdf_em = emitterPdf(sc, em, l);
                w_mats = st.pdf_mat + pdf_em > 0.f ? fdiv(st.pdf_mat, st.pdf_mat + pdf_em) : st.pdf_mat;
            }
            st.rad = st.rad + st.thr * w_mats * emitterEval(sc, em, l) * mediumTr(med, its.p, l.p);
        }
    }
    const V3 ref = medium ? mp : its.p;
    const nori_gpu_emitter &light = sc.emitters[randomEmitter(sc, st.rng.next1D())].pod;
    ERec e = makeERec(ref);
    V3 Li = emitterSample(sc, light, e, st.rng.next2D()) * (float) sc.n_emitters;
    if (medium) { st.thr = st.thr * sampled; ++nClosest; } else ++nShadow;
    // the medium vertex's query is a closest-hit one in the reference (volumetric.cpp:63); only its boolean is
    // used, which an any-hit query answers identically -- the closest-hit form is kept when the counters are on
    Hit tmp; bool occluded;
    if (COUNT && medium) occluded = traverse<false, COUNT>(sc, e.shadow.o, e.shadow.d, e.shadow.mint, e.shadow.maxt, tmp, cnt);
    else occluded = traverse<true, COUNT>(sc, e.shadow.o, e.shadow.d, e.shadow.mint, e.shadow.maxt, tmp, cnt);
    V3 wiLocal = mk(0.f);
    if (!medium) wiLocal = toLocal(its.sh, -st.d);
    if (!occluded) {
        const V3 Tr = mediumTr(med, ref, e.p);
        if (medium) st.rad = st.rad + st.thr * Tr * Li * NORI_INV_FOURPI;
        else {
            float pdf_em = emitterPdf(sc, light, e);
            V3 woLocal = toLocal(its.sh, e.wi);
            float theta = fmaxf(0.0f, woLocal.z);
            P2 uv0; uv0.x = 0.f; uv0.y = 0.f;                       // bRec.uv is not set (volumetric.cpp:103): Point2f() = 0
            BRec b = mkBRec(sc, *bsdf, wiLocal, M_SOLID_ANGLE, uv0); b.wo = woLocal;
            V3 f = bsdfEvalDyn(*bsdf, b);
            float pdf_mat = bsdfPdfDyn(*bsdf, b);
            float w_ems = (pdf_mat + pdf_em) > 0.0f ? fdiv(pdf_em, pdf_mat + pdf_em) : pdf_em;
            st.rad = st.rad + st.thr * w_ems * f * theta * Li * Tr;
        }
    }
    float p = fminf(st.thr.x, 0.80f);
    if (st.rng.next1D() > p) { st.flags = PF_TERMINATE; return; }
    st.thr = st.thr / p;
    if (medium) {
        next = mkray(mp, normalized(wo));
        st.pdf_mat = NORI_INV_FOURPI; st.flags = PF_ALIVE;
    } else {
        P2 uv1; uv1.x = 0.f; uv1.y = 0.f;
        BRec b = mkBRec(sc, *bsdf, wiLocal, M_UNKNOWN, uv1);
        V3 w = bsdfSampleDyn(*bsdf, b, st.rng.next2D());
        st.thr = st.thr * w;
        st.pdf_mat = bsdfPdfDyn(*bsdf, b);
        next = mkray(its.p, toWorld(its.sh, b.wo));
        st.flags = PF_ALIVE | (b.measure == M_DISCRETE ? PF_DISCRETE : 0u);
    }
}

// volumetric.cpp:18-156
template <bool COUNT>
__device__ V3 liVolumetric(const DScene &sc, Pcg32 &rng, Ray cur, RayStats &rs) {
    const nori_gpu_medium &med = sc.medium;
    V3 color = mk(0.f), att = mk(1.f); float w_mats = 1.0f;
    Hit h; Its its;
    bool intersection = closestHit<COUNT>(sc, cur, h, rs);
    if (intersection) hitInfo(sc, cur.o, cur.d, h, its);
    while (true) {
        float tmax = intersection ? norm(its.p - cur.o) : h.t;
        bool hitObject; V3 mp = mk(0.f);
        V3 sampled = mediumSample(med, cur, rng, tmax, hitObject, mp);
        if (!hitObject) {
            V3 wo = squareToUniformSphere(rng.next2D()); float pdf_mat = NORI_INV_FOURPI;   // phasefunction.cpp:13-16
            const nori_gpu_emitter &light = sc.emitters[randomEmitter(sc, rng.next1D())].pod;
            ERec e = makeERec(mp);
            V3 Li = emitterSample(sc, light, e, rng.next2D()) * (float) sc.n_emitters;
            att = att * sampled;
            Hit tmp;
            if (!closestHit<COUNT>(sc, e.shadow, tmp, rs))            // a closest-hit query in the reference (volumetric.cpp:63)
                color = color + att * mediumTr(med, mp, e.p) * Li * pdf_mat;
            float p = fminf(att.x, 0.80f);
            if (rng.next1D() > p) return color;
            att = att / p;
            cur = mkray(mp, normalized(wo));
            intersection = closestHit<COUNT>(sc, cur, h, rs);
            if (intersection) {
                hitInfo(sc, cur.o, cur.d, h, its);
                const DShape &shp = sc.shapes[its.shape];
                if (shp.emitter >= 0) {
                    ERec l = makeERec(cur.o, its.p, its.sh.n);
                    float pdf_em = emitterPdf(sc, sc.emitters[shp.emitter].pod, l);
                    w_mats = pdf_mat + pdf_em > 0.f ? fdiv(pdf_mat, pdf_mat + pdf_em) : pdf_mat;
                }
            }
        } else if (intersection) {
            const DShape &shp = sc.shapes[its.shape];
            const nori_gpu_bsdf &bsdf = sc.bsdfs[shp.bsdf];
            if (shp.emitter >= 0) {
                ERec e = makeERec(cur.o, its.p, its.sh.n);
                color = color + att * w_mats * emitterEval(sc, sc.emitters[shp.emitter].pod, e) * mediumTr(med, its.p, e.p);
            }
            const nori_gpu_emitter &light = sc.emitters[randomEmitter(sc, rng.next1D())].pod;
            ERec e = makeERec(its.p);
            V3 Li = emitterSample(sc, light, e, rng.next2D()) * (float) sc.n_emitters;
            const V3 wiLocal = toLocal(its.sh, -cur.d);
            if (!anyHit<COUNT>(sc, e.shadow, rs)) {
                float pdf_em = emitterPdf(sc, light, e);
                V3 woLocal = toLocal(its.sh, e.wi);
                float theta = fmaxf(0.0f, woLocal.z);
                P2 uv0; uv0.x = 0.f; uv0.y = 0.f;                       // bRec.uv is not set (volumetric.cpp:103): Point2f() = 0
            BRec b = mkBRec(sc, bsdf, wiLocal, M_SOLID_ANGLE, uv0); b.wo = woLocal;
                V3 f = bsdfEvalDyn(bsdf, b);
                float pdf_mat = bsdfPdfDyn(bsdf, b);
                float w_ems = (pdf_mat + pdf_em) > 0.0f ? fdiv(pdf_em, pdf_mat + pdf_em) : pdf_em;
                color = color + att * w_ems * f * theta * Li * mediumTr(med, its.p, e.p);
            }
            float p = fminf(att.x, 0.80f);
            if (rng.next1D() > p) return color;
            att = att / p;
            P2 uv1; uv1.x = 0.f; uv1.y = 0.f;
        BRec b = mkBRec(sc, bsdf, wiLocal, M_UNKNOWN, uv1);
            V3 w = bsdfSampleDyn(bsdf, b, rng.next2D());
            att = att * w;
            float pdf_mat = bsdfPdfDyn(bsdf, b);
            cur = mkray(its.p, toWorld(its.sh, b.wo));
            intersection = closestHit<COUNT>(sc, cur, h, rs);
            if (intersection) {
                hitInfo(sc, cur.o, cur.d, h, its);
                const DShape &shp2 = sc.shapes[its.shape];
                if (shp2.emitter >= 0) {
                    ERec l = makeERec(cur.o, its.p, its.sh.n);
                    float pdf_em = emitterPdf(sc, sc.emitters[shp2.emitter].pod, l);
                    w_mats = pdf_mat + pdf_em > 0.f ? fdiv(pdf_mat, pdf_mat + pdf_em) : pdf_mat;
                }
                if (b.measure == M_DISCRETE) w_mats = 1.0f;
            }
        } else break;
    }
    return color;
}

template <bool COUNT>
__device__ V3 liDispatch(const DScene &sc, Pcg32 &rng, const Ray &ray, RayStats &rs) {
    switch (sc.integrator) {
    case NORI_INTEGRATOR_NORMALS: return liNormals<COUNT>(sc, ray, rs);
    case NORI_INTEGRATOR_PATH_MIS: return liPath<COUNT, true>(sc, rng, ray, rs);
    case NORI_INTEGRATOR_PATH_MATS: return liPath<COUNT, false>(sc, rng, ray, rs);
    case NORI_INTEGRATOR_AV: return liAv<COUNT>(sc, rng, ray, rs);
    case NORI_INTEGRATOR_VOLUMETRIC: return liVolumetric<COUNT>(sc, rng, ray, rs);
    default: return liDirect<COUNT>(sc, rng, ray, sc.integrator, rs);
    }
}
