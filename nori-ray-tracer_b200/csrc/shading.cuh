// shading.cuh -- device restatement of the per-vertex work of the reference's integrators:
// hit information, frames, warps, BSDF sample/eval/pdf, emitter sample/eval/pdf, cameras.
// Each function cites the reference lines whose behaviour it reproduces (quirks included; see
// SURVEY.md appendix A).  None of this is shared with oracle/ -- the oracle is an independent
// scalar C++ restatement used only by the tests.
#pragma once
#include "device_common.cuh"
#include "traverse.cuh"

struct Frame { V3 s, t, n; };
struct Its {                      // shape.h:38-67 (geoFrame is never read by the hot-path integrators)
    V3 p; P2 uv; Frame sh; int shape;
};

// common.cpp:274-283 + frame.h:49-51
__device__ __forceinline__ Frame makeFrame(V3 a) {
    Frame f; f.n = a;
    V3 c;
    if (fabsf(a.x) > fabsf(a.y)) {
        float invLen = frsqrt(fma_(a.x, a.x, a.z * a.z));
        c = mk(a.z * invLen, 0.0f, -a.x * invLen);
    } else {
        float invLen = frsqrt(fma_(a.y, a.y, a.z * a.z));
        c = mk(0.0f, a.z * invLen, -a.y * invLen);
    }
    f.s = cross(c, a); f.t = c;
    return f;
}
__device__ __forceinline__ V3 toLocal(const Frame &f, V3 v) { return mk(dot(v, f.s), dot(v, f.t), dot(v, f.n)); }
__device__ __forceinline__ V3 toWorld(const Frame &f, V3 v) { return madd(madd(f.s * v.x, v.y, f.t), v.z, f.n); }

// ImageTexture::getData / NormalMap::getData (imagetexture.cpp:98-116, normalmap.cpp:98-120): the texel at the
// truncated coordinates, wrapped with C's % (repeat) or clamped.  A negative remainder reads in front of the
// array in the reference (undefined behaviour); here it is wrapped into range.
__device__ __forceinline__ V3 imageTexel(const DImage &im, float xf, float yf) {
    int x, y;
    if (im.wrap == NORI_WRAP_REPEAT) {
        x = (int) xf % im.width; y = (int) yf % im.height;
        if (x < 0) x += im.width;
        if (y < 0) y += im.height;
    } else {
        x = min(max((int) xf, 0), im.width - 1); y = min(max((int) yf, 0), im.height - 1);
    }
    const uint8_t *t = &im.rgb[((size_t) x + (size_t) im.width * y) * 3];
    return mk(fdiv((float) __ldg(t), 255.f), fdiv((float) __ldg(t + 1), 255.f), fdiv((float) __ldg(t + 2), 255.f));
}
// ImageTexture::eval / NormalMap::eval (imagetexture.cpp:118-136, normalmap.cpp:122-139).  In the reference
// `x` is the un-floored float uv.x * width, so dstdx = uv.x * width - x is exactly 0 and the "bilinear" blend
// 1*1*v00 + 1*0*v01 + 0*1*v10 + 0*0*v11 is the texel at the truncated coordinates, bit for bit (the other
// three texels are finite 8-bit values times zero).  NORMAL applies normalmap.cpp:115-119 (2c - 1).
template <bool NORMAL>
__device__ __forceinline__ V3 imageEval(const DImage &im, P2 uv) {
    V3 c = imageTexel(im, uv.x * (float) im.width, uv.y * (float) im.height);
    if (NORMAL) c = mk(2.0f * c.x - 1.0f, 2.0f * c.y - 1.0f, 2.0f * c.z - 1.0f);
    return c;
}

// mesh.cpp:122-170, sphere.cpp:78-93
__device__ __forceinline__ void hitInfo(const DScene &sc, V3 o, V3 d, const Hit &h, Its &its) {
    NORI_CHECK(h.leafpos < sc.n_prims);
    const float4 r0 = __ldg(&sc.prims[3 * h.leafpos]);
    const float4 r1 = __ldg(&sc.prims[3 * h.leafpos + 1]);
    const float4 r2 = __ldg(&sc.prims[3 * h.leafpos + 2]);
    its.shape = (int) __float_as_uint(r1.w);
    NORI_CHECK((uint32_t) its.shape < sc.n_shapes);
    const DShape &m = sc.shapes[its.shape];
    if (__float_as_uint(r2.w) == 0u) {
        const uint32_t prim = __float_as_uint(r0.w);
        NORI_CHECK(prim < m.n_triangles);
        const float b1 = h.u, b2 = h.v, b0 = 1 - (b1 + b2);
        const uint32_t i0 = __ldg(&m.F[3 * prim]), i1 = __ldg(&m.F[3 * prim + 1]), i2 = __ldg(&m.F[3 * prim + 2]);
        const V3 p0 = ld3(&m.V[3 * i0]), p1 = ld3(&m.V[3 * i1]), p2 = ld3(&m.V[3 * i2]);
        its.p = bary(b0, p0, b1, p1, b2, p2);
        its.uv.x = b1; its.uv.y = b2;
        if (m.has_uv) {
            its.uv.x = (b0 * __ldg(&m.UV[2 * i0]) + b1 * __ldg(&m.UV[2 * i1])) + b2 * __ldg(&m.UV[2 * i2]);
            its.uv.y = (b0 * __ldg(&m.UV[2 * i0 + 1]) + b1 * __ldg(&m.UV[2 * i1 + 1])) + b2 * __ldg(&m.UV[2 * i2 + 1]);
        }
        V3 n;
        if (m.has_n) n = normalizedDyn(bary(b0, ld3(&m.N[3 * i0]), b1, ld3(&m.N[3 * i1]), b2, ld3(&m.N[3 * i2])));
        else n = normalized(cross(p1 - p0, p2 - p0));
        its.sh = makeFrame(n);
        if (m.has_n && m.normal_map > 0)                          // mesh.cpp:149-154
            its.sh = makeFrame(toWorld(its.sh, normalized(imageEval<true>(sc.images[m.normal_map - 1], its.uv))));
    } else {
        const V3 c = mk(r0.x, r0.y, r0.z);
        its.p = madd(o, h.t, d);
        const V3 n = normalized(its.p - c);
        its.sh = makeFrame(n);
        its.uv.x = 0.f; its.uv.y = 0.f;
        // its.uv is only ever read by a textured albedo (mkBRec; normal maps exist on meshes only): the spherical
        // coordinates (acos + atan2) are computed for those shapes alone
        if (sc.bsdfs[m.bsdf].albedo_texture != NORI_TEXTURE_CONSTANT) {
            float th = acosf(n.z), ph = atan2f(n.y, n.x);            // common.cpp:264-272
            if (ph < 0) ph += 2 * NORI_PI;
            if (__float_as_uint(r2.w) == 1u) {                          // sphere.cpp:88-91
                its.uv.x = (float) (0.5 + th / (2 * NORI_PI));
                its.uv.y = ph / NORI_PI;
            } else {                                                    // perlinnoise.cpp:71-74
                its.uv.x = (float) (0.5 + (double) (th * 0.15915494309189533577f));
                its.uv.y = ph * NORI_INV_PI;
            }
        }
    }
}

// ------------------------------------------------------------------------------ warps (warp.cpp)
__device__ __forceinline__ V3 sphericalDir(float theta, float phi) {
    float st, ct, sp, cp; sincosf(theta, &st, &ct); sincosf(phi, &sp, &cp);
    return mk(st * cp, st * sp, ct);
}
#if NORI_FAST_SHADING
// The reference's warps compute theta = acos / atan(..) and then sin(theta), cos(theta) again.  Here the direction
// is built from cos(theta) AND sin(theta), each derived directly from the warped sample (never sin from 1 - cos^2,
// which loses everything near the pole), and the azimuth 2 pi y goes through the MUFU sine / cosine after an exact
// shift into [-pi, pi) (absolute error < 4e-7): no acosf / atan, no sincosf range reduction.
__device__ __forceinline__ V3 sphericalDirCS(float ct, float st, float y) {
    float sp, cp; __sincosf(__fmul_rn(2.f * NORI_PI, __fsub_rn(y, 0.5f)), &sp, &cp);   // sin / cos(phi - pi) = -sin / -cos(phi)
    return mk(-__fmul_rn(st, cp), -__fmul_rn(st, sp), ct);
}
__device__ __forceinline__ V3 squareToUniformSphere(P2 s) {                     // cos = 1 - 2 (1 - x), sin^2 = (1 - cos)(1 + cos)
    const float ct = 1 - 2 * (1 - s.x);
    return sphericalDirCS(ct, fsqrt(fmaxf(0.0f, (1.0f - ct) * (1.0f + ct))), s.y);
}
__device__ __forceinline__ V3 squareToCosineHemisphere(P2 s) {                  // cos^2 = 1 - (1 - x), sin^2 = 1 - cos^2
    const float c2 = 1 - (1 - s.x);
    return sphericalDirCS(fsqrt(c2), fsqrt(1.0f - c2), s.y);
}
__device__ __forceinline__ V3 squareToBeckmann(P2 s, float alpha) {            // tan^2 = -alpha^2 log(1 - x); cos = 1 / sqrt(1 + tan^2), sin = tan cos
    const float t2 = -(alpha * alpha) * logf(1 - s.x);
    const float ct = frsqrt(1.0f + t2);
    return sphericalDirCS(ct, fsqrt(t2) * ct, s.y);
}
#else
__device__ __forceinline__ V3 squareToUniformSphere(P2 s) { return sphericalDir(acosf(1 - 2 * (1 - s.x)), 2.f * NORI_PI * s.y); }   // :86-91
__device__ __forceinline__ V3 squareToCosineHemisphere(P2 s) { return sphericalDir(acosf(sqrtf(1 - (1 - s.x))), 2.f * NORI_PI * s.y); }  // :110-115
__device__ __forceinline__ V3 squareToBeckmann(P2 s, float alpha) {                                                      // :122-127
    float theta = (float) atan(sqrt(-((double) alpha * (double) alpha) * (double) logf(1 - s.x)));
    return sphericalDir(theta, 2 * NORI_PI * s.y);
}
#endif
__device__ __forceinline__ V3 squareToUniformTriangle(P2 s) {                                                            // :135-140
    float su1 = fsqrt(s.x); float u = 1.f - su1, v = s.y * su1;
    return mk(u, v, 1.f - u - v);
}
__device__ __forceinline__ P2 squareToConcentricDisk(P2 s) {                                                             // :143-162
    float ox = 2.f * s.x - 1.f, oy = 2.f * s.y - 1.f; P2 r; r.x = 0.f; r.y = 0.f;
    if (ox == 0.f && oy == 0.f) return r;
    float theta, rad;
    if (fabsf(ox) > fabsf(oy)) { rad = ox; theta = NORI_PI * 0.25f * (oy / ox); }
    else { rad = oy; theta = NORI_PI * 0.5f - NORI_PI * 0.25f * (ox / oy); }
    r.x = rad * cosf(theta); r.y = rad * sinf(theta); return r;
}
#if NORI_FAST_SHADING
__device__ __forceinline__ V3 squareToGTR2(P2 s, float alpha) {
    // The reference's theta = acos(c), c = sqrt((1 - x) / (1 + (a2 - 1) x)) in float: for a glossy lobe c is within a few
    // ulps of 1 and its LAST BITS decide theta (alpha = 1e-3: one ulp of c is 10 % of theta).  c is therefore computed
    // with the reference's own roundings (IEEE division and square root), and sin(theta) = sin(acos(c)) from that c as
    // sqrt((1 - c)(1 + c)) (1 - c is exact).
    const float a2 = alpha * alpha;
    const float c = __fsqrt_rn(__fdiv_rn(1.0f - s.x, __fadd_rn(1.0f, __fmul_rn(a2 - 1.0f, s.x))));
    return sphericalDirCS(c, fsqrt(fmaxf(0.0f, (1.0f - c) * (1.0f + c))), s.y);
}
__device__ __forceinline__ float squareToGTR2Pdf(V3 m, float alpha) {
    // the reference evaluates 1 + (double) (a2 - 1.0f) * c^2 in double, where the FLOAT a2 - 1.0f has already rounded a2
    // to a2' = fl(a2 - 1) + 1 (for alpha = 1e-3 up to 3 % off a2); in float the same quantity without the cancellation
    // is sin^2 + a2' cos^2 with sin^2 = (1 - c)(1 + c)
    const float a2 = alpha * alpha, a2r = __fadd_rn(__fsub_rn(a2, 1.0f), 1.0f), c = m.z;
    const float den = __fmaf_rn(a2r, c * c, (1.0f - c) * (1.0f + c));
    const float pdf = fdiv(a2 * c * NORI_INV_PI, den * den);
    return (c >= 0 && fabsf(sqnorm(m) - 1.0f) < 1.0f) ? pdf : 0.0f;
}
#else
__device__ __forceinline__ V3 squareToGTR2(P2 s, float alpha) {                                                          // :180-185
    float a2 = (float) ((double) alpha * (double) alpha);
    return sphericalDir(acosf(sqrtf((1.0f - s.x) / (1.0f + (a2 - 1.0f) * s.x))), 2 * NORI_PI * s.y);
}
__device__ __forceinline__ float squareToGTR2Pdf(V3 m, float alpha) {                                                    // :187-193
    float a2 = (float) ((double) alpha * (double) alpha);
    float c = m.z;
    double den = 1 + (double) (a2 - 1.0f) * ((double) c * (double) c);
    float pdf = (float) ((double) (a2 * c * NORI_INV_PI) / (den * den));
    return (c >= 0 && fabsf(sqnorm(m) - 1.0f) < 1.0f) ? pdf : 0.0f;
}
#endif

// common.cpp:285-314
__device__ __forceinline__ float fresnel(float cosThetaI, float extIOR, float intIOR) {
    float etaI = extIOR, etaT = intIOR;
    if (extIOR == intIOR) return 0.0f;
    if (cosThetaI < 0.0f) { float t = etaI; etaI = etaT; etaT = t; cosThetaI = -cosThetaI; }
    float eta = fdiv(etaI, etaT), sinThetaTSqr = eta * eta * (1 - cosThetaI * cosThetaI);
    if (sinThetaTSqr > 1.0f) return 1.0f;
    float cosThetaT = fsqrt(1.0f - sinThetaTSqr);
    float Rs = fdiv(etaI * cosThetaI - etaT * cosThetaT, etaI * cosThetaI + etaT * cosThetaT);
    float Rp = fdiv(etaT * cosThetaI - etaI * cosThetaT, etaT * cosThetaI + etaI * cosThetaT);
    return (Rs * Rs + Rp * Rp) * 0.5f;
}
__device__ __forceinline__ float tanTheta(V3 v) { float t = 1 - v.z * v.z; if (t <= 0.0f) return 0.0f; return fdiv(fsqrt(t), v.z); }   // frame.h:81-86

// ------------------------------------------------------------------------------ BSDFs
enum { M_UNKNOWN = 0, M_SOLID_ANGLE = 1, M_DISCRETE = 2 };
struct BRec { V3 wi, wo; int measure; P2 uv; V3 albedo; };       // bsdf.h:30-58; albedo = m_albedo->eval(uv), looked up once per vertex

__device__ __forceinline__ V3 albedoAt(const DScene &sc, const nori_gpu_bsdf &b, P2 uv) {
    if (b.albedo_texture == NORI_TEXTURE_IMAGE) return imageEval<false>(sc.images[b.albedo_image], uv);   // imagetexture.cpp:118-136
    if (b.albedo_texture == NORI_TEXTURE_CHECKERBOARD) {          // checkerboard.cpp:31-37
        int x = (int) fabsf(floorf(uv.x / b.tex_scale[0] - b.tex_delta[0]));
        int y = (int) fabsf(floorf(uv.y / b.tex_scale[1] - b.tex_delta[1]));
        return x % 2 == y % 2 ? arr3(b.albedo) : arr3(b.albedo2);
    }
    return arr3(b.albedo);                                        // consttexture.cpp:30-32
}
__device__ __forceinline__ float evalBeckmann(float alpha, V3 m) {                                    // microfacet.cpp:52-58
    float temp = fdiv(tanTheta(m), alpha), ct = m.z, ct2 = ct * ct;
    return fdiv(expf(-temp * temp), NORI_PI * alpha * alpha * ct2 * ct2);
}
__device__ __forceinline__ float smithBeckmannG1(float alpha, V3 v, V3 m) {                            // microfacet.cpp:61-82
    float tt = tanTheta(v);
    if (tt == 0.0f) return 1.0f;
    if (dot(m, v) * v.z <= 0) return 0.0f;
    float a = frcp(alpha * tt);
    if (a >= 1.6f) return 1.0f;
    float a2 = a * a;
    return fdiv(3.535f * a + 2.181f * a2, 1.0f + 2.276f * a + 2.577f * a2);
}
__device__ __forceinline__ float schlickFresnel(float u) {                                             // disney.cpp:26-30
    float m = fminf(1.0f, fmaxf(0.0f, 1 - u));
#if NORI_FAST_SHADING
    const float m2 = m * m; return m2 * m2 * m;
#else
    double md = m; return (float) (md * md * md * md * md);
#endif
}
__device__ __forceinline__ float ggx(float NdotV, float alphaG) { float a = alphaG * alphaG, b = NdotV * NdotV; return frcp(NdotV + fsqrt(a + b - a * b)); }   // disney.cpp:32-37
__device__ __forceinline__ V3 lerp3(float t, V3 a, V3 b) { return (1.0f - t) * a + t * b; }                 // disney.cpp:40-43
__device__ __forceinline__ float luminance(V3 c) { return c.x * 0.212671f + c.y * 0.715160f + c.z * 0.072169f; }   // common.cpp:233-235

template <int TYPE>
__device__ __forceinline__ V3 bsdfEval(const nori_gpu_bsdf &b, const BRec &r) {
    if (TYPE == NORI_BSDF_DIFFUSE) {                              // diffuse.cpp:72-82
        if (r.measure != M_SOLID_ANGLE || r.wi.z <= 0 || r.wo.z <= 0) return mk(0.f);
        return r.albedo * NORI_INV_PI;
    } else if (TYPE == NORI_BSDF_MICROFACET) {                    // microfacet.cpp:84-94
        V3 n = xnormalized(xadd(r.wi, r.wo));                     // exact: 1 - n.z^2 (tanTheta) cancels for glossy lobes
        float D = evalBeckmann(b.alpha, n);
        float F = fresnel(dot(n, r.wi), b.extIOR, b.intIOR);
        float G = smithBeckmannG1(b.alpha, r.wi, n) * smithBeckmannG1(b.alpha, r.wo, n);
        float denom = 4.0f * r.wi.z * r.wo.z;
        float spec = fdiv(b.ks * D * F * G, denom);
        V3 kd = arr3(b.kd) * NORI_INV_PI;
        return mk(kd.x + spec, kd.y + spec, kd.z + spec);
    } else if (TYPE == NORI_BSDF_DISNEY) {                        // disney.cpp:63-105
        float NdotV = r.wi.z, NdotL = r.wo.z;
        if (NdotV < 0 || NdotL < 0) return mk(0.f);
        V3 wh = xnormalized(xadd(r.wi, r.wo));
        float LdotH = dot(r.wo, wh), VdotH = dot(r.wi, wh);
        V3 base = arr3(b.baseColor), white = mk(1.f);
        float lum = luminance(base);
        V3 Ctint = lum > 0.f ? base / lum : mk(1.0f);
#if NORI_FAST_SHADING
        V3 CtintMix = (b.specular * 0.08f) * lerp3(b.specularTint, white, Ctint);
        float fd90 = __fmaf_rn(2 * b.roughness, VdotH * VdotH, 0.5f);
#else
        V3 CtintMix = (float) ((double) b.specular * 0.08) * lerp3(b.specularTint, white, Ctint);
        float fd90 = (float) (0.5 + (double) (2 * b.roughness) * ((double) VdotH * (double) VdotH));
#endif
        V3 Cspec = lerp3(b.metallic, CtintMix, base);
        float fl = schlickFresnel(NdotL), fv = schlickFresnel(NdotV);
        V3 diffuse = base * NORI_INV_PI * (1.f + (fd90 - 1.f) * fl) * (1.f + (fd90 - 1.f) * fv);
        float alpha = fmaxf(0.001f, b.roughness * b.roughness);
        float Ds = squareToGTR2Pdf(wh, alpha);
        float FH = schlickFresnel(LdotH);
        V3 Fs = lerp3(FH, Cspec, white);
        float Gs = ggx(NdotL, alpha) * ggx(NdotV, alpha);
        V3 specular = Gs * Fs * Ds;
        V3 Fsheen = FH * b.sheen * lerp3(b.sheenTint, white, Ctint);
        return (1 - b.metallic) * (diffuse + Fsheen) + specular;
    }
    return mk(0.f);                                               // mirror.cpp:29-32, dielectric.cpp:35-38
}

template <int TYPE>
__device__ __forceinline__ float bsdfPdf(const nori_gpu_bsdf &b, const BRec &r) {
    if (TYPE == NORI_BSDF_DIFFUSE) {                              // diffuse.cpp:85-101
        if (r.measure != M_SOLID_ANGLE || r.wi.z <= 0 || r.wo.z <= 0) return 0.0f;
        return NORI_INV_PI * r.wo.z;
    } else if (TYPE == NORI_BSDF_MICROFACET) {                    // microfacet.cpp:97-111
        float c = r.wo.z; if (c <= 0.0f) return 0.0f;
        V3 n = xnormalized(xadd(r.wi, r.wo));
        float metallicTerm = fdiv(evalBeckmann(b.alpha, n) * n.z, 4.0f * fabsf(dot(n, r.wo)));
        return b.ks * metallicTerm + (1 - b.ks) * (c * NORI_INV_PI);
    } else if (TYPE == NORI_BSDF_DISNEY) {                        // disney.cpp:108-121
        float c = r.wo.z; if (c <= 0.0f) return 0.0f;
        V3 n = xnormalized(xadd(r.wi, r.wo));
        float metallicTerm = fdiv(squareToGTR2Pdf(n, b.alpha) * n.z, 4.0f * fabsf(dot(n, r.wo)));
        return (1 - b.metallic) * (c * NORI_INV_PI) + b.metallic * metallicTerm;
    }
    return 0.0f;
}

// BSDFQueryRecord for a surface vertex: the diffuse albedo texture (constant / checkerboard / image) is evaluated
// here, once, instead of inside every eval / sample call (it is a pure function of uv)
__device__ __forceinline__ BRec mkBRec(const DScene &sc, const nori_gpu_bsdf &bsdf, V3 wi, int measure, P2 uv) {
    BRec b; b.wi = wi; b.wo = mk(0.f); b.measure = measure; b.uv = uv;
    b.albedo = bsdf.type == NORI_BSDF_DIFFUSE ? albedoAt(sc, bsdf, uv) : mk(0.f);
    return b;
}

// returns the importance weight; on failure wo stays (0,0,0) like the reference's zero-filled
// TVector (vector.h:49), which makes the next ray miss everything
template <int TYPE>
__device__ __forceinline__ V3 bsdfSample(const nori_gpu_bsdf &b, BRec &r, P2 s) {
    r.wo = mk(0.f);
    if (TYPE == NORI_BSDF_DIFFUSE) {                              // diffuse.cpp:104-120
        if (r.wi.z <= 0) return mk(0.f);
        r.measure = M_SOLID_ANGLE; r.wo = squareToCosineHemisphere(s);
        return r.albedo;
    } else if (TYPE == NORI_BSDF_MIRROR) {                        // mirror.cpp:39-55
        if (r.wi.z <= 0) return mk(0.f);
        r.wo = mk(-r.wi.x, -r.wi.y, r.wi.z); r.measure = M_DISCRETE;
        return mk(1.f);
    } else if (TYPE == NORI_BSDF_DIELECTRIC) {                    // dielectric.cpp:45-73
        float theta = r.wi.z; V3 nv = mk(0.f, 0.f, 1.0f);
        if (fresnel(theta, b.extIOR, b.intIOR) > s.x) r.wo = mk(-r.wi.x, -r.wi.y, r.wi.z);
        else {
            float factor = fdiv(b.extIOR, b.intIOR);
            if (theta < 0.0f) { factor = frcp(factor); nv.z *= -1; }
            float win = dot(r.wi, nv);
            V3 part1 = -factor * (r.wi - win * nv);
#if NORI_FAST_SHADING
            V3 part2 = -nv * fsqrt(__fmaf_rn(-(factor * factor), __fmaf_rn(-win, win, 1.0f), 1.0f));
#else
            double f2 = (double) factor * (double) factor, w2 = (double) win * (double) win;
            V3 part2 = -nv * (float) sqrt(1 - f2 * (1 - w2));
#endif
            r.wo = normalized(part1 + part2);
        }
        r.measure = M_DISCRETE;
        return mk(1.f);
    } else if (TYPE == NORI_BSDF_MICROFACET) {                    // microfacet.cpp:114-137
        if (r.wi.z <= 0.0f) return mk(0.f);
        if (s.x < b.ks) {
            P2 ns; ns.x = fdiv(s.x, b.ks); ns.y = s.y;
            V3 n = squareToBeckmann(ns, b.alpha);
            r.wo = normalized((2.0f * dot(r.wi, n) * n) - r.wi);
        } else {
            P2 ns; ns.x = fdiv(s.x - b.ks, 1.f - b.ks); ns.y = s.y;
            r.wo = squareToCosineHemisphere(ns);
        }
        float c = r.wo.z; if (c <= 0.f) return mk(0.f);
        return bsdfEval<TYPE>(b, r) * c / bsdfPdf<TYPE>(b, r);
    } else {                                                      // disney.cpp:124-145
        if (r.wi.z <= 0.0f) return mk(0.f);
        if (s.x <= b.metallic) {
            P2 ns; ns.x = fdiv(s.x, b.metallic); ns.y = s.y;
            V3 n = squareToGTR2(ns, b.alpha);
            r.wo = normalized((2.0f * dot(r.wi, n) * n) - r.wi);
        } else {
            P2 ns; ns.x = fdiv(s.x - b.metallic, 1 - b.metallic); ns.y = s.y;
            r.wo = squareToCosineHemisphere(ns);
        }
        float c = r.wo.z; if (c <= 0.0f) return mk(0.f);
        return bsdfEval<TYPE>(b, r) * c / bsdfPdf<TYPE>(b, r);
    }
}

// runtime-dispatched versions: a switch on the BSDF type (warp-uniform in k_shade, whose queues are sorted by
// type).  Out of line in the one-thread-per-sample kernel (many call sites), inline in k_shade.
#ifndef NORI_DYN_INLINE
#define NORI_DYN_INLINE 0
#endif
#if NORI_DYN_INLINE
#define NORI_DYN __forceinline__
#else
#define NORI_DYN __noinline__
#endif
static __device__ NORI_DYN V3 bsdfEvalDyn(const nori_gpu_bsdf &b, const BRec &r) {
    switch (b.type) {
    case NORI_BSDF_DIFFUSE: return bsdfEval<NORI_BSDF_DIFFUSE>(b, r);
    case NORI_BSDF_MICROFACET: return bsdfEval<NORI_BSDF_MICROFACET>(b, r);
    case NORI_BSDF_DISNEY: return bsdfEval<NORI_BSDF_DISNEY>(b, r);
    default: return mk(0.f);
    }
}
static __device__ NORI_DYN float bsdfPdfDyn(const nori_gpu_bsdf &b, const BRec &r) {
    switch (b.type) {
    case NORI_BSDF_DIFFUSE: return bsdfPdf<NORI_BSDF_DIFFUSE>(b, r);
    case NORI_BSDF_MICROFACET: return bsdfPdf<NORI_BSDF_MICROFACET>(b, r);
    case NORI_BSDF_DISNEY: return bsdfPdf<NORI_BSDF_DISNEY>(b, r);
    default: return 0.f;
    }
}
static __device__ NORI_DYN V3 bsdfSampleDyn(const nori_gpu_bsdf &b, BRec &r, P2 s) {
    switch (b.type) {
    case NORI_BSDF_DIFFUSE: return bsdfSample<NORI_BSDF_DIFFUSE>(b, r, s);
    case NORI_BSDF_MIRROR: return bsdfSample<NORI_BSDF_MIRROR>(b, r, s);
    case NORI_BSDF_DIELECTRIC: return bsdfSample<NORI_BSDF_DIELECTRIC>(b, r, s);
    case NORI_BSDF_MICROFACET: return bsdfSample<NORI_BSDF_MICROFACET>(b, r, s);
    default: return bsdfSample<NORI_BSDF_DISNEY>(b, r, s);
    }
}

// ------------------------------------------------------------------------------ emitters
// Emitter functions take the compile-time flag AO ("area lights only"): scenes whose emitters are all area lights
// -- most of them -- run kernels in which the point / spot / environment-map code does not exist (k_shade is
// instruction-fetch bound: dropping that code is worth 7 % of the kernel on the Cornell box).  Same area code, same bits.
struct ERec { V3 ref, p, n, wi; float pdf; Ray shadow; };        // emitter.h:31-59
__device__ __forceinline__ ERec makeERec(V3 ref, V3 p, V3 n) {
    ERec e; e.ref = ref; e.p = p; e.n = n; e.wi = normalized(p - ref); e.pdf = 0.f; return e;
}
__device__ __forceinline__ ERec makeERec(V3 ref) {
    ERec e; e.ref = ref; e.p = mk(0.f); e.n = mk(0.f); e.wi = mk(0.f); e.pdf = 0.f; return e;
}

// dpdf.h:119-157 : std::lower_bound over m_cdf, then sample reuse
__device__ __forceinline__ uint32_t cdfSampleReuse(const float *cdf, uint32_t nEntries /* = F+1 */, float &s) {
    uint32_t lo = 0, len = nEntries;
    while (len > 0) {                                             // first element not less than s
        uint32_t half = len >> 1;
        if (__ldg(&cdf[lo + half]) < s) { lo += half + 1; len -= half + 1; } else len = half;
    }
    int idx = (int) lo - 1; if (idx < 0) idx = 0;
    if ((uint32_t) idx > nEntries - 2) idx = (int) nEntries - 2;
    NORI_CHECK(nEntries >= 2 && idx >= 0 && (uint32_t) idx + 1 < nEntries);
    float c0 = __ldg(&cdf[idx]), c1 = __ldg(&cdf[idx + 1]);
    s = fdiv(s - c0, c1 - c0);
    return (uint32_t) idx;
}

// mesh.cpp:40-61, sphere.cpp:95-105
__device__ __forceinline__ void sampleSurface(const DShape &m, P2 s, V3 &p, V3 &n, float &pdf) {
    if (m.type == NORI_SHAPE_MESH) {
        uint32_t idT = cdfSampleReuse(m.cdf, m.n_triangles + 1, s.x);
        V3 bc = squareToUniformTriangle(s);
        uint32_t i0 = __ldg(&m.F[3 * idT]), i1 = __ldg(&m.F[3 * idT + 1]), i2 = __ldg(&m.F[3 * idT + 2]);
        V3 p0 = ld3(&m.V[3 * i0]), p1 = ld3(&m.V[3 * i1]), p2 = ld3(&m.V[3 * i2]);
        p = bary(bc.x, p0, bc.y, p1, bc.z, p2);
        if (m.has_n) n = normalizedDyn(bary(bc.x, ld3(&m.N[3 * i0]), bc.y, ld3(&m.N[3 * i1]), bc.z, ld3(&m.N[3 * i2])));
        else n = normalized(cross(p1 - p0, p2 - p0));
        pdf = m.area_normalization;
#if NORI_WITH_PERLIN
    } else if (m.type == NORI_SHAPE_PERLIN) {                       // perlinnoise.cpp:77-86
        V3 q = squareToUniformSphere(s);
        p = mk(m.cx, m.cy, m.cz) + m.radius * q;
        const float r = perlinNoisedRadius(m.radius, m.perlin_height, m.perlin_scale, p);
        p = mk(m.cx, m.cy, m.cz) + r * q; n = q;
        const double ir = (double) (1.f / r);
        pdf = (float) (ir * ir * (double) (0.25f * NORI_INV_PI));
#endif
    } else {
        V3 q = squareToUniformSphere(s);
        p = mk(m.cx, m.cy, m.cz) + m.radius * q; n = q;
        pdf = m.sphere_pdf;
    }
}
__device__ __forceinline__ float pdfSurface(const DShape &m, V3 p) {
#if NORI_WITH_PERLIN
    if (m.type == NORI_SHAPE_PERLIN) {                              // perlinnoise.cpp:88-91
        const double ir = (double) (1.f / perlinNoisedRadius(m.radius, m.perlin_height, m.perlin_scale, p));
        return (float) (ir * ir * (double) (0.25f * NORI_INV_PI));
    }
#endif
    return m.type == NORI_SHAPE_MESH ? m.area_normalization : m.sphere_pdf;
}

// envmap.cpp:60-76
__device__ __forceinline__ P2 envMapIntersect(const nori_gpu_emitter &e, V3 vec) {
    float th = acosf(vec.z), ph = atan2f(vec.y, vec.x);
    if (ph < 0) ph += 2 * NORI_PI;
    P2 r; r.x = th * (e.env_rows - 1) * NORI_INV_PI; r.y = (float) ((double) ph * 0.5 * (e.env_cols - 1) * NORI_INV_PI);
    if (isnan(r.x) || isnan(r.y)) { r.x = 0.f; r.y = 0.f; }
    return r;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }

template <bool AO = false>
__device__ __forceinline__ V3 emitterEval(const DScene &sc, const nori_gpu_emitter &e, const ERec &l) {
    switch (AO ? (int) NORI_EMITTER_AREA : (int) e.type) {
    case NORI_EMITTER_AREA:                                       // arealight.cpp:39-44
        return dot(l.n, -l.wi) > 0.0f ? arr3(e.radiance) : mk(0.f);
    case NORI_EMITTER_POINT:                                      // pointlight.cpp:26-29
        return arr3(e.radiance) / (4.f * NORI_PI * sqnorm(arr3(e.position) - l.ref));
    case NORI_EMITTER_SPOT: {                                     // spotlight.cpp:38-42
        V3 c = arr3(e.radiance) / (4.f * NORI_PI);
        return c * 2.f * NORI_PI * (float) (1 - 0.5 * (double) (e.cosFalloffStart + e.cosTotalWidth));
    }
    default: {                                                    // envmap.cpp:124-156
        P2 uv = envMapIntersect(e, normalized(l.wi));
        int W = e.env_rows, H = e.env_cols;
        int u = clampi((int) uv.x, 0, W - 1), v = clampi((int) uv.y, 0, H - 1);
        int us = (u + 1) % W, vs = (v + 1) % H;
        V3 BL = ld3(&e.env_image[((size_t) u * H + v) * 3]), UL = ld3(&e.env_image[((size_t) u * H + vs) * 3]);
        V3 BR = ld3(&e.env_image[((size_t) us * H + v) * 3]), UR = ld3(&e.env_image[((size_t) us * H + vs) * 3]);
        int dusu = us - u, dvsv = vs - v;
        float dusum = us - uv.x, dumu = uv.x - u, dvmv = uv.y - v, dvsvm = vs - uv.y;
        float k = (float) (1.0 / (double) (dusu * dvsv));
        return e.weight * (k * ((BL * dusum * dvsvm) + (BR * dumu * dvsvm) + (UL * dusum * dvmv) + (UR * dumu * dvmv)));
    }
    }
}

template <bool AO = false>
__device__ __forceinline__ float emitterPdf(const DScene &sc, const nori_gpu_emitter &e, const ERec &l) {
    switch (AO ? (int) NORI_EMITTER_AREA : (int) e.type) {
    case NORI_EMITTER_AREA: return dot(l.n, -l.wi) > 0.0f ? pdfSurface(sc.shapes[e.shape], l.p) : 0.0f;   // arealight.cpp:64-76
    case NORI_EMITTER_POINT: return 1.0f;                         // pointlight.cpp:30-33
    case NORI_EMITTER_SPOT: return l.pdf;                         // spotlight.cpp:44-47
    default: {                                                    // envmap.cpp:184-192
        P2 its = envMapIntersect(e, normalized(l.wi));
        int i = clampi((int) its.x, 0, e.env_rows - 1), j = clampi((int) its.y, 0, e.env_cols - 1);
        return __ldg(&e.env_pmarginal[i]) * __ldg(&e.env_pdf[(size_t) i * e.env_cols + j]);
    }
    }
}

// envmap.cpp:112-121.  The reference scans i = 0,1,.. for the first interval with Pf[i] <= s < Pf[i+1];
// Pf is non-decreasing (a running sum of non-negative terms, last entry 1), so that interval is
// upper_bound(s) - 1 and a binary search returns the same index in O(log n) instead of O(n) -- it
// matters because the reference's table puts almost every sample in the LAST interval.
// When no interval matches (s >= 1) the reference reads past the table; here the last interval is used.
__device__ __forceinline__ void envSample1D(const float *pfRow, const float *PfRow, int nPf, float s, float &x, float &prob) {
    int lo = 0, hi = nPf - 1;                                   // first index in [0, nPf-1] with Pf > s (or nPf-1)
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(&PfRow[mid]) > s) hi = mid; else lo = mid + 1;
    }
    int i = lo - 1;
    if (i < 0) i = 0;
    if (i > nPf - 2) i = nPf - 2;
    const float P0 = __ldg(&PfRow[i]), P1 = __ldg(&PfRow[i + 1]);
    const float t = (P1 - s) / (P1 - P0);
    x = (1 - t) * i + t * (i + 1);
    prob = __ldg(&pfRow[i]);
}

template <bool AO = false>
__device__ __forceinline__ V3 emitterSample(const DScene &sc, const nori_gpu_emitter &e, ERec &l, P2 s) {
    switch (AO ? (int) NORI_EMITTER_AREA : (int) e.type) {
    case NORI_EMITTER_AREA: {                                     // arealight.cpp:46-62
        float spdf;
        sampleSurface(sc.shapes[e.shape], s, l.p, l.n, spdf);
        l.wi = normalized(l.p - l.ref);
        l.shadow = mkray(l.ref, l.wi, NORI_EPS, norm(l.p - l.ref) - NORI_EPS);
        l.pdf = emitterPdf<AO>(sc, e, l);
        float att = fdiv(dot(l.n, -l.wi), sqnorm(l.p - l.ref));
        return l.pdf > 0.0f ? emitterEval<AO>(sc, e, l) * att / l.pdf : mk(0.f);
    }
    case NORI_EMITTER_POINT: {                                    // pointlight.cpp:15-24
        V3 pos = arr3(e.position);
        l.wi = normalized(pos - l.ref); l.p = pos; l.pdf = 1.0f;
        l.shadow = mkray(l.ref, l.wi, NORI_EPS, norm(pos - l.ref) - NORI_EPS);
        return arr3(e.radiance) / (4.f * NORI_PI * sqnorm(pos - l.ref));
    }
    case NORI_EMITTER_SPOT: {                                     // spotlight.cpp:19-36
        V3 pos = arr3(e.position), dir = arr3(e.direction);
        l.wi = normalized(pos - l.ref); l.p = pos; l.pdf = 1.0f; l.n = dir;
        l.shadow = mkray(l.ref, l.wi, NORI_EPS, norm(pos - l.ref) - NORI_EPS);
        float cosTheta = dot(dir, normalized(-l.wi)), fall;
        if (cosTheta < e.cosTotalWidth) fall = 0;
        else if (cosTheta > e.cosFalloffStart) fall = 1;
        else fall = (acosf(e.cosTotalWidth) - acosf(cosTheta)) / (acosf(e.cosTotalWidth) - acosf(e.cosFalloffStart));
        return arr3(e.radiance) * fall / (4.f * NORI_PI * sqnorm(l.ref - l.p));
    }
    default: {                                                    // envmap.cpp:158-181
        int W = e.env_rows, H = e.env_cols;
        float st2 = 1.0f - l.wi.z * l.wi.z, sinTheta = st2 <= 0.0f ? 0.0f : sqrtf(st2);   // lRec.wi is still (0,0,0) here => 1
        float jacobian = (float) ((double) ((H - 1) * (W - 1)) / (2 * ((double) NORI_PI * (double) NORI_PI) * (double) sinTheta));
        float u, v, up, vp;
        envSample1D(e.env_pmarginal, e.env_cmarginal, W + 1, s.x, u, up);
        int row = clampi((int) u, 0, W - 1);
        envSample1D(&e.env_pdf[(size_t) row * H], &e.env_cdf[(size_t) row * (H + 1)], H + 1, s.y, v, vp);
        float theta = u * NORI_PI / (W - 1), phi = v * 2 * NORI_PI / (H - 1);
        l.wi = normalized(mk(sinf(theta) * cosf(phi), sinf(theta) * sinf(phi), cosf(theta)));
        l.shadow = mkray(l.ref, l.wi, NORI_EPS, 100000.f);
        vp = emitterPdf<AO>(sc, e, l) * jacobian;
        return emitterEval<AO>(sc, e, l) / vp;
    }
    }
}

// scene.h:68-74
__device__ __forceinline__ int randomEmitter(const DScene &sc, float rnd) {
    uint32_t n = sc.n_emitters;
    uint32_t idx = (uint32_t) floorf((float) n * rnd);
    return (int) (idx < n - 1 ? idx : n - 1);
}

// ------------------------------------------------------------------------------ cameras
// NR selects the slow-path-free IEEE sequences (device_common.cuh); NR = false the compiler's own __fdiv_rn / __fsqrt_rn.
// Same bits either way (nori_gpu_selftest); the state-machine kernels use NR = false: inlined into their refill step the
// fast sequences cost registers their node / leaf steps need (k_extend_sm spilled, 10M-triangle scene 136 -> 149 ms).
template <bool NR> __device__ __forceinline__ V3 cnorm(V3 a) { return NR ? xnormalized_nr(a) : xnormalized(a); }
template <bool NR> __device__ __forceinline__ float cdiv(float a, float b) { return NR ? xdiv_nr(a, b) : __fdiv_rn(a, b); }
template <bool NR = true>
__device__ __forceinline__ V3 xfPoint(const float *m, V3 p) {                // transform.h:78-81
    float r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = ((m[4 * i] * p.x + m[4 * i + 1] * p.y) + m[4 * i + 2] * p.z) + m[4 * i + 3] * 1.0f;
    return NR ? xdivs_nr(mk(r[0], r[1], r[2]), r[3]) : xdivs(mk(r[0], r[1], r[2]), r[3]);
}
__device__ __forceinline__ V3 xfVector(const float *m, V3 v) {               // transform.h:68-70
    return mk((m[0] * v.x + m[1] * v.y) + m[2] * v.z, (m[4] * v.x + m[5] * v.y) + m[6] * v.z, (m[8] * v.x + m[9] * v.y) + m[10] * v.z);
}
__device__ __forceinline__ P2 squareToUniformDisk(P2 s) {                    // warp.cpp:53-58
    float angle = 2 * s.x * NORI_PI, size = sqrtf(s.y);
    P2 r; r.x = cosf(angle) * size; r.y = sinf(angle) * size; return r;
}
__device__ __forceinline__ bool hasChromaticAberrations(const nori_gpu_camera &c) {   // advancedCamera.cpp:230-232
    return c.type == NORI_CAMERA_ADVANCED && !(c.chromatic[0] == 0.f && c.chromatic[1] == 0.f && c.chromatic[2] == 0.f);
}
// Camera rays are built in the EXACT arithmetic in every build (x* operations; IEEE division / square root through the
// slow-path-free sequences xdiv_nr / xsqrt_nr, whose operands here -- clip distances, homogeneous w, ray lengths -- are of
// ordinary magnitude): the first vertex of every path -- hit primitive, hit point -- then equals the reference's for the
// same film sample.
// perspective.cpp:90-112, thinlens.cpp:126-171, advancedCamera.cpp:133-228.  `weight` is the camera's importance
// weight: Color3f(1), or the unit colour of `channel` when chromatic aberration is on (advancedCamera.cpp:176-183).
template <bool NR = true>
__device__ __forceinline__ Ray cameraRay(const nori_gpu_camera &c, P2 ps, P2 as, int channel, V3 &weight) {
    V3 nearP = xfPoint<NR>(c.sampleToCamera, mk(ps.x * c.invOutputSize[0], ps.y * c.invOutputSize[1], 0.0f));
    V3 d = cnorm<NR>(nearP);
    weight = mk(1.f);
    Ray ray;
    if (c.type == NORI_CAMERA_ADVANCED) {
        const bool chroma = hasChromaticAberrations(c);
        if (!(c.distortion[0] == 0.f && c.distortion[1] == 0.f)) {           // advancedCamera.cpp:145-170
            const float qx = nearP.x / nearP.z, qy = nearP.y / nearP.z;
            const float y = sqrtf(qx * qx + qy * qy);
            float r = y, r2, f, df; int i = 0;
            while (true) {
                r2 = r * r;
                f = r * (1 + (c.distortion[0] * r2) + c.distortion[1] * (r2 * r2)) - y;
                df = 1 + (3 * c.distortion[0] * r2) + (5 * c.distortion[1] * r2 * r2);
                r = r - f / df;
                if ((double) fabsf(f) < 1e-6 || i++ > 4) break;
            }
            const float distortionFactor = r / y;
            nearP.x *= distortionFactor; nearP.y *= distortionFactor;
            d = cnorm<NR>(nearP);
        }
        float w = 0.0f;
        if (chroma) { w = c.chromatic[channel]; weight = mk(channel == 0 ? 1.f : 0.f, channel == 1 ? 1.f : 0.f, channel == 2 ? 1.f : 0.f); }
        const float invZ = cdiv<NR>(1.0f, d.z);
        if (c.lensRadius > 0.0f || chroma) {                                  // advancedCamera.cpp:192-216
            P2 disk = squareToUniformDisk(as);
            float lx = c.lensRadius * disk.x, ly = c.lensRadius * disk.y;
            float ft = cdiv<NR>(c.focalDistance, d.z);
            V3 pFocus = xadd(mk(0.f), xscale(d, ft));
            float spx = ps.x - (0.5f * (float) c.width), spy = ps.y - (0.5f * (float) c.height);
            const float mx = (float) max(c.width, c.height);
            spx /= mx; spy /= mx;
            const float sq = spx * spx + spy * spy;
            const float dx = spx * sq * w, dy = spy * sq * w;
            pFocus = xadd(pFocus, mk(-dx, dy, 0.0f));
            V3 o = mk(lx, ly, 0.0f);
            V3 dir = cnorm<NR>(xsub(pFocus, o));
            ray.o = xfPoint<NR>(c.cameraToWorld, o); ray.d = xfVector(c.cameraToWorld, dir);
        } else {
            ray.o = xfPoint<NR>(c.cameraToWorld, mk(0.f)); ray.d = xfVector(c.cameraToWorld, d);
        }
        ray.mint = c.nearClip * invZ; ray.maxt = c.farClip * invZ;
        return ray;
    }
    float invZ = cdiv<NR>(1.0f, d.z);
    if (c.type == NORI_CAMERA_THINLENS && c.lensRadius > 0.0f) {
        P2 disk = squareToConcentricDisk(as);
        float lx = c.lensRadius * disk.x, ly = c.lensRadius * disk.y;
        float ft = cdiv<NR>(c.focalDistance, d.z);
        V3 pFocus = xadd(mk(0.f), xscale(d, ft));
        V3 o = mk(lx, ly, 0.0f);
        V3 dir = cnorm<NR>(xsub(pFocus, o));
        ray.o = xfPoint<NR>(c.cameraToWorld, o); ray.d = xfVector(c.cameraToWorld, dir);
    } else {
        ray.o = xfPoint<NR>(c.cameraToWorld, mk(0.f)); ray.d = xfVector(c.cameraToWorld, d);
    }
    ray.mint = c.nearClip * invZ; ray.maxt = c.farClip * invZ;
    return ray;
}

__device__ __forceinline__ bool validColor(V3 c) {                           // common.cpp:224-231
    return !(c.x < 0 || !isfinite(c.x) || c.y < 0 || !isfinite(c.y) || c.z < 0 || !isfinite(c.z));
}
