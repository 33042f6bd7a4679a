/* gpu_binding.h -- the reference-side binding of INTEGRATION.md, written out in full.
 *
 * GpuBinding::flatten walks a scene the REFERENCE loaded (its parser, OBJ loader and SAH BVH builder) and fills the
 * POD description of include/nori_gpu.h with pointers into the reference's own containers (BVH nodes / indices /
 * shape offsets, mesh V / N / UV / F, area CDFs) plus the few tables it has to rebuild because the reference keeps
 * them private to a .cpp file (camera matrices, environment-map tables, texture texels, filter table).
 *
 * This TU is compiled with -fno-access-control (the reference has no accessor for BVH::m_nodes etc., bvh.h:165-170;
 * a maintainer would add `friend struct GpuBinding;`).  Plugin parameters are private members of classes that live
 * in .cpp files; the binding reads them from the PropertyList each object was constructed from, captured by
 * wrapping NoriObjectFactory::m_constructors BEFORE the scene is parsed (installFactoryHook).
 *
 * Used by oracle/ref_tools/nori_gpu_main.cpp (the reference's headless front-end rendering through libnori_gpu.so).
 * Test / integration tooling: linked against the unmodified reference objects, built only into oracle/_ref/. */
#pragma once
#include <nori/parser.h>
#include <nori/scene.h>
#include <nori/camera.h>
#include <nori/integrator.h>
#include <nori/sampler.h>
#include <nori/mesh.h>
#include <nori/bsdf.h>
#include <nori/emitter.h>
#include <nori/rfilter.h>
#include <nori/bitmap.h>
#include <nori/medium.h>
#include <filesystem/resolver.h>
#include <stb_image.h>
#include <Eigen/Geometry>
#include <cstring>
#include <deque>
#include <map>
#include "nori_gpu.h"

namespace gpubind {
using namespace nori;

/* ---- what every NoriObject was constructed from ------------------------------------------------------------ */
struct Created { std::string type; PropertyList props; int seq; int depth; };
inline std::map<const NoriObject *, Created> &created() { static std::map<const NoriObject *, Created> m; return m; }
inline void installFactoryHook() {
    static int seq = 0, depth = 0;
    for (auto &kv : *NoriObjectFactory::m_constructors) {
        NoriObjectFactory::Constructor orig = kv.second;
        std::string name = kv.first;
        kv.second = [orig, name](const PropertyList &p) -> NoriObject * {
            ++depth; NoriObject *o = orig(p); --depth;
            created()[o] = Created{name, p, seq++, depth};
            return o;
        };
    }
}
inline const Created &info(const NoriObject *o) {
    auto it = created().find(o);
    if (it == created().end()) throw NoriException("gpu binding: object was not created through the factory");
    return it->second;
}

inline void copy3(float *dst, const Eigen::Array3f &v) { dst[0] = v[0]; dst[1] = v[1]; dst[2] = v[2]; }
inline void copy3(float *dst, const Eigen::Vector3f &v) { dst[0] = v[0]; dst[1] = v[1]; dst[2] = v[2]; }
inline void copyMat(float *dst, const Eigen::Matrix4f &m) { for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) dst[4 * r + c] = m(r, c); }

/* ---- environment map tables: private to envmap.cpp, rebuilt with the arithmetic of its constructor
 *      (envmap.cpp:31-58, 90-110), quirks included ------------------------------------------------ */
struct EnvTables { int rows = 0, cols = 0; std::vector<float> image, pdf, cdf, pmarg, cmarg; };
inline float envPrecompute1D(int row, const Matf &f, Matf &pf, Matf &Pf) {
    float res = 0; int i;
    for (i = 0; i < f.cols(); i++) res = i + f(row, i);
    if (res == 0) return res;
    for (int j = 0; j < f.cols(); j++) pf(row, j) = f(row, j) / res;
    Pf(row, 0) = 0;
    for (i = 1; i < f.cols(); i++) Pf(row, i) = Pf(row, i - 1) + pf(i - 1);
    Pf(row, i) = 1;
    return res;
}
inline EnvTables buildEnvTables(const PropertyList &props) {
    std::string fn = getFileResolver()->resolve(props.getString("filename", "textures/envmaptext.exr")).str();
    Vector3f lumScale = props.getVector3("luminanceScale", Vector3f(0.3f, 0.6f, 0.1f));
    Bitmap img(fn);
    EnvTables t; t.cols = (int) img.cols(); t.rows = (int) img.rows();
    int width = t.rows, height = t.cols;      /* the reference's swapped names */
    Matf lum(width, height), pdf = Matf::Zero(width, height), cdf = Matf::Zero(width, height + 1),
         pm = Matf::Zero(1, width), cm = Matf::Zero(1, width + 1);
    for (int i = 0; i < width; i++)
        for (int j = 0; j < height; j++)
            lum(i, j) = sqrt(lumScale.x() * img(i, j).r() + lumScale.y() * img(i, j).g() + lumScale.z() * img(i, j).b())
                        + Epsilon / 10000000;
    Matf sum(1, width);
    for (int i = 0; i < pdf.rows(); ++i) sum(0, i) = envPrecompute1D(i, lum, pdf, cdf);
    envPrecompute1D(0, sum, pm, cm);
    t.image.resize((size_t) width * height * 3);
    for (int i = 0; i < width; i++) for (int j = 0; j < height; j++) for (int k = 0; k < 3; ++k)
        t.image[((size_t) i * height + j) * 3 + k] = img(i, j)[k];
    t.pdf.assign(pdf.data(), pdf.data() + pdf.size()); t.cdf.assign(cdf.data(), cdf.data() + cdf.size());
    t.pmarg.assign(pm.data(), pm.data() + pm.size()); t.cmarg.assign(cm.data(), cm.data() + cm.size());
    return t;
}

/* ---- everything the description points to that the reference does not already own ------------------------- */
struct Storage {
    std::vector<nori_gpu_shape> shapes;
    std::vector<nori_gpu_bsdf> bsdfs;
    std::vector<nori_gpu_emitter> emitters;
    std::vector<nori_gpu_image> images;
    std::deque<std::vector<uint8_t>> texels;       /* stbi_load(.., STBI_rgb) arrays (imagetexture.cpp:73-80) */
    std::deque<EnvTables> env;
    std::map<const BSDF *, int> bsdfIndex;

    int addImage(const PropertyList &props, const char *defaultFile) {
        std::string fn = getFileResolver()->resolve(props.getString("fileName", defaultFile)).str();
        int W = 0, H = 0, C = 0;
        uint8_t *data = stbi_load(fn.c_str(), &W, &H, &C, STBI_rgb);
        if (!data) throw NoriException("gpu binding: cannot load image '%s'", fn);
        texels.emplace_back(data, data + (size_t) W * H * 3); stbi_image_free(data);
        nori_gpu_image im; memset(&im, 0, sizeof(im));
        im.width = W; im.height = H;
        im.wrap = wrapTypeFromString(props.getString("wrap", "repeat")) == ImageWrap::Repeat ? NORI_WRAP_REPEAT : NORI_WRAP_CLAMP;
        im.rgb = texels.back().data();
        images.push_back(im);
        return (int) images.size() - 1;
    }
};

struct GpuBinding {
    /* fills `out`; `keep` must outlive every use of `out` (and so must the scene) */
    static void flatten(const Scene *scene, nori_gpu_scene &out, Storage &keep) {
        memset(&out, 0, sizeof(out));
        out.abi_version = NORI_GPU_ABI_VERSION;
        static const std::map<std::string, int> integrators = {
            {"normals", NORI_INTEGRATOR_NORMALS}, {"path_mis", NORI_INTEGRATOR_PATH_MIS},
            {"path_mats", NORI_INTEGRATOR_PATH_MATS}, {"direct_ems", NORI_INTEGRATOR_DIRECT_EMS},
            {"direct_mats", NORI_INTEGRATOR_DIRECT_MATS}, {"direct_mis", NORI_INTEGRATOR_DIRECT_MIS},
            {"direct", NORI_INTEGRATOR_DIRECT}, {"av", NORI_INTEGRATOR_AV}, {"volumetric", NORI_INTEGRATOR_VOLUMETRIC}};
        const Created &ii = info(scene->getIntegrator());
        if (!integrators.count(ii.type)) throw NoriException("gpu binding: integrator '%s' is outside the hot path", ii.type);
        out.integrator = integrators.at(ii.type);
        out.av_length = ii.type == "av" ? ii.props.getFloat("length") : 0.f;

        /* ---- BVH: the reference's arrays, verbatim (bvh.h:127-170) */
        const BVH *bvh = scene->getBVH();
        static_assert(sizeof(BVH::BVHNode) == 32 && sizeof(nori_gpu_bvh_node) == 32, "node layout");
        out.nodes = reinterpret_cast<const nori_gpu_bvh_node *>(bvh->m_nodes.data()); out.n_nodes = (uint32_t) bvh->m_nodes.size();
        out.indices = bvh->m_indices.data(); out.n_indices = (uint32_t) bvh->m_indices.size();
        out.shape_offset = bvh->m_shapeOffset.data();

        /* ---- shapes + their BSDFs */
        const auto &shapes = bvh->m_shapes;
        keep.shapes.assign(shapes.size(), nori_gpu_shape());
        for (size_t s = 0; s < shapes.size(); ++s) {
            nori_gpu_shape &p = keep.shapes[s]; memset(&p, 0, sizeof(p));
            const Shape *sh = shapes[s];
            if (const Mesh *m = dynamic_cast<const Mesh *>(sh)) {      /* mesh.h:121-124: column-major, used in place */
                p.type = NORI_SHAPE_MESH;
                p.n_vertices = m->getVertexCount(); p.n_triangles = m->getPrimitiveCount();
                p.V = m->m_V.data(); p.F = m->m_F.data();
                p.N = m->m_N.size() > 0 ? m->m_N.data() : nullptr;
                p.UV = m->m_UV.size() > 0 ? m->m_UV.data() : nullptr;
                p.area_cdf = m->m_pdf.m_cdf.data();                    /* dpdf.h:194, built by Mesh::activate (mesh.cpp:32-38) */
                p.area_normalization = m->m_pdf.getNormalization();
                if (sh->m_normalMap) {                                 /* shape.cpp:59-66, used by mesh.cpp:147-155 */
                    const Created &ni = info(sh->m_normalMap);
                    if (ni.type != "NormalMap") throw NoriException("gpu binding: normal texture '%s' is not supported", ni.type);
                    p.normal_map = 1 + keep.addImage(ni.props, "textures/default.png");
                }
            } else if (info(sh).type == "sphere" || info(sh).type == "perlinsphere") {
                const PropertyList &pl = info(sh).props;
                p.type = info(sh).type == "sphere" ? NORI_SHAPE_SPHERE : NORI_SHAPE_PERLIN; p.n_triangles = 1;
                copy3(p.center, pl.getPoint3("center", Point3f()));
                p.radius = pl.getFloat("radius", 1.f);
                if (p.type == NORI_SHAPE_PERLIN) { p.perlin_height = pl.getFloat("height", 1.0f); p.perlin_scale = pl.getFloat("scale", 1.0f); }
            } else throw NoriException("gpu binding: shape '%s' is outside the hot-path scope", info(sh).type);
            p.bsdf = bsdfOf(sh->getBSDF(), keep);
            p.emitter = -1;
        }

        /* ---- emitters, in Scene::m_emitters order (scene.cpp:63-76) */
        const auto &lights = scene->getLights();
        keep.emitters.assign(lights.size(), nori_gpu_emitter());
        for (size_t e = 0; e < lights.size(); ++e) {
            nori_gpu_emitter &q = keep.emitters[e]; memset(&q, 0, sizeof(q)); q.shape = -1;
            const Created &ei = info(lights[e]); const PropertyList &pl = ei.props;
            for (size_t s = 0; s < shapes.size(); ++s)
                if (shapes[s]->getEmitter() == lights[e]) { q.shape = (int) s; keep.shapes[s].emitter = (int) e; }
            if (ei.type == "area") { q.type = NORI_EMITTER_AREA; copy3(q.radiance, pl.getColor("radiance")); }
            else if (ei.type == "point") {
                q.type = NORI_EMITTER_POINT; copy3(q.position, pl.getPoint3("position", Point3f()));
                copy3(q.radiance, pl.getColor("power", Color3f()));
            } else if (ei.type == "spotlight") {
                q.type = NORI_EMITTER_SPOT; copy3(q.position, pl.getPoint3("position"));
                copy3(q.radiance, pl.getColor("color"));
                Vector3f d = pl.getVector3("direction").normalized(); copy3(q.direction, d);
                q.cosFalloffStart = std::cos(M_PI / 180 * pl.getFloat("falloffStart"));
                q.cosTotalWidth = std::cos(M_PI / 180 * pl.getFloat("totalWidth"));
            } else if (ei.type == "envmap") {
                q.type = NORI_EMITTER_ENVMAP; q.weight = pl.getFloat("weight", 1.0f);
                keep.env.push_back(buildEnvTables(pl));
                const EnvTables &t = keep.env.back();
                q.env_rows = t.rows; q.env_cols = t.cols;
                q.env_image = t.image.data(); q.env_pdf = t.pdf.data(); q.env_cdf = t.cdf.data();
                q.env_pmarginal = t.pmarg.data(); q.env_cmarginal = t.cmarg.data();
            } else throw NoriException("gpu binding: emitter '%s' is outside the hot-path scope", ei.type);
        }
        out.shapes = keep.shapes.data(); out.n_shapes = (uint32_t) keep.shapes.size();
        out.bsdfs = keep.bsdfs.data(); out.n_bsdfs = (uint32_t) keep.bsdfs.size();
        out.emitters = keep.emitters.data(); out.n_emitters = (uint32_t) keep.emitters.size();
        out.images = keep.images.data(); out.n_images = (uint32_t) keep.images.size();

        /* ---- camera: the matrices are private to perspective.cpp / thinlens.cpp / advancedCamera.cpp, rebuilt with
         *      the same Eigen expressions (perspective.cpp:53-80) */
        const Camera *cam = scene->getCamera();
        const Created &ci = info(cam);
        nori_gpu_camera &c = out.camera;
        if (ci.type == "perspective") c.type = NORI_CAMERA_PERSPECTIVE;
        else if (ci.type == "thinlens") c.type = NORI_CAMERA_THINLENS;
        else if (ci.type == "advancedCamera") {                        /* advancedCamera.cpp:34-57 */
            c.type = NORI_CAMERA_ADVANCED;
            Vector2f dist = ci.props.getVector2("distortion", Vector2f::Zero());
            Vector3f chroma = ci.props.getVector3("chromaticAberation", Vector3f::Zero());
            c.distortion[0] = dist.x(); c.distortion[1] = dist.y();
            c.chromatic[0] = chroma.x(); c.chromatic[1] = chroma.y(); c.chromatic[2] = chroma.z();
        } else throw NoriException("gpu binding: camera '%s' is outside the hot-path scope", ci.type);
        c.width = cam->getOutputSize().x(); c.height = cam->getOutputSize().y();
        Vector2f inv = cam->getOutputSize().cast<float>().cwiseInverse();
        c.invOutputSize[0] = inv.x(); c.invOutputSize[1] = inv.y();
        Transform toWorld = ci.props.getTransform("toWorld", Transform());
        float fov = ci.props.getFloat("fov", 30.0f);
        c.nearClip = ci.props.getFloat("nearClip", 1e-4f); c.farClip = ci.props.getFloat("farClip", 1e4f);
        c.focalDistance = ci.props.getFloat("focalDist", 1.0f); c.lensRadius = ci.props.getFloat("lensRadius", 0.0f);
        {
            float aspect = c.width / (float) c.height;
            float recip = 1.0f / (c.farClip - c.nearClip), cot = 1.0f / std::tan(degToRad(fov / 2.0f));
            Eigen::Matrix4f perspective;
            perspective << cot, 0, 0, 0,  0, cot, 0, 0,  0, 0, c.farClip * recip, -c.nearClip * c.farClip * recip,  0, 0, 1, 0;
            Transform s2c = Transform(Eigen::DiagonalMatrix<float, 3>(Vector3f(0.5f, -0.5f * aspect, 1.0f)) *
                Eigen::Translation<float, 3>(1.0f, -1.0f / aspect, 0.0f) * perspective).inverse();
            copyMat(c.sampleToCamera, s2c.getMatrix());
            copyMat(c.cameraToWorld, toWorld.getMatrix());
        }

        /* ---- filter table exactly as ImageBlock::init tabulates it (block.cpp:54-64) */
        const ReconstructionFilter *rf = cam->getReconstructionFilter();
        out.filter.radius = rf->getRadius();
        for (int i = 0; i < NORI_FILTER_RESOLUTION; ++i) out.filter.table[i] = rf->eval((out.filter.radius * i) / NORI_FILTER_RESOLUTION);
        out.filter.table[NORI_FILTER_RESOLUTION] = 0.f;

        /* ---- medium (medium.cpp:8-20); Scene::m_medium is uninitialised without <medium> */
        if (ii.type == "volumetric") {
            const Created &mi = info(scene->getMedium());
            out.medium.present = 1;
            copy3(out.medium.sigma_a, mi.props.getColor("sigma_a")); copy3(out.medium.sigma_s, mi.props.getColor("sigma_s"));
            Vector3f sz = mi.props.getVector3("box_size").cwiseAbs(), org = mi.props.getVector3("box_origin");
            copy3(out.medium.bounds_min, Vector3f(org - sz)); copy3(out.medium.bounds_max, Vector3f(org + sz));
        }
    }

  private:
    static int bsdfOf(const BSDF *b, Storage &keep) {
        auto it = keep.bsdfIndex.find(b);
        if (it != keep.bsdfIndex.end()) return it->second;
        nori_gpu_bsdf q; memset(&q, 0, sizeof(q));
        const Created &bi = info(b); const PropertyList &pl = bi.props;
        if (bi.type == "diffuse") {
            q.type = NORI_BSDF_DIFFUSE; q.albedo_texture = NORI_TEXTURE_CONSTANT;
            if (pl.has("albedo")) copy3(q.albedo, pl.getColor("albedo"));
            else {
                /* a <texture name="albedo"> child is parsed (hence created) right before its BSDF */
                const Created *tex = nullptr;
                for (auto &kv : created())
                    if (kv.second.seq == bi.seq - 1 && kv.second.depth == 0 && kv.first->getClassType() == NoriObject::ETexture
                        && kv.first->getIdName() == "albedo") tex = &kv.second;
                if (!tex) { q.albedo[0] = q.albedo[1] = q.albedo[2] = 0.5f; }   /* diffuse.cpp:62-68 */
                else if (tex->type == "constant_color") copy3(q.albedo, tex->props.getColor("value", Color3f(0.f)));
                else if (tex->type == "checkerboard_color") {
                    q.albedo_texture = NORI_TEXTURE_CHECKERBOARD;
                    copy3(q.albedo, tex->props.getColor("value1", Color3f(0)));
                    copy3(q.albedo2, tex->props.getColor("value2", Color3f(1)));
                    Point2f d = tex->props.getPoint2("delta", Point2f(0)); Vector2f sc = tex->props.getVector2("scale", Vector2f(1));
                    q.tex_delta[0] = d.x(); q.tex_delta[1] = d.y(); q.tex_scale[0] = sc.x(); q.tex_scale[1] = sc.y();
                } else if (tex->type == "ImageTexture") {
                    q.albedo_texture = NORI_TEXTURE_IMAGE;
                    q.albedo_image = keep.addImage(tex->props, "textures/default.png");
                } else throw NoriException("gpu binding: texture '%s' is outside the hot-path scope (SURVEY 8f)", tex->type);
            }
        } else if (bi.type == "mirror") q.type = NORI_BSDF_MIRROR;
        else if (bi.type == "dielectric") {
            q.type = NORI_BSDF_DIELECTRIC;
            q.intIOR = pl.getFloat("intIOR", 1.5046f); q.extIOR = pl.getFloat("extIOR", 1.000277f);
        } else if (bi.type == "microfacet") {
            q.type = NORI_BSDF_MICROFACET;
            q.alpha = pl.getFloat("alpha", 0.1f);
            q.intIOR = pl.getFloat("intIOR", 1.5046f); q.extIOR = pl.getFloat("extIOR", 1.000277f);
            Color3f kd = pl.getColor("kd", Color3f(0.5f)); copy3(q.kd, kd);
            q.ks = 1 - kd.maxCoeff();                                  /* microfacet.cpp:48 */
        } else if (bi.type == "disney") {
            q.type = NORI_BSDF_DISNEY;
            q.metallic = pl.getFloat("metallic", 0.0f); q.specular = pl.getFloat("specular", 0.0f);
            q.roughness = pl.getFloat("roughness", 0.0f); q.sheen = pl.getFloat("sheen", 0.0f);
            q.sheenTint = pl.getFloat("sheenTint", 0.0f); q.specularTint = pl.getFloat("specularTint", 0.0f);
            copy3(q.baseColor, pl.getColor("baseColor", Color3f(0.0f)));
            q.alpha = std::max(1e-3, std::pow(q.roughness, 2));        /* disney.cpp:59 */
        } else throw NoriException("gpu binding: bsdf '%s' is outside the hot-path scope", bi.type);
        keep.bsdfs.push_back(q);
        return keep.bsdfIndex[b] = (int) keep.bsdfs.size() - 1;
    }
};
} // namespace gpubind
