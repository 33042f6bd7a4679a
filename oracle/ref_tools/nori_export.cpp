/* nori_export -- TEST TOOL, built only into oracle/_ref/ and linked against the UNMODIFIED reference
 * objects (compiled with -fno-access-control so that it can read the private BVH arrays,
 * bvh.h:165-170, which the reference exposes through no accessor).
 *
 *   nori_export scene.xml out.nscene [--rays N] [--seed S]
 *
 * 1. loads the scene with the reference's own parser / OBJ loader / SAH BVH builder
 *    (parser.cpp:28, obj.cpp:32, bvh.cpp:329);
 * 2. flattens it into the POD description of include/nori_gpu.h (this is the same walk the
 *    reference-side binding of INTEGRATION.md performs before nori_gpu_upload_scene);
 * 3. optionally generates ray batches (camera rays, cosine-hemisphere secondary rays, shadow rays to
 *    emitter samples) and records the reference's answers:  BVH::rayIntersect itself (t, shape, p,
 *    uv, frames) plus a replay of its loop (bvh.cpp:404-462) that calls the reference's own
 *    TBoundingBox::rayIntersect / Shape::rayIntersect to recover what the public API hides:
 *    the winning primitive index and the node-visit / primitive-test counts.  The replay is checked
 *    against the real function on every ray.
 *
 * Container: "NSCN0001", u32 count, then per entry: u32 name_len, name, u32 dtype
 * (0 f32, 1 u32, 2 i32, 3 u8), u32 ndim, u64 dims[ndim], raw little-endian data padded to 8 bytes.
 */
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
#include <nori/warp.h>
#include <nori/medium.h>
#include <nori/block.h>
#include <filesystem/resolver.h>
#include <pcg32.h>
#include <stb_image.h>
#include <Eigen/Geometry>
#include <fstream>
#include <map>
#include <cstring>
#include "nori_gpu.h"

using namespace nori;

/* ------------------------------------------------------------------ factory hook ------------ */
struct Created { std::string type; PropertyList props; int seq; int depth; };
static std::map<const NoriObject *, Created> g_created;
static int g_seq = 0, g_depth = 0;

static void installFactoryHook() {
    for (auto &kv : *NoriObjectFactory::m_constructors) {
        NoriObjectFactory::Constructor orig = kv.second;
        std::string name = kv.first;
        kv.second = [orig, name](const PropertyList &p) -> NoriObject * {
            ++g_depth;
            NoriObject *o = orig(p);
            --g_depth;
            g_created[o] = Created{name, p, g_seq++, g_depth};
            return o;
        };
    }
}

static const Created &info(const NoriObject *o) {
    auto it = g_created.find(o);
    if (it == g_created.end()) throw NoriException("nori_export: object was not created through the factory");
    return it->second;
}

/* ------------------------------------------------------------------ container writer -------- */
struct Writer {
    std::vector<uint8_t> buf; uint32_t count = 0;
    Writer() { buf.insert(buf.end(), {'N','S','C','N','0','0','0','1'}); buf.resize(12); }
    void raw(const void *p, size_t n) { buf.insert(buf.end(), (const uint8_t *) p, (const uint8_t *) p + n); }
    void add(const std::string &name, uint32_t dtype, std::vector<uint64_t> dims, const void *data) {
        size_t n = dtype == 3 ? 1 : 4; for (auto d : dims) n *= d;
        uint32_t nl = (uint32_t) name.size(), nd = (uint32_t) dims.size();
        raw(&nl, 4); raw(name.data(), nl); raw(&dtype, 4); raw(&nd, 4);
        for (auto d : dims) raw(&d, 8);
        if (n) raw(data, n);
        while (buf.size() % 8) buf.push_back(0);
        ++count;
    }
    void f32(const std::string &n, std::vector<uint64_t> d, const float *p) { add(n, 0, d, p); }
    void u32(const std::string &n, std::vector<uint64_t> d, const uint32_t *p) { add(n, 1, d, p); }
    void i32(const std::string &n, std::vector<uint64_t> d, const int32_t *p) { add(n, 2, d, p); }
    void bytes(const std::string &n, const void *p, size_t len) { add(n, 3, {len}, p); }
    void save(const std::string &fn) {
        memcpy(&buf[8], &count, 4);
        std::ofstream os(fn, std::ios::binary); os.write((const char *) buf.data(), buf.size());
    }
};

static void copy3(float *dst, const Eigen::Array3f &v) { dst[0] = v[0]; dst[1] = v[1]; dst[2] = v[2]; }
static void copy3(float *dst, const Eigen::Vector3f &v) { dst[0] = v[0]; dst[1] = v[1]; dst[2] = v[2]; }
static void copyMat(float *dst, const Eigen::Matrix4f &m) { for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) dst[4 * r + c] = m(r, c); }

/* ------------------------------------------------------------------ environment map tables --- *
 * EnvironmentMap keeps its tables private inside envmap.cpp, so the exporter rebuilds them with the
 * same arithmetic as its constructor (envmap.cpp:31-58, 90-110), quirks included. */
struct EnvTables { int rows, cols; std::vector<float> image, pdf, cdf, pmarg, cmarg; };

static float envPrecompute1D(int row, const Matf &f, Matf &pf, Matf &Pf) {
    float res = 0; int i;
    for (i = 0; i < f.cols(); i++) res = i + f(row, i);
    if (res == 0) return res;
    for (int j = 0; j < f.cols(); j++) pf(row, j) = f(row, j) / res;
    Pf(row, 0) = 0;
    for (i = 1; i < f.cols(); i++) Pf(row, i) = Pf(row, i - 1) + pf(i - 1);
    Pf(row, i) = 1;
    return res;
}

static EnvTables buildEnvTables(const PropertyList &props) {
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

/* ------------------------------------------------------------------ image textures ----------- *
 * ImageTexture / NormalMap are classes local to imagetexture.cpp / normalmap.cpp; both load their file
 * with stbi_load(path, .., STBI_rgb) (imagetexture.cpp:73-80, normalmap.cpp:73-80).  The exporter loads
 * the same file with the same call and stores the 8-bit RGB texels verbatim. */
struct ImageTable {
    std::vector<std::vector<uint8_t>> rgb; std::vector<int> w, h, wrap;
    int add(const PropertyList &props, const char *defaultFile) {
        std::string fn = getFileResolver()->resolve(props.getString("fileName", defaultFile)).str();
        int wrapMode = wrapTypeFromString(props.getString("wrap", "repeat")) == ImageWrap::Repeat ? NORI_WRAP_REPEAT : NORI_WRAP_CLAMP;
        int W = 0, H = 0, C = 0;
        uint8_t *data = stbi_load(fn.c_str(), &W, &H, &C, STBI_rgb);
        if (!data) throw NoriException("nori_export: cannot load image '%s'", fn);
        rgb.emplace_back(data, data + (size_t) W * H * 3); stbi_image_free(data);
        w.push_back(W); h.push_back(H); wrap.push_back(wrapMode);
        return (int) rgb.size() - 1;
    }
    void write(Writer &wr) const {
        for (size_t i = 0; i < rgb.size(); ++i) {
            wr.add("image." + std::to_string(i) + ".rgb", 3, {(uint64_t) h[i], (uint64_t) w[i], 3}, rgb[i].data());
            int32_t wm = wrap[i]; wr.i32("image." + std::to_string(i) + ".wrap", {1}, &wm);
        }
    }
};

/* ------------------------------------------------------------------ traversal replay --------- */
struct Replay { bool hit; float t, u, v; uint32_t shape, prim, nodes, prims; };

static Replay replay(const BVH *bvh, const Ray3f &_ray, bool shadowRay) {
    Replay r{false, std::numeric_limits<float>::infinity(), 0, 0, 0xffffffffu, 0xffffffffu, 0, 0};
    uint32_t node_idx = 0, stack_idx = 0, stack[64];
    Ray3f ray(_ray);
    if (ray.mint == Epsilon)
        ray.mint = std::max(ray.mint, ray.mint * ray.o.array().abs().maxCoeff());
    if (bvh->m_nodes.empty() || ray.maxt < ray.mint) return r;
    while (true) {
        const BVH::BVHNode &node = bvh->m_nodes[node_idx];
        ++r.nodes;
        bool descend = node.bbox.rayIntersect(ray);
        if (descend && node.isInner()) { stack[stack_idx++] = node.inner.rightChild; node_idx++; continue; }
        if (descend) {
            for (uint32_t i = node.start(), end = node.end(); i < end; ++i) {
                uint32_t idx = bvh->m_indices[i];
                uint32_t s = bvh->findShape(idx);
                float u = 0, v = 0, t;
                ++r.prims;
                if (bvh->m_shapes[s]->rayIntersect(idx, ray, u, v, t)) {
                    if (shadowRay) { r.hit = true; r.t = 0; return r; }
                    r.hit = true; ray.maxt = r.t = t; r.shape = s; r.prim = idx;
                    /* Sphere::rayIntersect never writes u,v (sphere.cpp:43-76); report 0 for them */
                    bool isMesh = dynamic_cast<const Mesh *>(bvh->m_shapes[s]) != nullptr;
                    r.u = isMesh ? u : 0.f; r.v = isMesh ? v : 0.f;
                }
            }
        }
        if (stack_idx == 0) break;
        node_idx = stack[--stack_idx];
    }
    return r;
}

int main(int argc, char **argv) {
    if (argc < 3) { cerr << "usage: nori_export scene.xml out.nscene [--rays N] [--special K] [--seed S] [--seq N] [--probe N]" << endl; return 1; }
    std::string xml = argv[1], out = argv[2];
    size_t nRays = 0, nSeq = 0, nProbe = 0, special = 0; uint64_t seed = 1;
    for (int i = 3; i + 1 < argc; i += 2) {
        if (!strcmp(argv[i], "--rays")) nRays = (size_t) atoll(argv[i + 1]);
        if (!strcmp(argv[i], "--special")) special = (size_t) atoll(argv[i + 1]);
        if (!strcmp(argv[i], "--seed")) seed = (uint64_t) atoll(argv[i + 1]);
        if (!strcmp(argv[i], "--seq")) nSeq = (size_t) atoll(argv[i + 1]);
        if (!strcmp(argv[i], "--probe")) nProbe = (size_t) atoll(argv[i + 1]);
    }
    try {
        installFactoryHook();
        filesystem::path path(xml);
        getFileResolver()->prepend(path.parent_path());
        std::unique_ptr<NoriObject> root(loadFromXML(xml));
        if (root->getClassType() != NoriObject::EScene) throw NoriException("root is not a scene");
        Scene *scene = static_cast<Scene *>(root.get());
        const BVH *bvh = scene->getBVH();
        Writer w;

        /* ---- header ---- */
        static const std::map<std::string, int> integrators = {
            {"normals", NORI_INTEGRATOR_NORMALS}, {"path_mis", NORI_INTEGRATOR_PATH_MIS},
            {"path_mats", NORI_INTEGRATOR_PATH_MATS}, {"direct_ems", NORI_INTEGRATOR_DIRECT_EMS},
            {"direct_mats", NORI_INTEGRATOR_DIRECT_MATS}, {"direct_mis", NORI_INTEGRATOR_DIRECT_MIS},
            {"direct", NORI_INTEGRATOR_DIRECT}, {"av", NORI_INTEGRATOR_AV}, {"volumetric", NORI_INTEGRATOR_VOLUMETRIC}};
        const Created &ii = info(scene->getIntegrator());
        if (!integrators.count(ii.type)) throw NoriException("nori_export: integrator '%s' is outside the hot path", ii.type);
        int32_t header[4] = {NORI_GPU_ABI_VERSION, integrators.at(ii.type), (int32_t) scene->getSampler()->getSampleCount(), 0};
        w.i32("header", {4}, header);
        float avLength = ii.type == "av" ? ii.props.getFloat("length") : 0.f;
        w.f32("av_length", {1}, &avLength);

        /* ---- BVH (bvh.h:127-170) ---- */
        static_assert(sizeof(BVH::BVHNode) == 32 && sizeof(nori_gpu_bvh_node) == 32, "node layout");
        w.u32("bvh.nodes", {bvh->m_nodes.size(), 8}, (const uint32_t *) bvh->m_nodes.data());
        w.u32("bvh.indices", {bvh->m_indices.size()}, bvh->m_indices.data());
        w.u32("bvh.shape_offset", {bvh->m_shapeOffset.size()}, bvh->m_shapeOffset.data());

        /* ---- shapes, bsdfs ---- */
        const auto &shapes = bvh->m_shapes;
        std::vector<nori_gpu_shape> pods(shapes.size());
        std::vector<nori_gpu_bsdf> bsdfs;
        std::map<const BSDF *, int> bsdfIndex;
        ImageTable images;
        for (size_t s = 0; s < shapes.size(); ++s) {
            nori_gpu_shape &p = pods[s]; memset(&p, 0, sizeof(p));
            const Shape *sh = shapes[s];
            std::string pre = "shape." + std::to_string(s) + ".";
            if (const Mesh *m = dynamic_cast<const Mesh *>(sh)) {
                p.type = NORI_SHAPE_MESH;
                p.n_vertices = m->getVertexCount(); p.n_triangles = m->getPrimitiveCount();
                w.f32(pre + "V", {p.n_vertices, 3}, m->m_V.data());
                if (m->m_N.size() > 0) w.f32(pre + "N", {p.n_vertices, 3}, m->m_N.data());
                if (m->m_UV.size() > 0) w.f32(pre + "UV", {p.n_vertices, 2}, m->m_UV.data());
                w.u32(pre + "F", {p.n_triangles, 3}, m->m_F.data());
                w.f32(pre + "area_cdf", {m->m_pdf.m_cdf.size()}, m->m_pdf.m_cdf.data());
                p.area_normalization = m->m_pdf.getNormalization();
                if (sh->m_normalMap) {                               /* shape.cpp:59-66, used by mesh.cpp:147-155 */
                    const Created &ni = info(sh->m_normalMap);
                    if (ni.type != "NormalMap") throw NoriException("nori_export: normal texture '%s' is not supported", ni.type);
                    p.normal_map = 1 + images.add(ni.props, "textures/default.png");
                }
            } else if (info(sh).type == "sphere") {
                p.type = NORI_SHAPE_SPHERE; p.n_triangles = 1;
                copy3(p.center, info(sh).props.getPoint3("center", Point3f()));
                p.radius = info(sh).props.getFloat("radius", 1.f);
            } else if (info(sh).type == "perlinsphere") {             /* perlinnoise.cpp:12-20 */
                p.type = NORI_SHAPE_PERLIN; p.n_triangles = 1;
                copy3(p.center, info(sh).props.getPoint3("center", Point3f()));
                p.radius = info(sh).props.getFloat("radius", 1.f);
                p.perlin_height = info(sh).props.getFloat("height", 1.0f);
                p.perlin_scale = info(sh).props.getFloat("scale", 1.0f);
            } else throw NoriException("nori_export: shape '%s' is outside the hot-path scope", info(sh).type);

            const BSDF *b = sh->getBSDF();
            if (!bsdfIndex.count(b)) {
                nori_gpu_bsdf q; memset(&q, 0, sizeof(q));
                const Created &bi = info(b); const PropertyList &pl = bi.props;
                if (bi.type == "diffuse") {
                    q.type = NORI_BSDF_DIFFUSE; q.albedo_texture = NORI_TEXTURE_CONSTANT;
                    if (pl.has("albedo")) copy3(q.albedo, pl.getColor("albedo"));
                    else {
                        /* a <texture name="albedo"> child is parsed (hence created) right before its BSDF */
                        const Created *tex = nullptr;
                        for (auto &kv : g_created)
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
                            q.albedo_image = images.add(tex->props, "textures/default.png");
                        } else throw NoriException("nori_export: texture '%s' is outside the hot-path scope (SURVEY 8f)", tex->type);
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
                    q.ks = 1 - kd.maxCoeff();
                } else if (bi.type == "disney") {
                    q.type = NORI_BSDF_DISNEY;
                    q.metallic = pl.getFloat("metallic", 0.0f); q.specular = pl.getFloat("specular", 0.0f);
                    q.roughness = pl.getFloat("roughness", 0.0f); q.sheen = pl.getFloat("sheen", 0.0f);
                    q.sheenTint = pl.getFloat("sheenTint", 0.0f); q.specularTint = pl.getFloat("specularTint", 0.0f);
                    copy3(q.baseColor, pl.getColor("baseColor", Color3f(0.0f)));
                    q.alpha = std::max(1e-3, std::pow(q.roughness, 2));           /* disney.cpp:59 */
                } else throw NoriException("nori_export: bsdf '%s' is outside the hot-path scope", bi.type);
                bsdfIndex[b] = (int) bsdfs.size(); bsdfs.push_back(q);
            }
            p.bsdf = bsdfIndex[b];
            p.emitter = -1;
        }

        /* ---- emitters, in Scene::m_emitters order (scene.cpp:63-76) ---- */
        const auto &lights = scene->getLights();
        std::vector<nori_gpu_emitter> ems(lights.size());
        for (size_t e = 0; e < lights.size(); ++e) {
            nori_gpu_emitter &q = ems[e]; memset(&q, 0, sizeof(q)); q.shape = -1;
            const Created &ei = info(lights[e]); const PropertyList &pl = ei.props;
            for (size_t s = 0; s < shapes.size(); ++s)
                if (shapes[s]->getEmitter() == lights[e]) { q.shape = (int) s; pods[s].emitter = (int) e; }
            std::string pre = "emitter." + std::to_string(e) + ".";
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
                EnvTables t = buildEnvTables(pl);
                q.env_rows = t.rows; q.env_cols = t.cols;
                w.f32(pre + "env_image", {(uint64_t) t.rows, (uint64_t) t.cols, 3}, t.image.data());
                w.f32(pre + "env_pdf", {(uint64_t) t.rows, (uint64_t) t.cols}, t.pdf.data());
                w.f32(pre + "env_cdf", {(uint64_t) t.rows, (uint64_t) t.cols + 1}, t.cdf.data());
                w.f32(pre + "env_pmarginal", {(uint64_t) t.rows}, t.pmarg.data());
                w.f32(pre + "env_cmarginal", {(uint64_t) t.rows + 1}, t.cmarg.data());
            } else throw NoriException("nori_export: emitter '%s' is outside the hot-path scope", ei.type);
        }
        w.bytes("shapes.pod", pods.data(), pods.size() * sizeof(nori_gpu_shape));
        w.bytes("bsdfs.pod", bsdfs.data(), bsdfs.size() * sizeof(nori_gpu_bsdf));
        w.bytes("emitters.pod", ems.data(), ems.size() * sizeof(nori_gpu_emitter));
        images.write(w);

        /* ---- camera: the matrices are private to perspective.cpp / thinlens.cpp, so rebuild them
         *      with the same Eigen expressions (perspective.cpp:53-80) ---- */
        const Camera *cam = scene->getCamera();
        const Created &ci = info(cam);
        nori_gpu_camera c; memset(&c, 0, sizeof(c));
        if (ci.type == "perspective") c.type = NORI_CAMERA_PERSPECTIVE;
        else if (ci.type == "thinlens") c.type = NORI_CAMERA_THINLENS;
        else if (ci.type == "advancedCamera") {                       /* advancedCamera.cpp:34-57 */
            c.type = NORI_CAMERA_ADVANCED;
            Vector2f dist = ci.props.getVector2("distortion", Vector2f::Zero());
            Vector3f chroma = ci.props.getVector3("chromaticAberation", Vector3f::Zero());
            c.distortion[0] = dist.x(); c.distortion[1] = dist.y();
            c.chromatic[0] = chroma.x(); c.chromatic[1] = chroma.y(); c.chromatic[2] = chroma.z();
        }
        else throw NoriException("nori_export: camera '%s' is outside the hot-path scope (SURVEY 8f)", ci.type);
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
        w.bytes("camera.pod", &c, sizeof(c));

        /* ---- filter table exactly as ImageBlock::init tabulates it (block.cpp:54-64) ---- */
        const ReconstructionFilter *rf = cam->getReconstructionFilter();
        nori_gpu_filter f; f.radius = rf->getRadius();
        for (int i = 0; i < NORI_FILTER_RESOLUTION; ++i) f.table[i] = rf->eval((f.radius * i) / NORI_FILTER_RESOLUTION);
        f.table[NORI_FILTER_RESOLUTION] = 0.f;
        w.bytes("filter.pod", &f, sizeof(f));

        /* ---- medium (medium.cpp:8-20); Scene::m_medium is uninitialised without <medium> (A.15) ---- */
        nori_gpu_medium md; memset(&md, 0, sizeof(md));
        if (ii.type == "volumetric") {
            const Created &mi = info(scene->getMedium());
            md.present = 1;
            copy3(md.sigma_a, mi.props.getColor("sigma_a")); copy3(md.sigma_s, mi.props.getColor("sigma_s"));
            Vector3f sz = mi.props.getVector3("box_size").cwiseAbs(), org = mi.props.getVector3("box_origin");
            copy3(md.bounds_min, Vector3f(org - sz)); copy3(md.bounds_max, Vector3f(org + sz));
        }
        w.bytes("medium.pod", &md, sizeof(md));

        /* ---- ray batches with the reference's answers ---- */
        if (nRays > 0) {
            pcg32 rng(seed, 7);
            std::vector<nori_gpu_ray> rays; std::vector<int32_t> shadowFlag;
            auto push = [&](const Ray3f &r, bool sh) {
                nori_gpu_ray q; copy3(q.o, r.o); copy3(q.d, r.d); q.mint = r.mint; q.maxt = r.maxt;
                rays.push_back(q); shadowFlag.push_back(sh ? 1 : 0);
            };
            while (rays.size() < nRays) {
                Point2f ps(rng.nextFloat() * c.width, rng.nextFloat() * c.height), as(rng.nextFloat(), rng.nextFloat());
                Ray3f ray; cam->sampleRay(ray, ps, as);
                push(ray, false);
                Intersection its;
                for (int bounce = 0; bounce < 3 && scene->rayIntersect(ray, its); ++bounce) {
                    /* shadow ray towards a sampled emitter (arealight.cpp:56 interval) */
                    if (!lights.empty()) {
                        EmitterQueryRecord eRec(its.p);
                        scene->getRandomEmitter(rng.nextFloat())->sample(eRec, Point2f(rng.nextFloat(), rng.nextFloat()));
                        push(eRec.shadowRay, true);
                        push(eRec.shadowRay, false);     /* same segment as a closest-hit query */
                    }
                    /* default-constructed secondary ray => adaptive epsilon path (bvh.cpp:410-412) */
                    Vector3f wo = Warp::squareToCosineHemisphere(Point2f(rng.nextFloat(), rng.nextFloat()));
                    if (rng.nextFloat() < 0.3f) wo = -wo;   /* also shoot below the surface (dielectric-like) */
                    ray = Ray3f(its.p, its.toWorld(wo));
                    push(ray, false);
                }
            }
            rays.resize(nRays); shadowFlag.resize(nRays);
            /* --special K: every K-th ray is turned into one of the special cases of the slab test (bbox.h:343-357):
               a zero (+0 / -0) or subnormal direction component (1/d infinite), for half of them with the origin
               exactly on a bounding plane of the scene or of a leaf, where (bound - o) * (1/d) is 0 * inf */
            if (special > 0) {
                const float comps[5] = { 0.0f, -0.0f, 1e-41f, -1e-41f, 1e-39f };
                for (size_t i = 0; i < nRays; i += special) {
                    const int a = (int) (rng.nextUInt() % 3u);
                    rays[i].d[a] = comps[rng.nextUInt() % 5u];
                    const uint32_t pick = rng.nextUInt() % 4u;
                    if (pick < 2u) {
                        const BVH::BVHNode &nd = bvh->m_nodes[pick == 0u ? 0u : rng.nextUInt() % (uint32_t) bvh->m_nodes.size()];
                        rays[i].o[a] = (rng.nextUInt() & 1u) ? nd.bbox.min[a] : nd.bbox.max[a];
                    }
                }
            }
            std::vector<nori_gpu_hit> hits(nRays);
            std::vector<float> hp(nRays * 3, 0.f), huv(nRays * 2, 0.f), hn(nRays * 3, 0.f), hg(nRays * 3, 0.f);
            size_t mismatches = 0;
            for (size_t i = 0; i < nRays; ++i) {
                Ray3f r(Point3f(rays[i].o[0], rays[i].o[1], rays[i].o[2]), Vector3f(rays[i].d[0], rays[i].d[1], rays[i].d[2]),
                        rays[i].mint, rays[i].maxt);
                Intersection its;
                bool hit = bvh->rayIntersect(r, its, shadowFlag[i] != 0);          /* the real thing */
                Replay rp = replay(bvh, r, shadowFlag[i] != 0);
                nori_gpu_hit &h = hits[i]; memset(&h, 0, sizeof(h));
                h.t = rp.t; h.u = rp.u; h.v = rp.v; h.shape = rp.shape; h.prim = rp.prim;
                h.nodes_visited = rp.nodes; h.prims_tested = rp.prims;
                if (hit != rp.hit) ++mismatches;
                if (hit && !shadowFlag[i]) {
                    if (its.t != rp.t || its.mesh != shapes[rp.shape]) ++mismatches;
                    copy3(&hp[3 * i], its.p); huv[2 * i] = its.uv.x(); huv[2 * i + 1] = its.uv.y();
                    copy3(&hn[3 * i], its.shFrame.n); copy3(&hg[3 * i], its.geoFrame.n);
                }
            }
            if (mismatches) throw NoriException("nori_export: traversal replay disagrees with BVH::rayIntersect on %i rays", (int) mismatches);
            w.f32("rays", {nRays, 8}, (const float *) rays.data());
            w.i32("rays.shadow", {nRays}, shadowFlag.data());
            w.bytes("rays.hits", hits.data(), hits.size() * sizeof(nori_gpu_hit));
            w.f32("rays.hit_p", {nRays, 3}, hp.data()); w.f32("rays.hit_uv", {nRays, 2}, huv.data());
            w.f32("rays.hit_n", {nRays, 3}, hn.data()); w.f32("rays.hit_ng", {nRays, 3}, hg.data());
        }

        /* ---- the reference's own sample sequence for block (0,0): renderBlock's loop (render.cpp:96-126)
         *      driven by the reference's Independent sampler, one row per sample:
         *      (pixelSample.x, pixelSample.y, r, g, b) -- pins the RNG consumption order ---- */
        if (nSeq > 0) {
            std::unique_ptr<Sampler> sampler(scene->getSampler()->clone());
            ImageBlock blk(Vector2i(NORI_BLOCK_SIZE), cam->getReconstructionFilter());
            blk.setOffset(Point2i(0, 0));
            sampler->prepare(blk);
            std::vector<float> seq; size_t done = 0;
            int bw = std::min(NORI_BLOCK_SIZE, c.width), bh = std::min(NORI_BLOCK_SIZE, c.height);
            while (done < nSeq)
                for (int y = 0; y < bh && done < nSeq; ++y)
                    for (int x = 0; x < bw && done < nSeq; ++x, ++done) {
                        Point2f pixelSample = Point2f((float) x, (float) y) + sampler->next2D();
                        Point2f apertureSample = sampler->next2D();
                        Color3f value(0.0f);
                        if (cam->hasChromaticAberrations()) {         /* render.cpp:106-121: one path per colour channel */
                            for (int ch = 0; ch < 3; ++ch) {
                                Ray3f rc;
                                Color3f vc = cam->sampleRay(rc, pixelSample, apertureSample, ch);
                                vc *= scene->getIntegrator()->Li(scene, sampler.get(), rc);
                                value += vc;
                            }
                        } else {
                            Ray3f ray;
                            value = cam->sampleRay(ray, pixelSample, apertureSample);
                            value *= scene->getIntegrator()->Li(scene, sampler.get(), ray);
                        }
                        seq.insert(seq.end(), {pixelSample.x(), pixelSample.y(), value[0], value[1], value[2]});
                    }
            w.f32("seq", {nSeq, 5}, seq.data());
        }

        /* ---- plugin probes: the reference's own BSDF::eval/pdf/sample and Emitter::sample/eval/pdf
         *      answers on seeded random queries (golden vectors for the per-function parity tests).
         *      bsdf rows:    in  (wi.xyz, wo.xyz, uv.xy, sample.xy)                        = 10 floats
         *                    out (eval.rgb, pdf, weight.rgb, sampled wo.xyz, measure, pdf(sampled)) = 12 floats
         *      emitter rows: in  (ref.xyz, sample.xy)                                      = 5 floats
         *                    out (Li.rgb, wi.xyz, pdf, shadow.mint, shadow.maxt, p.xyz, eval.rgb) = 15 floats */
        if (nProbe > 0) {
            pcg32 rng(seed, 11);
            auto unit = [&]() { return Warp::squareToUniformSphere(Point2f(rng.nextFloat(), rng.nextFloat())); };
            std::vector<const BSDF *> order(bsdfs.size());
            for (auto &kv : bsdfIndex) order[kv.second] = kv.first;
            for (size_t b = 0; b < order.size(); ++b) {
                std::vector<float> in, out;
                for (size_t i = 0; i < nProbe; ++i) {
                    Vector3f wi = unit(), wo = unit();
                    if (i % 4 != 3) wi.z() = std::abs(wi.z());
                    if (i % 8 < 6) wo.z() = std::abs(wo.z());
                    Point2f uv(rng.nextFloat(), rng.nextFloat()), smp(rng.nextFloat(), rng.nextFloat());
                    BSDFQueryRecord q(wi, wo, ESolidAngle); q.uv = uv;
                    Color3f ev = order[b]->eval(q); float pdf = order[b]->pdf(q);
                    BSDFQueryRecord r(wi); r.uv = uv;
                    Color3f wgt = order[b]->sample(r, smp);
                    float pdf2 = order[b]->pdf(r);
                    in.insert(in.end(), {wi.x(), wi.y(), wi.z(), wo.x(), wo.y(), wo.z(), uv.x(), uv.y(), smp.x(), smp.y()});
                    out.insert(out.end(), {ev[0], ev[1], ev[2], pdf, wgt[0], wgt[1], wgt[2], r.wo.x(), r.wo.y(), r.wo.z(), (float) r.measure, pdf2});
                }
                w.f32("probe.bsdf." + std::to_string(b) + ".in", {nProbe, 10}, in.data());
                w.f32("probe.bsdf." + std::to_string(b) + ".out", {nProbe, 12}, out.data());
            }
            BoundingBox3f bb = scene->getBoundingBox();
            for (size_t e = 0; e < lights.size(); ++e) {
                std::vector<float> in, out;
                for (size_t i = 0; i < nProbe; ++i) {
                    Point3f ref;
                    for (int k = 0; k < 3; ++k) ref[k] = bb.min[k] + rng.nextFloat() * (bb.max[k] - bb.min[k]);
                    Point2f smp(rng.nextFloat(), rng.nextFloat());
                    EmitterQueryRecord q(ref);
                    Color3f Li = lights[e]->sample(q, smp);
                    float pdf = lights[e]->pdf(q);
                    Color3f ev = lights[e]->eval(q);
                    in.insert(in.end(), {ref.x(), ref.y(), ref.z(), smp.x(), smp.y()});
                    out.insert(out.end(), {Li[0], Li[1], Li[2], q.wi.x(), q.wi.y(), q.wi.z(), pdf, q.shadowRay.mint, q.shadowRay.maxt,
                                           q.p.x(), q.p.y(), q.p.z(), ev[0], ev[1], ev[2]});
                }
                w.f32("probe.emitter." + std::to_string(e) + ".in", {nProbe, 5}, in.data());
                w.f32("probe.emitter." + std::to_string(e) + ".out", {nProbe, 15}, out.data());
            }
        }
        w.save(out);
        cout << "nori_export: wrote " << out << " (" << w.buf.size() << " bytes, " << bvh->m_nodes.size() << " nodes, "
             << bvh->m_indices.size() << " prims, " << nRays << " rays)" << endl;
    } catch (const std::exception &e) {
        cerr << "nori_export: " << e.what() << endl;
        return 2;
    }
    return 0;
}
