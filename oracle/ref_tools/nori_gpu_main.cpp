/* nori_ref_gpu -- the reference's headless front-end (src/main_euler.cpp) with the render-thread lambda of
 * src/render.cpp:173-284 replaced by calls into libnori_gpu.so, exactly as INTEGRATION.md section 3 describes.
 *
 *   nori_ref_gpu scene.xml [--devices 0,1,...] [--chunk SPP] [--spp N] [--seed S] [--variance] [--describe]
 *
 * Everything around the loop is the reference's own code, linked unmodified from oracle/_ref/libnori_ref.a: the XML
 * parser, the NoriObject registry, OBJ loading, the SAH BVH build (Scene::activate), Integrator::preprocess,
 * ImageBlock / Bitmap and the EXR writer.  What changes is ONE function: GpuRenderThread::renderScene below keeps
 * render.cpp:135-171 (load, init m_block, output names) and render.cpp:250-284 (toBitmap, save, status) and swaps the
 * spp loop (TBB over blocks, renderBlock, m_block.put) for  flatten -> nori_gpu_upload_scene -> nori_gpu_render per spp
 * chunk -> nori_gpu_download_film INTO m_block's own memory.
 *
 * --describe prints the flattened description (scalars + a checksum per array) and exits without touching a device:
 * tests/test_reference_binding.py compares it with the committed fixture of the same scene on the CPU.
 * Integration tooling: built only into oracle/_ref/ (oracle/Makefile: ref). */
#include "gpu_binding.h"
#include <nori/block.h>
#include <nori/render.h>
#include <nori/timer.h>
#include <filesystem/path.h>
#include <iomanip>
#include <thread>
#include <unistd.h>

using namespace nori;

struct Options {
    std::vector<int> devices{0};
    uint32_t chunk = 64, spp = 0; uint64_t seed = 0; bool variance = false, describe = false;
};

static uint32_t crc32(const void *data, size_t n) {
    static uint32_t table[256]; static bool init = false;
    if (!init) { for (uint32_t i = 0; i < 256; ++i) { uint32_t c = i; for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1; table[i] = c; } init = true; }
    uint32_t c = 0xffffffffu; const uint8_t *p = (const uint8_t *) data;
    for (size_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xff] ^ (c >> 8);
    return c ^ 0xffffffffu;
}

static void describe(const nori_gpu_scene &d) {
    cout << "{\"integrator\": " << d.integrator << ", \"n_nodes\": " << d.n_nodes << ", \"n_indices\": " << d.n_indices
         << ", \"n_shapes\": " << d.n_shapes << ", \"n_bsdfs\": " << d.n_bsdfs << ", \"n_emitters\": " << d.n_emitters
         << ", \"n_images\": " << d.n_images << ", \"width\": " << d.camera.width << ", \"height\": " << d.camera.height
         << ", \"nodes_crc\": " << crc32(d.nodes, (size_t) d.n_nodes * 32) << ", \"indices_crc\": " << crc32(d.indices, (size_t) d.n_indices * 4)
         << ", \"shape_offset_crc\": " << crc32(d.shape_offset, ((size_t) d.n_shapes + 1) * 4)
         << ", \"bsdfs_crc\": " << crc32(d.bsdfs, (size_t) d.n_bsdfs * sizeof(nori_gpu_bsdf))
         << ", \"camera_crc\": " << crc32(&d.camera, sizeof(d.camera)) << ", \"filter_crc\": " << crc32(&d.filter, sizeof(d.filter))
         << ", \"medium_crc\": " << crc32(&d.medium, sizeof(d.medium)) << ", \"shapes\": [";
    for (uint32_t s = 0; s < d.n_shapes; ++s) {
        const nori_gpu_shape &p = d.shapes[s];
        cout << (s ? ", " : "") << "{\"type\": " << p.type << ", \"bsdf\": " << p.bsdf << ", \"emitter\": " << p.emitter
             << ", \"n_vertices\": " << p.n_vertices << ", \"n_triangles\": " << p.n_triangles;
        if (p.type == NORI_SHAPE_MESH)
            cout << ", \"V_crc\": " << crc32(p.V, (size_t) p.n_vertices * 12) << ", \"F_crc\": " << crc32(p.F, (size_t) p.n_triangles * 12)
                 << ", \"cdf_crc\": " << crc32(p.area_cdf, ((size_t) p.n_triangles + 1) * 4);
        cout << "}";
    }
    cout << "], \"emitters\": [";
    for (uint32_t e = 0; e < d.n_emitters; ++e)
        cout << (e ? ", " : "") << "{\"type\": " << d.emitters[e].type << ", \"shape\": " << d.emitters[e].shape << "}";
    cout << "]}" << endl;
}

class GpuRenderThread : public RenderThread {
public:
    GpuRenderThread(ImageBlock &block, const Options &opt) : RenderThread(block), m_opt(opt) {}
    std::string error;

    void renderScene(const std::string &filename) {
        /* ---- render.cpp:137-169, unchanged in substance: load, preprocess, size the film, name the outputs */
        filesystem::path path(filename);
        getFileResolver()->prepend(path.parent_path());
        gpubind::installFactoryHook();                               /* before parsing: the binding reads PropertyLists */
        NoriObject *root = loadFromXML(filename);
        if (root->getClassType() != NoriObject::EScene) { delete root; return; }
        m_scene = static_cast<Scene *>(root);
        const Camera *camera_ = m_scene->getCamera();
        m_scene->getIntegrator()->preprocess(m_scene);
        m_block.init(camera_->getOutputSize(), camera_->getReconstructionFilter());
        m_block.clear();
        std::string stem = filename;
        size_t lastdot = stem.find_last_of(".");
        if (lastdot != std::string::npos) stem.erase(lastdot, std::string::npos);
        const std::string outputName = stem + ".exr", outputNameVariance = stem + "_variance.exr";

        if (m_opt.describe) {
            nori_gpu_scene desc; gpubind::Storage keep;
            gpubind::GpuBinding::flatten(m_scene, desc, keep);
            describe(desc);
            delete m_scene; m_scene = nullptr;                       /* no render thread was started: status stays 0 */
            return;
        }

        m_render_status = 1;
        m_render_thread = std::thread([this, outputName, outputNameVariance] {
            /* ---- replaces render.cpp:174-250 (BlockGenerator, samplers, tbb::parallel_for over renderBlock, put) */
            nori_gpu_ctx *gpu = nullptr;
            auto fail = [&](const char *what) {
                error = std::string(what) + ": " + nori_gpu_last_error(gpu);
                cerr << error << endl;
                if (gpu) nori_gpu_destroy(gpu);
                delete m_scene; m_scene = nullptr; m_render_status = 3;
            };
            const int rc = m_opt.devices.size() > 1 ? nori_gpu_init_multi(m_opt.devices.data(), (int) m_opt.devices.size(), &gpu)
                                                    : nori_gpu_init(m_opt.devices[0], &gpu);
            if (rc) return fail("nori_gpu_init");
            nori_gpu_scene desc; gpubind::Storage keep;
            try { gpubind::GpuBinding::flatten(m_scene, desc, keep); }
            catch (const std::exception &e) { error = e.what(); cerr << error << endl; nori_gpu_destroy(gpu); delete m_scene; m_scene = nullptr; m_render_status = 3; return; }
            if (nori_gpu_upload_scene(gpu, &desc)) return fail("nori_gpu_upload_scene");
            if (m_opt.variance && nori_gpu_set_option(gpu, "variance", 1)) return fail("nori_gpu_set_option(variance)");

            cout << "Rendering .. "; cout.flush();
            Timer timer;
            const uint32_t numSamples = m_opt.spp ? m_opt.spp : (uint32_t) m_scene->getSampler()->getSampleCount();
            for (uint32_t k = 0; k < numSamples; k += m_opt.chunk) {
                m_progress = k / float(numSamples);
                if (m_render_status == 2) break;                     /* the reference's cancel point (render.cpp:196-197) */
                if (nori_gpu_render(gpu, k, std::min(m_opt.chunk, numSamples - k), m_opt.seed)) return fail("nori_gpu_render");
            }
            cout << "done. (took " << timer.elapsedString() << ")" << endl;

            /* ---- render.cpp:254-261: the device film IS ImageBlock's memory image (block.h:48) */
            m_block.lock();
            static_assert(sizeof(Color4f) == 16, "Color4f is (r,g,b,w)");
            if (nori_gpu_download_film(gpu, reinterpret_cast<float *>(m_block.data()))) { m_block.unlock(); return fail("nori_gpu_download_film"); }
            std::unique_ptr<Bitmap> bitmap(m_block.toBitmap());
            m_block.unlock();
            bitmap->save(outputName);

            /* ---- render.cpp:263-278: the per-pixel variance estimate */
            if (m_opt.variance) {
                Bitmap var(m_scene->getCamera()->getOutputSize());
                static_assert(sizeof(Color3f) == 12, "Color3f is (r,g,b)");
                if (nori_gpu_download_variance(gpu, reinterpret_cast<float *>(var.data()))) return fail("nori_gpu_download_variance");
                var.save(outputNameVariance);
            }
            nori_gpu_stats st;
            if (!nori_gpu_get_stats(gpu, &st))
                cout << "gpu: " << st.samples << " samples, " << st.rays << " rays, " << st.devices << " device(s), last chunk " << st.render_ms << " ms" << endl;
            nori_gpu_destroy(gpu);
            delete m_scene; m_scene = nullptr;
            m_render_status = 3;
        });
    }
private:
    Options m_opt;
};

int main(int argc, char **argv) {
    Options opt; std::string filename;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "--devices" && i + 1 < argc) {
            opt.devices.clear();
            std::string list = argv[++i]; size_t pos = 0;
            while (pos <= list.size()) { size_t c = list.find(',', pos); if (c == std::string::npos) c = list.size(); opt.devices.push_back(atoi(list.substr(pos, c - pos).c_str())); pos = c + 1; }
        } else if (a == "--chunk" && i + 1 < argc) opt.chunk = (uint32_t) std::max(1, atoi(argv[++i]));
        else if (a == "--spp" && i + 1 < argc) opt.spp = (uint32_t) atoi(argv[++i]);
        else if (a == "--seed" && i + 1 < argc) opt.seed = (uint64_t) atoll(argv[++i]);
        else if (a == "--variance") opt.variance = true;
        else if (a == "--describe") opt.describe = true;
        else filename = a;
    }
    if (filename.empty()) { cerr << "usage: nori_ref_gpu scene.xml [--devices 0,1,..] [--chunk SPP] [--spp N] [--seed S] [--variance] [--describe]" << endl; return 1; }
    try {
        /* main_euler.cpp:35-36: a dummy block, (re)initialised by renderScene */
        ImageBlock block(Vector2i(720, 720), nullptr);
        GpuRenderThread thread(block, opt);
        filesystem::path path(filename);
        if (path.extension() != "xml") { cerr << "expected an .xml scene" << endl; return 1; }
        thread.renderScene(filename);
        while (thread.isBusy()) usleep(20000);                       /* main_euler.cpp:48-52 (progress polling); isBusy joins when done */
        return thread.error.empty() ? 0 : 2;
    } catch (const std::exception &e) {
        cerr << "Fatal error: " << e.what() << endl;
        return 1;
    }
}
