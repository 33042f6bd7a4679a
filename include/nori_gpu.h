/* nori_gpu.h -- the C-ABI boundary of the B200-native rendering hot path.
 *
 * This header is the drop-in seam for the reference renderer (francois141/nori-ray-tracer).
 * The reference has no C API: its seam is the body of the render-thread lambda in
 * src/render.cpp:173-284 (spp-major loop -> tbb::parallel_for over 32x32 blocks -> renderBlock()
 * src/render.cpp:80-133 -> ImageBlock::put / merge), consuming a `const Scene*`
 * (include/nori/scene.h:44-115) and producing the full-image `ImageBlock m_block`
 * (include/nori/block.h:48, row-major Color4f (r,g,b,weight), (H+2b) x (W+2b)).
 * Everything below replaces exactly that: plain pointers and sizes in, the same film array out.
 * No C++ types, no torch types, no exceptions cross this boundary (status code + last_error).
 *
 * Every struct is plain-old-data with natural alignment; all arrays are HOST pointers owned by the
 * caller and are copied during nori_gpu_upload_scene().  Layouts follow the reference's own
 * containers (cited per field) so that the reference-side exporter is a memcpy.
 */
#ifndef NORI_GPU_H
#define NORI_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NORI_GPU_ABI_VERSION 3
#define NORI_FILTER_RESOLUTION 32          /* include/nori/rfilter.h:25 */
#define NORI_BLOCK_SIZE 32                 /* include/nori/block.h:30   */

typedef struct nori_gpu_ctx nori_gpu_ctx;  /* opaque; owns all device memory */

/* ---- enumerations (values are part of the ABI) ------------------------------------------- */
enum { NORI_SHAPE_MESH = 0, NORI_SHAPE_SPHERE = 1, NORI_SHAPE_PERLIN = 2 };   /* src/mesh.cpp, src/sphere.cpp, src/perlinnoise.cpp */
enum { NORI_BSDF_DIFFUSE = 0, NORI_BSDF_MIRROR = 1, NORI_BSDF_DIELECTRIC = 2,
       NORI_BSDF_MICROFACET = 3, NORI_BSDF_DISNEY = 4, NORI_BSDF_COUNT = 5 };
enum { NORI_EMITTER_AREA = 0, NORI_EMITTER_POINT = 1, NORI_EMITTER_SPOT = 2, NORI_EMITTER_ENVMAP = 3 };
enum { NORI_CAMERA_PERSPECTIVE = 0, NORI_CAMERA_THINLENS = 1, NORI_CAMERA_ADVANCED = 2 };   /* src/advancedCamera.cpp */
enum { NORI_INTEGRATOR_NORMALS = 0, NORI_INTEGRATOR_PATH_MIS = 1, NORI_INTEGRATOR_PATH_MATS = 2,
       NORI_INTEGRATOR_DIRECT_EMS = 3, NORI_INTEGRATOR_DIRECT_MATS = 4, NORI_INTEGRATOR_DIRECT_MIS = 5,
       NORI_INTEGRATOR_DIRECT = 6, NORI_INTEGRATOR_AV = 7, NORI_INTEGRATOR_VOLUMETRIC = 8 };
enum { NORI_TEXTURE_CONSTANT = 0, NORI_TEXTURE_CHECKERBOARD = 1, NORI_TEXTURE_IMAGE = 2 };      /* src/imagetexture.cpp */
enum { NORI_WRAP_REPEAT = 0, NORI_WRAP_CLAMP = 1 };                      /* common.h:273-283 */

/* ---- 8-bit RGB images: ImageTexture::m_data / NormalMap::m_data as stbi_load(.., STBI_rgb) returns
 *      them (imagetexture.cpp:73-87, normalmap.cpp:73-87): texel (x, y) at rgb[(x + width*y)*3] ------ */
typedef struct {
    int32_t width, height;
    int32_t wrap;               /* NORI_WRAP_*                                               */
    int32_t reserved;
    const uint8_t *rgb;
} nori_gpu_image;

/* ---- BVH: the reference's 32-byte node, verbatim (include/nori/bvh.h:127-164) ------------ *
 * data[0]: bit 0 = leaf flag, bits 1..31 = leaf size (leaf) or split axis (inner, unused)
 * data[1]: leaf: first index into `indices`; inner: index of the right child (left = self+1)   */
typedef struct {
    uint32_t data[2];
    float    bmin[3];
    float    bmax[3];
} nori_gpu_bvh_node;

/* ---- shapes (include/nori/mesh.h:121-124; src/sphere.cpp:30-37) -------------------------- */
typedef struct {
    int32_t  type;              /* NORI_SHAPE_*                                             */
    int32_t  bsdf;              /* index into scene.bsdfs                                    */
    int32_t  emitter;           /* index into scene.emitters, or -1 (Shape::isEmitter)       */
    uint32_t n_vertices;        /* mesh: columns of m_V                                      */
    uint32_t n_triangles;       /* mesh: columns of m_F; sphere: 1 primitive                 */
    int32_t  normal_map;        /* 1 + index into scene.images of the mesh's NormalMap (shape.cpp:59-66), 0 = none */
    const float    *V;          /* 3*n_vertices, column-major m_V: vertex i = V[3i..3i+2]    */
    const float    *N;          /* 3*n_vertices or NULL (m_N)                                */
    const float    *UV;         /* 2*n_vertices or NULL (m_UV)                               */
    const uint32_t *F;          /* 3*n_triangles, column-major m_F                           */
    const float    *area_cdf;   /* n_triangles+1 floats: DiscretePDF::m_cdf (dpdf.h:194)     */
    float    area_normalization;/* DiscretePDF::getNormalization() = 1/total area            */
    float    center[3];         /* sphere / perlin sphere                                    */
    float    radius;            /* sphere / perlin sphere                                    */
    float    perlin_height;     /* perlinnoise.cpp:15 "height": base frequency is 1/height   */
    float    perlin_scale;      /* perlinnoise.cpp:16 "scale": radius += scale * clamp(noise, 0, 1) */
    uint32_t reserved2;
} nori_gpu_shape;

/* ---- BSDFs (src/diffuse.cpp, mirror.cpp, dielectric.cpp, microfacet.cpp, disney.cpp) ----- */
typedef struct {
    int32_t type;               /* NORI_BSDF_*                                               */
    int32_t albedo_texture;     /* NORI_TEXTURE_* for the diffuse albedo                     */
    float   albedo[3];          /* diffuse: constant albedo / checkerboard value1            */
    float   albedo2[3];         /* checkerboard value2 (src/checkerboard.cpp:31-37)          */
    float   tex_scale[2];       /* checkerboard scale                                        */
    float   tex_delta[2];       /* checkerboard delta                                        */
    float   intIOR, extIOR;     /* dielectric.cpp:26-29, microfacet.cpp:31-34                */
    float   alpha;              /* microfacet.cpp:28 ; disney: m_alpha (disney.cpp:59)       */
    float   kd[3];              /* microfacet.cpp:37                                         */
    float   ks;                 /* microfacet.cpp:48 : 1 - max(kd)                           */
    float   baseColor[3];       /* disney.cpp:57                                             */
    float   metallic, specular, roughness, sheen, sheenTint, specularTint; /* disney.cpp:50-55 */
    int32_t albedo_image;       /* NORI_TEXTURE_IMAGE: index into scene.images               */
    float   reserved;
} nori_gpu_bsdf;

/* ---- emitters (src/arealight.cpp, pointlight.cpp, spotlight.cpp, envmap.cpp) ------------- */
typedef struct {
    int32_t type;               /* NORI_EMITTER_*                                            */
    int32_t shape;              /* area / envmap: index of the shape it is attached to, else -1 */
    float   radiance[3];        /* area: m_radiance; point: power; spot: "color"             */
    float   position[3];        /* point / spot                                              */
    float   direction[3];       /* spot (normalised, spotlight.cpp:14)                       */
    float   cosFalloffStart;    /* spotlight.cpp:16                                          */
    float   cosTotalWidth;      /* spotlight.cpp:17                                          */
    float   weight;             /* envmap m_weight (envmap.cpp:14)                           */
    int32_t env_rows;           /* envmap "m_width"  = bitmap rows (envmap.cpp:33, sic)      */
    int32_t env_cols;           /* envmap "m_height" = bitmap cols (envmap.cpp:32, sic)      */
    const float *env_image;     /* rows*cols*3 row-major RGB (Bitmap, bitmap.h:32)           */
    const float *env_pdf;       /* rows*cols      row-major  m_pdf       (envmap.cpp:36)     */
    const float *env_cdf;       /* rows*(cols+1)  row-major  m_cdf       (envmap.cpp:37)     */
    const float *env_pmarginal; /* rows                     m_pmarginal  (envmap.cpp:38)     */
    const float *env_cmarginal; /* rows+1                   m_cmarginal  (envmap.cpp:39)     */
} nori_gpu_emitter;

/* ---- camera (src/perspective.cpp:53-112, src/thinlens.cpp:66-171) ------------------------ */
typedef struct {
    int32_t type;               /* NORI_CAMERA_*                                             */
    int32_t width, height;      /* m_outputSize                                              */
    float   sampleToCamera[16]; /* row-major 4x4 (m_sampleToCamera.getMatrix())              */
    float   cameraToWorld[16];  /* row-major 4x4 (m_cameraToWorld.getMatrix())               */
    float   invOutputSize[2];
    float   nearClip, farClip;
    float   lensRadius, focalDistance;   /* thinlens.cpp:49-50                               */
    float   distortion[2];      /* advancedCamera.cpp:54: radial distortion coefficients     */
    float   chromatic[3];       /* advancedCamera.cpp:55: per-channel chromatic aberration strength; non-zero =>
                                   three camera paths per sample, one per colour channel (render.cpp:106-121) */
    float   reserved;
} nori_gpu_camera;

/* ---- reconstruction filter, pre-tabulated by the host exactly as ImageBlock::init does
 *      (src/block.cpp:54-64): table[i] = filter->eval(radius*i/32), table[32] = 0 ---------- */
typedef struct {
    float radius;
    float table[NORI_FILTER_RESOLUTION + 1];
} nori_gpu_filter;

/* ---- homogeneous medium (src/medium.cpp:8-20) -------------------------------------------- */
typedef struct {
    int32_t present;
    float   sigma_a[3], sigma_s[3];
    float   bounds_min[3], bounds_max[3];
} nori_gpu_medium;

/* ---- the whole scene --------------------------------------------------------------------- */
typedef struct {
    uint32_t abi_version;       /* NORI_GPU_ABI_VERSION                                      */
    int32_t  integrator;        /* NORI_INTEGRATOR_*                                         */
    float    av_length;         /* averagevisibility.cpp:13 "length"                         */
    uint32_t n_nodes;
    uint32_t n_indices;         /* == total primitive count                                  */
    uint32_t n_shapes, n_bsdfs, n_emitters;
    const nori_gpu_bvh_node *nodes;        /* BVH::m_nodes (bvh.h:168)                       */
    const uint32_t          *indices;      /* BVH::m_indices (bvh.h:169)                     */
    const uint32_t          *shape_offset; /* BVH::m_shapeOffset, n_shapes+1 entries (bvh.h:167) */
    const nori_gpu_shape    *shapes;       /* in BVH::m_shapes order                         */
    const nori_gpu_bsdf     *bsdfs;
    const nori_gpu_emitter  *emitters;     /* in Scene::m_emitters order (scene.cpp:63-76)   */
    nori_gpu_camera camera;
    nori_gpu_filter filter;
    nori_gpu_medium medium;
    uint32_t n_images;
    uint32_t reserved;
    const nori_gpu_image *images;          /* image textures and normal maps                 */
} nori_gpu_scene;

/* ---- test hooks ---------------------------------------------------------------------------- */
typedef struct {                /* Ray3f (include/nori/ray.h:43-47) without dRcp              */
    float o[3]; float mint;
    float d[3]; float maxt;
} nori_gpu_ray;

typedef struct {                /* what BVH::rayIntersect decides before setHitInformation     */
    float    t;                 /* its.t, +inf if no hit                                      */
    float    u, v;              /* barycentric (u,v) of the winning triangle; 0 for spheres   */
    uint32_t shape;             /* index into scene.shapes, 0xffffffff if no hit              */
    uint32_t prim;              /* primitive index local to the shape (findShape, bvh.h:105)  */
    uint32_t nodes_visited;     /* BVHNode fetches (bvh.cpp:421)                              */
    uint32_t prims_tested;      /* Shape::rayIntersect calls (bvh.cpp:440)                    */
    uint32_t reserved;
} nori_gpu_hit;

typedef struct {
    uint64_t samples;           /* camera paths started (render.cpp:98-130 iterations)        */
    uint64_t rays;              /* BVH::rayIntersect queries, closest + shadow                */
    uint64_t shadow_rays;       /* the any-hit subset                                         */
    uint64_t nodes_visited;     /* only counted when stats collection is on                   */
    uint64_t prims_tested;      /*   "                                                        */
    uint64_t invalid_samples;   /* dropped like block.cpp:94-98                               */
    uint64_t iterations;        /* wavefront iterations executed                              */
    uint64_t kernel_launches;   /* CUDA kernels launched by render / trace calls              */
    double   render_ms;         /* device time of the last nori_gpu_render (CUDA events)      */
    double   trace_ms;          /* device time of the last nori_gpu_trace kernel              */
    uint64_t max_stack_depth;   /* with "stats" on: deepest per-ray traversal stack of the large-scene kernels (entries) */
    uint64_t guard_retraces;    /* with "stats" on: near-child-first closest-hit queries in which a second candidate appeared
                                   within 2^-11 of the best hit and that were answered again in the reference's child order */
    double   reduce_ms;         /* multi-device contexts: device time of the film sum of the last nori_gpu_render (part of render_ms) */
    uint64_t devices;           /* devices the context renders on (nori_gpu_init_multi), 1 otherwise */
} nori_gpu_stats;

/* per-kernel-class accounting of nori_gpu_render since the last reset (ms only with option
 * "kernel_timing" = 1: every launch is bracketed by CUDA events on the context's stream) */
enum { NORI_K_GENERATE = 0, NORI_K_EXTEND = 1, NORI_K_SHADE = 2, NORI_K_SHADOW = 3, NORI_K_FILM = 4,
       NORI_K_SINGLE = 5, NORI_K_COUNT = 6 };
typedef struct {
    double   ms;                /* summed device time of the launches of this class            */
    uint64_t launches;
    uint64_t rays;              /* BVH queries issued by this class                            */
    uint64_t nodes_visited;     /* with option "stats" = 1                                     */
    uint64_t prims_tested;      /*   "                                                         */
} nori_gpu_kernel_stats;

/* ---- entry points ---------------------------------------------------------------------------
 * All return 0 on success, non-zero on error (then nori_gpu_last_error() explains).
 * One host thread per ctx; a ctx is bound to one CUDA device.                                  */

/* Create a context on CUDA device `device`. Replaces nothing in the reference (there is no device). */
int nori_gpu_init(int device, nori_gpu_ctx **out);
/* One context over `n` devices of one node ("multi-GPU inside render").  The host keeps calling the same
 * entry points: upload_scene / set_option / clear_film act on every device (the scene is replicated); nori_gpu_render
 * shards the sample indices [spp_begin, spp_begin + spp_count) over the devices -- device g renders [g*count/n,
 * (g+1)*count/n), disjoint pcg32 initstate ranges, one host thread per device -- and sums the accumulation buffers onto
 * devices[0] with one kernel that reads the peers' buffers in place over NVLink (cudaDeviceEnablePeerAccess; staged
 * with cudaMemcpyPeerAsync where peer access is unavailable).  Film / resolve / stats calls then answer for the whole
 * render from devices[0].  render.cpp's block loop (render.cpp:173-284) has no counterpart: TBB shares one image among
 * the cores of one host.  The variance statistic and the test hooks (trace, probes, render_samples) stay
 * single-device: they run on devices[0]; set_option("variance", 1) is refused.
 * With torch.distributed (one process per GPU) use nori_gpu_init per rank and reduce nori_gpu_film_device_ptr instead. */
int nori_gpu_init_multi(const int *devices, int n, nori_gpu_ctx **out);
int nori_gpu_device_count(const nori_gpu_ctx *ctx);
void nori_gpu_destroy(nori_gpu_ctx *ctx);
const char *nori_gpu_last_error(const nori_gpu_ctx *ctx);   /* ctx may be NULL: last init error */

/* Copy + flatten the scene to the device.  Replaces Scene ownership by the render thread
 * (render.cpp:147-156: loadFromXML -> m_block.init(size, filter)); also (re)allocates and clears
 * the film. */
int nori_gpu_upload_scene(nori_gpu_ctx *ctx, const nori_gpu_scene *scene);

/* Tunables: "pool" (resident path slots), "results_mb" (MiB of the per-sample buffer: bounds the sample layers of one film batch),
 * "stats" (1: count node visits / primitive tests), "kernel_timing" (1: CUDA events around every
 * launch), "megakernel" (1: one thread per sample for every integrator), "poll" (wavefront iterations between two looks at the
 * device counters), "variance" (1: the running-mean variance statistic of render.cpp:238-278, see nori_gpu_download_variance),
 * "flush_l2" (bench only: overwrite that many MiB to evict L2), "order" (0: the reference's child
 * order, left child first, bvh.cpp:430-433 -- node-visit / primitive-test counters equal the
 * reference's; 1: the child on the ray's side of the node's split axis first -- same hits, same
 * primitive ids incl. the reference's tie rule, fewer node visits; 2 (default): 1 when rendering
 * scenes with more than 4096 primitives, 0 otherwise and always for nori_gpu_trace),
 * "wide" (1: with the near-first order, the large-scene kernels walk a 4-wide layout -- every inner
 * node merged with its inner children, the reference's boxes unchanged -- instead of child-box pairs),
 * "area_only" (default 1: scenes whose emitters are all
 * area lights are shaded by kernels compiled without the point / spot / environment-map code),
 * "emitter_sort" (path_mis, emitters of several types: shade from
 * (material, emitter type)-sorted queues; 0 off, 1 (default) when an environment map is among them, 2 always),
 * "drain" (finish a batch with one
 * thread per remaining path once at most this many paths are alive and none is left to start; default
 * 32768, 0 = never),
 * "shadow_pass" (path_mis NEE rays: 1 their own state-machine pass, 2 inside the shade kernel,
 * 0 (default) by scene size), "traversal" (1 plain per-lane loops, 2 warp state machine, 0 by size),
 * "trace_kernel" (test hook routing of nori_gpu_trace: 0 (default) one thread per ray over the reference nodes in the
 * order "order" selects; 2: the rays are loaded into path-pool slots and answered by the kernels that render large
 * scenes -- the warp-state-machine closest-hit / any-hit kernels with the configured "order" / "wide" layout; per-ray
 * counters are then 0, the totals are in nori_gpu_get_kernel_stats, and any-hit rays start at Epsilon like the NEE
 * rays those kernels trace, arealight.cpp:56),
 * "drain_mode" (0 (default): the tail by one warp per remaining path, 1: one thread per path),
 * "film_sep" (default 1: radius-2 filters use the film kernel with per-sample tabulated row / column weights; 0: the generic one),
 * "film_tma" (default 1: that kernel's sample tiles are staged by the TMA unit; 0: per-thread loads),
 * "l2_window" (MiB of the 4-wide records held by an L2 access-policy window; default 0),
 * "reset_options" (every scheduling option back to its default),
 * "wavefronts" (1..4, default 2: the sample layers of a batch are split between that many independent wavefronts --
 * own part of the pool, own counters, own stream -- whose kernels overlap on the device; the samples, the film and the
 * counters do not depend on it; "kernel_timing" renders with one). */
int nori_gpu_set_option(nori_gpu_ctx *ctx, const char *name, int64_t value);

/* Render sample indices [spp_begin, spp_begin+spp_count) for every pixel and ACCUMULATE them into
 * the film.  Replaces `for k in spp: tbb::parallel_for(blocks, renderBlock + m_block.put)`
 * (render.cpp:194-233).  Callable repeatedly (progress / cancel between calls, render.cpp:195-197).
 * Path (pixel p=(x,y), sample k) uses pcg32.seed(initstate = seed + k, initseq = y*W + x). */
int nori_gpu_render(nori_gpu_ctx *ctx, uint32_t spp_begin, uint32_t spp_count, uint64_t seed);

/* Same paths, but return the per-sample radiance instead of splatting: out[(k*H + y)*W + x] =
 * (r,g,b,valid).  The test hook behind the t-test fixtures (ttest.cpp:151-193). */
int nori_gpu_render_samples(nori_gpu_ctx *ctx, uint32_t spp_begin, uint32_t spp_count, uint64_t seed,
                            float *out_rgba);

int nori_gpu_clear_film(nori_gpu_ctx *ctx);                      /* ImageBlock::clear (render.cpp:155) */
/* (H+2b)*(W+2b)*4 floats, row-major (r,g,b,w) -- the exact memory image of ImageBlock m_block. */
int nori_gpu_download_film(nori_gpu_ctx *ctx, float *rgbaw);
int nori_gpu_upload_film(nori_gpu_ctx *ctx, const float *rgbaw); /* resume / multi-GPU merge      */
/* Device address of that same array, for zero-copy collectives (torch.distributed / NCCL). */
int nori_gpu_film_device_ptr(nori_gpu_ctx *ctx, void **dptr, uint64_t *n_floats);
int nori_gpu_film_dims(const nori_gpu_ctx *ctx, int32_t *rows, int32_t *cols, int32_t *border);
/* ImageBlock::toBitmap (block.cpp:76-82): H*W*3 floats = rgb / w (0 where w == 0). */
int nori_gpu_resolve(nori_gpu_ctx *ctx, float *rgb);

/* The reference's second output, <scene>_variance.exr (render.cpp:190-192,225,238-247,263-278): per
 * pixel mean_k(m_k^2) - (mean_k m_k)^2 over the running means m_k after each spp pass.
 * Enable with nori_gpu_set_option(ctx, "variance", 1) after upload_scene and before rendering;
 * clear_film resets it.  rgb: H*W*3 floats. */
int nori_gpu_download_variance(nori_gpu_ctx *ctx, float *rgb);

/* BVH::rayIntersect on a caller-supplied ray batch (bvh.cpp:404-462), shadow != 0 => any-hit.
 * For shadow rays only `t` (0 = occluded, +inf = free) and the counters are meaningful. */
int nori_gpu_trace(nori_gpu_ctx *ctx, const nori_gpu_ray *rays, uint64_t n, int shadow, nori_gpu_hit *out);

/* BSDF::eval/pdf/sample (bsdf.h:63-110) and Emitter::sample/pdf/eval (emitter.h:64-105) of one scene
 * plugin on caller-supplied queries.  Row layouts (floats):
 *   bsdf:    in  wi.xyz wo.xyz uv.xy sample.xy (10)   out eval.rgb pdf weight.rgb wo'.xyz measure pdf(wo') (12)
 *   emitter: in  ref.xyz sample.xy (5)                out Li.rgb wi.xyz pdf shadow.mint shadow.maxt p.xyz eval.rgb (15) */
int nori_gpu_probe_bsdf(nori_gpu_ctx *ctx, uint32_t bsdf, uint64_t n, const float *in, float *out);
int nori_gpu_probe_emitter(nori_gpu_ctx *ctx, uint32_t emitter, uint64_t n, const float *in, float *out);

/* n floats of pcg32(initstate, initseq).nextFloat() generated on the device (pcg32.h:51-110). */
int nori_gpu_pcg32(nori_gpu_ctx *ctx, uint64_t initstate, uint64_t initseq, uint64_t n, float *out);
/* n raw nextUInt() outputs, for the published pcg32-demo known answers. */
int nori_gpu_pcg32_uint(nori_gpu_ctx *ctx, uint64_t initstate, uint64_t initseq, uint64_t n, uint32_t *out);

/* ---- host-side helpers (no device needed) -------------------------------------------------------
 * For scenes that do not come from the reference's loader (synthetic benchmark scenes, host-authored
 * scenes).  nori_gpu_build_bvh restates the reference's SAH builder (bvh.cpp:54-382: 16 centroid bins
 * along the largest axis, serial sort-and-sweep below 32 primitives, node compaction) with the chunk
 * order of a single-threaded run, so the tree is one the reference itself could have produced.
 * nodes_out: capacity 2 * (total primitives); indices_out: total primitives; shape_offset_out: n_shapes+1. */
int nori_gpu_build_bvh(const nori_gpu_shape *shapes, uint32_t n_shapes, nori_gpu_bvh_node *nodes_out,
                       uint32_t *indices_out, uint32_t *shape_offset_out, uint32_t *n_nodes_out, int threads);
/* The same job on the GPU, for scenes whose host build time matters: a linear BVH (Morton codes,
 * radix sort, Karras' parallel radix tree, bottom-up box fitting; subtrees of at most leaf_size primitives
 * become one leaf) written in the SAME format -- reference nodes in depth-first order, left child behind its
 * parent, split axis recorded -- so it can be uploaded, traversed and checked like a reference tree.  An
 * alternative, not a replacement: the tree differs from the reference's, so exact ties between primitives may
 * resolve differently (traverse.cuh).  Same output capacities as nori_gpu_build_bvh; build_ms_out (optional)
 * receives the device time of the build kernels. */
int nori_gpu_build_bvh_device(int device, const nori_gpu_shape *shapes, uint32_t n_shapes, nori_gpu_bvh_node *nodes_out,
                              uint32_t *indices_out, uint32_t *shape_offset_out, uint32_t *n_nodes_out,
                              uint32_t leaf_size, float *build_ms_out);
/* Mesh::activate (mesh.cpp:30-38): per-triangle area CDF (n_triangles+1 floats) and 1/total area. */
int nori_gpu_mesh_area_cdf(const float *V, const uint32_t *F, uint32_t n_triangles, float *cdf_out,
                           float *normalization_out);

/* Test hook: the 4-wide node records nori_gpu_upload_scene derives from a reference-format tree for its
 * large-scene kernels (option "wide").  records_out: 32 words per record -- slot k = words 8k..8k+7 =
 * {bmin[3], ref}{bmax[3], 0}; ref: bit 31 = leaf (size in bits 30..25, first primitive in 24..0; 0x80000000 =
 * unused slot), else the index of the child record; record 0 is the root.  capacity in records (n_nodes / 2 + 1
 * always suffices).  *n_records_out = 0 when the layout is not built for this tree (root is a leaf, a leaf with
 * more than 63 primitives, more than 2^25 primitives, or a ray's stack could exceed its 96 entries). */
int nori_gpu_wide_layout(const nori_gpu_bvh_node *nodes, uint32_t n_nodes, uint32_t n_indices, uint32_t *records_out,
                         uint32_t capacity, uint32_t *n_records_out);

/* sizeof() of every ABI struct as compiled into the library (binding self-check); returns the count. */
/* Arithmetic self-test (test hook): the library's slow-path-free IEEE sequences -- division and square root for
 * normal-range operands (camera rays), the reciprocal of the triangle determinant (mesh.cpp:93) -- against the
 * compiler's correctly rounded operations on n pseudo-random operand pairs with exponents in [-60, 60].
 * mismatch[0..2] = operands whose division / square root / reciprocal differs in any bit (all must be 0). */
int nori_gpu_selftest(nori_gpu_ctx *ctx, uint64_t n, uint64_t *mismatch /* [3] */);

int nori_gpu_abi_sizes(uint32_t *out, int n);

int nori_gpu_get_stats(nori_gpu_ctx *ctx, nori_gpu_stats *out);
int nori_gpu_get_kernel_stats(nori_gpu_ctx *ctx, nori_gpu_kernel_stats *out /* [NORI_K_COUNT] */);
int nori_gpu_reset_stats(nori_gpu_ctx *ctx);
int nori_gpu_synchronize(nori_gpu_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* NORI_GPU_H */
