"""ctypes mirror of include/nori_gpu.h (the C-ABI boundary).  Field order, types and enum values
must match the header exactly; tests/test_abi.py checks every sizeof against the compiled library."""
import ctypes as C

ABI_VERSION = 3
FILTER_RESOLUTION = 32
BLOCK_SIZE = 32

SHAPE_MESH, SHAPE_SPHERE, SHAPE_PERLIN = 0, 1, 2
BSDF_DIFFUSE, BSDF_MIRROR, BSDF_DIELECTRIC, BSDF_MICROFACET, BSDF_DISNEY = range(5)
EMITTER_AREA, EMITTER_POINT, EMITTER_SPOT, EMITTER_ENVMAP = range(4)
CAMERA_PERSPECTIVE, CAMERA_THINLENS, CAMERA_ADVANCED = 0, 1, 2
(INTEGRATOR_NORMALS, INTEGRATOR_PATH_MIS, INTEGRATOR_PATH_MATS, INTEGRATOR_DIRECT_EMS,
 INTEGRATOR_DIRECT_MATS, INTEGRATOR_DIRECT_MIS, INTEGRATOR_DIRECT, INTEGRATOR_AV,
 INTEGRATOR_VOLUMETRIC) = range(9)
TEXTURE_CONSTANT, TEXTURE_CHECKERBOARD, TEXTURE_IMAGE = 0, 1, 2
WRAP_REPEAT, WRAP_CLAMP = 0, 1

INTEGRATOR_NAMES = {
    "normals": INTEGRATOR_NORMALS, "path_mis": INTEGRATOR_PATH_MIS, "path_mats": INTEGRATOR_PATH_MATS,
    "direct_ems": INTEGRATOR_DIRECT_EMS, "direct_mats": INTEGRATOR_DIRECT_MATS,
    "direct_mis": INTEGRATOR_DIRECT_MIS, "direct": INTEGRATOR_DIRECT, "av": INTEGRATOR_AV,
    "volumetric": INTEGRATOR_VOLUMETRIC,
}

f32, u32, i32, u64 = C.c_float, C.c_uint32, C.c_int32, C.c_uint64
pf32, pu32 = C.POINTER(C.c_float), C.POINTER(C.c_uint32)


class BvhNode(C.Structure):
    _fields_ = [("data", u32 * 2), ("bmin", f32 * 3), ("bmax", f32 * 3)]


class Shape(C.Structure):
    _fields_ = [("type", i32), ("bsdf", i32), ("emitter", i32), ("n_vertices", u32),
                ("n_triangles", u32), ("normal_map", i32),
                ("V", pf32), ("N", pf32), ("UV", pf32), ("F", pu32), ("area_cdf", pf32),
                ("area_normalization", f32), ("center", f32 * 3), ("radius", f32),
                ("perlin_height", f32), ("perlin_scale", f32), ("reserved2", u32)]


class Bsdf(C.Structure):
    _fields_ = [("type", i32), ("albedo_texture", i32), ("albedo", f32 * 3), ("albedo2", f32 * 3),
                ("tex_scale", f32 * 2), ("tex_delta", f32 * 2), ("intIOR", f32), ("extIOR", f32),
                ("alpha", f32), ("kd", f32 * 3), ("ks", f32), ("baseColor", f32 * 3),
                ("metallic", f32), ("specular", f32), ("roughness", f32), ("sheen", f32),
                ("sheenTint", f32), ("specularTint", f32), ("albedo_image", i32), ("reserved", f32)]


class Emitter(C.Structure):
    _fields_ = [("type", i32), ("shape", i32), ("radiance", f32 * 3), ("position", f32 * 3),
                ("direction", f32 * 3), ("cosFalloffStart", f32), ("cosTotalWidth", f32),
                ("weight", f32), ("env_rows", i32), ("env_cols", i32),
                ("env_image", pf32), ("env_pdf", pf32), ("env_cdf", pf32),
                ("env_pmarginal", pf32), ("env_cmarginal", pf32)]


class Camera(C.Structure):
    _fields_ = [("type", i32), ("width", i32), ("height", i32), ("sampleToCamera", f32 * 16),
                ("cameraToWorld", f32 * 16), ("invOutputSize", f32 * 2), ("nearClip", f32),
                ("farClip", f32), ("lensRadius", f32), ("focalDistance", f32),
                ("distortion", f32 * 2), ("chromatic", f32 * 3), ("reserved", f32)]


class Image(C.Structure):
    _fields_ = [("width", i32), ("height", i32), ("wrap", i32), ("reserved", i32), ("rgb", C.POINTER(C.c_uint8))]


class Filter(C.Structure):
    _fields_ = [("radius", f32), ("table", f32 * (FILTER_RESOLUTION + 1))]


class Medium(C.Structure):
    _fields_ = [("present", i32), ("sigma_a", f32 * 3), ("sigma_s", f32 * 3),
                ("bounds_min", f32 * 3), ("bounds_max", f32 * 3)]


class Scene(C.Structure):
    _fields_ = [("abi_version", u32), ("integrator", i32), ("av_length", f32), ("n_nodes", u32),
                ("n_indices", u32), ("n_shapes", u32), ("n_bsdfs", u32), ("n_emitters", u32),
                ("nodes", C.POINTER(BvhNode)), ("indices", pu32), ("shape_offset", pu32),
                ("shapes", C.POINTER(Shape)), ("bsdfs", C.POINTER(Bsdf)),
                ("emitters", C.POINTER(Emitter)),
                ("camera", Camera), ("filter", Filter), ("medium", Medium),
                ("n_images", u32), ("reserved", u32), ("images", C.POINTER(Image))]


class Ray(C.Structure):
    _fields_ = [("o", f32 * 3), ("mint", f32), ("d", f32 * 3), ("maxt", f32)]


class Hit(C.Structure):
    _fields_ = [("t", f32), ("u", f32), ("v", f32), ("shape", u32), ("prim", u32),
                ("nodes_visited", u32), ("prims_tested", u32), ("reserved", u32)]


class Stats(C.Structure):
    _fields_ = [("samples", u64), ("rays", u64), ("shadow_rays", u64), ("nodes_visited", u64),
                ("prims_tested", u64), ("invalid_samples", u64), ("iterations", u64), ("kernel_launches", u64),
                ("render_ms", C.c_double), ("trace_ms", C.c_double), ("max_stack_depth", u64), ("guard_retraces", u64),
                ("reduce_ms", C.c_double), ("devices", u64)]


class KernelStats(C.Structure):
    _fields_ = [("ms", C.c_double), ("launches", u64), ("rays", u64), ("nodes_visited", u64), ("prims_tested", u64)]


K_GENERATE, K_EXTEND, K_SHADE, K_SHADOW, K_FILM, K_SINGLE, K_COUNT = range(7)
K_NAMES = ["generate", "extend", "shade", "shadow", "film", "single"]

import numpy as _np

HIT_DTYPE = _np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("shape", "<u4"), ("prim", "<u4"),
                       ("nodes_visited", "<u4"), ("prims_tested", "<u4"), ("reserved", "<u4")])
RAY_DTYPE = _np.dtype([("o", "<f4", 3), ("mint", "<f4"), ("d", "<f4", 3), ("maxt", "<f4")])
assert HIT_DTYPE.itemsize == C.sizeof(Hit) == 32 and RAY_DTYPE.itemsize == C.sizeof(Ray) == 32

# every entry point include/nori_gpu.h declares (tests check the library exports all of them)
ENTRY_POINTS = [
    "nori_gpu_init", "nori_gpu_init_multi", "nori_gpu_device_count", "nori_gpu_destroy", "nori_gpu_last_error", "nori_gpu_upload_scene",
    "nori_gpu_set_option", "nori_gpu_render", "nori_gpu_render_samples", "nori_gpu_clear_film",
    "nori_gpu_download_film", "nori_gpu_upload_film", "nori_gpu_film_device_ptr",
    "nori_gpu_film_dims", "nori_gpu_resolve", "nori_gpu_download_variance", "nori_gpu_trace", "nori_gpu_probe_bsdf", "nori_gpu_probe_emitter", "nori_gpu_pcg32",
    "nori_gpu_pcg32_uint", "nori_gpu_build_bvh", "nori_gpu_build_bvh_device", "nori_gpu_mesh_area_cdf", "nori_gpu_wide_layout", "nori_gpu_abi_sizes", "nori_gpu_selftest", "nori_gpu_get_stats", "nori_gpu_get_kernel_stats", "nori_gpu_reset_stats", "nori_gpu_synchronize",
]
