"""Host-side scene construction without the reference's loader.

Builds the same flat description (nscene.SceneData) that the reference-side exporter writes, from numpy
arrays: meshes / spheres, BSDF and emitter parameter blocks, a perspective or thin-lens camera, the
tabulated reconstruction filter, and the SAH BVH built by the library's host builder
(nori_gpu_build_bvh, a restatement of bvh.cpp:54-382).  Used for the synthetic benchmark scenes
(BASELINE config 4: a 10 M-triangle height field) and by tests."""
import ctypes as C
import math

import numpy as np

from . import abi, gpu, nscene

f32 = np.float32


def gaussian_filter_table(radius=2.0, stddev=0.5):
    """GaussianFilter::eval (rfilter.cpp:37-42) tabulated like ImageBlock::init (block.cpp:59-63), in fp32."""
    radius, stddev = f32(radius), f32(stddev)
    alpha = f32(-1.0) / (f32(2.0) * stddev * stddev)
    tab = np.zeros(abi.FILTER_RESOLUTION + 1, f32)
    for i in range(abi.FILTER_RESOLUTION):
        x = (radius * f32(i)) / f32(abi.FILTER_RESOLUTION)
        tab[i] = max(f32(0), f32(np.exp(alpha * x * x)) - f32(np.exp(alpha * radius * radius)))
    return float(radius), tab


def look_at(origin, target, up):
    """parser.cpp lookat: camera-to-world with columns (left, newUp, dir, origin)."""
    o, t, u = (np.asarray(v, np.float64) for v in (origin, target, up))
    d = (t - o) / np.linalg.norm(t - o)
    left = np.cross(u / np.linalg.norm(u), d); left /= np.linalg.norm(left)
    new_up = np.cross(d, left)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = left, new_up, d, o
    return m


def perspective_sample_to_camera(width, height, fov, near, far):
    """PerspectiveCamera::activate (perspective.cpp:53-80)."""
    aspect = width / float(height)
    recip, cot = 1.0 / (far - near), 1.0 / math.tan(math.radians(fov / 2.0))
    persp = np.array([[cot, 0, 0, 0], [0, cot, 0, 0], [0, 0, far * recip, -near * far * recip], [0, 0, 1, 0]], np.float64)
    scale = np.diag([0.5, -0.5 * aspect, 1.0, 1.0])
    trans = np.eye(4); trans[0, 3], trans[1, 3] = 1.0, -1.0 / aspect
    return np.linalg.inv(scale @ trans @ persp)


class SceneBuilder:
    def __init__(self, integrator="path_mis"):
        self.integrator = integrator
        self.shapes, self.bsdfs, self.emitters = [], [], []
        self.arrays = {}
        self.camera = None
        self.filter = gaussian_filter_table()
        self.sample_count = 1

    # ---- materials / lights -----------------------------------------------------------------
    def diffuse(self, albedo=(0.5, 0.5, 0.5)):
        b = abi.Bsdf(); b.type = abi.BSDF_DIFFUSE; b.albedo_texture = abi.TEXTURE_CONSTANT
        b.albedo[:] = [float(a) for a in albedo]
        self.bsdfs.append(b); return len(self.bsdfs) - 1

    def mirror(self):
        b = abi.Bsdf(); b.type = abi.BSDF_MIRROR
        self.bsdfs.append(b); return len(self.bsdfs) - 1

    def dielectric(self, int_ior=1.5046, ext_ior=1.000277):
        b = abi.Bsdf(); b.type = abi.BSDF_DIELECTRIC; b.intIOR, b.extIOR = int_ior, ext_ior
        self.bsdfs.append(b); return len(self.bsdfs) - 1

    def microfacet(self, alpha=0.1, kd=(0.5, 0.5, 0.5), int_ior=1.5046, ext_ior=1.000277):
        b = abi.Bsdf(); b.type = abi.BSDF_MICROFACET; b.alpha, b.intIOR, b.extIOR = alpha, int_ior, ext_ior
        b.kd[:] = [float(k) for k in kd]; b.ks = float(f32(1) - f32(max(kd)))        # microfacet.cpp:48
        self.bsdfs.append(b); return len(self.bsdfs) - 1

    def area_light(self, radiance):
        e = abi.Emitter(); e.type = abi.EMITTER_AREA; e.shape = -1
        e.radiance[:] = [float(r) for r in radiance]
        self.emitters.append(e); return len(self.emitters) - 1

    def point_light(self, position, power):
        e = abi.Emitter(); e.type = abi.EMITTER_POINT; e.shape = -1
        e.position[:] = [float(p) for p in position]; e.radiance[:] = [float(p) for p in power]
        self.emitters.append(e); return len(self.emitters) - 1

    # ---- shapes ------------------------------------------------------------------------------
    def add_mesh(self, V, F, bsdf, emitter=-1, N=None, UV=None):
        i = len(self.shapes)
        s = abi.Shape(); s.type = abi.SHAPE_MESH; s.bsdf, s.emitter = bsdf, emitter
        V = np.ascontiguousarray(V, f32); F = np.ascontiguousarray(F, np.uint32)
        s.n_vertices, s.n_triangles = len(V), len(F)
        self.arrays[f"shape.{i}.V"], self.arrays[f"shape.{i}.F"] = V, F
        if N is not None:
            self.arrays[f"shape.{i}.N"] = np.ascontiguousarray(N, f32)
        if UV is not None:
            self.arrays[f"shape.{i}.UV"] = np.ascontiguousarray(UV, f32)
        lib = gpu.load_library()
        cdf = np.zeros(len(F) + 1, f32); norm = C.c_float()
        lib.nori_gpu_mesh_area_cdf(V.ctypes.data, F.ctypes.data, len(F), cdf.ctypes.data, C.byref(norm))
        self.arrays[f"shape.{i}.area_cdf"] = cdf
        s.area_normalization = norm.value
        if emitter >= 0:
            self.emitters[emitter].shape = i
        self.shapes.append(s); return i

    def add_sphere(self, center, radius, bsdf, emitter=-1):
        i = len(self.shapes)
        s = abi.Shape(); s.type = abi.SHAPE_SPHERE; s.bsdf, s.emitter = bsdf, emitter
        s.n_triangles = 1; s.center[:] = [float(c) for c in center]; s.radius = radius
        if emitter >= 0:
            self.emitters[emitter].shape = i
        self.shapes.append(s); return i

    # ---- camera ------------------------------------------------------------------------------
    def perspective(self, width, height, fov, origin, target, up, near=1e-4, far=1e4, scale_x=-1.0,
                    lens_radius=0.0, focal_dist=1.0):
        c = abi.Camera()
        c.type = abi.CAMERA_THINLENS if lens_radius > 0 else abi.CAMERA_PERSPECTIVE
        c.width, c.height = width, height
        s2c = perspective_sample_to_camera(width, height, fov, near, far).astype(f32)
        c2w = (look_at(origin, target, up) @ np.diag([scale_x, 1, 1, 1])).astype(f32)   # <scale value="-1,1,1"/> then lookat
        c.sampleToCamera[:] = s2c.reshape(-1).tolist(); c.cameraToWorld[:] = c2w.reshape(-1).tolist()
        c.invOutputSize[0], c.invOutputSize[1] = f32(1) / f32(width), f32(1) / f32(height)
        c.nearClip, c.farClip, c.lensRadius, c.focalDistance = near, far, lens_radius, focal_dist
        self.camera = c

    # ---- finish -------------------------------------------------------------------------------
    def build(self, threads=0, builder="sah", leaf_size=4, device=0):
        """builder="sah": the host restatement of the reference's SAH builder (reference-identical trees);
        builder="lbvh": the GPU builder (nori_gpu_build_bvh_device) -- same node format, different tree."""
        lib = gpu.load_library()
        self.build_ms = None
        n_shapes = len(self.shapes)
        shapes = (abi.Shape * max(n_shapes, 1))(*self.shapes)
        keep = []
        for i in range(n_shapes):
            for field, ct in (("V", C.c_float), ("F", C.c_uint32)):
                arr = self.arrays.get(f"shape.{i}.{field}")
                if arr is not None:
                    setattr(shapes[i], field, arr.ctypes.data_as(C.POINTER(ct))); keep.append(arr)
        total = sum(s.n_triangles for s in self.shapes)
        nodes = np.zeros((max(2 * total, 1), 8), np.uint32)
        indices = np.zeros(max(total, 1), np.uint32)
        offsets = np.zeros(n_shapes + 1, np.uint32)
        n_nodes = C.c_uint32()
        if builder == "lbvh":
            ms = C.c_float()
            rc = lib.nori_gpu_build_bvh_device(device, shapes, n_shapes, nodes.ctypes.data, indices.ctypes.data, offsets.ctypes.data,
                                               C.byref(n_nodes), leaf_size, C.byref(ms))
            self.build_ms = ms.value
        else:
            rc = lib.nori_gpu_build_bvh(shapes, n_shapes, nodes.ctypes.data, indices.ctypes.data, offsets.ctypes.data,
                                        C.byref(n_nodes), threads)
        if rc != 0:
            raise RuntimeError(f"BVH build ({builder}) failed")
        for s in shapes:                                   # the container stores PODs with null pointers
            s.V = s.N = s.UV = s.area_cdf = None; s.F = None
        e = dict(self.arrays)
        e["header"] = np.array([abi.ABI_VERSION, abi.INTEGRATOR_NAMES[self.integrator], self.sample_count, 0], np.int32)
        e["av_length"] = np.zeros(1, f32)
        e["bvh.nodes"] = nodes[:n_nodes.value].copy()
        e["bvh.indices"] = indices[:total].copy()
        e["bvh.shape_offset"] = offsets
        e["shapes.pod"] = np.frombuffer(bytes(shapes), np.uint8)[:n_shapes * C.sizeof(abi.Shape)].copy()
        e["bsdfs.pod"] = np.frombuffer(b"".join(bytes(b) for b in self.bsdfs), np.uint8).copy()
        ems = b"".join(bytes(x) for x in self.emitters)
        e["emitters.pod"] = np.frombuffer(ems, np.uint8).copy() if ems else np.zeros(0, np.uint8)
        e["camera.pod"] = np.frombuffer(bytes(self.camera), np.uint8).copy()
        flt = abi.Filter(); flt.radius = self.filter[0]; flt.table[:] = self.filter[1].tolist()
        e["filter.pod"] = np.frombuffer(bytes(flt), np.uint8).copy()
        e["medium.pod"] = np.frombuffer(bytes(abi.Medium()), np.uint8).copy()
        return nscene.SceneData(e)


def heightfield_scene(n=2237, width=3840, height=2160, seed=0, integrator="path_mis", builder="sah", leaf_size=4, return_builder=False):
    """BASELINE config 4: an n x n-vertex height field (2*(n-1)^2 triangles; n = 2237 -> 9,999,392),
    z = seeded value noise, diffuse, lit by one area-light quad, viewed from above at an angle."""
    rng = np.random.RandomState(seed)
    coarse = rng.rand(65, 65).astype(f32)
    xs = np.linspace(0, 64, n, dtype=f32)
    i0 = np.minimum(xs.astype(np.int32), 63); fr = xs - i0
    sm = fr * fr * (3 - 2 * fr)
    rows = coarse[i0][:, :] * (1 - sm)[:, None] + coarse[i0 + 1][:, :] * sm[:, None]        # n x 65
    z = rows[:, i0] * (1 - sm)[None, :] + rows[:, i0 + 1] * sm[None, :]                       # n x n
    fine = rng.rand(n, n).astype(f32) * f32(0.002)
    gx, gy = np.meshgrid(np.linspace(-1, 1, n, dtype=f32), np.linspace(-1, 1, n, dtype=f32), indexing="ij")
    V = np.stack([gx, gy, f32(0.25) * z + fine], -1).reshape(-1, 3)
    idx = np.arange(n * n, dtype=np.uint32).reshape(n, n)
    a, b, c, d = idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]
    F = np.concatenate([np.stack([a, b, c], -1).reshape(-1, 3), np.stack([a, c, d], -1).reshape(-1, 3)], 0)
    sb = SceneBuilder(integrator)
    ground = sb.diffuse((0.6, 0.55, 0.5))
    sb.add_mesh(V, F, ground)
    light = sb.area_light((30, 30, 30))
    lv = np.array([[-0.4, -0.4, 1.5], [0.4, -0.4, 1.5], [0.4, 0.4, 1.5], [-0.4, 0.4, 1.5]], f32)
    sb.add_mesh(lv, np.array([[0, 2, 1], [0, 3, 2]], np.uint32), sb.diffuse((0, 0, 0)), emitter=light)   # faces down
    sb.perspective(width, height, 40.0, origin=(0.0, -2.2, 1.6), target=(0, 0, 0.1), up=(0, 0, 1))
    sc = sb.build(builder=builder, leaf_size=leaf_size)
    return (sc, sb) if return_builder else sc


def rebuild_bvh(scene, builder="lbvh", leaf_size=4, device=0, threads=0):
    """A copy of `scene` (nscene.SceneData) whose BVH was rebuilt from its shapes by the host SAH builder
    ("sah": nori_gpu_build_bvh) or the GPU linear-BVH builder ("lbvh": nori_gpu_build_bvh_device).  Returns
    (new scene, device build time in ms or None)."""
    lib = gpu.load_library()
    n_shapes = scene.pod.n_shapes
    total = int(scene.indices.size)
    nodes = np.zeros((max(2 * total, 1), 8), np.uint32)
    indices = np.zeros(max(total, 1), np.uint32)
    offsets = np.zeros(n_shapes + 1, np.uint32)
    n_nodes = C.c_uint32(); ms = C.c_float()
    shapes = C.cast(scene.shapes, C.POINTER(abi.Shape))
    if builder == "lbvh":
        rc = lib.nori_gpu_build_bvh_device(device, shapes, n_shapes, nodes.ctypes.data, indices.ctypes.data, offsets.ctypes.data,
                                           C.byref(n_nodes), leaf_size, C.byref(ms))
    else:
        rc = lib.nori_gpu_build_bvh(shapes, n_shapes, nodes.ctypes.data, indices.ctypes.data, offsets.ctypes.data, C.byref(n_nodes), threads)
    if rc != 0:
        raise RuntimeError(f"BVH build ({builder}) failed")
    e = dict(scene.entries)
    e["bvh.nodes"] = nodes[:n_nodes.value].copy()
    e["bvh.indices"] = indices[:total].copy()
    e["bvh.shape_offset"] = offsets
    for k in [k for k in e if k.startswith("rays")]:       # reference-answered ray batches belong to the reference tree's counters
        del e[k]
    return nscene.SceneData(e), (ms.value if builder == "lbvh" else None)
