"""Flat scene container (.nscene) <-> the nori_gpu_scene POD of include/nori_gpu.h.

A .nscene file is what the reference-side exporter writes (oracle/ref_tools/nori_export.cpp walks a
loaded nori::Scene: BVH::m_nodes / m_indices / m_shapeOffset (bvh.h:165-170), Mesh::m_V/m_N/m_UV/m_F
(mesh.h:121-124), BSDF / emitter / camera parameters) and what the host scene builder
(host_scene.py) produces.  Layout: magic "NSCN0001", u32 entry count, then per entry
u32 name_len, name, u32 dtype (0 f32, 1 u32, 2 i32, 3 u8), u32 ndim, u64 dims[], data padded to 8.
"""
import ctypes as C
import struct

import numpy as np

from . import abi

_DTYPES = {0: np.float32, 1: np.uint32, 2: np.int32, 3: np.uint8}
_CODES = {np.dtype(np.float32): 0, np.dtype(np.uint32): 1, np.dtype(np.int32): 2, np.dtype(np.uint8): 3}


def read_container(path):
    """Return {name: ndarray} for every entry of a .nscene file."""
    buf = open(path, "rb").read()
    if buf[:8] != b"NSCN0001":
        raise ValueError(f"{path}: not a .nscene container")
    (count,) = struct.unpack_from("<I", buf, 8)
    pos, out = 12, {}
    for _ in range(count):
        (nl,) = struct.unpack_from("<I", buf, pos); pos += 4
        name = buf[pos:pos + nl].decode(); pos += nl
        dtype, ndim = struct.unpack_from("<II", buf, pos); pos += 8
        dims = struct.unpack_from(f"<{ndim}Q", buf, pos); pos += 8 * ndim
        dt = np.dtype(_DTYPES[dtype])
        n = int(np.prod(dims)) if ndim else 1
        out[name] = np.frombuffer(buf, dtype=dt, count=n, offset=pos).reshape(dims).copy()
        pos += n * dt.itemsize
        pos = (pos + 7) & ~7
    return out


def write_container(path, entries):
    """Inverse of read_container; `entries` is {name: ndarray (f32/u32/i32/u8)}."""
    parts = [b"NSCN0001", struct.pack("<I", len(entries))]
    size = 12
    for name, arr in entries.items():
        arr = np.ascontiguousarray(arr)
        nb = name.encode()
        head = struct.pack("<I", len(nb)) + nb + struct.pack("<II", _CODES[arr.dtype], arr.ndim)
        head += struct.pack(f"<{arr.ndim}Q", *arr.shape)
        data = arr.tobytes()
        pad = (-(size + len(head) + len(data))) % 8
        parts += [head, data, b"\0" * pad]
        size += len(head) + len(data) + pad
    with open(path, "wb") as f:
        f.write(b"".join(parts))


def _ptr(arr, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype)) if arr is not None else None


class SceneData:
    """Owns the host arrays of one scene and exposes them as a ctypes `abi.Scene` (self.pod).

    The numpy arrays are kept alive by this object; the POD only borrows pointers, exactly like the
    reference-side binding borrows pointers into nori::Mesh / nori::BVH storage."""

    def __init__(self, entries):
        self.entries = entries
        e = entries
        hdr = e["header"]
        # containers written for ABI 2 are what ABI 3 reads (only nori_gpu_stats and the entry points grew);
        # version-1 containers lack images and the advanced camera fields
        if int(hdr[0]) not in (1, 2, abi.ABI_VERSION):
            raise ValueError("ABI version mismatch in .nscene")
        self.sample_count = int(hdr[2])
        self.nodes = np.ascontiguousarray(e["bvh.nodes"], dtype=np.uint32)
        self.indices = np.ascontiguousarray(e["bvh.indices"], dtype=np.uint32)
        self.shape_offset = np.ascontiguousarray(e["bvh.shape_offset"], dtype=np.uint32)
        n_shapes = e["shapes.pod"].size // C.sizeof(abi.Shape)
        n_bsdfs = e["bsdfs.pod"].size // C.sizeof(abi.Bsdf)
        n_emitters = e["emitters.pod"].size // C.sizeof(abi.Emitter)
        self.shapes = (abi.Shape * max(n_shapes, 1)).from_buffer_copy(
            e["shapes.pod"].tobytes().ljust(C.sizeof(abi.Shape), b"\0"))
        self.bsdfs = (abi.Bsdf * max(n_bsdfs, 1)).from_buffer_copy(
            e["bsdfs.pod"].tobytes().ljust(C.sizeof(abi.Bsdf), b"\0"))
        self.emitters = (abi.Emitter * max(n_emitters, 1)).from_buffer_copy(
            e["emitters.pod"].tobytes().ljust(C.sizeof(abi.Emitter), b"\0"))
        self._keep = []
        for i in range(n_shapes):
            s = self.shapes[i]
            for field, ctype in (("V", C.c_float), ("N", C.c_float), ("UV", C.c_float),
                                 ("F", C.c_uint32), ("area_cdf", C.c_float)):
                arr = e.get(f"shape.{i}.{field}")
                if arr is not None:
                    arr = np.ascontiguousarray(arr)
                    self._keep.append(arr)
                setattr(s, field, _ptr(arr, ctype))
        for i in range(n_emitters):
            em = self.emitters[i]
            for field in ("env_image", "env_pdf", "env_cdf", "env_pmarginal", "env_cmarginal"):
                arr = e.get(f"emitter.{i}.{field}")
                if arr is not None:
                    arr = np.ascontiguousarray(arr, dtype=np.float32)
                    self._keep.append(arr)
                setattr(em, field, _ptr(arr, C.c_float))
        # image textures / normal maps: "image.<i>.rgb" (H, W, 3) u8 + "image.<i>.wrap" (1,) i32
        n_images = 0
        while f"image.{n_images}.rgb" in e:
            n_images += 1
        self.images = (abi.Image * max(n_images, 1))()
        for i in range(n_images):
            rgb = np.ascontiguousarray(e[f"image.{i}.rgb"], dtype=np.uint8)
            self._keep.append(rgb)
            im = self.images[i]
            im.height, im.width = rgb.shape[0], rgb.shape[1]
            im.wrap = int(e[f"image.{i}.wrap"][0])
            im.rgb = rgb.ctypes.data_as(C.POINTER(C.c_uint8))
        pod = abi.Scene()
        pod.abi_version = abi.ABI_VERSION
        pod.n_images = n_images
        pod.images = C.cast(self.images, C.POINTER(abi.Image))
        pod.integrator = int(hdr[1])
        pod.av_length = float(e["av_length"][0])
        pod.n_nodes = self.nodes.shape[0]
        pod.n_indices = self.indices.size
        pod.n_shapes, pod.n_bsdfs, pod.n_emitters = n_shapes, n_bsdfs, n_emitters
        pod.nodes = C.cast(self.nodes.ctypes.data, C.POINTER(abi.BvhNode))
        pod.indices = _ptr(self.indices, C.c_uint32)
        pod.shape_offset = _ptr(self.shape_offset, C.c_uint32)
        pod.shapes = C.cast(self.shapes, C.POINTER(abi.Shape))
        pod.bsdfs = C.cast(self.bsdfs, C.POINTER(abi.Bsdf))
        pod.emitters = C.cast(self.emitters, C.POINTER(abi.Emitter))
        pod.camera = abi.Camera.from_buffer_copy(e["camera.pod"].tobytes().ljust(C.sizeof(abi.Camera), b"\0"))
        pod.filter = abi.Filter.from_buffer_copy(e["filter.pod"].tobytes())
        pod.medium = abi.Medium.from_buffer_copy(e["medium.pod"].tobytes())
        self.pod = pod

    # ---- conveniences -------------------------------------------------------------------
    @property
    def width(self):
        return self.pod.camera.width

    @property
    def height(self):
        return self.pod.camera.height

    @property
    def border(self):
        import math
        return int(math.ceil(self.pod.filter.radius - 0.5))      # block.cpp:57

    @property
    def film_shape(self):
        b = self.border
        return (self.height + 2 * b, self.width + 2 * b, 4)

    def set_integrator(self, name):
        self.pod.integrator = abi.INTEGRATOR_NAMES[name]

    def set_resolution(self, width, height):
        """Change the output size the way PerspectiveCamera's constructor would (perspective.cpp:36-38).
        Only valid when the aspect ratio is unchanged (sampleToCamera depends on it)."""
        cam = self.pod.camera
        if abs(width / height - cam.width / cam.height) > 1e-6:
            raise ValueError("set_resolution must preserve the aspect ratio")
        cam.width, cam.height = width, height
        cam.invOutputSize[0] = np.float32(1.0) / np.float32(width)
        cam.invOutputSize[1] = np.float32(1.0) / np.float32(height)

    def set_film(self, width, height):
        """Output size with ANY aspect ratio: rebuilds sampleToCamera the way PerspectiveCamera's constructor does
        (perspective.cpp:53-80: inverse of scale(0.5, -0.5*aspect, 1) * translate(1, -1/aspect, 0) * perspective(fov))
        from the field of view encoded in the current matrix.  For throughput configurations at resolutions no
        fixture was exported at (BASELINE config 5: 3840 x 2160); float64 arithmetic, rounded once."""
        cam = self.pod.camera
        s2c = np.array(list(cam.sampleToCamera), np.float64).reshape(4, 4)
        cot = 2.0 / s2c[0, 0]                                         # M^-1[0][0] = 1 / (0.5 * cot)
        near, far = float(cam.nearClip), float(cam.farClip)
        aspect = width / float(height)
        recip = 1.0 / (far - near)
        persp = np.array([[cot, 0, 0, 0], [0, cot, 0, 0], [0, 0, far * recip, -near * far * recip], [0, 0, 1, 0]], np.float64)
        scale = np.diag([0.5, -0.5 * aspect, 1.0, 1.0])
        trans = np.eye(4); trans[0, 3] = 1.0; trans[1, 3] = -1.0 / aspect
        inv = np.linalg.inv(scale @ trans @ persp)
        inv[np.abs(inv) < 1e-12] = 0.0
        inv = inv.astype(np.float32).reshape(-1)
        for i in range(16):
            cam.sampleToCamera[i] = inv[i]
        cam.width, cam.height = width, height
        cam.invOutputSize[0] = np.float32(1.0) / np.float32(width)
        cam.invOutputSize[1] = np.float32(1.0) / np.float32(height)

    def ray_batch(self):
        """Reference-answered ray batch stored by nori_export --rays (None if absent)."""
        e = self.entries
        if "rays" not in e:
            return None
        rays = np.ascontiguousarray(e["rays"]).view(abi.RAY_DTYPE).reshape(-1)
        hits = np.ascontiguousarray(e["rays.hits"]).view(abi.HIT_DTYPE).reshape(-1)
        return {"rays": rays, "shadow": e["rays.shadow"], "hits": hits, "p": e["rays.hit_p"],
                "uv": e["rays.hit_uv"], "n": e["rays.hit_n"], "ng": e["rays.hit_ng"]}


def load_scene(path):
    return SceneData(read_container(path))
