"""Thin ctypes binding over csrc/libnori_gpu.so (the C ABI of include/nori_gpu.h).

This is the product path: there is NO CPU fallback.  A missing library or a missing GPU raises."""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libnori_gpu.so")
_lib = None


class NoriGpuError(RuntimeError):
    pass


def load_library():
    """dlopen libnori_gpu.so and declare every entry point; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NoriGpuError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a). There is no CPU fallback for the rendering hot path.")
    lib = C.CDLL(LIB_PATH)
    vp, u32, u64, i64 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int64
    sig = {
        "nori_gpu_init": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "nori_gpu_init_multi": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]),
        "nori_gpu_device_count": (C.c_int, [vp]),
        "nori_gpu_destroy": (None, [vp]),
        "nori_gpu_last_error": (C.c_char_p, [vp]),
        "nori_gpu_upload_scene": (C.c_int, [vp, C.POINTER(abi.Scene)]),
        "nori_gpu_set_option": (C.c_int, [vp, C.c_char_p, i64]),
        "nori_gpu_render": (C.c_int, [vp, u32, u32, u64]),
        "nori_gpu_render_samples": (C.c_int, [vp, u32, u32, u64, vp]),
        "nori_gpu_clear_film": (C.c_int, [vp]),
        "nori_gpu_download_film": (C.c_int, [vp, vp]),
        "nori_gpu_upload_film": (C.c_int, [vp, vp]),
        "nori_gpu_film_device_ptr": (C.c_int, [vp, C.POINTER(vp), C.POINTER(u64)]),
        "nori_gpu_film_dims": (C.c_int, [vp] + [C.POINTER(C.c_int32)] * 3),
        "nori_gpu_resolve": (C.c_int, [vp, vp]),
        "nori_gpu_download_variance": (C.c_int, [vp, vp]),
        "nori_gpu_trace": (C.c_int, [vp, vp, u64, C.c_int, vp]),
        "nori_gpu_probe_bsdf": (C.c_int, [vp, u32, u64, vp, vp]),
        "nori_gpu_probe_emitter": (C.c_int, [vp, u32, u64, vp, vp]),
        "nori_gpu_pcg32": (C.c_int, [vp, u64, u64, u64, vp]),
        "nori_gpu_pcg32_uint": (C.c_int, [vp, u64, u64, u64, vp]),
        "nori_gpu_abi_sizes": (C.c_int, [C.POINTER(u32), C.c_int]),
        "nori_gpu_selftest": (C.c_int, [vp, u64, C.POINTER(u64)]),
        "nori_gpu_build_bvh": (C.c_int, [C.POINTER(abi.Shape), u32, vp, vp, vp, C.POINTER(u32), C.c_int]),
        "nori_gpu_build_bvh_device": (C.c_int, [C.c_int, C.POINTER(abi.Shape), u32, vp, vp, vp, C.POINTER(u32), u32, C.POINTER(C.c_float)]),
        "nori_gpu_mesh_area_cdf": (C.c_int, [vp, vp, u32, vp, C.POINTER(C.c_float)]),
        "nori_gpu_wide_layout": (C.c_int, [vp, u32, u32, vp, u32, C.POINTER(u32)]),
        "nori_gpu_get_stats": (C.c_int, [vp, C.POINTER(abi.Stats)]),
        "nori_gpu_get_kernel_stats": (C.c_int, [vp, C.POINTER(abi.KernelStats)]),
        "nori_gpu_reset_stats": (C.c_int, [vp]),
        "nori_gpu_synchronize": (C.c_int, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError here = the library does not export the ABI
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


class _CudaArray:
    """Exposes a raw device pointer through __cuda_array_interface__ (for torch.as_tensor)."""

    def __init__(self, ptr, shape, owner):
        self._owner = owner
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class NoriGpu:
    """One rendering context on one CUDA device (mirrors what RenderThread owns, render.h:30-52)."""

    def __init__(self, device=0, devices=None):
        """`devices` (a list of device indices): one context over several GPUs of the node (nori_gpu_init_multi:
        the scene is replicated, render() shards the sample indices and sums the films onto devices[0])."""
        self.lib = load_library()
        self.ctx = C.c_void_p()
        if devices is not None:
            arr = (C.c_int * len(devices))(*devices)
            rc = self.lib.nori_gpu_init_multi(arr, len(devices), C.byref(self.ctx))
            device = devices[0] if len(devices) else 0
        else:
            rc = self.lib.nori_gpu_init(device, C.byref(self.ctx))
        if rc != 0:
            raise NoriGpuError(self.lib.nori_gpu_last_error(None).decode())
        self.device = device
        self.devices = list(devices) if devices is not None else [device]
        self.scene = None

    def close(self):
        if getattr(self, "ctx", None) and self.ctx.value:
            self.lib.nori_gpu_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise NoriGpuError(self.lib.nori_gpu_last_error(self.ctx).decode())

    # ---- scene / options --------------------------------------------------------------------
    def upload_scene(self, scene):
        self._check(self.lib.nori_gpu_upload_scene(self.ctx, C.byref(scene.pod)))
        self.scene = scene

    def set_option(self, name, value):
        self._check(self.lib.nori_gpu_set_option(self.ctx, name.encode(), int(value)))

    # ---- rendering --------------------------------------------------------------------------
    def render(self, spp_begin, spp_count, seed=0):
        self._check(self.lib.nori_gpu_render(self.ctx, spp_begin, spp_count, seed))

    def render_samples(self, spp_begin, spp_count, seed=0):
        out = np.empty((spp_count, self.scene.height, self.scene.width, 4), np.float32)
        self._check(self.lib.nori_gpu_render_samples(self.ctx, spp_begin, spp_count, seed, out.ctypes.data))
        return out

    def clear_film(self):
        self._check(self.lib.nori_gpu_clear_film(self.ctx))

    def film_dims(self):
        r, c, b = C.c_int32(), C.c_int32(), C.c_int32()
        self._check(self.lib.nori_gpu_film_dims(self.ctx, C.byref(r), C.byref(c), C.byref(b)))
        return r.value, c.value, b.value

    def download_film(self, out=None):
        rows, cols, _ = self.film_dims()
        if out is None:
            out = np.empty((rows, cols, 4), np.float32)
        self._check(self.lib.nori_gpu_download_film(self.ctx, out.ctypes.data))
        return out

    def upload_film(self, film):
        film = np.ascontiguousarray(film, np.float32)
        self._check(self.lib.nori_gpu_upload_film(self.ctx, film.ctypes.data))

    def film_device_array(self):
        """Zero-copy view of the device film for torch.distributed collectives."""
        ptr, n = C.c_void_p(), C.c_uint64()
        self._check(self.lib.nori_gpu_film_device_ptr(self.ctx, C.byref(ptr), C.byref(n)))
        rows, cols, _ = self.film_dims()
        return _CudaArray(ptr.value, (rows, cols, 4), self)

    def resolve(self):
        rgb = np.empty((self.scene.height, self.scene.width, 3), np.float32)
        self._check(self.lib.nori_gpu_resolve(self.ctx, rgb.ctypes.data))
        return rgb

    def variance(self):
        """<scene>_variance.exr of the reference (needs set_option('variance', 1) before rendering)."""
        rgb = np.empty((self.scene.height, self.scene.width, 3), np.float32)
        self._check(self.lib.nori_gpu_download_variance(self.ctx, rgb.ctypes.data))
        return rgb

    # ---- test hooks -------------------------------------------------------------------------
    def trace(self, rays, shadow):
        rays = np.ascontiguousarray(rays)
        assert rays.dtype == abi.RAY_DTYPE
        hits = np.zeros(rays.shape[0], dtype=abi.HIT_DTYPE)
        self._check(self.lib.nori_gpu_trace(self.ctx, rays.ctypes.data, rays.shape[0], int(shadow), hits.ctypes.data))
        return hits

    def probe_bsdf(self, index, queries):
        q = np.ascontiguousarray(queries, np.float32)
        out = np.zeros((len(q), 12), np.float32)
        self._check(self.lib.nori_gpu_probe_bsdf(self.ctx, index, len(q), q.ctypes.data, out.ctypes.data))
        return out

    def probe_emitter(self, index, queries):
        q = np.ascontiguousarray(queries, np.float32)
        out = np.zeros((len(q), 15), np.float32)
        self._check(self.lib.nori_gpu_probe_emitter(self.ctx, index, len(q), q.ctypes.data, out.ctypes.data))
        return out

    def pcg32(self, initstate, initseq, n):
        out = np.empty(n, np.float32)
        self._check(self.lib.nori_gpu_pcg32(self.ctx, initstate, initseq, n, out.ctypes.data))
        return out

    def pcg32_uint(self, initstate, initseq, n):
        out = np.empty(n, np.uint32)
        self._check(self.lib.nori_gpu_pcg32_uint(self.ctx, initstate, initseq, n, out.ctypes.data))
        return out

    def selftest(self, n):
        """(division, square root, reciprocal) operand counts on which the slow-path-free IEEE sequences differ from the
        compiler's correctly rounded operations; all zero when the library is sound."""
        out = (C.c_uint64 * 3)()
        self._check(self.lib.nori_gpu_selftest(self.ctx, n, out))
        return tuple(int(v) for v in out)

    def stats(self):
        s = abi.Stats()
        self._check(self.lib.nori_gpu_get_stats(self.ctx, C.byref(s)))
        return s

    def kernel_stats(self):
        arr = (abi.KernelStats * abi.K_COUNT)()
        self._check(self.lib.nori_gpu_get_kernel_stats(self.ctx, arr))
        return {abi.K_NAMES[i]: dict(ms=arr[i].ms, launches=arr[i].launches, rays=arr[i].rays,
                                     nodes=arr[i].nodes_visited, prims=arr[i].prims_tested) for i in range(abi.K_COUNT)}

    def reset_stats(self):
        self._check(self.lib.nori_gpu_reset_stats(self.ctx))

    def synchronize(self):
        self._check(self.lib.nori_gpu_synchronize(self.ctx))
