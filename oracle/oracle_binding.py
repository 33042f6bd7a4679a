"""ctypes binding for oracle/libnori_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product (nori-ray-tracer_b200/) never does."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libnori_oracle.so")


def build():
    subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)


def _load(abi):
    if not os.path.exists(_LIB):
        build()
    lib = C.CDLL(_LIB)
    lib.nori_oracle_create.restype = C.c_void_p
    lib.nori_oracle_create.argtypes = [C.POINTER(abi.Scene)]
    lib.nori_oracle_destroy.argtypes = [C.c_void_p]
    lib.nori_oracle_film_dims.argtypes = [C.c_void_p] + [C.POINTER(C.c_int32)] * 3
    lib.nori_oracle_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p] + [C.c_void_p] * 4
    lib.nori_oracle_trace_ordered.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
    lib.nori_oracle_pcg32.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
    lib.nori_oracle_pcg32_uint.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
    lib.nori_oracle_render_samples.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p]
    lib.nori_oracle_render.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
    lib.nori_oracle_render_var.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.nori_oracle_splat.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.nori_oracle_resolve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.nori_oracle_stats.argtypes = [C.c_void_p, C.POINTER(abi.Stats)]
    lib.nori_oracle_reset_stats.argtypes = [C.c_void_p]
    lib.nori_oracle_bsdf_probe.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p]
    lib.nori_oracle_emitter_probe.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p]
    lib.nori_oracle_block_sequence.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    return lib


class Oracle:
    """CPU restatement of the reference hot path over a SceneData (see nori_oracle.cpp)."""

    def __init__(self, scene, abi):
        self.abi, self.scene = abi, scene
        self.lib = _load(abi)
        self.h = self.lib.nori_oracle_create(C.byref(scene.pod))

    def close(self):
        if self.h:
            self.lib.nori_oracle_destroy(self.h)
            self.h = None

    __del__ = close

    def trace(self, rays, shadow, full=False):
        rays = np.ascontiguousarray(rays)
        n = rays.shape[0]
        hits = np.zeros(n, dtype=self.abi.HIT_DTYPE)
        if not full:
            self.lib.nori_oracle_trace(self.h, rays.ctypes.data, n, int(shadow), hits.ctypes.data, None, None, None, None)
            return hits
        p, uv = np.zeros((n, 3), np.float32), np.zeros((n, 2), np.float32)
        ns, ng = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
        self.lib.nori_oracle_trace(self.h, rays.ctypes.data, n, int(shadow), hits.ctypes.data,
                                   p.ctypes.data, uv.ctypes.data, ns.ctypes.data, ng.ctypes.data)
        return hits, p, uv, ns, ng

    def trace_ordered(self, rays, order):
        """Closest hits with the children of every inner node visited in another order (test of order independence)."""
        rays = np.ascontiguousarray(rays)
        hits = np.zeros(rays.shape[0], dtype=self.abi.HIT_DTYPE)
        self.lib.nori_oracle_trace_ordered(self.h, rays.ctypes.data, rays.shape[0], int(order), hits.ctypes.data)
        return hits

    def pcg32(self, initstate, initseq, n):
        out = np.zeros(n, np.float32)
        self.lib.nori_oracle_pcg32(initstate, initseq, n, out.ctypes.data)
        return out

    def pcg32_uint(self, initstate, initseq, n):
        out = np.zeros(n, np.uint32)
        self.lib.nori_oracle_pcg32_uint(initstate, initseq, n, out.ctypes.data)
        return out

    def render_samples(self, spp_begin, spp_count, seed=0):
        out = np.zeros((spp_count, self.scene.height, self.scene.width, 4), np.float32)
        self.lib.nori_oracle_render_samples(self.h, spp_begin, spp_count, seed, out.ctypes.data)
        return out

    def render(self, spp_begin, spp_count, seed=0, mode=0, film=None):
        if film is None:
            film = np.zeros(self.scene.film_shape, np.float32)
        self.lib.nori_oracle_render(self.h, spp_begin, spp_count, seed, mode, film.ctypes.data, None)
        return film

    def render_with_variance(self, spp, seed=0, mode=0):
        """(film, variance image): the reference's two outputs for a render of `spp` passes from scratch."""
        film = np.zeros(self.scene.film_shape, np.float32)
        vs = np.zeros((self.scene.height, self.scene.width, 3), np.float32); vs2 = np.zeros_like(vs)
        self.lib.nori_oracle_render_var(self.h, 0, spp, seed, mode, film.ctypes.data, None, vs.ctypes.data, vs2.ctypes.data)
        n = np.float32(spp)
        return film, vs2 / n - (vs / n) ** 2

    def splat(self, samples, spp_begin, seed=0, variance=False):
        """Film (and optionally the variance image) of the given per-sample radiance values (the layout of
        render_samples) at the film positions of the per-path streams: ImageBlock::put on its own."""
        samples = np.ascontiguousarray(samples, np.float32)
        film = np.zeros(self.scene.film_shape, np.float32)
        vs = np.zeros((self.scene.height, self.scene.width, 3), np.float32) if variance else None
        vs2 = np.zeros_like(vs) if variance else None
        self.lib.nori_oracle_splat(self.h, spp_begin, samples.shape[0], seed, samples.ctypes.data, film.ctypes.data,
                                   vs.ctypes.data if variance else None, vs2.ctypes.data if variance else None)
        if not variance:
            return film
        n = np.float32(samples.shape[0])
        return film, vs2 / n - (vs / n) ** 2

    def resolve(self, film):
        rgb = np.zeros((self.scene.height, self.scene.width, 3), np.float32)
        self.lib.nori_oracle_resolve(self.h, np.ascontiguousarray(film).ctypes.data, rgb.ctypes.data)
        return rgb

    def block_sequence(self, n):
        out = np.zeros((n, 5), np.float32)
        self.lib.nori_oracle_block_sequence(self.h, n, out.ctypes.data)
        return out

    def bsdf_probe(self, index, queries):
        q = np.ascontiguousarray(queries, np.float32); out = np.zeros((len(q), 12), np.float32)
        self.lib.nori_oracle_bsdf_probe(self.h, index, len(q), q.ctypes.data, out.ctypes.data)
        return out

    def emitter_probe(self, index, queries):
        q = np.ascontiguousarray(queries, np.float32); out = np.zeros((len(q), 15), np.float32)
        self.lib.nori_oracle_emitter_probe(self.h, index, len(q), q.ctypes.data, out.ctypes.data)
        return out

    def stats(self):
        s = self.abi.Stats()
        self.lib.nori_oracle_stats(self.h, C.byref(s))
        return s

    def reset_stats(self):
        self.lib.nori_oracle_reset_stats(self.h)
