"""Host-side mirror of the reference's RenderThread (include/nori/render.h:30-52, src/render.cpp:135-290).

Same surface -- renderScene(), isBusy(), stopRendering(), getProgress(), isRenderingDone() -- but the
body of the render-thread lambda (render.cpp:173-284: spp-major loop, TBB over blocks, block merge) is
replaced by calls into the C ABI (nori_gpu_render accumulates spp chunks into the device film, so the
host loop keeps the reference's progress / cancel points, render.cpp:195-197).

Multi-GPU (one process per GPU, torch.distributed): the scene is replicated, sample indices are
partitioned [g*spp/G, (g+1)*spp/G) with disjoint pcg32 initstate ranges, and the float accumulation
buffers are summed onto rank 0 with ONE reduce (NCCL over NVLink) -- the path has no other exchange."""
import os
import threading
import time

import numpy as np

from . import imageio


def shard_spp(spp, rank, world):
    """Sample-index range of `rank`: [rank*spp/world, (rank+1)*spp/world)."""
    begin = (rank * spp) // world
    end = ((rank + 1) * spp) // world
    return begin, end - begin


def reduce_film(film, dst=0):
    """Sum the (H+2b, W+2b, 4) accumulation buffers of all ranks onto `dst` (the single collective)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM)
    return film


def resolve_film(film, border):
    """ImageBlock::toBitmap (block.cpp:76-82): rgb / w inside the border, 0 where w == 0."""
    f = np.asarray(film)
    h, w = f.shape[0] - 2 * border, f.shape[1] - 2 * border
    c = f[border:border + h, border:border + w]
    wgt = c[..., 3:4]
    return np.where(wgt != 0, c[..., :3] / np.where(wgt != 0, wgt, 1), 0).astype(np.float32)


class RenderThread:
    def __init__(self, device=0, gpu=None):
        from .gpu import NoriGpu
        self.gpu = gpu if gpu is not None else NoriGpu(device)
        self._status = 0            # 0 free, 1 busy, 2 interruption requested, 3 done (render.h:49)
        self._progress = 0.0
        self._thread = None
        self.result = None
        self.error = None
        self.variance_image = None

    # ---- the reference's polling surface ----------------------------------------------------
    def isBusy(self):
        return self._status in (1, 2)

    def isRenderingDone(self):
        return self._status == 3

    def getProgress(self):
        return self._progress if self.isBusy() else 1.0

    def stopRendering(self):
        if self.isBusy():
            self._status = 2
            self._thread.join()
            self._status = 0

    # ---- synchronous core -------------------------------------------------------------------
    def render(self, scene, spp=None, spp_chunk=None, seed=0, distributed=False, progress=None, variance=False,
               chunk_seconds=0.25):
        """Render `scene` (a nscene.SceneData); returns (rgb bitmap, film) on rank 0, (None, film) elsewhere.

        The spp loop runs in chunks so that stopRendering() and getProgress() act between passes like the
        reference's per-pass checks (render.cpp:195-197): `spp_chunk` passes per nori_gpu_render call, or
        (None) as many as take about `chunk_seconds` -- the first chunk is 1 pass, later ones are sized from the
        measured rate.  One chunk = one wavefront batch, so very small chunks cost throughput (the pool drains
        once per call); benchmarks pass spp_chunk=spp."""
        g = self.gpu
        spp = scene.sample_count if spp is None else spp
        rank, world = 0, 1
        if distributed:
            import torch.distributed as dist
            rank, world = dist.get_rank(), dist.get_world_size()
        if variance and world > 1:
            # the statistic is the running mean after every pass (render.cpp:238-247): it depends on the
            # order of ALL passes and does not decompose over sample-index shards
            raise ValueError("variance=True is not available for sharded (distributed) renders")
        begin, count = shard_spp(spp, rank, world)
        g.upload_scene(scene)
        if variance:                                   # the reference's <scene>_variance.exr (render.cpp:263-278)
            g.set_option("variance", 1)
        g.clear_film()
        chunk = max(1, int(spp_chunk)) if spp_chunk else 1
        done = 0
        while done < count:
            self._progress = done / max(count, 1)
            if self._status == 2:                      # render.cpp:196-197
                break
            n = min(chunk, count - done)
            t0 = time.perf_counter()
            g.render(begin + done, n, seed)
            dt = time.perf_counter() - t0
            done += n
            if not spp_chunk:                          # aim at chunk_seconds per call, at most doubling per step
                chunk = max(1, min(2 * chunk + 1, int(n * chunk_seconds / max(dt, 1e-6))))
            if progress:
                progress(done / count)
        self._progress = done / max(count, 1)
        if distributed and world > 1:
            import torch
            film_t = torch.as_tensor(g.film_device_array(), device=f"cuda:{g.device}")
            g.synchronize()
            reduce_film(film_t, dst=0)
            torch.cuda.synchronize(g.device)
        film = g.download_film()
        rgb = g.resolve() if rank == 0 else None
        self.variance_image = g.variance() if (variance and done > 0) else None
        return rgb, film

    # ---- the reference's asynchronous entry point ---------------------------------------------
    def renderScene(self, scene, output=None, **kw):
        """Start rendering in a background thread; writes `output` (.exr) when done (render.cpp:256-261)."""
        self._status, self._progress, self.result, self.error = 1, 0.0, None, None

        def work():
            try:
                t0 = time.time()
                rgb, film = self.render(scene, **kw)
                self.result = (rgb, film, time.time() - t0)
                if output and rgb is not None:
                    imageio.write_exr(output, rgb)
                    if self.variance_image is not None:
                        imageio.write_exr(output[:-4] + "_variance.exr", self.variance_image)
            except Exception as e:                       # surfaced through .error, like NoriException
                self.error = e
            finally:
                self._status = 3
        self._thread = threading.Thread(target=work, daemon=True)
        self._thread.start()
        return self._thread
