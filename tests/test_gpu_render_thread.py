"""The host-side mirror of the reference's RenderThread (render.h:30-52) on the device: progress and cancel points between
spp chunks like render.cpp:195-197, chunking does not change the image, the variance output rules."""
import os
import time

import numpy as np
import pytest

from conftest import GOLDEN
from nori_ray_tracer_b200 import nscene, render

pytestmark = pytest.mark.gpu


def _scene():
    return nscene.load_scene(os.path.join(GOLDEN, "cbox_path_mis.nscene"))


def test_chunked_render_reports_progress_and_equals_one_call(gpu):
    sc = _scene()
    rt = render.RenderThread(gpu=gpu)
    seen = []
    rgb, film = rt.render(sc, spp=24, seed=3, progress=seen.append, chunk_seconds=1e-4)     # tiny time budget: many chunks
    assert len(seen) >= 3 and seen[-1] == 1.0 and all(b > a for a, b in zip(seen, seen[1:]))
    rgb1, film1 = rt.render(sc, spp=24, seed=3, spp_chunk=24)
    assert np.abs(film - film1).max() <= 1e-5 * np.abs(film1).max()
    assert np.allclose(rgb, rgb1, rtol=1e-4, atol=1e-6)


def test_stop_rendering_interrupts_between_chunks(gpu):
    sc = _scene()
    sc.set_resolution(800, 600)
    rt = render.RenderThread(gpu=gpu)
    rt.renderScene(sc, spp=1 << 16, spp_chunk=4)                    # would take minutes
    time.sleep(0.3)
    assert rt.isBusy() and 0.0 <= rt.getProgress() < 1.0
    t0 = time.time()
    rt.stopRendering()
    assert time.time() - t0 < 5.0 and not rt.isBusy()
    assert rt.error is None and rt.result is not None            # the partial image is kept, like the reference's m_block
    assert np.isfinite(rt.result[0]).all()


def test_variance_output_and_its_limits(gpu):
    sc = _scene()
    rt = render.RenderThread(gpu=gpu)
    rt.render(sc, spp=6, seed=1, variance=True, spp_chunk=2)
    assert rt.variance_image is not None and rt.variance_image.shape == (sc.height, sc.width, 3) and rt.variance_image.mean() > 0
    rt.render(sc, spp=0, seed=1, variance=True)                  # nothing rendered: no statistic, no error
    assert rt.variance_image is None
