"""Shared fixtures.  `-m "not gpu"` runs everything that needs no device (oracle pinned against the
reference's golden vectors, host logic, ABI surface); `-m gpu` runs the parity tests proper, which
call the CUDA path through the C ABI and check it against the oracle and the committed fixtures."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from __graft_entry__ import import_package  # noqa: E402

import_package()
from nori_ray_tracer_b200 import abi, nscene  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def meta():
    return json.load(open(os.path.join(GOLDEN, "meta.json")))


def load_golden_scene(name):
    return nscene.load_scene(os.path.join(GOLDEN, f"{name}.nscene"))


@pytest.fixture(scope="session")
def golden_scene():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden_scene(name)
        return cache[name]
    return get


@pytest.fixture(scope="session")
def make_oracle():
    from oracle_binding import Oracle
    return lambda scene: Oracle(scene, abi)


@pytest.fixture(scope="session")
def gpu():
    from nori_ray_tracer_b200.gpu import NoriGpu
    g = NoriGpu(0)          # raises loudly when the library or the device is missing: no CPU fallback
    yield g
    g.close()


@pytest.fixture(autouse=True)
def _restore_gpu_options(request):
    """Tests change scheduling options on the session-scoped context; whatever a test does (or wherever it
    fails), the next one starts from the library's defaults."""
    yield
    if "gpu" in request.fixturenames:
        request.getfixturevalue("gpu").set_option("reset_options", 0)


SCENE_NAMES = sorted(json.load(open(os.path.join(GOLDEN, "meta.json")))["scenes"]) if os.path.exists(
    os.path.join(GOLDEN, "meta.json")) else []


def downsample(img, f=16):
    """f x f box filter (the image tolerance of SURVEY 8(d) is defined on 16x16-downsampled images)."""
    h, w = (img.shape[0] // f) * f, (img.shape[1] // f) * f
    return img[:h, :w].reshape(h // f, f, w // f, f, -1).mean((1, 3))


def rel_mse(a, b, f=16):
    a, b = downsample(a, f), downsample(b, f)
    return float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))


def students_t_accept(values, reference, significance, num_tests):
    """hypothesis::students_t_test (ext/hypothesis/hypothesis.h:313-345) on luminance samples."""
    from scipy import stats
    n = len(values)
    mean, var = float(np.mean(values, dtype=np.float64)), float(np.var(values, ddof=1, dtype=np.float64))
    t = abs(mean - reference) * np.sqrt(n / max(var, 1e-5))
    pval = 2 * (1 - stats.t.cdf(t, n - 1))
    alpha = 1.0 - (1.0 - significance) ** (1.0 / num_tests)
    return bool(np.isfinite(pval) and pval >= alpha), mean, pval


def luminance(rgb):
    return rgb[..., 0] * 0.212671 + rgb[..., 1] * 0.715160 + rgb[..., 2] * 0.072169   # common.cpp:233-235
