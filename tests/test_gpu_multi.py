"""nori_gpu_init_multi: one context over several devices of the node ("multi-GPU inside render").  The
scene is replicated, nori_gpu_render shards the sample indices and sums the films onto devices[0] with one kernel that
reads the peers' accumulation buffers over NVLink.  The two-device cases need `gpurun --gpus 2`; with one device only the
single-entry device list runs."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from nori_ray_tracer_b200 import nscene
from nori_ray_tracer_b200.gpu import NoriGpu, NoriGpuError

pytestmark = pytest.mark.gpu


def _n_devices():
    import torch
    return torch.cuda.device_count()


def _scene(name="cbox_path_mis"):
    return nscene.load_scene(os.path.join(GOLDEN, f"{name}.nscene"))


def test_single_entry_device_list_equals_plain_context(gpu):
    sc = _scene()
    gpu.upload_scene(sc); gpu.clear_film(); gpu.render(3, 6, seed=4)
    want = gpu.download_film()
    m = NoriGpu(devices=[0])
    try:
        m.upload_scene(sc); m.render(3, 6, seed=4)
        assert np.array_equal(m.download_film(), want)
        assert m.stats().devices == 1
    finally:
        m.close()
    with pytest.raises(NoriGpuError):
        NoriGpu(devices=[0, 0])
    with pytest.raises(NoriGpuError):
        NoriGpu(devices=[])


@pytest.mark.parametrize("name", ["cbox_path_mis", "table_path_mis", "c5_volumetric"])
def test_sharded_render_equals_single_device_render(name, gpu):
    n = _n_devices()
    if n < 2:
        pytest.skip("needs two devices")
    sc = _scene(name)
    spp = 11                                                    # not a multiple of the device count: uneven shards
    gpu.upload_scene(sc); gpu.clear_film(); gpu.render(2, spp, seed=9)
    want = gpu.download_film(); s1 = gpu.stats()
    m = NoriGpu(devices=list(range(min(n, 4))))
    try:
        m.upload_scene(sc)
        m.set_option("pool", 1 << 18)
        m.reset_stats()
        m.render(2, spp, seed=9)
        got = m.download_film(); st = m.stats()
        # same paths, other summation order: fp32 round-off of the film sums
        assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max(), name
        assert np.array_equal(got[..., 3] > 0, want[..., 3] > 0)
        assert st.devices == min(n, 4) and st.samples == sc.width * sc.height * spp and st.reduce_ms > 0
        # a second call accumulates on top (the peers' buffers were handed back zeroed)
        m.render(2 + spp, 5, seed=9)
        gpu.render(2 + spp, 5, seed=9)
        assert np.abs(m.download_film() - gpu.download_film()).max() <= 2e-5 * np.abs(want).max()
        m.clear_film()
        assert not m.download_film().any()
        with pytest.raises(NoriGpuError):
            m.set_option("variance", 1)
    finally:
        m.close()


def test_reference_front_end_on_two_devices(tmp_path):
    import shutil
    import subprocess
    from conftest import ROOT, rel_mse
    exe = os.path.join(ROOT, "oracle", "_ref", "nori_ref_gpu")
    if _n_devices() < 2 or not os.path.exists(exe):
        pytest.skip("needs two devices and oracle/_ref/nori_ref_gpu")
    d = os.path.join(str(tmp_path), "cbox")
    shutil.copytree(os.path.join(GOLDEN, "scenes", "cbox"), d)
    out = subprocess.run([exe, os.path.join(d, "cbox_path_mis.xml"), "--devices", "0,1", "--chunk", "64"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "2 device(s)" in out.stdout, (out.stdout[-1000:], out.stderr[-1000:])
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    import cv2
    img = np.ascontiguousarray(cv2.imread(os.path.join(d, "cbox_path_mis.exr"), cv2.IMREAD_UNCHANGED)[..., 2::-1], np.float32)
    ref = np.load(os.path.join(GOLDEN, "cbox_path_mis.ref128.npy"))
    assert rel_mse(img, ref) < 1e-3
