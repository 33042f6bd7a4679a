"""The drop-in, compiled: oracle/_ref/nori_ref_gpu is the reference's own front-end (main_euler.cpp + the RenderThread of
render.cpp, linked unmodified from oracle/_ref/libnori_ref.a) with the spp loop of render.cpp:173-284 replaced by calls
into libnori_gpu.so through oracle/ref_tools/gpu_binding.h.

CPU: `--describe` flattens the scene the reference loaded and prints it; it must equal the committed fixture of the same
scene (written by nori_export, the other walk over the same reference objects).
GPU: the binary renders the Cornell box and the table scene from their XML files and writes the EXR like the reference;
the image must agree with the reference's own output (nori_ref render / the reference's shipped golden)."""
import json
import os
import shutil
import subprocess
import zlib

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, downsample, rel_mse
from nori_ray_tracer_b200 import abi, nscene

EXE = os.path.join(ROOT, "oracle", "_ref", "nori_ref_gpu")
needs_binary = pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/nori_ref_gpu is built where /root/reference exists (oracle/Makefile: ref)")


def _scene_copy(tmp_path, name):
    dst = os.path.join(str(tmp_path), name)
    shutil.copytree(os.path.join(GOLDEN, "scenes", name), dst)
    return dst


def _crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xffffffff


@needs_binary
def test_binding_flattens_the_reference_scene_like_the_fixture(tmp_path):
    d = _scene_copy(tmp_path, "cbox")
    out = subprocess.run([EXE, os.path.join(d, "cbox_path_mis.xml"), "--describe"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-2000:]
    got = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    sc = nscene.load_scene(os.path.join(GOLDEN, "cbox_path_mis.nscene"))
    e = sc.entries
    assert got["integrator"] == abi.INTEGRATOR_NAMES["path_mis"]
    assert (got["n_nodes"], got["n_indices"], got["n_shapes"], got["n_bsdfs"], got["n_emitters"], got["n_images"]) == \
           (sc.pod.n_nodes, sc.pod.n_indices, sc.pod.n_shapes, sc.pod.n_bsdfs, sc.pod.n_emitters, sc.pod.n_images)
    assert (got["width"], got["height"]) == (sc.width, sc.height)
    # the reference's arrays, byte for byte (14 primitives: the SAH build is deterministic at this size)
    assert got["nodes_crc"] == _crc(e["bvh.nodes"]) and got["indices_crc"] == _crc(e["bvh.indices"])
    assert got["shape_offset_crc"] == _crc(e["bvh.shape_offset"])
    import ctypes as C
    for k, t in (("bsdfs", None), ("camera", abi.Camera), ("filter", abi.Filter), ("medium", abi.Medium)):
        raw = e[f"{k}.pod"].tobytes()
        if t is not None:                                    # fixtures written before the struct grew end in zeros
            raw = raw.ljust(C.sizeof(t), b"\0")
        assert got[f"{k}_crc"] == zlib.crc32(raw) & 0xffffffff, k
    for i, s in enumerate(got["shapes"]):
        p = sc.shapes[i]
        assert (s["type"], s["bsdf"], s["emitter"], s["n_vertices"], s["n_triangles"]) == (p.type, p.bsdf, p.emitter, p.n_vertices, p.n_triangles)
        if p.type == abi.SHAPE_MESH:
            assert s["V_crc"] == _crc(e[f"shape.{i}.V"]) and s["F_crc"] == _crc(e[f"shape.{i}.F"]) and s["cdf_crc"] == _crc(e[f"shape.{i}.area_cdf"])
    assert [(q["type"], q["shape"]) for q in got["emitters"]] == [(sc.emitters[i].type, sc.emitters[i].shape) for i in range(sc.pod.n_emitters)]


def _read_exr(path):
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    import cv2
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    assert img is not None, path
    return np.ascontiguousarray(img[..., 2::-1], np.float32)


@needs_binary
@pytest.mark.gpu
def test_reference_front_end_renders_the_cornell_box_through_the_library(tmp_path):
    """cbox_path_mis.xml (200 x 150, 128 spp) -> cbox_path_mis.exr, against the image nori_ref wrote for the same file."""
    d = _scene_copy(tmp_path, "cbox")
    out = subprocess.run([EXE, os.path.join(d, "cbox_path_mis.xml"), "--chunk", "32"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-1500:])
    assert "Rendering .. done." in out.stdout
    img = _read_exr(os.path.join(d, "cbox_path_mis.exr"))
    ref = np.load(os.path.join(GOLDEN, "cbox_path_mis.ref128.npy"))
    assert img.shape == ref.shape
    assert rel_mse(img, ref) < 1e-3, rel_mse(img, ref)
    assert abs(img.mean() - ref.mean()) < 0.02 * ref.mean()


@needs_binary
@pytest.mark.gpu
def test_reference_front_end_renders_the_table_scene_through_the_library(tmp_path):
    """table_path_mis.xml as shipped (800 x 600, 512 spp; OBJ loading and the SAH build run in the reference's code)
    against the golden the reference ships for it (scenes/pa4/table/ref/table_path_mis_512spp.exr), and the per-pixel
    variance output next to it."""
    d = _scene_copy(tmp_path, "table")
    out = subprocess.run([EXE, os.path.join(d, "table_path_mis.xml"), "--chunk", "64", "--variance"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-1500:])
    img = _read_exr(os.path.join(d, "table_path_mis.exr"))
    gold = np.load(os.path.join(GOLDEN, "ref_goldens.npz"))["table_path_mis_512spp"]
    a = downsample(img, 16)
    assert a.shape == gold.shape
    err = float(np.mean((a - gold) ** 2 / (gold ** 2 + 1e-2)))
    assert err < 1e-3, err
    var = _read_exr(os.path.join(d, "table_path_mis_variance.exr"))
    assert var.shape == img.shape and np.isfinite(var).all() and var.mean() > 0


@needs_binary
@pytest.mark.gpu
def test_reference_front_end_reports_library_errors(tmp_path):
    """No exception crosses the ABI: a device index that does not exist comes back as an error string and exit code 2."""
    d = _scene_copy(tmp_path, "cbox")
    out = subprocess.run([EXE, os.path.join(d, "cbox_path_mis.xml"), "--devices", "99"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 2 and "device index out of range" in out.stderr
