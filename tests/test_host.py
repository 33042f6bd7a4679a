"""Host-side logic: the scene container, the RenderThread mirror's sharding/resolve helpers, EXR I/O."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from nori_ray_tracer_b200 import abi, imageio, nscene, render


def test_container_round_trip(tmp_path):
    e = nscene.read_container(os.path.join(GOLDEN, "cbox_path_mis.nscene"))
    p = tmp_path / "copy.nscene"
    nscene.write_container(str(p), e)
    e2 = nscene.read_container(str(p))
    assert list(e) == list(e2)
    for k in e:
        assert e[k].dtype == e2[k].dtype and np.array_equal(e[k], e2[k], equal_nan=True)


def test_scene_pod(golden_scene):
    sc = golden_scene("cbox_path_mis")
    assert (sc.width, sc.height, sc.border) == (200, 150, 2)         # gaussian r=2 -> border 2 (block.cpp:57)
    assert sc.film_shape == (154, 204, 4)
    assert sc.pod.n_indices == 14 and sc.pod.n_shapes == 6 and sc.pod.n_emitters == 1
    assert sc.pod.integrator == abi.INTEGRATOR_PATH_MIS
    types = sorted(sc.bsdfs[i].type for i in range(sc.pod.n_bsdfs))
    assert abi.BSDF_MIRROR in types and abi.BSDF_DIELECTRIC in types and abi.BSDF_DIFFUSE in types
    # gaussian table as ImageBlock::init tabulates it (block.cpp:59-63)
    t = np.array(sc.pod.filter.table[:])
    assert t[32] == 0 and t[0] > t[16] > t[31] >= 0 and sc.pod.filter.radius == 2.0


def test_set_resolution_keeps_aspect(golden_scene):
    sc = nscene.load_scene(os.path.join(GOLDEN, "cbox_path_mis.nscene"))
    sc.set_resolution(800, 600)
    assert sc.film_shape == (604, 804, 4)
    assert sc.pod.camera.invOutputSize[0] == np.float32(1) / np.float32(800)
    with pytest.raises(ValueError):
        sc.set_resolution(800, 800)


@pytest.mark.parametrize("spp,world", [(1024, 8), (1024, 3), (5, 8), (0, 4), (7, 1)])
def test_shard_spp_partitions_exactly(spp, world):
    ranges = [render.shard_spp(spp, r, world) for r in range(world)]
    covered = []
    for b, n in ranges:
        assert n >= 0
        covered += list(range(b, b + n))
    assert covered == list(range(spp))                       # disjoint, ordered, complete
    assert max(n for _, n in ranges) - min(n for _, n in ranges) <= 1


def test_resolve_film_divides_by_weight():
    film = np.zeros((8, 9, 4), np.float32)
    film[2:6, 2:7] = (2.0, 4.0, 6.0, 2.0)
    film[3, 3] = (1.0, 1.0, 1.0, 0.0)                       # zero weight -> black (color.h:84-89)
    rgb = render.resolve_film(film, 2)
    assert rgb.shape == (4, 5, 3)
    assert np.allclose(rgb[0, 0], (1, 2, 3)) and np.all(rgb[1, 1] == 0)


def test_exr_round_trip(tmp_path):
    img = np.random.RandomState(0).rand(7, 5, 3).astype(np.float32)
    p = str(tmp_path / "a.exr")
    imageio.write_exr(p, img)
    assert np.array_equal(imageio.read_exr(p), img)
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    cv2 = pytest.importorskip("cv2")
    back = cv2.imread(p, cv2.IMREAD_UNCHANGED)
    if back is not None:                                    # cv2 builds without EXR support return None
        assert np.array_equal(back[..., ::-1], img)


def test_set_film_rebuilds_the_camera_matrix_of_the_fixture():
    """SceneData.set_film (any aspect ratio) rebuilds sampleToCamera like perspective.cpp:53-80: at the fixture's own size it
    reproduces the matrix the reference computed, and a 16:9 film keeps the horizontal field of view."""
    import numpy as np
    from conftest import load_golden_scene
    sc = load_golden_scene("c5_volumetric")
    before = np.array(list(sc.pod.camera.sampleToCamera), np.float32)
    w, h = sc.width, sc.height
    sc.set_film(w, h)
    after = np.array(list(sc.pod.camera.sampleToCamera), np.float32)
    assert np.abs(before - after).max() < 1e-6
    sc.set_film(3840, 2160)
    wide = np.array(list(sc.pod.camera.sampleToCamera), np.float32).reshape(4, 4)
    assert (sc.width, sc.height) == (3840, 2160)
    assert abs(wide[0, 0] - before.reshape(4, 4)[0, 0]) < 1e-6            # 2 / cot(fov / 2): the horizontal extent
    assert abs(wide[1, 1] * (3840 / 2160) - wide[0, 0] * -1) < 1e-5       # vertical extent scales with the aspect ratio
