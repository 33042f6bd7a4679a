"""The GPU BVH builder (nori_gpu_build_bvh_device): its output is a tree in the reference's own format, so it
is checked like one -- structurally on the host, and by tracing it on the GPU against the oracle walking the same tree
(bit-exact, counters included) and against the reference-identical SAH tree (same closest hits)."""
import numpy as np
import pytest

from conftest import load_golden_scene
from nori_ray_tracer_b200 import abi, host_scene

pytestmark = pytest.mark.gpu


def _check_structure(sc):
    nodes = sc.nodes; boxes = nodes.view(np.float32)
    n_prims = sc.indices.size
    seen = np.zeros(n_prims, np.int32)
    stack = [0]; visited = 0
    while stack:
        i = stack.pop(); visited += 1
        w0, w1 = int(nodes[i, 0]), int(nodes[i, 1])
        if w0 & 1:
            size = w0 >> 1
            assert 1 <= size <= 63 and w1 + size <= n_prims
            seen[w1:w1 + size] += 1
        else:
            assert (w0 >> 1) in (0, 1, 2)                          # split axis
            for c in (i + 1, w1):                                  # left child right behind its parent
                assert i < c < len(nodes)
                assert (boxes[c, 2:5] >= boxes[i, 2:5]).all() and (boxes[c, 5:8] <= boxes[i, 5:8]).all()
                stack.append(c)
    assert visited == len(nodes) and (seen == 1).all()
    assert sorted(sc.indices.tolist()) == list(range(n_prims))     # a permutation of the global primitive indices


def _rays(sc, n, seed):
    rng = np.random.RandomState(seed)
    b = sc.nodes.view(np.float32)[0]
    lo, hi = b[2:5], b[5:8]
    rays = np.zeros(n, abi.RAY_DTYPE)
    rays["o"] = lo + (hi - lo) * (rng.rand(n, 3).astype(np.float32) * 1.4 - 0.2)
    d = rng.randn(n, 3).astype(np.float32)
    rays["d"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays["mint"], rays["maxt"] = np.float32(1e-4), np.float32(np.inf)
    # the special cases of the slab test (bbox.h:340-346): zero (+0 / -0) and subnormal direction components,
    # origins exactly on a bounding plane -- the kernels route these rays around their branch-free box test
    k = n // 8
    axis = rng.randint(0, 3, k)
    rays["d"][np.arange(k), axis] = rng.choice(np.array([0.0, -0.0, 1e-41, -1e-41, 1e-39], np.float32), k)
    on = rng.rand(k) < 0.5
    rays["o"][np.arange(k)[on], axis[on]] = np.where(rng.rand(int(on.sum())) < 0.5, lo[axis[on]], hi[axis[on]])
    return rays


@pytest.mark.parametrize("name,leaf", [("table_path_mis", 4), ("cbox_path_mis", 1), ("veach_mis", 2), ("sphere_mesh_normals", 8)])
def test_lbvh_is_a_valid_reference_format_tree_and_traces_like_one(name, leaf, gpu, make_oracle):
    sah = load_golden_scene(name)
    lb, ms = host_scene.rebuild_bvh(sah, "lbvh", leaf_size=leaf)
    assert ms is not None and ms >= 0
    _check_structure(lb)
    rays = _rays(lb, 20000, 7)
    gpu.upload_scene(lb)
    gpu.set_option("order", 0)
    got = gpu.trace(rays, 0)
    want = make_oracle(lb).trace(rays, 0)                          # the oracle walking the SAME tree
    for f in ("t", "u", "v", "shape", "prim", "nodes_visited", "prims_tested"):
        assert np.array_equal(got[f], want[f]), (name, f)
    sh = gpu.trace(rays, 1)
    gpu.upload_scene(sah)
    ref = gpu.trace(rays, 0)                                       # the reference-identical SAH tree
    gpu.set_option("order", 2)
    p = slice(len(rays) // 8, None)                                # not the special rays of _rays: a NaN in the slab test
    #                                                                (origin on a plane, 1/d infinite) rejects the box, whichever it is
    assert np.array_equal(got["t"][p], ref["t"][p])                # the closest hit distance does not depend on the tree
    same = (got["prim"][p] == ref["prim"][p]) & (got["shape"][p] == ref["shape"][p])
    assert same.mean() > 0.9999                                    # only exact ties may pick another primitive
    assert np.array_equal(np.isinf(sh["t"][p]), (np.isinf(ref["t"]) | (ref["t"] > rays["maxt"]))[p])


def test_lbvh_render_matches_sah_render(gpu):
    sah = load_golden_scene("table_path_mis")
    lb, _ = host_scene.rebuild_bvh(sah, "lbvh", leaf_size=4)
    imgs = []
    for sc in (sah, lb):
        gpu.upload_scene(sc)
        got = gpu.render_samples(0, 2, seed=4)
        imgs.append(got)
    differ = (np.abs(imgs[0] - imgs[1]).max(-1) > 0).mean()        # same paths unless a tie resolved differently somewhere
    assert differ < 2e-3, differ


def _soup(seed, n_tri=3000, degenerate=True):
    """A seeded triangle soup with the cases real meshes contain: slivers, zero-area and duplicated triangles,
    axis-aligned (flat-box) triangles, very different scales, plus a few spheres."""
    rng = np.random.RandomState(seed)
    c = rng.rand(n_tri, 1, 3).astype(np.float32) * 4 - 2
    V = (c + (rng.rand(n_tri, 3, 3).astype(np.float32) - 0.5) * rng.choice([0.02, 0.2, 1.0], (n_tri, 1, 1)).astype(np.float32))
    if degenerate:
        V[0:20, 2] = V[0:20, 1]                                   # zero area: two equal vertices
        V[20:40, :, 2] = np.float32(0.5)                          # axis-aligned: flat bounding boxes
        V[40:60] = V[60:80]                                       # exact duplicates (ties at equal t)
        V[80:90, 1] = V[80:90, 0] + np.float32(1e-7)              # slivers
    V = V.reshape(-1, 3)
    F = np.arange(3 * n_tri, dtype=np.uint32).reshape(-1, 3)
    sb = host_scene.SceneBuilder("normals")
    d = sb.diffuse((0.5, 0.5, 0.5))
    sb.add_mesh(V[: 3 * (n_tri // 2)], F[: n_tri // 2], d)
    sb.add_mesh(V[3 * (n_tri // 2):], F[: n_tri - n_tri // 2], d)
    for k in range(5):
        sb.add_sphere(rng.rand(3) * 3 - 1.5, 0.1 + 0.3 * rng.rand(), d)
    sb.perspective(64, 48, 40.0, origin=(0, -6, 1), target=(0, 0, 0), up=(0, 0, 1))
    return sb


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_soups_trace_like_the_oracle(seed, gpu, make_oracle):
    """Host SAH builder and GPU LBVH builder on random geometry incl. degenerate triangles: the GPU traversal
    equals the oracle's on the same tree bit for bit (reference child order: counters too; near child first: hits),
    and both trees give the same closest distances."""
    sb = _soup(seed)
    sah = sb.build()
    lb, _ = host_scene.rebuild_bvh(sah, "lbvh", leaf_size=1 + seed)
    _check_structure(lb)
    rays = _rays(sah, 20000, 100 + seed)
    ts = []
    for sc in (sah, lb):
        o = make_oracle(sc)
        gpu.upload_scene(sc)
        for order in (0, 1):
            gpu.set_option("order", order)
            for shadow in (0, 1):
                got, want = gpu.trace(rays, shadow), o.trace(rays, shadow)
                fields = ("t",) if shadow else ("t", "u", "v", "shape", "prim")
                if order == 0:
                    fields += ("nodes_visited", "prims_tested")
                for f in fields:
                    assert np.array_equal(got[f], want[f]), (seed, order, shadow, f, int((got[f] != want[f]).sum()))
                if not shadow and order == 0:
                    ts.append(got["t"])
        gpu.set_option("order", 2)
    assert np.array_equal(ts[0][len(rays) // 8:], ts[1][len(rays) // 8:])      # (plain rays only, see above)
