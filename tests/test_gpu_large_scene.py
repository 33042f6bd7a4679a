"""Parity of the LARGE-SCENE kernels with the oracle, on their own scene.

The 10 M-triangle height field of BASELINE config 4 (and a 1 M-triangle one) is rendered by k_extend_sm /
k_shadow_sm (wave_extend.cu): the warp state machine over the reference nodes (order 0), over child-box pairs
(order 1, wide 0) or over 4-wide records (order 1, wide 1 = what produces the reported C4 numbers).  Option
"trace_kernel" = 2 routes nori_gpu_trace through exactly those kernels, so the oracle (reference order,
bvh.cpp:404-462, mesh.cpp:83-120) can answer the same ray batches:

  * camera rays, hemisphere secondaries from their hit points, NEE-style shadow rays, axis-parallel rays (zero
    direction components: the rayPlain() fallback of the 4-wide kernels): bit-exact (t, u, v, shape, prim) in every
    layout; in reference order also the node-visit / primitive-test totals;
  * rays aimed at height-field vertices and edges, also at grazing angles -- a height field consists of shared edges
    only, and there the near-child-first answer can depend on the order through the distance cull: bit-exact in
    every layout, because the order guard (traverse.cuh) answers those rays again in the reference's order;
  * a hand-made 60-level chain tree: every layout's traversal stack spills out of shared memory (LaneStack::ovf /
    LaneStack2::ovf) and still returns the oracle's answers;
  * leaves with more than 63 primitives: the compact layouts are not built and the kernels walk the reference nodes;
  * a whole wavefront render of the field against the oracle, sample by sample.
"""
import numpy as np
import pytest

from nori_ray_tracer_b200 import abi, host_scene, nscene

pytestmark = pytest.mark.gpu

FIELDS = {}
# (order, wide): reference order over the reference nodes / near child first over child-box pairs / over 4-wide records
LAYOUTS = [(0, 0), (1, 0), (1, 1)]
HIT_FIELDS = ("t", "u", "v", "shape", "prim")


def field(n):
    if n not in FIELDS:
        FIELDS[n] = host_scene.heightfield_scene(n=n, width=960, height=540)
    return FIELDS[n]


def _rays(o, d, mint=1e-4, maxt=np.inf):
    r = np.zeros(len(o), abi.RAY_DTYPE)
    r["o"], r["d"] = np.asarray(o, np.float32), np.asarray(d, np.float32)
    r["mint"], r["maxt"] = np.float32(mint), np.asarray(maxt, np.float32)
    return r


def _unit(v):
    v = np.asarray(v, np.float32)
    return (v / np.linalg.norm(v, axis=1, keepdims=True).astype(np.float32)).astype(np.float32)


def camera_rays(n_rays, rng):
    """Rays from the scene's camera position towards the field (what the first wavefront iteration traces)."""
    eye = np.array([0.0, -2.2, 1.6], np.float32)
    tgt = np.stack([rng.uniform(-1.1, 1.1, n_rays), rng.uniform(-1.1, 1.1, n_rays), rng.uniform(0.0, 0.3, n_rays)], -1).astype(np.float32)
    d = _unit(tgt - eye)
    return _rays(np.broadcast_to(eye, d.shape), d, mint=np.float32(1e-4), maxt=np.float32(1e4))


def secondary_rays(hits, prim_rays, rng):
    """Hemisphere rays from the hit points: incoherent, many of them graze the surface."""
    m = hits["prim"] != 0xFFFFFFFF
    p = (prim_rays["o"][m] + hits["t"][m, None] * prim_rays["d"][m]).astype(np.float32)
    d = rng.randn(len(p), 3).astype(np.float32)
    d[:, 2] = np.abs(d[:, 2]) * rng.choice([1.0, 1.0, 1.0, 0.05], len(p)).astype(np.float32)
    return _rays(p, _unit(d)), p


def shadow_rays(p, rng):
    """NEE rays towards the light quad (z = 1.5, |x|, |y| <= 0.4): [Epsilon, dist - Epsilon] (arealight.cpp:56)."""
    q = np.stack([rng.uniform(-0.4, 0.4, len(p)), rng.uniform(-0.4, 0.4, len(p)), np.full(len(p), 1.5)], -1).astype(np.float32)
    v = (q - p).astype(np.float32)
    dist = np.linalg.norm(v, axis=1).astype(np.float32)
    return _rays(p, (v / dist[:, None]).astype(np.float32), maxt=(dist - np.float32(1e-4)).astype(np.float32))


def axis_parallel_rays(sc, n_rays, rng):
    """Straight down onto grid vertices (two zero direction components, x / y exactly a vertex's) and along the
    grid rows: rays outside rayPlain(), answered by the reference-node traversal inside the 4-wide kernels."""
    V = sc.entries["shape.0.V"]
    v = V[rng.randint(0, len(V), n_rays)]
    o = np.stack([v[:, 0], v[:, 1], np.full(n_rays, 1.0, np.float32)], -1).astype(np.float32)
    d = np.zeros((n_rays, 3), np.float32); d[:, 2] = -1
    k = n_rays // 3
    o[:k] = np.stack([np.full(k, -1.5, np.float32), v[:k, 1], v[:k, 2] + np.float32(0.01)], -1)
    d[:k] = (1, 0, 0)
    d[k:2 * k, 0] = rng.choice([-0.0, 0.0], k)                     # signed zeros (bbox.h:344-346 compares with == 0)
    return _rays(o, d)


def edge_rays(sc, n_rays, rng):
    """Rays aimed at mesh vertices and at points on mesh edges, from above: every hit lies on an edge shared by
    triangles that usually sit in different leaves."""
    V, F = sc.entries["shape.0.V"], sc.entries["shape.0.F"]
    f = F[rng.randint(0, len(F), n_rays)]
    a, b = V[f[:, 0]], V[f[:, 1]]
    s = rng.choice([0.0, 0.25, 0.5, 1.0], n_rays).astype(np.float32)[:, None]       # 0 / 1: a vertex
    tgt = (a + s * (b - a)).astype(np.float32)
    o = (tgt + np.stack([rng.uniform(-0.5, 0.5, n_rays), rng.uniform(-0.5, 0.5, n_rays), rng.uniform(0.3, 1.2, n_rays)], -1)).astype(np.float32)
    return _rays(o, _unit(tgt - o))


def _configure(gpu, order, wide, stats):
    gpu.set_option("trace_kernel", 2)
    gpu.set_option("order", order)
    gpu.set_option("wide", wide)
    gpu.set_option("stats", stats)
    gpu.set_option("pool", 1 << 18)


def _assert_same_hits(got, want, what):
    for f in HIT_FIELDS:
        assert np.array_equal(got[f], want[f], equal_nan=(f in "tuv")), (what, f, int((got[f] != want[f]).sum()))


@pytest.mark.parametrize("n", [708, 2237])
def test_large_scene_kernels_bit_exact_vs_oracle(n, gpu, make_oracle):
    sc = field(n)
    o = make_oracle(sc)
    rng = np.random.RandomState(n)
    cam = camera_rays(80000, rng)
    cam_ref = o.trace(cam, 0)
    sec, p = secondary_rays(cam_ref, cam, rng)
    axis = axis_parallel_rays(sc, 30000, rng)
    closest = np.concatenate([cam, sec, axis])
    closest_ref = o.trace(closest, 0)
    shadow = np.concatenate([shadow_rays(p, rng), secondary_rays(cam_ref, cam, rng)[0]])
    shadow_ref = o.trace(shadow, 1)
    assert len(closest) + len(shadow) >= 200000
    assert 0.3 < (closest_ref["prim"] != 0xFFFFFFFF).mean() < 0.95 and 0.02 < (shadow_ref["t"] == 0).mean() < 0.9
    gpu.upload_scene(sc)
    for order, wide in LAYOUTS:
        for stats in (0, 1):                                        # the kernels that render, and their counting variants
            _configure(gpu, order, wide, stats)
            gpu.reset_stats()
            _assert_same_hits(gpu.trace(closest, 0), closest_ref, (n, order, wide, stats, "closest"))
            got = gpu.trace(shadow, 1)
            assert np.array_equal(got["t"], shadow_ref["t"]), (n, order, wide, stats, "any-hit", int((got["t"] != shadow_ref["t"]).sum()))
            if stats:
                ks = gpu.kernel_stats()
                assert ks["extend"]["rays"] == len(closest) and ks["shadow"]["rays"] == len(shadow)
                if order == 0:                                      # reference order: the reference's own counters
                    assert ks["extend"]["nodes"] == int(closest_ref["nodes_visited"].sum())
                    assert ks["extend"]["prims"] == int(closest_ref["prims_tested"].sum())
                    assert ks["shadow"]["nodes"] == int(shadow_ref["nodes_visited"].sum())
                    assert ks["shadow"]["prims"] == int(shadow_ref["prims_tested"].sum())
                else:                                               # near child first: fewer boxes and primitives
                    assert ks["extend"]["prims"] < int(closest_ref["prims_tested"].sum())
    # the one-thread-per-ray hook over the same tree, both orders (per-ray counters in reference order)
    gpu.set_option("trace_kernel", 0)
    gpu.set_option("order", 0)
    got = gpu.trace(closest, 0)
    _assert_same_hits(got, closest_ref, (n, "k_trace"))
    assert np.array_equal(got["nodes_visited"], closest_ref["nodes_visited"]) and np.array_equal(got["prims_tested"], closest_ref["prims_tested"])
    gpu.set_option("order", 1)
    _assert_same_hits(gpu.trace(closest, 0), closest_ref, (n, "k_trace near-first"))


def grazing_edge_rays(sc, n_rays, rng, height):
    """The same targets from just above the surface: the triangle test is worst conditioned when the ray runs almost
    inside the triangle's plane, which is where the two distances of a shared edge differ most."""
    V, F = sc.entries["shape.0.V"], sc.entries["shape.0.F"]
    f = F[rng.randint(0, len(F), n_rays)]
    a, b = V[f[:, 0]], V[f[:, 1]]
    s = rng.choice([0.0, 0.25, 0.5, 1.0], n_rays).astype(np.float32)[:, None]
    tgt = (a + s * (b - a)).astype(np.float32)
    o = (tgt + np.stack([rng.uniform(-1, 1, n_rays), rng.uniform(-1, 1, n_rays), rng.uniform(height / 2, height, n_rays)], -1)).astype(np.float32)
    return _rays(o, _unit(tgt - o))


@pytest.mark.parametrize("n", [708, 2237])
def test_rays_aimed_at_shared_edges(n, gpu, make_oracle):
    """A height field consists of shared edges only.  Visiting the near child first, which of two triangles hit at
    (almost) the same distance is ever tested can depend on the order: a box is culled by a hit closer than its entry
    distance, and a primitive lying in a face of its box can be a few ulps in front of it (with the oracle walking near
    child first: 1.2 % of these rays return the other triangle of the edge).  The order guard (traverse.cuh) answers
    such rays again in the reference's order: every layout returns the reference's primitive on every ray."""
    sc = field(n)
    rng = np.random.RandomState(7 * n)
    rays = np.concatenate([edge_rays(sc, 200000, rng), grazing_edge_rays(sc, 60000, rng, 0.1), grazing_edge_rays(sc, 60000, rng, 0.02),
                           grazing_edge_rays(sc, 60000, rng, 0.005)])
    ref = make_oracle(sc).trace(rays, 0)
    assert (ref["prim"] != 0xFFFFFFFF).mean() > 0.9
    gpu.upload_scene(sc)
    for order, wide in LAYOUTS:
        _configure(gpu, order, wide, 1)
        gpu.reset_stats()
        _assert_same_hits(gpu.trace(rays, 0), ref, (n, "edges", order, wide))
        redo = gpu.stats().guard_retraces
        print(f"edge-aimed rays, n={n}, order={order}, wide={wide}: {redo} of {len(rays)} answered again in reference order")
        assert (redo == 0) if order == 0 else (0 < redo < 0.5 * len(rays))
    gpu.set_option("trace_kernel", 0); gpu.set_option("order", 1)       # the one-thread-per-ray traversal carries the same guard
    _assert_same_hits(gpu.trace(rays, 0), ref, (n, "edges", "k_trace near-first"))
    # generic rays rarely meet the guard (here: rays that cross one of the field's millimetre-sized ridges, whose front
    # and back faces are less than 2^-11 of the distance apart)
    cam = camera_rays(100000, rng)
    _configure(gpu, 1, 1, 1)
    gpu.reset_stats(); gpu.trace(cam, 0)
    print(f"camera rays, n={n}: {gpu.stats().guard_retraces} of {len(cam)} answered again in reference order")
    assert gpu.stats().guard_retraces < 2e-2 * len(cam)


def chain_tree_scene(levels=60):
    """`levels` parallel triangles stacked along x under a hand-made chain tree: inner node k = {left: inner node k+1,
    right: leaf of triangle k}, so a ray along +x descends the whole chain with one pending sibling per level --
    the deepest stack a valid tree (<= 64 levels, bvh.cpp:405) can ask for."""
    sb = host_scene.SceneBuilder("path_mis")
    N = levels
    xs = np.linspace(0, 1, N, dtype=np.float32)
    V = np.zeros((3 * N, 3), np.float32)
    for k in range(N):
        V[3 * k:3 * k + 3] = [(xs[k], -1, -1), (xs[k], 1, -1), (xs[k], 0, 1.5)]
    F = np.arange(3 * N, dtype=np.uint32).reshape(N, 3)
    light = sb.area_light((1, 1, 1))
    sb.add_mesh(V, F, sb.diffuse((0.5, 0.5, 0.5)), emitter=light)
    sb.perspective(64, 36, 40.0, origin=(-2, 0, 0), target=(0, 0, 0), up=(0, 0, 1))
    sc = sb.build()
    nodes = np.zeros((2 * N - 1, 8), np.uint32); nf = nodes.view(np.float32)
    tri_box = lambda k: (V[3 * k:3 * k + 3].min(0), V[3 * k:3 * k + 3].max(0))
    for k in range(N - 1):                                          # inner chain: node k, left = k + 1, right = leaf of triangle k
        nodes[k, 0] = 0 << 1                                        # split axis x
        nodes[k, 1] = (N - 1) + (N - 1 - k)
        nf[k, 2:5], nf[k, 5:8] = V[3 * k:].min(0), V[3 * k:].max(0)
    order = [N - 1] + list(range(N - 2, -1, -1))                    # leaves in depth-first order
    for j, k in enumerate(order):
        i = N - 1 + j
        nodes[i, 0] = (1 << 1) | 1; nodes[i, 1] = j
        nf[i, 2:5], nf[i, 5:8] = tri_box(k)
    e = dict(sc.entries)
    e["bvh.nodes"], e["bvh.indices"] = nodes, np.array(order, np.uint32)
    return nscene.SceneData(e)


def test_deep_tree_spills_the_traversal_stack(gpu, make_oracle):
    sc = chain_tree_scene(60)
    rng = np.random.RandomState(3)
    n_rays = 20000
    o = np.stack([np.full(n_rays, -0.5), rng.uniform(-0.8, 0.8, n_rays), rng.uniform(-0.9, 1.2, n_rays)], -1).astype(np.float32)
    d = _unit(np.stack([np.ones(n_rays), rng.uniform(-0.2, 0.2, n_rays), rng.uniform(-0.2, 0.2, n_rays)], -1))
    o[n_rays // 2:, 0] = 1.5; d[n_rays // 2:, 0] *= -1             # from the other end: the far child is the chain
    rays = _rays(o, d)
    far = rays.copy(); far["mint"] = np.float32(0.7)                # explicit mint: the nearest triangles are skipped
    rays = np.concatenate([rays, far])
    orc = make_oracle(sc)
    ref, sref = orc.trace(rays, 0), orc.trace(rays, 1)
    assert (ref["prim"] != 0xFFFFFFFF).mean() > 0.5
    gpu.upload_scene(sc)
    for order, wide in LAYOUTS:
        _configure(gpu, order, wide, 1)
        gpu.set_option("traversal", 2)
        gpu.reset_stats()
        _assert_same_hits(gpu.trace(rays, 0), ref, ("chain", order, wide))
        shadow = rays[rays["mint"] == np.float32(1e-4)]             # the any-hit kernel starts its rays at Epsilon
        assert np.array_equal(gpu.trace(shadow, 1)["t"], sref[rays["mint"] == np.float32(1e-4)]["t"])
        depth = gpu.stats().max_stack_depth
        shared = 32 if order == 0 else 16                           # NORI_SM_STACK / NORI_SM_STACK2 entries live in shared memory
        assert depth > shared, (order, wide, depth)
    # and a wavefront render through those kernels equals the one-thread-per-sample kernel bit for bit
    gpu.set_option("trace_kernel", 0)
    gpu.set_option("megakernel", 1)
    want = gpu.render_samples(0, 2, seed=5)
    gpu.set_option("megakernel", 0)
    for order, wide in LAYOUTS:
        gpu.set_option("order", order); gpu.set_option("wide", wide); gpu.set_option("traversal", 2); gpu.set_option("shadow_pass", 1)
        assert np.array_equal(gpu.render_samples(0, 2, seed=5), want, equal_nan=True), (order, wide)


def big_leaf_scene():
    """A 9x9-vertex height field (128 triangles) under a hand-made tree: root + two leaves of 70 and 58 triangles."""
    sc = host_scene.heightfield_scene(n=9, width=64, height=36)
    V, F = sc.entries["shape.0.V"], sc.entries["shape.0.F"]
    n_field = len(F)
    order = np.argsort(V[F].mean(1)[:, 0], kind="stable").astype(np.uint32)     # split the field by centroid x
    lightV = sc.entries["shape.1.V"]
    groups = [order[:70], np.concatenate([order[70:], np.arange(n_field, n_field + 2, dtype=np.uint32)])]   # + the light's two triangles
    pts = [V[F[groups[0]]].reshape(-1, 3), np.concatenate([V[F[groups[1][:-2]]].reshape(-1, 3), lightV])]
    nodes = np.zeros((3, 8), np.uint32); nf = nodes.view(np.float32)
    nodes[0, 0], nodes[0, 1] = 0 << 1, 2
    allp = np.concatenate(pts); nf[0, 2:5], nf[0, 5:8] = allp.min(0), allp.max(0)
    for k in (0, 1):
        nodes[1 + k, 0] = (len(groups[k]) << 1) | 1
        nodes[1 + k, 1] = 0 if k == 0 else len(groups[0])
        nf[1 + k, 2:5], nf[1 + k, 5:8] = pts[k].min(0), pts[k].max(0)
    e = dict(sc.entries)
    e["bvh.nodes"], e["bvh.indices"] = nodes, np.concatenate(groups).astype(np.uint32)
    return nscene.SceneData(e)


def test_leaves_beyond_the_compact_encoding_walk_the_reference_nodes(gpu, make_oracle):
    """A leaf reference of the pair / 4-wide layouts holds at most 63 primitives (host_layout.h): for a tree with
    larger leaves the layouts are not built and the state-machine kernels walk the reference's own nodes."""
    sc = big_leaf_scene()
    assert (sc.nodes[1:, 0] >> 1).max() > 63
    rng = np.random.RandomState(11)
    cam = camera_rays(20000, rng)
    ref = make_oracle(sc).trace(cam, 0)
    assert (ref["prim"] != 0xFFFFFFFF).mean() > 0.5
    gpu.upload_scene(sc)
    for order in (0, 1):
        _configure(gpu, order, 1, 1)
        gpu.set_option("traversal", 2)
        gpu.reset_stats()
        _assert_same_hits(gpu.trace(cam, 0), ref, ("big leaves", order))
    gpu.set_option("trace_kernel", 0); gpu.set_option("megakernel", 1)
    want = gpu.render_samples(0, 1, seed=2)
    gpu.set_option("megakernel", 0); gpu.set_option("shadow_pass", 1)
    assert np.array_equal(gpu.render_samples(0, 1, seed=2), want, equal_nan=True)


@pytest.mark.parametrize("order,wide", LAYOUTS)
def test_wavefront_render_of_the_field_vs_oracle(order, wide, gpu, make_oracle):
    """The 1 M-triangle field rendered by the wavefront (k_extend_sm + k_shade + k_shadow_sm), 960x540, 1 spp, against
    the oracle sample by sample (same per-path pcg32 streams).  Tolerance as in test_per_sample_radiance_vs_oracle."""
    sc = field(708)
    gpu.upload_scene(sc)
    gpu.set_option("order", order); gpu.set_option("wide", wide); gpu.set_option("pool", 1 << 18)
    gpu.reset_stats()
    got = gpu.render_samples(0, 1, seed=3)
    want = make_oracle(sc).render_samples(0, 1, seed=3)
    assert np.array_equal(got[..., 3], want[..., 3])
    rel = np.abs(got - want) / (np.abs(want) + 1e-3)
    assert float((rel.max(-1) > 1e-3).mean()) < 2e-3
    assert abs(got[..., :3].mean() - want[..., :3].mean()) < 2e-3 * want[..., :3].mean()
    s = gpu.stats()
    assert s.samples == 960 * 540 and 3.0 < s.rays / s.samples < 6.0
