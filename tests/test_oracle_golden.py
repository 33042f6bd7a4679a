"""The oracle (oracle/nori_oracle.cpp) against everything the reference pins for this path:
pcg32 known answers, ray batches answered by the reference's own BVH::rayIntersect, the reference's
per-plugin eval/pdf/sample answers, the reference's sample sequence, whole renders of the reference
binary (pixel by pixel, thanks to the block-sequential RNG mode) and the reference's t-test values."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, SCENE_NAMES, luminance, students_t_accept


def test_pcg32_known_answers(meta, golden_scene, make_oracle):
    demo = meta["pcg32_demo"]                      # ext/pcg32/pcg32-demo.out:8
    o = make_oracle(golden_scene("cbox_path_mis"))
    got = o.pcg32_uint(demo["initstate"], demo["initseq"], len(demo["uint"]))
    assert got.tolist() == demo["uint"]
    f = o.pcg32(demo["initstate"], demo["initseq"], 4)
    expect = ((np.array(demo["uint"][:4], np.uint32) >> 9) | 0x3f800000).view(np.float32) - 1.0
    assert np.array_equal(f, expect)               # pcg32.h:101-110


@pytest.mark.parametrize("name", SCENE_NAMES)
def test_sample_sequence_bit_exact(name, golden_scene, make_oracle):
    """First samples of block (0,0) in the reference's own order: pixel sample positions and radiance."""
    sc = golden_scene(name)
    ref = sc.entries["seq"]
    mine = make_oracle(sc).block_sequence(len(ref))
    assert np.array_equal(ref, mine)


@pytest.mark.parametrize("name", [n for n in SCENE_NAMES])
def test_ray_batches_bit_exact(name, golden_scene, make_oracle):
    sc = golden_scene(name)
    rb = sc.ray_batch()
    if rb is None or len(rb["rays"]) == 0:
        pytest.skip("fixture carries no ray batch")
    o = make_oracle(sc)
    for shadow in (0, 1):
        m = rb["shadow"] == shadow
        if not m.any():
            continue
        hits, p, uv, n, ng = o.trace(rb["rays"][m], shadow, full=True)
        ref = rb["hits"][m]
        for f in ("t", "u", "v", "shape", "prim", "nodes_visited", "prims_tested"):
            assert np.array_equal(hits[f], ref[f]), (name, shadow, f)
        if not shadow:
            hit = ref["shape"] != 0xffffffff
            assert np.array_equal(p[hit], rb["p"][m][hit])
            assert np.array_equal(n[hit], rb["n"][m][hit])
            assert np.array_equal(ng[hit], rb["ng"][m][hit])
            assert np.allclose(uv[hit], rb["uv"][m][hit], rtol=0, atol=2e-7)


@pytest.mark.parametrize("name", ["cbox_path_mis", "sphere_mesh_normals", "veach_mis", "odyssey_mis"])
def test_special_case_rays_bit_exact(name, golden_scene, make_oracle):
    """The special cases of the slab test (bbox.h:343-357) -- a zero (+0 / -0) or subnormal direction component,
    the origin exactly on a bounding plane ((bound - o) * (1/d) = 0 * inf = NaN rejects that box) -- answered by the
    reference's BVH::rayIntersect (tests/golden/make_special_rays.py) on the tree of that export."""
    from nori_ray_tracer_b200 import abi, nscene
    fx = np.load(os.path.join(GOLDEN, f"special_rays_{name}.npz"))
    entries = dict(golden_scene(name).entries)
    entries["bvh.nodes"], entries["bvh.indices"] = fx["nodes"], fx["indices"]      # same geometry, the export's own tree
    o = make_oracle(nscene.SceneData(entries))
    rays = np.ascontiguousarray(fx["rays"]).view(abi.RAY_DTYPE).reshape(-1)
    ref = np.ascontiguousarray(fx["hits"]).view(abi.HIT_DTYPE).reshape(-1)
    d = rays["d"]
    special = ((d == 0) | (np.abs(d) < np.float32(1.1754944e-38))).any(axis=1)
    assert special.sum() > len(rays) // 3
    for shadow in (0, 1):
        m = fx["shadow"] == shadow
        hits = o.trace(rays[m], shadow)
        for f in ("t", "u", "v", "shape", "prim", "nodes_visited", "prims_tested"):
            assert np.array_equal(hits[f], ref[m][f]), (name, shadow, f, int((hits[f] != ref[m][f]).sum()))


@pytest.mark.parametrize("name", SCENE_NAMES)
def test_plugin_probes_bit_exact(name, golden_scene, make_oracle):
    """BSDF::eval/pdf/sample and Emitter::sample/pdf/eval of every plugin instance in the scene."""
    sc = golden_scene(name)
    o = make_oracle(sc)
    for b in range(sc.pod.n_bsdfs):
        ref = sc.entries[f"probe.bsdf.{b}.out"]
        got = o.bsdf_probe(b, sc.entries[f"probe.bsdf.{b}.in"])
        assert np.array_equal(ref, got, equal_nan=True), (name, "bsdf", b)
    for e in range(sc.pod.n_emitters):
        ref = sc.entries[f"probe.emitter.{e}.out"]
        got = o.emitter_probe(e, sc.entries[f"probe.emitter.{e}.in"])
        assert np.array_equal(ref, got, equal_nan=True), (name, "emitter", e)


@pytest.mark.parametrize("name", SCENE_NAMES)
def test_render_matches_reference_binary(name, meta, golden_scene, make_oracle):
    """Oracle render with the reference's sampler mapping vs the image nori_ref wrote: every pixel."""
    sc = golden_scene(name)
    spp = meta["scenes"][name]["ref_spp"][0]
    o = make_oracle(sc)
    img = o.resolve(o.render(0, spp, mode=1))
    ref = np.load(os.path.join(GOLDEN, f"{name}.ref{spp}.npy"))
    rel = np.abs(img - ref) / (np.abs(ref) + 1e-3)
    # the only tolerated difference is float summation order in the block merge (block.cpp:124-133)
    assert rel.max() < 1e-4, (name, float(rel.max()))


def _ttest_cases(meta_path=os.path.join(GOLDEN, "meta.json")):
    import json
    if not os.path.exists(meta_path):
        return []
    m = json.load(open(meta_path))["ttests"]
    return [(k, i) for k in sorted(m) for i in range(len(m[k]["scenes"]))]


@pytest.mark.parametrize("test,idx", _ttest_cases())
def test_reference_ttests(test, idx, meta, make_oracle):
    """scenes/pa*/tests/*.xml known answers, accepted at the reference's own significance level."""
    from nori_ray_tracer_b200 import nscene
    t = meta["ttests"][test]
    sc = nscene.load_scene(os.path.join(GOLDEN, t["scenes"][idx]))
    o = make_oracle(sc)
    n = t["sampleCount"]
    assert sc.width == 1 and sc.height == 1
    vals = o.render_samples(0, n, seed=1234)[:, 0, 0, :3]
    ok, mean, pval = students_t_accept(luminance(vals.astype(np.float64)), t["references"][idx],
                                       t["significance"], len(t["references"]))
    assert ok, (test, idx, mean, t["references"][idx], pval)


@pytest.mark.parametrize("name", ["cbox_path_mis", "sphere_mesh_normals"])
def test_variance_output_matches_reference_binary(name, golden_scene, make_oracle):
    """<scene>_variance.exr of nori_ref (render.cpp:263-278) vs the oracle, pixel by pixel."""
    sc = golden_scene(name)
    film, var = make_oracle(sc).render_with_variance(4, mode=1)
    ref = np.load(os.path.join(GOLDEN, f"{name}.refvar4.npy"))
    # var = E[m^2] - E[m]^2 cancels catastrophically in fp32: tolerance relative to the largest m^2
    scale = max(np.abs(ref).max(), 1e-6)
    assert np.abs(var - ref).max() < 2e-5 * scale + 1e-6


def _aimed_rays(sc, n, seed, aim):
    """Rays from random origins towards random points of the scene box ("box"), or exactly at mesh vertices and at
    points on mesh edges ("edges"): the only place where two primitives of different leaves are hit at distances one
    ulp apart."""
    from nori_ray_tracer_b200 import abi
    rng = np.random.RandomState(seed)
    b = sc.nodes.view(np.float32)[0]
    lo, hi = b[2:5], b[5:8]
    rays = np.zeros(n, abi.RAY_DTYPE)
    o = (lo + (hi - lo) * (rng.rand(n, 3) * 1.2 - 0.1)).astype(np.float32)
    tgt = (lo + (hi - lo) * rng.rand(n, 3)).astype(np.float32)
    if aim == "edges":
        V = np.asarray(sc.entries["shape.0.V"], np.float32).reshape(-1, 3)
        F = np.asarray(sc.entries["shape.0.F"]).reshape(-1, 3)
        tgt[: n // 2] = V[rng.randint(0, len(V), n // 2)]
        tri = F[rng.randint(0, len(F), n - n // 2)]
        s = rng.rand(n - n // 2, 1).astype(np.float32)
        tgt[n // 2:] = V[tri[:, 0]] + (V[tri[:, 1]] - V[tri[:, 0]]) * s
    d = tgt - o
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-20)
    rays["o"], rays["d"] = o, d.astype(np.float32)
    rays["mint"], rays["maxt"] = np.float32(1e-4), np.float32(np.inf)
    return rays


@pytest.mark.parametrize("name", ["table_path_mis", "veach_mis", "sphere_mesh_normals", "cbox_path_mis"])
def test_closest_hit_does_not_depend_on_the_visiting_order(name, golden_scene, make_oracle):
    """What lets the GPU kernels visit the near child first on large scenes ("What same hits in another
    visiting order rests on"): with the tie rule, the reference-order answer is reproduced by the near-first order, by
    the far-first order and by right-child-first.  Generic rays: identical, every field.  Rays aimed exactly at
    vertices and edges expose the one gap -- two primitives of different leaves hit at the same distance (exactly, or
    up to the conditioning of the triangle test), where the distance cull (bvh.cpp:423) decides by rounding which of
    them is ever tested: a few percent even of those rays, and both answers are the same point on the shared edge."""
    sc = golden_scene(name)
    o = make_oracle(sc)
    rays = _aimed_rays(sc, 300000, 5, "box")
    ref = o.trace(rays, 0)
    for order in (1, 2, 3):
        got = o.trace_ordered(rays, order)
        for f in ("t", "u", "v", "shape", "prim"):
            assert np.array_equal(got[f], ref[f]), (name, order, f)
    rays = _aimed_rays(sc, 100000, 6, "edges")
    ref = o.trace(rays, 0)
    for order in (1, 2, 3):
        got = o.trace_ordered(rays, order)
        bad = np.zeros(len(rays), bool)
        for f in ("t", "u", "v", "shape", "prim"):
            bad |= got[f] != ref[f]
        assert bad.mean() < 6e-2, (name, order, float(bad.mean()))
        both = np.isfinite(ref["t"][bad]) & np.isfinite(got["t"][bad])
        assert both.all()                                           # never a hit against a miss
        assert (np.abs(got["t"][bad] - ref["t"][bad]) <= 1e-3 * ref["t"][bad]).all()
