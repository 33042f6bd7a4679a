"""The 4-wide node records nori_gpu_upload_scene derives from a reference-format tree (csrc/host_bvh.cpp,
host_layout.h; read by the large-scene kernels of wave_extend.cu), checked on the host through the
nori_gpu_wide_layout hook: structure, and the property the kernels rely on -- for rays that cannot meet a NaN in
the slab test, walking the records reaches exactly the leaves the reference's traversal (bvh.cpp:404-462 without
the distance cull) reaches, although the merged nodes' own boxes are never tested."""
import ctypes as C

import numpy as np
import pytest

from nori_ray_tracer_b200 import gpu, host_scene

EMPTY = 0x80000000


def _records(sc):
    lib = gpu.load_library()
    nodes = np.ascontiguousarray(sc.nodes, np.uint32)
    cap = len(nodes) // 2 + 1
    out = np.zeros((cap, 32), np.uint32); n = C.c_uint32()
    assert lib.nori_gpu_wide_layout(nodes.ctypes.data, len(nodes), int(sc.indices.size), out.ctypes.data, cap, C.byref(n)) == 0
    return out[:n.value]


def _scene(name, golden_scene):
    if name == "heightfield":
        return host_scene.heightfield_scene(n=65, width=64, height=36)
    return golden_scene(name)


def _box_hits(mn, mx, o, rcp):
    """bbox.h:336-363 for rays without zero direction components, float32 like the kernels: (boxes, rays) -> bool"""
    with np.errstate(over="ignore", invalid="ignore"):
        t1 = (mn[:, None, :] - o[None, :, :]) * rcp[None, :, :]
        t2 = (mx[:, None, :] - o[None, :, :]) * rcp[None, :, :]
    near = np.minimum(t1, t2).max(axis=2)
    far = np.maximum(t1, t2).min(axis=2)
    return (near <= far) & (far >= np.float32(1e-4))


SCENES = ["cbox_path_mis", "table_path_mis", "sphere_mesh_normals", "veach_mis", "heightfield"]


@pytest.mark.parametrize("name", SCENES)
def test_wide_records_cover_the_tree(name, golden_scene):
    sc = _scene(name, golden_scene)
    nodes = sc.nodes; boxes = nodes.view(np.float32)
    rec = _records(sc)
    assert len(rec) > 0
    leaf = (nodes[:, 0] & 1) == 1
    size = nodes[:, 0] >> 1
    assert len(rec) <= int((~leaf).sum())                          # every record absorbs at least one inner node
    by_start = {int(nodes[i, 1]): i for i in np.nonzero(leaf & (size > 0))[0]}
    seen_leaves, seen_records = set(), set()
    stack = [(0, 1, boxes[0, 2:8])]
    max_depth = 0
    while stack:
        r, depth, parent_box = stack.pop()
        assert r not in seen_records; seen_records.add(r)
        max_depth = max(max_depth, depth)
        slots = rec[r].reshape(4, 8)
        used = [k for k in range(4) if slots[k, 3] != EMPTY]
        assert used == list(range(len(used))) and len(used) >= 1   # filled from slot 0
        inner_slots = 0
        for k in used:
            ref = int(slots[k, 3]); b = slots[k].view(np.float32)
            box = np.concatenate([b[0:3], b[4:7]])
            assert (box[:3] >= parent_box[:3]).all() and (box[3:] <= parent_box[3:]).all()
            if ref & EMPTY:
                n_prims, start = (ref >> 25) & 63, ref & 0x1ffffff
                i = by_start[start]
                assert n_prims == size[i] and start not in seen_leaves
                assert np.array_equal(box, boxes[i, 2:8])          # the reference's own leaf box
                seen_leaves.add(start)
            else:
                inner_slots += 1
                assert 0 < ref < len(rec)
                stack.append((ref, depth + 1, box))
        if len(used) < 4:
            assert inner_slots == 0                                # greedy: an inner slot would have been opened
    assert seen_records == set(range(len(rec)))
    assert seen_leaves == set(by_start)                            # every non-empty leaf exactly once
    # the spare word of every slot ranks the slots of its record in the reference's depth-first (= leaf) order
    first_leaf = {}

    def first(ref):                                                # first leaf position below a reference
        if ref & EMPTY:
            return ref & 0x1ffffff
        if ref not in first_leaf:
            first_leaf[ref] = min(first(int(x)) for x in rec[ref].reshape(4, 8)[:, 3] if x != EMPTY)
        return first_leaf[ref]
    import sys
    sys.setrecursionlimit(10000)
    for r in range(len(rec)):
        slots = rec[r].reshape(4, 8)
        used = [k for k in range(4) if slots[k, 3] != EMPTY]
        ranks = [int(slots[k, 7]) for k in used]
        assert sorted(ranks) == list(range(len(used)))
        by_rank = [first(int(slots[k, 3])) for k in sorted(used, key=lambda k: slots[k, 7])]
        assert by_rank == sorted(by_rank)
    assert 3 * max_depth <= 96                                     # the kernels' per-ray stack (NORI_STACK2_MAX)


@pytest.mark.parametrize("name", ["cbox_path_mis", "table_path_mis", "veach_mis", "heightfield"])
def test_wide_records_reach_the_same_leaves_as_the_reference_tree(name, golden_scene):
    sc = _scene(name, golden_scene)
    nodes = sc.nodes; boxes = nodes.view(np.float32)
    rec = _records(sc)
    rng = np.random.RandomState(5)
    n_rays = 1500
    lo, hi = boxes[0, 2:5], boxes[0, 5:8]
    o = (lo + (hi - lo) * (rng.rand(n_rays, 3) * 1.6 - 0.3)).astype(np.float32)
    d = rng.randn(n_rays, 3).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[: n_rays // 4] *= np.float32(1e-3) ** rng.rand(n_rays // 4, 3).astype(np.float32)   # some nearly axis-parallel rays
    o[n_rays // 2:: 7, 0] = lo[0]                                  # some origins exactly on a bounding plane
    assert (d != 0).all()
    rcp = (np.float32(1) / d).astype(np.float32)
    # reference tree: a node is visited when its parent was entered; entered = visited and box hit
    hit = _box_hits(boxes[:, 2:5], boxes[:, 5:8], o, rcp)
    leaf = (nodes[:, 0] & 1) == 1
    entered = np.zeros_like(hit)
    visited = np.zeros_like(hit); visited[0] = True
    for i in range(len(nodes)):                                    # parents precede their children (depth-first order)
        entered[i] = visited[i] & hit[i]
        if not leaf[i]:
            visited[i + 1] |= entered[i]; visited[int(nodes[i, 1])] |= entered[i]
    want = {int(nodes[i, 1]): entered[i] for i in np.nonzero(leaf & ((nodes[:, 0] >> 1) > 0))[0]}
    # records: the root's box is tested once (smStart), then only slot boxes
    slots = rec.reshape(-1, 4, 8)
    sf = slots.view(np.float32)
    reach = np.zeros((len(rec), n_rays), bool); reach[0] = hit[0]
    got = {}
    for r in range(len(rec)):                                      # a record precedes the records it refers to
        h = _box_hits(sf[r, :, 0:3], sf[r, :, 4:7], o, rcp)
        for k in range(4):
            ref = int(slots[r, k, 3])
            if ref == EMPTY:
                continue
            if ref & EMPTY:
                got[ref & 0x1ffffff] = reach[r] & h[k]
            else:
                assert ref > r
                reach[ref] = reach[r] & h[k]
    assert got.keys() == want.keys()
    for start in want:
        assert np.array_equal(got[start], want[start]), (name, start)


def test_wide_layout_is_not_built_for_a_single_leaf_tree():
    nodes = np.zeros((1, 8), np.uint32); nodes[0, 0] = (3 << 1) | 1
    out = np.zeros((4, 32), np.uint32); n = C.c_uint32(7)
    assert gpu.load_library().nori_gpu_wide_layout(nodes.ctypes.data, 1, 3, out.ctypes.data, 4, C.byref(n)) == 0
    assert n.value == 0
