#!/usr/bin/env python
"""Ray batches for the special cases of the reference's slab test (bbox.h:343-357), answered by the reference itself.

    python tests/golden/make_special_rays.py          (build container only: needs /root/reference and oracle/_ref)

Runs oracle/_ref/nori_export --rays N --special 2 (every second ray gets a zero (+0 / -0) or subnormal direction
component, half of those an origin exactly on a bounding plane of the scene or of a BVH node) on a few scenes and
keeps, per scene, the rays, the reference's answers (hit record + node-visit / primitive-test counters) and the tree
of THAT export (the reference's parallel build is not deterministic on the larger meshes) as
tests/golden/special_rays_<scene>.npz.  tests/test_oracle_golden.py replays them through the oracle.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE); sys.path.insert(0, ROOT)
import make_fixtures as mf                                         # noqa: E402
from __graft_entry__ import import_package                         # noqa: E402
import_package()
from nori_ray_tracer_b200 import nscene                            # noqa: E402

SCENES = ["cbox_path_mis", "sphere_mesh_normals", "veach_mis", "odyssey_mis"]
N_RAYS = 6000


def main():
    fixtures = {fx["name"]: fx for fx in mf.FIXTURES}
    with tempfile.TemporaryDirectory() as tmp:
        for name in SCENES:
            fx = fixtures[name]
            work = mf.mirror_scene_dir(os.path.dirname(fx["src"]), os.path.join(tmp, name))
            xml = mf.rewrite(open(os.path.join(work, os.path.basename(fx["src"]))).read(), fx, 4)
            xml_path = os.path.join(work, f"{name}.xml")
            open(xml_path, "w").write(xml)
            out = os.path.join(tmp, f"{name}.nscene")
            mf.run([mf.EXPORT, xml_path, out, "--rays", str(N_RAYS), "--special", "2", "--seed", "11"], cwd=work)
            e = nscene.read_container(out)
            ref = nscene.read_container(os.path.join(HERE, f"{name}.nscene"))
            assert np.array_equal(e["bvh.shape_offset"], ref["bvh.shape_offset"])          # same geometry as the committed fixture
            d = e["rays"][:, 4:7]                                   # nori_gpu_ray: o[3], mint, d[3], maxt
            n_special = int(((d == 0) | (np.abs(d) < 1e-38)).any(axis=1).sum())
            np.savez_compressed(os.path.join(HERE, f"special_rays_{name}.npz"), nodes=e["bvh.nodes"], indices=e["bvh.indices"],
                                rays=e["rays"], shadow=e["rays.shadow"], hits=e["rays.hits"])
            print(name, "rays", len(e["rays"]), "special", n_special, "nodes", len(e["bvh.nodes"]),
                  os.path.getsize(os.path.join(HERE, f"special_rays_{name}.npz")) // 1024, "KB")


if __name__ == "__main__":
    main()
