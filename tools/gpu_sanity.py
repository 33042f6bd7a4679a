"""Tiny renders of several fixtures covering every kernel family -- meant to be run under compute-sanitizer."""
import sys, numpy as np
sys.path.insert(0, '.')
from __graft_entry__ import import_package
import_package()
from nori_ray_tracer_b200 import nscene, host_scene
from nori_ray_tracer_b200.gpu import NoriGpu
g = NoriGpu(0)
for name, opts in (('cbox_path_mis', {}), ('cbox_path_mats', {}), ('c5_volumetric', {}), ('c3_project', {'emitter_sort': 2}),
                   ('table_path_mis', {'traversal': 2, 'shadow_pass': 1, 'order': 1}), ('table_textured', {'order': 0, 'traversal': 2}),
                   ('cbox_advcam', {}), ('cbox_perlin', {}), ('sphere_mesh_normals', {}), ('veach_mis', {'megakernel': 1})):
    sc = nscene.load_scene(f'tests/golden/{name}.nscene')
    sc.set_resolution(sc.width // 4 * 2, sc.height // 4 * 2) if False else None
    g.upload_scene(sc)
    g.set_option('pool', 1 << 13)
    for k, v in opts.items(): g.set_option(k, v)
    g.set_option('variance', 1)
    g.render(0, 2, seed=1)
    img = g.resolve(); var = g.variance()
    rays = np.zeros(256, dtype=[('o', '<f4', 3), ('mint', '<f4'), ('d', '<f4', 3), ('maxt', '<f4')]); rays['d'][:, 2] = 1; rays['maxt'] = np.inf
    g.trace(rays, 0); g.trace(rays, 1)
    for k in opts: g.set_option(k, {'emitter_sort': 1, 'traversal': 0, 'shadow_pass': 0, 'order': 2, 'megakernel': 0}[k])
    print(name, 'ok', float(img.mean()), flush=True)
sc = nscene.load_scene('tests/golden/table_path_mis.nscene')
lb, ms = host_scene.rebuild_bvh(sc, 'lbvh', leaf_size=3)
g.upload_scene(lb); g.render(0, 1, seed=2); print('lbvh ok', ms, float(g.resolve().mean()))
