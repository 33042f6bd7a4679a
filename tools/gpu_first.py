import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
from __graft_entry__ import import_package
import_package()
from nori_ray_tracer_b200 import abi, nscene
from nori_ray_tracer_b200.gpu import NoriGpu
from oracle_binding import Oracle
sc = nscene.load_scene('tests/golden/cbox_tmp.nscene')
g = NoriGpu(0); g.upload_scene(sc)
o = Oracle(sc, abi)
print('pcg', g.pcg32_uint(42, 54, 6), o.pcg32_uint(42, 54, 6))
rb = sc.ray_batch()
for sh in (0, 1):
    m = rb['shadow'] == sh
    hg = g.trace(rb['rays'][m], sh); ref = rb['hits'][m]
    for f in ['t','u','v','shape','prim','nodes_visited','prims_tested']:
        print('trace', sh, f, int((hg[f] != ref[f]).sum()), 'mismatch of', int(m.sum()))
# per-sample parity at reduced res
sc.set_resolution(200, 150)
g.upload_scene(sc); o2 = Oracle(sc, abi)
for integ in ['path_mis', 'path_mats', 'normals', 'direct_mis', 'direct_ems', 'direct_mats', 'direct']:
    sc.set_integrator(integ); g.upload_scene(sc); o2 = Oracle(sc, abi)
    for mega in ([0, 1] if integ.startswith('path') else [1]):
        g.set_option('megakernel', mega); g.set_option('pool', 16384)
        a = g.render_samples(0, 4, seed=7); b = o2.render_samples(0, 4, seed=7)
        d = np.abs(a - b); rel = d / (np.abs(b) + 1e-3)
        print(integ, 'mega' if mega else 'wave', 'mean', a[..., :3].mean(), b[..., :3].mean(), 'frac rel>1e-3', float((rel.max(-1) > 1e-3).mean()), 'max', d.max())
sc.set_integrator('path_mis'); g.upload_scene(sc); g.set_option('megakernel', 0); g.set_option('pool', 1 << 20)
o2 = Oracle(sc, abi)
g.clear_film(); g.render(0, 8, seed=3); fg = g.download_film(); fo = o2.render(0, 8, seed=3, mode=0)
print('film maxdiff', np.abs(fg - fo).max(), 'rel', (np.abs(fg - fo) / (np.abs(fo) + 1e-3)).max(), 'sum w', fg[..., 3].sum(), fo[..., 3].sum())
s = g.stats(); print('stats', s.samples, s.rays, s.shadow_rays, s.iterations, s.render_ms)
# throughput
sc2 = nscene.load_scene('tests/golden/cbox_tmp.nscene'); g.upload_scene(sc2)
for pool in (1 << 19, 1 << 20, 1 << 21, 1 << 22):
    g.set_option('pool', pool); g.clear_film(); g.reset_stats()
    g.render(0, 16, seed=1); g.reset_stats()
    t = time.time(); g.render(0, 64, seed=1); dt = time.time() - t
    s = g.stats()
    print('pool', pool, 'Msamples/s', s.samples / dt / 1e6, 'Mrays/s', s.rays / dt / 1e6, 'ms', s.render_ms, 'iters', s.iterations)
