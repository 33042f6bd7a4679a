import sys, time, numpy as np
sys.path.insert(0, '.')
from __graft_entry__ import import_package
import_package()
from nori_ray_tracer_b200 import host_scene
from nori_ray_tracer_b200.gpu import NoriGpu
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2237
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 8
integ = sys.argv[3] if len(sys.argv) > 3 else 'path_mis'
import os
builder = os.environ.get('NORI_BUILDER', 'sah'); leaf = int(os.environ.get('NORI_LEAF', '4'))
t = time.time(); sc, sb = host_scene.heightfield_scene(n=n, integrator=integ, builder=builder, leaf_size=leaf, return_builder=True); print('builder', builder, 'leaf', leaf, 'scene+build s', round(time.time() - t, 2), 'device build ms', sb.build_ms, 'prims', sc.pod.n_indices, 'nodes', sc.pod.n_nodes, flush=True)
g = NoriGpu(0); t = time.time(); g.upload_scene(sc); print('upload s', round(time.time() - t, 2))
g.set_option('pool', 1 << 22); import os; g.set_option('order', int(os.environ.get('NORI_ORDER', '2'))); g.set_option('l2_window', int(os.environ.get('NORI_L2_WINDOW', '0')))
for wide in [int(x) for x in os.environ.get('NORI_WIDE', '1').split(',')]:
    g.set_option('wide', wide); print('== wide', wide)
    g.render(0, 2, seed=1)
    g.set_option('stats', 1); g.reset_stats(); g.clear_film(); g.render(0, 2, seed=1); s = g.stats(); kc = g.kernel_stats(); g.set_option('stats', 0)
    print('guard retraces', s.guard_retraces, 'of', kc['extend']['rays'], 'closest-hit queries; deepest stack', s.max_stack_depth)
    print('rays/sample', s.rays / s.samples, 'shadow/sample', s.shadow_rays / s.samples, 'nodes/ray', s.nodes_visited / s.rays, 'prims/ray', s.prims_tested / s.rays)
    for k in ('extend', 'shadow', 'single'):
        c = kc[k]
        if c['rays']: print(k, 'nodes/ray', c['nodes'] / c['rays'], 'prims/ray', c['prims'] / c['rays'], 'B_ray', 32 * c['nodes'] / c['rays'] + 48 * c['prims'] / c['rays'] + 48)
    g.set_option('kernel_timing', 1); g.reset_stats(); g.clear_film(); g.render(0, spp, seed=1); s = g.stats(); ks = g.kernel_stats(); g.set_option('kernel_timing', 0)
    print('spp', spp, 'ms', round(s.render_ms, 1), 'Msamples/s', round(s.samples / s.render_ms / 1e3, 1), 'Mrays/s', round(s.rays / s.render_ms / 1e3, 1), 'iters', s.iterations, {k: round(v['ms'], 1) for k, v in ks.items() if v['ms']})
    for k in ('extend', 'shadow', 'single'):
        c, t_ = kc[k], ks[k]
        if c['rays'] and t_['ms']:
            b = 32 * c['nodes'] / c['rays'] + 48 * c['prims'] / c['rays'] + 48
            print(k, 'achieved GB/s', round(t_['rays'] * b / (t_['ms'] * 1e-3) / 1e9, 1), 'Grays/s', round(t_['rays'] / t_['ms'] / 1e6, 2))
    img = g.resolve(); print('image mean', img.mean(), 'finite', np.isfinite(img).all())
