import sys, numpy as np
sys.path.insert(0, '.')
from __graft_entry__ import import_package
import_package()
from nori_ray_tracer_b200 import nscene
from nori_ray_tracer_b200.gpu import NoriGpu
g = NoriGpu(0)
for arg in sys.argv[1:]:
    name, w, h, spp = arg.split(':'); w, h, spp = int(w), int(h), int(spp)
    sc = nscene.load_scene(f'tests/golden/{name}.nscene'); sc.set_resolution(w, h)
    g.upload_scene(sc); g.set_option('pool', 1 << 22); import os; g.set_option('order', int(os.environ.get('NORI_ORDER', '2')))
    if 'NORI_SHADOW_PASS' in os.environ: g.set_option('shadow_pass', int(os.environ['NORI_SHADOW_PASS']))
    if 'NORI_TRAVERSAL' in os.environ: g.set_option('traversal', int(os.environ['NORI_TRAVERSAL']))
    if 'NORI_DRAIN' in os.environ: g.set_option('drain', int(os.environ['NORI_DRAIN']))
    g.render(0, 4, seed=1)
    g.set_option('stats', 1); g.reset_stats(); g.clear_film(); g.render(0, 2, seed=1); s0 = g.stats(); g.set_option('stats', 0)
    g.set_option('kernel_timing', 1); g.reset_stats(); g.clear_film(); g.render(0, spp, seed=1); s = g.stats(); ks = g.kernel_stats(); g.set_option('kernel_timing', 0)
    print(name, f'{w}x{h}@{spp}', 'ms', round(s.render_ms, 1), 'Msamples/s', round(s.samples / s.render_ms / 1e3, 1), 'Mrays/s', round(s.rays / s.render_ms / 1e3, 1),
          'rays/sample', round(s.rays / s.samples, 2), 'nodes/ray', round(s0.nodes_visited / s0.rays, 1), 'prims/ray', round(s0.prims_tested / s0.rays, 1),
          'iters', s.iterations, {k: round(v['ms'], 1) for k, v in ks.items() if v['ms']}, flush=True)
