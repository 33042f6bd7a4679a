import sys
sys.path.insert(0, '.')
from __graft_entry__ import import_package
import_package()
from nori_ray_tracer_b200 import host_scene
from nori_ray_tracer_b200.gpu import NoriGpu
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2237
sc = host_scene.heightfield_scene(n=n)          # 3840 x 2160, pool 4 Mi: the launches of bench.py's config-4 run
g = NoriGpu(0); g.upload_scene(sc); g.set_option('pool', 1 << 22); g.set_option('wavefronts', 1)   # one wavefront, as in bench.py's kernel-timing pass
import os; g.set_option('order', int(os.environ.get('NORI_ORDER', '2')))
g.render(0, 4, seed=1)
s = g.stats(); print('ms', s.render_ms, 'rays', s.rays, 'iters', s.iterations)
