"""Small deterministic render for profiling: Cornell box 800x600 path_mis, a few spp."""
import sys
sys.path.insert(0, '.')
from __graft_entry__ import import_package
import_package()
from nori_ray_tracer_b200 import nscene
from nori_ray_tracer_b200.gpu import NoriGpu
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 16
pool = int(sys.argv[2]) if len(sys.argv) > 2 else (1 << 21)
scene = sys.argv[3] if len(sys.argv) > 3 else 'cbox_path_mis'
sc = nscene.load_scene(f'tests/golden/{scene}.nscene')
if scene.startswith(('cbox', 'table', 'disney', 'c3', 'c5', 'volumetric')):
    sc.set_resolution(800, 600)
g = NoriGpu(0); g.upload_scene(sc); g.set_option('pool', pool)
g.render(0, spp, seed=0)
s = g.stats()
print('samples', s.samples, 'rays', s.rays, 'ms', s.render_ms, 'iters', s.iterations, 'launches', s.kernel_launches)
