import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
from __graft_entry__ import import_package
import_package()
from nori_ray_tracer_b200 import abi, nscene
from nori_ray_tracer_b200.gpu import NoriGpu
sc = nscene.load_scene('tests/golden/cbox_tmp.nscene')
g = NoriGpu(0); g.upload_scene(sc)
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
for pool in (1 << 18, 1 << 19, 1 << 20, 1 << 21):
    for poll in (8, 32):
        g.set_option('pool', pool); g.set_option('poll', poll); g.clear_film(); g.reset_stats()
        g.render(0, 8, seed=1); g.reset_stats()
        t = time.time(); g.render(0, spp, seed=1); dt = time.time() - t
        s = g.stats()
        print('pool', pool, 'poll', poll, 'spp', spp, 'Msamples/s', round(s.samples / dt / 1e6, 1), 'Mrays/s', round(s.rays / dt / 1e6, 1), 'ms', round(s.render_ms, 1), 'iters', s.iterations, flush=True)
