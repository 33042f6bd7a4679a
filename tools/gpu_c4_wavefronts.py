"""10 M-triangle scene (BASELINE config 4) at 16 spp per pool size and number of concurrent wavefronts."""
import sys
sys.path.insert(0, '.')
from __graft_entry__ import import_package
import_package()
from nori_ray_tracer_b200 import host_scene
from nori_ray_tracer_b200.gpu import NoriGpu
sc = host_scene.heightfield_scene(n=2237)
g = NoriGpu(0); g.upload_scene(sc)
g.render(0, 2, seed=1)
for pool in (1 << 22, 1 << 23, 3 << 22):
    for wf in (1, 2, 3, 4):
        g.set_option('pool', pool); g.set_option('wavefronts', wf); g.clear_film(); g.render(0, 4, seed=1)
        ms = []
        for _ in range(2):
            g.reset_stats(); g.render(0, 16, seed=1); s = g.stats(); ms.append(s.render_ms)
        print('pool', pool, 'wavefronts', wf, 'ms', ' '.join('%.1f' % m for m in ms), 'Msamples/s %.1f' % (s.samples / min(ms) / 1e3), flush=True)
