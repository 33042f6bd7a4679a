"""Option "wavefronts" (concurrent wavefronts inside one render call): device time of the headline job per setting."""
import sys
sys.path.insert(0, '.')
from __graft_entry__ import import_package
import_package()
from nori_ray_tracer_b200 import nscene
from nori_ray_tracer_b200.gpu import NoriGpu
sc = nscene.load_scene('tests/golden/cbox_path_mis.nscene'); sc.set_resolution(800, 600)
g = NoriGpu(0); g.upload_scene(sc)
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
g.render(0, spp, seed=1)
for pool in [int(x) for x in (sys.argv[2].split(',') if len(sys.argv) > 2 else ['4194304', '8388608', '16777216'])]:
    for wf in (1, 2, 3, 4):
        g.set_option('pool', pool); g.set_option('wavefronts', wf); g.clear_film(); g.render(0, 8, seed=1)
        ms = []
        for _ in range(3):
            g.reset_stats(); g.render(0, spp, seed=1); s = g.stats(); ms.append(s.render_ms)
        print('pool', pool, 'wavefronts', wf, 'ms', ' '.join('%.1f' % m for m in ms), 'Msamples/s %.1f' % (s.samples / min(ms) / 1e3), 'iters', s.iterations, flush=True)
