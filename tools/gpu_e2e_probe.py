"""Where does host time go in one bench step?  Times every call of the loop with perf_counter."""
import sys, time, numpy as np
sys.path.insert(0, '.')
from __graft_entry__ import import_package
import_package()
from nori_ray_tracer_b200 import nscene
from nori_ray_tracer_b200.gpu import NoriGpu
sc = nscene.load_scene('tests/golden/cbox_path_mis.nscene'); sc.set_resolution(800, 600)
g = NoriGpu(0); g.upload_scene(sc); g.set_option('pool', 1 << 22)
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
host_film = np.empty(sc.film_shape, np.float32)
for it in range(4):
    t = [time.perf_counter()]
    g.upload_scene(sc); t.append(time.perf_counter())
    g.clear_film(); t.append(time.perf_counter())
    g.render(0, spp, seed=0); t.append(time.perf_counter())
    ms = g.stats().render_ms; t.append(time.perf_counter())
    g.download_film(host_film); t.append(time.perf_counter())
    g.set_option('flush_l2', 256); g.synchronize(); t.append(time.perf_counter())
    d = [round(1e3 * (b - a), 2) for a, b in zip(t, t[1:])]
    print('upload', d[0], 'clear', d[1], 'render(host)', d[2], 'render(device)', round(ms, 2), 'stats', d[3], 'download', d[4], 'flush', d[5], flush=True)
