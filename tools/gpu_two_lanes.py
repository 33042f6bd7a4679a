"""Experiment: does running two independent wavefronts on ONE device (two contexts of the library, two host threads, two
streams) hide the per-iteration launch gaps and kernel tails?  Compares one context rendering 1024 spp with an 8 Mi pool
against two contexts rendering 512 spp each with 4 Mi pools, host wall clock around the synchronous render calls."""
import sys, time, threading
sys.path.insert(0, '.')
from __graft_entry__ import import_package
import_package()
from nori_ray_tracer_b200 import nscene
from nori_ray_tracer_b200.gpu import NoriGpu
sc = nscene.load_scene('tests/golden/cbox_path_mis.nscene'); sc.set_resolution(800, 600)
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
one = NoriGpu(0); one.upload_scene(sc); one.set_option('pool', 1 << 23)
two = [NoriGpu(0), NoriGpu(0)]
for g in two: g.upload_scene(sc); g.set_option('pool', 1 << 22)
def single():
    t = time.perf_counter(); one.render(0, spp, seed=1); return (time.perf_counter() - t) * 1e3
def double():
    th = [threading.Thread(target=g.render, args=(i * spp // 2, spp // 2), kwargs=dict(seed=1)) for i, g in enumerate(two)]
    t = time.perf_counter()
    for x in th: x.start()
    for x in th: x.join()
    return (time.perf_counter() - t) * 1e3
for _ in range(2): single(); double()
for _ in range(3): print('one context, 8 Mi pool: %.1f ms    two contexts, 4 Mi pools: %.1f ms' % (single(), double()), flush=True)
