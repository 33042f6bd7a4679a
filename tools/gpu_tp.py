import sys, time, numpy as np
sys.path.insert(0, '.')
from __graft_entry__ import import_package
import_package()
from nori_ray_tracer_b200 import abi, nscene
from nori_ray_tracer_b200.gpu import NoriGpu
sc = nscene.load_scene('tests/golden/cbox_path_mis.nscene'); sc.set_resolution(800, 600)
g = NoriGpu(0); g.upload_scene(sc)
import os
if 'NORI_DRAIN' in os.environ: g.set_option('drain', int(os.environ['NORI_DRAIN']))
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
pools = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [1 << 19, 1 << 20, 1 << 21]
g.render(0, spp, seed=1)
for pool in pools:
    g.set_option('pool', pool); g.clear_film(); g.render(0, 8, seed=1); g.reset_stats()
    g.set_option('kernel_timing', 1)
    g.render(0, spp, seed=1)
    s = g.stats(); ks = g.kernel_stats()
    g.set_option('kernel_timing', 0); g.reset_stats(); g.render(0, spp, seed=1); s2 = g.stats()
    print('pool', pool, 'ms', round(s2.render_ms, 1), 'Msamples/s', round(s2.samples / s2.render_ms / 1e3, 1), 'iters', s2.iterations, {k: round(v['ms'], 1) for k, v in ks.items() if v['ms']}, flush=True)
