"""Host clock around nori_gpu_render vs the device time between its two events: what the call costs beyond the kernels."""
import sys, time
sys.path.insert(0, '.')
from __graft_entry__ import import_package
import_package()
from nori_ray_tracer_b200 import nscene
from nori_ray_tracer_b200.gpu import NoriGpu
sc = nscene.load_scene('tests/golden/cbox_path_mis.nscene'); sc.set_resolution(800, 600)
g = NoriGpu(0); g.upload_scene(sc); g.set_option('pool', 1 << 23)
for spp in (1024, 128, 16):
    g.render(0, spp, seed=1)
    for _ in range(3):
        g.reset_stats(); t = time.perf_counter(); g.render(0, spp, seed=1); host = (time.perf_counter() - t) * 1e3
        print('spp', spp, 'host ms %.2f' % host, 'device ms %.2f' % g.stats().render_ms, 'difference %.2f' % (host - g.stats().render_ms), flush=True)
