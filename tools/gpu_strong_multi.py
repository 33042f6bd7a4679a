"""Strong scaling of ONE Cornell-box render (path_mis 800x600, 1024 spp) through nori_gpu_init_multi: one process, one
context over N devices, the films summed by the library's own kernel over NVLink peer memory (no torch, no NCCL).
    python tools/gpu_strong_multi.py [N ...]        (default 1 2 4 8, capped at the devices present)"""
import json, sys
sys.path.insert(0, '.')
from __graft_entry__ import import_package
import_package()
import torch
from nori_ray_tracer_b200 import nscene
from nori_ray_tracer_b200.gpu import NoriGpu
have = torch.cuda.device_count()
ns = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]
sc = nscene.load_scene('tests/golden/cbox_path_mis.nscene'); sc.set_resolution(800, 600)
t1 = None
for n in ns:
    if n > have: break
    g = NoriGpu(devices=list(range(n))); g.upload_scene(sc); g.set_option('pool', 1 << 22)
    for _ in range(2): g.clear_film(); g.render(0, 1024, seed=0)
    ms = []
    for _ in range(3):
        g.clear_film(); g.reset_stats(); g.render(0, 1024, seed=0); st = g.stats(); ms.append((st.render_ms, st.reduce_ms))
    t, r = min(ms)
    t1 = t1 or t
    print(json.dumps({'n_gpus': n, 'ms': t, 'reduce_ms': r, 'msamples_per_s': 800 * 600 * 1024 / t / 1e3, 'speedup': t1 / t, 'efficiency': t1 / t / n}), flush=True)
    g.close()
