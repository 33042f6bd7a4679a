"""Throughput of every BASELINE.json config on one GPU (device time of nori_gpu_render, CUDA events)."""
import sys, time, json, numpy as np
sys.path.insert(0, '.')
from __graft_entry__ import import_package
import_package()
from nori_ray_tracer_b200 import nscene, host_scene
from nori_ray_tracer_b200.gpu import NoriGpu
g = NoriGpu(0)
rows = []

def run(label, sc, spp, pool=1 << 22, warm=2):
    g.upload_scene(sc); g.set_option('pool', pool)
    g.render(0, warm, seed=1)
    g.reset_stats(); g.clear_film(); g.render(0, spp, seed=1)
    s = g.stats()
    img = g.resolve()
    r = dict(config=label, res=f'{sc.width}x{sc.height}', spp=spp, ms=round(s.render_ms, 1), msamples_s=round(s.samples / s.render_ms / 1e3, 1),
             mrays_s=round(s.rays / s.render_ms / 1e3, 1), rays_per_sample=round(s.rays / s.samples, 2), finite=bool(np.isfinite(img).all()), mean=float(img.mean()))
    rows.append(r); print(json.dumps(r), flush=True)

def golden(name, w, h):
    sc = nscene.load_scene(f'tests/golden/{name}.nscene'); sc.set_resolution(w, h); return sc

run('C1 sphere-mesh normals (5120 triangles, primary rays)', golden('sphere_mesh_normals', 768, 768), 32)
run('C2 cornell box path_mis', golden('cbox_path_mis', 800, 600), 1024)
run('C3 disney+microfacet+envmap+thinlens path_mis', golden('c3_project', 800, 600), 2048)
for spp in (4, 16, 64):
    if spp == 4: c4 = host_scene.heightfield_scene(n=2237)
    run('C4 10M-triangle height field path_mis', c4, spp)
c5 = nscene.load_scene('tests/golden/c5_volumetric.nscene'); c5.set_film(3840, 2160)
run('C5 volumetric + spot + envmap (per-GPU share of the 8-GPU job: 512 of 4096 spp)', c5, 512)
run('table (22k triangles) path_mis', golden('table_path_mis', 800, 600), 256)
