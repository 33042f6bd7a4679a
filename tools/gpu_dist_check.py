"""torchrun --nproc-per-node N tools/gpu_dist_check.py : the public multi-GPU path (render.RenderThread.render(distributed=True):
sample indices sharded across ranks, ONE NCCL reduce of the films) equals a single-GPU render of the same sample indices."""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '.')
from __graft_entry__ import import_package
import_package()
from nori_ray_tracer_b200 import nscene, render
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
rank, world = dist.get_rank(), dist.get_world_size()
for name, spp in (('cbox_path_mis', 64), ('c5_volumetric', 32), ('table_path_mis', 16)):
    sc = nscene.load_scene(f'tests/golden/{name}.nscene')
    rt = render.RenderThread(device=local)
    rgb, film = rt.render(sc, spp=spp, seed=3, distributed=True)
    if rank == 0:
        rgb1, film1 = rt.render(sc, spp=spp, seed=3, distributed=False)
        err = np.abs(film - film1).max() / np.abs(film1).max()
        print(name, 'world', world, 'spp', spp, 'max |sharded - single| / max =', float(err), 'weights equal:', bool(np.allclose(film[..., 3], film1[..., 3], rtol=1e-5)), flush=True)
        assert err < 1e-5
    dist.barrier()
dist.destroy_process_group()
