"""BASELINE config 4 sweep: the 10 M-triangle height field, path_mis 3840x2160, sharded by sample index over N GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/gpu_c4_dist.py [spp ...]
    python tools/gpu_c4_dist.py --devices 0,1,2,3 [spp ...]        (one process: nori_gpu_init_multi, no torch)

For every sample count (default 4 16 64): the job's device time (max over ranks, render + the one 133 MB film reduce),
Msamples/s, and speed-up / efficiency against the same job on one GPU (every rank times that on its own device).
The public path: render.RenderThread.render(distributed=True) shards with render.shard_spp exactly like this."""
import os, sys, json, time, numpy as np
sys.path.insert(0, '.')
from __graft_entry__ import import_package
import_package()
from nori_ray_tracer_b200 import host_scene, render
from nori_ray_tracer_b200.gpu import NoriGpu

args = sys.argv[1:]
devices = None
if '--devices' in args:
    i = args.index('--devices'); devices = [int(x) for x in args[i + 1].split(',')]; del args[i:i + 2]
spps = [int(a) for a in args] or [4, 16, 64]
W, H = 3840, 2160
sc = host_scene.heightfield_scene(n=2237)

if devices is not None:                                     # one process, multi-GPU inside the library
    g1 = NoriGpu(devices[0]); g1.upload_scene(sc); g1.set_option('pool', 1 << 22); g1.render(0, 2, seed=1)
    gm = NoriGpu(devices=devices); gm.upload_scene(sc); gm.set_option('pool', 1 << 22); gm.render(0, 2 * len(devices), seed=1)
    for spp in spps:
        g1.clear_film(); g1.reset_stats(); g1.render(0, spp, seed=0); t1 = g1.stats().render_ms
        gm.clear_film(); gm.reset_stats(); gm.render(0, spp, seed=0); st = gm.stats()
        err = np.abs(gm.download_film() - g1.download_film()).max() / np.abs(g1.download_film()).max()
        print(json.dumps({'mode': 'nori_gpu_init_multi', 'n_gpus': len(devices), 'spp': spp, 'ms': st.render_ms, 'reduce_ms': st.reduce_ms,
                          'msamples_per_s': W * H * spp / st.render_ms / 1e3, 'single_gpu_ms': t1, 'speedup': t1 / st.render_ms,
                          'efficiency': t1 / st.render_ms / len(devices), 'film_max_rel_diff_vs_single': float(err)}), flush=True)
    sys.exit(0)

import torch, torch.distributed as dist
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
rank, world = dist.get_rank(), dist.get_world_size()
g = NoriGpu(local); g.upload_scene(sc); g.set_option('pool', 1 << 22)
film = torch.as_tensor(g.film_device_array(), device=f'cuda:{local}')
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

def mx(x):
    t = torch.tensor([x], dtype=torch.float64, device=f'cuda:{local}'); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())

def job(spp):
    b, c = render.shard_spp(spp, rank, world)
    g.clear_film(); g.render(b, c, seed=0)
    ms = g.stats().render_ms if c else 0.0
    ev0.record(); dist.reduce(film, dst=0, op=dist.ReduceOp.SUM); ev1.record(); torch.cuda.synchronize()
    return ms, ev0.elapsed_time(ev1)

g.render(0, 2, seed=1); job(world)
for spp in spps:
    dist.barrier(); torch.cuda.synchronize()
    r, red = job(spp)
    tn, tr = mx(r + red), mx(red)
    g.clear_film(); g.render(0, spp, seed=0); t1 = mx(g.stats().render_ms)
    if rank == 0:
        print(json.dumps({'mode': 'torchrun + NCCL reduce', 'n_gpus': world, 'spp': spp, 'ms': tn, 'reduce_ms': tr, 'msamples_per_s': W * H * spp / tn / 1e3,
                          'single_gpu_ms': t1, 'speedup': t1 / tn, 'efficiency': t1 / tn / world}), flush=True)
dist.barrier(); dist.destroy_process_group()
