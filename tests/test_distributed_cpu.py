"""The N>1 path on CPU (gloo, world_size 2): sample-index sharding + the single film reduce.
The renderer behind it is stood in for by the oracle (tests may use it; the product may not)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from __graft_entry__ import import_package
    import_package()
    from nori_ray_tracer_b200 import abi, nscene, render
    from oracle_binding import Oracle
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = nscene.load_scene(os.path.join(GOLDEN, "sphere2_mats.nscene"))
    sc.set_resolution(48, 48)
    o = Oracle(sc, abi)
    begin, count = render.shard_spp(6, rank, world)
    film = torch.from_numpy(o.render(begin, count, seed=3, mode=0))
    render.reduce_film(film, dst=0)
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), film.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_render_equals_single_process(tmp_path, make_oracle):
    from nori_ray_tracer_b200 import nscene
    world, port = 2, 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    reduced = np.load(tmp_path / "reduced.npy")
    sc = nscene.load_scene(os.path.join(GOLDEN, "sphere2_mats.nscene"))
    sc.set_resolution(48, 48)
    single = make_oracle(sc).render(0, 6, seed=3, mode=0)
    # same samples, different summation order across the two partial films
    assert np.allclose(reduced, single, rtol=1e-5, atol=1e-6)
    assert reduced[..., 3].sum() > 0
