"""CPU, gloo, world_size 2: the N>1 host logic (ray sharding + tile all-gather, flat gradient all-reduce)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neuray_b200 import dist as nrd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_render(que, ref, is_train):
    c = que["coords"]
    return {"pixel_colors_nr": torch.stack([c[..., 0] * 2, c[..., 1] + 1, c[..., 0] - c[..., 1]], -1),
            "ray_mask": (c[..., 0] % 2) == 0, "render_depth": c.sum(-1)}


def _worker(rank, world, port, n_rays, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    coords = torch.stack([torch.arange(n_rays, dtype=torch.float32), torch.arange(n_rays, dtype=torch.float32) * 3], -1)[None]
    que = {"coords": coords}
    full = nrd.render_sharded(_fake_render, que, {}, False)
    single = _fake_render(que, {}, False)
    ok = all(torch.equal(full[k], single[k]) for k in single)
    # gradient all-reduce: rank r holds grads filled with r+1 -> mean = (1+2)/2
    lin = torch.nn.Linear(3, 2)
    extra = torch.nn.Parameter(torch.zeros(4))          # never receives a grad on rank 1
    lin.weight.grad = torch.full_like(lin.weight, rank + 1.0)
    lin.bias.grad = torch.full_like(lin.bias, rank + 1.0)
    if rank == 0:
        extra.grad = torch.ones(4)
    nrd.allreduce_gradients(list(lin.parameters()) + [extra])
    ok = ok and torch.allclose(lin.weight.grad, torch.full_like(lin.weight, 1.5)) and torch.allclose(extra.grad, torch.full((4,), 0.5))
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_ray_range_partitions_everything():
    for n in (0, 1, 7, 640000, 762048):
        for world in (1, 2, 3, 4, 8):
            spans = [nrd.ray_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1


def test_sharded_render_and_grad_allreduce_gloo():
    world = 2
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), 37, ret), nprocs=world, join=True)   # 37 rays: ragged split
        assert dict(ret) == {0: True, 1: True}


def test_fewer_rays_than_ranks_does_not_hang():
    """One ray, two ranks: rank 1's shard is empty; every rank must still issue the collective and get the full result."""
    world = 2
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), 1, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}
