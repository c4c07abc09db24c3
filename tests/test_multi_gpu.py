"""GPU (>= 2 devices): ray-sharded rendering over NCCL equals the single-GPU render bit for bit."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
CFG = {"use_hierarchical_sampling": True, "depth_sample_num": 32, "fine_depth_sample_num": 32, "agg_net_cfg": {"sample_num": 32},
       "fine_agg_net_cfg": {"sample_num": 32}, "render_depth": True, "dist_decoder_cfg": {"use_vis": False}, "ray_batch_num": 2048}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    from neuray_b200 import dist as nrd
    from neuray_b200 import renderer, synthetic
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    que, ref = synthetic.make_scene(64, 80, 6, seed=12, smooth=2)
    que = synthetic.slice_rays(que, 0, 5001)                      # odd count: ragged split
    net = renderer.NeuralRayRenderPath(CFG)
    net.load_state_dict(synthetic.make_weights(CFG, seed=2), strict=True)
    net.cuda()
    dq, dr = synthetic.to_device(que, f"cuda:{rank}"), synthetic.to_device(ref, f"cuda:{rank}")
    full = nrd.render_sharded(lambda q, r, t: net.render(q, r, t), dq, dr, False)
    single = net.render(dq, dr, False)
    torch.cuda.synchronize()
    ret[rank] = all(torch.equal(full[k], single[k]) for k in single) and set(full) == set(single)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_render_equals_single_gpu():
    world = 2
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}
