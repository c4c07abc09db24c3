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


def _train_worker(rank, world, port, ret):
    """Data-parallel step: every rank renders its own ray batch in training mode, backward through the
    native path, one flat-bucket all-reduce; the result must equal the mean of the per-batch gradients."""
    from neuray_b200 import dist as nrd
    from neuray_b200 import renderer, synthetic
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    que, ref = synthetic.make_scene(64, 80, 6, seed=12, smooth=2)
    dr = synthetic.to_device(ref, f"cuda:{rank}")
    W = synthetic.make_weights(CFG, seed=2)
    torch.manual_seed(0)                      # same fine-sampling quantiles on every rank: batches differ only in the rays

    def grads_of(batch_rank):
        net = renderer.NeuralRayRenderPath(CFG)
        net.load_state_dict(W, strict=True)
        net.cuda()
        q = synthetic.to_device(synthetic.slice_rays(que, 300 * batch_rank + 100, 300 * batch_rank + 164), f"cuda:{rank}")
        torch.manual_seed(7)
        out = net.render(q, dr, True)
        loss = (out["pixel_colors_nr"] ** 2).mean() + (out["pixel_colors_nr_fine"] ** 2).mean() + out["hit_prob_nr"].pow(2).sum() * 0.01
        loss.backward()
        return net

    net = grads_of(rank)
    nrd.allreduce_gradients(net.parameters())
    torch.cuda.synchronize()
    ok, bad = True, []
    if rank == 0:
        singles = [grads_of(r) for r in range(world)]
        for (k, p), *others in zip(net.named_parameters(), *[s.parameters() for s in singles]):
            if p.grad is None:
                continue
            if all(o.grad is None for o in others):          # e.g. a head no pass evaluates: the collective filled in zeros
                if float(p.grad.abs().max()) != 0.0:
                    bad.append(f"{k}: expected a zero gradient")
                continue
            mean = sum((o.grad if o.grad is not None else torch.zeros_like(o)) for o in others) / world
            err, scale = float((p.grad - mean).abs().max()), float(mean.abs().max())
            # fp32 atomics: the summation order differs from run to run; rgb_fc.4.bias has a zero true gradient (softmax
            # is shift invariant), so it only ever holds rounding noise -> absolute floor
            if err > 1e-4 * scale + 1e-7:
                bad.append(f"{k}: err {err:.3e} scale {scale:.3e}")
        ok = True if not bad else "; ".join(bad)
    ret[rank] = ok
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_data_parallel_gradients_equal_the_mean_of_single_rank_gradients():
    world = 2
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_train_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}
