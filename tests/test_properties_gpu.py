"""GPU: size-independent properties of the rendering path at BASELINE.json's full sizes (black_800 geometry, 8 views,
64+64 samples) where the CPU oracle is far too slow: chunking invariance (bit-exact), ray-order invariance (bit-exact),
probability mass and monotonicity of the resampled depths."""
import pytest
import torch

import bench
from neuray_b200 import renderer, synthetic

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    w, (dn_c, dn_f) = bench.WORKLOADS["black_800"]["scene"]["w"], bench.WORKLOADS["black_800"]["dn"]
    cfg = bench.model_cfg(dn_c, dn_f)
    que, ref = bench.make_workload("black_800", seed=0)
    n = que["coords"].shape[1]
    start = (n // 2 // w) * w + 123
    que = synthetic.slice_rays(que, start, start + 12345)          # ragged: not a multiple of the 32-point tile or the chunk
    W = synthetic.make_weights(cfg, seed=0)
    net = renderer.NeuralRayRenderPath(cfg)
    net.load_state_dict(W, strict=True)
    net.cuda()
    return net, synthetic.to_device(que, "cuda"), synthetic.to_device(ref, "cuda")


def render(net, que, ref, chunk):
    net.cfg["ray_batch_num"] = chunk
    return net.render(dict(que), ref, False)


def test_chunking_is_bit_exact(setup):
    net, que, ref = setup
    a = render(net, que, ref, 1 << 20)
    b = render(net, que, ref, 4096)
    c = render(net, que, ref, 1000)
    torch.cuda.synchronize()
    for k in a:
        assert torch.equal(a[k], b[k]) and torch.equal(a[k], c[k]), k


def test_ray_order_is_irrelevant(setup):
    net, que, ref = setup
    perm = torch.randperm(que["coords"].shape[1], device="cuda")
    q2 = dict(que)
    q2["coords"] = que["coords"][:, perm].contiguous()
    a = render(net, que, ref, 1 << 20)
    b = render(net, q2, ref, 1 << 20)
    torch.cuda.synchronize()
    for k in a:
        assert torch.equal(a[k][:, perm], b[k]), k


def test_probability_mass_and_resampled_depths(setup):
    net, que, ref = setup
    net.cfg["ray_batch_num"] = 1 << 20
    depth, _ = renderer.sample_depth(que["depth_range"], que["coords"], net.cfg["depth_sample_num"], False)
    res = renderer.run_pass(net, depth, que, ref, False, fine=renderer._fine_request(net, depth.shape[1], False, depth.device))
    torch.cuda.synchronize()
    hit, fd = res["hit_prob"], res["fine_depth"]
    assert bool((hit >= 0).all()) and float(hit.sum(-1).max()) <= 1.0 + 1e-5
    assert bool((fd[..., 1:] >= fd[..., :-1]).all())                     # sorted
    near, far = float(que["depth_range"][0, 0]), float(que["depth_range"][0, 1])
    assert float(fd.min()) >= near * (1 - 1e-5) and float(fd.max()) <= far * (1 + 1e-5)
    assert torch.isfinite(res["pixel_colors"]).all() and float(res["pixel_colors"].min()) >= -1e-5 and float(res["pixel_colors"].max()) <= 1 + 1e-4
