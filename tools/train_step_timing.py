"""Training-step timing on the cfg5 shape of SURVEY.md 8d (DTU-like: 8 reference views 300x400 padded to 304x400, 512 rays,
64+64 samples, is_train=True): forward through the CUDA kernels, native backward (nr_render_pass_bwd + nr_tape_gemms), Adam step.  usage: python tools/train_step_timing.py [rays] [steps]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neuray_b200 import renderer, synthetic  # noqa: E402

rays = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cfg = {"use_hierarchical_sampling": True, "fine_dist_decoder_cfg": {"use_vis": True}, "dist_decoder_cfg": {"use_vis": False},
       "render_depth": True, "ray_batch_num": rays}
que, ref = synthetic.make_scene(304, 400, 8, seed=5, smooth=2)
n = que["coords"].shape[1]
W = synthetic.make_weights(cfg, seed=1)
net = renderer.NeuralRayRenderPath(cfg)
net.load_state_dict(W, strict=True)
net.cuda()
opt = torch.optim.Adam(net.parameters(), lr=1e-4)
dr = synthetic.to_device(ref, "cuda")
gen = torch.Generator().manual_seed(0)


def batch():
    idx = torch.randperm(n, generator=gen)[:rays]
    q = {k: (v[:, idx] if k in ("coords",) else v) for k, v in que.items() if k != "imgs"}
    q["imgs"] = que["imgs"]
    return synthetic.to_device(q, "cuda")


def fwd_only():
    with torch.no_grad():
        return net.render(batch(), dr, True)


def train_step():
    out = net.render(batch(), dr, True)
    loss = ((out["pixel_colors_nr"] - out["pixel_colors_gt"]) ** 2).mean() + ((out["pixel_colors_nr_fine"] - out["pixel_colors_gt_fine"]) ** 2).mean()
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return loss


def timeit(fn, k):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(k):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / k * 1e3, r


if os.environ.get("NR_TRAIN_PROFILE"):
    from torch.profiler import ProfilerActivity, profile
    for _ in range(3):
        train_step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            train_step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
    sys.exit(0)
ms_f, _ = timeit(fwd_only, steps)
ms_t, loss = timeit(train_step, steps)
samples = rays * 128
print(f"cfg5 shape, {rays} rays x (64+64) samples, 8 views 304x400")
print(f"forward only (kernels, is_train=True): {ms_f:.2f} ms/step  ({samples / ms_f / 1e3:.2f} M ray-samples/s)")
print(f"training step (kernel forward + native backward + Adam): {ms_t:.2f} ms/step  ({samples / ms_t / 1e3:.3f} M ray-samples/s), loss {float(loss):.5f}")
