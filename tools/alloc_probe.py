"""Which device blocks pile up when the render loop reads its results back (development probe)."""
import collections
import os
import sys

import torch

torch.set_grad_enabled(False)      # inference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from neuray_b200 import renderer, synthetic  # noqa: E402

h, w, rfn, dn_c, dn_f, _ = bench.WORKLOADS["black_800"]
cfg = bench.model_cfg(dn_c, dn_f)
cfg["ray_batch_num"] = 65536
que, ref = synthetic.make_scene(h, w, rfn, seed=0, smooth=2, with_que_imgs=False)
W = synthetic.make_weights(cfg, seed=0)
net = renderer.NeuralRayRenderPath(cfg)
net.load_state_dict(W, strict=True)
net.cuda()
dq, dr = synthetic.to_device(que, "cuda"), synthetic.to_device(ref, "cuda")
keys = ("pixel_colors_nr", "pixel_colors_nr_fine", "render_depth_fine", "ray_mask_fine")
out = net.render(dict(dq), dr, False)
ho = {k: torch.empty(out[k].shape, dtype=out[k].dtype).pin_memory() for k in keys}
del out


def hist(tag):
    act, free = collections.Counter(), collections.Counter()
    for seg in torch.cuda.memory_snapshot():
        for b in seg["blocks"]:
            (act if b["state"] == "active_allocated" else free)[round(b["size"] / 2**20, 1)] += 1
    m = torch.cuda.memory_stats()
    print(f"--- {tag}: reserved {m['reserved_bytes.all.current'] >> 20} MiB, allocated {m['allocated_bytes.all.current'] >> 20} MiB, cudaMalloc {m['num_device_alloc']}")
    print("  active (MiB: count):", dict(sorted(act.items(), reverse=True)[:12]))
    print("  cached free (MiB: count):", dict(sorted(free.items(), reverse=True)[:16]))


def step(d2h):
    o = net.render(dict(dq), dr, False)
    if d2h:
        for k in keys:
            ho[k].copy_(o[k], non_blocking=True)


torch.cuda.synchronize()
hist("start")
for i in range(3):
    step(False)
torch.cuda.synchronize()
hist("after 3 steps without readback")
for i in range(3):
    step(True)
    torch.cuda.synchronize()
    hist(f"after readback step {i}")
