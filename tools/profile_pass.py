"""Runs a few render_impl calls of one ray chunk (for ncu captures).  usage: python tools/profile_pass.py [rays] [workload]"""
import os
import sys

import torch

torch.set_grad_enabled(False)      # inference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from neuray_b200 import renderer, synthetic  # noqa: E402

rays = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
wl = sys.argv[2] if len(sys.argv) > 2 else "black_800"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
w, (dn_c, dn_f) = bench.WORKLOADS[wl]["scene"]["w"], bench.WORKLOADS[wl]["dn"]
cfg = bench.model_cfg(dn_c, dn_f)
que, ref = bench.make_workload(wl, seed=0)
n = que["coords"].shape[1]
start = (n // 2 // w) * w
que = synthetic.slice_rays(que, start, start + rays)
W = synthetic.make_weights(cfg, seed=0)
net = renderer.NeuralRayRenderPath(cfg)
net.load_state_dict(W, strict=True)
net.cuda()
dq, dr = synthetic.to_device(que, "cuda"), synthetic.to_device(ref, "cuda")
for i in range(iters):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = net.render_impl(dq, dr, False)
    e1.record()
    torch.cuda.synchronize()
    print(f"iter {i}: {e0.elapsed_time(e1):.2f} ms  -> {rays * (dn_c + dn_f) / e0.elapsed_time(e1) / 1e3:.2f} M ray-samples/s")
