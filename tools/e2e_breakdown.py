"""Where the end-to-end step of bench.py spends its time beyond the device-resident step (one B200)."""
import os
import sys
import time

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
hq = {k: v.pin_memory() for k, v in que.items()}
hr = {k: v.pin_memory() for k, v in ref.items()}
print("pinned:", all(v.is_pinned() for v in list(hq.values()) + list(hr.values())))


def t(fn, n=3):
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3, r


ms, (dq, dr) = t(lambda: ({k: v.to("cuda", non_blocking=True) for k, v in hq.items()}, {k: v.to("cuda", non_blocking=True) for k, v in hr.items()}))
nbytes = sum(v.numel() * v.element_size() for v in list(hq.values()) + list(hr.values()))
print(f"H2D {nbytes / 1e6:.1f} MB: {ms:.2f} ms ({nbytes / ms / 1e6:.1f} GB/s)")
ms, out = t(lambda: net.render(dict(dq), dict(dr), False))
print(f"render (fresh dicts, frame re-packed): {ms:.2f} ms")
dr2 = dict(dr)
net.render(dict(dq), dr2, False)
ms, out = t(lambda: net.render(dict(dq), dr2, False))
print(f"render (frame pack cached): {ms:.2f} ms")
keys = ("pixel_colors_nr", "pixel_colors_nr_fine", "render_depth_fine", "ray_mask_fine")
ho = {k: torch.empty(out[k].shape, dtype=out[k].dtype).pin_memory() for k in keys}
ms, _ = t(lambda: [ho[k].copy_(out[k], non_blocking=True) for k in keys])
print(f"D2H {sum(v.numel() * v.element_size() for v in ho.values()) / 1e6:.1f} MB: {ms:.2f} ms")
t0 = time.perf_counter()
pack = renderer.frame_pack(dict(dr))
torch.cuda.synchronize()
print(f"frame_pack alone: {(time.perf_counter() - t0) * 1e3:.2f} ms")


def loop(fn, steps=4, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def up():
    return {k: v.to("cuda", non_blocking=True) for k, v in hq.items()}, {k: v.to("cuda", non_blocking=True) for k, v in hr.items()}


def a_resident():
    dr.pop(renderer.PACK_KEY, None)
    return net.render(dict(dq), dr, False)


def b_upload_render():
    q, r = up()
    return net.render(q, r, False)


def c_upload_render_d2h():
    o = b_upload_render()
    for k in keys:
        ho[k].copy_(o[k], non_blocking=True)
    return o


def stats():
    m = torch.cuda.memory_stats()
    return m["num_device_alloc"], m["num_device_free"], m["reserved_bytes.all.current"] >> 20


for name, fn in (("resident", a_resident), ("upload+render", b_upload_render), ("upload+render+d2h", c_upload_render_d2h), ("resident again", a_resident)):
    s0 = stats()
    ms = loop(fn)
    s1 = stats()
    print(f"{name}: {ms:.2f} ms/step   cudaMalloc +{s1[0] - s0[0]} cudaFree +{s1[1] - s0[1]} reserved {s0[2]} -> {s1[2]} MiB")


def d_resident_d2h():
    o = a_resident()
    for k in keys:
        ho[k].copy_(o[k], non_blocking=True)
    return o


side = torch.cuda.Stream()


def e_side_stream_d2h():
    o = b_upload_render()
    ev = torch.cuda.Event()
    ev.record()
    with torch.cuda.stream(side):
        side.wait_event(ev)
        for k in keys:
            ho[k].copy_(o[k], non_blocking=True)
            o[k].record_stream(side)
    return o


def f_sync_each():
    o = c_upload_render_d2h()
    torch.cuda.synchronize()
    return o


def nalloc():
    return torch.cuda.memory_stats()["num_device_alloc"]


for name, fn in (("resident+d2h", d_resident_d2h), ("upload+render+d2h(side stream)", e_side_stream_d2h), ("upload+render+d2h+sync", f_sync_each),
                 ("upload+render+d2h", c_upload_render_d2h)):
    n0 = nalloc()
    ms = loop(fn)
    print(f"{name}: {ms:.2f} ms/step   cudaMalloc calls during loop: {nalloc() - n0}")
print({k: (tuple(out[k].shape), out[k].dtype, out[k].is_contiguous()) for k in keys})
