#!/usr/bin/env python
"""bench.py -- ray-samples/s of the per-ray rendering hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload black_800|black_400|cfg1]

A "step" renders ONE full synthetic query image through the hot path (sample_depth -> coarse pass -> fused
resampling -> fine pass), encoder outputs given.  Workload (config.workload): the geometry BASELINE.json's metric is
quoted on -- NeRF-synthetic lego "black_800": 800x800 query, 8 reference views 800x800 (feature maps 32ch @ 200x200),
64 coarse + 64 fine samples, neuray_gen_depth cfg (coarse decoder use_vis:false), random-init weights, synthetic maps.

  value     ray-samples/s, whole job (N ranks, one image per rank = weak scaling), inputs resident in HBM, device-timed
  e2e       same metric through NeuralRayRenderPath.render() from pinned HOST buffers: every step copies all inputs
            host->device and the rendered tiles device->host inside the timed region
  roofline  the dominant kernel (point kernel): algorithmic FLOPs per launch / CUDA-event time, vs MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (port of the reference's PyTorch path) on this box's host cores, bounded ray sample

--impl reference times the reference's CPU implementation of the path (the oracle port: /root/reference does not
exist on the GPU box) on all host cores, same workload/metric, each step a bounded sample of rays.
"""
import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (ref h, ref w, rfn, dn coarse, dn fine, description)
    "black_800": (800, 800, 8, 64, 64, "nerf_synthetic/lego/black_800 geometry: 800x800 query, 8 ref views, 64+64 samples, neuray_gen_depth cfg"),
    "black_400": (400, 400, 8, 64, 64, "nerf_synthetic/lego/black_400 geometry: 400x400 query, 8 ref views, 64+64 samples, neuray_gen_depth cfg"),
    "cfg1": (64, 64, 3, 32, 32, "64x64 query, 3 ref views, 32+32 samples"),
}
METRIC = "ray-samples/sec (800x800x64 coarse+64 fine, 8 ref views)"


def model_cfg(dn_c, dn_f):
    return {"use_hierarchical_sampling": True, "dist_decoder_cfg": {"use_vis": False}, "depth_sample_num": dn_c,
            "fine_depth_sample_num": dn_f, "agg_net_cfg": {"sample_num": dn_c}, "fine_agg_net_cfg": {"sample_num": dn_f},
            "render_depth": True, "ray_batch_num": 65536}


def point_kernel_flops_per_sample(rfn, use_vis_head):
    """Algorithmic FLOPs (2*MAC of the reference's Linear layers, no hoist credit) executed by the point kernel per
    ray-sample: SURVEY.md 8a -- per (point,view): dist decoder 12,608 (+4,160 with a vis head) + agg per-view 44,320;
    per point: geometry_fc 2*(65*64+64*16)."""
    return rfn * (12608 + (4160 if use_vis_head else 0) + 44320) + 2 * (65 * 64 + 64 * 16)


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons = index, False, [], set()
        self.max_mhz = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                for bit, n in names.items():
                    if r & bit:
                        self.reasons.add(n)
                time.sleep(0.1)
        except Exception as e:  # NVML missing: report that rather than fail the bench
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def training_step(dev, rays=512, steps=8):
    """Auxiliary number (not the headline metric): one optimisation step on the DTU-train shape of SURVEY.md 8d cfg5
    (8 reference views 304x400, 512 rays, 64+64 samples): kernel forward + native backward + Adam, inputs resident."""
    from neuray_b200 import renderer, synthetic
    cfg = {"use_hierarchical_sampling": True, "fine_dist_decoder_cfg": {"use_vis": True}, "dist_decoder_cfg": {"use_vis": False},
           "render_depth": True, "ray_batch_num": rays}
    que, ref = synthetic.make_scene(304, 400, 8, seed=5, smooth=2)
    net = renderer.NeuralRayRenderPath(cfg)
    net.load_state_dict(synthetic.make_weights(cfg, seed=1), strict=True)
    net.to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)
    dr = synthetic.to_device(ref, dev)
    gen = torch.Generator().manual_seed(0)
    n = que["coords"].shape[1]
    batches = []
    for _ in range(4):
        idx = torch.randperm(n, generator=gen)[:rays]
        q = dict(que)
        q["coords"] = que["coords"][:, idx]
        batches.append(synthetic.to_device(q, dev))

    def step(i):
        out = net.render(batches[i % 4], dr, True)
        loss = ((out["pixel_colors_nr"] - out["pixel_colors_gt"]) ** 2).mean() + ((out["pixel_colors_nr_fine"] - out["pixel_colors_gt_fine"]) ** 2).mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()

    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / steps * 1e3
    return {"ms_per_step": ms, "ray_samples_per_s": rays * 128 / ms * 1e3, "rays": rays, "samples_per_ray": 128, "ref_views": 8,
            "backward": os.environ.get("NR_BACKWARD", "native"),
            "note": "wall clock incl. host work; kernel forward + nr_render_pass_bwd + nr_tape_gemms + Adam"}


def eager_gpu_baseline(net, dq, dr, dn, dev, rays=4096, reps=3):
    """Auxiliary GPU-vs-GPU number: the same pass as plain PyTorch-eager tensor ops on the same GPU
    (neuray_b200/autograd_path.render_pass_torch, the restatement of reference renderer.py:168-203 that the tests use as
    the A/B reference of the backward), coarse pass only, forward only, on a slice of the workload's rays."""
    from neuray_b200 import render_ops
    from neuray_b200.autograd_path import render_pass_torch
    from neuray_b200.weights import posenc_table
    n = dq["coords"].shape[1]
    start = max(0, n // 2 - rays // 2)                 # a stretch of rays through the middle of the image
    coords = dq["coords"][:, start:start + rays].contiguous()
    depth = render_ops.sample_depth(dq["depth_range"], coords, dn, False)[0]
    P = {f"dist_decoder.{k}": v for k, v in net.dist_decoder.named_parameters()}
    P.update({f"agg_net.{k}": v for k, v in net.agg_net.named_parameters()})
    cfgv = {"use_vis_prob": bool(net.dist_decoder.cfg["use_vis"]), "var_bias": float(net.dist_decoder.cfg["bias_val"])}
    ref = {k: dr[k] for k in ("poses", "Ks", "depth_range", "imgs", "ray_feats", "img_feats")}
    pe = posenc_table(dn).to(dev)

    def run():
        with torch.no_grad():
            return render_pass_torch(P, "dist_decoder", "agg_net", cfgv, depth, coords, dq["poses"], dq["Ks"], dq["depth_range"], ref, pe)

    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return {"value": rays * dn / ms * 1e3, "unit": "ray-samples/s", "ms_per_pass": ms, "rays": rays, "samples_per_ray": dn,
            "what": "one coarse pass as PyTorch-eager ops on the same GPU (autograd_path.render_pass_torch), forward only"}


def cpu_threads():
    """Threads for the CPU arm: the measured optimum on the GPU box's 128-core host is 16 (8: 61 k, 16: 67 k, 32: 63 k,
    64: 22 k, 128: 1.6 k ray-samples/s on the same rays, profiles/README.md): beyond that torch's intra-op pool only adds
    contention on this op mix of many small tensors."""
    return max(1, min(os.cpu_count() or 1, int(os.environ.get("NR_CPU_THREADS", "16"))))


def oracle_throughput(wl, n_rays, steps, warmup, threads):
    """ray-samples/s of the CPU oracle on `n_rays` rays of the workload (same maps, same weights)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import neuray_oracle as orc
    from gen_golden import flat_cfg
    from neuray_b200 import synthetic, renderer
    h, w, rfn, dn_c, dn_f, _ = WORKLOADS[wl]
    torch.set_num_threads(threads)
    cfg = model_cfg(dn_c, dn_f)
    que, ref = synthetic.make_scene(h, w, rfn, seed=0, smooth=2, with_que_imgs=False)
    W = synthetic.make_weights(cfg, seed=0)
    ocfg = flat_cfg({**renderer.base_cfg, **cfg})
    total = que["coords"].shape[1]
    start = (total // 2 // w) * w + w // 4         # a stretch of rays through the middle of the image
    q = synthetic.slice_rays(que, start, start + n_rays)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            orc.render(W, ocfg, q, ref, False, ray_batch_num=4096)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    return n_rays * (dn_c + dn_f) / t, t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    h, w, rfn, dn_c, dn_f, desc = WORKLOADS[args.workload]
    threads = cpu_threads()
    n_rays = args.ref_rays
    v, t = oracle_throughput(args.workload, n_rays, args.steps, max(1, min(args.warmup, 1)), threads)
    sample = f"{n_rays} rays x ({dn_c}+{dn_f}) samples of the workload per step (a contiguous stretch of image rows)"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "ray-samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "rays_per_step": n_rays, "note": "reference's PyTorch CPU path (oracle port), host cores"},
        "cpu_baseline": {"value": v, "unit": "ray-samples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "ray-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_b200(args):
    import torch.distributed as dist
    from neuray_b200 import _lib, renderer, synthetic
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    h, w, rfn, dn_c, dn_f, desc = WORKLOADS[args.workload]
    cfg = model_cfg(dn_c, dn_f)
    cfg["ray_batch_num"] = args.ray_batch
    # one image per rank: same reference views, a different query pose per rank (seeded)
    que, ref = synthetic.make_scene(h, w, rfn, seed=0, smooth=2, with_que_imgs=False)
    if rank > 0:
        q2, _ = synthetic.make_scene(h, w, rfn, seed=0, smooth=2, with_que_imgs=False, arc_deg=60.0 + 3.0 * rank)
        que["poses"] = q2["poses"]
    W = synthetic.make_weights(cfg, seed=0)
    net = renderer.NeuralRayRenderPath(cfg)
    net.load_state_dict(W, strict=True)
    net.to(dev)
    rays = que["coords"].shape[1]
    samples_per_step = rays * (dn_c + dn_f)

    host_que = {k: v.pin_memory() for k, v in que.items()}
    host_ref = {k: v.pin_memory() for k, v in ref.items()}
    h2d = sum(v.numel() * v.element_size() for v in list(host_que.values()) + list(host_ref.values()))
    out_keys = ("pixel_colors_nr", "pixel_colors_nr_fine", "render_depth_fine", "ray_mask_fine")
    host_out = {}
    gathered = [torch.empty(1, rays, 3, device=dev) for _ in range(world)] if world > 1 else None

    @torch.no_grad()                      # inference, like the reference's render.py (renderer(data) under torch.no_grad())
    def step_device(dq, dr):
        out = net.render(dq, dr, False)
        if world > 1:   # the rendered tiles of all ranks are gathered (north_star: all-gather of the final tiles)
            dist.all_gather(gathered, out["pixel_colors_nr_fine"])
        return out

    # end-to-end arm: double-buffered uploads.  The inputs of step i+1 go up on a copy stream while step i renders (what a
    # frame loop does); every step still uploads all of its inputs and reads its result back inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    pending = {}

    def upload():
        with torch.cuda.stream(copy_stream):
            dq = {k: v.to(dev, non_blocking=True) for k, v in host_que.items()}
            dr = {k: v.to(dev, non_blocking=True) for k, v in host_ref.items()}
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return dq, dr, ev

    def step_e2e():
        dq, dr, ev = pending.pop("next") if "next" in pending else upload()
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        for t in list(dq.values()) + list(dr.values()):
            t.record_stream(cur)
        pending["next"] = upload()
        out = step_device(dq, dr)
        for k in out_keys:
            if k not in host_out:
                host_out[k] = torch.empty(out[k].shape, dtype=out[k].dtype).pin_memory()
            host_out[k].copy_(out[k], non_blocking=True)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / 1e3

    # ---- device-resident arm (value) with per-kernel event timing of the dominant kernel ----
    dq = {k: v.to(dev) for k, v in que.items()}
    dr = {k: v.to(dev) for k, v in ref.items()}

    def step_resident():
        dr.pop(renderer.PACK_KEY, None)       # re-pack the maps every frame: it is part of the per-frame path
        return step_device(dict(dq), dr)

    sampler = ClockSampler(local)
    for _ in range(args.warmup):
        step_resident()
    barrier()
    _lib.PROFILE = []
    l0 = _lib.LAUNCHES
    sampler.start()
    t_res = timed(step_resident, args.steps, 0)
    sampler.stop_flag = True
    launches = _lib.LAUNCHES - l0
    prof, _lib.PROFILE = _lib.PROFILE, None
    pk_ms = [a.elapsed_time(b) for a, b, _ in prof]
    pk_samples = [n for _, _, n in prof]
    value = world * samples_per_step * args.steps / t_res

    # ---- end-to-end arm ----
    t_e2e = timed(step_e2e, args.steps, args.warmup)
    d2h = sum(v.numel() * v.element_size() for v in host_out.values())
    e2e_value = world * samples_per_step * args.steps / t_e2e

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    peaks_src = "fallback"
    pth = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pth):
        peaks = json.load(open(pth))
        peaks_src = "measured"
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)      # kernel timed inside a long step -> sustained figure
    peak_hbm = peaks.get("hbm_gbs", 6650.0)
    fl = point_kernel_flops_per_sample(rfn, False)
    pk_time = sum(pk_ms) / 1e3
    achieved_tf = sum(pk_samples) * fl / pk_time / 1e12
    gather_gbs = sum(pk_samples) * rfn * 1072 / pk_time / 1e9
    n_launch = len(pk_ms)

    cpu = None
    if not args.no_cpu_baseline and world == 1:      # rank 0 at N=1 only (torchrun pins OMP_NUM_THREADS=1)
        threads = cpu_threads()
        v, t = oracle_throughput(args.workload, args.cpu_rays, 1, 1, threads)
        cpu = {"value": v, "unit": "ray-samples/s", "cores": threads, "kind": "port",
               "sample": f"{args.cpu_rays} rays x ({dn_c}+{dn_f}) samples of the same workload, 1 warm-up + 1 timed pass, {t:.1f} s"}

    # DRAM bytes of one point-kernel launch from the ncu --set full capture of this kernel at this launch size
    # (profiles/r1_point_kernel_pm3_v1_ncu.md: dram__bytes_read 57.7 MB + dram__bytes_write 385.1 MB per 65 536-ray launch;
    # the algorithmic gather of that launch is 36 GB -- the maps stay L2 resident -- and the record written is 335 MB)
    traffic = 57697280 + 385050624 if (args.workload == "black_800" and args.ray_batch == 65536) else None
    line = {
        "metric": METRIC, "value": value, "unit": "ray-samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_res / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "rays_per_image": rays, "images_per_step_per_gpu": 1, "ray_batch_num": args.ray_batch,
                   "parallelism": f"one image per rank x{world}" + (", NCCL all-gather of rendered tiles" if world > 1 else ""),
                   "e2e_pipeline": "inputs of step i+1 are uploaded on a copy stream while step i renders", "l2": "inputs (143 MB of maps at black_800) exceed the 126 MB L2; no explicit flush"},
        "per_gpu": value / world,
        "e2e": {"value": e2e_value, "unit": "ray-samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": t_e2e / args.steps * 1e3},
        "gpu_launches": launches,
        "clocks": sampler.summary(),
        "roofline": {"bound": "tensor", "kernel": "nr::pkt::pm3::point_kernel_pm3", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved_tf / peak_tf, "traffic": traffic, "traffic_unit": "DRAM bytes per launch (ncu --set full, profiles/r1_point_kernel_pm3_v1_ncu.md)", "peak_source": f"{peaks_src} bf16_tflops_sustained",
                     "flops_per_ray_sample": fl, "launches": n_launch, "avg_launch_ms": sum(pk_ms) / max(n_launch, 1),
                     "share_of_step": pk_time / t_res,
                     "note": "dense layers on tcgen05 with 3xTF32 (3 MMAs per algorithmic one); the kernel is epilogue/latency bound, not tensor-pipe bound (profiles/README.md)"},
        "roofline_gather": {"bound": "hbm", "achieved": gather_gbs, "peak": peak_hbm, "unit": "GB/s", "frac": gather_gbs / peak_hbm,
                            "bytes_per_ray_sample": rfn * 1072,
                            "note": "gather runs inside the point kernel; algorithmic bytes / point-kernel time"},
        "cpu_baseline": cpu,
    }
    if world == 1 and not args.no_train_step:
        try:          # auxiliary numbers: never let them cost the headline line
            line["eager_gpu_baseline"] = eager_gpu_baseline(net, dq, dr, dn_c, dev)
        except Exception as e:
            line["eager_gpu_baseline"] = {"error": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()
        try:
            line["training_step"] = training_step(dev)
        except Exception as e:
            line["training_step"] = {"error": f"{type(e).__name__}: {e}"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="black_800", choices=sorted(WORKLOADS))
    ap.add_argument("--ray-batch", type=int, default=65536)
    ap.add_argument("--cpu-rays", type=int, default=512, help="rays of the workload timed on the CPU oracle (cpu_baseline)")
    ap.add_argument("--ref-rays", type=int, default=512, help="rays per step for --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train-step", action="store_true", help="skip the auxiliary training-step timing")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
