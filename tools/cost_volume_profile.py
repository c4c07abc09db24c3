"""Runs the native CostVolumeInitNet on the reference views of a bench workload (for ncu launch lists) and prints
bench.cost_volume_bench's JSON.  usage: python tools/cost_volume_profile.py [workload] [reps]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "black_800"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
_, ref = bench.make_workload(wl, seed=0)
dr = {k: v.to("cuda") for k, v in ref.items() if torch.is_tensor(v)}


def timed(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print(json.dumps(bench.cost_volume_bench(dr, torch.device("cuda"), reps, timed)))
