"""Runs the native encoders on the reference views of a bench workload (for ncu launch lists / captures) and prints
bench.encoders_bench's JSON.  usage: python tools/encoders_profile.py [workload] [reps]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "black_800"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
_, ref = bench.make_workload(wl, seed=0)
dr = {k: v.to("cuda") for k, v in ref.items() if torch.is_tensor(v)}
print(json.dumps(bench.encoders_bench(dr, torch.device("cuda"), reps=reps)))
