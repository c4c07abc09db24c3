"""TEST INFRASTRUCTURE ONLY -- tests/golden/diff_feats.npz: inputs and output of the UNMODIFIED reference's
network.init_net.get_diff_feats (init_net.py:29-61) on a small seeded scene (4 views 24x32, per-pixel random depths), run on
the CPU in the build container through oracle/ref_import.py.

    python oracle/gen_golden_diff_feats.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_import  # noqa: E402
from neuray_b200 import synthetic  # noqa: E402


def scene(rfn=4, h=24, w=32, seed=9):
    _, ref = synthetic.make_scene(h, w, rfn, seed=seed, smooth=2, pad=8, depth_range=(2.0, 6.0), arc_deg=40.0)
    rs = np.random.RandomState(seed)
    # smooth-ish depth in normalised inverse depth + noise, inside [0,1]
    base = torch.from_numpy(rs.uniform(0.2, 0.8, (rfn, 1, 6, 8)).astype(np.float32))
    depth_in = torch.nn.functional.interpolate(base, size=ref["imgs"].shape[-2:], mode="bilinear", align_corners=True)
    depth_in = (depth_in + torch.from_numpy(rs.uniform(-0.02, 0.02, depth_in.shape).astype(np.float32))).clamp(0, 1)
    return {k: ref[k] for k in ("imgs", "poses", "Ks", "depth_range")}, depth_in


def main():
    ref_import.load_reference()
    import network.init_net as ini
    ref, depth_in = scene()
    with torch.no_grad():
        out = ini.get_diff_feats({k: v.clone() for k, v in ref.items()}, depth_in.clone())
    path = os.path.join(ROOT, "tests", "golden", "diff_feats.npz")
    np.savez_compressed(path, depth_in=depth_in.numpy(), out=out.numpy(), **{"ref_" + k: v.numpy() for k, v in ref.items()})
    print(path, "%.1f KB" % (os.path.getsize(path) / 1024), "out", tuple(out.shape), "mean", float(out.mean()), "valid-ish", float((out[:, 6] > 0).float().mean()))


if __name__ == "__main__":
    main()
