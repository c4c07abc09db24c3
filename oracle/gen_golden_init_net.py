"""TEST INFRASTRUCTURE ONLY -- tests/golden/depth_init_net.npz: inputs and output of the UNMODIFIED reference's DepthInitNet
(network/init_net.py:76-101: extract_depth_for_init + get_diff_feats + ResEncoder + depth_skip + conv_out) on a small seeded
scene, run on the CPU in the build container through oracle/ref_import.py with the seeded parameters of
neuray_oracle.encoder_test_weights (the golden stores state-dict shapes, inputs and the output).

    python oracle/gen_golden_init_net.py
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import neuray_oracle as orc  # noqa: E402
import ref_import  # noqa: E402
from neuray_b200 import synthetic  # noqa: E402


def scene(rfn=3, h=64, w=80, seed=17):
    _, ref = synthetic.make_scene(h, w, rfn, seed=seed, smooth=2, pad=16, depth_range=(2.0, 6.0), arc_deg=40.0)
    rs = np.random.RandomState(seed)
    hh, ww = ref["imgs"].shape[-2:]
    base = torch.from_numpy(rs.uniform(2.2, 5.5, (rfn, 1, hh // 8, ww // 8)).astype(np.float32))
    depth = torch.nn.functional.interpolate(base, size=(hh, ww), mode="bilinear", align_corners=True)
    depth = depth + torch.from_numpy(rs.uniform(-0.03, 0.03, depth.shape).astype(np.float32))
    out = {k: ref[k] for k in ("imgs", "poses", "Ks", "depth_range")}
    out["depth"] = depth
    return out


def main():
    ref_import.load_reference()
    import network.init_net as ini
    torch.manual_seed(0)
    net = ini.DepthInitNet({}).eval()
    shapes = {k: list(v.shape) for k, v in net.state_dict().items()}
    net.load_state_dict(orc.encoder_test_weights(shapes, 13), strict=True)
    ref = scene()
    with torch.no_grad():
        out = net({k: v.clone() for k, v in ref.items()}, None, False)
    path = os.path.join(ROOT, "tests", "golden", "depth_init_net.npz")
    np.savez_compressed(path, shapes=json.dumps(shapes), out=out.numpy(), **{"ref_" + k: v.numpy() for k, v in ref.items()})
    print(path, "%.1f KB" % (os.path.getsize(path) / 1024), tuple(ref["imgs"].shape), "->", tuple(out.shape), "std %.3f" % float(out.std()), len(shapes), "tensors")


if __name__ == "__main__":
    main()
