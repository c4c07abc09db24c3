"""TEST INFRASTRUCTURE ONLY -- tests/golden/mvsnet.npz: inputs and outputs of the UNMODIFIED reference's
construct_cost_volume_with_src (network/init_net.py:113-160) over its MVSNet (network/mvsnet/mvsnet.py), run on the CPU in the
build container through oracle/ref_import.py with seeded parameters (neuray_oracle.mvs_test_weights; the golden stores the
state-dict shapes, the inputs and the two outputs).  Two cases: 64x80 images in training mode, and a small scene run through the
reference's own code path with the evaluation resize disabled (sizes below 800).

    python oracle/gen_golden_mvsnet.py
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


def scene(rfn=2, sn=3, h=64, w=96, seed=23):
    _, ref = synthetic.make_scene(h, w, rfn, seed=seed, smooth=2, pad=32, depth_range=(2.0, 6.0), arc_deg=30.0)
    _, src = synthetic.make_scene(h, w, sn, seed=seed + 1, smooth=2, pad=32, depth_range=(2.0, 6.0), arc_deg=36.0)
    rs = np.random.RandomState(seed)
    r = {k: ref[k] for k in ("imgs", "poses", "Ks", "depth_range")}
    s = {k: src[k] for k in ("imgs", "poses", "Ks")}
    r["nn_ids"] = torch.from_numpy(np.stack([rs.permutation(sn)[:2] for _ in range(rfn)]).astype(np.int64))
    return r, s


def main():
    ref_import.load_reference()
    import network.init_net as ini
    from inplace_abn import ABN
    from network.mvsnet.mvsnet import MVSNet
    torch.manual_seed(0)
    net = MVSNet(ABN).eval()
    shapes = {k: list(v.shape) for k, v in net.state_dict().items()}
    net.load_state_dict(orc.mvs_test_weights(shapes, 31), strict=True)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    ref, src = scene()
    dn = 16
    out = {"shapes": json.dumps(shapes), "dn": dn}
    for tag, is_train in (("train", True), ("eval", False)):
        with torch.no_grad():
            cost, depth = ini.construct_cost_volume_with_src({k: v.clone() for k, v in ref.items()}, {k: v.clone() for k, v in src.items()}, net, dn, mean, std, is_train)
        out[f"{tag}_cost"], out[f"{tag}_depth"] = cost.numpy(), depth.numpy()
        print(tag, tuple(cost.shape), tuple(depth.shape), "depth mean %.3f" % float(depth.mean()), "max prob %.3f" % float(cost.max()))
    out.update({"ref_" + k: v.numpy() for k, v in ref.items()})
    out.update({"src_" + k: v.numpy() for k, v in src.items()})
    path = os.path.join(ROOT, "tests", "golden", "mvsnet.npz")
    np.savez_compressed(path, **out)
    print(path, "%.1f KB" % (os.path.getsize(path) / 1024), len(shapes), "tensors")


if __name__ == "__main__":
    main()
