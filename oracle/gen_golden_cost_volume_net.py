"""TEST INFRASTRUCTURE ONLY -- tests/golden/cost_volume_init_net.npz: output of the UNMODIFIED reference's CostVolumeInitNet.forward
(network/init_net.py:205-254, cost_volume_sn = 16) on the scene of tests/golden/mvsnet.npz, run on the CPU in the build container
through oracle/ref_import.py; parameters seeded (MVSNet: neuray_oracle.mvs_test_weights(…, 31), the rest:
encoder_test_weights(…, 32)); the golden stores shapes and the output (inputs are mvsnet.npz's).

    python oracle/gen_golden_cost_volume_net.py
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
from gen_golden_mvsnet import scene  # noqa: E402


def weights(shapes):
    mv = {k[7:]: v for k, v in shapes.items() if k.startswith("mvsnet.")}
    rest = {k: v for k, v in shapes.items() if not k.startswith("mvsnet.") and not k.startswith("imagenet_")}
    W = {"mvsnet." + k: v for k, v in orc.mvs_test_weights(mv, 31).items()}
    W.update(orc.encoder_test_weights(rest, 32))
    return W


def main():
    ref_import.load_reference()
    import network.init_net as ini
    torch.manual_seed(0)
    cwd = os.getcwd()
    os.chdir(ref_import.REFERENCE_ROOT)          # the constructor opens network/mvsnet/mvsnet_pl.ckpt relative to the reference root
    try:
        net = ini.CostVolumeInitNet({"cost_volume_sn": 16}).eval()
    finally:
        os.chdir(cwd)
    shapes = {k: list(v.shape) for k, v in net.state_dict().items()}
    missing, unexpected = net.load_state_dict(weights(shapes), strict=False)
    assert not unexpected and set(missing) == {"imagenet_mean", "imagenet_std"}, (missing, unexpected)
    ref, src = scene()
    out = {}
    for tag, is_train in (("train", True), ("eval", False)):
        with torch.no_grad():
            out[tag] = net({k: v.clone() for k, v in ref.items()}, {k: v.clone() for k, v in src.items()}, is_train).numpy()
        print(tag, out[tag].shape, "std %.3f" % float(out[tag].std()))
    path = os.path.join(ROOT, "tests", "golden", "cost_volume_init_net.npz")
    np.savez_compressed(path, shapes=json.dumps(shapes), **out)
    print(path, "%.1f KB" % (os.path.getsize(path) / 1024), len(shapes), "tensors")


if __name__ == "__main__":
    main()
