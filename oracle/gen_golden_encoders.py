"""TEST INFRASTRUCTURE ONLY -- tests/golden/encoders.npz: inputs and outputs of the UNMODIFIED reference's
image_encoder = ResUNetLight(3, [1,2,6,4], 32, inplanes=16) (network/ops.py:150-230, renderer.py:59) and
vis_encoder = DefaultVisEncoder (network/vis_encoder.py:6-21), run on the CPU in the build container through
oracle/ref_import.py with the seeded parameters of neuray_oracle.encoder_test_weights (the golden stores the state-dict
shapes, the inputs and the outputs; the 2 M parameters are regenerated from the seed).

    python oracle/gen_golden_encoders.py
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


def main():
    ref_import.load_reference()
    from network.ops import ResUNetLight
    from network.vis_encoder import DefaultVisEncoder
    torch.manual_seed(0)
    img_net = ResUNetLight(3, [1, 2, 6, 4], 32, inplanes=16).eval()
    vis_net = DefaultVisEncoder({}).eval()
    img_shapes = {k: list(v.shape) for k, v in img_net.state_dict().items()}
    vis_shapes = {k: list(v.shape) for k, v in vis_net.state_dict().items()}
    img_net.load_state_dict(orc.encoder_test_weights(img_shapes, 11), strict=True)
    vis_net.load_state_dict(orc.encoder_test_weights(vis_shapes, 12), strict=True)
    rs = np.random.RandomState(5)
    out = {"image_shapes": json.dumps(img_shapes), "vis_shapes": json.dumps(vis_shapes)}
    # two image sizes: multiples of 16 (the datasets' padded sizes) and a size whose skip connections need padding
    for tag, (n, h, w) in {"a": (2, 64, 80), "b": (1, 40, 52)}.items():
        imgs = torch.from_numpy(rs.uniform(0, 1, (n, 3, h, w)).astype(np.float32))
        with torch.no_grad():
            img_feats = img_net(imgs)
        ray_in = torch.from_numpy(rs.standard_normal(tuple(img_feats.shape)).astype(np.float32))
        with torch.no_grad():
            ray_feats = vis_net(ray_in, img_feats)
        out.update({f"{tag}_imgs": imgs.numpy(), f"{tag}_img_feats": img_feats.numpy(), f"{tag}_ray_in": ray_in.numpy(),
                    f"{tag}_ray_feats": ray_feats.numpy()})
        print(tag, tuple(imgs.shape), "->", tuple(img_feats.shape), "img_feats std %.3f" % float(img_feats.std()), "ray_feats std %.3f" % float(ray_feats.std()))
    path = os.path.join(ROOT, "tests", "golden", "encoders.npz")
    np.savez_compressed(path, **out)
    print(path, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
