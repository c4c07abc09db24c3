"""Gradient goldens from the UNMODIFIED reference: tests/golden/grads_<case>.npz.

Test infrastructure (like gen_golden.py).  For a golden case, the reference renderer's `render_by_depth` (network/
renderer.py:168-203) is run with autograd enabled on the case's stage rays -- coarse pass on the stored coarse depths, fine
pass on the stored fine depths (so that no sampling discontinuity is involved) -- and a fixed linear loss over
pixel_colors_nr / hit_prob_nr / render_depth is back-propagated through the reference's own modules.  Saved: the loss
weights and the gradient of every dist_decoder / agg_net / fine_* parameter and of ref ray_feats / img_feats.

    python oracle/gen_golden_grads.py [case ...]        # default: train8, views10
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_import  # noqa: E402
from golden_io import GoldenCase  # noqa: E402

HOT = ("dist_decoder", "agg_net", "fine_dist_decoder", "fine_agg_net")


def loss_weights(rn, dn_c, seed):
    g = torch.Generator().manual_seed(seed)
    return {"gw_c": torch.randn(1, rn, 3, generator=g), "gh_c": torch.randn(1, rn, dn_c, generator=g) * 0.3,
            "gw_f": torch.randn(1, rn, 3, generator=g) * 0.5, "gd_f": torch.randn(1, rn, generator=g) * 0.2}


def total_loss(out_c, out_f, lw):
    return (out_c["pixel_colors_nr"] * lw["gw_c"]).sum() + (out_c["hit_prob_nr"] * lw["gh_c"]).sum() \
        + (out_f["pixel_colors_nr"] * lw["gw_f"]).sum() + (out_f["render_depth"] * lw["gd_f"]).sum()


def run(name, mod):
    g = GoldenCase(name)
    net = mod.NeuralRayBaseRenderer(g.cfg)
    missing, unexpected = net.load_state_dict(g.W, strict=False)
    assert not unexpected and not [k for k in missing if k.split(".")[0] in HOT]
    net.eval()
    sel = g.stage_sel
    q = g.stage_que()
    r = {k: v.clone() for k, v in g.ref.items()}
    r["ray_feats"].requires_grad_(True)
    r["img_feats"].requires_grad_(True)
    dc, df = g.que_depth[:, sel].contiguous(), g.que_depth_fine[:, sel].contiguous()
    lw = loss_weights(len(sel), dc.shape[-1], 77)
    out_c = net.render_by_depth(dc, q, r, True, False)
    out_f = net.render_by_depth(df, q, r, True, True)
    total_loss(out_c, out_f, lw).backward()
    blob = {"lw_" + k: v.numpy() for k, v in lw.items()}
    n = 0
    for k, p in net.named_parameters():
        if k.split(".")[0] in HOT and p.grad is not None:
            blob["grad_" + k] = p.grad.numpy()
            n += 1
    blob["grad_ref_ray_feats"] = r["ray_feats"].grad.numpy()
    blob["grad_ref_img_feats"] = r["img_feats"].grad.numpy()
    # predict_self_hit_prob (renderer.py:137-155) on the same rays with a seeded query feature map
    gen = torch.Generator().manual_seed(91)
    qh, qw = q["imgs"].shape[-2:]
    qmap = torch.randn(1, 32, (qh + 3) // 4, (qw + 3) // 4, generator=gen).requires_grad_(True)
    gs = torch.randn(1, len(sel), dc.shape[-1], generator=gen)
    net.zero_grad()
    qq = dict(q, ray_feats=qmap)
    hit_self = net.predict_self_hit_prob(qq, dc, mod.depth2inv_dists(dc, q["depth_range"]), False)
    (hit_self * gs).sum().backward()
    blob["self_map"], blob["self_gs"] = qmap.detach().numpy(), gs.numpy()
    blob["self_hit"] = hit_self.detach().numpy()
    blob["self_grad_map"] = qmap.grad.numpy()
    for k, p in net.named_parameters():
        if k.startswith("dist_decoder.") and p.grad is not None and float(p.grad.abs().max()) > 0:
            blob["selfgrad_" + k] = p.grad.numpy()
    blob["out_c_pixel_colors_nr"] = out_c["pixel_colors_nr"].detach().numpy()
    blob["out_f_pixel_colors_nr"] = out_f["pixel_colors_nr"].detach().numpy()
    path = os.path.join(ROOT, "tests", "golden", f"grads_{name}.npz")
    np.savez_compressed(path, **blob)
    print(f"{name} -> {path} {os.path.getsize(path) / 1024:.1f} KB, {n} parameter gradients")


def main():
    mod = ref_import.load_reference()
    for name in (sys.argv[1:] or ["train8", "views10"]):
        run(name, mod)


if __name__ == "__main__":
    main()
