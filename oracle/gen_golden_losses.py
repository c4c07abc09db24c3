"""TEST INFRASTRUCTURE ONLY -- tests/golden/losses.npz: inputs, outputs and input gradients of the UNMODIFIED reference's
RenderLoss / DepthLoss / ConsistencyLoss (network/loss.py:17-132) and of the two calls predict_mean_for_depth_loss makes per
decoder (interpolate_feature_map + MixtureLogisticsDistDecoder.predict_mean, network/renderer.py:293-294), run on the CPU
in the build container through oracle/ref_import.py.

    python oracle/gen_golden_losses.py
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


def main():
    ref_import.load_reference()
    import network.loss as L
    from network.dist_decoder import MixtureLogisticsDistDecoder
    from network.render_ops import interpolate_feature_map
    rs = np.random.RandomState(21)
    t = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))
    out = {}

    # ---- losses ----
    qn, rn, dn, rfn, pn, h, w = 1, 50, 8, 3, 40, 24, 32
    pr = t(rs.uniform(0, 1, (qn, rn, 3))).requires_grad_(True)
    gt = t(rs.uniform(0, 1, (qn, rn, 3)))
    mask = torch.from_numpy(rs.uniform(0, 1, (qn, rn)) > 0.3)
    for tag, use in (("masked", True), ("plain", False)):
        loss = L.RenderLoss({"use_ray_mask": use})({"pixel_colors_gt": gt, "pixel_colors_nr": pr, "ray_mask": mask}, {}, 0)["loss_rgb_nr"]
        g, = torch.autograd.grad(loss.sum() * 1.7, pr)
        out[f"render_{tag}_loss"], out[f"render_{tag}_grad"] = loss.detach().numpy(), g.numpy()
    out.update(render_pr=pr.detach().numpy(), render_gt=gt.numpy(), render_mask=mask.numpy())

    depth_range = t([[2.0, 6.0], [1.5, 5.0], [2.5, 7.0]])
    true_depth = t(rs.uniform(1.0, 8.0, (rfn, 1, h, w)))
    aug_depth = true_depth * t(rs.uniform(0.97, 1.03, (rfn, 1, h, w)))
    coords = torch.from_numpy(np.stack([rs.randint(0, h, (rfn, pn)), rs.randint(0, w, (rfn, pn))], -1))      # (row, col), as the reference builds them
    depth_pr = t(rs.uniform(0, 1, (rfn, pn))).requires_grad_(True)
    gscale = t([1.0, 0.5, 2.0])
    for tag, cfg, scene in (("l2", {}, "dtu_train/scan1"), ("smooth", {"depth_loss_type": "smooth_l1"}, "dtu_train/scan1"), ("gso", {}, "gso/obj")):
        data_gt = {"ref_imgs_info": {"true_depth": true_depth, "depth": aug_depth, "depth_range": depth_range}, "scene_name": scene}
        loss = L.DepthLoss(cfg)({"depth_coords": coords, "depth_mean": depth_pr, "pixel_colors_nr": pr}, data_gt, 0)["loss_depth"]
        g, = torch.autograd.grad((loss * gscale).sum(), depth_pr)
        out[f"depth_{tag}_loss"], out[f"depth_{tag}_grad"] = loss.detach().numpy(), g.numpy()
    out.update(depth_pr=depth_pr.detach().numpy(), depth_coords=coords.numpy(), true_depth=true_depth.numpy(), aug_depth=aug_depth.numpy(),
               depth_range=depth_range.numpy(), depth_gscale=gscale.numpy())

    p0 = t(rs.uniform(0, 1, (qn, rn, dn)))
    p1 = t(rs.uniform(0, 1, (qn, rn, dn))).requires_grad_(True)
    loss = L.ConsistencyLoss({})({"hit_prob_nr": p0, "hit_prob_self": p1}, {}, 0)["loss_prob"]
    g, = torch.autograd.grad(loss.sum() * 0.3, p1)
    out.update(consist_p0=p0.numpy(), consist_p1=p1.detach().numpy(), consist_loss=loss.detach().numpy(), consist_grad=g.numpy())

    # ---- predict_mean (one decoder; quarter-resolution ray_feats) ----
    torch.manual_seed(3)
    dec = MixtureLogisticsDistDecoder({"use_vis": False})
    fh, fw = h // 4, w // 4
    ray_feats = t(rs.standard_normal((rfn, 32, fh, fw))).requires_grad_(True)
    ones = torch.ones(rfn, pn)
    f = interpolate_feature_map(ray_feats, coords, ones, h, w)
    mean = dec.predict_mean(f)
    gm = t(rs.standard_normal(tuple(mean.shape)))
    params = [p for n, p in dec.named_parameters() if n.startswith("mean_decoder.")]
    names = [n for n, _ in dec.named_parameters() if n.startswith("mean_decoder.")]
    grads = torch.autograd.grad((mean * gm).sum(), [ray_feats] + params)
    out.update(mean_ray_feats=ray_feats.detach().numpy(), mean_out=mean.detach().numpy(), mean_gout=gm.numpy(), mean_d_ray_feats=grads[0].numpy())
    for n, p in dec.state_dict().items():
        out["mean_w_" + n] = p.numpy()
    for n, g in zip(names, grads[1:]):
        out["mean_g_" + n] = g.numpy()
    path = os.path.join(ROOT, "tests", "golden", "losses.npz")
    np.savez_compressed(path, **out)
    print(path, "%.1f KB" % (os.path.getsize(path) / 1024), {k: float(np.asarray(v).mean()) for k, v in out.items() if k.endswith("_loss")})


if __name__ == "__main__":
    main()
