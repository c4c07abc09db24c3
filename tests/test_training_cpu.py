"""CPU: the interim PyTorch restatement used for gradients (tests/torch_restatement.py, the A/B reference of the native backward) equals the oracle, in value and
in gradient, so that the GPU training tests only have to check the plumbing."""
import torch

import neuray_oracle as orc
from gen_golden import flat_cfg
import torch_restatement as autograd_path
from neuray_b200 import renderer, synthetic
from neuray_b200.weights import posenc_table

CFG = {"use_hierarchical_sampling": True, "depth_sample_num": 16, "fine_depth_sample_num": 16, "agg_net_cfg": {"sample_num": 16},
       "fine_agg_net_cfg": {"sample_num": 16}, "render_depth": True}


def test_torch_pass_matches_oracle_values_and_grads():
    que, ref = synthetic.make_scene(32, 40, 4, seed=1, smooth=2)
    que = synthetic.slice_rays(que, 100, 148)
    W = synthetic.make_weights(CFG, seed=3)
    ocfg = flat_cfg({**renderer.base_cfg, **CFG})
    depth, _ = orc.sample_depth(que["depth_range"], que["coords"], 16, False)
    gw = torch.randn(1, 48, 3)

    def run(fn):
        P = {k: v.clone().requires_grad_(True) for k, v in W.items() if k.startswith(("dist_decoder.", "agg_net."))}
        r = dict(ref)
        r["ray_feats"] = ref["ray_feats"].clone().requires_grad_(True)
        r["img_feats"] = ref["img_feats"].clone().requires_grad_(True)
        pix, hit = fn(P, r)
        ((pix * gw).sum() + 0.1 * hit.pow(2).sum()).backward()
        return pix, hit, P, r

    def oracle(P, r):
        o = orc.render_by_depth({**W, **P}, ocfg, depth, que, r, True, False)
        return o["pixel_colors_nr"], o["hit_prob_nr"]

    def product(P, r):
        pix, hit, _ = autograd_path.render_pass_torch(P, "dist_decoder", "agg_net", {"use_vis_prob": True, "var_bias": 0.05}, depth,
                                                      que["coords"], que["poses"], que["Ks"], que["depth_range"], r, posenc_table(16))
        return pix, hit
    po, ho, Po, ro = run(oracle)
    pp, hp, Pp, rp = run(product)
    assert torch.allclose(pp, po, atol=1e-5) and torch.allclose(hp, ho, atol=1e-5)
    for k in Po:
        assert torch.allclose(Pp[k].grad, Po[k].grad, atol=2e-5, rtol=1e-3), k
    for k in ("ray_feats", "img_feats"):
        assert torch.allclose(rp[k].grad, ro[k].grad, atol=2e-5, rtol=1e-3), k


def test_self_hit_prob_matches_oracle():
    que, ref = synthetic.make_scene(32, 40, 3, seed=2, smooth=2)
    que = synthetic.slice_rays(que, 50, 90)
    que["ray_feats"] = torch.randn(1, 32, 8, 10)
    W = synthetic.make_weights(CFG, seed=4)
    ocfg = flat_cfg({**renderer.base_cfg, **CFG})
    depth, _ = orc.sample_depth(que["depth_range"], que["coords"], 16, False)
    dists = orc.depth2inv_dists(depth, que["depth_range"])
    gold = orc.predict_self_hit_prob(W, ocfg, que, depth, dists, False)
    P = {k: v for k, v in W.items() if k.startswith("dist_decoder.")}
    h, w = que["imgs"].shape[-2:]
    got = autograd_path.self_hit_prob_torch(P, "dist_decoder", True, 0.05, que["ray_feats"], que["coords"], h, w, depth, que["depth_range"])
    assert torch.allclose(got, gold, atol=1e-6), (got - gold).abs().max()
