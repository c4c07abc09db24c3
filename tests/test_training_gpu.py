"""GPU: training-mode plumbing.  Forward values from the CUDA kernels, gradients through the interim autograd path; both
checked against the CPU oracle's autograd on the same inputs."""
import pytest
import torch

import neuray_oracle as orc
from gen_golden import flat_cfg
import torch_restatement as tr
from neuray_b200 import render_ops, renderer, synthetic
from neuray_b200.weights import posenc_table

pytestmark = pytest.mark.gpu
CFG = {"use_hierarchical_sampling": True, "depth_sample_num": 24, "fine_depth_sample_num": 24, "agg_net_cfg": {"sample_num": 24},
       "fine_agg_net_cfg": {"sample_num": 24}, "render_depth": True, "dist_decoder_cfg": {"use_vis": False}}


def test_gradients_match_oracle():
    que, ref = synthetic.make_scene(40, 48, 5, seed=6, smooth=2)
    que = synthetic.slice_rays(que, 300, 396)
    W = synthetic.make_weights(CFG, seed=8)
    ocfg = flat_cfg({**renderer.base_cfg, **CFG})
    gw = torch.randn(1, 96, 3)
    # oracle (CPU autograd); fine pass on the oracle's own fine depths so that both sides see identical samples
    Wo = {k: v.clone().requires_grad_(True) for k, v in W.items()}
    ro = dict(ref)
    ro["ray_feats"] = ref["ray_feats"].clone().requires_grad_(True)
    ro["img_feats"] = ref["img_feats"].clone().requires_grad_(True)
    depth, _ = orc.sample_depth(que["depth_range"], que["coords"], 24, False)
    oc = orc.render_by_depth(Wo, ocfg, depth, que, ro, True, False)
    fd = torch.sort(orc.sample_fine_depth(depth, oc["hit_prob_nr"].detach(), que["depth_range"], 24, False), -1)[0]
    of = orc.render_by_depth(Wo, ocfg, fd, que, ro, True, True)
    ((oc["pixel_colors_nr"] * gw).sum() + (of["pixel_colors_nr"] * gw).sum() * 0.5 + oc["hit_prob_nr"].pow(2).sum() * 0.1).backward()

    net = renderer.NeuralRayRenderPath(CFG)
    net.load_state_dict(W, strict=True)
    net.cuda()
    dq, dr = synthetic.to_device(que, "cuda"), synthetic.to_device(ref, "cuda")
    dr["ray_feats"].requires_grad_(True)
    dr["img_feats"].requires_grad_(True)
    pc = net.render_by_depth(depth.cuda(), dq, dr, True, False)
    pf = net.render_by_depth(fd.cuda(), dq, dr, True, True)
    g = gw.cuda()
    ((pc["pixel_colors_nr"] * g).sum() + (pf["pixel_colors_nr"] * g).sum() * 0.5 + pc["hit_prob_nr"].pow(2).sum() * 0.1).backward()
    torch.cuda.synchronize()
    assert torch.allclose(pc["pixel_colors_nr"].detach().cpu(), oc["pixel_colors_nr"].detach(), atol=1e-4)
    named = dict(net.named_parameters())
    checked = 0
    for k, p in named.items():
        go = Wo[k].grad
        if go is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        assert p.grad is not None, k
        assert torch.allclose(p.grad.cpu(), go, atol=5e-5, rtol=5e-3), (k, (p.grad.cpu() - go).abs().max())
        checked += 1
    assert checked > 80
    for k in ("ray_feats", "img_feats"):
        assert torch.allclose(dr[k].grad.cpu(), ro[k].grad, atol=5e-5, rtol=5e-3), k


def test_training_step_runs_and_updates():
    """render() in training mode (random fine quantiles, hit_prob outputs kept) + backward + optimizer step + re-render:
    the packed weight buffers must follow the parameter update."""
    cfg = dict(CFG, ray_batch_num=4096)
    que, ref = synthetic.make_scene(40, 48, 8, seed=9, smooth=2)
    que = synthetic.slice_rays(que, 0, 512)
    net = renderer.NeuralRayRenderPath(cfg)
    net.load_state_dict(synthetic.make_weights(cfg, seed=1), strict=True)
    net.cuda()
    dq, dr = synthetic.to_device(que, "cuda"), synthetic.to_device(ref, "cuda")
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    out = net.render(dq, dr, True)
    assert {"hit_prob_nr", "hit_prob_nr_fine", "pixel_colors_nr_fine", "pixel_colors_gt"} <= set(out)
    loss = ((out["pixel_colors_nr"] - out["pixel_colors_gt"]) ** 2).mean() + ((out["pixel_colors_nr_fine"] - out["pixel_colors_gt_fine"]) ** 2).mean()
    loss.backward()
    before = out["pixel_colors_nr"].detach().clone()
    opt.step()
    with torch.no_grad():
        after = net.render(dq, dr, False)["pixel_colors_nr"]
    torch.cuda.synchronize()
    assert float((after - before).abs().max()) > 1e-6          # weights re-packed after the step
    assert all(torch.isfinite(p.grad).all() for p in net.parameters() if p.grad is not None)


def _torch_pass(net, dec, agg, use_vis_prob, depth, dq, dr_leaf):
    """One pass through the PyTorch restatement (tests/torch_restatement.py) on the GPU, differentiable."""
    P = {k: v for k, v in net.named_parameters() if k.startswith(dec + ".") or k.startswith(agg + ".")}
    decm = getattr(net, dec)
    cfgv = {"use_vis_prob": use_vis_prob, "var_bias": float(decm.cfg["bias_val"])}
    ref = {k: dr_leaf[k] for k in ("poses", "Ks", "depth_range", "imgs", "ray_feats", "img_feats")}
    return tr.render_pass_torch(P, dec, agg, cfgv, depth, dq["coords"], dq["poses"], dq["Ks"], dq["depth_range"], ref,
                                posenc_table(depth.shape[-1]).to(depth.device))


def test_native_backward_matches_torch_recompute():
    """nr_render_pass_bwd (hand-written kernels + GEMMs over the tapes) against autograd over the PyTorch restatement of
    the same pass, on the GPU, at a size with thousands of rows per weight gradient (8 views, 48 rays x 24 samples, both
    passes, gradients into pixel colours, hit probabilities and depth)."""
    que, ref = synthetic.make_scene(64, 80, 8, seed=3, smooth=2)
    que = synthetic.slice_rays(que, 1000, 1048)
    cfg = dict(CFG, dist_decoder_cfg={"use_vis": True})
    W = synthetic.make_weights(cfg, seed=4)
    gw, gh = torch.randn(1, 48, 3, device="cuda"), torch.randn(1, 48, 24, device="cuda") * 0.3
    results = {}
    for mode in ("native", "torch"):
        net = renderer.NeuralRayRenderPath(cfg)
        net.load_state_dict(W, strict=True)
        net.cuda()
        dq, dr = synthetic.to_device(que, "cuda"), synthetic.to_device(ref, "cuda")
        dr["ray_feats"].requires_grad_(True)
        dr["img_feats"].requires_grad_(True)
        depth = renderer.sample_depth(dq["depth_range"], dq["coords"], 24, False)[0]
        if mode == "native":
            pc = net.render_by_depth(depth, dq, dr, True, False)
            pix_c, hit_c = pc["pixel_colors_nr"], pc["hit_prob_nr"]
        else:
            pix_c, hit_c, _ = _torch_pass(net, "dist_decoder", "agg_net", True, depth, dq, dr)
        fd = torch.sort(render_ops.sample_fine_depth(depth, hit_c.detach(), dq["depth_range"], 24, False), -1)[0]
        if mode == "native":
            pf = net.render_by_depth(fd, dq, dr, True, True)
            pix_f, dep_f = pf["pixel_colors_nr"], pf["render_depth"]
        else:
            pix_f, _, dep_f = _torch_pass(net, "fine_dist_decoder", "fine_agg_net", True, fd, dq, dr)
        loss = (pix_c * gw).sum() + (pix_f * gw).sum() * 0.5 + (hit_c * gh).sum() + dep_f.sum() * 0.2
        loss.backward()
        torch.cuda.synchronize()
        results[mode] = ({k: (p.grad.clone() if p.grad is not None else None) for k, p in net.named_parameters()},
                         dr["ray_feats"].grad.clone(), dr["img_feats"].grad.clone())
    gn, gt = results["native"], results["torch"]
    checked = 0
    for k in gt[0]:
        a, b = gn[0][k], gt[0][k]
        if b is None or float(b.abs().max()) == 0.0:
            assert a is None or float(a.abs().max()) < 1e-6, k
            continue
        assert a is not None, k
        scale = float(b.abs().max())
        assert float((a - b).abs().max()) <= 1e-3 * scale + 1e-6, (k, float((a - b).abs().max()), scale)
        checked += 1
    assert checked > 120
    for a, b in ((gn[1], gt[1]), (gn[2], gt[2])):
        assert float((a - b).abs().max()) <= 1e-3 * float(b.abs().max()) + 1e-6


@pytest.mark.parametrize("use_vis", [True, False])
def test_self_hit_prob_matches_the_oracle(use_vis):
    """use_self_hit_prob (fine-tuning configs, reference renderer.py:137-155, 188-189): nr_self_hit_prob on the GPU against
    the CPU ORACLE (oracle/neuray_oracle.py predict_self_hit_prob, itself pinned to the unmodified reference's values and
    gradients in tests/test_backward_cpu.py) -- values of hit_prob_self and the gradients its autograd sends into every
    decoder parameter and into the query view's feature map."""
    cfg = dict(CFG, use_self_hit_prob=True, dist_decoder_cfg={"use_vis": use_vis})
    que, ref = synthetic.make_scene(64, 80, 4, seed=21, smooth=2)
    que = synthetic.slice_rays(que, 500, 564)
    W = synthetic.make_weights(cfg, seed=6)
    torch.manual_seed(1)
    gw = torch.randn(1, 64, 24)
    fmap = torch.randn(1, 32, 16, 20)
    # oracle, CPU autograd
    Wo = {k: v.clone().requires_grad_(True) for k, v in W.items()}
    fo = fmap.clone().requires_grad_(True)
    depth = orc.sample_depth(que["depth_range"], que["coords"], 24, False)[0]
    qo = dict(que, ray_feats=fo)
    ho = orc.predict_self_hit_prob(Wo, flat_cfg({**renderer.base_cfg, **cfg}), qo, depth, orc.depth2inv_dists(depth, que["depth_range"]), False)
    (ho * gw).sum().backward()
    # CUDA
    net = renderer.NeuralRayRenderPath(cfg)
    net.load_state_dict(W, strict=True)
    net.cuda()
    dq, dr = synthetic.to_device(que, "cuda"), synthetic.to_device(ref, "cuda")
    dq["ray_feats"] = fmap.cuda().requires_grad_(True)
    out = net.render_by_depth(depth.cuda(), dq, dr, True, False)
    (out["hit_prob_self"] * gw.cuda()).sum().backward()
    torch.cuda.synchronize()
    assert torch.allclose(out["hit_prob_self"].detach().cpu(), ho.detach(), atol=5e-6)
    gm = fo.grad
    assert float((dq["ray_feats"].grad.cpu() - gm).abs().max()) <= 1e-3 * float(gm.abs().max()) + 1e-7
    checked = 0
    for k, p in net.named_parameters():
        g = Wo[k].grad
        if g is None or float(g.abs().max()) == 0.0:
            continue
        assert p.grad is not None, k
        assert float((p.grad.cpu() - g).abs().max()) <= 1e-3 * float(g.abs().max()) + 1e-7, k
        checked += 1
    assert checked == (24 if use_vis else 18)
