"""GPU: oracle parity AT the configurations BASELINE.json names, on ray slices the CPU oracle finishes
in seconds -- the kernel instantiations, map sizes, sample counts and weights that bench.py and the scaling runs time:

  cfg2       lego/black_400: 400x400, maps 100x100, 8 views, 64+64, neuray_gen_depth cfg (fine decoder keeps use_vis)
  black_800  the benched workload: 800x800, maps 200x200 (143 MB of maps), 8 views, 64+64, bench.model_cfg + its weights
  cfg3       black_800 with 64 coarse + fine_depth_use_all -> 128 fine samples (configs/gen/neuray_gen_cost_volume.yaml shape)
  cfg4       fern/high: query 1008x756, refs padded to 1008x768, 10 views (16 lanes per point), depth range (1.2, 12)
  cfg5       one DTU-shape training step: 8 views 300x400 padded to 304x400, depth (0.8, 4.0), 512 rays, 64+64, is_train:
             values + every parameter gradient + both map gradients against the oracle's autograd

Each slice = a stretch of rays through the image centre + a stretch along an image edge (where most views are masked).
Tolerance: BASELINE.json north_star, 1e-4 abs / 1e-3 rel; ray_mask exact.  The fine pass is compared on the oracle's own
fine depths (searchsorted is discontinuous) and, end to end, by PSNR + error quantiles.
"""
import os
import sys

import pytest
import torch

import neuray_oracle as orc
from gen_golden import flat_cfg
from neuray_b200 import renderer, synthetic

pytestmark = pytest.mark.gpu
ATOL, RTOL = 1e-4, 1e-3
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def close(a, b, what, atol=ATOL, rtol=RTOL):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = (a - b).abs()
    bad = int((err > atol + rtol * b.abs()).sum()) + int(torch.isnan(a).sum())
    print(f"{what}: max abs err {err.max().item():.3e} (ref max {b.abs().max().item():.3e}), {bad}/{err.numel()} outside tol")
    assert bad == 0, f"{what}: {bad} values outside tolerance, max err {err.max().item():.3e}"


def pick_rays(que, w, h, n=128):
    """n rays through the image centre + n rays along the top-left edge region."""
    coords = que["coords"]
    centre = (h // 2) * w + w // 2 - n // 2
    edge = 2 * w + 1
    idx = torch.cat([torch.arange(centre, centre + n), torch.arange(edge, edge + n)])
    return dict(que, coords=coords[:, idx].contiguous())


# synthetic.make_weights' default density head (gain 8, bias 0.3 -- the bench's) makes every ray terminate within its first
# few samples; a softer head spreads the termination over the ray (max hit probability ~0.2, render depths 2.0 .. 3.8), so
# that the later samples, the resampling and the fine pass carry weight in the comparison
SPREAD = dict(sigma_gain=1.0, sigma_bias=0.02)


def check(scene, cfg, seed_w, qw, qh, e2e=True, head=SPREAD):
    que, ref = synthetic.make_scene(**scene)
    que = pick_rays(que, qw, qh)
    W = synthetic.make_weights(cfg, seed=seed_w, **head)
    ocfg = flat_cfg({**renderer.base_cfg, **cfg})
    torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))
    with torch.no_grad():
        gold = orc.render_impl(W, ocfg, que, ref, False)
    net = renderer.NeuralRayRenderPath(cfg)
    net.load_state_dict(W, strict=True)
    net.cuda()
    dq, dr = synthetic.to_device(que, "cuda"), synthetic.to_device(ref, "cuda")
    with torch.no_grad():
        out = net.render_impl(dq, dr, False)
        out_f = net.render_by_depth(gold["que_depth_fine"].cuda(), dq, dr, False, True)
    torch.cuda.synchronize()
    term = float(gold["hit_prob_nr"].sum(-1).mean())
    print(f"mean ray opacity {term:.3f}, ray_mask fraction {float(gold['ray_mask'].float().mean()):.2f}")
    assert term > 0.05, "degenerate test: rays do not terminate"
    for k in ("pixel_colors_nr", "hit_prob_nr", "render_depth"):
        close(out[k], gold[k], k)
        close(out_f[k], gold[k + "_fine"], k + "_fine (oracle's fine depths)")
    assert torch.equal(out["ray_mask"].cpu(), gold["ray_mask"])
    assert torch.equal(out_f["ray_mask"].cpu(), gold["ray_mask_fine"])
    if e2e:
        fine, g = out["pixel_colors_nr_fine"].cpu(), gold["pixel_colors_nr_fine"]
        err = (fine - g).abs().flatten()
        mse = float((err ** 2).mean())
        psnr = 10 * torch.log10(torch.tensor(1.0 / max(mse, 1e-20))).item()
        bad = float((err > ATOL + RTOL * g.abs().flatten()).float().mean())
        print(f"end-to-end fine colours: PSNR {psnr:.1f} dB, {bad * 100:.2f} % outside tol, q99.9 {torch.quantile(err, 0.999).item():.2e}, max {err.max().item():.2e}")
        assert psnr > 60.0 and bad <= 0.01 and torch.quantile(err, 0.999).item() < 2e-3 and err.max().item() < 2e-2
    return out, gold


GEN_DEPTH = {"use_hierarchical_sampling": True, "dist_decoder_cfg": {"use_vis": False}, "render_depth": True}


def test_cfg2_black_400():
    check(dict(h=400, w=400, rfn=8, seed=0, smooth=2, with_que_imgs=False), dict(GEN_DEPTH), 0, 400, 400)


def test_benched_workload_black_800():
    sys.path.insert(0, ROOT)
    import bench
    wl = bench.WORKLOADS["black_800"]
    cfg = bench.model_cfg(*wl["dn"])
    scene = dict(wl["scene"], rfn=wl["rfn"], seed=0, smooth=2, with_que_imgs=False)                     # = bench.make_workload("black_800")
    check(scene, cfg, 0, 800, 800, head={})       # exactly what bench.py renders
    check(scene, cfg, 0, 800, 800)                # same maps, softer density head


def test_cfg3_black_800_fine_use_all_128():
    cfg = dict(GEN_DEPTH, fine_depth_use_all=True, fine_depth_sample_num=64, fine_agg_net_cfg={"sample_num": 128})
    out, gold = check(dict(h=800, w=800, rfn=8, seed=1, smooth=2, with_que_imgs=False), cfg, 3, 800, 800)
    assert gold["que_depth_fine"].shape[-1] == 128


def test_cfg4_fern_high_10_views():
    scene = dict(h=756, w=1008, rfn=10, seed=7, smooth=2, depth_range=(1.2, 12.0), arc_deg=100.0, focal=0.83 * 1008, with_que_imgs=False)
    out, gold = check(scene, dict(GEN_DEPTH), 5, 1008, 756)


def test_cfg5_training_step_values_and_gradients():
    cfg = dict(GEN_DEPTH, fine_dist_decoder_cfg={"use_vis": True}, ray_batch_num=512)
    que, ref = synthetic.make_scene(300, 400, 8, seed=5, smooth=2, depth_range=(0.8, 4.0), radius=2.4)
    assert ref["imgs"].shape[-2:] == (304, 400)
    gen = torch.Generator().manual_seed(0)
    idx = torch.randperm(que["coords"].shape[1], generator=gen)[:512]
    que = dict(que, coords=que["coords"][:, idx].contiguous())
    W = synthetic.make_weights(cfg, seed=1, **SPREAD)
    ocfg = flat_cfg({**renderer.base_cfg, **cfg})
    torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))
    u = torch.rand(1, 512, 64, generator=gen)
    # oracle: training mode (recorded random quantiles), the reference's render loss on both passes
    Wo = {k: v.clone().requires_grad_(True) for k, v in W.items()}
    ro = dict(ref, ray_feats=ref["ray_feats"].clone().requires_grad_(True), img_feats=ref["img_feats"].clone().requires_grad_(True))
    gold = orc.render_impl(Wo, ocfg, que, ro, True, fine_u=u)
    loss_o = ((gold["pixel_colors_nr"] - gold["pixel_colors_gt"]) ** 2).mean() + ((gold["pixel_colors_nr_fine"] - gold["pixel_colors_gt_fine"]) ** 2).mean()
    loss_o.backward()
    # CUDA: coarse pass + fine pass on the oracle's fine depths (identical samples on both sides), same loss
    net = renderer.NeuralRayRenderPath(cfg)
    net.load_state_dict(W, strict=True)
    net.cuda()
    dq, dr = synthetic.to_device(que, "cuda"), synthetic.to_device(ref, "cuda")
    dr["ray_feats"].requires_grad_(True)
    dr["img_feats"].requires_grad_(True)
    depth = renderer.sample_depth(dq["depth_range"], dq["coords"], 64, False)[0]
    pc = net.render_by_depth(depth, dq, dr, True, False)
    pf = net.render_by_depth(gold["que_depth_fine"].detach().cuda(), dq, dr, True, True)
    loss = ((pc["pixel_colors_nr"] - pc["pixel_colors_gt"]) ** 2).mean() + ((pf["pixel_colors_nr"] - pf["pixel_colors_gt"]) ** 2).mean()
    loss.backward()
    torch.cuda.synchronize()
    for k in ("pixel_colors_nr", "hit_prob_nr", "render_depth"):
        close(pc[k], gold[k], k)
        close(pf[k], gold[k + "_fine"], k + "_fine")
    assert abs(float(loss) - float(loss_o)) < 1e-5
    checked = 0
    for k, p in net.named_parameters():
        g = Wo[k].grad
        if g is None or float(g.abs().max()) == 0.0:
            assert p.grad is None or float(p.grad.abs().max()) < 1e-9, k
            continue
        assert p.grad is not None, k
        scale, err = float(g.abs().max()), float((p.grad.cpu() - g).abs().max())
        assert err <= 2e-3 * scale + 1e-8, (k, err, scale)
        checked += 1
    assert checked > 120, checked
    for k in ("ray_feats", "img_feats"):
        g = ro[k].grad
        scale, err = float(g.abs().max()), float((dr[k].grad.cpu() - g).abs().max())
        assert err <= 2e-3 * scale + 1e-9, (k, err, scale)
