"""Pins the CPU oracle (oracle/neuray_oracle.py) against golden vectors produced by the unmodified reference
(oracle/gen_golden.py).  CPU only."""
import pytest
import torch

import neuray_oracle as orc
from golden_io import GoldenCase

CASES = ["cfg1", "train8", "views10"]
ATOL, RTOL = 1e-4, 1e-3      # BASELINE.json north_star: 1e-4 abs / 1e-3 rel fp32


def close(a, b, atol=ATOL, rtol=RTOL, what=""):
    a, b = a.float(), b.float()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = (a - b).abs()
    ok = err <= atol + rtol * b.abs()
    assert bool(ok.all()), f"{what}: max abs err {err.max().item():.3e}, {int((~ok).sum())}/{ok.numel()} outside tol"


@pytest.mark.parametrize("name", CASES)
def test_stage_intermediates_match_reference(name):
    g = GoldenCase(name)
    cfg = g.flat_cfg()
    q = g.stage_que()
    for tag, depth, is_fine in (("c", g.que_depth[:, g.stage_sel], False), ("f", g.que_depth_fine[:, g.stage_sel], True)):
        keep = {}
        orc.render_by_depth(g.W, cfg, depth, q, dict(g.ref), g.is_train, is_fine, keep=keep)
        gold = g.stage[tag]
        for k in ("que_dists", "que_pts", "que_dir", "prj_dir", "prj_pts", "prj_depth", "prj_mask", "prj_ray_feats",
                  "prj_rgb", "prj_img_feats", "prj_vis", "prj_hit_prob", "density", "colors"):
            # tighter than the product tolerance: the oracle is the same arithmetic in the same precision
            close(keep[k], gold[k], atol=2e-5, rtol=1e-4, what=f"{name}/{tag}/{k}")


@pytest.mark.parametrize("name", CASES)
def test_sampling_matches_reference(name):
    g = GoldenCase(name)
    cfg = g.flat_cfg()
    depth, _ = orc.sample_depth(g.que["depth_range"], g.que["coords"], cfg["depth_sample_num"], False)
    assert torch.equal(depth, g.que_depth)
    fd = orc.sample_fine_depth(g.que_depth, g.out["hit_prob_nr"], g.que["depth_range"], cfg["fine_depth_sample_num"],
                               g.is_train, g.fine_u)
    if cfg["fine_depth_use_all"]:
        fd = torch.cat([g.que_depth, fd], -1)
    fd = torch.sort(fd, -1)[0]
    close(fd, g.que_depth_fine, atol=1e-6, rtol=1e-6, what="fine depth (golden hit_prob injected)")


@pytest.mark.parametrize("name", CASES)
def test_render_impl_matches_reference(name):
    g = GoldenCase(name)
    cfg = g.flat_cfg()
    n = min(512, g.que["coords"].shape[1])
    q = dict(g.que)
    q["coords"] = g.que["coords"][:, :n].contiguous()
    # fine pass with the reference's own fine depths injected (searchsorted is discontinuous)
    out = orc.render_impl(g.W, cfg, q, dict(g.ref), g.is_train, fine_depth_override=g.que_depth_fine[:, :n])
    for k, v in g.out.items():
        if k not in out:
            continue
        if v.dtype == torch.bool:
            assert torch.equal(out[k], v[:, :n]), k
        else:
            close(out[k], v[:, :n], what=f"{name}/{k}")
    assert set(g.out) <= set(out), set(g.out) - set(out)


@pytest.mark.parametrize("name", ["train8", "views10"])
def test_oracle_autograd_matches_the_reference_gradients(name):
    """Gradients of a fixed linear loss over both passes, produced by the unmodified reference's autograd
    (tests/golden/grads_*.npz, oracle/gen_golden_grads.py), against autograd over the oracle restatement."""
    import os

    import numpy as np

    import neuray_oracle as orc
    from golden_io import GOLDEN_DIR, GoldenCase
    g = GoldenCase(name)
    z = np.load(os.path.join(GOLDEN_DIR, f"grads_{name}.npz"))
    lw = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("lw_")}
    gold = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad_")}
    W = {k: v.clone().requires_grad_(True) for k, v in g.W.items()}
    ref = dict(g.ref)
    ref["ray_feats"] = g.ref["ray_feats"].clone().requires_grad_(True)
    ref["img_feats"] = g.ref["img_feats"].clone().requires_grad_(True)
    q, sel, cfg = g.stage_que(), g.stage_sel, g.flat_cfg()
    oc = orc.render_by_depth(W, cfg, g.que_depth[:, sel], q, ref, True, False)
    of = orc.render_by_depth(W, cfg, g.que_depth_fine[:, sel], q, ref, True, True)
    loss = (oc["pixel_colors_nr"] * lw["gw_c"]).sum() + (oc["hit_prob_nr"] * lw["gh_c"]).sum() \
        + (of["pixel_colors_nr"] * lw["gw_f"]).sum() + (of["render_depth"] * lw["gd_f"]).sum()
    loss.backward()
    for k, gref in gold.items():
        got = ref[k[4:]].grad if k.startswith("ref_") else W[k].grad
        assert got is not None, k
        err, scale = float((got - gref).abs().max()), float(gref.abs().max())
        assert err <= 2e-4 * max(scale, 1e-3) + 2e-6, (k, err, scale)
