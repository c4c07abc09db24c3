"""GPU parity: the CUDA path (through the C-ABI) against the golden vectors of the unmodified reference and
against the CPU oracle.  Tolerance = BASELINE.json north_star: 1e-4 abs / 1e-3 rel, fp32."""
import ctypes as C
import os

import pytest
import torch

import neuray_oracle as orc
from golden_io import GoldenCase
from neuray_b200 import _lib, renderer, render_ops, synthetic

pytestmark = pytest.mark.gpu
ATOL, RTOL = 1e-4, 1e-3
# Projected pixel coordinates are O(100) values computed through x/z in fp32 on both sides: 1e-4 abs on a coordinate of
# 50 px is 2e-6 relative, i.e. a handful of ulps; measured max difference 4e-5 px (cfg1), 3x that is the gate.
PTS_ATOL = 1.2e-4
CASES = ["cfg1", "train8", "views10"]


FAILS = []


def close(a, b, atol=ATOL, rtol=RTOL, what="", max_bad_frac=0.0, defer=False):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = (a - b).abs()
    ok = err <= atol + rtol * b.abs()
    bad = int((~ok).sum()) + int(torch.isnan(a).sum())
    msg = f"{what}: max abs err {err.max().item():.3e} (ref max {b.abs().max().item():.3e}), {bad}/{ok.numel()} outside tol"
    print(("OK   " if bad <= max_bad_frac * ok.numel() else "FAIL ") + msg)
    if bad > max_bad_frac * ok.numel():
        if defer:
            FAILS.append(msg)
        else:
            raise AssertionError(msg)


def flush_fails():
    if FAILS:
        msg = "\n".join(FAILS)
        FAILS.clear()
        raise AssertionError(msg)


def build_path(g, device="cuda"):
    net = renderer.NeuralRayRenderPath(g.cfg)
    missing, unexpected = net.load_state_dict(g.W, strict=True)
    return net.to(device)


def dev(d):
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in d.items()}


@pytest.mark.parametrize("name", CASES)
def test_point_kernel_stage_tap(name):
    """project_points_dict + predict_proj_ray_prob + get_img_feats, field by field (rows a4-a9)."""
    g = GoldenCase(name)
    net = build_path(g)
    ref = dev(g.ref)
    que = dev(g.stage_que())
    for tag, depth, is_fine in (("c", g.que_depth[:, g.stage_sel], False), ("f", g.que_depth_fine[:, g.stage_sel], True)):
        depth = depth.cuda().contiguous()
        _, rn, dn = depth.shape
        pack = renderer.frame_pack(ref)
        wp, wr, pe, wt = renderer.pass_weights(net, is_fine, dn, depth.device)
        from neuray_b200.weights import camera_blocks
        cam, _ = camera_blocks(que, None)
        rec = torch.empty(rn * dn * 20, device="cuda")
        dbg = torch.zeros(pack.rfn, rn * dn, 76, device="cuda")
        p = _lib.NrPassParams()
        coords = que["coords"][0].contiguous()
        p.coords, p.que_depth, p.que_cam, p.rn, p.dn = coords.data_ptr(), depth.data_ptr(), cam.data_ptr(), rn, dn
        p.feat, p.rgb, p.view_params = pack.feat.data_ptr(), pack.rgb.data_ptr(), pack.view_params.data_ptr()
        p.rfn, p.h, p.w, p.fh, p.fw = pack.rfn, pack.h, pack.w, pack.fh, pack.fw
        p.w_point, p.w_ray, p.pos_enc = wp.data_ptr(), wr.data_ptr(), pe.data_ptr()
        dec = net.fine_dist_decoder if is_fine else net.dist_decoder
        p.use_vis = 1 if net.dist_decoder.cfg["use_vis"] else 0
        p.var_bias = dec.cfg["bias_val"]
        p.point_rec = rec.data_ptr()
        p.w_tc = wt.data_ptr()
        _lib.check(_lib.lib().nr_point_kernel_debug(C.byref(p), dbg.data_ptr(), None), "debug")
        torch.cuda.synchronize()
        d = dbg.reshape(pack.rfn, 1, rn, dn, 76)
        gold = g.stage[tag]
        close(d[..., 0:1], gold["prj_mask"], what=f"{name}/{tag}/mask", atol=0, rtol=0, defer=True)
        close(d[..., 1:2], gold["prj_depth"], what=f"{name}/{tag}/depth", defer=True)
        close(d[..., 4:6], gold["prj_pts"], what=f"{name}/{tag}/pts", atol=PTS_ATOL, rtol=2e-6, defer=True)   # pixel units
        close(d[..., 6:9], gold["prj_dir"], what=f"{name}/{tag}/dir", defer=True)
        close(d[..., 9:12], gold["prj_rgb"], what=f"{name}/{tag}/rgb", defer=True)
        close(d[..., 12:44], gold["prj_ray_feats"], what=f"{name}/{tag}/ray_feats", defer=True)
        close(d[..., 44:76], gold["prj_img_feats"], what=f"{name}/{tag}/img_feats", defer=True)
        close(d[..., 2:3], gold["prj_hit_prob"], what=f"{name}/{tag}/hit_prob", defer=True)
        close(d[..., 3:4], gold["prj_vis"], what=f"{name}/{tag}/vis", defer=True)
    flush_fails()


@pytest.mark.parametrize("name", CASES)
def test_render_by_depth_matches_reference(name):
    """Whole passes (rows a8-a14) against the reference's render_impl outputs; the fine pass runs on the reference's
    own fine depths because searchsorted is discontinuous."""
    g = GoldenCase(name)
    net = build_path(g)
    ref, que = dev(g.ref), dev(g.que)
    out_c = net.render_by_depth(g.que_depth.cuda(), que, ref, g.is_train, False)
    out_f = net.render_by_depth(g.que_depth_fine.cuda(), que, ref, g.is_train, True)
    torch.cuda.synchronize()
    for suffix, out in (("", out_c), ("_fine", out_f)):
        for k in ("pixel_colors_nr", "hit_prob_nr", "render_depth", "pixel_colors_gt"):
            close(out[k], g.out[k + suffix], what=f"{name}/{k}{suffix}", defer=True)
        assert torch.equal(out["ray_mask"].cpu(), g.out["ray_mask" + suffix]), "ray_mask" + suffix
    flush_fails()


@pytest.mark.parametrize("name", CASES)
def test_fused_resampling(name):
    """sample_fine_depth + sort (row a15): stand-alone kernel on the reference's hit_prob (tight), and the fused
    emission of the coarse pass (searchsorted flips allowed on a tiny fraction of samples)."""
    g = GoldenCase(name)
    net = build_path(g)
    ref, que = dev(g.ref), dev(g.que)
    fdn = g.cfg["fine_depth_sample_num"]
    rn = g.que_depth.shape[1]
    use_all = bool(g.cfg.get("fine_depth_use_all", False))
    m = fdn + (g.que_depth.shape[2] if use_all else 0)
    if g.is_train:
        u, stride = g.fine_u[0].cuda().contiguous(), fdn
    else:
        u, stride = render_ops.fine_sample_u(fdn, "cuda"), 0
    out = torch.empty(rn, m, device="cuda")
    depth, hit = g.que_depth[0].cuda().contiguous(), g.out["hit_prob_nr"][0].cuda().contiguous()
    dr = g.que["depth_range"][0].cuda().contiguous()
    _lib.check(_lib.lib().nr_sample_fine_depth(depth.data_ptr(), hit.data_ptr(), dr.data_ptr(), rn,
                                               depth.shape[1], fdn, u.data_ptr(), stride, int(use_all), 1, out.data_ptr(), None), "fine")
    torch.cuda.synchronize()
    close(out[None], g.que_depth_fine, atol=2e-5, rtol=2e-5, what="standalone resample", max_bad_frac=2e-3)
    res = renderer.run_pass(net, g.que_depth.cuda(), que, ref, False,
                            fine={"dn": fdn, "use_all": use_all, "u": u, "u_stride": stride})
    torch.cuda.synchronize()
    close(res["fine_depth"], g.que_depth_fine, atol=1e-4, rtol=1e-3, what="fused resample", max_bad_frac=5e-3)


def psnr(a, b):
    mse = ((a - b) ** 2).mean().item()
    return 10 * torch.log10(torch.tensor(1.0 / max(mse, 1e-20))).item()


@pytest.mark.parametrize("name", CASES)
def test_render_end_to_end(name):
    """render() (chunk loop, both passes, fused resampling) against the reference's outputs.  End-to-end fine-pass
    values may jump where a quantile crosses a CDF knot, so this is judged by PSNR and by the bulk of the pixels."""
    g = GoldenCase(name)
    net = build_path(g)
    ref, que = dev(g.ref), dev(g.que)
    if g.is_train:
        torch.manual_seed(1234)          # render_impl draws the fine quantiles from torch's CPU generator like the reference
    net.cfg["ray_batch_num"] = 1 << 20 if g.is_train else 1500   # train: one chunk so the single rand() draw lines up
    out = net.render(que, ref, g.is_train)
    torch.cuda.synchronize()
    close(out["pixel_colors_nr"], g.out["pixel_colors_nr"], what="coarse colours")
    fine, gold = out["pixel_colors_nr_fine"].cpu(), g.out["pixel_colors_nr_fine"]
    assert psnr(fine, gold) > 60.0, psnr(fine, gold)
    close(fine, gold, what="fine colours", max_bad_frac=0.01)
    # how far the <= 1 % of pixels may be off where a quantile crossed a CDF knot: one resampled depth moves by at most one
    # coarse bin, i.e. one of the 16..64 fine samples of the ray changes its colour contribution
    err = (fine - gold).abs().flatten()
    q999 = torch.quantile(err, 0.999).item()
    assert q999 < 2e-3 and err.max().item() < 2e-2, (q999, err.max().item())
    expect = set(g.out) if g.is_train else {k for k in g.out if not k.startswith("hit_prob")}
    assert set(out) == expect, (set(out) ^ expect)


def test_against_oracle_mid_size():
    """Seeded mid-size case (8 views 96x128, 256 rays, 64+64 samples, gen_depth cfg) checked against the CPU oracle
    computed on the spot -- sizes the oracle finishes in seconds."""
    cfg = {"use_hierarchical_sampling": True, "dist_decoder_cfg": {"use_vis": False}, "render_depth": True}
    que, ref = synthetic.make_scene(96, 128, 8, seed=11, smooth=2)
    que = synthetic.slice_rays(que, 4000, 4256)
    W = synthetic.make_weights(cfg, seed=7)
    from gen_golden import flat_cfg
    ocfg = flat_cfg({**renderer.base_cfg, **cfg})
    gold = orc.render_impl(W, ocfg, que, ref, False)
    net = renderer.NeuralRayRenderPath(cfg)
    net.load_state_dict(W, strict=True)
    net.cuda()
    out = net.render_impl(dev(que), dev(ref), False)
    out_f = net.render_by_depth(gold["que_depth_fine"].cuda(), dev(que), dev(ref), False, True)
    torch.cuda.synchronize()
    for k in ("pixel_colors_nr", "hit_prob_nr", "render_depth"):
        close(out[k], gold[k], what=k)
        close(out_f[k], gold[k + "_fine"], what=k + "_fine(injected depths)")
    assert torch.equal(out["ray_mask"].cpu(), gold["ray_mask"])


@pytest.mark.parametrize("rfn", [1, 2, 3, 5, 6, 8, 9, 10, 16, 20, 32])
def test_view_counts(rfn):
    """Every lanes-per-point instantiation of the point kernel (4, 8, 16, 32 lanes; padding lanes when the view count
    is not a power of two; groups of 5, 6 and 10 lanes that are reduced through shared memory -- cfg4 of SURVEY.md 8d renders with 10 views)
    against the CPU oracle on a small seeded case."""
    cfg = {"use_hierarchical_sampling": True, "dist_decoder_cfg": {"use_vis": False}, "render_depth": True}
    que, ref = synthetic.make_scene(48, 64, rfn, seed=100 + rfn, smooth=2)
    que = synthetic.slice_rays(que, 1000, 1048)
    W = synthetic.make_weights(cfg, seed=3)
    from gen_golden import flat_cfg
    ocfg = flat_cfg({**renderer.base_cfg, **cfg})
    gold = orc.render_impl(W, ocfg, que, ref, False)
    net = renderer.NeuralRayRenderPath(cfg)
    net.load_state_dict(W, strict=True)
    net.cuda()
    out = net.render_impl(dev(que), dev(ref), False)
    out_f = net.render_by_depth(gold["que_depth_fine"].cuda(), dev(que), dev(ref), False, True)
    torch.cuda.synchronize()
    for k in ("pixel_colors_nr", "hit_prob_nr", "render_depth"):
        close(out[k], gold[k], what=f"{k} rfn={rfn}")
        close(out_f[k], gold[k + "_fine"], what=f"{k}_fine(injected depths) rfn={rfn}")
    assert torch.equal(out["ray_mask"].cpu(), gold["ray_mask"])


def _edge_case(rn, dn, rfn, look_away=False, seed=0):
    cfg = {"use_hierarchical_sampling": True, "depth_sample_num": dn, "fine_depth_sample_num": dn, "agg_net_cfg": {"sample_num": dn},
           "fine_agg_net_cfg": {"sample_num": dn}, "dist_decoder_cfg": {"use_vis": False}, "render_depth": True}
    que, ref = synthetic.make_scene(48, 64, rfn, seed=40 + seed, smooth=2)
    que = synthetic.slice_rays(que, 700, 700 + rn)
    if look_away:   # turn the query camera around: every sample projects behind / outside every reference view
        que["poses"] = que["poses"].clone()
        que["poses"][:, :, :3] = que["poses"][:, :, :3] * torch.tensor([1.0, 1.0, -1.0])[None, :, None]
    W = synthetic.make_weights(cfg, seed=seed)
    from gen_golden import flat_cfg
    gold = orc.render_impl(W, flat_cfg({**renderer.base_cfg, **cfg}), que, ref, False)
    net = renderer.NeuralRayRenderPath(cfg)
    net.load_state_dict(W, strict=True)
    net.cuda()
    with torch.no_grad():
        out = net.render_impl(dev(que), dev(ref), False)
        out_f = net.render_by_depth(gold["que_depth_fine"].cuda(), dev(que), dev(ref), False, True)
    torch.cuda.synchronize()
    for k in ("pixel_colors_nr", "hit_prob_nr", "render_depth"):
        close(out[k], gold[k], what=f"{k} rn={rn} dn={dn}")
        close(out_f[k], gold[k + "_fine"], what=f"{k}_fine rn={rn} dn={dn}")
    assert torch.equal(out["ray_mask"].cpu(), gold["ray_mask"])
    return out, gold


@pytest.mark.parametrize("rn,dn", [(1, 3), (7, 5), (130, 16), (3, 200)])
def test_ragged_and_extreme_sizes(rn, dn):
    """One ray, ray counts that are not a multiple of any tile, the minimum of 3 samples (sample_depth asserts dn > 2,
    render_ops.py:147) and 200 samples per ray (limit 256), against the oracle."""
    _edge_case(rn, dn, 3, seed=rn)


def test_rays_looking_away_from_the_scene():
    """Query camera turned around: most (point, view) pairs are masked or project from behind the reference cameras
    (negative depth stays valid in the reference, render_ops.py:97-103); no ray passes the ray_mask test
    (renderer.py:195-200).  Parity with the oracle and no NaN from nearly empty pools / masked softmax rows."""
    out, gold = _edge_case(24, 8, 4, look_away=True, seed=5)
    assert not bool(gold["ray_mask"].any())
    for v in out.values():
        assert bool(torch.isfinite(v.float()).all())


def test_empty_chunk():
    """Zero rays: every entry point returns empty outputs instead of launching."""
    cfg = {"use_hierarchical_sampling": True, "render_depth": True}
    que, ref = synthetic.make_scene(48, 64, 3, seed=2, smooth=2)
    que = synthetic.slice_rays(que, 10, 10)
    net = renderer.NeuralRayRenderPath(cfg)
    net.load_state_dict(synthetic.make_weights(cfg, seed=0), strict=True)
    net.cuda()
    with torch.no_grad():
        out = net.render_impl(dev(que), dev(ref), False)
    assert out["pixel_colors_nr"].shape == (1, 0, 3) and out["pixel_colors_nr_fine"].shape == (1, 0, 3)
