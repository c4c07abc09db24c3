"""GPU: the library-side packers (csrc/nr_pack.cu) and a complete render pass driven through the raw C-ABI only.

  * nr_pack_weights  == the round-1 PyTorch packers (tests/ref_packers.py), bit for bit, all three buffers
  * nr_camera_blocks == the reference's torch expressions (render_ops.py:14-20, 95, 112) to a few ulp
  * one coarse pass + fused resampling + one fine pass with nothing but ctypes calls into libneuray_b200.so (no
    neuray_b200.weights / renderer code on the path): a host in any language can do the same
  * one process, two devices: the library keeps no per-process / per-device state
"""
import ctypes as C

import pytest
import torch

import ref_packers
from neuray_b200 import _lib, renderer, synthetic
from neuray_b200 import weights as nr_weights

pytestmark = pytest.mark.gpu


def split(W, dec, agg):
    return {k: v for k, v in W.items() if k.startswith(dec + ".") or k.startswith(agg + ".")}


@pytest.mark.parametrize("use_vis", [True, False])
def test_pack_weights_equals_the_pytorch_packers(use_vis):
    cfg = {"use_hierarchical_sampling": True, "dist_decoder_cfg": {"use_vis": use_vis}}
    W = synthetic.make_weights(cfg, seed=2)
    for dec, agg in (("dist_decoder", "agg_net"), ("fine_dist_decoder", "fine_agg_net")):
        params = split(W, dec, agg)
        rp, rr = ref_packers.pack_pass_weights(params, dec, agg, torch.device("cpu"))
        rt = ref_packers.pack_tc_weights(params, dec, agg, torch.device("cpu"))
        wp, wr, wt = nr_weights.pack_pass({k: v.cuda() for k, v in params.items()}, dec, agg)
        torch.cuda.synchronize()
        assert torch.equal(wp.cpu(), rp), (wp.cpu() != rp).nonzero()[:5]
        assert torch.equal(wr.cpu(), rr)
        diff = (wt.cpu() != rt).nonzero().flatten()
        # the 8x32 composed block (neuray_fc.0 @ prob_embed.2, fp64 accumulate, rounded once) may differ from torch's fp64
        # matmul in the last bit of an element if the summation order differs: allow that there and nowhere else
        T = _lib.tc_layout()
        assert all(T.pe1 <= int(i) < T.pe1 + 3072 for i in diff), diff[:8]
        assert torch.allclose(wt.cpu(), rt, rtol=2e-7, atol=0)
        assert diff.numel() <= 4


def test_pack_weights_follows_parameter_updates():
    """Re-pack after an in-place parameter update (what an optimizer step does) picks up the new values."""
    W = {k: v.cuda() for k, v in synthetic.make_weights({}, seed=3).items()}
    a = nr_weights.pack_pass(W, "dist_decoder", "agg_net")
    for v in W.values():
        v.mul_(1.5).add_(0.01)
    b = nr_weights.pack_pass(W, "dist_decoder", "agg_net")
    cpu = {k: v.cpu() for k, v in W.items()}
    rp, rr = ref_packers.pack_pass_weights(cpu, "dist_decoder", "agg_net", torch.device("cpu"))
    assert torch.equal(b[0].cpu(), rp) and torch.equal(b[1].cpu(), rr) and not torch.equal(a[0], b[0])


def test_camera_blocks_match_the_reference_expressions():
    que, ref = synthetic.make_scene(96, 128, 10, seed=4, depth_range=(1.2, 12.0), arc_deg=100.0)
    cam, vp = nr_weights.camera_blocks(synthetic.to_device(que, "cuda"), synthetic.to_device(ref, "cuda"))
    torch.cuda.synchronize()
    want_cam = ref_packers.camera_block(que["poses"][0], que["Ks"][0], que["depth_range"][0])
    want_vp = ref_packers.view_param_block(ref["poses"], ref["Ks"], ref["depth_range"])
    # fp64 evaluation rounded once vs the reference's fp32 matmul / LU inverse: a few ulp of the entry's scale
    assert torch.allclose(cam.cpu(), want_cam, rtol=5e-7, atol=2e-7), (cam.cpu() - want_cam).abs().max()
    scale = want_vp[:, :12].abs().max()
    assert float((vp.cpu()[:, :12] - want_vp[:, :12]).abs().max()) <= 4e-7 * float(scale)
    assert torch.allclose(vp.cpu()[:, 12:], want_vp[:, 12:], rtol=5e-7, atol=2e-7)


# ---- a host that only knows the header ------------------------------------------------------------------------------

HEADS = ("mean_decoder", "var_decoder", "aw_decoder", "vis_decoder")
FIELDS = (("ray_dir_fc", (0, 2)), ("neuray_fc", (0, 2)), ("base_fc", (0, 2)), ("vis_fc", (0, 2)), ("vis_fc2", (0, 2)), ("rgb_fc", (0, 2, 4)),
          ("geometry_fc", (0, 2)), ("out_geometry_fc", (0, 2)))


def weight_struct(P, dec, agg):
    """NrPassWeights from a state dict, written against include/neuray_b200.h only."""
    w = _lib.NrPassWeights()
    for h, head in enumerate(HEADS):
        if f"{dec}.{head}.0.weight" in P:
            for l, i in enumerate((0, 2, 4)):
                w.dist_decoder[h][l].w, w.dist_decoder[h][l].b = P[f"{dec}.{head}.{i}.weight"].data_ptr(), P[f"{dec}.{head}.{i}.bias"].data_ptr()
    for l, i in enumerate((0, 2)):
        w.prob_embed[l].w, w.prob_embed[l].b = P[f"{agg}.prob_embed.{i}.weight"].data_ptr(), P[f"{agg}.prob_embed.{i}.bias"].data_ptr()
    for name, idx in FIELDS:
        for l, i in enumerate(idx):
            getattr(w, name)[l].w = P[f"{agg}.agg_impl.{name}.{i}.weight"].data_ptr()
            getattr(w, name)[l].b = P[f"{agg}.agg_impl.{name}.{i}.bias"].data_ptr()
    at = f"{agg}.agg_impl.ray_attention"
    w.w_qs, w.w_ks, w.w_vs = (P[f"{at}.{n}.weight"].data_ptr() for n in ("w_qs", "w_ks", "w_vs"))
    w.attn_fc, w.layer_norm_w, w.layer_norm_b = P[f"{at}.fc.weight"].data_ptr(), P[f"{at}.layer_norm.weight"].data_ptr(), P[f"{at}.layer_norm.bias"].data_ptr()
    return w


def raw_abi_render(W, cfg, que, ref, dev):
    """coarse pass (+ fused resampling) and fine pass with ctypes calls only; returns dict of tensors."""
    lib = _lib.lib()
    ck = _lib.check
    f32 = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        P = {k: v.to(dev).contiguous() for k, v in W.items()}
        q = {k: v.to(dev).contiguous() for k, v in que.items()}
        r = {k: v.to(dev).contiguous() for k, v in ref.items()}
        L, T = _lib.weight_layout(), _lib.tc_layout()
        rfn, _, h, w = r["imgs"].shape
        fh, fw = r["ray_feats"].shape[-2:]
        rn = q["coords"].shape[1]
        dn, fdn = cfg["depth_sample_num"], cfg["fine_depth_sample_num"]
        feat, rgb = f32(rfn, fh, fw, 64), f32(rfn, h, w, 4)
        ck(lib.nr_pack_feature_maps(r["ray_feats"].data_ptr(), r["img_feats"].data_ptr(), r["imgs"].data_ptr(), rfn, h, w, fh, fw,
                                    feat.data_ptr(), rgb.data_ptr(), stream), "pack maps")
        cam, vp = f32(24), f32(rfn, 20)
        ck(lib.nr_camera_blocks(q["poses"].data_ptr(), q["Ks"].data_ptr(), q["depth_range"].data_ptr(), r["poses"].data_ptr(),
                                r["Ks"].data_ptr(), r["depth_range"].data_ptr(), rfn, cam.data_ptr(), vp.data_ptr(), stream), "camera blocks")
        packed = {}
        for tag, dec, agg in (("c", "dist_decoder", "agg_net"), ("f", "fine_dist_decoder", "fine_agg_net")):
            bufs = (f32(L.total_point), f32(L.total_ray), f32(T.total))
            ws = weight_struct(P, dec, agg)
            ck(lib.nr_pack_weights(C.byref(ws), bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr(), stream), "pack weights")
            packed[tag] = bufs
        depth = f32(rn, dn)
        ck(lib.nr_sample_depth(q["depth_range"].data_ptr(), rn, dn, None, depth.data_ptr(), None, stream), "sample_depth")
        u = (0.5 * (1 / fdn) + torch.arange(fdn) * (1 / fdn)).to(dev).contiguous()      # render_ops.py:199-202
        fine_depth = f32(rn, fdn)
        out = {}
        for tag, qd, sample_num in (("c", depth, dn), ("f", fine_depth, fdn)):
            pe = nr_posenc(sample_num).to(dev)
            rec = f32(rn * sample_num * _lib.NR_POINT_REC)
            o = {"pixel_colors": f32(rn, 3), "hit_prob": f32(rn, sample_num), "render_depth": f32(rn),
                 "ray_mask": torch.empty(rn, dtype=torch.uint8, device=dev)}
            p = _lib.NrPassParams()
            p.coords, p.que_depth, p.que_cam, p.rn, p.dn = q["coords"].data_ptr(), qd.data_ptr(), cam.data_ptr(), rn, sample_num
            p.feat, p.rgb, p.view_params = feat.data_ptr(), rgb.data_ptr(), vp.data_ptr()
            p.rfn, p.h, p.w, p.fh, p.fw = rfn, h, w, fh, fw
            p.w_point, p.w_ray, p.w_tc = (b.data_ptr() for b in packed[tag])
            p.pos_enc = pe.data_ptr()
            p.use_vis, p.var_bias = 0, 0.05
            p.ray_mask_view_num, p.ray_mask_point_num = 2, 8
            p.point_rec = rec.data_ptr()
            p.pixel_colors, p.hit_prob = o["pixel_colors"].data_ptr(), o["hit_prob"].data_ptr()
            p.render_depth, p.ray_mask = o["render_depth"].data_ptr(), o["ray_mask"].data_ptr()
            if tag == "c":
                p.fine_dn, p.fine_use_all, p.fine_u, p.fine_u_stride, p.fine_depth = fdn, 0, u.data_ptr(), 0, fine_depth.data_ptr()
            ck(lib.nr_render_pass_fwd(C.byref(p), stream), "render pass")
            out[tag] = o
        torch.cuda.synchronize()
    return out, fine_depth


def nr_posenc(n):
    """sinusoid table of ibrnet.py:305-313 (what NrPassParams.pos_enc expects), written out here"""
    import numpy as np
    pos = np.arange(n, dtype=np.float64)[:, None]
    j = np.arange(16)[None, :]
    t = pos / np.power(10000, 2 * (j // 2) / 16)
    t[:, 0::2] = np.sin(t[:, 0::2])
    t[:, 1::2] = np.cos(t[:, 1::2])
    return torch.from_numpy(t).float().contiguous()


CFG = {"use_hierarchical_sampling": True, "dist_decoder_cfg": {"use_vis": False}, "depth_sample_num": 32, "fine_depth_sample_num": 32,
       "agg_net_cfg": {"sample_num": 32}, "fine_agg_net_cfg": {"sample_num": 32}, "render_depth": True}


def test_full_render_through_the_raw_c_abi():
    que, ref = synthetic.make_scene(48, 64, 5, seed=12, smooth=2, with_que_imgs=False)
    que = synthetic.slice_rays(que, 900, 1100)
    W = synthetic.make_weights(CFG, seed=2)
    out, fine_depth = raw_abi_render(W, CFG, que, ref, torch.device("cuda", 0))
    net = renderer.NeuralRayRenderPath(CFG)
    net.load_state_dict(W, strict=True)
    net.cuda()
    with torch.no_grad():
        want = net.render_impl(synthetic.to_device(que, "cuda"), synthetic.to_device(ref, "cuda"), False)
    torch.cuda.synchronize()
    # the same kernels on the same packed inputs: identical bits
    assert torch.equal(out["c"]["pixel_colors"][None], want["pixel_colors_nr"])
    assert torch.equal(out["f"]["pixel_colors"][None], want["pixel_colors_nr_fine"])
    assert torch.equal(out["f"]["render_depth"][None], want["render_depth_fine"])
    assert torch.equal(out["f"]["ray_mask"][None].bool(), want["ray_mask_fine"])
    # and the oracle agrees
    import neuray_oracle as orc
    from gen_golden import flat_cfg
    gold = orc.render_impl(W, flat_cfg({**renderer.base_cfg, **CFG}), que, ref, False)
    assert float((out["c"]["pixel_colors"][None].cpu() - gold["pixel_colors_nr"]).abs().max()) < 1e-4


def test_two_devices_in_one_process():
    """cuda:1 after cuda:0 from one process and one thread: same bits on both (per-device kernel attributes, current-device
    handling).  Needs a box with >= 2 GPUs (gpurun --gpus 2)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    que, ref = synthetic.make_scene(48, 64, 8, seed=13, smooth=2, with_que_imgs=False)
    que = synthetic.slice_rays(que, 100, 400)
    W = synthetic.make_weights(CFG, seed=5)
    outs = []
    for d in (0, 1, 0):
        net = renderer.NeuralRayRenderPath(CFG)
        net.load_state_dict(W, strict=True)
        net.to(f"cuda:{d}")
        with torch.no_grad():     # the current device stays cuda:0 throughout: the package switches per call
            o = net.render(synthetic.to_device(que, f"cuda:{d}"), synthetic.to_device(ref, f"cuda:{d}"), False)
        outs.append({k: v.cpu() for k, v in o.items()})
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]) and torch.equal(outs[0][k], outs[2][k]), k
