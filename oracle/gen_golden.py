"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by executing the UNMODIFIED reference
(/root/reference, imported through oracle/ref_import.py) on seeded synthetic inputs.

Run in the build container (the reference tree does not exist on the GPU box):

    python oracle/gen_golden.py            # writes tests/golden/case_*.npz

Each file stores the inputs (que/ref imgs_info tensors, weights under state-dict names, cfg as json) and
the reference's outputs: whole-chunk outputs of NeuralRayBaseRenderer.render_impl (renderer.py:217-226) and,
for a small ray subset, every intermediate of render_by_depth (prj_dict fields, density, colours) so that
both the oracle (tests/test_oracle_golden.py) and the CUDA path (tests/test_parity_gpu.py) can be checked
stage by stage.
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

import ref_import  # noqa: E402
from neuray_b200 import synthetic  # noqa: E402

CASES = {
    # BASELINE.json configs[0]: 64x64 query, 3 ref views, 32 coarse samples (+32 fine), gen_depth-style cfg
    "cfg1": dict(
        scene=dict(h=64, w=64, rfn=3, focal=80.0, seed=0, smooth=1),
        cfg={"use_hierarchical_sampling": True, "dist_decoder_cfg": {"use_vis": False}, "depth_sample_num": 32,
             "fine_depth_sample_num": 32, "agg_net_cfg": {"sample_num": 32}, "fine_agg_net_cfg": {"sample_num": 32},
             "render_depth": True, "ray_batch_num": 4096},
        is_train=False, stage_rays=64, seed=0),
    # ragged / train-mode variant: 8 views, non-square padded refs, use_vis on both decoders, fine_depth_use_all,
    # random fine quantiles (training), smooth feature maps
    "train8": dict(
        scene=dict(h=40, w=52, rfn=8, que_h=24, que_w=28, seed=3, smooth=2, depth_range=(2.5, 5.5), arc_deg=80.0),
        cfg={"use_hierarchical_sampling": True, "depth_sample_num": 24, "fine_depth_sample_num": 16,
             "fine_depth_use_all": True, "agg_net_cfg": {"sample_num": 24}, "fine_agg_net_cfg": {"sample_num": 40},
             "render_depth": True, "ray_batch_num": 1024},
        is_train=True, stage_rays=32, seed=5),
    # SURVEY.md 8d cfg4 in miniature: 10 reference views (not a power of two: padding lanes in the point kernel), wide
    # COLMAP-like depth range, non-square images
    "views10": dict(
        scene=dict(h=48, w=64, rfn=10, que_h=20, que_w=24, seed=7, smooth=2, depth_range=(1.2, 12.0), arc_deg=100.0),
        cfg={"use_hierarchical_sampling": True, "dist_decoder_cfg": {"use_vis": False}, "depth_sample_num": 16,
             "fine_depth_sample_num": 16, "agg_net_cfg": {"sample_num": 16}, "fine_agg_net_cfg": {"sample_num": 16},
             "render_depth": True, "ray_batch_num": 512},
        is_train=False, stage_rays=24, seed=11),
}


def flat_cfg(cfg):
    """The oracle's flat cfg (oracle/neuray_oracle.py DEFAULT_CFG) from a reference-style nested cfg."""
    from neuray_oracle import DEFAULT_CFG
    out = dict(DEFAULT_CFG)
    for k in out:
        if k in cfg:
            out[k] = cfg[k]
    out["dist_decoder_use_vis"] = cfg.get("dist_decoder_cfg", {}).get("use_vis", True)
    out["fine_dist_decoder_use_vis"] = cfg.get("fine_dist_decoder_cfg", {}).get("use_vis", True)
    out["dist_decoder_bias_val"] = cfg.get("dist_decoder_cfg", {}).get("bias_val", 0.05)
    out["fine_dist_decoder_bias_val"] = cfg.get("fine_dist_decoder_cfg", {}).get("bias_val", 0.05)
    out["agg_sample_num"] = cfg.get("agg_net_cfg", {}).get("sample_num", 64)
    out["fine_agg_sample_num"] = cfg.get("fine_agg_net_cfg", {}).get("sample_num", 64)
    return out


def run_case(name, spec, ref_renderer_mod):
    que, ref = synthetic.make_scene(**spec["scene"])
    W = synthetic.make_weights(spec["cfg"], seed=spec["seed"])
    torch.manual_seed(0)
    net = ref_renderer_mod.NeuralRayBaseRenderer(spec["cfg"])
    missing, unexpected = net.load_state_dict(W, strict=False)
    assert not unexpected, unexpected
    hot = [k for k in missing if k.split(".")[0] in ("dist_decoder", "agg_net", "fine_dist_decoder", "fine_agg_net")]
    assert not hot, hot
    net.eval()
    is_train = spec["is_train"]
    out = {}
    with torch.no_grad():
        q = {k: v.clone() for k, v in que.items()}
        r = {k: v.clone() for k, v in ref.items()}
        fdn = spec["cfg"]["fine_depth_sample_num"]
        rn = q["coords"].shape[1]
        if is_train:
            torch.manual_seed(1234)
            u = torch.rand([1, rn, fdn])          # the draw sample_fine_depth will make (render_ops.py:205)
            torch.manual_seed(1234)
            out["fine_u"] = u
        res = net.render_impl(q, r, is_train)
        for k, v in res.items():
            out["out_" + k] = v
        # fine-pass depths as the reference computed them (renderer.py:205-213)
        que_depth, _ = ref_renderer_mod.sample_depth(q["depth_range"], q["coords"], spec["cfg"]["depth_sample_num"], False)
        out["que_depth"] = que_depth
        if is_train:
            torch.manual_seed(1234)
        fd = ref_renderer_mod.sample_fine_depth(que_depth, res["hit_prob_nr"], q["depth_range"], fdn, is_train)
        if spec["cfg"].get("fine_depth_use_all", False):
            fd = torch.cat([que_depth, fd], -1)
        out["que_depth_fine"] = torch.sort(fd, -1)[0]

        # stage-level intermediates on a ray subset, coarse and fine pass
        n = spec["stage_rays"]
        sel = torch.linspace(0, rn - 1, n).long()
        out["stage_sel"] = sel
        qs = {k: v.clone() for k, v in que.items()}
        qs["coords"] = que["coords"][:, sel]
        for tag, depth, is_fine in (("c", que_depth[:, sel], False), ("f", out["que_depth_fine"][:, sel], True)):
            r2 = {k: v.clone() for k, v in ref.items()}
            que_dists = ref_renderer_mod.depth2inv_dists(depth, qs["depth_range"])
            que_pts, que_dir = ref_renderer_mod.depth2points(qs, depth)
            prj = ref_renderer_mod.project_points_dict(r2, que_pts)
            prj = net.predict_proj_ray_prob(prj, r2, que_dists, is_fine)
            prj = net.get_img_feats(r2, prj)
            agg = net.fine_agg_net if is_fine else net.agg_net
            density, colors = agg(prj, que_dir)
            st = {"que_dists": que_dists, "que_pts": que_pts, "que_dir": que_dir, "density": density, "colors": colors}
            st.update({"prj_" + k: v for k, v in prj.items() if k != "alpha"})
            for k, v in st.items():
                out[f"stage_{tag}_{k}"] = v
    blob = {}
    for k, v in que.items():
        blob["que_" + k] = v.numpy()
    for k, v in ref.items():
        blob["ref_" + k] = v.numpy()
    for k, v in W.items():
        blob["W_" + k] = v.numpy()
    for k, v in out.items():
        blob[k] = v.numpy()
    blob["cfg_json"] = np.frombuffer(json.dumps(spec["cfg"]).encode(), dtype=np.uint8)
    blob["is_train"] = np.array(int(is_train))
    path = os.path.join(ROOT, "tests", "golden", f"case_{name}.npz")
    np.savez_compressed(path, **blob)
    print(name, "->", path, "%.1f KB" % (os.path.getsize(path) / 1024), "keys", len(blob))
    print("   pixel_colors_nr mean %.4f  hit sum mean %.4f  ray_mask frac %.3f" % (
        out["out_pixel_colors_nr"].mean(), out["out_hit_prob_nr"].sum(-1).mean(), out["out_ray_mask"].float().mean()))


def main():
    mod = ref_import.load_reference()
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name, spec in CASES.items():
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        run_case(name, spec, mod)


if __name__ == "__main__":
    main()
