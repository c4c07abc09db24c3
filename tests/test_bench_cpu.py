"""CPU: bench.py's workload table against BASELINE.json / SURVEY.md section 8d (shapes only; nothing is rendered)."""
import json
import os

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_headline_workload_is_the_configuration_the_metric_is_quoted_on():
    metric = json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]
    assert "800" in metric and "64 coarse+64 fine" in metric and "8 ref views" in metric
    wl = bench.WORKLOADS["black_800"]
    assert (wl["scene"]["h"], wl["scene"]["w"], wl["rfn"], wl["dn"]) == (800, 800, 8, (64, 64))
    cfg = bench.model_cfg(*wl["dn"])
    assert cfg["dist_decoder_cfg"] == {"use_vis": False} and cfg["use_hierarchical_sampling"]      # configs/gen/neuray_gen_depth.yaml


def test_workload_shapes():
    """cfg4: 1008x756 query, references padded to 1008x768, 10 views, depth (1.2, 12); cfg5: 300x400 padded to 304x400, (0.8, 4.0)."""
    que, ref = bench.make_workload("fern_high", seed=7)
    assert que["coords"].shape == (1, 1008 * 756, 2) and tuple(ref["imgs"].shape) == (10, 3, 768, 1008)
    assert tuple(ref["ray_feats"].shape) == (10, 32, 192, 252) and tuple(que["depth_range"][0].tolist()) == (1.2000000476837158, 12.0)
    que, ref = bench.make_workload("train_dtu", seed=5, with_que_imgs=True)
    assert tuple(ref["imgs"].shape) == (8, 3, 304, 400) and tuple(que["imgs"].shape) == (1, 3, 300, 400)
    assert abs(float(que["depth_range"][0, 0]) - 0.8) < 1e-6 and bench.TRAIN_RAYS == 512
    que, ref = bench.make_workload("cfg1", seed=0)
    assert que["coords"].shape[1] == 64 * 64 and ref["imgs"].shape[0] == 3


def test_flop_model_matches_survey():
    assert bench.point_kernel_flops_per_sample(8, False) == 465792            # SURVEY.md 8a / DESIGN.md 3.2
    assert bench.point_kernel_flops_per_sample(8, True) == 465792 + 8 * 4160
