"""CPU: patch.install() rebinds the reference's API onto the B200 path (only where /root/reference exists)."""
import pytest
import torch

import ref_import
from neuray_b200 import _lib, patch, render_ops, renderer


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
def test_install_rebinds_reference_api():
    ref_renderer = ref_import.load_reference()
    orig = {k: getattr(ref_renderer.NeuralRayBaseRenderer, k) for k in ("render_by_depth", "render_impl", "fine_render_impl", "render")}
    import network.render_ops as ref_ops
    orig_ops = {n: getattr(ref_ops, n) for n in render_ops.__all__ if hasattr(ref_ops, n)}
    orig_ren = {n: getattr(ref_renderer, n) for n in render_ops.__all__ if hasattr(ref_renderer, n)}
    import network.loss as ref_loss
    from neuray_b200 import losses
    orig_loss = dict(ref_loss.name2loss)
    orig_pm = ref_renderer.NeuralRayGenRenderer.predict_mean_for_depth_loss
    try:
        base = patch.install()
        assert base.render_impl is renderer.render_impl and base.render is renderer.render
        assert ref_loss.name2loss == losses.name2loss and ref_loss.DepthLoss is losses.DepthLoss
        assert ref_renderer.NeuralRayGenRenderer.predict_mean_for_depth_loss is losses.predict_mean_for_depth_loss
        for n in orig_ops:
            assert getattr(ref_ops, n) is getattr(render_ops, n), n
        # the reference's own constructor still builds the network; the hot path now refuses CPU tensors (no fallback)
        net = base({"use_hierarchical_sampling": True})
        assert set(k.split(".")[0] for k in net.state_dict()) >= {"dist_decoder", "agg_net", "fine_dist_decoder", "fine_agg_net"}
        from neuray_b200 import synthetic
        que, ref = synthetic.make_scene(32, 32, 3, seed=0)
        with pytest.raises(_lib.NeurayB200Error):
            net.render_impl(que, dict(ref), False)
    finally:
        patch.uninstall()
    for k, v in orig.items():
        assert getattr(ref_renderer.NeuralRayBaseRenderer, k) is v, k
    assert ref_loss.name2loss == orig_loss and ref_renderer.NeuralRayGenRenderer.predict_mean_for_depth_loss is orig_pm
    for n, v in orig_ops.items():
        assert getattr(ref_ops, n) is v, n
    for n, v in orig_ren.items():
        assert getattr(ref_renderer, n) is v, n
