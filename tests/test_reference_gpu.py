"""GPU: the drop-in proof.  The UNMODIFIED reference (baseline/_ref, the verbatim copy made by baseline/install_ref.py; or
/root/reference in the build container) is imported, its own NeuralRayGenRenderer is built and run under CUDA twice -- as
is, and after neuray_b200.patch.install() -- on the same seeded inputs, and everything the reference's callers read is
compared: every output key of forward() in eval and training mode, and after backward() of the reference's kind of loss
(render + depth terms, self-hit-prob term) the gradient of EVERY parameter of the network (hot-path modules, image
encoder, vis encoder, init net) -- i.e. also what flows out of the CUDA path into the PyTorch encoders in front of it.

This is the test of SURVEY.md section 8b: `render.py` / `run_training.py` construct the network by name, call
`network(data)` and `loss.backward()`; nothing else touches the renderer.
"""
import numpy as np
import pytest
import torch

import ref_import
from neuray_b200 import patch, synthetic

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def fp32_convolutions():
    """cuDNN runs fp32 convolutions through TF32 by default; the encoders in front of the path would then turn a 1e-6
    difference in the feature-map gradients into a 1e-2 difference of their own weight gradients (sums with heavy
    cancellation), which says nothing about the path under test."""
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old

DN = 24
CFG = {"init_net_type": "depth", "use_hierarchical_sampling": True, "use_depth_loss": True, "dist_decoder_cfg": {"use_vis": False},
       "depth_sample_num": DN, "fine_depth_sample_num": DN, "agg_net_cfg": {"sample_num": DN}, "fine_agg_net_cfg": {"sample_num": DN},
       "ray_batch_num": 192, "render_depth": True, "depth_loss_coords_num": 256}


@pytest.fixture(scope="module")
def ref_mod():
    if not ref_import.available():
        pytest.fail("the reference tree is missing: run `python baseline/install_ref.py` before gpurun (baseline/_ref travels to the box)")
    mod = ref_import.load_reference()
    yield mod
    patch.uninstall()


def make_data(rfn=6, rays=384, self_feats=False, seed=3):
    que, ref = synthetic.make_scene(64, 80, rfn, seed=seed, smooth=2)
    ref = dict(ref)
    ref.pop("ray_feats"), ref.pop("img_feats")
    rs = np.random.RandomState(seed)
    ref["depth"] = torch.from_numpy(rs.uniform(2.5, 5.0, (rfn, 1, 64, 80)).astype(np.float32))
    ref["true_depth"] = ref["depth"]
    n = que["coords"].shape[1]
    idx = torch.from_numpy(rs.permutation(n)[:rays])
    que = dict(que, coords=que["coords"][:, idx].contiguous())
    if self_feats:
        que["ray_feats"] = torch.from_numpy(rs.standard_normal((1, 32, 16, 20)).astype(np.float32))
    return que, ref


def build(ref_mod, cfg, seed=0):
    torch.manual_seed(seed)
    net = ref_mod.NeuralRayGenRenderer(cfg)
    # hot-path modules: random init with a density head that terminates rays gradually (the default head of make_weights
    # multiplies the last density layer by 8, which also multiplies fp32 summation-order noise between cuBLAS and the kernels)
    W = synthetic.make_weights(cfg, seed=seed, sigma_gain=1.0, sigma_bias=0.02)
    missing, unexpected = net.load_state_dict(W, strict=False)
    assert not unexpected
    return net.cuda()


def run(net, que, ref, is_train, seed=77):
    data = {"que_imgs_info": synthetic.to_device(que, "cuda"), "ref_imgs_info": synthetic.to_device(ref, "cuda")}
    if not is_train:
        data["eval"] = True
    torch.manual_seed(seed)            # randperm of the depth-loss coordinates, torch.rand of the fine quantiles
    return net(data)


def loss_of(out):
    """The shape of the reference's training loss (network/loss.py:46-132): render terms on both passes + a depth term on
    the decoder means + the self-hit-prob consistency term when present."""
    loss = ((out["pixel_colors_nr"] - out["pixel_colors_gt"]) ** 2).mean() + ((out["pixel_colors_nr_fine"] - out["pixel_colors_gt_fine"]) ** 2).mean()
    loss = loss + 0.1 * out["depth_mean"].abs().mean() + 0.1 * out["depth_mean_fine"].abs().mean()
    if "hit_prob_self" in out:
        loss = loss + ((out["hit_prob_self"] - out["hit_prob_nr"].detach()) ** 2).mean() + ((out["hit_prob_self_fine"] - out["hit_prob_nr_fine"].detach()) ** 2).mean()
    return loss


def compare_outputs(a, b, fine_bad_frac):
    assert set(a) == set(b), set(a) ^ set(b)
    for k in a:
        x, y = a[k].detach().float().cpu(), b[k].detach().float().cpu()
        assert x.shape == y.shape, k
        if a[k].dtype == torch.bool:
            assert (x != y).float().mean().item() <= (fine_bad_frac if k.endswith("_fine") else 0.0), k
            continue
        bad = ((x - y).abs() > 1e-4 + 1e-3 * y.abs()).float().mean().item()
        # fine-pass quantities sit behind searchsorted: isolated samples may move by a bin
        allowed = fine_bad_frac if ("fine" in k) else 0.0
        assert bad <= allowed, (k, bad, float((x - y).abs().max()))


def test_eval_forward_patched_equals_unpatched(ref_mod):
    que, ref = make_data()
    net = build(ref_mod, CFG).eval()
    with torch.no_grad():
        want = run(net, que, ref, False)
        patch.install()
        try:
            got = run(net, que, ref, False)
        finally:
            patch.uninstall()
    torch.cuda.synchronize()
    assert "hit_prob_nr" not in got and "depth_mean_fine" in got
    compare_outputs(got, want, fine_bad_frac=0.01)


class GradTap:
    """Captures (and optionally perturbs) the gradients that flow OUT of the rendering path into the encoders in front of it:
    d loss / d output of every image_encoder / vis_encoder call (the path's inputs img_feats / ray_feats)."""

    def __init__(self, net, noise=None):
        self.grads, self.noise, self.handles = [], noise, []
        for name in ("image_encoder", "vis_encoder"):
            self.handles.append(getattr(net, name).register_forward_hook(self._fwd(name)))

    def _fwd(self, name):
        def hook(module, inputs, output):
            idx = len(self.grads)
            self.grads.append(None)

            def on_grad(g):
                self.grads[idx] = (name, g.detach().clone())
                if self.noise is not None:      # additive noise of a given size relative to the map's largest entry
                    gen = torch.Generator(device=g.device).manual_seed(1000 + idx)
                    r = torch.rand(g.shape, device=g.device, generator=gen) * 2 - 1
                    return g + self.noise[idx] * g.abs().max() * r
            output.register_hook(on_grad)
        return hook

    def close(self):
        for h in self.handles:
            h.remove()


@pytest.mark.parametrize("self_hit", [False, True])
def test_training_forward_and_every_gradient(ref_mod, self_hit):
    """forward() in training mode + backward() of the reference's kind of loss, unpatched vs patched:
      * every output key;
      * every parameter of the hot-path modules (dist_decoder, agg_net and their fine twins), directly;
      * the gradients the path hands to the encoders in front of it (d ray_feats, d img_feats), directly;
      * every parameter upstream of the path (image encoder, vis encoder, init net).  Their gradients come out of the
        reference's own PyTorch backward through up to ~20 convolution + InstanceNorm layers, which amplifies a last-bit
        difference in the feature-map gradients by orders of magnitude (two fp32 implementations of the same sum differ in
        the last bits); the yardstick is therefore the reference itself: its backward is run once more with the measured
        feature-map gradient difference injected as noise, and the patched network has to stay within a small multiple of
        what that does to each parameter."""
    cfg = dict(CFG, use_self_hit_prob=self_hit)
    que, ref = make_data(self_feats=self_hit)
    net = build(ref_mod, cfg).train()
    grads, outs, taps = {}, {}, {}

    def run_mode(mode, noise=None):
        tap = GradTap(net, noise)
        if mode == "patched":
            patch.install()
        try:
            net.zero_grad(set_to_none=True)
            out = run(net, que, ref, True)
            loss_of(out).backward()
            torch.cuda.synchronize()
        finally:
            patch.uninstall()
            tap.close()
        outs[mode] = {k: v.detach().clone() for k, v in out.items()}
        grads[mode] = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in net.named_parameters()}
        taps[mode] = tap.grads

    run_mode("reference")
    run_mode("patched")
    compare_outputs(outs["patched"], outs["reference"], fine_bad_frac=0.01)
    # what the path hands upstream
    assert len(taps["patched"]) == len(taps["reference"]) >= 2
    delta = []
    for (name, g_ref), (_, g) in zip(taps["reference"], taps["patched"]):
        rel = float((g - g_ref).abs().max()) / float(g_ref.abs().max())
        delta.append(rel)
        print(f"d loss / d {name} output: relative difference {rel:.2e}")
        assert rel <= 2e-3, (name, rel)
    run_mode("reference_noisy", noise=delta)          # the reference's own sensitivity to a difference of that size
    module_scale = {}
    for k, g_ref in grads["reference"].items():
        if g_ref is not None:
            m = k.split(".")[0]
            module_scale[m] = max(module_scale.get(m, 0.0), float(g_ref.abs().max()))
    checked, flows_upstream, worst = 0, 0, (0.0, "")
    for k, g_ref in grads["reference"].items():
        g = grads["patched"][k]
        if g_ref is None or float(g_ref.abs().max()) == 0.0:
            assert g is None or float(g.abs().max()) < 1e-7, k
            continue
        assert g is not None, f"{k}: the patched network sends no gradient here"
        scale = float(g_ref.abs().max())
        err = float((g - g_ref).abs().max())
        # parameters whose true gradient is zero (a conv bias in front of an InstanceNorm, the bias in front of a softmax) hold
        # rounding noise only, in the reference too: measured against the module's largest gradient entry
        noise = 1e-4 * module_scale[k.split(".")[0]]
        upstream = k.split(".")[0] in ("image_encoder", "vis_encoder", "init_net")
        sens = float((grads["reference_noisy"][k] - g_ref).abs().max()) if upstream else 0.0
        if scale > noise:
            worst = max(worst, (err / scale, k))
        # 5e-3: the fine pass runs on resampled depths that differ in the last bits between the two implementations
        assert err <= 5e-3 * scale + noise + 6 * sens, (k, err, scale, sens, noise)
        checked += 1
        flows_upstream += upstream
    print(f"worst relative gradient difference {worst[0]:.2e} ({worst[1]}), {checked} parameters, {flows_upstream} upstream of the path")
    assert checked > 250 and flows_upstream > 100, (checked, flows_upstream)


def test_render_ops_refuse_to_cut_a_graph(ref_mod):
    """A drop-in without a backward raises when handed an input that requires grad (instead of silently detaching)."""
    from neuray_b200 import _lib, render_ops
    d = torch.rand(1, 8, 16, device="cuda", requires_grad=True)
    with pytest.raises(_lib.NeurayB200Error):
        render_ops.depth2dists(d)
    with torch.no_grad():
        render_ops.depth2dists(d)
    fm = torch.randn(2, 32, 12, 16, device="cuda", requires_grad=True)
    pts = torch.rand(2, 50, 2, device="cuda") * torch.tensor([60.0, 44.0], device="cuda")
    mask = (torch.rand(2, 50, device="cuda") > 0.2).float()
    out = render_ops.interpolate_feature_map(fm, pts, mask, 48, 64)
    out.square().sum().backward()
    import network.render_ops as ref_ops          # the reference's own torch implementation
    fm2 = fm.detach().clone().requires_grad_(True)
    want = ref_ops.interpolate_feature_map(fm2, pts, mask, 48, 64)
    want.square().sum().backward()
    assert torch.allclose(out, want, atol=1e-5) and torch.allclose(fm.grad, fm2.grad, atol=1e-4, rtol=1e-4)
