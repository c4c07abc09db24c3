"""Diagnostics for tests/test_reference_gpu.py: per-parameter gradient differences patched vs reference, by loss term."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import ref_import
from neuray_b200 import patch
import test_reference_gpu as T

torch.backends.cudnn.allow_tf32 = False
mod = ref_import.load_reference()
cfg = dict(T.CFG, use_self_hit_prob=False)
que, ref = T.make_data()
net = T.build(mod, cfg).train()

def losses(out):
    return {"render_c": ((out["pixel_colors_nr"] - out["pixel_colors_gt"]) ** 2).mean(),
            "render_f": ((out["pixel_colors_nr_fine"] - out["pixel_colors_gt_fine"]) ** 2).mean(),
            "depth": 0.1 * out["depth_mean"].abs().mean() + 0.1 * out["depth_mean_fine"].abs().mean()}

res = {}
for mode in ("reference", "reference_again", "patched"):
    if mode == "patched":
        patch.install()
    try:
        for term in ("render_c", "render_f", "depth"):
            net.zero_grad(set_to_none=True)
            out = T.run(net, que, ref, True)
            losses(out)[term].backward()
            torch.cuda.synchronize()
            res[(mode, term)] = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in net.named_parameters()}
    finally:
        patch.uninstall()
for term in ("render_c", "render_f", "depth"):
    rows = []
    for k, g in res[("reference", term)].items():
        if g is None or float(g.abs().max()) == 0:
            continue
        gp, ga = res[("patched", term)][k], res[("reference_again", term)][k]
        scale = float(g.abs().max())
        rows.append((float((gp - g).abs().max()) / scale if gp is not None else float("nan"), float((ga - g).abs().max()) / scale, k))
    rows.sort(reverse=True)
    print(f"== {term}: {len(rows)} params; worst patched-vs-ref (rel), ref-vs-ref floor (rel)")
    for r in rows[:8]:
        print("   %.3e  %.3e  %s" % r)
    top = {}
    for e, f, k in rows:
        m = k.split(".")[0]
        top[m] = max(top.get(m, (0, 0)), (e, f))
    print("   by module:", {m: ("%.2e" % v[0], "%.2e" % v[1]) for m, v in top.items()})
