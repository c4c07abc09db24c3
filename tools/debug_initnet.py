"""Diagnostics: where does the init_net gradient difference (patched vs reference) come from?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import ref_import
from neuray_b200 import patch
import test_reference_gpu as T

torch.backends.cudnn.allow_tf32 = False
torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
mod = ref_import.load_reference()
import network.init_net as ini
cfg = dict(T.CFG, use_self_hit_prob=False)
que, ref = T.make_data()
net = T.build(mod, cfg).train()
res = {}

def run(mode):
    taps = {}
    hs = []
    def mk(name):
        def fwd(m, i, o):
            o.register_hook(lambda g: taps.__setitem__(name, g.detach().clone()))
            taps[name + "_out"] = o.detach().clone()
        return fwd
    hs.append(net.init_net.register_forward_hook(mk("init_net")))
    hs.append(net.vis_encoder.register_forward_hook(mk("vis_encoder")))
    hs.append(net.init_net.res_net.register_forward_hook(mk("res_net")))
    if mode != "reference":
        patch.install()
        if mode == "patched_no_initnet":
            for obj, name, val in list(patch._ORIGINALS):
                if getattr(obj, "__name__", "") == "network.init_net":
                    setattr(obj, name, val)
    try:
        net.zero_grad(set_to_none=True)
        out = T.run(net, que, ref, True)
        T.loss_of(out).backward()
        torch.cuda.synchronize()
    finally:
        patch.uninstall()
        for h in hs:
            h.remove()
    res[mode] = ({k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}, taps)

for m in ("reference", "patched", "patched_no_initnet"):
    run(m)
g0, t0 = res["reference"]
for m in ("patched", "patched_no_initnet"):
    g, t = res[m]
    print("==", m)
    for k in ("init_net", "vis_encoder", "res_net"):
        for suffix in ("", "_out"):
            a, b = t[k + suffix], t0[k + suffix]
            print(f"   {k+suffix:18s} rel diff {float((a - b).abs().max()) / float(b.abs().max()):.3e}")
    for k in ("init_net.res_net.conv1.weight", "init_net.conv_out.weight", "init_net.depth_skip.0.weight", "vis_encoder.out_conv.0.weight",
              "init_net.res_net.layer1.0.conv1.weight", "init_net.res_net.layer3.1.conv2.weight"):
        if k in g0:
            print(f"   {k:45s} rel err {float((g[k] - g0[k]).abs().max()) / float(g0[k].abs().max()):.3e}")
