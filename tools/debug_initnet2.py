"""Diagnostics: is the unmodified reference's init_net forward reproducible around a patched run? are parameters untouched?"""
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
cfg = dict(T.CFG, use_self_hit_prob=False)
que, ref = T.make_data()
net = T.build(mod, cfg).train()

def snapshot():
    return {k: v.detach().clone() for k, v in net.state_dict().items()}

def same(a, b):
    return [k for k in a if not torch.equal(a[k], b[k])]

def fwd(mode, backward=True):
    cap = {}
    def h_init(m, i, o):
        cap["init"] = o.detach().clone()
    def h_img(m, i, o):
        cap["img"] = o.detach().clone()
    def h_conv1(m, i, o):
        cap["conv1_in"] = i[0].detach().clone()
        cap["conv1_out"] = o.detach().clone()
    hs = [net.init_net.register_forward_hook(h_init), net.image_encoder.register_forward_hook(h_img),
          net.init_net.res_net.conv1.register_forward_hook(h_conv1)]
    if mode == "patched":
        patch.install()
    try:
        net.zero_grad(set_to_none=True)
        out = T.run(net, que, ref, True)
        if backward:
            T.loss_of(out).backward()
        torch.cuda.synchronize()
    finally:
        patch.uninstall()
        for h in hs:
            h.remove()
    return cap

s0 = snapshot()
a = fwd("reference")
print("params changed after reference run:", same(s0, snapshot()))
b = fwd("reference")
print("ref vs ref: ", {k: float((a[k] - b[k]).abs().max()) for k in a})
c = fwd("patched")
print("params changed after patched run:", same(s0, snapshot()))
print("ref vs patched: ", {k: float((a[k] - c[k]).abs().max()) for k in a})
d = fwd("reference")
print("ref vs ref-after-patched: ", {k: float((a[k] - d[k]).abs().max()) for k in a})
c2 = fwd("patched", backward=False)
print("patched vs patched(no backward): ", {k: float((c[k] - c2[k]).abs().max()) for k in c})
