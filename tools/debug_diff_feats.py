"""Diagnostics: the reference's get_diff_feats (init_net.py:29-61) with and without patch.install()."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import ref_import
from neuray_b200 import patch, synthetic
import test_reference_gpu as T
mod = ref_import.load_reference()
import network.init_net as ini
que, ref = T.make_data()
r = synthetic.to_device(ref, "cuda")
depth = ini.extract_depth_for_init(r)
with torch.no_grad():
    a = ini.get_diff_feats(r, depth)
    pa = ini.project_points_ref_views(r, torch.randn(5000, 3, device="cuda") * 2)
    patch.install()
    b = ini.get_diff_feats(r, depth)
    pb = ini.project_points_ref_views(r, (torch.manual_seed(0), torch.randn(5000, 3, device="cuda") * 2)[1])
    patch.uninstall()
d = (a - b).abs()
print("diff_feats max abs diff per channel", d.amax(dim=(0, 2, 3)).tolist())
print("fraction > 1e-4:", float((d > 1e-4).float().mean()), " > 1e-2:", float((d > 1e-2).float().mean()))
print("channel scales", a.abs().amax(dim=(0, 2, 3)).tolist())
